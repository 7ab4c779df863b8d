"""`BSQP` — the wrapper class of the reference's python/bsqp/interface.py:6-237, same constructor keywords, methods and
`stats` dictionary, without the pinocchio dependency (state sizes come from the plant table; `ee_pos` uses the solver's own
forward kinematics, which is the one the cost is built on)."""
import numpy as np

from .. import native
from . import load_module


class BSQP:
    def __init__(self, model_path, batch_size, N, dt, max_sqp_iters=10, kkt_tol=1e-4, max_pcg_iters=100, pcg_tol=1e-4, solve_ratio=1.0, mu=1.0, q_cost=2.0, qd_cost=1e-4,
                 u_cost=1e-6, N_cost=50.0, q_lim_cost=1e-3, vel_lim_cost=0.0, ctrl_lim_cost=0.0, rho=0.0, rho_batch=None, mu_batch=None, pcg_tol_batch=None, adapt_rho=True,
                 plant_type="indy7"):
        if model_path and str(model_path).endswith(".gmdl"):
            # a robot given as a data file (gato_model_load / gato_model_register): its solves run the table-driven kernels.  (The reference takes
            # a URDF here, for pinocchio only -- its solver is compiled per robot.)
            plant_type = native.Model.load(model_path).register()
        if plant_type is None:  # interface.py:37-41
            plant_type = "iiwa14" if (model_path and "iiwa" in str(model_path).lower()) else "indy7"
        try:
            base = load_module(N, plant_type)
        except ImportError as e:
            raise ValueError(f"Number of knots {N} not supported (could not import bsqp.bsqpN{N}_{plant_type}): {e}")
        self.lib, self.plant_type = base, plant_type
        self.solver_class = getattr(base, f"BSQP_{batch_size}_float")
        self.solver = self.solver_class(dt, max_sqp_iters, kkt_tol, max_pcg_iters, pcg_tol, solve_ratio, mu, q_cost, qd_cost, u_cost, N_cost, q_lim_cost, vel_lim_cost,
                                        ctrl_lim_cost, rho)
        self.batch_size, self.N, self.dt = batch_size, N, dt
        self.nq = self.nv = native.NQ[plant_type]
        self.nx, self.nu = 2 * self.nq, self.nq
        self.f_ext_B = np.zeros((batch_size, 6), dtype=np.float32)
        self.set_f_ext_B(self.f_ext_B)
        self.XU_B = np.zeros((batch_size, N * (self.nx + self.nu) - self.nu), dtype=np.float32)
        self.stats = {k: np.array([]) for k in ("sqp_time_us", "sqp_iters", "kkt_converged", "pcg_iters", "pcg_times_us", "min_merit", "step_size", "initial_merit", "best_initial_merit")}
        if rho_batch is not None:
            self.solver.set_rho_penalty_batch(np.asarray(rho_batch, dtype=np.float32).reshape(batch_size), True)
        self.solver.set_rho_adaptation(bool(adapt_rho))
        if mu_batch is not None:
            self.solver.set_mu_batch(np.asarray(mu_batch, dtype=np.float32).reshape(batch_size))
        if pcg_tol_batch is not None:
            self.solver.set_pcg_tol_batch(np.asarray(pcg_tol_batch, dtype=np.float32).reshape(batch_size))

    def solve(self, xcur_B, eepos_goals_B, XU_B=None):
        xcur_B = np.asarray(xcur_B, dtype=np.float32)
        eepos_goals_B = np.asarray(eepos_goals_B, dtype=np.float32)
        XU_B = self.XU_B if XU_B is None else np.asarray(XU_B, dtype=np.float32)
        XU_B[:, : self.nx] = xcur_B  # interface.py:130
        result = self.solver.solve(XU_B, self.dt, xcur_B, eepos_goals_B)
        B = self.batch_size
        self.XU_B = np.asarray(result["XU"], dtype=np.float32)
        st = self.stats
        st["sqp_time_us"] = int(result["sqp_time_us"])
        st["sqp_iters"] = result["sqp_iters"].reshape(B)
        st["kkt_converged"] = result["kkt_converged"].reshape(B)
        st["final_merit"] = result["final_merit"].reshape(B)
        st["initial_merit"] = result["initial_merit"].reshape(B)
        st["best_initial_merit"] = float(np.min(st["initial_merit"])) if st["initial_merit"].size else np.array([], dtype=np.float32)
        st["ls_num_iters"] = int(result["ls_num_iters"])
        st["pcg_iters"], st["pcg_times_us"] = result["pcg_iters"], result["pcg_times_us"]
        st["min_merit"], st["step_size"] = result["ls_min_merit"], result["ls_step_size"]
        ls = st["min_merit"]
        if ls.size and ls.ndim == 2:
            best = np.min(ls.astype(np.float32), axis=1)
            st["best_merit_per_iter"], st["best_merit_iter1"] = best, float(best[0])
        else:
            st["best_merit_per_iter"], st["best_merit_iter1"] = np.array([], dtype=np.float32), float("nan")
        denom = st["best_initial_merit"] if np.size(st["best_initial_merit"]) else None
        st["best_merit_per_iter_normalized"] = st["best_merit_per_iter"] / denom if (denom and st["best_merit_per_iter"].size) else st["best_merit_per_iter"]
        return self.XU_B, result["sqp_time_us"]

    def ee_pos(self, q):  # interface.py:212-214 (pinocchio forward kinematics there; the solver's own kinematics here)
        return self.solver.ee_pos(np.asarray(q, dtype=np.float32).reshape(1, self.nq))[0]

    def reset(self):  # interface.py:216-219
        self.reset_dual()
        self.set_f_ext_B(np.zeros((self.batch_size, 6)))
        self.XU_B = np.zeros((self.batch_size, self.N * (self.nx + self.nu) - self.nu), dtype=np.float32)

    def sim_forward(self, xk, uk, sim_dt):
        return self.solver.sim_forward(np.asarray(xk, dtype=np.float32), np.asarray(uk, dtype=np.float32), sim_dt)

    def set_f_ext_B(self, f_ext_B):
        self.f_ext_B = np.asarray(f_ext_B, dtype=np.float32)
        self.solver.set_f_ext_batch(self.f_ext_B)

    def reset_rho(self):
        self.solver.reset_rho()

    def reset_dual(self):
        self.solver.reset_dual()

    def get_stats(self):
        return self.stats
