"""Drop-in Python surface of the reference's `bsqp` package, served by the B200-native solver.

* `gato_b200.bsqp.bsqpN{N}_{plant}` modules with classes `BSQP_{B}_float` and attribute `KNOT_POINTS` — the names,
  constructor signatures, methods, argument meaning and result dictionary of the pybind11 modules the reference builds
  from python/bindings.cu (class PyBSQP :10-209, registration :224-265).  Any knot count >= 3 the kernels support and any
  batch size are available (the reference compiles a fixed list); modules and classes are synthesised on import.
* `gato_b200.bsqp.interface.BSQP` — the wrapper class of python/bsqp/interface.py:6-237 without its pinocchio dependency.

    from gato_b200.bsqp import bsqpN32_iiwa14
    solver = bsqpN32_iiwa14.BSQP_512_float(dt, max_sqp_iters, kkt_tol, max_pcg_iters, pcg_tol, solve_ratio, mu,
                                           q_cost, qd_cost, u_cost, N_cost, q_lim_cost, vel_lim_cost, ctrl_lim_cost, rho)
    result = solver.solve(xu, timestep, x_s, reference)      # dict: XU, sqp_time_us, sqp_iters, kkt_converged, ...
"""
import importlib.abc
import importlib.machinery
import re
import sys
import types

import numpy as np

from .. import native

_MOD_RE = re.compile(r"^bsqpN(\d+)_(iiwa14|indy7)$")
_CLS_RE = re.compile(r"^BSQP_(\d+)_float$")
# default constructor of the reference (bsqp.cuh:24-28)
_DEFAULTS = (0.01, 5, 0.0001, 100, 1e-5, 1.0, 10.0, 1.0, 1e-3, 1e-6, 50.0, 1e-3, 0.0, 0.0, 1e-3)


def _make_class(plant, N, B):
    class _BSQP:
        __doc__ = f"BSQP solver, {plant}, {N} knot points, batch {B}, float32 (mirror of PyBSQP<float,{B}>, python/bindings.cu:10-209)"

        def __init__(self, *args, **kw):
            names = native.PARAM_ORDER
            if not args and not kw:
                args = _DEFAULTS
            if len(args) + len(kw) != 15:
                raise TypeError(f"expected 0 or 15 arguments ({', '.join(names)}), got {len(args) + len(kw)}")
            p = dict(zip(names, args))
            p.update(kw)
            self._s = native.Solver(plant, N, B, p)
            self._d = self._s.d

        def _arr(self, a, n, what):
            a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
            if a.size != n:
                raise ValueError(f"{what}: expected {n} float32 values, got {a.size}")
            return a

        def solve(self, xu_traj_batch, timestep, x_s_batch, reference_traj_batch):
            d = self._d
            r = self._s.solve(self._arr(xu_traj_batch, B * d["traj"], "xu_traj_batch"), self._arr(x_s_batch, B * d["nx"], "x_s_batch"),
                              self._arr(reference_traj_batch, B * 6 * N, "reference_traj_batch"), float(timestep))
            n_ls = r["n_ls"]
            # result dictionary of bindings.cu:96-147 (pcg_iters is cut to the number of line searches, :110-127)
            return {
                "XU": r["XU"].reshape(B, d["traj"]),
                "sqp_time_us": float(r["sqp_time_us"]),
                "sqp_iters": r["sqp_iters"],
                "kkt_converged": r["kkt_converged"],
                "final_merit": r["final_merit"],
                "initial_merit": r["initial_merit"],
                "ls_num_iters": int(n_ls),
                "pcg_times_us": self._pcg_times(n_ls),
                "pcg_iters": r["pcg_iters"][:n_ls].astype(np.int32),
                "ls_min_merit": r["ls_min_merit"],
                "ls_step_size": r["ls_step_size"],
            }

        def set_kernel_timing(self, enabled=True):
            """Extension: with timing on, `pcg_times_us` holds the measured device time of each PCG kernel
            (the reference fills it with zeros, bsqp.cuh:138)."""
            self._timing = bool(enabled)
            self._s.set_kernel_timing(self._timing)

        def _pcg_times(self, n_ls):
            out = np.zeros(n_ls, np.float32)
            if getattr(self, "_timing", False):
                t = [1e3 * ms for k, ms in self._s.launch_times() if k == "k_pcg"][:n_ls]
                out[:len(t)] = t
            return out

        def reset_dual(self):
            self._s.reset("dual")

        def reset_rho(self):
            self._s.reset("rho")

        def set_rho_adaptation(self, enabled):
            self._s.set_rho_adaptation(bool(enabled))

        def set_f_ext_batch(self, f_ext_batch):
            self._s.set_batch("f_ext", self._arr(f_ext_batch, 6 * B, "f_ext_batch"))

        def set_rho_penalty_batch(self, rho_batch, set_as_reset_default=True):
            self._s.set_batch("rho", self._arr(rho_batch, B, "rho_batch"), set_as_reset_default)

        def set_drho_batch(self, drho_batch, set_as_reset_default=True):
            self._s.set_batch("drho", self._arr(drho_batch, B, "drho_batch"), set_as_reset_default)

        def set_mu_batch(self, mu_batch):
            self._s.set_batch("mu", self._arr(mu_batch, B, "mu_batch"))

        def set_pcg_tol_batch(self, pcg_tol_batch):
            self._s.set_batch("pcg_tol", self._arr(pcg_tol_batch, B, "pcg_tol_batch"))

        def sim_forward(self, xk, uk, dt):
            return self._s.sim_forward(self._arr(xk, self._d["nx"], "xk"), self._arr(uk, self._d["nu"], "uk"), float(dt))

        def ee_pos(self, q):
            """Extension: end-effector xyz of joint configurations q[n][nq] by the solver's own forward kinematics (what interface.py:212-214
            asks pinocchio for in the reference)."""
            return self._s.ee_pos(q)

        def set_kkt_residual_log(self, enabled=True):
            """Extension: with the log on, `kkt_residuals()` returns the per-iteration KKT residual norms
            (q_max, c_max) the reference computes and discards (bsqp.cuh:149-150)."""
            self._s.set_kkt_residual_log(enabled)

        def kkt_residuals(self):
            return self._s.kkt_residuals()

    _BSQP.__name__ = _BSQP.__qualname__ = f"BSQP_{B}_float"
    return _BSQP


class _BsqpModule(types.ModuleType):
    def __init__(self, name, plant, N):
        super().__init__(name, f"BSQP solver, {plant}, {N} knot points (gato_b200 mirror of bsqpN{N}_{plant})")
        self.KNOT_POINTS, self._plant, self._cache = N, plant, {}

    def __getattr__(self, item):
        m = _CLS_RE.match(item)
        if not m or int(m.group(1)) < 1:
            raise AttributeError(item)
        B = int(m.group(1))
        if B not in self._cache:
            self._cache[B] = _make_class(self._plant, self.KNOT_POINTS, B)
        return self._cache[B]

    def __dir__(self):
        return ["KNOT_POINTS"] + [f"BSQP_{b}_float" for b in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024)]


def load_module(N, plant):
    """Return the module object `bsqpN{N}_{plant}` (same as `import gato_b200.bsqp.bsqpN{N}_{plant}`)."""
    name = f"{__name__}.bsqpN{N}_{plant}"
    if name not in sys.modules:
        if plant not in native.NQ or N < 3:
            raise ImportError(f"no such solver module: bsqpN{N}_{plant}")
        sys.modules[name] = _BsqpModule(name, plant, N)
    return sys.modules[name]


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if not fullname.startswith(__name__ + "."):
            return None
        if _MOD_RE.match(fullname[len(__name__) + 1:]):
            return importlib.machinery.ModuleSpec(fullname, self)
        return None

    def create_module(self, spec):
        m = _MOD_RE.match(spec.name[len(__name__) + 1:])
        return load_module(int(m.group(1)), m.group(2))

    def exec_module(self, module):
        pass


sys.meta_path.insert(0, _Finder())


def __getattr__(item):
    m = _MOD_RE.match(item)
    if m:
        return load_module(int(m.group(1)), m.group(2))
    raise AttributeError(item)
