"""Closed-loop MPC step on the device — the solver side of the reference's `MPC_GATO.run_mpc` loop
(python/bsqp/mpc_controller.py:233-253: broadcast state and reference window, reset_rho, solve, pick the hypothesis whose
`sim_forward` prediction is closest to the measured state (:294-309), adopt its trajectory as everybody's warm start).

The reference does this through four host round trips per control step (H2D of the whole batch, solve, D2H of the whole batch,
sim_forward + numpy argmin, numpy tiling); here it is one call, `gato_mpc_step` (include/gato_b200.h), with the batch of warm
starts resident on the device.  `host_mpc_step` is the same step composed from the plain solver calls exactly the way the reference
composes it; the parity tests run both (and the CPU oracle) side by side.
"""
import numpy as np


class DeviceMPC:
    """solver: gato_b200.native.Solver.  dt: the solver's knot spacing (MPC_GATO.dt)."""

    def __init__(self, solver, dt):
        self.s, self.dt = solver, float(dt)
        self.nx, self.nu, self.traj = solver.d["nx"], solver.d["nu"], solver.d["traj"]
        self.N = solver.N

    def warm_start(self, x0, ref_window):
        """mpc_controller.py:168-182: every knot = x0, u = 0; reset_dual; one warm-up solve (no reset_rho, no scoring)."""
        xu = np.zeros(self.traj, np.float32)
        for i in range(self.N):
            xu[i * (self.nx + self.nu): i * (self.nx + self.nu) + self.nx] = x0
        self.s.reset("dual")
        self.s.mpc_set_warm_start(xu)
        return self.s.mpc_step(x0, ref_window, None, None, 0.0, self.dt, reset_rho=False)

    def step(self, x_curr, ref_window, x_last, u_last, sim_dt):
        return self.s.mpc_step(x_curr, ref_window, x_last, u_last, sim_dt, self.dt, reset_rho=True)


def host_mpc_step(solver, XU_batch, x_curr, ref_window, x_last, u_last, sim_dt, dt, reset_rho=True, offsets=None):
    """The reference's composition (mpc_controller.py:238-253, 294-303) on any solver object with solve / sim_forward / reset
    (gato_b200.native.Solver or the CPU oracle's).  XU_batch [B, traj] is updated in place.  Returns (result dict, best_id, errors)."""
    B = XU_batch.shape[0]
    nx = np.asarray(x_curr).size
    xs = np.tile(np.asarray(x_curr, np.float32), (B, 1))
    if offsets is not None:
        xs = (xs + np.asarray(offsets, np.float32)).astype(np.float32)
    ref = np.tile(np.asarray(ref_window, np.float32), (B, 1))
    XU_batch[:, :nx] = xs
    if reset_rho:
        solver.reset("rho")
    res = solver.solve(XU_batch, xs, ref, dt)
    best, errors = 0, np.zeros(B)
    if x_last is not None:
        x_next = solver.sim_forward(x_last, u_last, sim_dt)
        errors = np.linalg.norm(x_next.astype(np.float64) - np.asarray(x_curr, np.float32).astype(np.float64)[None, :], axis=1)
        best = int(np.argmin(errors))
    XU_batch[:, :] = res["XU"][best, :]
    return res, best, errors
