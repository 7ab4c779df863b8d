"""Host-side logic of the closed-loop MPC step (gato_b200/bsqp/mpc.py: host_mpc_step), exercised on the CPU oracle: the composition
the reference controller uses (python/bsqp/mpc_controller.py:233-253, 294-309)."""
import numpy as np

from gato_b200.bsqp.mpc import host_mpc_step
from gato_b200.workloads import DEFAULT_SOLVER_PARAMS, figure8
from oracle.pyapi import Backend, ensure_oracle_built


def test_host_mpc_step_scores_hypotheses_and_adopts_the_best():
    ensure_oracle_built()
    plant, N, B, dt = "iiwa14", 8, 4, 0.01
    p = dict(DEFAULT_SOLVER_PARAMS)
    p.update(max_sqp_iters=1, max_pcg_iters=50, dt=dt)
    s = Backend("oracle", plant, N).solver(B, p)
    fext = np.zeros((B, 6), np.float32)
    fext[1:, 2] = [4.0, -6.0, 9.0]
    s.set_batch("f_ext", fext)
    nx, nu = 14, 7
    traj = (nx + nu) * N - nu
    fig = figure8(dt).reshape(-1, 6)
    x = np.zeros(nx, np.float32)
    XU = np.zeros((B, traj), np.float32)
    s.reset("dual")
    res, best, err = host_mpc_step(s, XU, x, fig[:N].reshape(-1), None, None, 0.0, dt, reset_rho=False)
    assert best == 0 and not err.any()
    assert np.array_equal(XU, np.tile(res["XU"][0], (B, 1)))  # everybody adopts the selected trajectory
    true_hyp = 2
    for step in range(1, 4):
        x_last, u_last = x.copy(), XU[0, nx:nx + nu].copy()
        x = s.sim_forward(x_last, u_last, dt)[true_hyp].copy()
        res, best, err = host_mpc_step(s, XU, x, fig[step:step + N].reshape(-1), x_last, u_last, dt, dt)
        assert best == true_hyp and err[true_hyp] == 0.0 and err.dtype == np.float64
        assert (np.delete(err, true_hyp) > 0).all()
        assert np.array_equal(XU[:, :nx], np.tile(res["XU"][best, :nx], (B, 1)))
        assert np.array_equal(XU, np.tile(res["XU"][best], (B, 1)))
