"""Closed-loop MPC step on the device against the reference's host-side composition of the same step
(python/bsqp/mpc_controller.py:233-253, 294-309) run on (a) the same CUDA solver through its plain calls and (b) the CPU oracle.
Everything is compared bit-for-bit: selected hypothesis, float64 errors (numpy's summation order), trajectories, iteration counts."""
import numpy as np
import pytest

from gato_b200.bsqp.mpc import DeviceMPC, host_mpc_step
from gato_b200.workloads import DEFAULT_SOLVER_PARAMS, figure8

pytestmark = pytest.mark.gpu


def _setup(plant, N, B, seed):
    from gato_b200 import native

    p = dict(DEFAULT_SOLVER_PARAMS)
    p.update(max_sqp_iters=2, max_pcg_iters=100, dt=0.01)
    rng = np.random.default_rng(seed)
    fext = rng.normal(0, 5.0, (B, 6)).astype(np.float32)
    fext[0] = 0
    rho = np.logspace(-4, 0, B).astype(np.float32)
    dt = 0.01
    fig = figure8(dt).reshape(-1, 6)

    def mk(factory):
        s = factory()
        s.set_batch("f_ext", fext)
        s.set_batch("rho", rho, True)
        return s

    dev = mk(lambda: native.Solver(plant, N, B, p, device=0))
    host = mk(lambda: native.Solver(plant, N, B, p, device=0))
    return p, dt, fig, fext, dev, host, mk


@pytest.mark.parametrize("plant,N,B", [("iiwa14", 8, 16), ("indy7", 16, 8)])
def test_device_mpc_step_equals_host_composition_and_oracle(plant, N, B):
    from oracle.pyapi import Backend, ensure_oracle_built

    ensure_oracle_built()
    p, dt, fig, fext, dev, host, mk = _setup(plant, N, B, 7)
    orc = mk(lambda: Backend("oracle", plant, N).solver(B, p))
    nx, nu, traj = dev.d["nx"], dev.d["nu"], dev.d["traj"]
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-0.3, 0.3, nx // 2), np.zeros(nx // 2)]).astype(np.float32)
    true_hyp = 3

    mpc = DeviceMPC(dev, dt)
    ref0 = fig[:N].reshape(-1)
    r_dev = mpc.warm_start(x, ref0)
    XU_h = np.zeros((B, traj), np.float32)
    XU_o = np.zeros((B, traj), np.float32)
    for XU in (XU_h, XU_o):
        for i in range(N):
            XU[:, i * (nx + nu): i * (nx + nu) + nx] = x
    host.reset("dual"), orc.reset("dual")
    r_h, _, _ = host_mpc_step(host, XU_h, x, ref0, None, None, 0.0, dt, reset_rho=False)
    r_o, _, _ = host_mpc_step(orc, XU_o, x, ref0, None, None, 0.0, dt, reset_rho=False)
    assert np.array_equal(r_dev["XU_best"], XU_h[0]) and np.array_equal(XU_h, XU_o)
    assert np.array_equal(dev.mpc_get_warm_start(), XU_h)
    best_ids = []
    for step in range(1, 6):
        x_last, u_last = x.copy(), XU_h[0, nx:nx + nu].copy()
        # the "real" plant is hypothesis `true_hyp` of the solver's own simulator (cfg 5)
        x = host.sim_forward(x_last, u_last, dt)[true_hyp].copy()
        ref_w = fig[step:step + N].reshape(-1)
        r_dev = mpc.step(x, ref_w, x_last, u_last, dt)
        r_h, b_h, e_h = host_mpc_step(host, XU_h, x, ref_w, x_last, u_last, dt, dt)
        r_o, b_o, e_o = host_mpc_step(orc, XU_o, x, ref_w, x_last, u_last, dt, dt)
        assert r_dev["best_id"] == b_h == b_o
        assert np.array_equal(r_dev["errors"], e_h) and np.array_equal(e_h, e_o), "float64 hypothesis errors differ from numpy's"
        assert np.array_equal(r_dev["XU_best"], XU_h[0]) and np.array_equal(XU_h, XU_o)
        assert np.array_equal(r_dev["pcg_iters"], r_h["pcg_iters"]) and np.array_equal(r_h["pcg_iters"], r_o["pcg_iters"])
        assert np.array_equal(r_dev["ls_step_size"], r_o["ls_step_size"])
        assert np.array_equal(dev.mpc_get_warm_start(), XU_h)
        best_ids.append(b_h)
    assert true_hyp in best_ids, "the scoring never identified the true hypothesis"


def test_device_mpc_state_offsets_match_host_composition():
    p, dt, fig, fext, dev, host, mk = _setup("iiwa14", 8, 16, 11)
    B, N = 16, 8
    nx, nu, traj = dev.d["nx"], dev.d["nu"], dev.d["traj"]
    rng = np.random.default_rng(5)
    off = rng.normal(0, 0.01, (B, nx)).astype(np.float32)
    x = np.zeros(nx, np.float32)
    XU_h = np.zeros((B, traj), np.float32)
    dev.reset("dual"), host.reset("dual")
    dev.mpc_set_warm_start(XU_h[0])
    dev.mpc_set_state_offsets(off)
    ref_w = fig[:N].reshape(-1)
    r_dev = dev.mpc_step(x, ref_w, x, np.zeros(nu, np.float32), dt, dt, reset_rho=True)
    r_h, b_h, e_h = host_mpc_step(host, XU_h, x, ref_w, x, np.zeros(nu, np.float32), dt, dt, offsets=off)
    assert r_dev["best_id"] == b_h and np.array_equal(r_dev["errors"], e_h)
    assert np.array_equal(r_dev["XU_best"], XU_h[0])


def test_device_mpc_scoring_follows_numpy_argmin_with_nan():
    """np.argmin returns the first NaN if there is one (mpc_controller.py:301): a hypothesis whose wrench is NaN must win on the device too."""
    p, dt, fig, fext, dev, host, mk = _setup("iiwa14", 8, 16, 21)
    B, N = 16, 8
    nx, nu, traj = dev.d["nx"], dev.d["nu"], dev.d["traj"]
    fe = fext.copy()
    fe[5, 2] = np.nan
    fe[11, 0] = np.nan
    dev.set_batch("f_ext", fe), host.set_batch("f_ext", fe)
    x = np.zeros(nx, np.float32)
    XU_h = np.zeros((B, traj), np.float32)
    dev.reset("dual"), host.reset("dual")
    dev.mpc_set_warm_start(XU_h[0])
    ref_w = fig[:N].reshape(-1)
    r_dev = dev.mpc_step(x, ref_w, x, np.zeros(nu, np.float32), dt, dt, reset_rho=True)
    r_h, b_h, e_h = host_mpc_step(host, XU_h, x, ref_w, x, np.zeros(nu, np.float32), dt, dt)
    assert b_h == 5 and r_dev["best_id"] == b_h
    assert np.array_equal(np.isnan(r_dev["errors"]), np.isnan(e_h)) and np.array_equal(r_dev["errors"][~np.isnan(e_h)], e_h[~np.isnan(e_h)])
