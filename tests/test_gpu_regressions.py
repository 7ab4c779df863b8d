"""GPU regression tests for defects found in review (run on the B200 box: pytest -m gpu); everything goes through the C ABI."""
import ctypes as C

import numpy as np
import pytest

from conftest import n_mismatch
from gato_b200.workloads import make_config

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def backends(oracle_built):
    from gato_b200.native import GatoBackend
    from oracle.pyapi import Backend

    return lambda plant, N: (Backend("oracle", plant, N), GatoBackend(plant, N))


def test_solvers_of_different_horizons_coexist(backends):
    """The dynamic shared-memory opt-in is global per kernel and device: creating a short-horizon solver (or running a stage call, which
    builds a temporary one) must not lower the limit under a long-horizon solver that is still alive."""
    o32, g32 = backends("iiwa14", 32)
    o8, g8 = backends("iiwa14", 8)
    w32, w8 = make_config(2, B=6), make_config(1, B=3)
    big = g32.solver(6, w32["params"])
    first = big.solve(w32["xu"], w32["xs"], w32["ref"], w32["dt"])
    small = g8.solver(3, w8["params"])  # created while `big` is alive
    rs = small.solve(w8["xu"], w8["xs"], w8["ref"], w8["dt"])
    g8.stage_kkt(3, w8["xu"], w8["xs"], w8["ref"], np.zeros((3, 6), np.float32), w8["dt"], w8["params"])  # temporary N=8 solver
    big.reset("dual"), big.reset("rho")
    again = big.solve(w32["xu"], w32["xs"], w32["ref"], w32["dt"])
    assert n_mismatch(again["XU"], first["XU"]) == 0 and np.array_equal(again["pcg_iters"], first["pcg_iters"])
    ro = o8.solver(3, w8["params"]).solve(w8["xu"], w8["xs"], w8["ref"], w8["dt"])
    assert n_mismatch(rs["XU"], ro["XU"]) == 0
    ro32 = o32.solver(6, w32["params"]).solve(w32["xu"], w32["xs"], w32["ref"], w32["dt"])
    assert n_mismatch(first["XU"], ro32["XU"]) == 0


@pytest.mark.parametrize("plant,N", [("iiwa14", 8), ("indy7", 16)])
def test_zero_sqp_iterations(backends, plant, N):
    """max_sqp_iters = 0 (bsqp.cuh:121): no iteration runs, the trajectory is untouched, only the initial / final merit is evaluated."""
    o, g = backends(plant, N)
    w = make_config(1 if plant == "iiwa14" else 3, B=4, N=N)
    p = dict(w["params"], max_sqp_iters=0)
    ro, rg = o.solver(4, p).solve(w["xu"], w["xs"], w["ref"], w["dt"]), g.solver(4, p).solve(w["xu"], w["xs"], w["ref"], w["dt"])
    assert (rg["n_pcg"], rg["n_ls"]) == (ro["n_pcg"], ro["n_ls"]) == (0, 0)
    assert n_mismatch(rg["XU"], w["xu"]) == 0 and n_mismatch(rg["XU"], ro["XU"]) == 0
    assert np.array_equal(rg["sqp_iters"], ro["sqp_iters"]) and np.array_equal(rg["kkt_converged"], ro["kkt_converged"])
    for k in ("final_merit", "initial_merit"):
        assert n_mismatch(rg[k], ro[k]) == 0, k
    assert n_mismatch(rg["final_merit"], rg["initial_merit"]) == 0


def test_drho_batch_setter(backends):
    """set_drho_batch (bsqp.cuh:75): the per-solve rho multiplier enters the first line search's rho update, is reset to its default after
    every solve (bsqp.cuh:189), and the default itself can be replaced."""
    o, g = backends("iiwa14", 8)
    B = 6
    w = make_config(1, B=B)
    p = dict(w["params"], max_sqp_iters=3)
    drho = np.array([0.25, 0.5, 1.0, 2.0, 4.0, 8.0], np.float32)
    for set_default in (False, True):
        so, sg = o.solver(B, p), g.solver(B, p)
        for rep in range(2):
            if rep == 0 or not set_default:
                so.set_batch("drho", drho, set_default), sg.set_batch("drho", drho, set_default)
            ro, rg = so.solve(w["xu"], w["xs"], w["ref"], w["dt"]), sg.solve(w["xu"], w["xs"], w["ref"], w["dt"])
            for k in ("XU", "ls_step_size", "ls_min_merit", "final_merit"):
                assert n_mismatch(rg[k], ro[k]) == 0, (set_default, rep, k)
            assert np.array_equal(rg["pcg_iters"], ro["pcg_iters"])
    # the multiplier matters: different drho -> different rho after the first line search -> different later iterates
    base = g.solver(B, p).solve(w["xu"], w["xs"], w["ref"], w["dt"])
    assert n_mismatch(base["XU"], rg["XU"]) > 0


def test_get_merits_is_ordered_and_rejects_pending_solves():
    import torch

    from gato_b200 import native

    w = make_config(2, B=8)
    s = native.Solver(w["plant"], w["N"], 8, w["params"])
    r = s.solve(w["xu"], w["xs"], w["ref"], w["dt"])
    fin, ini = np.zeros(8, np.float32), np.zeros(8, np.float32)
    lib = native.load()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    assert lib.gato_get_merits(s.h, vp(fin), vp(ini)) == 0
    assert n_mismatch(fin, r["final_merit"]) == 0 and n_mismatch(ini, r["initial_merit"]) == 0
    xu, xs, rf = torch.from_numpy(w["xu"]).cuda(), torch.from_numpy(w["xs"]).cuda(), torch.from_numpy(w["ref"]).cuda()
    torch.cuda.synchronize()
    s.solve_async(xu.data_ptr(), xs.data_ptr(), rf.data_ptr(), w["dt"])
    assert lib.gato_get_merits(s.h, vp(fin), vp(ini)) == -1 and b"pending" in lib.gato_last_error(s.h)
    s.solve_wait()
    assert lib.gato_get_merits(s.h, vp(fin), vp(ini)) == 0


@pytest.mark.parametrize("plant", ["iiwa14", "indy7"])
def test_ee_pos_matches_the_oracle_kinematics(backends, plant):
    """gato_ee_pos / BSQP.ee_pos (reference python/bsqp/interface.py:212-214): the solver's forward kinematics, bit-for-bit the oracle's
    end-effector position (the one the tracking cost is built on)."""
    from gato_b200.bsqp.interface import BSQP

    o, g = backends(plant, 8)
    nq = o.d["nq"]
    rng = np.random.default_rng(3)
    q = rng.uniform(-2.5, 2.5, (37, nq)).astype(np.float32)
    x = np.concatenate([q, np.zeros_like(q)], 1)
    want = o.dyn_dump(x, np.zeros((37, nq), np.float32), np.zeros((37, 6), np.float32))["ee"][:, :3]
    w = make_config(1 if plant == "iiwa14" else 3, B=2, N=8)
    got = g.solver(2, w["params"]).ee_pos(q)
    assert n_mismatch(got, want) == 0
    b = BSQP(None, 2, 8, 0.01, plant_type=plant)
    assert n_mismatch(b.ee_pos(q[5]), want[5]) == 0


def test_kkt_residual_log(backends):
    """Optional q_max / c_max outputs (the norms bsqp.cuh:149-150 computes and discards): equal to max |.| of the oracle's stage outputs, and
    switching the log on does not change the solve."""
    o, g = backends("iiwa14", 8)
    B = 5
    w = make_config(1, B=B)
    p = dict(w["params"], max_sqp_iters=1)
    xu = w["xu"] + np.random.default_rng(2).normal(0, 0.05, w["xu"].shape).astype(np.float32)
    plain = g.solver(B, p).solve(xu, w["xs"], w["ref"], w["dt"])
    s = g.solver(B, p)
    s.set_kkt_residual_log(True)
    r = s.solve(xu, w["xs"], w["ref"], w["dt"])
    assert n_mismatch(r["XU"], plain["XU"]) == 0
    qm, cm = s.kkt_residuals()
    assert qm.shape == (1, B) and cm.shape == (1, B)
    # the same iteration through the oracle's stages
    kk = o.stage_kkt(B, xu, w["xs"], w["ref"], np.zeros((B, 6), np.float32), w["dt"], p)
    sc = o.stage_schur(B, kk, np.full(B, p["rho"], np.float32))
    lam, _ = o.stage_pcg(B, sc["S"], sc["Pinv"], sc["gamma"], np.zeros((B, o.d["vecp"]), np.float32), np.full(B, p["pcg_tol"], np.float32), int(p["max_pcg_iters"]))
    _, qres, _ = o.stage_dz(B, lam, sc["Qinv"], sc["Rinv"], kk["q"], kk["r"], kk["A"], kk["Bm"])
    assert n_mismatch(qm[0], np.abs(qres.reshape(B, -1)).max(1)) == 0
    assert n_mismatch(cm[0], np.abs(kk["c"].reshape(B, -1)).max(1)) == 0


@pytest.mark.parametrize("plant,N", [("iiwa14", 35), ("iiwa14", 64), ("iiwa14", 65), ("iiwa14", 97), ("iiwa14", 128), ("iiwa14", 144), ("indy7", 41), ("indy7", 80), ("indy7", 121), ("indy7", 168)])
def test_cluster_pcg_kernel_long_horizons(backends, plant, N, monkeypatch):
    """k_pcg_cluster (thread-block cluster per solve, rows in registers across 2 ... 5 CTAs, halos and dot products over distributed shared
    memory): whole solves and the PCG stage bit-for-bit against the oracle, full and partial last CTAs, one and two elements per virtual
    thread of the reference's 1024-thread dot product; and identical to the streaming kernel it replaces (GATO_PCG_NO_CLUSTER=1)."""
    o, g = backends(plant, N)
    B = 3
    w = make_config(2 if plant == "iiwa14" else 3, B=B, N=N)
    p = dict(w["params"], max_sqp_iters=2, max_pcg_iters=60, pcg_tol=1e-5)
    ro, rg = o.solver(B, p).solve(w["xu"], w["xs"], w["ref"], w["dt"]), g.solver(B, p).solve(w["xu"], w["xs"], w["ref"], w["dt"])
    assert np.array_equal(rg["pcg_iters"], ro["pcg_iters"]) and np.array_equal(rg["ls_step_size"], ro["ls_step_size"])
    for k in ("XU", "final_merit"):
        assert n_mismatch(rg[k], ro[k]) == 0, k
    monkeypatch.setenv("GATO_PCG_NO_CLUSTER", "1")
    rs = g.solver(B, p).solve(w["xu"], w["xs"], w["ref"], w["dt"])
    monkeypatch.delenv("GATO_PCG_NO_CLUSTER")
    assert n_mismatch(rs["XU"], rg["XU"]) == 0 and np.array_equal(rs["pcg_iters"], rg["pcg_iters"])
    # fixed cap (no early exit) and a flagged solve
    p2 = dict(p, max_sqp_iters=1, max_pcg_iters=25, pcg_tol=-1.0)
    ro, rg = o.solver(B, p2).solve(w["xu"], w["xs"], w["ref"], w["dt"]), g.solver(B, p2).solve(w["xu"], w["xs"], w["ref"], w["dt"])
    assert (rg["pcg_iters"] == 25).all() and n_mismatch(rg["XU"], ro["XU"]) == 0


@pytest.mark.parametrize("plant", ["iiwa14", "indy7"])
def test_kkt_large_grid_path_bit_exact(backends, plant):
    """k_kkt (the large-grid kernel: rolled halves, two gradient columns per pass as the lanes of packed FFMA2 pairs) against the oracle at stage
    level; the small stage tests all take k_kkt_fine.  (This is the test that would have caught ptxas fusing mul.rn.f32x2 + add.rn.f32x2.)"""
    o, g = backends(plant, 32)
    B = 80  # 80 x 32 items: beyond k_kkt_fine's grid limit
    w = make_config(2 if plant == "iiwa14" else 3, B=B, N=32)
    rng = np.random.default_rng(1)
    xu = (w["xu"] + rng.normal(0, 0.3, w["xu"].shape)).astype(np.float32)
    fext = rng.normal(0, 2, (B, 6)).astype(np.float32)
    p = dict(w["params"], vel_lim_cost=0.002, ctrl_lim_cost=0.001)
    ko, kg = o.stage_kkt(B, xu, w["xs"], w["ref"], fext, w["dt"], p), g.stage_kkt(B, xu, w["xs"], w["ref"], fext, w["dt"], p)
    for k in ko:
        assert n_mismatch(kg[k], ko[k]) == 0, k
