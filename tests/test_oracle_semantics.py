"""Named tests for the must-preserve semantics of SURVEY.md §8(c), exercised on the oracle (CPU)."""
import numpy as np
import pytest

from conftest import n_mismatch
from gato_b200.workloads import DEFAULT_SOLVER_PARAMS, figure8, make_config, traj_size, warm_start
from oracle.pyapi import Backend


def test_workload_shapes():
    for cfg in (1, 2, 3, 4, 5, "bench"):
        w = make_config(cfg, B=4)
        nq = 7 if w["plant"] == "iiwa14" else 6
        assert w["xu"].shape == (4, traj_size(w["plant"], w["N"])) and w["xs"].shape == (4, 2 * nq) and w["ref"].shape == (4, 6 * w["N"])
        assert w["xu"].dtype == np.float32
    f = figure8(0.01).reshape(-1, 6)
    assert f.shape == (3000, 6) and np.allclose(f[:, 3:], 0)
    xu = warm_start(np.ones((2, 14), np.float32), 8, 7)
    assert xu.shape == (2, 161) and np.all(xu.reshape(2, -1)[:, 14:21] == 0)


def test_rho_regularises_position_block_only(oracle_built):
    """item 1: rho is added to the first nx/2 diagonal entries only (linalg.cuh:91)."""
    be = Backend("oracle", "iiwa14", 8)
    w = make_config(1, B=1)
    k = be.stage_kkt(1, w["xu"], w["xs"], w["ref"], np.zeros((1, 6), np.float32), 0.01, w["params"])
    s0 = be.stage_schur(1, k, np.array([0.0], np.float32))
    s1 = be.stage_schur(1, k, np.array([0.5], np.float32))
    P0 = -s0["Pinv"][0, 0].reshape(14, 42)[:, 14:28]  # row 0 main = -(Q_0 + rho I~)
    P1 = -s1["Pinv"][0, 0].reshape(14, 42)[:, 14:28]
    d = np.diag(P1 - P0)
    assert np.allclose(d[:7], 0.5) and np.all(d[7:] == 0)


def test_terminal_block_uses_previous_state_and_q_cost(oracle_built):
    """item 2: Q_{N-1}, q_{N-1} are evaluated at x_{N-2} against ref_{N-1} with weight q_cost (setup_kkt.cuh:90-91)."""
    be = Backend("oracle", "iiwa14", 8)
    w = make_config(1, B=1)
    xu = w["xu"].copy().reshape(1, -1)
    rng = np.random.default_rng(1)
    xu += rng.normal(0, 0.1, xu.shape).astype(np.float32)
    fe = np.zeros((1, 6), np.float32)
    base = be.stage_kkt(1, xu, w["xs"], w["ref"], fe, 0.01, w["params"])
    x2 = xu.copy()
    x2[0, 7 * 21 : 7 * 21 + 14] += 0.3  # perturb x_{N-1}: the terminal block must not change
    pert = be.stage_kkt(1, x2, w["xs"], w["ref"], fe, 0.01, w["params"])
    assert n_mismatch(pert["Q"][0, 7], base["Q"][0, 7]) == 0 and n_mismatch(pert["q"][0, 7], base["q"][0, 7]) == 0
    p2 = dict(w["params"], N_cost=500.0)  # N_cost must not enter the KKT blocks ...
    assert n_mismatch(be.stage_kkt(1, xu, w["xs"], w["ref"], fe, 0.01, p2)["Q"], base["Q"]) == 0
    m1 = be.stage_merit(1, xu, np.zeros_like(xu), w["xs"], w["ref"], np.array([10], np.float32), fe, 0.01, w["params"], 1)
    m2 = be.stage_merit(1, xu, np.zeros_like(xu), w["xs"], w["ref"], np.array([10], np.float32), fe, 0.01, p2, 1)
    assert m2[0, 0] > m1[0, 0]  # ... but it does weight the last knot of the merit (plant:315)


def test_cost_hessian_is_rank_one(oracle_built):
    """item 3: Q[0:nq,0:nq] = w (J^T e)(J^T e)^T + diag(barrier'') — rank one apart from the diagonal barrier term."""
    be = Backend("oracle", "iiwa14", 8)
    w = make_config(1, B=1)
    p = dict(w["params"], q_lim_cost=0.0)
    xu = w["xu"] + np.random.default_rng(2).normal(0, 0.2, w["xu"].shape).astype(np.float32)
    k = be.stage_kkt(1, xu, w["xs"], w["ref"], np.zeros((1, 6), np.float32), 0.01, p)
    Q = k["Q"][0, 2].reshape(14, 14)[:7, :7].astype(np.float64)
    sv = np.linalg.svd(Q, compute_uv=False)
    assert sv[1] < 1e-5 * sv[0]


def test_convergence_flag_and_iteration_counting(oracle_built):
    """items 5/6: flag = 'PCG performed 0 iterations'; sqp_iters counts outer iterations for every solve; early exit skips the line search."""
    be = Backend("oracle", "iiwa14", 8)
    w = make_config(1, B=4)
    p = dict(w["params"], max_sqp_iters=3, max_pcg_iters=0)  # 0 PCG iterations => every solve is flagged in the first iteration
    o = be.solver(4, p).solve(w["xu"], w["xs"], w["ref"], w["dt"])
    assert o["n_pcg"] == 1 and o["n_ls"] == 0 and (o["kkt_converged"] == 1).all() and (o["sqp_iters"] == 1).all()
    assert n_mismatch(o["XU"], w["xu"]) == 0  # nothing moved
    p = dict(w["params"], max_sqp_iters=3, solve_ratio=0.0)  # threshold 0 => break after the first PCG, no line search
    o = be.solver(4, p).solve(w["xu"], w["xs"], w["ref"], w["dt"])
    assert o["n_pcg"] == 1 and o["n_ls"] == 0 and (o["sqp_iters"] == 1).all() and (o["kkt_converged"] == 0).all()


def test_line_search_rules(oracle_built):
    """item 7: first strict minimum wins; accept iff strictly below the running merit; rho / drho updates and clamps."""
    be = Backend("oracle", "iiwa14", 8)
    B = 5
    xu = np.zeros((B, 161), np.float32)
    dz = np.ones((B, 161), np.float32)
    m8 = np.tile(np.array([5, 3, 3, 4, 9, 9, 9, 9], np.float32), (B, 1))
    m8[1] = [np.nan] * 8  # all-NaN merits: treated as 1e38 -> failure
    m8[2, :] = 7.0  # tie everywhere: index 0
    mi = np.array([4.0, 4.0, 7.5, 3.0, 4.0], np.float32)  # solve 3: min == 3.0 is not strictly below 3.0 -> failure
    rho = np.array([0.01, 0.01, 9.0, 1e-8, 0.01], np.float32)
    drho = np.array([1.0, 1.0, 1.0, 1.0, 3.0], np.float32)
    o = be.stage_linesearch(B, xu, dz, m8, mi, rho, drho, 1)
    assert o["step"].tolist() == [0.5, -1.0, 1.0, -1.0, 0.5]
    assert np.allclose(o["xu"][0], 0.5) and np.all(o["xu"][1] == 0) and np.allclose(o["xu"][2], 1.0)
    f = np.float32(1.2)
    assert o["drho"][0] == np.float32(1.0) / f and o["drho"][1] == f and o["drho"][4] == min(np.float32(3.0) / f, np.float32(1.0) / f)
    assert o["rho"][3] == np.float32(max(np.float32(1e-8) * f, np.float32(1e-8)))
    assert o["merit_init"].tolist() == [3.0, 4.0, 7.0, 3.0, 3.0]
    o2 = be.stage_linesearch(B, xu, dz, m8, mi, rho, drho, 0)  # adaptation off: rho, drho untouched
    assert n_mismatch(o2["rho"], rho) == 0 and n_mismatch(o2["drho"], drho) == 0


def test_state_persistence_and_resets(oracle_built):
    """item 8: lambda and rho persist across solves, drho is reset after each solve, reset_rho restores the defaults."""
    be = Backend("oracle", "iiwa14", 8)
    w = make_config(1, B=2)
    p = dict(w["params"], max_sqp_iters=2)
    s = be.solver(2, p)
    a = s.solve(w["xu"], w["xs"], w["ref"], w["dt"])
    b = s.solve(w["xu"], w["xs"], w["ref"], w["dt"])  # same inputs, warm dual + adapted rho: a different path
    assert n_mismatch(a["XU"], b["XU"]) > 0
    s.reset("dual")
    s.reset("rho")
    c = s.solve(w["xu"], w["xs"], w["ref"], w["dt"])
    assert n_mismatch(a["XU"], c["XU"]) == 0 and np.array_equal(a["pcg_iters"], c["pcg_iters"])
    s.set_batch("rho", np.array([0.5, 0.5], np.float32), set_default=True)
    s.reset("dual")
    d1 = s.solve(w["xu"], w["xs"], w["ref"], w["dt"])
    s.reset("dual")
    s.reset("rho")  # now restores 0.5
    d2 = s.solve(w["xu"], w["xs"], w["ref"], w["dt"])
    assert n_mismatch(d1["XU"], d2["XU"]) == 0


def test_padded_vectors_and_pcg_edge_cases(oracle_built):
    """items 6/9: leading/trailing nx zeros of lambda/gamma stay zero; converged solves skip PCG; zero rhs exits with 0 iterations."""
    be = Backend("oracle", "iiwa14", 8)
    w = make_config(1, B=2)
    k = be.stage_kkt(2, w["xu"] + 0.05, w["xs"], w["ref"], np.zeros((2, 6), np.float32), 0.01, w["params"])
    s = be.stage_schur(2, k, np.full(2, 0.01, np.float32))
    assert np.all(s["gamma"][:, :14] == 0) and np.all(s["gamma"][:, -14:] == 0)
    lam0 = np.zeros((2, 140), np.float32)
    lam, it = be.stage_pcg(2, s["S"], s["Pinv"], s["gamma"], lam0, np.full(2, 1e-6, np.float32), 100, kkt_conv=np.array([0, 1], np.int32))
    assert it[0] > 0 and it[1] == 0 and np.all(lam[1] == 0) and np.all(lam[:, :14] == 0) and np.all(lam[:, -14:] == 0)
    lam, it = be.stage_pcg(2, s["S"], s["Pinv"], np.zeros_like(s["gamma"]), lam0, np.full(2, 1e-6, np.float32), 100)
    assert (it == 0).all()
    # PCG actually solves S lambda = gamma
    lam, it = be.stage_pcg(2, s["S"], s["Pinv"], s["gamma"], lam0, np.full(2, 1e-10, np.float32), 300)
    S = s["S"][0].reshape(8, 14, 42).astype(np.float64)
    full = np.zeros((112, 140))
    for r in range(8):
        full[14 * r : 14 * r + 14, 14 * r : 14 * r + 42] = S[r]
    res = full @ lam[0].astype(np.float64) - s["gamma"][0, 14:-14]
    assert np.abs(res).max() < 5e-2 * np.abs(s["gamma"][0]).max()  # fp32 PCG on an ill-conditioned Schur complement


def test_indy7_barrier_variant(oracle_built):
    """item 10: indy7 uses barrier'·barrier' (full outer product on the joint block) instead of barrier''."""
    be = Backend("oracle", "indy7", 16)
    w = make_config(3, B=1, N=16)
    p = dict(w["params"], q_cost=0.0, q_lim_cost=0.5)
    k = be.stage_kkt(1, w["xu"], w["xs"], w["ref"], np.zeros((1, 6), np.float32), w["dt"], p)
    Q = k["Q"][0, 0].reshape(12, 12)[:6, :6].astype(np.float64)
    assert np.abs(Q - Q.T).max() < 1e-6 and np.abs(Q[0, 1]) > 0  # off-diagonal barrier coupling exists
    assert np.linalg.svd(Q, compute_uv=False)[1] < 1e-4 * np.linalg.svd(Q, compute_uv=False)[0]
