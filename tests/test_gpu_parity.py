"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI (gato_b200/lib/libgato_b200.so).

Bar: the CUDA path must reproduce the CPU oracle BIT-FOR-BIT (floats and integers) on every stage and on whole solves — the
kernels evaluate the same expression trees (explicit fmaf, -fmad=false, reference reduction trees).  Against the reference's own
outputs (tests/golden/, fast-math build) the tolerance is 1e-4 relative on trajectories with exact integer outcomes where the
oracle tests established them.
"""
import numpy as np
import pytest

from conftest import load_golden, n_mismatch, params_of, rel_err
from gato_b200.workloads import DEFAULT_SOLVER_PARAMS, make_config

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def backends(oracle_built):
    from gato_b200.native import GatoBackend
    from oracle.pyapi import Backend

    def get(plant, N):
        return Backend("oracle", plant, N), GatoBackend(plant, N)

    return get


def _inputs(cfg, B, N, seed=11):
    rng = np.random.default_rng(seed)
    w = make_config(cfg, B=B, N=N)
    xu = w["xu"] + rng.normal(0, 0.05, w["xu"].shape).astype(np.float32)
    fext = rng.normal(0, 2, (B, 6)).astype(np.float32)
    fext[0] = 0
    return w, xu, fext


@pytest.mark.parametrize("plant,N,cfg,B", [("iiwa14", 8, 1, 3), ("iiwa14", 32, 2, 5), ("indy7", 32, 3, 4), ("indy7", 16, 3, 2), ("iiwa14", 16, 2, 33), ("iiwa14", 128, 4, 2), ("indy7", 64, 3, 2), ("iiwa14", 9, 2, 3), ("indy7", 7, 3, 2), ("iiwa14", 3, 1, 2)])
def test_every_stage_bit_exact_vs_oracle(backends, plant, N, cfg, B):
    o, g = backends(plant, N)
    d = o.d
    w, xu, fext = _inputs(cfg, B, N)
    p = dict(w["params"], vel_lim_cost=0.002, ctrl_lim_cost=0.001)
    rho = np.full(B, p["rho"], np.float32)
    rho[1::2] = 1e-3
    mu = np.full(B, 10, np.float32)
    ko, kg = o.stage_kkt(B, xu, w["xs"], w["ref"], fext, w["dt"], p), g.stage_kkt(B, xu, w["xs"], w["ref"], fext, w["dt"], p)
    for k in ko:
        assert n_mismatch(kg[k], ko[k]) == 0, f"kkt {k}"
    so, sg = o.stage_schur(B, ko, rho), g.stage_schur(B, ko, rho)
    for k in so:
        assert n_mismatch(sg[k], so[k]) == 0, f"schur {k}"
    lam0 = np.zeros((B, d["vecp"]), np.float32)
    for eps, cap in ((1e-4, 200), (-1.0, 20), (1e-4, 0)):
        lo, io = o.stage_pcg(B, so["S"], so["Pinv"], so["gamma"], lam0, np.full(B, eps, np.float32), cap)
        lg, ig = g.stage_pcg(B, so["S"], so["Pinv"], so["gamma"], lam0, np.full(B, eps, np.float32), cap)
        assert np.array_equal(ig, io) and n_mismatch(lg, lo) == 0
    conv = np.zeros(B, np.int32)
    conv[0] = 1  # flagged solves skip PCG (pcg.cuh:29)
    lo, io = o.stage_pcg(B, so["S"], so["Pinv"], so["gamma"], lam0, np.full(B, 1e-4, np.float32), 200, conv)
    lg, ig = g.stage_pcg(B, so["S"], so["Pinv"], so["gamma"], lam0, np.full(B, 1e-4, np.float32), 200, conv)
    assert ig[0] == 0 and np.array_equal(ig, io) and n_mismatch(lg, lo) == 0
    dzo = o.stage_dz(B, lo, so["Qinv"], so["Rinv"], ko["q"], ko["r"], ko["A"], ko["Bm"])
    dzg = g.stage_dz(B, lo, so["Qinv"], so["Rinv"], ko["q"], ko["r"], ko["A"], ko["Bm"])
    for a, b in zip(dzg, dzo):
        assert n_mismatch(a, b) == 0
    for na in (1, 8):
        mo = o.stage_merit(B, xu, dzo[0], w["xs"], w["ref"], mu, fext, w["dt"], p, na)
        mg = g.stage_merit(B, xu, dzo[0], w["xs"], w["ref"], mu, fext, w["dt"], p, na)
        assert n_mismatch(mg, mo) == 0, f"merit x{na}"
    mi = o.stage_merit(B, xu, np.zeros_like(dzo[0]), w["xs"], w["ref"], mu, fext, w["dt"], p, 1)[:, 0].copy()
    mi[-1] = -1e30
    for adapt in (1, 0):
        lso = o.stage_linesearch(B, xu, dzo[0], mo, mi, rho, np.ones(B, np.float32), adapt)
        lsg = g.stage_linesearch(B, xu, dzo[0], mo, mi, rho, np.ones(B, np.float32), adapt)
        for k in lso:
            assert n_mismatch(lsg[k], lso[k]) == 0, f"line search {k}"


@pytest.mark.parametrize("plant,N,cfg,B", [("iiwa14", 8, 1, 1), ("iiwa14", 32, 2, 16), ("indy7", 32, 3, 16), ("iiwa14", 32, 5, 16), ("iiwa14", 128, 4, 4), ("iiwa14", 9, 2, 5), ("indy7", 33, 3, 3), ("iiwa14", 34, 2, 3), ("indy7", 40, 3, 2)])
def test_whole_solve_bit_exact_vs_oracle(backends, plant, N, cfg, B):
    o, g = backends(plant, N)
    w = make_config(cfg, B=B, N=N)
    for p in (w["params"], dict(DEFAULT_SOLVER_PARAMS, dt=float(w["dt"]), max_sqp_iters=3), dict(w["params"], solve_ratio=0.5, max_pcg_iters=3)):
        so, sg = o.solver(B, p), g.solver(B, p)
        for key in ("rho", "mu"):
            if key in w["extra"]:
                so.set_batch(key, w["extra"][key])
                sg.set_batch(key, w["extra"][key])
        xin = w["xu"]
        for rep in range(2):  # the second solve runs on the persisted lambda / rho
            ro, rg = so.solve(xin, w["xs"], w["ref"], w["dt"]), sg.solve(xin, w["xs"], w["ref"], w["dt"])
            assert rg["n_pcg"] == ro["n_pcg"] and rg["n_ls"] == ro["n_ls"]
            for k in ("pcg_iters", "sqp_iters", "kkt_converged"):
                assert np.array_equal(rg[k], ro[k]), k
            for k in ("XU", "ls_step_size", "ls_min_merit", "final_merit", "initial_merit"):
                assert n_mismatch(rg[k], ro[k]) == 0, k
            xin = ro["XU"]
        fe = np.random.default_rng(5).normal(0, 2, (B, 6)).astype(np.float32)
        so.set_batch("f_ext", fe)
        sg.set_batch("f_ext", fe)
        uk = np.random.default_rng(6).uniform(-5, 5, o.d["nq"]).astype(np.float32)
        assert n_mismatch(sg.sim_forward(w["xs"][0], uk, w["dt"]), so.sim_forward(w["xs"][0], uk, w["dt"])) == 0
        so.reset("dual"), sg.reset("dual"), so.reset("rho"), sg.reset("rho")
        ro, rg = so.solve(w["xu"], w["xs"], w["ref"], w["dt"]), sg.solve(w["xu"], w["xs"], w["ref"], w["dt"])
        assert n_mismatch(rg["XU"], ro["XU"]) == 0  # wrench hypotheses enter KKT and merit
        so.close(), sg.close()


def test_edge_cases(backends):
    o, g = backends("iiwa14", 8)
    w = make_config(1, B=4)
    for p in (dict(w["params"], max_sqp_iters=3, max_pcg_iters=0), dict(w["params"], max_sqp_iters=3, solve_ratio=0.0), dict(w["params"], max_sqp_iters=1, mu=0.0)):
        ro, rg = o.solver(4, p).solve(w["xu"], w["xs"], w["ref"], w["dt"]), g.solver(4, p).solve(w["xu"], w["xs"], w["ref"], w["dt"])
        assert (rg["n_pcg"], rg["n_ls"]) == (ro["n_pcg"], ro["n_ls"])
        assert np.array_equal(rg["kkt_converged"], ro["kkt_converged"]) and np.array_equal(rg["sqp_iters"], ro["sqp_iters"])
        assert n_mismatch(rg["XU"], ro["XU"]) == 0
    # rho adaptation off + NaN in one solve's input: the NaN solve fails its line search, the others are unaffected
    xu = w["xu"].copy()
    xu[2, 5] = np.nan
    so, sg = o.solver(4, w["params"]), g.solver(4, w["params"])
    so.set_rho_adaptation(False), sg.set_rho_adaptation(False)
    ro, rg = so.solve(xu, w["xs"], w["ref"], w["dt"]), sg.solve(xu, w["xs"], w["ref"], w["dt"])
    assert np.array_equal(rg["ls_step_size"], ro["ls_step_size"]) and (rg["ls_step_size"][:, 2] == -1).all()
    for b in (0, 1, 3):
        assert n_mismatch(rg["XU"][b], ro["XU"][b]) == 0


def test_against_reference_golden_solves(backends):
    """Reference's own whole-solve outputs (fast-math build on B200): trajectories within 1e-4, integer outcomes exact (1 SQP iteration)."""
    from gato_b200.native import GatoBackend

    for N, Bs in ((8, 16), (32, 16), (32, 128)):
        G = load_golden("iiwa14", N, "fast")
        g = GatoBackend("iiwa14", N)
        base = f"solve_B{Bs}_"
        if base + "xu" in G.files:
            xu, xs, ref = G[base + "xu"], G[base + "xs"], G[base + "ref"]
        else:
            w = make_config(2, B=Bs, N=N)
            xu, xs, ref = w["xu"], w["xs"], w["ref"]
            assert abs(float(G[base + "input_checksum"]) - (xu.astype(np.float64).sum() + ref.astype(np.float64).sum())) < 1e-6
        r = g.solver(Bs, params_of(G, base + "d_params")).solve(xu, xs, ref, float(G[base + "dt"]))
        n = G[base + "d_XU"].shape[0]
        same_pcg = (r["pcg_iters"] == G[base + "d_pcg_iters"]).mean()
        same_step = (r["ls_step_size"] == G[base + "d_ls_step_size"]).mean()
        assert same_pcg >= 0.9 and same_step >= 0.95, (same_pcg, same_step)  # threshold crossings vs a fast-math build
        ok = (r["ls_step_size"][0, :n] == G[base + "d_ls_step_size"][0, :n])
        err = np.abs(r["XU"][:n] - G[base + "d_XU"]).max(axis=1) / np.abs(G[base + "d_XU"]).max(axis=1)
        assert np.median(err[ok]) < 1e-4 and (err[ok] < 1e-2).all()
        assert rel_err(r["initial_merit"], G[base + "d_initial_merit"]) < 1e-5


def test_full_size_properties(backends):
    """BASELINE.json headline shape (iiwa14, N=32, B=512): size-independent properties instead of an oracle run."""
    from gato_b200.native import GatoBackend

    g = GatoBackend("iiwa14", 32)
    w = make_config("bench", B=512)
    s = g.solver(512, w["params"])
    a = s.solve(w["xu"], w["xs"], w["ref"], w["dt"])
    assert (a["pcg_iters"] == 50).all() and a["n_pcg"] == 4 and a["n_ls"] == 4  # fixed caps: every PCG runs exactly max_pcg_iters
    assert np.isfinite(a["XU"]).all() and (a["final_merit"] <= a["initial_merit"] + 1e-3).all()  # accepted steps never increase the merit
    s.reset("dual"), s.reset("rho")
    b = s.solve(w["xu"], w["xs"], w["ref"], w["dt"])
    assert n_mismatch(a["XU"], b["XU"]) == 0 and np.array_equal(a["ls_step_size"], b["ls_step_size"])  # deterministic
    # batch independence: a solve gives the same bits alone, in a small batch, and in a permuted batch (=> sharding across GPUs is exact)
    perm = np.random.default_rng(0).permutation(512)
    s.reset("dual"), s.reset("rho")
    c = s.solve(w["xu"][perm], w["xs"][perm], w["ref"][perm], w["dt"])
    assert n_mismatch(c["XU"], a["XU"][perm]) == 0
    s16 = g.solver(16, w["params"])
    d = s16.solve(w["xu"][100:116], w["xs"][100:116], w["ref"][100:116], w["dt"])
    assert n_mismatch(d["XU"], a["XU"][100:116]) == 0 and np.array_equal(d["ls_step_size"], a["ls_step_size"][:, 100:116])
    assert s.kernel_launches() >= 3 * 17


def test_device_pointer_entry_and_async(backends):
    """gato_solve on caller-owned device memory (the BSQP::solve path) equals the host-buffer path."""
    import torch

    from gato_b200.native import GatoBackend

    g = GatoBackend("iiwa14", 32)
    w = make_config(2, B=32)
    ref_out = g.solver(32, w["params"]).solve(w["xu"], w["xs"], w["ref"], w["dt"])
    s = g.solver(32, w["params"])
    xu, xs, rf = torch.from_numpy(w["xu"]).cuda(), torch.from_numpy(w["xs"]).cuda(), torch.from_numpy(w["ref"]).cuda()
    torch.cuda.synchronize()
    st = s.solve_device(xu.data_ptr(), xs.data_ptr(), rf.data_ptr(), w["dt"])
    assert n_mismatch(xu.cpu().numpy(), ref_out["XU"]) == 0 and np.array_equal(st["pcg_iters"], ref_out["pcg_iters"])
    assert st["device_time_ms"] > 0


def test_kernel_timing_instrumentation_is_transparent():
    """gato_set_kernel_timing / gato_get_kernel_times / gato_get_launch_times: event-per-launch timing does not change results, and the
    Python mirror's opt-in `pcg_times_us` carries one measured PCG time per line search."""
    from gato_b200.native import GatoBackend

    g = GatoBackend("iiwa14", 32)
    w = make_config(2, B=32)
    a = g.solver(32, w["params"]).solve(w["xu"], w["xs"], w["ref"], w["dt"])
    s = g.solver(32, w["params"])
    s.set_kernel_timing(True)
    b = s.solve(w["xu"], w["xs"], w["ref"], w["dt"])
    assert n_mismatch(a["XU"], b["XU"]) == 0 and np.array_equal(a["pcg_iters"], b["pcg_iters"])
    kt, lt = s.kernel_times(), s.launch_times()
    n_it = int(w["params"]["max_sqp_iters"])
    assert [kt[k][1] for k in ("k_kkt", "k_schur", "k_pcg", "k_merit_ls<8>", "k_merit_ls<1>")] == [n_it, n_it, n_it, n_it, 1]
    assert all(ms > 0 for ms, _ in kt.values()) and len(lt) == 4 * n_it + 1 and lt[0][0] == "k_merit_ls<1>" and lt[1][0] == "k_kkt"
    assert abs(sum(ms for _, ms in lt) - sum(ms for ms, _ in kt.values())) < 1e-3
    from gato_b200.bsqp import bsqpN32_iiwa14 as mod
    from gato_b200.native import PARAM_ORDER

    p = w["params"]

    slv = mod.BSQP_32_float(*[p[k] for k in PARAM_ORDER])
    slv.set_kernel_timing(True)
    r = slv.solve(w["xu"], float(w["dt"]), w["xs"], w["ref"])
    assert r["pcg_times_us"].shape == (r["ls_num_iters"],) and (r["pcg_times_us"] > 0).all()
    assert n_mismatch(r["XU"], a["XU"]) == 0


@pytest.mark.parametrize("cfg,subset", [(3, 6), (4, 3), (5, 6)])
def test_baseline_configs_at_full_size(backends, cfg, subset):
    """BASELINE.json configs 3 (indy7, N=32, B=512), 4 (iiwa14, N=128, B=1024) and 5 (iiwa14, N=32, B=1024 per GPU, per-solve rho / mu)
    at their full sizes: determinism, merit never increased by accepted steps, and -- because solves are independent (checked bit-for-bit
    in test_full_size_properties) -- a handful of solves out of the full batch compared bit-for-bit with the CPU oracle run on just those."""
    w = make_config(cfg)
    B, N, plant = w["B"], w["N"], w["plant"]
    o, g = backends(plant, N)

    def setup(s, sl):
        for key in ("rho", "mu"):
            if key in w["extra"]:
                s.set_batch(key, w["extra"][key][sl])

    sg = g.solver(B, w["params"])
    setup(sg, slice(None))
    a = sg.solve(w["xu"], w["xs"], w["ref"], w["dt"])
    assert np.isfinite(a["XU"]).all() and (a["final_merit"] <= a["initial_merit"] * (1 + 1e-6) + 1e-6).all()
    sg.reset("dual"), sg.reset("rho")
    b = sg.solve(w["xu"], w["xs"], w["ref"], w["dt"])
    assert n_mismatch(a["XU"], b["XU"]) == 0 and np.array_equal(a["pcg_iters"], b["pcg_iters"])
    idx = np.linspace(0, B - 1, subset).astype(int)
    p1 = dict(w["params"], solve_ratio=2.0)  # the early-exit count is per batch: disable it for the subset comparison if it never fired
    if a["n_pcg"] == int(w["params"]["max_sqp_iters"]) and a["n_ls"] == a["n_pcg"]:
        so = o.solver(len(idx), p1)
        setup(so, idx)
        ro = so.solve(w["xu"][idx], w["xs"][idx], w["ref"][idx], w["dt"])
        assert n_mismatch(ro["XU"], a["XU"][idx]) == 0
        assert np.array_equal(ro["pcg_iters"], a["pcg_iters"][:, idx]) and np.array_equal(ro["ls_step_size"], a["ls_step_size"][:, idx])


def test_randomized_whole_solves_bit_exact(backends):
    """Randomised sweep (fixed seeds): horizon, batch, plant, cost weights, rho / mu / f_ext batches and large-amplitude states (sin/cos range
    reduction, barrier logs near the joint limits, tiny pivots in the elimination) -- whole solves bit-for-bit against the oracle."""
    for seed in range(12):
        rng = np.random.default_rng(1000 + seed)
        plant = "iiwa14" if rng.random() < 0.6 else "indy7"
        N = int(rng.integers(3, 41))
        B = int(rng.integers(1, 10))
        o, g = backends(plant, N)
        nq = o.d["nq"]
        w = make_config(2 if plant == "iiwa14" else 3, B=B, N=N)
        p = dict(w["params"], max_sqp_iters=int(rng.integers(1, 4)), max_pcg_iters=int(rng.integers(1, 60)), pcg_tol=float(10 ** rng.uniform(-6, -2)),
                 mu=float(10 ** rng.uniform(-1, 2)), q_cost=float(10 ** rng.uniform(-1, 1)), qd_cost=float(10 ** rng.uniform(-3, -1)), u_cost=float(10 ** rng.uniform(-7, -4)),
                 N_cost=float(10 ** rng.uniform(0, 2)), q_lim_cost=float(rng.choice([0.0, 0.01, 0.1])), vel_lim_cost=float(rng.choice([0.0, 0.002])),
                 ctrl_lim_cost=float(rng.choice([0.0, 0.001])), rho=float(10 ** rng.uniform(-6, 0)), solve_ratio=float(rng.choice([1.0, 0.5])))
        scale = float(rng.choice([0.05, 0.5, 2.0]))
        xu = (w["xu"] + rng.normal(0, scale, w["xu"].shape)).astype(np.float32)
        xs = (w["xs"] + rng.normal(0, 0.1, w["xs"].shape)).astype(np.float32)
        so, sg = o.solver(B, p), g.solver(B, p)
        fext = rng.normal(0, 3, (B, 6)).astype(np.float32) if seed % 2 else np.zeros((B, 6), np.float32)
        mu = rng.choice([1.0, 10.0, 100.0], B).astype(np.float32)
        for s in (so, sg):
            s.set_batch("f_ext", fext)
            s.set_batch("rho", np.logspace(-6, 0, B).astype(np.float32), True)
            s.set_batch("mu", mu)
        ro, rg = so.solve(xu, xs, w["ref"], w["dt"]), sg.solve(xu, xs, w["ref"], w["dt"])
        tag = f"seed {seed}: {plant} N={N} B={B}"
        assert (rg["n_pcg"], rg["n_ls"]) == (ro["n_pcg"], ro["n_ls"]), tag
        for k in ("pcg_iters", "sqp_iters", "kkt_converged"):
            assert np.array_equal(rg[k], ro[k]), f"{tag} {k}"
        for k in ("XU", "ls_step_size", "ls_min_merit", "final_merit", "initial_merit"):
            assert n_mismatch(rg[k], ro[k]) == 0, f"{tag} {k}"


def test_cabi_error_paths_and_concurrent_solvers():
    """C-ABI behaviour on a GPU box: unsupported horizons and bad arguments return error codes with a message (the reference aborts or ignores
    CUDA errors, cuda.cuh:7-19); two solvers driven from two host threads on their own streams give the same bits as one after the other."""
    import ctypes as C
    import threading

    from gato_b200 import native

    lib = native.load()
    h = C.c_void_p()
    prm = native.make_params(dict(DEFAULT_SOLVER_PARAMS, dt=0.01))
    assert lib.gato_create(C.byref(h), 1, 400, 4, 0, None, C.byref(prm)) == -3  # GATO_ERR_UNSUPPORTED: (N+2)*nx > 4096
    assert b"knot_points too large" in lib.gato_last_error(None)
    assert lib.gato_create(C.byref(h), 1, 32, 4, 99, None, C.byref(prm)) == -2  # GATO_ERR_CUDA: no such device
    assert lib.gato_create(C.byref(h), 1, 32, 4, 0, None, C.byref(prm)) == 0
    assert lib.gato_solve(h, None, None, None, C.c_float(0.01), None) == -1  # GATO_ERR_ARG
    assert lib.gato_reset(h, 7) == -1 and b"unknown reset field" in lib.gato_last_error(h)
    out = native.GatoMpcOut()
    x = np.zeros(14, np.float32)
    ref = np.zeros(6 * 32, np.float32)
    assert lib.gato_mpc_step(h, x, ref, None, None, C.c_float(0.0), C.c_float(0.01), 0, C.byref(out), None) == -1  # no warm start set
    assert b"gato_mpc_set_warm_start" in lib.gato_last_error(h)
    lib.gato_destroy(h)

    w = make_config(2, B=64)
    serial = [native.Solver(w["plant"], w["N"], 64, w["params"], device=0).solve(w["xu"], w["xs"], w["ref"], w["dt"]) for _ in range(2)]
    results = [None, None]

    def run(i):
        s = native.Solver(w["plant"], w["N"], 64, w["params"], device=0)
        for _ in range(3):
            s.reset("dual"), s.reset("rho")
            results[i] = s.solve(w["xu"], w["xs"], w["ref"], w["dt"])

    ts = [threading.Thread(target=run, args=(i,)) for i in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for i in range(2):
        assert n_mismatch(results[i]["XU"], serial[0]["XU"]) == 0 and np.array_equal(results[i]["pcg_iters"], serial[1]["pcg_iters"])
