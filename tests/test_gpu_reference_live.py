"""Live three-way parity on the GPU box: the UNMODIFIED reference (IEEE build, oracle/_ref/libgref_iiwa14_N32_ieee.so, compiled from
/root/reference by oracle/build_ref.sh and shipped with the snapshot), the CPU oracle and the CUDA path solve the same batch -- every
trajectory bit, PCG iteration count and line-search step must agree.  (The merit values are compared with 1e-6: the reference sums the
knots with unordered float atomics, merit.cuh:88-91.)  Skipped when the reference library did not travel."""
from pathlib import Path

import numpy as np
import pytest

from conftest import n_mismatch
from gato_b200.workloads import make_config

pytestmark = pytest.mark.gpu
LIB = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "libgref_iiwa14_N32_ieee.so"


@pytest.mark.skipif(not LIB.exists(), reason="oracle/_ref/libgref_iiwa14_N32_ieee.so not built (needs /root/reference at build time)")
@pytest.mark.parametrize("iters,extra", [(1, {}), (4, {}), (3, dict(vel_lim_cost=0.002, ctrl_lim_cost=0.001, max_pcg_iters=200, pcg_tol=1e-4))])
def test_cuda_path_equals_reference_ieee_build_bit_for_bit(oracle_built, iters, extra):
    from gato_b200.native import GatoBackend
    from oracle.pyapi import Backend

    N, B = 32, 16
    w = make_config(2, B=B, N=N)
    p = dict(w["params"], max_sqp_iters=iters, **extra)
    rng = np.random.default_rng(17)
    xu = (w["xu"] + rng.normal(0, 0.05, w["xu"].shape)).astype(np.float32)
    r = Backend("ref", "iiwa14", N, "ieee").solver(B, p).solve(xu, w["xs"], w["ref"], w["dt"])
    o = Backend("oracle", "iiwa14", N).solver(B, p).solve(xu, w["xs"], w["ref"], w["dt"])
    g = GatoBackend("iiwa14", N).solver(B, p).solve(xu, w["xs"], w["ref"], w["dt"])
    for name, x in (("oracle", o), ("cuda", g)):
        assert n_mismatch(x["XU"], r["XU"]) == 0, f"{name}: trajectory bits differ from the reference"
        assert np.array_equal(x["pcg_iters"], r["pcg_iters"]) and np.array_equal(x["ls_step_size"], r["ls_step_size"]), name
        assert np.array_equal(x["sqp_iters"], r["sqp_iters"]) and np.array_equal(x["kkt_converged"], r["kkt_converged"]), name
        assert np.allclose(x["final_merit"], r["final_merit"], rtol=1e-6) and np.allclose(x["initial_merit"], r["initial_merit"], rtol=1e-6), name


LIB_INDY7 = LIB.parent / "libgref_indy7_N32_ieee.so"


@pytest.mark.skipif(not LIB_INDY7.exists(), reason="oracle/_ref/libgref_indy7_N32_ieee.so not built (needs /root/reference at build time)")
def test_indy7_stage_chain_three_way_live(oracle_built):
    """indy7 has no whole-solve comparison with the reference (its merit kernel faults on B200 through the reference's own launcher): the
    stage chain of the UNMODIFIED reference (IEEE build), the oracle and the CUDA path on the same inputs, bit-for-bit, with limit weights
    and wrenches on; the merit stage through the harness's extra-shared-memory launch of the same kernel (atomic sum: 2e-6)."""
    import os

    from gato_b200.native import GatoBackend
    from oracle.pyapi import Backend

    os.environ["GREF_MERIT_EXTRA_SMEM"] = "4096"
    plant, N, B = "indy7", 32, 16
    w = make_config(3, B=B, N=N)
    ref, orc, gpu = Backend("ref", plant, N, "ieee"), Backend("oracle", plant, N), GatoBackend(plant, N)
    rng = np.random.default_rng(9)
    xu = (w["xu"] + rng.normal(0, 0.05, w["xu"].shape)).astype(np.float32)
    p = dict(w["params"], vel_lim_cost=0.002, ctrl_lim_cost=0.001)
    fext = rng.normal(0, 2, (B, 6)).astype(np.float32)
    rho = np.full(B, p["rho"], np.float32)
    rho[1::2] = 1e-3
    mu = np.full(B, 10.0, np.float32)
    kr = ref.stage_kkt(B, xu, w["xs"], w["ref"], fext, w["dt"], p)
    for name, be in (("oracle", orc), ("cuda", gpu)):
        k = be.stage_kkt(B, xu, w["xs"], w["ref"], fext, w["dt"], p)
        for key in kr:
            assert n_mismatch(k[key], kr[key]) == 0, f"{name} kkt {key}"
    sr = ref.stage_schur(B, kr, rho)
    for name, be in (("oracle", orc), ("cuda", gpu)):
        s = be.stage_schur(B, kr, rho)
        for key in sr:
            assert n_mismatch(s[key], sr[key]) == 0, f"{name} schur {key}"
    lam0 = np.zeros((B, orc.d["vecp"]), np.float32)
    for eps, cap in ((1e-4, 200), (-1.0, 30)):
        lr, ir = ref.stage_pcg(B, sr["S"], sr["Pinv"], sr["gamma"], lam0, np.full(B, eps, np.float32), cap)
        for name, be in (("oracle", orc), ("cuda", gpu)):
            lx, ix = be.stage_pcg(B, sr["S"], sr["Pinv"], sr["gamma"], lam0, np.full(B, eps, np.float32), cap)
            assert np.array_equal(ix, ir) and n_mismatch(lx, lr) == 0, f"{name} pcg eps={eps}"
    dr = ref.stage_dz(B, lr, sr["Qinv"], sr["Rinv"], kr["q"], kr["r"], kr["A"], kr["Bm"])
    for name, be in (("oracle", orc), ("cuda", gpu)):
        dx = be.stage_dz(B, lr, sr["Qinv"], sr["Rinv"], kr["q"], kr["r"], kr["A"], kr["Bm"])
        for a, b in zip(dx, dr):
            assert n_mismatch(a, b) == 0, f"{name} dz"
    for na in (1, 8):
        mr = ref.stage_merit(B, xu, dr[0], w["xs"], w["ref"], mu, fext, w["dt"], p, na)
        for name, be in (("oracle", orc), ("cuda", gpu)):
            mx = be.stage_merit(B, xu, dr[0], w["xs"], w["ref"], mu, fext, w["dt"], p, na)
            assert np.allclose(mx, mr, rtol=2e-6), f"{name} merit x{na}"
    m8 = orc.stage_merit(B, xu, dr[0], w["xs"], w["ref"], mu, fext, w["dt"], p, 8)
    mi = orc.stage_merit(B, xu, np.zeros_like(dr[0]), w["xs"], w["ref"], mu, fext, w["dt"], p, 1)[:, 0].copy()
    for adapt in (1, 0):
        lsr = ref.stage_linesearch(B, xu, dr[0], m8, mi, rho, np.ones(B, np.float32), adapt)
        for name, be in (("oracle", orc), ("cuda", gpu)):
            lsx = be.stage_linesearch(B, xu, dr[0], m8, mi, rho, np.ones(B, np.float32), adapt)
            for key in lsr:
                assert n_mismatch(lsx[key], lsr[key]) == 0, f"{name} line search {key}"
