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
