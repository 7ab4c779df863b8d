"""The product's per-item math (gato_b200/csrc/rbd.cuh, items.cuh) compiled for the HOST must agree bit-for-bit with the oracle
(on the GPU the same code runs one item per thread; sin/cos/log are CUDA's, of which the oracle holds bit-exact restatements)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT, n_mismatch
from gato_b200.workloads import make_config
from oracle.pyapi import PLANT_ID, Backend, cost7, f32p


@pytest.fixture(scope="module")
def hostlib(oracle_built):
    d = ROOT / "tests" / "host"
    subprocess.check_call(["make", "-C", str(d)], stdout=subprocess.DEVNULL)
    lib = C.CDLL(str(d / "librbd_host_check.so"))
    lib.hostchk_dyn_dump.argtypes = [C.c_int, C.c_int] + [f32p] * 7
    lib.hostchk_stage_kkt.argtypes = [C.c_int] * 3 + [f32p] * 4 + [C.c_float, f32p] + [f32p] * 7
    lib.hostchk_stage_merit.argtypes = [C.c_int] * 3 + [f32p] * 6 + [C.c_float, f32p, C.c_int, f32p]
    lib.hostchk_stage_merit_split.argtypes = [C.c_int] * 3 + [f32p] * 6 + [C.c_float, f32p, C.c_int, f32p]
    lib.hostchk_stage_kkt_kernel_paths.argtypes = [C.c_int] * 4 + [f32p] * 2 + [C.c_float] + [f32p] * 3
    return lib


@pytest.mark.parametrize("plant,N,cfg", [("iiwa14", 8, 1), ("iiwa14", 32, 2), ("indy7", 32, 3), ("indy7", 16, 3)])
def test_product_item_math_bit_exact(hostlib, plant, N, cfg):
    be = Backend("oracle", plant, N)
    nq, pid = be.d["nq"], PLANT_ID[plant]
    rng = np.random.default_rng(3)
    n = 24
    x = rng.uniform(-2, 2, (n, 2 * nq)).astype(np.float32)
    u = rng.uniform(-20, 20, (n, nq)).astype(np.float32)
    fe = rng.normal(0, 3, (n, 6)).astype(np.float32)
    fe[:6] = 0
    o = be.dyn_dump(x, u, fe)
    got = {k: np.zeros_like(v) for k, v in o.items()}
    hostlib.hostchk_dyn_dump(pid, n, x.ravel(), u.ravel(), fe.ravel(), got["qdd"].ravel(), got["dqdd"].ravel(), got["ee"].ravel(), got["dee"].ravel())
    for k in o:
        assert n_mismatch(got[k], o[k]) == 0, k
    B = 3
    w = make_config(cfg, B=B, N=N)
    xu = w["xu"] + rng.normal(0, 0.1, w["xu"].shape).astype(np.float32)
    fext = rng.normal(0, 2, (B, 6)).astype(np.float32)
    p = dict(w["params"], vel_lim_cost=0.003, ctrl_lim_cost=0.002)  # exercise every barrier term
    k0 = be.stage_kkt(B, xu, w["xs"], w["ref"], fext, w["dt"], p)
    k1 = {k: np.zeros_like(v) for k, v in k0.items()}
    hostlib.hostchk_stage_kkt(pid, N, B, xu.ravel(), w["xs"].ravel(), w["ref"].ravel(), fext.ravel(), np.float32(w["dt"]), cost7(p), *[k1[k].reshape(-1) for k in ("Q", "R", "q", "r", "A", "Bm", "c")])
    for k in k0:
        assert n_mismatch(k1[k], k0[k]) == 0, k
    # the code paths the kernels run: k_kkt's rolled halves (1) and k_kkt_fine's prologue + base + columns (2)
    for variant in (1, 2):
        k2 = {k: np.zeros_like(k0[k]) for k in ("A", "Bm", "c")}
        hostlib.hostchk_stage_kkt_kernel_paths(pid, variant, N, B, xu.ravel(), fext.ravel(), np.float32(w["dt"]), *[k2[k].reshape(-1) for k in ("A", "Bm", "c")])
        assert n_mismatch(k2["A"][:, :-1], k0["A"][:, :-1]) == 0 and n_mismatch(k2["Bm"][:, :-1], k0["Bm"][:, :-1]) == 0, variant
        assert n_mismatch(k2["c"][:, 1:], k0["c"][:, 1:]) == 0, variant
    dz = rng.normal(0, 0.05, xu.shape).astype(np.float32)
    mu = np.full(B, 10, np.float32)
    # every barrier weight non-zero, then the default weights (zero velocity / control barriers are skipped by the product, not by the oracle)
    for na, pm in ((1, p), (8, p), (8, dict(w["params"])), (8, dict(w["params"], q_lim_cost=0.0))):
        m0 = be.stage_merit(B, xu, dz, w["xs"], w["ref"], mu, fext, w["dt"], pm, na)
        m1 = np.zeros_like(m0)
        hostlib.hostchk_stage_merit(pid, N, B, xu.ravel(), dz.ravel(), w["xs"].ravel(), w["ref"].ravel(), mu, fext.ravel(), np.float32(w["dt"]), cost7(pm), na, m1.reshape(-1))
        assert n_mismatch(m1, m0) == 0
        m2 = np.zeros_like(m0)
        hostlib.hostchk_stage_merit_split(pid, N, B, xu.ravel(), dz.ravel(), w["xs"].ravel(), w["ref"].ravel(), mu, fext.ravel(), np.float32(w["dt"]), cost7(pm), na, m2.reshape(-1))
        assert n_mismatch(m2, m0) == 0


def test_linearisation_matches_finite_differences(oracle_built):
    """A_k, B_k are the Jacobians of the integrator map x_{k+1} = f(x_k,u_k) whose defect is c_{k+1} (integrator.cuh:59,143-184)."""
    be = Backend("oracle", "iiwa14", 8)
    d = be.d
    nx, nu = d["nx"], d["nu"]
    w = make_config(1, B=2)
    rng = np.random.default_rng(0)
    xu = w["xu"] + rng.normal(0, 0.2, w["xu"].shape).astype(np.float32)
    fext = rng.normal(0, 1.0, (2, 6)).astype(np.float32)
    k0 = be.stage_kkt(2, xu, w["xs"], w["ref"], fext, 0.01, w["params"])
    eps = 1e-3
    for j in range(nx + nu):
        xp, xm = xu.copy(), xu.copy()
        xp[:, j] += eps
        xm[:, j] -= eps
        cp = be.stage_kkt(2, xp, w["xs"], w["ref"], fext, 0.01, w["params"])["c"][:, 1]
        cm = be.stage_kkt(2, xm, w["xs"], w["ref"], fext, 0.01, w["params"])["c"][:, 1]
        fd = -(cp - cm) / (2 * eps)
        an = k0["A"][:, 0].reshape(2, nx, nx)[:, j, :] if j < nx else k0["Bm"][:, 0].reshape(2, nu, nx)[:, j - nx, :]
        assert np.abs(fd - an).max() / max(1e-3, np.abs(an).max()) < 2e-2
