"""Run-time robot models (include/gato_b200.h: gato_model_*), checked without a GPU:
the model file round trip, the registration checks, and the PRODUCT's table-driven per-item math (gato_b200/csrc/rbd_rt.cuh, items_rt.cuh --
what the kernels of a registered model execute) compiled for the host against the CPU oracle, bit for bit -- for the two compiled robots'
own tables (so table-driven == compiled == oracle) and for robots that exist only as data."""
import ctypes as C
import subprocess

import numpy as np
import pytest

from conftest import ROOT, n_mismatch
from gato_b200 import native
from gato_b200.workloads import make_config
from oracle import pyapi
from oracle.pyapi import Backend, cost7, f32p


def derived_robot(base, name, style, seed):
    """A robot that exists only as data: the base robot's kinematic structure with other link offsets, inertias and limits."""
    m = native.Model.builtin(base)
    rng = np.random.default_rng(seed)
    m.raw.name = name.encode()
    m.raw.style = style
    m.array("I")[:] *= rng.uniform(0.8, 1.3, (m.nq, 1))  # heavier / lighter links (a scaled spatial inertia stays positive definite)
    X = m.array("X")
    off = (np.abs(X) > 0) & (np.abs(X) < 1)  # the link offsets in the bottom-left blocks
    X[off] *= rng.uniform(0.9, 1.1, off.sum())
    for i in range(m.raw.n_x_trig):
        if abs(m.raw.x_trig[i].coef) != 1.0:
            m.raw.x_trig[i].coef *= float(rng.uniform(0.9, 1.1))
    Xh = m.array("Xhom")
    offh = (np.abs(Xh) > 0) & (np.abs(Xh) < 1)
    Xh[offh] *= rng.uniform(0.9, 1.1, offh.sum())
    for f in ("joint_limit", "vel_limit", "ctrl_limit"):
        m.array(f)[:] *= rng.uniform(0.9, 1.2, m.nq)
    return m


def eight_joint_robot(name="chain8", seed=21):
    """An 8-joint robot (no reference robot has eight joints): iiwa14 with one more link, a scaled copy of its sixth, appended at the tip."""
    base = derived_robot("iiwa14", name, 1, seed)
    m = native.Model()
    m.raw.name, m.raw.nq, m.raw.style = name.encode(), 8, 1
    src = 5  # the link that is copied to the tip
    for f, per in (("X", 36), ("I", 36), ("Xhom", 16), ("dXhom", 16)):
        a = np.ctypeslib.as_array(getattr(m.raw, f))
        b = np.ctypeslib.as_array(getattr(base.raw, f))
        a[: 7 * per] = b[: 7 * per]
        a[7 * per : 8 * per] = b[src * per : (src + 1) * per] * (0.8 if f == "I" else 1.0)
    for f in ("joint_limit", "vel_limit", "ctrl_limit"):
        a, b = np.ctypeslib.as_array(getattr(m.raw, f)), np.ctypeslib.as_array(getattr(base.raw, f))
        a[:7], a[7] = b[:7], b[src]
    for which, per in (("x", 36), ("xh", 16), ("dxh", 16)):
        old, n = getattr(base.raw, f"{which}_trig"), int(getattr(base.raw, f"n_{which}_trig"))
        new, cnt = getattr(m.raw, f"{which}_trig"), 0
        def put(idx, k_joint, is_cos, coef):
            nonlocal cnt
            new[cnt].idx, new[cnt].k, new[cnt].coef = idx, k_joint + (8 if is_cos else 0), coef
            cnt += 1
        for i in range(n):  # t[k < nq] = sin(q_k), t[k >= nq] = cos(q_{k - nq}): re-encode for nq = 8
            e = old[i]
            put(e.idx, e.k % 7, e.k >= 7, e.coef)
        for i in range(n):
            e = old[i]
            if e.idx // per == src:
                put(e.idx + (7 - src) * per, 7, e.k >= 7, e.coef)
        setattr(m.raw, f"n_{which}_trig", cnt)
    return m


def random_workload(nq, N, B, seed=0):
    """inputs of a solve for a robot the synthetic BASELINE workloads do not know (they are tied to the two reference robots)"""
    rng = np.random.default_rng(seed)
    nx, nu = 2 * nq, nq
    xu = np.zeros((B, N, nx + nu), np.float32)
    xu[..., :nq] = rng.uniform(-0.8, 0.8, (B, 1, nq)) + rng.normal(0, 0.05, (B, N, nq))
    xu[..., nq:nx] = rng.normal(0, 0.2, (B, N, nq))
    xu[..., nx:] = rng.normal(0, 2.0, (B, N, nu))
    xu = xu.reshape(B, -1)[:, : (nx + nu) * N - nu].copy()
    ref = np.zeros((B, N, 6), np.float32)
    ref[..., :3] = rng.uniform(-0.4, 0.4, (B, 1, 3)) + rng.normal(0, 0.02, (B, N, 3))
    return dict(xu=xu, xs=xu[:, :nx].copy(), ref=ref.reshape(B, -1), dt=0.01)


@pytest.fixture(scope="module")
def rtlib(oracle_built):
    d = ROOT / "tests" / "host"
    subprocess.check_call(["make", "-C", str(d)], stdout=subprocess.DEVNULL)
    lib = C.CDLL(str(d / "librt_host_check.so"))
    lib.hostchk_rt_set_model.argtypes = [C.POINTER(native.GatoModel), C.c_char_p]
    lib.hostchk_rt_stage_kkt.argtypes = [C.c_int] * 3 + [f32p] * 4 + [C.c_float, f32p] + [f32p] * 7
    lib.hostchk_rt_stage_merit.argtypes = [C.c_int] * 3 + [f32p] * 6 + [C.c_float, f32p, C.c_int, f32p]
    lib.hostchk_rt_dyn.argtypes = [C.c_int, C.c_int] + [f32p] * 5
    return lib


def test_model_file_round_trip_and_registration_checks(tmp_path):
    lib = native.load()
    for plant in ("iiwa14", "indy7"):
        m = native.Model.builtin(plant)
        assert m.nq == native.NQ[plant] and m.name == plant
        p = tmp_path / f"{plant}.gmdl"
        m.save(p)
        m2 = native.Model.load(p)
        assert bytes(m.raw) == bytes(m2.raw), "save / load must reproduce every table bit for bit"
    # the shipped data files are these tables
    for plant in ("iiwa14", "indy7"):
        shipped = native.Model.load(ROOT / "gato_b200" / "models" / f"{plant}.gmdl")
        assert bytes(shipped.raw) == bytes(native.Model.builtin(plant).raw)
    bad = native.Model.builtin("iiwa14")
    bad.array("X")[2, 6 * 4 + 1] = 0.5  # a non-zero entry in the top-right block of X_2
    assert lib.gato_model_register(C.byref(bad.raw)) == -3 and b"top-right" in lib.gato_last_error(None)
    bad = native.Model.builtin("iiwa14")
    bad.raw.nq = 5
    assert lib.gato_model_register(C.byref(bad.raw)) == -3 and b"nq must be 6, 7 or 8" in lib.gato_last_error(None)
    bad = native.Model.builtin("indy7")
    bad.raw.x_trig[0].idx = 36 * 6 + 1
    assert lib.gato_model_register(C.byref(bad.raw)) == -3
    (tmp_path / "junk.gmdl").write_text("gato_model 1\nnq 7\nfoo 1 2 3\nend\n")
    with pytest.raises(native.GatoError, match="unknown record"):
        native.Model.load(tmp_path / "junk.gmdl")
    # ids: the same tables always get the same id; gato_dims knows a registered model
    a = native.Model.builtin("indy7").register("indy7_as_data")
    assert native.Model.builtin("indy7").register("indy7_as_data") == a and native.PLANT_ID[a] >= 2
    nx, nu, traj = C.c_int(), C.c_int(), C.c_int()
    assert lib.gato_dims(native.PLANT_ID[a], 16, C.byref(nx), C.byref(nu), C.byref(traj)) == 0 and (nx.value, nu.value, traj.value) == (12, 6, 18 * 16 - 6)
    assert lib.gato_dims(2 + 7, 16, C.byref(nx), C.byref(nu), C.byref(traj)) == -1


CASES = [("iiwa14", None, 8, 1), ("indy7", None, 16, 3), ("iiwa14", ("custom7", 0, 11), 32, 2), ("indy7", ("custom6", 1, 12), 8, 3), ("chain8", ("chain8",), 12, None)]


@pytest.mark.parametrize("base,derive,N,cfg", CASES)
def test_table_driven_item_math_bit_exact(rtlib, base, derive, N, cfg):
    if derive is None:
        model, oplant = native.Model.builtin(base), base  # the compiled robot's own tables against the oracle's built-in robot
    elif base == "chain8":
        model = eight_joint_robot()
        oplant = pyapi.register_model("chain8", model)
    else:
        model = derived_robot(base, *derive)
        oplant = pyapi.register_model(derive[0], model)
    why = C.create_string_buffer(256)
    assert rtlib.hostchk_rt_set_model(C.byref(model.raw), why) == 0, why.value
    be = Backend("oracle", oplant, N)
    nq = be.d["nq"]
    rng = np.random.default_rng(5)
    n = 23
    x = rng.uniform(-2, 2, (n, 2 * nq)).astype(np.float32)
    u = rng.uniform(-20, 20, (n, nq)).astype(np.float32)
    fe = rng.normal(0, 3, (n, 6)).astype(np.float32)
    fe[:5] = 0
    o = be.dyn_dump(x, u, fe)
    for packed in (0, 1):  # one item per thread, and two items as the lanes of packed pairs
        qdd, ee = np.zeros((n, nq), np.float32), np.zeros((n, 3), np.float32)
        rtlib.hostchk_rt_dyn(packed, n, x.ravel(), u.ravel(), fe.ravel(), qdd.ravel(), ee.ravel())
        assert n_mismatch(qdd, o["qdd"]) == 0 and n_mismatch(ee, o["ee"][:, :3]) == 0, packed
    B = 3
    if cfg is None:
        from gato_b200.workloads import DEFAULT_SOLVER_PARAMS

        w = dict(random_workload(nq, N, B), params=dict(DEFAULT_SOLVER_PARAMS, dt=0.01))
    else:
        w = make_config(cfg, B=B, N=N)
    xu = w["xu"] + rng.normal(0, 0.1, w["xu"].shape).astype(np.float32)
    fext = rng.normal(0, 2, (B, 6)).astype(np.float32)
    p = dict(w["params"], vel_lim_cost=0.003, ctrl_lim_cost=0.002)  # exercise every barrier term
    k0 = be.stage_kkt(B, xu, w["xs"], w["ref"], fext, w["dt"], p)
    for variant in (1, 2):  # k_kkt's halves, k_kkt_fine's columns
        k1 = {k: np.zeros_like(v) for k, v in k0.items()}
        rtlib.hostchk_rt_stage_kkt(variant, N, B, xu.ravel(), w["xs"].ravel(), w["ref"].ravel(), fext.ravel(), np.float32(w["dt"]), cost7(p), *[k1[k].reshape(-1) for k in ("Q", "R", "q", "r", "A", "Bm", "c")])
        for k in k0:
            assert n_mismatch(k1[k], k0[k]) == 0, (variant, k)
    dz = rng.normal(0, 0.05, xu.shape).astype(np.float32)
    mu = np.full(B, 10, np.float32)
    for na, pm in ((1, p), (8, p), (8, dict(w["params"])), (8, dict(w["params"], q_lim_cost=0.0))):
        m0 = be.stage_merit(B, xu, dz, w["xs"], w["ref"], mu, fext, w["dt"], pm, na)
        for split in (0, 1):
            m1 = np.zeros_like(m0)
            rtlib.hostchk_rt_stage_merit(split, N, B, xu.ravel(), dz.ravel(), w["xs"].ravel(), w["ref"].ravel(), mu, fext.ravel(), np.float32(w["dt"]), cost7(pm), na, m1.reshape(-1))
            assert n_mismatch(m1, m0) == 0, (na, split)
