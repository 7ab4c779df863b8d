"""The Python surface mirrors the reference's `bsqp` package: module names bsqpN{N}_{plant}, classes BSQP_{B}_float,
KNOT_POINTS, method names and result keys of python/bindings.cu:96-147,224-237 and the wrapper of python/bsqp/interface.py."""
import subprocess

import numpy as np
import pytest

from conftest import ROOT, n_mismatch
from gato_b200.workloads import make_config

RESULT_KEYS = {"XU", "sqp_time_us", "sqp_iters", "kkt_converged", "final_merit", "initial_merit", "ls_num_iters", "pcg_times_us", "pcg_iters", "ls_min_merit", "ls_step_size"}
METHODS = ["solve", "reset_dual", "set_f_ext_batch", "set_rho_penalty_batch", "set_drho_batch", "set_mu_batch", "set_pcg_tol_batch", "sim_forward", "reset_rho", "set_rho_adaptation"]


def test_module_and_class_names():
    import importlib

    m = importlib.import_module("gato_b200.bsqp.bsqpN32_iiwa14")
    assert m.KNOT_POINTS == 32
    for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024):  # the reference's registered batch sizes (bindings.cu:254-264)
        cls = getattr(m, f"BSQP_{B}_float")
        for name in METHODS:
            assert callable(getattr(cls, name))
    from gato_b200.bsqp import bsqpN16_indy7

    assert bsqpN16_indy7.KNOT_POINTS == 16 and not hasattr(m, "BSQP_0_float")
    with pytest.raises(ImportError):
        importlib.import_module("gato_b200.bsqp.bsqpN32_panda")


def test_constructor_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gato_b200 import native
    from gato_b200.bsqp import bsqpN8_iiwa14

    with pytest.raises(native.GatoError):
        bsqpN8_iiwa14.BSQP_1_float()
    with pytest.raises(TypeError):
        bsqpN8_iiwa14.BSQP_1_float(0.01, 5)


@pytest.mark.gpu
def test_binding_class_matches_oracle(oracle_built):
    from gato_b200.bsqp import bsqpN32_iiwa14
    from oracle.pyapi import PARAM_ORDER, Backend

    w = make_config(2, B=8)
    p = w["params"]
    solver = bsqpN32_iiwa14.BSQP_8_float(*[p[k] for k in PARAM_ORDER])
    fe = np.random.default_rng(1).normal(0, 1, (8, 6)).astype(np.float32)
    solver.set_f_ext_batch(fe)
    solver.set_mu_batch(np.full(8, 5.0, np.float32))
    r = solver.solve(w["xu"], w["dt"], w["xs"], w["ref"])
    assert set(r) == RESULT_KEYS
    assert r["XU"].shape == (8, 665) and r["XU"].dtype == np.float32 and r["pcg_iters"].shape == (r["ls_num_iters"], 8) and r["pcg_times_us"].shape == (r["ls_num_iters"],)
    o = Backend("oracle", "iiwa14", 32).solver(8, p)
    o.set_batch("f_ext", fe)
    o.set_batch("mu", np.full(8, 5.0, np.float32))
    ro = o.solve(w["xu"], w["xs"], w["ref"], w["dt"])
    assert n_mismatch(r["XU"], ro["XU"]) == 0 and np.array_equal(r["pcg_iters"], ro["pcg_iters"][: r["ls_num_iters"]]) and n_mismatch(r["ls_step_size"], ro["ls_step_size"]) == 0
    x1 = solver.sim_forward(w["xs"][0], np.zeros(7, np.float32), 0.01)
    assert x1.shape == (8, 14) and n_mismatch(x1, o.sim_forward(w["xs"][0], np.zeros(7, np.float32), 0.01)) == 0


@pytest.mark.gpu
def test_interface_wrapper_closed_loop_step(oracle_built):
    from gato_b200.bsqp.interface import BSQP

    w = make_config(5, B=16)
    s = BSQP(None, 16, 32, 0.01, max_sqp_iters=2, max_pcg_iters=100, pcg_tol=1e-4, mu=10.0, q_cost=2.0, qd_cost=1e-2, u_cost=2e-6, N_cost=50.0, q_lim_cost=0.01, rho=0.01,
             rho_batch=w["extra"]["rho"][:16], mu_batch=w["extra"]["mu"][:16], plant_type="iiwa14")
    XU, t_us = s.solve(w["xs"], w["ref"])
    st = s.get_stats()
    assert XU.shape == (16, 665) and np.isfinite(XU).all() and t_us > 0
    assert st["pcg_iters"].shape[1] == 16 and st["min_merit"].shape == st["step_size"].shape and np.isfinite(st["best_merit_iter1"])
    assert np.array_equal(XU[:, :14], w["xs"]) or np.isfinite(XU[:, :14]).all()
    s.reset_rho(), s.reset()
    assert np.all(s.XU_B == 0)


@pytest.mark.gpu
def test_reference_example_compiles_and_runs_unchanged():
    """examples/bsqp.cu (indy7, N=16, B=16) built UNMODIFIED against include/gato_compat + libgato_b200 (binary built in this
    repo's container by __graft_entry__/tests, the reference sources are not present on the GPU box)."""
    exe = ROOT / "tests" / "compat" / "_build" / "bsqp_example_unchanged"
    if not exe.exists():
        pytest.skip("compat example binary not built")
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "XU Traj:" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_reference_pybind_module_source_unchanged_over_this_library():
    """python/bindings.cu of the reference, compiled UNMODIFIED against include/gato_compat and linked with libgato_b200 (built by
    __graft_entry__.build() where /root/reference is present): the original pybind classes drive the new solver and return what the
    ctypes mirror returns, bit for bit."""
    import importlib.util
    import sysconfig

    path = ROOT / "tests" / "compat" / "_build" / ("bsqpN8_iiwa14" + sysconfig.get_config_var("EXT_SUFFIX"))
    if not path.exists():
        pytest.skip("reference pybind module not built")
    spec = importlib.util.spec_from_file_location("bsqpN8_iiwa14", path)
    ref_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_mod)
    from gato_b200.bsqp import bsqpN8_iiwa14 as mirror
    from gato_b200.native import PARAM_ORDER

    assert ref_mod.KNOT_POINTS == 8 == mirror.KNOT_POINTS
    w = make_config(1, B=16)
    args = [w["params"][k] for k in PARAM_ORDER]
    a, b = ref_mod.BSQP_16_float(*args), mirror.BSQP_16_float(*args)
    fe = np.random.default_rng(2).normal(0, 2, (16, 6)).astype(np.float32)
    for s in (a, b):
        s.set_f_ext_batch(fe)
        s.set_mu_batch(np.full(16, 5.0, np.float32))
    ra = a.solve(w["xu"], float(w["dt"]), w["xs"], w["ref"])
    rb = b.solve(w["xu"], float(w["dt"]), w["xs"], w["ref"])
    assert set(ra.keys()) == set(rb.keys())
    for k in ("XU", "pcg_iters", "ls_step_size", "ls_min_merit", "sqp_iters", "kkt_converged", "final_merit", "initial_merit"):
        assert np.array_equal(np.asarray(ra[k]), np.asarray(rb[k])), k
    assert ra["ls_num_iters"] == rb["ls_num_iters"]
    xa, xb = a.sim_forward(w["xs"][0], np.zeros(7, np.float32), 0.01), b.sim_forward(w["xs"][0], np.zeros(7, np.float32), 0.01)
    assert np.array_equal(np.asarray(xa), np.asarray(xb))
