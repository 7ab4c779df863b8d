"""Run-time robot models on the GPU (pytest -m gpu): a robot registered from DATA (gato_model_register / a .gmdl file) runs the table-driven
kernels (k_kkt / k_kkt_fine / k_merit_ls / k_sim_forward / k_ee_pos instantiated for RtPlant<nq>, gato_b200/csrc/rbd_rt.cuh) behind the same
C ABI.  Bar: bit-for-bit equal to the CPU oracle given the same tables -- and, for the compiled robots' own tables, bit-for-bit equal to the
compiled kernels."""
import numpy as np
import pytest

from conftest import ROOT, n_mismatch
from gato_b200.workloads import DEFAULT_SOLVER_PARAMS, make_config
from test_models_host import derived_robot

pytestmark = pytest.mark.gpu


def _register(base, derive):
    from gato_b200 import native
    from oracle import pyapi

    if derive is None:
        model = native.Model.load(ROOT / "gato_b200" / "models" / f"{base}.gmdl")  # the compiled robot's tables, from the shipped data file
        return model.register(f"{base}_as_data"), base
    model = derived_robot(base, *derive)
    return model.register(derive[0]), pyapi.register_model(derive[0], model)


CASES = [("iiwa14", None, 8, 1, 3), ("iiwa14", None, 32, 2, 40), ("indy7", None, 16, 3, 5), ("iiwa14", ("custom7", 0, 11), 32, 2, 6), ("indy7", ("custom6", 1, 12), 12, 3, 4)]


@pytest.mark.parametrize("base,derive,N,cfg,B", CASES)
def test_table_driven_kernels_bit_exact(oracle_built, base, derive, N, cfg, B):
    from gato_b200.native import GatoBackend
    from oracle.pyapi import Backend

    gplant, oplant = _register(base, derive)
    o, g = Backend("oracle", oplant, N), GatoBackend(gplant, N)
    rng = np.random.default_rng(11)
    w = make_config(cfg, B=B, N=N)
    xu = w["xu"] + rng.normal(0, 0.05, w["xu"].shape).astype(np.float32)
    fext = rng.normal(0, 2, (B, 6)).astype(np.float32)
    fext[0] = 0
    p = dict(w["params"], vel_lim_cost=0.002, ctrl_lim_cost=0.001)
    mu = np.full(B, 10, np.float32)
    # stages that touch the robot: KKT (k_kkt_fine for these batch sizes, k_kkt for B = 40 x N = 32), merit x1 / x8 (split and plain kernels)
    ko, kg = o.stage_kkt(B, xu, w["xs"], w["ref"], fext, w["dt"], p), g.stage_kkt(B, xu, w["xs"], w["ref"], fext, w["dt"], p)
    for k in ko:
        assert n_mismatch(kg[k], ko[k]) == 0, f"kkt {k}"
    dz = rng.normal(0, 0.05, xu.shape).astype(np.float32)
    for na in (1, 8):
        mo = o.stage_merit(B, xu, dz, w["xs"], w["ref"], mu, fext, w["dt"], p, na)
        mg = g.stage_merit(B, xu, dz, w["xs"], w["ref"], mu, fext, w["dt"], p, na)
        assert n_mismatch(mg, mo) == 0, f"merit x{na}"
    if derive is None:
        # the compiled robot's tables through the table-driven kernels == the compiled kernels
        kc = GatoBackend(base, N).stage_kkt(B, xu, w["xs"], w["ref"], fext, w["dt"], p)
        for k in kc:
            assert n_mismatch(kg[k], kc[k]) == 0, f"kkt {k} (table-driven vs compiled)"
    # whole solves, state setters, sim_forward, ee_pos
    for prm in (w["params"], dict(DEFAULT_SOLVER_PARAMS, dt=float(w["dt"]), max_sqp_iters=3)):
        so, sg = o.solver(B, prm), g.solver(B, prm)
        so.set_batch("f_ext", fext), sg.set_batch("f_ext", fext)
        xin = w["xu"]
        for rep in range(2):
            ro, rg = so.solve(xin, w["xs"], w["ref"], w["dt"]), sg.solve(xin, w["xs"], w["ref"], w["dt"])
            assert rg["n_pcg"] == ro["n_pcg"] and rg["n_ls"] == ro["n_ls"]
            for k in ("pcg_iters", "sqp_iters", "kkt_converged"):
                assert np.array_equal(rg[k], ro[k]), k
            for k in ("XU", "ls_step_size", "ls_min_merit", "final_merit", "initial_merit"):
                assert n_mismatch(rg[k], ro[k]) == 0, k
            xin = ro["XU"]
        uk = rng.uniform(-5, 5, o.d["nq"]).astype(np.float32)
        assert n_mismatch(sg.sim_forward(w["xs"][0], uk, w["dt"]), so.sim_forward(w["xs"][0], uk, w["dt"])) == 0
        q = rng.uniform(-2, 2, (9, o.d["nq"])).astype(np.float32)
        dd = o.dyn_dump(np.concatenate([q, np.zeros_like(q)], axis=1), np.zeros_like(q), np.zeros((9, 6), np.float32))
        assert n_mismatch(sg.ee_pos(q), dd["ee"][:, :3]) == 0
        so.close(), sg.close()


def test_headline_shape_with_a_registered_model(oracle_built):
    """B = 512, N = 32 through the large-grid kernels (k_kkt's three kinds, the overlapped line search): the iiwa14 tables as data give
    the compiled iiwa14 kernels' result bit for bit."""
    from gato_b200.native import Solver

    gplant, _ = _register("iiwa14", None)
    w = make_config("bench")
    sc, sr = Solver("iiwa14", w["N"], w["B"], w["params"]), Solver(gplant, w["N"], w["B"], w["params"])
    rc, rr = sc.solve(w["xu"], w["xs"], w["ref"], w["dt"]), sr.solve(w["xu"], w["xs"], w["ref"], w["dt"])
    for k in ("pcg_iters", "sqp_iters", "kkt_converged"):
        assert np.array_equal(rc[k], rr[k]), k
    for k in ("XU", "ls_step_size", "ls_min_merit", "final_merit", "initial_merit"):
        assert n_mismatch(rc[k], rr[k]) == 0, k
    print(f"device ms: compiled {rc['device_time_ms']:.3f}  table-driven {rr['device_time_ms']:.3f}")
    sc.close(), sr.close()


@pytest.mark.parametrize("N,B", [(12, 4), (30, 3), (32, 2), (40, 2)])
def test_eight_joint_robot_every_stage_and_whole_solves(oracle_built, N, B):
    """nx = 16: no reference robot has eight joints, so k_schur / k_pcg / k_pcg_stream / k_pcg_cluster are instantiated for this size only for
    run-time models.  Every stage and whole solves against the oracle, bit for bit (N = 30: register-resident k_pcg at 512 threads; N = 32, 40:
    (N + 2) nx > 512 -> the cluster kernel with one / two CTAs per solve after k_pcg_stream's K2 phase)."""
    from gato_b200.native import GatoBackend
    from gato_b200.workloads import DEFAULT_SOLVER_PARAMS
    from oracle import pyapi
    from oracle.pyapi import Backend
    from test_models_host import eight_joint_robot, random_workload

    model = eight_joint_robot()
    gplant, oplant = model.register("chain8"), pyapi.register_model("chain8", model)
    o, g = Backend("oracle", oplant, N), GatoBackend(gplant, N)
    d = o.d
    assert d["nx"] == 16
    w = random_workload(8, N, B, seed=N)
    rng = np.random.default_rng(3)
    xu, fext = w["xu"], rng.normal(0, 2, (B, 6)).astype(np.float32)
    p = dict(DEFAULT_SOLVER_PARAMS, dt=0.01, vel_lim_cost=0.002, ctrl_lim_cost=0.001)
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
        assert np.array_equal(ig, io) and n_mismatch(lg, lo) == 0, (eps, cap)
    dzo = o.stage_dz(B, lo, so["Qinv"], so["Rinv"], ko["q"], ko["r"], ko["A"], ko["Bm"])
    dzg = g.stage_dz(B, lo, so["Qinv"], so["Rinv"], ko["q"], ko["r"], ko["A"], ko["Bm"])
    for a, b in zip(dzg, dzo):
        assert n_mismatch(a, b) == 0
    for na in (1, 8):
        mo = o.stage_merit(B, xu, dzo[0], w["xs"], w["ref"], mu, fext, w["dt"], p, na)
        mg = g.stage_merit(B, xu, dzo[0], w["xs"], w["ref"], mu, fext, w["dt"], p, na)
        assert n_mismatch(mg, mo) == 0, f"merit x{na}"
    for prm in (dict(p, max_sqp_iters=3), dict(p, max_sqp_iters=4, max_pcg_iters=30, pcg_tol=-1.0)):
        so_, sg_ = o.solver(B, prm), g.solver(B, prm)
        so_.set_batch("f_ext", fext), sg_.set_batch("f_ext", fext)
        xin = xu
        for rep in range(2):
            ro, rg = so_.solve(xin, w["xs"], w["ref"], w["dt"]), sg_.solve(xin, w["xs"], w["ref"], w["dt"])
            for k in ("pcg_iters", "sqp_iters", "kkt_converged"):
                assert np.array_equal(rg[k], ro[k]), k
            for k in ("XU", "ls_step_size", "ls_min_merit", "final_merit", "initial_merit"):
                assert n_mismatch(rg[k], ro[k]) == 0, k
            xin = ro["XU"]
        so_.close(), sg_.close()


def test_closed_loop_mpc_step_with_a_registered_model(oracle_built):
    """The device-side MPC step (gato_mpc_step: prepare, solve, sim_forward scoring, winner adoption) with a robot that exists only as data:
    the whole comparison of tests/test_gpu_mpc.py -- device step vs host composition vs the oracle, bit for bit -- under the registered plant name."""
    from test_gpu_mpc import test_device_mpc_step_equals_host_composition_and_oracle as mpc_case

    gplant, oplant = _register("iiwa14", ("custom7", 0, 11))
    assert gplant == oplant == "custom7"
    mpc_case("custom7", 8, 16)
