#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — mint golden vectors by RUNNING THE REFERENCE ITSELF on the B200 box.

Runs on the GPU box under gpurun (needs only oracle/_ref/*.so built from the unmodified reference by
`oracle/build_ref.sh all`, numpy and gato_b200.workloads; /root/reference is NOT read).  Writes
gpurun_out/golden/golden_<plant>_N<N>_<mode>.npz (+ meta.json, + reference timings); the files are then
copied into tests/golden/ and committed.

  python oracle/gen_golden.py [--out gpurun_out/golden] [--timing]
"""
import argparse
import json
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from gato_b200.workloads import make_config, DEFAULT_SOLVER_PARAMS  # noqa: E402
from oracle.pyapi import Backend, dims  # noqa: E402

LIMS = {  # generous sampling ranges inside the joint limits
    "iiwa14": (np.array([2.3, 1.6, 2.3, 1.6, 2.3, 1.6, 2.4]), 1.0, 20.0),
    "indy7": (np.array([2.4, 2.4, 2.4, 2.4, 2.4, 3.0]), 1.0, 20.0),
}

# which BASELINE.json config each library serves, stage-level batch and how many solves to keep
LIBS = [
    ("iiwa14", 8, "fast", 1), ("iiwa14", 8, "ieee", 1),
    ("iiwa14", 32, "fast", 2), ("iiwa14", 32, "ieee", 2),
    ("indy7", 32, "fast", 3), ("indy7", 32, "ieee", 3),
    ("iiwa14", 128, "fast", 4),
    ("indy7", 16, "fast", 3),
]


def stage_inputs(plant, N, B, cfg, seed):
    w = make_config(cfg, B=B, N=N)
    rng = np.random.default_rng(1000 + seed)
    d = dims(plant, N)
    nx, nu = d["nx"], d["nu"]
    xu = w["xu"].reshape(B, -1).copy()
    full = np.zeros((B, N, nx + nu), np.float32)
    full.reshape(B, -1)[:, : d["traj"]] = xu
    full[:, :, :nx] += rng.normal(0, 0.05, (B, N, nx)).astype(np.float32)
    full[:, :, nx:] += rng.normal(0, 1.0, (B, N, nu)).astype(np.float32)
    xu = full.reshape(B, -1)[:, : d["traj"]].copy()
    fext = np.zeros((B, 6), np.float32)
    fext[B // 2 :] = rng.normal(0, 2.0, (B - B // 2, 6)).astype(np.float32)
    return w, xu, fext


def gen_for_lib(plant, N, mode, cfg, out_dir, keep=2, stages_only=False):
    """stages_only: the dynamics dump and the per-stage chain only, no whole solves, with the merit stage launched through the harness's
    GREF_MERIT_EXTRA_SMEM switch (the SAME unmodified kernel with enough dynamic shared memory).  This is how indy7 is pinned: its merit kernel
    writes past the shared memory the reference's own launcher requests and faults on B200, so BSQP::solve cannot run."""
    import os

    if stages_only:
        os.environ["GREF_MERIT_EXTRA_SMEM"] = "4096"
    be = Backend("ref", plant, N, mode)
    d = dims(plant, N)
    nq = d["nq"]
    G = {}
    rng = np.random.default_rng(7)
    # ---- plant-level dynamics dump -------------------------------------------------------------
    n = 48
    qlim, qdl, ul = LIMS[plant]
    x = np.concatenate([rng.uniform(-1, 1, (n, nq)) * qlim, rng.uniform(-qdl, qdl, (n, nq))], 1).astype(np.float32)
    u = rng.uniform(-ul, ul, (n, nq)).astype(np.float32)
    fe = rng.normal(0, 3.0, (n, 6)).astype(np.float32)
    fe[: n // 2] = 0
    dd = be.dyn_dump(x, u, fe)
    G.update(dyn_x=x, dyn_u=u, dyn_fext=fe, **{"dyn_" + k: v for k, v in dd.items()})
    # ---- per-stage chain (each stage fed with the reference's own previous outputs) -----------------
    B = min(be.batches)
    w, xu, fext = stage_inputs(plant, N, B, cfg, N)
    p = dict(w["params"])
    rho = np.full(B, p["rho"], np.float32)
    rho[1::2] = 1e-3
    mu = np.full(B, p["mu"], np.float32)
    kk = be.stage_kkt(B, xu, w["xs"], w["ref"], fext, w["dt"], p)
    sc = be.stage_schur(B, kk, rho)
    lam0 = np.zeros((B, d["vecp"]), np.float32)
    lam_tol, it_tol = be.stage_pcg(B, sc["S"], sc["Pinv"], sc["gamma"], lam0, np.full(B, 1e-4, np.float32), 200)
    lam_cap, it_cap = be.stage_pcg(B, sc["S"], sc["Pinv"], sc["gamma"], lam0, np.full(B, -1.0, np.float32), 20)
    dz, qres, rres = be.stage_dz(B, lam_tol, sc["Qinv"], sc["Rinv"], kk["q"], kk["r"], kk["A"], kk["Bm"])
    m1 = be.stage_merit(B, xu, np.zeros_like(dz), w["xs"], w["ref"], mu, fext, w["dt"], p, 1)
    m8 = be.stage_merit(B, xu, dz, w["xs"], w["ref"], mu, fext, w["dt"], p, 8)
    mi = m1[:, 0].copy()
    mi[B - 1] = -1e30  # forces a line-search failure on the last solve
    ls = be.stage_linesearch(B, xu, dz, m8, mi, rho, np.ones(B, np.float32), 1)
    s = slice(0, keep)
    sl = lambda a: np.ascontiguousarray(a[s])  # noqa: E731
    G.update(st_B=np.int32(min(keep, B)), st_dt=np.float32(w["dt"]), st_params=json.dumps(p), st_xu=sl(xu), st_xs=sl(w["xs"]), st_ref=sl(w["ref"]), st_fext=sl(fext), st_rho=sl(rho), st_mu=sl(mu))
    G.update({"st_kkt_" + k: sl(v) for k, v in kk.items()})
    G.update({"st_schur_" + k: sl(v) for k, v in sc.items()})
    G.update(st_pcg_lam_tol=sl(lam_tol), st_pcg_it_tol=sl(it_tol), st_pcg_lam_cap=sl(lam_cap), st_pcg_it_cap=sl(it_cap))
    G.update(st_dz=sl(dz), st_dz_qres=sl(qres), st_dz_rres=sl(rres), st_merit1=sl(m1), st_merit8=sl(m8))
    # the line-search golden keeps ALL B rows but only compact outputs (+ the failure row)
    G.update(st_ls_merit8=m8, st_ls_merit_init=mi, st_ls_rho_in=rho, st_ls_step=ls["step"], st_ls_rho=ls["rho"], st_ls_drho=ls["drho"], st_ls_merit_out=ls["merit_init"],
             st_ls_xu_in=xu, st_ls_dz=dz, st_ls_xu_out=ls["xu"])
    if stages_only:
        G["stages_only"] = np.int32(1)
        np.savez_compressed(out_dir / f"golden_{plant}_N{N}_{mode}.npz", **G)
        print("wrote (stages only)", plant, N, mode, flush=True)
        return
    # ---- whole solves --------------------------------------------------------------------------
    for Bs in be.batches:
        ws = make_config(cfg, B=Bs, N=N)
        ps = dict(ws["params"])
        if Bs > 64 and mode == "ieee":
            continue
        sv = be.solver(Bs, ps)
        for key in ("rho", "mu"):
            if key in ws["extra"]:
                sv.set_batch(key, ws["extra"][key])
        o1 = sv.solve(ws["xu"], ws["xs"], ws["ref"], ws["dt"])
        if Bs <= 64:
            # run-to-run determinism of the reference (its Schur kernel 1 races on d_Q): repeat on fresh
            # solver objects and record, per solve, how many repeats reproduce the first run bit-for-bit
            same = np.zeros(Bs, np.int32)
            reps = 5
            for _ in range(reps):
                sx = be.solver(Bs, ps)
                for key in ("rho", "mu"):
                    if key in ws["extra"]:
                        sx.set_batch(key, ws["extra"][key])
                ox = sx.solve(ws["xu"], ws["xs"], ws["ref"], ws["dt"])
                same += (ox["XU"] == o1["XU"]).all(axis=1).astype(np.int32)
                sx.close()
            G[f"solve_B{Bs}_a_repro_count"] = same
            G[f"solve_B{Bs}_a_repro_reps"] = np.int32(reps)
            print("  determinism B", Bs, "solves reproduced in all repeats:", int((same == reps).sum()), "/", Bs, flush=True)
        # second solve WITHOUT reset, warm-started from the first result: pins lambda/rho persistence (bsqp.cuh:81-87,189)
        o2 = sv.solve(o1["XU"], ws["xs"], ws["ref"], ws["dt"])
        tag = f"solve_B{Bs}_"
        G[tag + "params"] = json.dumps(ps)
        G[tag + "dt"] = np.float32(ws["dt"])
        big = Bs > 64  # inputs are regenerated from the seeded make_config(); keep big-batch fixtures compact
        if not big:
            for k in ("xu", "xs", "ref"):
                G[tag + k] = ws[k]
        G[tag + "input_checksum"] = np.float64(ws["xu"].astype(np.float64).sum() + ws["ref"].astype(np.float64).sum())
        for nm, o in (("a_", o1), ("b_", o2)):
            for k in ("XU", "sqp_iters", "kkt_converged", "pcg_iters", "ls_min_merit", "ls_step_size", "final_merit", "initial_merit"):
                G[tag + nm + k] = o[k][:32] if (big and k == "XU") else o[k]
        # default-parameter variant (tolerance-terminated PCG, 1 SQP iteration) on the same inputs
        pd = dict(DEFAULT_SOLVER_PARAMS, dt=float(ws["dt"]))
        sd = be.solver(Bs, pd)
        od = sd.solve(ws["xu"], ws["xs"], ws["ref"], ws["dt"])
        G[tag + "d_params"] = json.dumps(pd)
        for k in ("XU", "sqp_iters", "kkt_converged", "pcg_iters", "ls_min_merit", "ls_step_size", "final_merit", "initial_merit"):
            G[tag + "d_" + k] = od[k][:32] if (big and k == "XU") else od[k]
        # sim_forward under per-solve wrench hypotheses
        fe2 = np.random.default_rng(5).normal(0, 2.0, (Bs, 6)).astype(np.float32)
        sd.set_batch("f_ext", fe2)
        xk = ws["xs"][0].copy()
        uk = np.random.default_rng(6).uniform(-5, 5, nq).astype(np.float32)
        G[tag + "sim_fext"], G[tag + "sim_xk"], G[tag + "sim_uk"] = fe2, xk, uk
        G[tag + "sim_out"] = sd.sim_forward(xk, uk, ws["dt"])
        sv.close()
        sd.close()
    np.savez_compressed(out_dir / f"golden_{plant}_N{N}_{mode}.npz", **G)
    print("wrote", plant, N, mode, "batches", be.batches, flush=True)


def timing(out_dir):
    """Reference CUDA build (sm_100, reference flags) timed on this B200: the R0 baseline row."""
    rows = []
    cases = [("bench", "iiwa14", 32, 512), (2, "iiwa14", 32, 128), (3, "indy7", 32, 512), (4, "iiwa14", 128, 1024), (1, "iiwa14", 8, 1)]
    for cfg, plant, N, B in cases:
        try:
            be = Backend("ref", plant, N, "fast")
        except FileNotFoundError:
            continue
        if B not in be.batches:
            continue
        w = make_config(cfg, B=B, N=N)
        for label, p in (("cfg", w["params"]), ("default", dict(DEFAULT_SOLVER_PARAMS, dt=float(w["dt"])))):
            sv = be.solver(B, p)
            sv.solve_timed(w["xu"], w["xs"], w["ref"], w["dt"], 3, True)  # warm-up
            reps = 20
            ev, us = sv.solve_timed(w["xu"], w["xs"], w["ref"], w["dt"], reps, True)
            o = sv.solve(w["xu"], w["xs"], w["ref"], w["dt"])
            row = dict(cfg=str(cfg), params=label, plant=plant, N=N, B=B, event_ms_p50=float(np.median(ev)), event_ms_min=float(ev.min()), ref_sqp_time_us_p50=float(np.median(us)),
                       solves_per_s=float(B / (np.median(ev) * 1e-3)), sqp_iters_mean=float(o["sqp_iters"].mean()), pcg_iters_mean=float(o["pcg_iters"].mean()) if o["n_pcg"] else 0.0,
                       n_pcg=o["n_pcg"], n_ls=o["n_ls"])
            rows.append(row)
            print(json.dumps(row), flush=True)
            sv.close()
    (out_dir / "reference_gpu_timing.json").write_text(json.dumps(rows, indent=1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "golden"))
    ap.add_argument("--timing", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--stages-only", action="store_true", help="with --only: dynamics + stage chain, merit through GREF_MERIT_EXTRA_SMEM (indy7)")
    a = ap.parse_args()
    out = Path(a.out)
    out.mkdir(parents=True, exist_ok=True)
    meta = dict(when=time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
                flags_fast="-std=c++17 -O3 -use_fast_math -DNDEBUG -gencode arch=compute_100,code=sm_100 (CMakeLists.txt:20-22, arch swapped)",
                flags_ieee="same without -use_fast_math")
    try:
        meta["gpu"] = subprocess.check_output(["nvidia-smi", "--query-gpu=name,driver_version", "--format=csv,noheader"], text=True).strip()
        meta["nvcc"] = "12.9.86"
    except Exception as e:  # pragma: no cover
        meta["gpu"] = f"unknown ({e})"
    (out / "meta.json").write_text(json.dumps(meta, indent=1))
    for plant, N, mode, cfg in LIBS:
        name = f"{plant}_N{N}_{mode}"
        if a.only:
            if a.only != name:
                continue
            try:
                gen_for_lib(plant, N, mode, cfg, out, keep=1 if N >= 128 else 2, stages_only=a.stages_only)
            except FileNotFoundError as e:
                print("skip (not built):", e, flush=True)
        else:
            # one process per library: a CUDA fault in the reference (e.g. the indy7 merit kernel) must not poison the rest
            rc = subprocess.call([sys.executable, __file__, "--out", str(out), "--only", name])
            print("lib", name, "exit", rc, flush=True)
    if a.timing:
        timing(out)


if __name__ == "__main__":
    main()
