#!/usr/bin/env python3
"""bench.py — batched SQP solves/s on the BASELINE.json headline shape (iiwa14, N=32, batch 512 per GPU).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W   (CPU arm: the oracle port on the host cores, same config)

One "step" = one gato_solve of the whole batch (max_sqp_iters=4, max_pcg_iters=50, pcg_tol=-1: fixed iteration
caps so every implementation does the same work, SURVEY.md §8(d) cfg 2 at batch 512).  `value` times the solve with
inputs resident in HBM (CUDA events on the solver stream, per step, L2 flushed between steps); `e2e` times the same
solve through the host-buffer entry point (pinned H2D of xu/x_s/ref + D2H of xu and stats inside the window).
Prints ONE JSON line on rank 0.  Extra keys (rank 0, N=1 unless noted): per-kernel rooflines against the FP32 peak measured
on this GPU, the reference CUDA build on the same GPU, parity against the stock (-use_fast_math) reference build, a
tolerance-terminated default-parameter workload, the CPU oracle, and (N>1) a strong-scaling figure.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from gato_b200.workloads import DEFAULT_SOLVER_PARAMS, make_config  # noqa: E402

WORKLOAD = "iiwa14_N32_B512_sqp4_pcg50_fixedcaps"
METRIC = "batched SQP solves/sec (iiwa14, N=32, batch 512)"
_N, _NX, _NU = 32, 14, 7
BYTES_PER_SOLVE = 10024  # compulsory HBM I/O per solve (xu, ref, x_s, f_ext, lambda in; xu, lambda, stats out), SURVEY.md §8(d)
FP32_NOMINAL_TFLOPS = 74.4  # 148 SM x 128 lanes x 2 x 1.965 GHz; used only if the on-box measurement is unavailable


def kernel_work(pcg_iters=50, N=_N, nx=_NX, nu=_NU):
    """Algorithmic flops and bytes ONE launch of each kernel performs per solve (flop closed forms of SURVEY.md §8(d),
    multiply and add counted separately).  Bytes: what the launch has to read and write once (fp32), independent of how it is cached."""
    kkt_floats = N * (2 * nx * nx + nu * nu + nx * nu + 2 * nx + nu)  # Q, A, R, B, q, c, r
    traj = (nx + nu) * N - nu
    vec = (N + 2) * nx
    schur_k1 = (N - 1) * (3 * nx * nx * (nx + 1) * 2 + 3 * nu * nu * (nu + 1) + 2 * nx ** 3 + 2 * nx * nu * nu + 2 * nx ** 3 + 2 * nx * nx * nu + (4 * nx * nx + 2 * nx * nu)
                          + 3 * nx * nx * (nx + 1)) + (3 * nx * nx * (nx + 1) + 2 * nx * nx)
    pcg_iter = 2 * (N * nx * 3 * nx * 2) + 2 * 2 * vec + 3 * 2 * vec
    return {
        "k_kkt": {"flop": (N - 1) * 44.8e3 + N * 12.6e3, "bytes": 4 * (traj + 6 * N + nx + 6 + kkt_floats)},
        "k_schur": {"flop": float(schur_k1), "bytes": 4 * (kkt_floats + 1 + 3 * N * nx * nx + N * nx * nx + vec + N * nx * nx + N * nu * nu)},
        # rows of S, main blocks of P^-1, gamma, lambda in/out, and for the primal step A, B, Q^-1, R^-1, q, r (read-modify-write) and dz
        "k_pcg": {"flop": (N - 1) * 2 * 2 * nx ** 3 + (2 * N * nx * 3 * nx * 2 + 2 * vec) + pcg_iters * pcg_iter + N * (4 * nx * nx + 2 * nx * nu + 2 * nu * nu + 3 * nx),
                  "bytes": 4 * (N * 3 * nx * nx + N * nx * nx + 3 * vec + (N - 1) * nx * nx + (N - 1) * nx * nu + N * nx * nx + (N - 1) * nu * nu + 2 * N * nx + 2 * (N - 1) * nu + traj)},
        "k_merit_ls<8>": {"flop": 8 * N * 13.8e3 + 8 + 2 * traj, "bytes": 4 * (3 * traj + 6 * N + nx + 6 + 16)},
        "k_merit_ls<1>": {"flop": N * 13.8e3, "bytes": 4 * (traj + 6 * N + nx + 6 + 2)},
    }


def flop_per_solve(sqp_iters, pcg_iters):
    w = kernel_work(pcg_iters)
    return sqp_iters * (w["k_kkt"]["flop"] + w["k_schur"]["flop"] + w["k_pcg"]["flop"] + w["k_merit_ls<8>"]["flop"]) + 2 * w["k_merit_ls<1>"]["flop"]


def sample_clocks(stop, out):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    idx = os.environ.get("LOCAL_RANK", "0")
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", idx], capture_output=True, text=True, timeout=5)
            parts = [x.strip() for x in r.stdout.strip().split(",")]
            if len(parts) >= 6:
                out.append(parts)
        except Exception:
            pass
        stop.wait(0.2)


def clocks_summary(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    sm = sorted(int(s[0]) for s in samples if s[0].isdigit())
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in samples)]
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(samples[0][1]) if samples[0][1].isdigit() else None, "reasons": reasons}


def cpu_oracle_rate(n_solves, reps, threads=None):
    """Oracle port (oracle/bsqp_oracle.cpp) on the host cores, same workload; returns (solves/s, cores, per-pass seconds)."""
    from oracle.pyapi import Backend, ensure_oracle_built

    ensure_oracle_built()
    w = make_config("bench", B=n_solves)
    be = Backend("oracle", w["plant"], w["N"])
    cores = threads or os.cpu_count() or 1
    be.set_threads(cores)
    s = be.solver(n_solves, w["params"])
    times = []
    for _ in range(reps):
        s.reset("dual")
        s.reset("rho")
        t0 = time.perf_counter()
        s.solve(w["xu"], w["xs"], w["ref"], w["dt"])
        times.append(time.perf_counter() - t0)
    return n_solves / float(np.median(times)), cores, times


def run_reference(args):
    """--impl reference: the reference ships no CPU solver; its CPU arm is the oracle port of the same math on all host
    threads, on this arm's config (512 solves per step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.batch
    rate, cores, times = cpu_oracle_rate(n, args.warmup + args.steps)
    t = times[args.warmup:]
    ms = 1e3 * float(np.mean(t))
    rate = n / float(np.mean(t))
    sample = f"{n} solves of the {WORKLOAD} workload per step (the whole seeded batch), {args.steps} steps after {args.warmup} warm-up"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "plant": "iiwa14", "knot_points": 32, "batch_per_gpu": n, "max_sqp_iters": 4, "max_pcg_iters": 50, "pcg_tol": -1.0},
        "cpu_baseline": {"value": rate, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def reference_gpu_row(w):
    """Reference CUDA build (unmodified sources, sm_100, -use_fast_math) on this GPU, if oracle/_ref travelled with the repo: device time of
    BSQP::solve (CUDA events, warm L2) and end to end through the harness's host-buffer call (H2D, solve, D2H, statistics: what the
    reference's pybind solve does, python/bindings.cu:68-148)."""
    try:
        from oracle.pyapi import Backend

        be = Backend("ref", w["plant"], w["N"], "fast")
        if w["B"] not in be.batches:
            return None
        sv = be.solver(w["B"], w["params"])
        sv.solve_timed(w["xu"], w["xs"], w["ref"], w["dt"], 3, True)
        ev, us = sv.solve_timed(w["xu"], w["xs"], w["ref"], w["dt"], 10, True)
        wall = []
        for _ in range(6):
            sv.reset("dual"), sv.reset("rho")
            t0 = time.perf_counter()
            sv.solve(w["xu"], w["xs"], w["ref"], w["dt"])
            wall.append(time.perf_counter() - t0)
        sv.close()
        ms = float(np.median(ev))
        e2e_ms = 1e3 * float(np.median(wall[1:]))
        return {"value": w["B"] / (ms * 1e-3), "unit": "solves/s", "ms_per_step": ms, "e2e": {"value": w["B"] / (e2e_ms * 1e-3), "unit": "solves/s", "ms_per_step": e2e_ms},
                "build": "reference sources unmodified, nvcc -O3 -use_fast_math -DNDEBUG sm_100 (oracle/build_ref.sh)"}
    except Exception as e:  # pragma: no cover
        return {"unavailable": str(e)[:200]}


def parity_vs_stock_reference(native, B=512):
    """The CUDA path against the reference's STOCK build (-use_fast_math) on this GPU, same inputs: integer-outcome mismatch rates and
    trajectory errors, on the bench workload (fixed caps) and on a tolerance-terminated one (default pcg_tol, 4 SQP iterations) where
    PCG counts are data dependent.  (Against the reference's sources compiled without -use_fast_math the path is bit-exact:
    tests/test_gpu_reference_live.py.)"""
    try:
        from oracle.pyapi import Backend

        out = {}
        be = Backend("ref", "iiwa14", 32, "fast")
        if B not in be.batches:
            return {"unavailable": f"no B={B} instantiation in oracle/_ref"}
        base = make_config("bench", B=B)
        for name, p in (("fixed_caps", base["params"]), ("tolerance_terminated", dict(base["params"], max_pcg_iters=200, pcg_tol=1e-4))):
            rs, gs = be.solver(B, p), native.Solver("iiwa14", 32, B, p)
            r, g = rs.solve(base["xu"], base["xs"], base["ref"], base["dt"]), gs.solve(base["xu"], base["xs"], base["ref"], base["dt"])
            rs.close(), gs.close()
            n_it = min(r["pcg_iters"].shape[0], g["pcg_iters"].shape[0])
            n_ls = min(r["ls_step_size"].shape[0], g["ls_step_size"].shape[0])
            err = np.abs(g["XU"] - r["XU"]).max(axis=1) / np.maximum(np.abs(r["XU"]).max(axis=1), 1e-30)
            same_all_steps = (g["ls_step_size"][:n_ls] == r["ls_step_size"][:n_ls]).all(axis=0)
            dp = np.abs(g["pcg_iters"][:n_it].astype(np.int64) - r["pcg_iters"][:n_it].astype(np.int64))
            out[name] = {
                "solves": int(B), "sqp_iters_compared": int(n_it),
                "pcg_count_mismatch_rate": float((g["pcg_iters"][:n_it] != r["pcg_iters"][:n_it]).mean()), "pcg_count_max_abs_diff": int(dp.max()) if dp.size else 0,
                "ls_step_mismatch_rate": float((g["ls_step_size"][:n_ls] != r["ls_step_size"][:n_ls]).mean()),
                "ls_step_mismatch_rate_first_iteration": float((g["ls_step_size"][0] != r["ls_step_size"][0]).mean()) if n_ls else None,
                "solves_with_identical_step_sequence": float(same_all_steps.mean()),
                "traj_rel_err_median": float(np.median(err)), "traj_rel_err_max": float(err.max()),
                "traj_rel_err_median_identical_steps": float(np.median(err[same_all_steps])) if same_all_steps.any() else None,
                "traj_rel_err_max_identical_steps": float(err[same_all_steps].max()) if same_all_steps.any() else None,
                "n_pcg": [int(g["n_pcg"]), int(r["n_pcg"])], "kkt_converged_equal": bool(np.array_equal(g["kkt_converged"], r["kkt_converged"])),
                "initial_merit_rel_err_max": float(np.abs(g["initial_merit"] - r["initial_merit"]).max() / np.abs(r["initial_merit"]).max()),
            }
        out["test_policy"] = {"pcg_count_match_min": 0.9, "ls_step_match_min": 0.95, "traj_rel_err_median_max": 1e-4, "where": "tests/test_gpu_parity.py::test_against_reference_golden_solves"}
        return out
    except Exception as e:  # pragma: no cover
        return {"unavailable": str(e)[:200]}


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of each kernel at batch 512 from the committed `ncu --set full` captures."""
    try:
        return json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())
    except Exception:
        return {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="gato_b200")
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference-CUDA-build row and the parity-vs-stock-reference block")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the default-parameter workload and the strong-scaling figure")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch

    from gato_b200 import native

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: gato_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch
    stream = torch.cuda.Stream()  # a real (non-default) stream shared by torch and the solver, so torch events see the solver's work
    torch.cuda.set_stream(stream)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")  # > 126 MB L2

    def device_timed(w, sl, nb, steps, warmup):
        """steps x (reset, L2 flush, one solve of rows sl of workload w) -> per-step CUDA-event ms, launches, last stats."""
        xu0 = torch.from_numpy(w["xu"][sl].copy()).cuda()
        xs = torch.from_numpy(w["xs"][sl].copy()).cuda()
        ref = torch.from_numpy(w["ref"][sl].copy()).cuda()
        xu = xu0.clone()
        solver = native.Solver(w["plant"], w["N"], nb, w["params"], device=local, stream=stream.cuda_stream)

        def one_step():
            xu.copy_(xu0)
            solver.reset("dual")
            solver.reset("rho")
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            solver.solve_async(xu.data_ptr(), xs.data_ptr(), ref.data_ptr(), float(w["dt"]))
            e1.record(stream)
            st = solver.solve_wait()
            return e0.elapsed_time(e1), st

        for _ in range(warmup):
            one_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        l0 = solver.kernel_launches()
        ms, st = [], None
        for _ in range(steps):
            m, st = one_step()
            ms.append(m)
        torch.cuda.synchronize()
        return ms, solver.kernel_launches() - l0, st, solver, one_step

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- headline: every rank solves its own shard of the (synthetic) global batch: independent problems, no data-path collective ----
    w = make_config("bench", B=B * world)
    sl = slice(rank * B, (rank + 1) * B)
    stop, samples = threading.Event(), []
    th = threading.Thread(target=sample_clocks, args=(stop, samples), daemon=True)
    th.start()
    step_ms, launches, st, solver, one_step = device_timed(w, sl, B, args.steps, args.warmup)
    stop.set()
    th.join(timeout=2)
    if world > 1:
        dist.barrier()
    total_ms = max_over_ranks(float(np.sum(step_ms)))
    value = B * world * args.steps / (total_ms * 1e-3)
    ms_per_step = total_ms / args.steps

    # ---- per-kernel durations: a few extra steps with the library's event-per-launch instrumentation (not part of the timed region) ----
    kernels = {}
    if rank == 0:
        solver.set_kernel_timing(True)
        acc = {}
        reps = 5
        for _ in range(reps):
            one_step()
            for k, (ms, n) in solver.kernel_times().items():
                a = acc.setdefault(k, [0.0, 0])
                a[0] += ms
                a[1] += n
        solver.set_kernel_timing(False)
        tot_k = sum(a[0] for a in acc.values())
        for k, (ms, n) in acc.items():
            if n:
                kernels[k] = {"us_per_launch": 1e3 * ms / n, "launches_per_step": n // reps, "share_of_step": ms / tot_k}

    # ---- e2e: host buffers, pinned, H2D + D2H inside the window ----
    e2e = None
    if not args.no_e2e:
        h_xu0 = torch.from_numpy(w["xu"][sl].copy()).pin_memory()
        h_xu = torch.empty_like(h_xu0).pin_memory()
        h_xs = torch.from_numpy(w["xs"][sl].copy()).pin_memory()
        h_ref = torch.from_numpy(w["ref"][sl].copy()).pin_memory()
        hnd = solver.h
        e_ms = []
        import ctypes as C

        fp = lambda t_: C.cast(t_.data_ptr(), C.c_void_p)  # noqa: E731
        lib = C.CDLL(str(native.lib_path()))  # private handle: raw-pointer signature for pinned torch buffers
        lib.gato_solve_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.POINTER(native.GatoStats)]
        for i in range(args.warmup + args.steps):
            h_xu.copy_(h_xu0)
            solver.reset("dual")
            solver.reset("rho")
            flush.zero_()
            torch.cuda.synchronize()
            stt = native.GatoStats()
            t0 = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            rc = lib.gato_solve_host(hnd, fp(h_xu), fp(h_xs), fp(h_ref), float(w["dt"]), C.byref(stt))
            e1.record(stream)
            e1.synchronize()
            assert rc == 0
            if i >= args.warmup:
                e_ms.append(max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)))
        tot = max_over_ranks(float(np.sum(e_ms)))
        n_it = int(w["params"]["max_sqp_iters"])
        e2e = {"value": B * world * args.steps / (tot * 1e-3), "unit": "solves/s",
               "h2d_bytes_per_step": int(4 * (h_xu0.numel() + h_xs.numel() + h_ref.numel())),
               "d2h_bytes_per_step": int(4 * h_xu0.numel() + 4 * B * (3 * n_it + 3) + 4 * n_it)}
    solver.close()

    # ---- extra workloads (all ranks take part so that the collectives match) ----
    extra = {}
    if not args.no_extra:
        # default solver parameters (python/bsqp/config.py:35-50): 1 SQP iteration, PCG terminated by tolerance 1e-4 (cap 200)
        wd = make_config("bench", B=B * world)
        wd["params"] = dict(DEFAULT_SOLVER_PARAMS, dt=float(wd["dt"]))
        d_ms, _, d_st, d_solver, _ = device_timed(wd, sl, B, 10, 3)
        d_solver.close()
        d_tot = max_over_ranks(float(np.sum(d_ms)))
        extra["default_params_workload"] = {"workload": "iiwa14_N32_B512_sqp1_pcgtol1e-4_cap200 (DEFAULT_SOLVER_PARAMS)", "value": B * world * 10 / (d_tot * 1e-3), "unit": "solves/s",
                                            "ms_per_step": d_tot / 10, "pcg_iters_mean_rank0": float(np.mean(d_st["pcg_iters"])), "pcg_iters_max_rank0": int(np.max(d_st["pcg_iters"]))}
        # run-time robot model: the iiwa14 tables registered as DATA run the table-driven kernels (gato_model_register) -- same
        # workload, same bits as the compiled kernels (checked), measured the same way
        wm = make_config("bench", B=B * world)
        wm["plant"] = native.Model.load(ROOT / "gato_b200" / "models" / "iiwa14.gmdl").register("iiwa14_as_data")
        m_ms, _, m_st, m_solver, _ = device_timed(wm, sl, B, 10, 3)
        m_solver.close()
        m_tot = max_over_ranks(float(np.sum(m_ms)))
        extra["table_driven_model"] = {"workload": WORKLOAD + ", robot loaded from gato_b200/models/iiwa14.gmdl at run time", "value": B * world * 10 / (m_tot * 1e-3), "unit": "solves/s",
                                       "ms_per_step": m_tot / 10, "same_integer_outcomes_as_compiled_kernels": bool(np.array_equal(m_st["pcg_iters"], st["pcg_iters"]) and
                                                                                                                      np.array_equal(m_st["ls_step_size"], st["ls_step_size"]))}
        if world > 1:
            # strong scaling: the SAME 512 solves split over the ranks
            ws = make_config("bench", B=B)
            nb = B // world
            s_ms, _, _, s_solver, _ = device_timed(ws, slice(rank * nb, (rank + 1) * nb), nb, 10, 3)
            s_solver.close()
            s_tot = max_over_ranks(float(np.sum(s_ms)))
            extra["strong_scaling"] = {"total_solves": nb * world, "solves_per_gpu": nb, "value": nb * world * 10 / (s_tot * 1e-3), "unit": "solves/s", "ms_per_step": s_tot / 10}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    try:
        fp32_peak, fp32_src = native.measure_fp32_peak(local, True), "measured on this GPU: packed FFMA2 loop at full occupancy (gato_measure_fp32_peak)"
        fp32_scalar = native.measure_fp32_peak(local, False)
    except Exception:
        fp32_peak, fp32_src, fp32_scalar = FP32_NOMINAL_TFLOPS, "nominal 148 SM x 128 lanes x 2 x 1.965 GHz", None
    per_gpu_solves_per_s = B / (ms_per_step * 1e-3)
    work, traffic = kernel_work(50), ncu_traffic()
    roof_k = {}
    for k, t in kernels.items():
        wk = work[k]
        sec = t["us_per_launch"] * 1e-6
        tf, gbs = B * wk["flop"] / sec / 1e12, B * wk["bytes"] / sec / 1e9
        roof_k[k] = {"us_per_launch": t["us_per_launch"], "share_of_step": t["share_of_step"], "flop_per_launch": B * wk["flop"], "fp32_tflops": tf, "fp32_frac": tf / fp32_peak,
                     "algorithmic_bytes_per_launch": B * wk["bytes"], "hbm_gbs": gbs, "hbm_frac": gbs / hbm_peak,
                     "traffic": traffic.get(k, {}).get("dram_bytes_per_launch") if B == 512 else None}
    dom = max(kernels, key=lambda k: kernels[k]["share_of_step"]) if kernels else "k_pcg"
    rd = roof_k.get(dom, {})
    fps = flop_per_solve(4, 50)
    line = {
        "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "plant": "iiwa14", "knot_points": 32, "batch_per_gpu": B, "max_sqp_iters": 4, "max_pcg_iters": 50, "pcg_tol": -1.0,
                   "timing": "CUDA events per step on the solver stream; 256 MB L2 flush + state reset between steps, outside the event windows",
                   "sharding": "independent solves, contiguous rows per rank, no data-path collective"},
        "sqp_iters_per_s": value * 4,
        "latency_ms_p50": float(np.median(step_ms)), "latency_ms_p95": float(np.percentile(step_ms, 95)),
        "gpu_launches": int(launches),
        "clocks": clocks_summary(samples),
        # Dominant kernel.  This path is bound by the FP32 CUDA-core / issue side, not by HBM (about 10 KB of compulsory traffic against
        # 45 Mflop per solve), so the primary roofline is the kernel's algorithmic flops against the FP32 peak measured on this GPU; the HBM
        # fraction of the same launch (algorithmic bytes / measured copy bandwidth) and ncu's DRAM traffic are given beside it.
        "roofline": {"bound": "fp32", "kernel": dom, "achieved": rd.get("fp32_tflops"), "peak": fp32_peak, "unit": "TFLOP/s", "frac": rd.get("fp32_frac"),
                     "traffic": rd.get("traffic"), "algorithmic_flop_per_launch": rd.get("flop_per_launch"), "us_per_launch": rd.get("us_per_launch"),
                     "share_of_step": rd.get("share_of_step"), "peak_source": fp32_src, "fp32_tflops_scalar_ffma_measured": fp32_scalar,
                     "hbm": {"achieved": rd.get("hbm_gbs"), "peak": hbm_peak, "unit": "GB/s", "frac": rd.get("hbm_frac"), "algorithmic_bytes_per_launch": rd.get("algorithmic_bytes_per_launch"),
                             "peak_source": hbm_src},
                     "limiter": "dependent latency: two block-wide dot-product reductions and two LSU-bound window loads per PCG iteration (DESIGN.md section 4)"},
        "roofline_kernels": roof_k,
        "roofline_fp32": {"scope": "whole step", "achieved": per_gpu_solves_per_s * fps / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                          "frac": per_gpu_solves_per_s * fps / 1e12 / fp32_peak, "flop_per_solve": fps, "peak_source": fp32_src},
        "roofline_hbm_whole_step": {"achieved": per_gpu_solves_per_s * BYTES_PER_SOLVE / 1e9, "unit": "GB/s", "frac": per_gpu_solves_per_s * BYTES_PER_SOLVE / 1e9 / hbm_peak,
                                    "note": "compulsory I/O only (10 024 B per solve)"},
        "kernels": kernels,
    }
    line.update(extra)
    if e2e:
        line["e2e"] = e2e
    if not args.no_ref_gpu and world == 1:
        line["reference_gpu"] = reference_gpu_row(make_config("bench", B=B))
        line["parity_vs_stock_reference"] = parity_vs_stock_reference(native, B)
    if not args.no_cpu and world == 1:
        n = 1024
        rate, cores, times = cpu_oracle_rate(n, 3)
        line["cpu_baseline"] = {"value": rate, "unit": "solves/s", "cores": cores, "kind": "port",
                                "sample": f"{n} solves of the same workload, median of 3 passes ({sum(times):.1f} s of wall time on {cores} threads)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
