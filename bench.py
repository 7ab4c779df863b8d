#!/usr/bin/env python3
"""bench.py — batched SQP solves/s on the BASELINE.json headline shape (iiwa14, N=32, batch 512 per GPU).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W   (CPU baseline arm: the oracle port on the host cores)

One "step" = one gato_solve of the whole batch (max_sqp_iters=4, max_pcg_iters=50, pcg_tol=-1: fixed iteration
caps so every implementation does the same work, SURVEY.md §8(d) cfg 2 at batch 512).  `value` times the solve with
inputs resident in HBM (CUDA events on the solver stream, per step, L2 flushed between steps); `e2e` times the same
solve through the host-buffer entry point (pinned H2D of xu/x_s/ref + D2H of xu and stats inside the window).
Prints ONE JSON line on rank 0.
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

from gato_b200.workloads import make_config  # noqa: E402

WORKLOAD = "iiwa14_N32_B512_sqp4_pcg50_fixedcaps"
METRIC = "batched SQP solves/sec (iiwa14, N=32, batch 512)"

# algorithmic work per solve of the headline workload: 4 SQP iterations x 50 PCG iterations
FLOP_PER_SQP_ITER = 7.06e6 + 0.08e6 * 51
FLOP_MERIT_EXTRA = 0.88e6
FLOP_PER_SOLVE = 4 * FLOP_PER_SQP_ITER + FLOP_MERIT_EXTRA
BYTES_PER_SOLVE = 10024  # compulsory HBM I/O per solve (xu, ref, x_s, f_ext, lambda in; xu, lambda, stats out)
# algorithmic bytes ONE k_pcg launch moves per solve (iiwa14, N=32): rows of S, main blocks of P^-1, gamma, lambda
# in/out, and for the primal step A, B, Q^-1, R^-1, q, r (read-modify-write) and dz
_N, _NX, _NU = 32, 14, 7
PCG_FLOATS_PER_SOLVE = (_N * 3 * _NX * _NX + _N * _NX * _NX + 3 * (_N + 2) * _NX + (_N - 1) * _NX * _NX + (_N - 1) * _NX * _NU + _N * _NX * _NX + (_N - 1) * _NU * _NU
                        + 2 * _N * _NX + 2 * (_N - 1) * _NU + (_N * (_NX + _NU) - _NU))
PCG_BYTES_PER_SOLVE = 4 * PCG_FLOATS_PER_SOLVE
# dram__bytes_read.sum + dram__bytes_write.sum of one k_pcg launch at batch 512 (ncu --set full, cold caches; profiles/r01_v14_k_pcg_raw.csv).
# Above the algorithmic bytes because the 56-byte main blocks of P^-1 sit inside 168-byte rows: DRAM sectors are fetched whole.
PCG_NCU_TRAFFIC_BYTES_B512 = 115.122432e6 + 4.426496e6
FP32_NOMINAL_TFLOPS = 74.4  # 148 SM x 128 lanes x 2 x 1.965 GHz (no fp32 number in MEASURED_PEAKS.json)


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
    """Oracle port (oracle/bsqp_oracle.cpp) on the host cores, same workload; returns (solves/s, cores)."""
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
    """--impl reference: the reference has no CPU solver; its CPU arm is the oracle port, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 128
    rate, cores, times = cpu_oracle_rate(n, args.warmup + args.steps)
    t = times[args.warmup:]
    ms = 1e3 * float(np.mean(t))
    rate = n / float(np.mean(t))
    sample = f"{n} solves of the {WORKLOAD} workload per step (first {n} rows of the seeded batch)"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "plant": "iiwa14", "knot_points": 32, "batch_per_step": n, "max_sqp_iters": 4, "max_pcg_iters": 50, "pcg_tol": -1.0},
        "cpu_baseline": {"value": rate, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def reference_gpu_row(w):
    """Reference CUDA build (unmodified sources, sm_100, -use_fast_math) on this GPU, if oracle/_ref travelled with the repo."""
    try:
        from oracle.pyapi import Backend

        be = Backend("ref", w["plant"], w["N"], "fast")
        if w["B"] not in be.batches:
            return None
        sv = be.solver(w["B"], w["params"])
        sv.solve_timed(w["xu"], w["xs"], w["ref"], w["dt"], 3, True)
        ev, us = sv.solve_timed(w["xu"], w["xs"], w["ref"], w["dt"], 10, True)
        sv.close()
        ms = float(np.median(ev))
        return {"value": w["B"] / (ms * 1e-3), "unit": "solves/s", "ms_per_step": ms, "build": "reference sources unmodified, nvcc -O3 -use_fast_math -DNDEBUG sm_100 (oracle/build_ref.sh)"}
    except Exception as e:  # pragma: no cover
        return {"unavailable": str(e)[:200]}


def roofline_pcg(kernels, B, hbm_peak, peak_source):
    k = kernels.get("k_pcg")
    if not k:
        return {"bound": "hbm", "achieved": None, "peak": hbm_peak, "unit": "GB/s", "frac": None, "traffic": None, "kernel": "k_pcg"}
    gbs = B * PCG_BYTES_PER_SOLVE / (k["us_per_launch"] * 1e-6) / 1e9
    return {"bound": "hbm", "kernel": "k_pcg", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
            "traffic": PCG_NCU_TRAFFIC_BYTES_B512 * B / 512 if B == 512 else None, "algorithmic_bytes_per_launch": B * PCG_BYTES_PER_SOLVE,
            "us_per_launch": k["us_per_launch"], "share_of_step": k["share_of_step"], "peak_source": peak_source,
            "limiter": "issue/latency (4 CTA-wide barriers and two dependent reduction trees per PCG iteration), not HBM"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="gato_b200")
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-ref-gpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
    # every rank solves its own shard of the (synthetic) global batch: independent problems, no data-path collective
    w = make_config("bench", B=B * world)
    sl = slice(rank * B, (rank + 1) * B)
    xu0 = torch.from_numpy(w["xu"][sl].copy()).cuda()
    xs = torch.from_numpy(w["xs"][sl].copy()).cuda()
    ref = torch.from_numpy(w["ref"][sl].copy()).cuda()
    xu = xu0.clone()
    stream = torch.cuda.Stream()  # a real (non-default) stream shared by torch and the solver, so torch events see the solver's work
    torch.cuda.set_stream(stream)
    solver = native.Solver(w["plant"], w["N"], B, w["params"], device=local, stream=stream.cuda_stream)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")  # > 126 MB L2

    def one_step(timed):
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

    for _ in range(args.warmup):
        one_step(False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    stop, samples = threading.Event(), []
    th = threading.Thread(target=sample_clocks, args=(stop, samples), daemon=True)
    th.start()
    l0 = solver.kernel_launches()
    step_ms = []
    for _ in range(args.steps):
        ms, st = one_step(True)
        step_ms.append(ms)
    torch.cuda.synchronize()
    launches = solver.kernel_launches() - l0
    stop.set()
    th.join(timeout=2)
    total_ms = float(np.sum(step_ms))
    if world > 1:
        dist.barrier()
        t = torch.tensor([total_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = B * world * args.steps / (total_ms * 1e-3)
    ms_per_step = total_ms / args.steps

    # ---- per-kernel durations: a few extra steps with the library's event-per-launch instrumentation (not part of the timed region) ----
    kernels = {}
    if rank == 0:
        solver.set_kernel_timing(True)
        acc = {}
        reps = 5
        for _ in range(reps):
            one_step(False)
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
        tot = float(np.sum(e_ms))
        if world > 1:
            t = torch.tensor([tot], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tot = float(t.item())
        n_it = int(w["params"]["max_sqp_iters"])
        e2e = {"value": B * world * args.steps / (tot * 1e-3), "unit": "solves/s",
               "h2d_bytes_per_step": int(4 * (h_xu0.numel() + h_xs.numel() + h_ref.numel())),
               "d2h_bytes_per_step": int(4 * h_xu0.numel() + 4 * B * (3 * n_it + 3) + 4 * n_it)}

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
    per_gpu_solves_per_s = B / (ms_per_step * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "plant": "iiwa14", "knot_points": 32, "batch_per_gpu": B, "max_sqp_iters": 4, "max_pcg_iters": 50, "pcg_tol": -1.0,
                   "timing": "CUDA events per step on the solver stream; 256 MB L2 flush + state reset between steps, outside the event windows",
                   "sharding": "independent solves, contiguous rows per rank, no data-path collective"},
        "sqp_iters_per_s": value * 4,
        "latency_ms_p50": float(np.median(step_ms)),
        "gpu_launches": int(launches),
        "clocks": clocks_summary(samples),
        # roofline of the dominant kernel (k_pcg, about half of the step): algorithmic bytes of one launch / its CUDA-event duration against
        # the measured HBM copy peak.  The kernel is issue/latency bound, not HBM bound: the whole path has ~10 KB of
        # compulsory HBM traffic and ~45 Mflop per solve, so the FP32 fraction of the whole step is reported next to it.
        "roofline": roofline_pcg(kernels, B, hbm_peak, "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s"),
        "roofline_fp32": {"scope": "whole step", "achieved": per_gpu_solves_per_s * FLOP_PER_SOLVE / 1e12, "peak": FP32_NOMINAL_TFLOPS, "unit": "TFLOP/s",
                          "frac": per_gpu_solves_per_s * FLOP_PER_SOLVE / 1e12 / FP32_NOMINAL_TFLOPS, "peak_source": "nominal 148 SM x 128 lanes x 2 x 1.965 GHz"},
        "roofline_hbm_whole_step": {"achieved": per_gpu_solves_per_s * BYTES_PER_SOLVE / 1e9, "unit": "GB/s", "frac": per_gpu_solves_per_s * BYTES_PER_SOLVE / 1e9 / hbm_peak,
                                    "note": "compulsory I/O only (10 024 B per solve)"},
        "kernels": kernels,
    }
    if e2e:
        line["e2e"] = e2e
    if not args.no_ref_gpu and world == 1:
        line["reference_gpu"] = reference_gpu_row(make_config("bench", B=B))
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
