#!/usr/bin/env python3
"""BASELINE.json config 5 for real: closed-loop figure-8 tracking, iiwa14 N=32, the hypothesis batch (per-solve rho / mu / f_ext / state
offsets) sharded over the ranks of a torchrun launch, one rank per GPU (gato_b200.sharding.ShardedMPC: one NCCL broadcast of the measurement
and one NCCL all-gather of the shards' winner records per control step; no collective inside the solve).

  torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/closed_loop_multi_gpu.py --per-gpu 1024 --steps 200
  ... --check : additionally run the SAME loop in one process on rank 0's GPU over the whole batch and require identical winners,
                trajectories and errors at every step (bit-for-bit)
Prints one JSON line on rank 0 (per-step wall time p50 / p95, device time, hypotheses per second)."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def problem(total, N, seed=4):
    from gato_b200.workloads import DEFAULT_SOLVER_PARAMS, figure8

    dt = 0.01
    p = dict(DEFAULT_SOLVER_PARAMS, dt=dt)  # 1 SQP iteration, PCG <= 200 at tolerance 1e-4 (python/bsqp/config.py:35-50)
    rng = np.random.default_rng(seed)
    fext = rng.normal(0, 5.0, (total, 6)).astype(np.float32)
    fext[0] = 0
    rho = np.logspace(-8, 1, total).astype(np.float32)
    mu = np.where(np.arange(total) % 2 == 0, 1.0, 10.0).astype(np.float32)
    off = rng.normal(0, 0.01, (total, 14)).astype(np.float32)
    off[0] = 0
    return p, dt, fext, rho, mu, off, figure8(dt).reshape(-1, 6)


def closed_loop(dist, rank, world, per_gpu, N, steps, true_hyp, saturate=True):
    import torch

    from gato_b200 import native
    from gato_b200.sharding import NativeShardEngine, ShardedMPC

    total = per_gpu * world
    p, dt, fext, rho, mu, off, fig = problem(total, N)
    sl = slice(rank * per_gpu, (rank + 1) * per_gpu)
    stream = torch.cuda.current_stream()
    s = native.Solver("iiwa14", N, per_gpu, p, device=torch.cuda.current_device(), stream=stream.cuda_stream)
    s.set_batch("f_ext", fext[sl])
    s.set_batch("rho", rho[sl], True)
    s.set_batch("mu", mu[sl])
    s.reset("dual")
    nx, nu, traj = s.d["nx"], s.d["nu"], s.d["traj"]
    s.mpc_set_warm_start(np.zeros(traj, np.float32))
    s.mpc_set_state_offsets(off[sl])
    # the "real" robot: hypothesis `true_hyp` of the solver's own simulator (SURVEY.md section 8(d) cfg 5), simulated on every rank alike
    plant = native.Solver("iiwa14", N, 1, p, device=torch.cuda.current_device(), stream=stream.cuda_stream)
    plant.set_batch("f_ext", fext[true_hyp:true_hyp + 1])
    tau_max = np.array([320, 320, 176, 176, 110, 40, 40], np.float32)
    mpc = ShardedMPC(NativeShardEngine(s), dt, dist)
    x = np.zeros(nx, np.float32)
    res = mpc.step(x, fig[:N].reshape(-1), None, None, 0.0, reset_rho=False)
    ids, xus, errs, wall, dev_ms, pcg = [res["best_id"]], [res["XU_best"]], [res["best_error"]], [], [], []
    for step in range(1, steps + 1):
        x_last = x.copy()
        u_last = res["XU_best"][nx:nx + nu].copy()
        if saturate:
            u_last = np.clip(u_last, -tau_max, tau_max)
        x = plant.sim_forward(x_last, u_last, dt)[0].copy()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        res = mpc.step(x, fig[step:step + N].reshape(-1), x_last, u_last, dt)
        wall.append(time.perf_counter() - t0)
        dev_ms.append(res["device_time_ms"])
        pcg.append(float(res["pcg_iters"].mean()))
        ids.append(res["best_id"]), xus.append(res["XU_best"]), errs.append(res["best_error"])
    return dict(ids=np.array(ids), xus=np.stack(xus), errs=np.array(errs), wall=np.array(wall), dev_ms=np.array(dev_ms), pcg=np.array(pcg), finite=bool(np.isfinite(np.stack(xus)).all()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--per-gpu", type=int, default=1024)
    ap.add_argument("--knots", type=int, default=32)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--true-hyp", type=int, default=5)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    torch.cuda.set_stream(torch.cuda.Stream())
    d = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        d = dist
    r = closed_loop(d, rank, world, a.per_gpu, a.knots, a.steps, a.true_hyp)
    w = r["wall"]
    if world > 1:
        t = torch.tensor([float(np.sum(w))], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot = float(t.item())
    else:
        tot = float(np.sum(w))
    line = None
    if rank == 0:
        line = {"workload": f"closed-loop figure-8 MPC, iiwa14 N={a.knots}, {a.per_gpu} hypotheses per GPU x {world} GPU(s), default parameters, reset_rho every step",
                "n_gpus": world, "hypotheses_total": a.per_gpu * world, "control_steps": a.steps, "ms_per_step_p50": 1e3 * float(np.median(w)), "ms_per_step_p95": 1e3 * float(np.percentile(w, 95)),
                "ms_per_step_mean_max_over_ranks": 1e3 * tot / a.steps, "device_ms_per_step_p50": float(np.median(r["dev_ms"])), "hypothesis_solves_per_s": a.per_gpu * world * a.steps / tot,
                "pcg_iters_mean": float(r["pcg"].mean()), "true_hypothesis": a.true_hyp, "true_hypothesis_selected_fraction": float((r["ids"][1:] == a.true_hyp).mean()),
                "all_finite": r["finite"], "collectives_per_step": "1 NCCL broadcast (227 floats) + 1 NCCL all-gather (one 2.7 KB winner record per rank)" if world > 1 else "none"}
    if a.check:
        ok = True
        if rank == 0:
            one = closed_loop(None, 0, 1, a.per_gpu * world, a.knots, a.steps, a.true_hyp)
            ok = bool(np.array_equal(one["ids"], r["ids"]) and np.array_equal(one["xus"], r["xus"]) and np.array_equal(one["errs"], r["errs"]))
            line["equals_single_process_bit_for_bit"] = ok
        if world > 1:
            dist.barrier()
        if rank == 0 and not ok:
            print(json.dumps(line), flush=True)
            raise SystemExit("sharded closed loop differs from the single-process loop")
    if rank == 0:
        print(json.dumps(line), flush=True)
        if a.out:
            Path(a.out).write_text(json.dumps(line) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
