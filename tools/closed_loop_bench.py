"""Closed-loop MPC benchmark (BASELINE.json config 5 / SURVEY.md section 8(d) cfg 5, per-GPU shard): iiwa14, N=32, one true state,
per-solve x0 noise, per-solve rho log-spaced and mu in {1,10}, figure-8 window advancing one knot per control step, reset_rho every
step, plant = the solver's own sim_forward.  Times the fused device step (gato_mpc_step) against the reference's host-side
composition of the same step through the plain calls of the same library (host_mpc_step); both produce identical trajectories.
usage: python tools/closed_loop_bench.py [--batch 1024] [--steps 200]      (prints one JSON line)"""
import argparse
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np

from gato_b200 import native
from gato_b200.bsqp.mpc import DeviceMPC, host_mpc_step
from gato_b200.workloads import figure8, make_config


def run(kind, w, steps, B):
    N, dt, p = w["N"], w["dt"], w["params"]
    s = native.Solver(w["plant"], N, B, p, device=0)
    s.set_batch("rho", w["extra"]["rho"], True)
    s.set_batch("mu", w["extra"]["mu"])
    fext = np.zeros((B, 6), np.float32)
    fext[:, 2] = np.linspace(-10, 10, B)
    s.set_batch("f_ext", fext)
    nx, nu, traj = s.d["nx"], s.d["nu"], s.d["traj"]
    fig = figure8(dt).reshape(-1, 6)
    off = (w["xs"] - w["xs"].mean(0, keepdims=True)).astype(np.float32)
    x = np.zeros(nx, np.float32)
    # the simulated plant saturates its actuators at the robot's torque limits (iiwa14_plant.cuh:57-70), as the real one does; without it
    # this synthetic scenario (start pose 0.8 m away from the figure-8, cheap control) runs away within ~60 steps
    u_max = np.array([320.0, 320.0, 176.0, 176.0, 110.0, 40.0, 40.0], np.float32)
    true_hyp = B // 2
    XU = np.zeros((B, traj), np.float32)
    lat, dev_ms, best_hist = [], [], []
    if kind == "device":
        mpc = DeviceMPC(s, dt)
        s.mpc_set_state_offsets(off)
        r = mpc.warm_start(x, fig[:N].reshape(-1))
        xu_best = r["XU_best"]
    else:
        s.reset("dual")
        r, _, _ = host_mpc_step(s, XU, x, fig[:N].reshape(-1), None, None, 0.0, dt, reset_rho=False, offsets=off)
        xu_best = XU[0].copy()
    for k in range(1, steps + 1):
        x_last, u_last = x.copy(), np.clip(xu_best[nx:nx + nu], -u_max, u_max).astype(np.float32)
        x = s.sim_forward(x_last, u_last, dt)[true_hyp].copy()
        ref_w = fig[k % 500:k % 500 + N].reshape(-1)
        t0 = time.perf_counter()
        if kind == "device":
            r = mpc.step(x, ref_w, x_last, u_last, dt)
            xu_best, best = r["XU_best"], r["best_id"]
        else:
            r, best, _ = host_mpc_step(s, XU, x, ref_w, x_last, u_last, dt, dt, offsets=off)
            xu_best = XU[0].copy()
        lat.append(1e3 * (time.perf_counter() - t0))
        dev_ms.append(r["device_time_ms"])
        best_hist.append(best)
    return np.array(lat), np.array(dev_ms), xu_best, best_hist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=200)
    a = ap.parse_args()
    w = make_config(5, B=a.batch)
    out = {"workload": f"cfg5 closed loop iiwa14 N=32 B={a.batch} (per-GPU shard), {a.steps} control steps, reset_rho each step"}
    res = {}
    for kind in ("host", "device"):
        lat, dev_ms, xu, best = run(kind, w, a.steps, a.batch)
        res[kind] = (xu, best)
        out[kind] = {"step_ms_p50": float(np.median(lat)), "step_ms_p95": float(np.percentile(lat, 95)), "device_ms_p50": float(np.median(dev_ms)),
                     "solves_per_s": a.batch / (np.median(lat) * 1e-3)}
    out["finite"] = bool(np.isfinite(res["host"][0]).all() and np.isfinite(res["device"][0]).all())
    out["identical_trajectories"] = bool(np.array_equal(res["host"][0], res["device"][0]) and res["host"][1] == res["device"][1])
    out["speedup_step_p50"] = out["host"]["step_ms_p50"] / out["device"]["step_ms_p50"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
