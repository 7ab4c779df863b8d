"""Small-batch latency (the MPC regime): wall-clock and device time per solve call for this library and, when oracle/_ref holds the
matching build, the unmodified reference CUDA build.  usage: python tools/latency_probe.py   (run on a GPU box)"""
import sys
import time

sys.path.insert(0, ".")
import numpy as np

from gato_b200 import native
from gato_b200.workloads import make_config


def timeit(fn, reps=200, warm=20):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(1e6 * (time.perf_counter() - t0))
    return float(np.median(ts)), float(np.percentile(ts, 95))


def main():
    rows = []
    for cfg, N, B in ((1, 8, 1), (1, 8, 16), (2, 32, 16), (2, 32, 128)):
        w = make_config(cfg, B=B, N=N)
        p = dict(w["params"])
        if cfg == 2:
            p.update(max_sqp_iters=1, max_pcg_iters=200, pcg_tol=1e-4)
        s = native.Solver(w["plant"], N, B, p, device=0)
        dev = []

        def mine():
            s.reset("dual")
            r = s.solve(w["xu"], w["xs"], w["ref"], w["dt"])
            dev.append(r["device_time_ms"])

        p50, p95 = timeit(mine)
        row = {"plant": w["plant"], "N": N, "B": B, "sqp_iters": int(p["max_sqp_iters"]), "gato_b200_wall_us_p50": p50, "gato_b200_wall_us_p95": p95,
               "gato_b200_device_us_p50": 1e3 * float(np.median(dev[-200:]))}
        try:
            from oracle.pyapi import Backend

            rs = Backend("ref", w["plant"], N, "fast").solver(B, p)

            def ref():
                rs.reset("dual")
                rs.solve(w["xu"], w["xs"], w["ref"], w["dt"])

            row["reference_wall_us_p50"], row["reference_wall_us_p95"] = timeit(ref)
        except Exception as e:  # the matching reference build did not travel
            row["reference"] = f"unavailable: {type(e).__name__}"
        rows.append(row)
        print(row, flush=True)


if __name__ == "__main__":
    main()
