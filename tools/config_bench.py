"""Device time per solve call for BASELINE.json configs 1-4 (inputs resident in HBM, CUDA events inside the library), this library
vs the unmodified reference CUDA build where oracle/_ref holds the matching library.  usage: python tools/config_bench.py  (GPU box)"""
import sys

sys.path.insert(0, ".")
import numpy as np

from gato_b200 import native
from gato_b200.workloads import make_config


def main():
    if len(sys.argv) < 2:  # one process per (config, implementation): the reference's indy7 / N=128 builds can fault and poison the CUDA context
        import subprocess

        for cfg in (1, 2, 3, 4):
            subprocess.run([sys.executable, __file__, str(cfg), "mine"])
            subprocess.run([sys.executable, __file__, str(cfg), "ref"])
        return
    which = sys.argv[2]
    for cfg in (int(sys.argv[1]),):
        w = make_config(cfg)
        B, N = w["B"], w["N"]
        if which == "ref":
            try:
                from oracle.pyapi import Backend

                rs = Backend("ref", w["plant"], N, "fast").solver(B, w["params"])
                t = []
                for i in range(6):
                    rs.reset("dual"), rs.reset("rho")
                    rr = rs.solve(w["xu"], w["xs"], w["ref"], w["dt"])
                    if i >= 2:
                        t.append(rr["sqp_time_us"] * 1e-3)
                print({"cfg": cfg, "reference_ms": float(np.median(t)), "reference_solves_per_s": B / (np.median(t) * 1e-3)}, flush=True)
            except BaseException as e:
                print({"cfg": cfg, "reference": f"unavailable ({type(e).__name__}: {str(e)[:60]})"}, flush=True)
            return
        s = native.Solver(w["plant"], N, B, w["params"], device=0)
        ms = []
        for i in range(13):
            s.reset("dual"), s.reset("rho")
            r = s.solve(w["xu"], w["xs"], w["ref"], w["dt"])
            if i >= 3:
                ms.append(r["device_time_ms"])
        row = {"cfg": cfg, "plant": w["plant"], "N": N, "B": B, "sqp_iters": int(w["params"]["max_sqp_iters"]), "pcg_iters_mean": float(r["pcg_iters"].mean()),
               "gato_b200_ms": float(np.median(ms)), "gato_b200_solves_per_s": float(B / (np.median(ms) * 1e-3))}
        print(row, flush=True)


if __name__ == "__main__":
    main()
