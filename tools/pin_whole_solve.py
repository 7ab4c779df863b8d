"""Whole solves of the UNMODIFIED reference (IEEE build, oracle/_ref) against the CPU oracle on a GPU box: number of solves whose trajectory
differs in any bit, per iteration count.  usage: python tools/pin_whole_solve.py"""
import sys

sys.path.insert(0, ".")
import numpy as np

from gato_b200.workloads import make_config
from oracle.pyapi import Backend, ensure_oracle_built

ensure_oracle_built()


def nm(a, b):
    return int((np.asarray(a).view(np.uint32) != np.asarray(b).view(np.uint32)).sum())


for plant, N, B, cfg in (("iiwa14", 3, 16, 1), ("iiwa14", 8, 16, 1), ("iiwa14", 9, 16, 2), ("iiwa14", 12, 16, 2), ("iiwa14", 32, 16, 2), ("iiwa14", 64, 16, 2)):
    w = make_config(cfg, B=B, N=N)
    ref, orc = Backend("ref", plant, N, "ieee"), Backend("oracle", plant, N)
    rng = np.random.default_rng(3)
    for label, xu in (("warm start", w["xu"]), ("perturbed", (w["xu"] + rng.normal(0, 0.05, w["xu"].shape)).astype(np.float32))):
        for its in (1, 2, 4):
            for extra in (dict(), dict(vel_lim_cost=0.002, ctrl_lim_cost=0.001, max_pcg_iters=200, pcg_tol=1e-4)):
                p = dict(w["params"], max_sqp_iters=its, **extra)
                r = ref.solver(B, p).solve(xu, w["xs"], w["ref"], w["dt"])
                o = orc.solver(B, p).solve(xu, w["xs"], w["ref"], w["dt"])
                bad = [b for b in range(B) if nm(r["XU"][b], o["XU"][b])]
                print(f"{plant} N={N} {label:10s} iters={its} {'limits+tol' if extra else 'default   '}: solves with a differing trajectory bit: {len(bad)}/{B} {bad}; "
                      f"pcg_iters equal {np.array_equal(r['pcg_iters'], o['pcg_iters'])}; steps equal {np.array_equal(r['ls_step_size'], o['ls_step_size'])}; "
                      f"max rel diff {np.max(np.abs(r['XU'] - o['XU'])) / np.max(np.abs(o['XU'])):.1e}", flush=True)
