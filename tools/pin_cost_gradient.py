"""Which FMA contraction did nvcc choose for the cost gradient "s_qk[i] = a*b; s_qk[i] += lim * barrier'" (iiwa14_plant.cuh:371-378,
indy7_plant.cuh) in the regular and in the terminal (computeR = false) instantiation?  Runs the UNMODIFIED reference's KKT stage (IEEE build,
oracle/_ref) on a GPU box with non-zero limit weights and scores the candidate expression trees, evaluated in numpy from the oracle's
end-effector Jacobian, against the reference's bits.  Result (B200, both plants): regular = fma(a, b, round(lim*barrier')), terminal =
fma(lim, barrier', round(a*b)); the oracle and the CUDA path implement exactly that.   usage: python tools/pin_cost_gradient.py"""
import sys, itertools
sys.path.insert(0, ".")
import numpy as np
from oracle.pyapi import Backend, ensure_oracle_built
from gato_b200.workloads import make_config
ensure_oracle_built()
f32 = np.float32
def fma(a, b, c): return f32(np.float64(f32(a)) * np.float64(f32(b)) + np.float64(f32(c)))
margin = float(f32(-0.1))
LIM = {"iiwa14": dict(J=[2.96706, 2.09440, 2.96706, 2.09440, 2.96706, 2.09440, 3.05433], V=[1.48353, 1.48353, 1.74533, 1.30900, 2.26893, 2.35619, 2.35619]),
       "indy7": dict(J=[3.0543, 3.0543, 3.0543, 3.0543, 3.0543, 3.7520], V=[2.61, 2.61, 2.61, 3.14, 3.14, 3.14])}
def bgrad(plant, q, lo, hi):
    dmin, dmax = f32(q - lo), f32(hi - q)
    if plant == "iiwa14":
        eps = f32(1e-6)
        dmin = max(dmin, eps) if dmin >= 0 else min(dmin, -eps)
        dmax = max(dmax, eps) if dmax >= 0 else min(dmax, -eps)
    else:
        dmin = f32(1e-6) if float(dmin) <= 1e-6 else dmin
        dmax = f32(1e-6) if float(dmax) <= 1e-6 else dmax
    return f32(f32(f32(-1.0) / dmin) + f32(f32(1.0) / dmax))
hv = {"a": lambda g, e: fma(g[2], e[2], fma(g[0], e[0], f32(g[1] * e[1]))), "b": lambda g, e: fma(g[2], e[2], fma(g[1], e[1], f32(g[0] * e[0]))),
      "c": lambda g, e: f32(f32(f32(g[0] * e[0]) + f32(g[1] * e[1])) + f32(g[2] * e[2])), "d": lambda g, e: fma(g[0], e[0], fma(g[1], e[1], f32(g[2] * e[2])))}
fv = {"A": lambda h, w, l, bg: fma(h, w, f32(l * bg)), "B": lambda h, w, l, bg: fma(l, bg, f32(h * w)), "C": lambda h, w, l, bg: f32(f32(h * w) + f32(l * bg))}
CASES = [("iiwa14", 2, n, 3) for n in (3, 4, 6, 8, 9, 10, 11, 12, 16, 32, 64, 128)] + [("indy7", 3, n, 3) for n in (8, 16, 32)]
for plant, cfg, N, pseed in CASES:
    nq = 7 if plant == "iiwa14" else 6
    nx, st = 2 * nq, 3 * nq
    B = 8 if N == 128 else 16
    w = make_config(cfg, B=B, N=N)
    try:
        ref, orc = Backend("ref", plant, N, "ieee"), Backend("oracle", plant, N)
    except Exception as ex:
        print(plant, N, "skipped:", type(ex).__name__)
        continue
    p = dict(w["params"], max_sqp_iters=1, vel_lim_cost=0.002, ctrl_lim_cost=0.001, q_lim_cost=0.013)
    rng = np.random.default_rng(pseed)
    xu = (w["xu"] + rng.normal(0, 0.05, w["xu"].shape)).astype(np.float32)
    fext = np.zeros((B, 6), np.float32)
    kr, ko = ref.stage_kkt(B, xu, w["xs"], w["ref"], fext, w["dt"], p), orc.stage_kkt(B, xu, w["xs"], w["ref"], fext, w["dt"], p)
    print(plant, "oracle vs ref mismatches:", {k: int((kr[k].view(np.uint32) != ko[k].view(np.uint32)).sum()) for k in ko})
    qr = kr["q"].reshape(B, N, nx)
    jl = [(f32(-J - margin), f32(J + margin)) for J in LIM[plant]["J"]]
    vl = [(f32(-V - margin), f32(V + margin)) for V in LIM[plant]["V"]]
    score = {}
    for kn, label in [(N - 1, "terminal")] + [(k_, "regular") for k_ in range(N - 1)]:
        for b in range(B):
            ks = N - 2 if kn == N - 1 else kn
            x = xu[b].reshape(-1)[ks * st: ks * st + nx]
            d = orc.dyn_dump(x[None, :], np.zeros((1, nq), np.float32), np.zeros((1, 6), np.float32))
            ee, dee = d["ee"][0], d["dee"][0].reshape(nq, 6)
            e = (ee[:3] - w["ref"][b].reshape(N, 6)[kn, :3]).astype(np.float32)
            for i in range(nq):
                g = dee[i, :3]
                bg = bgrad(plant, x[i], *jl[i])
                for hk, fk in itertools.product(hv, fv):
                    val = fv[fk](hv[hk](g, e), f32(p["q_cost"]), f32(p["q_lim_cost"]), bg)
                    score[(label, "pos", hk + fk)] = score.get((label, "pos", hk + fk), 0) + int(val.view(np.uint32) == qr[b, kn, i].view(np.uint32))
                bv = bgrad(plant, x[nq + i], *vl[i])
                for fk in fv:
                    val = fv[fk](f32(p["qd_cost"]), x[nq + i], f32(p["vel_lim_cost"]), bv)
                    score[(label, "vel", fk)] = score.get((label, "vel", fk), 0) + int(val.view(np.uint32) == qr[b, kn, nq + i].view(np.uint32))
    for label in ("terminal", "regular"):
        for part in ("pos", "vel"):
            print(" ", label, part, sorted(((v, k[2]) for k, v in score.items() if k[0] == label and k[1] == part), reverse=True)[:4], "of", B * nq * (1 if label == "terminal" else N - 1))
