"""indy7: the reference's merit kernel overruns its shared memory and faults on B200 (profiles/r01_reference_indy7_sanitizer.txt), so whole
solves cannot be compared; every other stage of the UNMODIFIED reference (IEEE build) is compared with the oracle here, bit for bit.
usage: python tools/pin_indy7_stages.py   (GPU box; needs oracle/_ref/libgref_indy7_N32_ieee.so from `oracle/build_ref.sh pin`)"""
import sys

sys.path.insert(0, ".")
import numpy as np

from gato_b200.workloads import make_config
from oracle.pyapi import Backend, ensure_oracle_built

ensure_oracle_built()


def nm(a, b):
    return int((np.asarray(a).view(np.uint32) != np.asarray(b).view(np.uint32)).sum())


plant, N, B = "indy7", 32, 16
w = make_config(3, B=B, N=N)
ref, orc = Backend("ref", plant, N, "ieee"), Backend("oracle", plant, N)
rng = np.random.default_rng(9)
xu = (w["xu"] + rng.normal(0, 0.05, w["xu"].shape)).astype(np.float32)
p = dict(w["params"], vel_lim_cost=0.002, ctrl_lim_cost=0.001)
fext = rng.normal(0, 2, (B, 6)).astype(np.float32)
rho = np.full(B, p["rho"], np.float32)
rho[1::2] = 1e-3
kr, ko = ref.stage_kkt(B, xu, w["xs"], w["ref"], fext, w["dt"], p), orc.stage_kkt(B, xu, w["xs"], w["ref"], fext, w["dt"], p)
print("kkt      ", {k: nm(kr[k], ko[k]) for k in ko})
sr, so = ref.stage_schur(B, ko, rho), orc.stage_schur(B, ko, rho)
print("schur    ", {k: nm(sr[k], so[k]) for k in so})
lam0 = np.zeros((B, orc.d["vecp"]), np.float32)
for eps, cap in ((1e-4, 200), (-1.0, 30)):
    lr, ir = ref.stage_pcg(B, so["S"], so["Pinv"], so["gamma"], lam0, np.full(B, eps, np.float32), cap)
    lo, io = orc.stage_pcg(B, so["S"], so["Pinv"], so["gamma"], lam0, np.full(B, eps, np.float32), cap)
    print(f"pcg eps={eps} cap={cap}: iteration counts equal {np.array_equal(ir, io)}, lambda mismatches {nm(lr, lo)}")
dzr = ref.stage_dz(B, lo, so["Qinv"], so["Rinv"], ko["q"], ko["r"], ko["A"], ko["Bm"])
dzo = orc.stage_dz(B, lo, so["Qinv"], so["Rinv"], ko["q"], ko["r"], ko["A"], ko["Bm"])
print("dz       ", [nm(a, b) for a, b in zip(dzr, dzo)])
merit8 = rng.uniform(1, 3, (B, 8)).astype(np.float32)
mi = rng.uniform(1.5, 2.5, B).astype(np.float32)
for adapt in (1, 0):
    lsr = ref.stage_linesearch(B, xu, dzo[0], merit8, mi, rho, np.ones(B, np.float32), adapt)
    lso = orc.stage_linesearch(B, xu, dzo[0], merit8, mi, rho, np.ones(B, np.float32), adapt)
    print(f"linesearch adapt={adapt}", {k: nm(lsr[k], lso[k]) for k in lso})
# merit: the unmodified kernel launched by the harness with enough dynamic shared memory (GREF_MERIT_EXTRA_SMEM, see oracle/ref_harness.cu);
# the reference accumulates the knots with atomics in arbitrary order, so the comparison is to float rounding of a 32-term sum
import os

os.environ["GREF_MERIT_EXTRA_SMEM"] = "4096"
mu = np.full(B, 10.0, np.float32)
for na in (1, 8):
    mr = ref.stage_merit(B, xu, dzo[0], w["xs"], w["ref"], mu, fext, w["dt"], p, na)
    mo = orc.stage_merit(B, xu, dzo[0], w["xs"], w["ref"], mu, fext, w["dt"], p, na)
    print(f"merit x{na}: max rel diff {float(np.max(np.abs(mr - mo) / np.abs(mo))):.2e}, bit-identical entries {int((mr.view(np.uint32) == mo.view(np.uint32)).sum())}/{mr.size}")
d_r, d_o = ref.dyn_dump(xu[:, :12], xu[:, 12:18], fext), orc.dyn_dump(xu[:, :12], xu[:, 12:18], fext)
# the pose's orientation part (roll, pitch, yaw and its Jacobian rows) never enters the solver (iiwa14_plant.cuh:313-319) and is not restated
ee_r, ee_o = d_r["ee"][:, :3], d_o["ee"][:, :3]
J_r, J_o = d_r["dee"].reshape(B, -1, 6)[:, :, :3], d_o["dee"].reshape(B, -1, 6)[:, :, :3]
print("dynamics ", {"qdd": nm(d_r["qdd"], d_o["qdd"]), "dqdd": nm(d_r["dqdd"], d_o["dqdd"]), "ee xyz": nm(ee_r, ee_o), "d ee xyz": nm(J_r, J_o)})
