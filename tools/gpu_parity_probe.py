#!/usr/bin/env python3
"""Dev probe (runs on the GPU box): product kernels vs CPU oracle, stage by stage and whole solves; quick timings."""
import sys, time, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from oracle.pyapi import Backend
from gato_b200.native import GatoBackend
from gato_b200.workloads import make_config, DEFAULT_SOLVER_PARAMS

def mism(a, b):
    a = np.asarray(a); b = np.asarray(b)
    bad = (a != b) & ~(np.isnan(a) & np.isnan(b))
    rel = float(np.abs(a.astype(np.float64) - b).max() / max(1e-30, np.abs(b).max())) if a.size else 0.0
    return f"{int(bad.sum())}/{a.size} rel={rel:.2e}"

def stages(plant, N, cfg, B=4):
    print(f"==== stages {plant} N={N}", flush=True)
    o, g = Backend('oracle', plant, N), GatoBackend(plant, N)
    d = o.d
    rng = np.random.default_rng(11)
    w = make_config(cfg, B=B, N=N)
    xu = w['xu'] + rng.normal(0, 0.05, w['xu'].shape).astype(np.float32)
    fext = rng.normal(0, 2, (B, 6)).astype(np.float32); fext[0] = 0
    p = dict(w['params'])
    rho = np.full(B, p['rho'], np.float32); rho[1::2] = 1e-3
    mu = np.full(B, 10, np.float32)
    ko, kg = o.stage_kkt(B, xu, w['xs'], w['ref'], fext, w['dt'], p), g.stage_kkt(B, xu, w['xs'], w['ref'], fext, w['dt'], p)
    print(' kkt', {k: mism(kg[k], ko[k]) for k in ko})
    so, sg = o.stage_schur(B, ko, rho), g.stage_schur(B, ko, rho)
    print(' schur', {k: mism(sg[k], so[k]) for k in so})
    lam0 = np.zeros((B, d['vecp']), np.float32)
    for eps, cap in ((1e-4, 200), (-1.0, 20)):
        lo, io = o.stage_pcg(B, so['S'], so['Pinv'], so['gamma'], lam0, np.full(B, eps, np.float32), cap)
        lg, ig = g.stage_pcg(B, so['S'], so['Pinv'], so['gamma'], lam0, np.full(B, eps, np.float32), cap)
        print(' pcg', eps, 'iters oracle', io.tolist(), 'gato', ig.tolist(), 'lam', mism(lg, lo))
    dzo = o.stage_dz(B, lo, so['Qinv'], so['Rinv'], ko['q'], ko['r'], ko['A'], ko['Bm'])
    dzg = g.stage_dz(B, lo, so['Qinv'], so['Rinv'], ko['q'], ko['r'], ko['A'], ko['Bm'])
    print(' dz', mism(dzg[0], dzo[0]), 'q', mism(dzg[1], dzo[1]), 'r', mism(dzg[2], dzo[2]))
    for na in (1, 8):
        mo = o.stage_merit(B, xu, dzo[0], w['xs'], w['ref'], mu, fext, w['dt'], p, na)
        mg = g.stage_merit(B, xu, dzo[0], w['xs'], w['ref'], mu, fext, w['dt'], p, na)
        print(' merit', na, mism(mg, mo))
    mi = o.stage_merit(B, xu, np.zeros_like(dzo[0]), w['xs'], w['ref'], mu, fext, w['dt'], p, 1)[:, 0].copy(); mi[-1] = -1e30
    lo_ = o.stage_linesearch(B, xu, dzo[0], mo, mi, rho, np.ones(B, np.float32), 1)
    lg_ = g.stage_linesearch(B, xu, dzo[0], mo, mi, rho, np.ones(B, np.float32), 1)
    print(' linesearch', {k: mism(lg_[k], lo_[k]) for k in lo_})

def solves(plant, N, cfg, B):
    print(f"==== solve {plant} N={N} B={B}", flush=True)
    o, g = Backend('oracle', plant, N), GatoBackend(plant, N)
    w = make_config(cfg, B=B, N=N)
    for label, p in (('cfg', w['params']), ('default', dict(DEFAULT_SOLVER_PARAMS, dt=float(w['dt'])))):
        so, sg = o.solver(B, p), g.solver(B, p)
        for rep in range(2):
            xin = w['xu'] if rep == 0 else ro['XU']
            ro = so.solve(xin, w['xs'], w['ref'], w['dt'])
            rg = sg.solve(xin, w['xs'], w['ref'], w['dt'])
            print(f'  {label} solve#{rep}: XU', mism(rg['XU'], ro['XU']), 'pcg', np.array_equal(rg['pcg_iters'], ro['pcg_iters']), 'step', np.array_equal(rg['ls_step_size'], ro['ls_step_size']),
                  'lsmerit', mism(rg['ls_min_merit'], ro['ls_min_merit']), 'sqp', np.array_equal(rg['sqp_iters'], ro['sqp_iters']), 'conv', np.array_equal(rg['kkt_converged'], ro['kkt_converged']),
                  'final', mism(rg['final_merit'], ro['final_merit']), 'init', mism(rg['initial_merit'], ro['initial_merit']), 'n', rg['n_pcg'], rg['n_ls'], ro['n_pcg'], ro['n_ls'],
                  f"dev_ms={rg['device_time_ms']:.3f}")
        fe = np.random.default_rng(5).normal(0, 2, (B, 6)).astype(np.float32)
        so.set_batch('f_ext', fe); sg.set_batch('f_ext', fe)
        xk = w['xs'][0]; uk = np.random.default_rng(6).uniform(-5, 5, o.d['nq']).astype(np.float32)
        print('  sim_forward', mism(sg.sim_forward(xk, uk, w['dt']), so.sim_forward(xk, uk, w['dt'])))
        so.close(); sg.close()

def timing():
    print("==== timing", flush=True)
    for cfg, B in (('bench', 512), (3, 512), (1, 1)):
        w = make_config(cfg, B=B)
        g = GatoBackend(w['plant'], w['N'])
        for label, p in (('cfg', w['params']), ('default', dict(DEFAULT_SOLVER_PARAMS, dt=float(w['dt'])))):
            s = g.solver(B, p)
            ts = []
            for r in range(8):
                s.reset('dual'); s.reset('rho')
                out = s.solve(w['xu'], w['xs'], w['ref'], w['dt'])
                ts.append(out['device_time_ms'])
            print(f"  {w['plant']} N={w['N']} B={B} {label}: device ms {np.median(ts[2:]):.3f} (min {min(ts):.3f}) solves/s {B/np.median(ts[2:])*1e3:.0f} n_pcg {out['n_pcg']} pcg_mean {out['pcg_iters'].mean():.1f} launches {s.kernel_launches()}", flush=True)
            s.close()

if __name__ == '__main__':
    what = sys.argv[1:] or ['stages', 'solves', 'timing']
    if 'stages' in what:
        stages('iiwa14', 8, 1, 2); stages('iiwa14', 32, 2, 4); stages('indy7', 32, 3, 4)
    if 'solves' in what:
        solves('iiwa14', 8, 1, 1); solves('iiwa14', 32, 2, 16); solves('indy7', 32, 3, 16)
    if 'timing' in what:
        timing()
