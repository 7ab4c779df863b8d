"""TEST INFRASTRUCTURE — ctypes drivers for the CPU oracle (oracle/libbsqp_oracle.so) and for the
reference harness (oracle/_ref/libgref_<plant>_N<N>_<mode>.so, built from the unmodified reference).

Both expose the same Python surface (`Backend`), so golden generation and parity tests use one driver.
Only tests/, bench.py (cpu_baseline / reference legs), __graft_entry__.smoke() and oracle/gen_golden.py
import this module; the product package gato_b200/ never does.
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
PLANT_ID = {"indy7": 0, "iiwa14": 1}
NQ = {"indy7": 6, "iiwa14": 7}

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")

PARAM_ORDER = ["dt", "max_sqp_iters", "kkt_tol", "max_pcg_iters", "pcg_tol", "solve_ratio", "mu", "q_cost", "qd_cost", "u_cost", "N_cost", "q_lim_cost", "vel_lim_cost", "ctrl_lim_cost", "rho"]
COST_ORDER = ["q_cost", "qd_cost", "u_cost", "N_cost", "q_lim_cost", "vel_lim_cost", "ctrl_lim_cost"]


def register_model(name, model):
    """Make a robot given as data tables known to the ORACLE under plant name `name`.  `model` is anything with the surface of
    gato_b200.native.Model: nq, raw.style, array(field), trig(which) -- plain data, no product code runs here."""
    if name in PLANT_ID:
        return name
    lib = C.CDLL(str(HERE / "libbsqp_oracle.so"))
    f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    lib.gato_oracle_register_model.argtypes = [C.c_int, C.c_int, f64p, f64p] + [C.c_int, i32p, f64p, i32p] * 3
    nq = int(model.nq)
    table = np.concatenate([np.asarray(model.array(k), np.float64).reshape(-1) for k in ("X", "I", "Xhom", "dXhom")])
    limits = np.concatenate([np.asarray(model.array(k), np.float64).reshape(-1) for k in ("joint_limit", "vel_limit", "ctrl_limit")])
    args = []
    for which in ("x", "xh", "dxh"):
        idx, coef, k = model.trig(which)
        args += [len(idx), np.ascontiguousarray(idx, np.int32), np.ascontiguousarray(coef, np.float64), np.ascontiguousarray(k, np.int32)]
    pid = lib.gato_oracle_register_model(nq, int(model.raw.style), np.ascontiguousarray(table), np.ascontiguousarray(limits), *args)
    assert pid >= 2, pid
    PLANT_ID[name], NQ[name] = pid, nq
    return name


def params15(p):
    return np.array([p[k] for k in PARAM_ORDER], dtype=np.float32)


def cost7(p):
    return np.array([p[k] for k in COST_ORDER], dtype=np.float32)


def dims(plant, N):
    nq = NQ[plant]
    nx, nu = 2 * nq, nq
    return dict(nq=nq, nx=nx, nu=nu, N=N, traj=(nx + nu) * N - nu, vecp=(N + 2) * nx, brow=3 * nx * nx)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class _Solver:
    def __init__(self, be, B, p):
        self.be, self.B, self.p = be, B, dict(p)
        self.h = be._create(B, params15(p))
        if not self.h:
            raise RuntimeError("solver creation failed")

    def close(self):
        if self.h:
            self.be._destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_batch(self, which, arr, set_default=True):
        idx = {"f_ext": 0, "rho": 1, "drho": 2, "mu": 3, "pcg_tol": 4}[which]
        rc = self.be._set_batch(self.h, idx, _f(arr).reshape(-1), int(set_default))
        assert rc == 0

    def reset(self, which):
        assert self.be._reset(self.h, {"dual": 0, "rho": 1}[which]) == 0

    def set_rho_adaptation(self, on):
        self.be._set_adapt(self.h, int(on))

    def solve(self, xu, xs, ref, dt):
        B, d = self.B, self.be.d
        xu = _f(xu).reshape(B, d["traj"]).copy()
        cap = int(self.p["max_sqp_iters"]) + 1
        out = dict(
            sqp_iters=np.zeros(B, np.int32),
            kkt_converged=np.zeros(B, np.int32),
            pcg_iters=np.zeros((cap, B), np.int32),
            ls_min_merit=np.zeros((cap, B), np.float32),
            ls_step_size=np.zeros((cap, B), np.float32),
            final_merit=np.zeros(B, np.float32),
            initial_merit=np.zeros(B, np.float32),
        )
        n_pcg, n_ls, t_us, ev_ms = C.c_int(0), C.c_int(0), C.c_double(0), C.c_float(0)
        rc = self.be._solve(self.h, xu, _f(xs).reshape(-1), _f(ref).reshape(-1), np.float32(dt), out["sqp_iters"], out["kkt_converged"], C.byref(n_pcg), C.byref(n_ls),
                            out["pcg_iters"].reshape(-1), out["ls_min_merit"].reshape(-1), out["ls_step_size"].reshape(-1), cap, out["final_merit"], out["initial_merit"],
                            C.byref(t_us), C.byref(ev_ms))
        assert rc == 0, rc
        out["XU"] = xu
        out["n_pcg"], out["n_ls"] = n_pcg.value, n_ls.value
        out["pcg_iters"] = out["pcg_iters"][: n_pcg.value]
        out["ls_min_merit"] = out["ls_min_merit"][: n_ls.value]
        out["ls_step_size"] = out["ls_step_size"][: n_ls.value]
        out["sqp_time_us"], out["event_ms"] = t_us.value, ev_ms.value
        return out

    def sim_forward(self, xk, uk, dt):
        out = np.zeros((self.B, self.be.d["nx"]), np.float32)
        assert self.be._sim(self.h, _f(xk), _f(uk), np.float32(dt), out.reshape(-1)) == 0
        return out

    def solve_timed(self, xu, xs, ref, dt, reps, reset_between=True):
        ev = np.zeros(reps, np.float32)
        us = np.zeros(reps, np.float64)
        rc = self.be._solve_timed(self.h, _f(xu).reshape(-1), _f(xs).reshape(-1), _f(ref).reshape(-1), np.float32(dt), reps, int(reset_between), ev, us)
        assert rc == 0
        return ev, us


class Backend:
    """Common surface over the oracle and the reference harness for one (plant, N)."""

    def __init__(self, kind, plant, N, mode="fast", path=None):
        self.kind, self.plant, self.N, self.mode = kind, plant, N, mode
        self.d = dims(plant, N)
        if kind == "oracle":
            path = path or HERE / "libbsqp_oracle.so"
            pre = "gato_oracle_"
        elif kind == "ref":
            path = path or HERE / "_ref" / f"libgref_{plant}_N{N}_{mode}.so"
            pre = "gref_"
        else:
            raise ValueError(kind)
        if not Path(path).exists():
            raise FileNotFoundError(path)
        self.lib = lib = C.CDLL(str(path))
        self.pre = pre
        pid = PLANT_ID[plant]
        is_o = kind == "oracle"
        lead = [C.c_int, C.c_int, C.c_int] if is_o else [C.c_int]
        self._lead = (lambda B: (pid, N, B)) if is_o else (lambda B: (B,))

        def fn(name, argtypes, restype=C.c_int):
            f = getattr(lib, pre + name)
            f.argtypes, f.restype = argtypes, restype
            return f

        if is_o:
            cr = fn("create", [C.c_int, C.c_int, C.c_int, f32p], C.c_void_p)
            self._create = lambda B, p: cr(pid, N, B, p)
            sv = fn("solve", [C.c_void_p, f32p, f32p, f32p, C.c_float, i32p, i32p, C.POINTER(C.c_int), C.POINTER(C.c_int), i32p, f32p, f32p, C.c_int, f32p, f32p, C.POINTER(C.c_double)])
            self._solve = lambda h, *a: sv(h, *a[:-1])  # oracle has no CUDA-event time
            self._solve_timed = None
        else:
            info = np.zeros(6, np.int32)
            fn("info", [i32p])(info)
            assert info[0] == N and info[1] == self.d["nx"] and info[3] == pid, (info, plant, N)
            cr = fn("create", [C.c_int, f32p], C.c_void_p)
            self._create = lambda B, p: cr(B, p)
            self._solve = fn("solve", [C.c_void_p, f32p, f32p, f32p, C.c_float, i32p, i32p, C.POINTER(C.c_int), C.POINTER(C.c_int), i32p, f32p, f32p, C.c_int, f32p, f32p,
                                       C.POINTER(C.c_double), C.POINTER(C.c_float)])
            self._solve_timed = fn("solve_timed", [C.c_void_p, f32p, f32p, f32p, C.c_float, C.c_int, C.c_int, f32p, np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")])
            nb = fn("batches", [i32p, C.c_int])
            buf = np.zeros(32, np.int32)
            self.batches = list(buf[: nb(buf, 32)])
        self._destroy = fn("destroy", [C.c_void_p], None)
        self._set_batch = fn("set_batch", [C.c_void_p, C.c_int, f32p, C.c_int])
        self._reset = fn("reset", [C.c_void_p, C.c_int])
        self._set_adapt = fn("set_rho_adaptation", [C.c_void_p, C.c_int], None)
        self._sim = fn("sim_forward", [C.c_void_p, f32p, f32p, C.c_float, f32p])
        self._kkt = fn("stage_kkt", lead + [f32p] * 4 + [C.c_float, f32p] + [f32p] * 7)
        self._schur = fn("stage_schur", lead + [f32p] * 11)
        self._pcg = fn("stage_pcg", lead + [f32p] * 5 + [C.c_int, i32p, i32p])
        self._dz = fn("stage_dz", lead + [f32p] * 8)
        self._merit = fn("stage_merit", lead + [f32p] * 6 + [C.c_float, f32p, C.c_int, f32p])
        self._ls = fn("stage_linesearch", lead + [f32p] * 7 + [C.c_int])
        if is_o:
            dd = fn("dyn_dump", [C.c_int, C.c_int] + [f32p] * 7)
            self._dyn = lambda n, *a: dd(pid, n, *a)
            lib.gato_oracle_set_threads.argtypes = [C.c_int]
        else:
            self._dyn = fn("dyn_dump", [C.c_int] + [f32p] * 7)

    def set_threads(self, n):
        if self.kind == "oracle":
            self.lib.gato_oracle_set_threads(int(n))

    def solver(self, B, p):
        return _Solver(self, B, p)

    # ---- stages: all arrays are [B, ...] float32, returned as new arrays ----
    def stage_kkt(self, B, xu, xs, ref, fext, dt, p):
        d, N = self.d, self.N
        nx, nu = d["nx"], d["nu"]
        o = dict(Q=np.zeros((B, N, nx * nx), np.float32), R=np.zeros((B, N, nu * nu), np.float32), q=np.zeros((B, N, nx), np.float32), r=np.zeros((B, N, nu), np.float32),
                 A=np.zeros((B, N, nx * nx), np.float32), Bm=np.zeros((B, N, nx * nu), np.float32), c=np.zeros((B, N, nx), np.float32))
        rc = self._kkt(*self._lead(B), _f(xu).reshape(-1), _f(xs).reshape(-1), _f(ref).reshape(-1), _f(fext).reshape(-1), np.float32(dt), cost7(p),
                       *[o[k].reshape(-1) for k in ("Q", "R", "q", "r", "A", "Bm", "c")])
        assert rc == 0, rc
        return o

    def stage_schur(self, B, kkt, rho):
        d, N = self.d, self.N
        nx = d["nx"]
        Q, R = _f(kkt["Q"]).copy(), _f(kkt["R"]).copy()
        o = dict(S=np.zeros((B, N, 3 * nx * nx), np.float32), Pinv=np.zeros((B, N, 3 * nx * nx), np.float32), gamma=np.zeros((B, (N + 2) * nx), np.float32))
        rc = self._schur(*self._lead(B), Q.reshape(-1), R.reshape(-1), _f(kkt["q"]).reshape(-1), _f(kkt["r"]).reshape(-1), _f(kkt["A"]).reshape(-1), _f(kkt["Bm"]).reshape(-1),
                         _f(kkt["c"]).reshape(-1), _f(rho).reshape(-1), o["S"].reshape(-1), o["Pinv"].reshape(-1), o["gamma"].reshape(-1))
        assert rc == 0, rc
        o["Qinv"], o["Rinv"] = Q, R
        return o

    def stage_pcg(self, B, S, Pinv, gamma, lam0, eps, max_iters, kkt_conv=None):
        lam = _f(lam0).copy()
        iters = np.zeros(B, np.int32)
        conv = np.zeros(B, np.int32) if kkt_conv is None else np.ascontiguousarray(kkt_conv, np.int32)
        rc = self._pcg(*self._lead(B), _f(S).reshape(-1), _f(Pinv).reshape(-1), _f(gamma).reshape(-1), lam.reshape(-1), _f(eps).reshape(-1), int(max_iters), conv, iters)
        assert rc == 0, rc
        return lam, iters

    def stage_dz(self, B, lam, Qinv, Rinv, q, r, A, Bm):
        q, r = _f(q).copy(), _f(r).copy()
        dz = np.zeros((B, self.d["traj"]), np.float32)
        rc = self._dz(*self._lead(B), _f(lam).reshape(-1), _f(Qinv).reshape(-1), _f(Rinv).reshape(-1), q.reshape(-1), r.reshape(-1), _f(A).reshape(-1), _f(Bm).reshape(-1), dz.reshape(-1))
        assert rc == 0, rc
        return dz, q, r

    def stage_merit(self, B, xu, dz, xs, ref, mu, fext, dt, p, num_alphas=8):
        m = np.zeros((B, num_alphas), np.float32)
        rc = self._merit(*self._lead(B), _f(xu).reshape(-1), _f(dz).reshape(-1), _f(xs).reshape(-1), _f(ref).reshape(-1), _f(mu).reshape(-1), _f(fext).reshape(-1), np.float32(dt), cost7(p),
                         int(num_alphas), m.reshape(-1))
        assert rc == 0, rc
        return m

    def stage_linesearch(self, B, xu, dz, merit8, merit_init, rho, drho, adapt=1):
        xu, mi, rho, drho = _f(xu).copy(), _f(merit_init).copy(), _f(rho).copy(), _f(drho).copy()
        step = np.zeros(B, np.float32)
        rc = self._ls(*self._lead(B), xu.reshape(-1), _f(dz).reshape(-1), _f(merit8).reshape(-1), mi, step, rho, drho, int(adapt))
        assert rc == 0, rc
        return dict(xu=xu, merit_init=mi, step=step, rho=rho, drho=drho)

    def dyn_dump(self, x, u, fext):
        n, nq = x.shape[0], self.d["nq"]
        qdd, dqdd = np.zeros((n, nq), np.float32), np.zeros((n, 3 * nq * nq), np.float32)
        ee, dee = np.zeros((n, 6), np.float32), np.zeros((n, 6 * nq), np.float32)
        rc = self._dyn(n, _f(x).reshape(-1), _f(u).reshape(-1), _f(fext).reshape(-1), qdd.reshape(-1), dqdd.reshape(-1), ee.reshape(-1), dee.reshape(-1))
        assert rc == 0, rc
        return dict(qdd=qdd, dqdd=dqdd, ee=ee, dee=dee)


def ensure_oracle_built():
    so = HERE / "libbsqp_oracle.so"
    src = HERE / "bsqp_oracle.cpp"
    if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        rc = os.system(f"make -C {HERE} >/dev/null")
        if rc != 0:
            raise RuntimeError("oracle build failed")
    return so
