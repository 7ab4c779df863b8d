"""ctypes binding of the C ABI in include/gato_b200.h (gato_b200/lib/libgato_b200.so).

This is the only place the Python side touches native code.  There is no CPU fallback: loading fails loudly
when the shared library is missing, and every call fails with the library's error text when no CUDA device /
kernel image is available.
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np

PLANT_ID = {"indy7": 0, "iiwa14": 1}
NQ = {"indy7": 6, "iiwa14": 7}
PARAM_ORDER = ["dt", "max_sqp_iters", "kkt_tol", "max_pcg_iters", "pcg_tol", "solve_ratio", "mu", "q_cost", "qd_cost", "u_cost", "N_cost", "q_lim_cost", "vel_lim_cost", "ctrl_lim_cost", "rho"]
COST_ORDER = ["q_cost", "qd_cost", "u_cost", "N_cost", "q_lim_cost", "vel_lim_cost", "ctrl_lim_cost"]
BATCH_FIELD = {"f_ext": 0, "rho": 1, "drho": 2, "mu": 3, "pcg_tol": 4}

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class GatoParams(C.Structure):
    _fields_ = [("dt", C.c_float), ("max_sqp_iters", C.c_uint32), ("kkt_tol", C.c_float), ("max_pcg_iters", C.c_uint32), ("pcg_tol", C.c_float), ("solve_ratio", C.c_float), ("mu", C.c_float),
                ("q_cost", C.c_float), ("qd_cost", C.c_float), ("u_cost", C.c_float), ("N_cost", C.c_float), ("q_lim_cost", C.c_float), ("vel_lim_cost", C.c_float), ("ctrl_lim_cost", C.c_float),
                ("rho", C.c_float)]


class GatoStats(C.Structure):
    _fields_ = [("solve_time_us", C.c_double), ("device_time_ms", C.c_float), ("batch", C.c_int32), ("n_pcg", C.c_int32), ("n_ls", C.c_int32), ("sqp_iters", C.POINTER(C.c_int32)),
                ("kkt_converged", C.POINTER(C.c_int32)), ("pcg_iters", C.POINTER(C.c_int32)), ("ls_min_merit", C.POINTER(C.c_float)), ("ls_step_size", C.POINTER(C.c_float)),
                ("final_merit", C.POINTER(C.c_float)), ("initial_merit", C.POINTER(C.c_float))]


class GatoMpcOut(C.Structure):
    _fields_ = [("best_id", C.c_int32), ("best_error", C.c_double), ("errors", C.POINTER(C.c_double)), ("xu_best", C.POINTER(C.c_float))]


class GatoError(RuntimeError):
    pass


MODEL_MAX_NQ, MODEL_MAX_TRIG, PLANT_MODEL0 = 8, 16, 2


class GatoTrig(C.Structure):
    _fields_ = [("idx", C.c_int32), ("k", C.c_int32), ("coef", C.c_double)]


class GatoModel(C.Structure):
    """gato_model of include/gato_b200.h: a robot as data tables."""

    _fields_ = [("name", C.c_char * 32), ("nq", C.c_int32), ("style", C.c_int32), ("X", C.c_double * (36 * MODEL_MAX_NQ)), ("I", C.c_double * (36 * MODEL_MAX_NQ)),
                ("Xhom", C.c_double * (16 * MODEL_MAX_NQ)), ("dXhom", C.c_double * (16 * MODEL_MAX_NQ)), ("joint_limit", C.c_double * MODEL_MAX_NQ), ("vel_limit", C.c_double * MODEL_MAX_NQ),
                ("ctrl_limit", C.c_double * MODEL_MAX_NQ), ("n_x_trig", C.c_int32), ("n_xh_trig", C.c_int32), ("n_dxh_trig", C.c_int32), ("x_trig", GatoTrig * (MODEL_MAX_TRIG * MODEL_MAX_NQ)),
                ("xh_trig", GatoTrig * (8 * MODEL_MAX_NQ)), ("dxh_trig", GatoTrig * (8 * MODEL_MAX_NQ))]


_LIB = None


def lib_path():
    """The in-tree shared library; GATO_B200_LIB overrides it (used to A/B kernel build variants on the GPU box)."""
    override = os.environ.get("GATO_B200_LIB")
    return Path(override) if override else Path(__file__).resolve().parent / "lib" / "libgato_b200.so"


def load():
    """Load libgato_b200.so (raises if it has not been built: run `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not p.exists():
        raise GatoError(f"{p} not found — the CUDA extension is not built; gato_b200 has no CPU fallback")
    lib = C.CDLL(str(p))
    vp = C.c_void_p
    lib.gato_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, vp, C.POINTER(GatoParams)]
    lib.gato_destroy.argtypes = [vp]
    lib.gato_destroy.restype = None
    lib.gato_last_error.argtypes = [vp]
    lib.gato_last_error.restype = C.c_char_p
    lib.gato_set_batch.argtypes = [vp, C.c_int, f32p, C.c_int]
    lib.gato_reset.argtypes = [vp, C.c_int]
    lib.gato_set_rho_adaptation.argtypes = [vp, C.c_int]
    lib.gato_solve.argtypes = [vp, vp, vp, vp, C.c_float, C.POINTER(GatoStats)]
    lib.gato_solve_host.argtypes = [vp, f32p, f32p, f32p, C.c_float, C.POINTER(GatoStats)]
    lib.gato_solve_async.argtypes = [vp, vp, vp, vp, C.c_float]
    lib.gato_solve_wait.argtypes = [vp, C.POINTER(GatoStats)]
    lib.gato_sim_forward.argtypes = [vp, vp, vp, vp, C.c_float]
    lib.gato_sim_forward_host.argtypes = [vp, f32p, f32p, f32p, C.c_float]
    lib.gato_get_merits.argtypes = [vp, vp, vp]
    lib.gato_dims.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.gato_kernel_launches.argtypes = [vp]
    lib.gato_kernel_launches.restype = C.c_long
    lib.gato_mpc_set_warm_start.argtypes = [vp, f32p, C.c_int]
    lib.gato_mpc_set_state_offsets.argtypes = [vp, C.c_void_p]
    lib.gato_mpc_get_warm_start.argtypes = [vp, f32p]
    lib.gato_mpc_step.argtypes = [vp, f32p, f32p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int, C.POINTER(GatoMpcOut), C.POINTER(GatoStats)]
    lib.gato_set_kernel_timing.argtypes = [vp, C.c_int]
    lib.gato_get_kernel_times.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_int)]
    lib.gato_get_launch_times.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_int]
    lib.gato_get_device_pointers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    if hasattr(lib, "gato_mpc_local_async"):
        lib.gato_mpc_record_floats.argtypes = [vp]
        lib.gato_mpc_input_floats.argtypes = [vp]
        lib.gato_mpc_local_async.argtypes = [vp, vp, C.c_int, C.c_float, C.c_float, C.c_int, vp]
        lib.gato_mpc_adopt_async.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
        lib.gato_mpc_wait.argtypes = [vp, C.POINTER(GatoMpcOut), C.POINTER(GatoStats)]
    if hasattr(lib, "gato_ee_pos"):
        lib.gato_ee_pos.argtypes = [vp, f32p, C.c_int, f32p]
        lib.gato_set_kkt_residual_log.argtypes = [vp, C.c_int]
        lib.gato_get_kkt_residuals.argtypes = [vp, f32p, f32p]
    if hasattr(lib, "gato_measure_fp32_peak"):  # absent from older builds loaded through GATO_B200_LIB for A/B runs
        lib.gato_measure_fp32_peak.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
    lib.gato_model_builtin.argtypes = [C.c_int, C.POINTER(GatoModel)]
    lib.gato_model_load.argtypes = [C.c_char_p, C.POINTER(GatoModel)]
    lib.gato_model_save.argtypes = [C.POINTER(GatoModel), C.c_char_p]
    lib.gato_model_register.argtypes = [C.POINTER(GatoModel)]
    lead = [C.c_int, C.c_int, C.c_int]
    lib.gato_stage_kkt.argtypes = lead + [f32p] * 4 + [C.c_float, f32p] + [f32p] * 7
    lib.gato_stage_schur.argtypes = lead + [f32p] * 11
    lib.gato_stage_pcg.argtypes = lead + [f32p] * 5 + [C.c_int, i32p, i32p]
    lib.gato_stage_dz.argtypes = lead + [f32p] * 8
    lib.gato_stage_merit.argtypes = lead + [f32p] * 6 + [C.c_float, f32p, C.c_int, f32p]
    lib.gato_stage_linesearch.argtypes = lead + [f32p] * 7 + [C.c_int]
    _LIB = lib
    return lib


def measure_fp32_peak(device=0, packed=True):
    """Measured FP32 FMA peak of the device in TFLOP/s (scalar FFMA or packed FFMA2)."""
    out = C.c_double()
    rc = load().gato_measure_fp32_peak(int(device), int(bool(packed)), C.byref(out))
    if rc != 0:
        raise GatoError(f"gato_measure_fp32_peak failed ({rc})")
    return float(out.value)


class Model:
    """A robot model as data (gato_model): the compiled robots' tables, a file, or tables edited in Python; `register` makes it a plant name
    that Solver / GatoBackend / gato_b200.bsqp accept like "iiwa14"."""

    def __init__(self, raw=None):
        self.raw = raw if raw is not None else GatoModel()

    @classmethod
    def builtin(cls, plant):
        m = cls()
        if load().gato_model_builtin(PLANT_ID[plant], C.byref(m.raw)) != 0:
            raise GatoError(f"no compiled robot '{plant}'")
        return m

    @classmethod
    def load(cls, path):
        m = cls()
        if load().gato_model_load(str(path).encode(), C.byref(m.raw)) != 0:
            raise GatoError(f"gato_model_load failed: {load().gato_last_error(None).decode()}")
        return m

    def save(self, path):
        if load().gato_model_save(C.byref(self.raw), str(path).encode()) != 0:
            raise GatoError(f"gato_model_save failed: {load().gato_last_error(None).decode()}")

    @property
    def nq(self):
        return int(self.raw.nq)

    @property
    def name(self):
        return self.raw.name.decode()

    def array(self, field):
        """numpy view (shared memory) of one of the double tables: X, I [nq, 36], Xhom, dXhom [nq, 16], joint_limit, vel_limit, ctrl_limit [nq]"""
        a = np.ctypeslib.as_array(getattr(self.raw, field))
        per = {"X": 36, "I": 36, "Xhom": 16, "dXhom": 16}.get(field)
        return a[: per * self.nq].reshape(self.nq, per) if per else a[: self.nq]

    def trig(self, which):
        """(idx, coef, k) arrays of the x / xh / dxh trig entries"""
        n = int(getattr(self.raw, f"n_{which}_trig"))
        t = getattr(self.raw, f"{which}_trig")
        return (np.array([t[i].idx for i in range(n)], np.int32), np.array([t[i].coef for i in range(n)], np.float64), np.array([t[i].k for i in range(n)], np.int32))

    def register(self, name=None):
        """-> plant name usable wherever "iiwa14" / "indy7" is (the library assigns the id; the same tables always get the same id)"""
        pid = load().gato_model_register(C.byref(self.raw))
        if pid < PLANT_MODEL0:
            raise GatoError(f"gato_model_register failed ({pid}): {load().gato_last_error(None).decode()}")
        name = name or self.name or f"model{pid}"
        if name in PLANT_ID and PLANT_ID[name] != pid:
            raise GatoError(f"plant name '{name}' is already taken")
        PLANT_ID[name], NQ[name] = pid, self.nq
        return name


def make_params(p):
    g = GatoParams()
    for k in PARAM_ORDER:
        setattr(g, k, int(p[k]) if k in ("max_sqp_iters", "max_pcg_iters") else float(p[k]))
    return g


def dims(plant, N):
    nq = NQ[plant]
    nx, nu = 2 * nq, nq
    return dict(nq=nq, nx=nx, nu=nu, N=N, traj=(nx + nu) * N - nu, vecp=(N + 2) * nx, brow=3 * nx * nx)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Solver:
    """One native solver object (== one reference BSQP<float,B> instance) on one device."""

    def __init__(self, plant, N, B, params, device=0, stream=None):
        self.lib = load()
        self.plant, self.N, self.B, self.params = plant, N, B, dict(params)
        self.d = dims(plant, N)
        self.h = C.c_void_p()
        gp = make_params(params)
        rc = self.lib.gato_create(C.byref(self.h), PLANT_ID[plant], N, B, device, stream, C.byref(gp))
        if rc != 0:
            raise GatoError(f"gato_create failed ({rc}): {self.lib.gato_last_error(None).decode()}")

    def _check(self, rc, what):
        if rc != 0:
            raise GatoError(f"{what} failed ({rc}): {self.lib.gato_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.gato_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_batch(self, which, arr, set_default=True):
        self._check(self.lib.gato_set_batch(self.h, BATCH_FIELD[which], _f(arr).reshape(-1), int(set_default)), "gato_set_batch")

    def reset(self, which):
        self._check(self.lib.gato_reset(self.h, {"dual": 0, "rho": 1}[which]), "gato_reset")

    def set_rho_adaptation(self, on):
        self._check(self.lib.gato_set_rho_adaptation(self.h, int(on)), "gato_set_rho_adaptation")

    def _stats(self, st):
        B = self.B
        arr = lambda p, n, dt: np.ctypeslib.as_array(p, shape=(n,)).astype(dt, copy=True) if n > 0 else np.zeros(0, dt)  # noqa: E731
        return dict(
            sqp_time_us=st.solve_time_us, device_time_ms=st.device_time_ms, n_pcg=st.n_pcg, n_ls=st.n_ls,
            sqp_iters=arr(st.sqp_iters, B, np.int32), kkt_converged=arr(st.kkt_converged, B, np.int32),
            pcg_iters=arr(st.pcg_iters, st.n_pcg * B, np.int32).reshape(st.n_pcg, B),
            ls_min_merit=arr(st.ls_min_merit, st.n_ls * B, np.float32).reshape(st.n_ls, B),
            ls_step_size=arr(st.ls_step_size, st.n_ls * B, np.float32).reshape(st.n_ls, B),
            final_merit=arr(st.final_merit, B, np.float32), initial_merit=arr(st.initial_merit, B, np.float32))

    def solve(self, xu, xs, ref, dt):
        """Host-buffer solve (H2D, solve, D2H) — the PyBSQP::solve path (python/bindings.cu:68-148)."""
        xu = _f(xu).reshape(self.B, self.d["traj"]).copy()
        st = GatoStats()
        self._check(self.lib.gato_solve_host(self.h, xu.reshape(-1), _f(xs).reshape(-1), _f(ref).reshape(-1), float(dt), C.byref(st)), "gato_solve_host")
        out = self._stats(st)
        out["XU"] = xu
        return out

    def solve_device(self, d_xu, d_xs, d_ref, dt):
        """Solve on device pointers (ints), in place — the BSQP::solve path (gato/bsqp/bsqp.cuh:103)."""
        st = GatoStats()
        self._check(self.lib.gato_solve(self.h, d_xu, d_xs, d_ref, float(dt), C.byref(st)), "gato_solve")
        return self._stats(st)

    def solve_async(self, d_xu, d_xs, d_ref, dt):
        self._check(self.lib.gato_solve_async(self.h, d_xu, d_xs, d_ref, float(dt)), "gato_solve_async")

    def solve_wait(self):
        st = GatoStats()
        self._check(self.lib.gato_solve_wait(self.h, C.byref(st)), "gato_solve_wait")
        return self._stats(st)

    def device_pointers(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self.lib.gato_get_device_pointers(self.h, C.byref(a), C.byref(b), C.byref(c)), "gato_get_device_pointers")
        return a.value, b.value, c.value

    def ee_pos(self, q):
        """End-effector xyz of joint configurations q[n][nq] (or [nq]) by the solver's forward kinematics."""
        q = _f(q).reshape(-1, self.d["nq"])
        out = np.zeros((q.shape[0], 3), np.float32)
        self._check(self.lib.gato_ee_pos(self.h, q.reshape(-1), q.shape[0], out.reshape(-1)), "gato_ee_pos")
        return out

    def set_kkt_residual_log(self, enable=True):
        self._check(self.lib.gato_set_kkt_residual_log(self.h, int(bool(enable))), "gato_set_kkt_residual_log")

    def kkt_residuals(self):
        """(q_max[n_pcg][B], c_max[n_pcg][B]) of the last completed solve (needs set_kkt_residual_log(True) before it)."""
        cap = max(int(self.params["max_sqp_iters"]), 1) * self.B
        qm, cm = np.zeros(cap, np.float32), np.zeros(cap, np.float32)
        n = self.lib.gato_get_kkt_residuals(self.h, qm, cm)
        if n < 0:
            self._check(n, "gato_get_kkt_residuals")
        return qm[: n * self.B].reshape(n, self.B), cm[: n * self.B].reshape(n, self.B)

    def sim_forward(self, xk, uk, dt):
        out = np.zeros((self.B, self.d["nx"]), np.float32)
        self._check(self.lib.gato_sim_forward_host(self.h, out.reshape(-1), _f(xk), _f(uk), float(dt)), "gato_sim_forward_host")
        return out

    # ---- closed-loop MPC step on the device (include/gato_b200.h: gato_mpc_*) ----
    def mpc_set_warm_start(self, xu):
        xu = _f(xu)
        per_solve = xu.size == self.B * self.d["traj"] and self.B > 1
        assert per_solve or xu.size == self.d["traj"]
        self._check(self.lib.gato_mpc_set_warm_start(self.h, xu.reshape(-1), int(per_solve)), "gato_mpc_set_warm_start")

    def mpc_set_state_offsets(self, off):
        if off is None:
            self._check(self.lib.gato_mpc_set_state_offsets(self.h, None), "gato_mpc_set_state_offsets")
        else:
            off = _f(off).reshape(self.B, self.d["nx"])
            self._check(self.lib.gato_mpc_set_state_offsets(self.h, off.ctypes.data_as(C.c_void_p)), "gato_mpc_set_state_offsets")

    def mpc_get_warm_start(self):
        out = np.zeros((self.B, self.d["traj"]), np.float32)
        self._check(self.lib.gato_mpc_get_warm_start(self.h, out.reshape(-1)), "gato_mpc_get_warm_start")
        return out

    def mpc_step(self, x_curr, ref_window, x_last=None, u_last=None, sim_dt=0.0, dt=0.01, reset_rho=True):
        """One control step: broadcast (x_curr, ref_window), solve, score the hypotheses with sim_forward, adopt the best trajectory.
        Returns the statistics dictionary plus best_id, errors[B] (float64) and XU_best[traj]."""
        x_curr, ref_window = _f(x_curr), _f(ref_window)
        xl = ul = None
        if x_last is not None:
            xl_a, ul_a = _f(x_last), _f(u_last)
            xl, ul = xl_a.ctypes.data_as(C.c_void_p), ul_a.ctypes.data_as(C.c_void_p)
        out, st = GatoMpcOut(), GatoStats()
        self._check(self.lib.gato_mpc_step(self.h, x_curr, ref_window, xl, ul, float(sim_dt), float(dt), 1 if reset_rho else 0, C.byref(out), C.byref(st)), "gato_mpc_step")
        res = self._stats(st)
        res["best_id"] = int(out.best_id)
        res["errors"] = np.ctypeslib.as_array(out.errors, shape=(self.B,)).copy()
        res["XU_best"] = np.ctypeslib.as_array(out.xu_best, shape=(self.d["traj"],)).copy()
        return res

    # ---- the same step in its multi-GPU phases (device pointers as ints; everything is enqueued on the solver's stream) ----
    def mpc_record_floats(self):
        return int(self.lib.gato_mpc_record_floats(self.h))

    def mpc_input_floats(self):
        return int(self.lib.gato_mpc_input_floats(self.h))

    def mpc_local_async(self, d_in, score, sim_dt, dt, reset_rho, d_record):
        self._check(self.lib.gato_mpc_local_async(self.h, d_in, int(bool(score)), float(sim_dt), float(dt), 1 if reset_rho else 0, d_record), "gato_mpc_local_async")

    def mpc_adopt_async(self, d_records, n_records, record_stride, id_stride):
        self._check(self.lib.gato_mpc_adopt_async(self.h, d_records, int(n_records), int(record_stride), int(id_stride)), "gato_mpc_adopt_async")

    def mpc_wait(self):
        out, st = GatoMpcOut(), GatoStats()
        self._check(self.lib.gato_mpc_wait(self.h, C.byref(out), C.byref(st)), "gato_mpc_wait")
        res = self._stats(st)
        res["best_id"] = int(out.best_id)
        res["best_error"] = float(out.best_error)
        res["errors"] = np.ctypeslib.as_array(out.errors, shape=(self.B,)).copy()
        res["XU_best"] = np.ctypeslib.as_array(out.xu_best, shape=(self.d["traj"],)).copy()
        return res

    def kernel_launches(self):
        return int(self.lib.gato_kernel_launches(self.h))

    KERNEL_CLASSES = ("k_kkt", "k_schur", "k_pcg", "k_merit_ls<8>", "k_merit_ls<1>")

    def set_kernel_timing(self, enable=True):
        self._check(self.lib.gato_set_kernel_timing(self.h, int(bool(enable))), "set_kernel_timing")

    def launch_times(self, cap=256):
        """[(kernel, ms), ...] of the last completed solve, in launch order."""
        cls = (C.c_int * cap)()
        ms = (C.c_float * cap)()
        n = self.lib.gato_get_launch_times(self.h, cls, ms, cap)
        if n < 0:
            self._check(n, "get_launch_times")
        return [(self.KERNEL_CLASSES[cls[i]], float(ms[i])) for i in range(n)]

    def kernel_times(self):
        """{kernel: (total ms, launches)} of the last completed solve (needs set_kernel_timing(True) before it)."""
        ms = (C.c_float * 5)()
        n = (C.c_int * 5)()
        self._check(self.lib.gato_get_kernel_times(self.h, ms, n), "get_kernel_times")
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(self.KERNEL_CLASSES)}


class GatoBackend:
    """Same Python surface as oracle.pyapi.Backend, served by the CUDA library (used by the parity tests)."""

    kind = "gato"

    def __init__(self, plant, N):
        self.lib = load()
        self.plant, self.N = plant, N
        self.d = dims(plant, N)
        self.pid = PLANT_ID[plant]

    def solver(self, B, p):
        return Solver(self.plant, self.N, B, p)

    @staticmethod
    def _c7(p):
        return np.array([p[k] for k in COST_ORDER], dtype=np.float32)

    def _ok(self, rc, what):
        if rc != 0:
            raise GatoError(f"{what} failed ({rc})")

    def stage_kkt(self, B, xu, xs, ref, fext, dt, p):
        d, N = self.d, self.N
        nx, nu = d["nx"], d["nu"]
        o = dict(Q=np.zeros((B, N, nx * nx), np.float32), R=np.zeros((B, N, nu * nu), np.float32), q=np.zeros((B, N, nx), np.float32), r=np.zeros((B, N, nu), np.float32),
                 A=np.zeros((B, N, nx * nx), np.float32), Bm=np.zeros((B, N, nx * nu), np.float32), c=np.zeros((B, N, nx), np.float32))
        self._ok(self.lib.gato_stage_kkt(self.pid, N, B, _f(xu).reshape(-1), _f(xs).reshape(-1), _f(ref).reshape(-1), _f(fext).reshape(-1), float(dt), self._c7(p),
                                         *[o[k].reshape(-1) for k in ("Q", "R", "q", "r", "A", "Bm", "c")]), "stage_kkt")
        return o

    def stage_schur(self, B, kkt, rho):
        N, nx = self.N, self.d["nx"]
        Q, R = _f(kkt["Q"]).copy(), _f(kkt["R"]).copy()
        o = dict(S=np.zeros((B, N, 3 * nx * nx), np.float32), Pinv=np.zeros((B, N, 3 * nx * nx), np.float32), gamma=np.zeros((B, (N + 2) * nx), np.float32))
        self._ok(self.lib.gato_stage_schur(self.pid, N, B, Q.reshape(-1), R.reshape(-1), _f(kkt["q"]).reshape(-1), _f(kkt["r"]).reshape(-1), _f(kkt["A"]).reshape(-1), _f(kkt["Bm"]).reshape(-1),
                                           _f(kkt["c"]).reshape(-1), _f(rho).reshape(-1), o["S"].reshape(-1), o["Pinv"].reshape(-1), o["gamma"].reshape(-1)), "stage_schur")
        o["Qinv"], o["Rinv"] = Q, R
        return o

    def stage_pcg(self, B, S, Pinv, gamma, lam0, eps, max_iters, kkt_conv=None):
        lam = _f(lam0).copy()
        iters = np.zeros(B, np.int32)
        conv = np.zeros(B, np.int32) if kkt_conv is None else np.ascontiguousarray(kkt_conv, np.int32)
        self._ok(self.lib.gato_stage_pcg(self.pid, self.N, B, _f(S).reshape(-1), _f(Pinv).reshape(-1), _f(gamma).reshape(-1), lam.reshape(-1), _f(eps).reshape(-1), int(max_iters), conv, iters),
                 "stage_pcg")
        return lam, iters

    def stage_dz(self, B, lam, Qinv, Rinv, q, r, A, Bm):
        q, r = _f(q).copy(), _f(r).copy()
        dz = np.zeros((B, self.d["traj"]), np.float32)
        self._ok(self.lib.gato_stage_dz(self.pid, self.N, B, _f(lam).reshape(-1), _f(Qinv).reshape(-1), _f(Rinv).reshape(-1), q.reshape(-1), r.reshape(-1), _f(A).reshape(-1), _f(Bm).reshape(-1),
                                        dz.reshape(-1)), "stage_dz")
        return dz, q, r

    def stage_merit(self, B, xu, dz, xs, ref, mu, fext, dt, p, num_alphas=8):
        m = np.zeros((B, num_alphas), np.float32)
        self._ok(self.lib.gato_stage_merit(self.pid, self.N, B, _f(xu).reshape(-1), _f(dz).reshape(-1), _f(xs).reshape(-1), _f(ref).reshape(-1), _f(mu).reshape(-1), _f(fext).reshape(-1),
                                           float(dt), self._c7(p), int(num_alphas), m.reshape(-1)), "stage_merit")
        return m

    def stage_linesearch(self, B, xu, dz, merit8, merit_init, rho, drho, adapt=1):
        xu, mi, rho, drho = _f(xu).copy(), _f(merit_init).copy(), _f(rho).copy(), _f(drho).copy()
        step = np.zeros(B, np.float32)
        self._ok(self.lib.gato_stage_linesearch(self.pid, self.N, B, xu.reshape(-1), _f(dz).reshape(-1), _f(merit8).reshape(-1), mi, step, rho, drho, int(adapt)), "stage_linesearch")
        return dict(xu=xu, merit_init=mi, step=step, rho=rho, drho=drho)
