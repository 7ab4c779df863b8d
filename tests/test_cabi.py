"""The C-ABI shared library loads and exports every symbol include/gato_b200.h declares; argument errors and the absence of a
GPU are reported through return codes and gato_last_error (never a crash, never a silent CPU fallback)."""
import ctypes as C
import re

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from gato_b200 import native

    if not native.lib_path().exists():
        import __graft_entry__ as g

        g.build()
    return native.load()


def test_exports_every_declared_symbol(lib):
    hdr = (ROOT / "include" / "gato_b200.h").read_text()
    names = sorted(set(re.findall(r"^\s*(?:int|void|long|const char\*)\s+(gato_[a-z0-9_]+)\s*\(", hdr, flags=re.M)))
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in gato_b200.h but not exported"


def test_dims(lib):
    nx, nu, tr = C.c_int(), C.c_int(), C.c_int()
    assert lib.gato_dims(1, 32, C.byref(nx), C.byref(nu), C.byref(tr)) == 0
    assert (nx.value, nu.value, tr.value) == (14, 7, 665)
    assert lib.gato_dims(0, 16, C.byref(nx), C.byref(nu), C.byref(tr)) == 0
    assert (nx.value, nu.value, tr.value) == (12, 6, 282)
    assert lib.gato_dims(5, 16, None, None, None) != 0


def test_create_rejects_bad_arguments_and_reports_missing_gpu(lib):
    from gato_b200 import native

    gp = native.make_params(dict(dt=0.01, max_sqp_iters=1, kkt_tol=0, max_pcg_iters=10, pcg_tol=1e-4, solve_ratio=1, mu=10, q_cost=1, qd_cost=0.01, u_cost=1e-6, N_cost=50,
                                 q_lim_cost=0, vel_lim_cost=0, ctrl_lim_cost=0, rho=0.01))
    h = C.c_void_p()
    assert lib.gato_create(C.byref(h), 7, 32, 4, 0, None, C.byref(gp)) == -1  # GATO_ERR_ARG
    assert b"invalid" in lib.gato_last_error(None)
    import torch

    if not torch.cuda.is_available():
        rc = lib.gato_create(C.byref(h), 1, 32, 4, 0, None, C.byref(gp))
        assert rc == -2 and not h.value  # GATO_ERR_CUDA: no CPU fallback
        assert b"no CUDA device" in lib.gato_last_error(None)
        with pytest.raises(native.GatoError):
            native.Solver("iiwa14", 32, 4, dict(dt=0.01, max_sqp_iters=1, kkt_tol=0, max_pcg_iters=10, pcg_tol=1e-4, solve_ratio=1, mu=10, q_cost=1, qd_cost=0.01, u_cost=1e-6,
                                                N_cost=50, q_lim_cost=0, vel_lim_cost=0, ctrl_lim_cost=0, rho=0.01))
