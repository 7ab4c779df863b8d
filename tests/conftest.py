import json
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_built():
    from oracle.pyapi import ensure_oracle_built

    return ensure_oracle_built()


def load_golden(plant, N, mode):
    p = GOLDEN / f"golden_{plant}_N{N}_{mode}.npz"
    if not p.exists():
        pytest.skip(f"golden fixture {p.name} not present")
    return np.load(p)


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(1e-30, np.abs(b).max()))


def n_mismatch(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return int(((a != b) & ~(np.isnan(a) & np.isnan(b))).sum())


def params_of(G, key):
    return json.loads(str(G[key]))
