import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from gato_b200 import native
from gato_b200.workloads import make_config, DEFAULT_SOLVER_PARAMS
from oracle.pyapi import Backend
name = native.Model.builtin("iiwa14").register("iiwa14_as_data")
w = make_config(1, B=3, N=8)
o = Backend("oracle", "iiwa14", 8).solver(3, w["params"])
g = native.Solver(name, 8, 3, w["params"])
gc = native.Solver("iiwa14", 8, 3, w["params"])
rng = np.random.default_rng(0)
for trial in range(3):
    xk = (rng.uniform(-1, 1, 14) * (trial > 0)).astype(np.float32)
    uk = rng.uniform(-5, 5, 7).astype(np.float32)
    a, b, c = g.sim_forward(xk, uk, 0.01), o.sim_forward(xk, uk, 0.01), gc.sim_forward(xk, uk, 0.01)
    print(trial, "rt==oracle", np.array_equal(a, b), "compiled==oracle", np.array_equal(c, b), a[0, :4], b[0, :4])
