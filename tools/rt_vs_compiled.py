"""Table-driven (run-time model) kernels against the compiled-plant kernels on the same workload (run on a GPU box): per-kernel CUDA-event
time per launch and whole-solve device time; the results of the two must be bit-identical (checked).
usage: python tools/rt_vs_compiled.py [batch [knot_points [plant]]]"""
import json
import sys

sys.path.insert(0, ".")
import numpy as np
import torch

from gato_b200 import native
from gato_b200.workloads import make_config

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = int(sys.argv[2]) if len(sys.argv) > 2 else 32
plant = sys.argv[3] if len(sys.argv) > 3 else "iiwa14"
w = make_config("bench" if plant == "iiwa14" else 3, B=B, N=N)
as_data = native.Model.builtin(plant).register(plant + "_as_data")
xu0 = torch.from_numpy(w["xu"].copy()).cuda()
xs = torch.from_numpy(w["xs"].copy()).cuda()
ref = torch.from_numpy(w["ref"].copy()).cuda()
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
out, res = {}, {}
for name in (plant, as_data):
    solver = native.Solver(name, N, B, w["params"], device=0, stream=stream.cuda_stream)
    xu = xu0.clone()
    per = {}
    for timing in (True, False):
        solver.set_kernel_timing(timing)
        tot, reps = 0.0, 10
        for it in range(3 + reps):
            xu.copy_(xu0)
            solver.reset("dual")
            solver.reset("rho")
            flush.zero_()
            solver.solve_async(xu.data_ptr(), xs.data_ptr(), ref.data_ptr(), float(w["dt"]))
            st = solver.solve_wait()
            if it >= 3:
                tot += st["device_time_ms"]
                if timing:
                    for k, (ms, n) in solver.kernel_times().items():
                        if n:
                            per.setdefault(k, []).append(1e3 * ms / n)
        per["solve_ms_timing" if timing else "solve_ms"] = [tot / reps]
    out[name] = {k: round(float(np.mean(v)), 3) for k, v in per.items()}
    res[name] = (xu.cpu().numpy().copy(), st["pcg_iters"].copy(), st["ls_step_size"].copy())
    solver.close()
same = all(np.array_equal(a, b) for a, b in zip(res[plant], res[as_data]))
print(json.dumps({"workload": f"{plant} N={N} B={B}", "bit_identical": bool(same), "compiled": out[plant], "table_driven": out[as_data]}))
