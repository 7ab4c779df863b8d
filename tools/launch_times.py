"""Per-launch CUDA-event durations of one solve of the bench workload (run on a GPU box): shows how each kernel's time changes
from the first SQP iteration (cold L2 after the flush) to the later ones.   usage: python tools/launch_times.py [batch [knot_points]]"""
import sys

sys.path.insert(0, ".")
import numpy as np
import torch

from gato_b200 import native
from gato_b200.workloads import make_config

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = int(sys.argv[2]) if len(sys.argv) > 2 else 32
w = make_config("bench", B=B, N=N)
xu0 = torch.from_numpy(w["xu"].copy()).cuda()
xs = torch.from_numpy(w["xs"].copy()).cuda()
ref = torch.from_numpy(w["ref"].copy()).cuda()
xu = xu0.clone()
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
solver = native.Solver(w["plant"], w["N"], B, w["params"], device=0, stream=stream.cuda_stream)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
solver.set_kernel_timing(True)
acc = None
reps = 8
for it in range(3 + reps):
    xu.copy_(xu0)
    solver.reset("dual")
    solver.reset("rho")
    flush.zero_()
    solver.solve_async(xu.data_ptr(), xs.data_ptr(), ref.data_ptr(), float(w["dt"]))
    solver.solve_wait()
    lt = solver.launch_times()
    if it >= 3:
        acc = [(k, 0.0) for k, _ in lt] if acc is None else acc
        acc = [(k, a + ms) for (k, a), (_, ms) in zip(acc, lt)]
print(f"batch {B}: mean over {reps} solves, microseconds per launch in launch order")
for k, a in acc:
    print(f"  {k:16s} {1e3 * a / reps:8.1f}")
