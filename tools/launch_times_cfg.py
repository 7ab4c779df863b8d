"""Per-kernel device time of one solve of a BASELINE.json config (event per launch; the cluster path launches k_pcg_stream's K2 phase and
k_pcg_cluster inside one 'k_pcg' bracket).   usage: python tools/launch_times_cfg.py <cfg>   (GPU box)"""
import sys

sys.path.insert(0, ".")
from gato_b200 import native
from gato_b200.workloads import make_config

w = make_config(int(sys.argv[1]))
s = native.Solver(w["plant"], w["N"], w["B"], w["params"], device=0)
s.set_kernel_timing(True)
for _ in range(3):
    s.reset("dual"), s.reset("rho")
    r = s.solve(w["xu"], w["xs"], w["ref"], w["dt"])
print({k: (round(ms, 3), n) for k, (ms, n) in s.kernel_times().items()}, "device ms", round(r["device_time_ms"], 3), "pcg mean", float(r["pcg_iters"].mean()))
