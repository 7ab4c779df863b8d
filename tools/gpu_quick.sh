#!/bin/bash
# quick GPU check (under gpurun): GPU tests, then the full bench line of the current build (and --impl reference if REF=1)
#   usage: tools/gpu_quick.sh <tag>
TAG=${1:-cur}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
python bench.py ${BENCH_FLAGS:-} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
f = "gpurun_out/${TAG}_bench.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(round(d["value"]), "solves/s", round(d["ms_per_step"], 4), "ms", {k: round(v["us_per_launch"], 1) for k, v in d["kernels"].items()}, "e2e", round(d.get("e2e", {}).get("value", 0)))
    for k in ("roofline", "reference_gpu", "parity_vs_stock_reference", "default_params_workload", "cpu_baseline", "latency_ms_p95"):
        print(k, json.dumps(d.get(k))[:900])
except Exception as e:
    print(f, "FAILED", e)
PY
if [ -n "${REF:-}" ]; then python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-400; fi
