#!/bin/bash
# GPU tests + bench with and without the overlapped line search
TAG=${1:-cur}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log
for MODE in overlap serial; do
  if [ $MODE = serial ]; then export GATO_NO_OVERLAP=1; fi
  python bench.py --no-cpu --no-ref-gpu --no-extra > gpurun_out/${TAG}_bench_$MODE.json 2> gpurun_out/${TAG}_bench_$MODE.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_$MODE.json").read().strip().splitlines()[-1])
print("$MODE", round(d["value"]), "solves/s", round(d["ms_per_step"], 4), "ms", {k: round(v["us_per_launch"], 1) for k, v in d["kernels"].items()}, "e2e", round(d.get("e2e", {}).get("value", 0)))
PY
done
