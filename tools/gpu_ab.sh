#!/bin/bash
# A/B bench of variant libraries (under gpurun): tools/gpu_ab.sh <tag> variant.so ...
TAG=${1:-ab}; shift
mkdir -p gpurun_out
for V in "$@"; do
    GATO_B200_LIB=gato_b200/lib/variants/$V python bench.py --no-cpu --no-ref-gpu --no-e2e > gpurun_out/${TAG}_bench_${V%.so}.json 2> gpurun_out/${TAG}_bench_${V%.so}.err
    python - <<PY
import json
f = "gpurun_out/${TAG}_bench_${V%.so}.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print("$V", round(d["value"]), "solves/s", round(d["ms_per_step"], 4), "ms", {k: round(v["us_per_launch"], 1) for k, v in d["kernels"].items()})
except Exception as e:
    print(f, "FAILED", e)
PY
done
