#!/bin/bash
# quick GPU check (under gpurun): GPU tests, then the bench line of the current build and optionally of variant libraries
#   usage: tools/gpu_quick.sh <tag> [variant.so ...]     (variants live in gato_b200/lib/variants/)
TAG=${1:-cur}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
python bench.py --no-cpu --no-ref-gpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
for f in ["gpurun_out/${TAG}_bench.json"]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "solves/s", round(d["ms_per_step"], 4), "ms", {k: round(v["us_per_launch"], 1) for k, v in d["kernels"].items()}, "e2e", round(d["e2e"]["value"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
for V in "$@"; do
    GATO_B200_LIB=gato_b200/lib/variants/$V python bench.py --no-cpu --no-ref-gpu --no-e2e > gpurun_out/${TAG}_bench_${V%.so}.json 2> gpurun_out/${TAG}_bench_${V%.so}.err
    python - <<PY
import json
f = "gpurun_out/${TAG}_bench_${V%.so}.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"]), "solves/s", round(d["ms_per_step"], 4), "ms", {k: round(v["us_per_launch"], 1) for k, v in d["kernels"].items()})
except Exception as e:
    print(f, "FAILED", e)
PY
done
