#!/bin/bash
# GPU job for the run-time model path: new tests first, then the whole GPU suite, then table-driven vs compiled timings
TAG=${1:-rt}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -x -q -s > gpurun_out/${TAG}_pytest_models.log 2>&1
tail -15 gpurun_out/${TAG}_pytest_models.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
python tools/rt_vs_compiled.py 512 32 iiwa14 > gpurun_out/${TAG}_rt_vs_compiled_iiwa14.json 2> gpurun_out/${TAG}_rt_vs_compiled.err; cat gpurun_out/${TAG}_rt_vs_compiled_iiwa14.json; tail -3 gpurun_out/${TAG}_rt_vs_compiled.err
python tools/rt_vs_compiled.py 512 32 indy7 > gpurun_out/${TAG}_rt_vs_compiled_indy7.json 2>> gpurun_out/${TAG}_rt_vs_compiled.err; cat gpurun_out/${TAG}_rt_vs_compiled_indy7.json
python bench.py --no-cpu --no-ref-gpu --no-extra > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), "solves/s", round(d["ms_per_step"], 4), "ms", {k: round(v["us_per_launch"], 1) for k, v in d["kernels"].items()}, "e2e", round(d.get("e2e", {}).get("value", 0)))
PY
