#!/bin/bash
# end-of-round GPU job: whole GPU suite, full bench line (all legs), reference arm, ncu evidence.  usage: tools/gpu_job_final.sh <tag>
TAG=${1:-final}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -2 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2>> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), "solves/s", round(d["ms_per_step"], 4), "ms", {k: round(v["us_per_launch"], 1) for k, v in d["kernels"].items()}, "e2e", round(d.get("e2e", {}).get("value", 0)))
for k in ("table_driven_model", "default_params_workload", "reference_gpu", "cpu_baseline", "clocks"):
    print(k, json.dumps(d.get(k))[:600])
PY
python tools/rt_vs_compiled.py 512 32 iiwa14 > gpurun_out/${TAG}_rt_vs_compiled_iiwa14.json 2>/dev/null; cat gpurun_out/${TAG}_rt_vs_compiled_iiwa14.json
python tools/rt_vs_compiled.py 512 32 indy7 > gpurun_out/${TAG}_rt_vs_compiled_indy7.json 2>/dev/null; cat gpurun_out/${TAG}_rt_vs_compiled_indy7.json
tools/ncu_capture_r02.sh ${TAG} 2>&1 | tail -12
tools/ncu_capture_rt.sh ${TAG} 2>&1 | tail -3
