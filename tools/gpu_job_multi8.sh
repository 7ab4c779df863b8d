#!/bin/bash
# 8-GPU job (under gpurun --gpus 8): BASELINE config 5 for real (8 x 1024 hypotheses, NCCL) with the bit-for-bit check on a shorter run, and bench.py under torchrun
NG=${1:-8}; TAG=${2:-r02_8gpu}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29655 tools/closed_loop_multi_gpu.py --per-gpu 1024 --steps 200 --out gpurun_out/${TAG}_closed_loop_${NG}gpu.json 2> gpurun_out/${TAG}_closed_loop.err | tail -1 | cut -c1-700
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29657 tools/closed_loop_multi_gpu.py --per-gpu 128 --steps 20 --check 2>> gpurun_out/${TAG}_closed_loop.err | tail -1 | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus $NG --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_${NG}gpu.json 2> gpurun_out/${TAG}_bench_${NG}gpu.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_${NG}gpu.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "e2e", "strong_scaling", "default_params_workload", "clocks")})
PY
tail -2 gpurun_out/${TAG}_closed_loop.err | cut -c1-300
