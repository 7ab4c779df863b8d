#!/bin/bash
# multi-GPU job (under gpurun --gpus N): GPU tests (incl. the NCCL closed-loop test), closed-loop config 5 on all GPUs, bench.py under torchrun
NG=${1:-2}; TAG=${2:-r02_multi}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29655 tools/closed_loop_multi_gpu.py --per-gpu 1024 --steps 200 --out gpurun_out/${TAG}_closed_loop_${NG}gpu.json 2> gpurun_out/${TAG}_closed_loop.err | tail -2
timeout 600 python tools/closed_loop_multi_gpu.py --per-gpu 1024 --steps 200 --out gpurun_out/${TAG}_closed_loop_1gpu.json 2>> gpurun_out/${TAG}_closed_loop.err | tail -1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus $NG --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_${NG}gpu.json 2> gpurun_out/${TAG}_bench_${NG}gpu.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_${NG}gpu.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "e2e", "strong_scaling", "default_params_workload")})
PY
tail -3 gpurun_out/${TAG}_closed_loop.err
