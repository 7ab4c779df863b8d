#!/bin/bash
# compute-sanitizer over the table-driven (run-time model) kernels (under gpurun): memcheck, racecheck, synccheck on the GPU model tests
mkdir -p gpurun_out
export GATO_NO_GRAPH=1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_models.py -m gpu -x -q -k "not headline" > gpurun_out/r02_rt_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/r02_rt_memcheck.log | tail -3
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_models.py -m gpu -x -q -k "iiwa14-None-8 or custom6" > gpurun_out/r02_rt_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/r02_rt_racecheck.log | tail -5
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_models.py -m gpu -x -q -k "iiwa14-None-8" > gpurun_out/r02_rt_synccheck.log 2>&1
echo "synccheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/r02_rt_synccheck.log | tail -3
