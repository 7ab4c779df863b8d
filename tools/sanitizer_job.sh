#!/bin/bash
# compute-sanitizer over the GPU parity tests (under gpurun): memcheck on a broad subset, racecheck + synccheck on a narrow one
mkdir -p gpurun_out
export GATO_NO_GRAPH=1   # the sanitizer instruments launches; keep them direct
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mpc.py tests/test_gpu_regressions.py -m gpu -x -q -k "not full_size and not golden and not baseline_configs and not randomized" > gpurun_out/r02_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/r02_memcheck.log | tail -3
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_regressions.py -m gpu -x -q -k "(every_stage and (iiwa14-8 or indy7-7 or iiwa14-128 or iiwa14-9)) or (whole_solve and iiwa14-8-1-1) or (cluster and (iiwa14-65 or indy7-41)) or kkt_large_grid" > gpurun_out/r02_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/r02_racecheck.log | tail -5
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_regressions.py -m gpu -x -q -k "(whole_solve and iiwa14-8-1-1) or (cluster and iiwa14-65)" > gpurun_out/r02_synccheck.log 2>&1
echo "synccheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/r02_synccheck.log | tail -3
