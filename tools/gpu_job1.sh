#!/bin/bash
# first GPU job of round 2: tests, FP32 peak, A/B bench of the new k_pcg against the round-1 library
mkdir -p gpurun_out
python tools/fp32_peak.py > gpurun_out/r02_fp32_peak.json 2>&1
cat gpurun_out/r02_fp32_peak.json
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_job1_pytest.log 2>&1
tail -15 gpurun_out/r02_job1_pytest.log
python bench.py --no-cpu --no-ref-gpu > gpurun_out/r02_job1_bench_new.json 2> gpurun_out/r02_job1_bench_new.err
cat gpurun_out/r02_job1_bench_new.json
GATO_B200_LIB=build_tmp/libgato_b200_r01.so python bench.py --no-cpu --no-ref-gpu > gpurun_out/r02_job1_bench_r01.json 2> gpurun_out/r02_job1_bench_r01.err
cat gpurun_out/r02_job1_bench_r01.json
