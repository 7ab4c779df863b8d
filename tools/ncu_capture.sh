#!/bin/bash
# Run on the GPU box (under gpurun): full ncu capture of one launch of each hot kernel inside the bench workload,
# exported to CSV under gpurun_out/ (the .ncu-rep files stay in /tmp: gpurun_out is capped at 64 MiB).
set -u
TAG=${1:-cur}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-ref-gpu --no-e2e"
for K in ${KERNELS:-k_pcg k_schur k_kkt k_merit_ls}; do
    # skip the first warm-up step's launches of this kernel, capture one launch
    ncu --set full --clock-control none --import-source on -k regex:"^${K}" -s 5 -c 1 -f -o /tmp/prof_${K} $CMD > /tmp/ncu_${K}.log 2>&1
    ncu -i /tmp/prof_${K}.ncu-rep --page raw --csv > gpurun_out/${TAG}_${K}_raw.csv 2>/dev/null
    ncu -i /tmp/prof_${K}.ncu-rep --page details --csv > gpurun_out/${TAG}_${K}_details.csv 2>/dev/null
    ncu -i /tmp/prof_${K}.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_${K}_source_sass.csv 2>/dev/null
    ncu -i /tmp/prof_${K}.ncu-rep --page source --csv --print-source cuda > gpurun_out/${TAG}_${K}_source_cuda.csv 2>/dev/null
done
ls -la gpurun_out | tail -20
