#!/bin/bash
# ncu captures of the table-driven (run-time model) kernels inside the bench-shaped workload (under gpurun).  usage: tools/ncu_capture_rt.sh <tag>
set -u
TAG=${1:-rt}
mkdir -p gpurun_out
CMD="python tools/rt_vs_compiled.py 512 32 iiwa14"
cap() {
    local NAME=$1 RE=$2 SKIP=$3
    ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$RE" -s $SKIP -c 1 -f -o /tmp/prof_${NAME} $CMD > /tmp/ncu_${NAME}.log 2>&1
    ncu -i /tmp/prof_${NAME}.ncu-rep --page raw --csv > gpurun_out/${TAG}_${NAME}_raw.csv 2>/dev/null
    ncu -i /tmp/prof_${NAME}.ncu-rep --page details --csv > gpurun_out/${TAG}_${NAME}_details.csv 2>/dev/null
    ncu -i /tmp/prof_${NAME}.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_${NAME}_source_sass.csv 2>/dev/null
    tail -1 /tmp/ncu_${NAME}.log | cut -c1-160
}
cap rt_k_kkt 'gato::k_kkt<gato::RtPlant' 5
cap rt_k_merit_ls8 'gato::k_merit_ls<gato::RtPlant<\(int\)7>, \(int\)8' 5
ls -la gpurun_out | grep ${TAG}_ | awk '{print $5, $9}'
