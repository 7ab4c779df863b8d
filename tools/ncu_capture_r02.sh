#!/bin/bash
# Run on the GPU box (under gpurun): ncu evidence of the current build -- the launch list of one bench step and one full capture of each hot
# kernel inside the bench workload (plus k_pcg_cluster inside BASELINE config 4), exported to CSV under gpurun_out/ (the .ncu-rep files stay
# in /tmp: gpurun_out is capped at 64 MiB).  usage: tools/ncu_capture_r02.sh <tag>
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-ref-gpu --no-e2e --no-extra"
ncu --metrics gpu__time_duration.sum --clock-control none -s 68 -c 40 --csv --log-file gpurun_out/${TAG}_launches_bench.csv $CMD > /tmp/ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches_bench.csv > gpurun_out/${TAG}_launches_bench_summary.txt 2>&1
cat gpurun_out/${TAG}_launches_bench_summary.txt
cap() {  # name, regex, skip, command...
    local NAME=$1 RE=$2 SKIP=$3; shift 3
    ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$RE" -s $SKIP -c 1 -f -o /tmp/prof_${NAME} "$@" > /tmp/ncu_${NAME}.log 2>&1
    ncu -i /tmp/prof_${NAME}.ncu-rep --page raw --csv > gpurun_out/${TAG}_${NAME}_raw.csv 2>/dev/null
    ncu -i /tmp/prof_${NAME}.ncu-rep --page details --csv > gpurun_out/${TAG}_${NAME}_details.csv 2>/dev/null
    ncu -i /tmp/prof_${NAME}.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_${NAME}_source_sass.csv 2>/dev/null
    ncu -i /tmp/prof_${NAME}.ncu-rep --page source --csv --print-source cuda > gpurun_out/${TAG}_${NAME}_source_cuda.csv 2>/dev/null
    tail -1 /tmp/ncu_${NAME}.log | cut -c1-160
}
cap k_pcg 'gato::k_pcg<' 5 $CMD
cap k_schur 'gato::k_schur<' 5 $CMD
cap k_kkt 'gato::k_kkt<' 5 $CMD
cap k_merit_ls8 'gato::k_merit_ls<gato::Iiwa14, \(int\)8' 5 $CMD
cap k_merit_ls1 'gato::k_merit_ls<gato::Iiwa14, \(int\)1' 2 $CMD
cap k_pcg_cluster 'gato::k_pcg_cluster<' 1 python tools/config_bench.py 4 mine
ls -la gpurun_out | grep ${TAG}_ | awk '{print $5, $9}'
