#!/bin/bash
# usage: tools/ncu_capture_one.sh <tag> <kernel-name-regex> [skip]   -- full ncu capture of one launch inside the bench workload, CSV into gpurun_out/
TAG=$1; RE=$2; SKIP=${3:-2}
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c 1 -f -o /tmp/prof_one python bench.py --steps 1 --warmup 1 --no-cpu --no-ref-gpu --no-e2e > /tmp/ncu_one.log 2>&1
ncu -i /tmp/prof_one.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/prof_one.ncu-rep --page details --csv > gpurun_out/${TAG}_details.csv 2>/dev/null
ncu -i /tmp/prof_one.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_source_sass.csv 2>/dev/null
tail -2 /tmp/ncu_one.log | cut -c1-200
