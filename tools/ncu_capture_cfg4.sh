#!/bin/bash
# full ncu capture of one k_pcg_stream launch inside BASELINE config 4 (iiwa14, N=128, B=1024); CSV exports into gpurun_out/
TAG=${1:-cfg4}
ncu --set full --clock-control none --import-source on -k regex:"^k_pcg_stream" -s 1 -c 1 -f -o /tmp/prof_stream python tools/config_bench.py 4 mine > /tmp/ncu_stream.log 2>&1
ncu -i /tmp/prof_stream.ncu-rep --page raw --csv > gpurun_out/${TAG}_k_pcg_stream_raw.csv 2>/dev/null
ncu -i /tmp/prof_stream.ncu-rep --page details --csv > gpurun_out/${TAG}_k_pcg_stream_details.csv 2>/dev/null
ncu -i /tmp/prof_stream.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_k_pcg_stream_source_sass.csv 2>/dev/null
tail -3 /tmp/ncu_stream.log
