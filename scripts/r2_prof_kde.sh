#!/bin/bash
mkdir -p gpurun_out
K=${2:-k_ring_kde}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 1 -c 1 -f -o /tmp/prof_k python scripts/prof_run.py ${1:-32} > gpurun_out/prof_k.log 2>&1
ncu -i /tmp/prof_k.ncu-rep --page raw --csv > gpurun_out/prof_k_raw.csv
ncu -i /tmp/prof_k.ncu-rep --page source --csv > gpurun_out/src_k.csv
tail -2 gpurun_out/prof_k.log
