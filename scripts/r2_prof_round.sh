#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tile_round_z|k_tile_round_fin" -s 4 -c 2 -o /tmp/prof_zr python scripts/prof_run.py ${1:-32} > gpurun_out/prof_zr.log 2>&1
ncu -i /tmp/prof_zr.ncu-rep --page raw --csv > gpurun_out/prof_zr_raw.csv
ncu -i /tmp/prof_zr.ncu-rep --page source --csv --kernel-name regex:k_tile_round_z > gpurun_out/src_zr.csv
ncu -i /tmp/prof_zr.ncu-rep --page source --csv --kernel-name regex:k_tile_round_fin > gpurun_out/src_zf.csv
tail -2 gpurun_out/prof_zr.log
