#!/bin/bash
# fallback rate + ncu full capture of the zone base kernels (one batch of 16 FFIs) + variants test
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "variants or golden or full_size or manual" 2>&1 | tail -4
timeout 300 python scripts/dev_zone.py 16 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tile_base" -s 2 -c 2 -o /tmp/prof_zb python scripts/prof_run.py 16 > gpurun_out/prof_zb.log 2>&1
ncu -i /tmp/prof_zb.ncu-rep --page raw --csv > gpurun_out/prof_zb_raw.csv
ncu -i /tmp/prof_zb.ncu-rep --page source --csv --kernel-name regex:k_tile_base_z > gpurun_out/src_zb.csv
tail -2 gpurun_out/prof_zb.log
timeout 300 python scripts/dev_perf.py 64 --no-stack 2>&1 | tail -1
