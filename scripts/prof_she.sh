#!/bin/bash
# ncu capture of the shenanigans median kernel (GPU box); exports compact CSVs
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_bkgshe_median" -s 1 -c 1 -o /tmp/prof_she python scripts/dev_she.py 8 > gpurun_out/prof_she.log 2>&1
ncu -i /tmp/prof_she.ncu-rep --page raw --csv > gpurun_out/prof_she_raw.csv
ncu -i /tmp/prof_she.ncu-rep --page source --csv > gpurun_out/src_she.csv
ls -la gpurun_out | tail -4
