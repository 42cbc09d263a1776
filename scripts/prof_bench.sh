#!/bin/bash
# Round profile on the GPU box: (1) launch list of a short bench run, (2) full ncu capture of the fit kernels.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_" --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --ffis 128 --no-cpu --no-prepare --e2e-ffis 32 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"^k_" -s 25 -c 25 -o gpurun_out/prof_fit \
    python scripts/prof_run.py 32 > gpurun_out/prof_fit.log 2>&1
ls -la gpurun_out
