#!/bin/bash
# source-level ncu capture of several kernels of one tbk_fit_batch: scripts/r2_prof_multi.sh NFFI kernelA kernelB ...
mkdir -p gpurun_out
N=$1; shift
for K in "$@"; do
	timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$K" -s 1 -c 1 -f -o /tmp/prof_$K python scripts/prof_run.py $N > gpurun_out/prof_$K.log 2>&1
	ncu -i /tmp/prof_$K.ncu-rep --page raw --csv > gpurun_out/prof_${K}_raw.csv
	ncu -i /tmp/prof_$K.ncu-rep --page source --csv > gpurun_out/src_$K.csv
	tail -1 gpurun_out/prof_$K.log
done
