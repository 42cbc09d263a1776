#!/bin/bash
# Full ncu capture of one tbk_fit_batch (all its launches) on the GPU box; keeps only CSV exports small enough to travel back.
# usage: scripts/prof_fit.sh TAG [NFFI] [LAUNCHES]
TAG=${1:-fit}; N=${2:-16}; L=${3:-32}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"^k_" -s $L -c $L -f -o /tmp/prof_$TAG python scripts/prof_run.py $N > gpurun_out/prof_$TAG.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv
