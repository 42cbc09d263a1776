#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python scripts/dev_perf.py 64 2>&1 | tail -8
TBK_FINAL_MINB=3 timeout 300 python scripts/dev_perf.py 64 --no-stack 2>&1 | tail -2
bash scripts/r2_times.sh 64 | tail -34
