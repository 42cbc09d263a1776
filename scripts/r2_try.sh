#!/bin/bash
# one GPU call per experiment: subset of GPU tests, fallback rates, kernel-class timing, per-launch times, TMA-staged variant timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python scripts/dev_zone.py 16 2>&1 | tail -3
timeout 300 python scripts/dev_perf.py 64 2>&1 | tail -8
[ -n "$WITH7" ] && TBK_TILE_KERNEL=7 timeout 300 python scripts/dev_perf.py 64 --no-stack 2>&1 | tail -2
bash scripts/r2_times.sh 64 | tail -34
