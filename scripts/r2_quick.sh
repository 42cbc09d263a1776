#!/bin/bash
# quick GPU check: GPU tests + fallback rates + per-kernel-class timing + per-launch times
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python scripts/dev_zone.py 16 2>&1 | tail -3
timeout 300 python scripts/dev_perf.py 64 2>&1 | tail -8
bash scripts/r2_times.sh 64 2>&1 | tail -32
