#!/bin/bash
# quick GPU check: GPU tests + fallback rates + per-kernel-class timing (config 2 and Mars frames)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python scripts/dev_zone.py 16 2>&1 | tail -3
timeout 300 python scripts/dev_perf.py 64 2>&1 | tail -8
timeout 300 python scripts/dev_perf.py 64 --mars --no-stack 2>&1 | tail -2
