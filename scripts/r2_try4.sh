#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/dev_zone.py 16 2>&1 | tail -4
timeout 300 python scripts/dev_perf.py 64 2>&1 | grep "chunk=64:\|nstreams=4 chunk=64"
timeout 600 python scripts/dev_crowded.py 128 2>&1 | tail -3
