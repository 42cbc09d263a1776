#!/bin/bash
# Round-2 GPU check, one call: GPU tests, parity report (small n), per-kernel-class timing (config 2 and Mars frames).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest.log; cat gpurun_out/r2_pytest.log
timeout 600 python tests/tools/parity_report.py --n ${1:-4} --out gpurun_out/r2_parity.txt > /dev/null 2>&1; cat gpurun_out/r2_parity.txt
timeout 300 python scripts/dev_perf.py 64 2>&1 | tail -8 | tee gpurun_out/r2_perf.txt
timeout 300 python scripts/dev_perf.py 64 --mars --no-stack 2>&1 | tail -3 | tee gpurun_out/r2_perf_mars.txt
