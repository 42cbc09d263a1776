#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "time_smooth or prepare or driver or accumulate" 2>&1 | tail -4
timeout 600 python scripts/dev_prep3.py 256 2>&1 | tail -13
