#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fit_stack or pack or driver or prepare" 2>&1 | tail -4
timeout 900 python bench.py --no-cpu --no-prepare --no-configs --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
b=json.loads(sys.stdin.read()); e=b['e2e']
print('value', b['value']); print({k:e[k] for k in ('value','copy_ceiling','frac_of_ceiling','copy_ceiling_gbs','h2d_bytes_per_step','d2h_bytes_per_step')})"
