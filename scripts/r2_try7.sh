#!/bin/bash
show='import json,sys
b=json.loads(sys.stdin.read()); print(sys.argv[1], b["e2e"]["value"], b["e2e"]["copy_ceiling"], b["prepare_path"]["value"], b["prepare_path"]["ms"])'
timeout 600 python bench.py --no-cpu 2>/dev/null | tail -1 | python -c "$show" nocpu
timeout 600 python bench.py 2>/dev/null | tail -1 | tee gpurun_out/bench_final.json | python -c "$show" full
timeout 600 python bench.py --no-cpu 2>/dev/null | tail -1 | python -c "$show" nocpu
