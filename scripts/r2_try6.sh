#!/bin/bash
for i in 1 2; do
timeout 600 python bench.py --no-cpu --no-configs --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
b=json.loads(sys.stdin.read()); print('pack', b['e2e']['value'], b['prepare_path']['value'], b['prepare_path']['ms'])"
done
TBK_E2E_PACK=0 timeout 600 python bench.py --no-cpu --no-configs --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
b=json.loads(sys.stdin.read()); print('nopack', b['e2e']['value'], b['prepare_path']['value'], b['prepare_path']['ms'])"
