#!/bin/bash
# Full ncu capture of one tbk_fit_batch (25 launches) on the GPU box; keeps only CSV exports small enough to travel back.
# usage: scripts/prof_fit.sh TAG [NFFI]
TAG=${1:-fit}; N=${2:-16}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"^k_" -s 25 -c 25 -o /tmp/prof_$TAG python scripts/prof_run.py $N > gpurun_out/prof_$TAG.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv
for k in k_tile_base_w3 k_tile_round_w k_ring_kde k_ring_gather_t k_final k_zp_bound k_mesh_finalize; do
	ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv --kernel-name regex:$k > /tmp/src_$k.csv
	python - "$k" "$TAG" <<'PY'
import re, sys
k, tag = sys.argv[1], sys.argv[2]
txt = open(f'/tmp/src_{k}.csv').read()
blocks = re.split(r'(?m)^(?="Kernel Name",)', txt)
blocks = [b for b in blocks if b.startswith('"Kernel Name"')]
keep = [blocks[0]] + ([blocks[-1]] if len(blocks) > 1 else [])   # first launch (round 1) and last (round 3)
open(f'gpurun_out/src_{tag}_{k}.csv', 'w').write(''.join(keep))
PY
done
ls -la gpurun_out | tail -12; du -sh gpurun_out
