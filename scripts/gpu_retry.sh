#!/bin/bash
# usage: scripts/gpu_retry.sh [--gpus N] TIMEOUT 'command'   -- retries while the pod answers busy
G=""
if [ "$1" = "--gpus" ]; then G="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 12); do
	out=$(/usr/local/graft/bin/gpurun $G --timeout $T -- "$@" 2>&1)
	if echo "$out" | grep -q "status=transient\|rc=3\|exit code 3"; then sleep 100; else echo "$out"; exit 0; fi
done
echo "gave up: pod busy"
