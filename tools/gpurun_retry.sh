#!/bin/bash
# usage: tools/gpurun_retry.sh [--gpus N] <timeout> '<command>'   — retries while the pod answers "busy" (nothing charged)
GP=""
if [ "$1" == "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun $GP --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "gave up after 40 tries"; echo "$out" | tail -3
