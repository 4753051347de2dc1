#!/bin/bash
# usage: tools/gpurun_retry.sh [--gpus N] TIMEOUT 'command'   -- retries while the pod answers "transient" (nothing charged)
G=""
if [ "$1" = "--gpus" ]; then G="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun $G --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then echo "[retry $i] transient, sleeping 90 s"; sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "gave up"; exit 3
