#!/bin/bash
# Local helper (authoring container): run a gpurun call, retrying while the pod answers "transient" (no slot free, nothing charged).
# usage: tools/gpurun_retry.sh <timeout-seconds> [--gpus N] -- '<command>'
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out"; exit 0
done
echo "gpurun: still transient after 30 tries"; exit 3
