#!/bin/bash
# usage: scripts/gpu_retry.sh <timeout> '<command>'  -- retries gpurun while the pod answers "retry in a few minutes" / busy
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$1" -- "$2" 2>&1)
  if echo "$out" | grep -q "retry in a few minutes\|no box or slot\|status=busy\|status=transient\|backing off"; then sleep 90; continue; fi
  echo "$out" | tail -6
  exit 0
done
echo "gave up"; exit 3
