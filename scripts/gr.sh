#!/bin/bash
# developer helper: gpurun with retries while the pod is busy.  usage: gr.sh <timeout-seconds> <command...>
T=$1; shift
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then sleep 90; continue; fi
  break
done
