#!/bin/bash
# usage: tools/gpu_retry.sh <timeout> '<command>'  - resubmits a gpurun call while the pod answers "busy" (exit code 3 / transient)
t=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  out=$(/usr/local/graft/bin/gpurun --timeout $t -- "$@" 2>&1)
  echo "$out" | tail -40
  if echo "$out" | grep -q "status=transient"; then sleep 150; continue; fi
  break
done
