#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <script> [--gpus N]   -- retries while the pod answers busy (nothing is charged for those)
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun $3 $4 --timeout $1 -- "bash $2" 2>&1)
  echo "$out" | tail -4
  if echo "$out" | grep -q "status=transient"; then sleep 150; continue; fi
  break
done
