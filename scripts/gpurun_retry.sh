#!/bin/bash
# usage: scripts/gpurun_retry.sh <logfile> <gpurun args...>   -- retries while the pod is busy
LOG=$1; shift
for i in $(seq 1 40); do
  gpurun "$@" > "$LOG" 2>&1
  if grep -q "status=transient" "$LOG"; then sleep 60; continue; fi
  break
done
tail -5 "$LOG"
