#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout> <script> <outfile>  -- retries while the pod answers busy (exit 3 / transient)
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "bash $2" > "$3" 2>&1
  rc=$?
  if grep -q "status=transient" "$3" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
exit $rc
