#!/bin/bash
# usage: retry.sh <log> <gpurun args...>   -- re-submits while the pod answers "transient" (rc 3), nothing is charged for those
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if grep -q "status=transient" "$log"; then sleep 90; continue; fi
  break
done
tail -30 "$log"
