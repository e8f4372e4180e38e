#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> [gpurun args...] -- '<command>'   (retries while the pod answers "transient" / busy)
log=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if grep -q "status=" "$log" && ! grep -q "status=transient" "$log"; then exit 0; fi   # any verdict but "busy, nothing charged" ends the loop
  sleep 90
done
exit 1
