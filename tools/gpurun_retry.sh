#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <logfile> <command...>   -- retries while the pod answers busy/transient (rc 3)
t=$1; log=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" $log || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
tail -12 $log
