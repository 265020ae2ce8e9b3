#!/bin/bash
# tools/gpu_retry.sh TIMEOUT LOGNAME 'command' -- run a gpurun call, retrying while the pod answers busy (exit 3 / transient)
t=$1; log=$2; shift 2
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@" > gpurun_out/$log 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" gpurun_out/$log || [ $rc -eq 3 ]; then sleep 90; continue; fi
  exit $rc
done
exit 3
