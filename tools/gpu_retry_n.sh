#!/bin/bash
# tools/gpu_retry_n.sh NGPUS TIMEOUT LOGNAME 'command' -- multi-GPU gpurun call with retries while the pod is busy
n=$1; t=$2; log=$3; shift 3
for i in 1 2 3 4 5 6 7 8 9 10 11 12 13 14 15 16; do
  /usr/local/graft/bin/gpurun --gpus $n --timeout $t -- "$@" > gpurun_out/$log 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" gpurun_out/$log || [ $rc -eq 3 ]; then sleep 120; continue; fi
  exit $rc
done
exit 3
