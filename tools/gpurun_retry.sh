#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3 = nothing charged).  Usage: tools/gpurun_retry.sh LOG TIMEOUT [--gpus N] -- 'command'
LOG=$1; TMO=$2; shift 2
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $TMO "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $LOG; then exit $rc; fi
  sleep 90
done
exit 3
