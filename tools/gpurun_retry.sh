#!/bin/bash
# gpurun with retries while the pod answers "transient" (nothing charged): tools/gpurun_retry.sh <timeout> [--gpus N] -- '<cmd>'
LOG=${GPURUN_LOG:-/tmp/gpurun_last.log}
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$@" > $LOG 2>&1
  if grep -q "status=transient" $LOG || grep -q "exit code 3" $LOG; then sleep 150; continue; fi
  break
done
tail -80 $LOG
