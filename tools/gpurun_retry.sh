#!/bin/bash
# usage: tools/gpurun_retry.sh LOGFILE [--gpus N] --timeout S -- 'command'
# retries while gpurun answers "transient" (no box / slot free: nothing charged)
LOG=$1; shift
for i in 1 2 3 4 5 6; do
  /usr/local/graft/bin/gpurun "$@" > $LOG 2>&1
  if grep -q "status=transient" $LOG || grep -q "exit code 3" $LOG; then sleep 60; continue; fi
  break
done
