#!/bin/bash
# usage: tools/gpurun_retry.sh LOGFILE [--gpus N] --timeout S -- 'command'
# retries while gpurun answers that nothing was charged (no box / slot free, pod draining): up to ~40 min
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > $LOG 2>&1
  if grep -q "nothing was charged" $LOG || grep -q "status=transient" $LOG || grep -q "exit code 3" $LOG; then sleep 60; continue; fi
  break
done
