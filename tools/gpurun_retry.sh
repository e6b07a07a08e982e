#!/bin/bash
# usage: [GPUS=N] tools/gpurun_retry.sh <logfile> <timeout> <command string>   -- retries while the pod answers "transient"/busy
log=$1; to=$2; shift 2
g=""; [ -n "$GPUS" ] && g="--gpus $GPUS"
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun $g --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient\|status=busy" $log; then sleep 45; continue; fi
  break
done
