#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>'   -- gpurun, retried while the pod has no free slot (exit 3)
T=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
