#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> [--gpus N] '<command>'   -- gpurun, retried while the pod has no free slot (exit 3)
T=$1; shift
G=""
if [ "$1" = "--gpus" ]; then G="--gpus $2"; shift 2; fi
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$T" $G -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
