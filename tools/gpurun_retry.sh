#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>' [extra gpurun flags]   -- retries while the pod has no free GPU slot (rc 3)
t=$1; shift; cmd=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" "$@" -- "$cmd"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
