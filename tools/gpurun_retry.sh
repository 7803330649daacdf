#!/bin/bash
# usage: gpurun_retry.sh TIMEOUT_S 'command'   -- retries while the pod answers "busy" (exit 3), up to ~40 min
T=$1; shift
for i in $(seq 1 14); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 170
done
exit 3
