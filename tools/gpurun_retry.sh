#!/bin/bash
# usage: gpurun_retry.sh TIMEOUT_S [--gpus N] 'command'   -- retries while the pod answers "busy" (exit 3), up to ~40 min
T=$1; shift
OPTS=()
if [ "$1" = "--gpus" ]; then OPTS=(--gpus "$2"); shift 2; fi
for i in $(seq 1 14); do
  /usr/local/graft/bin/gpurun "${OPTS[@]}" --timeout "$T" -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 170
done
exit 3
