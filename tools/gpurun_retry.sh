#!/bin/bash
# usage: gpurun_retry.sh <log> <gpurun args...>   retries while the pod answers "busy" (exit code 3)
log=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1; rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 150
done
exit 3
