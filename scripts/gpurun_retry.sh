#!/bin/sh
# usage: scripts/gpurun_retry.sh <gpurun args...>   -- retries while the pod answers "busy" (rc 3, nothing charged)
n=0
while :; do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  [ "$rc" -ne 3 ] && exit $rc
  n=$((n + 1))
  [ "$n" -ge 12 ] && exit 3
  sleep 120
done
