#!/bin/bash
# usage: gpurun_retry.sh <out-file> <gpurun args...>: retries while the pod answers "busy" (exit 3, nothing charged)
out=$1; shift
for i in $(seq 1 30); do
  gpurun "$@" > "$out" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
