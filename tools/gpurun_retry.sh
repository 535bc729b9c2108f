#!/usr/bin/env bash
# usage: tools/gpurun_retry.sh <out-file> <gpurun args...>   — retries while the pod answers "transient" (exit 3), nothing is charged for those
out="$1"; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
