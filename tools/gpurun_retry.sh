#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun options] -- <command>      retries while the pod answers "transient" (no box / slot free)
for attempt in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  rc=$?
  if echo "$out" | grep -q "status=transient"; then
    sleep 45
    continue
  fi
  echo "$out"
  exit $rc
done
echo "gpurun_retry: still transient after 40 attempts"
exit 3
