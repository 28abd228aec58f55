#!/bin/bash
# usage: scripts/gpurun_retry.sh <logfile> <timeout_s> [--gpus N] -- <command>   (retries while the pod answers busy)
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to "$@" > $log 2>&1
  rc=$?
  if ! grep -q "status=transient" $log; then exit $rc; fi
  sleep 90
done
