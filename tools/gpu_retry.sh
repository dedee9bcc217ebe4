#!/bin/bash
# Retries a gpurun call while the pod answers "busy" (exit 3: nothing charged).
#   tools/gpu_retry.sh <tag> <timeout_s> [--gpus N] -- '<command>'
tag=$1; shift; to=$1; shift
for n in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$to" "$@" > gpurun_out/$tag.out 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
