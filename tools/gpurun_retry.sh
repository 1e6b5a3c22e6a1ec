#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <log> '<command>'   (retries while the pod answers "busy", exit code 3)
t=$1; log=$2; shift 2
# the GPU box runs the in-tree .so: make sure it is current
(cd "$(dirname "$0")/.." && python -c "from multi_stylegan_b200 import _lib; _lib.build()") || exit 9
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
