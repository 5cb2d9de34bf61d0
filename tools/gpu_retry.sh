#!/bin/bash
# Retry a gpurun call until the pod has a slot (exit 3 / "transient" = nothing charged).
#   tools/gpu_retry.sh <log file> <timeout s> [--gpus N] -- '<command>'
log=$1; shift
to=$1; shift
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" "${extra[@]}" -- "$1" > "$log" 2>&1
  rc=$?
  if ! grep -q "status=transient" "$log" && [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
