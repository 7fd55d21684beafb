#!/bin/bash
# gpurun with retries while the pod answers "transient / busy" (nothing is charged for those).
# usage: tools/gpu.sh <timeout-seconds> [--gpus N] -- '<command>'
TO=$1; shift
for i in $(seq 1 30); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout "$TO" "$@" 2>&1)
  if echo "$OUT" | grep -q "status=transient\|answers busy\|no box\|status=busy"; then
    echo "[gpu.sh] attempt $i: busy, retrying in 60 s" >&2
    sleep 60
    continue
  fi
  echo "$OUT"
  exit 0
done
echo "$OUT"
exit 3
