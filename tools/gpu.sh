#!/bin/bash
# usage: tools/gpu.sh [--gpus N] <timeout_s> '<command>'   -- gpurun with retries while the pod answers busy/transient
GP=""
if [ "$1" = "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun $GP --timeout "$T" -- "$@" 2>&1)
  rc=$?
  if echo "$out" | grep -q "status=transient\|status=busy" || [ $rc -eq 3 ]; then
    echo "[gpu.sh] attempt $i: busy/transient (rc=$rc), retrying in 90 s" >&2
    sleep 90
    continue
  fi
  echo "$out"
  exit $rc
done
echo "[gpu.sh] gave up"; exit 3
