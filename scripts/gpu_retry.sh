#!/bin/bash
# usage: scripts/gpu_retry.sh <timeout_s> '<command>'   (extra gpurun flags through GPURUN_FLAGS)
# retries while the pod answers "busy" (exit code 3 / status=transient); nothing is charged for those
t=$1; shift
for i in $(seq 1 40); do
    out=$(/usr/local/graft/bin/gpurun $GPURUN_FLAGS --timeout "$t" -- "$@" 2>&1); rc=$?
    if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
    if [ $rc -eq 3 ]; then sleep 45; continue; fi
    echo "$out"; exit $rc
done
echo "gpu_retry: still busy after 40 tries"; exit 3
