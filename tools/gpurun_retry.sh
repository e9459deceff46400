#!/bin/bash
# retry a gpurun call while the pod answers "busy" (exit 3 / status=transient); usage: tools/gpurun_retry.sh <timeout_s> '<command>'
T=$1; shift
for attempt in $(seq 1 40); do
    out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1); rc=$?
    if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
    echo "$out"; exit $rc
done
echo "gpurun: still busy after 40 attempts"; exit 3
