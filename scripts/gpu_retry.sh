#!/bin/bash
# Run a command on the GPU box, retrying while the pool answers "busy" (nothing is charged for
# those).  usage: scripts/gpu_retry.sh <log file> <gpurun args...>
log=$1; shift
for i in $(seq 1 20); do
    /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
    grep -q "status=transient" "$log" || break
    sleep 60
done
