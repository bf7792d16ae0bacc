#!/bin/bash
# compute-sanitizer over a small-N run of every kernel (scripts/sanitize_run.py); logs go to
# gpurun_out/sanitize_<tool>.log, summaries are copied to profiles/ by hand.
# Usage (on the GPU box): bash scripts/sanitize.sh [memcheck racecheck synccheck]
set -u
mkdir -p gpurun_out
tools=${@:-memcheck racecheck synccheck}
for tool in $tools; do
    for part in integrator others; do
        log=gpurun_out/sanitize_${tool}_${part}.log
        timeout 900 compute-sanitizer --tool $tool --print-limit 20 \
            python scripts/sanitize_run.py $part > $log 2>&1
        echo "== $tool $part: exit $? ==" | tee -a $log
        grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|ok" $log | tail -12
    done
done
