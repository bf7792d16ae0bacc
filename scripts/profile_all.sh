#!/bin/bash
# ncu evidence for every kernel of the path (run on the GPU box through gpurun); CSV exports go
# to gpurun_out/ and are summarised into profiles/ by profiles/summarize_r02.py.
# 1. launch list of the default bench command; 2. one `--set full` capture per kernel.
set -u
mkdir -p gpurun_out
python scripts/profile_kernels.py 8 > gpurun_out/r02_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_raw.csv \
    python scripts/profile_kernels.py 4 > /dev/null 2>&1
for k in integrate_kernel gbs_restore_kernel derivatives_grain_kernel derivatives_gbm_kernel voigt_average_kernel \
         mindex_pairs_reflect_kernel mindex_quats_kernel pathlines_kernel resample_kernel fp64_peak_kernel pack_soa_kernel; do
    skip=0
    [ $k = integrate_kernel ] && skip=6
    [ $k = gbs_restore_kernel ] && skip=6
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f \
        -o gpurun_out/r02_$k python scripts/profile_kernels.py 8 > gpurun_out/r02_ncu_$k.log 2>&1
    if [ -f gpurun_out/r02_$k.ncu-rep ]; then
        ncu -i gpurun_out/r02_$k.ncu-rep --page raw --csv > gpurun_out/r02_${k}_raw.csv 2>/dev/null
        if [ $k = integrate_kernel ]; then
            ncu -i gpurun_out/r02_$k.ncu-rep --page source --csv > gpurun_out/r02_${k}_src.csv 2>/dev/null
        fi
        rm -f gpurun_out/r02_$k.ncu-rep
    fi
    echo "$k: $(tail -1 gpurun_out/r02_ncu_$k.log)"
done
