#!/bin/bash
# Round-end evidence run on the GPU box (through gpurun): GPU test suite, both bench arms, the ncu
# launch list of the bench command, and fresh full captures of the kernels that changed last.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/final_tests.log
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "reference arm rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_bench_launches_raw.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/final_ncu_bench.log 2>&1; echo "launch list rc=$?"
for k in integrate_kernel derivatives_grain_kernel voigt_average_kernel; do
    skip=0
    [ $k = integrate_kernel ] && skip=6
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f \
        -o gpurun_out/r02_$k python scripts/profile_kernels.py 8 > gpurun_out/r02_ncu_$k.log 2>&1
    if [ -f gpurun_out/r02_$k.ncu-rep ]; then
        ncu -i gpurun_out/r02_$k.ncu-rep --page raw --csv > gpurun_out/r02_${k}_raw.csv 2>/dev/null
        [ $k = integrate_kernel ] && ncu -i gpurun_out/r02_$k.ncu-rep --page source --csv > gpurun_out/r02_${k}_src.csv 2>/dev/null
        rm -f gpurun_out/r02_$k.ncu-rep
    fi
    echo "$k: $(tail -1 gpurun_out/r02_ncu_$k.log)"
done
