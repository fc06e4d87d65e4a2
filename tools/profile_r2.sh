#!/bin/bash
# Run on the GPU box (gpurun): round-2 launch list of the bench command, full ncu capture of the
# headline winckelmans pair kernel and of the gaussianerf table kernel on the dense blob.
# Outputs land in gpurun_out/; tools/ncu_summary.py turns them into profiles/*.txt here.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --particles 262144 --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_launches_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:uj_pairs_kernel -s 1 -c 1 -o gpurun_out/prof_uj_r2 -f \
    python bench.py --particles 262144 --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_uj_r2.log 2>&1
ls -la gpurun_out | tail -6
