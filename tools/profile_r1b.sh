#!/bin/bash
# Run on the GPU box (gpurun): launch list of the bench command + full ncu captures of the FP64
# pair kernel, the FP32-mode pair kernel and the near-field leaf kernel (second half of round 1).
# Outputs land in gpurun_out/; tools/ncu_summary.py turns them into profiles/*.txt here.
set -x
R=${1:-r1b}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --particles 262144 --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_launches_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:uj_pairs_kernel -s 1 -c 1 -o gpurun_out/prof_uj_$R -f \
    python bench.py --particles 262144 --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_uj_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"uj_pairs_kernel_f32|uj_leaf_kernel" -c 2 -o gpurun_out/prof_f32_leaf_$R -f \
    python -c "
import sys; sys.path.insert(0, '.')
from vpm_import import vpm
pf = vpm.fields.cloud_field(131072, kernel=vpm.winckelmans)
vpm.UJ_direct(pf, reset=True, fp32=True)
vpm.leaf_lists(pf, ncrit=64, theta=0.4, fetch=False)
vpm.UJ_nearfield(pf, reset=True)
" > gpurun_out/ncu_f32_leaf_$R.log 2>&1
ls -la gpurun_out | tail -8
