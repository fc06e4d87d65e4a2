#!/bin/bash
# Run on the GPU box (gpurun): launch list + full ncu captures of the two pair kernels.
# Outputs land in gpurun_out/; tools/ncu_summary.py turns them into profiles/*.txt here.
set -x
R=${1:-r1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --particles 262144 --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_launches_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:uj_pairs -s 1 -c 1 -o gpurun_out/prof_uj_$R -f \
    python bench.py --particles 262144 --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_uj_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sfs_pairs -c 1 -o gpurun_out/prof_sfs_$R -f \
    python -c "
import sys; sys.path.insert(0, '.')
from vpm_import import vpm
pf = vpm.fields.cloud_field(131072, kernel=vpm.winckelmans)
vpm.UJ_direct(pf, sfs=True, reset=True, reset_sfs=True)
" > gpurun_out/ncu_sfs_$R.log 2>&1
# full-size capture of the bench kernel for roofline.traffic (one launch, ~40 replays of 3 s)
ncu --set full --clock-control none -k regex:uj_pairs -s 1 -c 1 -o gpurun_out/prof_uj_1M_$R -f \
    python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_uj_1M_$R.log 2>&1
ls -la gpurun_out
