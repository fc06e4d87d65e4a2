#!/usr/bin/env python
"""rVPM step time (BASELINE.json metric, second half): one rungekutta3 step of the reformulated
VPM (f=0, g=1/5) with the DynamicSFS pseudo-3-level procedure (backscatter clipping,
force_positive) and corrected-Pedrizzetti relaxation = 5 U/J + 4 SFS sweeps, entirely on the
device(s) through vpm_field_step.  usage: rvpm_step_bench.py [log2N] [ngpu] [steps]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpm_import import load  # noqa: E402

vpm = load()
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ngpu = int(sys.argv[2]) if len(sys.argv) > 2 else 1
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
n = 1 << logn
h = vpm.Handle(ngpu)
pf = vpm.fields.cloud_field(n, kernel=vpm.gaussianerf)
t = time.perf_counter()
rf = vpm.ResidentField(pf, handle=h)
t_up = time.perf_counter() - t
kw = dict(integration="rungekutta3", f=0.0, g=0.2, sfs="dynamic", clip_backscatter=True, force_positive=True,
          alpha=0.999, sfs_rlxf=0.005, minC=0.0, maxC=1.0, relaxation="correctedpedrizzetti", relax=True, rlxf=0.3)
dt = 1e-4
times = []
for _ in range(steps + 1):   # first one is the warm-up (allocations, NCCL communicators)
    t = time.perf_counter()
    rf.nextstep(dt, **kw)
    times.append(time.perf_counter() - t)
t = time.perf_counter()
rf.download()
t_down = time.perf_counter() - t
import numpy as np
ok = bool(np.isfinite(pf.particles[:, :n]).all())
print(json.dumps({"n_particles": n, "gpus": ngpu, "kernel": "gaussianerf", "what": "RK3 + rVPM + DynamicSFS (pseudo3level, clipping) + "
                  "corrected Pedrizzetti relaxation: 5 U/J + 4 SFS sweeps, device-resident", "warmup_step_s": times[0],
                  "step_s": float(np.mean(times[1:])), "upload_s": t_up, "download_s": t_down,
                  "sweep_interactions_per_step": 9 * n * n, "finite": ok,
                  "C_mean": float(np.abs(pf.particles[36, :n]).mean())}))
