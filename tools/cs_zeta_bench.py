#!/usr/bin/env python
"""Tuning aid (GPU box): cost of CoreSpreading's basis evaluation on the resident field,
zeta_direct (O(N^2)) against zeta_fmm (near field of device-built leaf lists), and of the RBF around it."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpm_import import load  # noqa: E402

vpm = load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
ncrit = int(sys.argv[2]) if len(sys.argv) > 2 else 50
pf = vpm.fields.cloud_field(n, kernel=vpm.gaussianerf)
pf.particles[7, :n] = (1.0 / n)            # volumes: the RBF's initial guess
rf = vpm.ResidentField(pf)
h = rf.h
out = {}
for method in ("direct", "fmm_reset"):
    rf.upload()
    rf.zeta_method(method, ncrit=ncrit, theta=0.4)
    rf.zeta()                                # warm-up (builds the lists)
    t = time.perf_counter()
    for _ in range(3):
        rf.zeta()
    dt = (time.perf_counter() - t) / 3
    rf.download()
    out[method] = pf.particles[15:18, :n].copy()
    pf.particles[33:36, :n] = pf.particles[15:18, :n]     # target vorticity = what the field represents
    rf.upload()
    t = time.perf_counter()
    it, res = rf.rbf_conjugategradient(itmax=10, tol=1e-6, iterror=False)
    dr = time.perf_counter() - t
    print(f"n={n} ncrit={ncrit} zeta={method:9s}: {dt*1e3:8.2f} ms per evaluation; RBF {it} iterations in {dr*1e3:8.1f} ms "
          f"(residuals {res})", flush=True)
rf.zeta_method("direct")
d = np.abs(out["fmm_reset"] - out["direct"]).max() / np.abs(out["direct"]).max()
print(f"max |zeta_fmm - zeta_direct| / max |zeta_direct| = {d:.2e}")
