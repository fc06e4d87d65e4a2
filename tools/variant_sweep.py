#!/usr/bin/env python
"""Tuning aid (GPU box): time the U/J pair kernel for each launch variant <T><unroll>."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpm_import import load  # noqa: E402

vpm = load()
from flowvpm_jl_b200 import sharding  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
kernels = sys.argv[2].split(",") if len(sys.argv) > 2 else ["winckelmans", "singular", "gaussianerf"]
variants = sys.argv[3].split(",") if len(sys.argv) > 3 else ["11", "12", "14", "21", "22"]
h = vpm.Handle(1)
pf = vpm.fields.cloud_field(n)
src8 = torch.from_numpy(np.ascontiguousarray(vpm.source_system_to_buffer(pf).T)).cuda()
for k in kernels:
    f = sharding.ShardedField(h, src8, n, 0, 1, vpm.KERNELS[k].id)
    for v in variants:
        h.set_option(vpm._cabi.OPT_UJ_VARIANT, int(v))
        f.uj(0)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            f.uj(0)
            torch.cuda.synchronize()
            best = min(best, h.timing()["uj_ms"])
        print(f"{k:12s} variant {v}: {best:9.3f} ms  {n * n / best / 1e6:8.1f} G/s", flush=True)
    if len(sys.argv) > 4:
        for v in variants:
            h.set_option(vpm._cabi.OPT_SFS_VARIANT, int(v) // 10 * 10)
            f.sfs(8)
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(3):
                f.sfs(8)
                torch.cuda.synchronize()
                best = min(best, h.timing()["sfs_ms"])
            print(f"{k:12s} SFS variant {v}: {best:9.3f} ms  {n * n / best / 1e6:8.1f} G/s", flush=True)
