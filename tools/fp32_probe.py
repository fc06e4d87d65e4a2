#!/usr/bin/env python
"""Tuning aid (GPU box): FFMA / FFMA2 issue rates and the FP32-mode U/J sweep rate next to FP64."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpm_import import load  # noqa: E402

vpm = load()
from flowvpm_jl_b200 import sharding  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
h = vpm.Handle(1)
names = ["FFMA2 invariant operands", "FFMA2 3 distinct registers", "FFMA invariant operands", "FFMA 3 distinct registers"]
for mode in range(4):
    v, ms = C.c_double(), C.c_double()
    h.check(h.lib.vpm_measure_ffma_peak(h.ptr, mode, C.byref(v), C.byref(ms)))
    print(f"{names[mode]:32s} {v.value:.3e} FMA/s  ({ms.value:.2f} ms)", flush=True)
v, ms = C.c_double(), C.c_double()
h.check(h.lib.vpm_measure_dfma_peak(h.ptr, C.byref(v), C.byref(ms)))
print(f"{'DFMA':32s} {v.value:.3e} FMA/s", flush=True)
pf = vpm.fields.cloud_field(n)
src8 = torch.from_numpy(np.ascontiguousarray(vpm.source_system_to_buffer(pf).T)).cuda()
for k in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["winckelmans", "singular", "gaussianerf", "gaussian"]):
    f = sharding.ShardedField(h, src8, n, 0, 1, vpm.KERNELS[k].id)
    for flags, label in ((0, "fp64"), (32, "fp32 unroll1"), (32, "fp32 unroll2")):
        h.set_option(vpm._cabi.OPT_UJ_VARIANT, 0 if flags == 0 else (22 if label.endswith("2") else 21))
        f.uj(flags)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            f.uj(flags)
            torch.cuda.synchronize()
            best = min(best, h.timing()["uj_ms"])
        print(f"{k:12s} {label:13s}: {best:9.3f} ms  {n * n / best / 1e6:8.1f} G interactions/s", flush=True)
