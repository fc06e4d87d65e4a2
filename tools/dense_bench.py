#!/usr/bin/env python
"""GPU box: U/J pair-kernel rates of the four families on fields where most pairs are INSIDE
the regularised range (rings of BASELINE config 2, a compact blob) next to the sparse C4 cloud,
far-field shortcut on and off.  usage: python tools/dense_bench.py [case,...] [kernel,...] [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpm_import import load  # noqa: E402

vpm = load()
from flowvpm_jl_b200 import sharding  # noqa: E402

F_UJ = {"singular": 68, "gaussian": 75, "gaussianerf": 78, "winckelmans": 82}


def make(case):
    if case == "c2":  # two leapfrogging rings, nc = 6: 33 800 particles, sigma = Rcross
        R = 0.7906
        return vpm.fields.ring_field(Nphi=100, nc=6, R=R, Rcross=0.1 * R, rings=2, dZ=0.7906)
    if case == "c1":
        return vpm.fields.ring_field(Nphi=100, nc=3)
    if case == "blob":  # compact cloud: 65 536 particles in a unit cube, sigma = 3 lattice spacings
        rng = np.random.Generator(np.random.PCG64(5))
        n = 65536
        pf = vpm.ParticleField(n)
        P = pf.particles
        P[0:3, :n] = rng.random((3, n))
        P[3:6, :n] = rng.standard_normal((3, n)) / n
        P[6, :n] = 3.0 / 40 * (1 + 0.1 * (rng.random(n) - 0.5))
        pf.np = n
        return pf
    if case == "c3":
        return vpm.fields.jet_field(100_000)
    if case.startswith("c4"):
        return vpm.fields.cloud_field(int(case[3:] or 262144) if len(case) > 3 else 262144)
    raise SystemExit(case)


def main():
    cases = (sys.argv[1] if len(sys.argv) > 1 else "c2,blob,c4").split(",")
    kernels = (sys.argv[2] if len(sys.argv) > 2 else "gaussianerf,gaussian,winckelmans,singular").split(",")
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    h = vpm.Handle(1)
    if os.environ.get("UJ_VARIANT"):
        h.set_option(vpm._cabi.OPT_UJ_VARIANT, int(os.environ["UJ_VARIANT"]))
    if os.environ.get("UJ_TABLE"):
        h.set_option(vpm._cabi.OPT_UJ_TABLE, int(os.environ["UJ_TABLE"]))
    import ctypes as C
    dfma, dms = C.c_double(), C.c_double()
    h.check(h.lib.vpm_measure_dfma_peak(h.ptr, C.byref(dfma), C.byref(dms)))
    print(f"DFMA/s {dfma.value:.4g}")
    for case in cases:
        pf = make(case)
        n = pf.np
        X, sig = pf.get_X()[:, :n], pf.particles[6, :n]
        # fraction of pairs inside s < 9 (sampled)
        rng = np.random.Generator(np.random.PCG64(1))
        i, j = rng.integers(0, n, 200000), rng.integers(0, n, 200000)
        s = np.linalg.norm(X[:, i] - X[:, j], axis=0) / sig[j]
        print(f"case {case}: n = {n}, pairs with s < 9: {np.mean(s < 9):.3f}, s < 3.45: {np.mean(s < 3.45):.3f}")
        src8 = torch.from_numpy(np.ascontiguousarray(vpm.source_system_to_buffer(pf).T)).cuda()
        for k in kernels:
            f = sharding.ShardedField(h, src8, n, 0, 1, vpm.KERNELS[k].id)
            for flags, tag in ((0, "shortcut on "), (vpm._cabi.FLAG_NO_FARFIELD_SHORTCUT, "shortcut off")):
                if flags and k in ("winckelmans", "singular"):
                    continue
                f.uj(flags)
                torch.cuda.synchronize()
                best = 1e30
                for _ in range(reps):
                    f.uj(flags)
                    torch.cuda.synchronize()
                    best = min(best, h.timing()["uj_ms"])
                rate = n * n / best / 1e6
                print(f"  {k:12s} {tag}: {best:9.3f} ms  {rate:7.1f} G/s  frac(all pairs at {F_UJ[k]} flop) "
                      f"{rate * 1e9 * F_UJ[k] / (2 * dfma.value):.3f}", flush=True)


if __name__ == "__main__":
    main()
