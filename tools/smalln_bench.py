#!/usr/bin/env python
"""GPU box: latency of one UJ_direct(pfield; sfs=true) call at the sizes the reference's own tests use
(100-900 particles: test/runtests_singlevortexring.jl:17-31; 200: runtests_leapfrog.jl:48-49) up to
BASELINE config 2 (33 800), host matrix pageable / pinned, next to the CPU port of the reference on
all host cores.  usage: python tools/smalln_bench.py [kernel]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpm_import import load  # noqa: E402

vpm = load()
from oracle import oracle  # noqa: E402  (bench/test infrastructure: the CPU column)


def field(n, kernel):
    if n == 33800:
        R = 0.7906
        return vpm.fields.ring_field(Nphi=100, nc=6, R=R, Rcross=0.1 * R, rings=2, dZ=0.7906, kernel=kernel)
    if n == 4900:
        return vpm.fields.ring_field(Nphi=100, nc=3, kernel=kernel)
    if n == 900:
        return vpm.fields.ring_field(Nphi=100, nc=1, kernel=kernel)
    if n == 200:
        return vpm.fields.ring_field(Nphi=100, nc=0, rings=2, dZ=0.79, kernel=kernel)
    return vpm.fields.cloud_field(n, kernel=kernel)


def measure(h, kname="gaussianerf", sizes=(200, 900, 4900, 33800), reps=30):
    kernel = vpm.KERNELS[kname]
    rows = []
    for n in sizes:
        pf = field(n, kernel)
        assert pf.np == n, (pf.np, n)
        row = {"n": n}
        for mode in ("pageable", "pinned"):
            if mode == "pinned":
                h.check(h.lib.vpm_pin_host(h.ptr, pf.particles.ctypes.data, pf.particles.nbytes))
            for _ in range(3):
                vpm.UJ_direct(pf, sfs=True, reset=True, reset_sfs=True, handle=h)
            t = time.perf_counter()
            for _ in range(reps):
                vpm.UJ_direct(pf, sfs=True, reset=True, reset_sfs=True, handle=h)
            row[f"gpu_{mode}_us"] = (time.perf_counter() - t) / reps * 1e6
            # device-side phases of the last call in microseconds (a replayed graph reports its total only)
            row[f"timing_{mode}"] = {k[:-3] + "_us": round(v * 1e3, 1) for k, v in h.timing().items()
                                     if k.endswith("_ms") and v > 0}
            if mode == "pinned":
                h.check(h.lib.vpm_unpin_host(h.ptr, pf.particles.ctypes.data))
        ref = pf.particles.copy(order="F")
        threads = oracle.num_procs()
        oracle.uj_direct(ref, n, kname, sfs=True, reset=True, reset_sfs=True, nthreads=threads)
        k = max(1, min(20, int(2e8 / (n * n))))
        t = time.perf_counter()
        for _ in range(k):
            oracle.uj_direct(ref, n, kname, sfs=True, reset=True, reset_sfs=True, nthreads=threads)
        row["cpu_port_us"] = (time.perf_counter() - t) / k * 1e6
        row["cpu_threads"] = threads
        rows.append(row)
    return rows


if __name__ == "__main__":
    h = vpm.Handle(1)
    if os.environ.get("SMALL_GRAPH"):
        h.set_option(vpm._cabi.OPT_SMALL_GRAPH, int(os.environ["SMALL_GRAPH"]))
    for r in measure(h, sys.argv[1] if len(sys.argv) > 1 else "gaussianerf"):
        print(r, flush=True)
