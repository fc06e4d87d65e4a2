#!/usr/bin/env python
"""Tuning aid (GPU box): FMM near-field hook (vpm_p2p_leafpairs) throughput vs leaf size."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpm_import import load  # noqa: E402

vpm = load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
kernel = vpm.KERNELS[sys.argv[2]] if len(sys.argv) > 2 else vpm.winckelmans
h = vpm.get_handle()
pf = vpm.fields.cloud_field(n, kernel=kernel)
for ncrit in [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["64", "512", "1600"])]:
    t = time.perf_counter()
    ll = vpm.fields.build_leaf_lists(pf.get_X(), pf.get_sigma(), ncrit=ncrit, theta=0.4)
    t_build = time.perf_counter() - t
    order = ll["sort_index"]
    sb = np.asfortranarray(vpm.source_system_to_buffer(pf)[:, order])
    tb = np.zeros((16, n), order="F")
    tb[0:3] = pf.get_X()[:, order]
    leaves = (ll["leaf_begin"], ll["leaf_end"])
    sizes = ll["leaf_end"] - ll["leaf_begin"]
    dl = ll["direct_list"]
    pairs = int((sizes[dl[:, 0]].astype(np.int64) * sizes[dl[:, 1]]).sum())
    for rep in range(2):
        tb[4:] = 0
        t = time.perf_counter()
        vpm.nearfield_device(tb, leaves, sb, leaves, dl, kernel)
        dt = time.perf_counter() - t
    tm = h.timing()
    print(f"n={n} ncrit={ncrit} leaves={len(sizes)} mean={sizes.mean():.0f} max={sizes.max()} list={len(dl)} "
          f"interactions={pairs:.3e} host-build={t_build:.1f}s | call {dt*1e3:.1f} ms (h2d {tm['h2d_ms']:.1f} kernel {tm['uj_ms']:.1f} d2h {tm['d2h_ms']:.1f}) "
          f"-> kernel {pairs / tm['uj_ms'] / 1e6:.1f} G/s, e2e {pairs / dt / 1e9:.1f} G/s", flush=True)
    if len(sys.argv) > 4 and sys.argv[4] == "estr":  # Estr_fmm! over the same list (SFS leaf kernel)
        vpm.fields.random_results(pf, scale=1e-2)
        for rep in range(2):
            t = time.perf_counter()
            vpm.Estr_fmm(pf, order, order, leaves, leaves, dl)
            dt = time.perf_counter() - t
        tm = h.timing()
        print(f"   Estr_fmm: call {dt*1e3:.1f} ms (h2d {tm['h2d_ms']:.1f} kernel {tm['sfs_ms']:.1f} d2h {tm['d2h_ms']:.1f}) "
              f"-> kernel {pairs / tm['sfs_ms'] / 1e6:.1f} G/s", flush=True)
    if len(sys.argv) > 4 and sys.argv[4] == "zeta":  # the zeta_fmm leaf kernel (SFS-family tile producer)
        for rep in range(2):
            t = time.perf_counter()
            vpm.zeta_fmm(pf, order, leaves, dl)
            dt = time.perf_counter() - t
        tm = h.timing()
        print(f"   zeta_fmm: call {dt*1e3:.1f} ms, timing {tm} -> e2e {pairs / dt / 1e9:.1f} G/s", flush=True)
