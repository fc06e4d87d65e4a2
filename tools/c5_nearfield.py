#!/usr/bin/env python
"""One-off measurement (GPU box): BASELINE.json configs[4] -- FMM near-field offload on a
2^24-particle cloud: uniform-octree leaves, theta = 0.4 near-field list built on the host,
evaluated by vpm_p2p_leafpairs (all GPUs of the handle).  usage: c5_nearfield.py [log2N] [ncrit] [ngpu]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpm_import import load  # noqa: E402

vpm = load()
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 24
ncrit = int(sys.argv[2]) if len(sys.argv) > 2 else 512
ngpu = int(sys.argv[3]) if len(sys.argv) > 3 else 1
n = 1 << logn
h = vpm.Handle(ngpu)
t = time.perf_counter()
X, Gamma, sigma = vpm.fields.cloud_arrays(n)
ll = vpm.fields.build_leaf_lists(X, sigma, ncrit=ncrit, theta=0.4)
t_build = time.perf_counter() - t
order = ll["sort_index"]
sb = np.zeros((8, n), order="F")
sb[0:3], sb[4:7], sb[3], sb[7] = X[:, order], Gamma[:, order], sigma[order], sigma[order]
tb = np.zeros((16, n), order="F")
tb[0:3] = X[:, order]
del X, Gamma
leaves = (ll["leaf_begin"], ll["leaf_end"])
sizes = ll["leaf_end"] - ll["leaf_begin"]
dl = ll["direct_list"]
pairs = int((sizes[dl[:, 0]].astype(np.int64) * sizes[dl[:, 1]]).sum())
res = {"n": n, "ncrit": ncrit, "gpus": ngpu, "leaves": int(len(sizes)), "mean_leaf": float(sizes.mean()),
       "list_pairs": int(len(dl)), "interactions": pairs, "host_tree_s": t_build}
for rep in range(2):
    tb[4:] = 0
    t = time.perf_counter()
    vpm.nearfield_device(tb, leaves, sb, leaves, dl, vpm.winckelmans, handle=h)
    dt = time.perf_counter() - t
tm = h.timing()
res.update(call_s=dt, kernel_ms_dev0=tm["uj_ms"], h2d_ms=tm["h2d_ms"], d2h_ms=tm["d2h_ms"],
           e2e_interactions_per_s=pairs / dt, finite=bool(np.isfinite(tb[4:]).all()))
# parity on a slice: three target leaves recomputed by the CPU oracle (test infrastructure)
from oracle import oracle  # noqa: E402
worst = 0.0
for leaf in (0, len(sizes) // 2, len(sizes) - 1):
    sel = dl[dl[:, 0] == leaf]
    ref = np.zeros((16, n), order="F") if False else None
    b, e = int(ll["leaf_begin"][leaf]), int(ll["leaf_end"][leaf])
    loc = np.zeros((16, e - b), order="F")
    loc[0:3] = tb[0:3, b:e]
    for _, sl in sel:
        oracle.direct_buffers(loc, 0, e - b, sb, int(ll["leaf_begin"][sl]), int(ll["leaf_end"][sl]), "winckelmans")
    worst = max(worst, float(np.abs(loc[4:] - tb[4:, b:e]).max() / np.abs(loc[4:]).max()))
res["slice_parity_rel_err"] = worst
print(json.dumps(res))
