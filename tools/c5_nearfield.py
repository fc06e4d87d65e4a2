#!/usr/bin/env python
"""One-off measurement (GPU box): BASELINE.json configs[4] -- FMM near-field offload on a
2^24-particle cloud.  Two paths through the C ABI, both from host buffers:
  hook3 : lists handed in by the caller (as FastMultipole would) -> vpm_p2p_leafpairs
  f3    : vpm_leaflists_build (tree + theta = 0.4 list on the device) -> vpm_uj_nearfield
usage: c5_nearfield.py [log2N] [ncrit] [ngpu]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpm_import import load  # noqa: E402

vpm = load()
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 24
ncrits = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [512]
ngpu = int(sys.argv[3]) if len(sys.argv) > 3 else 1
n = 1 << logn
h = vpm.Handle(ngpu)
pf = vpm.fields.cloud_field(n, kernel=vpm.winckelmans)


def pin(a):
    h.check(h.lib.vpm_pin_host(h.ptr, a.ctypes.data, a.nbytes))


pin(pf.particles)  # as a Julia caller would page-lock pfield.particles once
sb = np.zeros((8, n), order="F")
tb = np.zeros((16, n), order="F")
pin(sb)
pin(tb)
for ncrit in ncrits:
  res = {"n": n, "ncrit": ncrit, "gpus": ngpu}
  # ---- f-3: tree and list on the device
  for rep in range(2):
      t = time.perf_counter()
      info = vpm.leaf_lists(pf, ncrit=ncrit, theta=0.4, handle=h, fetch=False)
      t_build = time.perf_counter() - t
  res.update(leaves=info["n_leaves"], list_pairs=info["n_pairs"], device_tree_s=t_build,
             device_tree_h2d_ms=h.timing()["h2d_ms"], device_tree_total_ms=h.timing()["total_ms"])
  for rep in range(2):
      t = time.perf_counter()
      vpm.UJ_nearfield(pf, reset=True, handle=h)
      dt = time.perf_counter() - t
  tm = h.timing()
  res.update(f3_call_s=dt, f3_kernel_ms_dev0=tm["uj_ms"], f3_h2d_ms=tm["h2d_ms"], f3_d2h_ms=tm["d2h_ms"],
             interactions=tm["uj_pairs"], f3_e2e_interactions_per_s=tm["uj_pairs"] / dt,
             f3_tree_plus_nearfield_interactions_per_s=tm["uj_pairs"] / (dt + t_build))
  U_f3 = pf.particles[9:12].copy()
  # ---- Hook 3: the same lists fetched to the host and handed back like FastMultipole's
  t = time.perf_counter()
  ll = vpm.leaf_lists(pf, ncrit=ncrit, theta=0.4, handle=h)
  res["device_tree_with_fetch_s"] = time.perf_counter() - t
  order = ll["sort_index"]
  sb[0:3], sb[4:7], sb[3], sb[7] = pf.get_X()[:, order], pf.get_Gamma()[:, order], pf.get_sigma()[order], pf.get_sigma()[order]
  tb[0:3] = sb[0:3]
  leaves = (ll["leaf_begin"], ll["leaf_end"])
  dl = ll["direct_list"]
  pairs = (ll["pair_tgt"], ll["pair_src"])   # contiguous columns: no de-interleaving copy in the wrapper
  for rep in range(2):
      tb[4:] = 0
      t = time.perf_counter()
      vpm.nearfield_device(tb, leaves, sb, leaves, pairs, vpm.winckelmans, handle=h)
      dt = time.perf_counter() - t
  tm = h.timing()
  res.update(hook3_call_s=dt, hook3_kernel_ms_dev0=tm["uj_ms"], hook3_h2d_ms=tm["h2d_ms"], hook3_d2h_ms=tm["d2h_ms"],
             hook3_e2e_interactions_per_s=tm["uj_pairs"] / dt, finite=bool(np.isfinite(tb[4:]).all()))
  res["f3_equals_hook3"] = bool(np.array_equal(U_f3[:, order], tb[4:7]))
  # ---- Hook 3 in the reference's call shape (vpm_nearfield_ranges): the tree-sorted SYSTEM (46 x N matrix,
  # page-locked) and per-target-leaf source ranges; the range tables are built vectorised here (a Julia
  # caller builds them from FastMultipole's branches), the call itself is the ctypes call of the ABI
  import ctypes as C
  S = np.asfortranarray(pf.particles[:, order])
  S[9:27] = 0.0
  pin(S)
  lb, le = ll["leaf_begin"], ll["leaf_end"]
  pt, ps = ll["pair_tgt"], ll["pair_src"]            # grouped by target leaf
  nl = len(lb)
  soff = np.searchsorted(pt, np.arange(nl + 1)).astype(np.int64)
  sbeg, send = np.ascontiguousarray(lb[ps]), np.ascontiguousarray(le[ps])
  for rep in range(2):
      S[9:27] = 0.0
      t = time.perf_counter()
      h.check(h.lib.vpm_nearfield_ranges(h.ptr, S.ctypes.data, S.shape[0], n, lb.ctypes.data, le.ctypes.data, nl,
                                         S.ctypes.data, S.shape[0], n, sbeg.ctypes.data, send.ctypes.data,
                                         soff.ctypes.data, vpm.winckelmans.id, 1, 1))
      dt = time.perf_counter() - t
  tm = h.timing()
  res.update(ranges_call_s=dt, ranges_kernel_ms_dev0=tm["uj_ms"], ranges_e2e_interactions_per_s=tm["uj_pairs"] / dt,
             ranges_equals_hook3=bool(np.array_equal(S[9:12], tb[4:7]) and np.array_equal(S[15:24], tb[7:16])))
  h.check(h.lib.vpm_unpin_host(h.ptr, S.ctypes.data))
  del S
  # parity on a slice: three target leaves recomputed by the CPU oracle (test infrastructure)
  from oracle import oracle  # noqa: E402
  worst = 0.0
  sizes = ll["leaf_end"] - ll["leaf_begin"]
  for leaf in (0, len(sizes) // 2, len(sizes) - 1):
      sel = dl[dl[:, 0] == leaf]
      b, e = int(ll["leaf_begin"][leaf]), int(ll["leaf_end"][leaf])
      loc = np.zeros((16, e - b), order="F")
      loc[0:3] = tb[0:3, b:e]
      for _, sl in sel:
          oracle.direct_buffers(loc, 0, e - b, sb, int(ll["leaf_begin"][sl]), int(ll["leaf_end"][sl]), "winckelmans")
      worst = max(worst, float(np.abs(loc[4:] - tb[4:, b:e]).max() / np.abs(loc[4:]).max()))
  res["slice_parity_rel_err"] = worst
  print(json.dumps(res), flush=True)
