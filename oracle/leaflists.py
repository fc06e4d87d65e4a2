"""CPU restatement (numpy) of the device-side leaf-list builder (csrc/vpm_tree.cuh).

TEST INFRASTRUCTURE ONLY.  FastMultipole.jl (the tree of UJ_fmm, src/FLOWVPM_UJ.jl:90-101) is an
un-vendored dependency of the reference (uuid ce07d0d3-..., compat "2", no Manifest), so there
is no reference tree arithmetic to pin against: "parity unpinned" for the tree itself.  What
is pinned is (i) bit-exact equality of the device lists with this restatement (integer / index
work) and (ii) that ANY valid list evaluated by the near-field kernels equals the oracle's
fmm.direct! arithmetic over the same list (tests/test_tree_gpu.py).

Recipe: uniform cell grid with mean occupancy ~ncrit/2 (thin directions padded to 1e-3 of the
largest extent; cells shrunk up to twice when the field does not fill its bounding box), stable
sort by cell key, one leaf per occupied cell, leaf sphere = middle of
the leaf's bounding box + largest distance + largest sigma, near-field list = leaf pairs failing
the MAC (r_i + r_j) <= theta * d (theta = 0.4: src/FLOWVPM_particlefield.jl:28-36), emitted in
(target, source) lexicographic order."""
import math

import numpy as np


def _dims(ext, h):
    c = np.maximum(1.0, np.ceil(ext / h))
    return c.astype(np.int64), float(c[0] * c[1] * c[2])


def grid_and_sort(X, ncrit):
    """cell size, grid dimensions and the stable sort by cell key; the cells shrink (at most twice)
    while the occupied ones hold more than 1.5 x the target ncrit/2 bodies on average"""
    N = X.shape[1]
    lo, hi = X.min(axis=1), X.max(axis=1)
    ext = hi - lo
    emax = float(ext.max())
    ext = np.maximum(np.maximum(ext, 1e-3 * emax), 1e-300)
    vol = float(ext[0] * ext[1] * ext[2])
    h = math.pow(vol * (ncrit / 2.0) / N, 1.0 / 3.0)
    if not (h > 0.0 and math.isfinite(h)):
        h = 1.0
    dims, _ = _dims(ext, h)
    it = 0
    while True:
        cell = np.minimum(((X - lo[:, None]) / h).astype(np.int64), (dims - 1)[:, None])
        key = (cell[0] * dims[1] + cell[1]) * dims[2] + cell[2]
        order = np.argsort(key, kind="stable")
        skey = key[order]
        nl = 1 + int(np.count_nonzero(skey[1:] != skey[:-1]))
        occ = N / nl
        if it >= 2 or occ <= 0.75 * ncrit:
            break
        h2 = h * math.pow((ncrit / 2.0) / occ, 1.0 / 3.0)
        dims2, ncell2 = _dims(ext, h2)
        if not (h2 > 0.0) or ncell2 > 1.0e9 or ncell2 > 64.0 * N + 4096.0:
            break
        h, dims = h2, dims2
        it += 1
    return lo, h, dims, order, skey


def build_leaf_lists(X, sigma, ncrit=64, theta=0.4):
    """Returns dict(sort_index, leaf_begin, leaf_end, direct_list)."""
    X = np.asarray(X, dtype=np.float64)
    sigma = np.asarray(sigma, dtype=np.float64)
    N = X.shape[1]
    lo, h, dims, order, skey = grid_and_sort(X, ncrit)
    uniq, begin = np.unique(skey, return_index=True)
    end = np.append(begin[1:], N)
    nl = len(uniq)
    Xs = X[:, order]
    ssig = sigma[order]
    # leaf spheres (reduceat keeps every operation an exactly-rounded min / max / mul / add)
    mn = np.minimum.reduceat(Xs, begin, axis=1)
    mx = np.maximum.reduceat(Xs, begin, axis=1)
    centers = 0.5 * (mn + mx)
    leaf_of = np.repeat(np.arange(nl), end - begin)
    d = Xs - centers[:, leaf_of]
    d2 = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]
    radii = np.sqrt(np.maximum.reduceat(d2, begin)) + np.maximum.reduceat(ssig, begin)
    cz = uniq % dims[2]
    cy = (uniq // dims[2]) % dims[1]
    cx = uniq // (dims[1] * dims[2])
    ncell = int(np.prod(dims))
    cell_to_leaf = np.full(ncell, -1, dtype=np.int64)
    cell_to_leaf[uniq] = np.arange(nl)
    reach = int(min(math.ceil(2.0 * float(radii.max()) / (theta * h)) + 1.0, 1.0e6))
    reach = min(reach, int(dims.max()))
    leaves = np.arange(nl)
    pairs = []
    for ddx in range(-reach, reach + 1):
        x = cx + ddx
        okx = (x >= 0) & (x < dims[0])
        if not okx.any():
            continue
        for ddy in range(-reach, reach + 1):
            y = cy + ddy
            okxy = okx & (y >= 0) & (y < dims[1])
            if not okxy.any():
                continue
            for ddz in range(-reach, reach + 1):
                z = cz + ddz
                ok = okxy & (z >= 0) & (z < dims[2])
                if not ok.any():
                    continue
                l = leaves[ok]
                m = cell_to_leaf[(x[ok] * dims[1] + y[ok]) * dims[2] + z[ok]]
                has = m >= 0
                l, m = l[has], m[has]
                dc = centers[:, l] - centers[:, m]
                dist = np.sqrt((dc[0] * dc[0] + dc[1] * dc[1]) + dc[2] * dc[2])
                near = (dist == 0) | ((radii[l] + radii[m]) > theta * dist)
                pairs.append(np.stack([l[near], m[near]], axis=1))
    pairs = np.concatenate(pairs) if pairs else np.zeros((0, 2), dtype=np.int64)
    pairs = pairs[np.lexsort((pairs[:, 1], pairs[:, 0]))]
    direct_list = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
    return dict(sort_index=order.astype(np.int64), leaf_begin=begin.astype(np.int64),
                leaf_end=end.astype(np.int64), direct_list=direct_list)
