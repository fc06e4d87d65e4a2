"""CPU-side checks: the C-ABI library loads and exports every symbol include/vpm_cuda.h
declares (no compute call -- there is no GPU here), the host mirror behaves like the
reference's container API, and the product never routes through the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import leaflists

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "vpm_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vpm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(vpm):
    from vpm_import import load_build
    lib_path = load_build().build()
    lib = ctypes.CDLL(lib_path)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/vpm_cuda.h but not exported"
    assert sorted(vpm._cabi.SYMBOLS) == syms
    assert lib.vpm_abi_version() == 1


def test_no_gpu_means_loud_failure_not_fallback(vpm):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(vpm.VpmError) as e:
        vpm.Handle(1)
    assert e.value.code == -4 and "no CPU fallback" in str(e.value)
    pf = vpm.fields.cloud_field(16)
    vpm.set_handle(None)
    with pytest.raises(vpm.VpmError):
        vpm.UJ_direct(pf)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "flowvpm.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "fields.py", f"{f} mentions the oracle"
    text = open(os.path.join(pkg, "fields.py")).read()
    assert "import oracle" not in text and "from oracle" not in text


def test_particlefield_container(vpm):
    pf = vpm.ParticleField(4, kernel=vpm.winckelmans)
    assert pf.particles.shape == (46, 4) and pf.particles.flags.f_contiguous
    pf.add_particle([1, 2, 3], [4, 5, 6], 0.5, vol=0.1, circulation=-2.0, static=True)
    pf.add_particle([0, 0, 0], [1, 0, 0], 0.2)
    assert pf.get_np() == 2
    col = pf.particles[:, 0]
    assert list(col[0:7]) == [1, 2, 3, 4, 5, 6, 0.5] and col[7] == 0.1 and col[8] == 2.0 and col[42] == 1.0
    # column stride equals Julia's Matrix{Float64}(46, n): 368 bytes per particle
    assert pf.particles.strides == (8, 368)
    pf.add_particle([9, 9, 9], [0, 0, 1], 0.3)
    pf.remove_particle(0)          # swap-remove
    assert pf.get_np() == 2 and list(pf.particles[0:3, 0]) == [9, 9, 9]
    pf.add_particle([0, 0, 0], [1, 0, 0], 0.2)
    pf.add_particle([0, 0, 0], [1, 0, 0], 0.2)
    with pytest.raises(RuntimeError):
        pf.add_particle([0, 0, 0], [1, 0, 0], 0.2)
    assert vpm.kernel_default is vpm.gaussianerf and pf.UJ is vpm.UJ_direct


def test_reset_container_functions(vpm):
    pf = vpm.fields.cloud_field(20, static_fraction=0.3)
    vpm.fields.random_results(pf)
    before = pf.particles.copy(order="F")
    st = pf.get_static()
    assert st.any() and (~st).any()
    vpm._reset_particles(pf)
    assert np.all(pf.particles[9:27][:, :20][:, ~st] == 0)
    assert np.array_equal(pf.particles[9:27][:, :20][:, st], before[9:27][:, :20][:, st])
    assert np.array_equal(pf.particles[39:42], before[39:42])
    vpm._reset_particles_sfs(pf)
    assert np.all(pf.particles[39:42][:, :20][:, ~st] == 0)


def test_field_generators(vpm):
    assert vpm.fields.number_particles(100, 3) == 4900 and vpm.fields.number_particles(100, 0) == 100
    pf = vpm.fields.ring_field(Nphi=100, nc=1)
    assert pf.np == 900
    # total vortex strength of a ring is tangential: sum Gamma = 0, circulation recovered
    assert np.abs(pf.get_Gamma().sum(axis=1)).max() < 1e-12
    G = np.linalg.norm(pf.get_Gamma(), axis=0).sum()
    assert abs(G / (2 * np.pi * 1.0) - 1.0) < 2e-2
    c = vpm.fields.cloud_field(1000)
    X = c.get_X()
    d = (7.0 / 1000) ** (1 / 3)
    assert X.min() > -0.3 * d and X[2].max() < 7.5
    # minimum separation of the jittered lattice: > 0.5 d
    from scipy.spatial import cKDTree
    dd, _ = cKDTree(X.T).query(X.T, k=2)
    assert dd[:, 1].min() > 0.45 * d
    assert abs(c.get_sigma().mean() / (0.65 * d) - 1) < 0.01


def test_leaf_lists_cover_all_near_pairs(vpm):
    c = vpm.fields.cloud_field(3000, seed=8)
    ll = leaflists.build_leaf_lists(c.get_X(), c.get_sigma(), ncrit=40, theta=0.4)
    b, e, dl = ll["leaf_begin"], ll["leaf_end"], ll["direct_list"]
    assert b[0] == 0 and e[-1] == 3000 and np.all(b[1:] == e[:-1])
    assert sorted(ll["sort_index"]) == list(range(3000))
    s = set(map(tuple, dl))
    assert all((i, i) in s for i in range(len(b)))       # self pairs are near field
    assert all((j, i) in s for (i, j) in list(s)[:2000])  # MAC is symmetric


def test_leaf_lists_refine_cells_for_sparse_fields(vpm):
    """a ring fills a small part of its bounding box: the restated builder shrinks the cells until
    the occupied ones hold about ncrit/2 bodies, and the lists stay complete and symmetric"""
    r = vpm.fields.ring_field(Nphi=100, nc=3)
    ll = leaflists.build_leaf_lists(r.get_X(), r.get_sigma(), ncrit=32, theta=0.4)
    sizes = ll["leaf_end"] - ll["leaf_begin"]
    assert sizes.sum() == r.np and 8 <= sizes.mean() <= 24
    s = set(map(tuple, ll["direct_list"]))
    assert all((i, i) in s for i in range(len(sizes))) and all((j, i) in s for (i, j) in s)
    # every pair of bodies closer than the smaller of their core sizes is covered by the list
    X, order = r.get_X(), ll["sort_index"]
    leaf_of = np.empty(r.np, dtype=np.int64)
    leaf_of[order] = np.repeat(np.arange(len(sizes)), sizes)
    from scipy.spatial import cKDTree
    for i, j in cKDTree(X.T).query_pairs(float(r.get_sigma().min())):
        assert (leaf_of[i], leaf_of[j]) in s


@pytest.mark.parametrize("field", ["cloud", "ring", "jet"])
@pytest.mark.parametrize("theta", [0.25, 0.4, 0.8])
def test_leaf_lists_equal_brute_force_mac(vpm, field, theta):
    """the stencil search of the builder (the recipe the device kernel restates) finds exactly the leaf pairs that
    fail the acceptance criterion (r_i + r_j) <= theta d: compared with all nl^2 pairs tested directly, with the
    leaf spheres recomputed independently from the sorted bodies"""
    pf = {"cloud": lambda: vpm.fields.cloud_field(2500, seed=3),
          "ring": lambda: vpm.fields.ring_field(Nphi=60, nc=2),
          "jet": lambda: vpm.fields.jet_field(2000)}[field]()
    n = pf.np
    X, sig = pf.get_X()[:, :n], pf.get_sigma()[:n]
    ll = leaflists.build_leaf_lists(X, sig, ncrit=24, theta=theta)
    b, e, order = ll["leaf_begin"], ll["leaf_end"], ll["sort_index"]
    nl = len(b)
    ctr, rad = np.empty((nl, 3)), np.empty(nl)
    for l in range(nl):
        xs = X[:, order[b[l]:e[l]]]
        c = 0.5 * (xs.min(axis=1) + xs.max(axis=1))
        ctr[l] = c
        rad[l] = np.sqrt(((xs - c[:, None]) ** 2).sum(axis=0).max()) + sig[order[b[l]:e[l]]].max()
    d = np.sqrt(((ctr[:, None, :] - ctr[None, :, :]) ** 2).sum(axis=2))
    near = (d == 0) | ((rad[:, None] + rad[None, :]) > theta * d)
    want = set(zip(*np.nonzero(near)))
    got = set(map(tuple, ll["direct_list"].tolist()))
    # pairs within rounding of the criterion may fall either way between the two evaluations
    margin = np.abs((rad[:, None] + rad[None, :]) - theta * d) <= 1e-12 * (rad[:, None] + rad[None, :])
    diff = (want ^ got)
    assert all(margin[i, j] for (i, j) in diff), (len(want), len(got), len(diff))
    assert len(got) > nl                                   # more than the self pairs


def test_direct_list_forms(vpm):
    """the list-taking wrappers accept the reference's (n, 2) direct_list or a tuple of its columns"""
    from flowvpm_jl_b200 import uj
    dl = np.array([[0, 1], [2, 3], [4, 5]], dtype=np.int64)
    a, b = uj._pairs(dl), uj._pairs((dl[:, 0], dl[:, 1]))
    for x, y in zip(a, b):
        assert x.dtype == np.int32 and x.flags.c_contiguous and np.array_equal(x, y)
    assert [v.tolist() for v in uj._pairs([[0, 1], [2, 3]])] == [[0, 2], [1, 3]]   # a list is an (n, 2) list
    cols = (np.arange(4, dtype=np.int32), np.arange(4, dtype=np.int32))
    assert uj._pairs(cols)[0] is cols[0]                                            # no copy
    with pytest.raises(ValueError):
        uj._pairs(np.zeros((3, 3)))
    with pytest.raises(ValueError):
        uj._pairs((np.arange(3), np.arange(4)))


def test_sharding_bounds(vpm):
    from flowvpm_jl_b200 import sharding
    for n, w in ((10, 4), (1 << 20, 8), (7, 8), (0, 2)):
        segs = [sharding.shard_bounds(n, w, r) for r in range(w)]
        assert segs[0][0] == 0 and segs[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(segs, segs[1:]))
        assert all(t1 - t0 <= sharding.shard_size(n, w) for t0, t1 in segs)


def test_save_xdmf_binary_roundtrip(vpm, tmp_path):
    """f-4: `save` (src/FLOWVPM_utils.jl:148-371) as XDMF + raw binary: names, shapes and values survive"""
    import numpy as np
    pf = vpm.fields.cloud_field(257, static_fraction=0.2, seed=3)
    vpm.fields.random_results(pf)
    pf.nt, pf.t = 7, 0.125
    ret = vpm.save(pf, "pfield", path=str(tmp_path))
    assert ret == "pfield.7.xmf;"
    got = vpm.io.read(str(tmp_path / "pfield.7.xmf"))
    n, P = pf.np, pf.particles
    assert (got["np"], got["nt"], got["t"]) == (n, 7, 0.125)
    assert np.array_equal(got["X"], P[0:3, :n].T) and np.array_equal(got["Gamma"], P[3:6, :n].T)
    assert np.array_equal(got["sigma"], P[6, :n]) and np.array_equal(got["static"], P[42, :n])
    assert np.array_equal(got["velocity"], P[9:12, :n].T) and np.array_equal(got["vorticity"], P[12:15, :n].T)
    assert np.array_equal(got["velocity_gradient_x"], P[15:18, :n].T)
    assert np.array_equal(got["velocity_gradient_z"], P[21:24, :n].T) and np.array_equal(got["C"], P[36:39, :n].T)
    # empty field -> one dummy particle, explicit number, no number
    empty = vpm.ParticleField(4)
    assert vpm.save(empty, "e", path=str(tmp_path), num=3) == "e.3.xmf;"
    assert vpm.io.read(str(tmp_path / "e.3.xmf"))["np"] == 1
    assert vpm.save(pf, "plain", path=str(tmp_path / "sub"), add_num=False, createpath=True) == "plain.xmf;"


# ---- launch plans (host arithmetic of the library, vpm_plan_query: no GPU needed) -------------------------------
def _plan(vpm, nt, ns, sm=148, kind=0):
    out = (ctypes.c_int64 * 8)()
    assert vpm._cabi.load().vpm_plan_query(int(nt), int(ns), int(sm), int(kind), out) == 0
    keys = ("targets_per_cta", "grid_x", "nsplit", "src_per_split", "tiles_per_split", "T", "unroll", "fills")
    return dict(zip(keys, list(out)))


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_launch_plans_cover_every_pair_exactly_once(vpm, kind):
    """every target belongs to one CTA column and every source to one non-empty split, for ragged sizes from one
    particle to 2^24, on 1 ... 148 SMs; the FP32 and table kernels split on tile boundaries, the FP64 kernels may
    split finer (small fields) but never below 16 sources unless the field is smaller than that"""
    rng = np.random.Generator(np.random.PCG64(11))
    sizes = [1, 2, 15, 16, 17, 31, 127, 128, 129, 200, 255, 257, 900, 1023, 1025, 4900, 8191, 8193, 33800, 100_000,
             300_000, 1 << 20, (1 << 22) + 5, 1 << 24] + [int(v) for v in rng.integers(1, 60_000, 40)]
    for sm in (1, 8, 132, 148):
        for nt in sizes:
            for ns in (sizes if nt in (200, 4900, 1 << 20) else [nt, max(1, nt // 3), 4900]):
                p = _plan(vpm, nt, ns, sm, kind)
                tpc = p["targets_per_cta"]
                assert p["grid_x"] * tpc >= nt > (p["grid_x"] - 1) * tpc, (nt, ns, sm, p)
                sps = p["src_per_split"]
                assert 1 <= p["nsplit"] <= 1024 and p["nsplit"] * sps >= ns > (p["nsplit"] - 1) * sps, (nt, ns, sm, p)
                if kind >= 2 or sps >= 128:
                    assert sps == p["tiles_per_split"] * 128, (nt, ns, sm, p)
                else:
                    assert p["tiles_per_split"] == 1 and (sps >= 16 or p["nsplit"] == 1), (nt, ns, sm, p)
                assert p["T"] in (1, 2) and p["unroll"] in (1, 2)
                assert p == _plan(vpm, nt, ns, sm, kind)          # a function of its arguments only


def test_small_field_plans_fill_one_wave_and_even_out_few_waves(vpm):
    """what the sub-tile splits are for: 200 and 900 particles get tens of CTAs instead of 4 and
    64, never more than 32 splits below one wave; 4 900 particles get 1.98 waves of 888 CTAs instead of 1.71; large
    fields keep tile-aligned splits of several tiles"""
    p = _plan(vpm, 200, 200)
    assert (p["grid_x"], p["nsplit"], p["src_per_split"]) == (2, 13, 16)
    p = _plan(vpm, 900, 900)
    assert p["grid_x"] == 8 and p["nsplit"] <= 32 and 16 <= p["src_per_split"] < 128
    p = _plan(vpm, 4900, 4900)
    ctas, wave = p["grid_x"] * p["nsplit"], 148 * 6
    assert 1.9 * wave <= ctas <= 2.0 * wave and p["src_per_split"] < 128
    p = _plan(vpm, 1 << 20, 1 << 20)
    assert p["T"] == 2 and p["tiles_per_split"] >= 4 and p["src_per_split"] == 128 * p["tiles_per_split"]
    assert _plan(vpm, 1 << 20, 1 << 20, kind=3)["fills"] == 1 and _plan(vpm, 4900, 4900, kind=3)["fills"] == 0
    lib = vpm._cabi.load()
    assert lib.vpm_plan_query(10, 10, 0, 0, (ctypes.c_int64 * 8)()) == -1          # VPM_EINVAL
    assert lib.vpm_plan_query(10, 10, 148, 7, (ctypes.c_int64 * 8)()) == -1


def test_header_is_plain_c_and_links_without_a_gpu(vpm, tmp_path):
    """include/vpm_cuda.h is the drop-in boundary: it must compile on its own as C11 and as C++ (no torch, no CUDA
    headers), and a plain-C program must link against libvpm_cuda.so and call its GPU-free entry points
    (vpm_abi_version, vpm_plan_query); vpm_create without a GPU returns VPM_ENODEV instead of aborting"""
    import subprocess
    from vpm_import import load_build
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_path = load_build().build()
    inc = os.path.join(root, "include")
    for compiler, std, suffix in (("/usr/bin/gcc", "-std=c11", ".c"), ("/usr/bin/g++", "-std=c++17", ".cpp")):
        src = tmp_path / ("hdr" + suffix)
        src.write_text('#include "vpm_cuda.h"\nint main(void) { return 0; }\n')
        subprocess.run([compiler, std, "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", inc, str(src)], check=True)
    prog = tmp_path / "plan.c"
    prog.write_text(r'''
#include <stdio.h>
#include "vpm_cuda.h"
int main(void) {
  int64_t out[8];
  if (vpm_abi_version() != 1) return 2;
  if (vpm_plan_query(4900, 4900, 148, 0, out) != VPM_OK) return 3;
  printf("%lld %lld %lld\n", (long long)out[1], (long long)out[2], (long long)out[3]);
  vpm_handle *h = 0;
  int rc = vpm_create(&h, 1, 0);
  if (rc == VPM_OK) { vpm_destroy(h); printf("gpu\n"); } else printf("rc %d\n", rc);
  return 0;
}
''')
    exe = str(tmp_path / "plan")
    libdir = os.path.dirname(lib_path)
    subprocess.run(["/usr/bin/gcc", "-std=c11", "-O1", "-I", inc, str(prog), "-L", libdir, "-lvpm_cuda",
                    "-Wl,-rpath," + libdir, "-o", exe], check=True)
    res = subprocess.run([exe], capture_output=True, text=True, check=True)
    lines = res.stdout.split("\n")
    assert lines[0].split() == ["39", "45", "109"]
    import torch
    assert lines[1] == ("gpu" if torch.cuda.is_available() else "rc -4")
