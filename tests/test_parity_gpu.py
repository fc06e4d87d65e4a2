"""GPU parity: libvpm_cuda (through the C ABI / host mirror) against the CPU oracle on the
same seeded inputs, and against the committed mpmath golden vectors.

Bar (BASELINE.json north_star): relative error <= 1e-12 in FP64 for U, J, stretching and
SFS, measured norm-wise per field (helpers.relerr); <= 1e-5 for the Float32 entry point.
"""
import os

import numpy as np
import pytest

from helpers import KERNELS, TOL_FP64, TOL_FP32, assert_parity, relerr, stretching
from oracle import leaflists, oracle

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "p2p_golden.npz"))


def oracle_uj(pf, **kw):
    ref = pf.particles.copy(order="F")
    oracle.uj_direct(ref, pf.np, pf.kernel.name, transposed=pf.transposed, **kw)
    return ref


# ------------------------------------------------------------ golden vectors
@pytest.mark.parametrize("kernel", KERNELS)
def test_uj_vs_mpmath_golden(vpm, handle, kernel):
    n = GOLD["X"].shape[1]
    pf = vpm.ParticleField(n, kernel=vpm.KERNELS[kernel])
    pf.particles[0:3] = GOLD["X"]
    pf.particles[3:6] = GOLD["Gamma"]
    pf.particles[6] = GOLD["sigma"]
    pf.np = n
    vpm.UJ_direct(pf)
    assert relerr(pf.get_U(), GOLD[f"U_{kernel}"]) < TOL_FP64
    assert relerr(pf.get_J(), GOLD[f"J_{kernel}"]) < TOL_FP64


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("transposed", [True, False])
def test_sfs_vs_mpmath_golden(vpm, handle, kernel, transposed):
    """Estr over a given J field: the device API takes J as an input, like the golden set-up"""
    import torch
    n = GOLD["X"].shape[1]
    src8 = np.zeros((8, n), order="F")
    src8[0:3], src8[4:7], src8[7] = GOLD["X"], GOLD["Gamma"], GOLD["sigma"]
    src8[3] = GOLD["sigma"]
    d_src = torch.from_numpy(np.ascontiguousarray(src8.T)).cuda()
    d_J = torch.from_numpy(np.ascontiguousarray(GOLD["Jin"].T)).cuda()
    d_stat = torch.from_numpy(GOLD["static"].copy()).cuda()
    d_out = torch.zeros((n, 3), dtype=torch.float64, device="cuda")
    flags = vpm._cabi.FLAG_SFS | (vpm._cabi.FLAG_TRANSPOSED if transposed else 0)
    torch.cuda.synchronize()
    handle.check(handle.lib.vpm_sfs_device(handle.ptr, d_src.data_ptr(), d_J.data_ptr(), d_stat.data_ptr(), n, 0, n,
                                           d_out.data_ptr(), vpm.KERNELS[kernel].id, flags,
                                           torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    out = d_out.cpu().numpy().T
    ref = GOLD[f"SFS_{kernel}_{'T' if transposed else 'C'}"]
    st = GOLD["static"] != 0
    if np.abs(ref).max() > 0:
        assert relerr(out[:, ~st], ref[:, ~st]) < TOL_FP64
    else:
        assert np.abs(out[:, ~st]).max() == 0


# ---------------------------------------------------- UJ_direct vs the oracle
@pytest.mark.parametrize("kernel", KERNELS)
def test_uj_direct_ring_c1(vpm, handle, kernel):
    """config C1: isolated vortex ring, Nphi=100, nc=3 -> 4900 particles"""
    pf = vpm.fields.ring_field(Nphi=100, nc=3, kernel=vpm.KERNELS[kernel])
    ref = oracle_uj(pf, sfs=True, reset=True, reset_sfs=True)
    vpm.UJ_direct(pf, sfs=True, reset=True, reset_sfs=True)
    errs = assert_parity(pf.particles, ref, pf.np, what=f"ring/{kernel}")
    assert relerr(stretching(pf.particles, pf.np), stretching(ref, pf.np)) < TOL_FP64
    print(kernel, errs)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("n", [1, 2, 31, 129, 1000, 3001])
def test_uj_direct_cloud_sizes(vpm, handle, kernel, n):
    """ragged sizes around the tile (128) and CTA boundaries, down to a single particle"""
    pf = vpm.fields.cloud_field(n, kernel=vpm.KERNELS[kernel], seed=100 + n)
    ref = oracle_uj(pf, sfs=True, reset=True, reset_sfs=True)
    vpm.UJ_direct(pf, sfs=True, reset=True, reset_sfs=True)
    assert_parity(pf.particles, ref, n, what=f"cloud{n}/{kernel}")


@pytest.mark.parametrize("kernel", ["gaussianerf", "gaussian"])
def test_farfield_shortcut_matches_full_evaluation(vpm, handle, kernel):
    """g == 1 beyond the cutoff: with and without the shortcut agree to 1e-14, both in parity"""
    pf = vpm.fields.cloud_field(6000, kernel=vpm.KERNELS[kernel], seed=5)
    ref = oracle_uj(pf, sfs=True, reset=True, reset_sfs=True)
    a = pf.particles.copy(order="F")
    vpm.UJ_direct(pf, sfs=True, reset=True, reset_sfs=True)
    fast = pf.particles.copy(order="F")
    pf.particles[:] = a
    vpm.UJ_direct(pf, sfs=True, reset=True, reset_sfs=True, no_farfield_shortcut=True)
    assert_parity(fast, ref, pf.np)
    assert_parity(pf.particles, ref, pf.np)
    assert_parity(fast, pf.particles, pf.np, tol=1e-14)


@pytest.mark.parametrize("kernel", ["winckelmans", "gaussianerf"])
@pytest.mark.parametrize("reset,reset_sfs,sfs", [(True, False, False), (False, False, True), (True, True, False),
                                                 (False, True, True), (False, False, False)])
def test_reset_static_accumulate_rules(vpm, handle, kernel, reset, reset_sfs, sfs):
    """static particles are targets of U/J but are never reset and take no part in SFS
    (src/FLOWVPM_particlefield.jl:464-511, src/FLOWVPM_subfilterscale_models.jl:63)"""
    pf = vpm.fields.cloud_field(1500, kernel=vpm.KERNELS[kernel], static_fraction=0.15, seed=9)
    vpm.fields.random_results(pf, scale=1e-3)
    ref = oracle_uj(pf, sfs=sfs, reset=reset, reset_sfs=reset_sfs)
    vpm.UJ_direct(pf, sfs=sfs, reset=reset, reset_sfs=reset_sfs)
    assert_parity(pf.particles, ref, pf.np, rows=("U", "J", "SFS", "W", "PSE"))
    # rows the path must not touch
    for rows in (slice(0, 9), slice(27, 39), slice(42, 46)):
        assert np.array_equal(pf.particles[rows], ref[rows])


def test_classic_scheme_sfs(vpm, handle):
    pf = vpm.fields.ring_field(Nphi=40, nc=2, kernel=vpm.winckelmans, transposed=False)
    ref = oracle_uj(pf, sfs=True, reset=True, reset_sfs=True)
    vpm.UJ_direct(pf, sfs=True, reset=True, reset_sfs=True)
    assert_parity(pf.particles, ref, pf.np)


def test_two_rings_c2_slice(vpm, handle):
    """config C2 geometry (two coaxial rings, R=0.7906, Rcross=0.1R, dZ=0.7906), nc=3 here"""
    R = 0.7906
    pf = vpm.fields.ring_field(Nphi=100, nc=3, R=R, Rcross=0.1 * R, rings=2, dZ=0.7906, kernel=vpm.gaussianerf)
    ref = oracle_uj(pf, sfs=True, reset=True, reset_sfs=True)
    vpm.UJ_direct(pf, sfs=True, reset=True, reset_sfs=True)
    assert_parity(pf.particles, ref, pf.np)


def test_jet_c3_with_static_inflow(vpm, handle):
    pf = vpm.fields.jet_field(n_target=6000, kernel=vpm.gaussianerf)
    assert pf.get_static().sum() > 0
    ref = oracle_uj(pf, sfs=True, reset=True, reset_sfs=True)
    vpm.UJ_direct(pf, sfs=True, reset=True, reset_sfs=True)
    assert_parity(pf.particles, ref, pf.np)


def test_empty_and_coincident(vpm, handle):
    pf = vpm.ParticleField(8, kernel=vpm.winckelmans)
    vpm.UJ_direct(pf, sfs=True)  # np == 0: nothing to do, no error
    assert not pf.particles.any()
    for k in KERNELS:
        pf = vpm.ParticleField(3, kernel=vpm.KERNELS[k])
        pf.add_particle([0.1, 0.2, 0.3], [0.1, 0.2, 0.3], 0.2)
        pf.add_particle([0.1, 0.2, 0.3], [0.0, -0.1, 0.2], 0.2)   # coincident: r2 == 0 -> skipped
        pf.add_particle([0.5, 0.2, 0.3], [0.3, 0.0, 0.1], 0.2)
        ref = oracle_uj(pf, sfs=True, reset=True, reset_sfs=True)
        vpm.UJ_direct(pf, sfs=True, reset=True, reset_sfs=True)
        assert np.all(np.isfinite(pf.particles))
        assert_parity(pf.particles, ref, 3, what=f"coincident/{k}")


def test_run_to_run_bit_identical(vpm, handle):
    pf = vpm.fields.cloud_field(5000, kernel=vpm.winckelmans)
    vpm.UJ_direct(pf, sfs=True, reset=True, reset_sfs=True)
    a = pf.particles.copy(order="F")
    vpm.UJ_direct(pf, sfs=True, reset=True, reset_sfs=True)
    assert np.array_equal(a, pf.particles)


def test_uj_direct_source_target(vpm, handle):
    """UJ_direct(source, target): probes accumulate U, J from another field"""
    src = vpm.fields.cloud_field(2000, kernel=vpm.gaussianerf, seed=1)
    tgt = vpm.fields.cloud_field(333, kernel=vpm.gaussianerf, seed=2)
    vpm.fields.random_results(tgt, scale=1e-2)
    sb = vpm.source_system_to_buffer(src)
    tb = np.zeros((16, tgt.np), order="F")
    tb[0:3] = tgt.get_X()
    oracle.direct_buffers(tb, 0, tgt.np, sb, 0, src.np, "gaussianerf")
    ref = tgt.particles.copy(order="F")
    ref[9:12] += tb[4:7]
    ref[15:24] += tb[7:16]
    vpm.UJ_direct(src, tgt)
    assert_parity(tgt.particles, ref, tgt.np, rows=("U", "J"))


@pytest.mark.parametrize("kernel", KERNELS)
def test_staged_api_device_residency(vpm, handle, kernel):
    pf = vpm.fields.cloud_field(2500, kernel=vpm.KERNELS[kernel], static_fraction=0.1)
    vpm.fields.random_results(pf, scale=1e-3)
    ref = oracle_uj(pf, sfs=True, reset=True, reset_sfs=True)
    P = pf.particles
    lib = handle.lib
    handle.check(lib.vpm_upload_state(handle.ptr, P.ctypes.data, P.shape[0], pf.np))
    flags = 1 | 2 | 4 | 8
    handle.check(lib.vpm_eval(handle.ptr, pf.kernel.id, flags))
    handle.check(lib.vpm_download_results(handle.ptr, P.ctypes.data, P.shape[0], pf.np, flags))
    assert_parity(P, ref, pf.np)
    # second evaluation on the resident state: static particles accumulate again
    oracle.uj_direct(ref, pf.np, kernel, sfs=True, reset=True, reset_sfs=True)
    handle.check(lib.vpm_eval(handle.ptr, pf.kernel.id, flags))
    handle.check(lib.vpm_download_results(handle.ptr, P.ctypes.data, P.shape[0], pf.np, flags))
    assert_parity(P, ref, pf.np)


def test_pinned_host_matrix(vpm, handle):
    pf = vpm.fields.cloud_field(4000, kernel=vpm.winckelmans)
    ref = oracle_uj(pf)
    P = pf.particles
    handle.check(handle.lib.vpm_pin_host(handle.ptr, P.ctypes.data, P.nbytes))
    try:
        vpm.UJ_direct(pf)
    finally:
        handle.check(handle.lib.vpm_unpin_host(handle.ptr, P.ctypes.data))
    assert_parity(P, ref, pf.np, rows=("U", "J"))


# ------------------------------------------------- Float32 entry point (1e-5)
@pytest.mark.parametrize("kernel", KERNELS)
def test_uj_direct_f32(vpm, handle, kernel):
    pf64 = vpm.fields.cloud_field(3000, kernel=vpm.KERNELS[kernel], static_fraction=0.1, seed=4)
    pf32 = vpm.fields.cloud_field(3000, kernel=vpm.KERNELS[kernel], static_fraction=0.1, seed=4, R=np.float32)
    # the oracle sees the same (float32-rounded) inputs
    pf64.particles[:] = pf32.particles.astype(np.float64)
    ref = oracle_uj(pf64, sfs=True, reset=True, reset_sfs=True)
    vpm.UJ_direct(pf32, sfs=True, reset=True, reset_sfs=True)
    assert pf32.particles.dtype == np.float32
    assert_parity(pf32.particles, ref, pf32.np, tol=TOL_FP32)


@pytest.mark.parametrize("kernel", KERNELS)
def test_uj_direct_f32_vs_float32_field_semantics(vpm, handle, kernel):
    """the Matrix{Float32} entry point against the reference's OWN Float32 behaviour (Julia promotion rules
    restated in oracle.direct_buffers_f32): both are within 1e-5 of each other on a compact field"""
    pf32 = vpm.fields.cloud_field(2000, kernel=vpm.KERNELS[kernel], seed=6, R=np.float32)
    pf32.particles[2] /= 7.0   # unit cube: positions O(1), so the reference's Float32 dx keeps ~1e-6
    src32 = np.asfortranarray(pf32.particles[[0, 1, 2, 6, 3, 4, 5, 6]][:, :pf32.np])
    tb32 = np.zeros((16, pf32.np), dtype=np.float32, order="F")
    tb32[0:3] = src32[0:3]
    oracle.direct_buffers_f32(tb32, 0, pf32.np, src32, 0, pf32.np, kernel)
    vpm.UJ_direct(pf32, reset=True)
    assert relerr(pf32.get_U(), tb32[4:7]) < TOL_FP32 and relerr(pf32.get_J(), tb32[7:16]) < TOL_FP32


# ------------------------------------------------ Hook 2: fmm.direct! buffers
@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("want_U,want_J", [(True, True), (True, False), (False, True)])
def test_direct_buffers_ranges(vpm, handle, kernel, want_U, want_J):
    src = vpm.fields.cloud_field(1200, kernel=vpm.KERNELS[kernel], seed=11)
    sb = vpm.source_system_to_buffer(src)
    rng = np.random.default_rng(3)
    tb = np.asfortranarray(rng.standard_normal((16, 700)) * 1e-3)
    tb[0:3] = src.get_X()[:, :700] + 0.003
    ref = tb.copy(order="F")
    oracle.direct_buffers(ref, 100, 650, sb, 37, 1111, kernel, want_U, want_J)
    vpm.direct_buffers(tb, (100, 650), sb, (37, 1111), vpm.KERNELS[kernel], want_U=want_U, want_J=want_J)
    assert relerr(tb[4:7], ref[4:7]) < TOL_FP64 and relerr(tb[7:16], ref[7:16]) < TOL_FP64
    assert np.array_equal(tb[:, :100], ref[:, :100]) and np.array_equal(tb[:, 650:], ref[:, 650:])
    assert np.array_equal(tb[0:4], ref[0:4])


# ---------------------------------------- Hook 3: FMM near field over leaf pairs
@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("ncrit", [16, 200])
def test_nearfield_leafpairs(vpm, handle, kernel, ncrit):
    pf = vpm.fields.cloud_field(5000, kernel=vpm.KERNELS[kernel], seed=21)
    ll = leaflists.build_leaf_lists(pf.get_X(), pf.get_sigma(), ncrit=ncrit, theta=0.4)
    order = ll["sort_index"]
    sb = np.asfortranarray(vpm.source_system_to_buffer(pf)[:, order])
    tb = np.zeros((16, pf.np), order="F")
    tb[0:3] = pf.get_X()[:, order]
    leaves = (ll["leaf_begin"], ll["leaf_end"])
    ref = tb.copy(order="F")
    oracle.direct_leafpairs(ref, sb, leaves, leaves, ll["direct_list"], kernel)
    vpm.nearfield_device(tb, leaves, sb, leaves, ll["direct_list"], vpm.KERNELS[kernel])
    assert relerr(tb[4:7], ref[4:7]) < TOL_FP64 and relerr(tb[7:16], ref[7:16]) < TOL_FP64
    # SFS over the same list (Estr_fmm!): J taken from the near-field result
    pf.particles[15:24, order] = ref[7:16]
    pf.particles[42, ::7] = 1.0  # static flags are NOT filtered in the list form
    refP = pf.particles.copy(order="F")
    oracle.estr_leafpairs(refP, order, order, leaves, leaves, ll["direct_list"], kernel, True)
    vpm.Estr_fmm(pf, order, order, leaves, leaves, ll["direct_list"])
    assert relerr(pf.particles[39:42], refP[39:42]) < TOL_FP64


def test_leafpairs_unsorted_list_and_empty_leaves(vpm, handle):
    pf = vpm.fields.cloud_field(900, kernel=vpm.winckelmans, seed=2)
    ll = leaflists.build_leaf_lists(pf.get_X(), pf.get_sigma(), ncrit=32)
    order = ll["sort_index"]
    rng = np.random.default_rng(0)
    dl = ll["direct_list"][rng.permutation(len(ll["direct_list"]))]
    # add an empty leaf referenced by the list
    lb = np.append(ll["leaf_begin"], 900)
    le = np.append(ll["leaf_end"], 900)
    dl = np.vstack([dl, [[len(lb) - 1, 0], [0, len(lb) - 1]]]).astype(np.int32)
    sb = np.asfortranarray(vpm.source_system_to_buffer(pf)[:, order])
    tb = np.zeros((16, pf.np), order="F")
    tb[0:3] = pf.get_X()[:, order]
    ref = tb.copy(order="F")
    oracle.direct_leafpairs(ref, sb, (lb, le), (lb, le), dl, "winckelmans")
    vpm.nearfield_device(tb, (lb, le), sb, (lb, le), dl, vpm.winckelmans)
    assert relerr(tb[4:7], ref[4:7]) < TOL_FP64 and relerr(tb[7:16], ref[7:16]) < TOL_FP64


def branch_table_like_fastmultipole(ll, n, seed=0):
    """a leaf table as a level-ordered tree would hand it over: interior branches first (root and
    a few unions of leaves, never referenced by the list), then the leaves in shuffled order"""
    rng = np.random.default_rng(seed)
    nl = len(ll["leaf_begin"])
    perm = rng.permutation(nl)
    interior_b = [0] + [int(ll["leaf_begin"][k]) for k in range(0, nl, 7)]
    interior_e = [n] + [int(ll["leaf_end"][min(k + 6, nl - 1)]) for k in range(0, nl, 7)]
    ni = len(interior_b)
    lb = np.array(interior_b + ll["leaf_begin"][perm].tolist(), dtype=np.int64)
    le = np.array(interior_e + ll["leaf_end"][perm].tolist(), dtype=np.int64)
    new_id = np.empty(nl, dtype=np.int64)
    new_id[perm] = ni + np.arange(nl)
    dl = new_id[ll["direct_list"]].astype(np.int32)
    return lb, le, dl[rng.permutation(len(dl))]


def test_leafpairs_level_ordered_branch_table(vpm, handle):
    """leaf table not in body order and with interior branches: same result as the oracle"""
    pf = vpm.fields.cloud_field(3000, kernel=vpm.winckelmans, seed=14)
    ll = leaflists.build_leaf_lists(pf.get_X(), pf.get_sigma(), ncrit=20)
    lb, le, dl = branch_table_like_fastmultipole(ll, pf.np)
    order = ll["sort_index"]
    sb = np.asfortranarray(vpm.source_system_to_buffer(pf)[:, order])
    tb = np.zeros((16, pf.np), order="F")
    tb[0:3] = pf.get_X()[:, order]
    ref = tb.copy(order="F")
    oracle.direct_leafpairs(ref, sb, (lb, le), (lb, le), dl, "winckelmans")
    vpm.nearfield_device(tb, (lb, le), sb, (lb, le), dl, vpm.winckelmans)
    assert relerr(tb[4:7], ref[4:7]) < TOL_FP64 and relerr(tb[7:16], ref[7:16]) < TOL_FP64


# ------------------------- second P2P: zeta_direct / zeta_fmm (vorticity basis sum)
@pytest.mark.parametrize("kernel", KERNELS)
def test_zeta_direct(vpm, handle, kernel):
    pf = vpm.fields.cloud_field(3333, kernel=vpm.KERNELS[kernel], static_fraction=0.1, seed=12)
    vpm.fields.random_results(pf)
    ref = pf.particles.copy(order="F")
    oracle.zeta_direct(ref, pf.np, kernel)
    vpm.zeta_direct(pf)
    assert relerr(pf.particles[15:18], ref[15:18]) < TOL_FP64
    keep = np.r_[0:15, 18:46]
    assert np.array_equal(pf.particles[keep], ref[keep])   # only J[1:3] is written


@pytest.mark.parametrize("kernel", ["gaussianerf", "winckelmans"])
def test_zeta_fmm_list(vpm, handle, kernel):
    pf = vpm.fields.cloud_field(4000, kernel=vpm.KERNELS[kernel], seed=13)
    ll = leaflists.build_leaf_lists(pf.get_X(), pf.get_sigma(), ncrit=50, theta=0.4)
    leaves = (ll["leaf_begin"], ll["leaf_end"])
    dl = ll["direct_list"][::-1].copy()           # unsorted, asymmetric order
    dl = dl[: len(dl) * 3 // 4]                   # and an asymmetric list
    vpm.fields.random_results(pf, scale=1e-2)     # zeta_fmm accumulates on what is there
    ref = pf.particles.copy(order="F")
    oracle.zeta_leafpairs(ref, ll["sort_index"], leaves, dl, kernel)
    vpm.zeta_fmm(pf, ll["sort_index"], leaves, dl)
    assert relerr(pf.particles[15:18], ref[15:18]) < TOL_FP64


# ------------------------------------------------------------ error behaviour
def test_errors_are_codes_with_messages(vpm, handle):
    pf = vpm.fields.cloud_field(10)
    P = pf.particles
    rc = handle.lib.vpm_uj_direct(handle.ptr, P.ctypes.data, 46, 10, 99, 1)
    assert rc == -1 and b"kernel_id" in handle.lib.vpm_last_error(handle.ptr)
    rc = handle.lib.vpm_uj_direct(handle.ptr, P.ctypes.data, 12, 10, 3, 1)
    assert rc == -1
    rc = handle.lib.vpm_eval(handle.ptr, 3, 1)
    assert rc in (-6, 0)  # ESTATE unless a previous test left state resident
    with pytest.raises(vpm.VpmError):
        vpm.Handle(device_ids=[10_000])


# ------------------------------- pageable host matrices: the pinned staging ring
def _with_pinned(handle, P, fn):
    handle.check(handle.lib.vpm_pin_host(handle.ptr, P.ctypes.data, P.nbytes))
    try:
        fn()
    finally:
        handle.check(handle.lib.vpm_unpin_host(handle.ptr, P.ctypes.data))


def test_pageable_matrix_goes_through_ring_bit_identical(vpm, handle):
    """A pageable matrix travels through the two pinned slots (h2d_strided / d2h_strided: several slots'
    worth of columns at this size); a page-locked one by 2-D DMA.  Same kernels, same operands: the
    results must be identical to the last bit, and rows the entry points do not own stay untouched."""
    n = 330_000                                    # 56-byte rows: one 16 MB slot holds 299 593 columns
    pf = vpm.fields.cloud_field(n, kernel=vpm.winckelmans, seed=21)
    vpm.fields.random_results(pf, scale=1e-2)
    ll = leaflists.build_leaf_lists(pf.get_X(), pf.get_sigma(), ncrit=64, theta=0.4)
    leaves = (ll["leaf_begin"], ll["leaf_end"])
    order, dl = ll["sort_index"], ll["direct_list"]
    start = pf.particles.copy(order="F")

    def run(pinned, call):
        pf.particles[...] = start
        if pinned:
            _with_pinned(handle, pf.particles, call)
        else:
            call()
        return pf.particles.copy(order="F")

    calls = {
        "zeta_fmm": lambda: vpm.zeta_fmm(pf, order, leaves, dl),
        "Estr_fmm": lambda: vpm.Estr_fmm(pf, order, order, leaves, leaves, dl),
        "nearfield_ranges": lambda: vpm.fmm_nearfield_device(
            pf, [range(0, 1000), range(150_000, 151_000), range(n - 500, n)], (False, True, True), pf,
            [[range(0, 3000)], [range(149_000, 152_000)], [range(n - 2000, n)]]),
    }
    for name, call in calls.items():
        a = run(False, call)
        b = run(True, call)
        assert np.array_equal(a, b), name
        assert not np.array_equal(a, start), name   # the call did write something


def test_pageable_f32_matrix_and_targets_through_ring(vpm, handle):
    pf = vpm.fields.cloud_field(40_000, kernel=vpm.gaussianerf, seed=22, R=np.float32)
    start = pf.particles.copy(order="F")
    vpm.UJ_direct(pf, sfs=True, reset_sfs=True)
    a = pf.particles.copy(order="F")
    pf.particles[...] = start
    _with_pinned(handle, pf.particles, lambda: vpm.UJ_direct(pf, sfs=True, reset_sfs=True))
    assert np.array_equal(a, pf.particles)
    # UJ_direct(source, target): targets accumulate
    src = vpm.fields.cloud_field(20_000, kernel=vpm.winckelmans, seed=23)
    tgt = vpm.fields.cloud_field(50_000, kernel=vpm.winckelmans, seed=24)
    vpm.fields.random_results(tgt, scale=1e-2)
    t0 = tgt.particles.copy(order="F")
    vpm.UJ_direct(src, tgt)
    a = tgt.particles.copy(order="F")
    tgt.particles[...] = t0
    _with_pinned(handle, tgt.particles, lambda: vpm.UJ_direct(src, tgt))
    assert np.array_equal(a, tgt.particles)
    # and the accumulated sums are right (oracle on the first 200 targets)
    ref = np.zeros((16, 200), order="F")
    ref[0:3] = t0[0:3, :200]
    sb = np.asfortranarray(vpm.source_system_to_buffer(src))
    oracle.direct_buffers(ref, 0, 200, sb, 0, src.np, "winckelmans")
    assert relerr(a[9:12, :200] - t0[9:12, :200], ref[4:7]) < 1e-10   # difference of O(1e-2) numbers


def test_pageable_buffers_and_field_through_ring(vpm, handle):
    """contiguous blocks from / to pageable memory (FastMultipole-style buffers and the near-field list of
    Hook 3, the whole matrix of vpm_field_upload / _download) also travel through the pinned ring"""
    n = 330_000
    pf = vpm.fields.cloud_field(n, kernel=vpm.winckelmans, seed=25)
    ll = leaflists.build_leaf_lists(pf.get_X(), pf.get_sigma(), ncrit=64, theta=0.4)
    leaves = (ll["leaf_begin"], ll["leaf_end"])
    order, dl = ll["sort_index"], ll["direct_list"]
    sb = np.asfortranarray(vpm.source_system_to_buffer(pf)[:, order])
    res = []
    for pinned in (False, True):
        tb = np.zeros((16, n), order="F")
        tb[0:3] = pf.get_X()[:, order]
        if pinned:
            for a in (tb, sb):
                handle.check(handle.lib.vpm_pin_host(handle.ptr, a.ctypes.data, a.nbytes))
        try:
            vpm.nearfield_device(tb, leaves, sb, leaves, dl, vpm.winckelmans)
        finally:
            if pinned:
                for a in (tb, sb):
                    handle.check(handle.lib.vpm_unpin_host(handle.ptr, a.ctypes.data))
        res.append(tb)
    assert np.array_equal(res[0], res[1]) and np.abs(res[0][4:16]).max() > 0
    # Hook 2 on the same buffers (a slab of targets against a slab of sources)
    out = []
    for pinned in (False, True):
        tb = np.zeros((16, n), order="F")
        tb[0:3] = pf.get_X()[:, order]
        if pinned:
            handle.check(handle.lib.vpm_pin_host(handle.ptr, tb.ctypes.data, tb.nbytes))
        try:
            vpm.direct_buffers(tb, (1000, 41000), sb, (0, 50000), vpm.winckelmans)
        finally:
            if pinned:
                handle.check(handle.lib.vpm_unpin_host(handle.ptr, tb.ctypes.data))
        out.append(tb)
    assert np.array_equal(out[0], out[1]) and np.abs(out[0][4:16, 1000:41000]).max() > 0
    assert not out[0][4:16, :1000].any() and not out[0][4:16, 41000:].any()
    # whole-matrix round trip of the resident field
    vpm.fields.random_results(pf, scale=1.0)
    before = pf.particles.copy(order="F")
    rf = vpm.ResidentField(pf)
    pf.particles[...] = 0.0
    rf.download()
    assert np.array_equal(pf.particles[:, :n], before[:, :n])
