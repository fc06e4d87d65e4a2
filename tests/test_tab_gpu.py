"""The bank-replicated log-spaced table kernel of the gaussianerf / gaussian U/J sweep
(csrc/vpm_kernels_tab.cuh, VPM_OPT_UJ_TABLE): device math against mpmath, parity against the
oracle with the kernel forced on small and ragged fields, the automatic choice on a field large
enough to take it, far-field shortcut on and off, agreement with the round-1 kernels."""
import mpmath as mp
import numpy as np
import pytest

from helpers import TOL_FP64, assert_parity, relerr, stretching
from oracle import hp_oracle, oracle

pytestmark = pytest.mark.gpu

FAMILIES = ["gaussianerf", "gaussian"]


@pytest.fixture()
def tab_forced(vpm, handle):
    handle.set_option(vpm._cabi.OPT_UJ_TABLE, 1)
    yield handle
    handle.set_option(vpm._cabi.OPT_UJ_TABLE, 0)


def oracle_uj(pf, **kw):
    ref = pf.particles.copy(order="F")
    oracle.uj_direct(ref, pf.np, pf.kernel.name, transposed=pf.transposed, **kw)
    return ref


@pytest.mark.parametrize("kernel,kid,smax", [("gaussianerf", 2, 9.0), ("gaussian", 1, 3.45)])
def test_table_scalars_vs_mpmath(handle, kernel, kid, smax):
    """A = g/r^3 and B = (dg/(sigma r) - 3g/r^2)/r^3 at sigma = 1 over the whole table range,
    including every row boundary neighbourhood that a coarse sweep would miss"""
    rng = np.random.default_rng(3)
    s = np.concatenate([np.linspace(1e-3, smax * 0.9999, 700), rng.uniform(0, smax, 300) * 0.9999,
                        [1e-6, 1e-4, 0.5, 1.0, 2.0],
                        # beyond the cut-off: the power-law rows (g == 1), several octaves, both exponent parities
                        smax * 2.0 ** np.linspace(0.001, 12, 400), [smax * 1.000001, 1e3, 12345.678, 1e6]])
    r2 = s * s
    A = np.empty_like(r2)
    B = np.empty_like(r2)
    handle.check(handle.lib.vpm_test_math(handle.ptr, 4, kid, r2.ctypes.data, A.ctypes.data, B.ctypes.data, r2.size))
    worstA = worstB = 0.0
    for i in range(r2.size):
        r2m = mp.mpf(float(r2[i]))
        r = mp.sqrt(r2m)
        g, dg = hp_oracle.g_dgdr(kernel, r)
        Am = g / r**3
        Bm = (dg / r - 3 * g / r2m) / r**3
        worstA = max(worstA, float(abs((mp.mpf(float(A[i])) - Am) / Am)))
        # B enters J as B c dx with |c dx| ~ r^2 |Gamma|: for the gaussian family B -> 0 like s, so measure
        # B r^2 against the size that product has where it matters (0.01; its maximum is ~0.5 at s ~ 1)
        if kernel == "gaussian":
            worstB = max(worstB, float(abs(mp.mpf(float(B[i])) - Bm) * r2m / max(abs(Bm) * r2m, mp.mpf("0.01"))))
        else:
            worstB = max(worstB, float(abs((mp.mpf(float(B[i])) - Bm) / Bm)))
    assert worstA < 2e-15, worstA
    # gaussian below s = 1/2 is off the table: there the device restates the reference's own formula, whose
    # E s^3 - g cancels like the reference's does
    assert worstB < (2e-13 if kernel == "gaussian" else 2e-14), worstB


@pytest.mark.parametrize("kid", [1, 2])
def test_table_zero_distance(handle, kid):
    r2 = np.array([0.0])
    A, B = np.empty(1), np.empty(1)
    handle.check(handle.lib.vpm_test_math(handle.ptr, 4, kid, r2.ctypes.data, A.ctypes.data, B.ctypes.data, 1))
    assert A[0] == 0.0 and np.isfinite(B[0])


@pytest.mark.parametrize("kernel", FAMILIES)
def test_forced_table_ring_c1(vpm, tab_forced, kernel):
    pf = vpm.fields.ring_field(Nphi=100, nc=3, kernel=vpm.KERNELS[kernel])
    ref = oracle_uj(pf, sfs=True, reset=True, reset_sfs=True)
    vpm.UJ_direct(pf, sfs=True, reset=True, reset_sfs=True)
    errs = assert_parity(pf.particles, ref, pf.np, what=f"tab ring/{kernel}")
    assert relerr(stretching(pf.particles, pf.np), stretching(ref, pf.np)) < TOL_FP64
    print(kernel, errs)


@pytest.mark.parametrize("kernel", FAMILIES)
@pytest.mark.parametrize("n", [1, 2, 31, 129, 1000, 1025, 3001])
def test_forced_table_cloud_sizes(vpm, tab_forced, kernel, n):
    """ragged sizes around the tile (128) and the 1024-target CTA, down to one particle"""
    pf = vpm.fields.cloud_field(n, kernel=vpm.KERNELS[kernel], seed=200 + n)
    ref = oracle_uj(pf, reset=True)
    vpm.UJ_direct(pf, reset=True)
    assert_parity(pf.particles, ref, n, rows=("U", "J"), what=f"tab cloud{n}/{kernel}")


@pytest.mark.parametrize("kernel", FAMILIES)
def test_forced_table_shortcut_on_off_and_round1_kernel(vpm, handle, kernel):
    """dense blob (about half the pairs inside the regularised range): table kernel with and without the
    far-field shortcut, and the round-1 kernel, all within 1e-12 of the oracle and 1e-13 of each other"""
    # jittered lattice with wide cores (sigma = 2.6 / 1.6 lattice spacings): no pair closer than half a
    # spacing -- the reference's own gaussian formula g = 1 - exp(-s^3) loses log10(1/s^3) digits, so on
    # a field with pairs at s ~ 0.01 the ORACLE is off by 1e-12 (checked against long double) while the
    # table is not; parity against the reference is only meaningful where the reference is accurate
    n = 5000
    pf = vpm.fields.cloud_field(n, kernel=vpm.KERNELS[kernel], seed=5)
    P = pf.particles
    P[6, :n] *= 4.0 if kernel == "gaussianerf" else 2.5
    ref = oracle_uj(pf, reset=True)
    base = P.copy(order="F")
    out = {}
    for tag, opt, nosc in (("tab", 1, False), ("tab_noshortcut", 1, True), ("round1", 2, False)):
        P[:] = base
        handle.set_option(vpm._cabi.OPT_UJ_TABLE, opt)
        try:
            vpm.UJ_direct(pf, reset=True, no_farfield_shortcut=nosc)
        finally:
            handle.set_option(vpm._cabi.OPT_UJ_TABLE, 0)
        out[tag] = P.copy(order="F")
        assert_parity(out[tag], ref, n, rows=("U", "J"), what=f"{tag}/{kernel}")
    assert_parity(out["tab"], out["tab_noshortcut"], n, rows=("U", "J"), tol=1e-13)
    assert_parity(out["tab"], out["round1"], n, rows=("U", "J"), tol=1e-13)


@pytest.mark.parametrize("kernel", FAMILIES)
def test_forced_table_coincident_and_static(vpm, tab_forced, kernel):
    """coincident particles (r2 == 0 pairs are skipped, src/FLOWVPM_fmm.jl:118) and static targets
    (never reset, src/FLOWVPM_particlefield.jl:468) through the table kernel"""
    pf = vpm.fields.cloud_field(700, kernel=vpm.KERNELS[kernel], static_fraction=0.2, seed=11)
    pf.particles[0:3, 100:110] = pf.particles[0:3, 0:10]
    vpm.fields.random_results(pf, scale=1e-3)
    ref = oracle_uj(pf, sfs=True, reset=True, reset_sfs=False)
    vpm.UJ_direct(pf, sfs=True, reset=True, reset_sfs=False)
    assert_parity(pf.particles, ref, pf.np, rows=("U", "J", "SFS", "W", "PSE"))


def _run_with(vpm, handle, pf, base, opt):
    pf.particles[:] = base
    handle.set_option(vpm._cabi.OPT_UJ_TABLE, opt)
    try:
        vpm.UJ_direct(pf, reset=True)
    finally:
        handle.set_option(vpm._cabi.OPT_UJ_TABLE, 0)
    return pf.particles[9:24].copy()


def test_automatic_choice_follows_the_sampled_near_fraction(vpm, handle):
    """20 000 particles fill the GPU with 1024-target CTAs, so the gaussian families are candidates for the table
    kernel; the automatic plan takes it only when >= 40 % of a deterministic sample of warps (32 consecutive
    targets x one source) see a pair inside the regularised range: a dense field (sigma x 8) runs the table
    kernel, the sparse cloud the round-1 kernel -- seen from outside as bit-identity with the forced runs.
    Every variant is in parity with the oracle."""
    for kernel, scale, expect in (("gaussianerf", 8.0, 1), ("gaussianerf", 0.4, 2), ("gaussian", 8.0, 1), ("gaussian", 0.4, 2)):
        pf = vpm.fields.cloud_field(20000, kernel=vpm.KERNELS[kernel], seed=77)
        pf.particles[0:3] *= np.array([[1.0], [1.0], [1.0 / 7.0]])   # unit cube
        pf.particles[6] *= scale
        base = pf.particles.copy(order="F")
        ref = oracle_uj(pf, reset=True)
        auto = _run_with(vpm, handle, pf, base, 0)
        assert_parity(pf.particles, ref, pf.np, rows=("U", "J"), what=f"auto {kernel} x{scale}")
        forced = {opt: _run_with(vpm, handle, pf, base, opt) for opt in (1, 2)}
        assert not np.array_equal(forced[1], forced[2])          # the two kernels round differently
        assert np.array_equal(auto, forced[expect]), (kernel, scale)


def test_option_validation(vpm, handle):
    for opt, bad in ((vpm._cabi.OPT_UJ_TABLE, 3), (vpm._cabi.OPT_UJ_VARIANT, 13), (vpm._cabi.OPT_UJ_VARIANT, 20),
                     (vpm._cabi.OPT_SFS_VARIANT, 21), (99, 1)):
        with pytest.raises(vpm.VpmError):
            handle.set_option(opt, bad)
    for v in (11, 12, 21, 22, 31, 32, 41, 42, 0):
        handle.set_option(vpm._cabi.OPT_UJ_VARIANT, v)


def test_device_entry_points_on_two_streams(vpm, handle):
    """vpm_uj_device / vpm_sfs_device return without synchronising and share the handle's scratch: two calls
    on DIFFERENT streams back to back (and a synchronous call right after) must not corrupt each other"""
    import torch
    n1, n2 = 40000, 3000
    f1 = vpm.fields.cloud_field(n1, kernel=vpm.winckelmans, seed=41)
    f2 = vpm.fields.cloud_field(n2, kernel=vpm.gaussianerf, seed=42)
    s1 = torch.from_numpy(np.ascontiguousarray(vpm.source_system_to_buffer(f1).T)).cuda()
    s2 = torch.from_numpy(np.ascontiguousarray(vpm.source_system_to_buffer(f2).T)).cuda()
    o1 = torch.zeros((n1, 12), dtype=torch.float64, device="cuda")
    o2 = torch.zeros((n2, 12), dtype=torch.float64, device="cuda")
    # references, one at a time
    r1, r2 = torch.zeros_like(o1), torch.zeros_like(o2)
    st0 = torch.cuda.current_stream().cuda_stream
    handle.check(handle.lib.vpm_uj_device(handle.ptr, s1.data_ptr(), n1, 0, n1, r1.data_ptr(), 3, 0, st0))
    torch.cuda.synchronize()
    handle.check(handle.lib.vpm_uj_device(handle.ptr, s2.data_ptr(), n2, 0, n2, r2.data_ptr(), 2, 0, st0))
    torch.cuda.synchronize()
    a, b = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(3):
        o1.zero_(); o2.zero_()
        torch.cuda.synchronize()
        handle.check(handle.lib.vpm_uj_device(handle.ptr, s1.data_ptr(), n1, 0, n1, o1.data_ptr(), 3, 0, a.cuda_stream))
        handle.check(handle.lib.vpm_uj_device(handle.ptr, s2.data_ptr(), n2, 0, n2, o2.data_ptr(), 2, 0, b.cuda_stream))
        vpm.UJ_direct(f2, reset=True)          # synchronous entry point on the handle's own stream
        torch.cuda.synchronize()
        assert torch.equal(o1, r1) and torch.equal(o2, r2)
        assert np.array_equal(f2.get_U(), r2[:, 0:3].cpu().numpy().T)


@pytest.mark.parametrize("kernel", ["winckelmans", "gaussianerf"])
@pytest.mark.parametrize("pinned", [False, True])
def test_small_field_graph_replay(vpm, kernel, pinned):
    """VPM_OPT_SMALL_GRAPH: the third and later calls with the same (matrix, np, kernel, flags) replay a
    captured CUDA graph; the field changes between calls (positions move, results accumulate, static flags
    appear) and every call must match the oracle like the ordinary path does"""
    h = vpm.Handle(1)
    try:
        h.set_option(vpm._cabi.OPT_SMALL_GRAPH, 1)
        pf = vpm.fields.cloud_field(900, kernel=vpm.KERNELS[kernel], seed=19)
        P = pf.particles
        if pinned:
            h.check(h.lib.vpm_pin_host(h.ptr, P.ctypes.data, P.nbytes))
        rng = np.random.default_rng(0)
        for call in range(6):
            P[0:3, :pf.np] += 1e-3 * rng.standard_normal((3, pf.np))
            if call == 4:
                P[42, 5:50:3] = 1.0          # static particles appear: a different key, captured separately
            ref = P.copy(order="F")
            kw = dict(sfs=True, reset=call % 2 == 0, reset_sfs=True)
            oracle.uj_direct(ref, pf.np, kernel, **kw)
            vpm.UJ_direct(pf, handle=h, **kw)
            assert_parity(P, ref, pf.np, rows=("U", "J", "SFS", "W", "PSE"), what=f"call {call}")
            assert h.timing()["kernel_launches"] > 0 and h.timing()["uj_pairs"] == pf.np ** 2
        # same field through a handle with the graph path switched off: bit-identical
        h2 = vpm.Handle(1)
        h2.set_option(vpm._cabi.OPT_SMALL_GRAPH, 0)
        try:
            a = P.copy(order="F")
            vpm.UJ_direct(pf, sfs=True, reset=True, reset_sfs=True, handle=h)
            b = P.copy(order="F")
            P[:] = a
            vpm.UJ_direct(pf, sfs=True, reset=True, reset_sfs=True, handle=h2)
            assert np.array_equal(P, b)
        finally:
            h2.close()
        if pinned:
            h.check(h.lib.vpm_unpin_host(h.ptr, P.ctypes.data))
    finally:
        h.close()
