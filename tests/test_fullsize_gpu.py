"""Parity at BASELINE.json's full size (config C4, N = 2^20) and size-independent properties.

At 2^20 particles the oracle cannot do the N^2 sweep in seconds, so (i) the GPU does the
whole sweep and the oracle recomputes a slice of targets against ALL sources, and (ii)
properties the domain offers are checked: linearity in Gamma, superposition of source sets,
translation invariance, SFS == 0 for a uniform velocity gradient."""
import numpy as np
import pytest

from helpers import TOL_FP64, relerr
from oracle import oracle

pytestmark = pytest.mark.gpu


def slice_targets(n, k_head=96, k_rand=160, seed=0):
    rng = np.random.default_rng(seed)
    return np.unique(np.concatenate([np.arange(min(n, k_head)), rng.choice(n, min(n, k_rand), replace=False),
                                     [n - 1]])).astype(np.int64)


def check_slice(pf, idx, kernel, what, sfs=True):
    """GPU did the whole field in pf.particles; the oracle re-does the targets `idx` against all sources
    (SFS over the J rows the GPU produced: identical inputs for the second sweep)."""
    P, n = pf.particles, pf.np
    U, J, S = oracle.uj_slice(P, n, idx, kernel, sfs=sfs, transposed=pf.transposed)
    errs = {"U": relerr(P[9:12, idx], U), "J": relerr(P[15:24, idx], J)}
    if sfs:
        errs["SFS"] = relerr(P[39:42, idx], S)
        if kernel != "singular":   # zeta_sing(r) = [r == 0]: only the self pair, whose JT - JS is 0
            assert np.abs(S).max() > 0
    assert all(v < TOL_FP64 for v in errs.values()), (what, errs)
    assert np.all(np.isfinite(P[9:27, :n])) and np.all(np.isfinite(P[39:42, :n]))
    return errs


@pytest.mark.parametrize("kernel", ["singular", "gaussian", "gaussianerf", "winckelmans"])
def test_c4_full_size_slice_parity_with_sfs(vpm, handle, kernel):
    """BASELINE config 4 at N = 2^20: U, J AND SFS of the whole cloud on the GPU, 257 targets re-done by the oracle"""
    n = 1 << 20
    pf = vpm.fields.cloud_field(n, kernel=vpm.KERNELS[kernel])
    vpm.UJ_direct(pf, sfs=True, reset=True, reset_sfs=True)
    print(kernel, check_slice(pf, slice_targets(n), kernel, f"C4/{kernel}"))


@pytest.mark.parametrize("kernel", ["gaussianerf", "winckelmans"])
def test_c2_leapfrog_full_field(vpm, handle, kernel):
    """BASELINE config 2 at its full size: two rings, Nphi = 100, nc = 6 -> 33 800 particles
    (test/runtests_leapfrog.jl:42-46 geometry), the WHOLE field against the oracle, SFS included"""
    R = 0.7906
    pf = vpm.fields.ring_field(Nphi=100, nc=6, R=R, Rcross=0.1 * R, rings=2, dZ=0.7906, kernel=vpm.KERNELS[kernel])
    assert pf.np == 33800
    ref = pf.particles.copy(order="F")
    oracle.uj_direct(ref, pf.np, kernel, sfs=True, reset=True, reset_sfs=True, nthreads=oracle.num_procs())
    vpm.UJ_direct(pf, sfs=True, reset=True, reset_sfs=True)
    from helpers import assert_parity, stretching
    print(kernel, assert_parity(pf.particles, ref, pf.np, what=f"C2/{kernel}"))
    assert relerr(stretching(pf.particles, pf.np), stretching(ref, pf.np)) < TOL_FP64


@pytest.mark.parametrize("kernel", ["gaussianerf", "winckelmans"])
def test_c3_jet_slice_with_static_and_sfs(vpm, handle, kernel):
    """BASELINE config 3 shape at its size: jet column of ~3e5 particles, 10 % static inflow, U/J + SFS;
    a slice of targets (static and free ones) re-done by the oracle"""
    pf = vpm.fields.jet_field(n_target=300_000, kernel=vpm.KERNELS[kernel])
    n = pf.np
    st = pf.get_static() != 0
    assert n >= 299_000 and 0.05 < st.mean() < 0.15
    vpm.UJ_direct(pf, sfs=True, reset=True, reset_sfs=True)
    idx = slice_targets(n, k_head=64, k_rand=192, seed=3)
    assert st[idx].any() and (~st[idx]).any()
    print(kernel, check_slice(pf, idx, kernel, f"C3/{kernel}"))
    assert np.all(pf.particles[39:42, :n][:, st] == 0)   # static particles take no part in the SFS sweep


def test_linearity_and_superposition(vpm, handle):
    n = 1 << 16
    pf = vpm.fields.cloud_field(n, kernel=vpm.winckelmans, seed=77)
    vpm.UJ_direct(pf)
    U1, J1 = pf.get_U().copy(), pf.get_J().copy()
    pf.particles[3:6] *= 2.0   # exact in FP64
    vpm.UJ_direct(pf)
    assert np.array_equal(pf.get_U(), 2.0 * U1) and np.array_equal(pf.get_J(), 2.0 * J1)
    # superposition: sources split in two halves, evaluated through UJ_direct(source, target)
    pf.particles[3:6] *= 0.5
    half = n // 2
    tgt = vpm.ParticleField(n, kernel=vpm.winckelmans)
    tgt.particles[:] = pf.particles
    tgt.particles[9:27] = 0
    tgt.np = n
    for lo, hi in ((0, half), (half, n)):
        src = vpm.ParticleField(hi - lo, kernel=vpm.winckelmans)
        src.particles[:] = pf.particles[:, lo:hi]
        src.np = hi - lo
        vpm.UJ_direct(src, tgt)
    assert relerr(tgt.get_U(), U1) < TOL_FP64 and relerr(tgt.get_J(), J1) < TOL_FP64


@pytest.mark.parametrize("kernel", ["winckelmans", "gaussianerf", "gaussian", "singular"])
def test_translation_invariance(vpm, handle, kernel):
    n = 20000
    pf = vpm.fields.cloud_field(n, kernel=vpm.KERNELS[kernel], seed=5)
    vpm.UJ_direct(pf, sfs=True, reset=True, reset_sfs=True)
    a = pf.particles.copy(order="F")
    pf.particles[0:3] += np.array([[0.5], [-0.25], [1.0]])   # exactly representable shifts keep dx to ~1 ulp
    vpm.UJ_direct(pf, sfs=True, reset=True, reset_sfs=True)
    for rows in (slice(9, 12), slice(15, 24), slice(39, 42)):
        assert relerr(pf.particles[rows], a[rows]) < 1e-11


def test_sfs_vanishes_for_uniform_gradient(vpm, handle):
    import torch
    n = 50000
    pf = vpm.fields.cloud_field(n, kernel=vpm.gaussianerf, seed=6)
    src8 = torch.from_numpy(np.ascontiguousarray(vpm.source_system_to_buffer(pf).T)).cuda()
    J = torch.from_numpy(np.tile(np.arange(1.0, 10.0), (n, 1))).cuda()
    out = torch.ones((n, 3), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    handle.check(handle.lib.vpm_sfs_device(handle.ptr, src8.data_ptr(), J.data_ptr(), None, n, 0, n, out.data_ptr(),
                                           pf.kernel.id, 4 | 8, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert float(out.abs().max()) == 0.0
