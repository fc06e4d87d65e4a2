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


@pytest.mark.parametrize("kernel", ["winckelmans", "gaussianerf"])
def test_full_size_slice_parity(vpm, handle, kernel):
    n = 1 << 20
    pf = vpm.fields.cloud_field(n, kernel=vpm.KERNELS[kernel])
    vpm.UJ_direct(pf, reset=True)
    sb = vpm.source_system_to_buffer(pf)
    rng = np.random.default_rng(0)
    idx = np.concatenate([np.arange(96), rng.choice(n, 160, replace=False)])
    tb = np.zeros((16, len(idx)), order="F")
    tb[0:3] = pf.get_X()[:, idx]
    oracle.direct_buffers(tb, 0, len(idx), sb, 0, n, kernel, True, True, oracle.max_threads())
    assert relerr(pf.get_U()[:, idx], tb[4:7]) < TOL_FP64
    assert relerr(pf.get_J()[:, idx], tb[7:16]) < TOL_FP64
    assert np.all(np.isfinite(pf.get_U())) and np.all(np.isfinite(pf.get_J()))


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
