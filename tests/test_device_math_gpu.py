"""Accuracy of the device math behind the pair loops (the "accuracy-checked polynomials"):
rsqrt_fp64 (MUFU.RSQ64H seed + one step), exp_fp64, and each family's (A, B) pair scalars
against mpmath.  A = g/r^3, B = (dg/(sigma r) - 3g/r^2)/r^3 with sigma = 1."""
import mpmath as mp
import numpy as np
import pytest

from oracle import hp_oracle

pytestmark = pytest.mark.gpu


def run(handle, op, arg, x, two=False):
    x = np.ascontiguousarray(x, dtype=np.float64)
    o1 = np.empty_like(x)
    o2 = np.empty_like(x)
    handle.check(handle.lib.vpm_test_math(handle.ptr, op, arg, x.ctypes.data, o1.ctypes.data,
                                          o2.ctypes.data if two else None, x.size))
    return (o1, o2) if two else o1


def test_rsqrt(handle):
    rng = np.random.default_rng(0)
    x = np.concatenate([10.0 ** rng.uniform(-200, 200, 200000), rng.uniform(1, 4, 200000)])
    y = run(handle, 0, 0, x)
    rel = np.abs(y * np.sqrt(x) - 1.0)   # sqrt is correctly rounded: error budget 1 ulp + ours
    assert rel.max() < 4.5e-16, rel.max()


def test_exp(handle):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-700, 0, 300000), rng.uniform(-1, 1, 100000), [0.0, -700.0]])
    y = run(handle, 1, 0, x)
    ref = np.exp(x)
    rel = np.abs(y - ref) / ref
    assert rel.max() < 4.5e-16, rel.max()
    # spot check against mpmath (libm is itself < 1 ulp)
    for xv in (-0.3, -17.25, -345.678, -699.9):
        yv = run(handle, 1, 0, np.array([xv]))[0]
        assert abs(yv - float(mp.exp(mp.mpf(xv)))) / float(mp.exp(mp.mpf(xv))) < 3e-16


@pytest.mark.parametrize("kernel,kid", [("singular", 0), ("gaussian", 1), ("gaussianerf", 2), ("winckelmans", 3)])
def test_pair_scalars_vs_mpmath(handle, kernel, kid):
    s = np.concatenate([np.linspace(0.3, 12, 400), [0.5, 1.0, 2.0, 3.4, 3.5, 8.9, 9.1, 30.0]])
    if kernel == "gaussianerf":   # the table ends at s = 9 (the kernel switches to g == 1 there)
        s = np.concatenate([s[s < 9.0], np.linspace(1e-3, 0.3, 50), np.sqrt(np.arange(0, 162) / 2 + 0.25 - 1e-9)])
    r2 = s * s
    A, B = run(handle, 2, kid, r2, two=True)
    worstA = worstB = 0.0
    for i, sv in enumerate(s):
        r2m = mp.mpf(float(r2[i]))
        r = mp.sqrt(r2m)
        g, dg = hp_oracle.g_dgdr(kernel, r)
        Am = g / r**3
        Bm = (dg / r - 3 * g / r2m) / r**3
        worstA = max(worstA, abs((mp.mpf(float(A[i])) - Am) / Am))
        worstB = max(worstB, abs((mp.mpf(float(B[i])) - Bm) / Bm))
    # A is a product of a few correctly-rounded-ish factors; B of the regularised families
    # inherits the reference's own cancellation (aux) for s < 1: still ~1e-15 at s >= 0.3
    # (gaussianerf comes from the dedicated G(u) table: no erf - aux cancellation on the device)
    assert worstA < 3e-15, worstA
    # gaussian: g = 1 - exp(-s^3) loses log10(1/s^3) digits in ANY FP64 evaluation (the reference's too)
    assert worstB < (3e-13 if kernel == "gaussian" else 4e-14), worstB


def test_zero_distance_is_masked(handle):
    for kid in range(4):
        A, B = run(handle, 2, kid, np.array([0.0]), two=True)
        assert A[0] == 0.0 and np.isfinite(B[0])


@pytest.mark.parametrize("kernel,kid", [("singular", 0), ("gaussian", 1), ("gaussianerf", 2), ("winckelmans", 3)])
def test_zeta_weights(handle, kernel, kid):
    s = np.concatenate([[0.0], np.linspace(0.01, 7, 300)])
    w = run(handle, 3, kid, s * s)
    for i, sv in enumerate(s):
        ref = float(hp_oracle.zeta(kernel, mp.sqrt(mp.mpf(float(s[i] * s[i])))))
        assert abs(w[i] - ref) <= 2e-15 * abs(ref) * (1 + abs(np.log(ref)) if ref > 0 else 1) + 1e-300
