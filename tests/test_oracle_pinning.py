"""Pins the CPU oracle (oracle/vpm_oracle.c) -- the checker of every GPU parity test --
against everything available without a Julia runtime:
  * the analytic two-particle known answers of the reference's scripts/check_fmm.jl:39-98,
  * the committed 50-digit mpmath golden vectors (tests/golden/p2p_golden.npz),
  * fdlibm-accuracy of custom_erf64 (src/FLOWVPM_gpu_erf.jl:159-191) against libm,
  * the reset / static / accumulate rules of src/FLOWVPM_particlefield.jl:464-511,
  * the physics assertion of test/runtests_singlevortexring.jl:127-143 (ring speed, 2 %).
"""
import math
import os

import numpy as np
import pytest

from oracle import oracle, hp_oracle
from helpers import KERNELS, relerr, TOL_FP64
import physics

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "p2p_golden.npz"))


def make_field(n=None):
    n = n or GOLD["X"].shape[1]
    P = np.zeros((46, n), order="F")
    P[0:3] = GOLD["X"][:, :n]
    P[3:6] = GOLD["Gamma"][:, :n]
    P[6] = GOLD["sigma"][:n]
    return P


# ---------------------------------------------------------------- custom_erf64
def test_erf64_matches_libm_to_2ulp():
    xs = np.concatenate([np.linspace(-7, 7, 20001), [0.0, 0.84375, 1.25, 1 / 0.35, 6.0, -0.84375, 5.999999],
                         np.nextafter([0.84375, 1.25, 2.857142857142857, 6.0], 0),
                         10.0 ** np.linspace(-300, -1, 200)])
    worst = 0.0
    for x in xs:
        a, b = oracle.erf64(x), math.erf(x)
        ulp = np.spacing(abs(b)) if b != 0 else 5e-324
        worst = max(worst, abs(a - b) / ulp)
    assert worst <= 2.0, worst
    assert oracle.erf64(6.0) == 1.0 and oracle.erf64(-6.0) == -1.0 and oracle.erf64(0.0) == 0.0


def _tail_tol(v, eps):
    v = abs(v)
    return 1e-300 if v == 0 else eps * v * (1 + abs(math.log(v))) + 1e-300


def test_erf32_restatement_matches_libm():
    """custom_erf32 (src/FLOWVPM_gpu_erf.jl:124-156) in Float32 arithmetic, Float64 `sb7` quirk included"""
    xs = np.concatenate([np.linspace(-7, 7, 8001), [0.84374, 0.84376, 1.2499, 1.2501, 2.857, 2.858, 5.99, 6.01]])
    worst = max(abs(oracle.erf32(np.float32(x)) - math.erf(float(np.float32(x)))) for x in xs)
    assert worst < 2e-7, worst
    assert oracle.erf32(0.0) == 0.0 and oracle.erf32(7.0) == 1.0 and oracle.erf32(-7.0) == -1.0


@pytest.mark.parametrize("kernel", KERNELS)
def test_float32_field_semantics_close_to_float64(kernel, vpm):
    """fmm.direct! on Float32 buffers under Julia's promotion rules (ParticleField{Float32}) against the
    Float64 restatement on the same (Float32-rounded) inputs: 1e-5 is the north star's FP32 bar.  Compact
    field (positions O(1), neighbour distances O(0.1)): the Float32 dx of the reference carries 1e-6."""
    pf = vpm.fields.cloud_field(600, kernel=vpm.KERNELS[kernel], seed=2)
    pf.particles[0:3] *= np.array([[1.0], [1.0], [1.0 / 7]])   # unit cube
    src32 = np.asfortranarray(vpm.source_system_to_buffer(pf).astype(np.float32))
    tb32 = np.zeros((16, pf.np), dtype=np.float32, order="F")
    tb32[0:3] = src32[0:3]
    oracle.direct_buffers_f32(tb32, 0, pf.np, src32, 0, pf.np, kernel)
    src64 = np.asfortranarray(src32.astype(np.float64))
    tb64 = np.zeros((16, pf.np), order="F")
    tb64[0:3] = src64[0:3]
    oracle.direct_buffers(tb64, 0, pf.np, src64, 0, pf.np, kernel)
    assert relerr(tb32[4:7], tb64[4:7]) < 1e-5 and relerr(tb32[7:16], tb64[7:16]) < 1e-5
    assert relerr(tb32[4:7], tb64[4:7]) > 1e-9   # it IS a Float32 evaluation


@pytest.mark.parametrize("kernel", KERNELS)
def test_g_dgdr_zeta_vs_mpmath(kernel):
    for s in [1e-3, 0.01, 0.1, 0.3, 0.7, 1.0, 1.5, 2.5, 4.0, 6.0, 8.4, 8.6, 12.0, 40.0]:
        g, dg = oracle.g_dgdr(kernel, s)
        G, DG = hp_oracle.g_dgdr(kernel, s)
        # absolute accuracy relative to the O(1) scale of g (the reference's own
        # formulas cancel for small s, e.g. g = erf - aux)
        assert abs(g - float(G)) <= 4e-16 * max(1.0, abs(float(G))) + 2e-16
        assert abs(dg - float(DG)) <= _tail_tol(float(DG), 1e-15)
        z, Z = oracle.zeta(kernel, s), float(hp_oracle.zeta(kernel, s))
        # exp(-x) carries a relative error ~ |x| eps (argument rounding) in any FP64 evaluation
        assert abs(z - Z) <= _tail_tol(Z, 2e-15)
    assert oracle.zeta("singular", 0.0) == 1.0 and oracle.zeta("singular", 1e-300) == 0.0


# ------------------------------------------- scripts/check_fmm.jl known answers
BODIES = np.array([[0.4, 0.1], [0.1, -0.5], [-0.3, 0.2], [1 / 8, 1 / 8], [0.3, -0.4], [-0.1, -0.2],
                   [0.08, 0.5]])  # scripts/check_fmm.jl:90-98: rows x y z sigma Gx Gy Gz


def _u(xt, xs, gs):  # scripts/check_fmm.jl:39-47
    dx = xt - xs
    n = np.linalg.norm(dx)
    return 1 / 4 / np.pi / n**3 * np.array([-dx[1] * gs[2] + dx[2] * gs[1], -dx[2] * gs[0] + dx[0] * gs[2],
                                            -dx[0] * gs[1] + dx[1] * gs[0]])


def _duidxj(xt, xs, gs):  # scripts/check_fmm.jl:57-71
    x, y, z = xt - xs
    xy, yz, xz = x * y, y * z, x * z
    gx, gy, gz = gs
    n = np.sqrt(x * x + y * y + z * z)
    m = np.array([
        [3 * xy * gz - 3 * xz * gy, (2 * y**2 - x**2 - z**2) * gz - 3 * yz * gy, 3 * yz * gz - (2 * z**2 - x**2 - y**2) * gy],
        [3 * xz * gx - (2 * x**2 - y**2 - z**2) * gz, 3 * yz * gx - 3 * xy * gz, (2 * z**2 - x**2 - y**2) * gx - 3 * xz * gz],
        [(2 * x**2 - y**2 - z**2) * gy - 3 * xy * gx, 3 * xy * gy - (2 * y**2 - x**2 - z**2) * gx, 3 * xz * gy - 3 * yz * gx]]) / n**5
    return m / 4 / np.pi


def _two_particle_field():
    P = np.zeros((46, 2), order="F")
    for i in range(2):
        P[0:3, i] = BODIES[0:3, i]
        P[6, i] = BODIES[3, i]
        P[3:6, i] = BODIES[4:7, i]
    return P


def test_check_fmm_known_answers_singular():
    P = _two_particle_field()
    oracle.uj_direct(P, 2, "singular")
    for t, s in ((0, 1), (1, 0)):
        u = _u(BODIES[0:3, t], BODIES[0:3, s], BODIES[4:7, s])
        Jm = _duidxj(BODIES[0:3, t], BODIES[0:3, s], BODIES[4:7, s])
        assert relerr(P[9:12, t], u) < 1e-14
        # reference J flat index i + 3 j = du_i/dx_j -> column-major of the 3x3
        assert relerr(P[15:24, t].reshape(3, 3, order="F"), Jm) < 1e-14
        # stretching (classic scheme) J * Gamma_target, scripts/check_fmm.jl:73-88
        st = Jm @ BODIES[4:7, t]
        Jo = P[15:24, t]
        st_o = np.array([Jo[k] * BODIES[4, t] + Jo[k + 3] * BODIES[5, t] + Jo[k + 6] * BODIES[6, t] for k in range(3)])
        assert relerr(st_o, st) < 1e-14


@pytest.mark.parametrize("kernel", ["gaussian", "gaussianerf", "winckelmans"])
def test_check_fmm_known_answers_regularised_far(kernel):
    """r/sigma = 6.7 here: every regularised family is within its tail of the singular answer"""
    P = _two_particle_field()
    oracle.uj_direct(P, 2, kernel)
    tail = {"gaussian": 1e-15, "gaussianerf": 1e-7, "winckelmans": 5e-3}[kernel]
    for t, s in ((0, 1), (1, 0)):
        u = _u(BODIES[0:3, t], BODIES[0:3, s], BODIES[4:7, s])
        assert relerr(P[9:12, t], u) < tail + 1e-14


# ------------------------------------------------------ mpmath golden vectors
@pytest.mark.parametrize("kernel", KERNELS)
def test_oracle_uj_vs_golden(kernel):
    P = make_field()
    n = P.shape[1]
    oracle.uj_direct(P, n, kernel, nthreads=1)
    assert relerr(P[9:12], GOLD[f"U_{kernel}"]) < 1e-14
    assert relerr(P[15:24], GOLD[f"J_{kernel}"]) < 1e-14
    # threaded form gives bit-identical results (same per-target source order)
    P2 = make_field()
    oracle.uj_direct(P2, n, kernel, nthreads=4)
    assert np.array_equal(P, P2)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("transposed", [True, False])
def test_oracle_sfs_vs_golden(kernel, transposed):
    P = make_field()
    n = P.shape[1]
    P[42] = GOLD["static"]
    P[15:24] = GOLD["Jin"]
    P[39:42] = 7.0  # must survive on static targets, be accumulated on elsewhere
    oracle.estr_direct(P, n, kernel, transposed)
    ref = GOLD[f"SFS_{kernel}_{'T' if transposed else 'C'}"]
    st = GOLD["static"] != 0
    assert np.all(P[39:42, st] == 7.0)
    assert relerr(P[39:42, ~st] - 7.0, ref[:, ~st]) < 2e-13 if np.abs(ref).max() > 0 else True


def test_oracle_zeta_direct_vs_mpmath():
    P = make_field(16)
    P[42, 3] = 1.0
    oracle.zeta_direct(P, 16, "gaussianerf", nthreads=1)
    import mpmath as mp
    for i in (0, 3, 9):
        acc = [mp.mpf(0)] * 3
        for j in range(16):
            d = [mp.mpf(float(P[k, i])) - mp.mpf(float(P[k, j])) for k in range(3)]
            r = mp.sqrt(sum(v * v for v in d))
            sg = mp.mpf(float(P[6, j]))
            z = hp_oracle.zeta("gaussianerf", r / sg) / sg**3
            for k in range(3):
                acc[k] += mp.mpf(float(P[3 + k, j])) * z
        assert relerr(P[15:18, i], [float(v) for v in acc]) < 1e-14


# ------------------------------------------------- reset / static / accumulate
def test_reset_and_static_rules():
    n = 12
    P = make_field(n)
    P[42, [2, 5]] = 1.0
    rng = np.random.default_rng(0)
    P[9:27] = rng.standard_normal((18, n))
    P[39:42] = rng.standard_normal((3, n))
    prior = P.copy(order="F")
    fresh = make_field(n)
    oracle.uj_direct(fresh, n, "winckelmans", sfs=False, reset=True)
    # reset=True: non-static rows are zeroed then accumulated; static accumulate on the old values
    oracle.uj_direct(P, n, "winckelmans", sfs=False, reset=True, reset_sfs=False)
    st = prior[42, :n] != 0
    assert np.allclose(P[9:12, ~st], fresh[9:12, ~st], rtol=0, atol=0)
    assert np.all(P[12:15, ~st] == 0) and np.all(P[24:27, ~st] == 0)
    assert np.allclose(P[9:12, st], prior[9:12, st] + fresh[9:12, st], rtol=1e-15)
    assert np.array_equal(P[12:15, st], prior[12:15, st])
    assert np.array_equal(P[39:42], prior[39:42])  # SFS untouched without sfs / reset_sfs
    # reset=False: everything accumulates
    P2 = prior.copy(order="F")
    oracle.uj_direct(P2, n, "winckelmans", reset=False)
    assert np.allclose(P2[15:24], prior[15:24] + fresh[15:24], rtol=1e-14)
    # reset_sfs zeroes non-static SFS only
    P3 = prior.copy(order="F")
    oracle.uj_direct(P3, n, "winckelmans", reset=True, reset_sfs=True)
    assert np.all(P3[39:42, ~st] == 0) and np.array_equal(P3[39:42, st], prior[39:42, st])


def test_self_pair_and_coincident_particles_are_skipped():
    P = np.zeros((46, 3), order="F")
    P[0:3, 0] = P[0:3, 1] = [0.1, 0.2, 0.3]   # coincident pair: r2 == 0 -> skipped
    P[0:3, 2] = [0.5, 0.2, 0.3]
    P[3:6] = [[0.1, 0.2, 0.3], [0.0, -0.1, 0.2], [0.3, 0.0, 0.1]]
    P[6] = 0.2
    for k in KERNELS:
        Q = P.copy(order="F")
        oracle.uj_direct(Q, 3, k)
        assert np.all(np.isfinite(Q))
        # particles 0 and 1 see only particle 2, identically
        assert np.array_equal(Q[9:12, 0], Q[9:12, 1]) and np.array_equal(Q[15:24, 0], Q[15:24, 1])


def test_empty_field():
    P = np.zeros((46, 4), order="F")
    oracle.uj_direct(P, 0, "gaussianerf", sfs=True)
    assert not P.any()


# ---------------------------------------------- physics regression (reference test)
class _VpmStub:
    """the physics driver only needs the host mirror (no GPU): import lazily"""


@pytest.mark.parametrize("integration", ["euler", "rk3"])
def test_single_ring_speed_within_2_percent(integration, vpm):
    def UJ(pf, reset=True):
        oracle.uj_direct(pf.particles, pf.np, pf.kernel.name, reset=reset, transposed=pf.transposed)
    U_vpm, U_ana = physics.run_single_ring(vpm, UJ, integration)
    err = (U_vpm - U_ana) / U_ana
    assert abs(err) < 0.02, (U_vpm, U_ana, err)  # test/runtests_singlevortexring.jl:143


@pytest.mark.parametrize("integration", ["euler", "rungekutta3"])
def test_oracle_step_equals_numpy_harness(integration, vpm):
    """two independent restatements of the reference integrators (C in oracle/, numpy in
    tests/physics.py) agree bit for bit"""
    pf = vpm.fields.ring_field(Nphi=40, nc=1, kernel=vpm.winckelmans)
    a = pf.particles.copy(order="F")

    def UJ(p, reset=True, **kw):
        oracle.uj_direct(p.particles, p.np, p.kernel.name, reset=reset, **kw)
    step = physics.euler_step if integration == "euler" else physics.rk3_step
    step(pf, 1e-2, UJ, f=0.0, g=0.2, relax=True)
    oracle.field_step(a, pf.np, "winckelmans", 1e-2, integration=integration, f=0.0, g=0.2, relax=True)
    assert np.array_equal(a[0:7], pf.particles[0:7])


@pytest.mark.parametrize("f,g,sfs", [(0.0, 0.0, False), (0.0, 0.2, False), (0.0, 0.2, "dynamic")])
def test_leapfrog_rings_vs_borisov_ode(vpm, f, g, sfs):
    """test/runtests_leapfrog.jl rows 1-3 (cVPM, rVPM, rVPM + DynamicSFS; RK3, corrected
    Pedrizzetti relaxation, 2 x 100 particles, 350 steps) with the oracle's integrator: end
    state within the reference's own tolerances (:169) of the Borisov-2013 ODE solution.
    (The reference runs these through UJ_fmm; here the sums are direct.)"""
    pf, dt, Rcross = physics.leapfrog_setup(vpm)
    for _ in range(physics.LEAPFROG["nsteps"]):
        oracle.field_step(pf.particles, pf.np, "winckelmans", dt, integration="rungekutta3", f=f, g=g, sfs=sfs,
                          relaxation="correctedpedrizzetti", relax=True, rlxf=0.3, alpha=0.667, sfs_rlxf=0.005,
                          minC=0.0, maxC=1.0, nthreads=4)
    Z1e, Z2e, R1e, R2e = physics.leapfrog_errors(vpm, pf, dt * physics.LEAPFROG["nsteps"], Rcross)
    assert abs(Z1e) < 0.05 and abs(Z2e) < 0.03 and abs(R1e) < 0.03 and abs(R2e) < 0.03, (Z1e, Z2e, R1e, R2e)


def test_oracle_rbf_recovers_strengths(vpm):
    """rbf_conjugategradient restatement (src/FLOWVPM_viscous.jl:309-478): fed the vorticity the
    field represents, the CG converges and returns the original strengths"""
    pf = vpm.fields.ring_field(Nphi=60, nc=1, kernel=vpm.gaussianerf, R=1.0, Rcross=0.15, sigma=0.12)
    n, P = pf.np, pf.particles
    G0 = P[3:6, :n].copy()
    oracle.zeta_direct(P, n, "gaussianerf")
    W = P[15:18, :n].copy()
    P[33:36, :n] = W
    it, res = oracle.rbf_conjugategradient(P, n, "gaussianerf", itmax=30, tol=1e-6)
    assert 2 <= it < 30 and res.max() < 1e-6
    assert relerr(P[3:6, :n], G0) < 1e-4
    oracle.zeta_direct(P, n, "gaussianerf")
    assert relerr(P[15:18, :n], W) < 1e-6


def test_oracle_pse_is_inviscid_plus_volumes(vpm):
    """in v4.0.3 nothing accumulates into the PSE rows (src/FLOWVPM_particlefield.jl:482-489 only zero them), so a
    step with ParticleStrengthExchange equals the inviscid step except for the recomputed volumes
    (src/FLOWVPM_viscous.jl:264-268)"""
    pf = vpm.fields.ring_field(Nphi=40, nc=1, kernel=vpm.winckelmans)
    a, b = pf.particles.copy(order="F"), pf.particles.copy(order="F")
    kw = dict(integration="rungekutta3", f=0.0, g=0.2, sfs=False, relaxation="pedrizzetti", relax=True)
    oracle.field_step(a, pf.np, "winckelmans", 1e-2, viscous=dict(scheme="pse", nu=0.5), **kw)
    oracle.field_step(b, pf.np, "winckelmans", 1e-2, **kw)
    rows = [r for r in range(46) if r != 7]
    assert np.array_equal(a[rows], b[rows])
    assert np.allclose(a[7, :pf.np], 4 / 3 * np.pi * a[6, :pf.np] ** 3, rtol=1e-15) and not np.array_equal(a[7], b[7])


def test_oracle_zeta_fmm_is_the_near_field_of_zeta_direct(vpm):
    """zeta_fmm (src/FLOWVPM_viscous.jl:523-558) = zeta_direct with the far field neglected: on a compact field whose
    leaf lists (theta = 0.4) cover every pair with a visible zeta the two agree to rounding; the reference's
    zeta_fmm ACCUMULATES on J[1:3] (it has no reset), which the restatement reproduces; a list entry (a, b) makes
    the bodies of leaf b receive from the bodies of leaf a"""
    pf = vpm.fields.ring_field(Nphi=60, nc=1, kernel=vpm.gaussianerf, R=1.0, Rcross=0.15, sigma=0.12)
    n = pf.np
    P = pf.particles.copy(order="F")
    oracle.zeta_direct(P, n, "gaussianerf")
    W = P[15:18, :n].copy()
    Q = pf.particles.copy(order="F")
    Q[15:18, :n] = 0.0
    z = oracle.cs_zeta_fmm(ncrit=20, theta=0.4, reset=False)
    z._eval(Q.ctypes.data, Q.shape[0], n, pf.kernel.id)
    assert relerr(Q[15:18, :n], W) < 1e-13
    z._eval(Q.ctypes.data, Q.shape[0], n, pf.kernel.id)          # no reset: the second call doubles the rows
    assert relerr(Q[15:18, :n], 2 * W) < 1e-13
    zr = oracle.cs_zeta_fmm(ncrit=20, theta=0.4, reset=True)
    zr._eval(Q.ctypes.data, Q.shape[0], n, pf.kernel.id)
    assert relerr(Q[15:18, :n], W) < 1e-13
    # a tight acceptance criterion drops visible pairs: then the list sum is NOT the direct sum
    zt = oracle.cs_zeta_fmm(ncrit=5, theta=4.0, reset=True)
    zt._eval(Q.ctypes.data, Q.shape[0], n, pf.kernel.id)
    assert relerr(Q[15:18, :n], W) > 1e-6


def test_oracle_rbf_with_zeta_fmm_context(vpm):
    """the RBF restatement with cs.zeta = zeta_fmm installed (what CoreSpreading(nu, sgm0, zeta_fmm) runs): with the
    rows zeroed per evaluation it converges like the zeta_direct one; with the reference's accumulating zeta_fmm
    the same iteration does not (residuals grow) -- the behaviour the device path is checked against"""
    pf = vpm.fields.ring_field(Nphi=60, nc=1, kernel=vpm.gaussianerf, R=1.0, Rcross=0.15, sigma=0.12)
    n = pf.np
    G0 = pf.particles[3:6, :n].copy()
    out = {}
    for reset in (True, False):
        P = pf.particles.copy(order="F")
        with oracle.cs_zeta_fmm(ncrit=20, theta=0.4, reset=reset) as z:
            P[15:18, :n] = 0.0
            z._eval(P.ctypes.data, P.shape[0], n, pf.kernel.id)
            P[33:36, :n] = P[15:18, :n]
            it, res = oracle.rbf_conjugategradient(P, n, "gaussianerf", itmax=6, tol=1e-6, iterror=False)
            assert z.calls == it + 2
        out[reset] = (it, res.max(), relerr(P[3:6, :n], G0))
    assert out[True][0] < 6 and out[True][1] < 1e-6 and out[True][2] < 1e-4
    assert out[False][0] == 6 and out[False][1] > 1e-2
    # the callback is gone after the context: the default is zeta_direct again
    P = pf.particles.copy(order="F")
    oracle.zeta_direct(P, n, "gaussianerf")
    P[33:36, :n] = P[15:18, :n]
    it, res = oracle.rbf_conjugategradient(P, n, "gaussianerf", itmax=30, tol=1e-6)
    assert res.max() < 1e-6
