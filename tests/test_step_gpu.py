"""the device-resident time step (vpm_field_*) against the CPU restatement of
the reference's euler / rungekutta3 (oracle.field_step), and the reference's own physics
assertion (ring speed within 2 %, test/runtests_singlevortexring.jl:143) run entirely on the GPU."""
import numpy as np
import pytest

from helpers import relerr
from oracle import oracle
import physics

pytestmark = pytest.mark.gpu

ROWS = {"X": slice(0, 3), "Gamma": slice(3, 6), "sigma": slice(6, 7), "U": slice(9, 12), "J": slice(15, 24),
        "M": slice(27, 36), "C": slice(36, 37), "SFS": slice(39, 42)}


def make_field(vpm, kernel):
    pf = vpm.fields.ring_field(Nphi=60, nc=1, kernel=kernel, R=1.0, Rcross=0.15)
    pf.particles[42, 5:pf.np:37] = 1.0          # a few static particles
    pf.particles[36, :pf.np] = 0.3              # stale C values must be overwritten by ConstantSFS
    return pf


@pytest.mark.parametrize("integration", ["euler", "rungekutta3"])
@pytest.mark.parametrize("sfs,clip", [(False, False), (True, False), (True, True)])
@pytest.mark.parametrize("relaxation", ["pedrizzetti", "correctedpedrizzetti", None])
def test_step_matches_oracle(vpm, handle, integration, sfs, clip, relaxation):
    kernel = vpm.gaussianerf if sfs else vpm.winckelmans
    pf = make_field(vpm, kernel)
    ref = pf.particles.copy(order="F")
    kw = dict(integration=integration, f=0.0, g=0.2, Uinf=(0.01, -0.02, 0.03), sfs=sfs, Cs=0.8,
              clip_backscatter=clip, relaxation=relaxation, relax=True, rlxf=0.3)
    rf = vpm.ResidentField(pf)
    dt = 2e-2
    for _ in range(2):
        rf.nextstep(dt, **kw)
        oracle.field_step(ref, pf.np, kernel.name, dt, transposed=pf.transposed, **kw)
    rf.download()
    for name, rows in ROWS.items():
        if not sfs and name in ("SFS", "C"):
            assert np.array_equal(pf.particles[rows], ref[rows])
            continue
        assert relerr(pf.particles[rows, :pf.np], ref[rows, :pf.np]) < 1e-11, name
    st = ref[42, :pf.np] != 0
    assert np.array_equal(pf.particles[0:9, :pf.np][:, st], ref[0:9, :pf.np][:, st])  # static particles do not move


@pytest.mark.parametrize("integration", ["euler", "rungekutta3"])
@pytest.mark.parametrize("force_positive,clip,alpha", [(False, False, 0.667), (True, True, 0.9), (False, True, 0.999)])
def test_dynamic_sfs_step_matches_oracle(vpm, handle, integration, force_positive, clip, alpha):
    """DynamicSFS pseudo-3-level procedure (src/FLOWVPM_subfilterscale.jl:447-673) on the device:
    three steps so that the Lagrangian averages <Gamma.L>, <Gamma.m> (C rows 2:3) evolve"""
    pf = vpm.fields.cloud_field(1500, kernel=vpm.gaussianerf, static_fraction=0.05, seed=31)
    ref = pf.particles.copy(order="F")
    kw = dict(integration=integration, f=0.0, g=0.2, sfs="dynamic", clip_backscatter=clip, relaxation="pedrizzetti",
              relax=True, rlxf=0.3, alpha=alpha, sfs_rlxf=0.3, minC=0.0, maxC=1.0, force_positive=force_positive)
    rf = vpm.ResidentField(pf)
    dt = 1e-3
    for _ in range(3):
        rf.nextstep(dt, **kw)
        oracle.field_step(ref, pf.np, "gaussianerf", dt, transposed=True, **kw)
    rf.download()
    rows = dict(ROWS, C=slice(36, 39))
    for name, r in rows.items():
        assert relerr(pf.particles[r, :pf.np], ref[r, :pf.np]) < 1e-9, name
    assert np.abs(ref[36, :pf.np]).max() > 0 and np.abs(ref[36, :pf.np]).max() <= 1.0


@pytest.mark.parametrize("sfs", ["constant", "dynamic"])
@pytest.mark.parametrize("directional,magnitude", [(True, False), (False, True), (True, True)])
def test_sfs_controls_match_oracle(vpm, handle, sfs, directional, magnitude):
    """control_directional / control_magnitude (src/FLOWVPM_subfilterscale.jl:300-397)"""
    pf = vpm.fields.cloud_field(1200, kernel=vpm.gaussianerf, static_fraction=0.05, seed=41)
    ref = pf.particles.copy(order="F")
    kw = dict(integration="rungekutta3", f=0.1, g=0.2, sfs=sfs, Cs=0.7, clip_backscatter=True, relaxation=None,
              relax=False, alpha=0.9, sfs_rlxf=0.4, control_directional=directional, control_magnitude=magnitude)
    rf = vpm.ResidentField(pf)
    dt = 1e-3
    for k in range(3):
        deltat = pf.t / pf.nt if pf.nt > 0 else 0.0
        rf.nextstep(dt, **kw)
        oracle.field_step(ref, pf.np, "gaussianerf", dt, transposed=True, deltat=deltat, **kw)
    rf.download()
    for name, r in dict(ROWS, C=slice(36, 39)).items():
        assert relerr(pf.particles[r, :pf.np], ref[r, :pf.np]) < 1e-9, name


def _vorticity_ring(vpm):
    # no static particles: their fixed strengths enter every basis evaluation of the CG, which
    # makes the reference's RBF system inhomogeneous (it does not converge there either)
    return vpm.fields.ring_field(Nphi=60, nc=1, kernel=vpm.gaussianerf, R=1.0, Rcross=0.15, sigma=0.12)


def test_rbf_conjugategradient_matches_oracle(vpm, handle):
    """rbf_conjugategradient (src/FLOWVPM_viscous.jl:309-478) with zeta_direct: the target is the
    vorticity the field itself represents, so the CG must give the strengths back"""
    pf = _vorticity_ring(vpm)
    n = pf.np
    G0 = pf.particles[3:6, :n].copy()
    oracle.zeta_direct(pf.particles, n, "gaussianerf")
    pf.particles[33:36, :n] = pf.particles[15:18, :n]      # M[7:9] <- target vorticity
    ref = pf.particles.copy(order="F")
    it_ref, res_ref = oracle.rbf_conjugategradient(ref, n, "gaussianerf", itmax=30, tol=1e-6)
    rf = vpm.ResidentField(pf)
    it, res = rf.rbf_conjugategradient(itmax=30, tol=1e-6)
    rf.download()
    assert it == it_ref and it >= 2
    assert relerr(pf.particles[3:6, :n], ref[3:6, :n]) < 1e-9
    assert relerr(pf.particles[27:33, :n], ref[27:33, :n]) < 1e-7          # solution and residual rows of M
    assert relerr(pf.particles[3:6, :n], G0) < 1e-4                        # strengths recovered
    with pytest.raises(vpm.VpmError):                                      # the reference throws when it does not converge
        rf.upload()
        rf.rbf_conjugategradient(itmax=1, tol=1e-12, iterror=True)


@pytest.mark.parametrize("integration", ["euler", "rungekutta3"])
def test_corespreading_step_matches_oracle(vpm, handle, integration):
    """viscousdiffusion(pfield, CoreSpreading, dt) (src/FLOWVPM_viscous.jl:152-223): core growth every
    stage, and an RBF reset when sigma/sgm0 reaches beta (forced every second step here)"""
    pf = _vorticity_ring(vpm)
    pf.particles[7, :pf.np] = 4 / 3 * np.pi * 0.05**3      # particle volumes feed the RBF initial guess
    ref = pf.particles.copy(order="F")
    vis = dict(nu=2e-3, sgm0=0.12, beta=1.02, itmax=20, tol=1e-4, iterror=True)
    vis_ref = dict(vis, t_sgm=0.0)
    kw = dict(integration=integration, f=0.0, g=0.2, sfs=False, relaxation="pedrizzetti", relax=True)
    rf = vpm.ResidentField(pf)
    resets = 0
    for k in range(4):
        rf.nextstep(5e-2, viscous=vis, **kw)
        oracle.field_step(ref, pf.np, "gaussianerf", 5e-2, transposed=True, viscous=vis_ref, **kw)
        assert abs(rf.t_sgm - vis_ref["t_sgm"]) < 1e-15
        resets += vis_ref["t_sgm"] == 0.0
    assert 1 <= resets < 4
    rf.download()
    for name, r in ROWS.items():
        if name in ("SFS", "C"):
            continue
        assert relerr(pf.particles[r, :pf.np], ref[r, :pf.np]) < 1e-8, name
    with pytest.raises(vpm.VpmError):   # CoreSpreading is only compatible with gaussianerf (src/FLOWVPM.jl:265-267)
        pf.kernel = vpm.winckelmans
        rf.nextstep(5e-2, viscous=vis, **kw)


# ---- CoreSpreading(nu, sgm0, zeta_fmm): what the reference's own tests and examples construct
# (test/runtests_singlevortexring.jl:27, runtests_leapfrog.jl:26)
@pytest.mark.parametrize("method", ["fmm", "fmm_reset"])
def test_field_zeta_fmm_matches_list_oracle(vpm, handle, method):
    """zeta_fmm (src/FLOWVPM_viscous.jl:523-558) on the resident matrix: near field of device-built lists.
    "fmm" accumulates on J[1:3] as the reference does, "fmm_reset" zeroes them first; nothing else is written."""
    pf = vpm.fields.cloud_field(5000, kernel=vpm.gaussianerf, static_fraction=0.1, seed=31)
    vpm.fields.random_results(pf, scale=1e-2)
    n = pf.np
    ref = pf.particles.copy(order="F")
    z = oracle.cs_zeta_fmm(ncrit=40, theta=0.4, reset=method == "fmm_reset")
    z._eval(ref.ctypes.data, ref.shape[0], n, pf.kernel.id)
    rf = vpm.ResidentField(pf)
    try:
        rf.zeta_method(method, ncrit=40, theta=0.4)
        rf.zeta()
        rf.zeta()                       # second evaluation reuses the lists (X, sigma unchanged)
        if method == "fmm":             # ... and accumulates again
            z._eval(ref.ctypes.data, ref.shape[0], n, pf.kernel.id)
        rf.download()
    finally:
        rf.zeta_method("direct")
    assert relerr(pf.particles[15:18, :n], ref[15:18, :n]) < 1e-12
    keep = np.r_[0:15, 18:pf.particles.shape[0]]
    assert np.array_equal(pf.particles[keep][:, :n], ref[keep][:, :n])
    with pytest.raises(vpm.VpmError):
        rf.zeta_method("fmm", ncrit=0)


@pytest.mark.parametrize("method,itmax,tol", [("fmm_reset", 30, 1e-6), ("fmm", 3, 1e-12)])
def test_rbf_with_zeta_fmm_matches_oracle(vpm, handle, method, itmax, tol):
    """rbf_conjugategradient with cs.zeta = zeta_fmm.  With J[1:3] zeroed per evaluation the CG converges and
    gives the strengths back; with the reference's accumulating zeta_fmm the iteration is what it is -- the
    resident path must still follow the oracle's arithmetic step by step (3 iterations compared)."""
    pf = _vorticity_ring(vpm)
    n = pf.np
    G0 = pf.particles[3:6, :n].copy()
    reset = method == "fmm_reset"
    with oracle.cs_zeta_fmm(ncrit=20, theta=0.4, reset=reset) as z:
        pf.particles[15:18, :n] = 0.0
        z._eval(pf.particles.ctypes.data, pf.particles.shape[0], n, pf.kernel.id)
        pf.particles[33:36, :n] = pf.particles[15:18, :n]      # M[7:9] <- target vorticity
        ref = pf.particles.copy(order="F")
        it_ref, res_ref = oracle.rbf_conjugategradient(ref, n, "gaussianerf", itmax=itmax, tol=tol, iterror=False)
        assert z.calls == it_ref + 2                            # target, initial residual, one per iteration
    rf = vpm.ResidentField(pf)
    try:
        rf.zeta_method(method, ncrit=20, theta=0.4)
        it, res = rf.rbf_conjugategradient(itmax=itmax, tol=tol, iterror=False)
        rf.download()
    finally:
        rf.zeta_method("direct")
    assert it == it_ref and it >= 2
    assert relerr(pf.particles[3:6, :n], ref[3:6, :n]) < 1e-9
    assert relerr(pf.particles[27:33, :n], ref[27:33, :n]) < 1e-7
    if reset:
        assert relerr(pf.particles[3:6, :n], G0) < 1e-4                    # strengths recovered


def test_corespreading_step_with_zeta_fmm_matches_oracle(vpm, handle):
    """CoreSpreading(nu, sgm0, zeta_fmm) inside nextstep: the lists are rebuilt whenever sigma has changed
    (core growth every stage, reset to sgm0 before the RBF) and reused across the CG iterations"""
    pf = _vorticity_ring(vpm)
    pf.particles[7, :pf.np] = 4 / 3 * np.pi * 0.05**3
    ref = pf.particles.copy(order="F")
    vis = dict(nu=2e-3, sgm0=0.12, beta=1.02, itmax=20, tol=1e-4, iterror=True)
    vis_ref = dict(vis, t_sgm=0.0)
    kw = dict(integration="rungekutta3", f=0.0, g=0.2, sfs=False, relaxation="pedrizzetti", relax=True)
    rf = vpm.ResidentField(pf)
    resets = 0
    try:
        with oracle.cs_zeta_fmm(ncrit=20, theta=0.4, reset=True):
            for k in range(4):
                rf.nextstep(5e-2, viscous=dict(vis, zeta="fmm_reset", ncrit=20, theta=0.4), **kw)
                oracle.field_step(ref, pf.np, "gaussianerf", 5e-2, transposed=True, viscous=vis_ref, **kw)
                assert abs(rf.t_sgm - vis_ref["t_sgm"]) < 1e-15
                resets += vis_ref["t_sgm"] == 0.0
    finally:
        rf.zeta_method("direct")
    assert 1 <= resets < 4
    rf.download()
    for name, r in ROWS.items():
        if name in ("SFS", "C"):
            continue
        assert relerr(pf.particles[r, :pf.np], ref[r, :pf.np]) < 1e-8, name


@pytest.mark.parametrize("integration", ["euler", "rungekutta3"])
@pytest.mark.parametrize("recalculate_vols", [True, False])
def test_pse_step_matches_oracle(vpm, handle, integration, recalculate_vols):
    """viscousdiffusion(pfield, ParticleStrengthExchange, dt) (src/FLOWVPM_viscous.jl:257-298), the per-particle
    part: volumes recomputed from sigma, strengths updated with nu * PSE rows (static particles keep whatever
    their never-reset PSE rows hold and are skipped by the iterator)"""
    pf = make_field(vpm, vpm.winckelmans)
    pf.particles[42, 3:pf.np:11] = 1.0
    pf.particles[24:27, :pf.np] = 0.3            # visible only if the reset rules were wrong
    pf.particles[7, :pf.np] = 1.0
    ref = pf.particles.copy(order="F")
    vis = dict(scheme="pse", nu=1e-2, recalculate_vols=recalculate_vols)
    kw = dict(integration=integration, f=0.0, g=0.2, sfs=False, relaxation="pedrizzetti", relax=True)
    rf = vpm.ResidentField(pf)
    for _ in range(2):
        rf.nextstep(2e-2, viscous=vis, **kw)
        oracle.field_step(ref, pf.np, "winckelmans", 2e-2, transposed=True, viscous=vis, **kw)
    rf.download()
    free = pf.particles[42, :pf.np] == 0
    for name, r in ROWS.items():
        if name in ("SFS", "C"):
            continue
        assert relerr(pf.particles[r, :pf.np], ref[r, :pf.np]) < 1e-11, name
    assert relerr(pf.particles[7, :pf.np], ref[7, :pf.np]) < 1e-11    # volumes follow sigma, itself at 1e-12
    sg = pf.particles[6, :pf.np][free]
    if recalculate_vols:
        assert np.allclose(pf.particles[7, :pf.np][free], 4 / 3 * np.pi * sg**3, rtol=1e-15)
    else:
        assert np.all(pf.particles[7, :pf.np] == 1.0)
    assert np.all(pf.particles[7, :pf.np][~free] == 1.0)


def test_formulations_and_classic_scheme(vpm, handle):
    for f, g, transposed in ((0.0, 0.0, True), (0.5, 0.0, True), (0.25, 0.25, False)):
        pf = make_field(vpm, vpm.gaussianerf)
        pf.transposed = transposed
        ref = pf.particles.copy(order="F")
        # (no clipping with the classic scheme on this symmetric ring: Gamma . SFS is zero to
        # rounding there, so the sign test of clipping_backscatter is decided by noise)
        kw = dict(integration="rungekutta3", f=f, g=g, sfs=True, Cs=1.0, clip_backscatter=transposed,
                  relaxation="pedrizzetti", relax=True)
        rf = vpm.ResidentField(pf)
        rf.nextstep(1e-2, **kw)
        oracle.field_step(ref, pf.np, "gaussianerf", 1e-2, transposed=transposed, **kw)
        rf.download()
        for name, rows in ROWS.items():
            assert relerr(pf.particles[rows, :pf.np], ref[rows, :pf.np]) < 1e-11, (f, g, name)


def test_resident_uj_equals_hook1(vpm, handle):
    pf = vpm.fields.cloud_field(3000, kernel=vpm.winckelmans, static_fraction=0.1)
    vpm.fields.random_results(pf, scale=1e-3)
    a = pf.particles.copy(order="F")
    rf = vpm.ResidentField(pf)
    rf.UJ(sfs=True, reset=True, reset_sfs=True)
    rf.download()
    b = pf.particles.copy(order="F")
    pf.particles[:] = a
    vpm.UJ_direct(pf, sfs=True, reset=True, reset_sfs=True)
    assert np.array_equal(b, pf.particles)   # same kernels, same order: bit-identical


@pytest.mark.parametrize("integration", ["euler", "rungekutta3"])
def test_single_ring_speed_on_device(vpm, handle, integration):
    """test/runtests_singlevortexring.jl rows 1-2, every step on the GPU"""
    Nphi, nc, R, Rtot, beta, faux, nsteps = 100, 0, 1.0, 2.0, 0.5, 0.25, 50
    Rcross = 0.15 * R
    Uref = vpm.fields.Uring(1.0, R, Rcross, beta)
    dt = (Rtot / Uref) / nsteps
    pf = vpm.ParticleField(vpm.fields.number_particles(Nphi, nc), kernel=vpm.winckelmans)
    vpm.fields.addvortexring(pf, 1.0, R, 1.0, faux * Rcross, Nphi, nc, Rcross)
    rf = vpm.ResidentField(pf)
    for _ in range(nsteps):
        rf.nextstep(dt, integration=integration, f=0.0, g=0.0, relaxation="pedrizzetti", relax=True, rlxf=0.3)
    rf.download()
    U_vpm = np.linalg.norm(physics.ring_centroid_weighted(pf)) / (dt * nsteps)
    assert abs((U_vpm - Uref) / Uref) < 0.02


def test_step_needs_resident_field(vpm):
    h = vpm.Handle(1)
    try:
        sp = vpm._cabi.VpmStepParams()
        sp.kernel_id = 3
        import ctypes
        assert h.lib.vpm_field_step(h.ptr, ctypes.byref(sp)) == -6
        assert b"vpm_field_upload" in h.lib.vpm_last_error(h.ptr)
    finally:
        h.close()


@pytest.mark.parametrize("f,g,sfs", [(0.0, 0.0, False), (0.0, 0.2, False), (0.0, 0.2, "dynamic")])
def test_leapfrog_rings_on_device(vpm, handle, f, g, sfs):
    """test/runtests_leapfrog.jl rows 1-3 run entirely on the GPU (vpm_field_step): end state
    within the reference's tolerances of the Borisov-2013 ODE solution (:169)"""
    pf, dt, Rcross = physics.leapfrog_setup(vpm)
    rf = vpm.ResidentField(pf)
    for _ in range(physics.LEAPFROG["nsteps"]):
        rf.nextstep(dt, integration="rungekutta3", f=f, g=g, sfs=sfs, relaxation="correctedpedrizzetti", relax=True,
                    rlxf=0.3, alpha=0.667, sfs_rlxf=0.005, minC=0.0, maxC=1.0)
    rf.download()
    Z1e, Z2e, R1e, R2e = physics.leapfrog_errors(vpm, pf, dt * physics.LEAPFROG["nsteps"], Rcross)
    assert abs(Z1e) < 0.05 and abs(Z2e) < 0.03 and abs(R1e) < 0.03 and abs(R2e) < 0.03, (Z1e, Z2e, R1e, R2e)
