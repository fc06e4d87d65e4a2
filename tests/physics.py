"""TEST HARNESS ONLY: restatement of the reference's time integration for the
physics regression of test/runtests_singlevortexring.jl rows 1-2 (Euler and RK3,
cVPM = ReformulatedVPM(0,0), inviscid, no SFS, Pedrizzetti relaxation every step).

  _euler                 src/FLOWVPM_timeintegration.jl:103-173
  rungekutta3            src/FLOWVPM_timeintegration.jl:388-461
  update_particle_states src/FLOWVPM_timeintegration.jl:463-534
  relax_pedrizzetti      src/FLOWVPM_relaxation.jl:41-60
  calc_rings_weighted!   examples/vortexrings/vortexrings_functions.jl:284-322

The UJ evaluation is a callable `UJ(pfield, reset=True)` so the same driver runs
on the oracle (CPU tests) and on libvpm_cuda (GPU tests)."""
import numpy as np


def _stretch(J, G, transposed):
    if transposed:
        return np.stack([J[3 * k] * G[0] + J[3 * k + 1] * G[1] + J[3 * k + 2] * G[2] for k in range(3)])
    return np.stack([J[k] * G[0] + J[k + 3] * G[1] + J[k + 6] * G[2] for k in range(3)])


def relax_pedrizzetti(P, n, rlxf=0.3):
    J = P[15:24, :n]
    G = P[3:6, :n]
    w = np.stack([J[5] - J[7], J[6] - J[2], J[1] - J[3]])
    nrmw = np.sqrt((w * w).sum(axis=0))
    nrmG = np.sqrt((G * G).sum(axis=0))
    ok = nrmw != 0
    G[:, ok] = (1 - rlxf) * G[:, ok] + rlxf * nrmG[ok] * w[:, ok] / nrmw[ok]


def euler_step(pf, dt, UJ, f=0.0, g=0.0, relax=True):
    n = pf.np
    P = pf.particles
    UJ(pf, reset=True)
    P[0:3, :n] += dt * P[9:12, :n]
    G = P[3:6, :n]
    S = _stretch(P[15:24, :n], G, pf.transposed)
    Gn2 = (G * G).sum(axis=0)
    Z = np.where(Gn2 > 0, (f + g) / (1 + 3 * f) * (S * G).sum(axis=0) / np.where(Gn2 > 0, Gn2, 1), 0.0)
    sig = P[6, :n].copy()
    G += dt * (S - 3 * Z * G)
    P[6, :n] -= dt * sig * Z
    if relax:
        relax_pedrizzetti(P, n)


def rk3_step(pf, dt, UJ, f=0.0, g=0.0, relax=True, sfs=False, zeta0=1.0):
    """rungekutta3 for ReformulatedVPM{f,g}; with sfs=True the UJ call also evaluates the SFS
    term and the per-particle coefficient C (row 37) multiplies it (ConstantSFS-style: the
    dynamic procedure of src/FLOWVPM_subfilterscale.jl:447-673 is not restated)."""
    n = pf.np
    P = pf.particles
    M = P[27:36, :n]
    M[:] = 0
    for a, b in ((0.0, 1 / 3), (-5 / 9, 15 / 16), (-153 / 128, 8 / 15)):
        if sfs:
            UJ(pf, reset=True, reset_sfs=True, sfs=True)
        else:
            UJ(pf, reset=True)
        G = P[3:6, :n]
        M[0:3] = a * M[0:3] + dt * P[9:12, :n]
        P[0:3, :n] += b * M[0:3]
        S = _stretch(P[15:24, :n], G, pf.transposed)
        Gn2 = (G * G).sum(axis=0)
        eps_ = P[36, :n] * P[39:42, :n] * P[6, :n] ** 3 / zeta0 if sfs else 0.0
        Z = (f + g) / (1 + 3 * f) * (S * G).sum(axis=0)
        if sfs:
            Z = Z - f / (1 + 3 * f) * (eps_ * G).sum(axis=0)
        Z = np.where(Gn2 > 0, Z / np.where(Gn2 > 0, Gn2, 1), 0.0)
        M[3:6] = a * M[3:6] + dt * (S - 3 * Z * G - eps_)
        M[7] = a * M[7] - dt * (P[6, :n] * Z)
        G += b * M[3:6]
        P[6, :n] += b * M[7]
    if relax:
        UJ(pf, reset=True)
        relax_pedrizzetti(P, n)


def ring_centroid_weighted(pf):
    n = pf.np
    P = pf.particles
    w = np.sqrt((P[3:6, :n] ** 2).sum(axis=0))
    return (P[0:3, :n] * w).sum(axis=1) / w.sum()


def run_single_ring(vpm, UJ, integration, nsteps=50, Nphi=100, nc=0, R=1.0, Rtot=2.0, beta=0.5, faux=0.25):
    """test/runtests_singlevortexring.jl:40-143; returns (U_vpm, U_ana)."""
    Rcross = 0.15 * R
    sigma = Rcross
    Uref = vpm.fields.Uring(1.0, R, Rcross, beta)
    dt = (Rtot / Uref) / nsteps
    pf = vpm.ParticleField(vpm.fields.number_particles(Nphi, nc), kernel=vpm.winckelmans, transposed=True)
    vpm.fields.addvortexring(pf, 1.0, R, 1.0, faux * Rcross, Nphi, nc, sigma)
    step = euler_step if integration == "euler" else rk3_step
    t = 0.0
    for _ in range(nsteps):
        step(pf, dt, UJ)
        t += dt
    Zc = ring_centroid_weighted(pf)
    return float(np.linalg.norm(Zc) / t), float(Uref)


# ---------------------------------------------------------------------------------------
# Leapfrogging rings: the analytic reference of test/runtests_leapfrog.jl
# (examples/vortexrings/vortexrings_postprocessing.jl:190-305): Borisov, Kilin & Mamaev 2013
# coaxial thin rings, forward Euler with dt = 1e-4, finite-difference dG/dR, dG/dZ (h = 1e-5),
# dynamica = false, Delta = 0 (Winckelmans kernel).
# ---------------------------------------------------------------------------------------
def _G(z, r, zt, rt):
    from scipy.special import ellipe, ellipk
    k = np.sqrt(4 * r * rt / ((z - zt) ** 2 + (r + rt) ** 2))
    return np.sqrt(r * rt) / (2 * np.pi) * ((2 / k - k) * ellipk(k * k) - 2 / k * ellipe(k * k))


def analytic_coaxialrings(Gammas, Rs, Zs, a_s, tend, dt=1e-4, h=1e-5):
    n = len(Gammas)
    R, Z = np.array(Rs, dtype=float), np.array(Zs, dtype=float)
    nst = int(np.floor(tend / dt + 1e-12))
    steps = [dt] * nst + ([tend - nst * dt] if tend - nst * dt > 0 else [])
    for this_dt in steps:
        dR, dZ = np.zeros(n), np.zeros(n)
        for i in range(n):
            dZ[i] = Gammas[i] / (4 * np.pi * R[i]) * (np.log(8 * R[i] / a_s[i]) - 0.5)
            for j in range(n):
                if i == j:
                    continue
                g0 = _G(Z[i], R[i], Z[j], R[j])
                dR[i] -= 1 / R[i] * Gammas[j] * (_G(Z[i] + h, R[i], Z[j], R[j]) - g0) / h
                dZ[i] += 1 / R[i] * Gammas[j] * (_G(Z[i], R[i] + h, Z[j], R[j]) - g0) / h
        R, Z = R + dR * this_dt, Z + dZ * this_dt
    return R, Z


def rings_weighted(pf, nrings, n_per_ring):
    """calc_rings_weighted! (examples/vortexrings/vortexrings_functions.jl:284-322): centroid
    and radius of each ring weighted by |Gamma|"""
    out = []
    for ri in range(nrings):
        P = pf.particles[:, ri * n_per_ring:(ri + 1) * n_per_ring]
        w = np.sqrt((P[3:6] ** 2).sum(axis=0))
        Zc = (P[0:3] * w).sum(axis=1) / w.sum()
        Rr = (w * np.sqrt(((P[0:3] - Zc[:, None]) ** 2).sum(axis=0))).sum() / w.sum()
        out.append((Zc, Rr))
    return out


LEAPFROG = dict(nsteps=350, R=0.7906, dZ=0.7906, Nphi=100, nc=0, beta=0.5)


def leapfrog_setup(vpm):
    """test/runtests_leapfrog.jl:38-75: two coaxial rings, Rcross = 0.1 R, sigma = Rcross"""
    c = LEAPFROG
    Rcross = 0.10 * c["R"]
    pf = vpm.fields.ring_field(Nphi=c["Nphi"], nc=c["nc"], R=c["R"], Rcross=Rcross, sigma=Rcross, rings=2, dZ=c["dZ"],
                               kernel=vpm.winckelmans)
    Uref = vpm.fields.Uring(1.0, c["R"], Rcross, c["beta"])
    dt = ((c["nsteps"] / 1000) / Uref) / c["nsteps"]
    return pf, dt, Rcross


def leapfrog_errors(vpm, pf, tend, Rcross):
    c = LEAPFROG
    (Z1, R1), (Z2, R2) = rings_weighted(pf, 2, vpm.fields.number_particles(c["Nphi"], c["nc"]))
    Ra, Za = analytic_coaxialrings([1.0, 1.0], [c["R"], c["R"]], [0.0, c["dZ"]], [Rcross, Rcross], tend)
    return ((Z1[2] - Za[0]) / Za[0], (Z2[2] - Za[1]) / Za[1], (R1 - Ra[0]) / Ra[0], (R2 - Ra[1]) / Ra[1])
