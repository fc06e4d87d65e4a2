"""Synthetic particle fields for the configurations of BASELINE.json.

Geometry restated from the reference's example generators so that the oracle,
the CPU baseline and the GPU path are fed identical, realistic inputs
(parity is measured on identical input arrays, not on identical generators):

  addvortexring  examples/vortexrings/vortexrings_functions.jl:73-227 (AR = 1;
                 the cell volumes HCubature integrates to 1e-8 are analytic here)
  addannulus     examples/roundjet/roundjet_functions.jl:52-131 (AR = 1)
  cloud          scripts/benchmark_fmm2.jl:11-34 (create_pfield): box 1x1x7 lattice,
                 Gamma = (0,0,1/N) + U(-.05,.05)^3, sigma = 0.65 d (1 + U(-.05,.05)),
                 plus a position jitter <= 0.25 d
  leaf lists     uniform-octree leaves + near-field list by the multipole
                 acceptance criterion theta (src/FLOWVPM_particlefield.jl:28-36):
                 a stand-in for FastMultipole's tree (external) to exercise the
                 FMM near-field hook with realistic (target leaf, source leaf) pairs
"""
import numpy as np

from .particlefield import ParticleField, kernel_default


def number_particles(Nphi, nc, extra_nc=0):
    """examples/vortexrings/vortexrings_functions.jl:34"""
    return int(Nphi * (1 + 8 * sum(range(1, nc + extra_nc + 1))))


def Uring(circulation, R, Rcross, beta):
    """Analytic self-induced velocity of an inviscid ring (vortexrings_functions.jl:47)"""
    return circulation / (4 * np.pi * R) * (np.log(8 * R / Rcross) - beta)


def addvortexring(pfield, circulation, R, AR, Rcross, Nphi, nc, sigma, *, extra_nc=0, O=None,
                  Oaxis=None):
    if AR != 1:
        raise NotImplementedError("only circular rings (AR = 1) are generated here")
    O = np.zeros(3) if O is None else np.asarray(O, dtype=float)
    Oaxis = np.eye(3) if Oaxis is None else np.asarray(Oaxis, dtype=float)
    rl = Rcross / (2 * nc + 1)
    dphi = 2 * np.pi / Nphi
    omega = circulation / (np.pi * Rcross**2)
    eps = np.finfo(float).eps
    zvec = np.array([0.0, 0.0, 1.0])

    def add(X, Gamma, vol, crcltn):
        pfield.add_particle(Oaxis @ X + O, Oaxis @ Gamma, sigma, vol=vol, circulation=crcltn)

    for N in range(Nphi):
        phi1, phi2 = dphi * N, dphi * (N + 1)
        phic = 0.5 * (phi1 + phi2)
        Xc = np.array([R * np.sin(phic), R * np.cos(phic), 0.0])
        T = -np.array([np.cos(phic), -np.sin(phic), 0.0])
        B = np.cross(zvec, T)
        for n in range(nc + extra_nc + 1):
            if n == 0:
                vol = (phi2 - phi1) * np.pi * R * rl**2
                Gamma = omega * vol * T
                length = R * (phi2 - phi1)
                add(Xc, Gamma, vol, np.linalg.norm(Gamma) / length)
            else:
                rc = (1 + 12 * n**2) / (6 * n) * rl
                r1, r2 = (2 * n - 1) * rl, (2 * n + 1) * rl
                ncells = 8 * n
                dtht = 2 * np.pi / ncells
                for j in range(ncells):
                    t1, t2 = dtht * j, dtht * (j + 1)
                    tc = 0.5 * (t1 + t2)
                    vol = (phi2 - phi1) * (R * (r2**2 - r1**2) / 2 * (t2 - t1)
                                           + (r2**3 - r1**3) / 3 * (np.sin(t2) - np.sin(t1)))
                    X = Xc + rc * np.cos(tc) * B + rc * np.sin(tc) * zvec
                    Gamma = omega * vol * T if n <= nc else eps * T
                    length = (R + rc * np.cos(tc)) * (phi2 - phi1)
                    add(X, Gamma, vol, np.linalg.norm(Gamma) / length)


def addannulus(pfield, circulation, R, Nphi, sigma, area, *, O=None, static=False):
    O = np.zeros(3) if O is None else np.asarray(O, dtype=float)
    dphi = 2 * np.pi / Nphi
    for N in range(Nphi):
        phic = dphi * (N + 0.5)
        X = np.array([R * np.sin(phic), R * np.cos(phic), 0.0])
        T = -np.array([np.cos(phic), -np.sin(phic), 0.0])
        length = R * dphi
        pfield.add_particle(X + O, circulation * length * T, sigma, vol=area * length,
                            circulation=circulation, static=static)


def ring_field(Nphi=100, nc=3, R=1.0, Rcross=0.15, sigma=None, circulation=1.0, kernel=None,
               rings=1, dZ=0.0, **kw):
    """C1 (one ring, Nphi=100, nc=3 -> 4900 particles) / C2 (two coaxial rings)."""
    sigma = Rcross if sigma is None else sigma
    n1 = number_particles(Nphi, nc)
    pf = ParticleField(n1 * rings, kernel=kernel or kernel_default, **kw)
    for ri in range(rings):
        addvortexring(pf, circulation, R, 1.0, Rcross, Nphi, nc, sigma, O=[0.0, 0.0, dZ * ri])
    return pf


def cloud_arrays(N, seed=20240607, Lx=1.0, Ly=1.0, Lz=7.0, overlap=1.3, jitter=0.25,
                 circulation=1.0):
    """X (3,N), Gamma (3,N), sigma (N,) of the C4 cloud, exactly N particles."""
    rng = np.random.Generator(np.random.PCG64(seed))
    d = (Lx * Ly * Lz / N) ** (1.0 / 3.0)
    nx, ny = max(1, int(np.ceil(Lx / d))), max(1, int(np.ceil(Ly / d)))
    nz = int(np.ceil(N / (nx * ny)))
    idx = np.arange(N)
    iz, rem = np.divmod(idx, nx * ny)
    iy, ix = np.divmod(rem, nx)
    X = np.empty((3, N))
    X[0] = (ix + 0.5) * d
    X[1] = (iy + 0.5) * d
    X[2] = (iz + 0.5) * d
    X += (rng.random((3, N)) - 0.5) * (2 * jitter * d)
    Gamma = np.zeros((3, N))
    Gamma[2] = circulation / N
    Gamma += (rng.random((3, N)) - 0.5) / 10
    sigma = d / 2 * overlap * (1 + (rng.random(N) - 0.5) / 10)
    del nz
    return X, Gamma, sigma


def cloud_field(N, seed=20240607, kernel=None, static_fraction=0.0, R=np.float64, **kw):
    """C4: jittered-lattice random vortex particle cloud with N particles."""
    X, Gamma, sigma = cloud_arrays(N, seed)
    pf = ParticleField(N, R, kernel=kernel or kernel_default, **kw)
    P = pf.particles
    P[0:3, :N] = X
    P[3:6, :N] = Gamma
    P[6, :N] = sigma
    P[8, :N] = 1.0
    pf.np = N
    if static_fraction > 0:
        rng = np.random.Generator(np.random.PCG64(seed + 1))
        P[42, :N] = (rng.random(N) < static_fraction).astype(P.dtype)
    return pf


def jet_field(n_target=300_000, seed=7, kernel=None, static_fraction=0.1, **kw):
    """C3: jet-like column of stacked annuli (addannulus geometry), sigma = 2.4 dx,
    ~10% static particles (the reference's jet inflow particles are static,
    examples/roundjet/roundjet_simulation.jl)."""
    Nphi = 100
    nann = max(1, n_target // Nphi)
    Rjet = 0.5
    dx = 2 * np.pi * Rjet / Nphi
    sigma = 2.4 * dx
    rng = np.random.Generator(np.random.PCG64(seed))
    pf = ParticleField(nann * Nphi, kernel=kernel or kernel_default, **kw)
    nstatic = int(static_fraction * nann)
    for k in range(nann):
        R = Rjet * (1 + 0.3 * np.tanh((k - nann / 3) / (nann / 6 + 1)) * rng.random())
        addannulus(pf, 1.0 * (1 + 0.1 * (rng.random() - 0.5)), R, Nphi, sigma, dx * dx,
                   O=[0.01 * rng.standard_normal(), 0.01 * rng.standard_normal(), k * dx],
                   static=(k < nstatic))
    return pf


def random_results(pfield, seed=3, scale=1.0):
    """Fill U, J, SFS rows with reproducible nonzero values (to test accumulate / reset rules)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    P = pfield.live()
    for rows in (slice(9, 27), slice(39, 42)):
        P[rows] = scale * rng.standard_normal(P[rows].shape)


# ----------------------------------------------------------------- leaf lists
def build_leaf_lists(X, sigma, ncrit=64, theta=0.4, handle=None):
    """Leaf lists for the FMM near-field hook, built ON THE GPU (vpm_leaflists_build, csrc/vpm_tree.cuh):
    dict(sort_index, leaf_begin, leaf_end, direct_list).  Stand-in for FastMultipole's tree
    (src/FLOWVPM_UJ.jl:90-101) when only positions and core sizes are at hand."""
    from .particlefield import ParticleField
    from .uj import leaf_lists
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[1]
    pf = ParticleField(max(n, 1))
    pf.particles[0:3, :n] = X
    pf.particles[6, :n] = np.asarray(sigma, dtype=np.float64)
    pf.np = n
    return leaf_lists(pf, ncrit=ncrit, theta=theta, handle=handle)
