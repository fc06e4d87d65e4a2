"""Host-side mirror of the reference's ParticleField container for the hot path.

Only what the P2P path needs: the 46-row column-major particle matrix with the
reference's row map (src/FLOWVPM_particlefield.jl:239-252), add/get accessors
and the kernel / UJ / transposed settings (src/FLOWVPM_particlefield.jl:67-141).
Time integration, SFS procedures, viscous schemes and I/O stay in the
reference's Julia code (out of scope).
"""
import numpy as np

from . import _cabi

NFIELDS = 46  # src/FLOWVPM_particlefield.jl:11

# 0-based slices of the reference's 1-based index constants (:239-252)
X_INDEX = slice(0, 3)
GAMMA_INDEX = slice(3, 6)
SIGMA_INDEX = 6
VOL_INDEX = 7
CIRCULATION_INDEX = 8
U_INDEX = slice(9, 12)
VORTICITY_INDEX = slice(12, 15)
J_INDEX = slice(15, 24)
PSE_INDEX = slice(24, 27)
M_INDEX = slice(27, 36)
C_INDEX = slice(36, 39)
SFS_INDEX = slice(39, 42)
STATIC_INDEX = 42
U_PREV_INDEX = 43


class Kernel:
    """Kernel family tag (reference: struct Kernel, src/FLOWVPM_kernel.jl:23-28).

    The reference stores four Julia closures; here the family is an id that the
    C ABI takes (include/vpm_cuda.h VPM_KERNEL_*)."""

    def __init__(self, name, kernel_id):
        self.name = name
        self.id = kernel_id

    def __repr__(self):
        return f"Kernel({self.name})"


# singletons and aliases: src/FLOWVPM.jl:129-160
kernel_singular = Kernel("singular", _cabi.KERNEL_SINGULAR)
kernel_gaussian = Kernel("gaussian", _cabi.KERNEL_GAUSSIAN)
kernel_gaussianerf = Kernel("gaussianerf", _cabi.KERNEL_GAUSSIANERF)
kernel_winckelmans = Kernel("winckelmans", _cabi.KERNEL_WINCKELMANS)
kernel_default = kernel_gaussianerf
singular, gaussian, gaussianerf, winckelmans = (kernel_singular, kernel_gaussian,
                                                kernel_gaussianerf, kernel_winckelmans)
KERNELS = {k.name: k for k in (singular, gaussian, gaussianerf, winckelmans)}


class ParticleField:
    """ParticleField(maxparticles, R=float64; kernel, UJ, transposed, useGPU)

    `particles` is the (46, maxparticles) Fortran-ordered matrix whose memory
    layout equals Julia's Matrix{R}(46, maxparticles); column i is particle i.
    """

    def __init__(self, maxparticles, R=np.float64, *, kernel=kernel_default, UJ=None,
                 transposed=True, useGPU=1, np_=0):
        if R not in (np.float64, np.float32):
            raise ValueError("R must be float64 or float32")
        self.maxparticles = int(maxparticles)
        self.particles = np.zeros((NFIELDS, self.maxparticles), dtype=R, order="F")
        self.np = int(np_)
        self.nt = 0
        self.t = 0.0
        self.kernel = kernel
        self.transposed = bool(transposed)
        self.useGPU = int(useGPU)
        if UJ is None:
            from .uj import UJ_direct
            UJ = UJ_direct
        self.UJ = UJ

    # --- container API (src/FLOWVPM_particlefield.jl:167-205, 257-333) ---
    def get_np(self):
        return self.np

    def add_particle(self, X, Gamma, sigma, *, vol=0.0, circulation=1.0, C=0.0, static=False):
        if self.np == self.maxparticles:
            raise RuntimeError(f"PARTICLE OVERFLOW. Max number of particles {self.maxparticles}"
                               " has been reached")  # :171-173
        i = self.np
        p = self.particles
        p[:, i] = 0
        p[X_INDEX, i] = X
        p[GAMMA_INDEX, i] = Gamma
        p[SIGMA_INDEX, i] = sigma
        p[VOL_INDEX, i] = vol
        p[CIRCULATION_INDEX, i] = abs(circulation)
        p[C_INDEX, i] = C
        p[STATIC_INDEX, i] = float(static)
        self.np += 1

    def remove_particle(self, i):
        """Swap-remove, as the reference (src/FLOWVPM_particlefield.jl:401-419); i is 0-based."""
        if i < 0 or i >= self.np:
            raise IndexError(f"Requested removal of invalid particle index {i}")
        last = self.np - 1
        if i != last:
            self.particles[:, i] = self.particles[:, last]
        self.np -= 1

    def live(self):
        return self.particles[:, : self.np]

    def get_X(self, i=None):
        return self.live()[X_INDEX] if i is None else self.particles[X_INDEX, i]

    def get_Gamma(self, i=None):
        return self.live()[GAMMA_INDEX] if i is None else self.particles[GAMMA_INDEX, i]

    def get_sigma(self, i=None):
        return self.live()[SIGMA_INDEX] if i is None else self.particles[SIGMA_INDEX, i]

    def get_U(self, i=None):
        return self.live()[U_INDEX] if i is None else self.particles[U_INDEX, i]

    def get_J(self, i=None):
        return self.live()[J_INDEX] if i is None else self.particles[J_INDEX, i]

    def get_SFS(self, i=None):
        return self.live()[SFS_INDEX] if i is None else self.particles[SFS_INDEX, i]

    def get_static(self, i=None):
        return self.live()[STATIC_INDEX] != 0 if i is None else bool(self.particles[STATIC_INDEX, i])

    def get_W(self):
        """vorticity from J: (J6-J8, J7-J3, J2-J4), src/FLOWVPM_particlefield.jl:277-279"""
        J = self.get_J()
        return np.stack([J[5] - J[7], J[6] - J[2], J[1] - J[3]])


def _reset_particles(pfield):
    """src/FLOWVPM_particlefield.jl:464-490 (container API; the UJ call resets on the device)"""
    P = pfield.live()
    m = P[STATIC_INDEX] == 0
    for rows in (U_INDEX, VORTICITY_INDEX, J_INDEX, PSE_INDEX):
        P[rows][:, m] = 0


def _reset_particles_sfs(pfield):
    """src/FLOWVPM_particlefield.jl:492-507"""
    P = pfield.live()
    P[SFS_INDEX][:, P[STATIC_INDEX] == 0] = 0
