"""flowvpm.jl_b200 -- B200-native rVPM particle-to-particle path behind the
reference's UJ interface (byuflowlab/FLOWVPM.jl v4.0.3).

The product is csrc/libvpm_cuda.so (hand-written sm_100a CUDA behind the C ABI
of include/vpm_cuda.h); this package is the thin host mirror of the reference's
operator interface for that path.  The directory name has a dot, so import it
through `vpm_import.load()` at the repo root (module name `flowvpm_jl_b200`).
"""
from . import _cabi
from ._cabi import Handle, VpmError
from .particlefield import (ParticleField, Kernel, KERNELS, NFIELDS, kernel_default,
                            kernel_singular, kernel_gaussian, kernel_gaussianerf,
                            kernel_winckelmans, singular, gaussian, gaussianerf, winckelmans,
                            _reset_particles, _reset_particles_sfs,
                            X_INDEX, GAMMA_INDEX, SIGMA_INDEX, U_INDEX, VORTICITY_INDEX, J_INDEX,
                            PSE_INDEX, M_INDEX, C_INDEX, SFS_INDEX, STATIC_INDEX)
from .uj import (ResidentField, UJ_direct, UJ_nearfield, leaf_lists, direct_buffers, nearfield_device, fmm_nearfield_device, Estr_fmm, zeta_direct, zeta_fmm, get_handle,
                 set_handle,
                 source_system_to_buffer, buffer_to_target_system, ROW_POS, ROW_GRAD, ROW_HESS)
from . import fields
from . import io
from .io import save
