// vpm_abi.cu -- C ABI of libvpm_cuda.so (see include/vpm_cuda.h).
//
// Host-side orchestration only: buffer management, strided host<->device
// copies of the rows of ParticleField.particles the path touches, launch
// planning and the per-call timing record.  All arithmetic of the hot path is
// in vpm_kernels.cuh.  There is deliberately no CPU implementation here: if a
// CUDA call fails the entry point returns VPM_ECUDA.
#include "../../include/vpm_cuda.h"

#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <chrono>
#include <string>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "vpm_kernels.cuh"
#include "vpm_kernels_f32.cuh"
#include "vpm_kernels_tab.cuh"
#include "vpm_leaf.cuh"
#include "vpm_leaf_f32.cuh"
#include "vpm_csr.cuh"
#include "vpm_tree.cuh"
#include "vpm_step.cuh"

using namespace vpm;

// The host side is one translation unit assembled from these parts, in dependency order
// (each part relies on the ones before it; none is a standalone header):
#include "vpm_host_base.cuh"    // handle, device buffers, error plumbing
#include "vpm_host_sweeps.cuh"  // launch plans, launchers, U/J and SFS sweeps
#include "vpm_host_hook1.cuh"   // row copies, Hook 1 on one device
#include "vpm_host_multi.cuh"   // NCCL, UJ_direct on G devices
#include "vpm_host_lists.cuh"   // direct_list regrouping, device-built leaf lists
#include "vpm_host_field.cuh"   // device-resident field (f-1)
// ============================================================== C ABI
#include "vpm_abi_core.cuh"     // lifetime, Hooks 1 and 2, device-pointer entry points
#include "vpm_abi_lists.cuh"    // Hook 3 and the leaf-list entry points
#include "vpm_abi_field.cuh"    // resident field / time step
#include "vpm_abi_instr.cuh"    // timing, probes
