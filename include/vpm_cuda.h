/*
 * vpm_cuda.h -- C ABI of libvpm_cuda.so, the B200 (sm_100a) implementation of
 * the rVPM particle-to-particle hot path of byuflowlab/FLOWVPM.jl v4.0.3.
 *
 * The entry points are what the reference's Julia host code binds through
 * `ccall` (see INTEGRATION.md and flowvpm.jl_b200/julia/FLOWVPMCuda.jl).  Each
 * one cites the reference interface (path:line under the FLOWVPM.jl tree) it
 * replaces.  Plain pointers and sizes only; no C++ or torch types cross it.
 *
 * Conventions
 *  - every function returns VPM_OK (0) or a negative VPM_E* code and never
 *    throws or aborts; vpm_last_error() gives the message.
 *  - matrices are column-major as Julia owns them; "particles" is the
 *    ParticleField.particles matrix (nfields x maxparticles, nfields >= 43,
 *    46 in v4.0.3: src/FLOWVPM_particlefield.jl:11,134) of which the first np
 *    columns are live.  Row map (1-based, src/FLOWVPM_particlefield.jl:239-252):
 *    X 1:3, Gamma 4:6, sigma 7, U 10:12, vorticity 13:15, J 16:24, PSE 25:27,
 *    SFS 40:42, static 43.
 *  - a handle is not re-entrant; calls block until results are in host memory
 *    unless the function name says _device (then they are stream-ordered).
 *  - there is no CPU fallback: without a usable CUDA device vpm_create fails.
 */
#ifndef VPM_CUDA_H
#define VPM_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VPM_ABI_VERSION 1

/* status codes */
#define VPM_OK 0
#define VPM_EINVAL (-1)  /* bad argument */
#define VPM_ECUDA (-2)   /* CUDA runtime error (message has the CUDA string) */
#define VPM_ENOMEM (-3)  /* device or host allocation failed */
#define VPM_ENODEV (-4)  /* no usable CUDA device */
#define VPM_ENCCL (-5)   /* NCCL error / NCCL not loadable */
#define VPM_ESTATE (-6)  /* call sequence error (e.g. eval before upload) */

/* kernel families: src/FLOWVPM.jl:129-133, src/FLOWVPM_kernel.jl:44-84 */
#define VPM_KERNEL_SINGULAR 0
#define VPM_KERNEL_GAUSSIAN 1
#define VPM_KERNEL_GAUSSIANERF 2
#define VPM_KERNEL_WINCKELMANS 3

/* flags of the UJ call: keywords of UJ_direct, src/FLOWVPM_UJ.jl:21-25, and
 * pfield.transposed, src/FLOWVPM_particlefield.jl:84,109 */
#define VPM_FLAG_RESET 1       /* reset=true   : _reset_particles      */
#define VPM_FLAG_RESET_SFS 2   /* reset_sfs    : _reset_particles_sfs  */
#define VPM_FLAG_SFS 4         /* sfs=true     : Estr_direct! sweep    */
#define VPM_FLAG_TRANSPOSED 8  /* pfield.transposed (default true)     */
/* library-only switches (no reference counterpart) */
#define VPM_FLAG_NO_FARFIELD_SHORTCUT 16 /* gaussian/gaussianerf: evaluate exp/erf for
                                            every pair instead of g==1 beyond the cutoff */

#define VPM_FLAG_FP32 32  /* optional FP32-arithmetic U/J sweep (north star: <= 1e-5): hi/lo split
                             positions, packed f32x2 pair loop, FP64 sums across tiles; the SFS
                             sweep stays FP64.  Without it every entry point computes in FP64.
                             Accuracy contract (tests/test_fp32_gpu.py; errors relative to the field
                             maximum against the FP64 oracle): U, J and the stretching term (Gamma.grad')U
                             <= 1e-5 for the three regularised families on every tested field (measured
                             2e-7 .. 1e-6).  STATED DEVIATIONS from the 1e-5 bar:
                              * the SFS term of the same call is computed in FP64 but FROM the FP32-mode J;
                                (JT - JS).Gamma cancels, so it inherits J's error amplified: asserted at
                                1e-4 (measured ~2e-5), not 1e-5;
                              * the singular kernel on overlapping cores (a ring of unregularised particles,
                                outside any reference use): J 5.6e-6 but stretching 1.7e-5, asserted at 5e-5.
                             Callers that need 1e-5 on SFS use the default FP64 sweep. */

typedef struct vpm_handle vpm_handle;

/* ---- lifetime ---------------------------------------------------------- */
/* n_gpus >= 1 devices driven by this one process (the Julia caller is a single
 * process: src/FLOWVPM_utils.jl:87-133).  device_ids may be NULL (0..n-1).
 * Replaces nothing in the reference (ParticleField(...; useGPU) only stores a
 * flag, src/FLOWVPM_particlefield.jl:91,137). */
int vpm_create(vpm_handle **out, int n_gpus, const int *device_ids);
int vpm_destroy(vpm_handle *h);
/* message of the last failing call on h (h == NULL: last vpm_create failure) */
const char *vpm_last_error(const vpm_handle *h);
int vpm_abi_version(void);
int vpm_num_devices(const vpm_handle *h);

/* handle options (library-only) */
#define VPM_OPT_NEARFIELD_FP32 1 /* value != 0: the U/J leaf-list kernels (vpm_p2p_leafpairs, vpm_uj_nearfield)
                                    compute in FP32 arithmetic (split-precision positions, FP64 sums across
                                    tiles; errors ~1e-6 of the field maximum): the near field of UJ_fmm is added
                                    to a far field FastMultipole truncates at 1e-3..1e-6 anyway.  Default 0. */
#define VPM_OPT_UJ_VARIANT 2      /* tuning aid: launch variant "<T><unroll>" of the FP64 U/J pair kernel, one of
                                    11, 12, 21, 22 (T targets per thread, unroll of the source loop), or of the
                                    table kernel of the gaussian families: 31, 32 (384 threads per CTA, unroll 1 / 2),
                                    41, 42 (512 threads); 0 = automatic (default).  Any other value: VPM_EINVAL. */
#define VPM_OPT_SFS_VARIANT 3     /* the same for the SFS pair kernel: 10, 20 or 0 (T only) */
#define VPM_OPT_UJ_CONST 4        /* value != 0: constant-bank form of the U/J sweep (records copied 768 at a time
                                    into __constant__ memory, one launch per chunk).  Default 0. */
#define VPM_OPT_UJ_TABLE 5        /* gaussianerf / gaussian U/J sweep: 0 = automatic (default: the bank-replicated
                                    log-spaced table kernel, csrc/vpm_kernels_tab.cuh, when the field is large enough
                                    to fill the GPU with 1024-target CTAs AND, of 2048 deterministically sampled warps
                                    (32 consecutive targets x one source), >= 40 % see a pair inside the regularised
                                    range; that costs one small kernel and one stream synchronisation per sweep of
                                    >= 1e8 pairs), 1 = always, 2 = never (the round-1 kernels). */
#define VPM_OPT_SMALL_GRAPH 6      /* value != 0 (default): vpm_uj_direct on one device replays a captured CUDA graph for fields of
                                    <= 8 192 particles (the second call with the same matrix, np, kernel and flags
                                    captures it): one launch + one synchronisation per call instead of ~20 API calls.
                                    The per-phase times of vpm_get_timing are 0 for replayed calls (total_ms is the
                                    wall time of the call).  Measured: 157 -> 119 us per UJ_direct(sfs=true) call at
                                    200 particles from a pageable matrix, 147 -> 100 us from a page-locked one. */
int vpm_set_option(vpm_handle *h, int option, int value);

/* ---- Hook 1: the UJ slot ---------------------------------------------- */
/* UJ_direct(pfield; sfs, reset, reset_sfs): src/FLOWVPM_UJ.jl:21-37.
 * Uploads rows X, Gamma, sigma (+ static, + previous U/J/SFS when they must be
 * accumulated on), evaluates U, J (and SFS) on the GPU(s) and writes rows
 * 10:27 (and 40:42) back in place with the reference's reset-then-accumulate
 * semantics (src/FLOWVPM_particlefield.jl:464-511, src/FLOWVPM_fmm.jl:170-176). */
int vpm_uj_direct(vpm_handle *h, double *particles, int64_t nfields, int64_t np,
                  int kernel_id, int flags);
/* same for a Matrix{Float32} field (ParticleField(n, Float32)): Float32 storage,
 * converted on the device, FP64 arithmetic; tolerance 1e-5 */
int vpm_uj_direct_f32(vpm_handle *h, float *particles, int64_t nfields, int64_t np,
                      int kernel_id, int flags);
/* UJ_direct(source, target): src/FLOWVPM_UJ.jl:48-50 -- the field `source`
 * induces U, J on every particle of `target` (accumulated, no reset, no SFS). */
int vpm_uj_direct_st(vpm_handle *h, const double *source_particles, int64_t nfields_s,
                     int64_t np_s, double *target_particles, int64_t nfields_t, int64_t np_t,
                     int kernel_id);

/* staged form of Hook 1 for device residency between RK stages */
int vpm_upload_state(vpm_handle *h, const double *particles, int64_t nfields, int64_t np);
int vpm_eval(vpm_handle *h, int kernel_id, int flags);
int vpm_download_results(vpm_handle *h, double *particles, int64_t nfields, int64_t np,
                         int flags);

/* page-lock a host range the caller keeps alive (pfield.particles is allocated
 * once at maxparticles: src/FLOWVPM_particlefield.jl:134), so the strided
 * host<->device copies run at PCIe speed */
int vpm_pin_host(vpm_handle *h, void *ptr, size_t bytes);
int vpm_unpin_host(vpm_handle *h, void *ptr);

/* ---- Hook 2: FastMultipole pair-loop overload --------------------------- */
/* fmm.direct!(target_buffer, target_index, switch, source_system, source_buffer,
 * source_index): src/FLOWVPM_fmm.jl:102-168.  target buffer: tgt_ld x (>= t1)
 * with position at rows row_pos..+2, velocity accumulated into rows
 * row_grad..+2 and J into rows row_hess..+8 (0-based row offsets; FastMultipole
 * owns that layout, so they are arguments).  source buffer: 8 x (>= s1),
 * [x y z rho Gx Gy Gz sigma] (src/FLOWVPM_fmm.jl:62-71).  Half-open 0-based
 * ranges [t0,t1), [s0,s1).  want_U / want_J are the VS / GS switches. */
int vpm_p2p_buffers(vpm_handle *h, double *tgt_buf, int64_t tgt_ld, int64_t t0, int64_t t1,
                    int row_pos, int row_grad, int row_hess, const double *src_buf, int64_t s0,
                    int64_t s1, int kernel_id, int want_U, int want_J);

/* ---- Hook 3: FMM near-field device hook -------------------------------- */
/* fmm.nearfield_device!(target_system, target_indices::Vector{UnitRange}, switch, source_system,
 * source_indices) in the ONLY call shape the reference shows (src/FLOWVPM_gpu.jl:637-643, reached from
 * UJ_fmm with useGPU > 0, src/FLOWVPM_UJ.jl:97): both systems are ParticleFields, i.e. 46 x N matrices
 * whose columns the ranges index; target range k = [tgt_begin[k], tgt_end[k]) (0-based, half-open)
 * receives the sources of the source ranges src_offsets[k] .. src_offsets[k+1]-1, which is what the
 * reference's combine_source_indices / expand_source_indices (:554-602) produced per target leaf.
 * U (rows 10:12) and J (rows 16:24) of the target columns are accumulated on -- UJ_fmm has reset them
 * before (src/FLOWVPM_UJ.jl:75-80).  want_U / want_J are the VS / GS switches of `switch`.
 * Multi-GPU handles cut the target ranges into contiguous runs of equal work (they must then be
 * increasing and non-overlapping, as tree leaves are; otherwise device 0 does everything). */
int vpm_nearfield_ranges(vpm_handle *h, double *target_particles, int64_t nfields_t, int64_t np_t,
                         const int64_t *tgt_begin, const int64_t *tgt_end, int64_t n_tgt_ranges,
                         const double *source_particles, int64_t nfields_s, int64_t np_s,
                         const int64_t *src_begin, const int64_t *src_end, const int64_t *src_offsets,
                         int kernel_id, int want_U, int want_J);
/* The same near field on FastMultipole's own buffers (as in Hook 2, tree-sorted) with the whole
 * direct_list in one call: leaves are half-open body ranges; pair k = (pair_tgt[k], pair_src[k]).
 * Every listed pair is evaluated with fmm.direct!'s arithmetic. */
int vpm_p2p_leafpairs(vpm_handle *h, double *tgt_buf, int64_t tgt_ld, int64_t n_tgt,
                      int row_pos, int row_grad, int row_hess, const double *src_buf,
                      int64_t n_src, const int64_t *tgt_leaf_begin, const int64_t *tgt_leaf_end,
                      int64_t n_tgt_leaves, const int64_t *src_leaf_begin,
                      const int64_t *src_leaf_end, int64_t n_src_leaves, const int32_t *pair_tgt,
                      const int32_t *pair_src, int64_t n_pairs, int kernel_id, int want_U,
                      int want_J);
/* Estr_fmm!(pfield, pfield, target_tree, source_tree, direct_list):
 * src/FLOWVPM_subfilterscale_models.jl:94-188.  sort_index maps sorted body ->
 * particle column (0-based); no static filtering (as the reference). */
int vpm_estr_leafpairs(vpm_handle *h, double *particles, int64_t nfields, int64_t np,
                       const int64_t *tgt_sort_index, const int64_t *src_sort_index,
                       const int64_t *tgt_leaf_begin, const int64_t *tgt_leaf_end,
                       int64_t n_tgt_leaves, const int64_t *src_leaf_begin,
                       const int64_t *src_leaf_end, int64_t n_src_leaves,
                       const int32_t *pair_tgt, const int32_t *pair_src, int64_t n_pairs,
                       int kernel_id, int flags);

/* ---- device-built leaf lists --------------------------------- */
/* What the near-field hook needs from FastMultipole's tree (src/FLOWVPM_UJ.jl:90-101: sort
 * index, leaf body ranges, direct_list), built on the GPU from rows X and sigma of the field:
 * uniform cell grid with mean occupancy ~ncrit/2, stable sort by cell, one leaf per occupied
 * cell, leaf spheres padded by the largest core size, and the list of leaf pairs that fail the
 * MAC (r_i + r_j) <= theta d (theta = 0.4 in the reference: src/FLOWVPM_particlefield.jl:28-36).
 * The list comes out grouped by target leaf, sources in increasing order.  The lists stay
 * resident on the first device of the handle until the next build. */
int vpm_leaflists_build(vpm_handle *h, const double *particles, int64_t nfields, int64_t np, int64_t ncrit,
                        double theta, int64_t *n_leaves, int64_t *n_pairs);
/* copy the resident lists to the host (any pointer may be NULL): sort_index[np] maps sorted body ->
 * particle column (0-based), leaf_begin/leaf_end[n_leaves] are half-open sorted-body ranges,
 * pair k = (pair_tgt[k], pair_src[k]).  They are valid inputs of vpm_p2p_leafpairs,
 * vpm_estr_leafpairs and vpm_zeta_leafpairs. */
int vpm_leaflists_get(vpm_handle *h, int64_t *sort_index, int64_t *leaf_begin, int64_t *leaf_end,
                      int32_t *pair_tgt, int32_t *pair_src);
/* near-field half of UJ_fmm (src/FLOWVPM_UJ.jl:62-129) over the resident lists, entirely on the
 * device(s): rows 10:12 and 16:24 of every particle receive the sum over the sources of its
 * near-field leaves (RESET: _reset_particles first; otherwise added to what is there, e.g. a
 * far field evaluated by the host).  Flags: VPM_FLAG_RESET, VPM_FLAG_NO_FARFIELD_SHORTCUT.
 * The lists are valid for ONE particle configuration: the call compares a fingerprint of the X and sigma
 * rows it uploads with the one taken by vpm_leaflists_build and returns VPM_ESTATE when they differ
 * (strengths may change freely; after particles have moved, rebuild -- 23 ms at 2^24 on 8 GPUs). */
int vpm_uj_nearfield(vpm_handle *h, double *particles, int64_t nfields, int64_t np, int kernel_id, int flags);

/* ---- second P2P of the reference: the basis-function (vorticity) sum -------- */
/* zeta_direct(pfield): src/FLOWVPM_viscous.jl:488-515.  Rows 16:18 (J[1:3]) of EVERY
 * particle (static included) <- sum_j Gamma_j zeta(|x_i-x_j|/sigma_j)/sigma_j^3 (self term
 * included), overwriting them as the reference does. */
int vpm_zeta_direct(vpm_handle *h, double *particles, int64_t nfields, int64_t np, int kernel_id);
/* zeta_fmm(pfield): src/FLOWVPM_viscous.jl:523-558 -- the same sum restricted to the
 * near-field list of one tree, ADDED to rows 16:18 (the caller zeroes them).  For a list
 * entry (a, b) the bodies of leaf b receive from the bodies of leaf a, as in the reference. */
int vpm_zeta_leafpairs(vpm_handle *h, double *particles, int64_t nfields, int64_t np,
                       const int64_t *sort_index, const int64_t *leaf_begin, const int64_t *leaf_end,
                       int64_t n_leaves, const int32_t *pair_a, const int32_t *pair_b, int64_t n_pairs,
                       int kernel_id);

/* ---- device-resident time step ------------------------------- */
/* The whole particle matrix is mirrored on the device; a step runs the reference's
 * integrator there, so the matrix only crosses PCIe when the caller wants it back.
 *   rungekutta3 / update_particle_states, ReformulatedVPM{f,g}: src/FLOWVPM_timeintegration.jl:388-534
 *   euler / _euler:                                             src/FLOWVPM_timeintegration.jl:23-37,103-173
 *   relaxation (pedrizzetti / correctedpedrizzetti):            src/FLOWVPM_relaxation.jl:62-142
 *   ConstantSFS hook + clipping_backscatter:                    src/FLOWVPM_subfilterscale.jl:110-135,287-296
 *   DynamicSFS pseudo-3-level procedure:                        src/FLOWVPM_subfilterscale.jl:447-673
 * Covered: any (f, g) incl. cVPM (0,0) and rVPM (0,1/5); NoSFS / ConstantSFS / DynamicSFS
 * (pseudo3level, force_positive, clipping_backscatter, control_directional, control_magnitude);
 * Inviscid or CoreSpreading (with zeta_direct as its basis evaluation); constant Uinf.
 * ParticleStrengthExchange stays in the reference's Julia code (it needs the FMM tree). */
typedef struct vpm_step_params {
  double dt;
  double f, g;        /* ReformulatedVPM{f,g}: src/FLOWVPM_formulation.jl:23-37 */
  double Uinf[3];     /* pfield.Uinf(t), constant over the step */
  double Cs;          /* ConstantSFS model coefficient */
  double rlxf;        /* relaxation factor (0.3 in the reference's presets, src/FLOWVPM.jl:169-170) */
  double alpha;       /* DynamicSFS: test-filter scaling (default 0.667) */
  double sfs_rlxf;    /* DynamicSFS: Lagrangian-average relaxation (default 0.005) */
  double minC, maxC;  /* DynamicSFS: bounds of |C| (defaults 0, 1) */
  double deltat;      /* pfield.t / pfield.nt, used by control_magnitude; <= 0 when pfield.nt == 0 */
  double nu, sgm0;    /* CoreSpreading(nu, sgm0, zeta): src/FLOWVPM_viscous.jl:63-141; zeta = vpm_field_zeta_method */
  double cs_beta;     /* maximum core growth sigma/sgm0 before the RBF reset (default 1.5) */
  double cs_tol;      /* RBF tolerance (default 1e-3) */
  int32_t kernel_id;
  int32_t integration;      /* 0 euler, 1 rungekutta3 */
  int32_t relaxation;       /* 0 none, 1 pedrizzetti, 2 correctedpedrizzetti */
  int32_t relax;            /* apply relaxation in this step (run_vpm!'s `relax`, src/FLOWVPM_utils.jl:94-96) */
  int32_t sfs;              /* 0 NoSFS, 1 ConstantSFS, 2 DynamicSFS (pseudo3level) */
  int32_t clip_backscatter; /* clippings = (clipping_backscatter,) */
  int32_t transposed;       /* pfield.transposed */
  int32_t force_positive;   /* DynamicSFS: pseudo3level_positive */
  int32_t controls;         /* SFS controls: bit 0 control_directional, bit 1 control_magnitude
                               (src/FLOWVPM_subfilterscale.jl:300-397) */
  int32_t viscous;          /* 0 Inviscid, 1 CoreSpreading (requires the gaussianerf kernel,
                               src/FLOWVPM.jl:265-267), 2 / 3 ParticleStrengthExchange(nu) with / without
                               recalculate_vols: the per-particle part of src/FLOWVPM_viscous.jl:257-298 (the
                               reference's scheme errors unless UJ == UJ_fmm and, in v4.0.3, nothing ever
                               accumulates into the PSE rows it reads) */
  int32_t cs_itmax;         /* maximum RBF iterations (default 15) */
  int32_t cs_iterror;       /* fail when the RBF does not converge (default 1) */
} vpm_step_params;
int vpm_field_upload(vpm_handle *h, const double *particles, int64_t nfields, int64_t np);
int vpm_field_download(vpm_handle *h, double *particles, int64_t nfields, int64_t np);
/* UJ_direct(pfield; ...) on the resident matrix (flags as vpm_uj_direct) */
int vpm_field_uj(vpm_handle *h, int kernel_id, int flags);
/* nextstep's integration call: one euler / rungekutta3 step on the resident matrix */
int vpm_field_step(vpm_handle *h, const vpm_step_params *params);
/* rbf_conjugategradient(pfield, cs) on the resident matrix (src/FLOWVPM_viscous.jl:309-478), cs.zeta = the method
 * of vpm_field_zeta_method (zeta_direct unless set): target vorticity in M[7:9], new strengths in Gamma.
 * iterations / residuals (3) may be NULL. */
int vpm_field_rbf(vpm_handle *h, int kernel_id, int itmax, double tol, int iterror, int *iterations,
                  double *residuals);
/* CoreSpreading's third constructor argument `zeta` (src/FLOWVPM_viscous.jl:63-141) for the resident field: which
 * basis-function evaluation vpm_field_rbf and the CoreSpreading branch of vpm_field_step call.
 *   VPM_ZETA_DIRECT (default)  zeta_direct (:488-515): J[1:3] zeroed, all pairs.
 *   VPM_ZETA_FMM               zeta_fmm (:523-558), what the reference's tests and examples pass: the near field
 *                              of leaf lists with leaf size `ncrit` and acceptance `theta` (pfield.fmm), the far field
 *                              neglected, and J[1:3] ACCUMULATED on -- the reference's zeta_fmm does not zero them.
 *   VPM_ZETA_FMM_RESET         the same sums on zeroed J[1:3] (zeta_direct's contract).
 * The lists are built on the device from the resident X and sigma (the recipe of vpm_leaflists_build, not
 * FastMultipole's octree: which far pairs are dropped differs; measured on the C4 cloud at ncrit 50, theta 0.4 the
 * list sum equals zeta_direct to 3e-15) and are reused until X or sigma change: the RBF's CG iterations
 * cost one O(N ncrit) sweep each.  They replace the lists vpm_leaflists_build left on the handle.
 * ncrit / theta are ignored for VPM_ZETA_DIRECT. */
#define VPM_ZETA_DIRECT 0
#define VPM_ZETA_FMM 1
#define VPM_ZETA_FMM_RESET 2
int vpm_field_zeta_method(vpm_handle *h, int method, int64_t ncrit, double theta);
/* cs.zeta(pfield) on the resident matrix with the selected method: results in J[1:3] (rows 16:18) */
int vpm_field_zeta(vpm_handle *h, int kernel_id);
/* CoreSpreading.t_sgm (time since the last core reset) kept with the resident field:
 * set != 0 stores *t_sgm, otherwise it is returned; vpm_field_upload resets it to 0 */
int vpm_field_tsgm(vpm_handle *h, double *t_sgm, int set);

/* ---- device-pointer entry points (one process per GPU; the caller owns the
 * collective, e.g. an NCCL all-gather of the 8 x N source buffer) ---------- */
/* targets [t0,t1) of the same 8 x ns buffer; out12 is 12 x (t1-t0): U then J.
 * All pointers are device memory on the handle's first device; `stream` is a
 * cudaStream_t used as given (NULL = CUDA's default stream, e.g. torch's current
 * stream handle 0).  Overwrites out12.
 * Stream ordering: these two calls are asynchronous (they return once the kernels are queued on
 * `stream`).  They use scratch buffers of the handle; the library records an event after the last
 * kernel and makes the next user of that scratch -- another _device call on ANY stream, or any of
 * the synchronous entry points -- wait on it, so calls may be issued on different streams without
 * a synchronisation in between.  Outputs are valid once `stream` has reached this call's kernels.
 * Exception: with the gaussian / gaussianerf kernels on fields large enough for the table kernel
 * (VPM_OPT_UJ_TABLE = 0, >= ~12 500 targets) vpm_uj_device synchronises `stream` once to read the 4-byte
 * sample its kernel choice depends on -- set VPM_OPT_UJ_TABLE to 1 or 2 to keep it fully asynchronous
 * (e.g. inside a stream capture). */
int vpm_uj_device(vpm_handle *h, const double *d_src8, int64_t ns, int64_t t0, int64_t t1,
                  double *d_out12, int kernel_id, int flags, void *stream);
/* SFS sweep for targets [t0,t1): d_J9 is 9 x ns (final J of every particle),
 * d_static ns flags (nonzero = static; NULL = none); out3 is 3 x (t1-t0). */
int vpm_sfs_device(vpm_handle *h, const double *d_src8, const double *d_J9,
                   const double *d_static, int64_t ns, int64_t t0, int64_t t1, double *d_out3,
                   int kernel_id, int flags, void *stream);

/* ---- instrumentation ---------------------------------------------------- */
typedef struct vpm_timing {
  double h2d_ms, prep_ms, uj_ms, sfs_ms, finish_ms, d2h_ms, total_ms; /* last call */
  int64_t uj_pairs, sfs_pairs; /* ordered pairs visited by the two sweeps */
  int32_t kernel_launches;     /* kernels of this library launched by the last call */
  int32_t n_gpus;
} vpm_timing;
int vpm_get_timing(const vpm_handle *h, vpm_timing *out);

/* FP64 FMA pipe peak of device 0, measured with a dependent-chain-free DFMA
 * loop: the roofline denominator. */
int vpm_measure_dfma_peak(vpm_handle *h, double *dfma_per_s, double *elapsed_ms);
/* FP32 FMA issue rate of device 0 (scalar FMAs/s), the roofline denominator of VPM_FLAG_FP32:
 * mode 0 FFMA2 with loop-invariant operands, 1 FFMA2 reading three distinct registers,
 * 2 / 3 the same with scalar FFMA. */
int vpm_measure_ffma_peak(vpm_handle *h, int mode, double *fma_per_s, double *elapsed_ms);
/* evaluate one device math routine on an array (accuracy tests):
 * op 0 rsqrt, 1 exp, 2 (g, dg) of kernel `arg` at s, 3 zeta of kernel `arg` */
int vpm_test_math(vpm_handle *h, int op, int arg, const double *in, double *out, double *out2,
                  int64_t n);

/* Launch plan of a direct sweep (host arithmetic only: needs no handle and no GPU; for tests and tuning aids).
 * kind 0 U/J FP64, 1 SFS FP64, 2 U/J FP32, 3 U/J table kernel (gaussianerf / gaussian).
 * out[0..7] = targets per CTA, grid.x, number of source splits (grid.y), sources per split, tiles per split,
 * targets per thread, loop unroll, 1 if the field fills the machine for the table kernel (kind 3) else 0. */
int vpm_plan_query(int64_t n_targets, int64_t n_sources, int sm_count, int kind, int64_t *out);

#ifdef __cplusplus
}
#endif
#endif /* VPM_CUDA_H */
