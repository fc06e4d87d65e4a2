// vpm_abi.cu -- C ABI of libvpm_cuda.so (see include/vpm_cuda.h).
//
// Host-side orchestration only: buffer management, strided host<->device
// copies of the rows of ParticleField.particles the path touches, launch
// planning and the per-call timing record.  All arithmetic of the hot path is
// in vpm_kernels.cuh.  There is deliberately no CPU implementation here: if a
// CUDA call fails the entry point returns VPM_ECUDA.
#include "../../include/vpm_cuda.h"

#include <dlfcn.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "vpm_kernels.cuh"
#include "vpm_kernels_f32.cuh"
#include "vpm_leaf.cuh"
#include "vpm_csr.cuh"
#include "vpm_tree.cuh"
#include "vpm_step.cuh"

using namespace vpm;

namespace {

thread_local std::string g_create_error;

// rows of ParticleField.particles, 0-based (src/FLOWVPM_particlefield.jl:239-252)
enum { R_X = 0, R_G = 3, R_SIGMA = 6, R_U = 9, R_W = 12, R_J = 15, R_PSE = 24, R_SFS = 39,
       R_STATIC = 42, MIN_FIELDS = 43 };
// rows inside the device-side result block res18 = particle rows 9..26
enum { RES_ROWS = 18, RES_U = 0, RES_W = 3, RES_J = 6, RES_PSE = 15 };

struct Buf {
  void *p = nullptr;
  size_t cap = 0;
};

struct Dev {
  int id = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[10] = {};  // 0..5 phases of a call, 6..7 pair kernel, 8..9 cross-device ordering
  Buf in7, stat, res18, sfs3, rec, srec, partial, tbuf, sbuf, ibuf, jbuf, fld, scr, scr2, cubtmp, tree, tlist;
};

struct Plan {
  int T = 1;
  int unroll = 2;
  int nsplit = 1;
  int tiles_per_split = 1;
  int64_t pstride = 0;
  dim3 grid;
};

}  // namespace

struct vpm_handle {
  std::vector<Dev> devs;
  std::string err;
  vpm_timing timing{};
  int64_t np_resident = -1;   // particles held by the staged API
  bool resident_static = false;
  bool resident_prior = false;
  double *h_stat = nullptr;   // pinned staging for compact static flags
  size_t h_stat_cap = 0;
  double *h_stage = nullptr;  // pinned staging for the strided rows of a pageable host matrix
  size_t h_stage_cap = 0;     // (doubles)
  std::vector<std::pair<void *, size_t>> pinned;  // ranges page-locked by vpm_pin_host
  int launches = 0;
  int64_t fld_nf = 0, fld_np = -1;  // device mirror of the whole particle matrix (vpm_field_*)
  double fld_t_sgm = 0.0;           // CoreSpreading.t_sgm of the resident field
  int device_timing = 0;  // 1/2: ev[6..7] bracket the last _device U/J / SFS pair kernel
  // device-built leaf lists (vpm_leaflists_build), resident on device 0
  int64_t tree_np = -1, tree_nl = 0, tree_npairs = 0;
  // single-process multi-GPU (n_gpus > 1): NCCL communicators, one per device
  void *nccl_lib = nullptr;
  std::vector<void *> comms;
};

namespace {

int fail(vpm_handle *h, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_create_error = buf;
  return code;
}

#define CK(h, call)                                                                        \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess)                                                                 \
      return fail(h, e_ == cudaErrorMemoryAllocation ? VPM_ENOMEM : VPM_ECUDA,             \
                  "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

#define TRY(expr)            \
  do {                       \
    int rc_ = (expr);        \
    if (rc_ != VPM_OK) return rc_; \
  } while (0)

int ensure(vpm_handle *h, Buf &b, size_t bytes) {
  if (bytes <= b.cap && b.p) return VPM_OK;
  if (b.p) CK(h, cudaFree(b.p));
  b.p = nullptr;
  b.cap = 0;
  size_t want = std::max<size_t>(bytes + bytes / 4, 4096);
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    want = bytes;
    e = cudaMalloc(&b.p, want);
  }
  if (e != cudaSuccess)
    return fail(h, VPM_ENOMEM, "cudaMalloc of %zu bytes failed: %s", want, cudaGetErrorString(e));
  b.cap = want;
  return VPM_OK;
}

int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

bool valid_kernel(int k) { return k >= 0 && k <= 3; }

// Launch plan.  One CTA = kThreads * T targets x one contiguous range of source tiles.
// The grid is sized to ~16 waves of resident CTAs so that the tail of the last wave is a
// few percent at most; when the target count alone cannot provide that, the sources are
// split (>= 4 tiles per split) and the finish kernel adds the splits in order.
enum PlanKind { PLAN_UJ = 0, PLAN_SFS = 1, PLAN_UJ_F32 = 2 };
Plan make_plan(int64_t nt, int64_t ns, int sm_count, PlanKind kind) {
  Plan p;
  const int64_t ntiles = std::max<int64_t>(1, (ns + kTile - 1) / kTile);
  // two targets per thread (5 % faster in steady state) only when that still leaves enough
  // CTAs to fill the machine a few times; small fields take T = 1 and splits down to one tile
  const int64_t nblk2 = std::max<int64_t>(1, (nt + 2 * kThreads - 1) / (2 * kThreads));
  const bool big = nblk2 * std::max<int64_t>(1, ntiles / 4) >= (int64_t)sm_count * 4 * 4;
  p.T = big ? 2 : 1;
  p.unroll = p.T == 2 ? 1 : 2;
  if (const char *v = getenv(kind == PLAN_SFS ? "VPM_SFS_VARIANT" : "VPM_UJ_VARIANT")) {
    int x = atoi(v);  // tuning aid: "<T><unroll>", e.g. 12, 21, 22
    if (x / 10 >= 1 && x / 10 <= 2) { p.T = x / 10; p.unroll = x % 10; }
  }
  if (kind == PLAN_UJ_F32) p.T = 2;  // the FP32 sweep packs the two targets of a thread into f32x2
  const int min_tiles = big ? 4 : 1;
  const int ctas_per_sm = p.T == 1 ? 6 : 4;
  const int64_t nblk = std::max<int64_t>(1, (nt + (int64_t)kThreads * p.T - 1) / ((int64_t)kThreads * p.T));
  const int64_t want_ctas = (int64_t)sm_count * ctas_per_sm * 16;
  int64_t nsplit = (want_ctas + nblk - 1) / nblk;
  nsplit = std::max<int64_t>(1, std::min<int64_t>(nsplit, std::max<int64_t>(1, ntiles / min_tiles)));
  nsplit = std::min<int64_t>(nsplit, 1024);
  p.tiles_per_split = (int)((ntiles + nsplit - 1) / nsplit);
  p.nsplit = (int)((ntiles + p.tiles_per_split - 1) / p.tiles_per_split);
  p.pstride = round_up(std::max<int64_t>(nt, 1), 32);
  p.grid = dim3((unsigned)nblk, (unsigned)p.nsplit, 1);
  return p;
}

template <int K>
void launch_uj_T(const Plan &p, const UjArgs &a, cudaStream_t st) {
  switch (p.T * 10 + p.unroll) {
    case 11: uj_pairs_kernel<K, 1, 1><<<p.grid, kThreads, 0, st>>>(a); break;
    case 12: uj_pairs_kernel<K, 1, 2><<<p.grid, kThreads, 0, st>>>(a); break;
    case 21: uj_pairs_kernel<K, 2, 1><<<p.grid, kThreads, 0, st>>>(a); break;
    default: uj_pairs_kernel<K, 2, 2><<<p.grid, kThreads, 0, st>>>(a); break;
  }
}
template <int K>
void launch_ujc_T(const Plan &p, const UjConstArgs &a, cudaStream_t st) {
  const dim3 grid(p.grid.x, 1, 1);
  switch (p.T * 10 + p.unroll) {
    case 11: uj_const_kernel<K, 1, 1><<<grid, kThreads, 0, st>>>(a); break;
    case 12: uj_const_kernel<K, 1, 2><<<grid, kThreads, 0, st>>>(a); break;
    case 22: uj_const_kernel<K, 2, 2><<<grid, kThreads, 0, st>>>(a); break;
    default: uj_const_kernel<K, 2, 1><<<grid, kThreads, 0, st>>>(a); break;
  }
}
void launch_ujc(int kernel, const Plan &p, const UjConstArgs &a, cudaStream_t st) {
  switch (kernel) {
    case K_SING: launch_ujc_T<K_SING>(p, a, st); break;
    case K_GAUS: launch_ujc_T<K_GAUS>(p, a, st); break;
    case K_GERF: launch_ujc_T<K_GERF>(p, a, st); break;
    default: launch_ujc_T<K_WINCK>(p, a, st); break;
  }
}
void launch_uj(int kernel, const Plan &p, const UjArgs &a, cudaStream_t st) {
  switch (kernel) {
    case K_SING: launch_uj_T<K_SING>(p, a, st); break;
    case K_GAUS: launch_uj_T<K_GAUS>(p, a, st); break;
    case K_GERF: launch_uj_T<K_GERF>(p, a, st); break;
    default: launch_uj_T<K_WINCK>(p, a, st); break;
  }
}
void launch_uj_f32(int kernel, const Plan &p, const UjArgsF &a, cudaStream_t st) {
  const bool u1 = p.unroll == 1;
  switch (kernel) {
    case K_SING: u1 ? uj_pairs_kernel_f32<K_SING, 1><<<p.grid, kThreads, 0, st>>>(a) : uj_pairs_kernel_f32<K_SING, 2><<<p.grid, kThreads, 0, st>>>(a); break;
    case K_GAUS: uj_pairs_kernel_f32<K_GAUS, 1><<<p.grid, kThreads, 0, st>>>(a); break;
    case K_GERF: uj_pairs_kernel_f32<K_GERF, 1><<<p.grid, kThreads, 0, st>>>(a); break;
    default: u1 ? uj_pairs_kernel_f32<K_WINCK, 1><<<p.grid, kThreads, 0, st>>>(a) : uj_pairs_kernel_f32<K_WINCK, 2><<<p.grid, kThreads, 0, st>>>(a); break;
  }
}
template <int K>
void launch_sfs_T(const Plan &p, const SfsArgs &a, cudaStream_t st, int mode) {
  if (mode == MODE_ZETA) {
    if (p.T == 2) sfs_pairs_kernel<K, 2, MODE_ZETA><<<p.grid, kThreads, 0, st>>>(a);
    else sfs_pairs_kernel<K, 1, MODE_ZETA><<<p.grid, kThreads, 0, st>>>(a);
  } else {
    if (p.T == 2) sfs_pairs_kernel<K, 2, MODE_SFS><<<p.grid, kThreads, 0, st>>>(a);
    else sfs_pairs_kernel<K, 1, MODE_SFS><<<p.grid, kThreads, 0, st>>>(a);
  }
}
void launch_sfs(int kernel, const Plan &p, const SfsArgs &a, cudaStream_t st, int mode = MODE_SFS) {
  switch (kernel) {
    case K_SING: launch_sfs_T<K_SING>(p, a, st, mode); break;
    case K_GAUS: launch_sfs_T<K_GAUS>(p, a, st, mode); break;
    case K_GERF: launch_sfs_T<K_GERF>(p, a, st, mode); break;
    default: launch_sfs_T<K_WINCK>(p, a, st, mode); break;
  }
}

unsigned blocks_for(int64_t n, int threads) { return (unsigned)std::max<int64_t>(1, (n + threads - 1) / threads); }

// U/J sweep: records from `src` columns [s0, s0+ns), targets tpos[0..nt), partial
// sums left in d.partial; the caller runs the finish kernel with its own output.
int uj_sweep(vpm_handle *h, Dev &d, cudaStream_t st, int kernel, const double *tpos, int64_t tld,
             int64_t nt, SrcView src, int64_t s0, int64_t ns, int flags, Plan &plan,
             bool time_pairs = false) {
  const int64_t ns_pad = round_up(std::max<int64_t>(ns, 1), kTile);
  if (flags & VPM_FLAG_FP32) {
    // optional FP32-arithmetic sweep (vpm_kernels_f32.cuh): FP32 records, FP64 partial sums in
    // the same layout, so the finish kernels are shared with the FP64 sweep
    TRY(ensure(h, d.rec, (size_t)ns_pad * kRecF * sizeof(float)));
    plan = make_plan(nt, ns, d.sm_count, PLAN_UJ_F32);
    TRY(ensure(h, d.partial, (size_t)plan.nsplit * kAcc * plan.pstride * sizeof(double)));
    prep_uj_records_f32<<<blocks_for(ns_pad, 256), 256, 0, st>>>(src, s0, ns, ns_pad, kernel, (float *)d.rec.p);
    h->launches++;
    if (nt > 0 && ns > 0) {
      UjArgsF a;
      a.tpos = tpos; a.tld = tld; a.nt = nt;
      a.rec = (const float *)d.rec.p; a.ns = ns;
      a.tiles_per_split = plan.tiles_per_split;
      a.partial = (double *)d.partial.p; a.pstride = plan.pstride;
      a.shortcut = (flags & VPM_FLAG_NO_FARFIELD_SHORTCUT) ? 0 : 1;
      if (time_pairs) CK(h, cudaEventRecord(d.ev[6], st));
      launch_uj_f32(kernel, plan, a, st);
      if (time_pairs) CK(h, cudaEventRecord(d.ev[7], st));
      h->launches++;
    } else {
      plan.nsplit = 0;
    }
    CK(h, cudaGetLastError());
    return VPM_OK;
  }
  TRY(ensure(h, d.rec, (size_t)ns_pad * kRec * sizeof(double)));
  plan = make_plan(nt, ns, d.sm_count, PLAN_UJ);
  TRY(ensure(h, d.partial, (size_t)plan.nsplit * kAcc * plan.pstride * sizeof(double)));
  prep_uj_records<<<blocks_for(ns_pad, 256), 256, 0, st>>>(src, s0, ns, ns_pad, kernel,
                                                           (double *)d.rec.p);
  h->launches++;
  const char *cenv = getenv("VPM_UJ_CONST");
  const bool use_const = cenv && atoi(cenv) != 0;
  if (nt > 0 && ns > 0 && use_const) {
    plan.nsplit = 1;
    UjConstArgs a;
    a.tpos = tpos; a.tld = tld; a.nt = nt;
    a.partial = (double *)d.partial.p; a.pstride = plan.pstride;
    a.shortcut = (flags & VPM_FLAG_NO_FARFIELD_SHORTCUT) ? 0 : 1;
    if (time_pairs) CK(h, cudaEventRecord(d.ev[6], st));
    for (int64_t c0 = 0; c0 < ns; c0 += kCChunk) {
      a.n = (int)std::min<int64_t>(kCChunk, ns - c0);
      a.first = c0 == 0;
      CK(h, cudaMemcpyToSymbolAsync(c_rec, (const double *)d.rec.p + c0 * kRec,
                                    (size_t)a.n * kRec * sizeof(double), 0, cudaMemcpyDeviceToDevice, st));
      launch_ujc(kernel, plan, a, st);
      h->launches++;
    }
    if (time_pairs) CK(h, cudaEventRecord(d.ev[7], st));
  } else if (nt > 0 && ns > 0) {
    UjArgs a;
    a.tpos = tpos; a.tld = tld; a.nt = nt;
    a.rec = (const double *)d.rec.p; a.ns = ns;
    a.tiles_per_split = plan.tiles_per_split;
    a.partial = (double *)d.partial.p; a.pstride = plan.pstride;
    a.shortcut = (flags & VPM_FLAG_NO_FARFIELD_SHORTCUT) ? 0 : 1;
    if (time_pairs) CK(h, cudaEventRecord(d.ev[6], st));
    launch_uj(kernel, plan, a, st);
    if (time_pairs) CK(h, cudaEventRecord(d.ev[7], st));
    h->launches++;
  } else {
    plan.nsplit = 0;
  }
  CK(h, cudaGetLastError());
  return VPM_OK;
}

int sfs_sweep(vpm_handle *h, Dev &d, cudaStream_t st, int kernel, const double *tpos, int64_t tld,
              const double *tJ, int64_t jld, const int64_t *tindex, int64_t nt, SrcView src,
              const double *sJ, int64_t sjld, int sjoff, const double *stat, int64_t sld,
              const int64_t *sindex, int64_t ns, int flags, Plan &plan, bool time_pairs = false,
              int mode = MODE_SFS) {
  const int64_t ns_pad = round_up(std::max<int64_t>(ns, 1), kTile);
  TRY(ensure(h, d.srec, (size_t)ns_pad * kSfsRec * sizeof(double)));
  plan = make_plan(nt, ns, d.sm_count, PLAN_SFS);
  TRY(ensure(h, d.partial, (size_t)std::max(1, plan.nsplit) * kAcc * plan.pstride * sizeof(double)));
  const int transposed = (flags & VPM_FLAG_TRANSPOSED) ? 1 : 0;
  prep_sfs_records<<<blocks_for(ns_pad, 256), 256, 0, st>>>(src, sJ, sjld, sjoff, stat, sld, sindex,
                                                            ns, ns_pad, kernel, transposed,
                                                            (double *)d.srec.p);
  h->launches++;
  if (nt > 0 && ns > 0) {
    SfsArgs a;
    a.tpos = tpos; a.tld = tld; a.tJ = tJ; a.jld = jld; a.tindex = tindex; a.nt = nt;
    a.rec = (const double *)d.srec.p; a.ns = ns;
    a.tiles_per_split = plan.tiles_per_split;
    a.partial = (double *)d.partial.p; a.pstride = plan.pstride;
    a.transposed = transposed;
    a.shortcut = (flags & VPM_FLAG_NO_FARFIELD_SHORTCUT) ? 0 : 1;
    if (time_pairs) CK(h, cudaEventRecord(d.ev[6], st));
    launch_sfs(kernel, plan, a, st, mode);
    if (time_pairs) CK(h, cudaEventRecord(d.ev[7], st));
    h->launches++;
  } else {
    plan.nsplit = 0;
  }
  CK(h, cudaGetLastError());
  return VPM_OK;
}

float ev_ms(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) { cudaGetLastError(); return 0.f; }
  return ms;
}


// ---- strided rows of the host matrix <-> compact device blocks ------------------------
// A 2-D copy straight from/to pageable host memory is staged row by row by the driver
// (measured: 37 ms up + 60 ms down for 262 144 particles against 2 + 1 ms from registered
// memory).  If the caller has not page-locked the matrix (vpm_pin_host), the rows are
// gathered into / scattered from one pinned staging block on the host instead, and the
// transfers themselves are contiguous.
bool host_is_pinned(const void *ptr) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

int ensure_stage(vpm_handle *h, size_t doubles) {
  if (doubles <= h->h_stage_cap) return VPM_OK;
  if (h->h_stage) cudaFreeHost(h->h_stage);
  h->h_stage = nullptr;
  h->h_stage_cap = 0;
  const size_t want = doubles + doubles / 4;
  CK(h, cudaMallocHost((void **)&h->h_stage, want * sizeof(double)));
  h->h_stage_cap = want;
  return VPM_OK;
}

// The O(N) host loops over the particle matrix (strided gathers / scatters, the static-flag
// scan) are memory-latency bound on one core: at 2^24 particles they cost 0.1 s each.  Split
// them over a few threads (chunks of >= 64 Ki particles; small fields stay on the caller's thread).
template <class F>
void parallel_chunks(int64_t n, F fn) {
  const int64_t min_chunk = 1 << 16;
  unsigned hw = std::thread::hardware_concurrency();
  int nt = (int)std::min<int64_t>(std::min<unsigned>(hw ? hw : 1, 8), n / min_chunk);
  if (nt <= 1) { fn((int64_t)0, n); return; }
  std::vector<std::thread> th;
  const int64_t chunk = (n + nt - 1) / nt;
  for (int t = 1; t < nt; ++t) th.emplace_back([=] { fn(t * chunk, std::min<int64_t>(n, (t + 1) * chunk)); });
  fn((int64_t)0, std::min<int64_t>(n, chunk));
  for (auto &t : th) t.join();
}
void gather_rows(double *dst, const double *P, int64_t nf, int row0, int nrows, int64_t np) {
  parallel_chunks(np, [=](int64_t a, int64_t b) {
    for (int64_t i = a; i < b; ++i) memcpy(dst + i * nrows, P + nf * i + row0, (size_t)nrows * sizeof(double));
  });
}
void scatter_rows(double *P, int64_t nf, int row0, int nrows, int64_t np, const double *src) {
  parallel_chunks(np, [=](int64_t a, int64_t b) {
    for (int64_t i = a; i < b; ++i) memcpy(P + nf * i + row0, src + i * nrows, (size_t)nrows * sizeof(double));
  });
}
// any particle with a non-zero static flag (row 43)?
template <class R>
bool any_static(const R *P, int64_t nf, int64_t np) {
  std::atomic<bool> found{false};
  parallel_chunks(np, [&](int64_t a, int64_t b) {
    for (int64_t i = a; i < b; ++i) {
      if (P[nf * i + R_STATIC] != (R)0) { found.store(true, std::memory_order_relaxed); return; }
      if ((i & 4095) == 0 && found.load(std::memory_order_relaxed)) return;
    }
  });
  return found.load();
}

// Strided rows [row.., row+nrows) of `np` columns of a host matrix (leading dimension nf)
// -> compact device block: a 2-D DMA.  (Measured alternative: page-locking the matrix as
// mapped memory and gathering the rows with a kernel reading host memory directly gave the
// same 13.8 GB/s for the 56-byte rows at N = 2^22, so the plain copy stays.)
int h2d_rows(vpm_handle *h, cudaStream_t st, double *dst, const double *src, int64_t nf, int nrows, int64_t np) {
  if (np <= 0) return VPM_OK;
  CK(h, cudaMemcpy2DAsync(dst, nrows * sizeof(double), src, nf * sizeof(double), nrows * sizeof(double), (size_t)np,
                          cudaMemcpyHostToDevice, st));
  return VPM_OK;
}

// ---- Hook 1 pieces (single device d; targets = all particles) ---------------

// host -> device: X, Gamma, sigma rows; static flags (compacted on the host,
// only if any is set); previous U..PSE and SFS rows when they are accumulated on.
int h1_upload(vpm_handle *h, Dev &d, const double *P, int64_t nf, int64_t np, bool need_prior,
              bool need_sfs_rows, bool &has_static) {
  cudaStream_t st = d.stream;
  CK(h, cudaSetDevice(d.id));
  const size_t n = (size_t)std::max<int64_t>(np, 1);
  TRY(ensure(h, d.in7, n * 7 * sizeof(double)));
  TRY(ensure(h, d.res18, n * RES_ROWS * sizeof(double)));
  TRY(ensure(h, d.sfs3, n * 3 * sizeof(double)));
  has_static = false;
  if (np == 0) return VPM_OK;
  has_static = any_static(P, nf, np);
  const bool pinned = host_is_pinned(P);
  double *stg = nullptr;
  if (!pinned) {
    TRY(ensure_stage(h, (size_t)np * (7 + RES_ROWS + 3)));
    stg = h->h_stage;
    gather_rows(stg, P, nf, R_X, 7, np);
    CK(h, cudaMemcpyAsync(d.in7.p, stg, (size_t)np * 7 * sizeof(double), cudaMemcpyHostToDevice, st));
  } else {
    TRY(h2d_rows(h, st, (double *)d.in7.p, P, nf, 7, np));
  }
  if (has_static) {
    if (h->h_stat_cap < (size_t)np) {
      if (h->h_stat) cudaFreeHost(h->h_stat);
      h->h_stat = nullptr;
      h->h_stat_cap = 0;
      CK(h, cudaMallocHost((void **)&h->h_stat, (size_t)np * sizeof(double)));
      h->h_stat_cap = (size_t)np;
    }
    for (int64_t i = 0; i < np; ++i) h->h_stat[i] = P[nf * i + R_STATIC];
    TRY(ensure(h, d.stat, (size_t)np * sizeof(double)));
    CK(h, cudaMemcpyAsync(d.stat.p, h->h_stat, (size_t)np * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  if (need_prior || has_static) {
    if (!pinned) {
      double *s18 = stg + (size_t)np * 7;
      gather_rows(s18, P, nf, R_U, RES_ROWS, np);
      CK(h, cudaMemcpyAsync(d.res18.p, s18, (size_t)np * RES_ROWS * sizeof(double), cudaMemcpyHostToDevice, st));
    } else {
      TRY(h2d_rows(h, st, (double *)d.res18.p, P + R_U, nf, RES_ROWS, np));
    }
  }
  if (need_sfs_rows) {
    if (!pinned) {
      double *s3 = stg + (size_t)np * (7 + RES_ROWS);
      gather_rows(s3, P, nf, R_SFS, 3, np);
      CK(h, cudaMemcpyAsync(d.sfs3.p, s3, (size_t)np * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    } else {
      TRY(h2d_rows(h, st, (double *)d.sfs3.p, P + R_SFS, nf, 3, np));
    }
  }
  return VPM_OK;
}

// device-resident evaluation: U/J sweep (+ SFS sweep) over all particles.
// `prior` says res18/sfs3 hold previous values that must be accumulated on.
int h1_eval(vpm_handle *h, Dev &d, int64_t np, int kernel, int flags, bool has_static, bool prior) {
  cudaStream_t st = d.stream;
  CK(h, cudaSetDevice(d.id));
  const double *stat = has_static ? (const double *)d.stat.p : nullptr;
  SrcView src{(const double *)d.in7.p, 7, 0, 3, 6};
  Plan plan;
  CK(h, cudaEventRecord(d.ev[1], st));
  TRY(uj_sweep(h, d, st, kernel, (const double *)d.in7.p, 7, np, src, 0, np, flags, plan));
  CK(h, cudaEventRecord(d.ev[2], st));
  if (np > 0) {
    UjFinishArgs f;
    f.partial = (const double *)d.partial.p; f.pstride = plan.pstride; f.nsplit = plan.nsplit;
    f.nt = np; f.out = (double *)d.res18.p; f.ld = RES_ROWS; f.urow = RES_U; f.jrow = RES_J;
    f.zrow0 = RES_W; f.zrow1 = RES_PSE; f.want_U = 1; f.want_J = 1;
    f.accumulate = prior ? 1 : 0;
    f.reset = (flags & VPM_FLAG_RESET) ? 1 : 0;
    f.stat = stat; f.sld = 1;
    if (!prior) {
      // nothing uploaded: vorticity / PSE rows of the block must still be defined
      CK(h, cudaMemsetAsync(d.res18.p, 0, (size_t)np * RES_ROWS * sizeof(double), st));
    }
    uj_finish_kernel<<<blocks_for(np, 256), 256, 0, st>>>(f);
    h->launches++;
    CK(h, cudaGetLastError());
  }
  CK(h, cudaEventRecord(d.ev[3], st));
  h->timing.uj_pairs = np * np;
  h->timing.sfs_pairs = 0;
  if (np > 0 && (flags & VPM_FLAG_SFS)) {
    Plan sp;
    const double *J = (const double *)d.res18.p + RES_J;
    TRY(sfs_sweep(h, d, st, kernel, (const double *)d.in7.p, 7, J, RES_ROWS, nullptr, np, src, J,
                  RES_ROWS, 0, stat, 1, nullptr, np, flags, sp));
    SfsFinishArgs f;
    f.partial = (const double *)d.partial.p; f.pstride = sp.pstride; f.nsplit = sp.nsplit;
    f.nt = np; f.tindex = nullptr; f.out = (double *)d.sfs3.p; f.ld = 3; f.row = 0;
    f.accumulate = 1;  // sfs3 holds either the uploaded rows or (below) zeros
    f.reset = (flags & VPM_FLAG_RESET_SFS) ? 1 : 0;
    f.filter_static = 1; f.stat = stat; f.sld = 1;
    sfs_finish_kernel<<<blocks_for(np, 256), 256, 0, st>>>(f);
    h->launches++;
    CK(h, cudaGetLastError());
    h->timing.sfs_pairs = np * np;
  } else if (np > 0 && (flags & VPM_FLAG_RESET_SFS)) {
    zero_rows_kernel<<<blocks_for(np, 256), 256, 0, st>>>((double *)d.sfs3.p, 3, 0, 3, np, stat, 1);
    h->launches++;
    CK(h, cudaGetLastError());
  }
  CK(h, cudaEventRecord(d.ev[4], st));
  return VPM_OK;
}

int h1_download(vpm_handle *h, Dev &d, double *P, int64_t nf, int64_t np, int flags) {
  cudaStream_t st = d.stream;
  CK(h, cudaSetDevice(d.id));
  const bool sfs_rows = flags & (VPM_FLAG_SFS | VPM_FLAG_RESET_SFS);
  if (np > 0 && !host_is_pinned(P)) {
    // pageable matrix: contiguous D2H into the pinned staging block, then scatter on the host
    TRY(ensure_stage(h, (size_t)np * (7 + RES_ROWS + 3)));
    double *s18 = h->h_stage + (size_t)np * 7, *s3 = h->h_stage + (size_t)np * (7 + RES_ROWS);
    CK(h, cudaMemcpyAsync(s18, d.res18.p, (size_t)np * RES_ROWS * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (sfs_rows) CK(h, cudaMemcpyAsync(s3, d.sfs3.p, (size_t)np * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(h, cudaEventRecord(d.ev[5], st));
    CK(h, cudaStreamSynchronize(st));
    scatter_rows(P, nf, R_U, RES_ROWS, np, s18);
    if (sfs_rows) scatter_rows(P, nf, R_SFS, 3, np, s3);
    return VPM_OK;
  }
  if (np > 0) {
    CK(h, cudaMemcpy2DAsync(P + R_U, nf * sizeof(double), d.res18.p, RES_ROWS * sizeof(double),
                            RES_ROWS * sizeof(double), (size_t)np, cudaMemcpyDeviceToHost, st));
    if (sfs_rows)
      CK(h, cudaMemcpy2DAsync(P + R_SFS, nf * sizeof(double), d.sfs3.p, 3 * sizeof(double),
                              3 * sizeof(double), (size_t)np, cudaMemcpyDeviceToHost, st));
  }
  CK(h, cudaEventRecord(d.ev[5], st));
  CK(h, cudaStreamSynchronize(st));
  return VPM_OK;
}

void h1_fill_timing(vpm_handle *h, Dev &d) {
  vpm_timing &t = h->timing;
  h->device_timing = 0;
  t.h2d_ms = ev_ms(d.ev[0], d.ev[1]);
  t.uj_ms = ev_ms(d.ev[1], d.ev[2]);
  t.finish_ms = ev_ms(d.ev[2], d.ev[3]);
  t.sfs_ms = ev_ms(d.ev[3], d.ev[4]);
  t.d2h_ms = ev_ms(d.ev[4], d.ev[5]);
  t.total_ms = ev_ms(d.ev[0], d.ev[5]);
  t.prep_ms = 0.0;
  t.kernel_launches = h->launches;
  t.n_gpus = (int32_t)h->devs.size();
}


// ---- NCCL, loaded lazily: only the single-process multi-GPU path needs it ----
typedef int (*nccl_comm_init_all_t)(void **comms, int ndev, const int *devlist);
typedef int (*nccl_all_gather_t)(const void *send, void *recv, size_t count, int dtype, void *comm,
                                 cudaStream_t stream);
typedef int (*nccl_broadcast_t)(const void *send, void *recv, size_t count, int dtype, int root, void *comm,
                                cudaStream_t stream);
typedef int (*nccl_sendrecv_t)(void *buf, size_t count, int dtype, int peer, void *comm, cudaStream_t stream);
typedef int (*nccl_group_t)(void);
typedef int (*nccl_comm_destroy_t)(void *comm);
typedef const char *(*nccl_err_t)(int);
struct NcclApi {
  nccl_comm_init_all_t comm_init_all = nullptr;
  nccl_all_gather_t all_gather = nullptr;
  nccl_broadcast_t broadcast = nullptr;
  nccl_sendrecv_t send = nullptr, recv = nullptr;
  nccl_group_t group_start = nullptr, group_end = nullptr;
  nccl_comm_destroy_t comm_destroy = nullptr;
  nccl_err_t err_string = nullptr;
};
NcclApi g_nccl;
constexpr int kNcclFloat64 = 8;  // ncclDouble (nccl.h ncclDataType_t)
constexpr int kNcclInt8 = 0;     // ncclChar

int nccl_load(vpm_handle *h) {
  if (h->nccl_lib) return VPM_OK;
  void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return fail(h, VPM_ENCCL, "cannot load libnccl.so.2: %s", dlerror());
  g_nccl.comm_init_all = (nccl_comm_init_all_t)dlsym(lib, "ncclCommInitAll");
  g_nccl.all_gather = (nccl_all_gather_t)dlsym(lib, "ncclAllGather");
  g_nccl.broadcast = (nccl_broadcast_t)dlsym(lib, "ncclBroadcast");
  g_nccl.send = (nccl_sendrecv_t)dlsym(lib, "ncclSend");
  g_nccl.recv = (nccl_sendrecv_t)dlsym(lib, "ncclRecv");
  g_nccl.group_start = (nccl_group_t)dlsym(lib, "ncclGroupStart");
  g_nccl.group_end = (nccl_group_t)dlsym(lib, "ncclGroupEnd");
  g_nccl.comm_destroy = (nccl_comm_destroy_t)dlsym(lib, "ncclCommDestroy");
  g_nccl.err_string = (nccl_err_t)dlsym(lib, "ncclGetErrorString");
  if (!g_nccl.comm_init_all || !g_nccl.all_gather || !g_nccl.broadcast || !g_nccl.group_start || !g_nccl.group_end ||
      !g_nccl.comm_destroy || !g_nccl.send || !g_nccl.recv)
    return fail(h, VPM_ENCCL, "libnccl.so.2 lacks a required symbol");
  h->nccl_lib = lib;
  return VPM_OK;
}

#define NCK(h, call)                                                                   \
  do {                                                                                 \
    int r_ = (call);                                                                   \
    if (r_ != 0)                                                                       \
      return fail(h, VPM_ENCCL, "%s failed: %s", #call,                                \
                  g_nccl.err_string ? g_nccl.err_string(r_) : "nccl error");           \
  } while (0)

int ensure_comms(vpm_handle *h) {
  if (!h->comms.empty()) return VPM_OK;
  TRY(nccl_load(h));
  const int G = (int)h->devs.size();
  std::vector<int> ids(G);
  for (int g = 0; g < G; ++g) ids[g] = h->devs[g].id;
  std::vector<void *> comms(G, nullptr);
  NCK(h, g_nccl.comm_init_all(comms.data(), G, ids.data()));
  h->comms = comms;  // only a fully initialised set is kept
  return VPM_OK;
}

// Replicate `bytes` of one buffer from device 0 to every device of the handle over NVLink
// (ncclBroadcast on each device's stream): the host uploads a replicated input once
// instead of G times over PCIe.
int bcast_from_dev0(vpm_handle *h, Buf Dev::*member, size_t bytes) {
  const int G = (int)h->devs.size();
  if (G < 2 || bytes == 0) return VPM_OK;
  TRY(ensure_comms(h));
  NCK(h, g_nccl.group_start());
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    NCK(h, g_nccl.broadcast((h->devs[0].*member).p, (d.*member).p, bytes, kNcclInt8, 0, h->comms[g], d.stream));
  }
  NCK(h, g_nccl.group_end());
  return VPM_OK;
}

// UJ_direct on G devices of this process: targets block-sharded, sources
// replicated by the host upload; with SFS the final J of every shard is
// all-gathered (NCCL over NVLink) before the second sweep.
int uj_direct_multi(vpm_handle *h, double *P, int64_t nf, int64_t np, int kernel, int flags) {
  const int G = (int)h->devs.size();
  h->launches = 0;
  if (np == 0) return VPM_OK;
  TRY(ensure_comms(h));
  const bool reset = flags & VPM_FLAG_RESET;
  const bool do_sfs = flags & VPM_FLAG_SFS;
  const bool sfs_rows = do_sfs || (flags & VPM_FLAG_RESET_SFS);
  const int64_t shard = (np + G - 1) / G;
  const int64_t np_pad = shard * G;
  bool has_static = false;
  has_static = any_static(P, nf, np);
  if (has_static) {
    if (h->h_stat_cap < (size_t)np) {
      if (h->h_stat) cudaFreeHost(h->h_stat);
      h->h_stat = nullptr; h->h_stat_cap = 0;
      CK(h, cudaMallocHost((void **)&h->h_stat, (size_t)np * sizeof(double)));
      h->h_stat_cap = (size_t)np;
    }
    for (int64_t i = 0; i < np; ++i) h->h_stat[i] = P[nf * i + R_STATIC];
  }
  const bool prior = !reset || has_static;
  const bool pinned = host_is_pinned(P);
  if (!pinned) TRY(ensure_stage(h, (size_t)np * (7 + RES_ROWS + 3)));
  double *stg18 = pinned ? nullptr : h->h_stage + (size_t)np * 7;
  double *stg3 = pinned ? nullptr : h->h_stage + (size_t)np * (7 + RES_ROWS);
  std::vector<Plan> plans(G);
  // sources (X, Gamma, sigma, static flags) go to device 0 once and are broadcast over NVLink
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    CK(h, cudaSetDevice(d.id));
    TRY(ensure(h, d.in7, (size_t)np * 7 * sizeof(double)));
    TRY(ensure(h, d.res18, (size_t)np_pad * RES_ROWS * sizeof(double)));
    TRY(ensure(h, d.sfs3, (size_t)np_pad * 3 * sizeof(double)));
    if (has_static) TRY(ensure(h, d.stat, (size_t)np * sizeof(double)));
  }
  {
    Dev &d0 = h->devs[0];
    CK(h, cudaSetDevice(d0.id));
    CK(h, cudaEventRecord(d0.ev[0], d0.stream));
    if (pinned) {
      TRY(h2d_rows(h, d0.stream, (double *)d0.in7.p, P, nf, 7, np));
    } else {  // pageable matrix: gather the strided rows into the pinned staging block (see h1_upload)
      TRY(ensure_stage(h, (size_t)np * (7 + RES_ROWS + 3)));
      gather_rows(h->h_stage, P, nf, R_X, 7, np);
      CK(h, cudaMemcpyAsync(d0.in7.p, h->h_stage, (size_t)np * 7 * sizeof(double), cudaMemcpyHostToDevice, d0.stream));
      if (prior) gather_rows(stg18, P, nf, R_U, RES_ROWS, np);
      if (sfs_rows) gather_rows(stg3, P, nf, R_SFS, 3, np);
    }
    if (has_static)
      CK(h, cudaMemcpyAsync(d0.stat.p, h->h_stat, (size_t)np * sizeof(double), cudaMemcpyHostToDevice, d0.stream));
  }
  TRY(bcast_from_dev0(h, &Dev::in7, (size_t)np * 7 * sizeof(double)));
  if (has_static) TRY(bcast_from_dev0(h, &Dev::stat, (size_t)np * sizeof(double)));
  // per-shard previous values + U/J sweep on every device
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    cudaStream_t st = d.stream;
    const int64_t t0 = std::min(np, g * shard), t1 = std::min(np, t0 + shard), nt = t1 - t0;
    CK(h, cudaSetDevice(d.id));
    double *res = (double *)d.res18.p + t0 * RES_ROWS;
    double *sfs = (double *)d.sfs3.p + t0 * 3;
    if (nt > 0) {
      if (prior && pinned)
        TRY(h2d_rows(h, st, (double *)res, P + nf * t0 + R_U, nf, RES_ROWS, nt));
      else if (prior)
        CK(h, cudaMemcpyAsync(res, stg18 + t0 * RES_ROWS, (size_t)nt * RES_ROWS * sizeof(double), cudaMemcpyHostToDevice, st));
      else
        CK(h, cudaMemsetAsync(res, 0, (size_t)nt * RES_ROWS * sizeof(double), st));
      if (sfs_rows && pinned)
        TRY(h2d_rows(h, st, (double *)sfs, P + nf * t0 + R_SFS, nf, 3, nt));
      else if (sfs_rows)
        CK(h, cudaMemcpyAsync(sfs, stg3 + t0 * 3, (size_t)nt * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    if (g == 0) CK(h, cudaEventRecord(d.ev[1], st));
    SrcView src{(const double *)d.in7.p, 7, 0, 3, 6};
    TRY(uj_sweep(h, d, st, kernel, (const double *)d.in7.p + t0 * 7, 7, nt, src, 0, np, flags, plans[g]));
    if (g == 0) CK(h, cudaEventRecord(d.ev[2], st));
    if (nt > 0) {
      UjFinishArgs f;
      f.partial = (const double *)d.partial.p; f.pstride = plans[g].pstride; f.nsplit = plans[g].nsplit;
      f.nt = nt; f.out = res; f.ld = RES_ROWS; f.urow = RES_U; f.jrow = RES_J;
      f.zrow0 = RES_W; f.zrow1 = RES_PSE; f.want_U = 1; f.want_J = 1;
      f.accumulate = prior ? 1 : 0; f.reset = reset ? 1 : 0;
      f.stat = has_static ? (const double *)d.stat.p + t0 : nullptr; f.sld = 1;
      uj_finish_kernel<<<blocks_for(nt, 256), 256, 0, st>>>(f);
      h->launches++;
      CK(h, cudaGetLastError());
    }
    if (g == 0) CK(h, cudaEventRecord(d.ev[3], st));
  }
  h->timing.uj_pairs = np * np;
  h->timing.sfs_pairs = 0;
  if (do_sfs) {
    // every device needs the final J of every particle: all-gather the result shards
    NCK(h, g_nccl.group_start());
    for (int g = 0; g < G; ++g) {
      Dev &d = h->devs[g];
      double *base = (double *)d.res18.p;
      NCK(h, g_nccl.all_gather(base + (int64_t)g * shard * RES_ROWS, base, (size_t)shard * RES_ROWS,
                               kNcclFloat64, h->comms[g], d.stream));
    }
    NCK(h, g_nccl.group_end());
    for (int g = 0; g < G; ++g) {
      Dev &d = h->devs[g];
      cudaStream_t st = d.stream;
      const int64_t t0 = std::min(np, g * shard), t1 = std::min(np, t0 + shard), nt = t1 - t0;
      CK(h, cudaSetDevice(d.id));
      const double *stat = has_static ? (const double *)d.stat.p : nullptr;
      SrcView src{(const double *)d.in7.p, 7, 0, 3, 6};
      const double *J = (const double *)d.res18.p + RES_J;
      Plan sp;
      TRY(sfs_sweep(h, d, st, kernel, (const double *)d.in7.p + t0 * 7, 7, J + t0 * RES_ROWS, RES_ROWS,
                    nullptr, nt, src, J, RES_ROWS, 0, stat, 1, nullptr, np, flags, sp));
      if (nt > 0) {
        SfsFinishArgs f;
        f.partial = (const double *)d.partial.p; f.pstride = sp.pstride; f.nsplit = sp.nsplit;
        f.nt = nt; f.tindex = nullptr; f.out = (double *)d.sfs3.p + t0 * 3; f.ld = 3; f.row = 0;
        f.accumulate = 1; f.reset = (flags & VPM_FLAG_RESET_SFS) ? 1 : 0;
        f.filter_static = 1; f.stat = stat ? stat + t0 : nullptr; f.sld = 1;
        sfs_finish_kernel<<<blocks_for(nt, 256), 256, 0, st>>>(f);
        h->launches++;
        CK(h, cudaGetLastError());
      }
    }
    h->timing.sfs_pairs = np * np;
  } else if (flags & VPM_FLAG_RESET_SFS) {
    for (int g = 0; g < G; ++g) {
      Dev &d = h->devs[g];
      const int64_t t0 = std::min(np, g * shard), t1 = std::min(np, t0 + shard), nt = t1 - t0;
      if (nt == 0) continue;
      CK(h, cudaSetDevice(d.id));
      zero_rows_kernel<<<blocks_for(nt, 256), 256, 0, d.stream>>>(
          (double *)d.sfs3.p + t0 * 3, 3, 0, 3, nt, has_static ? (const double *)d.stat.p + t0 : nullptr, 1);
      h->launches++;
      CK(h, cudaGetLastError());
    }
  }
  CK(h, cudaSetDevice(h->devs[0].id));
  CK(h, cudaEventRecord(h->devs[0].ev[4], h->devs[0].stream));
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    const int64_t t0 = std::min(np, g * shard), t1 = std::min(np, t0 + shard), nt = t1 - t0;
    if (nt == 0) continue;
    CK(h, cudaSetDevice(d.id));
    if (pinned) {
      CK(h, cudaMemcpy2DAsync(P + nf * t0 + R_U, nf * sizeof(double), (double *)d.res18.p + t0 * RES_ROWS,
                              RES_ROWS * sizeof(double), RES_ROWS * sizeof(double), (size_t)nt,
                              cudaMemcpyDeviceToHost, d.stream));
      if (sfs_rows)
        CK(h, cudaMemcpy2DAsync(P + nf * t0 + R_SFS, nf * sizeof(double), (double *)d.sfs3.p + t0 * 3,
                                3 * sizeof(double), 3 * sizeof(double), (size_t)nt, cudaMemcpyDeviceToHost,
                                d.stream));
    } else {  // contiguous D2H of each shard into the pinned staging block, scattered below
      CK(h, cudaMemcpyAsync(stg18 + t0 * RES_ROWS, (double *)d.res18.p + t0 * RES_ROWS,
                            (size_t)nt * RES_ROWS * sizeof(double), cudaMemcpyDeviceToHost, d.stream));
      if (sfs_rows)
        CK(h, cudaMemcpyAsync(stg3 + t0 * 3, (double *)d.sfs3.p + t0 * 3, (size_t)nt * 3 * sizeof(double),
                              cudaMemcpyDeviceToHost, d.stream));
    }
  }
  for (int g = G - 1; g >= 0; --g) {
    Dev &d = h->devs[g];
    CK(h, cudaSetDevice(d.id));
    if (g == 0) CK(h, cudaEventRecord(d.ev[5], d.stream));
    CK(h, cudaStreamSynchronize(d.stream));
  }
  if (!pinned) {
    scatter_rows(P, nf, R_U, RES_ROWS, np, stg18);
    if (sfs_rows) scatter_rows(P, nf, R_SFS, 3, np, stg3);
  }
  h1_fill_timing(h, h->devs[0]);
  h->np_resident = -1;
  return VPM_OK;
}

// ---- leaf-pair list (Hook 3): CSR by target leaf, built on device 0 (vpm_csr.cuh) ----
// carve aligned sub-arrays out of one device allocation
struct Carver {
  char *base;
  size_t off = 0;
  explicit Carver(void *p) : base((char *)p) {}
  template <class T>
  T *take(size_t n) {
    off = (off + 15) / 16 * 16;
    T *r = (T *)(base + off);
    off += n * sizeof(T);
    return r;
  }
};

struct DevCsr {
  LeafCsr csr;              // pointers into device 0's ibuf
  size_t bcast_bytes = 0;   // leading bytes of ibuf every device needs (tables + sort indices)
  int nt = kThreads;        // CTA width (targets per work item): 32, 64 or 128
  int64_t nwi = 0;          // work items
  int64_t pairs = 0;        // pair visits of the whole list
  std::vector<int64_t> cut;                     // [G + 1] work-item cuts
  std::vector<int64_t> first_leaf, first_off;   // [G + 1] item cut[g]     (first item of device g)
  std::vector<int64_t> last_leaf, last_off;     // [G + 1] item cut[g] - 1 (last item of device g - 1)
  const int64_t *d_tsort = nullptr, *d_ssort = nullptr;
};

// rebase the table pointers of device 0 onto another device's copy of ibuf
LeafCsr rebase_csr(const LeafCsr &c, const void *from, const void *to) {
  const ptrdiff_t shift = (const char *)to - (const char *)from;
  auto rb = [shift](auto *p) { return (decltype(p))((const char *)p + shift); };
  LeafCsr r;
  r.wi_leaf = rb(c.wi_leaf); r.wi_off = rb(c.wi_off);
  r.tleaf_begin = rb(c.tleaf_begin); r.tleaf_end = rb(c.tleaf_end);
  r.csr_ptr = rb(c.csr_ptr); r.csr_src = rb(c.csr_src);
  r.sleaf_begin = rb(c.sleaf_begin); r.sleaf_end = rb(c.sleaf_end);
  return r;
}

int build_csr_device(vpm_handle *h, const char *fn, const int64_t *tb, const int64_t *te, int64_t ntl,
                     int64_t n_tgt, const int64_t *sb, const int64_t *se, int64_t nsl, int64_t n_src,
                     const int32_t *pt, const int32_t *ps, int64_t npairs, int G, const int64_t *tsort,
                     int64_t n_tsort, const int64_t *ssort, int64_t n_ssort, DevCsr &out, bool dev_in = false) {
  // O(leaves) checks stay on the host; everything O(list entries) runs on the device.
  // dev_in: the tables are device arrays produced by vpm_leaflists_build (already valid).
  int64_t max_wi = dev_in ? n_tgt / 32 + ntl : 0;
  for (int64_t l = 0; l < ntl && !dev_in; ++l) {
    if (tb[l] < 0 || te[l] < tb[l] || te[l] > n_tgt)
      return fail(h, VPM_EINVAL, "%s: target leaf %lld range [%lld,%lld) outside 0..%lld", fn, (long long)l, (long long)tb[l], (long long)te[l], (long long)n_tgt);
    max_wi += (te[l] - tb[l] + 31) / 32;
  }
  for (int64_t l = 0; l < nsl && !dev_in; ++l)
    if (sb[l] < 0 || se[l] < sb[l] || se[l] > n_src)
      return fail(h, VPM_EINVAL, "%s: source leaf %lld range [%lld,%lld) outside 0..%lld", fn, (long long)l, (long long)sb[l], (long long)se[l], (long long)n_src);
  max_wi = std::max<int64_t>(max_wi, 1);
  Dev &d = h->devs[0];
  cudaStream_t st = d.stream;
  CK(h, cudaSetDevice(d.id));
  // replicated tables (ibuf) ...
  const size_t ibytes = 16 * 12 + (size_t)max_wi * 8 + (size_t)ntl * 16 + ((size_t)ntl + 1) * 8 + (size_t)npairs * 4 +
                        (size_t)nsl * 16 + (size_t)(n_tsort + n_ssort) * 8;
  TRY(ensure(h, d.ibuf, ibytes));
  Carver cv(d.ibuf.p);
  int32_t *wl = cv.take<int32_t>((size_t)max_wi), *wo = cv.take<int32_t>((size_t)max_wi);
  int64_t *dtb = cv.take<int64_t>((size_t)ntl), *dte = cv.take<int64_t>((size_t)ntl);
  u64 *dptr = cv.take<u64>((size_t)ntl + 1);
  int32_t *dsrc = cv.take<int32_t>((size_t)npairs);
  int64_t *dsb = cv.take<int64_t>((size_t)nsl), *dse = cv.take<int64_t>((size_t)nsl);
  int64_t *dts = cv.take<int64_t>((size_t)n_tsort), *dss = cv.take<int64_t>((size_t)n_ssort);
  out.bcast_bytes = cv.off;
  // ... and device-0 scratch
  const size_t sbytes = 16 * 8 + (size_t)npairs * 4 + (size_t)ntl * 8 * 3 + (size_t)max_wi * 8 + CS_SLOTS * 8 + (size_t)(G + 1) * 40;
  TRY(ensure(h, d.scr, sbytes));
  Carver sc(d.scr.p);
  int32_t *dpt = sc.take<int32_t>((size_t)npairs);
  u64 *srcw = sc.take<u64>((size_t)ntl), *wcnt = sc.take<u64>((size_t)ntl), *wofs = sc.take<u64>((size_t)ntl);
  u64 *wiw = sc.take<u64>((size_t)max_wi);
  u64 *stats = sc.take<u64>(CS_SLOTS);
  int64_t *dcut = sc.take<int64_t>((size_t)(G + 1) * 5);
  // cub temporary storage: the largest of the four calls below
  const int key_bits = std::max(1, (int)std::ceil(std::log2((double)std::max<int64_t>(ntl, 2))));
  size_t tmp = 0, t1 = 0;
  cub::DeviceScan::InclusiveSum(nullptr, t1, dptr, dptr, (int64_t)ntl + 1, st); tmp = std::max(tmp, t1);
  cub::DeviceScan::ExclusiveSum(nullptr, t1, wcnt, wofs, (int64_t)ntl, st); tmp = std::max(tmp, t1);
  cub::DeviceScan::InclusiveSum(nullptr, t1, wiw, wiw, max_wi, st); tmp = std::max(tmp, t1);
  cub::DeviceRadixSort::SortPairs(nullptr, t1, (const int32_t *)nullptr, (int32_t *)nullptr, (const int32_t *)nullptr,
                                  (int32_t *)nullptr, npairs, 0, key_bits, st);
  tmp = std::max(tmp, t1);
  TRY(ensure(h, d.cubtmp, tmp + 16));

  auto put = [&](auto *dst, const auto *srcp, size_t n) -> cudaError_t {
    if (n == 0) return cudaSuccess;
    return cudaMemcpyAsync((void *)dst, (const void *)srcp, n * sizeof(*srcp), cudaMemcpyDefault, st);
  };
  CK(h, put(dtb, tb, (size_t)ntl));
  CK(h, put(dte, te, (size_t)ntl));
  CK(h, put(dsb, sb, (size_t)nsl));
  CK(h, put(dse, se, (size_t)nsl));
  CK(h, put(dpt, pt, (size_t)npairs));
  CK(h, put(dsrc, ps, (size_t)npairs));  // already the CSR column array when the list is grouped
  if (n_tsort) CK(h, put(dts, tsort, (size_t)n_tsort));
  if (n_ssort) CK(h, put(dss, ssort, (size_t)n_ssort));
  csr_init_stats_kernel<<<1, 32, 0, st>>>(stats);
  csr_zero_kernel<<<blocks_for(ntl + 1, 256), 256, 0, st>>>(dptr, ntl + 1);
  csr_zero_kernel<<<blocks_for(ntl, 256), 256, 0, st>>>(srcw, ntl);
  csr_count_kernel<<<blocks_for(npairs, 256), 256, 0, st>>>(dpt, dsrc, npairs, ntl, nsl, dsb, dse, dptr, srcw, stats);
  csr_cand_kernel<<<blocks_for(ntl, 256), 256, 0, st>>>(dtb, dte, ntl, srcw, stats);
  t1 = d.cubtmp.cap;
  CK(h, cub::DeviceScan::InclusiveSum(d.cubtmp.p, t1, dptr, dptr, (int64_t)ntl + 1, st));
  h->launches += 6;
  u64 hs[CS_SLOTS];
  CK(h, cudaMemcpyAsync(hs, stats, sizeof hs, cudaMemcpyDeviceToHost, st));
  CK(h, cudaStreamSynchronize(st));
  CK(h, cudaGetLastError());
  if (hs[CS_BAD] != ~0ull) {
    const int64_t k = (int64_t)hs[CS_BAD] - 1;
    return fail(h, VPM_EINVAL, "%s: pair %lld = (%d,%d) outside the leaf tables", fn, (long long)k,
                dev_in ? -1 : pt[k], dev_in ? -1 : ps[k]);
  }
  // CTA width: minimise the padded lane-work  sum_leaf ceil(size/NT)*NT * (its source bodies);
  // wider CTAs amortise the tile traffic better: require a 10 % gain to go narrower
  double best = -1.0;
  const int cands[3] = {128, 64, 32};
  const u64 wsum[3] = {hs[CS_W128], hs[CS_W64], hs[CS_W32]};
  for (int c = 0; c < 3; ++c)
    if (best < 0.0 || (double)wsum[c] < 0.9 * best) { best = (double)wsum[c]; out.nt = cands[c]; }
  out.pairs = (int64_t)hs[CS_PAIRS];
  if (hs[CS_UNSORTED]) {
    // stable radix sort by target leaf keeps the list order inside each group
    TRY(ensure(h, d.scr2, (size_t)npairs * 8 + 32));
    Carver s2(d.scr2.p);
    int32_t *keys_out = s2.take<int32_t>((size_t)npairs), *vals_out = s2.take<int32_t>((size_t)npairs);
    t1 = d.cubtmp.cap;
    CK(h, cub::DeviceRadixSort::SortPairs(d.cubtmp.p, t1, (const int32_t *)dpt, keys_out, (const int32_t *)dsrc, vals_out,
                                          npairs, 0, key_bits, st));
    CK(h, cudaMemcpyAsync(dsrc, vals_out, (size_t)npairs * 4, cudaMemcpyDeviceToDevice, st));
    h->launches += 1;
  }
  csr_wi_count_kernel<<<blocks_for(ntl, 256), 256, 0, st>>>(dtb, dte, ntl, dptr, out.nt, wcnt);
  t1 = d.cubtmp.cap;
  CK(h, cub::DeviceScan::ExclusiveSum(d.cubtmp.p, t1, wcnt, wofs, (int64_t)ntl, st));
  csr_zero_kernel<<<blocks_for(max_wi, 256), 256, 0, st>>>(wiw, max_wi);
  csr_wi_fill_kernel<<<blocks_for(ntl, 256), 256, 0, st>>>(dtb, dte, ntl, wofs, wcnt, srcw, out.nt, wl, wo, wiw, stats);
  t1 = d.cubtmp.cap;
  CK(h, cub::DeviceScan::InclusiveSum(d.cubtmp.p, t1, wiw, wiw, max_wi, st));
  h->launches += 5;
  CK(h, cudaMemcpyAsync(hs, stats, sizeof hs, cudaMemcpyDeviceToHost, st));
  CK(h, cudaStreamSynchronize(st));
  out.nwi = (int64_t)hs[CS_NWI];
  out.cut.assign((size_t)G + 1, out.nwi);
  out.cut[0] = 0;
  for (auto *v : {&out.first_leaf, &out.first_off, &out.last_leaf, &out.last_off}) v->assign((size_t)G + 1, 0);
  if (out.nwi > 0) {
    if (G + 1 > 32) return fail(h, VPM_EINVAL, "%s: more than 31 devices", fn);
    csr_cut_kernel<<<1, 32, 0, st>>>(wiw, wl, wo, out.nwi, G, dcut);
    h->launches += 1;
    std::vector<int64_t> hc((size_t)(G + 1) * 5);
    CK(h, cudaMemcpyAsync(hc.data(), dcut, hc.size() * 8, cudaMemcpyDeviceToHost, st));
    CK(h, cudaStreamSynchronize(st));
    for (int g = 0; g <= G; ++g) {
      const int64_t *c = &hc[(size_t)5 * g];
      out.cut[(size_t)g] = c[0];
      out.first_leaf[(size_t)g] = c[1]; out.first_off[(size_t)g] = c[2];
      out.last_leaf[(size_t)g] = c[3]; out.last_off[(size_t)g] = c[4];
    }
  }
  CK(h, cudaGetLastError());
  out.csr.wi_leaf = wl; out.csr.wi_off = wo; out.csr.tleaf_begin = dtb; out.csr.tleaf_end = dte;
  out.csr.csr_ptr = (const int64_t *)dptr; out.csr.csr_src = dsrc; out.csr.sleaf_begin = dsb; out.csr.sleaf_end = dse;
  out.d_tsort = dts; out.d_ssort = dss;
  return VPM_OK;
}


// ---- device-built leaf lists (vpm_tree.cuh) --------------------------------------
// Builds sort index, leaf ranges and the near-field list from the rows X (3) and sigma of a
// device-resident column-major matrix view.  Results stay on device 0 (d.tree, d.tlist).
struct TreeView {
  int64_t *sidx, *lbegin, *lend;  // [np], [nl], [nl]
  int32_t *pt, *ps;               // [npairs]
};
TreeView tree_view(vpm_handle *h) {
  Dev &d = h->devs[0];
  TreeView v;
  Carver cv(d.tree.p);
  const size_t n = (size_t)std::max<int64_t>(h->tree_np, 1);
  v.sidx = cv.take<int64_t>(n); v.lbegin = cv.take<int64_t>(n); v.lend = cv.take<int64_t>(n);
  Carver cl(d.tlist.p);
  const size_t m = (size_t)std::max<int64_t>(h->tree_npairs, 1);
  v.pt = cl.take<int32_t>(m); v.ps = cl.take<int32_t>(m);
  return v;
}

int tree_build(vpm_handle *h, const double *d_P, int64_t ld, int osig, int64_t np, int64_t ncrit, double theta) {
  const char *fn = "vpm_leaflists_build";
  Dev &d = h->devs[0];
  cudaStream_t st = d.stream;
  CK(h, cudaSetDevice(d.id));
  h->tree_np = -1;
  const size_t n = (size_t)np;
  TRY(ensure(h, d.tree, 3 * n * 8 + 64));
  // scratch: bb[8] | keys | idx | skeys | rank (u64) | lkey | ctr[3] | rad | cnt | ofs
  TRY(ensure(h, d.scr, 16 * 16 + 8 * 8 + n * 8 * 11));
  Carver sc(d.scr.p);
  long long *bb = sc.take<long long>(8);
  int64_t *keys = sc.take<int64_t>(n), *idx0 = sc.take<int64_t>(n), *skeys = sc.take<int64_t>(n);
  u64 *rank = sc.take<u64>(n);
  int64_t *lkey = sc.take<int64_t>(n);
  double *ctr = sc.take<double>(3 * n), *rad = sc.take<double>(n);
  u64 *cnt = sc.take<u64>(n), *ofs = sc.take<u64>(n);
  Carver tv(d.tree.p);
  int64_t *sidx = tv.take<int64_t>(n), *lbegin = tv.take<int64_t>(n), *lend = tv.take<int64_t>(n);

  tree_bbox_init_kernel<<<1, 32, 0, st>>>(bb);
  tree_bbox_kernel<<<blocks_for(np, 256), 256, 0, st>>>(d_P, ld, np, bb);
  h->launches += 2;
  long long hb[8];
  CK(h, cudaMemcpyAsync(hb, bb, sizeof hb, cudaMemcpyDeviceToHost, st));
  CK(h, cudaStreamSynchronize(st));
  CK(h, cudaGetLastError());
  // grid: cell size for a mean occupancy of ncrit/2; thin directions are padded to 1e-3 of
  // the largest extent so that planar / linear fields do not explode the cell count
  TreeGrid g;
  double ext[3], emax = 0.0;
  for (int a = 0; a < 3; ++a) {
    g.lo[a] = ordered_to_dbl(hb[a]);
    ext[a] = ordered_to_dbl(hb[3 + a]) - g.lo[a];
    if (!std::isfinite(ext[a])) return fail(h, VPM_EINVAL, "%s: non-finite particle positions", fn);
    emax = std::max(emax, ext[a]);
  }
  for (int a = 0; a < 3; ++a) ext[a] = std::max(std::max(ext[a], 1e-3 * emax), 1e-300);
  const double vol = ext[0] * ext[1] * ext[2];
  g.h = std::pow(vol * ((double)ncrit / 2.0) / (double)np, 1.0 / 3.0);
  if (!(g.h > 0.0) || !std::isfinite(g.h)) g.h = 1.0;
  g.theta = theta;
  auto dims_of = [&](double hh, int64_t dims[3]) {
    double ncell_d = 1.0;
    for (int a = 0; a < 3; ++a) {
      const double c = std::max(1.0, std::ceil(ext[a] / hh));
      dims[a] = (int64_t)std::min(c, 4.0e18);
      ncell_d *= c;
    }
    return ncell_d;
  };
  if (dims_of(g.h, g.dims) > 1.0e9)
    return fail(h, VPM_EINVAL, "%s: too many grid cells (field too anisotropic for ncrit = %lld)", fn, (long long)ncrit);
  size_t tmp = 0, t1 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, t1, (const int64_t *)nullptr, (int64_t *)nullptr, (const int64_t *)nullptr,
                                  (int64_t *)nullptr, np, 0, 64, st);
  tmp = std::max(tmp, t1);
  cub::DeviceScan::InclusiveSum(nullptr, t1, rank, rank, np, st); tmp = std::max(tmp, t1);
  cub::DeviceScan::ExclusiveSum(nullptr, t1, cnt, ofs, np, st); tmp = std::max(tmp, t1);
  TRY(ensure(h, d.cubtmp, tmp + 16));
  // The first cell size assumes the field fills its bounding box.  Fields that do not (rings,
  // jets) leave most cells empty and the occupied ones overfull: shrink the cells by the
  // cube root of the overfill and sort again (at most twice; a sort is ~1 ms per million).
  int64_t nl = 0, ncell = 0;
  for (int iter = 0;; ++iter) {
    ncell = g.dims[0] * g.dims[1] * g.dims[2];
    const int key_bits = std::max(1, (int)std::ceil(std::log2((double)std::max<int64_t>(ncell, 2))));
    tree_keys_kernel<<<blocks_for(np, 256), 256, 0, st>>>(d_P, ld, np, g, keys, idx0);
    t1 = d.cubtmp.cap;
    CK(h, cub::DeviceRadixSort::SortPairs(d.cubtmp.p, t1, (const int64_t *)keys, skeys, (const int64_t *)idx0, sidx, np,
                                          0, key_bits, st));
    tree_heads_kernel<<<blocks_for(np, 256), 256, 0, st>>>(skeys, np, rank);
    t1 = d.cubtmp.cap;
    CK(h, cub::DeviceScan::InclusiveSum(d.cubtmp.p, t1, rank, rank, np, st));
    h->launches += 4;
    u64 nl64 = 0;
    CK(h, cudaMemcpyAsync(&nl64, rank + (np - 1), 8, cudaMemcpyDeviceToHost, st));
    CK(h, cudaStreamSynchronize(st));
    nl = (int64_t)nl64;
    const double occ = (double)np / (double)nl;
    if (iter >= 2 || occ <= 0.75 * (double)ncrit) break;
    const double h2 = g.h * std::pow(((double)ncrit / 2.0) / occ, 1.0 / 3.0);
    int64_t dims2[3];
    const double ncell2 = dims_of(h2, dims2);
    if (!(h2 > 0.0) || ncell2 > 1.0e9 || ncell2 > 64.0 * (double)np + 4096.0) break;
    g.h = h2;
    for (int a = 0; a < 3; ++a) g.dims[a] = dims2[a];
  }
  TRY(ensure(h, d.scr2, (size_t)ncell * 4 + 64));
  int32_t *cell_to_leaf = (int32_t *)d.scr2.p;
  tree_fill_i32_kernel<<<blocks_for(ncell, 256), 256, 0, st>>>(cell_to_leaf, ncell, -1);
  tree_leaves_kernel<<<blocks_for(np, 256), 256, 0, st>>>(skeys, rank, np, lbegin, lend, lkey, cell_to_leaf);
  h->launches += 2;
  tree_spheres_kernel<<<blocks_for(nl * 32, 256), 256, 0, st>>>(d_P, ld, osig, sidx, lbegin, lend, nl, ctr, rad, bb);
  h->launches++;
  CK(h, cudaMemcpyAsync(hb, bb, sizeof hb, cudaMemcpyDeviceToHost, st));
  CK(h, cudaStreamSynchronize(st));
  CK(h, cudaGetLastError());
  const double rmax = ordered_to_dbl(hb[6]);
  if (!std::isfinite(rmax)) return fail(h, VPM_EINVAL, "%s: non-finite leaf radius (core sizes)", fn);
  const double reach_d = std::ceil(2.0 * rmax / (theta * g.h)) + 1.0;
  int reach = (int)std::min(reach_d, 1.0e6);
  // no leaf is further than the grid itself
  reach = (int)std::min<int64_t>(reach, std::max(std::max(g.dims[0], g.dims[1]), g.dims[2]));
  {
    // the stencil search costs nl * prod_a min(2 reach + 1, dims_a) MAC tests: refuse fields whose
    // leaf radii (core sizes) are so large against the cell size that this would run for minutes
    double cand = (double)nl;
    for (int a = 0; a < 3; ++a) cand *= (double)std::min<int64_t>(2 * (int64_t)reach + 1, 2 * g.dims[a] - 1);
    if (cand > 1.0e11)
      return fail(h, VPM_EINVAL, "%s: %.2g leaf-pair tests (largest leaf radius %.3g against cell size %.3g): "
                  "core sizes too large for ncrit = %lld, use a larger ncrit", fn, cand, rmax, g.h, (long long)ncrit);
  }
  tree_list_kernel<0><<<blocks_for(nl * 32, 256), 256, 0, st>>>(g, reach, lkey, cell_to_leaf, ctr, rad, nl, cnt, nullptr,
                                                            nullptr, nullptr);
  t1 = d.cubtmp.cap;
  CK(h, cub::DeviceScan::ExclusiveSum(d.cubtmp.p, t1, cnt, ofs, nl, st));
  h->launches += 2;
  u64 last[2] = {0, 0};
  CK(h, cudaMemcpyAsync(&last[0], cnt + (nl - 1), 8, cudaMemcpyDeviceToHost, st));
  CK(h, cudaMemcpyAsync(&last[1], ofs + (nl - 1), 8, cudaMemcpyDeviceToHost, st));
  CK(h, cudaStreamSynchronize(st));
  const int64_t npairs = (int64_t)(last[0] + last[1]);
  TRY(ensure(h, d.tlist, (size_t)std::max<int64_t>(npairs, 1) * 8 + 64));
  Carver cl(d.tlist.p);
  int32_t *pt = cl.take<int32_t>((size_t)std::max<int64_t>(npairs, 1)), *ps = cl.take<int32_t>((size_t)std::max<int64_t>(npairs, 1));
  tree_list_kernel<1><<<blocks_for(nl * 32, 256), 256, 0, st>>>(g, reach, lkey, cell_to_leaf, ctr, rad, nl, cnt, ofs, pt, ps);
  h->launches++;
  CK(h, cudaStreamSynchronize(st));
  CK(h, cudaGetLastError());
  h->tree_np = np; h->tree_nl = nl; h->tree_npairs = npairs;
  return VPM_OK;
}

template <int K>
void launch_uj_leaf_K(int nt, unsigned nwi, const LeafUjArgs &a, cudaStream_t st) {
  if (nt == 32) uj_leaf_kernel<K, 32, 64><<<nwi, 32, 0, st>>>(a);
  else if (nt == 64) uj_leaf_kernel<K, 64, 64><<<nwi, 64, 0, st>>>(a);
  else uj_leaf_kernel<K, 128, 128><<<nwi, 128, 0, st>>>(a);
}
void launch_uj_leaf(int kernel, int nt, unsigned nwi, const LeafUjArgs &a, cudaStream_t st) {
  switch (kernel) {
    case K_SING: launch_uj_leaf_K<K_SING>(nt, nwi, a, st); break;
    case K_GAUS: launch_uj_leaf_K<K_GAUS>(nt, nwi, a, st); break;
    case K_GERF: launch_uj_leaf_K<K_GERF>(nt, nwi, a, st); break;
    default: launch_uj_leaf_K<K_WINCK>(nt, nwi, a, st); break;
  }
}
template <int K, int MODE>
void launch_sfs_leaf_K(int nt, unsigned nwi, const LeafSfsArgs &a, cudaStream_t st) {
  if (nt == 32) sfs_leaf_kernel<K, 32, 64, MODE><<<nwi, 32, 0, st>>>(a);
  else if (nt == 64) sfs_leaf_kernel<K, 64, 64, MODE><<<nwi, 64, 0, st>>>(a);
  else sfs_leaf_kernel<K, 128, 128, MODE><<<nwi, 128, 0, st>>>(a);
}
template <int MODE>
void launch_sfs_leaf_M(int kernel, int nt, unsigned nwi, const LeafSfsArgs &a, cudaStream_t st) {
  switch (kernel) {
    case K_SING: launch_sfs_leaf_K<K_SING, MODE>(nt, nwi, a, st); break;
    case K_GAUS: launch_sfs_leaf_K<K_GAUS, MODE>(nt, nwi, a, st); break;
    case K_GERF: launch_sfs_leaf_K<K_GERF, MODE>(nt, nwi, a, st); break;
    default: launch_sfs_leaf_K<K_WINCK, MODE>(nt, nwi, a, st); break;
  }
}
void launch_sfs_leaf(int kernel, int nt, unsigned nwi, const LeafSfsArgs &a, cudaStream_t st,
                     int mode = MODE_SFS) {
  if (mode == MODE_ZETA) launch_sfs_leaf_M<MODE_ZETA>(kernel, nt, nwi, a, st);
  else launch_sfs_leaf_M<MODE_SFS>(kernel, nt, nwi, a, st);
}



// ---- device-resident field: UJ_direct on the mirror of the whole matrix ----
// With G devices every device holds the whole mirror (np_pad = G * shard columns); device g
// sweeps the targets of its shard and the shards' columns are all-gathered in place over
// NVLink (columns = particles are contiguous in the column-major matrix), so all mirrors
// stay identical and the O(N) step kernels simply run on every device.
int64_t field_shard(const vpm_handle *h) {
  const int64_t G = (int64_t)h->devs.size();
  return (h->fld_np + G - 1) / G;
}

int field_allgather(vpm_handle *h) {
  const int G = (int)h->devs.size();
  if (G < 2) return VPM_OK;
  TRY(ensure_comms(h));
  const size_t count = (size_t)field_shard(h) * h->fld_nf;
  NCK(h, g_nccl.group_start());
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    double *F = (double *)d.fld.p;
    NCK(h, g_nccl.all_gather(F + (size_t)g * count, F, count, kNcclFloat64, h->comms[g], d.stream));
  }
  NCK(h, g_nccl.group_end());
  return VPM_OK;
}

int field_uj(vpm_handle *h, int kernel, int flags) {
  const int64_t nf = h->fld_nf, np = h->fld_np;
  if (np == 0) return VPM_OK;
  const int G = (int)h->devs.size();
  const int64_t shard = field_shard(h);
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    cudaStream_t st = d.stream;
    CK(h, cudaSetDevice(d.id));
    double *F = (double *)d.fld.p;
    const int64_t t0 = std::min(np, g * shard), t1 = std::min(np, t0 + shard), nt = t1 - t0;
    SrcView src{F, nf, 0, 3, 6};
    Plan plan;
    TRY(uj_sweep(h, d, st, kernel, F + t0 * nf, nf, nt, src, 0, np, flags, plan));
    if (nt > 0) {
      UjFinishArgs f;
      f.partial = (const double *)d.partial.p; f.pstride = plan.pstride; f.nsplit = plan.nsplit;
      f.nt = nt; f.out = F + t0 * nf; f.ld = nf; f.urow = R_U; f.jrow = R_J; f.zrow0 = R_W; f.zrow1 = R_PSE;
      f.want_U = 1; f.want_J = 1; f.accumulate = 1; f.reset = (flags & VPM_FLAG_RESET) ? 1 : 0;
      f.stat = F + t0 * nf + R_STATIC; f.sld = nf;
      uj_finish_kernel<<<blocks_for(nt, 256), 256, 0, st>>>(f);
      h->launches++;
      CK(h, cudaGetLastError());
    }
  }
  TRY(field_allgather(h));
  if (flags & VPM_FLAG_SFS) {
    for (int g = 0; g < G; ++g) {
      Dev &d = h->devs[g];
      cudaStream_t st = d.stream;
      CK(h, cudaSetDevice(d.id));
      double *F = (double *)d.fld.p;
      const int64_t t0 = std::min(np, g * shard), t1 = std::min(np, t0 + shard), nt = t1 - t0;
      SrcView src{F, nf, 0, 3, 6};
      Plan sp;
      TRY(sfs_sweep(h, d, st, kernel, F + t0 * nf, nf, F + t0 * nf + R_J, nf, nullptr, nt, src, F, nf, R_J,
                    F + R_STATIC, nf, nullptr, np, flags, sp));
      if (nt > 0) {
        SfsFinishArgs q;
        q.partial = (const double *)d.partial.p; q.pstride = sp.pstride; q.nsplit = sp.nsplit;
        q.nt = nt; q.tindex = nullptr; q.out = F + t0 * nf; q.ld = nf; q.row = R_SFS; q.accumulate = 1;
        q.reset = (flags & VPM_FLAG_RESET_SFS) ? 1 : 0; q.filter_static = 1;
        q.stat = F + t0 * nf + R_STATIC; q.sld = nf;
        sfs_finish_kernel<<<blocks_for(nt, 256), 256, 0, st>>>(q);
        h->launches++;
        CK(h, cudaGetLastError());
      }
    }
    TRY(field_allgather(h));
  } else if (flags & VPM_FLAG_RESET_SFS) {
    for (int g = 0; g < G; ++g) {  // O(N): every device does all particles, no exchange needed
      Dev &d = h->devs[g];
      CK(h, cudaSetDevice(d.id));
      double *F = (double *)d.fld.p;
      zero_rows_kernel<<<blocks_for(np, 256), 256, 0, d.stream>>>(F, nf, R_SFS, 3, np, F + R_STATIC, nf);
      h->launches++;
      CK(h, cudaGetLastError());
    }
  }
  return VPM_OK;
}


// zeta_direct on the resident mirror(s): J[1:3] of every particle <- sum_j Gamma_j zeta_sigma_j
int field_zeta(vpm_handle *h, int kernel) {
  const int64_t nf = h->fld_nf, np = h->fld_np;
  if (np == 0) return VPM_OK;
  const int G = (int)h->devs.size();
  const int64_t shard = field_shard(h);
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    cudaStream_t st = d.stream;
    CK(h, cudaSetDevice(d.id));
    double *F = (double *)d.fld.p;
    const int64_t t0 = std::min(np, g * shard), t1 = std::min(np, t0 + shard), nt = t1 - t0;
    SrcView src{F, nf, 0, 3, 6};
    Plan sp;
    TRY(sfs_sweep(h, d, st, kernel, F + t0 * nf, nf, F + t0 * nf + R_J, nf, nullptr, nt, src, F, nf, R_J, nullptr, 1,
                  nullptr, np, VPM_FLAG_TRANSPOSED, sp, false, MODE_ZETA));
    if (nt > 0) {
      SfsFinishArgs q;
      q.partial = (const double *)d.partial.p; q.pstride = sp.pstride; q.nsplit = sp.nsplit;
      q.nt = nt; q.tindex = nullptr; q.out = F + t0 * nf; q.ld = nf; q.row = R_J; q.accumulate = 0; q.reset = 0;
      q.filter_static = 0; q.stat = nullptr; q.sld = 1;
      sfs_finish_kernel<<<blocks_for(nt, 256), 256, 0, st>>>(q);
      h->launches++;
      CK(h, cudaGetLastError());
    }
  }
  return field_allgather(h);
}

StepArgs step_args_of(vpm_handle *h, Dev &d) {
  StepArgs a{};
  a.P = (double *)d.fld.p; a.nf = h->fld_nf; a.np = h->fld_np;
  return a;
}

// launch one O(N) kernel on every device's mirror
template <class L>
int field_on_all(vpm_handle *h, L launch) {
  for (Dev &d : h->devs) {
    CK(h, cudaSetDevice(d.id));
    launch(d);
    h->launches++;
  }
  CK(h, cudaGetLastError());
  return VPM_OK;
}

// sum over the non-static particles of r_k^2 (mode 0) or Gamma_k J_k (mode 1), from device 0
int field_reduce3(vpm_handle *h, int mode, double out[3]) {
  Dev &d = h->devs[0];
  CK(h, cudaSetDevice(d.id));
  TRY(ensure(h, d.jbuf, (size_t)(kRedBlocks * 3 + 3) * sizeof(double)));
  double *partial = (double *)d.jbuf.p, *res = partial + kRedBlocks * 3;
  rbf_reduce_partial<<<kRedBlocks, 256, 0, d.stream>>>(step_args_of(h, d), mode, partial);
  rbf_reduce_final<<<1, kRedBlocks, 0, d.stream>>>(partial, res);
  h->launches += 2;
  CK(h, cudaGetLastError());
  CK(h, cudaMemcpyAsync(out, res, 3 * sizeof(double), cudaMemcpyDeviceToHost, d.stream));
  CK(h, cudaStreamSynchronize(d.stream));
  return VPM_OK;
}

// rbf_conjugategradient (src/FLOWVPM_viscous.jl:309-478) with cs.zeta = zeta_direct
int field_rbf(vpm_handle *h, int kernel, int itmax, double tol, int iterror, int *iterations, double *residuals) {
  const double eps = 2.220446049250313e-16;
  const unsigned nb = blocks_for(h->fld_np, 256);
  auto stage = [&](int st, const double c[3]) {
    return field_on_all(h, [&](Dev &d) { rbf_stage<<<nb, 256, 0, d.stream>>>(step_args_of(h, d), st, c[0], c[1], c[2]); });
  };
  const double zero3[3] = {0, 0, 0};
  double rr0s[3], rrs[3], prev_rrs[3], pAps[3], alphas[3], betas[3];
  bool flags[3];
  TRY(stage(0, zero3));
  TRY(field_zeta(h, kernel));
  TRY(stage(1, zero3));
  TRY(field_reduce3(h, 0, rr0s));
  for (int k = 0; k < 3; ++k) {
    rrs[k] = rr0s[k];
    flags[k] = sqrt(rr0s[k]) > tol || sqrt(rrs[k] / rr0s[k]) > tol;
  }
  int it_done = 0;
  bool failed = false;
  for (int it = 1; it <= itmax; ++it) {
    if (!(flags[0] || flags[1] || flags[2])) break;
    it_done = it;
    TRY(field_zeta(h, kernel));
    TRY(field_reduce3(h, 1, pAps));
    for (int k = 0; k < 3; ++k) {
      alphas[k] = flags[k] ? rrs[k] / pAps[k] : 0.0;  // Julia: x * false == 0 (strong zero)
      prev_rrs[k] = rrs[k];
    }
    TRY(stage(2, alphas));
    TRY(field_reduce3(h, 0, rrs));
    for (int k = 0; k < 3; ++k) {
      betas[k] = rrs[k] / prev_rrs[k];
      if (fabs(prev_rrs[k]) <= 2 * eps) betas[k] = 1;
    }
    TRY(stage(3, betas));
    for (int k = 0; k < 3; ++k)
      flags[k] = flags[k] && (fabs(rr0s[k]) <= 2 * eps ? false : sqrt(rrs[k] / rr0s[k]) > tol);
    if (it == itmax && (flags[0] || flags[1] || flags[2])) failed = true;
  }
  TRY(stage(4, zero3));
  if (iterations) *iterations = it_done;
  if (residuals)
    for (int k = 0; k < 3; ++k) residuals[k] = rr0s[k] > 0 ? sqrt(rrs[k] / rr0s[k]) : 0.0;
  if (failed && iterror)
    return fail(h, VPM_ESTATE, "Maximum number of iterations %d reached before convergence. Errors: %g %g %g, tolerance: %g",
                itmax, sqrt(rrs[0] / rr0s[0]), sqrt(rrs[1] / rr0s[1]), sqrt(rrs[2] / rr0s[2]), tol);
  return VPM_OK;
}

// viscousdiffusion(pfield, CoreSpreading, dt; aux1, aux2): src/FLOWVPM_viscous.jl:152-223
int field_corespreading(vpm_handle *h, const vpm_step_params *sp, double aux1, double aux2) {
  const unsigned nb = blocks_for(h->fld_np, 256);
  const int rk = sp->integration == 1;
  TRY(field_on_all(h, [&](Dev &d) {
    StepArgs a = step_args_of(h, d);
    a.a = aux1; a.b = aux2; a.dt = sp->dt;
    cs_spread<<<nb, 256, 0, d.stream>>>(a, sp->nu, rk);
  }));
  const bool proceed = !rk || fabs(aux2 - 8.0 / 15) <= 1e-7;
  if (!proceed) return VPM_OK;
  h->fld_t_sgm += sp->dt;
  const double beta_cur = sqrt(2 * sp->nu * h->fld_t_sgm / (sp->sgm0 * sp->sgm0) + 1);
  if (beta_cur >= sp->cs_beta) {
    TRY(field_zeta(h, sp->kernel_id));
    TRY(field_on_all(h, [&](Dev &d) { cs_reset<<<nb, 256, 0, d.stream>>>(step_args_of(h, d), sp->sgm0); }));
    TRY(field_rbf(h, sp->kernel_id, sp->cs_itmax, sp->cs_tol, sp->cs_iterror, nullptr, nullptr));
    h->fld_t_sgm = 0.0;
  }
  return VPM_OK;
}

double zeta0_of(int kernel) {  // kernel.zeta(0): src/FLOWVPM_kernel.jl:45,51,60,69-74
  const double pi = 3.14159265358979323846;
  switch (kernel) {
    case K_SING: return 1.0;
    case K_GAUS: return 3.0 / (4.0 * pi);
    case K_GERF: return 1.0 / pow(2.0 * pi, 1.5);
    default: return 1.0 / (4.0 * pi) * 7.5 / sqrt(1.0);
  }
}

}  // namespace

// ============================================================== C ABI
extern "C" {

int vpm_abi_version(void) { return VPM_ABI_VERSION; }

const char *vpm_last_error(const vpm_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int vpm_num_devices(const vpm_handle *h) { return h ? (int)h->devs.size() : 0; }

int vpm_create(vpm_handle **out, int n_gpus, const int *device_ids) {
  if (!out) return fail(nullptr, VPM_EINVAL, "vpm_create: out is NULL");
  *out = nullptr;
  if (n_gpus < 1 || n_gpus > 64) return fail(nullptr, VPM_EINVAL, "vpm_create: n_gpus=%d", n_gpus);
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count < 1) {
    cudaGetLastError();
    return fail(nullptr, VPM_ENODEV, "vpm_create: no usable CUDA device (%s); there is no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  vpm_handle *h = new (std::nothrow) vpm_handle();
  if (!h) return fail(nullptr, VPM_ENOMEM, "vpm_create: out of host memory");
  for (int g = 0; g < n_gpus; ++g) {
    Dev d;
    d.id = device_ids ? device_ids[g] : g;
    if (d.id < 0 || d.id >= count) {
      int rc = fail(nullptr, VPM_ENODEV, "vpm_create: device %d not present (%d visible)", d.id, count);
      delete h;
      return rc;
    }
    cudaDeviceProp prop;
    if (cudaSetDevice(d.id) != cudaSuccess || cudaGetDeviceProperties(&prop, d.id) != cudaSuccess) {
      int rc = fail(nullptr, VPM_ECUDA, "vpm_create: cannot open device %d: %s", d.id,
                    cudaGetErrorString(cudaGetLastError()));
      delete h;
      return rc;
    }
    if (prop.major < 10) {
      int rc = fail(nullptr, VPM_ENODEV,
                    "vpm_create: device %d is sm_%d%d; libvpm_cuda is built for sm_100a only", d.id,
                    prop.major, prop.minor);
      delete h;
      return rc;
    }
    d.sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking) != cudaSuccess) {
      int rc = fail(nullptr, VPM_ECUDA, "vpm_create: stream: %s", cudaGetErrorString(cudaGetLastError()));
      delete h;
      return rc;
    }
    for (auto &ev : d.ev) cudaEventCreate(&ev);
    h->devs.push_back(d);
  }
  *out = h;
  return VPM_OK;
}

int vpm_destroy(vpm_handle *h) {
  if (!h) return VPM_OK;
  for (auto &p : h->pinned) cudaHostUnregister(p.first);
  if (g_nccl.comm_destroy)
    for (void *c : h->comms) if (c) g_nccl.comm_destroy(c);
  for (Dev &d : h->devs) {
    cudaSetDevice(d.id);
    cudaStreamSynchronize(d.stream);
    for (Buf *b : {&d.in7, &d.stat, &d.res18, &d.sfs3, &d.rec, &d.srec, &d.partial, &d.tbuf, &d.sbuf,
                   &d.ibuf, &d.jbuf, &d.fld, &d.scr, &d.scr2, &d.cubtmp, &d.tree, &d.tlist})
      if (b->p) cudaFree(b->p);
    for (auto &ev : d.ev) if (ev) cudaEventDestroy(ev);
    if (d.stream) cudaStreamDestroy(d.stream);
  }
  if (h->h_stat) cudaFreeHost(h->h_stat);
  if (h->h_stage) cudaFreeHost(h->h_stage);
  cudaGetLastError();
  delete h;
  return VPM_OK;
}

int vpm_pin_host(vpm_handle *h, void *ptr, size_t bytes) {
  if (!h || !ptr || bytes == 0) return fail(h, VPM_EINVAL, "vpm_pin_host: bad argument");
  CK(h, cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
  h->pinned.push_back({ptr, bytes});
  return VPM_OK;
}

int vpm_unpin_host(vpm_handle *h, void *ptr) {
  if (!h || !ptr) return fail(h, VPM_EINVAL, "vpm_unpin_host: bad argument");
  auto it = std::find_if(h->pinned.begin(), h->pinned.end(), [ptr](const std::pair<void *, size_t> &r) { return r.first == ptr; });
  if (it == h->pinned.end()) return fail(h, VPM_EINVAL, "vpm_unpin_host: pointer was not pinned by this handle");
  CK(h, cudaHostUnregister(ptr));
  h->pinned.erase(it);
  return VPM_OK;
}

static int check_field(vpm_handle *h, const char *fn, const void *P, int64_t nf, int64_t np, int kernel) {
  if (!h) return VPM_EINVAL;
  if (np < 0 || nf < MIN_FIELDS) return fail(h, VPM_EINVAL, "%s: need nfields >= 43 and np >= 0 (got %lld, %lld)", fn, (long long)nf, (long long)np);
  if (!P && np > 0) return fail(h, VPM_EINVAL, "%s: particles is NULL", fn);
  if (!valid_kernel(kernel)) return fail(h, VPM_EINVAL, "%s: unknown kernel_id %d", fn, kernel);
  return VPM_OK;
}

int vpm_uj_direct(vpm_handle *h, double *P, int64_t nf, int64_t np, int kernel, int flags) {
  TRY(check_field(h, "vpm_uj_direct", P, nf, np, kernel));
  if (h->devs.size() > 1) return uj_direct_multi(h, P, nf, np, kernel, flags);
  Dev &d = h->devs[0];
  h->launches = 0;
  CK(h, cudaSetDevice(d.id));
  CK(h, cudaEventRecord(d.ev[0], d.stream));
  const bool reset = flags & VPM_FLAG_RESET;
  const bool sfs_rows = (flags & VPM_FLAG_SFS) || (flags & VPM_FLAG_RESET_SFS);
  bool has_static = false;
  // previous SFS rows are needed unless every one of them is overwritten
  TRY(h1_upload(h, d, P, nf, np, !reset, sfs_rows, has_static));
  TRY(h1_eval(h, d, np, kernel, flags, has_static, !reset || has_static));
  TRY(h1_download(h, d, P, nf, np, flags));
  h1_fill_timing(h, d);
  h->np_resident = -1;
  return VPM_OK;
}


int vpm_uj_direct_f32(vpm_handle *h, float *P, int64_t nf, int64_t np, int kernel, int flags) {
  TRY(check_field(h, "vpm_uj_direct_f32", P, nf, np, kernel));
  Dev &d = h->devs[0];
  cudaStream_t st = d.stream;
  h->launches = 0;
  CK(h, cudaSetDevice(d.id));
  CK(h, cudaEventRecord(d.ev[0], st));
  const bool reset = flags & VPM_FLAG_RESET;
  const bool sfs_rows = (flags & VPM_FLAG_SFS) || (flags & VPM_FLAG_RESET_SFS);
  const size_t n = (size_t)std::max<int64_t>(np, 1);
  TRY(ensure(h, d.in7, n * 7 * sizeof(double)));
  TRY(ensure(h, d.res18, n * RES_ROWS * sizeof(double)));
  TRY(ensure(h, d.sfs3, n * 3 * sizeof(double)));
  TRY(ensure(h, d.jbuf, n * (7 + RES_ROWS + 3) * sizeof(float) + 64));
  float *f_in7 = (float *)d.jbuf.p, *f_res = f_in7 + n * 7, *f_sfs = f_res + n * RES_ROWS;
  bool has_static = false;
  has_static = any_static(P, nf, np);
  const bool prior = !reset || has_static;
  if (np > 0) {
    CK(h, cudaMemcpy2DAsync(f_in7, 7 * sizeof(float), P, nf * sizeof(float), 7 * sizeof(float), (size_t)np,
                            cudaMemcpyHostToDevice, st));
    cvt_f32_to_f64_kernel<<<blocks_for(np * 7, 256), 256, 0, st>>>(f_in7, (double *)d.in7.p, np * 7);
    h->launches++;
    if (has_static) {
      if (h->h_stat_cap < (size_t)np) {
        if (h->h_stat) cudaFreeHost(h->h_stat);
        h->h_stat = nullptr; h->h_stat_cap = 0;
        CK(h, cudaMallocHost((void **)&h->h_stat, (size_t)np * sizeof(double)));
        h->h_stat_cap = (size_t)np;
      }
      for (int64_t i = 0; i < np; ++i) h->h_stat[i] = (double)P[nf * i + R_STATIC];
      TRY(ensure(h, d.stat, (size_t)np * sizeof(double)));
      CK(h, cudaMemcpyAsync(d.stat.p, h->h_stat, (size_t)np * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    if (prior) {
      CK(h, cudaMemcpy2DAsync(f_res, RES_ROWS * sizeof(float), P + R_U, nf * sizeof(float),
                              RES_ROWS * sizeof(float), (size_t)np, cudaMemcpyHostToDevice, st));
      cvt_f32_to_f64_kernel<<<blocks_for(np * RES_ROWS, 256), 256, 0, st>>>(f_res, (double *)d.res18.p, np * RES_ROWS);
      h->launches++;
    }
    if (sfs_rows) {
      CK(h, cudaMemcpy2DAsync(f_sfs, 3 * sizeof(float), P + R_SFS, nf * sizeof(float), 3 * sizeof(float),
                              (size_t)np, cudaMemcpyHostToDevice, st));
      cvt_f32_to_f64_kernel<<<blocks_for(np * 3, 256), 256, 0, st>>>(f_sfs, (double *)d.sfs3.p, np * 3);
      h->launches++;
    }
    CK(h, cudaGetLastError());
  }
  TRY(h1_eval(h, d, np, kernel, flags, has_static, prior));
  if (np > 0) {
    cvt_f64_to_f32_kernel<<<blocks_for(np * RES_ROWS, 256), 256, 0, st>>>((const double *)d.res18.p, f_res, np * RES_ROWS);
    h->launches++;
    CK(h, cudaMemcpy2DAsync(P + R_U, nf * sizeof(float), f_res, RES_ROWS * sizeof(float),
                            RES_ROWS * sizeof(float), (size_t)np, cudaMemcpyDeviceToHost, st));
    if (sfs_rows) {
      cvt_f64_to_f32_kernel<<<blocks_for(np * 3, 256), 256, 0, st>>>((const double *)d.sfs3.p, f_sfs, np * 3);
      h->launches++;
      CK(h, cudaMemcpy2DAsync(P + R_SFS, nf * sizeof(float), f_sfs, 3 * sizeof(float), 3 * sizeof(float),
                              (size_t)np, cudaMemcpyDeviceToHost, st));
    }
    CK(h, cudaGetLastError());
  }
  CK(h, cudaEventRecord(d.ev[5], st));
  CK(h, cudaStreamSynchronize(st));
  h1_fill_timing(h, d);
  h->np_resident = -1;
  return VPM_OK;
}

int vpm_upload_state(vpm_handle *h, const double *P, int64_t nf, int64_t np) {
  TRY(check_field(h, "vpm_upload_state", P, nf, np, 0));
  Dev &d = h->devs[0];
  CK(h, cudaSetDevice(d.id));
  CK(h, cudaEventRecord(d.ev[0], d.stream));
  bool has_static = false;
  TRY(h1_upload(h, d, P, nf, np, true, true, has_static));
  CK(h, cudaStreamSynchronize(d.stream));
  h->np_resident = np;
  h->resident_static = has_static;
  h->resident_prior = true;
  return VPM_OK;
}

int vpm_eval(vpm_handle *h, int kernel, int flags) {
  if (!h) return VPM_EINVAL;
  if (h->np_resident < 0) return fail(h, VPM_ESTATE, "vpm_eval: no resident state (call vpm_upload_state first)");
  if (!valid_kernel(kernel)) return fail(h, VPM_EINVAL, "vpm_eval: unknown kernel_id %d", kernel);
  Dev &d = h->devs[0];
  h->launches = 0;
  CK(h, cudaSetDevice(d.id));
  CK(h, cudaEventRecord(d.ev[0], d.stream));
  TRY(h1_eval(h, d, h->np_resident, kernel, flags, h->resident_static, true));
  CK(h, cudaEventRecord(d.ev[5], d.stream));
  CK(h, cudaStreamSynchronize(d.stream));
  h1_fill_timing(h, d);
  return VPM_OK;
}

int vpm_download_results(vpm_handle *h, double *P, int64_t nf, int64_t np, int flags) {
  TRY(check_field(h, "vpm_download_results", P, nf, np, 0));
  if (h->np_resident != np) return fail(h, VPM_ESTATE, "vpm_download_results: np=%lld but %lld particles are resident", (long long)np, (long long)h->np_resident);
  return h1_download(h, h->devs[0], P, nf, np, flags | VPM_FLAG_SFS);
}

int vpm_uj_direct_st(vpm_handle *h, const double *S, int64_t nfs, int64_t nps, double *Tg, int64_t nft,
                     int64_t npt, int kernel) {
  TRY(check_field(h, "vpm_uj_direct_st(source)", S, nfs, nps, kernel));
  TRY(check_field(h, "vpm_uj_direct_st(target)", Tg, nft, npt, kernel));
  Dev &d = h->devs[0];
  cudaStream_t st = d.stream;
  h->launches = 0;
  CK(h, cudaSetDevice(d.id));
  if (npt == 0) return VPM_OK;
  TRY(ensure(h, d.in7, (size_t)std::max<int64_t>(nps, 1) * 7 * sizeof(double)));
  TRY(ensure(h, d.tbuf, (size_t)npt * 3 * sizeof(double)));
  TRY(ensure(h, d.res18, (size_t)npt * RES_ROWS * sizeof(double)));
  CK(h, cudaEventRecord(d.ev[0], st));
  if (nps > 0)
    TRY(h2d_rows(h, st, (double *)d.in7.p, S, nfs, 7, nps));
  TRY(h2d_rows(h, st, (double *)d.tbuf.p, Tg, nft, 3, npt));
  TRY(h2d_rows(h, st, (double *)d.res18.p, Tg + R_U, nft, RES_ROWS, npt));
  CK(h, cudaEventRecord(d.ev[1], st));
  SrcView src{(const double *)d.in7.p, 7, 0, 3, 6};
  Plan plan;
  TRY(uj_sweep(h, d, st, kernel, (const double *)d.tbuf.p, 3, npt, src, 0, nps, 0, plan));
  CK(h, cudaEventRecord(d.ev[2], st));
  UjFinishArgs f;
  f.partial = (const double *)d.partial.p; f.pstride = plan.pstride; f.nsplit = plan.nsplit;
  f.nt = npt; f.out = (double *)d.res18.p; f.ld = RES_ROWS; f.urow = RES_U; f.jrow = RES_J;
  f.zrow0 = -1; f.zrow1 = -1; f.want_U = 1; f.want_J = 1; f.accumulate = 1; f.reset = 0;
  f.stat = nullptr; f.sld = 1;
  uj_finish_kernel<<<blocks_for(npt, 256), 256, 0, st>>>(f);
  h->launches++;
  CK(h, cudaGetLastError());
  CK(h, cudaEventRecord(d.ev[3], st));
  CK(h, cudaEventRecord(d.ev[4], st));
  CK(h, cudaMemcpy2DAsync(Tg + R_U, nft * sizeof(double), d.res18.p, RES_ROWS * sizeof(double),
                          RES_ROWS * sizeof(double), (size_t)npt, cudaMemcpyDeviceToHost, st));
  CK(h, cudaEventRecord(d.ev[5], st));
  CK(h, cudaStreamSynchronize(st));
  h->timing.uj_pairs = nps * npt;
  h->timing.sfs_pairs = 0;
  h1_fill_timing(h, d);
  h->np_resident = -1;
  return VPM_OK;
}

int vpm_p2p_buffers(vpm_handle *h, double *tgt, int64_t ld, int64_t t0, int64_t t1, int row_pos,
                    int row_grad, int row_hess, const double *src, int64_t s0, int64_t s1, int kernel,
                    int want_U, int want_J) {
  if (!h) return VPM_EINVAL;
  if (!valid_kernel(kernel)) return fail(h, VPM_EINVAL, "vpm_p2p_buffers: unknown kernel_id %d", kernel);
  if (t0 < 0 || t1 < t0 || s0 < 0 || s1 < s0 || ld < 3)
    return fail(h, VPM_EINVAL, "vpm_p2p_buffers: bad ranges [%lld,%lld) [%lld,%lld) ld=%lld", (long long)t0, (long long)t1, (long long)s0, (long long)s1, (long long)ld);
  if (row_pos < 0 || row_pos + 3 > ld || (want_U && (row_grad < 0 || row_grad + 3 > ld)) ||
      (want_J && (row_hess < 0 || row_hess + 9 > ld)))
    return fail(h, VPM_EINVAL, "vpm_p2p_buffers: row offsets outside the %lld-row target buffer", (long long)ld);
  const int64_t nt = t1 - t0, ns = s1 - s0;
  if (nt == 0 || ns == 0 || (!want_U && !want_J)) return VPM_OK;
  if (!tgt || !src) return fail(h, VPM_EINVAL, "vpm_p2p_buffers: NULL buffer");
  Dev &d = h->devs[0];
  cudaStream_t st = d.stream;
  h->launches = 0;
  CK(h, cudaSetDevice(d.id));
  TRY(ensure(h, d.tbuf, (size_t)nt * ld * sizeof(double)));
  TRY(ensure(h, d.sbuf, (size_t)ns * 8 * sizeof(double)));
  CK(h, cudaEventRecord(d.ev[0], st));
  CK(h, cudaMemcpyAsync(d.tbuf.p, tgt + t0 * ld, (size_t)nt * ld * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(h, cudaMemcpyAsync(d.sbuf.p, src + s0 * 8, (size_t)ns * 8 * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(h, cudaEventRecord(d.ev[1], st));
  SrcView sv{(const double *)d.sbuf.p, 8, 0, 4, 7};
  Plan plan;
  TRY(uj_sweep(h, d, st, kernel, (const double *)d.tbuf.p + row_pos, ld, nt, sv, 0, ns, 0, plan));
  CK(h, cudaEventRecord(d.ev[2], st));
  UjFinishArgs f;
  f.partial = (const double *)d.partial.p; f.pstride = plan.pstride; f.nsplit = plan.nsplit;
  f.nt = nt; f.out = (double *)d.tbuf.p; f.ld = ld; f.urow = row_grad; f.jrow = row_hess;
  f.zrow0 = -1; f.zrow1 = -1; f.want_U = want_U; f.want_J = want_J; f.accumulate = 1; f.reset = 0;
  f.stat = nullptr; f.sld = 1;
  uj_finish_kernel<<<blocks_for(nt, 256), 256, 0, st>>>(f);
  h->launches++;
  CK(h, cudaGetLastError());
  CK(h, cudaEventRecord(d.ev[3], st));
  CK(h, cudaEventRecord(d.ev[4], st));
  CK(h, cudaMemcpyAsync(tgt + t0 * ld, d.tbuf.p, (size_t)nt * ld * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(h, cudaEventRecord(d.ev[5], st));
  CK(h, cudaStreamSynchronize(st));
  h->timing.uj_pairs = nt * ns;
  h->timing.sfs_pairs = 0;
  h1_fill_timing(h, d);
  h->np_resident = -1;
  return VPM_OK;
}

int vpm_uj_device(vpm_handle *h, const double *d_src8, int64_t ns, int64_t t0, int64_t t1,
                  double *d_out12, int kernel, int flags, void *stream) {
  if (!h) return VPM_EINVAL;
  if (!valid_kernel(kernel)) return fail(h, VPM_EINVAL, "vpm_uj_device: unknown kernel_id %d", kernel);
  if (ns < 0 || t0 < 0 || t1 < t0 || t1 > ns) return fail(h, VPM_EINVAL, "vpm_uj_device: bad target range [%lld,%lld) of %lld", (long long)t0, (long long)t1, (long long)ns);
  const int64_t nt = t1 - t0;
  if (nt == 0) return VPM_OK;
  if (!d_src8 || !d_out12) return fail(h, VPM_EINVAL, "vpm_uj_device: NULL device pointer");
  Dev &d = h->devs[0];
  cudaStream_t st = (cudaStream_t)stream;  // as given: NULL is CUDA's default stream
  h->launches = 0;
  CK(h, cudaSetDevice(d.id));
  SrcView sv{d_src8, 8, 0, 4, 7};
  Plan plan;
  TRY(uj_sweep(h, d, st, kernel, d_src8 + t0 * 8, 8, nt, sv, 0, ns, flags, plan, true));
  h->device_timing = 1;
  UjFinishArgs f;
  f.partial = (const double *)d.partial.p; f.pstride = plan.pstride; f.nsplit = plan.nsplit;
  f.nt = nt; f.out = d_out12; f.ld = 12; f.urow = 0; f.jrow = 3;
  f.zrow0 = -1; f.zrow1 = -1; f.want_U = 1; f.want_J = 1; f.accumulate = 0; f.reset = 0;
  f.stat = nullptr; f.sld = 1;
  uj_finish_kernel<<<blocks_for(nt, 256), 256, 0, st>>>(f);
  h->launches++;
  CK(h, cudaGetLastError());
  h->timing.uj_pairs = nt * ns;
  h->timing.kernel_launches = h->launches;
  return VPM_OK;
}

int vpm_sfs_device(vpm_handle *h, const double *d_src8, const double *d_J9, const double *d_static,
                   int64_t ns, int64_t t0, int64_t t1, double *d_out3, int kernel, int flags,
                   void *stream) {
  if (!h) return VPM_EINVAL;
  if (!valid_kernel(kernel)) return fail(h, VPM_EINVAL, "vpm_sfs_device: unknown kernel_id %d", kernel);
  if (ns < 0 || t0 < 0 || t1 < t0 || t1 > ns) return fail(h, VPM_EINVAL, "vpm_sfs_device: bad target range");
  const int64_t nt = t1 - t0;
  if (nt == 0) return VPM_OK;
  if (!d_src8 || !d_J9 || !d_out3) return fail(h, VPM_EINVAL, "vpm_sfs_device: NULL device pointer");
  Dev &d = h->devs[0];
  cudaStream_t st = (cudaStream_t)stream;  // as given: NULL is CUDA's default stream
  h->launches = 0;
  CK(h, cudaSetDevice(d.id));
  SrcView sv{d_src8, 8, 0, 4, 7};
  Plan plan;
  TRY(sfs_sweep(h, d, st, kernel, d_src8 + t0 * 8, 8, d_J9 + t0 * 9, 9, nullptr, nt, sv, d_J9, 9, 0,
                d_static, 1, nullptr, ns, flags, plan, true));
  h->device_timing = 2;
  SfsFinishArgs f;
  f.partial = (const double *)d.partial.p; f.pstride = plan.pstride; f.nsplit = plan.nsplit;
  f.nt = nt; f.tindex = nullptr; f.out = d_out3; f.ld = 3; f.row = 0; f.accumulate = 0; f.reset = 0;
  f.filter_static = 0;  // static targets get an (ignored) value; the caller masks them
  f.stat = nullptr; f.sld = 1;
  sfs_finish_kernel<<<blocks_for(nt, 256), 256, 0, st>>>(f);
  h->launches++;
  CK(h, cudaGetLastError());
  h->timing.sfs_pairs = nt * ns;
  h->timing.kernel_launches = h->launches;
  return VPM_OK;
}


int vpm_p2p_leafpairs(vpm_handle *h, double *tgt, int64_t ld, int64_t n_tgt, int row_pos, int row_grad,
                      int row_hess, const double *src, int64_t n_src, const int64_t *tb,
                      const int64_t *te, int64_t ntl, const int64_t *sb, const int64_t *se, int64_t nsl,
                      const int32_t *pt, const int32_t *ps, int64_t npairs, int kernel, int want_U,
                      int want_J) {
  if (!h) return VPM_EINVAL;
  const char *fn = "vpm_p2p_leafpairs";
  if (!valid_kernel(kernel)) return fail(h, VPM_EINVAL, "%s: unknown kernel_id %d", fn, kernel);
  if (n_tgt < 0 || n_src < 0 || ntl < 0 || nsl < 0 || npairs < 0 || ld < 3)
    return fail(h, VPM_EINVAL, "%s: negative size or ld < 3", fn);
  if (row_pos < 0 || row_pos + 3 > ld || (want_U && (row_grad < 0 || row_grad + 3 > ld)) ||
      (want_J && (row_hess < 0 || row_hess + 9 > ld)))
    return fail(h, VPM_EINVAL, "%s: row offsets outside the %lld-row target buffer", fn, (long long)ld);
  if (npairs == 0 || n_tgt == 0 || n_src == 0 || (!want_U && !want_J)) return VPM_OK;
  if (!tgt || !src || !tb || !te || !sb || !se || !pt || !ps) return fail(h, VPM_EINVAL, "%s: NULL argument", fn);
  h->launches = 0;
  // Multi-GPU: target leaves are sharded into G contiguous runs of work items
  // with balanced  sum nt*ns ; sources are replicated.  Needs the leaves in increasing,
  // non-overlapping body order (tree-sorted buffers) so that a device's targets are one
  // contiguous column range; otherwise device 0 does everything.
  int G = (int)h->devs.size();
  for (int64_t l = 0; l + 1 < ntl && G > 1; ++l)
    if (tb[l + 1] < te[l]) G = 1;
  const int64_t ns_pad = round_up(n_src, kTile);
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    CK(h, cudaSetDevice(d.id));
    TRY(ensure(h, d.tbuf, (size_t)n_tgt * ld * sizeof(double)));
    TRY(ensure(h, d.sbuf, (size_t)n_src * 8 * sizeof(double)));
    TRY(ensure(h, d.rec, (size_t)ns_pad * kRec * sizeof(double)));
  }
  // replicated inputs (source buffer, CSR tables): device 0 gets them from the host, the
  // other devices over NVLink.  The list is regrouped on device 0 (vpm_csr.cuh).
  Dev &d0 = h->devs[0];
  CK(h, cudaSetDevice(d0.id));
  CK(h, cudaEventRecord(d0.ev[0], d0.stream));
  CK(h, cudaMemcpyAsync(d0.sbuf.p, src, (size_t)n_src * 8 * sizeof(double), cudaMemcpyHostToDevice, d0.stream));
  DevCsr c;
  TRY(build_csr_device(h, fn, tb, te, ntl, n_tgt, sb, se, nsl, n_src, pt, ps, npairs, G, nullptr, 0, nullptr, 0, c));
  if (c.nwi == 0) return VPM_OK;
  std::vector<LeafCsr> csr(G);
  csr[0] = c.csr;
  for (int g = 1; g < G; ++g) {
    Dev &d = h->devs[g];
    CK(h, cudaSetDevice(d.id));
    TRY(ensure(h, d.ibuf, h->devs[0].ibuf.cap));
    csr[g] = rebase_csr(c.csr, h->devs[0].ibuf.p, d.ibuf.p);
  }
  TRY(bcast_from_dev0(h, &Dev::sbuf, (size_t)n_src * 8 * sizeof(double)));
  TRY(bcast_from_dev0(h, &Dev::ibuf, c.bcast_bytes));
  std::vector<std::pair<int64_t, int64_t>> cols(G, {0, 0});
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    cudaStream_t st = d.stream;
    const int64_t k0 = c.cut[g], k1 = c.cut[g + 1];
    if (k1 <= k0) continue;
    CK(h, cudaSetDevice(d.id));
    // this device's target columns: first target of its first item .. last target of its last
    const int64_t lf = c.first_leaf[g], ll = c.last_leaf[g + 1];
    const int64_t col0 = G == 1 ? 0 : tb[lf] + c.first_off[g];
    const int64_t col1 = G == 1 ? n_tgt : std::min<int64_t>(te[ll], tb[ll] + c.last_off[g + 1] + c.nt);
    CK(h, cudaMemcpyAsync((double *)d.tbuf.p + col0 * ld, tgt + col0 * ld, (size_t)(col1 - col0) * ld * sizeof(double),
                          cudaMemcpyHostToDevice, st));
    LeafUjArgs a;
    a.csr = csr[g];
    a.csr.wi_leaf += k0;
    a.csr.wi_off += k0;
    if (g == 0) CK(h, cudaEventRecord(d.ev[1], st));
    SrcView sv{(const double *)d.sbuf.p, 8, 0, 4, 7};
    prep_uj_records<<<blocks_for(ns_pad, 256), 256, 0, st>>>(sv, 0, n_src, ns_pad, kernel, (double *)d.rec.p);
    h->launches++;
    a.tpos = (const double *)d.tbuf.p + row_pos; a.tld = ld; a.rec = (const double *)d.rec.p;
    a.out = (double *)d.tbuf.p; a.urow = row_grad; a.jrow = row_hess; a.want_U = want_U; a.want_J = want_J;
    a.shortcut = 1;
    launch_uj_leaf(kernel, c.nt, (unsigned)(k1 - k0), a, st);
    h->launches++;
    CK(h, cudaGetLastError());
    if (g == 0) {
      CK(h, cudaEventRecord(d.ev[2], st));
      CK(h, cudaEventRecord(d.ev[3], st));
      CK(h, cudaEventRecord(d.ev[4], st));
    }
    cols[g] = {col0, col1};
  }
  // downloads in a second pass: a D2H into pageable memory blocks the host, and every
  // device must have its kernel in flight before that happens
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    if (cols[g].second <= cols[g].first) continue;
    CK(h, cudaSetDevice(d.id));
    const int64_t col0 = cols[g].first, col1 = cols[g].second;
    CK(h, cudaMemcpyAsync(tgt + col0 * ld, (double *)d.tbuf.p + col0 * ld, (size_t)(col1 - col0) * ld * sizeof(double),
                          cudaMemcpyDeviceToHost, d.stream));
    if (g == 0) CK(h, cudaEventRecord(d.ev[5], d.stream));
  }
  for (int g = G - 1; g >= 0; --g) {
    CK(h, cudaSetDevice(h->devs[g].id));
    CK(h, cudaStreamSynchronize(h->devs[g].stream));
  }
  h->timing.uj_pairs = c.pairs;
  h->timing.sfs_pairs = 0;
  h1_fill_timing(h, h->devs[0]);
  h->np_resident = -1;
  return VPM_OK;
}

static int leafpairs_field(vpm_handle *h, const char *fn, int mode, double *P, int64_t nf, int64_t np,
                           const int64_t *tsort, const int64_t *ssort, const int64_t *tb, const int64_t *te,
                           int64_t ntl, const int64_t *sb, const int64_t *se, int64_t nsl, const int32_t *pt,
                           const int32_t *ps, int64_t npairs, int kernel, int flags) {
  TRY(check_field(h, fn, P, nf, np, kernel));
  if (ntl < 0 || nsl < 0 || npairs < 0) return fail(h, VPM_EINVAL, "%s: negative size", fn);
  if (npairs == 0 || np == 0) return VPM_OK;
  if (!tsort || !ssort || !tb || !te || !sb || !se || !pt || !ps) return fail(h, VPM_EINVAL, "%s: NULL argument", fn);
  for (int64_t i = 0; i < np; ++i)
    if (tsort[i] < 0 || tsort[i] >= np || ssort[i] < 0 || ssort[i] >= np)
      return fail(h, VPM_EINVAL, "%s: sort index %lld out of range", fn, (long long)i);
  Dev &d = h->devs[0];
  cudaStream_t st = d.stream;
  h->launches = 0;
  // Multi-GPU: as Hook 3 -- target leaves sharded over the devices in contiguous runs of work
  // items; needs increasing, non-overlapping target leaves (tree-sorted), else device 0 alone
  int G = (int)h->devs.size();
  for (int64_t l = 0; l + 1 < ntl && G > 1; ++l)
    if (tb[l + 1] < te[l]) G = 1;
  const int64_t ns_pad = round_up(np, kTile);
  for (int g = 0; g < G; ++g) {
    Dev &dg = h->devs[g];
    CK(h, cudaSetDevice(dg.id));
    TRY(ensure(h, dg.in7, (size_t)np * 7 * sizeof(double)));
    TRY(ensure(h, dg.jbuf, (size_t)np * 9 * sizeof(double)));
    TRY(ensure(h, dg.srec, (size_t)ns_pad * kSfsRec * sizeof(double)));
    if (G > 1) TRY(ensure(h, dg.tbuf, (size_t)np * 3 * sizeof(double)));
  }
  CK(h, cudaSetDevice(d.id));
  TRY(ensure(h, d.sfs3, (size_t)np * 3 * sizeof(double)));
  CK(h, cudaEventRecord(d.ev[0], st));
  TRY(h2d_rows(h, st, (double *)d.in7.p, P, nf, 7, np));
  TRY(h2d_rows(h, st, (double *)d.jbuf.p, P + R_J, nf, 9, np));
  const int out_row = mode == MODE_ZETA ? R_J : R_SFS;
  TRY(h2d_rows(h, st, (double *)d.sfs3.p, P + out_row, nf, 3, np));
  DevCsr c;
  TRY(build_csr_device(h, fn, tb, te, ntl, np, sb, se, nsl, np, pt, ps, npairs, G, tsort, np, ssort, np, c));
  if (c.nwi == 0) return VPM_OK;
  CK(h, cudaEventRecord(d.ev[1], st));
  CK(h, cudaEventRecord(d.ev[2], st));
  CK(h, cudaEventRecord(d.ev[3], st));
  for (int g = 1; g < G; ++g) {
    CK(h, cudaSetDevice(h->devs[g].id));
    TRY(ensure(h, h->devs[g].ibuf, d.ibuf.cap));
  }
  TRY(bcast_from_dev0(h, &Dev::in7, (size_t)np * 7 * sizeof(double)));
  TRY(bcast_from_dev0(h, &Dev::jbuf, (size_t)np * 9 * sizeof(double)));
  TRY(bcast_from_dev0(h, &Dev::ibuf, c.bcast_bytes));
  const int transposed = (flags & VPM_FLAG_TRANSPOSED) ? 1 : 0;
  std::vector<std::pair<int64_t, int64_t>> cols(G, {0, 0});
  for (int g = 0; g < G; ++g) {
    Dev &dg = h->devs[g];
    cudaStream_t sg = dg.stream;
    const int64_t k0 = c.cut[g], k1 = c.cut[g + 1];
    if (k1 <= k0) continue;
    CK(h, cudaSetDevice(dg.id));
    const ptrdiff_t shift = (const char *)dg.ibuf.p - (const char *)d.ibuf.p;
    const int64_t *dts = (const int64_t *)((const char *)c.d_tsort + shift);
    const int64_t *dss = (const int64_t *)((const char *)c.d_ssort + shift);
    SrcView sv{(const double *)dg.in7.p, 7, 0, 3, 6};
    prep_sfs_records<<<blocks_for(ns_pad, 256), 256, 0, sg>>>(sv, (const double *)dg.jbuf.p, 9, 0, nullptr, 1, dss, np,
                                                              ns_pad, kernel, transposed, (double *)dg.srec.p);
    LeafSfsArgs a;
    a.csr = g == 0 ? c.csr : rebase_csr(c.csr, d.ibuf.p, dg.ibuf.p);
    a.csr.wi_leaf += k0;
    a.csr.wi_off += k0;
    a.tpos = (const double *)dg.in7.p; a.tld = 7; a.tJ = (const double *)dg.jbuf.p; a.jld = 9;
    a.tindex = dts; a.rec = (const double *)dg.srec.p; a.old = 3; a.orow = 0;
    a.transposed = transposed;
    a.shortcut = (flags & VPM_FLAG_NO_FARFIELD_SHORTCUT) ? 0 : 1;
    if (G == 1) {
      a.out = (double *)dg.sfs3.p;  // particle-indexed, accumulated in place
    } else {
      // sums land in a zeroed buffer indexed by sorted body; the columns return to device 0
      const int64_t lf = c.first_leaf[g], ll = c.last_leaf[g + 1];
      const int64_t col0 = tb[lf] + c.first_off[g];
      const int64_t col1 = std::min<int64_t>(te[ll], tb[ll] + c.last_off[g + 1] + c.nt);
      cols[g] = {col0, col1};
      CK(h, cudaMemsetAsync((double *)dg.tbuf.p + col0 * 3, 0, (size_t)(col1 - col0) * 3 * sizeof(double), sg));
      a.out = (double *)dg.tbuf.p;
      a.obody = 1;
    }
    launch_sfs_leaf(kernel, c.nt, (unsigned)(k1 - k0), a, sg, mode);
    h->launches += 2;
    CK(h, cudaGetLastError());
  }
  if (G > 1) {
    NCK(h, g_nccl.group_start());
    for (int g = 1; g < G; ++g) {
      const int64_t col0 = cols[g].first, col1 = cols[g].second;
      if (col1 <= col0) continue;
      const size_t cnt = (size_t)(col1 - col0) * 3;
      NCK(h, g_nccl.send((double *)h->devs[g].tbuf.p + col0 * 3, cnt, kNcclFloat64, 0, h->comms[g], h->devs[g].stream));
      NCK(h, g_nccl.recv((double *)d.tbuf.p + col0 * 3, cnt, kNcclFloat64, g, h->comms[0], st));
    }
    NCK(h, g_nccl.group_end());
    CK(h, cudaSetDevice(d.id));
    for (int g = 0; g < G; ++g) {
      const int64_t col0 = cols[g].first, col1 = cols[g].second;
      if (col1 <= col0) continue;
      add_sorted3_kernel<<<blocks_for(col1 - col0, 256), 256, 0, st>>>((const double *)d.tbuf.p, c.d_tsort, col0, col1,
                                                                      (double *)d.sfs3.p);
      h->launches++;
    }
    CK(h, cudaGetLastError());
  }
  CK(h, cudaSetDevice(d.id));
  CK(h, cudaEventRecord(d.ev[4], st));
  CK(h, cudaMemcpy2DAsync(P + out_row, nf * sizeof(double), d.sfs3.p, 3 * sizeof(double), 3 * sizeof(double),
                          (size_t)np, cudaMemcpyDeviceToHost, st));
  CK(h, cudaEventRecord(d.ev[5], st));
  for (int g = G - 1; g >= 0; --g) {
    CK(h, cudaSetDevice(h->devs[g].id));
    CK(h, cudaStreamSynchronize(h->devs[g].stream));
  }
  h->timing.uj_pairs = 0;
  h->timing.sfs_pairs = c.pairs;
  h1_fill_timing(h, d);
  h->np_resident = -1;
  return VPM_OK;
}


int vpm_estr_leafpairs(vpm_handle *h, double *P, int64_t nf, int64_t np, const int64_t *tsort,
                       const int64_t *ssort, const int64_t *tb, const int64_t *te, int64_t ntl,
                       const int64_t *sb, const int64_t *se, int64_t nsl, const int32_t *pt,
                       const int32_t *ps, int64_t npairs, int kernel, int flags) {
  return leafpairs_field(h, "vpm_estr_leafpairs", MODE_SFS, P, nf, np, tsort, ssort, tb, te, ntl, sb, se, nsl, pt,
                         ps, npairs, kernel, flags);
}

int vpm_zeta_leafpairs(vpm_handle *h, double *P, int64_t nf, int64_t np, const int64_t *sort_index,
                       const int64_t *lb, const int64_t *le, int64_t nl, const int32_t *pair_a,
                       const int32_t *pair_b, int64_t npairs, int kernel) {
  // zeta_fmm (src/FLOWVPM_viscous.jl:523-558): for a list entry (a, b) the bodies of leaf b
  // RECEIVE Gamma_j zeta_j from the bodies j of leaf a -> receivers are indexed by the second
  // element, givers by the first.
  return leafpairs_field(h, "vpm_zeta_leafpairs", MODE_ZETA, P, nf, np, sort_index, sort_index, lb, le, nl, lb, le,
                         nl, pair_b, pair_a, npairs, kernel, 0);
}

int vpm_zeta_direct(vpm_handle *h, double *P, int64_t nf, int64_t np, int kernel) {
  TRY(check_field(h, "vpm_zeta_direct", P, nf, np, kernel));
  if (np == 0) return VPM_OK;
  Dev &d = h->devs[0];
  cudaStream_t st = d.stream;
  h->launches = 0;
  CK(h, cudaSetDevice(d.id));
  TRY(ensure(h, d.in7, ((size_t)np * 7 + 16) * sizeof(double)));
  TRY(ensure(h, d.sfs3, (size_t)np * 3 * sizeof(double)));
  CK(h, cudaEventRecord(d.ev[0], st));
  TRY(h2d_rows(h, st, (double *)d.in7.p, P, nf, 7, np));
  CK(h, cudaEventRecord(d.ev[1], st));
  CK(h, cudaEventRecord(d.ev[2], st));
  CK(h, cudaEventRecord(d.ev[3], st));
  SrcView src{(const double *)d.in7.p, 7, 0, 3, 6};
  Plan sp;
  // the J operand is unused in zeta mode: the record builder reads 9 doubles per source
  // from it, so point it at in7 (stride 7; the allocation has 16 doubles of slack)
  TRY(sfs_sweep(h, d, st, kernel, (const double *)d.in7.p, 7, (const double *)d.in7.p, 7, nullptr, np, src,
                (const double *)d.in7.p, 7, 0, nullptr, 1, nullptr, np, VPM_FLAG_TRANSPOSED, sp, false, MODE_ZETA));
  SfsFinishArgs f;
  f.partial = (const double *)d.partial.p; f.pstride = sp.pstride; f.nsplit = sp.nsplit;
  f.nt = np; f.tindex = nullptr; f.out = (double *)d.sfs3.p; f.ld = 3; f.row = 0;
  f.accumulate = 0; f.reset = 0;  // zeta_direct zeroes J[1:3] of every particle first (:487-489)
  f.filter_static = 0; f.stat = nullptr; f.sld = 1;
  sfs_finish_kernel<<<blocks_for(np, 256), 256, 0, st>>>(f);
  h->launches++;
  CK(h, cudaGetLastError());
  CK(h, cudaEventRecord(d.ev[4], st));
  CK(h, cudaMemcpy2DAsync(P + R_J, nf * sizeof(double), d.sfs3.p, 3 * sizeof(double), 3 * sizeof(double),
                          (size_t)np, cudaMemcpyDeviceToHost, st));
  CK(h, cudaEventRecord(d.ev[5], st));
  CK(h, cudaStreamSynchronize(st));
  h->timing.uj_pairs = 0;
  h->timing.sfs_pairs = np * np;
  h1_fill_timing(h, d);
  h->np_resident = -1;
  return VPM_OK;
}

int vpm_leaflists_build(vpm_handle *h, const double *P, int64_t nf, int64_t np, int64_t ncrit, double theta,
                        int64_t *n_leaves, int64_t *n_pairs) {
  TRY(check_field(h, "vpm_leaflists_build", P, nf, np, 0));
  if (ncrit < 1 || !(theta > 0.0)) return fail(h, VPM_EINVAL, "vpm_leaflists_build: ncrit >= 1 and theta > 0 required");
  h->launches = 0;
  h->tree_np = -1;
  if (n_leaves) *n_leaves = 0;
  if (n_pairs) *n_pairs = 0;
  if (np == 0) { h->tree_np = 0; h->tree_nl = 0; h->tree_npairs = 0; return VPM_OK; }
  Dev &d = h->devs[0];
  bool has_static = false;
  CK(h, cudaSetDevice(d.id));
  CK(h, cudaEventRecord(d.ev[0], d.stream));
  TRY(h1_upload(h, d, P, nf, np, false, false, has_static));
  CK(h, cudaEventRecord(d.ev[1], d.stream));
  TRY(tree_build(h, (const double *)d.in7.p, 7, 6, np, ncrit, theta));
  for (int e = 2; e <= 5; ++e) CK(h, cudaEventRecord(d.ev[e], d.stream));
  CK(h, cudaStreamSynchronize(d.stream));
  h->timing.uj_pairs = 0; h->timing.sfs_pairs = 0;
  h1_fill_timing(h, d);
  h->np_resident = -1;
  if (n_leaves) *n_leaves = h->tree_nl;
  if (n_pairs) *n_pairs = h->tree_npairs;
  return VPM_OK;
}

int vpm_leaflists_get(vpm_handle *h, int64_t *sort_index, int64_t *leaf_begin, int64_t *leaf_end, int32_t *pair_tgt,
                      int32_t *pair_src) {
  if (!h) return VPM_EINVAL;
  if (h->tree_np < 0) return fail(h, VPM_ESTATE, "vpm_leaflists_get: no leaf lists (call vpm_leaflists_build first)");
  if (h->tree_np == 0) return VPM_OK;
  Dev &d = h->devs[0];
  CK(h, cudaSetDevice(d.id));
  const TreeView v = tree_view(h);
  if (sort_index) CK(h, cudaMemcpyAsync(sort_index, v.sidx, (size_t)h->tree_np * 8, cudaMemcpyDeviceToHost, d.stream));
  if (leaf_begin) CK(h, cudaMemcpyAsync(leaf_begin, v.lbegin, (size_t)h->tree_nl * 8, cudaMemcpyDeviceToHost, d.stream));
  if (leaf_end) CK(h, cudaMemcpyAsync(leaf_end, v.lend, (size_t)h->tree_nl * 8, cudaMemcpyDeviceToHost, d.stream));
  if (pair_tgt && h->tree_npairs) CK(h, cudaMemcpyAsync(pair_tgt, v.pt, (size_t)h->tree_npairs * 4, cudaMemcpyDeviceToHost, d.stream));
  if (pair_src && h->tree_npairs) CK(h, cudaMemcpyAsync(pair_src, v.ps, (size_t)h->tree_npairs * 4, cudaMemcpyDeviceToHost, d.stream));
  CK(h, cudaStreamSynchronize(d.stream));
  return VPM_OK;
}

int vpm_uj_nearfield(vpm_handle *h, double *P, int64_t nf, int64_t np, int kernel, int flags) {
  const char *fn = "vpm_uj_nearfield";
  TRY(check_field(h, fn, P, nf, np, kernel));
  if (h->tree_np != np) return fail(h, VPM_ESTATE, "%s: leaf lists were built for %lld particles, field has %lld (call vpm_leaflists_build)", fn, (long long)h->tree_np, (long long)np);
  if (np == 0) return VPM_OK;
  h->launches = 0;
  const int G = (int)h->devs.size();
  Dev &d0 = h->devs[0];
  CK(h, cudaSetDevice(d0.id));
  CK(h, cudaEventRecord(d0.ev[0], d0.stream));
  const bool reset = flags & VPM_FLAG_RESET;
  bool has_static = false;
  TRY(h1_upload(h, d0, P, nf, np, !reset, false, has_static));
  const bool prior = !reset || has_static;
  if (!prior) CK(h, cudaMemsetAsync(d0.res18.p, 0, (size_t)np * RES_ROWS * sizeof(double), d0.stream));
  const TreeView tv = tree_view(h);
  DevCsr c;
  TRY(build_csr_device(h, fn, tv.lbegin, tv.lend, h->tree_nl, np, tv.lbegin, tv.lend, h->tree_nl, np, tv.pt, tv.ps,
                       h->tree_npairs, G, nullptr, 0, nullptr, 0, c, true));
  CK(h, cudaEventRecord(d0.ev[1], d0.stream));
  const int64_t ns_pad = round_up(np, kTile);
  std::vector<LeafCsr> csr(G);
  csr[0] = c.csr;
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    CK(h, cudaSetDevice(d.id));
    TRY(ensure(h, d.tbuf, (size_t)np * 16 * sizeof(double)));
    TRY(ensure(h, d.sbuf, (size_t)np * 8 * sizeof(double)));
    TRY(ensure(h, d.rec, (size_t)ns_pad * kRec * sizeof(double)));
    if (g > 0) {
      TRY(ensure(h, d.in7, (size_t)np * 7 * sizeof(double)));
      TRY(ensure(h, d.tree, h->devs[0].tree.cap));
      TRY(ensure(h, d.ibuf, h->devs[0].ibuf.cap));
      csr[g] = rebase_csr(c.csr, h->devs[0].ibuf.p, d.ibuf.p);
    }
  }
  // replicate state, sort index and list tables over NVLink; every device gathers its own
  // tree-sorted buffers
  TRY(bcast_from_dev0(h, &Dev::in7, (size_t)np * 7 * sizeof(double)));
  TRY(bcast_from_dev0(h, &Dev::tree, (size_t)np * 8));
  TRY(bcast_from_dev0(h, &Dev::ibuf, c.bcast_bytes));
  std::vector<std::pair<int64_t, int64_t>> cols(G, {0, 0});
  // leaf tables on the host are not available (device-built): the column range of a device is
  // [begin of its first item, end of its last item), read from the cut records
  std::vector<int64_t> hb((size_t)h->tree_nl), he((size_t)h->tree_nl);
  if (G > 1) {
    CK(h, cudaSetDevice(d0.id));
    CK(h, cudaMemcpyAsync(hb.data(), tv.lbegin, hb.size() * 8, cudaMemcpyDeviceToHost, d0.stream));
    CK(h, cudaMemcpyAsync(he.data(), tv.lend, he.size() * 8, cudaMemcpyDeviceToHost, d0.stream));
    CK(h, cudaStreamSynchronize(d0.stream));
  }
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    cudaStream_t st = d.stream;
    const int64_t k0 = c.cut[g], k1 = c.cut[g + 1];
    if (k1 <= k0) continue;
    CK(h, cudaSetDevice(d.id));
    const int64_t lf = c.first_leaf[g], ll = c.last_leaf[g + 1];
    const int64_t col0 = G == 1 ? 0 : hb[(size_t)lf] + c.first_off[g];
    const int64_t col1 = G == 1 ? np : std::min<int64_t>(he[(size_t)ll], hb[(size_t)ll] + c.last_off[g + 1] + c.nt);
    tree_gather_kernel<<<blocks_for(np, 256), 256, 0, st>>>((const double *)d.in7.p, 7, 0, 3, 6, (const int64_t *)d.tree.p,
                                                            np, (double *)d.sbuf.p, (double *)d.tbuf.p);
    SrcView sv{(const double *)d.sbuf.p, 8, 0, 4, 7};
    prep_uj_records<<<blocks_for(ns_pad, 256), 256, 0, st>>>(sv, 0, np, ns_pad, kernel, (double *)d.rec.p);
    LeafUjArgs a;
    a.csr = csr[g];
    a.csr.wi_leaf += k0;
    a.csr.wi_off += k0;
    a.tpos = (const double *)d.tbuf.p; a.tld = 16; a.rec = (const double *)d.rec.p;
    a.out = (double *)d.tbuf.p; a.urow = 4; a.jrow = 7; a.want_U = 1; a.want_J = 1;
    a.shortcut = (flags & VPM_FLAG_NO_FARFIELD_SHORTCUT) ? 0 : 1;
    if (g == 0) CK(h, cudaEventRecord(d.ev[6], st));
    launch_uj_leaf(kernel, c.nt, (unsigned)(k1 - k0), a, st);
    if (g == 0) CK(h, cudaEventRecord(d.ev[7], st));
    h->launches += 3;
    CK(h, cudaGetLastError());
    cols[g] = {col0, col1};
  }
  // the devices return their columns of the sorted result to device 0 over NVLink
  // (ncclSend / ncclRecv; the receives are ordered on device 0's stream after its own
  // gather + pair kernel and before the scatter)
  if (G > 1) {
    NCK(h, g_nccl.group_start());
    for (int g = 1; g < G; ++g) {
      const int64_t col0 = cols[g].first, col1 = cols[g].second;
      if (col1 <= col0) continue;
      Dev &d = h->devs[g];
      const size_t cnt = (size_t)(col1 - col0) * 16;
      NCK(h, g_nccl.send((double *)d.tbuf.p + col0 * 16, cnt, kNcclFloat64, 0, h->comms[g], d.stream));
      NCK(h, g_nccl.recv((double *)d0.tbuf.p + col0 * 16, cnt, kNcclFloat64, g, h->comms[0], d0.stream));
    }
    NCK(h, g_nccl.group_end());
  }
  CK(h, cudaSetDevice(d0.id));
  CK(h, cudaEventRecord(d0.ev[2], d0.stream));
  tree_scatter_kernel<<<blocks_for(np, 256), 256, 0, d0.stream>>>((const double *)d0.tbuf.p, (const int64_t *)d0.tree.p, np,
                                                                 (double *)d0.res18.p, RES_ROWS, RES_U, RES_J, RES_W,
                                                                 RES_PSE, reset ? 1 : 0,
                                                                 has_static ? (const double *)d0.stat.p : nullptr, 1);
  h->launches++;
  CK(h, cudaGetLastError());
  CK(h, cudaEventRecord(d0.ev[3], d0.stream));
  CK(h, cudaEventRecord(d0.ev[4], d0.stream));
  TRY(h1_download(h, d0, P, nf, np, 0));
  for (int g = G - 1; g >= 0; --g) {
    CK(h, cudaSetDevice(h->devs[g].id));
    CK(h, cudaStreamSynchronize(h->devs[g].stream));
  }
  h->timing.uj_pairs = c.pairs;
  h->timing.sfs_pairs = 0;
  h1_fill_timing(h, d0);
  h->timing.uj_ms = ev_ms(d0.ev[6], d0.ev[7]);  // the pair kernel of device 0 alone
  h->np_resident = -1;
  return VPM_OK;
}

int vpm_field_upload(vpm_handle *h, const double *P, int64_t nf, int64_t np) {
  TRY(check_field(h, "vpm_field_upload", P, nf, np, 0));
  const int G = (int)h->devs.size();
  const int64_t shard = (np + G - 1) / G, np_pad = std::max<int64_t>(shard * G, 1);
  for (Dev &d : h->devs) {
    CK(h, cudaSetDevice(d.id));
    TRY(ensure(h, d.fld, (size_t)np_pad * nf * sizeof(double)));
  }
  Dev &d0 = h->devs[0];
  CK(h, cudaSetDevice(d0.id));
  if (np_pad > np)
    CK(h, cudaMemsetAsync((double *)d0.fld.p + np * nf, 0, (size_t)(np_pad - np) * nf * sizeof(double), d0.stream));
  if (np > 0) CK(h, cudaMemcpyAsync(d0.fld.p, P, (size_t)np * nf * sizeof(double), cudaMemcpyHostToDevice, d0.stream));
  h->fld_nf = nf;
  h->fld_np = np;
  h->fld_t_sgm = 0.0;
  TRY(bcast_from_dev0(h, &Dev::fld, (size_t)np_pad * nf * sizeof(double)));
  for (int g = G - 1; g >= 0; --g) {
    CK(h, cudaSetDevice(h->devs[g].id));
    CK(h, cudaStreamSynchronize(h->devs[g].stream));
  }
  return VPM_OK;
}

int vpm_field_download(vpm_handle *h, double *P, int64_t nf, int64_t np) {
  TRY(check_field(h, "vpm_field_download", P, nf, np, 0));
  if (h->fld_np != np || h->fld_nf != nf)
    return fail(h, VPM_ESTATE, "vpm_field_download: a %lld x %lld field is resident, not %lld x %lld",
                (long long)h->fld_nf, (long long)h->fld_np, (long long)nf, (long long)np);
  Dev &d = h->devs[0];
  CK(h, cudaSetDevice(d.id));
  if (np > 0) CK(h, cudaMemcpyAsync(P, d.fld.p, (size_t)np * nf * sizeof(double), cudaMemcpyDeviceToHost, d.stream));
  CK(h, cudaStreamSynchronize(d.stream));
  return VPM_OK;
}

static int field_sync_all(vpm_handle *h) {
  for (int g = (int)h->devs.size() - 1; g >= 0; --g) {
    CK(h, cudaSetDevice(h->devs[g].id));
    CK(h, cudaStreamSynchronize(h->devs[g].stream));
  }
  return VPM_OK;
}

int vpm_field_uj(vpm_handle *h, int kernel, int flags) {
  if (!h) return VPM_EINVAL;
  if (h->fld_np < 0) return fail(h, VPM_ESTATE, "vpm_field_uj: no resident field (call vpm_field_upload first)");
  if (!valid_kernel(kernel)) return fail(h, VPM_EINVAL, "vpm_field_uj: unknown kernel_id %d", kernel);
  Dev &d = h->devs[0];
  h->launches = 0;
  CK(h, cudaSetDevice(d.id));
  CK(h, cudaEventRecord(d.ev[0], d.stream));
  CK(h, cudaEventRecord(d.ev[1], d.stream));
  TRY(field_uj(h, kernel, flags));
  CK(h, cudaSetDevice(d.id));
  for (int k = 2; k <= 5; ++k) CK(h, cudaEventRecord(d.ev[k], d.stream));
  TRY(field_sync_all(h));
  h1_fill_timing(h, d);
  h->timing.uj_ms = h->timing.total_ms;
  return VPM_OK;
}

int vpm_field_step(vpm_handle *h, const vpm_step_params *sp) {
  if (!h || !sp) return VPM_EINVAL;
  if (h->fld_np < 0) return fail(h, VPM_ESTATE, "vpm_field_step: no resident field (call vpm_field_upload first)");
  if (!valid_kernel(sp->kernel_id)) return fail(h, VPM_EINVAL, "vpm_field_step: unknown kernel_id %d", sp->kernel_id);
  if (sp->integration < 0 || sp->integration > 1 || sp->relaxation < 0 || sp->relaxation > 2)
    return fail(h, VPM_EINVAL, "vpm_field_step: integration must be 0 (euler) or 1 (rungekutta3), relaxation 0..2");
  if (h->fld_nf < 44) return fail(h, VPM_EINVAL, "vpm_field_step: the resident field needs >= 44 rows");
  if (sp->sfs < 0 || sp->sfs > 2) return fail(h, VPM_EINVAL, "vpm_field_step: sfs must be 0 (none), 1 (constant) or 2 (dynamic)");
  if (sp->sfs == 2 && (sp->minC < 0 || sp->maxC < 0 || sp->minC > sp->maxC || sp->alpha <= 0))
    return fail(h, VPM_EINVAL, "vpm_field_step: invalid DynamicSFS parameters (minC=%g maxC=%g alpha=%g)", sp->minC, sp->maxC, sp->alpha);
  if (sp->viscous < 0 || sp->viscous > 1) return fail(h, VPM_EINVAL, "vpm_field_step: viscous must be 0 (Inviscid) or 1 (CoreSpreading)");
  if (sp->viscous == 1 && sp->kernel_id != K_GERF)
    return fail(h, VPM_EINVAL, "vpm_field_step: kernel %d is not compatible with viscous scheme CoreSpreading; compatible kernels are gaussianerf", sp->kernel_id);  // src/FLOWVPM_utils.jl:58-64
  if (sp->viscous == 1 && (sp->sgm0 <= 0 || sp->nu < 0 || sp->cs_itmax < 0))
    return fail(h, VPM_EINVAL, "vpm_field_step: invalid CoreSpreading parameters (nu=%g sgm0=%g itmax=%d)", sp->nu, sp->sgm0, sp->cs_itmax);
  h->launches = 0;
  const int64_t np = h->fld_np;
  if (np == 0) return VPM_OK;
  const int G = (int)h->devs.size();
  const int tr = sp->transposed ? VPM_FLAG_TRANSPOSED : 0;
  const int uj_flags = VPM_FLAG_RESET | tr | (sp->sfs ? (VPM_FLAG_SFS | VPM_FLAG_RESET_SFS) : 0);
  const unsigned nb = blocks_for(np, 256);
  std::vector<StepArgs> args(G);
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    CK(h, cudaSetDevice(d.id));
    StepArgs &a = args[g];
    a.P = (double *)d.fld.p; a.nf = h->fld_nf; a.np = np;
    a.a = 1.0; a.b = 1.0; a.dt = sp->dt; a.Ux = sp->Uinf[0]; a.Uy = sp->Uinf[1]; a.Uz = sp->Uinf[2];
    a.f = sp->f; a.g = sp->g; a.zeta0 = zeta0_of(sp->kernel_id); a.Cs = sp->Cs; a.rlxf = sp->rlxf;
    a.transposed = sp->transposed; a.sfs = sp->sfs; a.clip = sp->clip_backscatter; a.relax_kind = sp->relaxation;
    a.alpha = sp->alpha; a.sfs_rlxf = sp->sfs_rlxf; a.minC = sp->minC; a.maxC = sp->maxC;
    a.force_positive = sp->force_positive;
    a.controls = sp->controls; a.deltat = sp->deltat;
    TRY(ensure(h, d.ibuf, 4096));
    a.nan_flag = (int *)d.ibuf.p;
    CK(h, cudaMemsetAsync(a.nan_flag, 0, sizeof(int), d.stream));
  }
  // an O(N) kernel on every device's mirror (all mirrors hold the same data)
  auto on_all = [&](auto launch) -> int {
    for (int g = 0; g < G; ++g) {
      CK(h, cudaSetDevice(h->devs[g].id));
      launch(args[g], h->devs[g].stream);
      h->launches++;
    }
    CK(h, cudaGetLastError());
    return VPM_OK;
  };
  // the SFS hooks around a UJ evaluation at an Euler step / the first RK substep
  // (src/FLOWVPM_subfilterscale.jl:110-135 ConstantSFS, :204-268 DynamicSFS)
  auto sfs_before = [&]() -> int {
    if (sp->sfs != 2) return VPM_OK;
    TRY(on_all([&](StepArgs &a, cudaStream_t st) { step_scale_sigma<<<nb, 256, 0, st>>>(a, 0); }));
    TRY(field_uj(h, sp->kernel_id, VPM_FLAG_RESET | VPM_FLAG_RESET_SFS | VPM_FLAG_SFS | tr));
    TRY(on_all([&](StepArgs &a, cudaStream_t st) { step_dyn_store<<<nb, 256, 0, st>>>(a); }));
    TRY(on_all([&](StepArgs &a, cudaStream_t st) { step_scale_sigma<<<nb, 256, 0, st>>>(a, 1); }));
    return VPM_OK;
  };
  auto sfs_after = [&]() -> int {
    if (sp->sfs == 1) TRY(on_all([&](StepArgs &a, cudaStream_t st) { step_sfs_coeff<<<nb, 256, 0, st>>>(a); }));
    if (sp->sfs == 2) TRY(on_all([&](StepArgs &a, cudaStream_t st) { step_dyn_coeff<<<nb, 256, 0, st>>>(a); }));
    if (sp->sfs && (sp->controls & 3))
      TRY(on_all([&](StepArgs &a, cudaStream_t st) { step_sfs_controls<<<nb, 256, 0, st>>>(a); }));
    return VPM_OK;
  };
  Dev &d0 = h->devs[0];
  CK(h, cudaSetDevice(d0.id));
  CK(h, cudaEventRecord(d0.ev[0], d0.stream));
  CK(h, cudaEventRecord(d0.ev[1], d0.stream));
  if (sp->integration == 0) {  // euler: src/FLOWVPM_timeintegration.jl:23-37
    TRY(sfs_before());
    TRY(field_uj(h, sp->kernel_id, uj_flags));
    TRY(sfs_after());
    const int relax = sp->relax ? 1 : 0;
    TRY(on_all([&](StepArgs &a, cudaStream_t st) { step_euler<<<nb, 256, 0, st>>>(a, relax); }));
    if (sp->viscous) TRY(field_corespreading(h, sp, 0.0, 0.0));
  } else {  // rungekutta3: src/FLOWVPM_timeintegration.jl:388-461
    TRY(on_all([&](StepArgs &a, cudaStream_t st) { step_reset_M<<<nb, 256, 0, st>>>(a); }));
    const double ab[3][2] = {{0.0, 1.0 / 3}, {-5.0 / 9, 15.0 / 16}, {-153.0 / 128, 8.0 / 15}};
    for (int k = 0; k < 3; ++k) {
      for (StepArgs &a : args) { a.a = ab[k][0]; a.b = ab[k][1]; }
      if (k == 0) TRY(sfs_before());
      TRY(field_uj(h, sp->kernel_id, uj_flags));
      if (k == 0) TRY(sfs_after());
      TRY(on_all([&](StepArgs &a, cudaStream_t st) { step_rk_stage<<<nb, 256, 0, st>>>(a); }));
      if (sp->viscous) TRY(field_corespreading(h, sp, ab[k][0], ab[k][1]));
    }
    if (sp->relax && sp->relaxation) {
      TRY(field_uj(h, sp->kernel_id, VPM_FLAG_RESET | tr));
      TRY(on_all([&](StepArgs &a, cudaStream_t st) { step_relax<<<nb, 256, 0, st>>>(a); }));
    }
  }
  CK(h, cudaSetDevice(d0.id));
  for (int k = 2; k <= 5; ++k) CK(h, cudaEventRecord(d0.ev[k], d0.stream));
  int nan_flag = 0;
  CK(h, cudaMemcpyAsync(&nan_flag, args[0].nan_flag, sizeof(int), cudaMemcpyDeviceToHost, d0.stream));
  TRY(field_sync_all(h));
  h1_fill_timing(h, d0);
  h->timing.uj_ms = h->timing.total_ms;
  if (nan_flag) return fail(h, VPM_ESTATE, "NaN in dynamicprocedure_pseudo3level_afterUJ");  // subfilterscale.jl:645-652
  return VPM_OK;
}

int vpm_field_rbf(vpm_handle *h, int kernel, int itmax, double tol, int iterror, int *iterations, double *residuals) {
  if (!h) return VPM_EINVAL;
  if (h->fld_np < 0) return fail(h, VPM_ESTATE, "vpm_field_rbf: no resident field (call vpm_field_upload first)");
  if (!valid_kernel(kernel)) return fail(h, VPM_EINVAL, "vpm_field_rbf: unknown kernel_id %d", kernel);
  if (itmax < 0 || !(tol >= 0)) return fail(h, VPM_EINVAL, "vpm_field_rbf: itmax >= 0 and tol >= 0 required");
  h->launches = 0;
  if (iterations) *iterations = 0;
  if (h->fld_np == 0) return VPM_OK;
  int rc = field_rbf(h, kernel, itmax, tol, iterror, iterations, residuals);
  int rs = field_sync_all(h);
  h->timing.kernel_launches = h->launches;
  return rc != VPM_OK ? rc : rs;
}

int vpm_field_tsgm(vpm_handle *h, double *t_sgm, int set) {
  if (!h || !t_sgm) return VPM_EINVAL;
  if (set) h->fld_t_sgm = *t_sgm; else *t_sgm = h->fld_t_sgm;
  return VPM_OK;
}

int vpm_get_timing(const vpm_handle *h, vpm_timing *out) {
  if (!h || !out) return VPM_EINVAL;
  *out = h->timing;
  if (h->device_timing) {
    // stream-ordered entry points: the pair kernel's own duration, valid once the
    // caller has synchronised the stream it passed
    const Dev &d = h->devs[0];
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, d.ev[6], d.ev[7]) != cudaSuccess) { cudaGetLastError(); ms = 0.f; }
    if (h->device_timing == 1) out->uj_ms = ms; else out->sfs_ms = ms;
    out->n_gpus = (int32_t)h->devs.size();
  }
  return VPM_OK;
}

int vpm_measure_dfma_peak(vpm_handle *h, double *dfma_per_s, double *elapsed_ms) {
  if (!h || !dfma_per_s) return VPM_EINVAL;
  Dev &d = h->devs[0];
  cudaStream_t st = d.stream;
  CK(h, cudaSetDevice(d.id));
  TRY(ensure(h, d.ibuf, 4096));
  const int threads = 256, blocks = d.sm_count * 8, iters = 4096;
  dfma_peak_kernel<<<blocks, threads, 0, st>>>((double *)d.ibuf.p, 64, 1.0);  // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(h, cudaEventRecord(d.ev[6], st));
    dfma_peak_kernel<<<blocks, threads, 0, st>>>((double *)d.ibuf.p, iters, 1.0);
    CK(h, cudaEventRecord(d.ev[7], st));
    CK(h, cudaStreamSynchronize(st));
    CK(h, cudaGetLastError());
    best = std::min(best, ev_ms(d.ev[6], d.ev[7]));
  }
  const double n = (double)blocks * threads * (double)iters * 16.0 * 8.0;
  *dfma_per_s = n / (best * 1e-3);
  if (elapsed_ms) *elapsed_ms = best;
  return VPM_OK;
}

int vpm_measure_ffma_peak(vpm_handle *h, int mode, double *fma_per_s, double *elapsed_ms) {
  if (!h || !fma_per_s || mode < 0 || mode > 3) return VPM_EINVAL;
  Dev &d = h->devs[0];
  cudaStream_t st = d.stream;
  CK(h, cudaSetDevice(d.id));
  TRY(ensure(h, d.ibuf, 4096));
  const int threads = 256, blocks = d.sm_count * 8, iters = 4096;
  auto run = [&](int it) {
    float *o = (float *)d.ibuf.p;
    switch (mode) {
      case 0: ffma_peak_kernel<0><<<blocks, threads, 0, st>>>(o, it, 1.0f); break;
      case 1: ffma_peak_kernel<1><<<blocks, threads, 0, st>>>(o, it, 1.0f); break;
      case 2: ffma_peak_kernel<2><<<blocks, threads, 0, st>>>(o, it, 1.0f); break;
      default: ffma_peak_kernel<3><<<blocks, threads, 0, st>>>(o, it, 1.0f); break;
    }
  };
  run(64);  // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(h, cudaEventRecord(d.ev[6], st));
    run(iters);
    CK(h, cudaEventRecord(d.ev[7], st));
    CK(h, cudaStreamSynchronize(st));
    CK(h, cudaGetLastError());
    best = std::min(best, ev_ms(d.ev[6], d.ev[7]));
  }
  // scalar FMAs per second: 8 chains x 16 x 2 lanes per thread and iteration in every mode
  const double n = (double)blocks * threads * (double)iters * 16.0 * 8.0 * 2.0;
  *fma_per_s = n / (best * 1e-3);
  if (elapsed_ms) *elapsed_ms = best;
  return VPM_OK;
}

int vpm_test_math(vpm_handle *h, int op, int arg, const double *in, double *out, double *out2, int64_t n) {
  if (!h || !in || !out || n < 0) return fail(h, VPM_EINVAL, "vpm_test_math: bad argument");
  if (n == 0) return VPM_OK;
  Dev &d = h->devs[0];
  cudaStream_t st = d.stream;
  CK(h, cudaSetDevice(d.id));
  TRY(ensure(h, d.tbuf, (size_t)n * sizeof(double)));
  TRY(ensure(h, d.sbuf, (size_t)n * 2 * sizeof(double)));
  CK(h, cudaMemcpyAsync(d.tbuf.p, in, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
  double *o1 = (double *)d.sbuf.p, *o2 = o1 + n;
  test_math_kernel<<<blocks_for(n, 256), 256, 0, st>>>(op, arg, (const double *)d.tbuf.p, o1, o2, n);
  CK(h, cudaGetLastError());
  CK(h, cudaMemcpyAsync(out, o1, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (out2) CK(h, cudaMemcpyAsync(out2, o2, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(h, cudaStreamSynchronize(st));
  return VPM_OK;
}

}  // extern "C"
