// vpm_host_base.cuh -- handle, per-device buffers, error plumbing of the C ABI host side.
// Part of the single translation unit vpm_abi.cu (included there in order; not a standalone header).
#pragma once
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
  // The sweep scratch (rec, srec, partial) is shared by every entry point.  The synchronous ones run on
  // `stream` and return with it idle; vpm_uj_device / vpm_sfs_device run on the CALLER's stream and
  // return without synchronising.  scratch_ev is recorded after their last kernel, and whoever touches
  // the scratch next -- on any stream -- waits on it first (scratch_acquire).
  cudaEvent_t scratch_ev = nullptr;
  bool scratch_pending = false;
  Buf in7, stat, res18, sfs3, rec, srec, partial, tbuf, sbuf, ibuf, jbuf, fld, scr, scr2, cubtmp, tree, tlist;
  Buf flg;  // the time step's NaN flag (its own block: ibuf and the scratch are rebuilt by list sweeps inside a step)
  // two pinned slots through which the strided rows of a PAGEABLE host matrix travel (h2d_strided /
  // d2h_strided in vpm_host_hook1.cuh); allocated on first use
  char *ring[2] = {nullptr, nullptr};
  cudaEvent_t ring_ev[2] = {};
  bool ring_busy[2] = {false, false};
  int ring_next = 0;
};

struct Plan {
  int tab = 0;  // != 0: uj_pairs_tab_kernel with `tab` threads x 2 targets per CTA
  int T = 1;
  int unroll = 2;
  int nsplit = 1;
  int tiles_per_split = 1;
  int64_t src_per_split = 0;  // sources per split of the 128-thread FP64 kernels (tiles_per_split * kTile, or finer)
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
  int opt_nearfield_fp32 = 0;  // VPM_OPT_NEARFIELD_FP32
  int opt_uj_variant = 0;      // VPM_OPT_UJ_VARIANT  (0 = automatic)
  int opt_sfs_variant = 0;     // VPM_OPT_SFS_VARIANT
  int opt_uj_const = 0;        // VPM_OPT_UJ_CONST
  int opt_uj_table = 0;        // VPM_OPT_UJ_TABLE
  int opt_graph = 1;           // VPM_OPT_SMALL_GRAPH (on by default)
  double last_near_fraction = -1.0;  // sampled by the last automatic gaussianerf kernel choice
  // small-field path of vpm_uj_direct: captured CUDA graphs of the device half of a call, keyed by
  // everything the captured nodes depend on; dropped whenever a buffer they point into moves
  struct GraphEntry {
    const double *P = nullptr;
    int64_t nf = 0, np = 0;
    int kernel = 0, flags = 0;
    bool has_static = false, pinned = false;
    int seen = 0;
    cudaGraphExec_t exec = nullptr;
    int launches = 0;
  };
  std::vector<GraphEntry> graphs;
  bool capturing = false;  // a stream capture of the small-field path is in progress
  uint64_t alloc_epoch = 0, graphs_epoch = 0;  // alloc_epoch: bumped by every (re)allocation
  int64_t fld_nf = 0, fld_np = -1;  // device mirror of the whole particle matrix (vpm_field_*)
  double fld_t_sgm = 0.0;           // CoreSpreading.t_sgm of the resident field
  int device_timing = 0;  // 1/2: ev[6..7] bracket the last _device U/J / SFS pair kernel
  // device-built leaf lists (vpm_leaflists_build), resident on device 0
  int64_t tree_np = -1, tree_nl = 0, tree_npairs = 0;
  unsigned long long tree_fingerprint = 0;  // of the X and sigma rows the lists were built from
  int64_t tree_ncrit = 0;                   // parameters of that build
  double tree_theta = 0.0;
  // CoreSpreading's `zeta` on the resident field (vpm_field_zeta_method): 0 zeta_direct, 1 zeta_fmm
  // (J[1:3] accumulated on, as the reference's does), 2 zeta_fmm with J[1:3] zeroed first
  int zeta_method = 0;
  int64_t zeta_ncrit = 50;
  double zeta_theta = 0.4;
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
  if (h) h->alloc_epoch++;
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

void drop_graphs(vpm_handle *h) {
  for (auto &g : h->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  h->graphs.clear();
}

// order stream `st` after the last stream-ordered (_device) use of the sweep scratch of `d`
int scratch_acquire(vpm_handle *h, Dev &d, cudaStream_t st) {
  if (h->capturing) return VPM_OK;  // ordered before the capture began / before every replay instead
  if (d.scratch_pending) CK(h, cudaStreamWaitEvent(st, d.scratch_ev, 0));
  return VPM_OK;
}
// the _device entry points: the scratch is in use on `st` until here
int scratch_release_async(vpm_handle *h, Dev &d, cudaStream_t st) {
  if (!d.scratch_ev) CK(h, cudaEventCreateWithFlags(&d.scratch_ev, cudaEventDisableTiming));
  CK(h, cudaEventRecord(d.scratch_ev, st));
  d.scratch_pending = true;
  return VPM_OK;
}

bool valid_kernel(int k) { return k >= 0 && k <= 3; }

}  // namespace
