// vpm_host_hook1.cuh -- host <-> device row copies and the pieces of Hook 1 (UJ slot) on one device.
// Part of the single translation unit vpm_abi.cu (included there in order; not a standalone header).
#pragma once
namespace {

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
  h->alloc_epoch++;
  CK(h, cudaMallocHost((void **)&h->h_stage, want * sizeof(double)));
  h->h_stage_cap = want;
  return VPM_OK;
}

// ---- host threads ----------------------------------------------------------------------
// The O(N) host loops over the particle matrix (strided gathers / scatters, the static-flag scan) and the copies
// into / out of the pinned ring are memory-latency bound on one core (2^24 particles: 0.1 s each).  They are split
// over a few threads.  Starting threads per loop costs ~120 us for seven (measured: the upload half of a
// 33 800-particle call went from 65 to 310 us when its gathers were threaded that way), so the workers are a
// persistent pool, one per process (handles on different host threads take turns), started on first use.
class HostPool {
 public:
  static HostPool &get() {
    static HostPool pool;
    return pool;
  }
  int workers() const { return (int)th_.size(); }
  // runs job(0) .. job(njobs - 1), the caller taking part; returns when all are done
  void run(int njobs, const std::function<void(int)> &job) {
    if (njobs <= 1 || th_.empty()) {
      for (int j = 0; j < njobs; ++j) job(j);
      return;
    }
    std::lock_guard<std::mutex> turn(turn_);
    {
      std::lock_guard<std::mutex> lk(m_);
      job_ = &job; njobs_ = njobs; next_ = 0; left_ = njobs; ++gen_;
    }
    go_.notify_all();
    work();
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [&] { return left_ == 0; });
    job_ = nullptr;
  }

 private:
  HostPool() : pid_(getpid()) {
    unsigned hw = std::thread::hardware_concurrency();
    const int n = (int)std::min<unsigned>(hw > 1 ? hw - 1 : 0, 7);
    for (int t = 0; t < n; ++t) th_.emplace_back([this] { loop(); });
  }
  ~HostPool() {
    if (getpid() != pid_) {  // a forked child has no workers to join (run() there does every job on the caller)
      for (auto &t : th_) t.detach();
      return;
    }
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
    }
    go_.notify_all();
    for (auto &t : th_) t.join();
  }
  void work() {
    for (;;) {
      int j;
      const std::function<void(int)> *job;
      {
        std::lock_guard<std::mutex> lk(m_);
        if (job_ == nullptr || next_ >= njobs_) return;
        j = next_++;
        job = job_;
      }
      (*job)(j);
      std::lock_guard<std::mutex> lk(m_);
      if (--left_ == 0) done_.notify_all();
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m_);
        go_.wait(lk, [&] { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_;
      }
      work();
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_, turn_;
  std::condition_variable go_, done_;
  const std::function<void(int)> *job_ = nullptr;
  int njobs_ = 0, next_ = 0, left_ = 0;
  uint64_t gen_ = 0;
  bool stop_ = false;
  pid_t pid_;
};

// fn(begin, end) over [0, n) in at most 8 contiguous pieces of >= min_chunk items (multiples of `align`)
template <class I, class F>
void parallel_ranges(I n, I min_chunk, I align, F fn) {
  HostPool &pool = HostPool::get();
  const int nt = (int)std::min<I>((I)(pool.workers() + 1), n / min_chunk);
  if (nt <= 1) { fn((I)0, n); return; }
  const I chunk = ((n + nt - 1) / nt + align - 1) / align * align;
  pool.run(nt, [&](int t) {
    const I a = (I)t * chunk;
    if (a < n) fn(a, std::min<I>(n, a + chunk));
  });
}
// particle columns of the host matrix: pieces of >= 64 Ki particles.  (Waking the pool for less does not pay: with
// 8 Ki-particle pieces the 1.9 MB gather of a 33 800-particle call took 335 us instead of 65 us on one thread.)
template <class F>
void parallel_chunks(int64_t n, F fn) { parallel_ranges<int64_t>(n, (int64_t)1 << 16, 1, fn); }
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

// Strided rows of `n` columns of a host matrix <-> device block, for every entry point other than Hook 1
// (which stages whole fields through h_stage, above).
//  * page-locked matrix (vpm_pin_host, or the caller's own pinned allocation): a 2-D DMA.  (Measured
//    alternative: mapping the matrix and gathering the rows with a kernel reading host memory directly
//    gave the same 13.8 GB/s for the 56-byte rows at N = 2^22, so the plain copy stays.)
//  * pageable matrix: the driver stages such a 2-D copy row by row (measured 0.7 GB/s: zeta_fmm on 2^20
//    particles spent 0.5 s uploading and 11 ms computing).  The rows go through a ring of two pinned
//    slots per device instead: host threads gather a slot while the DMA of the other one runs.
constexpr size_t kRingSlot = 16u << 20;
constexpr size_t kRingMinBytes = 256u << 10;  // smaller copies: the 2-D copy costs no more than the ring's bookkeeping

Dev *dev_of_stream(vpm_handle *h, cudaStream_t st) {
  for (Dev &d : h->devs)
    if (d.stream == st) return &d;
  return nullptr;
}
int ring_ready(vpm_handle *h, Dev &d) {
  if (d.ring[0]) return VPM_OK;
  CK(h, cudaSetDevice(d.id));
  for (int b = 0; b < 2; ++b) {
    CK(h, cudaEventCreateWithFlags(&d.ring_ev[b], cudaEventDisableTiming));
    CK(h, cudaMallocHost((void **)&d.ring[b], kRingSlot));
  }
  return VPM_OK;
}
bool use_ring(vpm_handle *h, Dev *d, const void *host, size_t bytes) {
  return d && !h->capturing && bytes >= kRingMinBytes && !host_is_pinned(host);
}
template <class F>
void parallel_cols(int64_t n, F fn) { parallel_ranges<int64_t>(n, (int64_t)1 << 14, 1, fn); }  // columns of a ring slot

// Device columns of `dpitch` bytes <- pieces of the host columns (pitch spitch): piece k is `width[k]` bytes
// from byte offset soff[k] of the host column to byte offset doff[k] of the device column.  The ring path
// needs the pieces to tile the device column (then a slot is a run of whole device columns and ONE
// contiguous copy moves it -- measured: a 2-D copy of 3e5 56-byte rows out of a pinned slot ran at
// 1.5 GB/s, the 1-D copy of the same bytes at PCIe rate: 106 -> 10 ms for the 19 rows of 2^20 particles).
struct RowPiece { size_t soff, doff, width; };
int h2d_pieces(vpm_handle *h, cudaStream_t st, void *dst, size_t dpitch, const void *src, size_t spitch,
               const RowPiece *pc, int npc, int64_t n) {
  if (n <= 0) return VPM_OK;
  Dev *d = dev_of_stream(h, st);
  size_t covered = 0;
  bool tiles = true;
  for (int k = 0; k < npc; ++k) {
    tiles = tiles && pc[k].doff == covered;
    covered += pc[k].width;
  }
  tiles = tiles && covered == dpitch;
  if (!tiles || !use_ring(h, d, src, dpitch * (size_t)n)) {
    for (int k = 0; k < npc; ++k)
      CK(h, cudaMemcpy2DAsync((char *)dst + pc[k].doff, dpitch, (const char *)src + pc[k].soff, spitch, pc[k].width,
                              (size_t)n, cudaMemcpyHostToDevice, st));
    return VPM_OK;
  }
  TRY(ring_ready(h, *d));
  const int64_t cols = (int64_t)(kRingSlot / dpitch);
  RowPiece p[4];
  if (npc > 4) return fail(h, VPM_EINVAL, "h2d_pieces: more than 4 pieces");
  for (int k = 0; k < npc; ++k) p[k] = pc[k];
  for (int64_t c0 = 0; c0 < n; c0 += cols) {
    const int64_t nc = std::min(cols, n - c0);
    const int b = d->ring_next;
    d->ring_next ^= 1;
    if (d->ring_busy[b]) CK(h, cudaEventSynchronize(d->ring_ev[b]));  // the DMA that last read this slot
    char *slot = d->ring[b];
    const char *s0 = (const char *)src + (size_t)c0 * spitch;
    const RowPiece p0 = p[0], p1 = p[1], p2 = p[2], p3 = p[3];
    parallel_cols(nc, [=](int64_t a, int64_t e) {
      const RowPiece q[4] = {p0, p1, p2, p3};
      for (int64_t i = a; i < e; ++i)
        for (int k = 0; k < npc; ++k)
          memcpy(slot + (size_t)i * dpitch + q[k].doff, s0 + (size_t)i * spitch + q[k].soff, q[k].width);
    });
    CK(h, cudaMemcpyAsync((char *)dst + (size_t)c0 * dpitch, slot, (size_t)nc * dpitch, cudaMemcpyHostToDevice, st));
    CK(h, cudaEventRecord(d->ring_ev[b], st));
    d->ring_busy[b] = true;
  }
  return VPM_OK;
}
int h2d_strided(vpm_handle *h, cudaStream_t st, void *dst, size_t dpitch, const void *src, size_t spitch, size_t width,
                int64_t n) {
  const RowPiece p{0, 0, width};
  return h2d_pieces(h, st, dst, dpitch, src, spitch, &p, 1, n);
}

// host columns (pitch dpitch) <- `width` bytes of each of n device columns (pitch spitch).  For a pageable
// destination the call returns with the data in place (the scatter is host work); otherwise it is
// asynchronous on `st` like the 2-D copy it wraps.  The ring path moves WHOLE device columns (one
// contiguous copy per slot, see h2d_pieces) and scatters the `width` bytes it was asked for: meant for
// spitch not much larger than width (the callers: equal, or 144 against 120 bytes).
int d2h_strided(vpm_handle *h, cudaStream_t st, void *dst, size_t dpitch, const void *src, size_t spitch, size_t width,
                int64_t n) {
  if (n <= 0) return VPM_OK;
  Dev *d = dev_of_stream(h, st);
  if (spitch > 2 * width || !use_ring(h, d, dst, width * (size_t)n)) {
    CK(h, cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, (size_t)n, cudaMemcpyDeviceToHost, st));
    return VPM_OK;
  }
  TRY(ring_ready(h, *d));
  const int64_t cols = (int64_t)(kRingSlot / spitch);
  int pb = -1;
  int64_t pc0 = 0, pnc = 0;
  auto scatter_prev = [&]() -> int {
    if (pb < 0) return VPM_OK;
    CK(h, cudaEventSynchronize(d->ring_ev[pb]));
    d->ring_busy[pb] = false;
    const char *slot = d->ring[pb];
    char *d0 = (char *)dst + (size_t)pc0 * dpitch;
    parallel_cols(pnc, [=](int64_t a, int64_t e) {
      for (int64_t i = a; i < e; ++i) memcpy(d0 + (size_t)i * dpitch, slot + (size_t)i * spitch, width);
    });
    return VPM_OK;
  };
  for (int64_t c0 = 0; c0 < n; c0 += cols) {
    const int64_t nc = std::min(cols, n - c0);
    const int b = d->ring_next;
    d->ring_next ^= 1;
    if (d->ring_busy[b]) CK(h, cudaEventSynchronize(d->ring_ev[b]));
    // the last column contributes only its `width` bytes (the device block may end there)
    CK(h, cudaMemcpyAsync(d->ring[b], (const char *)src + (size_t)c0 * spitch, (size_t)(nc - 1) * spitch + width,
                          cudaMemcpyDeviceToHost, st));
    CK(h, cudaEventRecord(d->ring_ev[b], st));
    d->ring_busy[b] = true;
    TRY(scatter_prev());  // overlaps the copy just issued
    pb = b; pc0 = c0; pnc = nc;
  }
  return scatter_prev();
}

// Contiguous blocks (FastMultipole's buffers, the near-field list, the whole matrix of vpm_field_upload).  The
// driver stages a copy from / to pageable memory through its own bounce buffer on the calling thread; for blocks
// of megabytes the ring with eight copying threads is faster.  Device pointers and page-locked memory go straight
// to cudaMemcpyAsync (cudaMemcpyDefault: the list entry points also accept device-resident tables).
bool host_is_pageable(const void *ptr) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}
constexpr size_t kRingMinContig = 4u << 20;
template <class F>
void parallel_bytes(size_t n, F fn) { parallel_ranges<size_t>(n, (size_t)1 << 20, 64, fn); }
int h2d_contig(vpm_handle *h, cudaStream_t st, void *dst, const void *src, size_t bytes) {
  if (bytes == 0) return VPM_OK;
  Dev *d = dev_of_stream(h, st);
  if (!d || h->capturing || bytes < kRingMinContig || !host_is_pageable(src)) {
    CK(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, st));
    return VPM_OK;
  }
  TRY(ring_ready(h, *d));
  for (size_t off = 0; off < bytes; off += kRingSlot) {
    const size_t nb = std::min(kRingSlot, bytes - off);
    const int b = d->ring_next;
    d->ring_next ^= 1;
    if (d->ring_busy[b]) CK(h, cudaEventSynchronize(d->ring_ev[b]));
    char *slot = d->ring[b];
    const char *s0 = (const char *)src + off;
    parallel_bytes(nb, [=](size_t a, size_t e) { memcpy(slot + a, s0 + a, e - a); });
    CK(h, cudaMemcpyAsync((char *)dst + off, slot, nb, cudaMemcpyHostToDevice, st));
    CK(h, cudaEventRecord(d->ring_ev[b], st));
    d->ring_busy[b] = true;
  }
  return VPM_OK;
}
// (pageable destination: returns with the data in place, like d2h_strided)
int d2h_contig(vpm_handle *h, cudaStream_t st, void *dst, const void *src, size_t bytes) {
  if (bytes == 0) return VPM_OK;
  Dev *d = dev_of_stream(h, st);
  if (!d || h->capturing || bytes < kRingMinContig || !host_is_pageable(dst)) {
    CK(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, st));
    return VPM_OK;
  }
  TRY(ring_ready(h, *d));
  int pb = -1;
  size_t poff = 0, pnb = 0;
  auto drain_prev = [&]() -> int {
    if (pb < 0) return VPM_OK;
    CK(h, cudaEventSynchronize(d->ring_ev[pb]));
    d->ring_busy[pb] = false;
    const char *slot = d->ring[pb];
    char *d0 = (char *)dst + poff;
    parallel_bytes(pnb, [=](size_t a, size_t e) { memcpy(d0 + a, slot + a, e - a); });
    return VPM_OK;
  };
  for (size_t off = 0; off < bytes; off += kRingSlot) {
    const size_t nb = std::min(kRingSlot, bytes - off);
    const int b = d->ring_next;
    d->ring_next ^= 1;
    if (d->ring_busy[b]) CK(h, cudaEventSynchronize(d->ring_ev[b]));
    CK(h, cudaMemcpyAsync(d->ring[b], (const char *)src + off, nb, cudaMemcpyDeviceToHost, st));
    CK(h, cudaEventRecord(d->ring_ev[b], st));
    d->ring_busy[b] = true;
    TRY(drain_prev());  // overlaps the copy just issued
    pb = b; poff = off; pnb = nb;
  }
  return drain_prev();
}

// rows [row.., row+nrows) of `np` columns of a host matrix (leading dimension nf) <-> compact device block
int h2d_rows(vpm_handle *h, cudaStream_t st, double *dst, const double *src, int64_t nf, int nrows, int64_t np) {
  return h2d_strided(h, st, dst, nrows * sizeof(double), src, nf * sizeof(double), nrows * sizeof(double), np);
}
int d2h_rows(vpm_handle *h, cudaStream_t st, double *dst, int64_t nf, const double *src, int nrows, int64_t np) {
  return d2h_strided(h, st, dst, nf * sizeof(double), src, nrows * sizeof(double), nrows * sizeof(double), np);
}

// ---- Hook 1 pieces (single device d; targets = all particles) ---------------

// host -> device: X, Gamma, sigma rows; static flags (compacted on the host,
// only if any is set); previous U..PSE and SFS rows when they are accumulated on.
// Split in a host half (gathers into the pinned staging block / the static-flag block) and a device
// half (the asynchronous copies) so that a captured CUDA graph of the device half can be replayed
// with only the host half redone (small-field path of vpm_uj_direct).
struct H1Rows {
  bool pinned = false, has_static = false, need_prior = false, need_sfs_rows = false;
};

int h1_upload_host(vpm_handle *h, const double *P, int64_t nf, int64_t np, const H1Rows &r) {
  if (np == 0) return VPM_OK;
  if (!r.pinned) {
    TRY(ensure_stage(h, (size_t)np * (7 + RES_ROWS + 3)));
    double *stg = h->h_stage;
    gather_rows(stg, P, nf, R_X, 7, np);
    if (r.need_prior || r.has_static) gather_rows(stg + (size_t)np * 7, P, nf, R_U, RES_ROWS, np);
    if (r.need_sfs_rows) gather_rows(stg + (size_t)np * (7 + RES_ROWS), P, nf, R_SFS, 3, np);
  }
  if (r.has_static) {
    if (h->h_stat_cap < (size_t)np) {
      if (h->h_stat) cudaFreeHost(h->h_stat);
      h->h_stat = nullptr;
      h->h_stat_cap = 0;
      h->alloc_epoch++;
      CK(h, cudaMallocHost((void **)&h->h_stat, (size_t)np * sizeof(double)));
      h->h_stat_cap = (size_t)np;
    }
    for (int64_t i = 0; i < np; ++i) h->h_stat[i] = P[nf * i + R_STATIC];
  }
  return VPM_OK;
}

int h1_upload_dev(vpm_handle *h, Dev &d, const double *P, int64_t nf, int64_t np, const H1Rows &r) {
  cudaStream_t st = d.stream;
  CK(h, cudaSetDevice(d.id));
  const size_t n = (size_t)std::max<int64_t>(np, 1);
  TRY(ensure(h, d.in7, n * 7 * sizeof(double)));
  TRY(ensure(h, d.res18, n * RES_ROWS * sizeof(double)));
  TRY(ensure(h, d.sfs3, n * 3 * sizeof(double)));
  if (np == 0) return VPM_OK;
  const double *stg = h->h_stage;
  if (!r.pinned) CK(h, cudaMemcpyAsync(d.in7.p, stg, (size_t)np * 7 * sizeof(double), cudaMemcpyHostToDevice, st));
  else TRY(h2d_rows(h, st, (double *)d.in7.p, P, nf, 7, np));
  if (r.has_static) {
    TRY(ensure(h, d.stat, (size_t)np * sizeof(double)));
    CK(h, cudaMemcpyAsync(d.stat.p, h->h_stat, (size_t)np * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  if (r.need_prior || r.has_static) {
    if (!r.pinned)
      CK(h, cudaMemcpyAsync(d.res18.p, stg + (size_t)np * 7, (size_t)np * RES_ROWS * sizeof(double), cudaMemcpyHostToDevice, st));
    else
      TRY(h2d_rows(h, st, (double *)d.res18.p, P + R_U, nf, RES_ROWS, np));
  }
  if (r.need_sfs_rows) {
    if (!r.pinned)
      CK(h, cudaMemcpyAsync(d.sfs3.p, stg + (size_t)np * (7 + RES_ROWS), (size_t)np * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    else
      TRY(h2d_rows(h, st, (double *)d.sfs3.p, P + R_SFS, nf, 3, np));
  }
  return VPM_OK;
}

int h1_upload(vpm_handle *h, Dev &d, const double *P, int64_t nf, int64_t np, bool need_prior,
              bool need_sfs_rows, bool &has_static) {
  H1Rows r;
  r.need_prior = need_prior; r.need_sfs_rows = need_sfs_rows;
  has_static = false;
  if (np > 0) {
    r.has_static = has_static = any_static(P, nf, np);
    r.pinned = host_is_pinned(P);
  }
  TRY(h1_upload_host(h, P, nf, np, r));
  return h1_upload_dev(h, d, P, nf, np, r);
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
    launch_uj_finish(f, st);
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
    launch_sfs_finish(f, st);
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

// device -> host of rows 10:27 (+ 40:42): the asynchronous half (into the staging block for a pageable
// matrix) and, after the stream has been synchronised, the host half (scatter from the staging block)
int h1_download_dev(vpm_handle *h, Dev &d, double *P, int64_t nf, int64_t np, int flags, bool pinned) {
  cudaStream_t st = d.stream;
  CK(h, cudaSetDevice(d.id));
  const bool sfs_rows = flags & (VPM_FLAG_SFS | VPM_FLAG_RESET_SFS);
  if (np > 0 && !pinned) {
    TRY(ensure_stage(h, (size_t)np * (7 + RES_ROWS + 3)));
    double *s18 = h->h_stage + (size_t)np * 7, *s3 = h->h_stage + (size_t)np * (7 + RES_ROWS);
    CK(h, cudaMemcpyAsync(s18, d.res18.p, (size_t)np * RES_ROWS * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (sfs_rows) CK(h, cudaMemcpyAsync(s3, d.sfs3.p, (size_t)np * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
  } else if (np > 0) {
    CK(h, cudaMemcpy2DAsync(P + R_U, nf * sizeof(double), d.res18.p, RES_ROWS * sizeof(double),
                            RES_ROWS * sizeof(double), (size_t)np, cudaMemcpyDeviceToHost, st));
    if (sfs_rows)
      CK(h, cudaMemcpy2DAsync(P + R_SFS, nf * sizeof(double), d.sfs3.p, 3 * sizeof(double),
                              3 * sizeof(double), (size_t)np, cudaMemcpyDeviceToHost, st));
  }
  CK(h, cudaEventRecord(d.ev[5], st));
  return VPM_OK;
}
void h1_download_host(vpm_handle *h, double *P, int64_t nf, int64_t np, int flags, bool pinned) {
  if (np <= 0 || pinned) return;
  const bool sfs_rows = flags & (VPM_FLAG_SFS | VPM_FLAG_RESET_SFS);
  scatter_rows(P, nf, R_U, RES_ROWS, np, h->h_stage + (size_t)np * 7);
  if (sfs_rows) scatter_rows(P, nf, R_SFS, 3, np, h->h_stage + (size_t)np * (7 + RES_ROWS));
}
int h1_download(vpm_handle *h, Dev &d, double *P, int64_t nf, int64_t np, int flags) {
  const bool pinned = np > 0 && host_is_pinned(P);
  TRY(h1_download_dev(h, d, P, nf, np, flags, pinned));
  CK(h, cudaStreamSynchronize(d.stream));
  h1_download_host(h, P, nf, np, flags, pinned);
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

}  // namespace
