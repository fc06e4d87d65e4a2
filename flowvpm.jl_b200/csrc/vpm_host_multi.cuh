// vpm_host_multi.cuh -- NCCL (loaded lazily), replication helpers and UJ_direct on the G devices of one process.
// Part of the single translation unit vpm_abi.cu (included there in order; not a standalone header).
#pragma once
namespace {

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

// Replicated upload of `nrows` strided rows of a PAGE-LOCKED host matrix: every device pulls its 1/G of the
// columns over its OWN PCIe link into its own copy of the buffer, then one in-place ncclAllGather over
// NVLink completes every copy -- instead of one device pulling everything (the single link was the floor of a
// C5 call: 0.94 GB of X, Gamma, sigma at 2^24 particles) and broadcasting it.  The buffer must hold
// G * ceil(np / G) columns.
int h2d_rows_sharded(vpm_handle *h, Buf Dev::*member, const double *P, int64_t nf, int nrows, int64_t np) {
  const int G = (int)h->devs.size();
  TRY(ensure_comms(h));
  const int64_t c = (np + G - 1) / G;
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    CK(h, cudaSetDevice(d.id));
    const int64_t c0 = std::min<int64_t>(np, g * c), c1 = std::min<int64_t>(np, (g + 1) * c);
    TRY(h2d_rows(h, d.stream, (double *)(d.*member).p + c0 * nrows, P + c0 * nf, nf, nrows, c1 - c0));
  }
  NCK(h, g_nccl.group_start());
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    double *buf = (double *)(d.*member).p;
    NCK(h, g_nccl.all_gather(buf + (size_t)g * c * nrows, buf, (size_t)c * nrows, kNcclFloat64, h->comms[g], d.stream));
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
      launch_uj_finish(f, st);
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
        launch_sfs_finish(f, st);
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

}  // namespace
