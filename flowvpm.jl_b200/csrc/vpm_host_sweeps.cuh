// vpm_host_sweeps.cuh -- launch plans, kernel launchers and the two O(N^2) sweeps.
// Part of the single translation unit vpm_abi.cu (included there in order; not a standalone header).
#pragma once
namespace {

// Launch plan.  One CTA = kThreads * T targets x one contiguous range of source tiles.
// The grid is sized to ~16 waves of resident CTAs so that the tail of the last wave is a
// few percent at most; when the target count alone cannot provide that, the sources are
// split (>= 4 tiles per split) and the finish kernel adds the splits in order.
enum PlanKind { PLAN_UJ = 0, PLAN_SFS = 1, PLAN_UJ_F32 = 2 };
// `variant`: 0 = automatic, else the validated value of VPM_OPT_UJ_VARIANT / VPM_OPT_SFS_VARIANT
// ("<T><unroll>"; the plan is built FROM it, so grid and kernel always agree).
constexpr int64_t kMinSrcPerSplit = 16, kMaxSmallSplits = 32;
Plan make_plan(int64_t nt, int64_t ns, int sm_count, PlanKind kind, int variant = 0) {
  Plan p;
  const int64_t ntiles = std::max<int64_t>(1, (ns + kTile - 1) / kTile);
  // two targets per thread (5 % faster in steady state) only when that still leaves enough
  // CTAs to fill the machine a few times; small fields take T = 1 and splits down to one tile
  const int64_t nblk2 = std::max<int64_t>(1, (nt + 2 * kThreads - 1) / (2 * kThreads));
  const bool big = nblk2 * std::max<int64_t>(1, ntiles / 4) >= (int64_t)sm_count * 4 * 4;
  p.T = big ? 2 : 1;
  p.unroll = p.T == 2 ? 1 : 2;
  if (variant / 10 >= 1 && variant / 10 <= 2) {
    p.T = variant / 10;
    if (variant % 10 >= 1 && variant % 10 <= 2) p.unroll = variant % 10;
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
  p.src_per_split = (int64_t)p.tiles_per_split * kTile;
  // Small fields: with one tile per split and still less than one wave of CTAs (900 particles: 8 x 8 CTAs on 148
  // SMs, each walking 128 sources with one warp per scheduler to hide its latencies) the FP64 kernels split the
  // sources finer than a tile, down to kMinSrcPerSplit per CTA, until one wave is full.
  // The same freedom evens out fields of a few waves: 4 900 particles are 39 x 39 = 1 521 CTAs = 1.7 waves of 888;
  // 45 splits of 109 sources make it 1 755 = 1.98 waves of smaller CTAs.
  const int64_t one_wave = (int64_t)sm_count * ctas_per_sm;
  if (kind != PLAN_UJ_F32 && p.tiles_per_split == 1 && nblk * p.nsplit < 4 * one_wave) {
    const int64_t waves = std::max<int64_t>(1, (nblk * p.nsplit + one_wave - 1) / one_wave);
    const int64_t ws = std::max<int64_t>(1, waves * one_wave / nblk);  // splits that fill `waves` waves
    int64_t sps = std::max<int64_t>(kMinSrcPerSplit, (std::max<int64_t>(ns, 1) + ws - 1) / ws);
    // below one wave more splits only lengthen the finish kernel's sums: at most kMaxSmallSplits of them
    // (measured: 900 particles 57 splits 93 us per call, 29-32 splits 86 us; 200 particles 13 splits 63 us, 7: 74 us)
    if (nblk * p.nsplit < one_wave) sps = std::max<int64_t>(sps, (std::max<int64_t>(ns, 1) + kMaxSmallSplits - 1) / kMaxSmallSplits);
    if (sps < kTile) {
      p.src_per_split = sps;
      p.nsplit = (int)((std::max<int64_t>(ns, 1) + sps - 1) / sps);
    }
  }
  p.pstride = round_up(std::max<int64_t>(nt, 1), 32);
  p.grid = dim3((unsigned)nblk, (unsigned)p.nsplit, 1);
  return p;
}

unsigned blocks_for(int64_t n, int threads) { return (unsigned)std::max<int64_t>(1, (n + threads - 1) / threads); }

// finish kernel of a U/J sweep: many splits on a small field -> the wide form (one thread per target and accumulator)
void launch_uj_finish(const UjFinishArgs &f, cudaStream_t st) {
  if (f.nt <= 0) return;
  if (f.nsplit >= 4 && f.nt <= (1 << 17))
    uj_finish_wide_kernel<<<blocks_for(f.nt, kFinishWideTargets), 256, 0, st>>>(f);
  else
    uj_finish_kernel<<<blocks_for(f.nt, 256), 256, 0, st>>>(f);
}

void launch_sfs_finish(const SfsFinishArgs &f, cudaStream_t st) {
  if (f.nt <= 0) return;
  if (f.nsplit >= 4 && f.nt <= (1 << 17))
    sfs_finish_wide_kernel<<<blocks_for(f.nt, 64), 256, 0, st>>>(f);
  else
    sfs_finish_kernel<<<blocks_for(f.nt, 256), 256, 0, st>>>(f);
}

constexpr double kTabNearFraction = 0.40;  // automatic choice: table kernel from this share of sampled warps with a near pair on

// Plan of the table kernel (vpm_kernels_tab.cuh): one CTA per SM, 512 (or 384) threads x 2 targets.
// `fills` tells whether the field is large enough to give every SM >= 8 CTAs even at one
// source tile per CTA -- below that the 128-thread kernels balance better.
// variant (VPM_OPT_UJ_VARIANT): 41 / 42 = 512 threads, unroll 1 / 2;  31 / 32 = 384 threads.
Plan make_plan_tab(int64_t nt, int64_t ns, int sm_count, bool *fills, int variant = 0) {
  Plan p;
  p.tab = (variant / 10 == 3) ? 384 : kTabThreads;
  p.T = kTabT;
  p.unroll = (variant % 10 == 2) ? 2 : 1;
  const int64_t per = (int64_t)p.tab * kTabT;
  const int64_t ntiles = std::max<int64_t>(1, (ns + kTile - 1) / kTile);
  const int64_t nblk = std::max<int64_t>(1, (nt + per - 1) / per);
  if (fills) *fills = nblk * ntiles >= (int64_t)sm_count * 8;
  const bool big = nblk * std::max<int64_t>(1, ntiles / 4) >= (int64_t)sm_count * 16;
  const int min_tiles = big ? 4 : 1;
  const int64_t want_ctas = (int64_t)sm_count * 16;
  int64_t nsplit = (want_ctas + nblk - 1) / nblk;
  nsplit = std::max<int64_t>(1, std::min<int64_t>(nsplit, std::max<int64_t>(1, ntiles / min_tiles)));
  nsplit = std::min<int64_t>(nsplit, 1024);
  p.tiles_per_split = (int)((ntiles + nsplit - 1) / nsplit);
  p.nsplit = (int)((ntiles + p.tiles_per_split - 1) / p.tiles_per_split);
  p.src_per_split = (int64_t)p.tiles_per_split * kTile;
  p.pstride = round_up(std::max<int64_t>(nt, 1), 32);
  p.grid = dim3((unsigned)nblk, (unsigned)p.nsplit, 1);
  return p;
}

template <int K, int THREADS, int UNROLL>
cudaError_t launch_uj_tab_V(const Plan &p, const UjArgs &a, cudaStream_t st) {
  const size_t smem = tab_smem_bytes<K>();
  cudaError_t e = cudaFuncSetAttribute(uj_pairs_tab_kernel<K, THREADS, UNROLL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  uj_pairs_tab_kernel<K, THREADS, UNROLL><<<p.grid, THREADS, smem, st>>>(a);
  return cudaSuccess;
}
template <int K>
cudaError_t launch_uj_tab_K(const Plan &p, const UjArgs &a, cudaStream_t st) {
  if (p.tab == 384) return p.unroll == 2 ? launch_uj_tab_V<K, 384, 2>(p, a, st) : launch_uj_tab_V<K, 384, 1>(p, a, st);
  return p.unroll == 2 ? launch_uj_tab_V<K, 512, 2>(p, a, st) : launch_uj_tab_V<K, 512, 1>(p, a, st);
}
cudaError_t launch_uj_tab(int kernel, const Plan &p, const UjArgs &a, cudaStream_t st) {
  return kernel == K_GERF ? launch_uj_tab_K<K_GERF>(p, a, st) : launch_uj_tab_K<K_GAUS>(p, a, st);
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


// U/J sweep: records from `src` columns [s0, s0+ns), targets tpos[0..nt), partial
// sums left in d.partial; the caller runs the finish kernel with its own output.
int uj_sweep(vpm_handle *h, Dev &d, cudaStream_t st, int kernel, const double *tpos, int64_t tld,
             int64_t nt, SrcView src, int64_t s0, int64_t ns, int flags, Plan &plan,
             bool time_pairs = false) {
  const int64_t ns_pad = round_up(std::max<int64_t>(ns, 1), kTile);
  TRY(scratch_acquire(h, d, st));
  if (flags & VPM_FLAG_FP32) {
    // optional FP32-arithmetic sweep (vpm_kernels_f32.cuh): FP32 records, FP64 partial sums in
    // the same layout, so the finish kernels are shared with the FP64 sweep
    TRY(ensure(h, d.rec, (size_t)ns_pad * kRecF * sizeof(float)));
    plan = make_plan(nt, ns, d.sm_count, PLAN_UJ_F32, h->opt_uj_variant);
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
  const bool use_const = h->opt_uj_const != 0;
  bool use_tab = false;
  if ((kernel == K_GERF || kernel == K_GAUS) && !use_const && h->opt_uj_table != 2 && nt > 0 && ns > 0) {
    bool fills = false;
    Plan pt = make_plan_tab(nt, ns, d.sm_count, &fills, h->opt_uj_variant);
    if (h->opt_uj_table == 1 || (fills && (flags & VPM_FLAG_NO_FARFIELD_SHORTCUT))) {
      plan = pt; use_tab = true;  // forced, or every pair is to take the regularised evaluation anyway
    } else if (fills) {
      // Automatic choice.  The table kernel wins where most WARPS see pairs inside the regularised range
      // (lanes that mix near and far run one code path, conflict-free gathers); where most warps are
      // entirely in the far field the round-1 kernel's smaller CTAs are ~5-10 % faster.  Decide on 2048
      // sampled (32 consecutive targets, source) warps (deterministic hash; one small kernel and a 4-byte
      // read-back, i.e. one synchronisation of `st` -- only reached for sweeps of >= 1e8 pairs).
      TRY(ensure(h, d.cubtmp, 256));
      unsigned int *cnt = (unsigned int *)d.cubtmp.p;
      CK(h, cudaMemsetAsync(cnt, 0, sizeof(unsigned int), st));
      sample_near_kernel<<<kSampleBlocks, kSampleThreads, 0, st>>>(src, s0, ns, tpos, tld, nt,
                                                                   kernel == K_GERF ? kFarU_gerf : kFarU_gaus, cnt);
      unsigned int near = 0;
      CK(h, cudaMemcpyAsync(&near, cnt, sizeof near, cudaMemcpyDeviceToHost, st));
      CK(h, cudaStreamSynchronize(st));
      h->launches++;
      h->last_near_fraction = (double)near / (double)(kSampleBlocks * kSampleThreads / 32);
      if (h->last_near_fraction >= kTabNearFraction) { plan = pt; use_tab = true; }
    }
  }
  if (!use_tab) plan = make_plan(nt, ns, d.sm_count, PLAN_UJ, h->opt_uj_variant < 30 ? h->opt_uj_variant : 0);
  if (use_const) plan.nsplit = 1;  // the constant-bank form keeps ONE set of sums across its launches
  TRY(ensure(h, d.partial, (size_t)plan.nsplit * kAcc * plan.pstride * sizeof(double)));
  if (use_tab || (kernel == K_GERF && !use_const))  // kPairsTab: uj_pairs_kernel<gaussianerf> reads the G(u) rows too
    prep_uj_records_tab<<<blocks_for(ns_pad, 256), 256, 0, st>>>(src, s0, ns, ns_pad, kernel, (double *)d.rec.p);
  else
    prep_uj_records<<<blocks_for(ns_pad, 256), 256, 0, st>>>(src, s0, ns, ns_pad, kernel, (double *)d.rec.p);
  h->launches++;
  if (nt > 0 && ns > 0 && use_const) {
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
    a.src_per_split = plan.src_per_split;
    a.partial = (double *)d.partial.p; a.pstride = plan.pstride;
    a.shortcut = (flags & VPM_FLAG_NO_FARFIELD_SHORTCUT) ? 0 : 1;
    if (time_pairs) CK(h, cudaEventRecord(d.ev[6], st));
    if (use_tab) CK(h, launch_uj_tab(kernel, plan, a, st));
    else launch_uj(kernel, plan, a, st);
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
  TRY(scratch_acquire(h, d, st));
  TRY(ensure(h, d.srec, (size_t)ns_pad * kSfsRec * sizeof(double)));
  plan = make_plan(nt, ns, d.sm_count, PLAN_SFS, h->opt_sfs_variant);
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
    a.src_per_split = plan.src_per_split;
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

}  // namespace
