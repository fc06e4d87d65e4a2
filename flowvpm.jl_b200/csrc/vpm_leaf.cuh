// vpm_leaf.cuh -- FMM near-field form of the two sweeps: the pair arithmetic of
// vpm_kernels.cuh evaluated over a list of (target leaf, source leaf) pairs
// (FastMultipole's direct_list; SURVEY 3.3, reference call shape
// src/FLOWVPM_gpu.jl:637-643 and src/FLOWVPM_subfilterscale_models.jl:94-188).
//
// The list is regrouped by target leaf (CSR, list order kept inside a group) so
// that one CTA owns up to kThreads targets of one leaf and walks all of that
// leaf's source leaves: one owner per target, no atomics, and the per-target
// summation order is the order of the pairs in the reference's list.
#pragma once
#include "vpm_kernels.cuh"

namespace vpm {

struct LeafCsr {
  const int32_t *wi_leaf;      // work item -> target leaf
  const int32_t *wi_off;       // work item -> offset of its first target inside the leaf
  const int64_t *tleaf_begin;  // half-open sorted-body ranges of the target leaves
  const int64_t *tleaf_end;
  const int64_t *csr_ptr;      // [n_tgt_leaves + 1] into csr_src
  const int32_t *csr_src;      // source leaves of each target leaf, list order
  const int64_t *sleaf_begin;
  const int64_t *sleaf_end;
};

// Warp-cooperative tile producer.  A tile is filled with up to TILE bodies taken in list order and may span
// several source leaves (small leaves -- ncrit ~ 64 means ~30 bodies -- would otherwise make tiles of a
// quarter of their capacity, each paying a barrier round trip and an exposed TMA latency); the per-target
// summation order is the list order either way.  Round 1 walked the list piece by piece with three
// dependent global loads per piece, on one thread, and every consumer thread repeated the walk to learn
// the tile sizes: for CPU-style trees (ncrit 10..50: a 64-record tile is ~5 leaves) that serial
// latency was as long as the arithmetic of the tile.  Here the 32 lanes of warp 0 each load ONE piece of
// the list (csr_src -> leaf range), a warp scan places the pieces in the tile, lane 0 announces the bytes
// of the step on the barrier (mbarrier.expect_tx, no arrival), every lane issues the bulk copy of its own
// piece, and when the tile is full (or the list exhausted) lane 0 publishes the record count in shared
// memory and arrives: one step of parallel loads per tile, and the consumers read the size instead of walking.
__device__ __forceinline__ void mbar_expect_tx_only(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
struct LeafTileProducer {
  int64_t li, le;  // position / end in csr_src (uniform over the warp)
  int off;         // records of piece `li` already consumed
  __device__ __forceinline__ void init(const LeafCsr &c, int leaf) {
    li = c.csr_ptr[leaf];
    le = c.csr_ptr[leaf + 1];
    off = 0;
  }
  // Called by all 32 lanes of ONE warp.  Fills the tile with up to TILE records of RECW doubles in list
  // order; returns their number (0: list exhausted -- the barrier still completes so that the consumers wake up).
  template <int TILE, int RECW>
  __device__ __forceinline__ int issue(const LeafCsr &c, const double *__restrict__ rec, double *tile_smem,
                                       uint64_t *bar, int *tile_n) {
    const int lane = threadIdx.x & 31;
    int filled = 0;
    while (filled < TILE && li < le) {
      const int64_t k = li + lane;
      const bool have = k < le;
      int64_t b = 0;
      int len = 0;
      if (have) {
        const int s = c.csr_src[k];
        b = c.sleaf_begin[s];
        len = (int)(c.sleaf_end[s] - b);
      }
      if (lane == 0) { b += off; len -= off; }   // the rest of a piece the previous tile cut
      int incl = len;                            // inclusive scan of the piece lengths
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const int excl = incl - len, room = TILE - filled;
      const int take = max(0, min(len, room - excl));
      const int step = min(room, __shfl_sync(0xffffffffu, incl, 31));  // records this step adds
      if (lane == 0) mbar_expect_tx_only(bar, (uint32_t)step * RECW * sizeof(double));
      __syncwarp();
      if (take > 0)
        tma_bulk_g2s(tile_smem + (size_t)(filled + excl) * RECW, rec + b * RECW, (uint32_t)take * RECW * sizeof(double), bar);
      // finished pieces (take == len, empty ones included) form a prefix of the lanes that hold a piece
      const unsigned fin = __ballot_sync(0xffffffffu, have && take == len);
      const int nfull = fin == 0xffffffffu ? 32 : __ffs(~fin) - 1;
      const int part = nfull < 32 ? __shfl_sync(0xffffffffu, take, nfull) : 0;  // cut piece: records taken now
      off = (nfull == 0 ? off : 0) + part;
      li += nfull;
      filled += step;
    }
    if (lane == 0) {
      *tile_n = filled;
      mbar_arrive(bar);
    }
    return filled;
  }
};

// gaussianerf in the leaf kernels: G(u) rows of vpm_kernels_tab.cuh read through L1 (TabGlobal) and the records of
// prep_uj_records_tab instead of round 1's degree-9 table (three 16-byte gathers and 9 DFMA per pair instead of
// five and 18).  Measured on the 2^20 cloud (G pairs/s): ncrit 24: 163.9 -> 205.6, 50: 215.0 -> 227.9,
// 128: 282.7 -> 300.6, 512: 296.0 -> 302.7, 1600: 328.9 -> 328.1 (mostly far pairs there).  The gaussian family
// measured 1 % slower on its table here (its regularised range ends at s = 3.45: few pairs) and keeps ab_gaus.
template <int K>
constexpr bool kLeafTab = K == K_GERF;

struct LeafUjArgs {
  LeafCsr csr;
  const double *tpos;  // sorted target buffer: tpos[i*tld + 0..2]
  int64_t tld;
  const double *rec;   // U/J records of the sorted source buffer
  double *out;         // sorted target buffer (same matrix as tpos), ld = tld
  int urow, jrow;
  int want_U, want_J;
  int shortcut;
};

// NT threads per CTA (= targets per work item) and TILE records per shared-memory tile are
// chosen on the host from the leaf-size distribution: small leaves (ncrit ~ 64) would leave
// most lanes of a 128-thread CTA idle.
template <int K, int NT, int TILE>
__global__ void __launch_bounds__(NT) uj_leaf_kernel(const LeafUjArgs a) {
  __shared__ __align__(128) double tiles[kStages][TILE * kRec];
  __shared__ __align__(8) uint64_t full[kStages];
  __shared__ int tile_n[kStages];  // records in each stage's tile, published by the producer
  // gaussianerf: the 13 KB G(u) table is staged in shared memory only by CTAs wide enough to amortise the
  // copy; a one- or two-warp CTA of a small leaf reads it through L1 instead (its whole work item is a few
  // thousand pairs: the copy alone cost a third of the kernel at ncrit 24: 101 -> 139 G pairs/s)
  constexpr bool kTabInSmem = K == K_GERF && NT >= 128 && !kLeafTab<K>;
  __shared__ __align__(16) double2 gtab_s[kTabInSmem ? kGerfIntervals * kGerfCoeffs / 2 : 1];
  if constexpr (kTabInSmem) load_gerf_table(gtab_s);  // visible after the __syncthreads below
  const double2 *gtab = kTabInSmem ? gtab_s : reinterpret_cast<const double2 *>(kGerfTable);
  const int tid = threadIdx.x;
  const int leaf = a.csr.wi_leaf[blockIdx.x];
  const int64_t tb = a.csr.tleaf_begin[leaf] + a.csr.wi_off[blockIdx.x];
  int64_t te = a.csr.tleaf_end[leaf];
  if (te > tb + NT) te = tb + NT;
  // A warp with <= 16 live targets (the tail of a leaf, or a whole small leaf: CPU-style trees
  // have ncrit 10..50) would idle most of its lanes: its lanes form nsplit groups that own the
  // same targets and share the sources of every tile between them (uj_tile<SPLIT>).
  const int wbase = (tid >> 5) << 5, lane = tid & 31;
  const int64_t wlive = te - (tb + wbase);  // live targets of this warp (may be <= 0 when NT > 32)
  const int nsplit = wlive > 16 ? 1 : wlive > 8 ? 2 : wlive > 4 ? 4 : 8;
  const int glanes = 32 / nsplit, phase = lane / glanes;
  const int64_t i = tb + wbase + (lane % glanes);
  const bool valid = i < te;
  const double *p = a.tpos + (valid ? i : te - 1) * a.tld;
  double tx[1] = {p[0]}, ty[1] = {p[1]}, tz[1] = {p[2]};
  double acc[1][kAcc];
#pragma unroll
  for (int k = 0; k < kAcc; ++k) acc[0][k] = 0.0;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  LeafTileProducer prod;  // state lives (uniformly) in the lanes of warp 0
  prod.init(a.csr, leaf);
  if (tid < 32)
    for (int s = 0; s < kStages; ++s) prod.issue<TILE, kRec>(a.csr, a.rec, &tiles[s][0], &full[s], &tile_n[s]);

  for (int it = 0;; ++it) {
    const int st = it % kStages;
    mbar_wait(&full[st], (uint32_t)((it / kStages) & 1));
    const int n = tile_n[st];
    if (n == 0) break;  // list exhausted (the same value for every thread of the CTA)
    const double2 *tile = reinterpret_cast<const double2 *>(&tiles[st][0]);
    if constexpr (kLeafTab<K>) {
      if (nsplit == 1) uj_tile_tab<K, 1, 2, false>(tile, n, tx, ty, tz, acc, a.shortcut, tab_global<K>());
      else uj_tile_tab<K, 1, 2, true>(tile, n, tx, ty, tz, acc, a.shortcut, tab_global<K>(), nsplit, phase);
    } else {
      if (nsplit == 1) uj_tile<K, 1, 2>(tile, n, tx, ty, tz, acc, a.shortcut, gtab);
      else uj_tile<K, 1, 2, false, true>(tile, n, tx, ty, tz, acc, a.shortcut, gtab, nsplit, phase);
    }
    __syncthreads();
    if (tid < 32) prod.issue<TILE, kRec>(a.csr, a.rec, &tiles[st][0], &full[st], &tile_n[st]);
  }
  // add the groups' sums (fixed order: deterministic); afterwards every group holds the total
  for (int o = glanes; o < 32; o <<= 1) {
#pragma unroll
    for (int k = 0; k < kAcc; ++k) acc[0][k] += __shfl_xor_sync(0xffffffffu, acc[0][k], o);
  }

  if (valid && phase == 0) {
    double U[3], J[9];
    finish_sums(acc[0], U, J);
    double *o = a.out + i * a.tld;
    if (a.want_U) {
#pragma unroll
      for (int k = 0; k < 3; ++k) o[a.urow + k] += U[k];
    }
    if (a.want_J) {
#pragma unroll
      for (int k = 0; k < 9; ++k) o[a.jrow + k] += J[k];
    }
  }
}

// One-warp CTAs (work items of <= 32 targets: the leaves of CPU-style trees, ncrit 10..50): flexible lane
// groups and the two-targets-per-lane mapping.
template <int K, int TILE>
__global__ void __launch_bounds__(32, 16) uj_leaf1w_kernel(const LeafUjArgs a) {
  constexpr int NT = 32;
  __shared__ __align__(128) double tiles[kStages][TILE * kRec];
  __shared__ __align__(8) uint64_t full[kStages];
  __shared__ int tile_n[kStages];  // records in each stage's tile, published by the producer
  // gaussianerf: the 13 KB G(u) table is staged in shared memory only by CTAs wide enough to amortise the
  // copy; a one- or two-warp CTA of a small leaf reads it through L1 instead (its whole work item is a few
  // thousand pairs: the copy alone cost a third of the kernel at ncrit 24: 101 -> 139 G pairs/s)
  constexpr bool kTabInSmem = K == K_GERF && NT >= 128 && !kLeafTab<K>;
  __shared__ __align__(16) double2 gtab_s[kTabInSmem ? kGerfIntervals * kGerfCoeffs / 2 : 1];
  if constexpr (kTabInSmem) load_gerf_table(gtab_s);  // visible after the __syncthreads below
  const double2 *gtab = kTabInSmem ? gtab_s : reinterpret_cast<const double2 *>(kGerfTable);
  const int tid = threadIdx.x;
  const int leaf = a.csr.wi_leaf[blockIdx.x];
  const int64_t tb = a.csr.tleaf_begin[leaf] + a.csr.wi_off[blockIdx.x];
  int64_t te = a.csr.tleaf_end[leaf];
  if (te > tb + NT) te = tb + NT;
  // Lane mapping, chosen per warp from its number of live targets L (<= 32; CPU-style trees have ncrit 10..50,
  // i.e. ~12 bodies per leaf, and the tail of any leaf is short).  The lanes form `nsplit` groups of G lanes;
  // every group owns ALL the warp's targets and takes the records j = phase, phase + nsplit, ... of each tile
  // (uj_tile<SPLIT>); the groups' sums are added at the end in a fixed order.  Two mappings:
  //   one target per lane:   G = L,          nsplit = 32 / L   -> L * nsplit pairs per ~100..120-cycle iteration
  //   two targets per lane:  G = ceil(L/2),  nsplit = 32 / G   -> L * nsplit pairs per ~215-cycle iteration
  // (the second shares every source operand between two pairs; L = 12: 30 busy lanes instead of 24).
  const int wbase = (tid >> 5) << 5, lane = tid & 31;
  const int64_t wl64 = te - (tb + wbase);
  const int L = wl64 <= 0 ? 0 : (wl64 > 32 ? 32 : (int)wl64);  // live targets of this warp (0 when NT > 32 and the item is short)
  const int G2 = (L + 1) >> 1;
  const int ns1 = L > 0 ? 32 / L : 1, ns2 = G2 > 0 ? 32 / G2 : 1;
  // measured costs per warp-iteration: ~100 cycles one target without lane groups, ~120 with, ~215 two targets
  // (only these one-warp CTAs carry the two-target path: wider CTAs serve bigger leaves, whose full warps take
  // the plain one-target loop, and the extra registers cost them a CTA per SM: measured 325 -> 313 G/s)
  constexpr int TM = 2;
  const bool two = ns1 == 1 ? ns2 * 100 > 215 : ns2 * 120 > ns1 * 215;
  const int G = two ? G2 : (L > 0 ? L : 32);
  const int nsplit = two ? ns2 : ns1;
  const int phase = lane / G, g = lane - phase * G;      // lanes >= nsplit * G idle: phase == nsplit
  const bool active = phase < nsplit && L > 0;
  const int64_t i0 = tb + wbase + g, i1 = i0 + G;       // second target only in the two-target mapping
  const bool valid0 = active && g < L, valid1 = active && two && g + G < L;
  const double *p0 = a.tpos + (valid0 ? i0 : te - 1) * a.tld;
  const double *p1 = a.tpos + (valid1 ? i1 : te - 1) * a.tld;
  double tx1[1] = {p0[0]}, ty1[1] = {p0[1]}, tz1[1] = {p0[2]};
  double tx2[2] = {p0[0], p1[0]}, ty2[2] = {p0[1], p1[1]}, tz2[2] = {p0[2], p1[2]};
  double acc[TM][kAcc];
#pragma unroll
  for (int t = 0; t < TM; ++t)
#pragma unroll
    for (int k = 0; k < kAcc; ++k) acc[t][k] = 0.0;
  // an idle lane joins group 0's schedule with a phase past the tile (A = B = 0 for every record)
  const int ph = active ? phase : 0x3fffffff;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  LeafTileProducer prod;  // state lives (uniformly) in the lanes of warp 0
  prod.init(a.csr, leaf);
  if (tid < 32)
    for (int s = 0; s < kStages; ++s) prod.issue<TILE, kRec>(a.csr, a.rec, &tiles[s][0], &full[s], &tile_n[s]);

  for (int it = 0;; ++it) {
    const int st = it % kStages;
    mbar_wait(&full[st], (uint32_t)((it / kStages) & 1));
    const int n = tile_n[st];
    if (n == 0) break;  // list exhausted (the same value for every thread of the CTA)
    const double2 *tile = reinterpret_cast<const double2 *>(&tiles[st][0]);
    if constexpr (kLeafTab<K>) {
      if (two) {
        uj_tile_tab<K, 2, 1, true>(tile, n, tx2, ty2, tz2, acc, a.shortcut, tab_global<K>(), nsplit, ph);
      } else {
        double (&acc1)[1][kAcc] = *reinterpret_cast<double (*)[1][kAcc]>(&acc[0]);
        if (nsplit == 1) uj_tile_tab<K, 1, 2, false>(tile, n, tx1, ty1, tz1, acc1, a.shortcut, tab_global<K>());
        else uj_tile_tab<K, 1, 2, true>(tile, n, tx1, ty1, tz1, acc1, a.shortcut, tab_global<K>(), nsplit, ph);
      }
    } else if (two) {
      uj_tile<K, 2, 1, false, true>(tile, n, tx2, ty2, tz2, acc, a.shortcut, gtab, nsplit, ph);
    } else {
      double (&acc1)[1][kAcc] = *reinterpret_cast<double (*)[1][kAcc]>(&acc[0]);
      if (nsplit == 1) uj_tile<K, 1, 2>(tile, n, tx1, ty1, tz1, acc1, a.shortcut, gtab);
      else uj_tile<K, 1, 2, false, true>(tile, n, tx1, ty1, tz1, acc1, a.shortcut, gtab, nsplit, ph);
    }
    __syncthreads();
    if (tid < 32) prod.issue<TILE, kRec>(a.csr, a.rec, &tiles[st][0], &full[st], &tile_n[st]);
  }
  // add the groups' sums into group 0, phase by phase (fixed order: deterministic)
  for (int pgrp = 1; pgrp < nsplit; ++pgrp) {
    const int srcl = (lane + pgrp * G) & 31;
#pragma unroll
    for (int t = 0; t < TM; ++t) {
      if (t == 1 && !two) break;  // uniform over the warp
#pragma unroll
      for (int k = 0; k < kAcc; ++k) {
        const double v = __shfl_sync(0xffffffffu, acc[t][k], srcl);
        if (phase == 0) acc[t][k] += v;  // only group 0 accumulates: the other groups' sums must stay as they are
      }
    }
  }

  if (phase == 0) {
#pragma unroll
    for (int t = 0; t < TM; ++t) {
      if (!(t == 0 ? valid0 : valid1)) continue;
      double U[3], J[9];
      finish_sums(acc[t], U, J);
      double *o = a.out + (t == 0 ? i0 : i1) * a.tld;
      if (a.want_U) {
#pragma unroll
        for (int k = 0; k < 3; ++k) o[a.urow + k] += U[k];
      }
      if (a.want_J) {
#pragma unroll
        for (int k = 0; k < 9; ++k) o[a.jrow + k] += J[k];
      }
    }
  }
}

struct LeafSfsArgs {
  LeafCsr csr;
  const double *tpos;  // particle-indexed: tpos[c*tld + 0..2]
  int64_t tld;
  const double *tJ;    // tJ[c*jld + 0..8]
  int64_t jld;
  const int64_t *tindex;  // sorted target body -> particle column
  const double *rec;   // SFS records in sorted source order
  double *out;         // out[c*old + orow + 0..2] +=
  int64_t old;
  int orow;
  int transposed;
  int shortcut;
  int obody = 0;       // 1: out is indexed by sorted body (multi-GPU: columns return to device 0)
};

template <int K, int NT, int TILE, int MODE = MODE_SFS>
__global__ void __launch_bounds__(NT) sfs_leaf_kernel(const LeafSfsArgs a) {
  __shared__ __align__(128) double tiles[kStages][TILE * kSfsRec];
  __shared__ __align__(8) uint64_t full[kStages];
  __shared__ int tile_n[kStages];  // records in each stage's tile, published by the producer
  const int tid = threadIdx.x;
  const int leaf = a.csr.wi_leaf[blockIdx.x];
  const int64_t tb = a.csr.tleaf_begin[leaf] + a.csr.wi_off[blockIdx.x];
  int64_t te = a.csr.tleaf_end[leaf];
  if (te > tb + NT) te = tb + NT;
  // sparse warps share the sources of a tile between lane groups, as in uj_leaf_kernel
  const int wbase = (tid >> 5) << 5, lane = tid & 31;
  const int64_t wlive = te - (tb + wbase);
  const int nsplit = wlive > 16 ? 1 : wlive > 8 ? 2 : wlive > 4 ? 4 : 8;
  const int glanes = 32 / nsplit, phase = lane / glanes;
  const int64_t i = tb + wbase + (lane % glanes);
  const bool valid = i < te;
  const int64_t c = a.tindex[valid ? i : te - 1];
  const double *p = a.tpos + c * a.tld;
  double tx[1] = {p[0]}, ty[1] = {p[1]}, tz[1] = {p[2]};
  double JT[1][9], acc[1][3] = {{0.0, 0.0, 0.0}};
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      if constexpr (MODE == MODE_SFS) {
        const double *j = a.tJ + c * a.jld;
        JT[0][3 * k + m] = a.transposed ? j[3 * k + m] : j[k + 3 * m];
      } else {
        JT[0][3 * k + m] = 0.0;
      }
    }

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  LeafTileProducer prod;  // state lives (uniformly) in the lanes of warp 0
  prod.init(a.csr, leaf);
  if (tid < 32)
    for (int s = 0; s < kStages; ++s) prod.issue<TILE, kSfsRec>(a.csr, a.rec, &tiles[s][0], &full[s], &tile_n[s]);

  for (int it = 0;; ++it) {
    const int st = it % kStages;
    mbar_wait(&full[st], (uint32_t)((it / kStages) & 1));
    const int n = tile_n[st];
    if (n == 0) break;  // list exhausted (the same value for every thread of the CTA)
    const double2 *tile = reinterpret_cast<const double2 *>(&tiles[st][0]);
    if (nsplit == 1) sfs_tile<K, 1, MODE>(tile, n, tx, ty, tz, JT, acc, a.shortcut);
    else sfs_tile<K, 1, MODE, true>(tile, n, tx, ty, tz, JT, acc, a.shortcut, nsplit, phase);
    __syncthreads();
    if (tid < 32) prod.issue<TILE, kSfsRec>(a.csr, a.rec, &tiles[st][0], &full[st], &tile_n[st]);
  }
  for (int o = glanes; o < 32; o <<= 1) {
#pragma unroll
    for (int k = 0; k < 3; ++k) acc[0][k] += __shfl_xor_sync(0xffffffffu, acc[0][k], o);
  }

  if (valid && phase == 0) {
    double *o = a.out + (a.obody ? i : c) * a.old + a.orow;
    o[0] += acc[0][0];
    o[1] += acc[0][1];
    o[2] += acc[0][2];
  }
}

// out[tindex[i]] += sorted[i] for sorted bodies [i0, i1)  (3 values per body)
__global__ void add_sorted3_kernel(const double *__restrict__ sorted, const int64_t *__restrict__ tindex, int64_t i0,
                                   int64_t i1, double *__restrict__ out) {
  const int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= i1) return;
  const int64_t c = tindex[i];
  out[c * 3] += sorted[i * 3];
  out[c * 3 + 1] += sorted[i * 3 + 1];
  out[c * 3 + 2] += sorted[i * 3 + 2];
}

// M[c*ld + row + k] (+)= v[c*3 + k], k = 0..2, for every particle c (zeta_fmm on the resident field)
__global__ void add_rows3_kernel(const double *__restrict__ v, int64_t np, double *__restrict__ M, int64_t ld, int row,
                                 int accumulate) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= np) return;
  double *o = M + c * ld + row;
#pragma unroll
  for (int k = 0; k < 3; ++k) o[k] = (accumulate ? o[k] : 0.0) + v[c * 3 + k];
}

}  // namespace vpm
