// vpm_csr.cuh -- device-side regrouping of FastMultipole's direct_list (Hook 3).
//
// The near-field kernels of vpm_leaf.cuh want the (target leaf, source leaf) list as a CSR
// by target leaf (list order kept inside a group), a table of work items (<= NT targets of
// one leaf each) and, with several GPUs, cuts of that table with balanced pair counts.  For
// the lists of SURVEY config C5 (2^24 particles, 2e7..9e7 list entries) doing this with host
// loops costs more than the pair kernel itself, so every O(n_pairs) step runs here:
//   count + validate + sortedness  ->  exclusive scan  ->  stable sort by target leaf (only
//   if the list is not already grouped)  ->  CTA-width choice  ->  work items  ->  cuts.
// All reductions are integer (atomicAdd on 64-bit counters), so the outcome is deterministic.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace vpm {

typedef unsigned long long u64;

// slots of the small statistics block read back by the host
enum { CS_BAD = 0,       // 1 + index of an out-of-range list entry (0 = none)
       CS_UNSORTED = 1,  // != 0 when the list is not grouped by target leaf already
       CS_W128 = 2, CS_W64 = 3, CS_W32 = 4,  // padded lane-work for the three CTA widths
       CS_PAIRS = 5,     // sum over list entries of nt * ns (pair visits)
       CS_NWI = 6,       // number of work items
       CS_OVERLAP = 7,   // != 0 when target leaves that carry work overlap or are out of body order
       CS_SLOTS = 8 };

// one thread per list entry: counts per target leaf (cnt[l + 1]), source bodies per target
// leaf (srcw[l]), validation and sortedness
__global__ void csr_count_kernel(const int32_t *__restrict__ pt, const int32_t *__restrict__ ps, int64_t npairs,
                                 int64_t ntl, int64_t nsl, const int64_t *__restrict__ sb,
                                 const int64_t *__restrict__ se, u64 *__restrict__ cnt, u64 *__restrict__ srcw,
                                 u64 *__restrict__ stats) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= npairs) return;
  const int32_t t = pt[k], s = ps[k];
  if (t < 0 || t >= ntl || s < 0 || s >= nsl) {
    atomicMin(&stats[CS_BAD], (u64)k + 1);
    return;
  }
  atomicAdd(&cnt[t + 1], 1ull);
  atomicAdd(&srcw[t], (u64)(se[s] - sb[s]));
  if (k > 0 && pt[k - 1] > t) stats[CS_UNSORTED] = 1;
}

// one thread per target leaf: padded lane-work of the three candidate CTA widths and the
// number of pair visits
__global__ void csr_cand_kernel(const int64_t *__restrict__ tb, const int64_t *__restrict__ te, int64_t ntl,
                                const u64 *__restrict__ srcw, u64 *__restrict__ stats) {
  const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  u64 w128 = 0, w64 = 0, w32 = 0, pairs = 0;
  if (l < ntl) {
    const u64 sz = (u64)(te[l] - tb[l]), w = srcw[l];
    w128 = (sz + 127) / 128 * 128 * w;
    w64 = (sz + 63) / 64 * 64 * w;
    w32 = (sz + 31) / 32 * 32 * w;
    pairs = sz * w;
  }
  // warp reduction, then one atomic per warp
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    w128 += __shfl_down_sync(0xffffffffu, w128, o);
    w64 += __shfl_down_sync(0xffffffffu, w64, o);
    w32 += __shfl_down_sync(0xffffffffu, w32, o);
    pairs += __shfl_down_sync(0xffffffffu, pairs, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (w128) atomicAdd(&stats[CS_W128], w128);
    if (w64) atomicAdd(&stats[CS_W64], w64);
    if (w32) atomicAdd(&stats[CS_W32], w32);
    if (pairs) atomicAdd(&stats[CS_PAIRS], pairs);
  }
}

// Work items are laid out in BODY order (rank r = position of a leaf when the target leaves are
// sorted by their first body; ord[r] = leaf), not in leaf-index order: the caller's leaf table
// may be level-ordered and contain interior branches (FastMultipole's branch array), but a
// device's share of the items must still be one contiguous range of target columns.
// work items of the leaf of rank r: ceil(size / nt) when the leaf has list entries, else none
__global__ void csr_wi_count_kernel(const int64_t *__restrict__ tb, const int64_t *__restrict__ te, int64_t ntl,
                                    const u64 *__restrict__ ptr, const int32_t *__restrict__ ord, int nt,
                                    u64 *__restrict__ wcnt) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= ntl) return;
  const int64_t l = ord[r];
  const int64_t sz = te[l] - tb[l];
  wcnt[r] = ptr[l + 1] > ptr[l] ? (u64)((sz + nt - 1) / nt) : 0ull;
}

// fill the work items of the leaf of rank r at wofs[r].. and their pair counts (for the multi-GPU
// cuts); flag leaves with work whose body range starts before the previous such leaf ended
__global__ void csr_wi_fill_kernel(const int64_t *__restrict__ tb, const int64_t *__restrict__ te, int64_t ntl,
                                   const u64 *__restrict__ wofs, const u64 *__restrict__ wcnt,
                                   const u64 *__restrict__ srcw, const int32_t *__restrict__ ord, int nt,
                                   int32_t *__restrict__ wi_leaf, int32_t *__restrict__ wi_off,
                                   u64 *__restrict__ wi_w, u64 *__restrict__ stats) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= ntl) return;
  const int64_t l = ord[r];
  const u64 n = wcnt[r], o = wofs[r];
  const int64_t sz = te[l] - tb[l];
  for (u64 j = 0; j < n; ++j) {
    const int64_t off = (int64_t)j * nt;
    wi_leaf[o + j] = (int32_t)l;
    wi_off[o + j] = (int32_t)off;
    const int64_t c = sz - off < nt ? sz - off : nt;
    wi_w[o + j] = (u64)c * srcw[l];
  }
  if (n > 0) {
    // previous leaf with work, in body order
    for (int64_t q = r - 1; q >= 0; --q) {
      if (wcnt[q] > 0) {
        if (te[ord[q]] > tb[l]) stats[CS_OVERLAP] = 1;
        break;
      }
    }
  }
  if (r == ntl - 1) stats[CS_NWI] = o + n;
}

// The list of vpm_nearfield_ranges generated in place: entry i belongs to the target range k with
// off[k] <= i < off[k+1] (binary search) and its source "leaf" is the i-th source range itself:
// pair_tgt[i] = owner[k], pair_src[i] = i.  (Built on the host and uploaded this was 0.7 GB at 2^24 / ncrit 128.)
__global__ void csr_gen_pairs_kernel(const int64_t *__restrict__ off, const int64_t *__restrict__ owner, int64_t ntr,
                                     int64_t n, int32_t *__restrict__ pair_tgt, int32_t *__restrict__ pair_src) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t lo = 0, hi = ntr;  // invariant: off[lo] <= i < off[hi]
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (off[mid] <= i) lo = mid; else hi = mid;
  }
  pair_tgt[i] = (int32_t)owner[lo];
  pair_src[i] = (int32_t)i;
}

__global__ void csr_iota_kernel(int32_t *p, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int32_t)i;
}

// cut g (0 <= g <= G): first work item k whose exclusive pair-count prefix reaches total * g / G
// (cut 0 = 0, cut G = nwi).  out[5g..5g+4] = (k, leaf and offset of item k, leaf and offset of
// item k - 1): the host turns them into the contiguous target-column range of each device.
__global__ void csr_cut_kernel(const u64 *__restrict__ wsum /* inclusive */, const int32_t *__restrict__ wi_leaf,
                               const int32_t *__restrict__ wi_off, int64_t nwi, int G, int64_t *__restrict__ out) {
  const int g = threadIdx.x;
  if (g > G || nwi <= 0) return;
  int64_t k;
  if (g == 0) k = 0;
  else if (g == G) k = nwi;
  else {
    const double target = (double)wsum[nwi - 1] * g / G;
    int64_t lo = 0, hi = nwi;
    while (lo < hi) {
      const int64_t mid = (lo + hi) / 2;
      const double w = mid == 0 ? 0.0 : (double)wsum[mid - 1];
      if (w < target) lo = mid + 1; else hi = mid;
    }
    k = lo;
  }
  const int64_t ka = k < nwi ? k : nwi - 1, kb = k > 0 ? k - 1 : 0;
  out[5 * g] = k;
  out[5 * g + 1] = wi_leaf[ka];
  out[5 * g + 2] = wi_off[ka];
  out[5 * g + 3] = wi_leaf[kb];
  out[5 * g + 4] = wi_off[kb];
}

__global__ void csr_zero_kernel(u64 *p, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0ull;
}
__global__ void csr_init_stats_kernel(u64 *stats) {
  if (threadIdx.x < CS_SLOTS) stats[threadIdx.x] = threadIdx.x == CS_BAD ? ~0ull : 0ull;
}

}  // namespace vpm
