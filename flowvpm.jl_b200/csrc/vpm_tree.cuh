// vpm_tree.cuh -- device-side leaf lists for the FMM near field.
//
// FastMultipole.jl owns the tree of UJ_fmm (src/FLOWVPM_UJ.jl:90-101) and is not part of the
// reference tree, so there is no reference arithmetic to restate here; what the near-field
// hook needs from a tree is (i) a permutation that makes every leaf a contiguous body range,
// (ii) the leaf ranges and (iii) the list of (target leaf, source leaf) pairs that fail the
// multipole acceptance criterion  (r_i + r_j) <= theta * d  (src/FLOWVPM_particlefield.jl:28-36
// sets theta = 0.4), leaf radii padded by the largest core size in the leaf.  This builder
// produces those on the GPU with a uniform cell grid (mean occupancy ~ ncrit/2, as the host
// stand-in it replaces): cell keys -> stable radix sort -> run-length leaves -> leaf spheres
// -> stencil search with the MAC, one warp per target leaf, list emitted in (target, source)
// lexicographic order, i.e. already grouped by target leaf.
// Every floating-point operation is explicitly rounded (__dadd_rn / __dmul_rn / __ddiv_rn /
// __dsqrt_rn, no contraction) in one fixed order, so the lists are
// bit-identical to the CPU restatement the tests hold (integer / index work: exact parity).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "vpm_csr.cuh"

namespace vpm {

struct TreeGrid {
  double lo[3];
  double h;        // cell size
  int64_t dims[3];
  double theta;
};

// ---- bounding box: exact min / max per axis (block reduction + ordered-int atomics) -------
__device__ __forceinline__ long long dbl_to_ordered(double x) {
  long long b = __double_as_longlong(x);
  return b >= 0 ? b : b ^ 0x7fffffffffffffffLL;
}
__host__ __device__ __forceinline__ double ordered_to_dbl(long long o) {
  long long b = o >= 0 ? o : o ^ 0x7fffffffffffffffLL;
#ifdef __CUDA_ARCH__
  return __longlong_as_double(b);
#else
  double d;
  memcpy(&d, &b, 8);
  return d;
#endif
}
// bb[0..2] = ordered min, bb[3..5] = ordered max, bb[6] = ordered max sigma
__global__ void tree_bbox_init_kernel(long long *bb) {
  if (threadIdx.x < 3) bb[threadIdx.x] = 0x7fffffffffffffffLL;
  else if (threadIdx.x < 7) bb[threadIdx.x] = (long long)0x8000000000000000ULL;
}
__global__ void tree_bbox_kernel(const double *__restrict__ P, int64_t ld, int64_t n, long long *bb) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  long long mn[3] = {0x7fffffffffffffffLL, 0x7fffffffffffffffLL, 0x7fffffffffffffffLL};
  long long mx[3] = {(long long)0x8000000000000000ULL, (long long)0x8000000000000000ULL, (long long)0x8000000000000000ULL};
  if (i < n) {
#pragma unroll
    for (int c = 0; c < 3; ++c) mn[c] = mx[c] = dbl_to_ordered(P[i * ld + c]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      mn[c] = min(mn[c], __shfl_down_sync(0xffffffffu, mn[c], o));
      mx[c] = max(mx[c], __shfl_down_sync(0xffffffffu, mx[c], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      atomicMin(&bb[c], mn[c]);
      atomicMax(&bb[3 + c], mx[c]);
    }
  }
}

// ---- cell keys ---------------------------------------------------------------------------
__global__ void tree_keys_kernel(const double *__restrict__ P, int64_t ld, int64_t n, TreeGrid g,
                                 int64_t *__restrict__ keys, int64_t *__restrict__ idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t c[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double q = __ddiv_rn(__dsub_rn(P[i * ld + a], g.lo[a]), g.h);
    int64_t k = (int64_t)q;  // truncation, q >= 0
    c[a] = k < g.dims[a] - 1 ? k : g.dims[a] - 1;
  }
  keys[i] = (c[0] * g.dims[1] + c[1]) * g.dims[2] + c[2];
  idx[i] = i;
}

// ---- leaves = runs of equal keys in the sorted key array --------------------------------------
__global__ void tree_heads_kernel(const int64_t *__restrict__ skeys, int64_t n, u64 *__restrict__ head) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  head[i] = (i == 0 || skeys[i] != skeys[i - 1]) ? 1ull : 0ull;
}
// rank[i] = inclusive scan of head -> leaf id of body i is rank[i] - 1
__global__ void tree_leaves_kernel(const int64_t *__restrict__ skeys, const u64 *__restrict__ rank, int64_t n,
                                   int64_t *__restrict__ lbegin, int64_t *__restrict__ lend,
                                   int64_t *__restrict__ lkey, int32_t *__restrict__ cell_to_leaf) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t l = (int64_t)rank[i] - 1;
  if (i == 0 || skeys[i] != skeys[i - 1]) {
    lbegin[l] = i;
    lkey[l] = skeys[i];
    cell_to_leaf[skeys[i]] = (int32_t)l;
  }
  if (i == n - 1 || skeys[i] != skeys[i + 1]) lend[l] = i + 1;
}

// ---- leaf spheres: centre = middle of the bounding box of the leaf's bodies, radius =
// largest distance to the centre + largest core size (one warp per leaf) ------------------------
__global__ void tree_spheres_kernel(const double *__restrict__ P, int64_t ld, int osig,
                                    const int64_t *__restrict__ sidx, const int64_t *__restrict__ lbegin,
                                    const int64_t *__restrict__ lend, int64_t nl, double *__restrict__ ctr /*[3][nl]*/,
                                    double *__restrict__ rad, long long *bb) {
  const int64_t l = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (l >= nl) return;
  const int64_t b = lbegin[l], e = lend[l];
  double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY}, sg = -INFINITY;
  for (int64_t i = b + lane; i < e; i += 32) {
    const double *p = P + sidx[i] * ld;
#pragma unroll
    for (int c = 0; c < 3; ++c) { mn[c] = fmin(mn[c], p[c]); mx[c] = fmax(mx[c], p[c]); }
    sg = fmax(sg, p[osig]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      mn[c] = fmin(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmax(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
    sg = fmax(sg, __shfl_xor_sync(0xffffffffu, sg, o));
  }
  double c3[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) c3[c] = __dmul_rn(0.5, __dadd_rn(mn[c], mx[c]));
  double d2 = 0.0;
  for (int64_t i = b + lane; i < e; i += 32) {
    const double *p = P + sidx[i] * ld;
    const double dx = __dsub_rn(p[0], c3[0]), dy = __dsub_rn(p[1], c3[1]), dz = __dsub_rn(p[2], c3[2]);
    d2 = fmax(d2, __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d2 = fmax(d2, __shfl_xor_sync(0xffffffffu, d2, o));
  if (lane == 0) {
    const double r = __dadd_rn(__dsqrt_rn(d2), sg);
    ctr[l] = c3[0]; ctr[nl + l] = c3[1]; ctr[2 * nl + l] = c3[2];
    rad[l] = r;
    atomicMax(&bb[6], dbl_to_ordered(r));
  }
}

// ---- near-field list: one warp per target leaf walks the (2 reach + 1)^3 cell stencil in
// lexicographic order (= increasing source leaf id); PASS 0 counts, PASS 1 fills ----------------
template <int PASS>
__global__ void tree_list_kernel(TreeGrid g, int reach, const int64_t *__restrict__ lkey,
                                 const int32_t *__restrict__ cell_to_leaf, const double *__restrict__ ctr,
                                 const double *__restrict__ rad, int64_t nl, u64 *__restrict__ cnt /*[nl]*/,
                                 const u64 *__restrict__ ofs, int32_t *__restrict__ pair_t,
                                 int32_t *__restrict__ pair_s) {
  const int64_t l = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (l >= nl) return;
  const int64_t key = lkey[l];
  const int64_t cz = key % g.dims[2], cy = (key / g.dims[2]) % g.dims[1], cx = key / (g.dims[1] * g.dims[2]);
  const double lx = ctr[l], ly = ctr[nl + l], lz = ctr[2 * nl + l], lr = rad[l];
  const int w = 2 * reach + 1;
  const int64_t ncand = (int64_t)w * w * w;
  u64 out = PASS ? ofs[l] : 0ull;
  for (int64_t q0 = 0; q0 < ncand; q0 += 32) {
    const int64_t q = q0 + lane;
    bool near = false;
    int32_t m = -1;
    if (q < ncand) {
      const int64_t x = cx + q / ((int64_t)w * w) - reach, y = cy + (q / w) % w - reach, z = cz + q % w - reach;
      if (x >= 0 && x < g.dims[0] && y >= 0 && y < g.dims[1] && z >= 0 && z < g.dims[2]) {
        m = cell_to_leaf[(x * g.dims[1] + y) * g.dims[2] + z];
        if (m >= 0) {
          const double dx = __dsub_rn(lx, ctr[m]), dy = __dsub_rn(ly, ctr[nl + m]), dz = __dsub_rn(lz, ctr[2 * nl + m]);
          const double dist = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
          near = dist == 0.0 || __dadd_rn(lr, rad[m]) > __dmul_rn(g.theta, dist);
        }
      }
    }
    const unsigned ball = __ballot_sync(0xffffffffu, near);
    if (PASS) {
      if (near) {
        const u64 pos = out + __popc(ball & ((1u << lane) - 1));
        pair_t[pos] = (int32_t)l;
        pair_s[pos] = m;
      }
    }
    out += __popc(ball);
  }
  if (!PASS && lane == 0) cnt[l] = out;
}

// ---- gathers / scatters between particle order and tree-sorted order ----------------------------
// sorted 8-row source buffer [x y z sigma Gx Gy Gz sigma] and 16-row target buffer (positions,
// zeros) from the particle matrix view (rows X, Gamma, sigma)
__global__ void tree_gather_kernel(const double *__restrict__ P, int64_t ld, int ox, int og, int osig,
                                   const int64_t *__restrict__ sidx, int64_t n, double *__restrict__ src8,
                                   double *__restrict__ tgt16) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double *p = P + sidx[i] * ld;
  double *s = src8 + i * 8, *t = tgt16 + i * 16;
  s[0] = p[ox]; s[1] = p[ox + 1]; s[2] = p[ox + 2]; s[3] = p[osig];
  s[4] = p[og]; s[5] = p[og + 1]; s[6] = p[og + 2]; s[7] = p[osig];
  t[0] = p[ox]; t[1] = p[ox + 1]; t[2] = p[ox + 2];
#pragma unroll
  for (int k = 3; k < 16; ++k) t[k] = 0.0;
}
// out[sidx[i]] (+)= sorted near-field result; reset / static rule as uj_finish_kernel
__global__ void tree_scatter_kernel(const double *__restrict__ tgt16, const int64_t *__restrict__ sidx, int64_t n,
                                    double *__restrict__ out, int64_t ld, int urow, int jrow, int zrow0, int zrow1,
                                    int reset, const double *__restrict__ stat, int64_t sld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t c = sidx[i];
  const double *t = tgt16 + i * 16;
  double *o = out + c * ld;
  const bool is_static = stat != nullptr && stat[c * sld] != 0.0;
  const bool zero = reset && !is_static;
#pragma unroll
  for (int k = 0; k < 3; ++k) o[urow + k] = (zero ? 0.0 : o[urow + k]) + t[4 + k];
#pragma unroll
  for (int k = 0; k < 9; ++k) o[jrow + k] = (zero ? 0.0 : o[jrow + k]) + t[7 + k];
  if (zero) {
    if (zrow0 >= 0) { o[zrow0] = 0.0; o[zrow0 + 1] = 0.0; o[zrow0 + 2] = 0.0; }
    if (zrow1 >= 0) { o[zrow1] = 0.0; o[zrow1 + 1] = 0.0; o[zrow1 + 2] = 0.0; }
  }
}

// Order-independent 64-bit fingerprint of the rows the leaf lists depend on (X and sigma of every particle):
// the lists stay valid only while these do not change.  out[0] += mix(x, y, z, sigma, i) over all particles.
__global__ void tree_fingerprint_kernel(const double *__restrict__ P, int64_t ld, int ox, int osig, int64_t n,
                                        unsigned long long *out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long v = 0;
  if (i < n) {
    const double *p = P + i * ld;
    auto bits = [](double x) { return (unsigned long long)__double_as_longlong(x); };
    auto rot = [](unsigned long long x, int k) { return (x << k) | (x >> (64 - k)); };
    unsigned long long x = bits(p[ox]) ^ rot(bits(p[ox + 1]), 13) ^ rot(bits(p[ox + 2]), 29) ^ rot(bits(p[osig]), 47);
    x += 0x9e3779b97f4a7c15ull * (unsigned long long)(i + 1);
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    v = x ^ (x >> 31);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, v);
}

__global__ void tree_fill_i32_kernel(int32_t *p, int64_t n, int32_t v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace vpm
