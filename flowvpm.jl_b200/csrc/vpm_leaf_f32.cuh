// vpm_leaf_f32.cuh -- FP32-arithmetic form of the near-field U/J leaf kernel (option
// VPM_OPT_NEARFIELD_FP32).  [r2-prep: written without a GPU at hand, NOT validated yet]
//
// Why: the near field of UJ_fmm is added to a far field that FastMultipole truncates at
// 1e-3..1e-6, so FP64 pair arithmetic buys nothing there.  Same numerics as the FP32 mode of
// the all-pairs sweep (vpm_kernels_f32.cuh): hi/lo split positions, the cancellation-free A/B
// forms, FP32 sums per tile flushed into FP64 sums.  One target per lane (leaves are small), so
// the instructions are scalar FFMA, not the packed f32x2 form; records are 16 floats.
#pragma once
#include "vpm_kernels_f32.cuh"
#include "vpm_leaf.cuh"

namespace vpm {

constexpr int kRecFS = 16;  // floats per scalar FP32 record (64 B = 4 x LDS.128)

// [-xh -yh -zh -xl | -yl -zl G'x G'y | G'z q0 q1 q2 | q3 0 0 0], q as in prep_uj_records_f32
__global__ void prep_uj_records_f32s(SrcView src, int64_t s0, int64_t ns, int64_t ns_pad, int kernel,
                                     float *__restrict__ rec) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ns_pad) return;
  float4 *r = reinterpret_cast<float4 *>(rec + i * kRecFS);
  if (i >= ns) {
    r[0] = r[1] = r[2] = r[3] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const double *p = src.p + (s0 + i) * src.ld;
  const double sigma = p[src.osig];
  const double isig = 1.0 / sigma, isig2 = isig * isig, isig3 = isig2 * isig;
  double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
  if (kernel == K_WINCK) {
    q0 = sigma * sigma; q1 = 1.5 * q0; q2 = -7.5 * q0;
  } else if (kernel == K_GERF) {
    q0 = isig2; q1 = isig3; q2 = (double)kFarU_gerf_f32 * (sigma * sigma); q3 = 2.0 * isig3 * isig2;
  } else if (kernel == K_GAUS) {
    q0 = isig2; q1 = isig; q2 = (double)kFarU_gaus_f32 * (sigma * sigma);
  }
  float xh, xl, yh, yl, zh, zl;
  split_hi_lo(-p[src.ox], xh, xl);
  split_hi_lo(-p[src.ox + 1], yh, yl);
  split_hi_lo(-p[src.ox + 2], zh, zl);
  r[0] = make_float4(xh, yh, zh, xl);
  r[1] = make_float4(yl, zl, (float)(-kConst4 * p[src.og]), (float)(-kConst4 * p[src.og + 1]));
  r[2] = make_float4((float)(-kConst4 * p[src.og + 2]), (float)q0, (float)q1, (float)q2);
  r[3] = make_float4((float)q3, 0.f, 0.f, 0.f);
}

__device__ __forceinline__ void ab_sing_f32(float r2, float &A, float &B) {
  float y = rsqrt_mufu(r2);
  const float e = fmaf(r2 * y, y, -1.0f);
  y = fmaf(-0.5f * y, e, y);
  const float y2 = y * y, a = y2 * y, b = (-3.0f * y2) * a;
  const bool z = r2 == 0.0f;
  A = z ? 0.0f : a;
  B = z ? 0.0f : b;
}

// one tile of n scalar FP32 records against the target of this lane (SPLIT as in uj_tile)
template <int K, bool SPLIT>
__device__ __forceinline__ void uj_tile_f32s(const float4 *__restrict__ tile, int n, const float (&th)[3],
                                             const float (&tl)[3], float (&acc)[kAcc], int shortcut,
                                             const float *__restrict__ gtab, int nsplit, int phase) {
  const int trips = SPLIT ? (n + nsplit - 1) / nsplit : n;
#pragma unroll 2
  for (int jj = 0; jj < trips; ++jj) {
    int j = jj;
    bool live = true;
    if constexpr (SPLIT) {
      j = jj * nsplit + phase;
      live = j < n;
      j = live ? j : n - 1;
    }
    const float4 v0 = tile[j * 4 + 0], v1 = tile[j * 4 + 1], v2 = tile[j * 4 + 2], v3 = tile[j * 4 + 3];
    const float dx = (th[0] + v0.x) + (tl[0] + v0.w);
    const float dy = (th[1] + v0.y) + (tl[1] + v1.x);
    const float dz = (th[2] + v0.z) + (tl[2] + v1.y);
    const float gx = v1.z, gy = v1.w, gz = v2.x, q0 = v2.y, q1 = v2.z, q2 = v2.w, q3 = v3.x;
    float A, B;
    if constexpr (K == K_WINCK) {
      const float b = fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, q0)));
      float y = rsqrt_mufu(b);
      const float e = fmaf(b * y, y, -1.0f);
      y = fmaf(-0.5f * y, e, y);
      const float y2 = y * y, y3 = y2 * y, y5 = y3 * y2;
      A = fmaf(q1, y5, y3);
      B = fmaf(q2, y2, -3.0f) * y5;
      if (dx == 0.0f && dy == 0.0f && dz == 0.0f) A = 0.0f;  // src/FLOWVPM_fmm.jl:118: only W needs it
    } else {
      const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
      if constexpr (K == K_SING) {
        ab_sing_f32(r2, A, B);
      } else {
        const bool nearlane = r2 <= q2;
        if (__any_sync(0xffffffffu, !shortcut || nearlane)) {
          if constexpr (K == K_GERF) {
            ab_gerf_tab_f32(r2, q0, q1, q3, gtab, A, B);
            if (r2 * q0 >= kFarU_gerf_f32) ab_sing_f32(r2, A, B);  // beyond the table: g == 1 to 8e-8
          } else {
            ab_gaus_f32(r2, q1, A, B);
          }
        } else {
          ab_sing_f32(r2, A, B);
        }
      }
    }
    if constexpr (SPLIT) {
      A = live ? A : 0.0f;
      B = live ? B : 0.0f;
    }
    const float cx = fmaf(dy, gz, -(dz * gy));
    const float cy = fmaf(dz, gx, -(dx * gz));
    const float cz = fmaf(dx, gy, -(dy * gx));
    acc[0] = fmaf(A, cx, acc[0]);
    acc[1] = fmaf(A, cy, acc[1]);
    acc[2] = fmaf(A, cz, acc[2]);
    acc[11] = fmaf(A, gx, acc[11]);
    acc[12] = fmaf(A, gy, acc[12]);
    acc[13] = fmaf(A, gz, acc[13]);
    const float bx = B * cx, by = B * cy, bz = B * cz;
    acc[3] = fmaf(bx, dx, acc[3]);
    acc[4] = fmaf(by, dx, acc[4]);
    acc[5] = fmaf(bz, dx, acc[5]);
    acc[6] = fmaf(bx, dy, acc[6]);
    acc[7] = fmaf(by, dy, acc[7]);
    acc[8] = fmaf(bz, dy, acc[8]);
    acc[9] = fmaf(bx, dz, acc[9]);
    acc[10] = fmaf(by, dz, acc[10]);
  }
}

struct LeafUjArgsF {
  LeafCsr csr;
  const double *tpos;  // sorted target buffer (FP64): tpos[i*tld + 0..2]
  int64_t tld;
  const float *rec;    // scalar FP32 records of the sorted source buffer
  double *out;         // sorted target buffer, ld = tld
  int urow, jrow;
  int want_U, want_J;
  int shortcut;
};

template <int K, int NT, int TILE>
__global__ void __launch_bounds__(NT) uj_leaf_kernel_f32(const LeafUjArgsF a) {
  __shared__ __align__(128) float tiles[kStages][TILE * kRecFS];
  __shared__ __align__(8) uint64_t full[kStages];
  __shared__ int tile_n[kStages];  // records in each stage's tile, published by the producer
  __shared__ __align__(16) float gtab[K == K_GERF ? kGerfIntervalsF * kGerfCoeffsF : 1];
  const int tid = threadIdx.x;
  if constexpr (K == K_GERF) {
    for (int i = tid; i < kGerfIntervalsF * kGerfCoeffsF; i += NT) gtab[i] = kGerfTableF[i];
  }
  const int leaf = a.csr.wi_leaf[blockIdx.x];
  const int64_t tb = a.csr.tleaf_begin[leaf] + a.csr.wi_off[blockIdx.x];
  int64_t te = a.csr.tleaf_end[leaf];
  if (te > tb + NT) te = tb + NT;
  const int wbase = (tid >> 5) << 5, lane = tid & 31;
  const int64_t wlive = te - (tb + wbase);
  const int nsplit = wlive > 16 ? 1 : wlive > 8 ? 2 : wlive > 4 ? 4 : 8;
  const int glanes = 32 / nsplit, phase = lane / glanes;
  const int64_t i = tb + wbase + (lane % glanes);
  const bool valid = i < te;
  const double *p = a.tpos + (valid ? i : te - 1) * a.tld;
  float th[3], tl[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) split_hi_lo(p[c], th[c], tl[c]);
  double dsum[kAcc];
#pragma unroll
  for (int k = 0; k < kAcc; ++k) dsum[k] = 0.0;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  // 64-byte records = 8 doubles for the (byte-counting) producer of vpm_leaf.cuh
  constexpr int kRecD = kRecFS * (int)sizeof(float) / (int)sizeof(double);
  const double *recd = reinterpret_cast<const double *>(a.rec);
  LeafTileProducer prod;  // state lives (uniformly) in the lanes of warp 0
  prod.init(a.csr, leaf);
  if (tid < 32)
    for (int s = 0; s < kStages; ++s)
      prod.issue<TILE, kRecD>(a.csr, recd, reinterpret_cast<double *>(&tiles[s][0]), &full[s], &tile_n[s]);

  for (int it = 0;; ++it) {
    const int st = it % kStages;
    mbar_wait(&full[st], (uint32_t)((it / kStages) & 1));
    const int n = tile_n[st];
    if (n == 0) break;  // list exhausted (the same value for every thread of the CTA)
    const float4 *tile = reinterpret_cast<const float4 *>(&tiles[st][0]);
    float acc[kAcc];
#pragma unroll
    for (int k = 0; k < kAcc; ++k) acc[k] = 0.0f;
    if (nsplit == 1) uj_tile_f32s<K, false>(tile, n, th, tl, acc, a.shortcut, gtab, 1, 0);
    else uj_tile_f32s<K, true>(tile, n, th, tl, acc, a.shortcut, gtab, nsplit, phase);
    // the tile's FP32 sums go into FP64 sums: the summation error does not grow with the list
#pragma unroll
    for (int k = 0; k < kAcc; ++k) dsum[k] += (double)acc[k];
    __syncthreads();
    if (tid < 32)
      prod.issue<TILE, kRecD>(a.csr, recd, reinterpret_cast<double *>(&tiles[st][0]), &full[st], &tile_n[st]);
  }
  for (int o = glanes; o < 32; o <<= 1) {
#pragma unroll
    for (int k = 0; k < kAcc; ++k) dsum[k] += __shfl_xor_sync(0xffffffffu, dsum[k], o);
  }

  if (valid && phase == 0) {
    double U[3], J[9];
    finish_sums(dsum, U, J);
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

}  // namespace vpm
