// vpm_kernels_f32.cuh -- optional FP32-arithmetic form of the U/J pair sweep (sm_100a).
//
// Same reference arithmetic as vpm_kernels.cuh (src/FLOWVPM_fmm.jl:102-168,
// src/FLOWVPM_kernel.jl:44-84), evaluated on the FP32 FMA pipe for callers that accept the
// north-star's FP32 bar (1e-5 norm-wise instead of 1e-12).  What makes that bar reachable
// where a plain Float32 evaluation (the reference's own ParticleField{Float32}) is not:
//  * positions are carried as hi + lo pairs of floats, so that dx = (xt_h - xs_h) + (xt_l - xs_l)
//    has the relative accuracy of FP32 even when |x| >> |dx| (the benchmark cloud: |x| ~ 7,
//    neighbour distance ~ 0.01, where (float)x alone already carries 2e-5 of dx);
//  * the FP32 partial sums of one tile (128 sources) are flushed into FP64 sums kept in
//    shared memory, so the summation error does not grow with N;
//  * the kernel scalars use the same cancellation-free A/B forms as the FP64 sweep.
// Throughput: every FP32 instruction of the pair loop is a packed f32x2 operation
// (FFMA2 / FADD2 / FMUL2, new on sm_100) over the TWO targets a thread owns; the source
// operands are stored pre-duplicated in the record, so one LDS.128 yields two packed operands.
#pragma once
#include "vpm_kernels.cuh"

namespace vpm {

constexpr int kRecF = 32;  // floats per FP32 source record (128 B = 8 x LDS.128)

// ---- packed helpers ------------------------------------------------------------------
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float rsqrt_mufu(float a) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a));
  return y;
}
__device__ __forceinline__ float ex2_mufu(float a) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a));
  return y;
}
// 1/sqrt(b) for both lanes: MUFU.RSQ seed (2^-22) + one Newton step written without negations
//   e = b y0^2 - 1,  y = y0 - (y0/2) e
__device__ __forceinline__ float2 rsqrt2(float2 b) {
  float2 y0 = make_float2(rsqrt_mufu(b.x), rsqrt_mufu(b.y));
  float2 t = mul2(b, y0);
  float2 e = fma2(t, y0, f2(-1.0f));
  float2 hn = mul2(y0, f2(-0.5f));
  return fma2(hn, e, y0);
}

// FP32 record of source i (every value duplicated into both lanes of a float2):
//   [-xh -yh | -zh -xl | -yl -zl | G'x G'y | G'z -G'x | -G'y -G'z | q0 q1 | q2 q3]
// G' = -Gamma/(4 pi); (xh, xl) = hi/lo split of the FP64 position;
//   winckelmans: q0 = sigma^2, q1 = 1.5 sigma^2, q2 = -7.5 sigma^2
//   gaussianerf: q0 = 1/sigma^2, q1 = 1/sigma^3, q2 = r^2 beyond which g == 1, q3 = 2/sigma^5
//   gaussian:    q0 = 1/sigma^2, q1 = 1/sigma,   q2 = r^2 beyond which g == 1
constexpr float kFarU_gerf_f32 = 36.0f;   // s >= 6: 1 - g < 8e-8
constexpr float kFarU_gaus_f32 = 6.8f;    // s^3 >= 17.7: e^{-s^3} < 2e-8

__device__ __forceinline__ void split_hi_lo(double x, float &hi, float &lo) {
  hi = (float)x;
  lo = (float)(x - (double)hi);
}

__global__ void prep_uj_records_f32(SrcView src, int64_t s0, int64_t ns, int64_t ns_pad, int kernel,
                                    float *__restrict__ rec) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ns_pad) return;
  float2 *r = reinterpret_cast<float2 *>(rec + i * kRecF);
  if (i >= ns) {
#pragma unroll
    for (int k = 0; k < kRecF / 2; ++k) r[k] = make_float2(0.f, 0.f);
    // padding records are never read (tiles are cut at ns); keep q0 finite anyway
    return;
  }
  const double *p = src.p + (s0 + i) * src.ld;
  const double sigma = p[src.osig];
  const double isig = 1.0 / sigma;
  const double isig2 = isig * isig;
  const double isig3 = isig2 * isig;
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
  const float gx = (float)(-kConst4 * p[src.og]), gy = (float)(-kConst4 * p[src.og + 1]),
              gz = (float)(-kConst4 * p[src.og + 2]);
  r[0] = f2(xh);  r[1] = f2(yh);  r[2] = f2(zh);  r[3] = f2(xl);
  r[4] = f2(yl);  r[5] = f2(zl);  r[6] = f2(gx);  r[7] = f2(gy);
  r[8] = f2(gz);  r[9] = f2(-gx); r[10] = f2(-gy); r[11] = f2(-gz);
  r[12] = f2((float)q0); r[13] = f2((float)q1); r[14] = f2((float)q2); r[15] = f2((float)q3);
}

// ---- per-lane kernel scalars of the families that need a table / exp ------------------
// gaussianerf: G(u) = g/s^3 from a degree-5 table on intervals of width 1/2 in u (kGerfTableF,
// tools/gen_coeffs.py; 72 intervals up to u = 36), value + derivative from one Horner pass.
__device__ __forceinline__ void ab_gerf_tab_f32(float r2, float q0, float q1, float q3,
                                                const float *__restrict__ tab, float &A, float &B) {
  const float kMagic = 12582912.0f;  // 1.5 * 2^23: low bits of 2u + magic = rint(2u)
  float u = r2 * q0;
  float kd = fmaf(u, 2.0f, kMagic);
  int idx = __float_as_int(kd) & 0x3fffff;
  idx = min(idx, kGerfIntervalsF - 1);
  float x = fmaf(kd - kMagic, -0.5f, u);
  const float *c = tab + idx * kGerfCoeffsF;
  float p = c[5], dp;
  dp = p;               p = fmaf(p, x, c[4]);
  dp = fmaf(dp, x, p);  p = fmaf(p, x, c[3]);
  dp = fmaf(dp, x, p);  p = fmaf(p, x, c[2]);
  dp = fmaf(dp, x, p);  p = fmaf(p, x, c[1]);
  dp = fmaf(dp, x, p);  p = fmaf(p, x, c[0]);
  A = r2 == 0.0f ? 0.0f : q1 * p;
  B = q3 * dp;
}

// gaussian (src/FLOWVPM_kernel.jl:63-66) with t = s^3:  A = g/r^3,  B = 3 (e^-t - g/t) / (sigma^3 r^2)
// g = 1 - e^-t loses its leading digits for small t: series there.
__device__ __forceinline__ void ab_gaus_f32(float r2, float q1, float &A, float &B) {
  float rinv = rsqrt_mufu(r2);
  float r = r2 * rinv;
  float s = r * q1;
  float t = s * s * s;
  float E = ex2_mufu(-1.4426950408889634f * fminf(t, 80.0f));
  float g = t < 0.0625f ? t * fmaf(t, fmaf(t, fmaf(t, -1.0f / 24.0f, 1.0f / 6.0f), -0.5f), 1.0f) : 1.0f - E;
  float dg = 3.0f * s * s * E;
  float rinv2 = rinv * rinv;
  float rinv3 = rinv2 * rinv;
  float a = g * rinv3;
  float b = (dg * q1 * rinv - 3.0f * g * rinv2) * rinv3;
  const bool z = r2 == 0.0f;
  A = z ? 0.0f : a;
  B = z ? 0.0f : b;
}

__device__ __forceinline__ void ab_sing2(float2 r2, float2 &A, float2 &B) {
  float2 y = make_float2(rsqrt_mufu(r2.x), rsqrt_mufu(r2.y));  // raw seed + Newton below
  float2 t = mul2(r2, y);
  float2 e = fma2(t, y, f2(-1.0f));
  float2 hn = mul2(y, f2(-0.5f));
  y = fma2(hn, e, y);
  float2 y2 = mul2(y, y);
  float2 a = mul2(y2, y);
  float2 b = mul2(mul2(y2, f2(-3.0f)), a);
  A = make_float2(r2.x == 0.0f ? 0.0f : a.x, r2.y == 0.0f ? 0.0f : a.y);
  B = make_float2(r2.x == 0.0f ? 0.0f : b.x, r2.y == 0.0f ? 0.0f : b.y);
}

// One tile of n FP32 records against the two targets of this thread; acc[k] holds the k-th
// partial sum of target 0 in .x and of target 1 in .y (same 14 sums as uj_tile).
template <int K, int UNROLL>
__device__ __forceinline__ void uj_tile_f32(const float4 *__restrict__ tile, int n, float2 txh, float2 tyh,
                                            float2 tzh, float2 txl, float2 tyl, float2 tzl,
                                            float2 (&acc)[kAcc], int shortcut,
                                            const float *__restrict__ gtab) {
#pragma unroll UNROLL
  for (int j = 0; j < n; ++j) {
    const float4 v0 = tile[j * 8 + 0], v1 = tile[j * 8 + 1], v2 = tile[j * 8 + 2], v3 = tile[j * 8 + 3];
    const float4 v4 = tile[j * 8 + 4], v5 = tile[j * 8 + 5], v6 = tile[j * 8 + 6], v7 = tile[j * 8 + 7];
    const float2 sxh = make_float2(v0.x, v0.y), syh = make_float2(v0.z, v0.w);
    const float2 szh = make_float2(v1.x, v1.y), sxl = make_float2(v1.z, v1.w);
    const float2 syl = make_float2(v2.x, v2.y), szl = make_float2(v2.z, v2.w);
    const float2 gx = make_float2(v3.x, v3.y), gy = make_float2(v3.z, v3.w);
    const float2 gz = make_float2(v4.x, v4.y), ngx = make_float2(v4.z, v4.w);
    const float2 ngy = make_float2(v5.x, v5.y), ngz = make_float2(v5.z, v5.w);
    const float2 q0 = make_float2(v6.x, v6.y), q1 = make_float2(v6.z, v6.w);
    const float2 q2 = make_float2(v7.x, v7.y), q3 = make_float2(v7.z, v7.w);

    // dx = (xt_h - xs_h) + (xt_l - xs_l): the record holds the negated source parts
    const float2 dx = add2(add2(txh, sxh), add2(txl, sxl));
    const float2 dy = add2(add2(tyh, syh), add2(tyl, syl));
    const float2 dz = add2(add2(tzh, szh), add2(tzl, szl));
    float2 A, B;
    if constexpr (K == K_WINCK) {
      const float2 b = fma2(dz, dz, fma2(dy, dy, fma2(dx, dx, q0)));
      const float2 y = rsqrt2(b);
      const float2 y2 = mul2(y, y);
      const float2 y3 = mul2(y2, y);
      const float2 y5 = mul2(y3, y2);
      const float2 y7 = mul2(y5, y2);
      A = fma2(q1, y5, y3);
      B = fma2(q2, y7, mul2(y5, f2(-3.0f)));
      // the reference skips r2 == 0 (src/FLOWVPM_fmm.jl:118): only the W term needs the mask
      const bool z0 = dx.x == 0.0f && dy.x == 0.0f && dz.x == 0.0f;
      const bool z1 = dx.y == 0.0f && dy.y == 0.0f && dz.y == 0.0f;
      A = make_float2(z0 ? 0.0f : A.x, z1 ? 0.0f : A.y);
    } else {
      const float2 r2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
      if constexpr (K == K_SING) {
        ab_sing2(r2, A, B);
      } else {
        const bool n0 = r2.x <= q2.x, n1 = r2.y <= q2.y;
        if (__any_sync(0xffffffffu, !shortcut || n0 || n1)) {
          float a0, b0, a1, b1;
          if constexpr (K == K_GERF) {
            ab_gerf_tab_f32(r2.x, q0.x, q1.x, q3.x, gtab, a0, b0);
            ab_gerf_tab_f32(r2.y, q0.x, q1.x, q3.x, gtab, a1, b1);
            // lanes beyond the table take the singular values (g == 1 to 8e-8 there)
            if (__any_sync(0xffffffffu, !(n0 && n1))) {
              float2 As, Bs;
              ab_sing2(r2, As, Bs);
              const bool f0 = r2.x * q0.x >= kFarU_gerf_f32, f1 = r2.y * q0.x >= kFarU_gerf_f32;
              a0 = f0 ? As.x : a0; b0 = f0 ? Bs.x : b0;
              a1 = f1 ? As.y : a1; b1 = f1 ? Bs.y : b1;
            }
          } else {
            ab_gaus_f32(r2.x, q1.x, a0, b0);
            ab_gaus_f32(r2.y, q1.x, a1, b1);
          }
          A = make_float2(a0, a1);
          B = make_float2(b0, b1);
        } else {
          ab_sing2(r2, A, B);
        }
      }
    }
    // c = dx x G'
    const float2 cx = fma2(dy, gz, mul2(dz, ngy));
    const float2 cy = fma2(dz, gx, mul2(dx, ngz));
    const float2 cz = fma2(dx, gy, mul2(dy, ngx));
    acc[0] = fma2(A, cx, acc[0]);
    acc[1] = fma2(A, cy, acc[1]);
    acc[2] = fma2(A, cz, acc[2]);
    acc[11] = fma2(A, gx, acc[11]);
    acc[12] = fma2(A, gy, acc[12]);
    acc[13] = fma2(A, gz, acc[13]);
    const float2 bx = mul2(B, cx), by = mul2(B, cy), bz = mul2(B, cz);
    acc[3] = fma2(bx, dx, acc[3]);
    acc[4] = fma2(by, dx, acc[4]);
    acc[5] = fma2(bz, dx, acc[5]);
    acc[6] = fma2(bx, dy, acc[6]);
    acc[7] = fma2(by, dy, acc[7]);
    acc[8] = fma2(bz, dy, acc[8]);
    acc[9] = fma2(bx, dz, acc[9]);
    acc[10] = fma2(by, dz, acc[10]);
  }
}

struct UjArgsF {
  const double *tpos;  // target positions (FP64): tpos[i*tld + 0..2]
  int64_t tld;
  int64_t nt;
  const float *rec;    // FP32 source records [ns_pad][kRecF]
  int64_t ns;
  int tiles_per_split;
  double *partial;     // [nsplit][kAcc][pstride]  (FP64, same layout as the FP64 sweep)
  int64_t pstride;
  int shortcut;
};

template <int K, int UNROLL>
__global__ void __launch_bounds__(kThreads, 4) uj_pairs_kernel_f32(const UjArgsF a) {
  __shared__ __align__(128) float tiles[kStages * kTile * kRecF];  // 32 KB
  __shared__ __align__(8) uint64_t full[kStages];
  __shared__ __align__(16) float gtab[K == K_GERF ? kGerfIntervalsF * kGerfCoeffsF : 1];
  const int tid = threadIdx.x;
  if constexpr (K == K_GERF) {
    for (int i = tid; i < kGerfIntervalsF * kGerfCoeffsF; i += kThreads) gtab[i] = kGerfTableF[i];
  }
  const int64_t tbase = (int64_t)blockIdx.x * (kThreads * 2);

  float th[3][2], tl[3][2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    int64_t i = tbase + (int64_t)t * kThreads + tid;
    if (i >= a.nt) i = a.nt - 1;
    const double *p = a.tpos + i * a.tld;
#pragma unroll
    for (int c = 0; c < 3; ++c) split_hi_lo(p[c], th[c][t], tl[c][t]);
  }
  const float2 txh = make_float2(th[0][0], th[0][1]), tyh = make_float2(th[1][0], th[1][1]),
               tzh = make_float2(th[2][0], th[2][1]);
  const float2 txl = make_float2(tl[0][0], tl[0][1]), tyl = make_float2(tl[1][0], tl[1][1]),
               tzl = make_float2(tl[2][0], tl[2][1]);
  double dsum[2][kAcc];  // FP64 sums of the two targets; the FP32 sums of each tile are flushed here
#pragma unroll
  for (int k = 0; k < kAcc; ++k) dsum[0][k] = dsum[1][k] = 0.0;

  const int64_t ntiles = (a.ns + kTile - 1) / kTile;
  const int64_t tile0 = (int64_t)blockIdx.y * a.tiles_per_split;
  int64_t tile1 = tile0 + a.tiles_per_split;
  if (tile1 > ntiles) tile1 = ntiles;
  const int ntl = tile1 > tile0 ? (int)(tile1 - tile0) : 0;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  auto issue = [&](int it) {
    const int64_t first = (tile0 + it) * kTile;
    const int n = (int)((a.ns - first) < kTile ? (a.ns - first) : kTile);
    const uint32_t bytes = (uint32_t)n * kRecF * sizeof(float);
    const int st = it % kStages;
    mbar_expect_tx(&full[st], bytes);
    tma_bulk_g2s(tiles + (size_t)st * kTile * kRecF, a.rec + first * kRecF, bytes, &full[st]);
  };
  if (tid == 0) {
    for (int s = 0; s < kStages && s < ntl; ++s) issue(s);
  }

  for (int it = 0; it < ntl; ++it) {
    const int st = it % kStages;
    mbar_wait(&full[st], (uint32_t)((it / kStages) & 1));
    const int64_t first = (tile0 + it) * kTile;
    const int n = (int)((a.ns - first) < kTile ? (a.ns - first) : kTile);
    float2 acc[kAcc];
#pragma unroll
    for (int k = 0; k < kAcc; ++k) acc[k] = make_float2(0.f, 0.f);
    uj_tile_f32<K, UNROLL>(reinterpret_cast<const float4 *>(tiles + (size_t)st * kTile * kRecF), n, txh, tyh,
                           tzh, txl, tyl, tzl, acc, a.shortcut, gtab);
#pragma unroll
    for (int k = 0; k < kAcc; ++k) {
      dsum[0][k] += (double)acc[k].x;
      dsum[1][k] += (double)acc[k].y;
    }
    __syncthreads();  // everyone is done reading stage st
    if (tid == 0 && it + kStages < ntl) issue(it + kStages);
  }

#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int64_t i = tbase + (int64_t)t * kThreads + tid;
    if (i < a.nt) {
      double *o = a.partial + (int64_t)blockIdx.y * kAcc * a.pstride + i;
#pragma unroll
      for (int k = 0; k < kAcc; ++k) o[(int64_t)k * a.pstride] = dsum[t][k];
    }
  }
}

// FFMA / FFMA2 issue-rate probe (roofline denominator of the FP32 mode): 8 independent chains,
// packed or scalar, with per-chain (3 distinct registers) or loop-invariant multiplicands.
template <int MODE>
__global__ void ffma_peak_kernel(float *out, int iters, float seed) {
  float2 a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = make_float2(seed + i, seed - i);
    b[i] = make_float2(0.999f - 1e-4f * i + 1e-6f * threadIdx.x, 0.998f + 1e-4f * i);
  }
  const float2 m = make_float2(0.999999f, 0.999998f), c = make_float2(1e-9f, 2e-9f);
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) a[i] = fma2(a[i], m, c);                       // FFMA2, invariant operands
        if (MODE == 1) a[i] = fma2(b[i], b[(i + 1) & 7], a[i]);       // FFMA2, 3 distinct registers
        if (MODE == 2) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }  // FFMA
        if (MODE == 3) { a[i].x = fmaf(b[i].x, b[(i + 1) & 7].x, a[i].x); a[i].y = fmaf(b[i].y, b[(i + 1) & 7].y, a[i].y); }
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y + b[i].x;
  if (s == 12345.678f) out[0] = s;
}

}  // namespace vpm
