// vpm_kernels_tab.cuh -- U/J pair sweep of the gaussianerf and gaussian families with the
// regularising functions read from a log-spaced, bank-replicated shared-memory table.
//
// Reference arithmetic: src/FLOWVPM_fmm.jl:102-168 (pair loop), src/FLOWVPM_kernel.jl:51-66
// (g, dg of gaussianerf and gaussian).  Per pair the loop needs (vpm_kernels.cuh)
//     A = g/r^3 = G/sigma^3,                G = g(s)/s^3
//     B = (dg/(sigma r) - 3 g/r^2)/r^3 = (1/s) dG/ds / sigma^5 = 2 dG/du / sigma^5,  u = s^2,
// i.e. ONE smooth function and its derivative.  Round 1 read G from a degree-9 table with five
// per-lane 16-byte gathers from a single shared-memory copy: on fields where most pairs are
// inside the regularised range the lanes of a warp hit different rows, the gathers bank-conflict
// (ncu: shared-memory wavefronts 94 % of peak, FP64 pipe 67 % active, 205 G pairs/s = 0.44 of the
// FP64 roofline -- profiles/r2_gerf_dense_before_ncu_summary.txt).  This version
//   * tabulates G(u), u = s^2, on intervals uniform in the BITS of t = u + c (c = 4 for gaussianerf; c = 0 and
//     rows from u = 1/4 for gaussian, whose G is not analytic at u = 0; 2^LOGN intervals per octave of t): the row index and the in-interval coordinate
//     xi in [-1/2, 1/2) come from the exponent/mantissa fields of t on the integer pipe -- no
//     FP64 magic-number rounding, and the relative interval width follows the function's scale;
//   * evaluates value and derivative with one joint Horner pass of degree 7 whose three highest
//     terms run on the otherwise idle FP32 pipe (they are < 2^-33 of the value): 9 DFMA instead
//     of 17, rows of 48 bytes (tools/gen_tab_coeffs.py: G 1.9e-16, dG 5e-15 against mpmath);
//   * keeps EIGHT copies of the table in shared memory, copy k living entirely in the 16-byte
//     bank group k: lane l reads copy l & 7, so the eight lanes of every quarter-warp phase of an
//     LDS.128 hit eight different bank groups whatever rows they ask for -- conflict-free by
//     construction, 3 x 4 wavefronts per pair-warp instead of ~50;
//   * runs as ONE 512-thread CTA per SM (two targets per thread) so that the 106 KB table is
//     shared by 16 warps;
//   * appends power-law rows for the far field (g == 1), so that a warp whose lanes straddle the
//     cut-off runs ONE code path instead of the table AND a rsqrt-based evaluation.
#pragma once
#include "vpm_kernels.cuh"
#include "vpm_tab_coeffs.cuh"

namespace vpm {

constexpr int kTabThreads = 512;  // default CTA size (VPM_OPT_UJ_VARIANT 31/32 take 384)
constexpr int kTabT = 2;
constexpr int kTabRowChunks = 3;                       // 16-byte chunks per row
constexpr int kTabRowBytes = kTabRowChunks * 8 * 16;   // one row of all eight copies: 384 B

template <int K> struct TabOf;
template <> struct TabOf<K_GERF> {
  static constexpr int offset = kTabGerfOffset, emin = kTabGerfEmin, logn = kTabGerfLogN, near = kTabGerfRows,
                       far = kTabGerfFarRows, rows = near + far;
  static __device__ __forceinline__ const uint64_t *words() { return kTabGerf; }
};
template <> struct TabOf<K_GAUS> {
  static constexpr int offset = kTabGausOffset, emin = kTabGausEmin, logn = kTabGausLogN, near = kTabGausRows,
                       far = kTabGausFarRows, rows = near + far;
  static __device__ __forceinline__ const uint64_t *words() { return kTabGaus; }
};

template <int K>
constexpr size_t tab_smem_bytes() {
  return (size_t)TabOf<K>::rows * kTabRowBytes + (size_t)kStages * kTile * kRec * sizeof(double) + 64;
}

// cooperative load: global packed rows (6 words) -> 8 bank-group-private copies
template <int K>
__device__ __forceinline__ void load_tab(unsigned char *tab) {
  const uint64_t *w = TabOf<K>::words();
  const int total = TabOf<K>::rows * kTabRowChunks * 8;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int copy = i & 7, rc = i >> 3;  // rc = row * 3 + chunk
    ulonglong2 v;
    v.x = w[2 * rc];
    v.y = w[2 * rc + 1];
    reinterpret_cast<ulonglong2 *>(tab)[rc * 8 + copy] = v;
  }
}

// 16-byte shared-memory load at a 32-bit shared address + constant offset (the table base of a
// thread is computed once per kernel; a generic pointer would be re-derived in every iteration)
template <int OFF>
__device__ __forceinline__ double2 lds_v2(uint32_t addr) {
  double2 v;
  asm("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(addr), "n"(OFF));
  return v;
}

// Where a thread reads the three 16-byte chunks of a table row from.
struct TabShared {  // the bank-group-private copies: base = shared address of the table + 16 * (lane & 7)
  uint32_t base;
  __device__ __forceinline__ void load(unsigned row, double2 &c01, double2 &c23, double2 &c4t) const {
    const uint32_t ra = base + row * kTabRowBytes;
    c4t = lds_v2<256>(ra); c23 = lds_v2<128>(ra); c01 = lds_v2<0>(ra);
  }
};
struct TabGlobal {  // the packed 48-byte rows in global memory, through L1: the leaf-list kernels, whose CTAs
  const double2 *rows;  // (one to four warps, a few thousand pairs each) cannot amortise staging eight copies
  __device__ __forceinline__ void load(unsigned row, double2 &c01, double2 &c23, double2 &c4t) const {
    const double2 *r = rows + row * kTabRowChunks;
    c4t = __ldg(r + 2); c23 = __ldg(r + 1); c01 = __ldg(r);
  }
};
template <int K>
__device__ __forceinline__ TabGlobal tab_global() {
  return TabGlobal{reinterpret_cast<const double2 *>(TabOf<K>::words())};
}

// Records (prep_uj_records_tab), the same for both families:  [x y z q0 | G'x G'y G'z q1 | q2 q3],
//   q0 = 1/sigma^2 (u = r^2 q0), q1 = 1/sigma^3 (A = q1 G), q2 = far cut-off in r^2, q3 = 2/sigma^5 (B = q3 dG/du)
__global__ void prep_uj_records_tab(SrcView src, int64_t s0, int64_t ns, int64_t ns_pad, int kernel,
                                    double *__restrict__ rec) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ns_pad) return;
  double *r = rec + i * kRec;
  if (i >= ns) {
    // padding (never read by the sweep): zero strength, q's that keep every intermediate finite
#pragma unroll
    for (int k = 0; k < kRec; ++k) r[k] = 0.0;
    r[3] = 1.0; r[7] = 1.0; r[8] = 1.0; r[9] = 1.0;
    return;
  }
  const double *p = src.p + (s0 + i) * src.ld;
  const double sigma = p[src.osig];
  const double isig = 1.0 / sigma;
  const double isig2 = isig * isig;
  const double isig3 = isig2 * isig;
  r[0] = p[src.ox]; r[1] = p[src.ox + 1]; r[2] = p[src.ox + 2]; r[3] = isig2;
  r[4] = -kConst4 * p[src.og]; r[5] = -kConst4 * p[src.og + 1]; r[6] = -kConst4 * p[src.og + 2];
  r[7] = isig3;
  r[8] = (kernel == K_GERF ? kFarU_gerf : kFarU_gaus) * (sigma * sigma);
  r[9] = 2.0 * isig3 * isig2;
}

// Value p and xi-derivative dp of the polynomial of row `row` at the in-interval coordinate
// given by the low (52 - LOGN) mantissa bits of t.  tab: TabShared or TabGlobal.
template <int LOGN, class TAB>
__device__ __forceinline__ void tab_poly(int hi, unsigned lo, unsigned row, const TAB &tab,
                                         double &p, double &dp) {
  constexpr int MB = 20 - LOGN;  // mantissa bits of the high word below the row index
  const unsigned mh = (unsigned)hi & ((1u << MB) - 1u);
  // xi = (low 52-LOGN mantissa bits) / 2^(52-LOGN) - 1/2, exact in FP64; its top 23 bits in FP32
  const double xi = __hiloint2double((int)(0x3ff00000u | (mh << LOGN) | (lo >> (32 - LOGN))), (int)(lo << LOGN)) - 1.5;
  const float xf = __uint_as_float(0x3f800000u | (mh << (3 + LOGN)) | (lo >> (29 - LOGN))) - 1.5f;
  double2 c01, c23, c4t;
  tab.load(row, c01, c23, c4t);
  const float t5 = __int_as_float(__double2loint(c4t.y)), t6 = __int_as_float(__double2hiint(c4t.y));
  const float t7 = __int_as_float(__double2loint(c4t.x) << 12);
  // Horner with derivative: b_j = c_j + xi b_{j+1},  d_j = b_{j+1} + xi d_{j+1}
  const float b6 = fmaf(xf, t7, t6);
  const float b5 = fmaf(xf, b6, t5);
  const float d5 = fmaf(xf, t7, b6);
  const float d4 = fmaf(xf, d5, b5);
  // (float -> double as F2F.F64.F32: five integer instructions doing the same were measured slower)
  double b = fma(xi, (double)b5, c4t.x);
  double d = fma(xi, (double)d4, b);
  b = fma(xi, b, c23.y); d = fma(xi, d, b);
  b = fma(xi, b, c23.x); d = fma(xi, d, b);
  b = fma(xi, b, c01.y); d = fma(xi, d, b);
  b = fma(xi, b, c01.x);
  p = b;
  dp = d;
}

// A = G/sigma^3 = q1 G and B = 2 dG/du / sigma^5 = q3 dG/du of one pair from the table of its family;
// `far` = the pair is beyond the regularised range (g == 1, dg == 0); `t` = u + offset for a near lane,
// u for a far lane (u = r^2 / sigma^2); no square root anywhere.
//   table rows [0, 128): the power law u^-3/2 (far lanes): row = exponent parity bit and the top mantissa bits
//     of t; the rest of the power of two is an exponent shift of q1 and q3;
//   rows [128, 128 + near): regularised range: row from exponent + top LOGN mantissa bits of t.
// G = p 2^ka, dG/du = dp 2^(LOGN - e + ka) with t = m 2^e; ka = 0 (near), -3 (e >> 1) (far).
// All of that is integer arithmetic on the high words (E = biased exponent field of t, in place):
//   kq = ka << 20 + C,  C = (1023 + LOGN) << 20;   hi(q1) += kq - C;   hi(q3) += kq - (E << 20).
// r2 == 0 (the reference skips those pairs, src/FLOWVPM_fmm.jl:118; c = dx x G' = 0 there, so only the
// W sums need A = 0): A's scale factor gets a zero high word, i.e. A ~ 1e-320 (an exact 0 to any
// tolerance); decided on the high word of r2 alone, so pairs closer than ~1e-154 count as coincident.
template <int K, class TAB>
__device__ __forceinline__ void ab_tab(double t, double r2, bool far, double q1, double q3,
                                       const TAB &tab, double &A, double &B) {
  using TB = TabOf<K>;
  constexpr int LOGN = TB::logn, MB = 20 - LOGN;
  static_assert(TB::far == 128, "far rows are indexed with a 7-bit mask");
  constexpr unsigned C = (unsigned)(1023 + LOGN) << 20;
  const bool z = __double2hiint(r2) == 0;
  const int hi = __double2hiint(t);
  const unsigned lo = (unsigned)__double2loint(t);
  const int top = hi >> MB;  // biased exponent and the top LOGN mantissa bits
  // t >= 2^EMIN for every finite near lane the table serves, so the difference is >= 0; the unsigned compare
  // also sends NaN (CUDA's canonical NaN has the sign bit set), anything past the range and the gaussian
  // family's u < 2^EMIN lanes (handled by the caller's rare path) to the last near row
  const unsigned row_near = min((unsigned)(top - ((1023 + TB::emin) << LOGN) + TB::far), (unsigned)(TB::rows - 1));
  const unsigned row = far ? (unsigned)(top & (TB::far - 1)) : row_near;
  double p, dp;
  tab_poly<LOGN>(hi, lo, row, tab, p, dp);
  const unsigned em = (unsigned)hi & 0xfff00000u;  // E << 20
  unsigned kq = ((unsigned)(hi + 0x100000) >> 21) * (unsigned)(-3 << 20) + ((1536u << 20) + C);  // -3 ((E - 1023) >> 1)
  kq = far ? kq : C;
  const unsigned hA = (unsigned)__double2hiint(q1) + kq - C;
  A = __hiloint2double((int)(z ? 0u : hA), __double2loint(q1)) * p;
  B = __hiloint2double((int)((unsigned)__double2hiint(q3) + kq - em), __double2loint(q3)) * dp;
}

// The gaussian family off its table (some lane of the warp has a pair closer than s = 1/2, the self pair
// included), for every lane of that warp-iteration.  With v = s^3:  G = g/s^3 = phi(v),  phi = (1 - e^-v)/v,  and
// 2 dG/du = 3 (e^-v v - (1 - e^-v))/s^5 = 3 s psi(v),  psi = (e^-v v - 1 + e^-v)/v^2.  For v >= 1/32 these are
// evaluated through exp as the reference does (src/FLOWVPM_kernel.jl:63-66; its 1 - e^-v loses at most five bits
// there); below, where the reference's form cancels (g == 0 for s < 1e-5), from their Taylor series
//   phi = sum_k (-v)^k / (k+1)!,   psi = sum_k (-1)^(k+1) (k+1)/(k+2)! v^k   (nine terms: < 1e-20 at v = 1/32).
// Rare: one pair per target and sweep plus physically overlapping neighbours.
__device__ __forceinline__ void ab_gaus_u(double r2, bool far, double q0, double q1, double q3, double &A, double &B) {
  const bool z = __double2hiint(r2) == 0;
  const double u0 = r2 * q0;
  const double u = __hiloint2double(z ? 0x3ff00000 : __double2hiint(u0), __double2loint(u0));  // 0 -> 1: finite
  const double y = rsqrt_fp64(u);      // 1/s
  const double s = u * y;
  const double v = u * s;
  const double y2 = y * y, y3 = y2 * y, y5 = y3 * y2;
  double G, H;                         // G and 2 dG/du
  if (far) {
    G = y3; H = -3.0 * y5;
  } else if (v < 0.03125) {
    double ph = 1.0 / 362880.0, ps = -9.0 / 3628800.0;  // k = 8
    ph = fma(ph, -v, 1.0 / 40320.0);  ps = fma(ps, -v, -8.0 / 362880.0);
    ph = fma(ph, -v, 1.0 / 5040.0);   ps = fma(ps, -v, -7.0 / 40320.0);
    ph = fma(ph, -v, 1.0 / 720.0);    ps = fma(ps, -v, -6.0 / 5040.0);
    ph = fma(ph, -v, 1.0 / 120.0);    ps = fma(ps, -v, -5.0 / 720.0);
    ph = fma(ph, -v, 1.0 / 24.0);     ps = fma(ps, -v, -4.0 / 120.0);
    ph = fma(ph, -v, 1.0 / 6.0);      ps = fma(ps, -v, -3.0 / 24.0);
    ph = fma(ph, -v, 1.0 / 2.0);      ps = fma(ps, -v, -2.0 / 6.0);
    ph = fma(ph, -v, 1.0);            ps = fma(ps, -v, -1.0 / 2.0);
    G = ph; H = 3.0 * s * ps;
  } else {
    const double E = exp_neg_fp64(v);
    const double g = 1.0 - E;
    G = g * y3; H = 3.0 * (E * v - g) * y5;
  }
  A = select_zero(z, q1 * G);
  B = select_zero(z, (0.5 * q3) * H);
}

// SPLIT (leaf kernels, as uj_tile): the lanes of a warp form nsplit groups that own the same targets and take
// every nsplit-th record of the tile; a group past the end re-reads the last record with A = B = 0.
template <int K, int T, int UNROLL, bool SPLIT, class TAB>
__device__ __forceinline__ void uj_tile_tab(const double2 *__restrict__ tile, int n,
                                            const double (&tx)[T], const double (&ty)[T],
                                            const double (&tz)[T], double (&acc)[T][kAcc],
                                            int shortcut, const TAB &tab, int nsplit = 1, int phase = 0) {
  const int trips = SPLIT ? (n + nsplit - 1) / nsplit : n;
#pragma unroll UNROLL
  for (int jj = 0; jj < trips; ++jj) {
    int j = jj;
    bool live = true;
    if constexpr (SPLIT) {
      j = jj * nsplit + phase;
      live = j < n;
      j = live ? j : n - 1;
    }
    double sx, sy, sz, q0, gx, gy, gz, q1, q2, q3;
    load_rec<false>(tile, j, sx, sy, sz, q0, gx, gy, gz, q1, q2, q3);
    double dx[T], dy[T], dz[T], A[T], B[T], r2[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      dx[t] = tx[t] - sx; dy[t] = ty[t] - sy; dz[t] = tz[t] - sz;
      r2[t] = fma(dz[t], dz[t], fma(dy[t], dy[t], dx[t] * dx[t]));
    }
    // far-field test on the integer pipe: r2 > cutoff  <=  hi word strictly greater
    const int far_hi = __double2hiint(q2);
    bool near = !shortcut;
#pragma unroll
    for (int t = 0; t < T; ++t) near |= __double2hiint(r2[t]) <= far_hi;
    if (__any_sync(0xffffffffu, near)) {
      // one code path for the whole warp: lanes beyond the cut-off read the power-law rows
      using TB = TabOf<K>;
      double tt[T];
      bool far[T], rare = false;
#pragma unroll
      for (int t = 0; t < T; ++t) {
        far[t] = __double2hiint(r2[t]) > far_hi;
        tt[t] = fma(r2[t], q0, far[t] ? 0.0 : (double)TB::offset);
        if constexpr (K == K_GAUS) rare |= !far[t] && __double2hiint(tt[t]) < ((1023 + TB::emin) << 20);
      }
      if (K == K_GAUS && __any_sync(0xffffffffu, rare)) {
#pragma unroll
        for (int t = 0; t < T; ++t) ab_gaus_u(r2[t], far[t], q0, q1, q3, A[t], B[t]);
      } else {
#pragma unroll
        for (int t = 0; t < T; ++t) ab_tab<K>(tt[t], r2[t], far[t], q1, q3, tab, A[t], B[t]);
      }
    } else {
      // every pair of the warp is in the far field: g = 1, dg = 0 (cheaper than the table)
#pragma unroll
      for (int t = 0; t < T; ++t) ab_sing(r2[t], A[t], B[t]);
    }
    if constexpr (SPLIT) {
#pragma unroll
      for (int t = 0; t < T; ++t) {
        A[t] = select_zero(!live, A[t]);
        B[t] = select_zero(!live, B[t]);
      }
    }
#pragma unroll
    for (int t = 0; t < T; ++t) {
      double cx = fma(dy[t], gz, -(dz[t] * gy));
      double cy = fma(dz[t], gx, -(dx[t] * gz));
      double cz = fma(dx[t], gy, -(dy[t] * gx));
      double *s = acc[t];
      s[0] = fma(A[t], cx, s[0]);
      s[1] = fma(A[t], cy, s[1]);
      s[2] = fma(A[t], cz, s[2]);
      s[11] = fma(A[t], gx, s[11]);
      s[12] = fma(A[t], gy, s[12]);
      s[13] = fma(A[t], gz, s[13]);
      double bx = B[t] * cx, by = B[t] * cy, bz = B[t] * cz;
      s[3] = fma(bx, dx[t], s[3]);
      s[4] = fma(by, dx[t], s[4]);
      s[5] = fma(bz, dx[t], s[5]);
      s[6] = fma(bx, dy[t], s[6]);
      s[7] = fma(by, dy[t], s[7]);
      s[8] = fma(bz, dy[t], s[8]);
      s[9] = fma(bx, dz[t], s[9]);
      s[10] = fma(by, dz[t], s[10]);
    }
  }
}

// hook of uj_pairs_kernel (vpm_kernels.cuh): 128-thread CTAs, table rows through L1
template <int K, int T, int UNROLL>
struct PairTileTab {
  static __device__ __forceinline__ void run(const double2 *__restrict__ tile, int n, const double (&tx)[T],
                                             const double (&ty)[T], const double (&tz)[T],
                                             double (&acc)[T][kAcc], int shortcut) {
    uj_tile_tab<K, T, UNROLL, false>(tile, n, tx, ty, tz, acc, shortcut, tab_global<K>());
  }
  static __device__ __forceinline__ void prefetch() {
    const char *base = reinterpret_cast<const char *>(TabOf<K>::words());
    constexpr int kLines = (TabOf<K>::rows * kTabRowChunks * 16 + 127) / 128;
    for (int i = threadIdx.x; i < kLines; i += blockDim.x)
      asm volatile("prefetch.global.L1 [%0];" ::"l"(base + (size_t)i * 128));
  }
};

// Same contract as uj_pairs_kernel (UjArgs, partial sums [split][kAcc][pstride]); one CTA =
// 512 threads x 2 targets against one contiguous range of source tiles.  Dynamic shared
// memory: [table copies][kStages tiles][mbarriers].
template <int K, int THREADS, int UNROLL>
__global__ void __launch_bounds__(THREADS, 1) uj_pairs_tab_kernel(const UjArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int T = kTabT;
  unsigned char *tab = smem;
  double *tiles = reinterpret_cast<double *>(smem + (size_t)TabOf<K>::rows * kTabRowBytes);
  uint64_t *full = reinterpret_cast<uint64_t *>(tiles + kStages * kTile * kRec);
  load_tab<K>(tab);  // visible after the __syncthreads below

  const int tid = threadIdx.x;
  const uint32_t lane_tab = smem_u32(tab) + 16 * (tid & 7);
  const int64_t tbase = (int64_t)blockIdx.x * (THREADS * T);

  double tx[T], ty[T], tz[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    int64_t i = tbase + (int64_t)t * THREADS + tid;
    if (i >= a.nt) i = a.nt - 1;
    const double *p = a.tpos + i * a.tld;
    tx[t] = p[0]; ty[t] = p[1]; tz[t] = p[2];
  }
  double acc[T][kAcc];
#pragma unroll
  for (int t = 0; t < T; ++t)
#pragma unroll
    for (int k = 0; k < kAcc; ++k) acc[t][k] = 0.0;

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
    const uint32_t bytes = (uint32_t)n * kRec * sizeof(double);
    const int st = it % kStages;
    mbar_expect_tx(&full[st], bytes);
    tma_bulk_g2s(tiles + st * kTile * kRec, a.rec + first * kRec, bytes, &full[st]);
  };
  if (tid == 0) {
    for (int s = 0; s < kStages && s < ntl; ++s) issue(s);
  }

  for (int it = 0; it < ntl; ++it) {
    const int st = it % kStages;
    mbar_wait(&full[st], (uint32_t)((it / kStages) & 1));
    const int64_t first = (tile0 + it) * kTile;
    const int n = (int)((a.ns - first) < kTile ? (a.ns - first) : kTile);
    const double2 *tile = reinterpret_cast<const double2 *>(tiles + st * kTile * kRec);
    uj_tile_tab<K, kTabT, UNROLL, false>(tile, n, tx, ty, tz, acc, a.shortcut, TabShared{lane_tab});
    __syncthreads();  // everyone is done reading stage st
    if (tid == 0 && it + kStages < ntl) issue(it + kStages);
  }

#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int64_t i = tbase + (int64_t)t * THREADS + tid;
    if (i < a.nt) {
      double *o = a.partial + (int64_t)blockIdx.y * kAcc * a.pstride + i;
#pragma unroll
      for (int k = 0; k < kAcc; ++k) o[(int64_t)k * a.pstride] = acc[t][k];
    }
  }
}

// How often does a warp of the sweep see a pair inside the regularised range?  2048 samples, each one warp
// = 32 CONSECUTIVE targets against one pseudo-random source (a fixed hash of the sample index: the same
// field gives the same count, so the kernel choice that depends on it is deterministic).
// out[0] += samples in which at least one of the 32 pairs has r^2 < cutoff_u * sigma_source^2.
// It is the warp-level share that matters, not the pair-level one: in a spatially ordered field (rings,
// sorted clouds) near pairs come in whole warps and the far-field shortcut still serves most warps, in an
// unordered dense field nearly every warp mixes near and far lanes.
__global__ void sample_near_kernel(SrcView src, int64_t s0, int64_t ns, const double *__restrict__ tpos, int64_t tld,
                                   int64_t nt, double cutoff_u, unsigned int *out) {
  const unsigned k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  auto mix = [](unsigned long long x) {  // splitmix64 finaliser
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
  };
  const int64_t nblk = (nt + 31) / 32;
  int64_t i = (int64_t)(mix(2ull * k) % (unsigned long long)nblk) * 32 + lane;
  if (i >= nt) i = nt - 1;
  const int64_t j = (int64_t)(mix(2ull * k + 1) % (unsigned long long)ns);
  const double *t = tpos + i * tld;
  const double *p = src.p + (s0 + j) * src.ld;
  const double dx = t[0] - p[src.ox], dy = t[1] - p[src.ox + 1], dz = t[2] - p[src.ox + 2];
  const double sg = p[src.osig];
  const bool near = dx * dx + dy * dy + dz * dz < cutoff_u * sg * sg;
  const unsigned m = __ballot_sync(0xffffffffu, near);
  if (lane == 0 && m != 0) atomicAdd(out, 1u);
}
constexpr int kSampleBlocks = 256, kSampleThreads = 256;  // 2048 warps

// device-math test hook (vpm_test_math op 4): (A, B) of the table path at r2 = in[i], sigma = 1
template <int K>
__global__ void test_tab_kernel(const double *in, double *out, double *out2, int64_t n) {
  extern __shared__ __align__(128) unsigned char smem[];
  load_tab<K>(smem);
  __syncthreads();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double A, B;
  const double x = in[i];  // r2 at sigma = 1: q0 = 1, q1 = 1, q3 = 2
  const bool far = x > (K == K_GERF ? kFarU_gerf : kFarU_gaus);
  const double t = fma(x, 1.0, far ? 0.0 : (double)TabOf<K>::offset);
  if (K == K_GAUS && !far && __double2hiint(t) < ((1023 + TabOf<K>::emin) << 20)) ab_gaus_u(x, far, 1.0, 1.0, 2.0, A, B);
  else ab_tab<K>(t, x, far, 1.0, 2.0, TabShared{smem_u32(smem) + 16 * (threadIdx.x & 7)}, A, B);
  out[i] = A;
  if (out2) out2[i] = B;
}

}  // namespace vpm
