/*
 * vpm_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, FP64, no FMA contraction) of the arithmetic of
 * byuflowlab/FLOWVPM.jl v4.0.3 on the particle-to-particle hot path.  It is
 * the checker the CUDA path is compared with; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The
 * product (libvpm_cuda.so and the flowvpm.jl_b200 host layer) never does.
 *
 * PARITY PINNING: the reference ships no golden vectors for this path and
 * Julia is not installable here, so the reference itself cannot be run.  The
 * oracle is pinned by (i) the analytic two-particle known-answer formulas of
 * the reference's scripts/check_fmm.jl:39-98, (ii) a 50-digit mpmath
 * evaluation of the reference formulas (oracle/hp_oracle.py), (iii) the
 * physics assertion of test/runtests_singlevortexring.jl:127-143.  Against
 * reference-run outputs it is UNPINNED (see DESIGN.md "Oracle").
 *
 * Every function cites the reference file:line (relative to /root/reference)
 * whose operation order it follows.  Build: oracle/Makefile
 * (gcc -O2 -ffp-contract=off -fopenmp): Julia does not contract a*b+c into an
 * FMA unless asked, so neither may the compiler here.
 *
 * Layouts (all column-major, as the Julia side owns them):
 *   particle field  : nfields(=46) x np, rows per src/FLOWVPM_particlefield.jl:239-252
 *                     (0-based here: X 0:3, Gamma 3:6, sigma 6, U 9:12, vort 12:15,
 *                      J 15:24, PSE 24:27, SFS 39:42, static 42)
 *   source buffer   : 8 x ns  [x y z rho Gx Gy Gz sigma]   src/FLOWVPM_fmm.jl:62-71
 *   target buffer   : ld x nt, rows 0:3 position, 3 scalar potential,
 *                     4:7 gradient (velocity), 7:16 hessian (J, column-major 3x3)
 *                     -- FastMultipole's convention, reached in the reference only
 *                     through get_position/set_gradient!/set_hessian!.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define VPM_KERNEL_SINGULAR 0
#define VPM_KERNEL_GAUSSIAN 1
#define VPM_KERNEL_GAUSSIANERF 2
#define VPM_KERNEL_WINCKELMANS 3

#define VPM_FLAG_RESET 1
#define VPM_FLAG_RESET_SFS 2
#define VPM_FLAG_SFS 4
#define VPM_FLAG_TRANSPOSED 8

/* rows of the particle matrix, 0-based (src/FLOWVPM_particlefield.jl:239-252) */
enum { R_X = 0, R_G = 3, R_SIGMA = 6, R_U = 9, R_W = 12, R_J = 15, R_PSE = 24,
       R_SFS = 39, R_STATIC = 42 };

/* ---- constants as the reference computes them (src/FLOWVPM.jl:62-66) ---- */
static double c_const1, c_const2, c_const3, c_const4, c_sqr2;
static int c_init_done = 0;
static void init_consts(void) {
  if (c_init_done) return;
  const double pi = 3.14159265358979323846; /* Julia's pi rounds to this double */
  c_const1 = 1.0 / pow(2.0 * pi, 1.5);
  c_const2 = sqrt(2.0 / pi);
  c_const3 = 3.0 / (4.0 * pi);
  c_const4 = 1.0 / (4.0 * pi);
  c_sqr2 = sqrt(2.0);
  c_init_done = 1;
}

/* ---- custom_erf64: src/FLOWVPM_gpu_erf.jl:159-191, coefficients :63-121 ---- */
static const double erx = 8.45062911510467529297e-01;
static const double pp0 = 1.28379167095512558561e-01, pp1 = -3.25042107247001499370e-01,
                    pp2 = -2.84817495755985104766e-02, pp3 = -5.77027029648944159157e-03,
                    pp4 = -2.37630166566501626084e-05;
static const double qq1 = 3.97917223959155352819e-01, qq2 = 6.50222499887672944485e-02,
                    qq3 = 5.08130628187576562776e-03, qq4 = 1.32494738004321644526e-04,
                    qq5 = -3.96022827877536812320e-06;
static const double pa0 = -2.36211856075265944077e-03, pa1 = 4.14856118683748331666e-01,
                    pa2 = -3.72207876035701323847e-01, pa3 = 3.18346619901161753674e-01,
                    pa4 = -1.10894694282396677476e-01, pa5 = 3.54783043256182359371e-02,
                    pa6 = -2.16637559486879084300e-03;
static const double qa1 = 1.06420880400844228286e-01, qa2 = 5.40397917702171048937e-01,
                    qa3 = 7.18286544141962662868e-02, qa4 = 1.26171219808761642112e-01,
                    qa5 = 1.36370839120290507362e-02, qa6 = 1.19844998467991074170e-02;
static const double ra0 = -9.86494403484714822705e-03, ra1 = -6.93858572707181764372e-01,
                    ra2 = -1.05586262253232909814e+01, ra3 = -6.23753324503260060396e+01,
                    ra4 = -1.62396669462573470355e+02, ra5 = -1.84605092906711035994e+02,
                    ra6 = -8.12874355063065934246e+01, ra7 = -9.81432934416914548592e+00;
static const double sa1 = 1.96512716674392571292e+01, sa2 = 1.37657754143519042600e+02,
                    sa3 = 4.34565877475229228821e+02, sa4 = 6.45387271733267880336e+02,
                    sa5 = 4.29008140027567833386e+02, sa6 = 1.08635005541779435134e+02,
                    sa7 = 6.57024977031928170135e+00, sa8 = -6.04244152148580987438e-02;
static const double rb0 = -9.86494292470009928597e-03, rb1 = -7.99283237680523006574e-01,
                    rb2 = -1.77579549177547519889e+01, rb3 = -1.60636384855821916062e+02,
                    rb4 = -6.37566443368389627722e+02, rb5 = -1.02509513161107724954e+03,
                    rb6 = -4.83519191608651397019e+02;
static const double sb1 = 3.03380607434824582924e+01, sb2 = 3.25792512996573918826e+02,
                    sb3 = 1.53672958608443695994e+03, sb4 = 3.19985821950859553908e+03,
                    sb5 = 2.55305040643316442583e+03, sb6 = 4.74528541206955367215e+02,
                    sb7 = -2.24409524465858183362e+01;

static double jl_sign(double x) { return x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : x); }

double vpm_oracle_erf64(double x) {
  double xabs = fabs(x), sgn = jl_sign(x), val = sgn * 1.0;
  if (xabs < 0.84375) { /* :166-170 */
    double z = x * x;
    double r = pp0 + z * (pp1 + z * (pp2 + z * (pp3 + z * pp4)));
    double s = 1.0 + z * (qq1 + z * (qq2 + z * (qq3 + z * (qq4 + z * qq5))));
    double y = r / s;
    val = sgn * (xabs + xabs * y);
  } else if (xabs < 1.25) { /* :171-175 */
    double s = xabs - 1.0;
    double P = pa0 + s * (pa1 + s * (pa2 + s * (pa3 + s * (pa4 + s * (pa5 + s * pa6)))));
    double Q = 1.0 + s * (qa1 + s * (qa2 + s * (qa3 + s * (qa4 + s * (qa5 + s * qa6)))));
    val = sgn * (erx + P / Q);
  } else if (xabs < 2.857142857142857) { /* :176-181 */
    double s = 1.0 / (x * x);
    double R = ra0 + s * (ra1 + s * (ra2 + s * (ra3 + s * (ra4 + s * (ra5 + s * (ra6 + s * ra7))))));
    double S = 1.0 + s * (sa1 + s * (sa2 + s * (sa3 + s * (sa4 + s * (sa5 + s * (sa6 + s * (sa7 + s * sa8)))))));
    double r = exp(-x * x - 0.5625 + R / S);
    val = sgn * (1.0 - r / xabs);
  } else if (xabs < 6.0) { /* :182-187 */
    double s = 1.0 / (x * x);
    double R = rb0 + s * (rb1 + s * (rb2 + s * (rb3 + s * (rb4 + s * (rb5 + s * rb6)))));
    double S = 1.0 + s * (sb1 + s * (sb2 + s * (sb3 + s * (sb4 + s * (sb5 + s * (sb6 + s * sb7))))));
    double r = exp(-x * x - 0.5625 + R / S);
    val = sgn * (1.0 - r / xabs);
  }
  return val;
}

/* ---- g_dgdr of the four families: src/FLOWVPM_kernel.jl:44-84 ---- */
static inline void g_dgdr(int kernel, double r, double *g, double *dg) {
  switch (kernel) {
    case VPM_KERNEL_SINGULAR: /* :48 */
      *g = 1.0; *dg = 0.0; break;
    case VPM_KERNEL_GAUSSIANERF: { /* :54-57 */
      double aux = c_const2 * r * exp(-r * r / 2);
      *g = vpm_oracle_erf64(r / c_sqr2) - aux;
      *dg = r * aux;
    } break;
    case VPM_KERNEL_GAUSSIAN: { /* :63-66 */
      double aux = exp(-r * r * r);
      *g = 1 - aux;
      *dg = 3 * r * r * aux;
    } break;
    default: { /* winckelmans :77-84 */
      double aux0 = r * r + 1;
      double aux02 = aux0 * aux0;
      aux0 = aux02 * aux02 * aux0;
      aux0 = sqrt(aux0);
      *g = r * r * r * (r * r + 2.5) / aux0;
      *dg = 7.5 * r * r / (aux0 * (r * r + 1));
    }
  }
}

/* zeta: src/FLOWVPM_kernel.jl:45,51,60,69-74 */
static inline double zeta_fn(int kernel, double r) {
  switch (kernel) {
    case VPM_KERNEL_SINGULAR: return r == 0.0 ? 1.0 : 0.0;
    case VPM_KERNEL_GAUSSIANERF: return c_const1 * exp(-r * r / 2);
    case VPM_KERNEL_GAUSSIAN: return c_const3 * exp(-r * r * r);
    default: {
      double temp = r * r + 1;
      double temp2 = temp * temp * temp;
      temp = temp2 * temp2 * temp;
      return c_const4 * 7.5 / sqrt(temp);
    }
  }
}

void vpm_oracle_g_dgdr(int kernel, double r, double *g, double *dg) {
  init_consts();
  g_dgdr(kernel, r, g, dg);
}
double vpm_oracle_zeta(int kernel, double r) {
  init_consts();
  return zeta_fn(kernel, r);
}

/*
 * The 6-argument fmm.direct! overload, src/FLOWVPM_fmm.jl:102-168: sources in
 * the outer loop, targets in the inner loop, pair skipped iff r2 == 0, the
 * setters accumulate.  Half-open 0-based ranges [t0,t1), [s0,s1).
 */
void vpm_oracle_direct_buffers(double *tgt, int64_t ld, int64_t t0, int64_t t1,
                               const double *src, int64_t s0, int64_t s1,
                               int kernel, int want_U, int want_J) {
  init_consts();
  for (int64_t is = s0; is < s1; ++is) {
    const double *S = src + 8 * is;
    const double gamma_x = S[4], gamma_y = S[5], gamma_z = S[6];
    const double source_x = S[0], source_y = S[1], source_z = S[2];
    const double sigma = S[7];
    for (int64_t jt = t0; jt < t1; ++jt) {
      double *T = tgt + ld * jt;
      double dx = T[0] - source_x, dy = T[1] - source_y, dz = T[2] - source_z;
      double r2 = dx * dx + dy * dy + dz * dz;
      if (r2 != 0.0) {
        double r = sqrt(r2);
        double g_sgm, dg_sgmdr;
        g_dgdr(kernel, r / sigma, &g_sgm, &dg_sgmdr);
        double r3inv = 1.0 / (r2 * r);
        double crss1 = -c_const4 * r3inv * (dy * gamma_z - dz * gamma_y);
        double crss2 = -c_const4 * r3inv * (dz * gamma_x - dx * gamma_z);
        double crss3 = -c_const4 * r3inv * (dx * gamma_y - dy * gamma_x);
        if (want_U) {
          T[4] += g_sgm * crss1;
          T[5] += g_sgm * crss2;
          T[6] += g_sgm * crss3;
        }
        if (want_J) {
          double aux = dg_sgmdr / (sigma * r) - 3 * g_sgm / r2;
          double aux2 = -c_const4 * g_sgm * r3inv;
          T[7] += aux * crss1 * dx;
          T[8] += aux * crss2 * dx - aux2 * gamma_z;
          T[9] += aux * crss3 * dx + aux2 * gamma_y;
          T[10] += aux * crss1 * dy + aux2 * gamma_z;
          T[11] += aux * crss2 * dy;
          T[12] += aux * crss3 * dy - aux2 * gamma_x;
          T[13] += aux * crss1 * dz - aux2 * gamma_y;
          T[14] += aux * crss2 * dz + aux2 * gamma_x;
          T[15] += aux * crss3 * dz;
        }
      }
    }
  }
}

/*
 * ---- Float32 fields (ParticleField(n, Float32): R is a type parameter, src/FLOWVPM_particlefield.jl:120,134) ----
 * Julia's promotion rules applied line by line to the SAME source: Float32 op Float32 stays Float32, an
 * Int literal keeps the other operand's type, a Float64 constant or literal (const4, const2, sqr2, 2.5,
 * 7.5, 3.0 ...) promotes to Float64; stores into the Float32 buffers round.  Below `float` expressions are
 * Float32 arithmetic and `double` ones Float64 (gcc -O2 without -ffast-math / FLT_EVAL_METHOD 0 on x86-64).
 *
 * custom_erf32, src/FLOWVPM_gpu_erf.jl:124-156, coefficients :2-60 (the Float32 roundings of the Float64
 * table, written here as casts).  Quirk restated: the last branch uses the Float64 `sb7` (:150), so S, the
 * exponent and the result of that branch are Float64.  NOTE: g_dgdr_gauserf calls custom_erf(r/sqr2) with
 * the Float64 constant sqr2, so a Float32 field reaches custom_erf64, not this function; it is restated
 * because it is part of the reference's file and is checked against libm in tests/test_oracle_pinning.py.
 */
double vpm_oracle_erf32(float x) {
  const float xabs = fabsf(x), sgn = x > 0.f ? 1.f : (x < 0.f ? -1.f : x), oneval = 1.f;
  float val = sgn * oneval;
#define F(c) ((float)(c))
  if (xabs < 0.84375f) {
    float z = x * x;
    float r = F(pp0) + z * (F(pp1) + z * (F(pp2) + z * (F(pp3) + z * F(pp4))));
    float s = oneval + z * (F(qq1) + z * (F(qq2) + z * (F(qq3) + z * (F(qq4) + z * F(qq5)))));
    float y = r / s;
    val = sgn * (xabs + xabs * y);
  } else if (xabs < 1.25f) {
    float s = xabs - oneval;
    float P = F(pa0) + s * (F(pa1) + s * (F(pa2) + s * (F(pa3) + s * (F(pa4) + s * (F(pa5) + s * F(pa6))))));
    float Q = oneval + s * (F(qa1) + s * (F(qa2) + s * (F(qa3) + s * (F(qa4) + s * (F(qa5) + s * F(qa6))))));
    val = sgn * (F(erx) + P / Q);
  } else if (xabs < 2.857142857142857f) {
    float s = oneval / (x * x);
    float R = F(ra0) + s * (F(ra1) + s * (F(ra2) + s * (F(ra3) + s * (F(ra4) + s * (F(ra5) + s * (F(ra6) + s * F(ra7)))))));
    float S = oneval + s * (F(sa1) + s * (F(sa2) + s * (F(sa3) + s * (F(sa4) + s * (F(sa5) + s * (F(sa6) + s * (F(sa7) + s * F(sa8))))))));
    float r = expf(-x * x - 0.5625f + R / S);
    val = sgn * (oneval - r / xabs);
  } else if (xabs < 6.0f) {
    float s = oneval / (x * x);
    float R = F(rb0) + s * (F(rb1) + s * (F(rb2) + s * (F(rb3) + s * (F(rb4) + s * (F(rb5) + s * F(rb6))))));
    /* s*sb7 with the Float64 sb7 (:150): everything outward of it is Float64 */
    double S = (double)oneval + (double)s * ((double)F(sb1) + (double)s * ((double)F(sb2) + (double)s * ((double)F(sb3) +
               (double)s * ((double)F(sb4) + (double)s * ((double)F(sb5) + (double)s * ((double)F(sb6) + (double)s * sb7))))));
    double r = exp((double)(-x * x - 0.5625f) + (double)R / S);
    return (double)sgn * ((double)oneval - r / (double)xabs);
  }
#undef F
  return (double)val;
}

/* g_dgdr(r::Float32) of the four families with Julia's promotion (src/FLOWVPM_kernel.jl:44-84) */
static inline void g_dgdr_f32(int kernel, float r, double *g, double *dg) {
  switch (kernel) {
    case VPM_KERNEL_SINGULAR: *g = 1.0; *dg = 0.0; return;
    case VPM_KERNEL_GAUSSIANERF: { /* const2, sqr2 are Float64; -r*r/2 and its exp are Float32 */
      float e = expf(-r * r / 2);
      double aux = c_const2 * (double)r * (double)e;
      *g = vpm_oracle_erf64((double)r / c_sqr2) - aux;
      *dg = (double)r * aux;
      return;
    }
    case VPM_KERNEL_GAUSSIAN: { /* integer literals only: everything stays Float32 */
      float aux = expf(-r * r * r);
      *g = (double)(1 - aux);
      *dg = (double)(3 * r * r * aux);
      return;
    }
    default: { /* winckelmans: aux0 chain Float32, the literals 2.5 and 7.5 promote */
      float aux0 = r * r + 1;
      float aux02 = aux0 * aux0;
      aux0 = aux02 * aux02 * aux0;
      aux0 = sqrtf(aux0);
      *g = (double)(r * r * r) * ((double)(r * r) + 2.5) / (double)aux0;
      *dg = 7.5 * (double)r * (double)r / (double)(aux0 * (r * r + 1));
      return;
    }
  }
}

/*
 * fmm.direct! (src/FLOWVPM_fmm.jl:102-168) on Float32 buffers: dx, r2, r, r3inv and the cross-product
 * factors are Float32; -const4 promotes crss, U, aux2 and the J entries to Float64; every setter adds
 * in Float64 and rounds into the Float32 buffer (`buffer[i] += val`).
 */
void vpm_oracle_direct_buffers_f32(float *tgt, int64_t ld, int64_t t0, int64_t t1, const float *src,
                                   int64_t s0, int64_t s1, int kernel, int want_U, int want_J) {
  init_consts();
  for (int64_t is = s0; is < s1; ++is) {
    const float *S = src + 8 * is;
    const float gamma_x = S[4], gamma_y = S[5], gamma_z = S[6];
    const float source_x = S[0], source_y = S[1], source_z = S[2];
    const float sigma = S[7];
    for (int64_t jt = t0; jt < t1; ++jt) {
      float *T = tgt + ld * jt;
      float dx = T[0] - source_x, dy = T[1] - source_y, dz = T[2] - source_z;
      float r2 = dx * dx + dy * dy + dz * dz;
      if (r2 != 0.0f) {
        float r = sqrtf(r2);
        double g_sgm, dg_sgmdr;
        g_dgdr_f32(kernel, r / sigma, &g_sgm, &dg_sgmdr);
        float r3inv = 1.0f / (r2 * r);
        double crss1 = -c_const4 * (double)r3inv * (double)(dy * gamma_z - dz * gamma_y);
        double crss2 = -c_const4 * (double)r3inv * (double)(dz * gamma_x - dx * gamma_z);
        double crss3 = -c_const4 * (double)r3inv * (double)(dx * gamma_y - dy * gamma_x);
        if (want_U) {
          T[4] = (float)((double)T[4] + g_sgm * crss1);
          T[5] = (float)((double)T[5] + g_sgm * crss2);
          T[6] = (float)((double)T[6] + g_sgm * crss3);
        }
        if (want_J) {
          double aux = dg_sgmdr / (double)(sigma * r) - 3 * g_sgm / (double)r2;
          double aux2 = -c_const4 * g_sgm * (double)r3inv;
          T[7] = (float)((double)T[7] + (aux * crss1 * (double)dx));
          T[8] = (float)((double)T[8] + (aux * crss2 * (double)dx - aux2 * (double)gamma_z));
          T[9] = (float)((double)T[9] + (aux * crss3 * (double)dx + aux2 * (double)gamma_y));
          T[10] = (float)((double)T[10] + (aux * crss1 * (double)dy + aux2 * (double)gamma_z));
          T[11] = (float)((double)T[11] + (aux * crss2 * (double)dy));
          T[12] = (float)((double)T[12] + (aux * crss3 * (double)dy - aux2 * (double)gamma_x));
          T[13] = (float)((double)T[13] + (aux * crss1 * (double)dz - aux2 * (double)gamma_y));
          T[14] = (float)((double)T[14] + (aux * crss2 * (double)dz + aux2 * (double)gamma_x));
          T[15] = (float)((double)T[15] + (aux * crss3 * (double)dz));
        }
      }
    }
  }
}

/*
 * Threaded form: contiguous target blocks per thread, every block sees all
 * sources in index order -- the per-target accumulation order is therefore the
 * same as the serial loop (FastMultipole's threaded driver is not under
 * /root/reference; the block split mirrors Estr_direct_multithreaded,
 * src/FLOWVPM_subfilterscale_models.jl:51-59).
 */
void vpm_oracle_direct_buffers_mt(double *tgt, int64_t ld, int64_t t0, int64_t t1,
                                  const double *src, int64_t s0, int64_t s1,
                                  int kernel, int want_U, int want_J, int nthreads) {
  if (nthreads <= 1 || t1 - t0 < 2) {
    vpm_oracle_direct_buffers(tgt, ld, t0, t1, src, s0, s1, kernel, want_U, want_J);
    return;
  }
  init_consts();
  int64_t nt = t1 - t0;
  /* contiguous target blocks, at least four per thread so that every thread is busy even on
   * a short target slice (one block per thread is the reference's split,
   * src/FLOWVPM_subfilterscale_models.jl:51-59; smaller blocks only balance the tail), and at
   * most 256 targets so that a block stays in L1/L2 while the sources stream */
  int64_t blk = (nt + 4 * (int64_t)nthreads - 1) / (4 * (int64_t)nthreads);
  if (blk > 256) blk = 256;
  if (blk < 1) blk = 1;
  int64_t nblk = (nt + blk - 1) / blk;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int64_t b = 0; b < nblk; ++b) {
    int64_t a = t0 + b * blk, e = a + blk < t1 ? a + blk : t1;
    vpm_oracle_direct_buffers(tgt, ld, a, e, src, s0, s1, kernel, want_U, want_J);
  }
}

/* reset rules: src/FLOWVPM_particlefield.jl:464-511 */
void vpm_oracle_reset_particles(double *P, int64_t nf, int64_t np) {
  for (int64_t i = 0; i < np; ++i) {
    double *p = P + nf * i;
    if (p[R_STATIC] == 0.0) {
      for (int k = 0; k < 3; ++k) p[R_U + k] = 0.0;
      for (int k = 0; k < 3; ++k) p[R_W + k] = 0.0;
      for (int k = 0; k < 9; ++k) p[R_J + k] = 0.0;
      for (int k = 0; k < 3; ++k) p[R_PSE + k] = 0.0;
    }
  }
}
void vpm_oracle_reset_particles_sfs(double *P, int64_t nf, int64_t np) {
  for (int64_t i = 0; i < np; ++i) {
    double *p = P + nf * i;
    if (p[R_STATIC] == 0.0)
      for (int k = 0; k < 3; ++k) p[R_SFS + k] = 0.0;
  }
}

/* Estr_direct pair term: src/FLOWVPM_subfilterscale_models.jl:16-41 */
static inline void estr_pair(double *tp, const double *sp, double r, int kernel, int transposed) {
  const double *GS = sp + R_G, *JS = sp + R_J, *JT = tp + R_J;
  double S1, S2, S3;
  if (transposed) {
    S1 = (JT[0] - JS[0]) * GS[0] + (JT[1] - JS[1]) * GS[1] + (JT[2] - JS[2]) * GS[2];
    S2 = (JT[3] - JS[3]) * GS[0] + (JT[4] - JS[4]) * GS[1] + (JT[5] - JS[5]) * GS[2];
    S3 = (JT[6] - JS[6]) * GS[0] + (JT[7] - JS[7]) * GS[1] + (JT[8] - JS[8]) * GS[2];
  } else {
    S1 = (JT[0] - JS[0]) * GS[0] + (JT[3] - JS[3]) * GS[1] + (JT[6] - JS[6]) * GS[2];
    S2 = (JT[1] - JS[1]) * GS[0] + (JT[4] - JS[4]) * GS[1] + (JT[7] - JS[7]) * GS[2];
    S3 = (JT[2] - JS[2]) * GS[0] + (JT[5] - JS[5]) * GS[1] + (JT[8] - JS[8]) * GS[2];
  }
  double sigma_inv = 1.0 / sp[R_SIGMA];
  double zeta_sgm = zeta_fn(kernel, r * sigma_inv) * sigma_inv * sigma_inv * sigma_inv;
  tp[R_SFS + 0] += zeta_sgm * S1;
  tp[R_SFS + 1] += zeta_sgm * S2;
  tp[R_SFS + 2] += zeta_sgm * S3;
}

/*
 * Estr_direct!: src/FLOWVPM_subfilterscale_models.jl:43-92.  Targets in the
 * outer loop (static targets skipped, :63,80), sources = iterator(pfield) =
 * non-static particles in index order (src/FLOWVPM_particlefield.jl:368-374).
 */
void vpm_oracle_estr_direct(double *P, int64_t nf, int64_t np, int kernel, int transposed,
                            int nthreads) {
  init_consts();
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads > 0 ? nthreads : 1)
  for (int64_t it = 0; it < np; ++it) {
    double *tp = P + nf * it;
    if (tp[R_STATIC] != 0.0) continue;
    double tx = tp[0], ty = tp[1], tz = tp[2];
    for (int64_t is = 0; is < np; ++is) {
      const double *sp = P + nf * is;
      if (sp[R_STATIC] != 0.0) continue;
      double dx = sp[0] - tx, dy = sp[1] - ty, dz = sp[2] - tz;
      double r = sqrt(dx * dx + dy * dy + dz * dz);
      estr_pair(tp, sp, r, kernel, transposed);
    }
  }
}

/*
 * The same sweep for a LIST of targets only (full-size parity tests: the GPU does the whole
 * field, the oracle re-does a few hundred targets against all sources).  `targets[k]` are
 * particle indices; the sums are added into P's own SFS rows of those particles.
 */
void vpm_oracle_estr_direct_targets(double *P, int64_t nf, int64_t np, const int64_t *targets,
                                    int64_t ntargets, int kernel, int transposed, int nthreads) {
  init_consts();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : 1)
  for (int64_t k = 0; k < ntargets; ++k) {
    double *tp = P + nf * targets[k];
    if (tp[R_STATIC] != 0.0) continue;
    double tx = tp[0], ty = tp[1], tz = tp[2];
    for (int64_t is = 0; is < np; ++is) {
      const double *sp = P + nf * is;
      if (sp[R_STATIC] != 0.0) continue;
      double dx = sp[0] - tx, dy = sp[1] - ty, dz = sp[2] - tz;
      double r = sqrt(dx * dx + dy * dy + dz * dz);
      estr_pair(tp, sp, r, kernel, transposed);
    }
  }
}

/*
 * UJ_direct(pfield; sfs, reset, reset_sfs): src/FLOWVPM_UJ.jl:21-37, with the
 * FastMultipole driver it calls restated from the callbacks it must use
 * (source_system_to_buffer! src/FLOWVPM_fmm.jl:62-71, zeroed target buffer,
 * direct! :102-168, buffer_to_target_system! :170-176).
 */
int vpm_oracle_uj_direct(double *P, int64_t nf, int64_t np, int kernel, int flags, int nthreads) {
  init_consts();
  if (nf < 43) return -1;
  if (flags & VPM_FLAG_RESET) vpm_oracle_reset_particles(P, nf, np);
  if (flags & VPM_FLAG_RESET_SFS) vpm_oracle_reset_particles_sfs(P, nf, np);
  if (np > 0) {
    double *src = (double *)malloc(sizeof(double) * 8 * np);
    double *tgt = (double *)calloc(16 * np, sizeof(double));
    if (!src || !tgt) { free(src); free(tgt); return -2; }
    for (int64_t i = 0; i < np; ++i) {
      const double *p = P + nf * i;
      double *s = src + 8 * i;
      s[0] = p[0]; s[1] = p[1]; s[2] = p[2];
      s[3] = p[R_SIGMA]; /* rho: regularisation radius, unused by direct! */
      s[4] = p[3]; s[5] = p[4]; s[6] = p[5];
      s[7] = p[R_SIGMA];
      double *t = tgt + 16 * i;
      t[0] = p[0]; t[1] = p[1]; t[2] = p[2];
    }
    vpm_oracle_direct_buffers_mt(tgt, 16, 0, np, src, 0, np, kernel, 1, 1, nthreads);
    for (int64_t i = 0; i < np; ++i) {
      double *p = P + nf * i;
      const double *t = tgt + 16 * i;
      for (int k = 0; k < 3; ++k) p[R_U + k] += t[4 + k];
      for (int k = 0; k < 9; ++k) p[R_J + k] += t[7 + k];
    }
    free(src);
    free(tgt);
  }
  if (flags & VPM_FLAG_SFS)
    vpm_oracle_estr_direct(P, nf, np, kernel, (flags & VPM_FLAG_TRANSPOSED) != 0, nthreads);
  return 0;
}

/*
 * FMM near field over a leaf-pair list ("direct_list"): for each
 * (target leaf, source leaf) the 6-argument direct! on the tree-sorted buffers
 * (call shape src/FLOWVPM_gpu.jl:637-643).  Leaves are half-open
 * 0-based body ranges of the sorted buffers.  Pairs are visited in list order.
 */
void vpm_oracle_direct_leafpairs(double *tgt, int64_t ld, const double *src,
                                 const int64_t *tleaf_begin, const int64_t *tleaf_end,
                                 const int64_t *sleaf_begin, const int64_t *sleaf_end,
                                 const int32_t *pair_t, const int32_t *pair_s, int64_t npairs,
                                 int kernel, int want_U, int want_J) {
  for (int64_t k = 0; k < npairs; ++k) {
    int32_t a = pair_t[k], b = pair_s[k];
    vpm_oracle_direct_buffers(tgt, ld, tleaf_begin[a], tleaf_end[a], src, sleaf_begin[b],
                              sleaf_end[b], kernel, want_U, want_J);
  }
}

/*
 * Estr_fmm! over the same list: src/FLOWVPM_subfilterscale_models.jl:157-188
 * (single-thread form; the threaded form :102-155 visits the same pairs).
 * Sorted body index -> particle column through sort_index (0-based here),
 * source-outer / target-inner, NO static filtering (:131-149).
 */
void vpm_oracle_estr_leafpairs(double *P, int64_t nf, const int64_t *tsort, const int64_t *ssort,
                               const int64_t *tleaf_begin, const int64_t *tleaf_end,
                               const int64_t *sleaf_begin, const int64_t *sleaf_end,
                               const int32_t *pair_t, const int32_t *pair_s, int64_t npairs,
                               int kernel, int transposed) {
  init_consts();
  for (int64_t k = 0; k < npairs; ++k) {
    int32_t a = pair_t[k], b = pair_s[k];
    for (int64_t is = sleaf_begin[b]; is < sleaf_end[b]; ++is) {
      const double *sp = P + nf * ssort[is];
      double sx = sp[0], sy = sp[1], sz = sp[2];
      for (int64_t it = tleaf_begin[a]; it < tleaf_end[a]; ++it) {
        double *tp = P + nf * tsort[it];
        double dx = sx - tp[0], dy = sy - tp[1], dz = sz - tp[2];
        double r = sqrt(dx * dx + dy * dy + dz * dz);
        estr_pair(tp, sp, r, kernel, transposed);
      }
    }
  }
}

/*
 * zeta_direct(pfield): src/FLOWVPM_viscous.jl:488-515.  J[1:3] of every particle (static
 * included) is zeroed, then J[1:3]_i += Gamma_j * (1/sigma_j^3 * zeta(r/sigma_j)) over all j
 * in index order (self term included).
 */
void vpm_oracle_zeta_direct(double *P, int64_t nf, int64_t np, int kernel, int nthreads) {
  init_consts();
  for (int64_t i = 0; i < np; ++i)
    for (int k = 0; k < 3; ++k) P[nf * i + R_J + k] = 0.0;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads > 0 ? nthreads : 1)
  for (int64_t i = 0; i < np; ++i) {
    double *pi = P + nf * i;
    for (int64_t j = 0; j < np; ++j) {
      const double *pj = P + nf * j;
      double dX1 = pi[0] - pj[0], dX2 = pi[1] - pj[1], dX3 = pi[2] - pj[2];
      double r = sqrt(dX1 * dX1 + dX2 * dX2 + dX3 * dX3);
      double sg = pj[R_SIGMA];
      double zeta_sgm = 1 / (sg * sg * sg) * zeta_fn(kernel, r / sg);
      pi[R_J + 0] += pj[R_G + 0] * zeta_sgm;
      pi[R_J + 1] += pj[R_G + 1] * zeta_sgm;
      pi[R_J + 2] += pj[R_G + 2] * zeta_sgm;
    }
  }
}

/*
 * zeta_fmm's list loop: src/FLOWVPM_viscous.jl:535-557.  For a list entry (i_target,
 * i_source): Pi runs over the bodies of the SOURCE branch and receives, Pj over the bodies of
 * the TARGET branch and gives (the reference's naming); no zeroing here.
 */
void vpm_oracle_zeta_leafpairs(double *P, int64_t nf, const int64_t *sort, const int64_t *lb,
                               const int64_t *le, const int32_t *pair_t, const int32_t *pair_s,
                               int64_t npairs, int kernel) {
  init_consts();
  for (int64_t k = 0; k < npairs; ++k) {
    int32_t bt = pair_t[k], bs = pair_s[k];
    for (int64_t is = lb[bs]; is < le[bs]; ++is) {
      double *pi = P + nf * sort[is];
      for (int64_t it = lb[bt]; it < le[bt]; ++it) {
        const double *pj = P + nf * sort[it];
        double dX1 = pi[0] - pj[0], dX2 = pi[1] - pj[1], dX3 = pi[2] - pj[2];
        double r = sqrt(dX1 * dX1 + dX2 * dX2 + dX3 * dX3);
        double sg = pj[R_SIGMA];
        double zeta_sgm = 1 / (sg * sg * sg) * zeta_fn(kernel, r / sg);
        pi[R_J + 0] += pj[R_G + 0] * zeta_sgm;
        pi[R_J + 1] += pj[R_G + 1] * zeta_sgm;
        pi[R_J + 2] += pj[R_G + 2] * zeta_sgm;
      }
    }
  }
}

/*
 * cs.zeta(pfield) as rbf_conjugategradient and CoreSpreading call it (src/FLOWVPM_viscous.jl:204,347,374):
 * zeta_direct unless the test installed another evaluation (zeta_fmm: the Python side builds the leaf lists
 * for the CURRENT X, sigma with oracle/leaflists.py and calls vpm_oracle_zeta_leafpairs).
 */
typedef void (*vpm_oracle_zeta_cb)(double *P, int64_t nf, int64_t np, int kernel);
static vpm_oracle_zeta_cb g_cs_zeta = NULL;
void vpm_oracle_set_cs_zeta(vpm_oracle_zeta_cb cb) { g_cs_zeta = cb; }
static void o_cs_zeta(double *P, int64_t nf, int64_t np, int kernel, int nthreads) {
  if (g_cs_zeta) g_cs_zeta(P, nf, np, kernel);
  else vpm_oracle_zeta_direct(P, nf, np, kernel, nthreads);
}

/* ===========================================================================
 * Time step on the host (SURVEY 8 f-1 checker).  ReformulatedVPM{f,g} only.
 *   rungekutta3            src/FLOWVPM_timeintegration.jl:388-461
 *   update_particle_states src/FLOWVPM_timeintegration.jl:463-534
 *   euler / _euler         src/FLOWVPM_timeintegration.jl:23-37,103-173
 *   relaxation             src/FLOWVPM_relaxation.jl:62-142
 *   ConstantSFS hook       src/FLOWVPM_subfilterscale.jl:110-135, clipping :287-296
 * ========================================================================== */
enum { R_M = 27, R_C = 36 };

static void o_stretch(const double *J, const double *G, int transposed, double *m) {
  if (transposed) {
    m[0] = J[0] * G[0] + J[1] * G[1] + J[2] * G[2];
    m[1] = J[3] * G[0] + J[4] * G[1] + J[5] * G[2];
    m[2] = J[6] * G[0] + J[7] * G[1] + J[8] * G[2];
  } else {
    m[0] = J[0] * G[0] + J[3] * G[1] + J[6] * G[2];
    m[1] = J[1] * G[0] + J[4] * G[1] + J[7] * G[2];
    m[2] = J[2] * G[0] + J[5] * G[1] + J[8] * G[2];
  }
}

static void o_relax(double *p, double rlxf, int kind) {
  const double *J = p + R_J;
  double *G = p + R_G;
  double nrmw = sqrt((J[5] - J[7]) * (J[5] - J[7]) + (J[6] - J[2]) * (J[6] - J[2]) + (J[1] - J[3]) * (J[1] - J[3]));
  if (nrmw != 0.0) {
    double nrmGamma = sqrt(G[0] * G[0] + G[1] * G[1] + G[2] * G[2]);
    double b2 = 1.0;
    if (kind == 2)
      b2 = 1 - 2 * (1 - rlxf) * rlxf *
                   (1 - (G[0] * (J[5] - J[7]) + G[1] * (J[6] - J[2]) + G[2] * (J[1] - J[3])) / (nrmGamma * nrmw));
    G[0] = (1 - rlxf) * G[0] + rlxf * nrmGamma * (J[5] - J[7]) / nrmw;
    G[1] = (1 - rlxf) * G[1] + rlxf * nrmGamma * (J[6] - J[2]) / nrmw;
    G[2] = (1 - rlxf) * G[2] + rlxf * nrmGamma * (J[1] - J[3]) / nrmw;
    if (kind == 2) {
      double sq = sqrt(b2);
      G[0] /= sq; G[1] /= sq; G[2] /= sq;
    }
  }
}

static void o_sfs_coeff(double *P, int64_t nf, int64_t np, double Cs, int clip) {
  for (int64_t i = 0; i < np; ++i) {
    double *p = P + nf * i;
    if (p[R_STATIC] != 0.0) continue;
    p[R_C] = Cs;
  }
  if (clip)
    for (int64_t i = 0; i < np; ++i) {
      double *p = P + nf * i;
      if (p[R_STATIC] != 0.0) continue;
      if (p[R_C] * (p[R_G] * p[R_SFS] + p[R_G + 1] * p[R_SFS + 1] + p[R_G + 2] * p[R_SFS + 2]) < 0) p[R_C] = 0;
    }
}

static void o_rates(const double *p, double f, double g, double zeta0, int transposed, double *MM, double *MM4,
                    double *eps) {
  const double *G = p + R_G;
  double C = p[R_C], sg = p[R_SIGMA];
  o_stretch(p + R_J, G, transposed, MM);
  double Gnorm2 = G[0] * G[0] + G[1] * G[1] + G[2] * G[2];
  if (Gnorm2 > 0) {
    *MM4 = (f + g) / (1 + 3 * f) * (MM[0] * G[0] + MM[1] * G[1] + MM[2] * G[2]);
    *MM4 -= f / (1 + 3 * f) * (C * p[R_SFS] * G[0] + C * p[R_SFS + 1] * G[1] + C * p[R_SFS + 2] * G[2]) *
            (sg * sg * sg) / zeta0;
    *MM4 /= Gnorm2;
  } else {
    *MM4 = 0;
  }
  for (int k = 0; k < 3; ++k) eps[k] = C * p[R_SFS + k] * (sg * sg * sg) / zeta0;
}

/* dynamicprocedure_pseudo3level_beforeUJ: src/FLOWVPM_subfilterscale.jl:447-539 */
static void o_dyn_before(double *P, int64_t nf, int64_t np, int kernel, int transposed, double alpha, int nthreads) {
  for (int64_t i = 0; i < np; ++i)
    if (P[nf * i + R_STATIC] == 0.0) P[nf * i + R_SIGMA] *= alpha;
  vpm_oracle_uj_direct(P, nf, np, kernel,
                       VPM_FLAG_RESET | VPM_FLAG_RESET_SFS | VPM_FLAG_SFS | (transposed ? VPM_FLAG_TRANSPOSED : 0),
                       nthreads);
  for (int64_t i = 0; i < np; ++i) {
    double *p = P + nf * i;
    if (p[R_STATIC] != 0.0) continue;
    double *M = p + R_M;
    for (int k = 0; k < 9; ++k) M[k] = 0.0;
    o_stretch(p + R_J, p + R_G, transposed, M);
    M[3] = p[R_SFS]; M[4] = p[R_SFS + 1]; M[5] = p[R_SFS + 2];
  }
  for (int64_t i = 0; i < np; ++i)
    if (P[nf * i + R_STATIC] == 0.0) P[nf * i + R_SIGMA] /= alpha;
}

/* dynamicprocedure_pseudo3level_afterUJ (:541-673) followed by the DynamicSFS clipping (:225-243) */
static int o_dyn_after(double *P, int64_t nf, int64_t np, int kernel, int transposed, double alpha, double rlxf,
                       double minC, double maxC, int force_positive, int clip) {
  const double zeta0 = zeta_fn(kernel, 0.0);
  for (int64_t i = 0; i < np; ++i) {
    double *p = P + nf * i;
    if (p[R_STATIC] != 0.0) continue;
    double *M = p + R_M, MM[3];
    o_stretch(p + R_J, p + R_G, transposed, MM);
    M[0] -= MM[0]; M[1] -= MM[1]; M[2] -= MM[2];
    M[3] -= p[R_SFS]; M[4] -= p[R_SFS + 1]; M[5] -= p[R_SFS + 2];
  }
  for (int64_t i = 0; i < np; ++i) {
    double *p = P + nf * i;
    if (p[R_STATIC] != 0.0) continue;
    double *M = p + R_M, *C_p = p + R_C;
    const double *Gamma = p + R_G;
    double nume = M[0] * Gamma[0] + M[1] * Gamma[1] + M[2] * Gamma[2];
    nume *= 3 * alpha - 2;
    double deno = M[3] * Gamma[0] + M[4] * Gamma[1] + M[5] * Gamma[2];
    deno /= zeta0 / (p[R_SIGMA] * p[R_SIGMA] * p[R_SIGMA]);
    if (C_p[2] == 0) {
      C_p[2] = deno;
      if (C_p[2] == 0) C_p[2] = 2.220446049250313e-16;
    }
    nume = rlxf * nume + (1 - rlxf) * C_p[1];
    deno = rlxf * deno + (1 - rlxf) * C_p[2];
    if (fabs(nume / deno) > maxC) {
      if (fabs(deno) < fabs(C_p[2])) deno = jl_sign(deno) * fabs(C_p[2]);
      if (fabs(nume / deno) >= maxC) nume = jl_sign(nume) * fabs(deno) * maxC;
    } else if (fabs(nume / deno) < minC) {
      nume = jl_sign(nume) * fabs(deno) * minC;
    }
    C_p[1] = nume;
    C_p[2] = deno;
    C_p[0] = C_p[1] / C_p[2];
    if (C_p[0] != C_p[0]) return -3;
    if (force_positive) C_p[0] *= jl_sign(C_p[0]);
  }
  for (int64_t i = 0; i < np; ++i)
    if (P[nf * i + R_STATIC] == 0.0)
      for (int k = 0; k < 9; ++k) P[nf * i + R_M + k] = 0.0;
  if (clip)
    for (int64_t i = 0; i < np; ++i) {
      double *p = P + nf * i;
      if (p[R_STATIC] != 0.0) continue;
      if (p[R_C] * (p[R_G] * p[R_SFS] + p[R_G + 1] * p[R_SFS + 1] + p[R_G + 2] * p[R_SFS + 2]) < 0) p[R_C] *= 0;
    }
  return 0;
}

/* SFS control strategies after the clippings: control_directional (src/FLOWVPM_subfilterscale.jl:319-334),
 * control_magnitude (:367-397), each applied to every non-static particle in turn (:245-265) */
static void o_controls(double *P, int64_t nf, int64_t np, int controls, double f, double zeta0, double deltat) {
  if (controls & 1)
    for (int64_t i = 0; i < np; ++i) {
      double *p = P + nf * i;
      if (p[R_STATIC] != 0.0) continue;
      double G1 = p[R_G], G2 = p[R_G + 1], G3 = p[R_G + 2], S1 = p[R_SFS], S2 = p[R_SFS + 1], S3 = p[R_SFS + 2];
      double aux = S1 * G1 + S2 * G2 + S3 * G3;
      aux /= (G1 * G1 + G2 * G2 + G3 * G3);
      p[R_SFS] = aux * G1; p[R_SFS + 1] = aux * G2; p[R_SFS + 2] = aux * G3;
    }
  if ((controls & 2) && deltat > 0)
    for (int64_t i = 0; i < np; ++i) {
      double *p = P + nf * i;
      if (p[R_STATIC] != 0.0) continue;
      double C = p[R_C];
      if (C != 0) {
        double G1 = p[R_G], G2 = p[R_G + 1], G3 = p[R_G + 2], S1 = p[R_SFS], S2 = p[R_SFS + 1], S3 = p[R_SFS + 2];
        double aux = S1 * G1 + S2 * G2 + S3 * G3;
        aux /= (G1 * G1 + G2 * G2 + G3 * G3);
        aux -= (1 + 3 * f) * (zeta0 / (p[R_SIGMA] * p[R_SIGMA] * p[R_SIGMA])) / deltat / C;
        if (aux > 0) { p[R_SFS] = -aux * G1; p[R_SFS + 1] = -aux * G2; p[R_SFS + 2] = -aux * G3; }
      }
    }
}

/*
 * rbf_conjugategradient(pfield, cs): src/FLOWVPM_viscous.jl:309-478 with cs.zeta = zeta_direct.
 * Target vorticity in M[7:9], solution built in M[1:3], residual in M[4:6], search direction in
 * Gamma, basis evaluation in J[1:3]; updates over iterator(pfield) (non-static), zeta over all.
 * Returns the number of CG iterations, or -4 when itmax is reached without convergence and
 * iterror is set (the reference throws).  info[0:3] = final sqrt(rrs/rr0s).
 */
enum { R_VOL = 7 };
int vpm_oracle_rbf_cg(double *P, int64_t nf, int64_t np, int kernel, int itmax, double tol, int iterror,
                      double *info, int nthreads) {
  init_consts();
  const double eps = 2.220446049250313e-16;
  double rr0s[3] = {0, 0, 0}, rrs[3], prev_rrs[3], pAps[3], alphas[3], betas[3];
  int flags[3];
  for (int64_t i = 0; i < np; ++i) {
    double *p = P + nf * i;
    if (p[R_STATIC] != 0.0) continue;
    for (int k = 0; k < 3; ++k) {
      p[R_M + k] = p[R_M + 6 + k] * p[R_VOL];
      p[R_G + k] = p[R_M + k];
    }
  }
  o_cs_zeta(P, nf, np, kernel, nthreads);
  for (int64_t i = 0; i < np; ++i) {
    double *p = P + nf * i;
    if (p[R_STATIC] != 0.0) continue;
    for (int k = 0; k < 3; ++k) {
      p[R_M + 3 + k] = p[R_M + 6 + k] - p[R_J + k];
      p[R_G + k] = p[R_M + 3 + k];
      rr0s[k] += p[R_M + 3 + k] * p[R_M + 3 + k];
    }
  }
  for (int k = 0; k < 3; ++k) {
    rrs[k] = rr0s[k];
    flags[k] = sqrt(rr0s[k]) > tol || sqrt(rrs[k] / rr0s[k]) > tol;
  }
  int it_done = 0, failed = 0;
  for (int it = 1; it <= itmax; ++it) {
    if (!(flags[0] || flags[1] || flags[2])) break;
    it_done = it;
    o_cs_zeta(P, nf, np, kernel, nthreads);
    for (int k = 0; k < 3; ++k) pAps[k] = 0;
    for (int64_t i = 0; i < np; ++i) {
      const double *p = P + nf * i;
      if (p[R_STATIC] != 0.0) continue;
      for (int k = 0; k < 3; ++k) pAps[k] += p[R_G + k] * p[R_J + k];
    }
    for (int k = 0; k < 3; ++k) {
      alphas[k] = flags[k] ? rrs[k] / pAps[k] : 0.0; /* Julia: x * false == 0 even for NaN (strong zero) */
      prev_rrs[k] = rrs[k];
      rrs[k] = 0;
    }
    for (int64_t i = 0; i < np; ++i) {
      double *p = P + nf * i;
      if (p[R_STATIC] != 0.0) continue;
      for (int k = 0; k < 3; ++k) {
        p[R_M + k] += alphas[k] * p[R_G + k];
        p[R_M + 3 + k] -= alphas[k] * p[R_J + k];
        rrs[k] += p[R_M + 3 + k] * p[R_M + 3 + k];
      }
    }
    for (int k = 0; k < 3; ++k) {
      betas[k] = rrs[k] / prev_rrs[k];
      if (fabs(prev_rrs[k]) <= 2 * eps) betas[k] = 1;
    }
    for (int64_t i = 0; i < np; ++i) {
      double *p = P + nf * i;
      if (p[R_STATIC] != 0.0) continue;
      for (int k = 0; k < 3; ++k) p[R_G + k] = p[R_M + 3 + k] + betas[k] * p[R_G + k];
    }
    for (int k = 0; k < 3; ++k)
      flags[k] = flags[k] && (fabs(rr0s[k]) <= 2 * eps ? 0 : sqrt(rrs[k] / rr0s[k]) > tol);
    if (it == itmax && (flags[0] || flags[1] || flags[2])) failed = 1;
  }
  for (int64_t i = 0; i < np; ++i) {
    double *p = P + nf * i;
    if (p[R_STATIC] != 0.0) continue;
    for (int k = 0; k < 3; ++k) p[R_G + k] = p[R_M + k];
  }
  if (info)
    for (int k = 0; k < 3; ++k) info[k] = rr0s[k] > 0 ? sqrt(rrs[k] / rr0s[k]) : 0.0;
  if (failed && iterror) return -4;
  return it_done;
}

/* viscousdiffusion(pfield, CoreSpreading, dt; aux1, aux2): src/FLOWVPM_viscous.jl:152-223.
 * vis[0..4] = nu, sgm0, beta, tol, t_sgm (in/out); returns < 0 on RBF failure. */
static int o_corespreading(double *P, int64_t nf, int64_t np, int kernel, int integration, double dt, double aux1,
                           double aux2, double *vis, int itmax, int iterror, int nthreads) {
  const double nu = vis[0], sgm0 = vis[1], beta = vis[2], tol = vis[3];
  int proceed = 0;
  if (integration == 0) {
    for (int64_t i = 0; i < np; ++i) {
      double *p = P + nf * i;
      if (p[R_STATIC] != 0.0) continue;
      p[R_SIGMA] = sqrt(p[R_SIGMA] * p[R_SIGMA] + 2 * nu * dt);
    }
    proceed = 1;
  } else {
    for (int64_t i = 0; i < np; ++i) {
      double *p = P + nf * i;
      if (p[R_STATIC] != 0.0) continue;
      p[R_M + 6] = aux1 * p[R_M + 6] + dt * 2 * nu;
      p[R_SIGMA] = sqrt(p[R_SIGMA] * p[R_SIGMA] + aux2 * p[R_M + 6]);
    }
    if (fabs(aux2 - 8.0 / 15) <= 1e-7) proceed = 1;
  }
  if (proceed) {
    vis[4] += dt;
    double beta_cur = sqrt(2 * nu * vis[4] / (sgm0 * sgm0) + 1);
    if (beta_cur >= beta) {
      o_cs_zeta(P, nf, np, kernel, nthreads);
      for (int64_t i = 0; i < np; ++i) {
        double *p = P + nf * i;
        if (p[R_STATIC] != 0.0) continue;
        for (int k = 0; k < 3; ++k) p[R_M + 6 + k] = p[R_J + k];
        p[R_SIGMA] = sgm0;
      }
      int rc = vpm_oracle_rbf_cg(P, nf, np, kernel, itmax, tol, iterror, NULL, nthreads);
      if (rc < 0) return rc;
      vis[4] = 0;
    }
  }
  return 0;
}

/* viscousdiffusion(pfield, ParticleStrengthExchange, dt; aux1, aux2): src/FLOWVPM_viscous.jl:257-298 (the
 * `pfield.UJ != UJ_fmm` error of :259-262 is the caller's business: this is the per-particle part) */
static void o_pse(double *P, int64_t nf, int64_t np, int integration, double dt, double aux2, double nu,
                  int recalculate_vols) {
  const double pi = 3.14159265358979323846;
  for (int64_t i = 0; i < np; ++i) {
    double *p = P + nf * i;
    if (p[R_STATIC] != 0.0) continue;
    if (recalculate_vols) p[7] = 4.0 / 3.0 * pi * (p[R_SIGMA] * p[R_SIGMA] * p[R_SIGMA]); /* 4/3*pi*sigma^3 */
    for (int k = 0; k < 3; ++k) {
      if (integration == 0) {
        p[R_G + k] += dt * nu * p[R_PSE + k];
      } else {
        p[R_M + 3 + k] += dt * nu * p[R_PSE + k];
        p[R_G + k] += aux2 * dt * nu * p[R_PSE + k];
      }
    }
  }
}

int vpm_oracle_field_step(double *P, int64_t nf, int64_t np, double *dp, const int *ip, int nthreads) {
  init_consts();
  const int viscous = ip[9], itmax = ip[10], iterror = ip[11];
  double *vis = dp + 13; /* nu, sgm0, beta, tol, t_sgm (in/out) */
  const double dt = dp[0], f = dp[1], g = dp[2], Uinf[3] = {dp[3], dp[4], dp[5]}, Cs = dp[6], rlxf = dp[7];
  const double alpha = dp[8], sfs_rlxf = dp[9], minC = dp[10], maxC = dp[11], deltat = dp[12];
  const int controls = ip[8];
  const int kernel = ip[0], integration = ip[1], relaxation = ip[2], relax = ip[3], sfs = ip[4], clip = ip[5],
            transposed = ip[6], force_positive = ip[7];
  const double zeta0 = zeta_fn(kernel, 0.0);
  const int tr = transposed ? VPM_FLAG_TRANSPOSED : 0;
  const int uj_flags = VPM_FLAG_RESET | tr | (sfs ? (VPM_FLAG_SFS | VPM_FLAG_RESET_SFS) : 0);
  if (integration == 0) {
    if (sfs == 2) o_dyn_before(P, nf, np, kernel, transposed, alpha, nthreads);
    vpm_oracle_uj_direct(P, nf, np, kernel, uj_flags, nthreads);
    if (sfs == 1) o_sfs_coeff(P, nf, np, Cs, clip);
    if (sfs == 2 && o_dyn_after(P, nf, np, kernel, transposed, alpha, sfs_rlxf, minC, maxC, force_positive, clip))
      return -3;
    if (sfs) o_controls(P, nf, np, controls, f, zeta0, deltat);
    for (int64_t i = 0; i < np; ++i) {
      double *p = P + nf * i;
      if (p[R_STATIC] != 0.0) continue;
      double *G = p + R_G, MM[3], MM4, eps[3];
      for (int k = 0; k < 3; ++k) p[k] += dt * (p[R_U + k] + Uinf[k]);
      o_rates(p, f, g, zeta0, transposed, MM, &MM4, eps);
      for (int k = 0; k < 3; ++k) G[k] += dt * (MM[k] - 3 * MM4 * G[k] - eps[k]);
      p[R_SIGMA] -= dt * (p[R_SIGMA] * MM4);
      if (relax && relaxation) o_relax(p, rlxf, relaxation);
    }
    if (viscous == 1) return o_corespreading(P, nf, np, kernel, 0, dt, 0.0, 0.0, vis, itmax, iterror, nthreads);
    if (viscous >= 2) o_pse(P, nf, np, 0, dt, 0.0, vis[0], viscous == 2);
    return 0;
  }
  for (int64_t i = 0; i < np; ++i)
    if (P[nf * i + R_STATIC] == 0.0)
      for (int k = 0; k < 9; ++k) P[nf * i + R_M + k] = 0.0;
  const double ab[3][2] = {{0.0, 1.0 / 3}, {-5.0 / 9, 15.0 / 16}, {-153.0 / 128, 8.0 / 15}};
  for (int s_ = 0; s_ < 3; ++s_) {
    const double a = ab[s_][0], b = ab[s_][1];
    if (sfs == 2 && a == 0.0) o_dyn_before(P, nf, np, kernel, transposed, alpha, nthreads);
    vpm_oracle_uj_direct(P, nf, np, kernel, uj_flags, nthreads);
    if (sfs == 1 && a == 0.0) o_sfs_coeff(P, nf, np, Cs, clip);
    if (sfs == 2 && a == 0.0 &&
        o_dyn_after(P, nf, np, kernel, transposed, alpha, sfs_rlxf, minC, maxC, force_positive, clip))
      return -3;
    if (sfs && a == 0.0) o_controls(P, nf, np, controls, f, zeta0, deltat);
    for (int64_t i = 0; i < np; ++i) {
      double *p = P + nf * i;
      if (p[R_STATIC] != 0.0) continue;
      double *M = p + R_M, *G = p + R_G, MM[3], MM4, eps[3];
      for (int k = 0; k < 3; ++k) M[k] = a * M[k] + dt * (p[R_U + k] + Uinf[k]);
      for (int k = 0; k < 3; ++k) p[k] += b * M[k];
      o_rates(p, f, g, zeta0, transposed, MM, &MM4, eps);
      for (int k = 0; k < 3; ++k) M[3 + k] = a * M[3 + k] + dt * (MM[k] - 3 * MM4 * G[k] - eps[k]);
      M[7] = a * M[7] - dt * (p[R_SIGMA] * MM4);
      for (int k = 0; k < 3; ++k) G[k] += b * M[3 + k];
      p[R_SIGMA] += b * M[7];
    }
    if (viscous == 1) {
      int rc = o_corespreading(P, nf, np, kernel, 1, dt, a, b, vis, itmax, iterror, nthreads);
      if (rc < 0) return rc;
    } else if (viscous >= 2) {
      o_pse(P, nf, np, 1, dt, b, vis[0], viscous == 2);
    }
  }
  if (relax && relaxation) {
    vpm_oracle_uj_direct(P, nf, np, kernel, VPM_FLAG_RESET | tr, nthreads);
    for (int64_t i = 0; i < np; ++i)
      if (P[nf * i + R_STATIC] == 0.0) o_relax(P + nf * i, rlxf, relaxation);
  }
  return 0;
}

/*
 * Timing helper for bench.py's cpu_baseline / --impl reference legs: all ns
 * sources against the target slice [t0,t1) on `nthreads` threads, returning
 * nothing but the accumulated buffer (the caller times the call).
 */
int vpm_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
/* cores this process may run on, whatever OMP_NUM_THREADS says (torchrun exports
 * OMP_NUM_THREADS=1 to its workers; the CPU baseline must still use the whole host) */
int vpm_oracle_num_procs(void) {
#ifdef _OPENMP
  return omp_get_num_procs();
#else
  return 1;
#endif
}
