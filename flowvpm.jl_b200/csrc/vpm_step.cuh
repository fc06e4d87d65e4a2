// vpm_step.cuh -- SURVEY 8(f-1): the O(N) per-particle kernels around the sweeps, so that a
// whole time step can run on a device-resident copy of ParticleField.particles.
//
// Reference being restated (FLOWVPM.jl v4.0.3):
//   rungekutta3 / update_particle_states for ReformulatedVPM{f,g}  src/FLOWVPM_timeintegration.jl:388-534
//   euler / _euler for ReformulatedVPM{f,g}                        src/FLOWVPM_timeintegration.jl:23-37,103-173
//   relax_pedrizzetti / relax_correctedpedrizzetti                 src/FLOWVPM_relaxation.jl:62-142
//   ConstantSFS AfterUJ hook + clipping_backscatter                src/FLOWVPM_subfilterscale.jl:110-135,287-296
//   DynamicSFS pseudo-3-level procedure (before / after UJ)         src/FLOWVPM_subfilterscale.jl:447-673
// Covered: cVPM / rVPM / any (f, g); NoSFS, ConstantSFS and DynamicSFS (pseudo3level, optional
// force_positive, backscatter clipping, directional / magnitude controls); Inviscid; constant
// Uinf.  Not covered (stays in Julia): the (stale) sensor-function procedure, viscous schemes.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace vpm {

struct StepArgs {
  double *P;  // device mirror of ParticleField.particles, nf x np column-major
  int64_t nf, np;
  double a, b, dt;
  double Ux, Uy, Uz;  // Uinf
  double f, g, zeta0;
  double Cs, rlxf;
  double alpha, sfs_rlxf, minC, maxC;  // DynamicSFS (src/FLOWVPM_subfilterscale.jl:167-202)
  int transposed, sfs, clip, relax_kind;  // sfs: 0 none, 1 constant, 2 dynamic; relax_kind: 0 none, 1 pedrizzetti, 2 corrected
  int force_positive;
  int controls;     // bit 0: control_directional, bit 1: control_magnitude (applied in that order)
  double deltat;    // pfield.t / pfield.nt for control_magnitude (<= 0: pfield.nt == 0, control skipped)
  int *nan_flag;  // set when the dynamic procedure produces a NaN coefficient (:645-652)
};

// rows, 0-based (src/FLOWVPM_particlefield.jl:239-252)
enum { S_X = 0, S_G = 3, S_SIGMA = 6, S_U = 9, S_J = 15, S_M = 27, S_C = 36, S_SFS = 39, S_STATIC = 42 };

__device__ __forceinline__ void stretching(const double *J, const double *G, int transposed, double &m1,
                                           double &m2, double &m3) {
  if (transposed) {
    m1 = J[0] * G[0] + J[1] * G[1] + J[2] * G[2];
    m2 = J[3] * G[0] + J[4] * G[1] + J[5] * G[2];
    m3 = J[6] * G[0] + J[7] * G[1] + J[8] * G[2];
  } else {
    m1 = J[0] * G[0] + J[3] * G[1] + J[6] * G[2];
    m2 = J[1] * G[0] + J[4] * G[1] + J[7] * G[2];
    m3 = J[2] * G[0] + J[5] * G[1] + J[8] * G[2];
  }
}

// relaxation(pfield, i): src/FLOWVPM_relaxation.jl:62-81 (kind 1), :117-142 (kind 2)
__device__ __forceinline__ void relax_particle(double *p, double rlxf, int kind) {
  const double *J = p + S_J;
  double *G = p + S_G;
  const double w1 = J[5] - J[7], w2 = J[6] - J[2], w3 = J[1] - J[3];
  const double nrmw = sqrt(w1 * w1 + w2 * w2 + w3 * w3);
  if (nrmw == 0.0) return;
  const double nrmG = sqrt(G[0] * G[0] + G[1] * G[1] + G[2] * G[2]);
  double b2 = 1.0;
  if (kind == 2) b2 = 1 - 2 * (1 - rlxf) * rlxf * (1 - (G[0] * w1 + G[1] * w2 + G[2] * w3) / (nrmG * nrmw));
  G[0] = (1 - rlxf) * G[0] + rlxf * nrmG * w1 / nrmw;
  G[1] = (1 - rlxf) * G[1] + rlxf * nrmG * w2 / nrmw;
  G[2] = (1 - rlxf) * G[2] + rlxf * nrmG * w3 / nrmw;
  if (kind == 2) {
    const double s = sqrt(b2);
    G[0] /= s; G[1] /= s; G[2] /= s;
  }
}

// zero the RK storage M of non-static particles (src/FLOWVPM_timeintegration.jl:402-414)
__global__ void step_reset_M(StepArgs a) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.np) return;
  double *p = a.P + i * a.nf;
  if (p[S_STATIC] != 0.0) return;
#pragma unroll
  for (int k = 0; k < 9; ++k) p[S_M + k] = 0.0;
}

// ConstantSFS AfterUJ at an Euler step or the first RK substep: C <- Cs, then backscatter
// clipping C <- 0 where C (Gamma . SFS) < 0 (src/FLOWVPM_subfilterscale.jl:110-135,287-296)
__global__ void step_sfs_coeff(StepArgs a) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.np) return;
  double *p = a.P + i * a.nf;
  if (p[S_STATIC] != 0.0) return;
  double C = a.Cs;
  if (a.clip) {
    const double d = p[S_G] * p[S_SFS] + p[S_G + 1] * p[S_SFS + 1] + p[S_G + 2] * p[S_SFS + 2];
    if (C * d < 0.0) C = 0.0;
  }
  p[S_C] = C;
}

// SFS control strategies, applied after the clippings at an Euler step / the first RK substep
// (src/FLOWVPM_subfilterscale.jl:121-155,245-265): control_directional (:319-334) keeps only
// the component of SFS along Gamma; control_magnitude (:367-397) limits forward scatter.
__global__ void step_sfs_controls(StepArgs a) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.np) return;
  double *p = a.P + i * a.nf;
  if (p[S_STATIC] != 0.0) return;
  const double G1 = p[S_G], G2 = p[S_G + 1], G3 = p[S_G + 2];
  if (a.controls & 1) {
    const double S1 = p[S_SFS], S2 = p[S_SFS + 1], S3 = p[S_SFS + 2];
    double aux = S1 * G1 + S2 * G2 + S3 * G3;
    aux /= (G1 * G1 + G2 * G2 + G3 * G3);
    p[S_SFS] = aux * G1; p[S_SFS + 1] = aux * G2; p[S_SFS + 2] = aux * G3;
  }
  if ((a.controls & 2) && a.deltat > 0.0) {
    const double C = p[S_C];
    if (C != 0.0) {
      const double S1 = p[S_SFS], S2 = p[S_SFS + 1], S3 = p[S_SFS + 2], sg = p[S_SIGMA];
      double aux = S1 * G1 + S2 * G2 + S3 * G3;
      aux /= (G1 * G1 + G2 * G2 + G3 * G3);
      aux -= (1 + 3 * a.f) * (a.zeta0 / (sg * sg * sg)) / a.deltat / C;
      if (aux > 0.0) { p[S_SFS] = -aux * G1; p[S_SFS + 1] = -aux * G2; p[S_SFS + 2] = -aux * G3; }
    }
  }
}

// ---- DynamicSFS, pseudo-3-level procedure -------------------------------------------
// sigma *= alpha (test filter) / sigma /= alpha (back to the domain filter), non-static only
// (src/FLOWVPM_subfilterscale.jl:464-475, 525-536)
__global__ void step_scale_sigma(StepArgs a, int divide) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.np) return;
  double *p = a.P + i * a.nf;
  if (p[S_STATIC] != 0.0) return;
  if (divide) p[S_SIGMA] /= a.alpha; else p[S_SIGMA] *= a.alpha;
}

// after the test-filter UJ: M <- 0, M[1:3] <- stretching, M[4:6] <- SFS (:480-521)
__global__ void step_dyn_store(StepArgs a) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.np) return;
  double *p = a.P + i * a.nf;
  if (p[S_STATIC] != 0.0) return;
  double *M = p + S_M;
#pragma unroll
  for (int k = 0; k < 9; ++k) M[k] = 0.0;
  stretching(p + S_J, p + S_G, a.transposed, M[0], M[1], M[2]);
  M[3] = p[S_SFS]; M[4] = p[S_SFS + 1]; M[5] = p[S_SFS + 2];
}

__device__ __forceinline__ double jl_sign(double x) { return x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : x); }

// after the domain-filter UJ: differences, Lagrangian-averaged C = <Gamma.L>/<Gamma.m> with
// clamps, M flushed (:556-670)
__global__ void step_dyn_coeff(StepArgs a) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.np) return;
  double *p = a.P + i * a.nf;
  if (p[S_STATIC] != 0.0) return;
  double *M = p + S_M, *Cp = p + S_C;
  const double *G = p + S_G;
  double m1, m2, m3;
  stretching(p + S_J, G, a.transposed, m1, m2, m3);
  M[0] -= m1; M[1] -= m2; M[2] -= m3;
  M[3] -= p[S_SFS]; M[4] -= p[S_SFS + 1]; M[5] -= p[S_SFS + 2];
  double nume = M[0] * G[0] + M[1] * G[1] + M[2] * G[2];
  nume *= 3 * a.alpha - 2;
  double deno = M[3] * G[0] + M[4] * G[1] + M[5] * G[2];
  const double sg = p[S_SIGMA];
  deno /= a.zeta0 / (sg * sg * sg);
  if (Cp[2] == 0.0) {
    Cp[2] = deno;
    if (Cp[2] == 0.0) Cp[2] = 2.220446049250313e-16;  // eps()
  }
  nume = a.sfs_rlxf * nume + (1 - a.sfs_rlxf) * Cp[1];
  deno = a.sfs_rlxf * deno + (1 - a.sfs_rlxf) * Cp[2];
  if (fabs(nume / deno) > a.maxC) {
    if (fabs(deno) < fabs(Cp[2])) deno = jl_sign(deno) * fabs(Cp[2]);
    if (fabs(nume / deno) >= a.maxC) nume = jl_sign(nume) * fabs(deno) * a.maxC;
  } else if (fabs(nume / deno) < a.minC) {
    nume = jl_sign(nume) * fabs(deno) * a.minC;
  }
  Cp[1] = nume;
  Cp[2] = deno;
  double C = Cp[1] / Cp[2];
  if (C != C) atomicExch(a.nan_flag, 1);
  if (a.force_positive) C *= jl_sign(C);
  if (a.clip) {
    const double d = G[0] * p[S_SFS] + G[1] * p[S_SFS + 1] + G[2] * p[S_SFS + 2];
    if (C * d < 0.0) C *= 0.0;
  }
  Cp[0] = C;
#pragma unroll
  for (int k = 0; k < 9; ++k) M[k] = 0.0;
}

// the Z and dGamma terms shared by the Euler and RK updates (:505-521, :145-160)
__device__ __forceinline__ void rvpm_rates(const double *p, const StepArgs &a, double &m1, double &m2, double &m3,
                                           double &m4, double &e1, double &e2, double &e3) {
  const double *G = p + S_G, *SFS = p + S_SFS;
  const double C = p[S_C], sg = p[S_SIGMA];
  stretching(p + S_J, G, a.transposed, m1, m2, m3);
  const double sigma3 = sg * sg * sg;
  const double Gn2 = G[0] * G[0] + G[1] * G[1] + G[2] * G[2];
  if (Gn2 > 0.0) {
    m4 = (a.f + a.g) / (1 + 3 * a.f) * (m1 * G[0] + m2 * G[1] + m3 * G[2]);
    m4 -= a.f / (1 + 3 * a.f) * (C * SFS[0] * G[0] + C * SFS[1] * G[1] + C * SFS[2] * G[2]) * sigma3 / a.zeta0;
    m4 /= Gn2;
  } else {
    m4 = 0.0;
  }
  e1 = C * SFS[0] * sigma3 / a.zeta0;
  e2 = C * SFS[1] * sigma3 / a.zeta0;
  e3 = C * SFS[2] * sigma3 / a.zeta0;
}

// update_particle_states, ReformulatedVPM: src/FLOWVPM_timeintegration.jl:463-534
__global__ void step_rk_stage(StepArgs a) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.np) return;
  double *p = a.P + i * a.nf;
  if (p[S_STATIC] != 0.0) return;
  double *M = p + S_M, *G = p + S_G, *X = p + S_X;
  const double *U = p + S_U;
  M[0] = a.a * M[0] + a.dt * (U[0] + a.Ux);
  M[1] = a.a * M[1] + a.dt * (U[1] + a.Uy);
  M[2] = a.a * M[2] + a.dt * (U[2] + a.Uz);
  X[0] += a.b * M[0];
  X[1] += a.b * M[1];
  X[2] += a.b * M[2];
  double m1, m2, m3, m4, e1, e2, e3;
  rvpm_rates(p, a, m1, m2, m3, m4, e1, e2, e3);
  M[3] = a.a * M[3] + a.dt * (m1 - 3 * m4 * G[0] - e1);
  M[4] = a.a * M[4] + a.dt * (m2 - 3 * m4 * G[1] - e2);
  M[5] = a.a * M[5] + a.dt * (m3 - 3 * m4 * G[2] - e3);
  M[7] = a.a * M[7] - a.dt * (p[S_SIGMA] * m4);
  G[0] += a.b * M[3];
  G[1] += a.b * M[4];
  G[2] += a.b * M[5];
  p[S_SIGMA] += a.b * M[7];
}

// _euler, ReformulatedVPM: src/FLOWVPM_timeintegration.jl:103-173 (relaxation inside the loop)
__global__ void step_euler(StepArgs a, int relax) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.np) return;
  double *p = a.P + i * a.nf;
  if (p[S_STATIC] != 0.0) return;
  double *G = p + S_G, *X = p + S_X;
  const double *U = p + S_U;
  X[0] += a.dt * (U[0] + a.Ux);
  X[1] += a.dt * (U[1] + a.Uy);
  X[2] += a.dt * (U[2] + a.Uz);
  double m1, m2, m3, m4, e1, e2, e3;
  rvpm_rates(p, a, m1, m2, m3, m4, e1, e2, e3);
  G[0] += a.dt * (m1 - 3 * m4 * G[0] - e1);
  G[1] += a.dt * (m2 - 3 * m4 * G[1] - e2);
  G[2] += a.dt * (m3 - 3 * m4 * G[2] - e3);
  p[S_SIGMA] -= a.dt * (p[S_SIGMA] * m4);
  if (relax && a.relax_kind) relax_particle(p, a.rlxf, a.relax_kind);
}

__global__ void step_relax(StepArgs a) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.np) return;
  double *p = a.P + i * a.nf;
  if (p[S_STATIC] != 0.0) return;
  relax_particle(p, a.rlxf, a.relax_kind);
}

// ---- CoreSpreading viscous scheme (src/FLOWVPM_viscous.jl:152-223) and the RBF conjugate-
// gradient re-discretisation it triggers (:309-478).  The scalar recurrences of the CG
// (alphas, betas, convergence flags) run on the host; these are its O(N) pieces.
enum { S_VOL = 7 };

// sigma <- sqrt(sigma^2 + 2 nu dt) (Euler) / the low-storage RK form with M[7] (:159-175)
__global__ void cs_spread(StepArgs a, double nu, int rk) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.np) return;
  double *p = a.P + i * a.nf;
  if (p[S_STATIC] != 0.0) return;
  if (rk) {
    p[S_M + 6] = a.a * p[S_M + 6] + a.dt * 2 * nu;
    p[S_SIGMA] = sqrt(p[S_SIGMA] * p[S_SIGMA] + a.b * p[S_M + 6]);
  } else {
    p[S_SIGMA] = sqrt(p[S_SIGMA] * p[S_SIGMA] + 2 * nu * a.dt);
  }
}

// ParticleStrengthExchange, the per-particle part of viscousdiffusion (src/FLOWVPM_viscous.jl:257-298):
// optionally vol <- 4/3 pi sigma^3, then Gamma += dt nu PSE (Euler) resp. M[4:6] += dt nu PSE,
// Gamma += aux2 dt nu PSE (RK3), over the non-static particles.  In v4.0.3 nothing ever accumulates
// into the PSE rows 25:27 (src/FLOWVPM_particlefield.jl:482-489 only zero them), so the strength update
// adds zeros for every particle that has been reset; it is restated as written.
enum { S_PSE = 24 };
__global__ void pse_update(StepArgs a, double nu, int rk, int recalculate_vols) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.np) return;
  double *p = a.P + i * a.nf;
  if (p[S_STATIC] != 0.0) return;
  if (recalculate_vols) {
    const double sg = p[S_SIGMA];
    p[S_VOL] = __dmul_rn(__dmul_rn(4.0 / 3.0, 3.14159265358979323846), __dmul_rn(__dmul_rn(sg, sg), sg));
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double q = __dmul_rn(__dmul_rn(a.dt, nu), p[S_PSE + k]);
    if (rk) {
      p[S_M + 3 + k] = __dadd_rn(p[S_M + 3 + k], q);
      p[S_G + k] = __dadd_rn(p[S_G + k], __dmul_rn(__dmul_rn(__dmul_rn(a.b, a.dt), nu), p[S_PSE + k]));
    } else {
      p[S_G + k] = __dadd_rn(p[S_G + k], q);
    }
  }
}

// overgrown cores: target vorticity M[7:9] <- J[1:3] (basis evaluation), sigma <- sgm0 (:205-212)
__global__ void cs_reset(StepArgs a, double sgm0) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.np) return;
  double *p = a.P + i * a.nf;
  if (p[S_STATIC] != 0.0) return;
#pragma unroll
  for (int k = 0; k < 3; ++k) p[S_M + 6 + k] = p[S_J + k];
  p[S_SIGMA] = sgm0;
}

// stage 0: initial guess Gamma = omega_targ * vol (:334-341); stage 1: r0 = omega_targ - omega_cur,
// p0 = r0 (:346-357); stage 2: x += alpha p, r -= alpha A p (:392-398); stage 3: p = r + beta p
// (:410-414); stage 4: Gamma <- solution (:449-453)
__global__ void rbf_stage(StepArgs a, int stage, double c0, double c1, double c2) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.np) return;
  double *p = a.P + i * a.nf;
  if (p[S_STATIC] != 0.0) return;
  double *M = p + S_M, *G = p + S_G;
  const double *J = p + S_J;
  const double c[3] = {c0, c1, c2};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (stage == 0) { M[k] = M[6 + k] * p[S_VOL]; G[k] = M[k]; }
    else if (stage == 1) { M[3 + k] = M[6 + k] - J[k]; G[k] = M[3 + k]; }
    else if (stage == 2) { M[k] += c[k] * G[k]; M[3 + k] -= c[k] * J[k]; }
    else if (stage == 3) { G[k] = M[3 + k] + c[k] * G[k]; }
    else { G[k] = M[k]; }
  }
}

// deterministic 3-component reductions over the non-static particles:
// mode 0: sum r_k^2 (r = M[4:6]); mode 1: sum Gamma_k J_k (the pAp product).
// Fixed launch shape (kRedBlocks x 256) and fixed tree order => run-to-run identical.
constexpr int kRedBlocks = 256;
__global__ void rbf_reduce_partial(StepArgs a, int mode, double *partial) {
  __shared__ double sh[3][256];
  double acc[3] = {0.0, 0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.np; i += (int64_t)gridDim.x * blockDim.x) {
    const double *p = a.P + i * a.nf;
    if (p[S_STATIC] != 0.0) continue;
#pragma unroll
    for (int k = 0; k < 3; ++k)
      acc[k] += mode == 0 ? p[S_M + 3 + k] * p[S_M + 3 + k] : p[S_G + k] * p[S_J + k];
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) sh[k][threadIdx.x] = acc[k];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s)
#pragma unroll
      for (int k = 0; k < 3; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x < 3) partial[blockIdx.x * 3 + threadIdx.x] = sh[threadIdx.x][0];
}
__global__ void rbf_reduce_final(const double *partial, double *out) {
  __shared__ double sh[3][kRedBlocks];
#pragma unroll
  for (int k = 0; k < 3; ++k) sh[k][threadIdx.x] = partial[threadIdx.x * 3 + k];
  __syncthreads();
  for (int s = kRedBlocks / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s)
#pragma unroll
      for (int k = 0; k < 3; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x < 3) out[threadIdx.x] = sh[threadIdx.x][0];
}

}  // namespace vpm
