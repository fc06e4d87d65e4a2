// vpm_host_field.cuh -- device-resident particle matrix: UJ, zeta, RBF, CoreSpreading (f-1).
// Part of the single translation unit vpm_abi.cu (included there in order; not a standalone header).
#pragma once
namespace {

// ---- device-resident field: UJ_direct on the mirror of the whole matrix ----
// With G devices every device holds the whole mirror (np_pad = G * shard columns); device g
// sweeps the targets of its shard and the shards' columns are all-gathered in place over
// NVLink (columns = particles are contiguous in the column-major matrix), so all mirrors
// stay identical and the O(N) step kernels simply run on every device.
int64_t field_shard(const vpm_handle *h) {
  const int64_t G = (int64_t)h->devs.size();
  return (h->fld_np + G - 1) / G;
}

int field_allgather(vpm_handle *h) {
  const int G = (int)h->devs.size();
  if (G < 2) return VPM_OK;
  TRY(ensure_comms(h));
  const size_t count = (size_t)field_shard(h) * h->fld_nf;
  NCK(h, g_nccl.group_start());
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    double *F = (double *)d.fld.p;
    NCK(h, g_nccl.all_gather(F + (size_t)g * count, F, count, kNcclFloat64, h->comms[g], d.stream));
  }
  NCK(h, g_nccl.group_end());
  return VPM_OK;
}

int field_uj(vpm_handle *h, int kernel, int flags) {
  const int64_t nf = h->fld_nf, np = h->fld_np;
  if (np == 0) return VPM_OK;
  const int G = (int)h->devs.size();
  const int64_t shard = field_shard(h);
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    cudaStream_t st = d.stream;
    CK(h, cudaSetDevice(d.id));
    double *F = (double *)d.fld.p;
    const int64_t t0 = std::min(np, g * shard), t1 = std::min(np, t0 + shard), nt = t1 - t0;
    SrcView src{F, nf, 0, 3, 6};
    Plan plan;
    TRY(uj_sweep(h, d, st, kernel, F + t0 * nf, nf, nt, src, 0, np, flags, plan));
    if (nt > 0) {
      UjFinishArgs f;
      f.partial = (const double *)d.partial.p; f.pstride = plan.pstride; f.nsplit = plan.nsplit;
      f.nt = nt; f.out = F + t0 * nf; f.ld = nf; f.urow = R_U; f.jrow = R_J; f.zrow0 = R_W; f.zrow1 = R_PSE;
      f.want_U = 1; f.want_J = 1; f.accumulate = 1; f.reset = (flags & VPM_FLAG_RESET) ? 1 : 0;
      f.stat = F + t0 * nf + R_STATIC; f.sld = nf;
      launch_uj_finish(f, st);
      h->launches++;
      CK(h, cudaGetLastError());
    }
  }
  TRY(field_allgather(h));
  if (flags & VPM_FLAG_SFS) {
    for (int g = 0; g < G; ++g) {
      Dev &d = h->devs[g];
      cudaStream_t st = d.stream;
      CK(h, cudaSetDevice(d.id));
      double *F = (double *)d.fld.p;
      const int64_t t0 = std::min(np, g * shard), t1 = std::min(np, t0 + shard), nt = t1 - t0;
      SrcView src{F, nf, 0, 3, 6};
      Plan sp;
      TRY(sfs_sweep(h, d, st, kernel, F + t0 * nf, nf, F + t0 * nf + R_J, nf, nullptr, nt, src, F, nf, R_J,
                    F + R_STATIC, nf, nullptr, np, flags, sp));
      if (nt > 0) {
        SfsFinishArgs q;
        q.partial = (const double *)d.partial.p; q.pstride = sp.pstride; q.nsplit = sp.nsplit;
        q.nt = nt; q.tindex = nullptr; q.out = F + t0 * nf; q.ld = nf; q.row = R_SFS; q.accumulate = 1;
        q.reset = (flags & VPM_FLAG_RESET_SFS) ? 1 : 0; q.filter_static = 1;
        q.stat = F + t0 * nf + R_STATIC; q.sld = nf;
        launch_sfs_finish(q, st);
        h->launches++;
        CK(h, cudaGetLastError());
      }
    }
    TRY(field_allgather(h));
  } else if (flags & VPM_FLAG_RESET_SFS) {
    for (int g = 0; g < G; ++g) {  // O(N): every device does all particles, no exchange needed
      Dev &d = h->devs[g];
      CK(h, cudaSetDevice(d.id));
      double *F = (double *)d.fld.p;
      zero_rows_kernel<<<blocks_for(np, 256), 256, 0, d.stream>>>(F, nf, R_SFS, 3, np, F + R_STATIC, nf);
      h->launches++;
      CK(h, cudaGetLastError());
    }
  }
  return VPM_OK;
}


int field_zeta_lists(vpm_handle *h, int kernel);
template <class L>
int field_on_all(vpm_handle *h, L launch);

// cs.zeta(pfield) on the resident mirror(s).  zeta_direct (the default): J[1:3] of every particle <-
// sum_j Gamma_j zeta_sigma_j; zeta_fmm: field_zeta_lists below.
int field_zeta(vpm_handle *h, int kernel) {
  if (h->zeta_method != 0) return field_zeta_lists(h, kernel);
  const int64_t nf = h->fld_nf, np = h->fld_np;
  if (np == 0) return VPM_OK;
  const int G = (int)h->devs.size();
  const int64_t shard = field_shard(h);
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    cudaStream_t st = d.stream;
    CK(h, cudaSetDevice(d.id));
    double *F = (double *)d.fld.p;
    const int64_t t0 = std::min(np, g * shard), t1 = std::min(np, t0 + shard), nt = t1 - t0;
    SrcView src{F, nf, 0, 3, 6};
    Plan sp;
    TRY(sfs_sweep(h, d, st, kernel, F + t0 * nf, nf, F + t0 * nf + R_J, nf, nullptr, nt, src, F, nf, R_J, nullptr, 1,
                  nullptr, np, VPM_FLAG_TRANSPOSED, sp, false, MODE_ZETA));
    if (nt > 0) {
      SfsFinishArgs q;
      q.partial = (const double *)d.partial.p; q.pstride = sp.pstride; q.nsplit = sp.nsplit;
      q.nt = nt; q.tindex = nullptr; q.out = F + t0 * nf; q.ld = nf; q.row = R_J; q.accumulate = 0; q.reset = 0;
      q.filter_static = 0; q.stat = nullptr; q.sld = 1;
      launch_sfs_finish(q, st);
      h->launches++;
      CK(h, cudaGetLastError());
    }
  }
  return field_allgather(h);
}

// zeta_fmm on the resident mirror(s) (src/FLOWVPM_viscous.jl:523-558): the near field of leaf lists built on
// the device from the resident X and sigma (csrc/vpm_tree.cuh; the reference builds a FastMultipole tree per
// call).  The lists depend on X and sigma only: the CG iterations of the RBF change Gamma alone, so a build
// serves every evaluation until positions or core sizes change (fingerprint of the rows, as vpm_uj_nearfield).
// For a list entry (a, b) the bodies of leaf b RECEIVE from the bodies of leaf a (as vpm_zeta_leafpairs).
// The O(N ncrit) sweep runs on device 0; with G devices the three sums per particle are broadcast over NVLink
// and every device adds them to its own mirror, which keeps the mirrors bit-identical.
int field_zeta_lists(vpm_handle *h, int kernel) {
  const char *fn = "zeta_fmm (resident field)";
  const int64_t nf = h->fld_nf, np = h->fld_np;
  if (np == 0) return VPM_OK;
  const int G = (int)h->devs.size();
  Dev &d = h->devs[0];
  cudaStream_t st = d.stream;
  CK(h, cudaSetDevice(d.id));
  double *F = (double *)d.fld.p;
  unsigned long long fp = 0;
  TRY(tree_fingerprint(h, d, F, nf, 6, np, &fp));
  if (h->tree_np != np || fp != h->tree_fingerprint || h->tree_ncrit != h->zeta_ncrit || h->tree_theta != h->zeta_theta)
    TRY(tree_build(h, F, nf, 6, np, h->zeta_ncrit, h->zeta_theta));
  const TreeView tv = tree_view(h);
  DevCsr c;
  TRY(build_csr_device(h, fn, tv.lbegin, tv.lend, h->tree_nl, np, tv.lbegin, tv.lend, h->tree_nl, np, tv.ps, tv.pt,
                       h->tree_npairs, 1, nullptr, 0, nullptr, 0, c, true));
  const int64_t ns_pad = round_up(np, kTile);
  TRY(scratch_acquire(h, d, st));
  TRY(ensure(h, d.srec, (size_t)ns_pad * kSfsRec * sizeof(double)));
  TRY(ensure(h, d.sfs3, (size_t)np * 3 * sizeof(double)));
  CK(h, cudaMemsetAsync(d.sfs3.p, 0, (size_t)np * 3 * sizeof(double), st));
  if (c.nwi > 0) {
    SrcView sv{F, nf, 0, 3, 6};
    // (the J operand of the record builder is not used in zeta mode)
    prep_sfs_records<<<blocks_for(ns_pad, 256), 256, 0, st>>>(sv, F, nf, R_J, nullptr, 1, tv.sidx, np, ns_pad, kernel, 1,
                                                              (double *)d.srec.p);
    LeafSfsArgs a;
    a.csr = c.csr;
    a.tpos = F; a.tld = nf; a.tJ = F + R_J; a.jld = nf; a.tindex = tv.sidx; a.rec = (const double *)d.srec.p;
    a.out = (double *)d.sfs3.p; a.old = 3; a.orow = 0; a.transposed = 1; a.shortcut = 1;
    launch_sfs_leaf(kernel, c.nt, (unsigned)c.nwi, a, st, MODE_ZETA);
    h->launches += 2;
    CK(h, cudaGetLastError());
  }
  for (int g = 1; g < G; ++g) {
    CK(h, cudaSetDevice(h->devs[g].id));
    TRY(ensure(h, h->devs[g].sfs3, (size_t)np * 3 * sizeof(double)));
  }
  if (G > 1) TRY(bcast_from_dev0(h, &Dev::sfs3, (size_t)np * 3 * sizeof(double)));
  const int accumulate = h->zeta_method == 1;
  return field_on_all(h, [&](Dev &dg) {
    add_rows3_kernel<<<blocks_for(np, 256), 256, 0, dg.stream>>>((const double *)dg.sfs3.p, np, (double *)dg.fld.p, nf, R_J,
                                                                accumulate);
  });
}

StepArgs step_args_of(vpm_handle *h, Dev &d) {
  StepArgs a{};
  a.P = (double *)d.fld.p; a.nf = h->fld_nf; a.np = h->fld_np;
  return a;
}

// launch one O(N) kernel on every device's mirror
template <class L>
int field_on_all(vpm_handle *h, L launch) {
  for (Dev &d : h->devs) {
    CK(h, cudaSetDevice(d.id));
    launch(d);
    h->launches++;
  }
  CK(h, cudaGetLastError());
  return VPM_OK;
}

// sum over the non-static particles of r_k^2 (mode 0) or Gamma_k J_k (mode 1), from device 0
int field_reduce3(vpm_handle *h, int mode, double out[3]) {
  Dev &d = h->devs[0];
  CK(h, cudaSetDevice(d.id));
  TRY(ensure(h, d.jbuf, (size_t)(kRedBlocks * 3 + 3) * sizeof(double)));
  double *partial = (double *)d.jbuf.p, *res = partial + kRedBlocks * 3;
  rbf_reduce_partial<<<kRedBlocks, 256, 0, d.stream>>>(step_args_of(h, d), mode, partial);
  rbf_reduce_final<<<1, kRedBlocks, 0, d.stream>>>(partial, res);
  h->launches += 2;
  CK(h, cudaGetLastError());
  CK(h, cudaMemcpyAsync(out, res, 3 * sizeof(double), cudaMemcpyDeviceToHost, d.stream));
  CK(h, cudaStreamSynchronize(d.stream));
  return VPM_OK;
}

// rbf_conjugategradient (src/FLOWVPM_viscous.jl:309-478) with cs.zeta = zeta_direct
int field_rbf(vpm_handle *h, int kernel, int itmax, double tol, int iterror, int *iterations, double *residuals) {
  const double eps = 2.220446049250313e-16;
  const unsigned nb = blocks_for(h->fld_np, 256);
  auto stage = [&](int st, const double c[3]) {
    return field_on_all(h, [&](Dev &d) { rbf_stage<<<nb, 256, 0, d.stream>>>(step_args_of(h, d), st, c[0], c[1], c[2]); });
  };
  const double zero3[3] = {0, 0, 0};
  double rr0s[3], rrs[3], prev_rrs[3], pAps[3], alphas[3], betas[3];
  bool flags[3];
  TRY(stage(0, zero3));
  TRY(field_zeta(h, kernel));
  TRY(stage(1, zero3));
  TRY(field_reduce3(h, 0, rr0s));
  for (int k = 0; k < 3; ++k) {
    rrs[k] = rr0s[k];
    flags[k] = sqrt(rr0s[k]) > tol || sqrt(rrs[k] / rr0s[k]) > tol;
  }
  int it_done = 0;
  bool failed = false;
  for (int it = 1; it <= itmax; ++it) {
    if (!(flags[0] || flags[1] || flags[2])) break;
    it_done = it;
    TRY(field_zeta(h, kernel));
    TRY(field_reduce3(h, 1, pAps));
    for (int k = 0; k < 3; ++k) {
      alphas[k] = flags[k] ? rrs[k] / pAps[k] : 0.0;  // Julia: x * false == 0 (strong zero)
      prev_rrs[k] = rrs[k];
    }
    TRY(stage(2, alphas));
    TRY(field_reduce3(h, 0, rrs));
    for (int k = 0; k < 3; ++k) {
      betas[k] = rrs[k] / prev_rrs[k];
      if (fabs(prev_rrs[k]) <= 2 * eps) betas[k] = 1;
    }
    TRY(stage(3, betas));
    for (int k = 0; k < 3; ++k)
      flags[k] = flags[k] && (fabs(rr0s[k]) <= 2 * eps ? false : sqrt(rrs[k] / rr0s[k]) > tol);
    if (it == itmax && (flags[0] || flags[1] || flags[2])) failed = true;
  }
  TRY(stage(4, zero3));
  if (iterations) *iterations = it_done;
  if (residuals)
    for (int k = 0; k < 3; ++k) residuals[k] = rr0s[k] > 0 ? sqrt(rrs[k] / rr0s[k]) : 0.0;
  if (failed && iterror)
    return fail(h, VPM_ESTATE, "Maximum number of iterations %d reached before convergence. Errors: %g %g %g, tolerance: %g",
                itmax, sqrt(rrs[0] / rr0s[0]), sqrt(rrs[1] / rr0s[1]), sqrt(rrs[2] / rr0s[2]), tol);
  return VPM_OK;
}

// viscousdiffusion(pfield, CoreSpreading, dt; aux1, aux2): src/FLOWVPM_viscous.jl:152-223
int field_corespreading(vpm_handle *h, const vpm_step_params *sp, double aux1, double aux2) {
  const unsigned nb = blocks_for(h->fld_np, 256);
  const int rk = sp->integration == 1;
  TRY(field_on_all(h, [&](Dev &d) {
    StepArgs a = step_args_of(h, d);
    a.a = aux1; a.b = aux2; a.dt = sp->dt;
    cs_spread<<<nb, 256, 0, d.stream>>>(a, sp->nu, rk);
  }));
  const bool proceed = !rk || fabs(aux2 - 8.0 / 15) <= 1e-7;
  if (!proceed) return VPM_OK;
  h->fld_t_sgm += sp->dt;
  const double beta_cur = sqrt(2 * sp->nu * h->fld_t_sgm / (sp->sgm0 * sp->sgm0) + 1);
  if (beta_cur >= sp->cs_beta) {
    TRY(field_zeta(h, sp->kernel_id));
    TRY(field_on_all(h, [&](Dev &d) { cs_reset<<<nb, 256, 0, d.stream>>>(step_args_of(h, d), sp->sgm0); }));
    TRY(field_rbf(h, sp->kernel_id, sp->cs_itmax, sp->cs_tol, sp->cs_iterror, nullptr, nullptr));
    h->fld_t_sgm = 0.0;
  }
  return VPM_OK;
}

// viscousdiffusion(pfield, ParticleStrengthExchange, dt; aux1, aux2): src/FLOWVPM_viscous.jl:257-298
// (viscous = 2: recalculate_vols = true, the default; 3: false)
int field_pse(vpm_handle *h, const vpm_step_params *sp, double aux1, double aux2) {
  const unsigned nb = blocks_for(h->fld_np, 256);
  const int rk = sp->integration == 1;
  return field_on_all(h, [&](Dev &d) {
    StepArgs a = step_args_of(h, d);
    a.a = aux1; a.b = aux2; a.dt = sp->dt;
    pse_update<<<nb, 256, 0, d.stream>>>(a, sp->nu, rk, sp->viscous == 2);
  });
}
int field_viscous(vpm_handle *h, const vpm_step_params *sp, double aux1, double aux2) {
  return sp->viscous == 1 ? field_corespreading(h, sp, aux1, aux2) : field_pse(h, sp, aux1, aux2);
}

double zeta0_of(int kernel) {  // kernel.zeta(0): src/FLOWVPM_kernel.jl:45,51,60,69-74
  const double pi = 3.14159265358979323846;
  switch (kernel) {
    case K_SING: return 1.0;
    case K_GAUS: return 3.0 / (4.0 * pi);
    case K_GERF: return 1.0 / pow(2.0 * pi, 1.5);
    default: return 1.0 / (4.0 * pi) * 7.5 / sqrt(1.0);
  }
}

}  // namespace
