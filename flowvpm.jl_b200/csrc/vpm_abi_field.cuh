// vpm_abi_field.cuh -- exports: device-resident field and time step.
// Part of the single translation unit vpm_abi.cu (included there in order; not a standalone header).
#pragma once
extern "C" {

int vpm_field_upload(vpm_handle *h, const double *P, int64_t nf, int64_t np) {
  TRY(check_field(h, "vpm_field_upload", P, nf, np, 0));
  const int G = (int)h->devs.size();
  const int64_t shard = (np + G - 1) / G, np_pad = std::max<int64_t>(shard * G, 1);
  for (Dev &d : h->devs) {
    CK(h, cudaSetDevice(d.id));
    TRY(ensure(h, d.fld, (size_t)np_pad * nf * sizeof(double)));
  }
  Dev &d0 = h->devs[0];
  CK(h, cudaSetDevice(d0.id));
  if (np_pad > np)
    CK(h, cudaMemsetAsync((double *)d0.fld.p + np * nf, 0, (size_t)(np_pad - np) * nf * sizeof(double), d0.stream));
  if (np > 0) TRY(h2d_contig(h, d0.stream, d0.fld.p, P, (size_t)np * nf * sizeof(double)));
  h->fld_nf = nf;
  h->fld_np = np;
  h->fld_t_sgm = 0.0;
  TRY(bcast_from_dev0(h, &Dev::fld, (size_t)np_pad * nf * sizeof(double)));
  for (int g = G - 1; g >= 0; --g) {
    CK(h, cudaSetDevice(h->devs[g].id));
    CK(h, cudaStreamSynchronize(h->devs[g].stream));
  }
  return VPM_OK;
}

int vpm_field_download(vpm_handle *h, double *P, int64_t nf, int64_t np) {
  TRY(check_field(h, "vpm_field_download", P, nf, np, 0));
  if (h->fld_np != np || h->fld_nf != nf)
    return fail(h, VPM_ESTATE, "vpm_field_download: a %lld x %lld field is resident, not %lld x %lld",
                (long long)h->fld_nf, (long long)h->fld_np, (long long)nf, (long long)np);
  Dev &d = h->devs[0];
  CK(h, cudaSetDevice(d.id));
  if (np > 0) TRY(d2h_contig(h, d.stream, P, d.fld.p, (size_t)np * nf * sizeof(double)));
  CK(h, cudaStreamSynchronize(d.stream));
  return VPM_OK;
}

static int field_sync_all(vpm_handle *h) {
  for (int g = (int)h->devs.size() - 1; g >= 0; --g) {
    CK(h, cudaSetDevice(h->devs[g].id));
    CK(h, cudaStreamSynchronize(h->devs[g].stream));
  }
  return VPM_OK;
}

int vpm_field_uj(vpm_handle *h, int kernel, int flags) {
  if (!h) return VPM_EINVAL;
  if (h->fld_np < 0) return fail(h, VPM_ESTATE, "vpm_field_uj: no resident field (call vpm_field_upload first)");
  if (!valid_kernel(kernel)) return fail(h, VPM_EINVAL, "vpm_field_uj: unknown kernel_id %d", kernel);
  Dev &d = h->devs[0];
  h->launches = 0;
  CK(h, cudaSetDevice(d.id));
  CK(h, cudaEventRecord(d.ev[0], d.stream));
  CK(h, cudaEventRecord(d.ev[1], d.stream));
  TRY(field_uj(h, kernel, flags));
  CK(h, cudaSetDevice(d.id));
  for (int k = 2; k <= 5; ++k) CK(h, cudaEventRecord(d.ev[k], d.stream));
  TRY(field_sync_all(h));
  h1_fill_timing(h, d);
  h->timing.uj_ms = h->timing.total_ms;
  return VPM_OK;
}

int vpm_field_step(vpm_handle *h, const vpm_step_params *sp) {
  if (!h || !sp) return VPM_EINVAL;
  if (h->fld_np < 0) return fail(h, VPM_ESTATE, "vpm_field_step: no resident field (call vpm_field_upload first)");
  if (!valid_kernel(sp->kernel_id)) return fail(h, VPM_EINVAL, "vpm_field_step: unknown kernel_id %d", sp->kernel_id);
  if (sp->integration < 0 || sp->integration > 1 || sp->relaxation < 0 || sp->relaxation > 2)
    return fail(h, VPM_EINVAL, "vpm_field_step: integration must be 0 (euler) or 1 (rungekutta3), relaxation 0..2");
  if (h->fld_nf < 44) return fail(h, VPM_EINVAL, "vpm_field_step: the resident field needs >= 44 rows");
  if (sp->sfs < 0 || sp->sfs > 2) return fail(h, VPM_EINVAL, "vpm_field_step: sfs must be 0 (none), 1 (constant) or 2 (dynamic)");
  if (sp->sfs == 2 && (sp->minC < 0 || sp->maxC < 0 || sp->minC > sp->maxC || sp->alpha <= 0))
    return fail(h, VPM_EINVAL, "vpm_field_step: invalid DynamicSFS parameters (minC=%g maxC=%g alpha=%g)", sp->minC, sp->maxC, sp->alpha);
  if (sp->viscous < 0 || sp->viscous > 3)
    return fail(h, VPM_EINVAL, "vpm_field_step: viscous must be 0 (Inviscid), 1 (CoreSpreading), 2 or 3 (ParticleStrengthExchange with / without recalculate_vols)");
  if (sp->viscous >= 2 && sp->nu < 0) return fail(h, VPM_EINVAL, "vpm_field_step: ParticleStrengthExchange needs nu >= 0");
  if (sp->viscous == 1 && sp->kernel_id != K_GERF)
    return fail(h, VPM_EINVAL, "vpm_field_step: kernel %d is not compatible with viscous scheme CoreSpreading; compatible kernels are gaussianerf", sp->kernel_id);  // src/FLOWVPM_utils.jl:58-64
  if (sp->viscous == 1 && (sp->sgm0 <= 0 || sp->nu < 0 || sp->cs_itmax < 0))
    return fail(h, VPM_EINVAL, "vpm_field_step: invalid CoreSpreading parameters (nu=%g sgm0=%g itmax=%d)", sp->nu, sp->sgm0, sp->cs_itmax);
  h->launches = 0;
  const int64_t np = h->fld_np;
  if (np == 0) return VPM_OK;
  const int G = (int)h->devs.size();
  const int tr = sp->transposed ? VPM_FLAG_TRANSPOSED : 0;
  const int uj_flags = VPM_FLAG_RESET | tr | (sp->sfs ? (VPM_FLAG_SFS | VPM_FLAG_RESET_SFS) : 0);
  const unsigned nb = blocks_for(np, 256);
  std::vector<StepArgs> args(G);
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    CK(h, cudaSetDevice(d.id));
    StepArgs &a = args[g];
    a.P = (double *)d.fld.p; a.nf = h->fld_nf; a.np = np;
    a.a = 1.0; a.b = 1.0; a.dt = sp->dt; a.Ux = sp->Uinf[0]; a.Uy = sp->Uinf[1]; a.Uz = sp->Uinf[2];
    a.f = sp->f; a.g = sp->g; a.zeta0 = zeta0_of(sp->kernel_id); a.Cs = sp->Cs; a.rlxf = sp->rlxf;
    a.transposed = sp->transposed; a.sfs = sp->sfs; a.clip = sp->clip_backscatter; a.relax_kind = sp->relaxation;
    a.alpha = sp->alpha; a.sfs_rlxf = sp->sfs_rlxf; a.minC = sp->minC; a.maxC = sp->maxC;
    a.force_positive = sp->force_positive;
    a.controls = sp->controls; a.deltat = sp->deltat;
    TRY(ensure(h, d.flg, 256));
    a.nan_flag = (int *)d.flg.p;
    CK(h, cudaMemsetAsync(a.nan_flag, 0, sizeof(int), d.stream));
  }
  // an O(N) kernel on every device's mirror (all mirrors hold the same data)
  auto on_all = [&](auto launch) -> int {
    for (int g = 0; g < G; ++g) {
      CK(h, cudaSetDevice(h->devs[g].id));
      launch(args[g], h->devs[g].stream);
      h->launches++;
    }
    CK(h, cudaGetLastError());
    return VPM_OK;
  };
  // the SFS hooks around a UJ evaluation at an Euler step / the first RK substep
  // (src/FLOWVPM_subfilterscale.jl:110-135 ConstantSFS, :204-268 DynamicSFS)
  auto sfs_before = [&]() -> int {
    if (sp->sfs != 2) return VPM_OK;
    TRY(on_all([&](StepArgs &a, cudaStream_t st) { step_scale_sigma<<<nb, 256, 0, st>>>(a, 0); }));
    TRY(field_uj(h, sp->kernel_id, VPM_FLAG_RESET | VPM_FLAG_RESET_SFS | VPM_FLAG_SFS | tr));
    TRY(on_all([&](StepArgs &a, cudaStream_t st) { step_dyn_store<<<nb, 256, 0, st>>>(a); }));
    TRY(on_all([&](StepArgs &a, cudaStream_t st) { step_scale_sigma<<<nb, 256, 0, st>>>(a, 1); }));
    return VPM_OK;
  };
  auto sfs_after = [&]() -> int {
    if (sp->sfs == 1) TRY(on_all([&](StepArgs &a, cudaStream_t st) { step_sfs_coeff<<<nb, 256, 0, st>>>(a); }));
    if (sp->sfs == 2) TRY(on_all([&](StepArgs &a, cudaStream_t st) { step_dyn_coeff<<<nb, 256, 0, st>>>(a); }));
    if (sp->sfs && (sp->controls & 3))
      TRY(on_all([&](StepArgs &a, cudaStream_t st) { step_sfs_controls<<<nb, 256, 0, st>>>(a); }));
    return VPM_OK;
  };
  Dev &d0 = h->devs[0];
  CK(h, cudaSetDevice(d0.id));
  CK(h, cudaEventRecord(d0.ev[0], d0.stream));
  CK(h, cudaEventRecord(d0.ev[1], d0.stream));
  if (sp->integration == 0) {  // euler: src/FLOWVPM_timeintegration.jl:23-37
    TRY(sfs_before());
    TRY(field_uj(h, sp->kernel_id, uj_flags));
    TRY(sfs_after());
    const int relax = sp->relax ? 1 : 0;
    TRY(on_all([&](StepArgs &a, cudaStream_t st) { step_euler<<<nb, 256, 0, st>>>(a, relax); }));
    if (sp->viscous) TRY(field_viscous(h, sp, 0.0, 0.0));
  } else {  // rungekutta3: src/FLOWVPM_timeintegration.jl:388-461
    TRY(on_all([&](StepArgs &a, cudaStream_t st) { step_reset_M<<<nb, 256, 0, st>>>(a); }));
    const double ab[3][2] = {{0.0, 1.0 / 3}, {-5.0 / 9, 15.0 / 16}, {-153.0 / 128, 8.0 / 15}};
    for (int k = 0; k < 3; ++k) {
      for (StepArgs &a : args) { a.a = ab[k][0]; a.b = ab[k][1]; }
      if (k == 0) TRY(sfs_before());
      TRY(field_uj(h, sp->kernel_id, uj_flags));
      if (k == 0) TRY(sfs_after());
      TRY(on_all([&](StepArgs &a, cudaStream_t st) { step_rk_stage<<<nb, 256, 0, st>>>(a); }));
      if (sp->viscous) TRY(field_viscous(h, sp, ab[k][0], ab[k][1]));
    }
    if (sp->relax && sp->relaxation) {
      TRY(field_uj(h, sp->kernel_id, VPM_FLAG_RESET | tr));
      TRY(on_all([&](StepArgs &a, cudaStream_t st) { step_relax<<<nb, 256, 0, st>>>(a); }));
    }
  }
  CK(h, cudaSetDevice(d0.id));
  for (int k = 2; k <= 5; ++k) CK(h, cudaEventRecord(d0.ev[k], d0.stream));
  int nan_flag = 0;
  CK(h, cudaMemcpyAsync(&nan_flag, args[0].nan_flag, sizeof(int), cudaMemcpyDeviceToHost, d0.stream));
  TRY(field_sync_all(h));
  h1_fill_timing(h, d0);
  h->timing.uj_ms = h->timing.total_ms;
  if (nan_flag) return fail(h, VPM_ESTATE, "NaN in dynamicprocedure_pseudo3level_afterUJ");  // subfilterscale.jl:645-652
  return VPM_OK;
}

int vpm_field_rbf(vpm_handle *h, int kernel, int itmax, double tol, int iterror, int *iterations, double *residuals) {
  if (!h) return VPM_EINVAL;
  if (h->fld_np < 0) return fail(h, VPM_ESTATE, "vpm_field_rbf: no resident field (call vpm_field_upload first)");
  if (!valid_kernel(kernel)) return fail(h, VPM_EINVAL, "vpm_field_rbf: unknown kernel_id %d", kernel);
  if (itmax < 0 || !(tol >= 0)) return fail(h, VPM_EINVAL, "vpm_field_rbf: itmax >= 0 and tol >= 0 required");
  h->launches = 0;
  if (iterations) *iterations = 0;
  if (h->fld_np == 0) return VPM_OK;
  int rc = field_rbf(h, kernel, itmax, tol, iterror, iterations, residuals);
  int rs = field_sync_all(h);
  h->timing.kernel_launches = h->launches;
  return rc != VPM_OK ? rc : rs;
}

int vpm_field_zeta_method(vpm_handle *h, int method, int64_t ncrit, double theta) {
  if (!h) return VPM_EINVAL;
  if (method < VPM_ZETA_DIRECT || method > VPM_ZETA_FMM_RESET)
    return fail(h, VPM_EINVAL, "vpm_field_zeta_method: method must be VPM_ZETA_DIRECT, VPM_ZETA_FMM or VPM_ZETA_FMM_RESET (got %d)", method);
  if (method != VPM_ZETA_DIRECT && (ncrit < 1 || !(theta > 0.0)))
    return fail(h, VPM_EINVAL, "vpm_field_zeta_method: ncrit >= 1 and theta > 0 required");
  h->zeta_method = method;
  if (method != VPM_ZETA_DIRECT) { h->zeta_ncrit = ncrit; h->zeta_theta = theta; }
  return VPM_OK;
}

int vpm_field_zeta(vpm_handle *h, int kernel) {
  if (!h) return VPM_EINVAL;
  if (h->fld_np < 0) return fail(h, VPM_ESTATE, "vpm_field_zeta: no resident field (call vpm_field_upload first)");
  if (!valid_kernel(kernel)) return fail(h, VPM_EINVAL, "vpm_field_zeta: unknown kernel_id %d", kernel);
  h->launches = 0;
  if (h->fld_np == 0) return VPM_OK;
  int rc = field_zeta(h, kernel);
  int rs = field_sync_all(h);
  h->timing.kernel_launches = h->launches;
  return rc != VPM_OK ? rc : rs;
}

int vpm_field_tsgm(vpm_handle *h, double *t_sgm, int set) {
  if (!h || !t_sgm) return VPM_EINVAL;
  if (set) h->fld_t_sgm = *t_sgm; else *t_sgm = h->fld_t_sgm;
  return VPM_OK;
}

}  // extern "C"
