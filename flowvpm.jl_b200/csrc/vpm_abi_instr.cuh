// vpm_abi_instr.cuh -- exports: timing record, pipe-peak probes, device-math test hook.
// Part of the single translation unit vpm_abi.cu (included there in order; not a standalone header).
#pragma once
extern "C" {

int vpm_get_timing(const vpm_handle *h, vpm_timing *out) {
  if (!h || !out) return VPM_EINVAL;
  *out = h->timing;
  if (h->device_timing) {
    // stream-ordered entry points: the pair kernel's own duration, valid once the
    // caller has synchronised the stream it passed
    const Dev &d = h->devs[0];
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, d.ev[6], d.ev[7]) != cudaSuccess) { cudaGetLastError(); ms = 0.f; }
    if (h->device_timing == 1) out->uj_ms = ms; else out->sfs_ms = ms;
    out->n_gpus = (int32_t)h->devs.size();
  }
  return VPM_OK;
}

int vpm_measure_dfma_peak(vpm_handle *h, double *dfma_per_s, double *elapsed_ms) {
  if (!h || !dfma_per_s) return VPM_EINVAL;
  Dev &d = h->devs[0];
  cudaStream_t st = d.stream;
  CK(h, cudaSetDevice(d.id));
  TRY(ensure(h, d.ibuf, 4096));
  const int threads = 256, blocks = d.sm_count * 8, iters = 4096;
  dfma_peak_kernel<<<blocks, threads, 0, st>>>((double *)d.ibuf.p, 64, 1.0);  // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(h, cudaEventRecord(d.ev[6], st));
    dfma_peak_kernel<<<blocks, threads, 0, st>>>((double *)d.ibuf.p, iters, 1.0);
    CK(h, cudaEventRecord(d.ev[7], st));
    CK(h, cudaStreamSynchronize(st));
    CK(h, cudaGetLastError());
    best = std::min(best, ev_ms(d.ev[6], d.ev[7]));
  }
  const double n = (double)blocks * threads * (double)iters * 16.0 * 8.0;
  *dfma_per_s = n / (best * 1e-3);
  if (elapsed_ms) *elapsed_ms = best;
  return VPM_OK;
}

int vpm_measure_ffma_peak(vpm_handle *h, int mode, double *fma_per_s, double *elapsed_ms) {
  if (!h || !fma_per_s || mode < 0 || mode > 3) return VPM_EINVAL;
  Dev &d = h->devs[0];
  cudaStream_t st = d.stream;
  CK(h, cudaSetDevice(d.id));
  TRY(ensure(h, d.ibuf, 4096));
  const int threads = 256, blocks = d.sm_count * 8, iters = 4096;
  auto run = [&](int it) {
    float *o = (float *)d.ibuf.p;
    switch (mode) {
      case 0: ffma_peak_kernel<0><<<blocks, threads, 0, st>>>(o, it, 1.0f); break;
      case 1: ffma_peak_kernel<1><<<blocks, threads, 0, st>>>(o, it, 1.0f); break;
      case 2: ffma_peak_kernel<2><<<blocks, threads, 0, st>>>(o, it, 1.0f); break;
      default: ffma_peak_kernel<3><<<blocks, threads, 0, st>>>(o, it, 1.0f); break;
    }
  };
  run(64);  // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(h, cudaEventRecord(d.ev[6], st));
    run(iters);
    CK(h, cudaEventRecord(d.ev[7], st));
    CK(h, cudaStreamSynchronize(st));
    CK(h, cudaGetLastError());
    best = std::min(best, ev_ms(d.ev[6], d.ev[7]));
  }
  // scalar FMAs per second: 8 chains x 16 x 2 lanes per thread and iteration in every mode
  const double n = (double)blocks * threads * (double)iters * 16.0 * 8.0 * 2.0;
  *fma_per_s = n / (best * 1e-3);
  if (elapsed_ms) *elapsed_ms = best;
  return VPM_OK;
}

int vpm_test_math(vpm_handle *h, int op, int arg, const double *in, double *out, double *out2, int64_t n) {
  if (!h || !in || !out || n < 0) return fail(h, VPM_EINVAL, "vpm_test_math: bad argument");
  if (n == 0) return VPM_OK;
  Dev &d = h->devs[0];
  cudaStream_t st = d.stream;
  CK(h, cudaSetDevice(d.id));
  TRY(ensure(h, d.tbuf, (size_t)n * sizeof(double)));
  TRY(ensure(h, d.sbuf, (size_t)n * 2 * sizeof(double)));
  CK(h, cudaMemcpyAsync(d.tbuf.p, in, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
  double *o1 = (double *)d.sbuf.p, *o2 = o1 + n;
  if (op == 4) {  // (A, B) of the bank-replicated table path (vpm_kernels_tab.cuh), arg = K_GAUS or K_GERF
    if (arg != K_GERF && arg != K_GAUS) return fail(h, VPM_EINVAL, "vpm_test_math: op 4 needs arg 1 or 2");
    const size_t smem = arg == K_GERF ? tab_smem_bytes<K_GERF>() : tab_smem_bytes<K_GAUS>();
    if (arg == K_GERF) {
      CK(h, cudaFuncSetAttribute(test_tab_kernel<K_GERF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      test_tab_kernel<K_GERF><<<blocks_for(n, 256), 256, smem, st>>>((const double *)d.tbuf.p, o1, o2, n);
    } else {
      CK(h, cudaFuncSetAttribute(test_tab_kernel<K_GAUS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      test_tab_kernel<K_GAUS><<<blocks_for(n, 256), 256, smem, st>>>((const double *)d.tbuf.p, o1, o2, n);
    }
  } else {
    test_math_kernel<<<blocks_for(n, 256), 256, 0, st>>>(op, arg, (const double *)d.tbuf.p, o1, o2, n);
  }
  CK(h, cudaGetLastError());
  CK(h, cudaMemcpyAsync(out, o1, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (out2) CK(h, cudaMemcpyAsync(out2, o2, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(h, cudaStreamSynchronize(st));
  return VPM_OK;
}

int vpm_plan_query(int64_t nt, int64_t ns, int sm_count, int kind, int64_t *out) {
  if (!out || nt < 0 || ns < 0 || sm_count < 1 || kind < 0 || kind > 3) return VPM_EINVAL;
  bool fills = false;
  const Plan p = kind == 3 ? make_plan_tab(nt, ns, sm_count, &fills)
                           : make_plan(nt, ns, sm_count, kind == 0 ? PLAN_UJ : kind == 1 ? PLAN_SFS : PLAN_UJ_F32);
  out[0] = (int64_t)(p.tab ? p.tab : kThreads) * p.T;
  out[1] = p.grid.x; out[2] = p.nsplit; out[3] = p.src_per_split; out[4] = p.tiles_per_split;
  out[5] = p.T; out[6] = p.unroll; out[7] = fills ? 1 : 0;
  return VPM_OK;
}

}  // extern "C"
