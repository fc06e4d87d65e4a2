// vpm_abi_lists.cuh -- exports: Hook 3 and the other leaf-list entry points, zeta_direct, device-built lists.
// Part of the single translation unit vpm_abi.cu (included there in order; not a standalone header).
#pragma once
extern "C" {

int vpm_p2p_leafpairs(vpm_handle *h, double *tgt, int64_t ld, int64_t n_tgt, int row_pos, int row_grad,
                      int row_hess, const double *src, int64_t n_src, const int64_t *tb,
                      const int64_t *te, int64_t ntl, const int64_t *sb, const int64_t *se, int64_t nsl,
                      const int32_t *pt, const int32_t *ps, int64_t npairs, int kernel, int want_U,
                      int want_J) {
  if (!h) return VPM_EINVAL;
  const char *fn = "vpm_p2p_leafpairs";
  if (!valid_kernel(kernel)) return fail(h, VPM_EINVAL, "%s: unknown kernel_id %d", fn, kernel);
  if (n_tgt < 0 || n_src < 0 || ntl < 0 || nsl < 0 || npairs < 0 || ld < 3)
    return fail(h, VPM_EINVAL, "%s: negative size or ld < 3", fn);
  if (row_pos < 0 || row_pos + 3 > ld || (want_U && (row_grad < 0 || row_grad + 3 > ld)) ||
      (want_J && (row_hess < 0 || row_hess + 9 > ld)))
    return fail(h, VPM_EINVAL, "%s: row offsets outside the %lld-row target buffer", fn, (long long)ld);
  if (npairs == 0 || n_tgt == 0 || n_src == 0 || (!want_U && !want_J)) return VPM_OK;
  if (!tgt || !src || !tb || !te || !sb || !se || !pt || !ps) return fail(h, VPM_EINVAL, "%s: NULL argument", fn);
  h->launches = 0;
  // Multi-GPU: target leaves are sharded into G contiguous runs of work items
  // with balanced  sum nt*ns ; sources are replicated.  Needs the leaves in increasing,
  // non-overlapping body order (tree-sorted buffers) so that a device's targets are one
  // contiguous column range; otherwise device 0 does everything.
  int G = (int)h->devs.size();
  const int64_t ns_pad = round_up(n_src, kTile);
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    CK(h, cudaSetDevice(d.id));
    TRY(ensure(h, d.tbuf, (size_t)n_tgt * ld * sizeof(double)));
    TRY(ensure(h, d.sbuf, (size_t)((n_src + G - 1) / G * G) * 8 * sizeof(double)));
    TRY(ensure(h, d.rec, (size_t)ns_pad * kRec * sizeof(double)));
  }
  // replicated inputs (source buffer, CSR tables): device 0 gets them from the host, the
  // other devices over NVLink.  The list is regrouped on device 0 (vpm_csr.cuh).
  Dev &d0 = h->devs[0];
  CK(h, cudaSetDevice(d0.id));
  CK(h, cudaEventRecord(d0.ev[0], d0.stream));
  // page-locked source buffer on a multi-GPU handle: every device pulls 1/G of it over its own PCIe link and
  // an all-gather over NVLink completes the copies (h2d_rows_sharded); otherwise device 0 + broadcast
  const bool src_sharded = G > 1 && host_is_pinned(src);
  if (src_sharded) {
    TRY(h2d_rows_sharded(h, &Dev::sbuf, src, 8, 8, n_src));
    CK(h, cudaSetDevice(d0.id));
  } else {
    TRY(h2d_contig(h, d0.stream, d0.sbuf.p, src, (size_t)n_src * 8 * sizeof(double)));
  }
  DevCsr c;
  TRY(build_csr_device(h, fn, tb, te, ntl, n_tgt, sb, se, nsl, n_src, pt, ps, npairs, G, nullptr, 0, nullptr, 0, c));
  if (c.nwi == 0) return VPM_OK;
  G = c.G_eff;  // 1 when the target leaves that carry work overlap
  std::vector<LeafCsr> csr(G);
  csr[0] = c.csr;
  for (int g = 1; g < G; ++g) {
    Dev &d = h->devs[g];
    CK(h, cudaSetDevice(d.id));
    TRY(ensure(h, d.ibuf, h->devs[0].ibuf.cap));
    csr[g] = rebase_csr(c.csr, h->devs[0].ibuf.p, d.ibuf.p);
  }
  if (G > 1 && !src_sharded) TRY(bcast_from_dev0(h, &Dev::sbuf, (size_t)n_src * 8 * sizeof(double)));
  if (G > 1) TRY(bcast_from_dev0(h, &Dev::ibuf, c.bcast_bytes));
  std::vector<std::pair<int64_t, int64_t>> cols(G, {0, 0});
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    cudaStream_t st = d.stream;
    const int64_t k0 = c.cut[g], k1 = c.cut[g + 1];
    if (k1 <= k0) continue;
    CK(h, cudaSetDevice(d.id));
    // this device's target columns: first target of its first item .. last target of its last
    const int64_t lf = c.first_leaf[g], ll = c.last_leaf[g + 1];
    const int64_t col0 = G == 1 ? 0 : tb[lf] + c.first_off[g];
    const int64_t col1 = G == 1 ? n_tgt : std::min<int64_t>(te[ll], tb[ll] + c.last_off[g + 1] + c.nt);
    TRY(h2d_contig(h, st, (double *)d.tbuf.p + col0 * ld, tgt + col0 * ld, (size_t)(col1 - col0) * ld * sizeof(double)));
    LeafUjArgs a;
    a.csr = csr[g];
    a.csr.wi_leaf += k0;
    a.csr.wi_off += k0;
    if (g == 0) CK(h, cudaEventRecord(d.ev[1], st));
    a.tpos = (const double *)d.tbuf.p + row_pos; a.tld = ld; a.rec = (const double *)d.rec.p;
    a.out = (double *)d.tbuf.p; a.urow = row_grad; a.jrow = row_hess; a.want_U = want_U; a.want_J = want_J;
    a.shortcut = 1;
    launch_uj_leaf_any(h, d, st, kernel, c.nt, (unsigned)(k1 - k0), a, (const double *)d.sbuf.p, n_src, ns_pad);
    CK(h, cudaGetLastError());
    if (g == 0) {
      CK(h, cudaEventRecord(d.ev[2], st));
      CK(h, cudaEventRecord(d.ev[3], st));
      CK(h, cudaEventRecord(d.ev[4], st));
    }
    cols[g] = {col0, col1};
  }
  // downloads in a second pass: a D2H into pageable memory blocks the host, and every
  // device must have its kernel in flight before that happens
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    if (cols[g].second <= cols[g].first) continue;
    CK(h, cudaSetDevice(d.id));
    const int64_t col0 = cols[g].first, col1 = cols[g].second;
    TRY(d2h_contig(h, d.stream, tgt + col0 * ld, (double *)d.tbuf.p + col0 * ld, (size_t)(col1 - col0) * ld * sizeof(double)));
    if (g == 0) CK(h, cudaEventRecord(d.ev[5], d.stream));
  }
  for (int g = G - 1; g >= 0; --g) {
    CK(h, cudaSetDevice(h->devs[g].id));
    CK(h, cudaStreamSynchronize(h->devs[g].stream));
  }
  h->timing.uj_pairs = c.pairs;
  h->timing.sfs_pairs = 0;
  h1_fill_timing(h, h->devs[0]);
  h->np_resident = -1;
  return VPM_OK;
}

// Hook 3 in the call shape the reference shows (src/FLOWVPM_gpu.jl:637-643, helpers :554-602):
//   fmm.nearfield_device!(target_system, target_indices::Vector{UnitRange}, switch, source_system, source_indices)
// Both systems are ParticleFields (46 x N matrices, columns in the order the ranges index); the k-th target
// range receives the sources of the source ranges src_offsets[k] .. src_offsets[k+1]-1 (what
// combine_source_indices / expand_source_indices produced per target leaf).  U (rows 10:12) and J (rows 16:24)
// of the targets are ACCUMULATED on (no reset: UJ_fmm resets before calling fmm!, src/FLOWVPM_UJ.jl:75-80).
// Sources are replicated over the devices, target ranges are cut into contiguous runs of equal work and every
// device moves its own target columns over its own PCIe link.
int vpm_nearfield_ranges(vpm_handle *h, double *TP, int64_t nf_t, int64_t np_t, const int64_t *tb, const int64_t *te,
                         int64_t ntr, const double *SP, int64_t nf_s, int64_t np_s, const int64_t *sb,
                         const int64_t *se, const int64_t *soff, int kernel, int want_U, int want_J) {
  const char *fn = "vpm_nearfield_ranges";
  if (!h) return VPM_EINVAL;
  if (!valid_kernel(kernel)) return fail(h, VPM_EINVAL, "%s: unknown kernel_id %d", fn, kernel);
  if (np_t < 0 || np_s < 0 || ntr < 0 || nf_t < MIN_FIELDS - 19 || nf_s < 7)
    return fail(h, VPM_EINVAL, "%s: negative size, target matrix with < 24 rows or source matrix with < 7 rows", fn);
  if (ntr == 0 || np_t == 0 || np_s == 0 || (!want_U && !want_J)) return VPM_OK;
  if (!TP || !SP || !tb || !te || !sb || !se || !soff) return fail(h, VPM_EINVAL, "%s: NULL argument", fn);
  if (soff[0] != 0) return fail(h, VPM_EINVAL, "%s: src_offsets[0] must be 0", fn);
  for (int64_t k = 0; k < ntr; ++k)
    if (soff[k + 1] < soff[k]) return fail(h, VPM_EINVAL, "%s: src_offsets not non-decreasing at %lld", fn, (long long)k);
  const int64_t nsr = soff[ntr];
  if (nsr == 0) return VPM_OK;
  if (nsr > INT32_MAX || ntr > INT32_MAX) return fail(h, VPM_EINVAL, "%s: more than 2^31 ranges", fn);
  // One owner per target column: the k-th range's CTAs add into its columns without atomics, so target
  // ranges that carry work must not overlap.  IDENTICAL ranges (the reference's own example of the call,
  // warmup_gpu at src/FLOWVPM_gpu.jl:637-643, lists `1:n` once per GPU) are merged: the later entries'
  // source ranges join the first one's list, in order.  Partial overlaps are refused.
  std::vector<int64_t> owner((size_t)ntr);
  {
    std::vector<int64_t> ord;
    for (int64_t k = 0; k < ntr; ++k) {
      owner[(size_t)k] = k;
      if (soff[k + 1] > soff[k] && te[k] > tb[k]) ord.push_back(k);
    }
    // (tree-ordered ranges arrive sorted: 270 912 of them at 2^24 / ncrit 128 -- sort only when they are not)
    auto before = [&](int64_t a, int64_t b) { return tb[a] != tb[b] ? tb[a] < tb[b] : te[a] < te[b]; };
    bool sorted = true;
    for (size_t i = 1; i < ord.size() && sorted; ++i) sorted = !before(ord[i], ord[i - 1]);
    if (!sorted) std::stable_sort(ord.begin(), ord.end(), before);
    for (size_t i = 1; i < ord.size(); ++i) {
      const int64_t a = ord[i - 1], b = ord[i];
      if (tb[a] == tb[b] && te[a] == te[b]) owner[(size_t)b] = owner[(size_t)a];
      else if (tb[b] < te[a])
        return fail(h, VPM_EINVAL, "%s: target ranges %lld [%lld,%lld) and %lld [%lld,%lld) overlap", fn, (long long)a,
                    (long long)tb[a], (long long)te[a], (long long)b, (long long)tb[b], (long long)te[b]);
    }
  }
  // the (target range, source range) list: entry i <-> (owner of k, i) for soff[k] <= i < soff[k+1]; generated on
  // the device from soff and owner (build_csr_device, csr_gen_pairs_kernel)
  h->launches = 0;
  int G = (int)h->devs.size();
  const int TLD = 18;  // device target buffer: rows 0:3 X, rows 3:18 = particle rows 10:24 (U, vorticity, J)
  const int64_t ns_pad = round_up(np_s, kTile);
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    CK(h, cudaSetDevice(d.id));
    TRY(ensure(h, d.tbuf, (size_t)np_t * TLD * sizeof(double)));
    TRY(ensure(h, d.in7, (size_t)((np_s + G - 1) / G * G) * 7 * sizeof(double)));
    TRY(ensure(h, d.rec, (size_t)ns_pad * kRec * sizeof(double)));
  }
  Dev &d0 = h->devs[0];
  CK(h, cudaSetDevice(d0.id));
  CK(h, cudaEventRecord(d0.ev[0], d0.stream));
  const bool src_sharded = G > 1 && host_is_pinned(SP);
  if (src_sharded) TRY(h2d_rows_sharded(h, &Dev::in7, SP, nf_s, 7, np_s));
  else TRY(h2d_rows(h, d0.stream, (double *)d0.in7.p, SP, nf_s, 7, np_s));
  CK(h, cudaSetDevice(d0.id));
  DevCsr c;
  TRY(build_csr_device(h, fn, tb, te, ntr, np_t, sb, se, nsr, np_s, nullptr, nullptr, nsr, G, nullptr, 0, nullptr, 0, c,
                       false, soff, owner.data()));
  if (c.nwi == 0) return VPM_OK;
  G = c.G_eff;
  std::vector<LeafCsr> csr(G);
  csr[0] = c.csr;
  for (int g = 1; g < G; ++g) {
    Dev &d = h->devs[g];
    CK(h, cudaSetDevice(d.id));
    TRY(ensure(h, d.ibuf, h->devs[0].ibuf.cap));
    csr[g] = rebase_csr(c.csr, h->devs[0].ibuf.p, d.ibuf.p);
  }
  if (G > 1 && !src_sharded) TRY(bcast_from_dev0(h, &Dev::in7, (size_t)np_s * 7 * sizeof(double)));
  if (G > 1) TRY(bcast_from_dev0(h, &Dev::ibuf, c.bcast_bytes));
  std::vector<std::pair<int64_t, int64_t>> cols(G, {0, 0});
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    cudaStream_t st = d.stream;
    const int64_t k0 = c.cut[g], k1 = c.cut[g + 1];
    if (k1 <= k0) continue;
    CK(h, cudaSetDevice(d.id));
    const int64_t lf = c.first_leaf[g], ll = c.last_leaf[g + 1];
    const int64_t col0 = G == 1 ? 0 : tb[lf] + c.first_off[g];
    const int64_t col1 = G == 1 ? np_t : std::min<int64_t>(te[ll], tb[ll] + c.last_off[g + 1] + c.nt);
    const size_t ncol = (size_t)(col1 - col0);
    double *dt = (double *)d.tbuf.p + col0 * TLD;
    const RowPiece pieces[2] = {{R_X * sizeof(double), 0, 3 * sizeof(double)},
                                {R_U * sizeof(double), 3 * sizeof(double), 15 * sizeof(double)}};
    TRY(h2d_pieces(h, st, dt, TLD * sizeof(double), TP + col0 * nf_t, nf_t * sizeof(double), pieces, 2, ncol));
    LeafUjArgs a;
    a.csr = csr[g];
    a.csr.wi_leaf += k0;
    a.csr.wi_off += k0;
    if (g == 0) CK(h, cudaEventRecord(d.ev[1], st));
    a.tpos = (const double *)d.tbuf.p; a.tld = TLD; a.rec = (const double *)d.rec.p;
    a.out = (double *)d.tbuf.p; a.urow = 3; a.jrow = 3 + (R_J - R_U); a.want_U = want_U; a.want_J = want_J;
    a.shortcut = 1;
    if (g == 0) CK(h, cudaEventRecord(d.ev[6], st));
    launch_uj_leaf_any(h, d, st, kernel, c.nt, (unsigned)(k1 - k0), a, SrcView{(const double *)d.in7.p, 7, 0, 3, 6}, np_s, ns_pad);
    if (g == 0) CK(h, cudaEventRecord(d.ev[7], st));
    CK(h, cudaGetLastError());
    if (g == 0) {
      CK(h, cudaEventRecord(d.ev[2], st));
      CK(h, cudaEventRecord(d.ev[3], st));
      CK(h, cudaEventRecord(d.ev[4], st));
    }
    cols[g] = {col0, col1};
  }
  // downloads in a second pass (a D2H into pageable memory blocks the host: every device must have its
  // kernel in flight before that happens)
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    if (cols[g].second <= cols[g].first) continue;
    CK(h, cudaSetDevice(d.id));
    const int64_t col0 = cols[g].first;
    const size_t ncol = (size_t)(cols[g].second - col0);
    TRY(d2h_strided(h, d.stream, TP + col0 * nf_t + R_U, nf_t * sizeof(double), (double *)d.tbuf.p + col0 * TLD + 3,
                    TLD * sizeof(double), 15 * sizeof(double), ncol));
    if (g == 0) CK(h, cudaEventRecord(d.ev[5], d.stream));
  }
  for (int g = G - 1; g >= 0; --g) {
    CK(h, cudaSetDevice(h->devs[g].id));
    CK(h, cudaStreamSynchronize(h->devs[g].stream));
  }
  h->timing.uj_pairs = c.pairs;
  h->timing.sfs_pairs = 0;
  h1_fill_timing(h, h->devs[0]);
  h->timing.uj_ms = ev_ms(d0.ev[6], d0.ev[7]);
  h->np_resident = -1;
  return VPM_OK;
}

static int leafpairs_field(vpm_handle *h, const char *fn, int mode, double *P, int64_t nf, int64_t np,
                           const int64_t *tsort, const int64_t *ssort, const int64_t *tb, const int64_t *te,
                           int64_t ntl, const int64_t *sb, const int64_t *se, int64_t nsl, const int32_t *pt,
                           const int32_t *ps, int64_t npairs, int kernel, int flags) {
  TRY(check_field(h, fn, P, nf, np, kernel));
  if (ntl < 0 || nsl < 0 || npairs < 0) return fail(h, VPM_EINVAL, "%s: negative size", fn);
  if (npairs == 0 || np == 0) return VPM_OK;
  if (!tsort || !ssort || !tb || !te || !sb || !se || !pt || !ps) return fail(h, VPM_EINVAL, "%s: NULL argument", fn);
  for (int64_t i = 0; i < np; ++i)
    if (tsort[i] < 0 || tsort[i] >= np || ssort[i] < 0 || ssort[i] >= np)
      return fail(h, VPM_EINVAL, "%s: sort index %lld out of range", fn, (long long)i);
  Dev &d = h->devs[0];
  cudaStream_t st = d.stream;
  h->launches = 0;
  // Multi-GPU: as Hook 3 -- target leaves sharded over the devices in contiguous runs of work
  // items; needs increasing, non-overlapping target leaves (tree-sorted), else device 0 alone
  int G = (int)h->devs.size();
  const int64_t ns_pad = round_up(np, kTile);
  for (int g = 0; g < G; ++g) {
    Dev &dg = h->devs[g];
    CK(h, cudaSetDevice(dg.id));
    TRY(ensure(h, dg.in7, (size_t)np * 7 * sizeof(double)));
    TRY(ensure(h, dg.jbuf, (size_t)np * 9 * sizeof(double)));
    TRY(scratch_acquire(h, dg, dg.stream));
    TRY(ensure(h, dg.srec, (size_t)ns_pad * kSfsRec * sizeof(double)));
    if (G > 1) TRY(ensure(h, dg.tbuf, (size_t)np * 3 * sizeof(double)));
  }
  CK(h, cudaSetDevice(d.id));
  TRY(ensure(h, d.sfs3, (size_t)np * 3 * sizeof(double)));
  CK(h, cudaEventRecord(d.ev[0], st));
  TRY(h2d_rows(h, st, (double *)d.in7.p, P, nf, 7, np));
  TRY(h2d_rows(h, st, (double *)d.jbuf.p, P + R_J, nf, 9, np));
  const int out_row = mode == MODE_ZETA ? R_J : R_SFS;
  TRY(h2d_rows(h, st, (double *)d.sfs3.p, P + out_row, nf, 3, np));
  DevCsr c;
  TRY(build_csr_device(h, fn, tb, te, ntl, np, sb, se, nsl, np, pt, ps, npairs, G, tsort, np, ssort, np, c));
  if (c.nwi == 0) return VPM_OK;
  G = c.G_eff;
  CK(h, cudaEventRecord(d.ev[1], st));
  CK(h, cudaEventRecord(d.ev[2], st));
  CK(h, cudaEventRecord(d.ev[3], st));
  for (int g = 1; g < G; ++g) {
    CK(h, cudaSetDevice(h->devs[g].id));
    TRY(ensure(h, h->devs[g].ibuf, d.ibuf.cap));
  }
  if (G > 1) TRY(bcast_from_dev0(h, &Dev::in7, (size_t)np * 7 * sizeof(double)));
  if (G > 1) TRY(bcast_from_dev0(h, &Dev::jbuf, (size_t)np * 9 * sizeof(double)));
  if (G > 1) TRY(bcast_from_dev0(h, &Dev::ibuf, c.bcast_bytes));
  const int transposed = (flags & VPM_FLAG_TRANSPOSED) ? 1 : 0;
  std::vector<std::pair<int64_t, int64_t>> cols(G, {0, 0});
  for (int g = 0; g < G; ++g) {
    Dev &dg = h->devs[g];
    cudaStream_t sg = dg.stream;
    const int64_t k0 = c.cut[g], k1 = c.cut[g + 1];
    if (k1 <= k0) continue;
    CK(h, cudaSetDevice(dg.id));
    const ptrdiff_t shift = (const char *)dg.ibuf.p - (const char *)d.ibuf.p;
    const int64_t *dts = (const int64_t *)((const char *)c.d_tsort + shift);
    const int64_t *dss = (const int64_t *)((const char *)c.d_ssort + shift);
    SrcView sv{(const double *)dg.in7.p, 7, 0, 3, 6};
    prep_sfs_records<<<blocks_for(ns_pad, 256), 256, 0, sg>>>(sv, (const double *)dg.jbuf.p, 9, 0, nullptr, 1, dss, np,
                                                              ns_pad, kernel, transposed, (double *)dg.srec.p);
    LeafSfsArgs a;
    a.csr = g == 0 ? c.csr : rebase_csr(c.csr, d.ibuf.p, dg.ibuf.p);
    a.csr.wi_leaf += k0;
    a.csr.wi_off += k0;
    a.tpos = (const double *)dg.in7.p; a.tld = 7; a.tJ = (const double *)dg.jbuf.p; a.jld = 9;
    a.tindex = dts; a.rec = (const double *)dg.srec.p; a.old = 3; a.orow = 0;
    a.transposed = transposed;
    a.shortcut = (flags & VPM_FLAG_NO_FARFIELD_SHORTCUT) ? 0 : 1;
    if (G == 1) {
      a.out = (double *)dg.sfs3.p;  // particle-indexed, accumulated in place
    } else {
      // sums land in a zeroed buffer indexed by sorted body; the columns return to device 0
      const int64_t lf = c.first_leaf[g], ll = c.last_leaf[g + 1];
      const int64_t col0 = tb[lf] + c.first_off[g];
      const int64_t col1 = std::min<int64_t>(te[ll], tb[ll] + c.last_off[g + 1] + c.nt);
      cols[g] = {col0, col1};
      CK(h, cudaMemsetAsync((double *)dg.tbuf.p + col0 * 3, 0, (size_t)(col1 - col0) * 3 * sizeof(double), sg));
      a.out = (double *)dg.tbuf.p;
      a.obody = 1;
    }
    launch_sfs_leaf(kernel, c.nt, (unsigned)(k1 - k0), a, sg, mode);
    h->launches += 2;
    CK(h, cudaGetLastError());
  }
  if (G > 1) {
    NCK(h, g_nccl.group_start());
    for (int g = 1; g < G; ++g) {
      const int64_t col0 = cols[g].first, col1 = cols[g].second;
      if (col1 <= col0) continue;
      const size_t cnt = (size_t)(col1 - col0) * 3;
      NCK(h, g_nccl.send((double *)h->devs[g].tbuf.p + col0 * 3, cnt, kNcclFloat64, 0, h->comms[g], h->devs[g].stream));
      NCK(h, g_nccl.recv((double *)d.tbuf.p + col0 * 3, cnt, kNcclFloat64, g, h->comms[0], st));
    }
    NCK(h, g_nccl.group_end());
    CK(h, cudaSetDevice(d.id));
    for (int g = 0; g < G; ++g) {
      const int64_t col0 = cols[g].first, col1 = cols[g].second;
      if (col1 <= col0) continue;
      add_sorted3_kernel<<<blocks_for(col1 - col0, 256), 256, 0, st>>>((const double *)d.tbuf.p, c.d_tsort, col0, col1,
                                                                      (double *)d.sfs3.p);
      h->launches++;
    }
    CK(h, cudaGetLastError());
  }
  CK(h, cudaSetDevice(d.id));
  CK(h, cudaEventRecord(d.ev[4], st));
  TRY(d2h_rows(h, st, P + out_row, nf, (const double *)d.sfs3.p, 3, np));
  CK(h, cudaEventRecord(d.ev[5], st));
  for (int g = G - 1; g >= 0; --g) {
    CK(h, cudaSetDevice(h->devs[g].id));
    CK(h, cudaStreamSynchronize(h->devs[g].stream));
  }
  h->timing.uj_pairs = 0;
  h->timing.sfs_pairs = c.pairs;
  h1_fill_timing(h, d);
  h->np_resident = -1;
  return VPM_OK;
}


int vpm_estr_leafpairs(vpm_handle *h, double *P, int64_t nf, int64_t np, const int64_t *tsort,
                       const int64_t *ssort, const int64_t *tb, const int64_t *te, int64_t ntl,
                       const int64_t *sb, const int64_t *se, int64_t nsl, const int32_t *pt,
                       const int32_t *ps, int64_t npairs, int kernel, int flags) {
  return leafpairs_field(h, "vpm_estr_leafpairs", MODE_SFS, P, nf, np, tsort, ssort, tb, te, ntl, sb, se, nsl, pt,
                         ps, npairs, kernel, flags);
}

int vpm_zeta_leafpairs(vpm_handle *h, double *P, int64_t nf, int64_t np, const int64_t *sort_index,
                       const int64_t *lb, const int64_t *le, int64_t nl, const int32_t *pair_a,
                       const int32_t *pair_b, int64_t npairs, int kernel) {
  // zeta_fmm (src/FLOWVPM_viscous.jl:523-558): for a list entry (a, b) the bodies of leaf b
  // RECEIVE Gamma_j zeta_j from the bodies j of leaf a -> receivers are indexed by the second
  // element, givers by the first.
  return leafpairs_field(h, "vpm_zeta_leafpairs", MODE_ZETA, P, nf, np, sort_index, sort_index, lb, le, nl, lb, le,
                         nl, pair_b, pair_a, npairs, kernel, 0);
}

int vpm_zeta_direct(vpm_handle *h, double *P, int64_t nf, int64_t np, int kernel) {
  TRY(check_field(h, "vpm_zeta_direct", P, nf, np, kernel));
  if (np == 0) return VPM_OK;
  Dev &d = h->devs[0];
  cudaStream_t st = d.stream;
  h->launches = 0;
  CK(h, cudaSetDevice(d.id));
  TRY(ensure(h, d.in7, ((size_t)np * 7 + 16) * sizeof(double)));
  TRY(ensure(h, d.sfs3, (size_t)np * 3 * sizeof(double)));
  CK(h, cudaEventRecord(d.ev[0], st));
  TRY(h2d_rows(h, st, (double *)d.in7.p, P, nf, 7, np));
  CK(h, cudaEventRecord(d.ev[1], st));
  CK(h, cudaEventRecord(d.ev[2], st));
  CK(h, cudaEventRecord(d.ev[3], st));
  SrcView src{(const double *)d.in7.p, 7, 0, 3, 6};
  Plan sp;
  // the J operand is unused in zeta mode: the record builder reads 9 doubles per source
  // from it, so point it at in7 (stride 7; the allocation has 16 doubles of slack)
  TRY(sfs_sweep(h, d, st, kernel, (const double *)d.in7.p, 7, (const double *)d.in7.p, 7, nullptr, np, src,
                (const double *)d.in7.p, 7, 0, nullptr, 1, nullptr, np, VPM_FLAG_TRANSPOSED, sp, false, MODE_ZETA));
  SfsFinishArgs f;
  f.partial = (const double *)d.partial.p; f.pstride = sp.pstride; f.nsplit = sp.nsplit;
  f.nt = np; f.tindex = nullptr; f.out = (double *)d.sfs3.p; f.ld = 3; f.row = 0;
  f.accumulate = 0; f.reset = 0;  // zeta_direct zeroes J[1:3] of every particle first (:487-489)
  f.filter_static = 0; f.stat = nullptr; f.sld = 1;
  launch_sfs_finish(f, st);
  h->launches++;
  CK(h, cudaGetLastError());
  CK(h, cudaEventRecord(d.ev[4], st));
  TRY(d2h_rows(h, st, P + R_J, nf, (const double *)d.sfs3.p, 3, np));
  CK(h, cudaEventRecord(d.ev[5], st));
  CK(h, cudaStreamSynchronize(st));
  h->timing.uj_pairs = 0;
  h->timing.sfs_pairs = np * np;
  h1_fill_timing(h, d);
  h->np_resident = -1;
  return VPM_OK;
}

int vpm_leaflists_build(vpm_handle *h, const double *P, int64_t nf, int64_t np, int64_t ncrit, double theta,
                        int64_t *n_leaves, int64_t *n_pairs) {
  TRY(check_field(h, "vpm_leaflists_build", P, nf, np, 0));
  if (ncrit < 1 || !(theta > 0.0)) return fail(h, VPM_EINVAL, "vpm_leaflists_build: ncrit >= 1 and theta > 0 required");
  h->launches = 0;
  h->tree_np = -1;
  if (n_leaves) *n_leaves = 0;
  if (n_pairs) *n_pairs = 0;
  if (np == 0) { h->tree_np = 0; h->tree_nl = 0; h->tree_npairs = 0; return VPM_OK; }
  Dev &d = h->devs[0];
  bool has_static = false;
  CK(h, cudaSetDevice(d.id));
  CK(h, cudaEventRecord(d.ev[0], d.stream));
  const int G = (int)h->devs.size();
  if (G > 1 && host_is_pinned(P)) {
    // every device pulls 1/G of the rows over its own PCIe link, all-gather over NVLink (h2d_rows_sharded)
    for (int g = 0; g < G; ++g) {
      CK(h, cudaSetDevice(h->devs[g].id));
      TRY(ensure(h, h->devs[g].in7, (size_t)((np + G - 1) / G * G) * 7 * sizeof(double)));
    }
    TRY(h2d_rows_sharded(h, &Dev::in7, P, nf, 7, np));
    CK(h, cudaSetDevice(d.id));
  } else {
    TRY(h1_upload(h, d, P, nf, np, false, false, has_static));
  }
  CK(h, cudaEventRecord(d.ev[1], d.stream));
  TRY(tree_build(h, (const double *)d.in7.p, 7, 6, np, ncrit, theta));
  for (int e = 2; e <= 5; ++e) CK(h, cudaEventRecord(d.ev[e], d.stream));
  CK(h, cudaStreamSynchronize(d.stream));
  h->timing.uj_pairs = 0; h->timing.sfs_pairs = 0;
  h1_fill_timing(h, d);
  h->np_resident = -1;
  if (n_leaves) *n_leaves = h->tree_nl;
  if (n_pairs) *n_pairs = h->tree_npairs;
  return VPM_OK;
}

int vpm_leaflists_get(vpm_handle *h, int64_t *sort_index, int64_t *leaf_begin, int64_t *leaf_end, int32_t *pair_tgt,
                      int32_t *pair_src) {
  if (!h) return VPM_EINVAL;
  if (h->tree_np < 0) return fail(h, VPM_ESTATE, "vpm_leaflists_get: no leaf lists (call vpm_leaflists_build first)");
  if (h->tree_np == 0) return VPM_OK;
  Dev &d = h->devs[0];
  CK(h, cudaSetDevice(d.id));
  const TreeView v = tree_view(h);
  if (sort_index) CK(h, cudaMemcpyAsync(sort_index, v.sidx, (size_t)h->tree_np * 8, cudaMemcpyDeviceToHost, d.stream));
  if (leaf_begin) CK(h, cudaMemcpyAsync(leaf_begin, v.lbegin, (size_t)h->tree_nl * 8, cudaMemcpyDeviceToHost, d.stream));
  if (leaf_end) CK(h, cudaMemcpyAsync(leaf_end, v.lend, (size_t)h->tree_nl * 8, cudaMemcpyDeviceToHost, d.stream));
  if (pair_tgt && h->tree_npairs) CK(h, cudaMemcpyAsync(pair_tgt, v.pt, (size_t)h->tree_npairs * 4, cudaMemcpyDeviceToHost, d.stream));
  if (pair_src && h->tree_npairs) CK(h, cudaMemcpyAsync(pair_src, v.ps, (size_t)h->tree_npairs * 4, cudaMemcpyDeviceToHost, d.stream));
  CK(h, cudaStreamSynchronize(d.stream));
  return VPM_OK;
}

int vpm_uj_nearfield(vpm_handle *h, double *P, int64_t nf, int64_t np, int kernel, int flags) {
  const char *fn = "vpm_uj_nearfield";
  TRY(check_field(h, fn, P, nf, np, kernel));
  if (h->tree_np != np) return fail(h, VPM_ESTATE, "%s: leaf lists were built for %lld particles, field has %lld (call vpm_leaflists_build)", fn, (long long)h->tree_np, (long long)np);
  if (np == 0) return VPM_OK;
  h->launches = 0;
  int G = (int)h->devs.size();
  Dev &d0 = h->devs[0];
  CK(h, cudaSetDevice(d0.id));
  CK(h, cudaEventRecord(d0.ev[0], d0.stream));
  const bool reset = flags & VPM_FLAG_RESET;
  bool has_static = any_static(P, nf, np);
  const bool prior = !reset || has_static;
  // Sharded transfers (multi-GPU handle, page-locked matrix, nothing to accumulate on): every device pulls
  // its 1/G of X, Gamma, sigma over its own PCIe link (+ one all-gather over NVLink) and, at the end, pushes
  // its 1/G of the result rows itself; otherwise everything goes through device 0.
  bool sharded = G > 1 && !prior && host_is_pinned(P);
  const int64_t shard = (np + G - 1) / G;
  if (sharded) {
    for (int g = 0; g < G; ++g) {
      Dev &d = h->devs[g];
      CK(h, cudaSetDevice(d.id));
      TRY(ensure(h, d.in7, (size_t)shard * G * 7 * sizeof(double)));
      TRY(ensure(h, d.res18, (size_t)np * RES_ROWS * sizeof(double)));
      CK(h, cudaMemsetAsync(d.res18.p, 0, (size_t)np * RES_ROWS * sizeof(double), d.stream));
    }
    TRY(h2d_rows_sharded(h, &Dev::in7, P, nf, 7, np));
    CK(h, cudaSetDevice(d0.id));
  } else {
    TRY(h1_upload(h, d0, P, nf, np, !reset, false, has_static));
    if (!prior) CK(h, cudaMemsetAsync(d0.res18.p, 0, (size_t)np * RES_ROWS * sizeof(double), d0.stream));
  }
  // The resident lists (sort index, leaf ranges, near-field list) are functions of X and sigma as they were at
  // vpm_leaflists_build: once particles have moved the leaf spheres no longer bound their bodies and near pairs
  // would be dropped silently.  Compare the fingerprint of the rows just uploaded with the build's.
  {
    unsigned long long fp = 0;
    TRY(tree_fingerprint(h, d0, (const double *)d0.in7.p, 7, 6, np, &fp));
    if (fp != h->tree_fingerprint)
      return fail(h, VPM_ESTATE, "%s: positions or core sizes changed since vpm_leaflists_build; rebuild the leaf lists (they are valid for one particle configuration)", fn);
  }
  const TreeView tv = tree_view(h);
  DevCsr c;
  TRY(build_csr_device(h, fn, tv.lbegin, tv.lend, h->tree_nl, np, tv.lbegin, tv.lend, h->tree_nl, np, tv.pt, tv.ps,
                       h->tree_npairs, G, nullptr, 0, nullptr, 0, c, true));
  G = c.G_eff;
  CK(h, cudaEventRecord(d0.ev[1], d0.stream));
  const int64_t ns_pad = round_up(np, kTile);
  std::vector<LeafCsr> csr(G);
  csr[0] = c.csr;
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    CK(h, cudaSetDevice(d.id));
    TRY(ensure(h, d.tbuf, (size_t)np * 16 * sizeof(double)));
    TRY(ensure(h, d.sbuf, (size_t)np * 8 * sizeof(double)));
    TRY(ensure(h, d.rec, (size_t)ns_pad * kRec * sizeof(double)));
    if (g > 0) {
      if (!sharded) TRY(ensure(h, d.in7, (size_t)np * 7 * sizeof(double)));
      TRY(ensure(h, d.tree, h->devs[0].tree.cap));
      TRY(ensure(h, d.ibuf, h->devs[0].ibuf.cap));
      csr[g] = rebase_csr(c.csr, h->devs[0].ibuf.p, d.ibuf.p);
    }
  }
  // replicate state, sort index and list tables over NVLink; every device gathers its own
  // tree-sorted buffers
  if (G < (int)h->devs.size()) sharded = false;  // the cuts did not use every device: return through device 0
  if (G > 1 && !sharded) TRY(bcast_from_dev0(h, &Dev::in7, (size_t)np * 7 * sizeof(double)));
  if (G > 1) TRY(bcast_from_dev0(h, &Dev::tree, (size_t)np * 8));
  if (G > 1) TRY(bcast_from_dev0(h, &Dev::ibuf, c.bcast_bytes));
  std::vector<std::pair<int64_t, int64_t>> cols(G, {0, 0});
  // leaf tables on the host are not available (device-built): the column range of a device is
  // [begin of its first item, end of its last item), read from the cut records
  std::vector<int64_t> hb((size_t)h->tree_nl), he((size_t)h->tree_nl);
  if (G > 1) {
    CK(h, cudaSetDevice(d0.id));
    CK(h, cudaMemcpyAsync(hb.data(), tv.lbegin, hb.size() * 8, cudaMemcpyDeviceToHost, d0.stream));
    CK(h, cudaMemcpyAsync(he.data(), tv.lend, he.size() * 8, cudaMemcpyDeviceToHost, d0.stream));
    CK(h, cudaStreamSynchronize(d0.stream));
  }
  for (int g = 0; g < G; ++g) {
    Dev &d = h->devs[g];
    cudaStream_t st = d.stream;
    const int64_t k0 = c.cut[g], k1 = c.cut[g + 1];
    if (k1 <= k0) continue;
    CK(h, cudaSetDevice(d.id));
    const int64_t lf = c.first_leaf[g], ll = c.last_leaf[g + 1];
    const int64_t col0 = G == 1 ? 0 : hb[(size_t)lf] + c.first_off[g];
    const int64_t col1 = G == 1 ? np : std::min<int64_t>(he[(size_t)ll], hb[(size_t)ll] + c.last_off[g + 1] + c.nt);
    tree_gather_kernel<<<blocks_for(np, 256), 256, 0, st>>>((const double *)d.in7.p, 7, 0, 3, 6, (const int64_t *)d.tree.p,
                                                            np, (double *)d.sbuf.p, (double *)d.tbuf.p);
    LeafUjArgs a;
    a.csr = csr[g];
    a.csr.wi_leaf += k0;
    a.csr.wi_off += k0;
    a.tpos = (const double *)d.tbuf.p; a.tld = 16; a.rec = (const double *)d.rec.p;
    a.out = (double *)d.tbuf.p; a.urow = 4; a.jrow = 7; a.want_U = 1; a.want_J = 1;
    a.shortcut = (flags & VPM_FLAG_NO_FARFIELD_SHORTCUT) ? 0 : 1;
    if (g == 0) CK(h, cudaEventRecord(d.ev[6], st));
    launch_uj_leaf_any(h, d, st, kernel, c.nt, (unsigned)(k1 - k0), a, (const double *)d.sbuf.p, np, ns_pad);
    if (g == 0) CK(h, cudaEventRecord(d.ev[7], st));
    h->launches += 1;
    CK(h, cudaGetLastError());
    cols[g] = {col0, col1};
  }
  if (sharded) {
    // every device receives every device's columns of the sorted result (one ncclBroadcast per owner),
    // scatters the whole field into its own result block and writes its 1/G of the rows to the host
    NCK(h, g_nccl.group_start());
    for (int r = 0; r < G; ++r) {
      const int64_t col0 = cols[r].first, col1 = cols[r].second;
      if (col1 <= col0) continue;
      const size_t cnt = (size_t)(col1 - col0) * 16;
      for (int g = 0; g < G; ++g) {
        Dev &d = h->devs[g];
        NCK(h, g_nccl.broadcast((double *)h->devs[r].tbuf.p + col0 * 16, (double *)d.tbuf.p + col0 * 16, cnt, kNcclFloat64, r,
                                h->comms[g], d.stream));
      }
    }
    NCK(h, g_nccl.group_end());
    CK(h, cudaSetDevice(d0.id));
    CK(h, cudaEventRecord(d0.ev[2], d0.stream));
    for (int g = 0; g < G; ++g) {
      Dev &d = h->devs[g];
      CK(h, cudaSetDevice(d.id));
      tree_scatter_kernel<<<blocks_for(np, 256), 256, 0, d.stream>>>((const double *)d.tbuf.p, (const int64_t *)d.tree.p, np,
                                                                   (double *)d.res18.p, RES_ROWS, RES_U, RES_J, RES_W,
                                                                   RES_PSE, 1, nullptr, 1);
      CK(h, cudaGetLastError());
      if (g == 0) {
        h->launches++;
        CK(h, cudaEventRecord(d.ev[3], d.stream));
        CK(h, cudaEventRecord(d.ev[4], d.stream));
      }
      const int64_t c0 = std::min<int64_t>(np, g * shard), c1 = std::min<int64_t>(np, (g + 1) * shard);
      if (c1 > c0)
        CK(h, cudaMemcpy2DAsync(P + c0 * nf + R_U, nf * sizeof(double), (double *)d.res18.p + c0 * RES_ROWS,
                                RES_ROWS * sizeof(double), RES_ROWS * sizeof(double), (size_t)(c1 - c0),
                                cudaMemcpyDeviceToHost, d.stream));
      if (g == 0) CK(h, cudaEventRecord(d.ev[5], d.stream));
    }
  } else {
  // the devices return their columns of the sorted result to device 0 over NVLink
  // (ncclSend / ncclRecv; the receives are ordered on device 0's stream after its own
  // gather + pair kernel and before the scatter)
  if (G > 1) {
    NCK(h, g_nccl.group_start());
    for (int g = 1; g < G; ++g) {
      const int64_t col0 = cols[g].first, col1 = cols[g].second;
      if (col1 <= col0) continue;
      Dev &d = h->devs[g];
      const size_t cnt = (size_t)(col1 - col0) * 16;
      NCK(h, g_nccl.send((double *)d.tbuf.p + col0 * 16, cnt, kNcclFloat64, 0, h->comms[g], d.stream));
      NCK(h, g_nccl.recv((double *)d0.tbuf.p + col0 * 16, cnt, kNcclFloat64, g, h->comms[0], d0.stream));
    }
    NCK(h, g_nccl.group_end());
  }
  CK(h, cudaSetDevice(d0.id));
  CK(h, cudaEventRecord(d0.ev[2], d0.stream));
  tree_scatter_kernel<<<blocks_for(np, 256), 256, 0, d0.stream>>>((const double *)d0.tbuf.p, (const int64_t *)d0.tree.p, np,
                                                                 (double *)d0.res18.p, RES_ROWS, RES_U, RES_J, RES_W,
                                                                 RES_PSE, reset ? 1 : 0,
                                                                 has_static ? (const double *)d0.stat.p : nullptr, 1);
  h->launches++;
  CK(h, cudaGetLastError());
  CK(h, cudaEventRecord(d0.ev[3], d0.stream));
  CK(h, cudaEventRecord(d0.ev[4], d0.stream));
  TRY(h1_download(h, d0, P, nf, np, 0));
  }
  for (int g = G - 1; g >= 0; --g) {
    CK(h, cudaSetDevice(h->devs[g].id));
    CK(h, cudaStreamSynchronize(h->devs[g].stream));
  }
  h->timing.uj_pairs = c.pairs;
  h->timing.sfs_pairs = 0;
  h1_fill_timing(h, d0);
  h->timing.uj_ms = ev_ms(d0.ev[6], d0.ev[7]);  // the pair kernel of device 0 alone
  h->np_resident = -1;
  return VPM_OK;
}

}  // extern "C"
