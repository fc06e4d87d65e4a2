// vpm_host_lists.cuh -- leaf lists: device-side regrouping of direct_list (Hook 3), device-built lists (f-3), leaf-kernel launchers.
// Part of the single translation unit vpm_abi.cu (included there in order; not a standalone header).
#pragma once
namespace {

// ---- leaf-pair list (Hook 3): CSR by target leaf, built on device 0 (vpm_csr.cuh) ----
// carve aligned sub-arrays out of one device allocation
struct Carver {
  char *base;
  size_t off = 0;
  explicit Carver(void *p) : base((char *)p) {}
  template <class T>
  T *take(size_t n) {
    off = (off + 15) / 16 * 16;
    T *r = (T *)(base + off);
    off += n * sizeof(T);
    return r;
  }
};

struct DevCsr {
  LeafCsr csr;              // pointers into device 0's ibuf
  size_t bcast_bytes = 0;   // leading bytes of ibuf every device needs (tables + sort indices)
  int nt = kThreads;        // CTA width (targets per work item): 32, 64 or 128
  int64_t nwi = 0;          // work items
  int64_t pairs = 0;        // pair visits of the whole list
  int G_eff = 1;            // devices the work items are cut over: G, or 1 when the target leaves that
                            // carry work overlap (a device's targets must be one contiguous column range)
  std::vector<int64_t> cut;                     // [G + 1] work-item cuts
  std::vector<int64_t> first_leaf, first_off;   // [G + 1] item cut[g]     (first item of device g)
  std::vector<int64_t> last_leaf, last_off;     // [G + 1] item cut[g] - 1 (last item of device g - 1)
  const int64_t *d_tsort = nullptr, *d_ssort = nullptr;
};

// rebase the table pointers of device 0 onto another device's copy of ibuf
LeafCsr rebase_csr(const LeafCsr &c, const void *from, const void *to) {
  const ptrdiff_t shift = (const char *)to - (const char *)from;
  auto rb = [shift](auto *p) { return (decltype(p))((const char *)p + shift); };
  LeafCsr r;
  r.wi_leaf = rb(c.wi_leaf); r.wi_off = rb(c.wi_off);
  r.tleaf_begin = rb(c.tleaf_begin); r.tleaf_end = rb(c.tleaf_end);
  r.csr_ptr = rb(c.csr_ptr); r.csr_src = rb(c.csr_src);
  r.sleaf_begin = rb(c.sleaf_begin); r.sleaf_end = rb(c.sleaf_end);
  return r;
}

int build_csr_device(vpm_handle *h, const char *fn, const int64_t *tb, const int64_t *te, int64_t ntl,
                     int64_t n_tgt, const int64_t *sb, const int64_t *se, int64_t nsl, int64_t n_src,
                     const int32_t *pt, const int32_t *ps, int64_t npairs, int G, const int64_t *tsort,
                     int64_t n_tsort, const int64_t *ssort, int64_t n_ssort, DevCsr &out, bool dev_in = false,
                     const int64_t *gen_off = nullptr, const int64_t *gen_owner = nullptr) {
  // gen_off / gen_owner (host, ntl + 1 / ntl entries): pt and ps are NULL and the list is generated on the
  // device (csr_gen_pairs_kernel: the call shape of vpm_nearfield_ranges)
  // O(leaves) checks stay on the host; everything O(list entries) runs on the device.
  // dev_in: the tables are device arrays produced by vpm_leaflists_build (already valid).
  int64_t max_wi = dev_in ? n_tgt / 32 + ntl : 0;
  for (int64_t l = 0; l < ntl && !dev_in; ++l) {
    if (tb[l] < 0 || te[l] < tb[l] || te[l] > n_tgt)
      return fail(h, VPM_EINVAL, "%s: target leaf %lld range [%lld,%lld) outside 0..%lld", fn, (long long)l, (long long)tb[l], (long long)te[l], (long long)n_tgt);
    max_wi += (te[l] - tb[l] + 31) / 32;
  }
  for (int64_t l = 0; l < nsl && !dev_in; ++l)
    if (sb[l] < 0 || se[l] < sb[l] || se[l] > n_src)
      return fail(h, VPM_EINVAL, "%s: source leaf %lld range [%lld,%lld) outside 0..%lld", fn, (long long)l, (long long)sb[l], (long long)se[l], (long long)n_src);
  max_wi = std::max<int64_t>(max_wi, 1);
  Dev &d = h->devs[0];
  cudaStream_t st = d.stream;
  CK(h, cudaSetDevice(d.id));
  // replicated tables (ibuf) ...
  const size_t ibytes = 16 * 12 + (size_t)max_wi * 8 + (size_t)ntl * 16 + ((size_t)ntl + 1) * 8 + (size_t)npairs * 4 +
                        (size_t)nsl * 16 + (size_t)(n_tsort + n_ssort) * 8;
  TRY(ensure(h, d.ibuf, ibytes));
  Carver cv(d.ibuf.p);
  int32_t *wl = cv.take<int32_t>((size_t)max_wi), *wo = cv.take<int32_t>((size_t)max_wi);
  int64_t *dtb = cv.take<int64_t>((size_t)ntl), *dte = cv.take<int64_t>((size_t)ntl);
  u64 *dptr = cv.take<u64>((size_t)ntl + 1);
  int32_t *dsrc = cv.take<int32_t>((size_t)npairs);
  int64_t *dsb = cv.take<int64_t>((size_t)nsl), *dse = cv.take<int64_t>((size_t)nsl);
  int64_t *dts = cv.take<int64_t>((size_t)n_tsort), *dss = cv.take<int64_t>((size_t)n_ssort);
  out.bcast_bytes = cv.off;
  // ... and device-0 scratch
  const size_t sbytes = 16 * 12 + (size_t)npairs * 4 + (size_t)ntl * 8 * 3 + (size_t)max_wi * 8 + CS_SLOTS * 8 +
                        (size_t)(G + 1) * 40 + (size_t)ntl * 16;
  TRY(ensure(h, d.scr, sbytes));
  Carver sc(d.scr.p);
  int32_t *dpt = sc.take<int32_t>((size_t)npairs);
  u64 *srcw = sc.take<u64>((size_t)ntl), *wcnt = sc.take<u64>((size_t)ntl), *wofs = sc.take<u64>((size_t)ntl);
  u64 *wiw = sc.take<u64>((size_t)max_wi);
  u64 *stats = sc.take<u64>(CS_SLOTS);
  int64_t *dcut = sc.take<int64_t>((size_t)(G + 1) * 5);
  int64_t *tb_sorted = sc.take<int64_t>((size_t)ntl);
  int32_t *iota = sc.take<int32_t>((size_t)ntl), *ord = sc.take<int32_t>((size_t)ntl);
  // cub temporary storage: the largest of the four calls below
  const int key_bits = std::max(1, (int)std::ceil(std::log2((double)std::max<int64_t>(ntl, 2))));
  size_t tmp = 0, t1 = 0;
  cub::DeviceScan::InclusiveSum(nullptr, t1, dptr, dptr, (int64_t)ntl + 1, st); tmp = std::max(tmp, t1);
  cub::DeviceScan::ExclusiveSum(nullptr, t1, wcnt, wofs, (int64_t)ntl, st); tmp = std::max(tmp, t1);
  cub::DeviceScan::InclusiveSum(nullptr, t1, wiw, wiw, max_wi, st); tmp = std::max(tmp, t1);
  cub::DeviceRadixSort::SortPairs(nullptr, t1, (const int32_t *)nullptr, (int32_t *)nullptr, (const int32_t *)nullptr,
                                  (int32_t *)nullptr, npairs, 0, key_bits, st);
  tmp = std::max(tmp, t1);
  cub::DeviceRadixSort::SortPairs(nullptr, t1, (const int64_t *)nullptr, (int64_t *)nullptr, (const int32_t *)nullptr,
                                  (int32_t *)nullptr, ntl, 0, 64, st);
  tmp = std::max(tmp, t1);
  TRY(ensure(h, d.cubtmp, tmp + 16));

  // (host tables of a pageable caller go through the pinned ring; device-resident tables are plain copies)
  auto put = [&](auto *dst, const auto *srcp, size_t n) -> int {
    return h2d_contig(h, st, (void *)dst, (const void *)srcp, n * sizeof(*srcp));
  };
  TRY(put(dtb, tb, (size_t)ntl));
  TRY(put(dte, te, (size_t)ntl));
  TRY(put(dsb, sb, (size_t)nsl));
  TRY(put(dse, se, (size_t)nsl));
  if (gen_off) {
    // (the two small tables live in d.scr2, idle at this point: the unsorted-list branch below takes it over only
    // after the stream has been synchronised)
    TRY(ensure(h, d.scr2, (size_t)(2 * ntl + 2) * 8 + 64));
    int64_t *goff = (int64_t *)d.scr2.p, *gown = goff + (ntl + 1);
    TRY(put(goff, gen_off, (size_t)ntl + 1));
    TRY(put(gown, gen_owner, (size_t)ntl));
    csr_gen_pairs_kernel<<<blocks_for(npairs, 256), 256, 0, st>>>(goff, gown, ntl, npairs, dpt, dsrc);
    h->launches++;
  } else {
    TRY(put(dpt, pt, (size_t)npairs));
    TRY(put(dsrc, ps, (size_t)npairs));  // already the CSR column array when the list is grouped
  }
  if (n_tsort) TRY(put(dts, tsort, (size_t)n_tsort));
  if (n_ssort) TRY(put(dss, ssort, (size_t)n_ssort));
  csr_init_stats_kernel<<<1, 32, 0, st>>>(stats);
  csr_zero_kernel<<<blocks_for(ntl + 1, 256), 256, 0, st>>>(dptr, ntl + 1);
  csr_zero_kernel<<<blocks_for(ntl, 256), 256, 0, st>>>(srcw, ntl);
  csr_count_kernel<<<blocks_for(npairs, 256), 256, 0, st>>>(dpt, dsrc, npairs, ntl, nsl, dsb, dse, dptr, srcw, stats);
  csr_cand_kernel<<<blocks_for(ntl, 256), 256, 0, st>>>(dtb, dte, ntl, srcw, stats);
  t1 = d.cubtmp.cap;
  CK(h, cub::DeviceScan::InclusiveSum(d.cubtmp.p, t1, dptr, dptr, (int64_t)ntl + 1, st));
  h->launches += 6;
  u64 hs[CS_SLOTS];
  CK(h, cudaMemcpyAsync(hs, stats, sizeof hs, cudaMemcpyDeviceToHost, st));
  CK(h, cudaStreamSynchronize(st));
  CK(h, cudaGetLastError());
  if (hs[CS_BAD] != ~0ull) {
    const int64_t k = (int64_t)hs[CS_BAD] - 1;
    return fail(h, VPM_EINVAL, "%s: pair %lld = (%d,%d) outside the leaf tables", fn, (long long)k,
                (dev_in || !pt) ? -1 : pt[k], (dev_in || !ps) ? -1 : ps[k]);
  }
  // CTA width: minimise the padded lane-work  sum_leaf ceil(size/NT)*NT * (its source bodies);
  // wider CTAs amortise the tile traffic better: require a 10 % gain to go narrower
  double best = -1.0;
  const int cands[3] = {128, 64, 32};
  const u64 wsum[3] = {hs[CS_W128], hs[CS_W64], hs[CS_W32]};
  for (int c = 0; c < 3; ++c)
    if (best < 0.0 || (double)wsum[c] < 0.9 * best) { best = (double)wsum[c]; out.nt = cands[c]; }
  out.pairs = (int64_t)hs[CS_PAIRS];
  if (hs[CS_UNSORTED]) {
    // stable radix sort by target leaf keeps the list order inside each group
    TRY(ensure(h, d.scr2, (size_t)npairs * 8 + 32));
    Carver s2(d.scr2.p);
    int32_t *keys_out = s2.take<int32_t>((size_t)npairs), *vals_out = s2.take<int32_t>((size_t)npairs);
    t1 = d.cubtmp.cap;
    CK(h, cub::DeviceRadixSort::SortPairs(d.cubtmp.p, t1, (const int32_t *)dpt, keys_out, (const int32_t *)dsrc, vals_out,
                                          npairs, 0, key_bits, st));
    CK(h, cudaMemcpyAsync(dsrc, vals_out, (size_t)npairs * 4, cudaMemcpyDeviceToDevice, st));
    h->launches += 1;
  }
  // target leaves in body order (stable: leaves with the same first body keep their index order)
  csr_iota_kernel<<<blocks_for(ntl, 256), 256, 0, st>>>(iota, ntl);
  t1 = d.cubtmp.cap;
  CK(h, cub::DeviceRadixSort::SortPairs(d.cubtmp.p, t1, (const int64_t *)dtb, tb_sorted, (const int32_t *)iota, ord, ntl, 0,
                                        64, st));
  csr_wi_count_kernel<<<blocks_for(ntl, 256), 256, 0, st>>>(dtb, dte, ntl, dptr, ord, out.nt, wcnt);
  t1 = d.cubtmp.cap;
  CK(h, cub::DeviceScan::ExclusiveSum(d.cubtmp.p, t1, wcnt, wofs, (int64_t)ntl, st));
  csr_zero_kernel<<<blocks_for(max_wi, 256), 256, 0, st>>>(wiw, max_wi);
  csr_wi_fill_kernel<<<blocks_for(ntl, 256), 256, 0, st>>>(dtb, dte, ntl, wofs, wcnt, srcw, ord, out.nt, wl, wo, wiw,
                                                          stats);
  h->launches += 2;
  t1 = d.cubtmp.cap;
  CK(h, cub::DeviceScan::InclusiveSum(d.cubtmp.p, t1, wiw, wiw, max_wi, st));
  h->launches += 5;
  CK(h, cudaMemcpyAsync(hs, stats, sizeof hs, cudaMemcpyDeviceToHost, st));
  CK(h, cudaStreamSynchronize(st));
  out.nwi = (int64_t)hs[CS_NWI];
  out.G_eff = hs[CS_OVERLAP] ? 1 : G;
  G = out.G_eff;
  out.cut.assign((size_t)G + 1, out.nwi);
  out.cut[0] = 0;
  for (auto *v : {&out.first_leaf, &out.first_off, &out.last_leaf, &out.last_off}) v->assign((size_t)G + 1, 0);
  if (out.nwi > 0) {
    if (G + 1 > 32) return fail(h, VPM_EINVAL, "%s: more than 31 devices", fn);
    csr_cut_kernel<<<1, 32, 0, st>>>(wiw, wl, wo, out.nwi, G, dcut);
    h->launches += 1;
    std::vector<int64_t> hc((size_t)(G + 1) * 5);
    CK(h, cudaMemcpyAsync(hc.data(), dcut, hc.size() * 8, cudaMemcpyDeviceToHost, st));
    CK(h, cudaStreamSynchronize(st));
    for (int g = 0; g <= G; ++g) {
      const int64_t *c = &hc[(size_t)5 * g];
      out.cut[(size_t)g] = c[0];
      out.first_leaf[(size_t)g] = c[1]; out.first_off[(size_t)g] = c[2];
      out.last_leaf[(size_t)g] = c[3]; out.last_off[(size_t)g] = c[4];
    }
  }
  CK(h, cudaGetLastError());
  out.csr.wi_leaf = wl; out.csr.wi_off = wo; out.csr.tleaf_begin = dtb; out.csr.tleaf_end = dte;
  out.csr.csr_ptr = (const int64_t *)dptr; out.csr.csr_src = dsrc; out.csr.sleaf_begin = dsb; out.csr.sleaf_end = dse;
  out.d_tsort = dts; out.d_ssort = dss;
  return VPM_OK;
}


// ---- device-built leaf lists (vpm_tree.cuh) --------------------------------------
// Builds sort index, leaf ranges and the near-field list from the rows X (3) and sigma of a
// device-resident column-major matrix view.  Results stay on device 0 (d.tree, d.tlist).
struct TreeView {
  int64_t *sidx, *lbegin, *lend;  // [np], [nl], [nl]
  int32_t *pt, *ps;               // [npairs]
};
TreeView tree_view(vpm_handle *h) {
  Dev &d = h->devs[0];
  TreeView v;
  Carver cv(d.tree.p);
  const size_t n = (size_t)std::max<int64_t>(h->tree_np, 1);
  v.sidx = cv.take<int64_t>(n); v.lbegin = cv.take<int64_t>(n); v.lend = cv.take<int64_t>(n);
  Carver cl(d.tlist.p);
  const size_t m = (size_t)std::max<int64_t>(h->tree_npairs, 1);
  v.pt = cl.take<int32_t>(m); v.ps = cl.take<int32_t>(m);
  return v;
}

// fingerprint of X and sigma of a device copy of the rows (compact 7-row layout: ld = 7, sigma at 6)
int tree_fingerprint(vpm_handle *h, Dev &d, const double *d_P, int64_t ld, int osig, int64_t np, unsigned long long *out) {
  TRY(ensure(h, d.cubtmp, 256));
  unsigned long long *acc = (unsigned long long *)d.cubtmp.p + 8;  // (the first words serve the near-fraction sample)
  CK(h, cudaMemsetAsync(acc, 0, sizeof *acc, d.stream));
  tree_fingerprint_kernel<<<blocks_for(np, 256), 256, 0, d.stream>>>(d_P, ld, 0, osig, np, acc);
  CK(h, cudaMemcpyAsync(out, acc, sizeof *out, cudaMemcpyDeviceToHost, d.stream));
  CK(h, cudaStreamSynchronize(d.stream));
  return VPM_OK;
}

int tree_build(vpm_handle *h, const double *d_P, int64_t ld, int osig, int64_t np, int64_t ncrit, double theta) {
  const char *fn = "vpm_leaflists_build";
  Dev &d = h->devs[0];
  cudaStream_t st = d.stream;
  CK(h, cudaSetDevice(d.id));
  h->tree_np = -1;
  const size_t n = (size_t)np;
  TRY(ensure(h, d.tree, 3 * n * 8 + 64));
  // scratch: bb[8] | keys | idx | skeys | rank (u64) | lkey | ctr[3] | rad | cnt | ofs
  TRY(ensure(h, d.scr, 16 * 16 + 8 * 8 + n * 8 * 11));
  Carver sc(d.scr.p);
  long long *bb = sc.take<long long>(8);
  int64_t *keys = sc.take<int64_t>(n), *idx0 = sc.take<int64_t>(n), *skeys = sc.take<int64_t>(n);
  u64 *rank = sc.take<u64>(n);
  int64_t *lkey = sc.take<int64_t>(n);
  double *ctr = sc.take<double>(3 * n), *rad = sc.take<double>(n);
  u64 *cnt = sc.take<u64>(n), *ofs = sc.take<u64>(n);
  Carver tv(d.tree.p);
  int64_t *sidx = tv.take<int64_t>(n), *lbegin = tv.take<int64_t>(n), *lend = tv.take<int64_t>(n);

  tree_bbox_init_kernel<<<1, 32, 0, st>>>(bb);
  tree_bbox_kernel<<<blocks_for(np, 256), 256, 0, st>>>(d_P, ld, np, bb);
  h->launches += 2;
  long long hb[8];
  CK(h, cudaMemcpyAsync(hb, bb, sizeof hb, cudaMemcpyDeviceToHost, st));
  CK(h, cudaStreamSynchronize(st));
  CK(h, cudaGetLastError());
  // grid: cell size for a mean occupancy of ncrit/2; thin directions are padded to 1e-3 of
  // the largest extent so that planar / linear fields do not explode the cell count
  TreeGrid g;
  double ext[3], emax = 0.0;
  for (int a = 0; a < 3; ++a) {
    g.lo[a] = ordered_to_dbl(hb[a]);
    ext[a] = ordered_to_dbl(hb[3 + a]) - g.lo[a];
    if (!std::isfinite(ext[a])) return fail(h, VPM_EINVAL, "%s: non-finite particle positions", fn);
    emax = std::max(emax, ext[a]);
  }
  for (int a = 0; a < 3; ++a) ext[a] = std::max(std::max(ext[a], 1e-3 * emax), 1e-300);
  const double vol = ext[0] * ext[1] * ext[2];
  g.h = std::pow(vol * ((double)ncrit / 2.0) / (double)np, 1.0 / 3.0);
  if (!(g.h > 0.0) || !std::isfinite(g.h)) g.h = 1.0;
  g.theta = theta;
  auto dims_of = [&](double hh, int64_t dims[3]) {
    double ncell_d = 1.0;
    for (int a = 0; a < 3; ++a) {
      const double c = std::max(1.0, std::ceil(ext[a] / hh));
      dims[a] = (int64_t)std::min(c, 4.0e18);
      ncell_d *= c;
    }
    return ncell_d;
  };
  if (dims_of(g.h, g.dims) > 1.0e9)
    return fail(h, VPM_EINVAL, "%s: too many grid cells (field too anisotropic for ncrit = %lld)", fn, (long long)ncrit);
  size_t tmp = 0, t1 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, t1, (const int64_t *)nullptr, (int64_t *)nullptr, (const int64_t *)nullptr,
                                  (int64_t *)nullptr, np, 0, 64, st);
  tmp = std::max(tmp, t1);
  cub::DeviceScan::InclusiveSum(nullptr, t1, rank, rank, np, st); tmp = std::max(tmp, t1);
  cub::DeviceScan::ExclusiveSum(nullptr, t1, cnt, ofs, np, st); tmp = std::max(tmp, t1);
  TRY(ensure(h, d.cubtmp, tmp + 16));
  // The first cell size assumes the field fills its bounding box.  Fields that do not (rings,
  // jets) leave most cells empty and the occupied ones overfull: shrink the cells by the
  // cube root of the overfill and sort again (at most twice; a sort is ~1 ms per million).
  int64_t nl = 0, ncell = 0;
  for (int iter = 0;; ++iter) {
    ncell = g.dims[0] * g.dims[1] * g.dims[2];
    const int key_bits = std::max(1, (int)std::ceil(std::log2((double)std::max<int64_t>(ncell, 2))));
    tree_keys_kernel<<<blocks_for(np, 256), 256, 0, st>>>(d_P, ld, np, g, keys, idx0);
    t1 = d.cubtmp.cap;
    CK(h, cub::DeviceRadixSort::SortPairs(d.cubtmp.p, t1, (const int64_t *)keys, skeys, (const int64_t *)idx0, sidx, np,
                                          0, key_bits, st));
    tree_heads_kernel<<<blocks_for(np, 256), 256, 0, st>>>(skeys, np, rank);
    t1 = d.cubtmp.cap;
    CK(h, cub::DeviceScan::InclusiveSum(d.cubtmp.p, t1, rank, rank, np, st));
    h->launches += 4;
    u64 nl64 = 0;
    CK(h, cudaMemcpyAsync(&nl64, rank + (np - 1), 8, cudaMemcpyDeviceToHost, st));
    CK(h, cudaStreamSynchronize(st));
    nl = (int64_t)nl64;
    const double occ = (double)np / (double)nl;
    if (iter >= 2 || occ <= 0.75 * (double)ncrit) break;
    const double h2 = g.h * std::pow(((double)ncrit / 2.0) / occ, 1.0 / 3.0);
    int64_t dims2[3];
    const double ncell2 = dims_of(h2, dims2);
    if (!(h2 > 0.0) || ncell2 > 1.0e9 || ncell2 > 64.0 * (double)np + 4096.0) break;
    g.h = h2;
    for (int a = 0; a < 3; ++a) g.dims[a] = dims2[a];
  }
  TRY(ensure(h, d.scr2, (size_t)ncell * 4 + 64));
  int32_t *cell_to_leaf = (int32_t *)d.scr2.p;
  tree_fill_i32_kernel<<<blocks_for(ncell, 256), 256, 0, st>>>(cell_to_leaf, ncell, -1);
  tree_leaves_kernel<<<blocks_for(np, 256), 256, 0, st>>>(skeys, rank, np, lbegin, lend, lkey, cell_to_leaf);
  h->launches += 2;
  tree_spheres_kernel<<<blocks_for(nl * 32, 256), 256, 0, st>>>(d_P, ld, osig, sidx, lbegin, lend, nl, ctr, rad, bb);
  h->launches++;
  CK(h, cudaMemcpyAsync(hb, bb, sizeof hb, cudaMemcpyDeviceToHost, st));
  CK(h, cudaStreamSynchronize(st));
  CK(h, cudaGetLastError());
  const double rmax = ordered_to_dbl(hb[6]);
  if (!std::isfinite(rmax)) return fail(h, VPM_EINVAL, "%s: non-finite leaf radius (core sizes)", fn);
  const double reach_d = std::ceil(2.0 * rmax / (theta * g.h)) + 1.0;
  int reach = (int)std::min(reach_d, 1.0e6);
  // no leaf is further than the grid itself
  reach = (int)std::min<int64_t>(reach, std::max(std::max(g.dims[0], g.dims[1]), g.dims[2]));
  {
    // the stencil search costs nl * prod_a min(2 reach + 1, dims_a) MAC tests: refuse fields whose
    // leaf radii (core sizes) are so large against the cell size that this would run for minutes
    double cand = (double)nl;
    for (int a = 0; a < 3; ++a) cand *= (double)std::min<int64_t>(2 * (int64_t)reach + 1, 2 * g.dims[a] - 1);
    if (cand > 1.0e11)
      return fail(h, VPM_EINVAL, "%s: %.2g leaf-pair tests (largest leaf radius %.3g against cell size %.3g): "
                  "core sizes too large for ncrit = %lld, use a larger ncrit", fn, cand, rmax, g.h, (long long)ncrit);
  }
  tree_list_kernel<0><<<blocks_for(nl * 32, 256), 256, 0, st>>>(g, reach, lkey, cell_to_leaf, ctr, rad, nl, cnt, nullptr,
                                                            nullptr, nullptr);
  t1 = d.cubtmp.cap;
  CK(h, cub::DeviceScan::ExclusiveSum(d.cubtmp.p, t1, cnt, ofs, nl, st));
  h->launches += 2;
  u64 last[2] = {0, 0};
  CK(h, cudaMemcpyAsync(&last[0], cnt + (nl - 1), 8, cudaMemcpyDeviceToHost, st));
  CK(h, cudaMemcpyAsync(&last[1], ofs + (nl - 1), 8, cudaMemcpyDeviceToHost, st));
  CK(h, cudaStreamSynchronize(st));
  const int64_t npairs = (int64_t)(last[0] + last[1]);
  TRY(ensure(h, d.tlist, (size_t)std::max<int64_t>(npairs, 1) * 8 + 64));
  Carver cl(d.tlist.p);
  int32_t *pt = cl.take<int32_t>((size_t)std::max<int64_t>(npairs, 1)), *ps = cl.take<int32_t>((size_t)std::max<int64_t>(npairs, 1));
  tree_list_kernel<1><<<blocks_for(nl * 32, 256), 256, 0, st>>>(g, reach, lkey, cell_to_leaf, ctr, rad, nl, cnt, ofs, pt, ps);
  h->launches++;
  CK(h, cudaStreamSynchronize(st));
  CK(h, cudaGetLastError());
  TRY(tree_fingerprint(h, d, d_P, ld, osig, np, &h->tree_fingerprint));
  h->tree_np = np; h->tree_nl = nl; h->tree_npairs = npairs;
  h->tree_ncrit = ncrit; h->tree_theta = theta;
  return VPM_OK;
}

template <int K>
void launch_uj_leaf_K(int nt, unsigned nwi, const LeafUjArgs &a, cudaStream_t st) {
  if (nt == 32) uj_leaf1w_kernel<K, 64><<<nwi, 32, 0, st>>>(a);
  else if (nt == 64) uj_leaf_kernel<K, 64, 64><<<nwi, 64, 0, st>>>(a);
  else uj_leaf_kernel<K, 128, 128><<<nwi, 128, 0, st>>>(a);
}
void launch_uj_leaf(int kernel, int nt, unsigned nwi, const LeafUjArgs &a, cudaStream_t st) {
  switch (kernel) {
    case K_SING: launch_uj_leaf_K<K_SING>(nt, nwi, a, st); break;
    case K_GAUS: launch_uj_leaf_K<K_GAUS>(nt, nwi, a, st); break;
    case K_GERF: launch_uj_leaf_K<K_GERF>(nt, nwi, a, st); break;
    default: launch_uj_leaf_K<K_WINCK>(nt, nwi, a, st); break;
  }
}
template <int K>
void launch_uj_leaf_f32_K(int nt, unsigned nwi, const LeafUjArgsF &a, cudaStream_t st) {
  if (nt == 32) uj_leaf_kernel_f32<K, 32, 64><<<nwi, 32, 0, st>>>(a);
  else if (nt == 64) uj_leaf_kernel_f32<K, 64, 64><<<nwi, 64, 0, st>>>(a);
  else uj_leaf_kernel_f32<K, 128, 128><<<nwi, 128, 0, st>>>(a);
}
void launch_uj_leaf_f32(int kernel, int nt, unsigned nwi, const LeafUjArgsF &a, cudaStream_t st) {
  switch (kernel) {
    case K_SING: launch_uj_leaf_f32_K<K_SING>(nt, nwi, a, st); break;
    case K_GAUS: launch_uj_leaf_f32_K<K_GAUS>(nt, nwi, a, st); break;
    case K_GERF: launch_uj_leaf_f32_K<K_GERF>(nt, nwi, a, st); break;
    default: launch_uj_leaf_f32_K<K_WINCK>(nt, nwi, a, st); break;
  }
}
// records + leaf kernel of one device's share of the work items, FP64 or (option) FP32 arithmetic
void launch_uj_leaf_any(vpm_handle *h, Dev &d, cudaStream_t st, int kernel, int nt, unsigned nwi, const LeafUjArgs &a,
                        SrcView sv, int64_t n_src, int64_t ns_pad) {
  if (d.scratch_pending && !h->capturing) cudaStreamWaitEvent(st, d.scratch_ev, 0);  // d.rec may still be read by a _device sweep
  if (h->opt_nearfield_fp32) {
    prep_uj_records_f32s<<<blocks_for(ns_pad, 256), 256, 0, st>>>(sv, 0, n_src, ns_pad, kernel, (float *)d.rec.p);
    LeafUjArgsF f;
    f.csr = a.csr; f.tpos = a.tpos; f.tld = a.tld; f.rec = (const float *)d.rec.p; f.out = a.out;
    f.urow = a.urow; f.jrow = a.jrow; f.want_U = a.want_U; f.want_J = a.want_J; f.shortcut = a.shortcut;
    launch_uj_leaf_f32(kernel, nt, nwi, f, st);
  } else {
    if (kernel == K_GERF)  // kLeafTab: the leaf kernels of this family read the G(u) table
      prep_uj_records_tab<<<blocks_for(ns_pad, 256), 256, 0, st>>>(sv, 0, n_src, ns_pad, kernel, (double *)d.rec.p);
    else
      prep_uj_records<<<blocks_for(ns_pad, 256), 256, 0, st>>>(sv, 0, n_src, ns_pad, kernel, (double *)d.rec.p);
    launch_uj_leaf(kernel, nt, nwi, a, st);
  }
  h->launches += 2;
}
// the same for FastMultipole's 8-row source buffer [x y z rho Gx Gy Gz sigma]
void launch_uj_leaf_any(vpm_handle *h, Dev &d, cudaStream_t st, int kernel, int nt, unsigned nwi, const LeafUjArgs &a,
                        const double *sbuf8, int64_t n_src, int64_t ns_pad) {
  launch_uj_leaf_any(h, d, st, kernel, nt, nwi, a, SrcView{sbuf8, 8, 0, 4, 7}, n_src, ns_pad);
}
template <int K, int MODE>
void launch_sfs_leaf_K(int nt, unsigned nwi, const LeafSfsArgs &a, cudaStream_t st) {
  if (nt == 32) sfs_leaf_kernel<K, 32, 64, MODE><<<nwi, 32, 0, st>>>(a);
  else if (nt == 64) sfs_leaf_kernel<K, 64, 64, MODE><<<nwi, 64, 0, st>>>(a);
  else sfs_leaf_kernel<K, 128, 128, MODE><<<nwi, 128, 0, st>>>(a);
}
template <int MODE>
void launch_sfs_leaf_M(int kernel, int nt, unsigned nwi, const LeafSfsArgs &a, cudaStream_t st) {
  switch (kernel) {
    case K_SING: launch_sfs_leaf_K<K_SING, MODE>(nt, nwi, a, st); break;
    case K_GAUS: launch_sfs_leaf_K<K_GAUS, MODE>(nt, nwi, a, st); break;
    case K_GERF: launch_sfs_leaf_K<K_GERF, MODE>(nt, nwi, a, st); break;
    default: launch_sfs_leaf_K<K_WINCK, MODE>(nt, nwi, a, st); break;
  }
}
void launch_sfs_leaf(int kernel, int nt, unsigned nwi, const LeafSfsArgs &a, cudaStream_t st,
                     int mode = MODE_SFS) {
  if (mode == MODE_ZETA) launch_sfs_leaf_M<MODE_ZETA>(kernel, nt, nwi, a, st);
  else launch_sfs_leaf_M<MODE_SFS>(kernel, nt, nwi, a, st);
}

}  // namespace
