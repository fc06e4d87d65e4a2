/* Plain-C driver of the C ABI (include/vpm_cuda.h): proves the boundary needs nothing but a
 * C compiler and dlopen-free linking.  Builds a small particle field in the reference's
 * 46-row column-major layout, calls vpm_uj_direct, and prints the rows it wrote so that
 * tests/test_c_driver_gpu.py can compare them with the oracle.
 *   gcc -std=c11 -I include tests/c_abi_driver.c -L flowvpm.jl_b200/csrc -lvpm_cuda -lm -o /tmp/c_abi_driver */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "vpm_cuda.h"

int main(int argc, char **argv) {
  const int64_t nf = 46, np = argc > 1 ? atoll(argv[1]) : 300;
  const int kernel = argc > 2 ? atoi(argv[2]) : VPM_KERNEL_WINCKELMANS;
  const int mode = argc > 3 ? atoi(argv[3]) : 0; /* 0: vpm_uj_direct; 1: device leaf lists + near field */
  double *P = calloc((size_t)(nf * np), sizeof(double));
  if (!P) return 2;
  /* a deterministic helix of particles: no RNG so that the Python side can rebuild it */
  for (int64_t i = 0; i < np; ++i) {
    double t = 0.05 * (double)i;
    double *p = P + nf * i;
    p[0] = cos(t); p[1] = sin(t); p[2] = 0.02 * t;              /* X */
    p[3] = -0.01 * sin(t); p[4] = 0.01 * cos(t); p[5] = 0.001;  /* Gamma */
    p[6] = 0.08 + 0.01 * sin(3 * t);                            /* sigma */
    p[42] = (i % 17 == 0) ? 1.0 : 0.0;                          /* static */
    p[9] = 0.5;                                                 /* a stale U that reset must clear */
  }
  vpm_handle *h = NULL;
  int rc = vpm_create(&h, 1, NULL);
  if (rc != VPM_OK) { fprintf(stderr, "vpm_create: %d %s\n", rc, vpm_last_error(NULL)); return 3; }
  if (mode == 1) {
    /* f-3: tree + near-field list on the device, then the near-field half of UJ_fmm over them */
    int64_t nl = 0, npairs = 0;
    rc = vpm_uj_nearfield(h, P, nf, np, kernel, VPM_FLAG_RESET);
    if (rc != VPM_ESTATE) { fprintf(stderr, "expected VPM_ESTATE before vpm_leaflists_build\n"); return 6; }
    rc = vpm_leaflists_build(h, P, nf, np, 16, 0.4, &nl, &npairs);
    if (rc != VPM_OK) { fprintf(stderr, "vpm_leaflists_build: %d %s\n", rc, vpm_last_error(h)); return 7; }
    int64_t *sidx = malloc((size_t)np * 8), *lb = malloc((size_t)nl * 8), *le = malloc((size_t)nl * 8);
    int32_t *pt = malloc((size_t)npairs * 4), *ps = malloc((size_t)npairs * 4);
    rc = vpm_leaflists_get(h, sidx, lb, le, pt, ps);
    if (rc != VPM_OK) { fprintf(stderr, "vpm_leaflists_get: %d %s\n", rc, vpm_last_error(h)); return 8; }
    int64_t covered = 0;
    for (int64_t l = 0; l < nl; ++l) covered += le[l] - lb[l];
    printf("lists %lld %lld %lld\n", (long long)nl, (long long)npairs, (long long)covered);
    free(sidx); free(lb); free(le); free(pt); free(ps);
    rc = vpm_uj_nearfield(h, P, nf, np, kernel, VPM_FLAG_RESET);
    if (rc != VPM_OK) { fprintf(stderr, "vpm_uj_nearfield: %d %s\n", rc, vpm_last_error(h)); return 9; }
  } else {
    rc = vpm_uj_direct(h, P, nf, np, kernel, VPM_FLAG_RESET | VPM_FLAG_RESET_SFS | VPM_FLAG_SFS | VPM_FLAG_TRANSPOSED);
    if (rc != VPM_OK) { fprintf(stderr, "vpm_uj_direct: %d %s\n", rc, vpm_last_error(h)); return 4; }
  }
  /* error path: an unknown kernel id must come back as a code with a message, not abort */
  rc = vpm_uj_direct(h, P, nf, np, 42, 0);
  if (rc != VPM_EINVAL || vpm_last_error(h)[0] == 0) { fprintf(stderr, "expected VPM_EINVAL\n"); return 5; }
  vpm_timing tm;
  vpm_get_timing(h, &tm);
  printf("abi %d gpus %d\n", vpm_abi_version(), vpm_num_devices(h));
  for (int64_t i = 0; i < np; ++i) {
    const double *p = P + nf * i;
    for (int r = 9; r < 12; ++r) printf("%.17g ", p[r]);
    for (int r = 15; r < 24; ++r) printf("%.17g ", p[r]);
    for (int r = 39; r < 42; ++r) printf("%.17g ", p[r]);
    printf("\n");
  }
  vpm_destroy(h);
  free(P);
  return 0;
}
