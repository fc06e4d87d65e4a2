// acc_probe.cu -- does the operand-reuse cache help the FP64 accumulate FMAs when several warps
// share a scheduler?  (the pair kernel behaves as if every DFMA with three
// distinct register operands cost 3 cycles, reuse flags ignored.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/acc_probe tools/acc_probe.cu && /tmp/acc_probe
// Per iteration a thread loads 16 fresh doubles from shared memory (as the pair kernel does with
// a source record) and issues 14 accumulate FMAs acc_k = fma(x, y, acc_k):
//   PATTERN 0: the pair kernel's operand pattern (U, W share A; J = b_i * d_j)
//   PATTERN 1: one shared multiplicand for all 14 (best case for the reuse cache)
//   PATTERN 2: 14 unrelated products (no reuse possible)
// Reported: FP64-pipe cycles per FMA per scheduler at 1, 2, 4 warps per scheduler.
#include <cstdio>
#include <cuda_runtime.h>

template <int PATTERN>
__global__ void __launch_bounds__(128) probe(double *out, int iters) {
  __shared__ double2 sm[8 * 64];
  // contents come from memory so that nothing about them is known at compile time
  for (int i = threadIdx.x; i < 8 * 64; i += blockDim.x) sm[i] = make_double2(out[8 + (i & 15)], out[24 + (i & 7)]);
  __syncthreads();
  double acc[14];
#pragma unroll
  for (int k = 0; k < 14; ++k) acc[k] = threadIdx.x * 1e-3 + k;
  const double t0 = 1.0 + threadIdx.x * 1e-6;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll 2
    for (int j = 0; j < 64; ++j) {
      const double2 v0 = sm[j * 8 + 0], v1 = sm[j * 8 + 1], v2 = sm[j * 8 + 2], v3 = sm[j * 8 + 3];
      const double2 v4 = sm[j * 8 + 4], v5 = sm[j * 8 + 5], v6 = sm[j * 8 + 6], v7 = sm[j * 8 + 7];
      // per-thread operands (as dx, c, b are in the pair kernel): one DADD each makes them thread-private
      const double A = v0.x + t0, cx = v0.y + t0, cy = v1.x + t0, cz = v1.y + t0;
      const double bx = v2.x + t0, by = v2.y + t0, bz = v3.x + t0, dx = v3.y + t0, dy = v4.x + t0, dz = v4.y + t0;
      const double gx = v5.x, gy = v5.y, gz = v6.x;  // per-source operands stay as loaded
      if (PATTERN == 0) {
        acc[0] = fma(A, cx, acc[0]); acc[1] = fma(A, cy, acc[1]); acc[2] = fma(A, cz, acc[2]);
        acc[11] = fma(A, gx, acc[11]); acc[12] = fma(A, gy, acc[12]); acc[13] = fma(A, gz, acc[13]);
        acc[3] = fma(bx, dx, acc[3]); acc[4] = fma(by, dx, acc[4]); acc[5] = fma(bz, dx, acc[5]);
        acc[6] = fma(bx, dy, acc[6]); acc[7] = fma(by, dy, acc[7]); acc[8] = fma(bz, dy, acc[8]);
        acc[9] = fma(bx, dz, acc[9]); acc[10] = fma(by, dz, acc[10]);
      } else if (PATTERN == 1) {
        const double o[14] = {cx, cy, cz, gx, gy, gz, bx, by, bz, dx, dy, dz, v6.y, v7.x};
#pragma unroll
        for (int k = 0; k < 14; ++k) acc[k] = fma(A, o[k], acc[k]);
      } else {
        const double p[14] = {A, cx, cy, cz, bx, by, bz, dx, dy, dz, gx, gy, gz, v6.y};
#pragma unroll
        for (int k = 0; k < 14; ++k) acc[k] = fma(p[k], p[(k + 5) % 14], acc[k]);  // 14 different products
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 14; ++k) s += acc[k];
  if (s == 12345.678) out[0] = s;
}

template <int PATTERN>
double cycles_per_fma(int blocks_per_sm, int sms, double khz, double *d_out) {
  const int iters = 512;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  probe<PATTERN><<<sms * blocks_per_sm, 128>>>(d_out, 8);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    probe<PATTERN><<<sms * blocks_per_sm, 128>>>(d_out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  // per scheduler: blocks_per_sm warps, each issuing iters*64*(14 FMA + 10 DADD) FP64 instructions
  const double cycles = best * 1e-3 * khz * 1e3;
  return cycles / ((double)blocks_per_sm * iters * 64.0 * 24.0);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  double *d_out;
  cudaMalloc(&d_out, 64 * sizeof(double));
  double h[64];
  for (int i = 0; i < 64; ++i) h[i] = 1.0 + 1e-7 * ((i * 37) % 64 - 32);
  cudaMemcpy(d_out, h, sizeof h, cudaMemcpyHostToDevice);
  printf("device %s, %d SMs, %d kHz; cycles per FP64 instruction per scheduler (14 accumulate DFMA + 10 DADD per source)\n",
         p.name, p.multiProcessorCount, p.clockRate);
  for (int wps = 1; wps <= 4; wps *= 2)
    printf("warps/scheduler %d: pair-kernel pattern %.3f   one shared multiplicand %.3f   unrelated products %.3f\n", wps,
           cycles_per_fma<0>(wps, p.multiProcessorCount, p.clockRate, d_out),
           cycles_per_fma<1>(wps, p.multiProcessorCount, p.clockRate, d_out),
           cycles_per_fma<2>(wps, p.multiProcessorCount, p.clockRate, d_out));
  return 0;
}
