// dfma_probe.cu -- B200 FP64 pipe microbenchmarks (tuning aid, run on the GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dfma_probe tools/dfma_probe.cu && /tmp/dfma_probe
// Measures the DFMA issue rate as a function of how many DISTINCT register operands the
// instruction reads, of warps per scheduler, and the dependent-issue latency.
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
#define INNER 16

// mode 0: a = fma(a, m, c)        m, c loop-invariant (uniform / reused)   -> 1 distinct reg
// mode 1: a_i = fma(b_i, c_i, a_i) b_i, c_i per chain                       -> 3 distinct regs
// mode 2: a_i = fma(a_i, b_i, c)   b_i per chain, c shared                  -> 2-3 distinct
// mode 3: a_i = b_i * c_i + ... DMUL with 2 distinct regs
template <int MODE>
__global__ void probe(double *out, int iters, double seed) {
  double a[CHAINS], b[CHAINS], c[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) {
    a[i] = seed + i + threadIdx.x * 1e-3;
    b[i] = 0.999999 - i * 1e-7 + threadIdx.x * 1e-9;
    c[i] = 1e-9 * (i + 1) + threadIdx.x * 1e-12;
  }
  const double m = 0.999999, cc = 1e-9;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < INNER; ++k) {
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) {
        if (MODE == 0) a[i] = fma(a[i], m, cc);
        if (MODE == 1) a[i] = fma(b[i], c[i], a[i]);
        if (MODE == 2) a[i] = fma(a[i], b[i], c[0]);
        if (MODE == 3) a[i] = a[i] * b[i];
        if (MODE == 4) a[i] = fma(b[i], c[(i + 1) % CHAINS], a[i]);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += a[i] + b[i] + c[i];
  if (s == 12345.678) out[0] = s;
}

// dependent chain latency: one warp per SM
__global__ void latency(double *out, int iters, double seed, long long *cycles) {
  double a = seed, b = 0.999999 + threadIdx.x * 1e-9, c = 1e-9;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 64; ++k) a = fma(a, b, c);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  if (a == 12345.678) out[0] = a;
}

template <int MODE>
double run(int blocks_per_sm, int threads, int sms, double *d_out) {
  int iters = 2048;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  probe<MODE><<<sms * blocks_per_sm, threads>>>(d_out, 64, 1.0);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    probe<MODE><<<sms * blocks_per_sm, threads>>>(d_out, iters, 1.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  double n = (double)sms * blocks_per_sm * threads * iters * INNER * CHAINS;
  return n / (best * 1e-3);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  double *d_out;
  cudaMalloc(&d_out, 64);
  long long *d_cyc;
  cudaMalloc(&d_cyc, 8);
  printf("device %s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
  const double theo = sms * 64.0 * p.clockRate * 1e3;
  printf("theoretical DFMA/s at max clock: %.3e\n", theo);
  for (int wps = 1; wps <= 8; wps *= 2) {  // warps per scheduler
    int threads = 128, blocks = wps;       // 4 warps per block -> 1 warp per SMSP per block
    printf("warps/SMSP %d:  mode0(1 reg) %.3e  mode1(3 regs) %.3e  mode4(3 regs, mixed) %.3e  mode2(2 regs) %.3e  mode3(DMUL 2 regs) %.3e\n",
           wps, run<0>(blocks, threads, sms, d_out), run<1>(blocks, threads, sms, d_out), run<4>(blocks, threads, sms, d_out),
           run<2>(blocks, threads, sms, d_out), run<3>(blocks, threads, sms, d_out));
  }
  latency<<<1, 32>>>(d_out, 256, 1.0, d_cyc);
  long long cyc;
  cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
  printf("dependent DFMA latency: %.2f cycles\n", (double)cyc / (256.0 * 64));
  return 0;
}
