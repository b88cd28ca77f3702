// tools/int_peak.cu -- calibration of the integer-pipe denominators of the ALU roofline (SURVEY.md 8d asks for a
// measured INT32 peak at the sustained clock, not an assumed one).  Three dependent-chain-free loops per thread,
// 8 independent accumulators each, on 148 x 8 blocks of 256 threads:
//   max  : VIMNMX      (integer max, the instruction the gap/no-gap recurrences are made of)
//   add  : IADD3       (the adds of open / extend / score; the compiler may move some to the FMA pipe as IMAD)
//   mix  : 1 max + 1 add + 1 compare/select per step, the per-cell mix of the fill
// Prints giga-operations per second (thread-level ops) and ops per clock per SM at the clock nvidia-smi reports.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define ITER 4096
template <int MODE> __global__ void __launch_bounds__(256) k(int *out, int a, int b) {
  int x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 8
  for (int i = 0; i < ITER; i++) {
    if (MODE == 0) {
      x0 = max(x0, a) ^ b; x1 = max(x1, a) ^ b; x2 = max(x2, a) ^ b; x3 = max(x3, a) ^ b;
      x4 = max(x4, a) ^ b; x5 = max(x5, a) ^ b; x6 = max(x6, a) ^ b; x7 = max(x7, a) ^ b;
    } else if (MODE == 1) {
      x0 = (x0 + a) ^ b; x1 = (x1 + a) ^ b; x2 = (x2 + a) ^ b; x3 = (x3 + a) ^ b;
      x4 = (x4 + a) ^ b; x5 = (x5 + a) ^ b; x6 = (x6 + a) ^ b; x7 = (x7 + a) ^ b;
    } else {
      x0 = max(x0 + a, x1); x1 = x1 > x2 ? x1 + b : x2; x2 = max(x2 + a, x3); x3 = x3 > x4 ? x3 + b : x4;
      x4 = max(x4 + a, x5); x5 = x5 > x6 ? x5 + b : x6; x6 = max(x6 + a, x7); x7 = x7 > x0 ? x7 + b : x0;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
}

template <int MODE> static double run(int *d, int grid, double ops_per_iter) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 3; w++) k<MODE><<<grid, 256>>>(d, 3, 5);
  cudaEventRecord(e0);
  const int reps = 20;
  for (int r = 0; r < reps; r++) k<MODE><<<grid, 256>>>(d, 3, 5);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return (double)grid * 256 * ITER * ops_per_iter * reps / (ms * 1e-3) / 1e9;
}

int main() {
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { fprintf(stderr, "no device\n"); return 1; }
  const int grid = p.multiProcessorCount * 8;
  int *d;
  cudaMalloc(&d, (size_t)grid * 256 * sizeof(int));
  const double gmax = run<0>(d, grid, 16), gadd = run<1>(d, grid, 16), gmix = run<2>(d, grid, 16);
  int mhz = 0;
  cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0);
  const double clk = mhz * 1e3;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz_attr\": %.0f, "
         "\"max_xor_gops\": %.1f, \"add_xor_gops\": %.1f, \"fill_mix_gops\": %.1f, "
         "\"max_xor_ops_per_clk_sm\": %.1f, \"add_xor_ops_per_clk_sm\": %.1f, \"fill_mix_ops_per_clk_sm\": %.1f}\n",
         p.name, p.multiProcessorCount, clk / 1e6, gmax, gadd, gmix,
         gmax * 1e9 / clk / p.multiProcessorCount, gadd * 1e9 / clk / p.multiProcessorCount, gmix * 1e9 / clk / p.multiProcessorCount);
  return 0;
}
