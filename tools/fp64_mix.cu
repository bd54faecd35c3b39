// Do DMMA (tensor sub-pipe) and DFMA (FP64 pipe) run concurrently on B200?  Half the warps of every SM issue DMMA
// chains, the other half DFMA chains; compare the combined rate with each alone.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void mix(double* out, int iters, int mode, double seed) {
  const int warp = threadIdx.x >> 5;
  const bool do_mma = (mode == 0) || (mode == 2 && (warp & 1) == 0);
  const bool do_fma = (mode == 1) || (mode == 2 && (warp & 1) == 1);
  double c[8][2]; double f[16];
  for (int j = 0; j < 8; ++j) c[j][0] = c[j][1] = 0; for (int j = 0; j < 16; ++j) f[j] = j;
  double a = seed * threadIdx.x, b = seed + threadIdx.x;
  if (do_mma) for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};" : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
  }
  if (do_fma) for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = fma(a, f[j], b);
  }
  double s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1]; for (int j = 0; j < 16; ++j) s += f[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* out; cudaMalloc(&out, 8 * 148 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, warps = 16;
  for (int mode = 0; mode < 3; ++mode) {
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); mix<<<148, warps * 32>>>(out, iters, mode, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    double mma_w = (mode == 0) ? warps : (mode == 2 ? warps / 2 : 0), fma_w = (mode == 1) ? warps : (mode == 2 ? warps / 2 : 0);
    double flop = (512.0 * 8 * mma_w + 64.0 * 16 * fma_w) * iters * 148;
    printf("mode %d (%s): %.3f ms  %.2f TFLOP/s total\n", mode, mode == 0 ? "DMMA only" : mode == 1 ? "DFMA only" : "half DMMA + half DFMA", best, flop / best * 1e-9);
  }
  return 0;
}
