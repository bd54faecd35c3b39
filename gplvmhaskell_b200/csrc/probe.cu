// probe.cu -- FP64 tensor-pipe peak probe: register-resident mma.sync.m8n8k4.f64 chains (SASS DMMA.8x8x4),
// 16 warps per SM, 8 independent accumulator chains per warp.  bench.py uses the result as the "measured"
// FP64 denominator of the roofline because MEASURED_PEAKS.json carries no FP64 entry.
#include "common.cuh"

namespace gplvm {
namespace {
__global__ void __launch_bounds__(512) dmma_probe_kernel(double* out, int iters, double seed) {
  double c[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) c[j][0] = c[j][1] = 0.0;
  const double a = seed * threadIdx.x, b = seed + threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
                   : "+d"(c[j][0]), "+d"(c[j][1])
                   : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace
}  // namespace gplvm

using namespace gplvm;

extern "C" int gplvm_probe_dmma_tflops(double* tflops_out) {
  if (!tflops_out) return fail(GPLVM_ERR_ARG, "gplvm_probe_dmma_tflops: null argument");
  GP_TRY(ensure_context());
  Context& c = ctx();
  std::lock_guard<std::recursive_mutex> lk(c.mu);
  CUDA_TRY(cudaSetDevice(c.shards[0].device));
  cudaStream_t st = c.shards[0].stream;
  int sms = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c.shards[0].device));
  double* out = nullptr;
  CUDA_TRY(cudaMalloc(&out, sizeof(double) * (size_t)sms * 512));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  const int iters = 4000;
  float best = 1e30f;
  for (int r = 0; r < 6; ++r) {
    CUDA_TRY(cudaEventRecord(e0, st));
    GP_LAUNCH(dmma_probe_kernel, sms, 512, 0, st, out, iters, 1e-9);
    CUDA_TRY(cudaEventRecord(e1, st));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (r > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  CUDA_TRY(cudaGetLastError());
  const double flop = 512.0 * 8.0 * iters * 16.0 * sms;  // 2*8*8*4 flop per DMMA.8x8x4
  *tflops_out = flop / (best * 1e-3) * 1e-12;
  return GPLVM_OK;
}
