// bulk_load_probe.cu -- latency / throughput of cp.async.bulk (1-D, global -> shared) from an L2-resident region that every
// SM reads (the B-operand stream of masked_gram_tc.cu).   nvcc -arch=sm_100a -O3 -o bulk_load_probe bulk_load_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// depth loads in flight, each `bytes`, split into `pieces` bulk instructions
// mode 0: idle, 1: 8 warps of 64-bit shuffles, 2: 8 warps of shared-memory stores + loads, 3: 8 warps of DFMA
__device__ volatile int g_stop;
__global__ void probe(const uint8_t* src, int region_chunks, int bytes, int depth, int pieces, int iters, long long* out, int mode = 0, double* sink = nullptr) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) unsigned long long bars[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __shared__ double scratch[8 * 32 * 9];
  __shared__ volatile int stop;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  if (threadIdx.x >= 32) {
    double v = threadIdx.x, acc = 0.0;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    while (!stop) {
#pragma unroll 16
      for (int k = 0; k < 64; ++k) {
        if (mode == 1) v += __shfl_sync(0xffffffffu, v, (l + k) & 31, 4);
        if (mode == 2) { scratch[(w - 1) * 288 + l * 9 + (k & 7)] = v; acc += scratch[(w - 1) * 288 + ((l + 1) & 31) * 9 + (k & 7)]; }
        if (mode == 3) acc = fma(acc, 1.0000001, v);
      }
    }
    if (sink) sink[blockIdx.x * blockDim.x + threadIdx.x] = v + acc;
    return;
  }
  if (threadIdx.x == 0) {
    const int pb = bytes / pieces;
    auto load = [&](int i) {
      const int b = i % depth;
      mbar_expect_tx(smem_u32(&bars[b]), bytes);
      for (int p = 0; p < pieces; ++p)
        bulk_load_1d(smem_u32(sm) + b * bytes + p * pb, src + (size_t)(i % region_chunks) * bytes + p * pb, pb, smem_u32(&bars[b]));
    };
    for (int i = 0; i < depth && i < iters; ++i) load(i);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      mbar_wait(smem_u32(&bars[i % depth]), (i / depth) & 1);
      if (i + depth < iters) load(i + depth);
    }
    out[blockIdx.x] = clock64() - t0;
    stop = 1;
  }
}
int main() {
  const int bytes = 28672, region_chunks = 9, iters = 900;
  uint8_t* src; cudaMalloc(&src, (size_t)bytes * region_chunks); cudaMemset(src, 1, (size_t)bytes * region_chunks);
  long long* out; cudaMalloc(&out, 148 * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  double* sink; cudaMalloc(&sink, 148 * 288 * 8);
  for (int mode : {0, 1, 2, 3})
    for (int depth : {1, 2}) {
      const int grid = 148, pieces = 1;
      probe<<<grid, 288, depth * bytes>>>(src, region_chunks, bytes, depth, pieces, iters, out, mode, sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
      printf("mode %d depth %d: %8.0f cycles per 28 KB load (%5.1f B/clk/SM) %s\n", mode, depth, avg / iters, bytes / (avg / iters), cudaGetErrorString(e));
    }
  return 0;
}
