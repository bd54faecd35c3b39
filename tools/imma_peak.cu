// imma_peak.cu -- legacy mma.sync INT8 (m16n8k32, s32 accumulate) issue rate on sm_100a: the budget for an exact
// fixed-point (Ozaki-style) version of the masked complement sums.  nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void imma(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int CHAINS>
__global__ void __launch_bounds__(256) probe(int iters, int* out, uint32_t seed) {
  int c[CHAINS][4];
  uint32_t a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = seed * (threadIdx.x + i + 1);
  for (int i = 0; i < 2; ++i) b[i] = seed ^ (threadIdx.x * 2654435761u + i);
  for (int k = 0; k < CHAINS; ++k)
    for (int i = 0; i < 4; ++i) c[k][i] = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) imma(c[k], a, b);
  }
  int s = 0;
  for (int k = 0; k < CHAINS; ++k)
    for (int i = 0; i < 4; ++i) s += c[k][i];
  if (s == 0x7fffffff) out[0] = s;
}

// the register pattern of masked_imma.cu's inner loop: 2 A fragments x 9 B fragments, 18 accumulators, no memory traffic
template <bool MT_OUTER>
__global__ void __launch_bounds__(256, 2) probe_tile(int iters, int* out, const uint32_t* __restrict__ src) {
  int c[2][9][4];
  uint32_t a[2][4], b[9][2];
  for (int m = 0; m < 2; ++m)
    for (int i = 0; i < 4; ++i) a[m][i] = src[threadIdx.x + 256 * (m * 4 + i)];
  for (int n = 0; n < 9; ++n)
    for (int i = 0; i < 2; ++i) b[n][i] = src[threadIdx.x + 256 * (8 + n * 2 + i)];
  for (int m = 0; m < 2; ++m)
    for (int n = 0; n < 9; ++n)
      for (int i = 0; i < 4; ++i) c[m][n][i] = 0;
  for (int it = 0; it < iters; ++it) {
    if (MT_OUTER) {
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < 9; ++n) imma(c[m][n], a[m], b[n]);
    } else {
#pragma unroll
      for (int n = 0; n < 9; ++n)
#pragma unroll
        for (int m = 0; m < 2; ++m) imma(c[m][n], a[m], b[n]);
    }
  }
  int s = 0;
  for (int m = 0; m < 2; ++m)
    for (int n = 0; n < 9; ++n)
      for (int i = 0; i < 4; ++i) s += c[m][n][i];
  if (s == 0x7fffffff) out[0] = s;
}

template <bool MT_OUTER>
void run_tile(int sms) {
  int* out;
  uint32_t* src;
  cudaMalloc(&out, 4);
  cudaMalloc(&src, 256 * 32 * 4);
  cudaMemset(src, 1, 256 * 32 * 4);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  probe_tile<MT_OUTER><<<sms * 2, 256>>>(100, out, src);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  probe_tile<MT_OUTER><<<sms * 2, 256>>>(iters, out, src);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)sms * 2 * 8 * iters * 18 * 16.0 * 8 * 32 * 2;
  printf("IMMA tile 2x9 (%s outer), 2 CTAs/SM: %.3f ms  %.1f TOPS  (%s)\n", MT_OUTER ? "m" : "n", ms, ops / ms * 1e-9,
         cudaGetErrorString(cudaGetLastError()));
}

template <int CHAINS>
void run(int ctas_per_sm, int sms) {
  int* out;
  cudaMalloc(&out, 4);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  probe<CHAINS><<<sms * ctas_per_sm, 256>>>(100, out, 3);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  probe<CHAINS><<<sms * ctas_per_sm, 256>>>(iters, out, 3);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)sms * ctas_per_sm * 8 * iters * CHAINS * 16.0 * 8 * 32 * 2;
  printf("IMMA m16n8k32 chains=%d ctas/sm=%d: %.3f ms  %.1f TOPS  (%s)\n", CHAINS, ctas_per_sm, ms, ops / ms * 1e-9,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  run<4>(1, p.multiProcessorCount);
  run<8>(1, p.multiProcessorCount);
  run<16>(1, p.multiProcessorCount);
  run<16>(2, p.multiProcessorCount);
  run_tile<false>(p.multiProcessorCount);
  run_tile<true>(p.multiProcessorCount);
  return 0;
}
