// tcgen05_i8_peak.cu -- issue rate of tcgen05.mma kind::i8 (s8 x s8 -> s32 in TMEM, M = 128, N = 256, K = 32) on sm_100a:
// the budget of an exact fixed-point (Ozaki-style) FP64 product on the 5th-generation tensor cores.
// One CTA per SM; operands are two zero-filled SWIZZLE_128B K-major tiles in shared memory that every MMA re-reads, the
// accumulator is a 128 x 256 TMEM tile.  nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
  // start address >> 4 | LBO (unused for swizzled K-major: 1) << 16 | SBO = 1024 B >> 4 << 32 | version 1 << 46 | SWIZZLE_128B (2) << 61
  return (uint64_t)((addr & 0x3ffff) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__global__ void __launch_bounds__(128, 1) probe(int batches, unsigned long long* cycles, int* err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) unsigned long long bar;
  uint8_t* A = smem;                 // 128 rows x 128 B
  uint8_t* B = smem + 128 * 128;     // 256 rows x 128 B
  for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  const uint32_t bar_u = smem_u32(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  // instruction descriptor: D = S32 (2 << 4), A = B = signed 8 bit (1 << 7, 1 << 10), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
  const uint64_t adesc = smem_desc_sw128(smem_u32(A)), bdesc = smem_desc_sw128(smem_u32(B));
  unsigned long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {
    t0 = clock64();
    uint32_t phase = 0;
    for (int b = 0; b < batches; ++b) {
#pragma unroll 1
      for (int i = 0; i < 16; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {   // 4 x K32 along the 128-byte swizzle row: + 32 B = + 2 in descriptor units
          const uint32_t acc = (b | i | k) ? 1u : 0u;
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem),
              "l"(adesc + 2 * k), "l"(bdesc + 2 * k), "r"(idesc), "r"(acc), "r"(0u)
              : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_u) : "memory");
      uint32_t done = 0;
      for (long long spin = 0; spin < (1ll << 26) && !done; ++spin)
        asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.b32 %0, 1, 0, q;\n\t}\n"
                     : "=r"(done)
                     : "r"(bar_u), "r"(phase)
                     : "memory");
      if (!done) {
        *err = 1;
        break;
      }
      phase ^= 1u;
    }
    t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount, smem = (128 + 256) * 128 + 1024;
  unsigned long long* cyc;
  int* err;
  cudaMalloc(&cyc, sizeof(unsigned long long) * sms);
  cudaMalloc(&err, 4);
  cudaMemset(err, 0, 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  probe<<<sms, 128, smem>>>(10, cyc, err);
  cudaError_t rc = cudaDeviceSynchronize();
  printf("%s, %d SMs; warm-up: %s\n", p.name, sms, cudaGetErrorString(rc));
  if (rc != cudaSuccess) return 1;
  const int batches = 2000;
  cudaEventRecord(e0);
  probe<<<sms, 128, smem>>>(batches, cyc, err);
  cudaEventRecord(e1);
  rc = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  int herr = 0;
  cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost);
  const double ops = (double)sms * batches * 64.0 * 128 * 256 * 32 * 2;
  printf("tcgen05.mma kind::i8 M128 N256 K32, 1 CTA/SM: %.3f ms  %.1f TOPS  (%s, timeout flag %d)\n", ms, ops / ms * 1e-9,
         cudaGetErrorString(rc), herr);
  return 0;
}
