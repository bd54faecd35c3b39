// tcgen05_i8_check.cu -- known-answer check of the building blocks an INT8 tcgen05 kernel needs: operands written by
// ordinary threads in the canonical SWIZZLE_128B K-major layout, tcgen05.mma kind::i8 (s8 x u8 -> s32), K advanced
// inside the swizzle row through the descriptor start address, accumulator read back with tcgen05.ld 32x32b.
// D[128 x N] = A[128 x 128] . B[N x 128]^T against a host reference.  nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 80, K = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
  return (uint64_t)((addr & 0x3ffff) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// byte (row r, k-byte b) of a K-major tile with 128-byte rows, 8-row groups of 1024 B, 16-byte units XOR-swizzled by r & 7
__host__ __device__ inline uint32_t sw128(int r, int b) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((b >> 4) ^ (r & 7))) << 4) + (b & 15)); }

// a_mn = 1: the SAME A tile is read MN-major (instruction-descriptor bit 15): D[m][n] = sum_k A[k][m] B[n][k], i.e. the
// tile is used transposed -- what X^T . Ez needs.  K then advances by 32 ROWS (4096 B) per MMA.
__global__ void __launch_bounds__(128, 1) check(const int8_t* __restrict__ Ag, const uint8_t* __restrict__ Bg, int* __restrict__ Dg, int* err, int a_mn) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) unsigned long long bar;
  uint8_t* A = smem;
  uint8_t* B = smem + M * 128;
  for (int i = threadIdx.x; i < M * K; i += 128) A[sw128(i / K, i % K)] = (uint8_t)Ag[i];
  for (int i = threadIdx.x; i < N * K; i += 128) B[sw128(i / K, i % K)] = Bg[i];
  const uint32_t bar_u = smem_u32(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  // D = S32, A signed 8 bit, B unsigned 8 bit, both K-major
  const uint32_t idesc = (2u << 4) | (1u << 7) | (0u << 10) | ((uint32_t)(a_mn ? 1u : 0u) << 15) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  const uint64_t adesc = smem_desc_sw128(smem_u32(A)), bdesc = smem_desc_sw128(smem_u32(B));
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K / 32; ++k) {
      const uint32_t acc = k ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem),
          "l"(adesc + (a_mn ? 256 * k : 2 * k)), "l"(bdesc + 2 * k), "r"(idesc), "r"(acc), "r"(0u)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_u) : "memory");
  }
  uint32_t done = 0;
  for (long long spin = 0; spin < (1ll << 24) && !done; ++spin)
    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.b32 %0, 1, 0, q;\n\t}\n"
                 : "=r"(done)
                 : "r"(bar_u), "r"(0u)
                 : "memory");
  if (!done) *err = 1;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (done) {
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16];
      const uint32_t taddr = tmem + ((uint32_t)(32 * w) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
            "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j) Dg[(32 * w + lane) * N + c0 + j] = (int)r[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

int main() {
  std::vector<int8_t> A(M * K);
  std::vector<uint8_t> B(N * K);
  srand(7);
  for (auto& v : A) v = (int8_t)(rand() % 3 - 1);      // -1, 0, 1 (the product kernel uses 0 / 1)
  for (auto& v : B) v = (uint8_t)(rand() & 255);
  int8_t* dA;
  uint8_t* dB;
  int *dD, *derr;
  cudaMalloc(&dA, A.size());
  cudaMalloc(&dB, B.size());
  cudaMalloc(&dD, sizeof(int) * M * N);
  cudaMalloc(&derr, 4);
  cudaMemset(derr, 0, 4);
  cudaMemset(dD, 0xff, sizeof(int) * M * N);
  cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
  const int smem = (M + N) * 128 + 1024;
  cudaFuncSetAttribute(check, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int fails = 0;
  for (int a_mn = 0; a_mn < 2; ++a_mn) {
    cudaMemset(dD, 0xff, sizeof(int) * M * N);
    check<<<1, 128, smem>>>(dA, dB, dD, derr, a_mn);
    cudaError_t rc = cudaDeviceSynchronize();
    std::vector<int> D(M * N);
    int herr = 0;
    cudaMemcpy(D.data(), dD, sizeof(int) * M * N, cudaMemcpyDeviceToHost);
    cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        int ref = 0;
        for (int k = 0; k < K; ++k) ref += (int)(a_mn ? A[k * K + m] : A[m * K + k]) * (int)B[n * K + k];
        if (ref != D[m * N + n]) {
          if (bad < 3) printf("mismatch D[%d][%d] = %d, expected %d\n", m, n, D[m * N + n], ref);
          ++bad;
        }
      }
    const bool pass = rc == cudaSuccess && !herr && bad == 0;
    printf("tcgen05 kind::i8 s8 x u8 known-answer check, A %s: %s, timeout flag %d, %ld mismatches of %d -> %s\n",
           a_mn ? "MN-major (tile used transposed)" : "K-major", cudaGetErrorString(rc), herr, bad, M * N, pass ? "PASS" : "FAIL");
    fails += pass ? 0 : 1;
  }
  return fails;
}
