// tcgen05_i8_tmem_a_check.cu -- known-answer check of tcgen05.mma kind::i8 with the A operand in TENSOR MEMORY
// (written with tcgen05.st 32x32b: lane = row m, 32-bit column q holds the K elements 4 q .. 4 q + 3, byte i = element 4 q + i)
// and B in shared memory (canonical SWIZZLE_128B K-major).  D[128 x N] = A[128 x 128] . B[N x 128]^T against the host.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tcgen05_i8_tmem_a_check tools/tcgen05_i8_tmem_a_check.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 80, K = 128;
constexpr int A_COL = 96;   // TMEM columns [96, 128): the A operand (K / 4 = 32 columns)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
  return (uint64_t)((addr & 0x3ffff) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ inline uint32_t sw128(int r, int b) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((b >> 4) ^ (r & 7))) << 4) + (b & 15)); }

__global__ void __launch_bounds__(128, 1) check(const int8_t* __restrict__ Ag, const uint8_t* __restrict__ Bg, int* __restrict__ Dg, int* err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) unsigned long long bar;
  uint8_t* B = smem;
  for (int i = threadIdx.x; i < N * K; i += 128) B[sw128(i / K, i % K)] = Bg[i];
  const uint32_t bar_u = smem_u32(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // A row m = threadIdx.x -> 32 words -> TMEM lane m, columns A_COL ..
  {
    uint32_t r[32];
    const uint32_t* row = reinterpret_cast<const uint32_t*>(Ag + (size_t)threadIdx.x * K);
#pragma unroll
    for (int q = 0; q < 32; ++q) r[q] = row[q];
    const uint32_t taddr = tmem + ((uint32_t)(32 * w) << 16) + (uint32_t)A_COL;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
        "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t idesc = (2u << 4) | (1u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  const uint64_t bdesc = smem_desc_sw128(smem_u32(B));
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K / 32; ++k) {
      const uint32_t acc = k ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem),
          "r"(tmem + (uint32_t)(A_COL + 8 * k)), "l"(bdesc + 2 * k), "r"(idesc), "r"(acc), "r"(0u)
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
  for (auto& v : A) v = (int8_t)(rand() % 3 - 1);
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
  const int smem = N * 128 + 1024;
  cudaFuncSetAttribute(check, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  check<<<1, 128, smem>>>(dA, dB, dD, derr);
  cudaError_t rc = cudaDeviceSynchronize();
  std::vector<int> D(M * N);
  int herr = 0;
  cudaMemcpy(D.data(), dD, sizeof(int) * M * N, cudaMemcpyDeviceToHost);
  cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost);
  long bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      int ref = 0;
      for (int k = 0; k < K; ++k) ref += (int)A[m * K + k] * (int)B[n * K + k];
      if (ref != D[m * N + n]) {
        if (bad < 5) printf("mismatch D[%d][%d] = %d, expected %d\n", m, n, D[m * N + n], ref);
        ++bad;
      }
    }
  const bool pass = rc == cudaSuccess && !herr && bad == 0;
  printf("tcgen05 kind::i8 with A in tensor memory: %s, timeout flag %d, %ld mismatches of %d -> %s\n", cudaGetErrorString(rc), herr, bad,
         M * N, pass ? "PASS" : "FAIL");
  return pass ? 0 : 1;
}
