// dense_i8_proto.cu -- feasibility prototype for DESIGN.md section 7 ("FP64-faithful dense statistics on INT8 tcgen05"):
// step 1 of the dense EM statistics, T (128 x 16) = X (128 x 256) . A (256 x 16), with BOTH factors cut into 7 signed
// byte planes (balanced digits, Ozaki splitting).  X plane p (weight 256^p) meets the A planes q >= 6 - p in ONE
// tcgen05.mma kind::i8 of N = 16 (p + 1): the B descriptor starts at row 16 (6 - p), the accumulator at TMEM column 0,
// so products of equal weight p + q accumulate in the same s32 columns (7 MMAs x 8 k32 steps per 128 observations).
// Part 1 checks the result against FP64 (long double) on the host; part 2 times the 56-MMA sequence on every SM with
// the operands resident in shared memory (is an N = 16 .. 112 MMA as cheap as N / 256 of an N = 256 one?).
// nvcc -gencode arch=compute_100a,code=sm_100a
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, K = 256, L = 16, NP = 7;   // observations per tile, features, latent dimensions, planes
constexpr int NB = L * NP;                        // 112 B rows: row 16 q + a = plane q of latent column a

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
  return (uint64_t)((addr & 0x3ffff) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// K-major SWIZZLE_128B tile with 128 B of K per row; K = 256 is two such tiles (k-blocks) `kb_stride` bytes apart
__host__ __device__ inline uint32_t sw128(int r, int b) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((b >> 4) ^ (r & 7))) << 4) + (b & 15)); }

__device__ __forceinline__ bool wait_bar(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (long long spin = 0; spin < (1ll << 24) && !done; ++spin)
    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.b32 %0, 1, 0, q;\n\t}\n"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
  return done != 0;
}

// Xp: [NP][M][K] int8 planes, Ap: [NP][L][K] int8 planes (K contiguous).  reps == 0: accuracy mode (the A buffer is
// refilled with plane p before its MMAs); reps > 0: timing mode (the 56-MMA sequence reps times on resident operands).
// Planes run p = 6 .. 0: the first MMA (p = 6, N = 112) overwrites all 112 accumulator columns, everything else adds.
__global__ void __launch_bounds__(128, 1) proto(const int8_t* __restrict__ Xp, const int8_t* __restrict__ Ap, int* __restrict__ Dg, int reps,
                                                unsigned long long* cycles, int* err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t raw_u = smem_u32(smem_raw), base_u = (raw_u + 1023u) & ~1023u;
  uint8_t* base = smem_raw + (base_u - raw_u);
  uint8_t* Xs = base;                       // one plane: 2 k-blocks x (128 rows x 128 B) = 32 KB
  uint8_t* Bs = base + 2 * 16384;           // 2 k-blocks x (112 rows x 128 B)
  constexpr int BKB = NB * 128;             // bytes of one B k-block (14 336 = 14 x 1024)
  auto fill_plane = [&](int p) {
    for (int i = threadIdx.x; i < M * K; i += 128) {
      const int r = i / K, k = i % K;
      Xs[(k >> 7) * 16384 + sw128(r, k & 127)] = (uint8_t)Xp[(size_t)p * M * K + i];
    }
  };
  fill_plane(NP - 1);
  for (int i = threadIdx.x; i < NB * K; i += 128) {
    const int r = i / K, k = i % K;        // r = 16 q + a
    Bs[(k >> 7) * BKB + sw128(r, k & 127)] = (uint8_t)Ap[i];
  }
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
  auto issue_plane = [&](int p) {   // one thread: 8 MMAs (K = 256) of X plane p against the A planes q = 6 - p .. 6
    const int n = L * (p + 1);
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
#pragma unroll
    for (int ks = 0; ks < K / 32; ++ks) {
      const uint64_t adesc = smem_desc_sw128(smem_u32(Xs) + (ks >> 2) * 16384) + 2 * (ks & 3);
      const uint64_t bdesc = smem_desc_sw128(smem_u32(Bs) + (ks >> 2) * BKB + (uint32_t)(L * (NP - 1 - p)) * 128) + 2 * (ks & 3);
      const uint32_t acc = (p == NP - 1 && ks == 0) ? 0u : 1u;
      asm volatile(
          "{\n\t.reg .pred pp;\n\tsetp.ne.b32 pp, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, pp;\n\t}\n" ::"r"(tmem),
          "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
          : "memory");
    }
  };
  bool ok = true;
  uint32_t phase = 0;
  if (reps == 0) {
    for (int p = NP - 1; p >= 0; --p) {
      if (p != NP - 1) {
        fill_plane(p);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
      }
      if (threadIdx.x == 0) {
        issue_plane(p);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_u) : "memory");
      }
      ok = wait_bar(bar_u, phase) && ok;   // every thread waits: the A buffer is free again
      phase ^= 1u;
      __syncthreads();
    }
    if (!ok && threadIdx.x == 0) *err = 1;
  } else if (threadIdx.x == 0) {
    const unsigned long long t0 = clock64();
    for (int rep = 0; rep < reps && ok; ++rep) {
#pragma unroll 1
      for (int p = NP - 1; p >= 0; --p) issue_plane(p);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_u) : "memory");
      ok = wait_bar(bar_u, phase);
      phase ^= 1u;
    }
    cycles[blockIdx.x] = clock64() - t0;
    if (!ok) *err = 1;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (blockIdx.x == 0 && reps == 0) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < NB; c0 += 16) {
      uint32_t r[16];
      const uint32_t taddr = tmem + ((uint32_t)(32 * w) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
            "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j) Dg[(32 * w + lane) * NB + c0 + j] = (int)r[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

// balanced base-256 digits of a 56-bit signed integer
static void cut(long long x, int8_t* d) {
  for (int s = 0; s < NP; ++s) {
    const int8_t lo = (int8_t)(x & 0xff);
    d[s] = lo;
    x = (x - lo) >> 8;
  }
}

int main() {
  srand(11);
  std::vector<double> X(M * K), A(K * L);
  for (auto& v : X) v = (rand() / (double)RAND_MAX * 2 - 1) * ((rand() % 7) ? 1.0 : 1e-3);
  for (auto& v : A) v = (rand() / (double)RAND_MAX * 2 - 1);
  // feature scales folded into A: X' = X / c_f, A' = c_f A; column scales of A'
  std::vector<double> c(K, 0.0), sa(L, 0.0);
  for (int i = 0; i < M; ++i)
    for (int f = 0; f < K; ++f) c[f] = fmax(c[f], fabs(X[i * K + f]));
  for (int f = 0; f < K; ++f)
    for (int a = 0; a < L; ++a) sa[a] = fmax(sa[a], fabs(c[f] * A[f * L + a]));
  std::vector<int8_t> Xp((size_t)NP * M * K), Ap((size_t)NP * L * K);
  std::vector<long long> Xi(M * K), Ai(K * L);
  const double S = ldexp(1.0, 54);
  for (int i = 0; i < M; ++i)
    for (int f = 0; f < K; ++f) {
      Xi[i * K + f] = llrint(X[i * K + f] / c[f] * S);
      int8_t d[NP];
      cut(Xi[i * K + f], d);
      for (int p = 0; p < NP; ++p) Xp[((size_t)p * M + i) * K + f] = d[p];
    }
  for (int f = 0; f < K; ++f)
    for (int a = 0; a < L; ++a) {
      Ai[f * L + a] = llrint(c[f] * A[f * L + a] / sa[a] * S);
      int8_t d[NP];
      cut(Ai[f * L + a], d);
      for (int q = 0; q < NP; ++q) Ap[((size_t)q * L + a) * K + f] = d[q];
    }
  int8_t *dX, *dA;
  int *dD, *derr;
  unsigned long long* dcyc;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  cudaMalloc(&dX, Xp.size());
  cudaMalloc(&dA, Ap.size());
  cudaMalloc(&dD, sizeof(int) * M * NB);
  cudaMalloc(&derr, 4);
  cudaMalloc(&dcyc, 8 * sms);
  cudaMemset(derr, 0, 4);
  cudaMemcpy(dX, Xp.data(), Xp.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dA, Ap.data(), Ap.size(), cudaMemcpyHostToDevice);
  const int smem = 32768 + 2 * NB * 128 + 1024;
  if (cudaFuncSetAttribute(proto, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
    printf("shared memory request of %d B refused\n", smem);
    return 1;
  }
  proto<<<1, 128, smem>>>(dX, dA, dD, 0, dcyc, derr);
  cudaError_t rc = cudaDeviceSynchronize();
  std::vector<int> D(M * NB);
  int herr = 0;
  cudaMemcpy(D.data(), dD, sizeof(int) * M * NB, cudaMemcpyDeviceToHost);
  cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost);
  // host: exact plane-pair sums for the kept pairs, and the FP64 comparison
  long bad = 0;
  double max_rel = 0.0, max_rel_full = 0.0;
  for (int i = 0; i < M; ++i)
    for (int a = 0; a < L; ++a) {
      __int128 T = 0;
      for (int w = 0; w < NP; ++w) {
        long long ref = 0;
        for (int p = 0; p < NP; ++p) {
          const int q = w + 6 - p;
          if (q < 0 || q >= NP) continue;
          for (int f = 0; f < K; ++f) ref += (long long)Xp[((size_t)p * M + i) * K + f] * Ap[((size_t)q * L + a) * K + f];
        }
        if (ref != D[i * NB + 16 * w + a]) ++bad;
        T += (__int128)D[i * NB + 16 * w + a] << (8 * w);
      }
      const long double t_fx = (long double)T * ldexpl(1.0L, 48 - 108) * sa[a];   // weight 256^(w + 6), operands scaled by 2^54 each
      long double t_ref = 0.0L, t_abs = 0.0L;
      for (int f = 0; f < K; ++f) {
        t_ref += (long double)X[i * K + f] * A[f * L + a];
        t_abs += fabsl((long double)X[i * K + f] * A[f * L + a]);
      }
      max_rel = fmax(max_rel, (double)(fabsl(t_fx - t_ref) / t_abs));          // relative to sum |x a| (dot-product error measure)
      max_rel_full = fmax(max_rel_full, (double)(fabsl(t_fx - t_ref) / fabsl(t_ref)));
    }
  printf("part 1 (%s, timeout flag %d): %ld plane-pair sums differ from the host integers; max |T - T_fp64| / sum|x a| = %.3e, / |T| = %.3e\n",
         cudaGetErrorString(rc), herr, bad, max_rel, max_rel_full);
  // part 2: issue rate of the 56-MMA tile sequence on every SM
  const int reps = 2000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  proto<<<sms, 128, smem>>>(dX, dA, dD, reps, dcyc, derr);
  cudaEventRecord(e1);
  rc = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<unsigned long long> cyc(sms);
  cudaMemcpy(cyc.data(), dcyc, 8 * sms, cudaMemcpyDeviceToHost);
  cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost);
  const double ops = (double)sms * reps * 2.0 * M * K * 16.0 * 28;   // 28 plane pairs of 128 x 256 x 16
  printf("part 2 (%s, timeout flag %d): %d tiles per SM in %.3f ms (includes the operand fill); %.0f cycles per 128-observation tile "
         "(commit + wait after every tile); %.1f TOPS on the 28 kept pairs\n",
         cudaGetErrorString(rc), herr, reps, ms, (double)cyc[0] / reps, ops / ms * 1e-9);
  return (rc == cudaSuccess && !herr && bad == 0) ? 0 : 1;
}
