// dense_i8_proto2.cu -- second feasibility prototype for DESIGN.md section 7: the COMPLETE per-tile data flow of the
// dense EM statistics on tcgen05 kind::i8 for one tile of 128 observations (D = 256 features, M = 16 latent dims):
//   step 1  T = X' . A'               X planes (K-major A operand) x projector planes          -> TMEM columns   0..111
//   hand-off: each thread (= one observation = one TMEM lane) reads its 112 sums, assembles the 128-bit integers,
//            rounds Ez to FP64 once, re-cuts it into 8 balanced byte planes and writes them as the K-major operand
//            tile [(plane q, latent a)][observation] in shared memory
//   step 2  XEz = X'^T . Ez           the SAME X tiles read MN-major (transposed), 2 feature blocks -> columns 128..351
//   EE      Ez^T Ez                   the Ez plane tile against itself (all plane pairs)         -> columns 352..479
// Every integer the device produces is checked against the host; the assembled FP64 statistics are compared with a
// long-double reference.  (X planes are refilled into one 32 KB buffer per plane: the streaming pipeline is not the
// subject here.)   nvcc -gencode arch=compute_100a,code=sm_100a
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, K = 256, L = 16, NPX = 7, NPE = 8;
constexpr int NB1 = L * NPX;    // 112 projector-plane rows: 16 q + a
constexpr int NB2 = L * NPE;    // 128 Ez-plane rows: 16 q + a
constexpr int C1 = 0, C2 = 128, C3 = 352;   // TMEM column bases

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
  return (uint64_t)((addr & 0x3ffff) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
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
__device__ __forceinline__ void mma_i8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred pp;\n\tsetp.ne.b32 pp, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, pp;\n\t}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// exact integer (|T| < 2^77) -> FP64 with one rounding
__host__ __device__ inline double i128_to_double(__int128 T) {
  const long long thi = (long long)(T >> 24), tlo = (long long)(T & 0xffffff);
  return fma((double)thi, 16777216.0, (double)tlo);
}

// s8 x s8 instruction descriptor: D = S32, N, M = 128, optional MN-major A
__device__ __forceinline__ uint32_t idesc_s8(int n, int a_mn) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(128, 1) proto2(const int8_t* __restrict__ Xp, const int8_t* __restrict__ Ap, const double* __restrict__ sa,
                                                 const double* __restrict__ se, int* __restrict__ D1g, double* __restrict__ Ezg,
                                                 long long* __restrict__ Eig, int* __restrict__ D2g, int* __restrict__ D3g, int* err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t raw_u = smem_u32(smem_raw), base_u = (raw_u + 1023u) & ~1023u;
  uint8_t* base = smem_raw + (base_u - raw_u);
  uint8_t* Xs = base;                         // one X plane: 2 feature blocks x (128 observation rows x 128 B)
  uint8_t* B1 = base + 32768;                 // projector planes: 2 k-blocks x (112 rows x 128 B)
  uint8_t* B2 = base + 32768 + 2 * NB1 * 128; // Ez planes: 128 rows x 128 B (K = observations)
  constexpr int B1KB = NB1 * 128;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  auto fill_plane = [&](int p) {
    for (int i = tid; i < M * K; i += 128) {
      const int r = i / K, k = i % K;
      Xs[(k >> 7) * 16384 + sw128(r, k & 127)] = (uint8_t)Xp[(size_t)p * M * K + i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  };
  for (int i = tid; i < NB1 * K; i += 128) {
    const int r = i / K, k = i % K;
    B1[(k >> 7) * B1KB + sw128(r, k & 127)] = (uint8_t)Ap[i];
  }
  const uint32_t bar_u = smem_u32(&bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  uint32_t phase = 0;
  bool ok = true;
  auto commit_wait = [&]() {   // thread 0 commits; everybody waits (the operand buffers are free again afterwards)
    if (tid == 0) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_u) : "memory");
    ok = wait_bar(bar_u, phase) && ok;
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();
  };

  // ---- step 1: planes p = 6 .. 0 against projector planes q >= 6 - p, class p + q - 6 -> columns 16 (p + q - 6) + a --------
  for (int p = NPX - 1; p >= 0; --p) {
    fill_plane(p);
    if (tid == 0) {
      const uint32_t idesc = idesc_s8(L * (p + 1), 0);
#pragma unroll
      for (int ks = 0; ks < K / 32; ++ks)
        mma_i8(tmem + C1, smem_desc_sw128(smem_u32(Xs) + (ks >> 2) * 16384) + 2 * (ks & 3),
               smem_desc_sw128(smem_u32(B1) + (ks >> 2) * B1KB + (uint32_t)(L * (NPX - 1 - p)) * 128) + 2 * (ks & 3), idesc,
               (p == NPX - 1 && ks == 0) ? 0u : 1u);
    }
    commit_wait();
  }
  // ---- hand-off: thread = observation i = TMEM lane ------------------------------------------------------------------
  {
    const int i = 32 * w + lane;
    __int128 T[L];
#pragma unroll
    for (int a = 0; a < L; ++a) T[a] = 0;
#pragma unroll
    for (int cls = 0; cls < NPX; ++cls) {
      uint32_t r[16];
      tmem_ld16(tmem + ((uint32_t)(32 * w) << 16) + (uint32_t)(C1 + 16 * cls), r);
#pragma unroll
      for (int a = 0; a < L; ++a) {
        D1g[i * NB1 + 16 * cls + a] = (int)r[a];
        T[a] += (__int128)(int)r[a] << (8 * cls);
      }
    }
#pragma unroll
    for (int a = 0; a < L; ++a) {
      // operands carry 2^54 each, the kept classes start at weight 256^6: 2^(48 - 108)
      const double ez = (i128_to_double(T[a]) * 0x1p-60) * sa[a];
      Ezg[i * L + a] = ez;
      long long ei = __double2ll_rn((ez / se[a]) * 4611686018427387904.0);   // 2^62
      Eig[i * L + a] = ei;
#pragma unroll
      for (int q = 0; q < NPE; ++q) {   // balanced base-256 digits
        const long long lo = (long long)(int8_t)(ei & 0xff);
        B2[sw128(16 * q + a, i)] = (uint8_t)lo;
        ei = (ei - lo) >> 8;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // ---- step 2: X^T (MN-major use of the same tiles; K = 128 observations = 4 x 32 rows) against Ez planes q >= 7 - p ------------
  for (int p = NPX - 1; p >= 0; --p) {
    fill_plane(p);
    if (tid == 0) {
      const uint32_t idesc = idesc_s8(L * (p + 1), 1);
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int ks = 0; ks < M / 32; ++ks)
          mma_i8(tmem + C2 + kb * NB1, smem_desc_sw128(smem_u32(Xs) + kb * 16384) + 256 * ks,
                 smem_desc_sw128(smem_u32(B2) + (uint32_t)(L * (NPE - 1 - p)) * 128) + 2 * ks, idesc, (p == NPX - 1 && ks == 0) ? 0u : 1u);
    }
    commit_wait();
  }
  // ---- Ez^T Ez: the Ez plane tile against itself ---------------------------------------------------------------------------
  if (tid == 0) {
    const uint32_t idesc = idesc_s8(NB2, 0);
#pragma unroll
    for (int ks = 0; ks < M / 32; ++ks)
      mma_i8(tmem + C3, smem_desc_sw128(smem_u32(B2)) + 2 * ks, smem_desc_sw128(smem_u32(B2)) + 2 * ks, idesc, ks ? 1u : 0u);
  }
  commit_wait();
  // ---- read-out ---------------------------------------------------------------------------------------------------------------
  {
    const int row = 32 * w + lane;
    for (int kb = 0; kb < 2; ++kb)
      for (int c0 = 0; c0 < NB1; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem + ((uint32_t)(32 * w) << 16) + (uint32_t)(C2 + kb * NB1 + c0), r);
        for (int j = 0; j < 16; ++j) D2g[((kb * 128 + row) * NB1) + c0 + j] = (int)r[j];
      }
    for (int c0 = 0; c0 < NB2; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem + ((uint32_t)(32 * w) << 16) + (uint32_t)(C3 + c0), r);
      for (int j = 0; j < 16; ++j) D3g[row * NB2 + c0 + j] = (int)r[j];
    }
  }
  if (!ok && tid == 0) *err = 1;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

static void cut(long long x, int8_t* d, int n) {
  for (int s = 0; s < n; ++s) {
    const int8_t lo = (int8_t)(x & 0xff);
    d[s] = lo;
    x = (x - lo) >> 8;
  }
}

int main() {
  srand(11);
  std::vector<double> X(M * K), A(K * L);
  for (auto& v : X) v = (rand() / (double)RAND_MAX * 2 - 1) * ((rand() % 7) ? 1.0 : 1e-3);
  for (auto& v : A) v = (rand() / (double)RAND_MAX * 2 - 1) * 0.05;
  std::vector<double> c(K, 0.0), sa(L, 0.0), se(L, 0.0);
  for (int i = 0; i < M; ++i)
    for (int f = 0; f < K; ++f) c[f] = fmax(c[f], fabs(X[i * K + f]));
  double rown = 0.0;
  for (int i = 0; i < M; ++i) {
    double s = 0.0;
    for (int f = 0; f < K; ++f) s += (X[i * K + f] / c[f]) * (X[i * K + f] / c[f]);
    rown = fmax(rown, sqrt(s));
  }
  for (int a = 0; a < L; ++a) {
    double s = 0.0;
    for (int f = 0; f < K; ++f) {
      sa[a] = fmax(sa[a], fabs(c[f] * A[f * L + a]));
      s += (c[f] * A[f * L + a]) * (c[f] * A[f * L + a]);
    }
    se[a] = rown * sqrt(s);   // |Ez_a| <= |x'| |A'_a|
  }
  std::vector<int8_t> Xp((size_t)NPX * M * K), Ap((size_t)NPX * L * K);
  const double S = ldexp(1.0, 54);
  for (int i = 0; i < M; ++i)
    for (int f = 0; f < K; ++f) {
      int8_t d[NPX];
      cut(llrint(X[i * K + f] / c[f] * S), d, NPX);
      for (int p = 0; p < NPX; ++p) Xp[((size_t)p * M + i) * K + f] = d[p];
    }
  for (int f = 0; f < K; ++f)
    for (int a = 0; a < L; ++a) {
      int8_t d[NPX];
      cut(llrint(c[f] * A[f * L + a] / sa[a] * S), d, NPX);
      for (int q = 0; q < NPX; ++q) Ap[((size_t)q * L + a) * K + f] = d[q];
    }
  int8_t *dX, *dA;
  double *dsa, *dse, *dEz;
  long long* dEi;
  int *dD1, *dD2, *dD3, *derr;
  cudaMalloc(&dX, Xp.size());
  cudaMalloc(&dA, Ap.size());
  cudaMalloc(&dsa, 8 * L);
  cudaMalloc(&dse, 8 * L);
  cudaMalloc(&dEz, 8 * M * L);
  cudaMalloc(&dEi, 8 * M * L);
  cudaMalloc(&dD1, 4 * M * NB1);
  cudaMalloc(&dD2, 4 * K * NB1);
  cudaMalloc(&dD3, 4 * NB2 * NB2);
  cudaMalloc(&derr, 4);
  cudaMemset(derr, 0, 4);
  cudaMemcpy(dX, Xp.data(), Xp.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dA, Ap.data(), Ap.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dsa, sa.data(), 8 * L, cudaMemcpyHostToDevice);
  cudaMemcpy(dse, se.data(), 8 * L, cudaMemcpyHostToDevice);
  const int smem = 32768 + 2 * NB1 * 128 + NB2 * 128 + 1024;
  cudaFuncSetAttribute(proto2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  proto2<<<1, 128, smem>>>(dX, dA, dsa, dse, dD1, dEz, dEi, dD2, dD3, derr);
  cudaError_t rc = cudaDeviceSynchronize();
  std::vector<int> D1(M * NB1), D2(K * NB1), D3(NB2 * NB2);
  std::vector<double> Ez(M * L);
  std::vector<long long> Ei(M * L);
  int herr = 0;
  cudaMemcpy(D1.data(), dD1, 4 * D1.size(), cudaMemcpyDeviceToHost);
  cudaMemcpy(D2.data(), dD2, 4 * D2.size(), cudaMemcpyDeviceToHost);
  cudaMemcpy(D3.data(), dD3, 4 * D3.size(), cudaMemcpyDeviceToHost);
  cudaMemcpy(Ez.data(), dEz, 8 * Ez.size(), cudaMemcpyDeviceToHost);
  cudaMemcpy(Ei.data(), dEi, 8 * Ei.size(), cudaMemcpyDeviceToHost);
  cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost);
  printf("kernel: %s, timeout flag %d\n", cudaGetErrorString(rc), herr);

  // (1) step-1 integers, (2) Ez, (3) Ez fixed point
  long bad1 = 0, badez = 0, badei = 0;
  double ez_err = 0.0;
  std::vector<long double> EzRef(M * L);
  for (int i = 0; i < M; ++i)
    for (int a = 0; a < L; ++a) {
      __int128 T = 0;
      for (int cls = 0; cls < NPX; ++cls) {
        long long ref = 0;
        for (int p = 0; p < NPX; ++p) {
          const int q = cls + 6 - p;
          if (q < 0 || q >= NPX) continue;
          for (int f = 0; f < K; ++f) ref += (long long)Xp[((size_t)p * M + i) * K + f] * Ap[((size_t)q * L + a) * K + f];
        }
        if (ref != D1[i * NB1 + 16 * cls + a]) ++bad1;
        T += (__int128)D1[i * NB1 + 16 * cls + a] << (8 * cls);
      }
      const double ez_h = (i128_to_double(T) * 0x1p-60) * sa[a];
      if (ez_h != Ez[i * L + a]) ++badez;
      const long long ei_h = llrint((ez_h / se[a]) * 4611686018427387904.0);
      if (llabs(ei_h - Ei[i * L + a]) > 0) ++badei;
      long double t = 0.0L, tabs = 0.0L;
      for (int f = 0; f < K; ++f) {
        t += (long double)X[i * K + f] * A[f * L + a];
        tabs += fabsl((long double)X[i * K + f] * A[f * L + a]);
      }
      EzRef[i * L + a] = t;
      ez_err = fmax(ez_err, (double)(fabsl(Ez[i * L + a] - t) / tabs));
    }
  printf("step 1: %ld integer mismatches; hand-off: %ld Ez values and %ld fixed-point values differ from the host restatement; "
         "max |Ez - Ez_ld| / sum|x a| = %.3e\n", bad1, badez, badei, ez_err);
  // Ez planes from the DEVICE's fixed-point values
  std::vector<int8_t> Ep((size_t)NPE * M * L);
  for (int i = 0; i < M; ++i)
    for (int a = 0; a < L; ++a) {
      int8_t d[NPE];
      cut(Ei[i * L + a], d, NPE);
      for (int q = 0; q < NPE; ++q) Ep[((size_t)q * M + i) * L + a] = d[q];
    }
  // (4) step-2 integers and the assembled XEz
  long bad2 = 0;
  double xez_err = 0.0;
  for (int f = 0; f < K; ++f)
    for (int a = 0; a < L; ++a) {
      __int128 T = 0;
      for (int cls = 0; cls < NPX; ++cls) {
        long long ref = 0;
        for (int p = 0; p < NPX; ++p) {
          const int q = cls + 7 - p;
          if (q < 0 || q >= NPE) continue;
          for (int i = 0; i < M; ++i) ref += (long long)Xp[((size_t)p * M + i) * K + f] * Ep[((size_t)q * M + i) * L + a];
        }
        if (ref != D2[f * NB1 + 16 * cls + a]) ++bad2;
        T += (__int128)D2[f * NB1 + 16 * cls + a] << (8 * cls);
      }
      const double xez = (i128_to_double(T) * ldexp(1.0, 56 - 54 - 62)) * c[f] * se[a];   // classes start at 256^7
      long double t = 0.0L, tabs = 0.0L;
      for (int i = 0; i < M; ++i) {
        t += (long double)X[i * K + f] * EzRef[i * L + a];
        tabs += fabsl((long double)X[i * K + f] * EzRef[i * L + a]);
      }
      xez_err = fmax(xez_err, (double)(fabsl(xez - t) / tabs));
    }
  // (5) Ez^T Ez integers (all plane pairs) and the assembled matrix from the 7 top classes
  long bad3 = 0;
  double ee_err = 0.0;
  for (int a = 0; a < L; ++a)
    for (int b = 0; b < L; ++b) {
      __int128 T = 0;
      for (int q = 0; q < NPE; ++q)
        for (int q2 = 0; q2 < NPE; ++q2) {
          long long ref = 0;
          for (int i = 0; i < M; ++i) ref += (long long)Ep[((size_t)q * M + i) * L + a] * Ep[((size_t)q2 * M + i) * L + b];
          if (ref != D3[(16 * q + a) * NB2 + 16 * q2 + b]) ++bad3;
          if (q + q2 >= 8) T += (__int128)D3[(16 * q + a) * NB2 + 16 * q2 + b] << (8 * (q + q2 - 8));
        }
      const double ee = (i128_to_double(T) * ldexp(1.0, 64 - 124)) * se[a] * se[b];
      long double t = 0.0L, tabs = 0.0L;
      for (int i = 0; i < M; ++i) {
        t += EzRef[i * L + a] * EzRef[i * L + b];
        tabs += fabsl(EzRef[i * L + a] * EzRef[i * L + b]);
      }
      ee_err = fmax(ee_err, (double)(fabsl(ee - t) / tabs));
    }
  printf("step 2 (MN-major X): %ld integer mismatches, max |XEz - ref| / sum|x ez| = %.3e;  Ez^T Ez: %ld integer mismatches, "
         "max |EE - ref| / sum|ez ez| = %.3e\n", bad2, xez_err, bad3, ee_err);
  // Integer results must be exact.  The FP64 figures are relative to sum |terms| against a long-double reference that uses
  // the UNROUNDED Ez; rounding Ez to FP64 (any FP64 pipeline does) already costs ~1e-16 sum|x a| / |Ez| in the second-order
  // statistics, so these bounds are an order above the dot-product rounding level, not a statement about the planes.
  const bool pass = rc == cudaSuccess && !herr && bad1 == 0 && bad2 == 0 && bad3 == 0 && badez == 0 && badei == 0 && ez_err < 1e-15 &&
                    xez_err < 1e-14 && ee_err < 1e-13;
  printf("%s\n", pass ? "PASS" : "FAIL");
  return pass ? 0 : 1;
}
