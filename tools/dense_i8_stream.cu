// dense_i8_stream.cu -- DRAFT.  Run ONCE on a B200 with the last seconds of the round-1 GPU budget, 1 tile per CTA
// (`./dense_i8_stream 1`): PASS (XEz within 1.0e-16, Ez^T Ez within 5.2e-16 of the reference, no handshake timeout).  The
// multi-tile path (ring reuse, S1(t + 1) / S2(t) interleave, read-outs every 64 tiles) and the timing are NOT yet exercised.
//
// Streaming prototype of the kernel blueprint in DESIGN.md section 7 (FP64-faithful dense EM statistics on tcgen05
// kind::i8): one persistent CTA per SM streams 128-observation tiles of pre-cut, pre-swizzled X planes from HBM and
// produces, per CTA, the exact integer class sums of  XEz = X'^T . Ez  (256 x 16) and  Ez^T Ez  (16 x 16).  It composes
// the pieces that tools/dense_i8_proto.cu and tools/dense_i8_proto2.cu validated one tile at a time:
//
//   load warp     : cp.async.bulk of 32 KB plane tiles into a 4-stage ring, in consumption order
//                   S1(0) | S1(1) S2(0) | S1(2) S2(1) | ...        (the second read of a tile's planes comes from L2)
//   MMA thread    : S1(t): 7 planes x 8 k32 MMAs -> D1[t & 1];   S2(t): 7 planes x 2 feature blocks x 4 MMAs (X tile read
//                   MN-major) -> D2, then 4 MMAs Ez-planes x Ez-planes(q >= 4) -> D3.  Order: S1(t + 1) before S2(t), so the
//                   tensor pipe is busy while the epilogue warps turn tile t's sums into Ez planes.
//   epilogue warps: thread = TMEM lane = observation: read D1, assemble 128-bit sums, round Ez once, cut into 8 planes
//                   -> B2[t & 1];  every FLUSH tiles (and at the end): add D2 / D3 (int32) into the CTA's int64 blocks.
//   TMEM map      : D1[0] 0..111, D1[1] 112..223, D2 224..447 (2 feature blocks x 112), D3 448..511   (512 columns)
//
// The host cuts the operands, assembles the FP64 statistics from the int64 class sums and compares them with a double
// reference; the kernel time gives cycles per tile against the 2 x 3 334-cycle budget measured for the MMAs alone.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/dense_i8_stream tools/dense_i8_stream.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int TM = 128, K = 256, L = 16, NPX = 7, NPE = 8;
constexpr int NB1 = L * NPX;               // 112 projector-plane rows (16 q + a) = columns of D1 and of D2 per feature block
constexpr int NB2 = L * NPE;               // 128 Ez-plane rows
constexpr int ND3 = 64;                    // D3 columns: Ez planes q' = 4..7
constexpr int C1 = 0, C2 = 2 * NB1, C3 = 4 * NB1;   // 0, 224, 448
constexpr int RS = 4;                      // ring stages of 32 KB
constexpr int PLANE_BYTES = 32768;
constexpr int B1KB = NB1 * 128;            // one k-block of the projector planes
constexpr int FLUSH = 64;                  // tiles between read-outs: 7 pairs x 128 x 2^14 x 64 < 2^31
constexpr int EPI = 128;                   // epilogue threads (warps 0..3); warp 4 loads, warp 5 issues
constexpr int THREADS = EPI + 64;
constexpr long long SPIN = 1ll << 20;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  return (uint64_t)((addr & 0x3ffff) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ inline uint32_t sw128(int r, int b) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((b >> 4) ^ (r & 7))) << 4) + (b & 15)); }
__device__ __forceinline__ bool wait_bar(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (long long spin = 0; spin < SPIN && !done; ++spin)
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
__device__ __forceinline__ void commit_to(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__host__ __device__ inline double i128_to_double(__int128 T) {   // |T| < 2^77, one rounding
  const long long thi = (long long)(T >> 24), tlo = (long long)(T & 0xffffff);
  return fma((double)thi, 16777216.0, (double)tlo);
}
__device__ __forceinline__ uint32_t idesc_s8(int n, int a_mn) {   // D = S32, A = B = s8, M = 128
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

// barrier slots (8 bytes each)
enum { B_FULL = 0, B_EMPTY = RS, B_D1 = 2 * RS, B_B2 = 2 * RS + 2, B_D2DONE = 2 * RS + 4, B_D2FREE = 2 * RS + 5, B_COUNT = 2 * RS + 6 };

// Xpl : [tile][plane p][2 feature blocks][128 rows x 128 B, SWIZZLE_128B]   (224 KB per tile)
// B1g : [2 k-blocks][112 rows x 128 B, SWIZZLE_128B]                         (projector planes, 28 KB)
// XEc : per CTA [256 features][112] int64 class sums;  EEc : per CTA [128 rows (q, a)][64] int64 plane-pair sums
__global__ void __launch_bounds__(THREADS, 1) stream_kernel(const uint8_t* __restrict__ Xpl, const uint8_t* __restrict__ B1g,
                                                            const double* __restrict__ sa, const double* __restrict__ se, int tiles_per_cta,
                                                            int ntiles, long long* __restrict__ XEc, long long* __restrict__ EEc, int* err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) unsigned long long bars[B_COUNT];
  const uint32_t raw_u = smem_u32(smem_raw), base_u = (raw_u + 1023u) & ~1023u;
  uint8_t* base = smem_raw + (base_u - raw_u);
  const uint32_t ring_u = base_u;                              // RS x 32 KB
  const uint32_t b1_u = base_u + RS * PLANE_BYTES;             // 28 KB
  const uint32_t b2_u = b1_u + 2 * B1KB;                       // 2 x 16 KB
  uint8_t* B2p = base + RS * PLANE_BYTES + 2 * B1KB;
  const uint32_t bar_u = smem_u32(bars);
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int t0 = (int)blockIdx.x * tiles_per_cta;
  const int T = max(0, min(tiles_per_cta, ntiles - t0));
  long long* xec = XEc + (size_t)blockIdx.x * K * NB1;
  long long* eec = EEc + (size_t)blockIdx.x * NB2 * ND3;

  // projector planes: plain copy (already swizzled); barriers; TMEM
  for (int i = tid; i < 2 * B1KB / 16; i += THREADS) reinterpret_cast<uint4*>(base + RS * PLANE_BYTES)[i] = reinterpret_cast<const uint4*>(B1g)[i];
  if (tid == 0) {
    for (int s = 0; s < RS; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u + 8 * (B_FULL + s)));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u + 8 * (B_EMPTY + s)));
    }
    for (int b = 0; b < 2; ++b) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u + 8 * (B_D1 + b)));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_u + 8 * (B_B2 + b)), "r"(EPI));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u + 8 * B_D2DONE));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_u + 8 * B_D2FREE), "r"(EPI));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (w == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  bool ok = true;

  if (w == 4) {
    // ------------------------------------------------------------------ load warp ---------------------------------------
    if (lane == 0 && T > 0) {
      long long j = 0;   // running plane-load index: stage j % RS, use j / RS
      auto load_tile_planes = [&](int t) {
        for (int p = NPX - 1; p >= 0 && ok; --p, ++j) {
          const int s = (int)(j % RS);
          const long long use = j / RS;
          if (use > 0) ok = wait_bar(bar_u + 8 * (B_EMPTY + s), (uint32_t)((use - 1) & 1));
          if (!ok) break;
          const uint32_t fb = bar_u + 8 * (B_FULL + s);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(PLANE_BYTES) : "memory");
          const uint8_t* src = Xpl + ((size_t)(t0 + t) * NPX + p) * PLANE_BYTES;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ring_u + s * PLANE_BYTES),
                       "l"(src), "r"(PLANE_BYTES), "r"(fb)
                       : "memory");
        }
      };
      load_tile_planes(0);                       // S1(0)
      for (int t = 0; t < T && ok; ++t) {
        if (t + 1 < T) load_tile_planes(t + 1);  // S1(t + 1)
        load_tile_planes(t);                     // S2(t)
      }
      if (!ok) *err = 2;
    }
  } else if (w == 5) {
    // ------------------------------------------------------------------ MMA thread --------------------------------------
    if (lane == 0 && T > 0) {
      long long j = 0;
      int flushes = 0;          // completed read-outs of D2 / D3
      bool fresh = true;        // D2 / D3 hold nothing yet (first MMAs overwrite)
      auto wait_stage = [&](int& s) {
        s = (int)(j % RS);
        ok = ok && wait_bar(bar_u + 8 * (B_FULL + s), (uint32_t)((j / RS) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      };
      auto step1 = [&](int t) {
        const uint32_t d1 = tmem + C1 + (uint32_t)((t & 1) * NB1);
        for (int p = NPX - 1; p >= 0 && ok; --p, ++j) {
          int s;
          wait_stage(s);
          if (!ok) break;
          const uint32_t idesc = idesc_s8(L * (p + 1), 0), xs = ring_u + s * PLANE_BYTES;
#pragma unroll
          for (int ks = 0; ks < K / 32; ++ks)
            mma_i8(d1, desc_sw128(xs + (ks >> 2) * 16384) + 2 * (ks & 3),
                   desc_sw128(b1_u + (ks >> 2) * B1KB + (uint32_t)(L * (NPX - 1 - p)) * 128) + 2 * (ks & 3), idesc,
                   (p == NPX - 1 && ks == 0) ? 0u : 1u);
          commit_to(bar_u + 8 * (B_EMPTY + s));
        }
        commit_to(bar_u + 8 * (B_D1 + (t & 1)));
      };
      auto step2 = [&](int t) {
        const int buf = t & 1;
        ok = ok && wait_bar(bar_u + 8 * (B_B2 + buf), (uint32_t)((t >> 1) & 1));       // Ez planes of tile t are in B2[buf]
        if (fresh && flushes > 0) ok = ok && wait_bar(bar_u + 8 * B_D2FREE, (uint32_t)((flushes - 1) & 1));   // D2 / D3 were read out
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ez = b2_u + buf * (NB2 * 128);
        for (int p = NPX - 1; p >= 0 && ok; --p, ++j) {
          int s;
          wait_stage(s);
          if (!ok) break;
          const uint32_t idesc = idesc_s8(L * (p + 1), 1), xs = ring_u + s * PLANE_BYTES;
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int ks = 0; ks < TM / 32; ++ks)   // MN-major A: K = observations advances by 32 rows = 4096 B = 256 units
              mma_i8(tmem + C2 + kb * NB1, desc_sw128(xs + kb * 16384) + 256 * ks, desc_sw128(ez + (uint32_t)(L * (NPE - 1 - p)) * 128) + 2 * ks,
                     idesc, (fresh && p == NPX - 1 && ks == 0) ? 0u : 1u);
          commit_to(bar_u + 8 * (B_EMPTY + s));
        }
        const uint32_t idesc3 = idesc_s8(ND3, 0);
#pragma unroll
        for (int ks = 0; ks < TM / 32; ++ks)      // rows (q, a) x columns (q' >= 4, b)
          mma_i8(tmem + C3, desc_sw128(ez) + 2 * ks, desc_sw128(ez + (NB2 - ND3) * 128) + 2 * ks, idesc3, (fresh && ks == 0) ? 0u : 1u);
        fresh = false;
        if ((t + 1) % FLUSH == 0 || t + 1 == T) {
          commit_to(bar_u + 8 * B_D2DONE);
          ++flushes;
          fresh = true;
        }
      };
      step1(0);
      for (int t = 0; t < T && ok; ++t) {
        if (t + 1 < T) step1(t + 1);
        step2(t);
      }
      if (!ok) *err = 3;
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps ----------------------------------
    const int row = 32 * w + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(32 * w) << 16);
    int flushes = 0;
    auto read_out = [&]() {    // add D2 (thread = feature, 2 blocks) and D3 (thread = Ez-plane row) into the int64 blocks
      ok = ok && wait_bar(bar_u + 8 * B_D2DONE, (uint32_t)(flushes & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int kb = 0; kb < 2; ++kb)
        for (int c0 = 0; c0 < NB1; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(lane_base + (uint32_t)(C2 + kb * NB1 + c0), r);
          long long* dst = xec + (size_t)(kb * 128 + row) * NB1 + c0;
#pragma unroll
          for (int q = 0; q < 16; ++q) dst[q] += (long long)(int)r[q];
        }
      for (int c0 = 0; c0 < ND3; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(lane_base + (uint32_t)(C3 + c0), r);
        long long* dst = eec + (size_t)row * ND3 + c0;
#pragma unroll
        for (int q = 0; q < 16; ++q) dst[q] += (long long)(int)r[q];
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_u + 8 * B_D2FREE) : "memory");
      ++flushes;
    };
    double sca[L], sce[L];
#pragma unroll
    for (int a = 0; a < L; ++a) sca[a] = sa[a], sce[a] = se[a];
    for (int t = 0; t < T; ++t) {
      const int buf = t & 1;
      ok = ok && wait_bar(bar_u + 8 * (B_D1 + buf), (uint32_t)((t >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      __int128 Tt[L];
#pragma unroll
      for (int a = 0; a < L; ++a) Tt[a] = 0;
#pragma unroll
      for (int cls = 0; cls < NPX; ++cls) {
        uint32_t r[16];
        tmem_ld16(lane_base + (uint32_t)(C1 + buf * NB1 + 16 * cls), r);
#pragma unroll
        for (int a = 0; a < L; ++a) Tt[a] += (__int128)(int)r[a] << (8 * cls);
      }
      uint8_t* ezt = B2p + buf * (NB2 * 128);
#pragma unroll
      for (int a = 0; a < L; ++a) {
        const double ez = (i128_to_double(Tt[a]) * 0x1p-60) * sca[a];
        long long ei = __double2ll_rn((ez / sce[a]) * 0x1p62);
#pragma unroll
        for (int q = 0; q < NPE; ++q) {
          const long long lo = (long long)(int8_t)(ei & 0xff);
          ezt[sw128(16 * q + a, row)] = (uint8_t)lo;
          ei = (ei - lo) >> 8;
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_u + 8 * (B_B2 + buf)) : "memory");
      // the read-out that follows S2(t - 1) is due once this tile's planes are handed over (the MMA thread issues S2(t - 1)
      // before it waits for them), and after the last tile
      if (t >= 1 && (t % FLUSH == 0)) read_out();
    }
    if (T > 0) read_out();
    if (!ok && lane == 0) *err = 4;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (w == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

static void cut(long long x, int8_t* d, int n) {
  for (int s = 0; s < n; ++s) {
    const int8_t lo = (int8_t)(x & 0xff);
    d[s] = lo;
    x = (x - lo) >> 8;
  }
}

int main(int argc, char** argv) {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  const int tiles_per_cta = argc > 1 ? atoi(argv[1]) : 8;
  const int ntiles = sms * tiles_per_cta;
  const long long N = (long long)ntiles * TM;
  printf("%s: %d CTAs x %d tiles = %lld observations\n", prop.name, sms, tiles_per_cta, N);
  srand(5);
  std::vector<double> X((size_t)N * K), A(K * L);
  for (auto& v : X) v = (rand() / (double)RAND_MAX * 2 - 1) * ((rand() % 5) ? 1.0 : 0.01);
  for (auto& v : A) v = (rand() / (double)RAND_MAX * 2 - 1) * 0.05;
  std::vector<double> c(K, 0.0), sa(L, 0.0), se(L, 0.0);
  for (long long i = 0; i < N; ++i)
    for (int f = 0; f < K; ++f) c[f] = fmax(c[f], fabs(X[i * K + f]));
  double rown = 0.0;
  for (long long i = 0; i < N; ++i) {
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
    se[a] = rown * sqrt(s);
  }
  // operands in the device layout
  std::vector<uint8_t> Xpl((size_t)ntiles * NPX * PLANE_BYTES), B1(2 * B1KB);
  const double S = ldexp(1.0, 54);
  for (long long i = 0; i < N; ++i)
    for (int f = 0; f < K; ++f) {
      int8_t d[NPX];
      cut(llrint(X[i * K + f] / c[f] * S), d, NPX);
      const long long t = i / TM;
      const int r = (int)(i % TM);
      for (int p = 0; p < NPX; ++p) Xpl[((size_t)t * NPX + p) * PLANE_BYTES + (f >> 7) * 16384 + sw128(r, f & 127)] = (uint8_t)d[p];
    }
  for (int f = 0; f < K; ++f)
    for (int a = 0; a < L; ++a) {
      int8_t d[NPX];
      cut(llrint(c[f] * A[f * L + a] / sa[a] * S), d, NPX);
      for (int q = 0; q < NPX; ++q) B1[(f >> 7) * B1KB + sw128(16 * q + a, f & 127)] = (uint8_t)d[q];
    }
  uint8_t *dX, *dB1;
  double *dsa, *dse;
  long long *dXE, *dEE;
  int* derr;
  cudaMalloc(&dX, Xpl.size());
  cudaMalloc(&dB1, B1.size());
  cudaMalloc(&dsa, 8 * L);
  cudaMalloc(&dse, 8 * L);
  cudaMalloc(&dXE, 8 * (size_t)sms * K * NB1);
  cudaMalloc(&dEE, 8 * (size_t)sms * NB2 * ND3);
  cudaMalloc(&derr, 4);
  cudaMemcpy(dX, Xpl.data(), Xpl.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dB1, B1.data(), B1.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dsa, sa.data(), 8 * L, cudaMemcpyHostToDevice);
  cudaMemcpy(dse, se.data(), 8 * L, cudaMemcpyHostToDevice);
  const int smem = RS * PLANE_BYTES + 2 * B1KB + 2 * NB2 * 128 + 1024;
  if (cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
    printf("shared memory request of %d B refused\n", smem);
    return 1;
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms = 0;
  cudaError_t rc = cudaSuccess;
  for (int rep = 0; rep < 2 && rc == cudaSuccess; ++rep) {   // second launch is the timed one (first touches HBM / L2 cold)
    cudaMemset(dXE, 0, 8 * (size_t)sms * K * NB1);
    cudaMemset(dEE, 0, 8 * (size_t)sms * NB2 * ND3);
    cudaMemset(derr, 0, 4);
    cudaEventRecord(e0);
    stream_kernel<<<sms, THREADS, smem>>>(dX, dB1, dsa, dse, tiles_per_cta, ntiles, dXE, dEE, derr);
    cudaEventRecord(e1);
    rc = cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms, e0, e1);
  }
  int herr = 0;
  cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost);
  printf("kernel: %s, handshake flag %d, %.3f ms -> %.0f cycles per tile at 1.965 GHz (MMA-only budget 2 x 3334)\n", cudaGetErrorString(rc), herr,
         ms, ms * 1e-3 * 1.965e9 / tiles_per_cta);
  if (rc != cudaSuccess || herr) return 1;
  std::vector<long long> XE((size_t)sms * K * NB1), EE((size_t)sms * NB2 * ND3);
  cudaMemcpy(XE.data(), dXE, 8 * XE.size(), cudaMemcpyDeviceToHost);
  cudaMemcpy(EE.data(), dEE, 8 * EE.size(), cudaMemcpyDeviceToHost);
  // reference in double (Ez rounded to double like any FP64 pipeline), with sum |terms| for the error scale
  std::vector<double> ez((size_t)N * L);
  for (long long i = 0; i < N; ++i)
    for (int a = 0; a < L; ++a) {
      long double t = 0.0L;
      for (int f = 0; f < K; ++f) t += (long double)X[i * K + f] * A[f * L + a];
      ez[i * L + a] = (double)t;
    }
  double xez_err = 0.0, ee_err = 0.0;
  for (int f = 0; f < K; ++f)
    for (int a = 0; a < L; ++a) {
      __int128 Tt = 0;
      for (int b = 0; b < sms; ++b)
        for (int cls = 0; cls < NPX; ++cls) Tt += (__int128)XE[((size_t)b * K + f) * NB1 + 16 * cls + a] << (8 * cls);
      const double got = (double)((long double)Tt * ldexpl(1.0L, 56 - 54 - 62)) * c[f] * se[a];
      long double t = 0.0L, tabs = 0.0L;
      for (long long i = 0; i < N; ++i) {
        t += (long double)X[i * K + f] * ez[i * L + a];
        tabs += fabsl((long double)X[i * K + f] * ez[i * L + a]);
      }
      xez_err = fmax(xez_err, (double)(fabsl(got - t) / tabs));
    }
  for (int a = 0; a < L; ++a)
    for (int b2 = 0; b2 < L; ++b2) {
      // EE[a][b] = sum over plane pairs (q, q') with q + q' >= 8; the device holds columns q' >= 4; a pair with q < 4 is also the
      // transposed pair (q', q) of entry (b, a)
      __int128 Tt = 0;
      for (int blk = 0; blk < sms; ++blk)
        for (int q = 0; q < NPE; ++q)
          for (int q2 = 4; q2 < NPE; ++q2) {
            if (q + q2 < 8) continue;
            const int sh = 8 * (q + q2 - 8);
            Tt += (__int128)EE[((size_t)blk * NB2 + 16 * q + a) * ND3 + 16 * (q2 - 4) + b2] << sh;
            if (q < 4) Tt += (__int128)EE[((size_t)blk * NB2 + 16 * q + b2) * ND3 + 16 * (q2 - 4) + a] << sh;
          }
      const double got = (double)((long double)Tt * ldexpl(1.0L, 64 - 124)) * se[a] * se[b2];
      long double t = 0.0L, tabs = 0.0L;
      for (long long i = 0; i < N; ++i) {
        t += (long double)ez[i * L + a] * ez[i * L + b2];
        tabs += fabsl((long double)ez[i * L + a] * ez[i * L + b2]);
      }
      ee_err = fmax(ee_err, (double)(fabsl(got - t) / tabs));
    }
  printf("max |XEz - ref| / sum|x ez| = %.3e,  max |EE - ref| / sum|ez ez| = %.3e\n", xez_err, ee_err);
  const bool pass = xez_err < 1e-13 && ee_err < 1e-13;
  printf("%s\n", pass ? "PASS" : "FAIL");
  return pass ? 0 : 1;
}
