// dense_i8.cu -- EXPERIMENTAL, OFF BY DEFAULT (GPLVM_DENSE_I8=1): the dense EM statistics of dense_dmma.cu
//   XEz (D x 16) = Xc^T . Ez,   Ez = Xc . A (A = W M^-1),   EE (16 x 16) = Ez^T Ez          (emStepsFast.stepOne, PPCA.hs:63-70)
// as an FP64-faithful product of BYTE PLANES on tcgen05 kind::i8.  Status: the kernel structure is
// the one of tools/dense_i8_stream.cu, which passed its single-tile-per-CTA run on a B200; the multi-tile pipeline and
// this integration (plane preparation, per-step projector planes, in-kernel assembly) have NOT been run yet, so the
// product path is unchanged unless the switch is set.  Numerical design: oracle/fixedpoint.py:dense_stats_i8.
//
//   once per data set : c_f = max |Xc[:, f]|,  X' = Xc / c_f cut into 7 balanced byte planes (|x'| <= 1, 2^54 grid), stored
//                       per 128-observation tile as [plane][2 feature blocks][128 rows x 128 B, SWIZZLE_128B]; rown = max |x'_i|
//   once per step     : A' = c_f A, column scales sa_a = max |A'_a|, se_a = rown |A'_a|_2 (bound on |Ez_a|), 7 planes of A' / sa
//   statistics kernel : one persistent CTA per SM.  load warp: cp.async.bulk of plane tiles into a 4-stage ring, order
//                       S1(0) | S1(1) S2(0) | ...;  MMA thread: S1(t) = 7 planes x 8 MMAs -> D1[t & 1] (class p + q - 6 ->
//                       columns 16 class + a), S2(t) = 7 planes x 2 feature blocks x 4 MMAs with the X tile read MN-major ->
//                       D2, plus 4 MMAs Ez planes x Ez planes (q' >= 4) -> D3;  4 epilogue warps (thread = TMEM lane):
//                       D1 -> 128-bit sums -> Ez (one rounding) -> 8 planes -> B2[t & 1]; D2 / D3 are added into the CTA's
//                       int64 scratch every 64 tiles; at the end the class sums are assembled into the SAME per-CTA FP64
//                       partial block as the DMMA kernel writes, so reduction / all-reduce / update are untouched.
#include "ppca.cuh"
#include "tma.cuh"

namespace gplvm {
namespace {

constexpr int TM = 128, KF = 256, L = 16, NPX = 7, NPE = 8;
constexpr int NB1 = L * NPX;               // 112: projector-plane rows; columns of D1 and of D2 per feature block
constexpr int NB2 = L * NPE;               // 128: Ez-plane rows
constexpr int ND3 = 64;                    // D3 columns: Ez planes q' = 4..7
constexpr int C1 = 0, C2 = 2 * NB1, C3 = 4 * NB1;   // TMEM column bases 0, 224, 448
constexpr int RS = 4;
constexpr int PLANE_BYTES = 32768;
constexpr int TILE_BYTES = NPX * PLANE_BYTES;
constexpr int B1KB = NB1 * 128;
constexpr int FLUSH = 64;                  // 7 pairs x 128 x 2^14 x 64 < 2^31
constexpr int EPI = 128;
constexpr int THREADS = EPI + 64;          // warps 0..3 epilogue, warp 4 loads, warp 5 issues
constexpr long long SPIN = 1ll << 24;
constexpr int SCRATCH_LL = KF * NB1 + NB2 * ND3;   // int64 words of scratch per CTA
constexpr int SMEM_BYTES = RS * PLANE_BYTES + 2 * B1KB + 2 * NB2 * 128 + 1024;

__device__ __forceinline__ uint64_t i8_desc(uint32_t addr) {   // K-major or MN-major SWIZZLE_128B tile, 1024-byte row groups
  return (uint64_t)((addr & 0x3ffff) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ inline uint32_t i8_sw128(int r, int b) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((b >> 4) ^ (r & 7))) << 4) + (b & 15)); }
__device__ __forceinline__ bool i8_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (long long spin = 0; spin < SPIN && !done; ++spin)
    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.b32 %0, 1, 0, q;\n\t}\n"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
  return done != 0;
}
__device__ __forceinline__ void i8_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred pp;\n\tsetp.ne.b32 pp, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, pp;\n\t}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void i8_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void i8_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ double i8_to_double(__int128 T) {   // |T| < 2^77, one rounding
  const long long thi = (long long)(T >> 24), tlo = (long long)(T & 0xffffff);
  return fma((double)thi, 16777216.0, (double)tlo);
}
__device__ __forceinline__ uint32_t i8_idesc(int n, int a_mn) {   // D = S32, A = B = s8, M = 128
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}
__device__ __forceinline__ double bits_to_scale(unsigned long long b) {   // a maximum kept as the bits of a non-negative double
  const double v = __longlong_as_double((long long)b);
  return v > 0.0 ? v : 1.0;
}

// ---- once per data set ---------------------------------------------------------------------------------------------------------
// cmax[f] = bits of max_i |Xc[i][f]| (non-negative doubles order like their bit patterns: atomicMax is exact and order-free)
__global__ void __launch_bounds__(KF) i8_colmax_kernel(const double* __restrict__ X, int64_t ldx, int64_t n, int64_t rows_per_cta,
                                                       unsigned long long* __restrict__ cmax) {
  const int f = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta, r1 = min(n, r0 + rows_per_cta);
  double m = 0.0;
  for (int64_t i = r0; i < r1; ++i) m = fmax(m, fabs(X[i * ldx + f]));
  atomicMax(&cmax[f], (unsigned long long)__double_as_longlong(m));
}
// rown2 = bits of max_i sum_f (Xc[i][f] / c_f)^2
__global__ void __launch_bounds__(256) i8_rownorm_kernel(const double* __restrict__ X, int64_t ldx, int64_t n, const unsigned long long* __restrict__ cmax,
                                                         unsigned long long* __restrict__ rown2) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * 8;
  double best = 0.0;
  for (int64_t i = w; i < n; i += nw) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < KF / 32; ++k) {
      const int f = lane + 32 * k;
      const double v = X[i * ldx + f] / bits_to_scale(cmax[f]);
      s = fma(v, v, s);
    }
    s = warp_sum(s);
    best = fmax(best, s);
  }
  if (lane == 0) atomicMax(rown2, (unsigned long long)__double_as_longlong(best));
}
// planes of one 128-observation tile: thread = feature, 7 byte stores per entry into the swizzled plane tiles
__global__ void __launch_bounds__(KF) i8_cut_kernel(const double* __restrict__ X, int64_t ldx, int64_t n, const unsigned long long* __restrict__ cmax,
                                                    uint8_t* __restrict__ planes) {
  const int f = threadIdx.x;
  const int64_t t = blockIdx.x;
  const double c = bits_to_scale(cmax[f]);
  uint8_t* tile = planes + (size_t)t * TILE_BYTES + (f >> 7) * 16384;
  for (int r = 0; r < TM; ++r) {
    const int64_t i = t * TM + r;
    long long xi = (i < n) ? __double2ll_rn((X[i * ldx + f] / c) * 0x1p54) : 0ll;
    const uint32_t off = i8_sw128(r, f & 127);
#pragma unroll
    for (int p = 0; p < NPX; ++p) {
      const long long lo = (long long)(int8_t)(xi & 0xff);
      tile[(size_t)p * PLANE_BYTES + off] = (uint8_t)lo;
      xi = (xi - lo) >> 8;
    }
  }
}

// ---- once per step: projector planes and scales (1 CTA, thread = feature) -----------------------------------------------------------
// scales: [0, 16) sa_a, [16, 32) se_a
__global__ void __launch_bounds__(KF) i8_proj_kernel(const double* __restrict__ proj, const unsigned long long* __restrict__ cmax,
                                                     const unsigned long long* __restrict__ rown2, uint8_t* __restrict__ B1g,
                                                     double* __restrict__ scales, const double* __restrict__ scal) {
  if (scal[S_DONE] != 0.0) return;
  __shared__ double red[40];
  __shared__ double s_sa[L];
  const int f = threadIdx.x;
  const double c = bits_to_scale(cmax[f]);
  const double rown = sqrt(__longlong_as_double((long long)*rown2));
  double a[L];
#pragma unroll
  for (int k = 0; k < L; ++k) a[k] = c * proj[f * L + k];
#pragma unroll
  for (int k = 0; k < L; ++k) {
    const double mx = block_max(fabs(a[k]), red);
    const double ss = block_sum(a[k] * a[k], red);
    if (f == 0) {
      const double sa = mx > 0.0 ? mx : 1.0, se = rown * sqrt(ss);
      s_sa[k] = sa;
      scales[k] = sa;
      scales[L + k] = se > 0.0 ? se : 1.0;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < L; ++k) {
    long long ai = __double2ll_rn((a[k] / s_sa[k]) * 0x1p54);
#pragma unroll
    for (int q = 0; q < NPX; ++q) {
      const long long lo = (long long)(int8_t)(ai & 0xff);
      B1g[(f >> 7) * B1KB + i8_sw128(16 * q + k, f & 127)] = (uint8_t)lo;
      ai = (ai - lo) >> 8;
    }
  }
}

// ---- statistics kernel -------------------------------------------------------------------------------------------------------------
enum { B_FULL = 0, B_EMPTY = RS, B_D1 = 2 * RS, B_B2 = 2 * RS + 2, B_D2DONE = 2 * RS + 4, B_D2FREE = 2 * RS + 5, B_COUNT = 2 * RS + 6 };

__global__ void __launch_bounds__(THREADS, 1) dense_stats_i8_kernel(const uint8_t* __restrict__ Xpl, const uint8_t* __restrict__ B1g,
                                                                    const double* __restrict__ scales, const unsigned long long* __restrict__ cmax,
                                                                    int tiles_per_cta, int ntiles, long long* __restrict__ scratch,
                                                                    double* __restrict__ partials, int64_t S, double* __restrict__ scal) {
  if (scal[S_DONE] != 0.0) return;
  extern __shared__ __align__(1024) uint8_t i8_smem_raw[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) unsigned long long bars[B_COUNT];
  const uint32_t raw_u = smem_u32(i8_smem_raw), base_u = (raw_u + 1023u) & ~1023u;
  uint8_t* base = i8_smem_raw + (base_u - raw_u);
  const uint32_t ring_u = base_u, b1_u = base_u + RS * PLANE_BYTES, b2_u = b1_u + 2 * B1KB;
  uint8_t* B2p = base + RS * PLANE_BYTES + 2 * B1KB;
  const uint32_t bar_u = smem_u32(bars);
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int t0 = (int)blockIdx.x * tiles_per_cta;
  const int T = max(0, min(tiles_per_cta, ntiles - t0));
  long long* xec = scratch + (size_t)blockIdx.x * SCRATCH_LL;   // [256 features][112 class sums]
  long long* eec = xec + KF * NB1;                               // [128 rows (q, a)][64 columns (q' - 4, b)]
  double* out = partials + (size_t)blockIdx.x * S;

  for (int i = tid; i < 2 * B1KB / 16; i += THREADS) reinterpret_cast<uint4*>(base + RS * PLANE_BYTES)[i] = reinterpret_cast<const uint4*>(B1g)[i];
  if (tid == 0) {
    for (int s = 0; s < RS; ++s) {
      mbar_init(bar_u + 8 * (B_FULL + s), 1);
      mbar_init(bar_u + 8 * (B_EMPTY + s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_u + 8 * (B_D1 + b), 1);
      mbar_init(bar_u + 8 * (B_B2 + b), EPI);
    }
    mbar_init(bar_u + 8 * B_D2DONE, 1);
    mbar_init(bar_u + 8 * B_D2FREE, EPI);
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
          if (use > 0) ok = i8_wait(bar_u + 8 * (B_EMPTY + s), (uint32_t)((use - 1) & 1));
          if (!ok) break;
          const uint32_t fb = bar_u + 8 * (B_FULL + s);
          mbar_expect_tx(fb, PLANE_BYTES);
          bulk_load_1d(ring_u + s * PLANE_BYTES, Xpl + ((size_t)(t0 + t) * NPX + p) * PLANE_BYTES, PLANE_BYTES, fb);
        }
      };
      load_tile_planes(0);                       // S1(0)
      for (int t = 0; t < T && ok; ++t) {
        if (t + 1 < T) load_tile_planes(t + 1);  // S1(t + 1)
        load_tile_planes(t);                     // S2(t): second read of the tile's planes (L2)
      }
      if (!ok) set_err(scal, 3.0);
    }
  } else if (w == 5) {
    // ------------------------------------------------------------------ MMA thread --------------------------------------
    if (lane == 0 && T > 0) {
      long long j = 0;
      int flushes = 0;          // read-outs of D2 / D3 requested so far
      bool fresh = true;        // D2 / D3 hold nothing (the next MMAs overwrite)
      auto wait_stage = [&](int& s) {
        s = (int)(j % RS);
        ok = ok && i8_wait(bar_u + 8 * (B_FULL + s), (uint32_t)((j / RS) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      };
      auto step1 = [&](int t) {
        const uint32_t d1 = tmem + C1 + (uint32_t)((t & 1) * NB1);
        for (int p = NPX - 1; p >= 0 && ok; --p, ++j) {
          int s;
          wait_stage(s);
          if (!ok) break;
          const uint32_t idesc = i8_idesc(L * (p + 1), 0), xs = ring_u + s * PLANE_BYTES;
#pragma unroll
          for (int ks = 0; ks < KF / 32; ++ks)
            i8_mma(d1, i8_desc(xs + (ks >> 2) * 16384) + 2 * (ks & 3),
                   i8_desc(b1_u + (ks >> 2) * B1KB + (uint32_t)(L * (NPX - 1 - p)) * 128) + 2 * (ks & 3), idesc,
                   (p == NPX - 1 && ks == 0) ? 0u : 1u);
          i8_commit(bar_u + 8 * (B_EMPTY + s));
        }
        i8_commit(bar_u + 8 * (B_D1 + (t & 1)));
      };
      auto step2 = [&](int t) {
        const int buf = t & 1;
        ok = ok && i8_wait(bar_u + 8 * (B_B2 + buf), (uint32_t)((t >> 1) & 1));                          // Ez planes of tile t are in B2[buf]
        if (fresh && flushes > 0) ok = ok && i8_wait(bar_u + 8 * B_D2FREE, (uint32_t)((flushes - 1) & 1));   // D2 / D3 were read out
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ez = b2_u + buf * (NB2 * 128);
        for (int p = NPX - 1; p >= 0 && ok; --p, ++j) {
          int s;
          wait_stage(s);
          if (!ok) break;
          const uint32_t idesc = i8_idesc(L * (p + 1), 1), xs = ring_u + s * PLANE_BYTES;
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int ks = 0; ks < TM / 32; ++ks)   // MN-major A: K = observations advances by 32 rows = 4096 B = 256 units
              i8_mma(tmem + C2 + kb * NB1, i8_desc(xs + kb * 16384) + 256 * ks, i8_desc(ez + (uint32_t)(L * (NPE - 1 - p)) * 128) + 2 * ks, idesc,
                     (fresh && p == NPX - 1 && ks == 0) ? 0u : 1u);
          i8_commit(bar_u + 8 * (B_EMPTY + s));
        }
        const uint32_t idesc3 = i8_idesc(ND3, 0);
#pragma unroll
        for (int ks = 0; ks < TM / 32; ++ks)      // rows (q, a) x columns (q' >= 4, b)
          i8_mma(tmem + C3, i8_desc(ez) + 2 * ks, i8_desc(ez + (NB2 - ND3) * 128) + 2 * ks, idesc3, (fresh && ks == 0) ? 0u : 1u);
        fresh = false;
        if ((t + 1) % FLUSH == 0 || t + 1 == T) {
          i8_commit(bar_u + 8 * B_D2DONE);
          ++flushes;
          fresh = true;
        }
      };
      step1(0);
      for (int t = 0; t < T && ok; ++t) {
        if (t + 1 < T) step1(t + 1);
        step2(t);
      }
      if (!ok) set_err(scal, 3.0);
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps ----------------------------------
    const int row = 32 * w + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(32 * w) << 16);
    int flushes = 0;
    auto read_out = [&]() {    // D2 (thread = feature, 2 blocks) and D3 (thread = Ez-plane row) into the int64 scratch
      ok = ok && i8_wait(bar_u + 8 * B_D2DONE, (uint32_t)(flushes & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const bool first = flushes == 0;
      for (int kb = 0; kb < 2; ++kb)
        for (int c0 = 0; c0 < NB1; c0 += 16) {
          uint32_t r[16];
          i8_ld16(lane_base + (uint32_t)(C2 + kb * NB1 + c0), r);
          long long* dst = xec + (size_t)(kb * 128 + row) * NB1 + c0;
#pragma unroll
          for (int q = 0; q < 16; ++q) dst[q] = (first ? 0ll : dst[q]) + (long long)(int)r[q];
        }
      for (int c0 = 0; c0 < ND3; c0 += 16) {
        uint32_t r[16];
        i8_ld16(lane_base + (uint32_t)(C3 + c0), r);
        long long* dst = eec + (size_t)row * ND3 + c0;
#pragma unroll
        for (int q = 0; q < 16; ++q) dst[q] = (first ? 0ll : dst[q]) + (long long)(int)r[q];
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_u + 8 * B_D2FREE) : "memory");
      ++flushes;
    };
    double sca[L], sce[L];
#pragma unroll
    for (int a = 0; a < L; ++a) sca[a] = scales[a], sce[a] = scales[L + a];
    for (int t = 0; t < T; ++t) {
      const int buf = t & 1;
      ok = ok && i8_wait(bar_u + 8 * (B_D1 + buf), (uint32_t)((t >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      __int128 Tt[L];
#pragma unroll
      for (int a = 0; a < L; ++a) Tt[a] = 0;
#pragma unroll
      for (int cls = 0; cls < NPX; ++cls) {
        uint32_t r[16];
        i8_ld16(lane_base + (uint32_t)(C1 + buf * NB1 + 16 * cls), r);
#pragma unroll
        for (int a = 0; a < L; ++a) Tt[a] += (__int128)(int)r[a] << (8 * cls);
      }
      uint8_t* ezt = B2p + buf * (NB2 * 128);
#pragma unroll
      for (int a = 0; a < L; ++a) {
        const double ez = (i8_to_double(Tt[a]) * 0x1p-60) * sca[a];   // operands carry 2^54 each, kept classes start at 256^6
        long long ei = __double2ll_rn((ez / sce[a]) * 0x1p62);
#pragma unroll
        for (int q = 0; q < NPE; ++q) {
          const long long lo = (long long)(int8_t)(ei & 0xff);
          ezt[i8_sw128(16 * q + a, row)] = (uint8_t)lo;
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
    if (!ok && lane == 0) set_err(scal, 3.0);
    // ---- assembly into the FP64 partial block [XEz (256 x 16) | EE (16 x 16)] -----------------------------------------------------
    // XEz[f][a] = c_f se_a 2^-(54 + 62) sum_cls 256^(cls + 7) xec[f][16 cls + a]: every thread owns the rows it wrote
    for (int kb = 0; kb < 2; ++kb) {
      const int f = kb * 128 + row;
      const double cf = bits_to_scale(cmax[f]);
#pragma unroll
      for (int a = 0; a < L; ++a) {
        __int128 Ts = 0;
        if (T > 0) {
#pragma unroll
          for (int cls = 0; cls < NPX; ++cls) Ts += (__int128)xec[(size_t)f * NB1 + 16 * cls + a] << (8 * cls);
        }
        out[f * L + a] = ((i8_to_double(Ts) * 0x1p-60) * cf) * sce[a];
      }
    }
    __threadfence_block();
    asm volatile("bar.sync 2, %0;" ::"n"(EPI) : "memory");   // the EE rows were written by other epilogue threads
    // EE[a][b] = se_a se_b 2^-124 sum over plane pairs q + q' >= 8 of 256^(q + q') P(q, a; q', b); the scratch holds the pairs
    // with q' >= 4, a pair with q < 4 is also the transposed pair of entry (b, a)
    for (int e = row; e < L * L; e += EPI) {
      const int a = e >> 4, b = e & 15;
      __int128 Ts = 0;
      if (T > 0) {
        for (int q = 0; q < NPE; ++q)
          for (int q2 = 4; q2 < NPE; ++q2) {
            if (q + q2 < 8) continue;
            const int sh = 8 * (q + q2 - 8);
            Ts += (__int128)eec[(size_t)(16 * q + a) * ND3 + 16 * (q2 - 4) + b] << sh;
            if (q < 4) Ts += (__int128)eec[(size_t)(16 * q + b) * ND3 + 16 * (q2 - 4) + a] << sh;
          }
      }
      out[KF * L + e] = ((i8_to_double(Ts) * 0x1p-60) * sce[a]) * sce[b];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (w == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

}  // namespace

bool dense_i8_enabled() {
  static const bool on = []() {
    const char* e = getenv("GPLVM_DENSE_I8");
    return e ? (atoi(e) != 0) : false;
  }();
  return on;
}

// planes of the centred observations (once per data set; the session must be centred)
int dense_i8_prepare(gplvm_ppca_session* s, ShardState& sh) {
  if (s->D != KF || s->MP != L) return GPLVM_OK;   // the experimental path covers D = 256, M = 16 only
  const int64_t ntiles = ceil_div(sh.n, (int64_t)TM);
  if (!sh.i8_planes) {
    CUDA_TRY(cudaMalloc(&sh.i8_planes, (size_t)ntiles * TILE_BYTES));
    CUDA_TRY(cudaMalloc(&sh.i8_aux, 4096 + 2 * B1KB));                                   // cmax[256], rown2, scales[32], then the projector planes
    CUDA_TRY(cudaMalloc(&sh.i8_scratch, sizeof(long long) * (size_t)sm_count(sh.dev) * SCRATCH_LL));
    CUDA_TRY(cudaFuncSetAttribute(dense_stats_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  }
  unsigned long long* cmax = reinterpret_cast<unsigned long long*>(sh.i8_aux);
  CUDA_TRY(cudaMemsetAsync(sh.i8_aux, 0, 4096, sh.st));
  const int ctas = 8 * sm_count(sh.dev);
  GP_LAUNCH(i8_colmax_kernel, ctas, KF, 0, sh.st, sh.X, sh.ldx, sh.n, ceil_div(sh.n, (int64_t)ctas), cmax);
  GP_LAUNCH(i8_rownorm_kernel, ctas, 256, 0, sh.st, sh.X, sh.ldx, sh.n, cmax, cmax + KF);
  GP_LAUNCH(i8_cut_kernel, (unsigned)ntiles, KF, 0, sh.st, sh.X, sh.ldx, sh.n, cmax, sh.i8_planes);
  return GPLVM_OK;
}

// statistics pass of one EM step (same contract as dense_dmma_stats)
int dense_i8_stats(gplvm_ppca_session* s, ShardState& sh, const double* proj) {
  unsigned long long* cmax = reinterpret_cast<unsigned long long*>(sh.i8_aux);
  double* scales = reinterpret_cast<double*>(sh.i8_aux + 8 * (KF + 8));
  uint8_t* B1g = sh.i8_aux + 4096;
  GP_LAUNCH(i8_proj_kernel, 1, KF, 0, sh.st, proj, cmax, cmax + KF, B1g, scales, sh.par.scal);
  const int ntiles = (int)ceil_div(sh.n, (int64_t)TM);
  int grid = std::min(sm_count(sh.dev), ntiles);
  const int per = (int)ceil_div((int64_t)ntiles, (int64_t)grid);
  grid = (int)ceil_div((int64_t)ntiles, (int64_t)per);
  if (grid > sh.nparts) return fail(GPLVM_ERR_STATE, "INT8 dense statistics: partial buffer too small");
  GP_LAUNCH(dense_stats_i8_kernel, grid, THREADS, SMEM_BYTES, sh.st, sh.i8_planes, B1g, scales, cmax, per, ntiles, sh.i8_scratch, sh.partials,
            s->S, sh.par.scal);
  return launch_reduce_partials(sh, s->S, grid);
}

}  // namespace gplvm
