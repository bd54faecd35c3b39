// masked_imma.cu -- complement sums of the NaN-masked M-step as an EXACT fixed-point product on the INT8 tensor
// cores (mma.sync.m16n8k32.s8.u8 -> s32; 1.14 POP/s on B200, tools/imma_peak.cu, against 37 TFLOP/s of DMMA).
// This is the legacy-path kernel (GPLVM_MISS_IMMA=1); the default is the tcgen05 kernel of masked_utc.cu, which
// computes the same integers.  The derivation below holds for both; oracle/fixedpoint.py restates it on the CPU.
//
//   miss (D x 152) = Miss^T (D x n) . [vech A | ez] (n x 152),   Miss[i][d] = 1 when observation i misses feature d
// (reference: the per-feature sums over the observations that miss / have the feature, emStepsMissed.stepTwo,
// src/GPLVM/PPCA.hs:119-147; complement form).
//
// One factor is exactly 0 / 1, so only the other needs splitting: every [vech A | ez] entry v is rounded ONCE to the
// fixed-point grid of its column class,  x = rint(v 2^(51 - E)),  |v| < 2^E  (E from the largest |entry| of the class --
// vech A or ez -- which the solver kernel tracks with an integer max), and  u = x + 2^51  in [0, 2^52] is cut into its
// 7 bytes.  Byte planes are u8 B-operands, the mask bits expand to s8 A-operands, the products accumulate exactly in
// s32, and a constant-1 column counts the offsets:  sum_i x_i = sum_s 256^s S_s - 2^51 cnt.  All sums are integers =>
// independent of the summation order and of the split of the observations over CTAs: bitwise reproducible.  The one
// rounding per entry is <= 2^-52 of the class maximum, i.e. no worse than the rounding of a single FP64 addition of that
// entry into a running sum; the accumulated sum itself is exact and is rounded to FP64 once at the end.
//
// Kernel: CTA = (column group, observation split).  The 152 columns form 8 groups of 10 and 8 groups of 9
// (10 x 7 + 1 = 71 -> 9 n-tiles, 9 x 7 + 1 = 64 -> 8 n-tiles: < 1 % padding); the groups get 20 / 17 observation splits
// so that the 296 CTAs (2 per SM) carry equal work.  Per 128-observation chunk: every thread loads 4 or 8 entries, cuts
// them and stores byte planes to shared memory (k-contiguous, pitch 160 B: conflict-free 64-bit fragment loads);
// then 4 k32 steps of 2 x NT IMMAs per warp (8 warps x 32 features).  Tiles are double buffered: one barrier per chunk.
// The totals row (sum over ALL observations) is accumulated by the cutting threads as integer sums of the two halves
// of u.  A small finalize kernel adds the splits in int64, assembles the 128-bit integers and rounds once.
// Precision guard: the cutting threads count (sampled) the entries more than 2^20 below their class maximum; above 1 %
// the finalize kernel hands the step to the FP64 product of masked_fused.cu on the device (masked_fx.cuh).
#include "masked_fx.cuh"
#include "ppca.cuh"
#include "tma.cuh"

namespace gplvm {
namespace {

using namespace fx;

constexpr int IM_PITCH = 160;       // bytes per tile row (128 + 32: rows 4 apart in g land in different bank groups)
constexpr int IM_LDP = 72;          // int32 columns per feature row in the per-CTA partial block
constexpr int IM_PF = 3;            // cp.async stages of the entries (prefetch distance: 2 chunks)
constexpr int IM_STAGE_BYTES = (256 + 64) * 32;
constexpr int IM_MS = 4;            // stages of the mask words (DD x 16 B each)
constexpr int IM_G10 = 8;           // groups 0..7: 10 columns each (0..79); groups 8..15: 9 columns each (80..151)

__device__ __forceinline__ void imma_s8u8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// cuts 4 consecutive observations of one column into byte planes and stores plane s as one 32-bit word
__device__ __forceinline__ void cut_store(const double (&v)[4], double sc, uint32_t thr, int64_t o0, int64_t n_obs, uint32_t dst,
                                          unsigned long long& tlo, unsigned long long& thi, uint32_t& small) {
  const int nvalid = (int)min((int64_t)4, max((int64_t)0, n_obs - o0));   // observations past the end: masked out, not totalled
  uint32_t lo[4], hi[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const double y = fma(v[r], sc, 6755399441055744.0);   // 2^52 + 2^51: the mantissa of y is the integer x + 2^51
    lo[r] = (uint32_t)__double2loint(y);
    hi[r] = (uint32_t)__double2hiint(y) - 0x43300000u;
  }
  // precision guard, sampled on the first observation of every quad: entries more than 2^IM_GUARD_BITS below the class
  // maximum keep fewer than 52 - IM_GUARD_BITS bits on the fixed-point grid
  small += (((uint32_t)__double2hiint(v[0]) & 0x7fffffffu) < thr) ? 1u : 0u;
  if (nvalid == 4) {   // hi <= 2^20: the four high words add in 32 bits
    tlo += ((unsigned long long)lo[0] + lo[1]) + ((unsigned long long)lo[2] + lo[3]);
    thi += (hi[0] + hi[1]) + (hi[2] + hi[3]);
  } else {
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (r < nvalid) {
        tlo += lo[r];
        thi += hi[r];
      }
  }
#pragma unroll
  for (int s = 0; s < NSL; ++s) {
    const uint32_t sel = (uint32_t)((s & 3) | (((s & 3) + 4) << 4));
    const uint32_t w01 = (s < 4) ? __byte_perm(lo[0], lo[1], sel) : __byte_perm(hi[0], hi[1], sel);
    const uint32_t w23 = (s < 4) ? __byte_perm(lo[2], lo[3], sel) : __byte_perm(hi[2], hi[3], sel);
    const uint32_t wv = __byte_perm(w01, w23, 0x5410u);
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(dst + (uint32_t)(s * IM_PITCH)), "r"(wv) : "memory");
  }
}

template <int DD, int NC>
__device__ __forceinline__ void miss_imma_body(const double* __restrict__ EA, const uint32_t* __restrict__ mT, int64_t nwT,
                                               int64_t n_obs, int col0, int64_t chunk_begin, int64_t chunk_end,
                                               uint32_t* __restrict__ eamax, int* __restrict__ ipart,
                                               unsigned long long* __restrict__ tpart, uint8_t* smem) {
  // warp tile: MT m16 tiles (DD / 8 features) x all NT n8 tiles.  (A 64-feature x NT / 2 warp tile halves the B-fragment
  // loads but doubles the A-fragment work and measured slower: 8.1 ms vs 6.8 ms.)
  constexpr int MT = DD / 128;
  constexpr int NT = (NC * NSL + 1 + 7) / 8;       // n8 tiles: NC x 7 byte planes + the count column
  constexpr int NROW = NT * 8;
  static_assert(NROW <= IM_LDP, "partial block pitch");
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  uint32_t* smem32 = reinterpret_cast<uint32_t*>(smem);
  for (int e = tid; e < 2 * NROW * (IM_PITCH / 4); e += 256) {
    const int row = (e / (IM_PITCH / 4)) % NROW;
    smem32[e] = (row == NC * NSL) ? 0x01010101u : 0u;   // count column: all ones; padding columns: zero
  }
  __syncthreads();
  int acc[MT][NT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0;

  // cutting assignment: thread = (column (tid & 3) + 4 (tid >> 7), observation quad (tid >> 2) & 31): a warp covers 4
  // columns x 8 consecutive quads, so its plane stores hit 32 different banks (column offsets {0, 24, 16, 8} words + the
  // 8 words of one k32 step) and each load instruction touches 8 rows x 32 contiguous bytes.  Threads 0..63 also take
  // column 8 + (tid & 1) for quad tid >> 1.  Every thread stages ITS OWN entries two chunks ahead with 8-byte cp.async
  // into private shared-memory slots ([row][thread]: conflict-free; no registers, no extra barrier).
  const int ca = (tid & 3) + 4 * (tid >> 7), kqa = (tid >> 2) & 31;
  const int cb = 8 + (tid & 1), kqb = (tid >> 1) & 31;
  const bool has_b = (tid < 64) && (cb < NC);
  const uint32_t hA = eamax[0], hE = eamax[1];
  const double sca = scale_up((col0 + ca < IM_NV) ? hA : hE);
  const double scb = scale_up((col0 + cb < IM_NV) ? hA : hE);
  unsigned long long tlo[2] = {0ull, 0ull}, thi[2] = {0ull, 0ull};
  const bool a_is_A = col0 + ca < IM_NV, b_is_A = col0 + cb < IM_NV;
  const uint32_t thra = (uint32_t)(class_exponent(a_is_A ? hA : hE) - IM_GUARD_BITS) << 20;
  const uint32_t thrb = (uint32_t)(class_exponent(b_is_A ? hA : hE) - IM_GUARD_BITS) << 20;
  uint32_t small[2] = {0u, 0u};
  // k-word (kq & 7) = 4 h + t' of k32 step (kq >> 3) is stored at word 2 t' + h: (b0, b1) of a fragment are one 64-bit load
  auto kpos = [](int kq) { return (uint32_t)(((kq >> 3) * 8 + 2 * (kq & 3) + ((kq >> 2) & 1)) * 4); };
  const uint32_t smem_u = smem_u32(smem);
  const uint32_t stage_u = smem_u + 2 * IM_LDP * IM_PITCH;          // IM_PF stages of (256 + 64) x 32 B private slots
  const uint32_t slot_a = (uint32_t)(tid * 8), slot_b = (uint32_t)(4 * 2048 + (tid & 63) * 8);   // + 2048 r / + 512 r
  // mask words of a chunk: one contiguous DD x 16 B block, copied 2 chunks ahead (16 B per thread) into one of IM_MS
  // stages; stage (i + 2) % 4 was last read in the MMA phase of chunk i - 2, which every warp left before barrier i - 1
  const uint32_t mask_u = stage_u + IM_PF * IM_STAGE_BYTES;
  (void)nwT;

  // EA is allocated in whole chunks (session.cu), so the last chunk may be read past n_obs: those entries meet zero mask
  // bits and are left out of the totals (nvalid in cut_store)
  const double* pa = EA + (chunk_begin * IM_CHUNK + 4 * kqa) * IM_EAS + col0 + ca;
  const double* pb = EA + (chunk_begin * IM_CHUNK + 4 * kqb) * IM_EAS + col0 + cb;
  const uint32_t* pm = mT + (chunk_begin * DD + (tid % DD)) * 4;
  int pf_stage = 0, pf_mstage = 0;
  auto prefetch = [&](int64_t chunk) {
    if (chunk < chunk_end) {
      const uint32_t st = stage_u + (uint32_t)pf_stage * IM_STAGE_BYTES;
#pragma unroll
      for (int r = 0; r < 4; ++r)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(st + slot_a + 2048 * r), "l"(pa + r * IM_EAS) : "memory");
      if (has_b) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(st + slot_b + 512 * r), "l"(pb + r * IM_EAS) : "memory");
      }
      if (tid < DD)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(mask_u + (uint32_t)(pf_mstage * DD * 16 + tid * 16)), "l"(pm)
                     : "memory");
      pa += IM_CHUNK * IM_EAS;
      pb += IM_CHUNK * IM_EAS;
      pm += DD * 4;
      pf_stage = (pf_stage + 1 == IM_PF) ? 0 : pf_stage + 1;
      pf_mstage = (pf_mstage + 1) & (IM_MS - 1);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");   // one group per chunk, empty past the end
  };
  prefetch(chunk_begin);
  prefetch(chunk_begin + 1);

  int cs = 0, cms = 0, cbuf = 0;
  for (int64_t chunk = chunk_begin; chunk < chunk_end; ++chunk) {
    const uint32_t tile_u = smem_u + (uint32_t)(cbuf * NROW * IM_PITCH);
    prefetch(chunk + 2);
    asm volatile("cp.async.wait_group 2;" ::: "memory");     // this thread's entries of `chunk` have landed
    {
      const uint32_t st = stage_u + (uint32_t)cs * IM_STAGE_BYTES;
      double va[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) va[r] = lds64(st + slot_a + 2048 * r);
      cut_store(va, sca, thra, chunk * IM_CHUNK + 4 * kqa, n_obs, tile_u + (uint32_t)(ca * NSL * IM_PITCH) + kpos(kqa), tlo[0], thi[0], small[0]);
      if (has_b) {
        double vb[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) vb[r] = lds64(st + slot_b + 512 * r);
        cut_store(vb, scb, thrb, chunk * IM_CHUNK + 4 * kqb, n_obs, tile_u + (uint32_t)(cb * NSL * IM_PITCH) + kpos(kqb), tlo[1], thi[1], small[1]);
      }
    }
    __syncthreads();   // tile + mask words complete; every warp finished the MMAs of the previous use of the other buffer
    uint32_t mk[MT][2];   // this lane's 32 mask bits of rows g and g + 8 (fragment order, see the transpose kernel)
    {
      const uint32_t mst = mask_u + (uint32_t)(cms * DD * 16);
      cs = (cs + 1 == IM_PF) ? 0 : cs + 1;
      cms = (cms + 1) & (IM_MS - 1);
      cbuf ^= 1;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int r = 0; r < 2; ++r)
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(mk[mt][r]) : "r"(mst + (uint32_t)((w * 16 * MT + mt * 16 + g + 8 * r) * 16 + 4 * t)));
    }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t a[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        a[mt][0] = (mk[mt][0] >> (2 * ks)) & 0x01010101u;
        a[mt][1] = (mk[mt][1] >> (2 * ks)) & 0x01010101u;
        a[mt][2] = (mk[mt][0] >> (2 * ks + 1)) & 0x01010101u;
        a[mt][3] = (mk[mt][1] >> (2 * ks + 1)) & 0x01010101u;
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        uint32_t b0, b1;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];"
                     : "=r"(b0), "=r"(b1)
                     : "r"(tile_u + (uint32_t)((nt * 8 + g) * IM_PITCH + ks * 32 + 8 * t)));
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) imma_s8u8(acc[mt][nt], a[mt], b0, b1);
      }
    }
  }
  // per-CTA exact partial sums
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int f = w * 16 * MT + mt * 16 + g;
      *reinterpret_cast<int2*>(&ipart[(int64_t)f * IM_LDP + nt * 8 + 2 * t]) = make_int2(acc[mt][nt][0], acc[mt][nt][1]);
      *reinterpret_cast<int2*>(&ipart[(int64_t)(f + 8) * IM_LDP + nt * 8 + 2 * t]) = make_int2(acc[mt][nt][2], acc[mt][nt][3]);
    }
  {  // precision-guard counters (integer atomics: order-independent)
    const uint32_t sA = __reduce_add_sync(0xffffffffu, (a_is_A ? small[0] : 0u) + (b_is_A ? small[1] : 0u));
    const uint32_t sE = __reduce_add_sync(0xffffffffu, (a_is_A ? 0u : small[0]) + (b_is_A ? 0u : small[1]));
    if (lane == 0) {
      if (sA) atomicAdd(guard_counts(eamax), (unsigned long long)sA);
      if (sE) atomicAdd(guard_counts(eamax) + 1, (unsigned long long)sE);
    }
  }
  // totals: within a warp the lanes with equal (lane & 3) share a column, warps 0..3 hold columns 0..3 and warps 4..7
  // columns 4..7; column 8 + (lane & 1) lives in warps 0 and 1.  Integer sums: any order gives the same result.
  __syncthreads();
  unsigned long long* tred = reinterpret_cast<unsigned long long*>(smem);   // [8 warps][6 columns][2]
#pragma unroll
  for (int o = 4; o <= 16; o <<= 1) {
    tlo[0] += __shfl_xor_sync(0xffffffffu, tlo[0], o);
    thi[0] += __shfl_xor_sync(0xffffffffu, thi[0], o);
  }
#pragma unroll
  for (int o = 2; o <= 16; o <<= 1) {
    tlo[1] += __shfl_xor_sync(0xffffffffu, tlo[1], o);
    thi[1] += __shfl_xor_sync(0xffffffffu, thi[1], o);
  }
  if (lane < 4) {
    tred[(w * 6 + lane) * 2] = tlo[0];
    tred[(w * 6 + lane) * 2 + 1] = thi[0];
  }
  if (lane < 2) {
    tred[(w * 6 + 4 + lane) * 2] = tlo[1];     // zero outside warps 0 and 1
    tred[(w * 6 + 4 + lane) * 2 + 1] = thi[1];
  }
  __syncthreads();
  if (tid < 20) {
    const int c = tid >> 1, h = tid & 1;
    unsigned long long acc64 = 0ull;
    if (c < 8) {
#pragma unroll
      for (int q = 0; q < 4; ++q) acc64 += tred[(((c >> 2) * 4 + q) * 6 + (c & 3)) * 2 + h];
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) acc64 += tred[((q * 6) + 4 + (c - 8)) * 2 + h];
    }
    tpart[tid] = acc64;
  }
}

// CTA index -> (column group, split): groups 0..7 (10 columns) have s10 splits each, groups 8..15 (9 columns) s9
struct ImmaPlan {
  int s10, s9;
};
__host__ __device__ inline void imma_locate(const ImmaPlan& pl, int cta, int& group, int& split, int& nsplit) {
  if (cta < IM_G10 * pl.s10) {
    group = cta / pl.s10;
    split = cta % pl.s10;
    nsplit = pl.s10;
  } else {
    cta -= IM_G10 * pl.s10;
    group = IM_G10 + cta / pl.s9;
    split = cta % pl.s9;
    nsplit = pl.s9;
  }
}

template <int DD>
__global__ void __launch_bounds__(256, 2) masked_miss_imma_kernel(const double* __restrict__ EA, const uint32_t* __restrict__ mT,
                                                                 int64_t nwT, int64_t n_obs, ImmaPlan pl,
                                                                 uint32_t* __restrict__ eamax, int* __restrict__ ipart,
                                                                 unsigned long long* __restrict__ tpart,
                                                                 const double* __restrict__ scal) {
  if (scal[S_DONE] != 0.0) return;
  extern __shared__ __align__(16) uint8_t im_smem[];
  int group, split, nsplit;
  imma_locate(pl, (int)blockIdx.x, group, split, nsplit);
  const int64_t nchunks = (n_obs + IM_CHUNK - 1) / IM_CHUNK;
  const int64_t c0 = nchunks * split / nsplit, c1 = nchunks * (split + 1) / nsplit;
  int* ip = ipart + (int64_t)blockIdx.x * DD * IM_LDP;
  unsigned long long* tp = tpart + (int64_t)blockIdx.x * 20;
  if (group < IM_G10)
    miss_imma_body<DD, 10>(EA, mT, nwT, n_obs, group * 10, c0, c1, eamax, ip, tp, im_smem);
  else
    miss_imma_body<DD, 9>(EA, mT, nwT, n_obs, 80 + (group - IM_G10) * 9, c0, c1, eamax, ip, tp, im_smem);
}

// dst[f][c] (f < D: complement sums of feature f; f = D: totals over all observations), each rounded to FP64 once
__global__ void __launch_bounds__(256) masked_miss_imma_finalize(const int* __restrict__ ipart, const unsigned long long* __restrict__ tpart,
                                                                 int D, ImmaPlan pl, int64_t n_obs, uint32_t* __restrict__ eamax,
                                                                 double* __restrict__ dst, const double* __restrict__ scal) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const bool done = scal[S_DONE] != 0.0;
  const bool fallback = !done && guard_fallback(eamax, n_obs);
  if (idx == 0) guard_scal(eamax)[S_DONE] = fallback ? 0.0 : 1.0;   // 0.0: the FP64 kernels that follow redo the sums
  if (done || fallback || idx >= (D + 1) * IM_EAS) return;
  const int f = idx / IM_EAS, c = idx % IM_EAS;
  const int group = (c < 80) ? c / 10 : IM_G10 + (c - 80) / 9;
  const int lc = (c < 80) ? c % 10 : (c - 80) % 9;
  const int nc = (group < IM_G10) ? 10 : 9;
  const int nsplit = (group < IM_G10) ? pl.s10 : pl.s9;
  const int cta0 = (group < IM_G10) ? group * pl.s10 : IM_G10 * pl.s10 + (group - IM_G10) * pl.s9;
  __int128 T = 0;
  if (f < D) {
    long long S[NSL], cnt = 0;
#pragma unroll
    for (int s = 0; s < NSL; ++s) S[s] = 0;
    for (int sp = 0; sp < nsplit; ++sp) {
      const int* base = ipart + ((int64_t)(cta0 + sp) * D + f) * IM_LDP;
#pragma unroll
      for (int s = 0; s < NSL; ++s) S[s] += base[lc * NSL + s];
      cnt += base[nc * NSL];
    }
#pragma unroll
    for (int s = 0; s < NSL; ++s) T += (__int128)S[s] << (8 * s);
    T -= (__int128)cnt << FXP;
  } else {
    unsigned long long lo = 0ull, hi = 0ull;
    for (int sp = 0; sp < nsplit; ++sp) {
      lo += tpart[(int64_t)(cta0 + sp) * 20 + lc * 2];
      hi += tpart[(int64_t)(cta0 + sp) * 20 + lc * 2 + 1];
    }
    T = ((__int128)hi << 32) + (__int128)lo - ((__int128)n_obs << FXP);
  }
  dst[idx] = int128_to_double(T) * scale_down((c < IM_NV) ? eamax[0] : eamax[1]);
}

// Transposed copy of the mask bits in MMA-fragment order (constant over the EM iterations): one contiguous D x 16 B
// block per 128-observation chunk; word t (0..3) of feature f holds the 32 observations that the lanes with
// threadID_in_group = t need for their A fragments: bit 8 b + 2 ks + h  <=>  observation 32 ks + 16 h + 4 t + b of the
// chunk.  (W >> (2 ks + h)) & 0x01010101 is then the s8 fragment register of k32 step ks, k half h.
__global__ void __launch_bounds__(256) masked_bits_transpose_kernel(const uint32_t* __restrict__ mbits, int wpr, int64_t n_obs,
                                                                    uint32_t* __restrict__ mT, int64_t nchunks) {
  const int lane = threadIdx.x & 31;
  const int64_t wi = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);   // warp = (chunk, word of 32 features)
  if (wi >= nchunks * wpr) return;
  const int64_t chunk = wi / wpr;
  const int wd = (int)(wi % wpr);
  uint32_t S[4];   // lane b: bit i of S[ks] = observation 32 ks + i of the chunk misses feature 32 wd + b
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int64_t n = chunk * IM_CHUNK + ks * 32 + lane;
    const uint32_t wv = (n < n_obs) ? mbits[n * wpr + wd] : 0u;
    uint32_t keep = 0u;
#pragma unroll
    for (int b = 0; b < 32; ++b) {
      const uint32_t m = __ballot_sync(0xffffffffu, (wv >> b) & 1u);
      if (lane == b) keep = m;
    }
    S[ks] = keep;
  }
  uint32_t out[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    uint32_t o = 0u;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int b = 0; b < 4; ++b) o |= ((S[ks] >> (16 * h + 4 * t + b)) & 1u) << (8 * b + 2 * ks + h);
    out[t] = o;
  }
  *reinterpret_cast<uint4*>(&mT[((chunk * wpr + wd) * 32 + lane) * 4]) = make_uint4(out[0], out[1], out[2], out[3]);
}

ImmaPlan make_plan(const ShardState& sh) {
  const int per8 = std::max(2, 2 * sm_count(sh.dev) / 8);
  const int64_t nchunks = ceil_div(sh.n, (int64_t)IM_CHUNK);
  ImmaPlan pl;
  pl.s9 = std::max(1, per8 * 8 / 17);
  pl.s10 = std::max(1, per8 - pl.s9);
  pl.s9 = (int)std::min<int64_t>(pl.s9, nchunks);
  pl.s10 = (int)std::min<int64_t>(pl.s10, nchunks);
  return pl;
}

template <int DD>
int launch_imma(ShardState& sh, double* dst) {
  const ImmaPlan pl = make_plan(sh);
  const int ncta = IM_G10 * pl.s10 + (16 - IM_G10) * pl.s9;
  const int smem = 2 * IM_LDP * IM_PITCH + IM_PF * IM_STAGE_BYTES + IM_MS * DD * 16;   // 23 040 + 30 720 + 16 384 B
  int* ipart = reinterpret_cast<int*>(sh.partials);
  unsigned long long* tpart = reinterpret_cast<unsigned long long*>(ipart + (int64_t)ncta * DD * IM_LDP);
  static bool attr_done[64] = {};
  if (!attr_done[sh.dev & 63]) {
    CUDA_TRY(cudaFuncSetAttribute((masked_miss_imma_kernel<DD>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done[sh.dev & 63] = true;
  }
  GP_LAUNCH((masked_miss_imma_kernel<DD>), ncta, 256, smem, sh.st, sh.EA, sh.mbitsT, sh.nwT, sh.n, pl, sh.eamax, ipart, tpart,
            sh.par.scal);
  GP_LAUNCH(masked_miss_imma_finalize, (unsigned)ceil_div((int64_t)(DD + 1) * IM_EAS, 256), 256, 0, sh.st, ipart, tpart, DD, pl, sh.n,
            sh.eamax, dst, sh.par.scal);
  return GPLVM_OK;
}

}  // namespace

double* masked_imma_guard_scal(ShardState& sh) { return guard_scal(sh.eamax); }

// transposed mask bits for the INT8 complement sums (once per data set)
int masked_imma_prepare(gplvm_ppca_session* s, ShardState& sh) {
  CUDA_TRY(cudaMemsetAsync(sh.eamax, 0, 256, sh.st));
  const int64_t nchunks = sh.nwT / 4;
  GP_LAUNCH(masked_bits_transpose_kernel, (unsigned)ceil_div(nchunks * s->WPR, (int64_t)8), 256, 0, sh.st, sh.mbits, s->WPR, sh.n,
            sh.mbitsT, nchunks);
  return GPLVM_OK;
}

// complement sums [G^miss | S1^miss] (D x 152) + totals (152) into dst (length (D + 1) x 152)
int masked_imma_miss_sums(gplvm_ppca_session* s, ShardState& sh, double* dst) {
  switch (s->D) {
    case 128: return launch_imma<128>(sh, dst);
    case 256: return launch_imma<256>(sh, dst);
    default: return fail(GPLVM_ERR_STATE, "INT8 masked statistics: unsupported D");
  }
}

}  // namespace gplvm
