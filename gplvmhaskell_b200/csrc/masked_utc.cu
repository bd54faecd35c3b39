// masked_utc.cu -- exact fixed-point complement sums of the masked M-step on the 5th-generation tensor cores:
// tcgen05.mma kind::i8 (s8 x u8 -> s32 accumulators in TMEM; 4.41 POP/s measured on B200, tools/tcgen05_i8_peak.cu,
// 3.9x the legacy mma.sync IMMA path).  Arithmetic and derivation: masked_fx.cuh.
//
//   D[f][n] (TMEM, 128 lanes x 144 columns per 128 features)  +=  A[f][k] . B[n][k]      k = observation
//   A = mask bytes (0 / 1),  B row 7 c + s = byte plane s of column c of [vech A | ez],  B row 133 = 1 (offset count)
//
// CTA = (one of 8 groups of 19 columns, observation split); one CTA per SM (2 x 144 accumulator + 2 x 64 A-operand TMEM
// columns for D = 256).  Warps 0..7 are PRODUCERS: per 128-observation chunk they expand the staged mask words into the
// A operand IN TENSOR MEMORY (tcgen05.st: lane = feature, 4 observations per 32-bit column; tools/tcgen05_i8_tmem_a_check.cu)
// and cut their staged entries into the B tile in shared memory, written in the canonical SWIZZLE_128B K-major layout (the
// layout TMA would write; tools/tcgen05_i8_check.cu pins it), then fence.proxy.async + tcgen05.fence + arrive on full[buffer].  Warp 8 lane 0 is the MMA
// ISSUER: waits full[b], issues (D / 128) x 4 MMAs (K = 32 each, the K offset goes into the descriptor start address),
// and tcgen05.commit's to empty[b], which the producers wait on before they overwrite the buffer.  Two operand
// buffers; entries (flattened, coalesced) and mask words are staged 2 chunks ahead by cp.async.  After the
// last chunk warps 0..3 read the accumulators with tcgen05.ld (warp w owns TMEM lanes 32 w ..) and store the per-CTA
// int32 block; the finalize kernel adds the splits in int64, assembles 128-bit integers and rounds once.
// Every wait is bounded (a failed handshake raises S_ERR instead of hanging the GPU).
#include "masked_fx.cuh"
#include "ppca.cuh"
#include "tma.cuh"

namespace gplvm {
namespace {

using namespace fx;

constexpr int UT_NC = 19;                    // columns per group (8 groups x 19 = 152)
constexpr int UT_GROUPS = IM_EAS / UT_NC;
constexpr int UT_N = 144;                    // MMA N: 19 x 7 planes + count row = 134, padded to a multiple of 16
constexpr int UT_CNT = UT_NC * NSL;          // row / column index of the offset count (133)
constexpr int UT_PROD = 320;                 // producer threads: 10 warps x 2 cutting tasks = the 20 (column quad, quad block) tasks
constexpr int UT_PW = UT_PROD / 32;          // producer warps; warp UT_PW issues the MMAs
constexpr int UT_THREADS = UT_PROD + 32;
constexpr int UT_TASKS = 2;                  // (column quad, quad block) tasks per producer thread: 5 x 4 = 20 over 10 warps (8 warps x 3 slots left
                                             // 4 warps one task short and everybody waiting for the others at the chunk barrier)
static_assert(UT_PW * UT_TASKS >= 20, "cutting tasks");
constexpr int UT_B_BYTES = UT_N * 128;       // 18 432
constexpr int UT_STAGE_BYTES = IM_CHUNK * UT_NC * 8;        // staged entries of a chunk: [128 observations][19 columns] FP64
constexpr int UT_PF = 3;                     // cp.async stages (prefetch distance 2)
constexpr long long UT_SPIN = 1ll << 26;
constexpr int UT_NB = 2;                     // operand buffers (B tile in shared memory + A columns in tensor memory); 3: 1.97 vs 1.87 ms at N = 2^22
static_assert(UT_GROUPS * UT_NC == IM_EAS, "column groups");

// byte (row r, k-byte b) of a K-major operand tile: 128-byte rows, 8-row groups of 1024 B, 16-byte units XOR r & 7
__device__ __forceinline__ uint32_t sw128(int r, int b) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((b >> 4) ^ (r & 7))) << 4) + (b & 15));
}
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
  // start address >> 4 | LBO (unused for swizzled K-major) | SBO = 1024 B >> 4 | descriptor version 1 | SWIZZLE_128B
  return (uint64_t)((addr & 0x3ffff) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (long long spin = 0; spin < UT_SPIN && !done; ++spin)
    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.b32 %0, 1, 0, q;\n\t}\n"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
  return done != 0;
}

template <int DD>
struct UtcCfg {
  static constexpr int MT = DD / 128;                       // M = 128 tiles
  // The A operand (mask bytes) lives in TENSOR MEMORY: lane = feature within its 128-row tile, 32-bit column q of a tile =
  // observations 4 q .. 4 q + 3 of the chunk (tools/tcgen05_i8_tmem_a_check.cu pins this layout); 32 columns per tile and
  // buffer.  The MMAs then read only B from shared memory (their A reads were half of the operand traffic that kept the
  // producers' loads and stores waiting), and the producers' 32 KB of A stores per chunk leave the shared-memory pipe.
  static constexpr int A_COLS = MT * 32;                    // TMEM columns of one A buffer
  static constexpr int A_BASE = MT * UT_N;                  // first A column (behind the accumulators)
  static constexpr int BUF_BYTES = UT_B_BYTES;
  static constexpr int TMEM_COLS = (MT * UT_N + UT_NB * A_COLS <= 256) ? 256 : 512;
  static_assert(MT * UT_N + UT_NB * A_COLS <= 512, "tensor memory budget");
  static constexpr int SMEM = UT_NB * BUF_BYTES + UT_PF * UT_STAGE_BYTES + UT_PF * DD * 16 + 1024;
};

template <int DD>
__global__ void __launch_bounds__(UT_THREADS, 1) masked_miss_utc_kernel(const double* __restrict__ EA, const uint32_t* __restrict__ mT,
                                                                        int64_t n_obs, int nsplit, uint32_t* __restrict__ eamax,
                                                                        int* __restrict__ ipart, unsigned long long* __restrict__ tpart,
                                                                        double* __restrict__ scal) {
  using C = UtcCfg<DD>;
  if (scal[S_DONE] != 0.0) return;
  extern __shared__ __align__(1024) uint8_t ut_smem_raw[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) unsigned long long bars[2 * UT_NB + 1];   // full[NB], empty[NB], done
  __shared__ unsigned long long tred[UT_PW * UT_TASKS * 4 * 2];
  const uint32_t raw_u = smem_u32(ut_smem_raw);
  const uint32_t base_u = (raw_u + 1023u) & ~1023u;      // operand tiles need 1024-byte alignment
  uint8_t* base = ut_smem_raw + (base_u - raw_u);
  const uint32_t stage_u = base_u + UT_NB * C::BUF_BYTES;
  const uint32_t mask_u = stage_u + UT_PF * UT_STAGE_BYTES;
  const uint32_t bar_u = smem_u32(bars);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int group = (int)blockIdx.x / nsplit, split = (int)blockIdx.x % nsplit;
  const int col0 = group * UT_NC;
  const int64_t nchunks = (n_obs + IM_CHUNK - 1) / IM_CHUNK;
  const int64_t chunk_begin = nchunks * split / nsplit, chunk_end = nchunks * (split + 1) / nsplit;
  const int64_t my_chunks = chunk_end - chunk_begin;
  int* ip = ipart + (int64_t)blockIdx.x * DD * UT_N;
  unsigned long long* tp = tpart + (int64_t)blockIdx.x * (2 * UT_NC);

  // operand buffers: zero everything (padding rows of B stay zero), count row of B = 1
  for (int e = tid; e < UT_NB * C::BUF_BYTES / 16; e += UT_THREADS) reinterpret_cast<uint4*>(base)[e] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  for (int e = tid; e < UT_NB * 32; e += UT_THREADS) {
    const int b = e >> 5, kw = e & 31;
    *reinterpret_cast<uint32_t*>(base + b * C::BUF_BYTES + sw128(UT_CNT, 4 * kw)) = 0x01010101u;
  }
  if (tid == 0) {
    for (int b = 0; b < UT_NB; ++b) {
      mbar_init(bar_u + 8 * b, UT_PROD);
      mbar_init(bar_u + 8 * (UT_NB + b), 1);
    }
    mbar_init(bar_u + 16 * UT_NB, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (w == UT_PW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;

  if (w == UT_PW) {
    // ---------------------------------------------------------------- MMA issuer -------------------------------------
    if (lane == 0) {
      // D = S32, A signed 8 bit, B unsigned 8 bit, both K-major, N = 144, M = 128
      constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (0u << 10) | ((uint32_t)(UT_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      bool ok = true;
      for (int64_t i = 0; i < my_chunks && ok; ++i) {
        const int b = (int)(i % UT_NB);
        ok = mbar_wait_bounded(bar_u + 8 * b, (uint32_t)((i / UT_NB) & 1));
        if (!ok) break;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t b_u = base_u + b * C::BUF_BYTES;
        const uint32_t a_t = tmem + (uint32_t)(C::A_BASE + b * C::A_COLS);
        const uint64_t bdesc = smem_desc_sw128(b_u);
#pragma unroll
        for (int mt = 0; mt < C::MT; ++mt) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {   // K = 32 per instruction: + 32 B inside the swizzle row = + 2 descriptor units
            const uint32_t acc = (i | k) ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem + (uint32_t)(mt * UT_N)),
                "r"(a_t + (uint32_t)(mt * 32 + 8 * k)), "l"(bdesc + 2 * k), "r"(idesc), "r"(acc), "r"(0u)
                : "memory");
          }
        }
        // arrives on empty[b] once these MMAs have finished reading the buffer (implies fence::before_thread_sync)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_u + 8 * (UT_NB + b)) : "memory");
      }
      if (!ok) set_err(scal, 3.0);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_u + 16 * UT_NB) : "memory");
    }
  } else {
    // ---------------------------------------------------------------- producers --------------------------------------
    // Cutting task j of warp w: id = w + 8 j < 20 -> column quad cq = id >> 2, quad block kb = id & 3 (observation quads
    // 8 kb + (lane >> 2)).  A column quad takes every second column (cq < 4: 8 (cq >> 1) + (cq & 1) + 2 (lane & 3); cq = 4:
    // 16 + (lane & 3)): rows 7 c + s of columns two apart differ by 6 mod 8, so the 4 x 8 plane words of a store
    // instruction fall into 32 different banks of the swizzled tile.
    int tcol[UT_TASKS], tkq[UT_TASKS];
    bool tval[UT_TASKS];
    double tsc[UT_TASKS];
    uint32_t tthr[UT_TASKS];
    const uint32_t hA = eamax[0], hE = eamax[1];
#pragma unroll
    for (int j = 0; j < UT_TASKS; ++j) {
      const int id = w + UT_PW * j, cq = id >> 2;
      tcol[j] = (cq < 4) ? 8 * (cq >> 1) + (cq & 1) + 2 * (lane & 3) : 16 + (lane & 3);
      tkq[j] = 8 * (id & 3) + (lane >> 2);
      tval[j] = id < 20 && tcol[j] < UT_NC;
      const bool isA = col0 + tcol[j] < IM_NV;
      tsc[j] = scale_up(isA ? hA : hE);
      tthr[j] = (uint32_t)(class_exponent(isA ? hA : hE) - IM_GUARD_BITS) << 20;
    }
    unsigned long long tlo[UT_TASKS], thi[UT_TASKS];
    uint32_t small[UT_TASKS];
#pragma unroll
    for (int j = 0; j < UT_TASKS; ++j) tlo[j] = thi[j] = 0ull, small[j] = 0u;

    // Staging: the 128 x 19 entries of a chunk are copied element-wise (8-byte cp.async) in flattened order, thread t
    // taking elements t, t + 256, ...: a warp instruction reads 32 consecutive entries (at most 3 rows of the column
    // group) and writes 256 contiguous bytes.  Mask words: thread t copies the 16 bytes of feature t.
    const int er0 = tid / UT_NC, ec0 = tid % UT_NC;   // element tid = (row er0, column ec0); + UT_PROD = (+ UT_PROD / 19 rows, + UT_PROD % 19 columns)
    const double* pe = EA + chunk_begin * IM_CHUNK * IM_EAS + col0;
    const uint32_t* pm = mT + (chunk_begin * DD + (tid % DD)) * 4;
    int pf_stage = 0;
    int64_t pf_chunk = chunk_begin;
    auto prefetch = [&]() {
      if (pf_chunk < chunk_end) {
        const uint32_t st = stage_u + (uint32_t)pf_stage * UT_STAGE_BYTES;
        int er = er0, ec = ec0;
#pragma unroll
        for (int m = 0; m < (IM_CHUNK * UT_NC + UT_PROD - 1) / UT_PROD; ++m) {
          if (tid + UT_PROD * m < IM_CHUNK * UT_NC)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(st + (uint32_t)((tid + UT_PROD * m) * 8)),
                         "l"(pe + er * IM_EAS + ec)
                         : "memory");
          er += UT_PROD / UT_NC;
          ec += UT_PROD % UT_NC;
          if (ec >= UT_NC) ec -= UT_NC, ++er;
        }
        if (tid < DD)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(mask_u + (uint32_t)(pf_stage * DD * 16 + tid * 16)), "l"(pm) : "memory");
        pe += IM_CHUNK * IM_EAS;
        pm += DD * 4;
        pf_stage = (pf_stage + 1 == UT_PF) ? 0 : pf_stage + 1;
        ++pf_chunk;
      }
      asm volatile("cp.async.commit_group;" ::: "memory");   // one group per chunk, empty past the end
    };
    prefetch();
    prefetch();
    int cs = 0;
    bool ok = true;
    for (int64_t i = 0; i < my_chunks; ++i) {
      const int b = (int)(i % UT_NB);
      const int64_t chunk = chunk_begin + i;
      asm volatile("cp.async.wait_group 1;" ::: "memory");   // this thread's copies of `chunk` have landed
      asm volatile("bar.sync 1, %0;" ::"n"(UT_PROD) : "memory");         // ... and everybody's; all producers are done cutting chunk - 1,
      prefetch();                                            // whose stage is the one this prefetch (chunk + 2) overwrites
      if (i >= UT_NB && ok) ok = mbar_wait_bounded(bar_u + 8 * (UT_NB + b), (uint32_t)(((i / UT_NB) - 1) & 1));   // MMAs of chunk i - NB released the buffer
      const uint32_t b_u = base_u + b * C::BUF_BYTES;
      const uint32_t st = stage_u + (uint32_t)cs * UT_STAGE_BYTES;
      // ---- mask: feature tid, 128 observations: stored word wq, bit 8 b + s  <=>  observation 32 wq + 4 s + b
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");   // the MMAs that read A buffer b have completed (empty[b])
      if (tid < DD) {
        uint32_t mw[4];
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(mw[0]), "=r"(mw[1]), "=r"(mw[2]), "=r"(mw[3])
                     : "r"(mask_u + (uint32_t)(cs * DD * 16 + tid * 16)));
        // 16 observations = 4 output words per u: word 4 u + jj = observations 16 u + 4 jj .. + 3 -> TMEM column 4 u + jj
        uint32_t o[32];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          o[4 * u + 0] = (mw[u >> 1] >> (4 * (u & 1) + 0)) & 0x01010101u;
          o[4 * u + 1] = (mw[u >> 1] >> (4 * (u & 1) + 1)) & 0x01010101u;
          o[4 * u + 2] = (mw[u >> 1] >> (4 * (u & 1) + 2)) & 0x01010101u;
          o[4 * u + 3] = (mw[u >> 1] >> (4 * (u & 1) + 3)) & 0x01010101u;
        }
        // warp w holds features 32 w .. 32 w + 31: TMEM lanes 32 (w & 3) .. of feature tile w >> 2
        const uint32_t taddr = tmem + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)(C::A_BASE + b * C::A_COLS + (w >> 2) * 32);
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
            "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
            "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]), "r"(o[8]), "r"(o[9]), "r"(o[10]),
            "r"(o[11]), "r"(o[12]), "r"(o[13]), "r"(o[14]), "r"(o[15]), "r"(o[16]), "r"(o[17]), "r"(o[18]), "r"(o[19]), "r"(o[20]),
            "r"(o[21]), "r"(o[22]), "r"(o[23]), "r"(o[24]), "r"(o[25]), "r"(o[26]), "r"(o[27]), "r"(o[28]), "r"(o[29]), "r"(o[30]),
            "r"(o[31])
            : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      // ---- entries: 4 consecutive observations of one column -> 7 plane words
#pragma unroll
      for (int j = 0; j < UT_TASKS; ++j) {
        if (tval[j]) {
          const int nvalid = (int)min((int64_t)4, max((int64_t)0, n_obs - (chunk * IM_CHUNK + 4 * tkq[j])));
          uint32_t lo[4], hi[4];
          double v0 = 0.0;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const double v = lds64(st + (uint32_t)(((4 * tkq[j] + r) * UT_NC + tcol[j]) * 8));
            if (r == 0) v0 = v;
            const double y = fma(v, tsc[j], 6755399441055744.0);   // 2^52 + 2^51: the mantissa of y is the integer x + 2^51
            lo[r] = (uint32_t)__double2loint(y);
            hi[r] = (uint32_t)__double2hiint(y) - 0x43300000u;
          }
          small[j] += (((uint32_t)__double2hiint(v0) & 0x7fffffffu) < tthr[j]) ? 1u : 0u;   // precision guard (sampled)
          if (nvalid == 4) {
            tlo[j] += ((unsigned long long)lo[0] + lo[1]) + ((unsigned long long)lo[2] + lo[3]);
            thi[j] += (hi[0] + hi[1]) + (hi[2] + hi[3]);
          } else {
#pragma unroll
            for (int r = 0; r < 4; ++r)
              if (r < nvalid) {
                tlo[j] += lo[r];
                thi[j] += hi[r];
              }
          }
#pragma unroll
          for (int s = 0; s < NSL; ++s) {
            const uint32_t sel = (uint32_t)((s & 3) | (((s & 3) + 4) << 4));
            const uint32_t w01 = (s < 4) ? __byte_perm(lo[0], lo[1], sel) : __byte_perm(hi[0], hi[1], sel);
            const uint32_t w23 = (s < 4) ? __byte_perm(lo[2], lo[3], sel) : __byte_perm(hi[2], hi[3], sel);
            const uint32_t wv = __byte_perm(w01, w23, 0x5410u);
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(b_u + sw128(tcol[j] * NSL + s, 4 * tkq[j])), "r"(wv) : "memory");
          }
        }
      }
      cs = (cs + 1 == UT_PF) ? 0 : cs + 1;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  // ... and the tensor-memory stores of the A operand
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_u + 8 * b) : "memory");
    }
    if (!ok) set_err(scal, 3.0);
    // ---- precision-guard counters and totals (integer sums: order-independent) -----------------------------------------
    uint32_t sA = 0u, sE = 0u;
#pragma unroll
    for (int j = 0; j < UT_TASKS; ++j) {
      const bool isA = col0 + tcol[j] < IM_NV;
      sA += (tval[j] && isA) ? small[j] : 0u;
      sE += (tval[j] && !isA) ? small[j] : 0u;
    }
    sA = __reduce_add_sync(0xffffffffu, sA);
    sE = __reduce_add_sync(0xffffffffu, sE);
    if (lane == 0) {
      if (sA) atomicAdd(guard_counts(eamax), (unsigned long long)sA);
      if (sE) atomicAdd(guard_counts(eamax) + 1, (unsigned long long)sE);
    }
#pragma unroll
    for (int j = 0; j < UT_TASKS; ++j) {
#pragma unroll
      for (int o = 4; o <= 16; o <<= 1) {   // lanes with equal (lane & 3) share the column
        tlo[j] += __shfl_xor_sync(0xffffffffu, tlo[j], o);
        thi[j] += __shfl_xor_sync(0xffffffffu, thi[j], o);
      }
      if (lane < 4) {
        tred[((w * UT_TASKS + j) * 4 + lane) * 2] = tval[j] ? tlo[j] : 0ull;
        tred[((w * UT_TASKS + j) * 4 + lane) * 2 + 1] = tval[j] ? thi[j] : 0ull;
      }
    }
  }
  // ------------------------------------------------------------------ epilogue ---------------------------------------
  __syncthreads();
  if (tid < 2 * UT_NC) {   // column c lives in column quad cq, lane slot cl; tasks id = 4 cq + kb -> (warp id % UT_PW, slot id / UT_PW)
    const int c = tid >> 1, h = tid & 1;
    const int cq = (c < 16) ? 2 * (c >> 3) + (c & 1) : 4, cl = (c < 16) ? (c & 7) >> 1 : c - 16;
    unsigned long long a = 0ull;
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
      const int id = 4 * cq + kb;
      a += tred[(((id % UT_PW) * UT_TASKS + (id / UT_PW)) * 4 + cl) * 2 + h];
    }
    tp[tid] = a;
  }
  if (w < 4) {
    const bool ok = mbar_wait_bounded(bar_u + 16 * UT_NB, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok) {
      if (lane == 0) set_err(scal, 3.0);
    } else {
#pragma unroll 1
      for (int mt = 0; mt < C::MT; ++mt) {
        int* out = ip + (int64_t)(mt * 128 + 32 * w + lane) * UT_N;
#pragma unroll 1
        for (int c0 = 0; c0 < UT_N; c0 += 16) {
          uint32_t r[16];
          const uint32_t taddr = tmem + ((uint32_t)(32 * w) << 16) + (uint32_t)(mt * UT_N + c0);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
              : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const bool live = my_chunks > 0;   // no MMA was issued for an empty split: the accumulator is undefined
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<int4*>(out + c0 + 4 * q) = live ? make_int4((int)r[4 * q], (int)r[4 * q + 1], (int)r[4 * q + 2], (int)r[4 * q + 3])
                                                              : make_int4(0, 0, 0, 0);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (w == UT_PW) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(C::TMEM_COLS) : "memory");
}

// dst[f][c] (f < D: complement sums of feature f; f = D: totals over all observations), each rounded to FP64 once
__global__ void __launch_bounds__(256) masked_miss_utc_finalize(const int* __restrict__ ipart, const unsigned long long* __restrict__ tpart,
                                                                int D, int nsplit, int64_t n_obs, uint32_t* __restrict__ eamax,
                                                                double* __restrict__ dst, const double* __restrict__ scal) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const bool done = scal[S_DONE] != 0.0;
  const bool fallback = !done && (guard_fallback(eamax, n_obs) || scal[S_ERR] == 3.0);
  if (idx == 0) guard_scal(eamax)[S_DONE] = fallback ? 0.0 : 1.0;   // 0.0: the FP64 kernels that follow redo the sums
  if (done || fallback || idx >= (D + 1) * IM_EAS) return;
  const int f = idx / IM_EAS, c = idx % IM_EAS;
  const int group = c / UT_NC, lc = c % UT_NC;
  __int128 T = 0;
  if (f < D) {
    long long S[NSL], cnt = 0;
#pragma unroll
    for (int s = 0; s < NSL; ++s) S[s] = 0;
    for (int sp = 0; sp < nsplit; ++sp) {
      const int* row = ipart + ((int64_t)(group * nsplit + sp) * D + f) * UT_N;
#pragma unroll
      for (int s = 0; s < NSL; ++s) S[s] += row[lc * NSL + s];
      cnt += row[UT_CNT];
    }
#pragma unroll
    for (int s = 0; s < NSL; ++s) T += (__int128)S[s] << (8 * s);
    T -= (__int128)cnt << FXP;
  } else {
    unsigned long long lo = 0ull, hi = 0ull;
    for (int sp = 0; sp < nsplit; ++sp) {
      lo += tpart[(int64_t)(group * nsplit + sp) * (2 * UT_NC) + lc * 2];
      hi += tpart[(int64_t)(group * nsplit + sp) * (2 * UT_NC) + lc * 2 + 1];
    }
    T = ((__int128)hi << 32) + (__int128)lo - ((__int128)n_obs << FXP);
  }
  dst[idx] = int128_to_double(T) * scale_down((c < IM_NV) ? eamax[0] : eamax[1]);
}

// Transposed copy of the mask bits for the A-tile expansion (constant over the EM iterations): one contiguous D x 16 B
// block per 128-observation chunk; word wq (0..3) of feature f, bit 8 b + s  <=>  observation 32 wq + 4 s + b of the
// chunk, so that (W >> s) & 0x01010101 is the 4 mask BYTES of observations 32 wq + 4 s .. + 3.
__global__ void __launch_bounds__(256) masked_bits_transpose_utc_kernel(const uint32_t* __restrict__ mbits, int wpr, int64_t n_obs,
                                                                        uint32_t* __restrict__ mT, int64_t nchunks) {
  const int lane = threadIdx.x & 31;
  const int64_t wi = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);   // warp = (chunk, word of 32 features)
  if (wi >= nchunks * wpr) return;
  const int64_t chunk = wi / wpr;
  const int wd = (int)(wi % wpr);
  uint32_t out[4];
#pragma unroll
  for (int wq = 0; wq < 4; ++wq) {
    const int64_t n = chunk * IM_CHUNK + wq * 32 + lane;
    const uint32_t wv = (n < n_obs) ? mbits[n * wpr + wd] : 0u;
    uint32_t keep = 0u;   // lane b: bit i = observation 32 wq + i misses feature 32 wd + b
#pragma unroll
    for (int b = 0; b < 32; ++b) {
      const uint32_t m = __ballot_sync(0xffffffffu, (wv >> b) & 1u);
      if (lane == b) keep = m;
    }
    uint32_t o = 0u;
#pragma unroll
    for (int s = 0; s < 8; ++s)
#pragma unroll
      for (int b = 0; b < 4; ++b) o |= ((keep >> (4 * s + b)) & 1u) << (8 * b + s);
    out[wq] = o;
  }
  *reinterpret_cast<uint4*>(&mT[((chunk * wpr + wd) * 32 + lane) * 4]) = make_uint4(out[0], out[1], out[2], out[3]);
}

template <int DD>
int launch_utc(ShardState& sh, double* dst) {
  using C = UtcCfg<DD>;
  const int64_t nchunks = ceil_div(sh.n, (int64_t)IM_CHUNK);
  const int nsplit = (int)std::max<int64_t>(1, std::min<int64_t>(sm_count(sh.dev) / UT_GROUPS, nchunks));
  const int ncta = UT_GROUPS * nsplit;
  int* ipart = reinterpret_cast<int*>(sh.partials);
  unsigned long long* tpart = reinterpret_cast<unsigned long long*>(ipart + (int64_t)ncta * DD * UT_N);
  static bool attr_done[64] = {};
  if (!attr_done[sh.dev & 63]) {
    CUDA_TRY(cudaFuncSetAttribute((masked_miss_utc_kernel<DD>), cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_done[sh.dev & 63] = true;
  }
  GP_LAUNCH((masked_miss_utc_kernel<DD>), ncta, UT_THREADS, C::SMEM, sh.st, sh.EA, sh.mbitsT, sh.n, nsplit, sh.eamax, ipart, tpart,
            sh.par.scal);
  GP_LAUNCH(masked_miss_utc_finalize, (unsigned)ceil_div((int64_t)(DD + 1) * IM_EAS, 256), 256, 0, sh.st, ipart, tpart, DD, nsplit, sh.n,
            sh.eamax, dst, sh.par.scal);
  return GPLVM_OK;
}

}  // namespace

// transposed mask bits in the layout of the tcgen05 kernel (once per data set)
int masked_utc_prepare(gplvm_ppca_session* s, ShardState& sh) {
  CUDA_TRY(cudaMemsetAsync(sh.eamax, 0, 256, sh.st));
  const int64_t nchunks = sh.nwT / 4;
  GP_LAUNCH(masked_bits_transpose_utc_kernel, (unsigned)ceil_div(nchunks * s->WPR, (int64_t)8), 256, 0, sh.st, sh.mbits, s->WPR, sh.n,
            sh.mbitsT, nchunks);
  return GPLVM_OK;
}

// complement sums [G^miss | S1^miss] (D x 152) + totals (152) into dst (length (D + 1) x 152)
int masked_utc_miss_sums(gplvm_ppca_session* s, ShardState& sh, double* dst) {
  switch (s->D) {
    case 128: return launch_utc<128>(sh, dst);
    case 256: return launch_utc<256>(sh, dst);
    default: return fail(GPLVM_ERR_STATE, "INT8 masked statistics: unsupported D");
  }
}

}  // namespace gplvm
