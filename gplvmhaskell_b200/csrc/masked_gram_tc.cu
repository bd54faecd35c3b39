// masked_gram_tc.cu -- per-observation solver of the NaN-masked E-step (M = 16, D = 256) whose Gram matrices come from the
// 5th-generation tensor cores instead of the FP64 DMMA product of masked_fused.cu.
//
// Reference: emStepsMissed.stepOne, src/GPLVM/PPCA.hs:106-117:  M_i = sigma^2 I + W_i^T W_i  (W_i = rows of W at the
// features observation i has), invMP = M_i^-1, expXi = M_i^-1 W_i^T (y_i - mu).  Complement form:
//   M_i = C0 - G_i,   C0 = sigma^2 I + W^T W,   G_i[a][c] = sum_d miss[i][d] W[d][a] W[d][c]
// and G (N x 136 unique entries) = Miss (N x D, 0 / 1) . P (D x 136),  P[d][(a,c)] = W[d][a] W[d][c]  is a GEMM with an
// exactly representable 0 / 1 factor, so only P needs splitting (the convention of masked_fx.cuh):
//   x = rint(P[d][(a,c)] 2^(51 - ea - ec)),  2^ea > max_d |W[d][a]|  (one exponent per latent column, i.e. the symmetric
//   diagonal scaling S M S with S = diag(2^-ea), exact in binary floating point),  u = x + 2^51 in [0, 2^52] cut into 7
//   bytes;  tcgen05.mma kind::i8 accumulates mask x byte products exactly in s32 (TMEM);  sum_d x = sum_s 256^s S_s -
//   2^51 cnt_i  is an exact 64-bit integer, converted to FP64 with ONE rounding and scaled by 2^(ea + ec - 51).
// The only other rounding is the rint of each product to 2^-52 of its column-pair scale -- the size of the rounding of a
// single FP64 addition of that product into a running sum (what the DMMA product does 51 times per entry).
// Precision guard: if more than 1 % of the non-zero entries of W lie 2^20 below their column maximum the prep kernel
// clears the gate and the FP64 DMMA solver of masked_fused.cu takes the step (launched right after, gated the other way).
//
// Kernel: persistent, one CTA per SM, tiles of 128 observations.
//   warp 8 lane 0  : streams the 9 pre-swizzled B chunks (16 entries x 7 planes = 112 rows, K = 256; 28 KB each, written
//                    by gram_planes_kernel in the canonical SWIZZLE_128B K-major image) through 3 shared-memory buffers with
//                    cp.async.bulk, and issues 8 MMAs (M = 128, N = 112, K = 32; A from tensor memory) per chunk into one of
//                    4 TMEM accumulator buffers.
//   warps 0..7     : (a) expand the mask words of the NEXT tile into the A operand IN TENSOR MEMORY (tcgen05.st: lane =
//                        observation, 4 features per 32-bit column; 64 columns behind the accumulators) -- the MMAs read only B
//                        from shared memory, and the 32 KB an A tile would take pay for the third B buffer;
//                    (b) per chunk: tcgen05.ld their TMEM lane quadrant (lane = observation), combine the planes, and
//                        store M_i entries into the packed lower-triangular staging rows (pitch 137: conflict-free);
//                        warps w and w + 4 share a quadrant and split the chunk's entries;
//                    (c) two rounds of 8 observations per warp: load the sweep layout (4 lanes per observation), invert
//                        by symmetric sweeps in registers (masked_sweeps.cuh), ez = M_i^-1 b_i,
//                        A_i = ez ez^T + sigma^2 M_i^-1, rows [vech A | ez] out through the same staging rows.
// Every mbarrier wait is bounded (tma.cuh: trap instead of a hang).
#include "masked_fx.cuh"
#include "masked_sweeps.cuh"
#include "ppca.cuh"
#include "tma.cuh"

namespace gplvm {
namespace {

constexpr int GT_M = 16;
constexpr int GT_NV = GT_M * (GT_M + 1) / 2;       // 136 unique entries
constexpr int GT_EAS = GT_NV + GT_M;               // 152
constexpr int GT_NSL = 7;                          // byte planes
constexpr int GT_FXP = 51;
constexpr int GT_CE = 16;                          // entries per chunk
constexpr int GT_N = GT_CE * GT_NSL;               // MMA N = 112
constexpr int GT_CHUNKS = (GT_NV + GT_CE - 1) / GT_CE;   // 9 (the last one holds 8 entries)
constexpr int GT_D = 256;
constexpr int GT_KB = GT_D / 128;                  // K blocks of 128 features (one swizzle row each)
constexpr int GT_KB_BYTES = GT_N * 128;            // 14 336
constexpr int GT_B_BYTES = GT_KB * GT_KB_BYTES;    // 28 672 per chunk
constexpr int GT_NBUF = 3;                         // B chunk buffers in shared memory
constexpr int GT_TILE = 128;                       // observations per tile
constexpr int GT_PITCH = 137;                      // doubles per staging row (136 entries; odd pitch: conflict-free transposed stores)
constexpr int GT_STAGE_BYTES = GT_TILE * GT_PITCH * 8;
constexpr int GT_TBUF = 4, GT_TCOLS = GT_N;        // TMEM accumulator buffers x columns (4 x 112 = 448)
constexpr int GT_A_COL = GT_TBUF * GT_TCOLS;       // the A operand (mask bytes) lives in TMEM columns [448, 512): lane = observation,
constexpr int GT_A_COLS = GT_D / 4;                //   32-bit column q = features 4 q .. 4 q + 3 (tools/tcgen05_i8_tmem_a_check.cu)
static_assert(GT_A_COL + GT_A_COLS <= 512, "tensor memory budget");
constexpr int GT_SOLVERS = 8;                      // solver warps
constexpr int GT_THREADS = (GT_SOLVERS + 4) * 32;   // 3 warpgroups: two of solver warps, one whose first warp issues the MMAs
// register budget (setmaxnreg, per warpgroup): the kernel starts at 168 registers per thread (64 512 for 384 threads);
// the solver warpgroups grow to 232, the issuer warpgroup shrinks to 40:  8 x 32 x 232 + 4 x 32 x 40 = 64 512
constexpr int GT_REGS_SOLVER = 232, GT_REGS_ISSUER = 40;
constexpr int GT_OFF_B = 0, GT_OFF_STAGE = GT_OFF_B + GT_NBUF * GT_B_BYTES, GT_OFF_BAR = GT_OFF_STAGE + GT_STAGE_BYTES;
constexpr int GT_SMEM = GT_OFF_BAR + 128;
// tab block (doubles): [0, 136) scale 2^(ea + ec - 51), [136, 272) C0 entries, [272] gate (1.0: this kernel takes the step)
constexpr int GT_TAB_C0 = GT_NV, GT_TAB_GATE = 2 * GT_NV, GT_TAB_DOUBLES = 2 * GT_NV + 8;
static_assert(GT_SMEM <= 232448, "shared memory budget of one CTA per SM");
static_assert(GT_KB_BYTES % 1024 == 0 && GT_B_BYTES % 1024 == 0, "operand tiles need 1024-byte alignment");

// byte (row r, k-byte b) of a K-major operand tile: 128-byte rows, 8-row groups of 1024 B, 16-byte units XOR r & 7
__host__ __device__ __forceinline__ uint32_t sw128(int r, int b) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((b >> 4) ^ (r & 7))) << 4) + (b & 15));
}
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
  // start address >> 4 | LBO (unused for swizzled K-major) | SBO = 1024 B >> 4 | descriptor version 1 | SWIZZLE_128B
  return (uint64_t)((addr & 0x3ffff) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

#define GT_LD8(r, o, addr)                                                                                                    \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"                                \
               : "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), \
                 "=r"(r[o + 7])                                                                                               \
               : "r"(addr))
#define GT_LD16(r, o, addr)                                                                                                   \
  asm volatile(                                                                                                               \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" \
      : "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]),        \
        "=r"(r[o + 7]), "=r"(r[o + 8]), "=r"(r[o + 9]), "=r"(r[o + 10]), "=r"(r[o + 11]), "=r"(r[o + 12]), "=r"(r[o + 13]),    \
        "=r"(r[o + 14]), "=r"(r[o + 15])                                                                                      \
      : "r"(addr))

// ------------------------------------------------------------------------------------------------------------------
// B operand image + tables, once per EM step.  grid = (9 chunks, 2 K blocks), 128 threads = the features of the K block.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) gram_planes_kernel(Par p, uint8_t* __restrict__ Bimg, double* __restrict__ tab) {
  __shared__ double wmax[4][GT_M];
  __shared__ int eexp[GT_M];
  __shared__ unsigned int small_cnt, nonzero_cnt;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int chunk = blockIdx.x, kb = blockIdx.y;
  if (tid == 0) small_cnt = nonzero_cnt = 0u;
  // column maxima over all D rows (every CTA recomputes them: 4096 loads)
  double wa[2][GT_M];
#pragma unroll
  for (int a = 0; a < GT_M; ++a) {
    wa[0][a] = p.W[(int64_t)tid * GT_M + a];
    wa[1][a] = p.W[(int64_t)(tid + 128) * GT_M + a];
    double v = fmax(fabs(wa[0][a]), fabs(wa[1][a]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) wmax[w][a] = v;
  }
  __syncthreads();
  if (tid < GT_M) {
    const double v = fmax(fmax(wmax[0][tid], wmax[1][tid]), fmax(wmax[2][tid], wmax[3][tid]));
    // 2^e > v ;  non-finite or zero columns get e = 0 (a non-finite W fails the pivots of every observation anyway)
    int e = 0;
    if (v > 0.0 && v < 1.7e308) e = ilogb(v) + 1;
    eexp[tid] = e;
  }
  __syncthreads();
  const double* wd = wa[kb];   // this thread's feature d = 128 kb + tid
  uint8_t* img = Bimg + (size_t)chunk * GT_B_BYTES + (size_t)kb * GT_KB_BYTES;
  for (int i = 0; i < GT_CE; ++i) {
    const int e = chunk * GT_CE + i;
    if (e >= GT_NV) break;
    int a = 0;
    while ((a + 1) * (a + 2) / 2 <= e) ++a;
    const int c = e - a * (a + 1) / 2;
    double wva = 0.0, wvc = 0.0;
#pragma unroll
    for (int q = 0; q < GT_M; ++q) {   // register arrays: select instead of a dynamic index
      wva = (q == a) ? wd[q] : wva;
      wvc = (q == c) ? wd[q] : wvc;
    }
    const double prod = wva * wvc;
    const double y = fma(prod, ldexp(1.0, GT_FXP - eexp[a] - eexp[c]), 6755399441055744.0);   // 2^52 + 2^51: mantissa = x + 2^51
    const uint32_t lo = (uint32_t)__double2loint(y), hi = (uint32_t)__double2hiint(y) - 0x43300000u;
#pragma unroll
    for (int s = 0; s < GT_NSL; ++s) img[sw128(GT_NSL * i + s, tid)] = (uint8_t)(((s < 4) ? (lo >> (8 * s)) : (hi >> (8 * (s - 4)))) & 0xffu);
  }
  if (chunk == 0 && kb == 0) {
    // precision guard over all of W, tables, gate
    unsigned int sm = 0u, nz = 0u;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int a = 0; a < GT_M; ++a) {
        const double v = fabs(wa[h][a]);
        if (v > 0.0) {
          ++nz;
          if (v < ldexp(1.0, eexp[a] - 1 - fx::IM_GUARD_BITS)) ++sm;
        }
      }
    atomicAdd(&small_cnt, sm);
    atomicAdd(&nonzero_cnt, nz);
    for (int e = tid; e < GT_NV; e += 128) {
      int a = 0;
      while ((a + 1) * (a + 2) / 2 <= e) ++a;
      const int c = e - a * (a + 1) / 2;
      tab[e] = ldexp(1.0, eexp[a] + eexp[c] - GT_FXP);
      tab[GT_TAB_C0 + e] = p.Minv[a * GT_M + c];
    }
    __syncthreads();
    if (tid == 0) {
      bool range_ok = true;   // 2^(51 - ea - ec) and 2^(ea + ec - 51) must be normal numbers
      for (int a = 0; a < GT_M; ++a) range_ok = range_ok && eexp[a] > -400 && eexp[a] < 400;
      tab[GT_TAB_GATE] = (!range_ok || small_cnt * 100u > nonzero_cnt) ? 0.0 : 1.0;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// the solver
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GT_THREADS, 1)
masked_solve_tc_kernel(const double* __restrict__ bsrc, const uint32_t* __restrict__ mbits, int64_t n_obs, Par p, double* __restrict__ EA,
                       uint32_t* __restrict__ eamax, const uint8_t* __restrict__ Bimg, const double* __restrict__ tab) {
  constexpr int WPRT = GT_D / 32;
  constexpr unsigned FULL = 0xffffffffu;
  if (p.scal[S_DONE] != 0.0) return;
  if (tab[GT_TAB_GATE] != 1.0) return;   // precision guard: the FP64 solver takes this step
  extern __shared__ __align__(1024) uint8_t gt_smem[];
  const uint32_t base_u = smem_u32(gt_smem);
  if (base_u & 1023u) {   // the swizzled operand tiles need 1024-byte alignment (uniform: every thread leaves)
    if (threadIdx.x == 0) set_err(p.scal, 3.0);
    return;
  }
  const uint32_t b_u = base_u + GT_OFF_B, bar_u = base_u + GT_OFF_BAR;
  double* stage = reinterpret_cast<double*>(gt_smem + GT_OFF_STAGE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gt_smem + GT_OFF_BAR + 120);
  // barriers: b_full[3] @0, b_empty[3] @24, t_full[4] @48, t_empty[4] @80, a_full @112
  const uint32_t b_full = bar_u, b_empty = bar_u + 24, t_full = bar_u + 48, t_empty = bar_u + 80, a_full = bar_u + 112;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int64_t ntiles = (n_obs + GT_TILE - 1) / GT_TILE;
  const int64_t iters = (ntiles > (int64_t)blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (tid == 0) {
    for (int i = 0; i < GT_NBUF; ++i) mbar_init(b_full + 8 * i, 1), mbar_init(b_empty + 8 * i, 1);
    for (int i = 0; i < GT_TBUF; ++i) mbar_init(t_full + 8 * i, 1), mbar_init(t_empty + 8 * i, GT_SOLVERS);
    mbar_init(a_full, GT_SOLVERS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (w == GT_SOLVERS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (w >= GT_SOLVERS) {
    // ================================================================ B stream + MMA issue ==============================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GT_REGS_ISSUER));
    if (w == GT_SOLVERS && lane == 0) {
      constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (0u << 10) | ((uint32_t)(GT_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const int64_t total = iters * GT_CHUNKS;
      auto load = [&](int64_t i) {
        const int buf = (int)(i % GT_NBUF), c = (int)(i % GT_CHUNKS);
        mbar_expect_tx(b_full + 8 * buf, (uint32_t)GT_B_BYTES);
        bulk_load_1d(b_u + buf * GT_B_BYTES, Bimg + (size_t)c * GT_B_BYTES, (uint32_t)GT_B_BYTES, b_full + 8 * buf);
      };
      for (int64_t i = 0; i < GT_NBUF - 1 && i < total; ++i) load(i);
      for (int64_t i = 0; i < total; ++i) {
        const int buf = (int)(i % GT_NBUF), tb = (int)(i & 3), c = (int)(i % GT_CHUNKS);
        const int64_t nx = i + GT_NBUF - 1;   // keep NBUF - 1 chunk loads in flight behind the one being multiplied
        if (nx < total) {
          if (nx >= GT_NBUF) mbar_wait(b_empty + 8 * (int)(nx % GT_NBUF), (uint32_t)(((nx / GT_NBUF) - 1) & 1));
          load(nx);
        }
        mbar_wait(b_full + 8 * buf, (uint32_t)((i / GT_NBUF) & 1));
        if (i >= GT_TBUF) mbar_wait(t_empty + 8 * tb, (uint32_t)(((i >> 2) - 1) & 1));
        if (c == 0) mbar_wait(a_full, (uint32_t)((i / GT_CHUNKS) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int kb = 0; kb < GT_KB; ++kb) {
          const uint64_t bdesc = smem_desc_sw128(b_u + buf * GT_B_BYTES + kb * GT_KB_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {   // K = 32 per instruction: + 32 B inside the swizzle row = + 2 descriptor units
            const uint32_t acc = (kb | k) ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem + (uint32_t)(tb * GT_TCOLS)),
                "r"(tmem + (uint32_t)(GT_A_COL + kb * 32 + 8 * k)), "l"(bdesc + 2 * k), "r"(idesc), "r"(acc), "r"(0u)
                : "memory");
          }
        }
        umma_commit(b_empty + 8 * buf);   // the B buffer may be refilled once these MMAs have read it
        umma_commit(t_full + 8 * tb);     // ... and the accumulator is complete
      }
    }
  } else {
    // ================================================================ solver warps =====================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(GT_REGS_SOLVER));
    const int q = w & 3, h = w >> 2;           // TMEM lane quadrant, half of the chunk's entries / of the quadrant's observations
    const int rw = lane >> 2, j = lane & 3;    // sweep coordinates: observation within the round, lane within its quad
    double* Sq = stage + (size_t)(32 * q) * GT_PITCH;
    const double s2 = p.scal[S_SIGMA2];
    uint32_t mxa = 0u, mxe = 0u;
    // Mask words of observation `tile * 128 + 32 q + lane` (the lane's conversion-role observation): their population count
    // is the missing-feature count, and this warp's half of them (features 128 h .. 128 h + 127) becomes the A operand bytes
    // (1 = missing) of TMEM lane 32 q + lane, columns GT_A_COL + 32 h .. + 31.  Fetched / written one tile ahead.
    auto build_A = [&](int64_t tile) -> int {
      const int64_t n = tile * GT_TILE + 32 * q + lane;
      uint4 m0 = make_uint4(0u, 0u, 0u, 0u), m1 = m0;
      if (n < n_obs) {
        const uint4* mp = reinterpret_cast<const uint4*>(mbits + n * WPRT);
        m0 = mp[0];
        m1 = mp[1];
      }
      const int c = __popc(m0.x) + __popc(m0.y) + __popc(m0.z) + __popc(m0.w) + __popc(m1.x) + __popc(m1.y) + __popc(m1.z) + __popc(m1.w);
      const uint4 mh = h ? m1 : m0;
      const uint32_t wds[4] = {mh.x, mh.y, mh.z, mh.w};
      uint32_t o[32];
#pragma unroll
      for (int wl = 0; wl < 4; ++wl)
#pragma unroll
        for (int gq = 0; gq < 8; ++gq) o[8 * wl + gq] = (((wds[wl] >> (4 * gq)) & 0xfu) * 0x00204081u) & 0x01010101u;
      const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(GT_A_COL + 32 * h);
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
          "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
          "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]), "r"(o[8]), "r"(o[9]), "r"(o[10]),
          "r"(o[11]), "r"(o[12]), "r"(o[13]), "r"(o[14]), "r"(o[15]), "r"(o[16]), "r"(o[17]), "r"(o[18]), "r"(o[19]), "r"(o[20]),
          "r"(o[21]), "r"(o[22]), "r"(o[23]), "r"(o[24]), "r"(o[25]), "r"(o[26]), "r"(o[27]), "r"(o[28]), "r"(o[29]), "r"(o[30]),
          "r"(o[31])
          : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full);
      return c;
    };
    int cnt_next = 0;
    if (iters > 0) cnt_next = build_A((int64_t)blockIdx.x);

    for (int64_t it = 0; it < iters; ++it) {
      const int64_t tile = (int64_t)blockIdx.x + it * gridDim.x;
      const int64_t nq0 = tile * GT_TILE + 32 * q;   // first observation of this quadrant
      // ---- (a) conversion role: lane = observation nq0 + lane -----------------------------------------------------------
      const int cnt = cnt_next;
      const long long off = (long long)cnt << GT_FXP;
      double* srow_c = Sq + lane * GT_PITCH;
#pragma unroll 1
      for (int c = 0; c < GT_CHUNKS; ++c) {
        const int64_t i = it * GT_CHUNKS + c;
        const int tb = (int)(i & 3);
        mbar_wait(t_full + 8 * tb, (uint32_t)((i >> 2) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const bool mine = (c < GT_CHUNKS - 1) || (h == 0);   // the last chunk holds 8 entries: the first half's
        uint32_t r[56];
        if (mine) {
          const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(tb * GT_TCOLS + 56 * h);
          GT_LD16(r, 0, taddr);
          GT_LD16(r, 16, taddr + 16);
          GT_LD16(r, 32, taddr + 32);
          GT_LD8(r, 48, taddr + 48);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(t_empty + 8 * tb);
        if (mine) {
          const int e0 = c * GT_CE + 8 * h;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t t0 = r[7 * k + 1] * 256u + r[7 * k], t1 = r[7 * k + 3] * 256u + r[7 * k + 2];
            const uint32_t t2 = r[7 * k + 5] * 256u + r[7 * k + 4];
            const uint32_t hi = r[7 * k + 6] * 65536u + t2;
            const unsigned long long lo = (unsigned long long)t1 * 65536ull + t0;
            const long long T = (long long)(lo + ((unsigned long long)hi << 32)) - off;
            srow_c[e0 + k] = fma(-(double)T, __ldg(tab + e0 + k), __ldg(tab + GT_TAB_C0 + e0 + k));
          }
        }
      }
      // every MMA of this tile has completed (t_full of its last chunk): the A operand is free for the next tile
      if (it + 1 < iters) cnt_next = build_A(tile + gridDim.x);
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");   // both warps of the quadrant have staged their entries

      // ---- (b) sweep role: two rounds of 8 observations -----------------------------------------------------------------
#pragma unroll 1
      for (int round = 0; round < 2; ++round) {
        const int os0 = 16 * h + 8 * round;               // first staging row (within the quadrant) of the round
        const int64_t nb = nq0 + os0;                     // first observation of the round
        const bool valid = nb + rw < n_obs;
        const int64_t n = valid ? nb + rw : n_obs - 1;
        if (j == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(bsrc + n * GT_M));   // b_i is needed right after the sweeps
        const int nmiss = __shfl_sync(FULL, cnt, os0 + rw);
        double* srow = Sq + (os0 + rw) * GT_PITCH;
        double m[4][GT_M];
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const int cc = j + 4 * qq;
          const int cb = cc * (cc + 1) / 2;
#pragma unroll
          for (int a = 0; a < GT_M; ++a) {
            int idx;
            if (a < 4 * qq) idx = cb + a;
            else if (a >= 4 * qq + 4) idx = a * (a + 1) / 2 + cc;
            else idx = (a >= cc) ? a * (a + 1) / 2 + cc : cb + a;
            m[qq][a] = srow[idx];
          }
        }
        __syncwarp();   // the staging rows of this round are free (reused for the output rows below)
        const bool empty = nmiss >= GT_D;   // every feature missing: unsupported by the reference (Util.hs:81); flagged, outputs zeroed
        if (empty && valid && j == 0) set_err(p.scal, 2.0);
        bool ok = true;
        double pprod = 1.0;
        sweep_invert16<4, false>(m, j, ok, pprod);
        if (!ok && valid && j == 0) set_err(p.scal, 1.0);
        // ez = M_i^-1 b ;  A_i = ez ez^T + s2 M_i^-1
        double ez[4] = {0.0, 0.0, 0.0, 0.0};
        {
          const double2* bp = reinterpret_cast<const double2*>(bsrc + n * GT_M);
#pragma unroll
          for (int hh = 0; hh < GT_M / 2; ++hh) {
            const double2 bv = bp[hh];
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              ez[qq] = fma(m[qq][2 * hh], bv.x, ez[qq]);
              ez[qq] = fma(m[qq][2 * hh + 1], bv.y, ez[qq]);
            }
          }
        }
#pragma unroll
        for (int a = 0; a < GT_M; ++a) {
          const double ea = __shfl_sync(FULL, ez[a / 4], a % 4, 4);
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            if (a < 4 * qq) continue;   // compile time: columns j + 4 qq > a lie above the diagonal for every lane
            const int cc = j + 4 * qq;
            const double v = empty ? 0.0 : fma(ea, ez[qq], s2 * m[qq][a]);
            if (a >= cc) srow[a * (a + 1) / 2 + cc] = v;
            // A_i is positive semi-definite: its largest |entry| is on the diagonal
            if (a == cc && valid) mxa = max(mxa, (uint32_t)__double2hiint(v) & 0x7fffffffu);
          }
        }
        if (valid) {
          double* ep = EA + n * GT_EAS + GT_NV + j;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            const double v = empty ? 0.0 : ez[qq];
            ep[4 * qq] = v;
            mxe = max(mxe, (uint32_t)__double2hiint(v) & 0x7fffffffu);
          }
        }
        __syncwarp();
        const int nrows = (int)min((int64_t)8, n_obs - nb);   // <= 0 past the end
        {
          const double* src = Sq + os0 * GT_PITCH;
          double* dst = EA + nb * GT_EAS;
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) {
            if (rr < nrows) {
#pragma unroll
              for (int ee = 0; ee < GT_NV; ee += 32)
                if (ee + 32 <= GT_NV || lane < GT_NV - ee) dst[rr * GT_EAS + ee + lane] = src[rr * GT_PITCH + ee + lane];
            }
          }
        }
        __syncwarp();
      }
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");   // the quadrant's staging rows are free for the next tile
    }
    if (eamax != nullptr) {
      mxa = __reduce_max_sync(FULL, mxa);
      mxe = __reduce_max_sync(FULL, mxe);
      if (lane == 0) {
        atomicMax(&eamax[0], mxa);   // max is order-independent: deterministic
        atomicMax(&eamax[1], mxe);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (w == GT_SOLVERS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

}  // namespace

size_t masked_gram_tc_image_bytes() { return (size_t)GT_CHUNKS * GT_B_BYTES; }
size_t masked_gram_tc_table_bytes() { return sizeof(double) * GT_TAB_DOUBLES; }
int masked_gram_tc_gate_index() { return GT_TAB_GATE; }

// B image, tables and the gate for this step's W, then the tensor-core solver (which returns at once when the gate is closed)
int masked_gram_tc_solve(gplvm_ppca_session* s, ShardState& sh, const double* bsrc) {
  if (s->D != GT_D || s->M != GT_M || !sh.gramB || !sh.gramT) return fail(GPLVM_ERR_STATE, "tensor-core Gram solver: unsupported shape");
  static bool attr_done[64] = {};
  if (!attr_done[sh.dev & 63]) {
    CUDA_TRY(cudaFuncSetAttribute(masked_solve_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM));
    attr_done[sh.dev & 63] = true;
  }
  GP_LAUNCH(gram_planes_kernel, dim3(GT_CHUNKS, GT_KB), 128, 0, sh.st, sh.par, sh.gramB, sh.gramT);
  if (sh.eamax) CUDA_TRY(cudaMemsetAsync(sh.eamax, 0, 24, sh.st));   // class maxima + precision-guard counters
  const int grid = (int)std::min<int64_t>(sm_count(sh.dev), ceil_div(sh.n, (int64_t)GT_TILE));
  GP_LAUNCH(masked_solve_tc_kernel, grid, GT_THREADS, GT_SMEM, sh.st, bsrc, sh.mbits, sh.n, sh.par, sh.EA, sh.eamax, sh.gramB, sh.gramT);
  return GPLVM_OK;
}

}  // namespace gplvm
