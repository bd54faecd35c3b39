// dense_dmma.cu -- fused dense E-step statistics kernel: TMA-staged FP64 tensor-core (DMMA) contractions.
//
// Replaces, per EM step, the reference's materialised E[z] and its three passes over the data
// (src/GPLVM/PPCA.hs:67-69,78) by ONE pass over the centred observations Xc (n x D, observation-major):
//     Ez  = Xc . A                 (n x 16;   A = W Minv, so Ez rows are E[z_n], PPCA.hs:67)
//     XEz = Xc^T . Ez              (D x 16;   PPCA.hs:68)
//     EE  = Ez^T . Ez              (16 x 16;  the data part of E[z z^T], PPCA.hs:69)
// Ez never leaves the SM.  Algorithmic cost per observation: 8 D bytes read, 4 D M + 2 M^2 flop.
//
// Hardware mapping (sm_100a).  tcgen05/TMEM has no FP64 kind, so the FP64 tensor path is mma.sync.m8n8k4.f64
// (SASS DMMA.8x8x4).  One CTA = 4 warps, 2 CTAs per SM (so one CTA's barriers hide under the other's math).
// Per pipeline stage a CTA owns RT = 16 observations:
//   TMA    : D/16 boxes of 16 doubles x 16 rows (128-byte rows, SWIZZLE_128B) land in shared memory and
//            complete on one mbarrier; a stage is re-armed right after the CTA barrier that proves every warp
//            is done with it (no consumer->producer mbarrier needed).
//   step 1 : T^T = A^T . Xc^T with the latent dimension on the MMA M axis.  Warp q owns the feature slice
//            [q D/4, (q+1) D/4) (split K); its A^T fragments live in registers for the whole kernel.  The four
//            partial 16 x 16 tiles are exchanged through shared memory and summed in fixed order -> Ez tile.
//   step 2 : XEz += Xc^T . Ez with features on the M axis; warp q owns the same feature slice as accumulators.
//            The k (observation) order inside an MMA is rows {r, r+2, r+4, r+6}, which makes the swizzled
//            shared-memory reads of Xc^T fragments bank-conflict free.
//   EE     : the B fragments of step 2 are also the A fragments of Ez^T . Ez; warp q handles k-step q.
// Every partial result is written per CTA and summed over CTAs in fixed order by reduce_partials_kernel
// (deterministic: the Haskell signature above this library is pure).
#include "ppca.cuh"
#include "tma.cuh"

namespace gplvm {
namespace {

constexpr int RT = 16;      // observations per stage
constexpr int MPF = 16;     // latent dimension of the fused kernel (M <= 16, zero padded)
constexpr int EZ_LD = 18;   // row stride (doubles) of the Ez and partial tiles: conflict-free fragment traffic
constexpr int NW = 4;       // warps per CTA

template <int DD>
struct Cfg {
  static constexpr int CHUNKS = DD / 16;            // 128-byte column chunks per row
  static constexpr int STAGE_BYTES = RT * DD * 8;
  static constexpr int STAGES = (DD >= 256) ? 3 : 4;
  static constexpr int FW = DD / NW;                // features per warp
  static constexpr int KS1 = FW / 4;                // k-steps per warp in step 1
  static constexpr int MT2 = FW / 8;                // m-tiles per warp in step 2
  static constexpr int TPART_BYTES = NW * RT * EZ_LD * 8;
  static constexpr int EZ_BYTES = RT * EZ_LD * 8;
  static constexpr int SMEM = 1024 + STAGES * STAGE_BYTES + TPART_BYTES + 2 * EZ_BYTES + 64 + NW * RT * 8;
};

enum StatsMode {
  MODE_DENSE = 0,      // Ez = Xc proj ; XEz += Xc^T Ez ; EE += Ez^T Ez                       (dense EM step / dense LL pass)
  MODE_MASKED_B = 1,   // b_i = sum_{d observed} w_d (y_id - mu_d), yy_i = sum_{d obs} (y_id - mu_d)^2 -> HBM (PPCA.hs:117)
  MODE_MASKED_S2 = 2,  // S2 += X0^T Ez with X0 = Y, NaN -> 0 and Ez read from HBM             (PPCA.hs:136)
  MODE_MASKED_BE = 3   // MODE_MASKED_B without yy_i: the E-step only needs b_i (yy_i enters the likelihood, PPCA.hs:149-159),
                       // and its two DFMA per k-step share the FP64 datapath with the DMMAs
};
__host__ __device__ constexpr bool mode_is_b(int mode) { return mode == MODE_MASKED_B || mode == MODE_MASKED_BE; }

// proj     : D x 16 row-major projection (DENSE: A = W Minv or W; MASKED_B: W)
// partials : gridDim.x x S doubles; per CTA [XEz D x 16 | EE 16 x 16] (DENSE) or [S2 D x 16] (MASKED_S2)
// mu       : D feature means (MASKED_B)         bout / yyout : n x 16 / n outputs of MASKED_B
// ezsrc    : E[z_i] rows with leading dimension ldez (MASKED_S2)
template <int DD, int MODE>
__global__ void __launch_bounds__(NW * 32, 2) dense_stats_dmma_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                     const double* __restrict__ proj, int64_t n_obs,
                                                                     int64_t tiles_per_cta, double* __restrict__ partials,
                                                                     int64_t S, const double* __restrict__ scal,
                                                                     const double* __restrict__ mu, double* __restrict__ bout,
                                                                     double* __restrict__ yyout, const double* __restrict__ ezsrc,
                                                                     int64_t ldez) {
  using C = Cfg<DD>;
  if (scal[S_DONE] != 0.0) return;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t tpart_u = base + C::STAGES * C::STAGE_BYTES;
  const uint32_t ez_u = tpart_u + C::TPART_BYTES;
  const uint32_t bar_u = ez_u + 2 * C::EZ_BYTES;
  double* tpart = reinterpret_cast<double*>(gen_base + C::STAGES * C::STAGE_BYTES);
  double* ezs = reinterpret_cast<double*>(gen_base + C::STAGES * C::STAGE_BYTES + C::TPART_BYTES);  // 2 buffers
  double* yys = reinterpret_cast<double*>(gen_base + C::STAGES * C::STAGE_BYTES + C::TPART_BYTES + 2 * C::EZ_BYTES + 64);

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, t = lane & 3;

  const int64_t ntiles_all = (n_obs + RT - 1) / RT;
  const int64_t tile0 = (int64_t)blockIdx.x * tiles_per_cta;
  int64_t ntiles = ntiles_all - tile0;
  if (ntiles > tiles_per_cta) ntiles = tiles_per_cta;
  if (ntiles < 0) ntiles = 0;

  if (tid == 0) {
    for (int s = 0; s < C::STAGES; ++s) mbar_init(bar_u + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
  }
  __syncthreads();

  auto issue = [&](int64_t i) {  // thread 0 only: arm stage i % STAGES with tile i
    const int s = (int)(i % C::STAGES);
    const uint32_t bar = bar_u + 8 * s;
    const uint32_t dst = base + s * C::STAGE_BYTES;
    const int row0 = (int)((tile0 + i) * RT);
    mbar_expect_tx(bar, C::STAGE_BYTES);
#pragma unroll
    for (int c = 0; c < C::CHUNKS; ++c) tma_load_2d(dst + c * (RT * 128), &tmap, bar, c * 16, row0);
  };
  if (tid == 0) {
    for (int64_t i = 0; i < C::STAGES && i < ntiles; ++i) issue(i);
  }

  // step-1 A^T fragments: proj[feature = w FW + 4 ks + t][latent = 8 mt + g], resident for the whole kernel
  constexpr int NP = (MODE == MODE_MASKED_S2) ? 1 : C::KS1;
  double preg[NP][2];
  double mureg[mode_is_b(MODE) ? C::KS1 : 1];
  if (MODE != MODE_MASKED_S2) {
#pragma unroll
    for (int ks = 0; ks < NP; ++ks) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) preg[ks][mt] = proj[(size_t)(w * C::FW + ks * 4 + t) * MPF + mt * 8 + g];
      if (mode_is_b(MODE)) mureg[ks] = mu[w * C::FW + ks * 4 + t];
    }
  }

  constexpr int NA = mode_is_b(MODE) ? 1 : C::MT2;
  double acc[NA][2][2];  // XEz[feature = w FW + 8 mt + g][latent = 8 nt + 2 t + {0,1}]
#pragma unroll
  for (int mt = 0; mt < NA; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
  double ee[2][2][2];        // EE[8 a + g][8 b + 2 t + {0,1}] (this warp's k-steps only)
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) ee[a][b][0] = ee[a][b][1] = 0.0;

  // per-thread address pieces (see the swizzle algebra in DESIGN.md):
  //   step 1 reads Xc[row 8 nt + rho1(g)][feature f + t] -> 16-byte unit (u0 ^ gx), gx = rho1(g) ^ (t >> 1)
  //   step 2 reads Xc[row rho(j, t)][feature f8 + g]    -> 16-byte unit (cc ^ pg), pg = (g >> 1) ^ (t << 1)
  // Step-1 n-tiles use the row order rho1(n) = 2 (n & 3) + (n >> 2), i.e. lanes g = 0..3 read rows 0, 2, 4, 6 and
  // g = 4..7 rows 1, 3, 5, 7 of an 8-row group: a 64-bit shared load is served per half-warp, and rows r, r ^ 1 keep
  // their two 16-byte units in the same pair of bank groups under SWIZZLE_128B (2-way conflict with rows 0..3).
  const int rho1g = 2 * (g & 3) + (g >> 2);
  const uint32_t s1_thr = (uint32_t)(rho1g * 128 + (t & 1) * 8);
  const uint32_t gx4 = (uint32_t)((rho1g ^ (t >> 1)) << 4);
  const int rho1c0 = 2 * ((2 * t) & 3) + ((2 * t) >> 2);          // row (within its group of 8) of C-fragment column 2 t
  const int rho1c1 = 2 * ((2 * t + 1) & 3) + ((2 * t + 1) >> 2);  // ... and of column 2 t + 1
  const uint32_t s2_thr = (uint32_t)(t * 256 + (g & 1) * 8);
  const uint32_t pg4 = (uint32_t)((((g >> 1) ^ (t << 1)) & 7) << 4);
  const uint32_t ez_thr = ez_u + (uint32_t)((t * 2 * EZ_LD + g) * 8);
  const int rrow = tid >> 3, rlp = (tid & 7) * 2;  // (row, latent pair) of the reduce / Ez staging role

  double2 eznext = make_double2(0.0, 0.0);
  if (MODE == MODE_MASKED_S2 && ntiles > 0) {
    const int64_t r = tile0 * RT + rrow;
    if (r < n_obs) eznext = *reinterpret_cast<const double2*>(&ezsrc[r * ldez + rlp]);
  }

  for (int64_t i = 0; i < ntiles; ++i) {
    const int s = (int)(i % C::STAGES);
    const uint32_t xs = base + s * C::STAGE_BYTES;
    uint32_t ezbuf_off = 0;

    if (MODE == MODE_MASKED_S2) {
      // Ez tile of this stage from HBM (prefetched one tile ahead); double-buffered => one CTA barrier per tile
      ezbuf_off = (uint32_t)((i & 1) * C::EZ_BYTES);
      *reinterpret_cast<double2*>(&ezs[(i & 1) * (RT * EZ_LD) + rrow * EZ_LD + rlp]) = eznext;
      eznext = make_double2(0.0, 0.0);
      if (i + 1 < ntiles) {
        const int64_t r = (tile0 + i + 1) * RT + rrow;
        if (r < n_obs) eznext = *reinterpret_cast<const double2*>(&ezsrc[r * ldez + rlp]);
      }
      mbar_wait(bar_u + 8 * s, (uint32_t)((i / C::STAGES) & 1));
      __syncthreads();  // Ez tile visible; every warp is done with step 2 of the previous tile
      if (tid == 0 && i >= 1 && i - 1 + C::STAGES < ntiles) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(i - 1 + C::STAGES);
      }
    } else {
      mbar_wait(bar_u + 8 * s, (uint32_t)((i / C::STAGES) & 1));

      // ---- step 1: partial T^T (16 latent x 16 rows) over this warp's feature slice -------------------
      double tp[2][2][2];
      double yacc0 = 0.0, yacc1 = 0.0;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) tp[mt][nt][0] = tp[mt][nt][1] = 0.0;
#pragma unroll
      for (int ks = 0; ks < C::KS1; ++ks) {
        const int f = w * C::FW + ks * 4;  // w is warp-uniform: the chunk offset is a runtime add, the rest folds
        const uint32_t chunk_off = (uint32_t)((f >> 4) * (RT * 128));
        const uint32_t u0 = (uint32_t)(((ks * 4) & 15) >> 1) << 4;  // FW % 16 == 0 => (f & 15) == (4 ks & 15)
        const uint32_t a0 = xs + chunk_off + s1_thr + (u0 ^ gx4);
        double x0 = lds64(a0);
        double x1 = lds64(a0 + 1024);
        if (mode_is_b(MODE)) {  // missing -> 0, observed -> y - mu   (PPCA.hs:112-113,117)
          x0 = (x0 == x0) ? x0 - mureg[ks] : 0.0;
          x1 = (x1 == x1) ? x1 - mureg[ks] : 0.0;
          if (MODE == MODE_MASKED_B) {
            yacc0 = fma(x0, x0, yacc0);
            yacc1 = fma(x1, x1, yacc1);
          }
        }
        dmma(tp[0][0], preg[ks][0], x0);
        dmma(tp[1][0], preg[ks][1], x0);
        dmma(tp[0][1], preg[ks][0], x1);
        dmma(tp[1][1], preg[ks][1], x1);
      }
      // tp[mt][nt][c] = T[row 8 nt + rho1(2 t + c)][latent 8 mt + g]  -> tpart[w][row][latent]
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          tpart[(w * RT + nt * 8 + rho1c0) * EZ_LD + mt * 8 + g] = tp[mt][nt][0];
          tpart[(w * RT + nt * 8 + rho1c1) * EZ_LD + mt * 8 + g] = tp[mt][nt][1];
        }
      if (MODE == MODE_MASKED_B) {
        yacc0 += __shfl_xor_sync(0xffffffffu, yacc0, 1);
        yacc0 += __shfl_xor_sync(0xffffffffu, yacc0, 2);
        yacc1 += __shfl_xor_sync(0xffffffffu, yacc1, 1);
        yacc1 += __shfl_xor_sync(0xffffffffu, yacc1, 2);
        if (t == 0) {
          yys[w * RT + rho1g] = yacc0;
          yys[w * RT + 8 + rho1g] = yacc1;
        }
      }
      __syncthreads();  // S1: partials complete; every warp is also done with step 2 of the previous tile

      if (mode_is_b(MODE)) {  // step 1 was the only reader of this stage: re-arm it right away
        if (tid == 0 && i + C::STAGES < ntiles) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          issue(i + C::STAGES);
        }
      } else if (tid == 0 && i >= 1 && i - 1 + C::STAGES < ntiles) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(i - 1 + C::STAGES);
      }
      {  // fixed-order sum of the 4 partials: thread -> (row, latent pair)
        double2 sacc = *reinterpret_cast<const double2*>(&tpart[(0 * RT + rrow) * EZ_LD + rlp]);
#pragma unroll
        for (int q = 1; q < NW; ++q) {
          const double2 v = *reinterpret_cast<const double2*>(&tpart[(q * RT + rrow) * EZ_LD + rlp]);
          sacc.x += v.x;
          sacc.y += v.y;
        }
        if (mode_is_b(MODE)) {
          const int64_t r = (tile0 + i) * RT + rrow;
          if (r < n_obs) {
            *reinterpret_cast<double2*>(&bout[r * MPF + rlp]) = sacc;
            if (MODE == MODE_MASKED_B && rlp == 0) yyout[r] = ((yys[rrow] + yys[RT + rrow]) + yys[2 * RT + rrow]) + yys[3 * RT + rrow];
          }
        } else {
          *reinterpret_cast<double2*>(&ezs[rrow * EZ_LD + rlp]) = sacc;
        }
      }
      __syncthreads();  // S2: Ez tile complete (MASKED_B: partial tiles free again)
    }

    if (mode_is_b(MODE)) continue;

    // ---- step 2: XEz += Xc^T . Ez ;  EE += Ez^T . Ez ----------------------------------------------------
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rowc = 8 * (j >> 1) + (j & 1);  // rows rho(j, t) = rowc + 2 t
      const double e0 = lds64(ez_thr + ezbuf_off + (uint32_t)((rowc * EZ_LD) * 8));
      const double e1 = lds64(ez_thr + ezbuf_off + (uint32_t)((rowc * EZ_LD + 8) * 8));
#pragma unroll
      for (int mt = 0; mt < NA; ++mt) {
        const uint32_t chunk_off = (uint32_t)(((w * C::FW) >> 4) + (mt >> 1)) * (RT * 128);
        const uint32_t cc4 = (uint32_t)((((mt & 1) << 2) ^ (j & 1)) << 4);
        double xa = lds64(xs + chunk_off + (uint32_t)(rowc * 128) + s2_thr + (cc4 ^ pg4));
        if (MODE == MODE_MASKED_S2) xa = (xa == xa) ? xa : 0.0;
        dmma(acc[mt][0], xa, e0);
        dmma(acc[mt][1], xa, e1);
      }
      if (MODE == MODE_DENSE && j == w) {
        dmma(ee[0][0], e0, e0);
        dmma(ee[0][1], e0, e1);
        dmma(ee[1][0], e1, e0);
        dmma(ee[1][1], e1, e1);
      }
    }
  }
  if (mode_is_b(MODE)) return;

  // ---- epilogue: per-CTA partial statistics ---------------------------------------------------------------
  double* out = partials + (size_t)blockIdx.x * S;
#pragma unroll
  for (int mt = 0; mt < NA; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int f = w * C::FW + mt * 8 + g;
      *reinterpret_cast<double2*>(&out[(size_t)f * MPF + nt * 8 + 2 * t]) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
    }
  if (MODE != MODE_DENSE) return;
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int c = 0; c < 2; ++c) tpart[(w * RT + a * 8 + g) * EZ_LD + b * 8 + 2 * t + c] = ee[a][b][c];
  __syncthreads();
  for (int e = tid; e < MPF * MPF; e += NW * 32) {
    const int r = e >> 4, cidx = e & 15;
    double sacc = 0.0;
#pragma unroll
    for (int q = 0; q < NW; ++q) sacc += tpart[(q * RT + r) * EZ_LD + cidx];
    out[(size_t)DD * MPF + e] = sacc;
  }
}

struct MaskedArgs {
  const double* mu = nullptr;
  double* bout = nullptr;
  double* yyout = nullptr;
  const double* ezsrc = nullptr;
  int64_t ldez = 0;
};

template <int DD, int MODE>
int launch_dmma(ShardState& sh, const double* proj, int64_t S, int grid, int64_t tiles_per_cta, const MaskedArgs& ma) {
  using C = Cfg<DD>;
  static bool attr_done[64] = {};
  if (!attr_done[sh.dev & 63]) {
    CUDA_TRY(cudaFuncSetAttribute((dense_stats_dmma_kernel<DD, MODE>), cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    CUDA_TRY(cudaFuncSetAttribute((dense_stats_dmma_kernel<DD, MODE>), cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_done[sh.dev & 63] = true;
  }
  const CUtensorMap* tm = static_cast<const CUtensorMap*>(sh.tmap);
  GP_LAUNCH((dense_stats_dmma_kernel<DD, MODE>), grid, NW * 32, C::SMEM, sh.st, *tm, proj, sh.n, tiles_per_cta, sh.partials, S,
            sh.par.scal, ma.mu, ma.bout, ma.yyout, ma.ezsrc, ma.ldez);
  return GPLVM_OK;
}

template <int MODE>
int dispatch_dmma(gplvm_ppca_session* s, ShardState& sh, const double* proj, int64_t S, const MaskedArgs& ma, int* grid_out) {
  const int64_t ntiles = ceil_div(sh.n, RT);
  int grid = (int)std::min<int64_t>(2 * sm_count(sh.dev), ntiles);
  if (grid > sh.nparts) grid = sh.nparts;
  const int64_t per = ceil_div(ntiles, grid);
  grid = (int)ceil_div(ntiles, per);
  if (grid_out) *grid_out = grid;
  switch (s->D) {
    case 64: return launch_dmma<64, MODE>(sh, proj, S, grid, per, ma);
    case 128: return launch_dmma<128, MODE>(sh, proj, S, grid, per, ma);
    case 256: return launch_dmma<256, MODE>(sh, proj, S, grid, per, ma);
    default: return fail(GPLVM_ERR_STATE, "fused DMMA kernel: unsupported D");
  }
}

}  // namespace

bool dense_dmma_eligible(const gplvm_ppca_session* s) {
  if (s->M > MPF) return false;
  if (s->missing && s->M != MPF) return false;  // the fused masked solver is specialised for M = 16
  return s->D == 64 || s->D == 128 || s->D == 256;
}

int dense_dmma_nparts(const gplvm_ppca_session*, const ShardState& sh) { return 2 * sm_count(sh.dev); }

int dense_dmma_prepare(gplvm_ppca_session* s) {
  for (ShardState& sh : s->sh) {
    CUDA_TRY(cudaSetDevice(sh.dev));
    if (!sh.tmap) {
      void* p = nullptr;
      if (posix_memalign(&p, 128, sizeof(CUtensorMap)) != 0) return fail(GPLVM_ERR_CUDA, "out of host memory");
      sh.tmap = p;
    }
    GP_TRY(encode_rows_map(static_cast<CUtensorMap*>(sh.tmap), sh.X, s->D, sh.n, sh.ldx, RT));
  }
  return GPLVM_OK;
}

int dense_dmma_stats(gplvm_ppca_session* s, ShardState& sh, const double* proj) {
  int grid = 0;
  GP_TRY(dispatch_dmma<MODE_DENSE>(s, sh, proj, s->S, MaskedArgs(), &grid));
  seg_mark(sh, 1);
  return launch_reduce_partials(sh, s->S, grid);
}

// b_i = W_obs^T (y_i - mu)_obs and, when yyout is given (masked LL), yy_i = |(y_i - mu)_obs|^2 for every local observation
int masked_dmma_b(gplvm_ppca_session* s, ShardState& sh, const double* W, const double* mu, double* bout, double* yyout) {
  MaskedArgs ma;
  ma.mu = mu; ma.bout = bout; ma.yyout = yyout;
  return yyout ? dispatch_dmma<MODE_MASKED_B>(s, sh, W, 0, ma, nullptr) : dispatch_dmma<MODE_MASKED_BE>(s, sh, W, 0, ma, nullptr);
}

// dst (D x 16) = sum_i (y_i, NaN -> 0) ez_i^T, ez_i = row i of `ez` (leading dimension ldez)
int masked_dmma_s2(gplvm_ppca_session* s, ShardState& sh, const double* ez, int64_t ldez, double* dst) {
  MaskedArgs ma;
  ma.ezsrc = ez; ma.ldez = ldez;
  int grid = 0;
  const int64_t S = (int64_t)s->D * MPF;
  GP_TRY(dispatch_dmma<MODE_MASKED_S2>(s, sh, nullptr, S, ma, &grid));
  return launch_reduce_partials(sh, S, grid, dst);
}

int dense_dmma_release(gplvm_ppca_session*) { return GPLVM_OK; }

}  // namespace gplvm
