// dense_v2.cu -- one-barrier-per-tile variant of the fused dense statistics kernel (see dense_dmma.cu for the maths,
// the TMA staging and the swizzle algebra of the observation tile).
//
// dense_dmma.cu synchronises twice per 16-observation tile: after the split-K partial products of step 1 and after
// their fixed-order sum into the Ez tile.  Here the sum is folded into step 2's fragment loads: every warp adds the four
// partial tiles (always in the order q = 0, 1, 2, 3 => deterministic) while it builds its B fragments, so the reduction
// phase and the second barrier disappear.  The partial tiles are double buffered (tile i + 2 reuses buffer i & 1 only
// after every warp has passed barrier i + 1, i.e. finished step 2 of tile i) and XOR-swizzled instead of padded
// (16-byte unit u of row r lives at u ^ X(r), X(r) = (r & 6) ^ ((r & 1) << 1): conflict-free for the step-1 stores
// and the step-2 loads), which is what makes two CTAs fit in one SM's shared memory:
//   3 x 32 KB stages + 2 x 8 KB partial tiles + barriers = 114 752 B per CTA (D = 256); no alignment slack -- the
//   dynamic shared memory base must be 1024-byte aligned (checked at run time; the kernel has no static shared memory).
#include "ppca.cuh"
#include "tma.cuh"

namespace gplvm {
namespace {

constexpr int RT = 16;
constexpr int MPF = 16;
constexpr int NW = 4;

template <int DD>
struct V2Cfg {
  static constexpr int CHUNKS = DD / 16;
  static constexpr int STAGE_BYTES = RT * DD * 8;
  static constexpr int STAGES = (DD >= 256) ? 3 : 4;
  static constexpr int FW = DD / NW;
  static constexpr int KS1 = FW / 4;
  static constexpr int MT2 = FW / 8;
  static constexpr int TP_BYTES = NW * RT * MPF * 8;   // 4 partial tiles of 16 x 16 doubles
  static constexpr int SMEM = STAGES * STAGE_BYTES + 2 * TP_BYTES + 64;
};

__device__ __forceinline__ void sts64(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }

template <int DD>
__global__ void __launch_bounds__(NW * 32, 2) dense_stats_v2_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                   const double* __restrict__ proj, int64_t n_obs,
                                                                   int64_t tiles_per_cta, double* __restrict__ partials, int64_t S,
                                                                   double* __restrict__ scal) {
  using C = V2Cfg<DD>;
  if (scal[S_DONE] != 0.0) return;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) {  // SWIZZLE_128B destinations must be 1024-byte aligned
    if (threadIdx.x == 0) set_err(scal, 4.0);
    return;
  }
  const uint32_t tp_u = base + C::STAGES * C::STAGE_BYTES;
  const uint32_t bar_u = tp_u + 2 * C::TP_BYTES;

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
  auto issue = [&](int64_t i) {
    const int s = (int)(i % C::STAGES);
    const uint32_t bar = bar_u + 8 * s;
    const uint32_t dst = base + s * C::STAGE_BYTES;
    const int row0 = (int)((tile0 + i) * RT);
    mbar_expect_tx(bar, C::STAGE_BYTES);
#pragma unroll
    for (int c = 0; c < C::CHUNKS; ++c) tma_load_2d(dst + c * (RT * 128), &tmap, bar, c * 16, row0);
  };
  if (tid == 0)
    for (int64_t i = 0; i < C::STAGES && i < ntiles; ++i) issue(i);

  double preg[C::KS1][2];
#pragma unroll
  for (int ks = 0; ks < C::KS1; ++ks)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) preg[ks][mt] = proj[(size_t)(w * C::FW + ks * 4 + t) * MPF + mt * 8 + g];
  double acc[C::MT2][2][2];
#pragma unroll
  for (int mt = 0; mt < C::MT2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
  double ee[2][2][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) ee[a][b][0] = ee[a][b][1] = 0.0;

  // observation-tile addressing (identical to dense_dmma.cu)
  const int rho1g = 2 * (g & 3) + (g >> 2);
  const uint32_t s1_thr = (uint32_t)(rho1g * 128 + (t & 1) * 8);
  const uint32_t gx4 = (uint32_t)((rho1g ^ (t >> 1)) << 4);
  const uint32_t s2_thr = (uint32_t)(t * 256 + (g & 1) * 8);
  const uint32_t pg4 = (uint32_t)((((g >> 1) ^ (t << 1)) & 7) << 4);
  // partial-tile addressing.  Stores: C-fragment column 2 t + c of n-tile nt is row 8 nt + rho1(2 t + c); the thread's
  // latent column is 8 mt + g.  Loads: row rho(j, t) = rowc + 2 t, latent column 8 nt + g.
  const int rl0 = 2 * ((2 * t) & 3) + ((2 * t) >> 2), rl1 = 2 * ((2 * t + 1) & 3) + ((2 * t + 1) >> 2);
  const uint32_t x0 = (uint32_t)((rl0 & 6) ^ ((rl0 & 1) << 1)), x1 = (uint32_t)((rl1 & 6) ^ ((rl1 & 1) << 1));
  const uint32_t st0 = (uint32_t)(w * 2048 + rl0 * 128 + (g & 1) * 8), st1 = (uint32_t)(w * 2048 + rl1 * 128 + (g & 1) * 8);
  const uint32_t gu = (uint32_t)(g >> 1);

  for (int64_t i = 0; i < ntiles; ++i) {
    const int s = (int)(i % C::STAGES);
    const uint32_t xs = base + s * C::STAGE_BYTES;
    const uint32_t tpb = tp_u + (uint32_t)(i & 1) * C::TP_BYTES;
    mbar_wait(bar_u + 8 * s, (uint32_t)((i / C::STAGES) & 1));

    // ---- step 1: partial T^T (16 latent x 16 rows) over this warp's feature slice -------------------------------
    double tp[2][2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) tp[mt][nt][0] = tp[mt][nt][1] = 0.0;
#pragma unroll
    for (int ks = 0; ks < C::KS1; ++ks) {
      const int f = w * C::FW + ks * 4;
      const uint32_t chunk_off = (uint32_t)((f >> 4) * (RT * 128));
      const uint32_t u0 = (uint32_t)(((ks * 4) & 15) >> 1) << 4;
      const uint32_t a0 = xs + chunk_off + s1_thr + (u0 ^ gx4);
      const double xa0 = lds64(a0);
      const double xa1 = lds64(a0 + 1024);
      dmma(tp[0][0], preg[ks][0], xa0);
      dmma(tp[1][0], preg[ks][1], xa0);
      dmma(tp[0][1], preg[ks][0], xa1);
      dmma(tp[1][1], preg[ks][1], xa1);
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        sts64(tpb + st0 + (uint32_t)(nt * 1024) + ((((uint32_t)(mt * 4) + gu) ^ x0) << 4), tp[mt][nt][0]);
        sts64(tpb + st1 + (uint32_t)(nt * 1024) + ((((uint32_t)(mt * 4) + gu) ^ x1) << 4), tp[mt][nt][1]);
      }
    __syncthreads();  // the only barrier of the tile: partials visible; every warp is done with step 2 of tile i - 1
    if (tid == 0 && i >= 1 && i - 1 + C::STAGES < ntiles) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(i - 1 + C::STAGES);
    }

    // ---- step 2: Ez fragments = fixed-order sum of the 4 partial tiles;  XEz += Xc^T Ez ;  EE += Ez^T Ez ---------------
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rowc = 8 * (j >> 1) + (j & 1);
      double e[2];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const uint32_t cc2 = (uint32_t)(((nt << 2) ^ ((j & 1) << 1)) << 4);
        const uint32_t a = tpb + (uint32_t)(rowc * 128) + s2_thr + (cc2 ^ pg4);
        double v = lds64(a);
        v += lds64(a + 2048);
        v += lds64(a + 4096);
        v += lds64(a + 6144);
        e[nt] = v;
      }
#pragma unroll
      for (int mt = 0; mt < C::MT2; ++mt) {
        const uint32_t chunk_off = (uint32_t)(((w * C::FW) >> 4) + (mt >> 1)) * (RT * 128);
        const uint32_t cc4 = (uint32_t)((((mt & 1) << 2) ^ (j & 1)) << 4);
        const double xa = lds64(xs + chunk_off + (uint32_t)(rowc * 128) + s2_thr + (cc4 ^ pg4));
        dmma(acc[mt][0], xa, e[0]);
        dmma(acc[mt][1], xa, e[1]);
      }
      if (j == w) {
        dmma(ee[0][0], e[0], e[0]);
        dmma(ee[0][1], e[0], e[1]);
        dmma(ee[1][0], e[1], e[0]);
        dmma(ee[1][1], e[1], e[1]);
      }
    }
  }

  // ---- epilogue: per-CTA partial statistics ------------------------------------------------------------------------
  double* out = partials + (size_t)blockIdx.x * S;
#pragma unroll
  for (int mt = 0; mt < C::MT2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int f = w * C::FW + mt * 8 + g;
      *reinterpret_cast<double2*>(&out[(size_t)f * MPF + nt * 8 + 2 * t]) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
    }
  __syncthreads();
  double* red = reinterpret_cast<double*>(smem_raw + C::STAGES * C::STAGE_BYTES);  // 4 x 16 x 16, plain layout
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int c = 0; c < 2; ++c) red[(w * 16 + a * 8 + g) * 16 + b * 8 + 2 * t + c] = ee[a][b][c];
  __syncthreads();
  for (int e = tid; e < MPF * MPF; e += NW * 32) {
    double sacc = 0.0;
#pragma unroll
    for (int q = 0; q < NW; ++q) sacc += red[q * 256 + e];
    out[(size_t)DD * MPF + e] = sacc;
  }
}

template <int DD>
int launch_v2(ShardState& sh, const double* proj, int64_t S, int grid, int64_t tiles_per_cta) {
  using C = V2Cfg<DD>;
  static bool attr_done[64] = {};
  if (!attr_done[sh.dev & 63]) {
    CUDA_TRY(cudaFuncSetAttribute((dense_stats_v2_kernel<DD>), cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    CUDA_TRY(cudaFuncSetAttribute((dense_stats_v2_kernel<DD>), cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_done[sh.dev & 63] = true;
  }
  const CUtensorMap* tm = static_cast<const CUtensorMap*>(sh.tmap);
  GP_LAUNCH((dense_stats_v2_kernel<DD>), grid, NW * 32, C::SMEM, sh.st, *tm, proj, sh.n, tiles_per_cta, sh.partials, S, sh.par.scal);
  return GPLVM_OK;
}

}  // namespace

// One-barrier dense statistics pass (same contract as dense_dmma_stats).
int dense_v2_stats(gplvm_ppca_session* s, ShardState& sh, const double* proj) {
  const int64_t ntiles = ceil_div(sh.n, RT);
  int grid = (int)std::min<int64_t>(2 * sm_count(sh.dev), ntiles);
  if (grid > sh.nparts) grid = sh.nparts;
  const int64_t per = ceil_div(ntiles, grid);
  grid = (int)ceil_div(ntiles, per);
  int rc;
  switch (s->D) {
    case 64: rc = launch_v2<64>(sh, proj, s->S, grid, per); break;
    case 128: rc = launch_v2<128>(sh, proj, s->S, grid, per); break;
    case 256: rc = launch_v2<256>(sh, proj, s->S, grid, per); break;
    default: return fail(GPLVM_ERR_STATE, "fused DMMA kernel: unsupported D");
  }
  GP_TRY(rc);
  return launch_reduce_partials(sh, s->S, grid);
}

}  // namespace gplvm
