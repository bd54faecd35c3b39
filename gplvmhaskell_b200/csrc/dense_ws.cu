// dense_ws.cu -- warp-specialised variant of the fused dense statistics kernel (see dense_dmma.cu for the maths and
// the shared-memory layout).  One CTA = 8 warps in two role groups that run a producer/consumer pipeline, so that the
// FP64 tensor pipe always has work from one group while the other synchronises:
//   group P (warps 0-3): step 1, Ez = Xc . proj  (split K over the 4 warps, fixed-order sum through shared memory)
//   group A (warps 4-7): step 2, XEz += Xc^T . Ez and EE += Ez^T . Ez, for the tile P finished one step earlier
// Hand-offs are mbarriers (Ez tile full / free, double buffered); each group synchronises internally with its own
// named barrier (bar.sync id, 128); the TMA stage of tile i is re-armed by group A's first thread once the whole
// group is done with it.  2 CTAs per SM (<= 128 registers per thread: each role keeps only its own accumulators).
#include "ppca.cuh"
#include "tma.cuh"

namespace gplvm {
namespace {

constexpr int RT = 16;
constexpr int MPF = 16;
constexpr int EZ_LD = 18;
constexpr int NW = 4;  // warps per role group

template <int DD>
struct WsCfg {
  static constexpr int CHUNKS = DD / 16;
  static constexpr int STAGE_BYTES = RT * DD * 8;
  static constexpr int STAGES = (DD >= 256) ? 3 : 4;
  static constexpr int FW = DD / NW;
  static constexpr int KS1 = FW / 4;
  static constexpr int MT2 = FW / 8;
  static constexpr int TPART_BYTES = NW * RT * EZ_LD * 8;
  static constexpr int EZ_BYTES = RT * EZ_LD * 8;
  static constexpr int SMEM = 1024 + STAGES * STAGE_BYTES + TPART_BYTES + 2 * EZ_BYTES + 128;
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

template <int DD>
__global__ void __launch_bounds__(256, 2) dense_stats_ws_kernel(const __grid_constant__ CUtensorMap tmap,
                                                               const double* __restrict__ proj, int64_t n_obs,
                                                               int64_t tiles_per_cta, double* __restrict__ partials, int64_t S,
                                                               const double* __restrict__ scal) {
  using C = WsCfg<DD>;
  if (scal[S_DONE] != 0.0) return;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t tpart_u = base + C::STAGES * C::STAGE_BYTES;
  const uint32_t ez_u = tpart_u + C::TPART_BYTES;
  const uint32_t bar_u = ez_u + 2 * C::EZ_BYTES;      // full[STAGES] | ezfull[2] | ezfree[2]
  const uint32_t ezfull_u = bar_u + 8 * C::STAGES, ezfree_u = ezfull_u + 16;
  double* tpart = reinterpret_cast<double*>(gen_base + C::STAGES * C::STAGE_BYTES);
  double* ezs = reinterpret_cast<double*>(gen_base + C::STAGES * C::STAGE_BYTES + C::TPART_BYTES);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int64_t ntiles_all = (n_obs + RT - 1) / RT;
  const int64_t tile0 = (int64_t)blockIdx.x * tiles_per_cta;
  int64_t ntiles = ntiles_all - tile0;
  if (ntiles > tiles_per_cta) ntiles = tiles_per_cta;
  if (ntiles < 0) ntiles = 0;

  if (tid == 0) {
    for (int s = 0; s < C::STAGES; ++s) mbar_init(bar_u + 8 * s, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(ezfull_u + 8 * b, 128);
      mbar_init(ezfree_u + 8 * b, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
  }
  __syncthreads();

  auto issue = [&](int64_t i) {  // one thread: arm stage i % STAGES with tile i
    const int s = (int)(i % C::STAGES);
    const uint32_t bar = bar_u + 8 * s;
    const uint32_t dst = base + s * C::STAGE_BYTES;
    const int row0 = (int)((tile0 + i) * RT);
    mbar_expect_tx(bar, C::STAGE_BYTES);
#pragma unroll
    for (int c = 0; c < C::CHUNKS; ++c) tma_load_2d(dst + c * (RT * 128), &tmap, bar, c * 16, row0);
  };
  if (tid == 128) {
    for (int64_t i = 0; i < C::STAGES && i < ntiles; ++i) issue(i);
  }

  if (warp < NW) {
    // =============================== group P: step 1 ===============================================================
    const int w = warp;
    double preg[C::KS1][2];
#pragma unroll
    for (int ks = 0; ks < C::KS1; ++ks)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) preg[ks][mt] = proj[(size_t)(w * C::FW + ks * 4 + t) * MPF + mt * 8 + g];
    const int rho1g = 2 * (g & 3) + (g >> 2);   // conflict-free row order of the step-1 n-tiles (see dense_dmma.cu)
    const uint32_t s1_thr = (uint32_t)(rho1g * 128 + (t & 1) * 8);
    const uint32_t gx4 = (uint32_t)((rho1g ^ (t >> 1)) << 4);
    const int rho1c0 = 2 * ((2 * t) & 3) + ((2 * t) >> 2);
    const int rho1c1 = 2 * ((2 * t + 1) & 3) + ((2 * t + 1) >> 2);
    const int rrow = tid >> 3, rlp = (tid & 7) * 2;
    for (int64_t i = 0; i < ntiles; ++i) {
      const int s = (int)(i % C::STAGES);
      const uint32_t xs = base + s * C::STAGE_BYTES;
      mbar_wait(bar_u + 8 * s, (uint32_t)((i / C::STAGES) & 1));
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
        const double x0 = lds64(a0);
        const double x1 = lds64(a0 + 1024);
        dmma(tp[0][0], preg[ks][0], x0);
        dmma(tp[1][0], preg[ks][1], x0);
        dmma(tp[0][1], preg[ks][0], x1);
        dmma(tp[1][1], preg[ks][1], x1);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          tpart[(w * RT + nt * 8 + rho1c0) * EZ_LD + mt * 8 + g] = tp[mt][nt][0];
          tpart[(w * RT + nt * 8 + rho1c1) * EZ_LD + mt * 8 + g] = tp[mt][nt][1];
        }
      group_sync(1);  // partial tiles complete
      const int b = (int)(i & 1);
      if (i >= 2) mbar_wait(ezfree_u + 8 * b, (uint32_t)(((i >> 1) - 1) & 1));  // group A is done with Ez[b] of tile i - 2
      {
        double2 sacc = *reinterpret_cast<const double2*>(&tpart[(0 * RT + rrow) * EZ_LD + rlp]);
#pragma unroll
        for (int q = 1; q < NW; ++q) {
          const double2 v = *reinterpret_cast<const double2*>(&tpart[(q * RT + rrow) * EZ_LD + rlp]);
          sacc.x += v.x;
          sacc.y += v.y;
        }
        *reinterpret_cast<double2*>(&ezs[b * (RT * EZ_LD) + rrow * EZ_LD + rlp]) = sacc;
      }
      mbar_arrive(ezfull_u + 8 * b);
      group_sync(1);  // partial tiles free again
    }
  } else {
    // =============================== group A: step 2 ===============================================================
    const int w = warp - NW;
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
    const uint32_t s2_thr = (uint32_t)(t * 256 + (g & 1) * 8);
    const uint32_t pg4 = (uint32_t)((((g >> 1) ^ (t << 1)) & 7) << 4);
    const uint32_t ez_thr = ez_u + (uint32_t)((t * 2 * EZ_LD + g) * 8);
    for (int64_t i = 0; i < ntiles; ++i) {
      const int s = (int)(i % C::STAGES);
      const uint32_t xs = base + s * C::STAGE_BYTES;
      const int b = (int)(i & 1);
      mbar_wait(bar_u + 8 * s, (uint32_t)((i / C::STAGES) & 1));
      mbar_wait(ezfull_u + 8 * b, (uint32_t)((i >> 1) & 1));
      const uint32_t ezb = ez_thr + (uint32_t)(b * C::EZ_BYTES);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int rowc = 8 * (j >> 1) + (j & 1);
        const double e0 = lds64(ezb + (uint32_t)((rowc * EZ_LD) * 8));
        const double e1 = lds64(ezb + (uint32_t)((rowc * EZ_LD + 8) * 8));
#pragma unroll
        for (int mt = 0; mt < C::MT2; ++mt) {
          const uint32_t chunk_off = (uint32_t)(((w * C::FW) >> 4) + (mt >> 1)) * (RT * 128);
          const uint32_t cc4 = (uint32_t)((((mt & 1) << 2) ^ (j & 1)) << 4);
          const double xa = lds64(xs + chunk_off + (uint32_t)(rowc * 128) + s2_thr + (cc4 ^ pg4));
          dmma(acc[mt][0], xa, e0);
          dmma(acc[mt][1], xa, e1);
        }
        if (j == w) {
          dmma(ee[0][0], e0, e0);
          dmma(ee[0][1], e0, e1);
          dmma(ee[1][0], e1, e0);
          dmma(ee[1][1], e1, e1);
        }
      }
      mbar_arrive(ezfree_u + 8 * b);
      group_sync(2);  // the whole group is done with stage s (group P finished with it before it signalled Ez full)
      if (tid == 128 && i + C::STAGES < ntiles) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(i + C::STAGES);
      }
    }
    // per-CTA partial statistics
    double* out = partials + (size_t)blockIdx.x * S;
#pragma unroll
    for (int mt = 0; mt < C::MT2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int f = w * C::FW + mt * 8 + g;
        *reinterpret_cast<double2*>(&out[(size_t)f * MPF + nt * 8 + 2 * t]) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
      }
    // EE: the four warps hold partial sums; combine them in fixed order through the partial-tile area (group P read
    // it for the last time before it signalled the last Ez tile)
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int bb = 0; bb < 2; ++bb)
#pragma unroll
        for (int c = 0; c < 2; ++c) tpart[(w * RT + a * 8 + g) * EZ_LD + bb * 8 + 2 * t + c] = ee[a][bb][c];
    group_sync(2);
    for (int e = tid - 128; e < MPF * MPF; e += 128) {
      const int r = e >> 4, cidx = e & 15;
      double sacc = 0.0;
#pragma unroll
      for (int q = 0; q < NW; ++q) sacc += tpart[(q * RT + r) * EZ_LD + cidx];
      out[(size_t)DD * MPF + e] = sacc;
    }
  }
}

template <int DD>
int launch_ws(ShardState& sh, const double* proj, int64_t S, int grid, int64_t tiles_per_cta) {
  using C = WsCfg<DD>;
  static bool attr_done[64] = {};
  if (!attr_done[sh.dev & 63]) {
    CUDA_TRY(cudaFuncSetAttribute((dense_stats_ws_kernel<DD>), cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    CUDA_TRY(cudaFuncSetAttribute((dense_stats_ws_kernel<DD>), cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_done[sh.dev & 63] = true;
  }
  const CUtensorMap* tm = static_cast<const CUtensorMap*>(sh.tmap);
  GP_LAUNCH((dense_stats_ws_kernel<DD>), grid, 256, C::SMEM, sh.st, *tm, proj, sh.n, tiles_per_cta, sh.partials, S, sh.par.scal);
  return GPLVM_OK;
}

}  // namespace

// Warp-specialised dense statistics pass (same contract as dense_dmma_stats); returns the grid size used.
int dense_ws_stats(gplvm_ppca_session* s, ShardState& sh, const double* proj) {
  const int64_t ntiles = ceil_div(sh.n, RT);
  int grid = (int)std::min<int64_t>(2 * sm_count(sh.dev), ntiles);
  if (grid > sh.nparts) grid = sh.nparts;
  const int64_t per = ceil_div(ntiles, grid);
  grid = (int)ceil_div(ntiles, per);
  int rc;
  switch (s->D) {
    case 64: rc = launch_ws<64>(sh, proj, s->S, grid, per); break;
    case 128: rc = launch_ws<128>(sh, proj, s->S, grid, per); break;
    case 256: rc = launch_ws<256>(sh, proj, s->S, grid, per); break;
    default: return fail(GPLVM_ERR_STATE, "fused DMMA kernel: unsupported D");
  }
  GP_TRY(rc);
  return launch_reduce_partials(sh, s->S, grid);
}

}  // namespace gplvm
