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
#include <cuda.h>

#include "ppca.cuh"

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
  static constexpr int SMEM = 1024 + STAGES * STAGE_BYTES + TPART_BYTES + EZ_BYTES + 64;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ double lds64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// proj: D x 16 row-major projection (A = W Minv for the E-step, W for the log-likelihood pass)
// partials: gridDim.x x S doubles; per CTA [XEz D x 16 | EE 16 x 16]
template <int DD>
__global__ void __launch_bounds__(NW * 32, 2) dense_stats_dmma_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                     const double* __restrict__ proj, int64_t n_obs,
                                                                     int64_t tiles_per_cta, double* __restrict__ partials,
                                                                     int64_t S, const double* __restrict__ scal) {
  using C = Cfg<DD>;
  if (scal[S_DONE] != 0.0) return;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t tpart_u = base + C::STAGES * C::STAGE_BYTES;
  const uint32_t ez_u = tpart_u + C::TPART_BYTES;
  const uint32_t bar_u = ez_u + C::EZ_BYTES;
  double* tpart = reinterpret_cast<double*>(gen_base + C::STAGES * C::STAGE_BYTES);
  double* ezs = reinterpret_cast<double*>(gen_base + C::STAGES * C::STAGE_BYTES + C::TPART_BYTES);

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
  double preg[C::KS1][2];
#pragma unroll
  for (int ks = 0; ks < C::KS1; ++ks)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) preg[ks][mt] = proj[(size_t)(w * C::FW + ks * 4 + t) * MPF + mt * 8 + g];

  double acc[C::MT2][2][2];  // XEz[feature = w FW + 8 mt + g][latent = 8 nt + 2 t + {0,1}]
#pragma unroll
  for (int mt = 0; mt < C::MT2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
  double ee[2][2][2];        // EE[8 a + g][8 b + 2 t + {0,1}] (this warp's k-steps only)
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) ee[a][b][0] = ee[a][b][1] = 0.0;

  // per-thread address pieces (see the swizzle algebra in DESIGN.md):
  //   step 1 reads Xc[row 8 nt + g][feature f + t]      -> 16-byte unit (u0 ^ gx), gx = g ^ (t >> 1)
  //   step 2 reads Xc[row rho(j, t)][feature f8 + g]    -> 16-byte unit (cc ^ pg), pg = (g >> 1) ^ (t << 1)
  const uint32_t s1_thr = (uint32_t)(g * 128 + (t & 1) * 8);
  const uint32_t gx4 = (uint32_t)((g ^ (t >> 1)) << 4);
  const uint32_t s2_thr = (uint32_t)(t * 256 + (g & 1) * 8);
  const uint32_t pg4 = (uint32_t)((((g >> 1) ^ (t << 1)) & 7) << 4);
  const uint32_t ez_thr = ez_u + (uint32_t)((t * 2 * EZ_LD + g) * 8);

  for (int64_t i = 0; i < ntiles; ++i) {
    const int s = (int)(i % C::STAGES);
    const uint32_t xs = base + s * C::STAGE_BYTES;
    mbar_wait(bar_u + 8 * s, (uint32_t)((i / C::STAGES) & 1));

    // ---- step 1: partial T^T (16 latent x 16 rows) over this warp's feature slice ---------------------
    double tp[2][2][2];
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
      const double x0 = lds64(a0);
      const double x1 = lds64(a0 + 1024);
      dmma(tp[0][0], preg[ks][0], x0);
      dmma(tp[1][0], preg[ks][1], x0);
      dmma(tp[0][1], preg[ks][0], x1);
      dmma(tp[1][1], preg[ks][1], x1);
    }
    // tp[mt][nt][c] = T[row 8 nt + 2 t + c][latent 8 mt + g]  -> tpart[w][row][latent]
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int c = 0; c < 2; ++c) tpart[(w * RT + nt * 8 + 2 * t + c) * EZ_LD + mt * 8 + g] = tp[mt][nt][c];
    __syncthreads();  // S1: partials complete; every warp is also done with step 2 of the previous tile

    if (tid == 0 && i >= 1 && i - 1 + C::STAGES < ntiles) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(i - 1 + C::STAGES);
    }
    {  // fixed-order sum of the 4 partials: thread -> (row, latent pair)
      const int row = tid >> 3, lp = (tid & 7) * 2;
      double2 sacc = *reinterpret_cast<const double2*>(&tpart[(0 * RT + row) * EZ_LD + lp]);
#pragma unroll
      for (int q = 1; q < NW; ++q) {
        const double2 v = *reinterpret_cast<const double2*>(&tpart[(q * RT + row) * EZ_LD + lp]);
        sacc.x += v.x;
        sacc.y += v.y;
      }
      *reinterpret_cast<double2*>(&ezs[row * EZ_LD + lp]) = sacc;
    }
    __syncthreads();  // S2: Ez tile complete

    // ---- step 2: XEz += Xc^T . Ez ;  EE += Ez^T . Ez ----------------------------------------------------
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rowc = 8 * (j >> 1) + (j & 1);  // rows rho(j, t) = rowc + 2 t
      const double e0 = lds64(ez_thr + (uint32_t)((rowc * EZ_LD) * 8));
      const double e1 = lds64(ez_thr + (uint32_t)((rowc * EZ_LD + 8) * 8));
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
  }

  // ---- epilogue: per-CTA partial statistics ---------------------------------------------------------------
  double* out = partials + (size_t)blockIdx.x * S;
#pragma unroll
  for (int mt = 0; mt < C::MT2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int f = w * C::FW + mt * 8 + g;
      *reinterpret_cast<double2*>(&out[(size_t)f * MPF + nt * 8 + 2 * t]) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
    }
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn* fn) {
  static EncodeTiledFn cached = nullptr;
  if (!cached) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !p) return fail(GPLVM_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    cached = reinterpret_cast<EncodeTiledFn>(p);
  }
  *fn = cached;
  return GPLVM_OK;
}

template <int DD>
int launch_dmma(ShardState& sh, const double* proj, int64_t S, int grid, int64_t tiles_per_cta) {
  using C = Cfg<DD>;
  static bool attr_done[64] = {};
  if (!attr_done[sh.dev & 63]) {
    CUDA_TRY(cudaFuncSetAttribute(dense_stats_dmma_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    CUDA_TRY(cudaFuncSetAttribute(dense_stats_dmma_kernel<DD>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_done[sh.dev & 63] = true;
  }
  const CUtensorMap* tm = static_cast<const CUtensorMap*>(sh.tmap);
  GP_LAUNCH(dense_stats_dmma_kernel<DD>, grid, NW * 32, C::SMEM, sh.st, *tm, proj, sh.n, tiles_per_cta, sh.partials, S, sh.par.scal);
  return GPLVM_OK;
}

int sm_count(int dev) {
  static int cache[64] = {};
  if (!cache[dev & 63]) cudaDeviceGetAttribute(&cache[dev & 63], cudaDevAttrMultiProcessorCount, dev);
  return cache[dev & 63] > 0 ? cache[dev & 63] : 148;
}

}  // namespace

bool dense_dmma_eligible(const gplvm_ppca_session* s) {
  if (s->missing) return false;
  if (s->M > MPF) return false;
  return s->D == 64 || s->D == 128 || s->D == 256;
}

int dense_dmma_nparts(const gplvm_ppca_session*, const ShardState& sh) { return 2 * sm_count(sh.dev); }

int dense_dmma_prepare(gplvm_ppca_session* s) {
  EncodeTiledFn enc = nullptr;
  GP_TRY(get_encode_fn(&enc));
  for (ShardState& sh : s->sh) {
    CUDA_TRY(cudaSetDevice(sh.dev));
    if (!sh.tmap) {
      void* p = nullptr;
      if (posix_memalign(&p, 128, sizeof(CUtensorMap)) != 0) return fail(GPLVM_ERR_CUDA, "out of host memory");
      sh.tmap = p;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)s->D, (cuuint64_t)sh.n};
    const cuuint64_t gstr[1] = {(cuuint64_t)sh.ldx * sizeof(double)};
    const cuuint32_t box[2] = {16u, (cuuint32_t)RT};
    const cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(static_cast<CUtensorMap*>(sh.tmap), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, sh.X, gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(GPLVM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  }
  return GPLVM_OK;
}

int dense_dmma_stats(gplvm_ppca_session* s, ShardState& sh, const double* proj) {
  const int64_t ntiles = ceil_div(sh.n, RT);
  int grid = (int)std::min<int64_t>(2 * sm_count(sh.dev), ntiles);
  if (grid > sh.nparts) grid = sh.nparts;
  const int64_t per = ceil_div(ntiles, grid);
  grid = (int)ceil_div(ntiles, per);
  int rc;
  switch (s->D) {
    case 64: rc = launch_dmma<64>(sh, proj, s->S, grid, per); break;
    case 128: rc = launch_dmma<128>(sh, proj, s->S, grid, per); break;
    case 256: rc = launch_dmma<256>(sh, proj, s->S, grid, per); break;
    default: return fail(GPLVM_ERR_STATE, "dense DMMA kernel: unsupported D");
  }
  GP_TRY(rc);
  return launch_reduce_partials(sh, s->S, grid);
}

int dense_dmma_release(gplvm_ppca_session*) { return GPLVM_OK; }

}  // namespace gplvm
