// masked.cu -- NaN-masked PPCA EM step (emStepsMissed, reference src/GPLVM/PPCA.hs:96-182).
//
// Sufficient-statistics formulation, per observation i with O_i the
// observed features (old W, mu, s2):
//   M_i  = (s2 I + W^T W) - sum_{d missing} w_d w_d^T                                  (PPCA.hs:114-115)
//   b_i  = sum_{d in O_i} w_d (y_id - mu_d);  ez_i = M_i^-1 b_i                         (PPCA.hs:116-117)
//   A_i  = ez_i ez_i^T + s2 M_i^-1
// statistics (complement form: the work is proportional to the MISSING entries):
//   tot = sum_i [vech A_i | ez_i],   miss_d = sum_{i : d missing} [vech A_i | ez_i],   S2_d = sum_i o_id y_id ez_i
//   => G_d = tot_A - miss_d,A ;  S1_d = tot_ez - miss_d,ez
// update (per feature d, replicated):
//   mu'_d = (sy_d - w_d . S1_d) / n_d                                                   (PPCA.hs:123)
//   r_d   = S2_d - mu'_d S1_d ;  w'_d = G_d^-1 r_d                                       (PPCA.hs:133-137)
//   s2'   = sum_d (syy_d - 2 mu'_d sy_d + mu'_d^2 n_d - 2 w'_d . r_d + w'_d^T G_d w'_d) / #known   (PPCA.hs:140-147)
// LL (PPCA.hs:149-159) and restore (PPCA.hs:161-168) are separate final passes.
#include "ppca.cuh"
#include "tma.cuh"

namespace gplvm {
int moments_pass(gplvm_ppca_session* s, int mode);

namespace {

constexpr int EW = 8;  // warps per CTA in the per-observation kernels

// In-place lower Cholesky of an n x n matrix in shared memory by ONE warp.  Returns false on a non
// positive pivot.
__device__ __forceinline__ bool warp_cholesky(double* a, int n, int lane) {
  bool ok = true;
  for (int k = 0; k < n; ++k) {
    __syncwarp();
    const double akk = a[k * n + k];
    if (!(akk > 0.0)) ok = false;
    const double lkk = sqrt(akk);
    __syncwarp();
    for (int i = k + lane; i < n; i += 32) a[i * n + k] = (i == k) ? lkk : a[i * n + k] / lkk;
    __syncwarp();
    const int r = n - k - 1;
    for (int idx = lane; idx < r * r; idx += 32) {
      const int i = k + 1 + idx / r, j = k + 1 + idx % r;
      if (j <= i) a[i * n + j] -= a[i * n + k] * a[j * n + k];
    }
  }
  __syncwarp();
  return ok;
}

// x <- L^-1 x (forward substitution, column oriented; x in shared memory, all lanes participate)
__device__ __forceinline__ void warp_forward(const double* L, double* x, int n, int lane) {
  for (int k = 0; k < n; ++k) {
    __syncwarp();
    const double xk = x[k] / L[k * n + k];
    __syncwarp();
    if (lane == 0) x[k] = xk;
    for (int i = k + 1 + lane; i < n; i += 32) x[i] -= L[i * n + k] * xk;
  }
  __syncwarp();
}
// x <- L^-T x
__device__ __forceinline__ void warp_backward(const double* L, double* x, int n, int lane) {
  for (int k = n - 1; k >= 0; --k) {
    __syncwarp();
    const double xk = x[k] / L[k * n + k];
    __syncwarp();
    if (lane == 0) x[k] = xk;
    for (int i = lane; i < k; i += 32) x[i] -= L[k * n + i] * xk;
  }
  __syncwarp();
}

// Builds, for observation row x (global, D entries), the masked system of one observation:
//   Mi (shared, M x M) = C0 - sum_{missing} w w^T,   b (shared, M) = sum_{observed} w_d (x_d - mu_d)
// returns the number of observed features; yy = sum_{observed} (x_d - mu_d)^2.
__device__ __forceinline__ int masked_build(const double* __restrict__ x, const Par& p, double* Mi, double* b, int lane,
                                            double& yy) {
  const int D = p.D, M = p.M;
  for (int e = lane; e < M * M; e += 32) Mi[e] = p.Minv[e];  // C0 = s2 I + W^T W
  double bacc[GPLVM_MAX_M];
#pragma unroll
  for (int a = 0; a < GPLVM_MAX_M; ++a) bacc[a] = 0.0;
  double yacc = 0.0;
  int cnt = 0;
  __syncwarp();
  for (int d0 = 0; d0 < D; d0 += 32) {
    const int d = d0 + lane;
    const double v = (d < D) ? x[d] : 0.0;
    const bool miss = (d < D) && isnan(v);
    const unsigned mm = __ballot_sync(0xffffffffu, miss);
    const unsigned vm = __ballot_sync(0xffffffffu, d < D);
    cnt += __popc(vm & ~mm);
    if (d < D && !miss) {
      const double t = v - p.mu[d];
      yacc = fma(t, t, yacc);
      const double* w = p.W + (int64_t)d * M;
#pragma unroll
      for (int a = 0; a < GPLVM_MAX_M; ++a)
        if (a < M) bacc[a] = fma(w[a], t, bacc[a]);
    }
    unsigned rem = mm;
    while (rem) {
      const int bit = __ffs(rem) - 1;
      rem &= rem - 1;
      const double* w = p.W + (int64_t)(d0 + bit) * M;
      for (int e = lane; e < M * M; e += 32) Mi[e] = fma(-w[e / M], w[e % M], Mi[e]);
    }
  }
#pragma unroll
  for (int a = 0; a < GPLVM_MAX_M; ++a)
    if (a < M) {
      const double t = warp_sum(bacc[a]);
      if (lane == 0) b[a] = t;
    }
  yy = warp_sum(yacc);
  __syncwarp();
  return cnt;
}

// E-step: one warp per observation; writes ez_i (M) and vech(A_i) (NV) to HBM.
__global__ void __launch_bounds__(EW * 32) masked_estep_kernel(const double* __restrict__ X, int64_t ldx, int64_t n_obs, Par p,
                                                              double* __restrict__ EA, int eas) {
  if (p.scal[S_DONE] != 0.0) return;
  extern __shared__ double sm[];
  const int M = p.M, NV = p.NV;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double* Mi = sm + (size_t)w * (2 * M * M + 3 * M);
  double* Inv = Mi + M * M;
  double* b = Inv + M * M;
  const double s2 = p.scal[S_SIGMA2];
  const int64_t nw = (int64_t)gridDim.x * EW;
  for (int64_t n = (int64_t)blockIdx.x * EW + w; n < n_obs; n += nw) {
    double yy;
    const int cnt = masked_build(X + n * ldx, p, Mi, b, lane, yy);
    if (cnt == 0) {
      if (lane == 0) set_err(p.scal, 2.0);
      for (int e = lane; e < eas; e += 32) EA[n * eas + e] = 0.0;
      continue;
    }  // all-NaN observation: unsupported (Util.hs:81)
    const bool ok = warp_cholesky(Mi, M, lane);
    if (!ok && lane == 0) set_err(p.scal, 1.0);
    // ez = (L L^T)^-1 b
    warp_forward(Mi, b, M, lane);
    warp_backward(Mi, b, M, lane);
    // Inv = (L L^T)^-1, one lane per column
    for (int j = lane; j < M; j += 32) {
      for (int i = 0; i < M; ++i) {
        double s = (i == j) ? 1.0 : 0.0;
        for (int k = j; k < i; ++k) s -= Mi[i * M + k] * Inv[k * M + j];
        Inv[i * M + j] = (i < j) ? 0.0 : s / Mi[i * M + i];
      }
      for (int i = M - 1; i >= 0; --i) {
        double s = Inv[i * M + j];
        for (int k = i + 1; k < M; ++k) s -= Mi[k * M + i] * Inv[k * M + j];
        Inv[i * M + j] = s / Mi[i * M + i];
      }
    }
    __syncwarp();
    for (int a = lane; a < M; a += 32) EA[n * eas + NV + a] = b[a];
    // vech (row-wise lower triangle): index(a, c) = a (a+1)/2 + c, c <= a
    for (int e = lane; e < NV; e += 32) {
      int a = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
      while ((a + 1) * (a + 2) / 2 <= e) ++a;
      while (a * (a + 1) / 2 > e) --a;
      const int c = e - a * (a + 1) / 2;
      EA[n * eas + e] = fma(b[a], b[c], s2 * Inv[a * M + c]);
    }
    __syncwarp();
  }
}

// mbits[n][word] bit j = feature 32 word + j of observation n is missing (NaN); constant over the EM iterations
__global__ void __launch_bounds__(256) masked_bits_kernel(const double* __restrict__ X, int64_t ldx, int64_t n_obs, int D, int wpr,
                                                          uint32_t* __restrict__ mbits) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= n_obs) return;
  for (int wd = 0; wd < wpr; ++wd) {
    const int d = wd * 32 + lane;
    const bool miss = (d < D) && isnan(X[n * ldx + d]);
    const unsigned m = __ballot_sync(0xffffffffu, miss);
    if (lane == 0) mbits[n * wpr + wd] = m;
  }
}

// Complement sums.  grid = (row blocks, feature groups); 8 warps per CTA; warp w of feature group fg owns the FPW
// features f0 = (8 fg + w) FPW ... and keeps miss_d[e] for them in REGISTERS: lane l owns e = l + 32 k, k < KPL.
// Tiles of TR observations ([vech A_i | ez_i] rows are contiguous in HBM, so a tile is ONE 1-D bulk TMA copy, and
// so is its block of mask words) are staged in shared memory through a 3-deep mbarrier pipeline; every warp reads
// each observation's slice once and adds it to the accumulators of the features that observation misses
// (warp-uniform tests on the packed mask bits) -- work ~ #missing, fixed order (rows ascending) => deterministic,
// no atomics.  Warp 0 of feature group 0 also accumulates the totals.
// partials: [row block][(D + 1) x eas]  (row D = totals)
constexpr int MS_STAGES = 3;

// acc[f][:] += a[:] for every feature f whose bit is set in m.  Written as a set-bit loop around a switch so that
// the compiler keeps REAL (warp-uniform) branches: an if-converted version executes all FPW x KPL additions per
// observation (dense work); this one executes ~ #missing x KPL.
template <int KPL, int FPW>
__device__ __forceinline__ void add_missing(double (&acc)[FPW][KPL], const double (&a)[KPL], unsigned m) {
#define GP_ADD_CASE(F)                                   \
  case F:                                                \
    if (F < FPW) {                                       \
      _Pragma("unroll") for (int k = 0; k < KPL; ++k) acc[F < FPW ? F : 0][k] += a[k]; \
    }                                                    \
    break;
  while (m) {
    const int f = __ffs(m) - 1;
    m &= m - 1;
    switch (f) {
      GP_ADD_CASE(0) GP_ADD_CASE(1) GP_ADD_CASE(2) GP_ADD_CASE(3) GP_ADD_CASE(4) GP_ADD_CASE(5) GP_ADD_CASE(6) GP_ADD_CASE(7)
      GP_ADD_CASE(8) GP_ADD_CASE(9) GP_ADD_CASE(10) GP_ADD_CASE(11) GP_ADD_CASE(12) GP_ADD_CASE(13) GP_ADD_CASE(14)
      GP_ADD_CASE(15)
      default: break;
    }
  }
#undef GP_ADD_CASE
}


template <int KPL, int FPW, int MS_TR, int MINB>
__global__ void __launch_bounds__(256, MINB) masked_miss_sums_kernel(const double* __restrict__ EA, int eas, const uint32_t* __restrict__ mbits,
                                                               int wpr, int64_t n_obs, int64_t rows_per_cta, int D,
                                                               double* __restrict__ partials, int64_t Spart,
                                                               const double* __restrict__ scal) {
  if (scal[S_DONE] != 0.0) return;
  extern __shared__ __align__(128) uint8_t ms_smem[];
  const int ea_bytes = (MS_TR * eas * 8 + 15) & ~15;
  const int mk_bytes = (MS_TR * wpr * 4 + 15) & ~15;
  const int stage_bytes = ea_bytes + mk_bytes;
  const uint32_t sm_u = smem_u32(ms_smem);
  const uint32_t bar_u = sm_u + MS_STAGES * stage_bytes;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int f0 = ((int)blockIdx.y * 8 + w) * FPW;
  const bool totals = (blockIdx.y == 0 && w == 0);
  const bool has_feat = f0 < D;
  double acc[FPW][KPL], tot[KPL];
#pragma unroll
  for (int f = 0; f < FPW; ++f)
#pragma unroll
    for (int k = 0; k < KPL; ++k) acc[f][k] = 0.0;
#pragma unroll
  for (int k = 0; k < KPL; ++k) tot[k] = 0.0;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = min(n_obs, r0 + rows_per_cta);
  const int64_t ntiles = (r1 > r0) ? (r1 - r0 + MS_TR - 1) / MS_TR : 0;
  const int word = f0 >> 5, shift = f0 & 31;
  const unsigned fmask = (FPW >= 32) ? 0xffffffffu : ((1u << FPW) - 1u);

  if (tid == 0) {
    for (int s = 0; s < MS_STAGES; ++s) mbar_init(bar_u + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int64_t i) {  // thread 0: stage i % STAGES <- tile i (mask block copied whole: mbits is padded by MS_TR rows)
    const int s = (int)(i % MS_STAGES);
    const int64_t ra = r0 + i * MS_TR;
    const int rows = (int)min((int64_t)MS_TR, r1 - ra);
    const uint32_t dst = sm_u + s * stage_bytes;
    const uint32_t eb = (uint32_t)((rows * eas * 8 + 15) & ~15);  // EA is padded: the rounded-up tail stays inside it
    mbar_expect_tx(bar_u + 8 * s, eb + (uint32_t)mk_bytes);
    bulk_load_1d(dst, EA + ra * eas, eb, bar_u + 8 * s);
    bulk_load_1d(dst + ea_bytes, mbits + ra * wpr, (uint32_t)mk_bytes, bar_u + 8 * s);
  };
  if (tid == 0)
    for (int64_t i = 0; i < MS_STAGES && i < ntiles; ++i) issue(i);

  for (int64_t i = 0; i < ntiles; ++i) {
    const int s = (int)(i % MS_STAGES);
    const int rows = (int)min((int64_t)MS_TR, r1 - (r0 + i * MS_TR));
    mbar_wait(bar_u + 8 * s, (uint32_t)((i / MS_STAGES) & 1));
    const double* tile = reinterpret_cast<const double*>(ms_smem + s * stage_bytes);
    const uint32_t* mk = reinterpret_cast<const uint32_t*>(ms_smem + s * stage_bytes + ea_bytes);
    if (has_feat || totals) {
      int r = 0;
      for (; r + 1 < rows; r += 2) {  // two observations in flight
        double a0[KPL], a1[KPL];
#pragma unroll
        for (int k = 0; k < KPL; ++k) {
          const int e = lane + 32 * k;
          a0[k] = (e < eas) ? tile[r * eas + e] : 0.0;
          a1[k] = (e < eas) ? tile[(r + 1) * eas + e] : 0.0;
        }
        const unsigned m0 = has_feat ? ((mk[r * wpr + word] >> shift) & fmask) : 0u;
        const unsigned m1 = has_feat ? ((mk[(r + 1) * wpr + word] >> shift) & fmask) : 0u;
        if (totals) {
#pragma unroll
          for (int k = 0; k < KPL; ++k) tot[k] += a0[k];
#pragma unroll
          for (int k = 0; k < KPL; ++k) tot[k] += a1[k];
        }
        add_missing<KPL, FPW>(acc, a0, m0);
        add_missing<KPL, FPW>(acc, a1, m1);
      }
      if (r < rows) {
        double a0[KPL];
#pragma unroll
        for (int k = 0; k < KPL; ++k) {
          const int e = lane + 32 * k;
          a0[k] = (e < eas) ? tile[r * eas + e] : 0.0;
        }
        const unsigned m0 = has_feat ? ((mk[r * wpr + word] >> shift) & fmask) : 0u;
        if (totals) {
#pragma unroll
          for (int k = 0; k < KPL; ++k) tot[k] += a0[k];
        }
        add_missing<KPL, FPW>(acc, a0, m0);
      }
    }
    __syncthreads();  // every warp is done with this stage
    if (tid == 0 && i + MS_STAGES < ntiles) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(i + MS_STAGES);
    }
  }
  double* out = partials + (int64_t)blockIdx.x * Spart;
  if (has_feat) {
#pragma unroll
    for (int f = 0; f < FPW; ++f) {
      if (f0 + f >= D) break;
#pragma unroll
      for (int k = 0; k < KPL; ++k) {
        const int e = lane + 32 * k;
        if (e < eas) out[(int64_t)(f0 + f) * eas + e] = acc[f][k];
      }
    }
  }
  if (totals) {
#pragma unroll
    for (int k = 0; k < KPL; ++k) {
      const int e = lane + 32 * k;
      if (e < eas) out[(int64_t)D * eas + e] = tot[k];
    }
  }
}

// per-feature update, one warp per feature.  stats layout (all-reduced):
//   [miss_d = (vech G^miss_d | S1^miss_d) : D x eas] [tot = (vech sum A | sum ez) : eas] [S2_d : D x M]
// update of one feature by one warp; leaves the new w_d in the warp's `r` slot of shared memory (sm + 2 M^2)
__device__ __forceinline__ void masked_update_feature(const Par& p, const double* __restrict__ stats, double* __restrict__ scratch,
                                                      double* G, int d, int lane) {
  const int M = p.M, NV = p.NV, D = p.D;
  double* G0 = G + M * M;
  double* r = G0 + M * M;
  double* wold = r + M;
  const int eas = NV + M;
  const double* miss = stats + (int64_t)d * eas;
  const double* tot = stats + (int64_t)D * eas;
  const double* S2 = stats + (int64_t)(D + 1) * eas + (int64_t)d * M;
  double* S1 = wold + M;  // M doubles of per-warp scratch (see warp_scratch_bytes)
  for (int e = lane; e < M * M; e += 32) {
    const int a = e / M, c = e % M;
    const int v_idx = (c <= a) ? a * (a + 1) / 2 + c : c * (c + 1) / 2 + a;
    const double v = tot[v_idx] - miss[v_idx];
    G[e] = v;
    G0[e] = v;
  }
  for (int a = lane; a < M; a += 32) S1[a] = tot[NV + a] - miss[NV + a];
  __syncwarp();
  double dot = 0.0;
  for (int a = lane; a < M; a += 32) {
    wold[a] = p.W[(int64_t)d * M + a];
    dot = fma(wold[a], S1[a], dot);
  }
  dot = warp_sum(dot);
  const double nd = p.nd[d], sy = p.sy[d], syy = p.syy[d];
  const double mu_new = (sy - dot) / nd;
  for (int a = lane; a < M; a += 32) r[a] = S2[a] - mu_new * S1[a];
  __syncwarp();
  double wr = 0.0;  // w' . r needs r before the solve overwrites it
  double rsave[(GPLVM_MAX_M + 31) / 32];
#pragma unroll
  for (int j = 0; j < (GPLVM_MAX_M + 31) / 32; ++j) rsave[j] = (lane + 32 * j < M) ? r[lane + 32 * j] : 0.0;
  const bool ok = warp_cholesky(G, M, lane);
  warp_forward(G, r, M, lane);
  warp_backward(G, r, M, lane);  // r now holds w'_d
  double quad = 0.0, mdw = 0.0;
#pragma unroll
  for (int j = 0; j < (GPLVM_MAX_M + 31) / 32; ++j) {
    const int a = lane + 32 * j;
    if (a < M) {
      wr = fma(r[a], rsave[j], wr);
      double t = 0.0;
      for (int c = 0; c < M; ++c) t = fma(G0[a * M + c], r[c], t);
      quad = fma(r[a], t, quad);
      mdw = fmax(mdw, fabs(wold[a] - r[a]));
      p.W[(int64_t)d * M + a] = r[a];
    }
  }
  wr = warp_sum(wr);
  quad = warp_sum(quad);
  mdw = warp_max(mdw);
  if (lane == 0) {
    const double cd = syy - 2.0 * mu_new * sy + mu_new * mu_new * nd;
    scratch[d] = cd - 2.0 * wr + quad;
    scratch[D + d] = mdw;
    p.mu[d] = mu_new;
    if (!ok) set_err(p.scal, 1.0);
  }
}


// wtw_part[blockIdx.x][M x M] receives the block's share of W'^T W' (summed in fixed order by the finalize kernel).
__global__ void __launch_bounds__(EW * 32) masked_update_kernel(Par p, const double* __restrict__ stats, double* __restrict__ scratch,
                                                               double* __restrict__ wtw_part) {
  if (p.scal[S_DONE] != 0.0) return;
  extern __shared__ double sm[];
  const int M = p.M, NV = p.NV, D = p.D;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int d = blockIdx.x * EW + w;
  if (d < D) masked_update_feature(p, stats, scratch, sm + (size_t)w * (2 * M * M + 3 * M), d, lane);
  __syncthreads();
  for (int e = threadIdx.x; e < M * M; e += blockDim.x) {
    const int a = e / M, c = e % M;
    double t = 0.0;
    for (int w2 = 0; w2 < EW; ++w2) {
      if (blockIdx.x * EW + w2 >= D) break;
      const double* wn = sm + (size_t)w2 * (2 * M * M + 3 * M) + 2 * M * M;   // this warp's new w_d (the solved r)
      t = fma(wn[a], wn[c], t);
    }
    wtw_part[(size_t)blockIdx.x * M * M + e] = t;
  }
}

// C0 = s2 I + W^T W into p.Minv (single CTA)
__device__ void masked_prep_block(const Par& p) {
  const int D = p.D, M = p.M;
  const double s2 = p.scal[S_SIGMA2];
  for (int e = threadIdx.x; e < M * M; e += blockDim.x) {
    const int a = e / M, b = e % M;
    double t = 0.0;
    for (int d = 0; d < D; ++d) t = fma(p.W[(int64_t)d * M + a], p.W[(int64_t)d * M + b], t);
    p.Minv[e] = t + ((a == b) ? s2 : 0.0);
  }
  __syncthreads();
}
__global__ void __launch_bounds__(256) masked_prep_kernel(Par p) { masked_prep_block(p); }

__global__ void __launch_bounds__(256) masked_finalize_kernel(Par p, const double* __restrict__ stats, const double* __restrict__ scratch,
                                                              const double* __restrict__ wtw_part, int nblocks) {
  if (p.scal[S_DONE] != 0.0) return;
  __shared__ double red[40];
  const int D = p.D, M = p.M, NV = p.NV;
  double a = 0.0, m = 0.0;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    a += scratch[d];
    m = fmax(m, scratch[D + d]);
  }
  a = block_sum(a, red);
  m = block_max(m, red);
  // mean of the E-step latent means, needed by the restore (meanColumn expX^T, PPCA.hs:162)
  const double* sez = stats + (int64_t)D * (NV + M) + NV;
  for (int b = threadIdx.x; b < M; b += blockDim.x) p.c[b] = sez[b] / (double)p.N;
  __syncthreads();
  if (threadIdx.x == 0) {
    const double s2 = p.scal[S_SIGMA2];
    const double news2 = a / p.scal[S_KNOWN];
    const double dvar = fabs(news2 - s2);
    p.scal[S_SIGMA2] = news2;
    p.scal[S_DVAR] = dvar;
    p.scal[S_MAXDW] = m;
    p.scal[S_STEPS] += 1.0;
    if (!(news2 > 0.0)) set_err(p.scal, 1.0);
    if (p.scal[S_USETOL] != 0.0 && !(fmax(dvar, m) > p.scal[S_TOL])) p.scal[S_DONE] = 1.0;
  }
  __syncthreads();
  // C0 = s2' I + W'^T W' for the next E-step, from the update kernel's per-block partial Gram matrices (fixed order)
  const double news2 = p.scal[S_SIGMA2];
  for (int e = threadIdx.x; e < M * M; e += blockDim.x) {
    double t = 0.0;
    for (int b = 0; b < nblocks; ++b) t += wtw_part[(size_t)b * M * M + e];
    p.Minv[e] = t + ((e / M == e % M) ? news2 : 0.0);
  }
}

// data constants: stats = [n_d | sy_d | syy_d] (all-reduced)
__global__ void masked_set_constants_kernel(Par p, const double* __restrict__ stats) {
  const int D = p.D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    p.nd[d] = stats[d];
    p.sy[d] = stats[D + d];
    p.syy[d] = stats[2 * D + d];
    p.mu[d] = 0.0;  // initMu = 0 (PPCA.hs:103)
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double k = 0.0;
    for (int d = 0; d < D; ++d) k += stats[d];
    p.scal[S_KNOWN] = k;
  }
}

// per-observation log-likelihood terms with the CURRENT parameters (PPCA.hs:149-159 via A.4)
__global__ void __launch_bounds__(EW * 32) masked_ll_kernel(const double* __restrict__ X, int64_t ldx, int64_t n_obs, Par p,
                                                           double* __restrict__ partials) {
  extern __shared__ double sm[];
  __shared__ double wsum[EW];
  const int M = p.M;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double* Mi = sm + (size_t)w * (2 * M * M + 3 * M);
  double* b = Mi + 2 * M * M;
  const double s2 = p.scal[S_SIGMA2];
  const double log2pi = log(2.0 * M_PI), logs2 = log(s2);
  const int64_t nw = (int64_t)gridDim.x * EW;
  double acc = 0.0;
  for (int64_t n = (int64_t)blockIdx.x * EW + w; n < n_obs; n += nw) {
    double yy;
    const int cnt = masked_build(X + n * ldx, p, Mi, b, lane, yy);
    const bool ok = warp_cholesky(Mi, M, lane);
    if (!ok && lane == 0) set_err(p.scal, 1.0);
    warp_forward(Mi, b, M, lane);  // t = L^-1 W_i^T yc
    double tt = 0.0, ld = 0.0;
    for (int a = lane; a < M; a += 32) {
      tt = fma(b[a], b[a], tt);
      ld += log(Mi[a * M + a]);
    }
    tt = warp_sum(tt);
    ld = 2.0 * warp_sum(ld);
    acc += (double)cnt * log2pi + ((double)(cnt - M) * logs2 + ld) + (yy - tt) / s2;
    __syncwarp();
  }
  if (lane == 0) wsum[w] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < EW; ++i) t += wsum[i];
    partials[blockIdx.x] = t;
  }
}

__global__ void masked_ll_final_kernel(Par p, const double* __restrict__ stats) {
  if (threadIdx.x == 0) p.scal[S_LL] = -1.0 * stats[0] / 2.0;
}

// B = W (W^T W)^-1 (W^T W + s2 I)  (D x M) into Wprev;  fm = mu + W zbar (D) into scratch   (PPCA.hs:162-167)
__global__ void __launch_bounds__(256) masked_restore_prep_kernel(Par p, double* __restrict__ fm) {
  extern __shared__ double sm[];
  const int D = p.D, M = p.M;
  double* WtW = sm;            // M x M -> Cholesky
  double* Inv = sm + M * M;    // (W^T W)^-1
  double* F = sm + 2 * M * M;  // Inv * (WtW + s2 I)
  const double s2 = p.scal[S_SIGMA2];
  for (int e = threadIdx.x; e < M * M; e += blockDim.x) {
    const int a = e / M, b = e % M;
    WtW[e] = p.Minv[e] - ((a == b) ? s2 : 0.0);  // C0 - s2 I
  }
  __syncthreads();
  const bool ok = block_cholesky(WtW, M, M);
  block_chol_inverse(WtW, Inv, M, M);
  for (int e = threadIdx.x; e < M * M; e += blockDim.x) {
    const int a = e / M, b = e % M;
    double t = 0.0;
    for (int k = 0; k < M; ++k) t = fma(Inv[a * M + k], p.Minv[k * M + b], t);  // p.Minv = W^T W + s2 I
    F[e] = t;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < D * M; e += blockDim.x) {
    const int d = e / M, b = e % M;
    double t = 0.0;
    for (int a = 0; a < M; ++a) t = fma(p.W[(int64_t)d * M + a], F[a * M + b], t);
    p.Wprev[e] = t;
  }
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    double t = p.mu[d];
    for (int a = 0; a < M; ++a) t = fma(p.W[(int64_t)d * M + a], p.c[a], t);
    fm[d] = t;
  }
  if (threadIdx.x == 0 && !ok) set_err(p.scal, 1.0);
}

// out[d][i] = B_d . ez_i + fm_d for a chunk of observations; out is feature-major (D x ld)
__global__ void __launch_bounds__(256) masked_restore_kernel(Par p, const double* __restrict__ Ez, int ldez, const double* __restrict__ fm,
                                                            double* __restrict__ out, int64_t ld, int64_t n_chunk) {
  const int M = p.M, D = p.D;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double ez[GPLVM_MAX_M];
  if (i < n_chunk)
    for (int a = 0; a < M; ++a) ez[a] = Ez[i * ldez + a];
  for (int d = 0; d < D; ++d) {
    if (i < n_chunk) {
      double t = fm[d];
      for (int a = 0; a < M; ++a) t = fma(p.Wprev[(int64_t)d * M + a], ez[a], t);
      out[(int64_t)d * ld + i] = t;
    }
  }
}

size_t warp_scratch_bytes(int M) { return sizeof(double) * (size_t)EW * (2 * M * M + 3 * M); }

}  // namespace

int masked_init_constants(gplvm_ppca_session* s) {
  for (ShardState& sh : s->sh) {
    CUDA_TRY(cudaSetDevice(sh.dev));
    GP_LAUNCH(masked_bits_kernel, (unsigned)ceil_div(sh.n, 8), 256, 0, sh.st, sh.X, sh.ldx, sh.n, (int)s->D, s->WPR, sh.mbits);
    if (sh.mbitsT) GP_TRY(masked_utc_prepare(s, sh));
  }
  GP_TRY(moments_pass(s, 2));
  GP_TRY(session_allreduce_stats(s, 3 * s->D));
  for (ShardState& sh : s->sh) {
    CUDA_TRY(cudaSetDevice(sh.dev));
    GP_LAUNCH(masked_set_constants_kernel, 1, 256, 0, sh.st, sh.par, sh.stats);
    GP_LAUNCH(masked_prep_kernel, 1, 256, 0, sh.st, sh.par);
  }
  static bool attr_done = false;
  if (!attr_done) {
    // opt in to > 48 KB of dynamic shared memory for M = 32
    for (ShardState& sh : s->sh) {
      CUDA_TRY(cudaSetDevice(sh.dev));
      CUDA_TRY(cudaFuncSetAttribute(masked_estep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      CUDA_TRY(cudaFuncSetAttribute(masked_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      CUDA_TRY(cudaFuncSetAttribute(masked_ll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    }
  }
  return GPLVM_OK;
}

static int launch_miss_sums(gplvm_ppca_session* s, ShardState& sh) {
  const int D = (int)s->D, eas = s->EAS;
  const int64_t Spart = (int64_t)(D + 1) * eas;
  const bool small = eas <= 160;   // M <= 16: 5 values per lane, 8 features per warp; else 18 per lane, 2 features
  const int fpw = small ? 8 : 2;
  const int tr = small ? 16 : 8;   // observations per pipeline stage
  const int fgroups = (int)ceil_div(D, 8 * fpw);
  int rb = (int)std::min<int64_t>(std::max<int64_t>(1, (small ? 2 : 1) * sm_count(sh.dev) / fgroups), sh.nparts);
  const int64_t rows_per_cta = round_up(std::max<int64_t>(ceil_div(sh.n, rb), 4), 4);  // 16-byte aligned tile starts
  rb = (int)ceil_div(sh.n, rows_per_cta);
  const int smem = MS_STAGES * (((tr * eas * 8 + 15) & ~15) + ((tr * s->WPR * 4 + 15) & ~15)) + 64;
  static bool attr_done[64] = {};
  if (!attr_done[sh.dev & 63]) {
    CUDA_TRY(cudaFuncSetAttribute((masked_miss_sums_kernel<5, 8, 16, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CUDA_TRY(cudaFuncSetAttribute((masked_miss_sums_kernel<18, 2, 8, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done[sh.dev & 63] = true;
  }
  if (smem > 200 * 1024) return fail(GPLVM_ERR_ARG, "missing-data PPCA: M too large for the statistics kernel");
  dim3 grid(rb, fgroups);
  if (small)
    GP_LAUNCH((masked_miss_sums_kernel<5, 8, 16, 2>), grid, 256, smem, sh.st, sh.EA, eas, sh.mbits, s->WPR, sh.n, rows_per_cta, D,
              sh.partials, Spart, sh.par.scal);
  else
    GP_LAUNCH((masked_miss_sums_kernel<18, 2, 8, 1>), grid, 256, smem, sh.st, sh.EA, eas, sh.mbits, s->WPR, sh.n, rows_per_cta, D,
              sh.partials, Spart, sh.par.scal);
  return launch_reduce_partials(sh, Spart, rb, sh.stats);
}

int masked_enqueue_step(gplvm_ppca_session* s) {
  const int M = (int)s->M, D = (int)s->D;
  const size_t smem = warp_scratch_bytes(M);
  for (ShardState& sh : s->sh) {
    CUDA_TRY(cudaSetDevice(sh.dev));
    if (sh.evk0) CUDA_TRY(cudaEventRecord(sh.evk0, sh.st));
    seg_mark(sh, 0);
    double* s2dst = sh.stats + (int64_t)(D + 1) * s->EAS;
    if (s->use_dmma) {
      // fused path (M = 16): b_i by TMA + DMMA, per-observation register solver, sparse complement sums, S2 by DMMA
      GP_TRY(masked_dmma_b(s, sh, sh.par.W, sh.par.mu, sh.Ez, nullptr));   // yy_i is only needed by the likelihood pass
      seg_mark(sh, 1);
      GP_TRY(masked_fused_solve(s, sh, sh.Ez, sh.yy, false, nullptr));
      seg_mark(sh, 2);
      if (D >= 128) GP_TRY(masked_fused_miss_sums(s, sh, sh.stats));
      else GP_TRY(launch_miss_sums(s, sh));
      seg_mark(sh, 3);
      GP_TRY(masked_dmma_s2(s, sh, sh.EA + s->NV, s->EAS, s2dst));
      seg_mark(sh, 4);
    } else {
      const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(sh.n, EW), 148 * 8);
      GP_LAUNCH(masked_estep_kernel, grid, EW * 32, smem, sh.st, sh.X, sh.ldx, sh.n, sh.par, sh.EA, s->EAS);
      GP_TRY(launch_miss_sums(s, sh));
      // S2_d = sum_i o_id y_id ez_i  (D x M)
      GP_TRY(launch_stats_gemm(sh, ROWS_MASKED_VALUES, D, D, D, M, M, sh.EA + s->NV, s->EAS, M, nullptr, 0, (int64_t)D * M, s2dst));
    }
    if (sh.evk1) CUDA_TRY(cudaEventRecord(sh.evk1, sh.st));
  }
  // one combination of the packed statistics per step: over NVLink peer memory when the mappings exist, else NCCL
  if (s->p2p) GP_TRY(session_peer_allreduce_stats(s));
  else GP_TRY(session_allreduce_stats(s, s->S));
  for (ShardState& sh : s->sh) {
    CUDA_TRY(cudaSetDevice(sh.dev));
    if (s->use_dmma) seg_mark(sh, 5);
    double* wtw_part = sh.scratch + (size_t)8 * D;   // ceil(D / EW) x M x M (session_create sizes the scratch block)
    GP_LAUNCH(masked_update_kernel, (unsigned)ceil_div(D, EW), EW * 32, smem, sh.st, sh.par, sh.stats, sh.scratch, wtw_part);
    GP_LAUNCH(masked_finalize_kernel, 1, 256, 0, sh.st, sh.par, sh.stats, sh.scratch, wtw_part, (int)ceil_div(D, EW));
    if (s->use_dmma) seg_mark(sh, 6);
  }
  return GPLVM_OK;
}

int masked_loglik(gplvm_ppca_session* s, double* ll) {
  const size_t smem = warp_scratch_bytes((int)s->M);
  for (ShardState& sh : s->sh) {
    CUDA_TRY(cudaSetDevice(sh.dev));
    int grid = 0;
    if (s->use_dmma) {
      GP_TRY(masked_dmma_b(s, sh, sh.par.W, sh.par.mu, sh.Ez, sh.yy));
      GP_TRY(masked_fused_solve(s, sh, sh.Ez, sh.yy, true, &grid));
    } else {
      grid = (int)std::min<int64_t>(ceil_div(sh.n, EW), (int64_t)sh.nparts);
      GP_LAUNCH(masked_ll_kernel, (unsigned)grid, EW * 32, smem, sh.st, sh.X, sh.ldx, sh.n, sh.par, sh.partials);
    }
    // partials are 1 double per CTA: reduce with S = 1
    GP_TRY(launch_reduce_partials(sh, 1, grid));
  }
  GP_TRY(session_allreduce_stats(s, 1));
  for (ShardState& sh : s->sh) {
    CUDA_TRY(cudaSetDevice(sh.dev));
    GP_LAUNCH(masked_ll_final_kernel, 1, 32, 0, sh.st, sh.par, sh.stats);
  }
  ShardState& s0 = s->sh[0];
  CUDA_TRY(cudaSetDevice(s0.dev));
  double scal[S_COUNT];
  CUDA_TRY(cudaMemcpyAsync(scal, s0.par.scal, sizeof scal, cudaMemcpyDeviceToHost, s0.st));
  CUDA_TRY(cudaStreamSynchronize(s0.st));
  if (scal[S_ERR] == 2.0) return fail(GPLVM_ERR_UNSUPPORTED, "missing-data PPCA: an observation has every feature missing");
  if (scal[S_ERR] != 0.0) return fail(GPLVM_ERR_NUMERIC, "missing-data PPCA: matrix not positive definite");
  *ll = scal[S_LL];
  return GPLVM_OK;
}

int masked_restore(gplvm_ppca_session* s, double* restored, int64_t ld) {
  const int64_t D = s->D;
  const int M = (int)s->M;
  const int64_t first = s->sh[0].n0;
  for (ShardState& sh : s->sh) {
    CUDA_TRY(cudaSetDevice(sh.dev));
    double* fm = sh.scratch + 2 * D;
    GP_LAUNCH(masked_restore_prep_kernel, 1, 256, sizeof(double) * 3 * M * M, sh.st, sh.par, fm);
    GP_TRY(session_d2h_rows(sh, D, sh.n, restored + (sh.n0 - first), ld, (int)s->sh.size(),
                            [&](int64_t c0, int64_t cn, double* stage, int64_t pitch) -> int {
                              GP_LAUNCH(masked_restore_kernel, (unsigned)ceil_div(cn, 256), 256, 0, sh.st, sh.par,
                                        sh.EA + c0 * s->EAS + s->NV, s->EAS, fm, stage, pitch, cn);
                              return GPLVM_OK;
                            }));
  }
  return GPLVM_OK;
}

}  // namespace gplvm
