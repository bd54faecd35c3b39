// masked_fused.cu -- batched per-observation solver of the NaN-masked E-step for M = 16 (north-star item 2):
// every observation applies its own NaN mask, builds and factors its own 16 x 16 system in REGISTERS with
// warp shuffles, and emits E[z_i] and A_i = E[z_i] E[z_i]^T + sigma^2 M_i^-1.
//
// Reference: emStepsMissed.stepOne, src/GPLVM/PPCA.hs:106-117 (M_i = sigma^2 I + W_i^T W_i, invMP, expXi) and the
// per-observation log-likelihood terms of PPCA.hs:149-159 (through the determinant lemma).
//
// Mapping: one HALF-WARP per observation, lane c (0..15) owns column c of the symmetric 16 x 16 matrix in 16
// registers.
//   build    : M_i = C0 - sum_{d missing} w_d w_d^T  (complement form: work ~ #missing); the missing features come
//              from the packed mask bits, w_d from shared memory (W is D x 16, 32 KB at D = 256).
//   factor   : right-looking Cholesky on the full symmetric column set (lane c reads L[c][k] = M[k][c]/L[k][k] from
//              its OWN registers by symmetry; only the pivot and column k travel by __shfl_sync); reciprocal
//              square roots by rsqrt.approx.f64 + 2 Newton steps.
//   invert   : T = L^-1 by forward substitution (column c in lane c), M_i^-1 = T^T T, ez = M_i^-1 b.
// b_i (and |y_c|^2 for the likelihood) come from the TMA + DMMA pass in dense_dmma.cu (MODE_MASKED_B).
#include "ppca.cuh"
#include "tma.cuh"

namespace gplvm {
namespace {

constexpr int MF = 16;
constexpr int NVF = MF * (MF + 1) / 2;   // 136
constexpr int EASF = NVF + MF;           // 152

__device__ __forceinline__ double rsqrt_nr(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const double e = fma(-x * y, y, 1.0);
    y = fma(0.5 * y, e, y);
  }
  return y;
}

// LL = false: writes EA[n] = [vech A_i | ez_i].   LL = true: accumulates the per-observation likelihood terms
//   cnt log 2pi + (cnt - M) log s2 + log det M_i + (yy_i - b_i . ez_i) / s2      into partials[blockIdx.x].
template <int DD, bool LL>
__global__ void __launch_bounds__(256, 2) masked_solve_kernel(const double* __restrict__ bsrc, const double* __restrict__ yy,
                                                             const uint32_t* __restrict__ mbits, int64_t n_obs, Par p,
                                                             double* __restrict__ EA, double* __restrict__ partials) {
  constexpr int WPRT = DD / 32;
  if (!LL && p.scal[S_DONE] != 0.0) return;
  extern __shared__ double sm[];
  double* Ws = sm;                 // DD x 16
  double* C0s = sm + DD * MF;      // 16 x 16 : s2 I + W^T W  (kept in p.Minv by the masked prep kernel)
  __shared__ double red[40];
  for (int e = threadIdx.x; e < DD * MF; e += blockDim.x) Ws[e] = p.W[e];
  for (int e = threadIdx.x; e < MF * MF; e += blockDim.x) C0s[e] = p.Minv[e];
  __syncthreads();
  const double s2 = p.scal[S_SIGMA2];
  const int lane = threadIdx.x & 31, c = lane & 15;
  constexpr unsigned hmask = 0xffffffffu;  // both half-warps run every shuffle together (width 16 keeps them apart)
  const int64_t pair0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2;  // warp-uniform
  const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 5) * 2;
  const uint32_t ws_u = smem_u32(Ws);
  double llacc = 0.0;
  const double log2pi = 1.8378770664093453, logs2 = LL ? log(s2) : 0.0;

  for (int64_t nb = pair0; nb < n_obs; nb += stride) {
    const bool valid = nb + (lane >> 4) < n_obs;              // an odd tail leaves the upper half-warp idle:
    const int64_t n = valid ? nb + (lane >> 4) : n_obs - 1;   // it recomputes the last observation and stores nothing
    // ---- build M_i -----------------------------------------------------------------------------------------
    double m[MF];
#pragma unroll
    for (int a = 0; a < MF; ++a) m[a] = C0s[a * MF + c];
    const uint32_t mw = (c < WPRT) ? mbits[n * WPRT + c] : 0u;
    const double bc = bsrc[n * MF + c];
    int nmiss = 0;
#pragma unroll
    for (int wd = 0; wd < WPRT; ++wd) {
      uint32_t bits = __shfl_sync(hmask, mw, wd, 16);
      nmiss += __popc(bits);
      while (bits) {
        const int j = __ffs(bits) - 1;
        bits &= bits - 1;
        const uint32_t wr = ws_u + (uint32_t)((wd * 32 + j) * (MF * 8));
        double wv[MF];
#pragma unroll
        for (int q = 0; q < MF / 2; ++q) lds128(wr + q * 16, wv[2 * q], wv[2 * q + 1]);
        const double wc = lds64(wr + c * 8);
#pragma unroll
        for (int a = 0; a < MF; ++a) m[a] = fma(-wv[a], wc, m[a]);
      }
      __syncwarp();
    }
    // every feature missing: unsupported by the reference (Util.hs:81); flagged, outputs zeroed below
    const bool empty = nmiss >= DD;
    if (empty && c == 0) set_err(p.scal, 2.0);
    // ---- Cholesky: lc[a] = L[a][c] (a > c), lc[c] = 1 / L[c][c] -------------------------------------------
    double lc[MF];
    bool ok = true;
    double yprod = 1.0;
#pragma unroll
    for (int k = 0; k < MF; ++k) {
      const double pk = __shfl_sync(hmask, m[k], k, 16);
      ok = ok && (pk > 0.0);
      const double y = rsqrt_nr(pk);
      if (LL) yprod *= y;
      const double lck = m[k] * y;
      if (c == k) lc[k] = y;
#pragma unroll
      for (int a = k + 1; a < MF; ++a) {
        const double sa = m[a] * y;
        if (c == k) lc[a] = sa;
        const double lak = __shfl_sync(hmask, sa, k, 16);
        m[a] = fma(-lak, lck, m[a]);
      }
    }
    if (!ok && c == 0) set_err(p.scal, 1.0);
    // ---- T = L^-1 (column c in lane c) ---------------------------------------------------------------------
    double tt[MF];
#pragma unroll
    for (int i = 0; i < MF; ++i) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < i; ++k) {
        const double lik = __shfl_sync(hmask, lc[i], k, 16);
        s = fma(lik, tt[k], s);
      }
      const double di = __shfl_sync(hmask, lc[i], i, 16);
      tt[i] = (i == c) ? di : -s * di;
    }
    // ---- M_i^-1 = T^T T (column c), ez = M_i^-1 b -----------------------------------------------------------
    double mi[MF];
#pragma unroll
    for (int a = 0; a < MF; ++a) {
      double s = 0.0;
#pragma unroll
      for (int i = a; i < MF; ++i) {
        const double tia = __shfl_sync(hmask, tt[i], a, 16);
        s = fma(tia, tt[i], s);
      }
      mi[a] = s;
    }
    double ez = 0.0;
#pragma unroll
    for (int a = 0; a < MF; ++a) ez = fma(mi[a], __shfl_sync(hmask, bc, a, 16), ez);

    if (!LL) {
      double* out = EA + n * EASF;
#pragma unroll
      for (int a = 0; a < MF; ++a) {
        const double ea = __shfl_sync(hmask, ez, a, 16);
        const double v = empty ? 0.0 : fma(ea, ez, s2 * mi[a]);
        if (valid && a >= c) out[a * (a + 1) / 2 + c] = v;
      }
      if (valid) out[NVF + c] = empty ? 0.0 : ez;
    } else {
      double dot = bc * ez;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) dot += __shfl_xor_sync(hmask, dot, o, 16);
      if (c == 0 && valid && !empty) {
        const double cnt = (double)(DD - nmiss);
        const double logdetM = -2.0 * log(yprod);
        llacc += cnt * log2pi + ((cnt - (double)MF) * logs2 + logdetM) + (yy[n] - dot) / s2;
      }
    }
  }
  if (LL) {
    const double tsum = block_sum(llacc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = tsum;
  }
}

template <int DD, bool LL>
int launch_solve(ShardState& sh, const double* bsrc, const double* yy, int grid) {
  const int smem = (DD * MF + MF * MF) * (int)sizeof(double);
  static bool attr_done[64] = {};
  if (!attr_done[sh.dev & 63]) {
    CUDA_TRY(cudaFuncSetAttribute((masked_solve_kernel<DD, LL>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done[sh.dev & 63] = true;
  }
  GP_LAUNCH((masked_solve_kernel<DD, LL>), grid, 256, smem, sh.st, bsrc, yy, sh.mbits, sh.n, sh.par, sh.EA, sh.partials);
  return GPLVM_OK;
}

}  // namespace

// E-step solve (ll = false) or likelihood terms (ll = true; per-CTA sums land in sh.partials[0 .. *grid_out))
int masked_fused_solve(gplvm_ppca_session* s, ShardState& sh, const double* bsrc, const double* yy, bool ll, int* grid_out) {
  int grid = (int)std::min<int64_t>(2 * sm_count(sh.dev), ceil_div(sh.n, 16));
  if (grid > sh.nparts) grid = sh.nparts;
  if (grid_out) *grid_out = grid;
  switch (s->D) {
    case 64: return ll ? launch_solve<64, true>(sh, bsrc, yy, grid) : launch_solve<64, false>(sh, bsrc, yy, grid);
    case 128: return ll ? launch_solve<128, true>(sh, bsrc, yy, grid) : launch_solve<128, false>(sh, bsrc, yy, grid);
    case 256: return ll ? launch_solve<256, true>(sh, bsrc, yy, grid) : launch_solve<256, false>(sh, bsrc, yy, grid);
    default: return fail(GPLVM_ERR_STATE, "fused masked solver: unsupported D");
  }
}

}  // namespace gplvm
