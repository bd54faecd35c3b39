// dense.cu -- dense PPCA EM step (emStepsFast, reference src/GPLVM/PPCA.hs:54-94) on one shard of
// observations: generic (any D, M <= GPLVM_MAX_M) kernels, the replicated parameter update and the
// log-likelihood.  The fused FP64 tensor-core statistics kernel for the benchmark shape lives in
// dense_dmma.cu; everything else (reduction of partials, update, LL) is shared.
//
// Single-pass formulation.  The resident observations are centred ONCE, in place
// (Xc = X - feature mean, PPCA.hs:61; the reference re-evaluates the delayed subtraction in every product):
//   prep   : Minv = (s2 I + W^T W)^-1,  A = W Minv                                   (PPCA.hs:65-66)
//   stats  : ez_n = xc_n^T A  (= Minv W^T xc_n, PPCA.hs:67);
//            XEz = sum_n xc_n ez_n^T, EE = sum_n ez_n ez_n^T                         (PPCA.hs:68-69)
//   update : Ezz = N s2 Minv + EE; W' = XEz Ezz^-1                                   (PPCA.hs:69,74)
//            s2' = (|Xc|^2 - 2 tr(W'^T XEz) + |W' U^T|^2) / (N D), U = chol(Ezz)     (PPCA.hs:75-80)
//   loglik : PPCA.hs:83-89 through the determinant lemma / Woodbury (no D x D matrix).
#include "ppca.cuh"

namespace gplvm {

// ------------------------------------------------------------------------------------------------
// generic split-K statistics GEMM:  out[r][q] = sum_n row(n, r) * col(n, q)
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int TG_R = 64, TG_Q = 64, TG_K = 16;  // tile: 64 rows x 64 cols, 16 observations per chunk

template <int KIND>
__device__ __forceinline__ double row_value(const double* __restrict__ X, int64_t ldx, const double* __restrict__ Ez,
                                            int ldez, int D, int MPd, int64_t n, int r) {
  if (KIND == ROWS_DENSE) {
    // rows: [0,D) features of Xc | [D, D+MP) components of ez
    if (r < D) return X[n * ldx + r];
    return Ez[n * ldez + (r - D)];
  } else {
    // rows: [0,D) observed value, missing (NaN) -> 0
    const double v = X[n * ldx + r];
    return isnan(v) ? 0.0 : v;
  }
}

template <int KIND>
__global__ void __launch_bounds__(256) stats_gemm_kernel(const double* __restrict__ X, int64_t ldx, int64_t n_obs, int D,
                                                         int MPd, int R1, int R, int Q1, int Q2,
                                                         const double* __restrict__ C1, int ldc1, int Cw1,
                                                         const double* __restrict__ C2, int ldc2,
                                                         double* __restrict__ partials, int64_t S, int64_t obs_per_split,
                                                         const double* __restrict__ scal) {
  if (scal[S_DONE] != 0.0) return;
  __shared__ double Rs[TG_K][TG_R];
  __shared__ double Cs[TG_K][TG_Q];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int r0 = blockIdx.x * TG_R, q0 = blockIdx.z * TG_Q;
  const int Q = (r0 < R1) ? Q1 : Q2;  // a tile that straddles R1 computes Q1 columns and stores per row
  if (q0 >= Q) return;
  const int64_t nb = (int64_t)blockIdx.y * obs_per_split;
  const int64_t ne = min(n_obs, nb + obs_per_split);
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  for (int64_t nc = nb; nc < ne; nc += TG_K) {
    // stage 16 observations x 64 rows and x 64 cols
    for (int e = threadIdx.x; e < TG_K * TG_R; e += 256) {
      const int o = e / TG_R, rr = e % TG_R;
      const int64_t n = nc + o;
      const int r = r0 + rr;
      Rs[o][rr] = (n < ne && r < R) ? row_value<KIND>(X, ldx, C1, ldc1, D, MPd, n, r) : 0.0;
    }
    for (int e = threadIdx.x; e < TG_K * TG_Q; e += 256) {
      const int o = e / TG_Q, qq = e % TG_Q;
      const int64_t n = nc + o;
      const int q = q0 + qq;
      double v = 0.0;
      if (n < ne && q < Q) v = (q < Cw1) ? C1[n * ldc1 + q] : C2[n * ldc2 + (q - Cw1)];
      Cs[o][qq] = v;
    }
    __syncthreads();
#pragma unroll
    for (int o = 0; o < TG_K; ++o) {
      double rv[4], cv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) rv[i] = Rs[o][4 * tx + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) cv[j] = Cs[o][4 * ty + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(rv[i], cv[j], acc[i][j]);
    }
    __syncthreads();
  }
  double* out = partials + (int64_t)blockIdx.y * S;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + 4 * tx + i;
    if (r >= R) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q = q0 + 4 * ty + j;
      if (q >= ((r < R1) ? Q1 : Q2)) continue;
      const int64_t off = (r < R1) ? (int64_t)r * Q1 + q : (int64_t)R1 * Q1 + (int64_t)(r - R1) * Q2 + q;
      out[off] = acc[i][j];
    }
  }
}

// stats[i] = sum_p partials[p][i] by a FIXED two-level tree (deterministic): thread (o, grp) of a 256-thread block sums
// the partials p = grp, grp + 8, ... of output 32 blockIdx.x + o, then the 8 group sums are added in order.
//
// With pa.peers != nullptr the kernel is also the ALL-REDUCE of the step -- fused reduction + collective over NVLink peer
// memory, no NCCL call: a block writes its 32 local sums into this rank's publication buffer, raises its epoch flag
// (system-scope release), waits (bounded) for the same block's flag of every other rank with system-scope acquire loads,
// then reads the `world` slices straight from the peers' HBM and adds them in RANK ORDER, so every rank forms the
// bit-identical total and the replicated update stays replicated.  A block publishes before it waits, so no rank can
// block another.  Two buffers (step parity) suffice: a rank overwrites buffer k & 1 at step k + 2, after all of its
// blocks have seen every peer's step k + 1 flags, which a peer raises only after its step-k kernel (and reads) finished.
// (partials and stats may alias: the in-place peer all-reduce of the masked statistics passes the same buffer)
__global__ void __launch_bounds__(256) reduce_partials_kernel(const double* partials, double* stats,
                                                              int64_t S, int nparts, double* __restrict__ scal, PeerArgs pa) {
  if (scal && scal[S_DONE] != 0.0) return;
  __shared__ double part[8][33];
  const int o = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + o;
  double a = 0.0;
  if (i < S) {
    // the loads of 8 consecutive terms are issued together (L2 latency overlaps); the additions keep their order
    int p = grp;
    for (; p + 56 < nparts; p += 64) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcg(&partials[(int64_t)(p + 8 * u) * S + i]);
#pragma unroll
      for (int u = 0; u < 8; ++u) a += v[u];
    }
    for (; p < nparts; p += 8) a += __ldcg(&partials[(int64_t)p * S + i]);
  }
  part[grp][o] = a;
  __syncthreads();
  double t = 0.0;
  if (grp == 0 && i < S) {
    t = part[0][o];
#pragma unroll
    for (int q = 1; q < 8; ++q) t += part[q][o];
  }
  if (pa.peers == nullptr) {
    if (grp == 0 && i < S) stats[i] = t;
    return;
  }
  double* mine = pa.peers[pa.rank];
  if (grp == 0 && i < S) mine[(int64_t)pa.parity * S + i] = t;
  __syncthreads();
  const int64_t flag_off = 2 * S;  // flags: [parity][block] of uint64, after the two data buffers
  if (threadIdx.x == 0) {
    __threadfence_system();
    unsigned long long* f = reinterpret_cast<unsigned long long*>(mine + flag_off) + (int64_t)pa.parity * gridDim.x + blockIdx.x;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(pa.epoch) : "memory");
  }
  if ((int)threadIdx.x < pa.world && (int)threadIdx.x != pa.rank) {
    const unsigned long long* f =
        reinterpret_cast<const unsigned long long*>(pa.peers[threadIdx.x] + flag_off) + (int64_t)pa.parity * gridDim.x + blockIdx.x;
    unsigned long long v = 0;
    long long spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
    } while (v < pa.epoch && ++spins < (1ll << 23) && reinterpret_cast<volatile double*>(scal)[S_ERR] != 3.0);
    if (v < pa.epoch) set_err(scal, 3.0);   // never hang the GPU: give up (after seconds of skew) and flag the session; once the
                                            // flag is up every later wait bails out at once, so a dead peer costs one timeout in total
  }
  __syncthreads();
  if (grp == 0 && i < S) {
    double acc = 0.0;
    for (int r = 0; r < pa.world; ++r) acc += (r == pa.rank) ? t : __ldcv(pa.peers[r] + (int64_t)pa.parity * S + i);
    stats[i] = acc;
  }
}

// per-feature moments of the local observations.  mode 0: sum x; 1: sum (x-mean)^2; 3: sum x^2;
// 2 (masked): [count | sum | sum of squares] over the known entries.
__global__ void __launch_bounds__(256) moments_kernel(const double* __restrict__ X, int64_t ldx, int64_t n_obs, int D, int mode,
                                                      const double* __restrict__ mean, double* __restrict__ partials,
                                                      int64_t S, int64_t obs_per_split) {
  __shared__ double red[3][8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int d = blockIdx.x * 32 + tx;
  const int64_t nb = (int64_t)blockIdx.y * obs_per_split, ne = min(n_obs, nb + obs_per_split);
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  if (d < D) {
    const double m = (mode == 1) ? mean[d] : 0.0;
    for (int64_t n = nb + ty; n < ne; n += 8) {
      const double v = X[n * ldx + d];
      if (mode == 0) a0 += v;
      else if (mode == 1) { const double t = v - m; a0 = fma(t, t, a0); }
      else if (mode == 3) a0 = fma(v, v, a0);
      else if (!isnan(v)) { a0 += 1.0; a1 += v; a2 = fma(v, v, a2); }
    }
  }
  red[0][ty][tx] = a0; red[1][ty][tx] = a1; red[2][ty][tx] = a2;
  __syncthreads();
  if (ty == 0 && d < D) {
    double* out = partials + (int64_t)blockIdx.y * S;
    const int nk = (mode == 2) ? 3 : 1;
    for (int k = 0; k < nk; ++k) {
      double t = 0.0;
      for (int j = 0; j < 8; ++j) t += red[k][j][tx];
      out[(int64_t)k * D + d] = t;
    }
  }
}

// Xc = X - mean, in place (PPCA.hs:61 -> Util.hs:175-178); sign = +1 undoes it (gplvm_session_read_data)
__global__ void __launch_bounds__(256) centre_kernel(double* __restrict__ X, int64_t ldx, int64_t n_obs, int D,
                                                     const double* __restrict__ mean, double sign) {
  const int64_t total = n_obs * D;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = e / D;
    const int d = (int)(e - n * D);
    X[n * ldx + d] = fma(sign, mean[d], X[n * ldx + d]);
  }
}

// ez_n = xc_n^T A for the generic path (materialised like the reference's expEznZ, PPCA.hs:67)
__global__ void __launch_bounds__(256) dense_ez_kernel(const double* __restrict__ X, int64_t ldx, int64_t n_obs, Par p,
                                                       const double* __restrict__ A, double* __restrict__ Ez) {
  if (p.scal[S_DONE] != 0.0) return;
  extern __shared__ double sm[];
  const int MP = p.MP, D = p.D;
  constexpr int DC = 64;                 // feature chunk
  double* As = sm;                       // DC x MP
  double* Xs = sm + DC * MP;             // 32 obs x (DC+1)
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // obs, column group
  const int64_t n = (int64_t)blockIdx.x * 32 + tx;
  double acc[GPLVM_MAX_M / 8 + 1];
  const int nj = (MP + 7) / 8;
#pragma unroll
  for (int j = 0; j < GPLVM_MAX_M / 8 + 1; ++j) acc[j] = 0.0;
  for (int d0 = 0; d0 < D; d0 += DC) {
    const int dc = min(DC, D - d0);
    for (int e = threadIdx.x; e < dc * MP; e += 256) As[e] = A[(int64_t)d0 * MP + e];
    for (int e = threadIdx.x; e < 32 * DC; e += 256) {
      const int o = e / DC, dd = e % DC;
      const int64_t nn = (int64_t)blockIdx.x * 32 + o;
      Xs[o * (DC + 1) + dd] = (nn < n_obs && dd < dc) ? X[nn * ldx + d0 + dd] : 0.0;
    }
    __syncthreads();
    for (int dd = 0; dd < dc; ++dd) {
      const double xv = Xs[tx * (DC + 1) + dd];
#pragma unroll
      for (int j = 0; j < GPLVM_MAX_M / 8 + 1; ++j)
        if (j < nj && ty + 8 * j < MP) acc[j] = fma(xv, As[dd * MP + ty + 8 * j], acc[j]);
    }
    __syncthreads();
  }
  if (n < n_obs) {
#pragma unroll
    for (int j = 0; j < GPLVM_MAX_M / 8 + 1; ++j)
      if (j < nj && ty + 8 * j < MP) Ez[n * MP + ty + 8 * j] = acc[j];
  }
}

// ------------------------------------------------------------------------------------------------
// replicated parameter kernels (single CTA; identical on every GPU after the all-reduce)
// ------------------------------------------------------------------------------------------------

// Minv = (s2 I + W^T W)^-1 (Cholesky), logdet over the first M pivots, A = W Minv.
// sm: 2 * MP * MP doubles.
__device__ void dense_prep_block(const Par& p, double* sm) {
  const int D = p.D, MP = p.MP, M = p.M;
  double* Mm = sm;
  double* Mi = sm + MP * MP;
  const double s2 = p.scal[S_SIGMA2];
  for (int e = threadIdx.x; e < MP * MP; e += blockDim.x) {
    const int a = e / MP, b = e % MP;
    double t = 0.0;
    for (int d = 0; d < D; ++d) t = fma(p.W[(int64_t)d * MP + a], p.W[(int64_t)d * MP + b], t);
    Mm[e] = t + ((a == b) ? s2 : 0.0);
  }
  __syncthreads();
  const bool ok = block_cholesky(Mm, MP, MP);
  if (threadIdx.x == 0) {
    double ld = 0.0;
    for (int i = 0; i < M; ++i) ld += log(Mm[i * MP + i]);
    p.scal[S_LOGDETM] = 2.0 * ld;
    if (!ok) set_err(p.scal, 1.0);
  }
  block_chol_inverse(Mm, Mi, MP, MP);
  for (int e = threadIdx.x; e < MP * MP; e += blockDim.x) p.Minv[e] = Mi[e];
  for (int e = threadIdx.x; e < D * MP; e += blockDim.x) {
    const int d = e / MP, b = e % MP;
    double t = 0.0;
    for (int a = 0; a < MP; ++a) t = fma(p.W[(int64_t)d * MP + a], Mi[a * MP + b], t);
    p.A[e] = t;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) dense_prep_kernel(Par p) {
  extern __shared__ double sm[];
  dense_prep_block(p, sm);
}

// stats layout: XEz [D x MP] | EE [MP x MP]
// sm: 2*MP*MP + 64 doubles
__global__ void __launch_bounds__(256) dense_update_kernel(Par p, const double* __restrict__ stats) {
  extern __shared__ double sm[];
  if (p.scal[S_DONE] != 0.0) return;
  const int D = p.D, MP = p.MP;
  double* Ezz = sm;                 // MP x MP -> Cholesky factor L (lower)
  double* red = sm + 2 * MP * MP;   // 33
  const double* XEz = stats;
  const double* EE = stats + (int64_t)D * MP;
  const double s2 = p.scal[S_SIGMA2];
  const double Nd = (double)p.N;
  for (int e = threadIdx.x; e < MP * MP; e += blockDim.x) Ezz[e] = Nd * s2 * p.Minv[e] + EE[e];
  __syncthreads();
  const bool ok = block_cholesky(Ezz, MP, MP);
  // one thread per feature row: w'_d = Ezz^-1 xezc_d ; accumulate the trace terms and max |dW|
  double second = 0.0, third = 0.0, maxdw = 0.0;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    double x[GPLVM_MAX_M], b[GPLVM_MAX_M];
    for (int a = 0; a < MP; ++a) { b[a] = XEz[(int64_t)d * MP + a]; x[a] = b[a]; }
    chol_solve_vec(Ezz, MP, MP, x);
    for (int a = 0; a < MP; ++a) {
      second = fma(x[a], b[a], second);
      maxdw = fmax(maxdw, fabs(p.W[(int64_t)d * MP + a] - x[a]));
    }
    // |U w|^2 with U = L^T:  (L^T w)_a = sum_{k>=a} L[k][a] w_k        (wr = newW U^T, PPCA.hs:76,79)
    for (int a = 0; a < MP; ++a) {
      double t = 0.0;
      for (int k = a; k < MP; ++k) t = fma(Ezz[k * MP + a], x[k], t);
      third = fma(t, t, third);
    }
    for (int a = 0; a < MP; ++a) p.W[(int64_t)d * MP + a] = x[a];
  }
  second = 2.0 * block_sum(second, red);
  third = block_sum(third, red);
  maxdw = block_max(maxdw, red);
  __syncthreads();
  if (threadIdx.x == 0) {
    const double news2 = (p.scal[S_TOTAL] - second + third) / (Nd * (double)D);
    const double dvar = fabs(news2 - s2);
    p.scal[S_SIGMA2] = news2;
    p.scal[S_DVAR] = dvar;
    p.scal[S_MAXDW] = maxdw;
    p.scal[S_STEPS] += 1.0;
    if (!ok || !(news2 > 0.0)) set_err(p.scal, 1.0);
    if (p.scal[S_USETOL] != 0.0 && !(fmax(dvar, maxdw) > p.scal[S_TOL])) p.scal[S_DONE] = 1.0;
  }
  __syncthreads();
  __threadfence_block();
  dense_prep_block(p, sm);  // parameters for the next E-step (also what the LL pass needs)
}

// In-place inverse of the symmetric positive definite 16 x 16 matrix A (shared memory, row-major) by ALL 256 threads of
// the block: 16 symmetric sweeps (see sweep_inverse16), thread (i, l) owns entry A[i][l].  Returns (in every thread) whether
// every pivot was positive; det = product of the pivots.  A plain loop: the code stays a few hundred instructions (the
// fully unrolled register version made the single-CTA update kernel 13 000 instructions = 208 KB, which streamed from L2
// on every launch and cost ~25 of its 36 microseconds).
__device__ __forceinline__ bool block_sweep_inverse16(double* A, int tid, double& det) {
  const int i = tid >> 4, l = tid & 15;
  bool ok = true;
  det = 1.0;
#pragma unroll 1
  for (int k = 0; k < 16; ++k) {
    const double pk = A[k * 16 + k], aik = A[i * 16 + k], akl = A[k * 16 + l], ail = A[tid];
    ok = ok && (pk > 0.0);
    det *= pk;
    const double d = 1.0 / pk;
    __syncthreads();
    double v;
    if (i == k) v = (l == k) ? -d : akl * d;
    else v = (l == k) ? aik * d : fma(-aik * d, akl, ail);
    A[tid] = v;
    __syncthreads();
  }
  A[tid] = -A[tid];
  __syncthreads();
  return ok;
}

// Fast path of the replicated update for MP = 16 (the fused-kernel shape).  Follows the reference literally:
// W' = XEz . inv(Ezz) (PPCA.hs:74); the third term |W' U^T|_F^2 with U = chol(Ezz) (PPCA.hs:75-76,79) equals
// sum_d w'_d^T Ezz w'_d and is evaluated in that form.  Also computes the next E-step's M^-1 = (s2' I + W'^T W')^-1,
// log det and A = W' M^-1 (dense_prep_block).  Loops over the latent index stay rolled (small code, see above); the
// operands of the inner, unrolled loops are rows of the 16 x 16 matrices in shared memory.
// Round-2 measurements (N = 2^20 per GPU, the 8-GPU shard): fully unrolled register version 13 008 instructions, 32-35 us;
// this version 1 544 instructions, 27 us; a 1024-thread variant (4 threads per feature) 31 us -- what is left is a chain of
// short dependent phases on one SM, not instruction fetch and not the arithmetic.
// sm: Ezz 256 | Einv 256 | Mi 256 | red 40 | Ws D x 17 | Bs D x 17
__global__ void __launch_bounds__(256, 1) dense_update16_kernel(Par p, const double* __restrict__ stats) {
  extern __shared__ double sm[];
  if (p.scal[S_DONE] != 0.0) return;
  constexpr int MP = 16;
  constexpr int WSP = 17;   // pitch of the per-feature rows in shared memory: one feature per thread => conflict-free
  const int D = p.D, tid = threadIdx.x;
  double* Ezz = sm;
  double* Einv = sm + 256;
  double* Mi = sm + 512;
  double* red = sm + 768;
  double* Ws = sm + 768 + 40;
  double* Bs = Ws + (size_t)D * WSP;   // XEz staged once, coalesced (a rolled loop over global loads would serialise their latency)
  const double* XEz = stats;
  const double* EE = stats + (int64_t)D * MP;
  const double s2 = p.scal[S_SIGMA2];
  const double Nd = (double)p.N;
  {
    const double e = Nd * s2 * p.Minv[tid] + EE[tid];
    for (int i = tid; i < D * 16; i += 256) Bs[(i >> 4) * WSP + (i & 15)] = XEz[i];
    Ezz[tid] = e;
    Einv[tid] = e;
  }
  __syncthreads();
  double det;
  const bool ok1 = block_sweep_inverse16(Einv, tid, det);
  double second = 0.0, third = 0.0, maxdw = 0.0;
  for (int d = tid; d < D; d += 256) {
    const double* b = Bs + d * WSP;
    double x[16];
#pragma unroll
    for (int a = 0; a < 16; ++a) x[a] = 0.0;
#pragma unroll 1
    for (int k = 0; k < 16; ++k) {   // row vector times inv(Ezz)
      const double bk = b[k];
      const double* er = Einv + k * 16;
#pragma unroll
      for (int a = 0; a < 16; ++a) x[a] = fma(bk, er[a], x[a]);
    }
    double* wrow = p.W + (int64_t)d * MP;
#pragma unroll
    for (int a = 0; a < 16; ++a) {
      maxdw = fmax(maxdw, fabs(wrow[a] - x[a]));
      wrow[a] = x[a];
      Ws[d * WSP + a] = x[a];
    }
    double y[16];   // Ezz w'_d
#pragma unroll
    for (int a = 0; a < 16; ++a) y[a] = 0.0;
#pragma unroll 1
    for (int k = 0; k < 16; ++k) {
      const double xk = Ws[d * WSP + k];   // own row: written by this thread
      second = fma(xk, b[k], second);
      const double* zr = Ezz + k * 16;
#pragma unroll
      for (int a = 0; a < 16; ++a) y[a] = fma(zr[a], xk, y[a]);
    }
#pragma unroll
    for (int a = 0; a < 16; ++a) third = fma(x[a], y[a], third);
  }
  second = 2.0 * block_sum(second, red);
  third = block_sum(third, red);
  maxdw = block_max(maxdw, red);
  __syncthreads();
  const double news2 = (p.scal[S_TOTAL] - second + third) / (Nd * (double)D);
  if (tid == 0) {
    const double dvar = fabs(news2 - s2);
    p.scal[S_SIGMA2] = news2;
    p.scal[S_DVAR] = dvar;
    p.scal[S_MAXDW] = maxdw;
    p.scal[S_STEPS] += 1.0;
    if (!ok1 || !(news2 > 0.0)) set_err(p.scal, 1.0);
    if (p.scal[S_USETOL] != 0.0 && !(fmax(dvar, maxdw) > p.scal[S_TOL])) p.scal[S_DONE] = 1.0;
  }
  {  // M = s2' I + W'^T W'
    const int a = tid >> 4, b = tid & 15;
    double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
    int d = 0;
#pragma unroll 2
    for (; d + 3 < D; d += 4) {
      t0 = fma(Ws[d * WSP + a], Ws[d * WSP + b], t0);
      t1 = fma(Ws[(d + 1) * WSP + a], Ws[(d + 1) * WSP + b], t1);
      t2 = fma(Ws[(d + 2) * WSP + a], Ws[(d + 2) * WSP + b], t2);
      t3 = fma(Ws[(d + 3) * WSP + a], Ws[(d + 3) * WSP + b], t3);
    }
    for (; d < D; ++d) t0 = fma(Ws[d * WSP + a], Ws[d * WSP + b], t0);
    Mi[tid] = ((t0 + t1) + (t2 + t3)) + ((a == b) ? news2 : 0.0);
  }
  __syncthreads();
  const bool ok2 = block_sweep_inverse16(Mi, tid, det);
  p.Minv[tid] = Mi[tid];
  if (tid == 0) {
    // log det over the first M pivots only matters when M < 16 (padded columns contribute log s2 each)
    p.scal[S_LOGDETM] = log(det) - (double)(16 - p.M) * log(news2);
    if (!ok2) set_err(p.scal, 1.0);
  }
  for (int d = tid; d < D; d += 256) {
    double acc[16];
#pragma unroll
    for (int b = 0; b < 16; ++b) acc[b] = 0.0;
#pragma unroll 1
    for (int a = 0; a < 16; ++a) {
      const double wa = Ws[d * WSP + a];
      const double* mr = Mi + a * 16;
#pragma unroll
      for (int b = 0; b < 16; ++b) acc[b] = fma(wa, mr[b], acc[b]);
    }
    double* arow = p.A + (int64_t)d * MP;
#pragma unroll
    for (int b = 0; b < 16; ++b) arow[b] = acc[b];
  }
}

// LL = -(N/2) (D log 2pi + log det C + tr(C^-1 S)),  S = Xc Xc^T / (N-1)            (PPCA.hs:87-89)
//   log det C = (D-M) log s2 + log det(s2 I + W^T W);  tr(C^-1 S) = (tr S - tr(Minv W^T S W)) / s2
// stats (from a statistics pass with A := W): EE = sum_n t_n t_n^T, t_n = W^T xc_n.
__global__ void dense_ll_kernel(Par p, const double* __restrict__ stats) {
  if (threadIdx.x != 0) return;
  const int D = p.D, MP = p.MP, M = p.M;
  const double* TT = stats + (int64_t)D * MP;
  const double s2 = p.scal[S_SIGMA2], Nd = (double)p.N;
  double tr = 0.0;
  for (int a = 0; a < MP; ++a)
    for (int b = 0; b < MP; ++b) tr = fma(p.Minv[a * MP + b], TT[b * MP + a], tr);
  const double trS = p.scal[S_TOTAL] / (Nd - 1.0);
  const double trCS = (trS - tr / (Nd - 1.0)) / s2;
  const double logdetC = (double)(D - M) * log(s2) + p.scal[S_LOGDETM];
  p.scal[S_LL] = -1.0 * (Nd / 2.0) * ((double)D * log(2.0 * M_PI) + logdetC + trCS);
}

// mean = sums / N ; c for the LL pass
__global__ void dense_set_mean_kernel(Par p, const double* __restrict__ sums) {
  for (int d = threadIdx.x + blockIdx.x * blockDim.x; d < p.D; d += blockDim.x * gridDim.x) p.mu[d] = sums[d] / (double)p.N;
}
__global__ void dense_set_total_kernel(Par p, const double* __restrict__ sums) {
  // sums[d] = sum_n (x_nd - mean_d)^2 ; fixed order over features
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int d = 0; d < p.D; ++d) t += sums[d];
    p.scal[S_TOTAL] = t;
  }
}
int splits_for(int64_t n_obs, int64_t unit, int tiles_xy) {
  // aim for ~4 CTAs per SM over the whole grid
  int64_t want = ceil_div(148 * 4, tiles_xy > 0 ? tiles_xy : 1);
  int64_t maxs = ceil_div(n_obs, unit);
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  return (int)want;
}

}  // namespace

int launch_reduce_partials(ShardState& sh, int64_t S, int nparts, double* dst, double* scal) {
  PeerArgs pa{nullptr, 1, 0, 0, 0ull, (long long)S};
  if (!dst && sh.peer_args.peers) pa = sh.peer_args;   // set by the dense step when the peer exchange is active
  GP_LAUNCH(reduce_partials_kernel, (unsigned)ceil_div(S, 32), 256, 0, sh.st, sh.partials, dst ? dst : sh.stats, S, nparts,
            scal ? scal : sh.par.scal, pa);
  return GPLVM_OK;
}

int launch_stats_gemm(ShardState& sh, RowKind kind, int D, int R1, int R, int Q1, int Q2, const double* Cmat, int ldc,
                      int Cw1, const double* Cmat2, int ldc2, int64_t S, double* dst) {
  const int rt = (int)ceil_div(R, TG_R), qt = (int)ceil_div(Q1, TG_Q);
  int ks = splits_for(sh.n, 1024, rt * qt);
  if (ks > sh.nparts) ks = sh.nparts;
  int64_t per = round_up(ceil_div(sh.n, ks), TG_K);
  ks = (int)ceil_div(sh.n, per);
  dim3 grid(rt, ks, qt);
  CUDA_TRY(cudaMemsetAsync(sh.partials, 0, sizeof(double) * S * ks, sh.st));
  if (kind == ROWS_DENSE)
    GP_LAUNCH(stats_gemm_kernel<ROWS_DENSE>, grid, 256, 0, sh.st, sh.X, sh.ldx, sh.n, D, sh.par.MP, R1, R, Q1, Q2, Cmat, ldc,
              Cw1, Cmat2, ldc2, sh.partials, S, per, sh.par.scal);
  else
    GP_LAUNCH(stats_gemm_kernel<ROWS_MASKED_VALUES>, grid, 256, 0, sh.st, sh.X, sh.ldx, sh.n, D, sh.par.MP, R1, R, Q1, Q2, Cmat, ldc,
              Cw1, Cmat2, ldc2, sh.partials, S, per, sh.par.scal);
  GP_TRY(launch_reduce_partials(sh, S, ks, dst));
  return GPLVM_OK;
}

static int launch_moments(gplvm_ppca_session* s, int mode) {
  for (ShardState& sh : s->sh) {
    CUDA_TRY(cudaSetDevice(sh.dev));
    const int ft = (int)ceil_div(s->D, 32);
    int ks = splits_for(sh.n, 256, ft);
    if (ks > sh.nparts) ks = sh.nparts;
    const int64_t per = ceil_div(sh.n, ks);
    ks = (int)ceil_div(sh.n, per);
    const int64_t S = 3 * s->D;
    dim3 grid(ft, ks);
    GP_LAUNCH(moments_kernel, grid, 256, 0, sh.st, sh.X, sh.ldx, sh.n, (int)s->D, mode, sh.par.mu, sh.partials, S, per);
    GP_LAUNCH(reduce_partials_kernel, (unsigned)ceil_div(S, 32), 256, 0, sh.st, sh.partials, sh.stats, S, ks,
              (double*)nullptr, (PeerArgs{nullptr, 1, 0, 0, 0ull, (long long)S}));
  }
  return GPLVM_OK;
}

int moments_pass(gplvm_ppca_session* s, int mode) { return launch_moments(s, mode); }

int dense_centre(gplvm_ppca_session* s, double sign) {
  for (ShardState& sh : s->sh) {
    CUDA_TRY(cudaSetDevice(sh.dev));
    GP_LAUNCH(centre_kernel, 148 * 8, 256, 0, sh.st, sh.X, sh.ldx, sh.n, (int)s->D, sh.par.mu, sign);
  }
  return GPLVM_OK;
}

int dense_init_constants(gplvm_ppca_session* s) {
  // feature means (Util.hs:159-164 on the transposed matrix), centring in place (PPCA.hs:61), then |Xc|_F^2 (PPCA.hs:77)
  if (!s->centred) {
    GP_TRY(launch_moments(s, 0));
    GP_TRY(session_allreduce_stats(s, s->D));
    for (ShardState& sh : s->sh) {
      CUDA_TRY(cudaSetDevice(sh.dev));
      GP_LAUNCH(dense_set_mean_kernel, 4, 256, 0, sh.st, sh.par, sh.stats);
    }
    GP_TRY(dense_centre(s, -1.0));
    s->centred = true;
    GP_TRY(launch_moments(s, 3));
    GP_TRY(session_allreduce_stats(s, s->D));
    for (ShardState& sh : s->sh) {
      CUDA_TRY(cudaSetDevice(sh.dev));
      GP_LAUNCH(dense_set_total_kernel, 1, 32, 0, sh.st, sh.par, sh.stats);
    }
  }
  const size_t smem = sizeof(double) * (2 * s->MP * s->MP + 64);
  for (ShardState& sh : s->sh) {
    CUDA_TRY(cudaSetDevice(sh.dev));
    GP_LAUNCH(dense_prep_kernel, 1, 256, smem, sh.st, sh.par);
  }
  return GPLVM_OK;
}

// statistics pass with projection matrix proj (D x MP): fills sh.stats (local sums)
static int dense_stats_pass(gplvm_ppca_session* s, ShardState& sh, const double* proj, bool em_step = false) {
  if (s->use_dmma) return dense_dmma_stats(s, sh, proj);
  const int MP = s->MP, D = (int)s->D;
  const size_t smem = sizeof(double) * (64 * MP + 32 * 65);
  GP_LAUNCH(dense_ez_kernel, (unsigned)ceil_div(sh.n, 32), 256, smem, sh.st, sh.X, sh.ldx, sh.n, sh.par, proj, sh.Ez);
  seg_mark(sh, 1);   // generic path: segment 0 = E[z] kernel, segment 1 = statistics GEMM + reduction
  const int R = D + MP;
  return launch_stats_gemm(sh, ROWS_DENSE, D, R, R, MP, MP, sh.Ez, MP, MP, nullptr, 0, s->S, sh.stats);
}

// arguments of the next exchange over peer memory (session_setup_p2p succeeded): one epoch per exchange, two buffers
static void next_peer_args(gplvm_ppca_session* s) {
  s->xchg_epoch += 1;
  for (ShardState& sh : s->sh) {
    sh.peer_args.peers = sh.peers_dev;
    sh.peer_args.world = ctx().world;
    sh.peer_args.rank = ctx().multiproc ? ctx().rank : (int)(&sh - &s->sh[0]);
    sh.peer_args.parity = (int)(s->xchg_epoch & 1);
    sh.peer_args.epoch = s->xchg_epoch;
    sh.peer_args.S = (long long)s->S;
  }
}

// In-place all-reduce of sh.stats (s->S doubles) over NVLink peer memory: the reduction kernel with a single "partial"
// (the local statistics themselves) publishes, waits for the peers' flags and adds the ranks in rank order.
int session_peer_allreduce_stats(gplvm_ppca_session* s) {
  if (!s->p2p) return fail(GPLVM_ERR_STATE, "peer exchange is not set up");
  next_peer_args(s);
  for (ShardState& sh : s->sh) {
    CUDA_TRY(cudaSetDevice(sh.dev));
    GP_LAUNCH(reduce_partials_kernel, (unsigned)ceil_div(s->S, 32), 256, 0, sh.st, sh.stats, sh.stats, s->S, 1, sh.par.scal, sh.peer_args);
    sh.peer_args.peers = nullptr;
  }
  return GPLVM_OK;
}

int dense_enqueue_step(gplvm_ppca_session* s) {
  const bool p2p_step = s->p2p;
  if (p2p_step) next_peer_args(s);
  for (ShardState& sh : s->sh) {
    CUDA_TRY(cudaSetDevice(sh.dev));
    if (sh.evk0) CUDA_TRY(cudaEventRecord(sh.evk0, sh.st));
    seg_mark(sh, 0);
    if (!p2p_step) sh.peer_args.peers = nullptr;
    GP_TRY(dense_stats_pass(s, sh, sh.par.A, true));
    sh.peer_args.peers = nullptr;
    if (sh.evk1) CUDA_TRY(cudaEventRecord(sh.evk1, sh.st));
  }
  const size_t smem = sizeof(double) * (2 * s->MP * s->MP + 64);
  const bool fast16 = (s->MP == 16 && s->D <= 512);
  if (!s->p2p) GP_TRY(session_allreduce_stats(s, s->S));
  const size_t smem16 = sizeof(double) * (768 + 40 + (size_t)s->D * 34);
  static bool attr_done[64] = {};
  for (ShardState& sh : s->sh) {
    CUDA_TRY(cudaSetDevice(sh.dev));
    if (fast16 && !attr_done[sh.dev & 63]) {
      CUDA_TRY(cudaFuncSetAttribute(dense_update16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      attr_done[sh.dev & 63] = true;
    }
    seg_mark(sh, 2);
    if (fast16) {
      GP_LAUNCH(dense_update16_kernel, 1, 256, smem16, sh.st, sh.par, sh.stats);
    } else {
      GP_LAUNCH(dense_update_kernel, 1, 256, smem, sh.st, sh.par, sh.stats);
    }
    seg_mark(sh, 3);
  }
  return GPLVM_OK;
}

int dense_loglik(gplvm_ppca_session* s, double* ll) {
  // one extra pass with the projection W (A := W): EE becomes W^T Xc Xc^T W
  for (ShardState& sh : s->sh) {
    CUDA_TRY(cudaSetDevice(sh.dev));
    GP_TRY(dense_stats_pass(s, sh, sh.par.W));
  }
  GP_TRY(session_allreduce_stats(s, s->S));
  for (ShardState& sh : s->sh) {
    CUDA_TRY(cudaSetDevice(sh.dev));
    GP_LAUNCH(dense_ll_kernel, 1, 32, 0, sh.st, sh.par, sh.stats);
  }
  ShardState& s0 = s->sh[0];
  CUDA_TRY(cudaSetDevice(s0.dev));
  double scal[S_COUNT];
  CUDA_TRY(cudaMemcpyAsync(scal, s0.par.scal, sizeof scal, cudaMemcpyDeviceToHost, s0.st));
  CUDA_TRY(cudaStreamSynchronize(s0.st));
  if (scal[S_ERR] != 0.0) return fail(GPLVM_ERR_NUMERIC, "dense PPCA: matrix not positive definite");
  *ll = scal[S_LL];
  return GPLVM_OK;
}

}  // namespace gplvm
