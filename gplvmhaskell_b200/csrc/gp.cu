// gp.cu -- GP kernel-matrix build fused with a blocked, tiled FP64 Cholesky, log-determinant and the
// log marginal likelihood; plus the stand-alone kernel builder, Cholesky and lower-triangular solve.
//
// Reference call sites (all relative to the reference checkout):
//   src/GPLVM/GPExample.hs:79-91     sqDist          (expanded square, result is len(v2) x len(v1))
//   src/GPLVM/GPExample.hs:93-100    kernelFunction  (exp(-0.5 * (1/0.1) * d2))
//   src/GPLVM/GaussianProcess.hs:67-68  cholSH (K + 5e-5 I)
//   src/GPLVM/GaussianProcess.hs:74,77  linearSolveS cholK ...   (triangular factor solves)
// Extension (BASELINE.json north_star / SURVEY.md 8a row G4): Q-dimensional ARD inputs and
//   lml = -1/2 |L^-1 y|^2 - sum log L_ii - N/2 log 2 pi.
//
// Factorisation layout: the matrix lives in HBM as N x N row-major, lower triangle only is touched.
// Right-looking blocked algorithm with panel width NB = 128:
//   panel k:  [diag]  L11 = chol(A11) in shared memory by one CTA, T = L11^-1 (triangular inverse),
//                     alpha_k = T y_k, logdet += 2 sum log diag(L11)
//             [panel] L21 = A21 T^T            (FP64 tensor-core GEMM, C = A B^T form), fused with the
//                     forward-substitution update y_rest -= L21 alpha_k
//             [syrk]  A22 -= L21 L21^T         (same GEMM kernel, lower tiles only).  For k = 0 the
//                     epilogue GENERATES the kernel-matrix tile instead of reading it, so the N x N kernel
//                     matrix is never written to and re-read from HBM before the factorisation
//                     touches it ("kernel build fused with the Cholesky").
// The GEMM uses mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4 on sm_100a; tcgen05 has no FP64 kind).
#include "common.cuh"

namespace gplvm {
namespace {

constexpr int NB = 128;      // panel width == GEMM tile edge
constexpr int KC = 16;       // K chunk staged in shared memory
constexpr int KP = KC + 4;   // padded row stride (doubles) of a staged chunk: conflict-free fragment loads
constexpr int STAGES = 3;
constexpr int GEMM_SMEM = STAGES * 2 * NB * KP * (int)sizeof(double);          // 128 x 128 tiles (panel)
constexpr int SYRK_SMEM = STAGES * (NB + 64) * KP * (int)sizeof(double);       // 128 x 64 tiles (trailing update)

struct KernelSpec {      // how to generate a kernel-matrix entry on the fly
  const double* X;       // N x Q row-major inputs (device)
  const double* inv_l2;  // Q
  double variance, jitter;
  int Q;
};

// variance * exp(-0.5 * sum_q inv_l2[q] * (a_q^2 + (b_q^2 - 2 a_q b_q)))  -- the association order of
// sqDist (GPExample.hs:83-87): v1^2 + (v2^2 - 2 v1 v2), with a = X1[i], b = X2[j].
__device__ __forceinline__ double rbf_entry(const double* __restrict__ a, const double* __restrict__ b,
                                            const double* __restrict__ inv_l2, int Q, double variance) {
  double acc = 0.0;
  for (int q = 0; q < Q; ++q) {
    const double v1 = a[q], v2 = b[q];
    const double d2 = __dadd_rn(__dmul_rn(v1, v1), __dsub_rn(__dmul_rn(v2, v2), __dmul_rn(__dmul_rn(2.0, v1), v2)));
    acc = __dadd_rn(acc, __dmul_rn(inv_l2[q], d2));
  }
  return variance * exp(-0.5 * acc);
}

// K_out[j * n1 + i] = k(X1[i], X2[j]); i is the fastest index => coalesced, 2 entries per thread.
__global__ void __launch_bounds__(256) rbf_build_kernel(const double* __restrict__ X1, int64_t n1, const double* __restrict__ X2,
                                                        int64_t n2, int Q, const double* __restrict__ inv_l2, double variance,
                                                        double* __restrict__ out, int64_t ldo, int64_t j_begin, int64_t j_count) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  const int64_t j = j_begin + blockIdx.y;
  if (i >= n1 || j >= j_begin + j_count) return;
  const double* b = X2 + j * Q;
  const double v0 = rbf_entry(X1 + i * Q, b, inv_l2, Q, variance);
  if (i + 1 < n1) {
    const double v1 = rbf_entry(X1 + (i + 1) * Q, b, inv_l2, Q, variance);
    if (((ldo | i) & 1) == 0) {
      *reinterpret_cast<double2*>(out + (j - j_begin) * ldo + i) = make_double2(v0, v1);
    } else {
      out[(j - j_begin) * ldo + i] = v0;
      out[(j - j_begin) * ldo + i + 1] = v1;
    }
  } else {
    out[(j - j_begin) * ldo + i] = v0;
  }
}

// first panel of the fused factorisation: A[i][j] for j < nb, i >= j (row-major, ld = N), with jitter
__global__ void __launch_bounds__(128) gp_first_panel_kernel(double* __restrict__ A, int64_t ld, int nb, KernelSpec ks) {
  const int64_t i = blockIdx.x;
  const int j = threadIdx.x;
  if (j >= nb || j > i) return;
  double v = rbf_entry(ks.X + (int64_t)j * ks.Q, ks.X + i * ks.Q, ks.inv_l2, ks.Q, ks.variance);
  if (i == j) v += ks.jitter;
  A[i * ld + j] = v;
}

// ------------------------------------------------------------------------------------------------
// diagonal block: Cholesky + triangular inverse + forward-substitution piece + log-determinant
// ------------------------------------------------------------------------------------------------
// state[0] = sum log diag(L) so far, state[1] = error flag (index of the failing panel + 1)
//
// Shared tile Ls (nb x LDS): lower triangle = L; after the factorisation the strict upper triangle receives the
// strictly-lower part of T = L^-1 TRANSPOSED (T[i][j], i > j, lives at Ls[j][i]); diag(T) = 1 / diag(L).
constexpr int LDS_DIAG = NB + 1;
constexpr int DIAG_SMEM = (NB * LDS_DIAG + NB) * (int)sizeof(double);

// T = L^-1 by forward substitution on e_j, one thread per column j (reads row j of the upper triangle as
// its own history => bank-conflict free, stride LDS_DIAG between threads).  dinv[i] = 1 / L[i][i].  Four
// independent partial sums per entry break the FMA dependency chain.
__device__ __forceinline__ void tri_inverse_in_place(double* Ls, int nb, const double* dinv) {
  for (int j = threadIdx.x; j < nb; j += blockDim.x) {
    const double tjj = dinv[j];
    const double* hist = Ls + j * LDS_DIAG;   // hist[k] = T[k][j] for k > j
    for (int i = j + 1; i < nb; ++i) {
      const double* Li = Ls + i * LDS_DIAG;
      double s0 = -Li[j] * tjj, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int k = j + 1;
      for (; k + 3 < i; k += 4) {
        s0 = fma(-Li[k], hist[k], s0);
        s1 = fma(-Li[k + 1], hist[k + 1], s1);
        s2 = fma(-Li[k + 2], hist[k + 2], s2);
        s3 = fma(-Li[k + 3], hist[k + 3], s3);
      }
      for (; k < i; ++k) s0 = fma(-Li[k], hist[k], s0);
      Ls[j * LDS_DIAG + i] = ((s0 + s1) + (s2 + s3)) * dinv[i];
    }
  }
  __syncthreads();
}
// The same inverse for a full 128 x 128 factor (identity padded), blocked by 32, in place and without scratch:
//   1. the four diagonal blocks T_II = L_II^-1 by substitution, one thread per column (128 threads, <= 465 multiply-adds each);
//   2. U_I = T_II L_I[:, 0 .. 32 I) overwrites the off-diagonal part of block row I (L has been copied out; one thread per
//      column keeps the 32 entries in registers, 192 threads);  then T_II [L_I | L_II] = [U_I | 1], so
//   3. T_IJ = - sum_{K = J .. I-1} U_IK T_KJ, sub-diagonal by sub-diagonal (2 x 2 entries per thread, 256 threads).
// The thread-per-column routine above gives its longest thread 8 000 dependent multiply-adds; here no thread has more than
// ~1 300 (gp_diag_kernel sits on the critical path of the factorisation's last outer steps).  The shared-memory footprint
// must not grow: the kernel has to fit on an SM beside ONE 92 KB trailing-update CTA (228 - 93 KB), or it waits for an SM to
// drain completely.  Storage as above: T[i][j], i > j, in Ls[j * LDS_DIAG + i]; the diagonal of T is dinv.
__device__ inline void tri_inverse_blocked_128(double* Ls, const double* dinv) {
  constexpr int LD = LDS_DIAG;
  const int tid = threadIdx.x;
  if (tid < NB) {
    const int b0 = tid & ~31, j = tid;
    const double tjj = dinv[j];
    const double* hist = Ls + j * LD;
    for (int i = j + 1; i < b0 + 32; ++i) {
      const double* Li = Ls + i * LD;
      double s0 = -Li[j] * tjj, s1 = 0.0;
      int k = j + 1;
      for (; k + 1 < i; k += 2) {
        s0 = fma(-Li[k], hist[k], s0);
        s1 = fma(-Li[k + 1], hist[k + 1], s1);
      }
      if (k < i) s0 = fma(-Li[k], hist[k], s0);
      Ls[j * LD + i] = (s0 + s1) * dinv[i];
    }
  }
  __syncthreads();
  if (tid < 192) {   // (I, m): I = 1: m < 32 (tid 0..31); I = 2: m < 64 (tid 32..95); I = 3: m < 96 (tid 96..191)
    const int I = (tid < 32) ? 1 : ((tid < 96) ? 2 : 3);
    const int m = tid - ((I == 1) ? 0 : ((I == 2) ? 32 : 96));
    const int i0 = 32 * I;
    double col[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) col[k] = Ls[(i0 + k) * LD + m];
#pragma unroll
    for (int r = 31; r >= 0; --r) {   // row r needs the original rows k <= r only: descending order, all in registers
      double u = dinv[i0 + r] * col[r];
#pragma unroll
      for (int k = 0; k < r; ++k) u = fma(Ls[(i0 + k) * LD + i0 + r], col[k], u);   // T_II[r][k], a broadcast read
      Ls[(i0 + r) * LD + m] = u;
    }
  }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;   // rows 2 ty + {0, 1}, columns 2 tx + {0, 1} of a 32 x 32 block
#pragma unroll 1
  for (int d = 1; d < NB / 32; ++d) {
#pragma unroll 1
    for (int J = 0; J + d < NB / 32; ++J) {
      const int i0 = 32 * (J + d), c0 = 32 * J + 2 * tx, c1 = c0 + 1;
      const double* Ur0 = Ls + (i0 + 2 * ty) * LD;
      const double* Ur1 = Ur0 + LD;
      const double* Tc0 = Ls + c0 * LD;   // T[m][c0] for m > c0
      const double* Tc1 = Tc0 + LD;
      double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
      {   // m = c0 and m = c1: the diagonal of T_JJ
        const double u00 = Ur0[c0], u10 = Ur1[c0], u01 = Ur0[c1], u11 = Ur1[c1];
        const double d0 = dinv[c0], d1 = dinv[c1], t10 = Tc0[c1];   // T[c1][c0]
        a00 = fma(u01, t10, u00 * d0); a10 = fma(u11, t10, u10 * d0);
        a01 = u01 * d1; a11 = u11 * d1;
      }
      for (int mm = c1 + 1; mm < i0; ++mm) {
        const double t0 = Tc0[mm], t1 = Tc1[mm], u0 = Ur0[mm], u1 = Ur1[mm];
        a00 = fma(u0, t0, a00); a01 = fma(u0, t1, a01);
        a10 = fma(u1, t0, a10); a11 = fma(u1, t1, a11);
      }
      // T_IJ goes to the transposed upper storage; nobody reads these entries in this sub-diagonal pass
      Ls[c0 * LD + i0 + 2 * ty] = -a00; Ls[c1 * LD + i0 + 2 * ty] = -a01;
      Ls[c0 * LD + i0 + 2 * ty + 1] = -a10; Ls[c1 * LD + i0 + 2 * ty + 1] = -a11;
    }
    __syncthreads();
  }
}
__device__ __forceinline__ double tinv_at(const double* Ls, int i, int j) {
  return (j < i) ? Ls[j * LDS_DIAG + i] : ((i == j) ? 1.0 / Ls[i * LDS_DIAG + i] : 0.0);
}

// In-place lower Cholesky of the 128 x 128 SPD matrix in shared memory (lower triangle valid on entry; the strict
// upper triangle is scratch), blocked by 32 columns:
//   factor : right-looking Cholesky of the 32 x 32 diagonal block in shared memory by all threads (rsqrt + Newton)
//   trsm   : one thread per row below solves x L11^T = a in place by substitution (L11 reads are shared-memory broadcasts)
//   syrk   : 16 x 16 threads, 6 x 6 register tiles, trailing block -= L21 L21^T
// dinv[j] receives 1 / L[j][j].  Returns (in every thread) false when a pivot is not positive.
__device__ inline bool block_cholesky_128(double* Ls, double* dinv, int* flag) {
  const int tid = threadIdx.x;
  const int LD = LDS_DIAG;
  if (tid == 0) *flag = 1;
  __syncthreads();
#pragma unroll 1
  for (int c0 = 0; c0 < NB; c0 += 32) {
    // factor: right-looking Cholesky of the 32 x 32 diagonal block in shared memory by all threads (rolled loops: the
    // register / shuffle version of round 1 unrolled to 8 544 instructions = 137 KB, refetched on each of the 256 launches)
    bool ok = true;
#pragma unroll 1
    for (int k = 0; k < 32; ++k) {
      const int kk = c0 + k;
      __syncthreads();   // the trailing update of step k - 1 (which touched the pivot) is complete
      const double pk = Ls[kk * LD + kk];
      ok = ok && (pk > 0.0);
      const double y = rsqrt_nr(pk);
      if (tid < 31 - k) Ls[(kk + 1 + tid) * LD + kk] *= y;   // column k below the pivot (nobody reads it before the barrier)
      __syncthreads();
      if (tid == 0) {   // the pivot itself is not read again before the factor is copied out
        dinv[kk] = y;
        Ls[kk * LD + kk] = pk * y;
      }
      const int r = 31 - k;
      for (int idx = tid; idx < r * r; idx += 256) {
        const int i = kk + 1 + idx / r, j = kk + 1 + idx % r;
        if (j <= i) Ls[i * LD + j] = fma(-Ls[i * LD + kk], Ls[j * LD + kk], Ls[i * LD + j]);
      }
    }
    if (!ok && tid == 0) *flag = 0;
    __syncthreads();
    const int rb = NB - c0 - 32;  // rows below this block column
    if (tid < rb) {   // trsm, in place: row i of the panel below, x L11^T = a by substitution
      double* xi = Ls + (c0 + 32 + tid) * LD + c0;
#pragma unroll 1
      for (int j = 0; j < 32; ++j) {
        const double* lj = Ls + (c0 + j) * LD + c0;
        double s0 = xi[j], s1 = 0.0;
        int k = 0;
        for (; k + 1 < j; k += 2) {
          s0 = fma(-xi[k], lj[k], s0);
          s1 = fma(-xi[k + 1], lj[k + 1], s1);
        }
        if (k < j) s0 = fma(-xi[k], lj[k], s0);
        xi[j] = (s0 + s1) * dinv[c0 + j];
      }
    }
    __syncthreads();
    if (rb > 0) {
      const int ty = tid >> 4, tx = tid & 15, nt = rb >> 4, base = c0 + 32;
      double acc[6][6];
#pragma unroll
      for (int p = 0; p < 6; ++p)
#pragma unroll
        for (int q = 0; q < 6; ++q) acc[p][q] = 0.0;
      for (int k = 0; k < 32; ++k) {
        double li[6], lj[6];
#pragma unroll
        for (int p = 0; p < 6; ++p) {
          li[p] = (p < nt) ? Ls[(base + ty + 16 * p) * LD + c0 + k] : 0.0;
          lj[p] = (p < nt) ? Ls[(base + tx + 16 * p) * LD + c0 + k] : 0.0;
        }
#pragma unroll
        for (int p = 0; p < 6; ++p)
#pragma unroll
          for (int q = 0; q < 6; ++q) acc[p][q] = fma(li[p], lj[q], acc[p][q]);
      }
#pragma unroll
      for (int p = 0; p < 6; ++p)
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const int i = base + ty + 16 * p, j = base + tx + 16 * q;
          if (p < nt && q < nt && j <= i) Ls[i * LD + j] -= acc[p][q];
        }
    }
    __syncthreads();
  }
  return *flag != 0;
}

__global__ void __launch_bounds__(256, 1) gp_diag_kernel(double* __restrict__ A, int64_t ld, int64_t k0, int nb, double* __restrict__ Tinv,
                                                      double* __restrict__ y, double* __restrict__ alpha, double* __restrict__ state) {
  extern __shared__ double sm[];
  double* Ls = sm;
  double* dinv = sm + NB * LDS_DIAG;
  __shared__ int flag;
  const int tid = threadIdx.x;
  // rows / columns past nb (last, incomplete panel) are padded with the identity
  for (int e = tid; e < NB * NB; e += blockDim.x) {
    const int i = e / NB, j = e % NB;
    double v = 0.0;
    if (j <= i) v = (i < nb) ? A[(k0 + i) * ld + (k0 + j)] : ((i == j) ? 1.0 : 0.0);
    Ls[i * LDS_DIAG + j] = v;
  }
  __syncthreads();
  const bool ok = block_cholesky_128(Ls, dinv, &flag);
  for (int e = tid; e < nb * nb; e += blockDim.x) {
    const int i = e / nb, j = e % nb;
    if (j <= i) A[(k0 + i) * ld + (k0 + j)] = Ls[i * LDS_DIAG + j];
  }
  __syncthreads();
  tri_inverse_blocked_128(Ls, dinv);   // rows / columns past nb are identity: their T entries are 0
  for (int e = tid; e < NB * NB; e += blockDim.x) {
    const int i = e / NB, j = e % NB;
    Tinv[e] = (i < nb && j < nb) ? tinv_at(Ls, i, j) : 0.0;
  }
  if (y) {
    for (int i = tid; i < nb; i += blockDim.x) {
      double s = 0.0;
      for (int k = 0; k <= i; ++k) s = fma(tinv_at(Ls, i, k), y[k0 + k], s);
      alpha[k0 + i] = s;
    }
  }
  __syncthreads();
  if (tid < nb) dinv[tid] = log(Ls[tid * LDS_DIAG + tid]);   // dinv is free again: one log per thread, summed in order below
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int i = 0; i < nb; ++i) s += dinv[i];
    state[0] += s;
    if (!ok && state[1] == 0.0) state[1] = (double)(k0 / NB + 1);
  }
}

// ------------------------------------------------------------------------------------------------
// FP64 tensor-core GEMM, C(128 x 128 tile) = op(C) - / = A(rows i) . B(rows j)^T over K
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = pred ? 16 : 0;  // zero-fill when predicated off
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

enum GemmMode { MODE_PANEL = 0, MODE_SYRK = 1, MODE_SYRK_GEN = 2 };

// MODE_PANEL : rows [r0, N) of the panel: C = A . B^T, A = C's own panel columns (in place), B = Tinv (NB x NB);
//              one CTA per 128-row strip (TN = 128); also y[i] -= sum_j C[i][j] alpha[j].
// MODE_SYRK  : trailing matrix, lower tiles: C -= A . B^T with A = panel rows of the tile's rows, B = panel rows of the
//              tile's columns.  128 x 64 tiles (TN = 64), 2 CTAs per SM so that one CTA's read-modify-write epilogue
//              hides under the other's MMA loop.
// MODE_SYRK_GEN: as SYRK but C is generated from the kernel function instead of being read.
// All row indices are global; `Kdim` = panel width (multiple of 4); A/B rows have leading dimensions lda/ldb.
// Tile numbering (SYRK): tile row ti covers rows r0 + 128 ti ..., tile column tj covers columns r0 + 64 tj ...;
// row ti has the 2 ti + 2 column tiles that touch the lower triangle => linear index b = ti (ti + 1) + tj.
// colmode: only the first 128-column block column (two column tiles) of the trailing matrix.
template <int MODE, int TN>
__global__ void __launch_bounds__(256, (TN == 64) ? 2 : 1)
    gp_gemm_kernel(double* __restrict__ C, int64_t ldc, const double* __restrict__ Apan, int64_t lda, const double* __restrict__ Bmat,
                   int64_t ldb, int64_t r0, int64_t N, int64_t ccol0, int Kdim, double* __restrict__ y,
                   const double* __restrict__ alpha, KernelSpec ks, int colmode) {
  constexpr int WN_WARPS = TN / 32;            // 4 | 2
  constexpr int WM_WARPS = 8 / WN_WARPS;       // 2 | 4
  constexpr int WROWS = NB / WM_WARPS;         // 64 | 32 rows per warp
  constexpr int MT = WROWS / 8;                // 8 | 4 m-tiles per warp (4 n-tiles always)
  constexpr int STAGE_D = (NB + TN) * KP;      // doubles per stage
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp / WN_WARPS, wn = warp % WN_WARPS;

  int64_t ti, tj;
  if (MODE == MODE_PANEL) {
    ti = blockIdx.x; tj = 0;
  } else if (colmode) {
    ti = blockIdx.x >> 1; tj = blockIdx.x & 1;
  } else {
    const int64_t b = blockIdx.x;
    ti = (int64_t)((sqrt(4.0 * (double)b + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) <= b) ++ti;
    while (ti * (ti + 1) > b) --ti;
    tj = b - ti * (ti + 1);
  }
  const int64_t row0 = r0 + ti * NB;                                   // first global row of the C tile
  const int64_t brow0 = (MODE == MODE_PANEL) ? 0 : r0 + tj * TN;      // first row of the B operand
  const int64_t col0 = (MODE == MODE_PANEL) ? ccol0 : r0 + tj * TN;   // first global column of the C tile
  const int64_t brows = (MODE == MODE_PANEL) ? NB : N;                // rows available in B

  double acc[MT][4][2];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int nchunks = (Kdim + KC - 1) / KC;
  auto load_chunk = [&](int c, int stage) {
    double* As = smem + (size_t)stage * STAGE_D;
    double* Bs = As + NB * KP;
    const int kc0 = c * KC;
    // rows x 16 doubles = rows x 8 16-byte pieces; 256 threads -> 4 pieces each for A (128 rows), TN / 32 for B
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int e = tid + it * 256;
      const int r = e >> 3, p = e & 7;
      const int kk = kc0 + p * 2;
      const bool kin = kk < Kdim;
      const int64_t ar = row0 + r;
      const bool pa = kin && ar < N;
      cp_async16(As + r * KP + p * 2, Apan + (pa ? ar * lda + kk : 0), pa);
      if (it < TN / 32) {
        const int64_t br = brow0 + r;
        const bool pb = kin && br < brows;
        cp_async16(Bs + r * KP + p * 2, Bmat + (pb ? br * ldb + kk : 0), pb);
      }
    }
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nchunks) load_chunk(s, s);
    cp_async_commit();
  }
  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int cn = c + STAGES - 1;
      if (cn < nchunks) load_chunk(cn, cn % STAGES);
      cp_async_commit();
    }
    const double* As = smem + (size_t)(c % STAGES) * STAGE_D;
    const double* Bs = As + NB * KP;
#pragma unroll
    for (int k4 = 0; k4 < KC; k4 += 4) {
      double af[MT], bf[4];
#pragma unroll
      for (int i = 0; i < MT; ++i) af[i] = As[(wm * WROWS + i * 8 + g) * KP + k4 + t];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = Bs[(wn * 32 + j * 8 + g) * KP + k4 + t];
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();  // every warp is done reading the staged operands (PANEL writes C == A in place)

  // epilogue: thread owns C[row0 + wm*WROWS + i*8 + g][col0 + wn*32 + j*8 + 2t + {0,1}]
  double ydot[MT];
#pragma unroll
  for (int i = 0; i < MT; ++i) ydot[i] = 0.0;
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    const int64_t r = row0 + wm * WROWS + i * 8 + g;
    if (r >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int lc = wn * 32 + j * 8 + 2 * t;
      const int64_t cc = col0 + lc;
      if (MODE == MODE_PANEL) {
        if (lc < Kdim) {
          double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
          *reinterpret_cast<double2*>(C + r * ldc + cc) = v;
          if (alpha) ydot[i] = fma(v.x, alpha[cc], fma(v.y, alpha[cc + 1], ydot[i]));
        }
      } else {
        if (cc > r) continue;  // strictly upper part of a diagonal tile (cc even, so cc+1 > r handled below)
        double2 v;
        if (MODE == MODE_SYRK_GEN) {
          v.x = rbf_entry(ks.X + cc * ks.Q, ks.X + r * ks.Q, ks.inv_l2, ks.Q, ks.variance);
          v.y = (cc + 1 <= r) ? rbf_entry(ks.X + (cc + 1) * ks.Q, ks.X + r * ks.Q, ks.inv_l2, ks.Q, ks.variance) : 0.0;
          if (cc == r) v.x += ks.jitter;
          if (cc + 1 == r) v.y += ks.jitter;
        } else {
          v = *reinterpret_cast<const double2*>(C + r * ldc + cc);
        }
        v.x -= acc[i][j][0];
        v.y -= acc[i][j][1];
        if (cc + 1 <= r) *reinterpret_cast<double2*>(C + r * ldc + cc) = v;
        else C[r * ldc + cc] = v.x;
      }
    }
  }
  if (MODE == MODE_PANEL && alpha) {
    // y[r] -= sum over the 4 column-warps and the 4 lanes of a quad: shared-memory reduction, fixed order
    double* red = smem;  // 128 rows x 16 partial sums (4 wn x 4 t)
#pragma unroll
    for (int i = 0; i < MT; ++i) red[(wm * WROWS + i * 8 + g) * 16 + wn * 4 + t] = ydot[i];
    __syncthreads();
    if (tid < NB) {
      const int64_t r = row0 + tid;
      if (r < N) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < 16; ++q) s += red[tid * 16 + q];
        y[r] -= s;
      }
    }
  }
}

__global__ void zero_upper_kernel(double* __restrict__ A, int64_t N, int64_t ld) {
  const int64_t i = blockIdx.y;
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < N && j > i) A[i * ld + j] = 0.0;
}

__global__ void gp_lml_kernel(const double* __restrict__ alpha, int64_t N, double* __restrict__ state) {
  __shared__ double red[40];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < N; i += blockDim.x) s = fma(alpha[i], alpha[i], s);
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    state[2] = s;                                                                 // |L^-1 y|^2
    state[3] = -0.5 * s - state[0] - 0.5 * (double)N * log(2.0 * M_PI);           // lml
  }
}

// X (n x R) <- Tinv (n x n used) . B_k ; then B_rest -= L21 X_k : simple CUDA-core kernels (the GP posterior
// path, GaussianProcess.hs:74,77, is a "next" row of SURVEY.md 8f; these keep gplvm_tri_solve_lower correct).
__global__ void __launch_bounds__(256) tri_diag_apply_kernel(const double* __restrict__ Tinv, int nb, const double* __restrict__ B,
                                                             double* __restrict__ Xo, int64_t R, int64_t k0) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  for (int i = 0; i < nb; ++i) {
    double s = 0.0;
    for (int k = 0; k <= i; ++k) s = fma(Tinv[i * NB + k], B[(k0 + k) * R + r], s);
    Xo[(k0 + i) * R + r] = s;
  }
}
__global__ void __launch_bounds__(256) tri_update_kernel(const double* __restrict__ L, int64_t ld, int64_t N, int64_t k0, int nb,
                                                         const double* __restrict__ Xo, double* __restrict__ B, int64_t R) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = k0 + nb + blockIdx.y;
  if (r >= R || i >= N) return;
  double s = 0.0;
  for (int k = 0; k < nb; ++k) s = fma(L[i * ld + k0 + k], Xo[(k0 + k) * R + r], s);
  B[i * R + r] -= s;
}
__global__ void __launch_bounds__(256) tri_inverse_kernel(const double* __restrict__ L, int64_t ld, int64_t k0, int nb,
                                                          double* __restrict__ Tinv, double* __restrict__ state) {
  extern __shared__ double sm[];
  double* Ls = sm;
  const int tid = threadIdx.x;
  for (int e = tid; e < nb * nb; e += blockDim.x) {
    const int i = e / nb, j = e % nb;
    Ls[i * LDS_DIAG + j] = (j <= i) ? L[(k0 + i) * ld + (k0 + j)] : 0.0;
  }
  __syncthreads();
  double* dinv = sm + NB * LDS_DIAG;
  for (int j = tid; j < nb; j += blockDim.x) {
    if (Ls[j * LDS_DIAG + j] == 0.0) state[1] = 1.0;
    dinv[j] = 1.0 / Ls[j * LDS_DIAG + j];
  }
  __syncthreads();
  tri_inverse_in_place(Ls, nb, dinv);
  for (int e = tid; e < NB * NB; e += blockDim.x) {
    const int i = e / NB, j = e % NB;
    Tinv[e] = (i < nb && j < nb) ? tinv_at(Ls, i, j) : 0.0;
  }
}

// Cached device workspace for the N x ld matrix (the largest request so far is kept until gplvm_dist_finalize or a
// larger request): an 8.6 GB cudaMalloc / cudaFree pair per call costs more than the panel factorisations.
struct Workspace {
  void* p = nullptr;
  size_t bytes = 0;
  int device = -1;
};
Workspace g_ws;

int workspace_get(size_t bytes, int device, double** out) {
  if (g_ws.p && (g_ws.bytes < bytes || g_ws.device != device)) {
    cudaFree(g_ws.p);
    g_ws = Workspace();
  }
  if (!g_ws.p) {
    CUDA_TRY(cudaMalloc(&g_ws.p, bytes));
    g_ws.bytes = bytes;
    g_ws.device = device;
  }
  *out = static_cast<double*>(g_ws.p);
  return GPLVM_OK;
}

// Leading dimension: even (16-byte aligned rows for the staged copies) and never a multiple of 256 doubles -- with a
// power-of-two row pitch the 128 rows of a tile fall into the same L2 slice / HBM channel (partition camping).
inline int64_t padded_ld(int64_t N) {
  int64_t ld = round_up(N, 2);
  if (ld % 256 == 0) ld += 16;
  return ld;
}

// second (high-priority) stream and events of the look-ahead factorisation, one set per device
int lookahead_resources(cudaStream_t* sp, cudaEvent_t* e0, cudaEvent_t* e1, cudaEvent_t* e2) {
  struct Res { cudaStream_t s = nullptr; cudaEvent_t e[3] = {nullptr, nullptr, nullptr}; };
  static Res res[64];
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  Res& r = res[dev & 63];
  if (!r.s) {
    int lo = 0, hi = 0;
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CUDA_TRY(cudaStreamCreateWithPriority(&r.s, cudaStreamNonBlocking, hi));
    for (int i = 0; i < 3; ++i) CUDA_TRY(cudaEventCreateWithFlags(&r.e[i], cudaEventDisableTiming));
  }
  *sp = r.s; *e0 = r.e[0]; *e1 = r.e[1]; *e2 = r.e[2];
  return GPLVM_OK;
}

// X (N x R) = L^-1 B on the device, blocked by NB rows: diagonal blocks are inverted in shared memory, the rows
// below are updated by a CUDA-core kernel (B is overwritten).  L is N x N lower, leading dimension ld.
int tri_solve_device(const double* dL, int64_t N, int64_t ld, double* dB, int64_t R, double* dX, double* dT, double* dstate,
                     cudaStream_t st) {
  for (int64_t k0 = 0; k0 < N; k0 += NB) {
    const int nb = (int)std::min<int64_t>(NB, N - k0);
    GP_LAUNCH(tri_inverse_kernel, 1, 256, DIAG_SMEM, st, dL, ld, k0, nb, dT, dstate);
    GP_LAUNCH(tri_diag_apply_kernel, (unsigned)ceil_div(R, 256), 256, 0, st, dT, nb, dB, dX, R, k0);
    const int64_t below = N - k0 - nb;
    if (below > 0) {
      for (int64_t b0 = 0; b0 < below; b0 += 32768) {
        dim3 grid((unsigned)ceil_div(R, 256), (unsigned)std::min<int64_t>(32768, below - b0));
        GP_LAUNCH(tri_update_kernel, grid, 256, 0, st, dL + b0 * ld, ld, N - b0, k0, nb, dX, dB + b0 * R, R);
      }
    }
  }
  return GPLVM_OK;
}

// X (N x R) = L^-T B (back substitution with the UPPER factor U = L^T: the solve hmatrix performs when cholSH hands
// linearSolveS the upper factor, see gplvm_gp_posterior_upper), blocked bottom-up.  Same simple CUDA-core kernels.
__global__ void __launch_bounds__(256) tri_diag_apply_t_kernel(const double* __restrict__ Tinv, int nb, const double* __restrict__ B,
                                                               double* __restrict__ Xo, int64_t R, int64_t k0) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  for (int i = 0; i < nb; ++i) {   // (L_kk^-T)[i][k] = Tinv[k][i], k >= i
    double s = 0.0;
    for (int k = i; k < nb; ++k) s = fma(Tinv[k * NB + i], B[(k0 + k) * R + r], s);
    Xo[(k0 + i) * R + r] = s;
  }
}
__global__ void __launch_bounds__(256) tri_update_t_kernel(const double* __restrict__ L, int64_t ld, int64_t k0, int nb,
                                                           const double* __restrict__ Xo, double* __restrict__ B, int64_t R, int64_t i0) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = i0 + blockIdx.y;   // a row above the block: B[i] -= sum_k L[k0 + k][i] X[k0 + k]
  if (r >= R || i >= k0) return;
  double s = 0.0;
  for (int k = 0; k < nb; ++k) s = fma(L[(k0 + k) * ld + i], Xo[(k0 + k) * R + r], s);
  B[i * R + r] -= s;
}
int tri_solve_t_device(const double* dL, int64_t N, int64_t ld, double* dB, int64_t R, double* dX, double* dT, double* dstate,
                       cudaStream_t st) {
  const int64_t nblocks = ceil_div(N, NB);
  for (int64_t b = nblocks - 1; b >= 0; --b) {
    const int64_t k0 = b * NB;
    const int nb = (int)std::min<int64_t>(NB, N - k0);
    GP_LAUNCH(tri_inverse_kernel, 1, 256, DIAG_SMEM, st, dL, ld, k0, nb, dT, dstate);
    GP_LAUNCH(tri_diag_apply_t_kernel, (unsigned)ceil_div(R, 256), 256, 0, st, dT, nb, dB, dX, R, k0);
    for (int64_t i0 = 0; i0 < k0; i0 += 32768) {
      dim3 grid((unsigned)ceil_div(R, 256), (unsigned)std::min<int64_t>(32768, k0 - i0));
      GP_LAUNCH(tri_update_t_kernel, grid, 256, 0, st, dL, ld, k0, nb, dX, dB, R, i0);
    }
  }
  return GPLVM_OK;
}

// posterior pieces of gpToPosteriorSample (GaussianProcess.hs:80-98)
// mean[i] = sum_k alpha[k] Lk[k][i]                                                     (:80)
__global__ void __launch_bounds__(256) gp_post_mean_kernel(const double* __restrict__ alpha, const double* __restrict__ Lk, int64_t nt,
                                                           int64_t no, double* __restrict__ mean) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= no) return;
  double s = 0.0;
  for (int64_t k = 0; k < nt; ++k) s = fma(alpha[k], Lk[k * no + i], s);
  mean[i] = s;
}
// LkT (no x ldt, zero padded columns) = Lk^T ;  C (no x ldc, lower) = K_ss + jitter I                        (:83-85)
__global__ void __launch_bounds__(256) gp_post_prepare_kernel(const double* __restrict__ Lk, int64_t nt, int64_t no, double* __restrict__ LkT,
                                                              int64_t ldt, double* __restrict__ Cm, int64_t ldc, KernelSpec ks) {
  const int64_t i = blockIdx.y;
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < ldt) LkT[i * ldt + j] = (j < nt) ? Lk[j * no + i] : 0.0;
  if (j <= i) {
    double v = rbf_entry(ks.X + j * ks.Q, ks.X + i * ks.Q, ks.inv_l2, ks.Q, ks.variance);
    if (i == j) v += ks.jitter;
    Cm[i * ldc + j] = v;
  }
}
// samples (no x S) = mean 1^T + P normals, P lower (no x ldc)                                                  (:89,95-98)
__global__ void __launch_bounds__(256) gp_post_sample_kernel(const double* __restrict__ P, int64_t ldc, int64_t no, const double* __restrict__ mean,
                                                             const double* __restrict__ normals, int64_t S, double* __restrict__ out) {
  const int64_t i = blockIdx.y;
  const int64_t sidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (sidx >= S) return;
  double acc = mean[i];
  for (int64_t k = 0; k <= i; ++k) acc = fma(P[i * ldc + k], normals[k * S + sidx], acc);
  out[i * S + sidx] = acc;
}

// samples (no x S) = mean 1^T + U normals with U = P^T upper (P lower, no x ldc)
__global__ void __launch_bounds__(256) gp_post_sample_upper_kernel(const double* __restrict__ P, int64_t ldc, int64_t no,
                                                                   const double* __restrict__ mean, const double* __restrict__ normals,
                                                                   int64_t S, double* __restrict__ out) {
  const int64_t i = blockIdx.y;
  const int64_t sidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (sidx >= S) return;
  double acc = mean[i];
  for (int64_t k = i; k < no; ++k) acc = fma(P[k * ldc + i], normals[k * S + sidx], acc);
  out[i * S + sidx] = acc;
}
// U (N x N dense, row-major) = L^T with the strict lower part zero
__global__ void gp_transpose_factor_kernel(const double* __restrict__ L, int64_t ld, int64_t N, double* __restrict__ U) {
  const int64_t i = blockIdx.y;
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < N) U[i * N + j] = (j >= i) ? L[j * ld + i] : 0.0;
}

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  template <class T> T* as() { return static_cast<T*>(p); }
};

int set_gemm_attrs() {
  static bool done_dev[64] = {};
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  bool& done = done_dev[dev & 63];
  if (done) return GPLVM_OK;
  CUDA_TRY(cudaFuncSetAttribute((gp_gemm_kernel<MODE_PANEL, 128>), cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
  CUDA_TRY(cudaFuncSetAttribute((gp_gemm_kernel<MODE_SYRK, 64>), cudaFuncAttributeMaxDynamicSharedMemorySize, SYRK_SMEM));
  CUDA_TRY(cudaFuncSetAttribute((gp_gemm_kernel<MODE_SYRK_GEN, 64>), cudaFuncAttributeMaxDynamicSharedMemorySize, SYRK_SMEM));
  CUDA_TRY(cudaFuncSetAttribute((gp_gemm_kernel<MODE_SYRK, 64>), cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  CUDA_TRY(cudaFuncSetAttribute((gp_gemm_kernel<MODE_SYRK_GEN, 64>), cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  const int diag_smem = DIAG_SMEM;
  CUDA_TRY(cudaFuncSetAttribute(gp_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, diag_smem));
  CUDA_TRY(cudaFuncSetAttribute(tri_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, diag_smem));
  done = true;
  return GPLVM_OK;
}

// Factorise the N x N matrix in dA (row-major, lower part) in place.  gen != nullptr: the matrix is the
// kernel matrix of gen->X and is generated on the fly (dA's lower part is written, never read before).
// dy (nullable): right-hand side overwritten by the forward substitution; dalpha receives L^-1 y.
int factorise(double* dA, int64_t N, int64_t ld, const KernelSpec* gen, double* dy, double* dalpha, double* dTinv, double* dstate,
              cudaStream_t st) {
  GP_TRY(set_gemm_attrs());
  const int diag_smem = DIAG_SMEM;
  KernelSpec ks{};
  if (gen) ks = *gen;
  if (gen) {
    const int nb0 = (int)std::min<int64_t>(NB, N);
    GP_LAUNCH(gp_first_panel_kernel, (unsigned)N, 128, 0, st, dA, ld, nb0, ks);
  }
  // Two-level blocking: outer panels of NI NB columns (NI = 8: 1024; N = 32768: 435 / 422 / 414 ms at NI = 2 / 4 / 8).  Inside an outer panel the NI inner panels are
  // factorised one after the other (block-column update with the inner panels before it, diag, panel GEMM); the trailing
  // matrix is then updated ONCE with K = NI NB, which divides the read-modify-write traffic of the big update and its
  // prologue / epilogue overhead per flop by NI (measured on the trailing update: 77 % of the DMMA peak at K = 256).
  // The panel chain of an outer step (NI diagonal blocks at 280 us each, NI panel GEMMs, block-column and head updates)
  // runs on the look-ahead stream and must hide under the tail update: 6.7 ms per step at NI = 8, which a K = 1024 tail
  // covers only while more than ~14 000 columns remain.  NI therefore shrinks with the remaining matrix: 8 / 4 / 2.
  static const int NI_FIXED = []() {
    const char* e = getenv("GPLVM_GP_INNER");
    const int v = e ? atoi(e) : 0;
    return (v >= 1 && v <= 32) ? v : 0;
  }();
  auto panel = [&](cudaStream_t sx, int64_t k0, int nb) -> int {   // L21 = A21 T^T for the rows below the diagonal block at k0
    const int64_t r0 = k0 + nb;
    const int64_t T = ceil_div(N - r0, NB);
    GP_LAUNCH((gp_gemm_kernel<MODE_PANEL, 128>), (unsigned)T, 256, GEMM_SMEM, sx, dA, ld, dA + k0, ld, dTinv, (int64_t)NB, r0, N, k0,
              (int)round_up(nb, 4), dy, dy ? dalpha : nullptr, ks, 0);
    return GPLVM_OK;
  };
  auto update = [&](cudaStream_t sx, int64_t kpan, int Kdim, int64_t r0, bool column_only, bool generate) -> int {
    const int64_t T = ceil_div(N - r0, NB);
    const int64_t ntiles = column_only ? 2 * T : T * (T + 1);
    if (generate)
      GP_LAUNCH((gp_gemm_kernel<MODE_SYRK_GEN, 64>), (unsigned)ntiles, 256, SYRK_SMEM, sx, dA, ld, dA + kpan, ld, dA + kpan, ld, r0, N, 0,
                Kdim, nullptr, nullptr, ks, column_only ? 1 : 0);
    else
      GP_LAUNCH((gp_gemm_kernel<MODE_SYRK, 64>), (unsigned)ntiles, 256, SYRK_SMEM, sx, dA, ld, dA + kpan, ld, dA + kpan, ld, r0, N, 0, Kdim,
                nullptr, nullptr, ks, column_only ? 1 : 0);
    return GPLVM_OK;
  };
  // Look-ahead over two streams.  `sp` (high priority) runs the latency-bound panel work of outer step k + 1 (NI
  // diagonal blocks, NI panel GEMMs, NI - 1 block-column updates) and the "head" of update k (the next NI block columns);
  // `st` runs the "tail" of update k (everything right of the head) -- the bulk of the flops -- back to back.
  //   head(k) needs tail(k-1) (same region) and panel(k);  tail(k) needs panel(k) and tail(k-1);
  //   panel(k+1) needs head(k) only, so it hides under tail(k).
  cudaStream_t sp = nullptr;
  cudaEvent_t ev_panel = nullptr, ev_tail = nullptr, ev_start = nullptr;
  GP_TRY(lookahead_resources(&sp, &ev_panel, &ev_tail, &ev_start));
  CUDA_TRY(cudaEventRecord(ev_start, st));
  CUDA_TRY(cudaStreamWaitEvent(sp, ev_start, 0));
  bool tail_pending = false;
  int NI = 2;
  for (int64_t K0 = 0; K0 < N; K0 += (int64_t)NI * NB) {
    const int64_t left = N - K0;
    static const int64_t T8 = getenv("GPLVM_GP_T8") ? atoll(getenv("GPLVM_GP_T8")) : 16384;
    static const int64_t T4 = getenv("GPLVM_GP_T4") ? atoll(getenv("GPLVM_GP_T4")) : 8192;
    static const int64_t T2 = getenv("GPLVM_GP_T2") ? atoll(getenv("GPLVM_GP_T2")) : 0;
    NI = NI_FIXED ? NI_FIXED : (left > T8 ? 8 : (left > T4 ? 4 : (left > T2 ? 2 : 1)));
    const bool generate = gen && K0 == 0;
    int64_t c0 = K0;
    bool done = false;
    for (int ip = 0; ip < NI; ++ip) {
      const int nb = (int)std::min<int64_t>(NB, N - c0);
      if (ip > 0) GP_TRY(update(sp, K0, (int)(c0 - K0), c0, true, generate));   // block column [c0, c0 + NB), K = ip NB
      GP_LAUNCH(gp_diag_kernel, 1, 256, diag_smem, sp, dA, ld, c0, nb, dTinv, dy, dalpha, dstate);
      c0 += nb;
      if (c0 >= N) { done = true; break; }
      GP_TRY(panel(sp, c0 - nb, nb));
    }
    if (done) break;
    const int Kdim = (int)(c0 - K0);   // = NI NB
    CUDA_TRY(cudaEventRecord(ev_panel, sp));
    // head: the NI block columns the next outer step factorises
    if (tail_pending) CUDA_TRY(cudaStreamWaitEvent(sp, ev_tail, 0));
    for (int ib = 0; ib < NI && c0 + (int64_t)ib * NB < N; ++ib) GP_TRY(update(sp, K0, Kdim, c0 + (int64_t)ib * NB, true, generate));
    // tail: the trailing matrix right of the head
    if (c0 + (int64_t)NI * NB < N) {
      CUDA_TRY(cudaStreamWaitEvent(st, ev_panel, 0));
      GP_TRY(update(st, K0, Kdim, c0 + (int64_t)NI * NB, false, generate));
      CUDA_TRY(cudaEventRecord(ev_tail, st));
      tail_pending = true;
    }
  }
  CUDA_TRY(cudaEventRecord(ev_panel, sp));
  CUDA_TRY(cudaStreamWaitEvent(st, ev_panel, 0));  // later work on `st` (LML, copies) sees the whole factor
  return GPLVM_OK;
}

int copy_lower_to_host(double* dA, int64_t N, int64_t ld, double* L_out, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(N, 256), (unsigned)N);
  GP_LAUNCH(zero_upper_kernel, grid, 256, 0, st, dA, N, ld);
  CUDA_TRY(cudaMemcpy2DAsync(L_out, sizeof(double) * N, dA, sizeof(double) * ld, sizeof(double) * N, N, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return GPLVM_OK;
}

}  // namespace

void gp_release_workspace() {
  if (g_ws.p) cudaFree(g_ws.p);
  g_ws = Workspace();
}

}  // namespace gplvm

using namespace gplvm;

extern "C" {

int gplvm_rbf_kernel(const double* X1, int64_t n1, const double* X2, int64_t n2, int64_t Q, const double* inv_lengthscale2,
                     double variance, double* K_out) {
  if (!X1 || !X2 || !inv_lengthscale2 || !K_out) return fail(GPLVM_ERR_ARG, "gplvm_rbf_kernel: null argument");
  if (n1 < 1 || n2 < 1 || Q < 1 || Q > 64) return fail(GPLVM_ERR_ARG, "gplvm_rbf_kernel: sizes must be positive (Q <= 64)");
  GP_TRY(ensure_context());
  Context& c = ctx();
  std::lock_guard<std::recursive_mutex> lk(c.mu);
  CUDA_TRY(cudaSetDevice(c.shards[0].device));
  cudaStream_t st = c.shards[0].stream;
  DevBuf d1, d2, dl, dk;
  // rows of K_out are produced in chunks so that any n2 x n1 fits next to the inputs
  const int64_t rows_per_chunk = std::max<int64_t>(1, std::min<int64_t>(n2, ((int64_t)1 << 28) / n1));
  CUDA_TRY(cudaMalloc(&d1.p, sizeof(double) * n1 * Q));
  CUDA_TRY(cudaMalloc(&d2.p, sizeof(double) * n2 * Q));
  CUDA_TRY(cudaMalloc(&dl.p, sizeof(double) * Q));
  CUDA_TRY(cudaMalloc(&dk.p, sizeof(double) * rows_per_chunk * n1));
  CUDA_TRY(cudaMemcpyAsync(d1.p, X1, sizeof(double) * n1 * Q, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d2.p, X2, sizeof(double) * n2 * Q, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(dl.p, inv_lengthscale2, sizeof(double) * Q, cudaMemcpyHostToDevice, st));
  for (int64_t j0 = 0; j0 < n2; j0 += rows_per_chunk) {
    const int64_t jc = std::min(rows_per_chunk, n2 - j0);
    for (int64_t jj = 0; jj < jc; jj += 32768) {  // grid.y limit
      const int64_t jn = std::min<int64_t>(32768, jc - jj);
      dim3 grid((unsigned)ceil_div(n1, 512), (unsigned)jn);
      GP_LAUNCH(rbf_build_kernel, grid, 256, 0, st, d1.as<double>(), n1, d2.as<double>(), n2, (int)Q, dl.as<double>(), variance,
                dk.as<double>() + jj * n1, n1, j0 + jj, jn);
    }
    CUDA_TRY(cudaMemcpyAsync(K_out + j0 * n1, dk.p, sizeof(double) * jc * n1, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  CUDA_TRY(cudaGetLastError());
  return GPLVM_OK;
}

int gplvm_gp_chol_lml(const double* X, int64_t N, int64_t Q, const double* y, const double* inv_lengthscale2, double variance,
                      double jitter, double* L_out, double* lml_out, double* logdet_out) {
  if (!X || !inv_lengthscale2) return fail(GPLVM_ERR_ARG, "gplvm_gp_chol_lml: null argument");
  if (N < 1 || Q < 1 || Q > 64) return fail(GPLVM_ERR_ARG, "gplvm_gp_chol_lml: sizes must be positive (Q <= 64)");
  if (lml_out && !y) return fail(GPLVM_ERR_ARG, "gplvm_gp_chol_lml: y is required for the log marginal likelihood");
  GP_TRY(ensure_context());
  Context& c = ctx();
  std::lock_guard<std::recursive_mutex> lk(c.mu);
  CUDA_TRY(cudaSetDevice(c.shards[0].device));
  cudaStream_t st = c.shards[0].stream;
  DevBuf dX, dl, dA, dy, dal, dT, dst;
  const int64_t ld = padded_ld(N);
  CUDA_TRY(cudaMalloc(&dX.p, sizeof(double) * N * Q));
  CUDA_TRY(cudaMalloc(&dl.p, sizeof(double) * Q));
  double* dAp = nullptr;
  GP_TRY(workspace_get(sizeof(double) * (size_t)N * ld, c.shards[0].device, &dAp));
  CUDA_TRY(cudaMalloc(&dy.p, sizeof(double) * N));
  CUDA_TRY(cudaMalloc(&dal.p, sizeof(double) * N));
  CUDA_TRY(cudaMalloc(&dT.p, sizeof(double) * NB * NB));
  CUDA_TRY(cudaMalloc(&dst.p, sizeof(double) * 8));
  CUDA_TRY(cudaMemsetAsync(dst.p, 0, sizeof(double) * 8, st));
  CUDA_TRY(cudaMemcpyAsync(dX.p, X, sizeof(double) * N * Q, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(dl.p, inv_lengthscale2, sizeof(double) * Q, cudaMemcpyHostToDevice, st));
  if (y) CUDA_TRY(cudaMemcpyAsync(dy.p, y, sizeof(double) * N, cudaMemcpyHostToDevice, st));
  KernelSpec ks{dX.as<double>(), dl.as<double>(), variance, jitter, (int)Q};
  GP_TRY(factorise(dAp, N, ld, &ks, y ? dy.as<double>() : nullptr, dal.as<double>(), dT.as<double>(), dst.as<double>(), st));
  if (y) GP_LAUNCH(gp_lml_kernel, 1, 256, 0, st, dal.as<double>(), N, dst.as<double>());
  double state[8];
  CUDA_TRY(cudaMemcpyAsync(state, dst.p, sizeof state, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  if (state[1] != 0.0)
    return fail(GPLVM_ERR_NUMERIC, "gplvm_gp_chol_lml: kernel matrix not positive definite (panel %d)", (int)state[1] - 1);
  if (lml_out) *lml_out = state[3];
  if (logdet_out) *logdet_out = 2.0 * state[0];
  if (L_out) GP_TRY(copy_lower_to_host(dAp, N, ld, L_out, st));
  return GPLVM_OK;
}

int gplvm_cholesky(const double* A, int64_t N, double* L_out) {
  if (!A || !L_out) return fail(GPLVM_ERR_ARG, "gplvm_cholesky: null argument");
  if (N < 1) return fail(GPLVM_ERR_ARG, "gplvm_cholesky: N must be positive");
  GP_TRY(ensure_context());
  Context& c = ctx();
  std::lock_guard<std::recursive_mutex> lk(c.mu);
  CUDA_TRY(cudaSetDevice(c.shards[0].device));
  cudaStream_t st = c.shards[0].stream;
  DevBuf dA, dT, dst;
  const int64_t ld = padded_ld(N);
  CUDA_TRY(cudaMalloc(&dA.p, sizeof(double) * (size_t)N * ld));
  CUDA_TRY(cudaMalloc(&dT.p, sizeof(double) * NB * NB));
  CUDA_TRY(cudaMalloc(&dst.p, sizeof(double) * 8));
  CUDA_TRY(cudaMemsetAsync(dst.p, 0, sizeof(double) * 8, st));
  CUDA_TRY(cudaMemcpy2DAsync(dA.p, sizeof(double) * ld, A, sizeof(double) * N, sizeof(double) * N, N, cudaMemcpyHostToDevice, st));
  GP_TRY(factorise(dA.as<double>(), N, ld, nullptr, nullptr, nullptr, dT.as<double>(), dst.as<double>(), st));
  double state[8];
  CUDA_TRY(cudaMemcpyAsync(state, dst.p, sizeof state, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  if (state[1] != 0.0) return fail(GPLVM_ERR_NUMERIC, "gplvm_cholesky: matrix not positive definite (panel %d)", (int)state[1] - 1);
  return copy_lower_to_host(dA.as<double>(), N, ld, L_out, st);
}

int gplvm_tri_solve_lower(const double* L, int64_t N, const double* B, int64_t R, double* X_out) {
  if (!L || !B || !X_out) return fail(GPLVM_ERR_ARG, "gplvm_tri_solve_lower: null argument");
  if (N < 1 || R < 1) return fail(GPLVM_ERR_ARG, "gplvm_tri_solve_lower: sizes must be positive");
  GP_TRY(ensure_context());
  Context& c = ctx();
  std::lock_guard<std::recursive_mutex> lk(c.mu);
  CUDA_TRY(cudaSetDevice(c.shards[0].device));
  cudaStream_t st = c.shards[0].stream;
  GP_TRY(set_gemm_attrs());
  DevBuf dL, dB, dX, dT, dst;
  CUDA_TRY(cudaMalloc(&dL.p, sizeof(double) * (size_t)N * N));
  CUDA_TRY(cudaMalloc(&dB.p, sizeof(double) * (size_t)N * R));
  CUDA_TRY(cudaMalloc(&dX.p, sizeof(double) * (size_t)N * R));
  CUDA_TRY(cudaMalloc(&dT.p, sizeof(double) * NB * NB));
  CUDA_TRY(cudaMalloc(&dst.p, sizeof(double) * 8));
  CUDA_TRY(cudaMemsetAsync(dst.p, 0, sizeof(double) * 8, st));
  CUDA_TRY(cudaMemcpyAsync(dL.p, L, sizeof(double) * (size_t)N * N, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(dB.p, B, sizeof(double) * (size_t)N * R, cudaMemcpyHostToDevice, st));
  GP_TRY(tri_solve_device(dL.as<double>(), N, N, dB.as<double>(), R, dX.as<double>(), dT.as<double>(), dst.as<double>(), st));
  double state[8];
  CUDA_TRY(cudaMemcpyAsync(state, dst.p, sizeof state, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(X_out, dX.p, sizeof(double) * (size_t)N * R, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  if (state[1] != 0.0) return fail(GPLVM_ERR_NUMERIC, "gplvm_tri_solve_lower: singular triangular factor");
  return GPLVM_OK;
}


}  // extern "C"

// upper = false: both Cholesky factors lower (K = L L^T), forward substitutions -- the textbook GP posterior.
// upper = true : the reference read literally with hmatrix's UPPER factor U = L^T (K = U^T U): linearSolveS cholK b solves
//                U x = b (GaussianProcess.hs:74,77), the mean is (U^-1 y)^T (U^-1 K_s) (:80), the posterior factor is the upper
//                factor of K_ss + jitter I - Z^T Z (:83-86) and the samples are mean + U_post normals (:89).
static int gp_posterior_impl(bool upper, const double* X_obs, int64_t n_obs, const double* X_train, const double* y_train, int64_t n_train,
                             int64_t Q, const double* inv_lengthscale2, double variance, double jitter_train, double jitter_post,
                             const double* normals, int64_t n_samples, double* mean_out, double* postfactor_out, double* samples_out) {
  if (!X_obs || !X_train || !y_train || !inv_lengthscale2 || !mean_out) return fail(GPLVM_ERR_ARG, "gplvm_gp_posterior: null argument");
  if (n_obs < 1 || n_train < 1 || Q < 1 || Q > 64) return fail(GPLVM_ERR_ARG, "gplvm_gp_posterior: sizes must be positive (Q <= 64)");
  if (samples_out && (!normals || n_samples < 1)) return fail(GPLVM_ERR_ARG, "gplvm_gp_posterior: normals required for samples");
  if (n_obs > 32768) return fail(GPLVM_ERR_ARG, "gplvm_gp_posterior: n_obs > 32768 is not supported");
  GP_TRY(ensure_context());
  Context& c = ctx();
  std::lock_guard<std::recursive_mutex> lk(c.mu);
  CUDA_TRY(cudaSetDevice(c.shards[0].device));
  cudaStream_t st = c.shards[0].stream;
  GP_TRY(set_gemm_attrs());
  const int64_t ldk = padded_ld(n_train), ldc = padded_ld(n_obs), ldt = round_up(n_train, 4);
  DevBuf dXo, dXt, dl, dK, dy, dal, dT, dst, dKs, dLk, dLkT, dC, dmean, dnorm, dsamp, dU;
  CUDA_TRY(cudaMalloc(&dXo.p, sizeof(double) * n_obs * Q));
  CUDA_TRY(cudaMalloc(&dXt.p, sizeof(double) * n_train * Q));
  CUDA_TRY(cudaMalloc(&dl.p, sizeof(double) * Q));
  CUDA_TRY(cudaMalloc(&dK.p, sizeof(double) * (size_t)n_train * ldk));
  CUDA_TRY(cudaMalloc(&dy.p, sizeof(double) * n_train));
  CUDA_TRY(cudaMalloc(&dal.p, sizeof(double) * n_train));
  CUDA_TRY(cudaMalloc(&dT.p, sizeof(double) * NB * NB));
  CUDA_TRY(cudaMalloc(&dst.p, sizeof(double) * 8));
  CUDA_TRY(cudaMalloc(&dKs.p, sizeof(double) * (size_t)n_train * n_obs));
  CUDA_TRY(cudaMalloc(&dLk.p, sizeof(double) * (size_t)n_train * n_obs));
  CUDA_TRY(cudaMalloc(&dLkT.p, sizeof(double) * (size_t)n_obs * ldt));
  CUDA_TRY(cudaMalloc(&dC.p, sizeof(double) * (size_t)n_obs * ldc));
  CUDA_TRY(cudaMalloc(&dmean.p, sizeof(double) * n_obs));
  CUDA_TRY(cudaMemsetAsync(dst.p, 0, sizeof(double) * 8, st));
  CUDA_TRY(cudaMemcpyAsync(dXo.p, X_obs, sizeof(double) * n_obs * Q, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(dXt.p, X_train, sizeof(double) * n_train * Q, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(dl.p, inv_lengthscale2, sizeof(double) * Q, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(dy.p, y_train, sizeof(double) * n_train, cudaMemcpyHostToDevice, st));
  // L = chol(K(train, train) + jitter I), alpha = L^-1 y                                   (GaussianProcess.hs:64-68,77)
  KernelSpec kt{dXt.as<double>(), dl.as<double>(), variance, jitter_train, (int)Q};
  if (!upper) {
    GP_TRY(factorise(dK.as<double>(), n_train, ldk, &kt, dy.as<double>(), dal.as<double>(), dT.as<double>(), dst.as<double>(), st));
  } else {
    GP_TRY(factorise(dK.as<double>(), n_train, ldk, &kt, nullptr, nullptr, dT.as<double>(), dst.as<double>(), st));
    // a = U^-1 y = L^-T y                                                                  (GaussianProcess.hs:77, upper reading)
    GP_TRY(tri_solve_t_device(dK.as<double>(), n_train, ldk, dy.as<double>(), 1, dal.as<double>(), dT.as<double>(), dst.as<double>(), st));
  }
  // K_s = k(observe, train): n_train x n_obs (orientation of sqDist, GPExample.hs:83)     (GaussianProcess.hs:71)
  for (int64_t jj = 0; jj < n_train; jj += 32768) {
    const int64_t jn = std::min<int64_t>(32768, n_train - jj);
    dim3 grid((unsigned)ceil_div(n_obs, 512), (unsigned)jn);
    GP_LAUNCH(rbf_build_kernel, grid, 256, 0, st, dXo.as<double>(), n_obs, dXt.as<double>(), n_train, (int)Q, dl.as<double>(), variance,
              dKs.as<double>() + jj * n_obs, n_obs, jj, jn);
  }
  // Lk = L^-1 K_s                                                                         (GaussianProcess.hs:74)
  if (!upper)
    GP_TRY(tri_solve_device(dK.as<double>(), n_train, ldk, dKs.as<double>(), n_obs, dLk.as<double>(), dT.as<double>(), dst.as<double>(), st));
  else   // Z = U^-1 K_s = L^-T K_s
    GP_TRY(tri_solve_t_device(dK.as<double>(), n_train, ldk, dKs.as<double>(), n_obs, dLk.as<double>(), dT.as<double>(), dst.as<double>(), st));
  GP_LAUNCH(gp_post_mean_kernel, (unsigned)ceil_div(n_obs, 256), 256, 0, st, dal.as<double>(), dLk.as<double>(), n_train, n_obs,
            dmean.as<double>());
  CUDA_TRY(cudaMemcpyAsync(mean_out, dmean.p, sizeof(double) * n_obs, cudaMemcpyDeviceToHost, st));
  if (postfactor_out || samples_out) {
    // P = chol(K_ss + jitter_post I - Lk^T Lk)                                            (GaussianProcess.hs:83-86)
    KernelSpec ko{dXo.as<double>(), dl.as<double>(), variance, jitter_post, (int)Q};
    {
      dim3 grid((unsigned)ceil_div(std::max(ldt, n_obs), 256), (unsigned)n_obs);
      GP_LAUNCH(gp_post_prepare_kernel, grid, 256, 0, st, dLk.as<double>(), n_train, n_obs, dLkT.as<double>(), ldt, dC.as<double>(), ldc,
                ko);
    }
    {
      const int64_t T = ceil_div(n_obs, NB);
      GP_LAUNCH((gp_gemm_kernel<MODE_SYRK, 64>), (unsigned)(T * (T + 1)), 256, SYRK_SMEM, st, dC.as<double>(), ldc, dLkT.as<double>(), ldt,
                dLkT.as<double>(), ldt, (int64_t)0, n_obs, (int64_t)0, (int)ldt, nullptr, nullptr, ko, 0);
    }
    GP_TRY(factorise(dC.as<double>(), n_obs, ldc, nullptr, nullptr, nullptr, dT.as<double>(), dst.as<double>() + 4, st));
    if (postfactor_out && !upper) GP_TRY(copy_lower_to_host(dC.as<double>(), n_obs, ldc, postfactor_out, st));
    if (postfactor_out && upper) {
      CUDA_TRY(cudaMalloc(&dU.p, sizeof(double) * (size_t)n_obs * n_obs));
      dim3 grid((unsigned)ceil_div(n_obs, 256), (unsigned)n_obs);
      GP_LAUNCH(gp_transpose_factor_kernel, grid, 256, 0, st, dC.as<double>(), ldc, n_obs, dU.as<double>());
      CUDA_TRY(cudaMemcpyAsync(postfactor_out, dU.p, sizeof(double) * (size_t)n_obs * n_obs, cudaMemcpyDeviceToHost, st));
    }
    if (samples_out) {
      CUDA_TRY(cudaMalloc(&dnorm.p, sizeof(double) * (size_t)n_obs * n_samples));
      CUDA_TRY(cudaMalloc(&dsamp.p, sizeof(double) * (size_t)n_obs * n_samples));
      CUDA_TRY(cudaMemcpyAsync(dnorm.p, normals, sizeof(double) * (size_t)n_obs * n_samples, cudaMemcpyHostToDevice, st));
      dim3 grid((unsigned)ceil_div(n_samples, 256), (unsigned)n_obs);
      if (!upper)
        GP_LAUNCH(gp_post_sample_kernel, grid, 256, 0, st, dC.as<double>(), ldc, n_obs, dmean.as<double>(), dnorm.as<double>(), n_samples,
                  dsamp.as<double>());
      else
        GP_LAUNCH(gp_post_sample_upper_kernel, grid, 256, 0, st, dC.as<double>(), ldc, n_obs, dmean.as<double>(), dnorm.as<double>(),
                  n_samples, dsamp.as<double>());
      CUDA_TRY(cudaMemcpyAsync(samples_out, dsamp.p, sizeof(double) * (size_t)n_obs * n_samples, cudaMemcpyDeviceToHost, st));
    }
  }
  double state[8];
  CUDA_TRY(cudaMemcpyAsync(state, dst.p, sizeof state, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  if (state[1] != 0.0 || state[5] != 0.0)
    return fail(GPLVM_ERR_NUMERIC, "gplvm_gp_posterior: matrix not positive definite (the reference's cholSH throws; Maybe result Nothing)");
  return GPLVM_OK;
}

extern "C" {

int gplvm_gp_posterior(const double* X_obs, int64_t n_obs, const double* X_train, const double* y_train, int64_t n_train, int64_t Q,
                       const double* inv_lengthscale2, double variance, double jitter_train, double jitter_post,
                       const double* normals, int64_t n_samples, double* mean_out, double* postfactor_out, double* samples_out) {
  return gp_posterior_impl(false, X_obs, n_obs, X_train, y_train, n_train, Q, inv_lengthscale2, variance, jitter_train, jitter_post, normals,
                           n_samples, mean_out, postfactor_out, samples_out);
}

int gplvm_gp_posterior_upper(const double* X_obs, int64_t n_obs, const double* X_train, const double* y_train, int64_t n_train, int64_t Q,
                             const double* inv_lengthscale2, double variance, double jitter_train, double jitter_post,
                             const double* normals, int64_t n_samples, double* mean_out, double* postfactor_out, double* samples_out) {
  return gp_posterior_impl(true, X_obs, n_obs, X_train, y_train, n_train, Q, inv_lengthscale2, variance, jitter_train, jitter_post, normals,
                           n_samples, mean_out, postfactor_out, samples_out);
}

}  // extern "C"
