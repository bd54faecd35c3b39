// pca.cu -- PCA by eigendecomposition (makePCA, reference src/GPLVM/PCA.hs:35-62; typed twin
// src/GPLVM/TypeSafe/PCA.hs:56-92), SURVEY.md 8(f) row 4.
//
//   (mean, C) = meanCovS X            mean over the ROWS of X (observations), C = Xc^T Xc / (N - 1)      (PCA.hs:42)
//   (lambda, E) = eigSH C             eigenvalues descending, eigenvectors in the columns of E            (PCA.hs:45)
//   E_d = first d columns             (all of them when d >= D)                                            (PCA.hs:48-53)
//   final    = Xc . E_d               N x d: data in the eigenvector basis                                 (PCA.hs:55-56)
//   restored = final . E_d^T + mean   pinvS of a matrix with orthonormal rows is its transpose             (PCA.hs:58-60)
//
// Unlike the PPCA entry points the reference's PCA takes the observations as ROWS (N x D, hmatrix meanCov), which is
// the device layout of this library, so the input is uploaded as is.
// Kernels: fixed-order column sums, centring, a split-K CUDA-core SYRK for the D x D covariance (the one pass over the
// data that costs flops: N D^2), a one-CTA cyclic Jacobi eigensolver (round-robin ordering: D/2 disjoint rotations per
// round, applied to rows, columns and the eigenvector matrix by all threads), and two streaming products for
// `final` / `restored`.  Eigenvector signs are a convention (LAPACK's is not reproducible): each column is normalised so
// that its entry of largest magnitude is positive.
#include <algorithm>
#include <numeric>

#include "common.cuh"

namespace gplvm {
namespace {

struct Buf {
  void* p = nullptr;
  ~Buf() { if (p) cudaFree(p); }
  template <class T> T* as() { return static_cast<T*>(p); }
};

// partial[split][d] = sum over the split's observations of X[n][d]   (fixed order inside a split)
__global__ void __launch_bounds__(256) pca_colsum_kernel(const double* __restrict__ X, int64_t ld, int64_t n_obs, int D,
                                                         double* __restrict__ partial, int64_t per_split) {
  __shared__ double red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int d = blockIdx.x * 32 + tx;
  const int64_t nb = (int64_t)blockIdx.y * per_split, ne = min(n_obs, nb + per_split);
  double a = 0.0;
  if (d < D)
    for (int64_t n = nb + ty; n < ne; n += 8) a += X[n * ld + d];
  red[ty][tx] = a;
  __syncthreads();
  if (ty == 0 && d < D) {
    double t = 0.0;
    for (int j = 0; j < 8; ++j) t += red[j][tx];
    partial[(int64_t)blockIdx.y * D + d] = t;
  }
}

// out[i] = scale * sum_s partial[s][i]   (fixed order)
__global__ void pca_reduce_kernel(const double* __restrict__ partial, int nsplit, int64_t len, double scale, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  double t = 0.0;
  for (int s = 0; s < nsplit; ++s) t += partial[(int64_t)s * len + i];
  out[i] = t * scale;
}

__global__ void __launch_bounds__(256) pca_centre_kernel(double* __restrict__ X, int64_t ld, int64_t n_obs, int D, const double* __restrict__ mean) {
  const int64_t total = n_obs * D;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = e / D;
    const int d = (int)(e - n * D);
    X[n * ld + d] -= mean[d];
  }
}

// partial[split][a][b] = sum over the split's observations of X[n][a] X[n][b] for the 64 x 64 tile (blockIdx.x, blockIdx.y);
// 256 threads, 4 x 4 outputs each, 16 observations staged per step.
constexpr int CT = 64, CK = 16;
__global__ void __launch_bounds__(256) pca_syrk_kernel(const double* __restrict__ X, int64_t ld, int64_t n_obs, int D,
                                                       double* __restrict__ partial, int64_t per_split) {
  __shared__ double As[CK][CT + 1];
  __shared__ double Bs[CK][CT + 1];
  const int a0 = blockIdx.x * CT, b0 = blockIdx.y * CT;
  if (b0 > a0) return;   // lower tiles only; the mirror is filled by pca_symmetrise_kernel
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t nb = (int64_t)blockIdx.z * per_split, ne = min(n_obs, nb + per_split);
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int64_t nc = nb; nc < ne; nc += CK) {
    for (int e = threadIdx.x; e < CK * CT; e += 256) {
      const int o = e / CT, c = e % CT;
      const int64_t n = nc + o;
      As[o][c] = (n < ne && a0 + c < D) ? X[n * ld + a0 + c] : 0.0;
      Bs[o][c] = (n < ne && b0 + c < D) ? X[n * ld + b0 + c] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int o = 0; o < CK; ++o) {
      double av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[o][4 * ty + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[o][4 * tx + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  double* out = partial + (int64_t)blockIdx.z * D * D;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a = a0 + 4 * ty + i;
    if (a >= D) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int b = b0 + 4 * tx + j;
      if (b < D) out[(int64_t)a * D + b] = acc[i][j];
    }
  }
}

// C[b][a] = C[a][b] for b > a in tiles above the diagonal (the SYRK kernel computed the lower tiles)
__global__ void pca_symmetrise_kernel(double* __restrict__ C, int D) {
  const int a = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.x * blockDim.x + threadIdx.x;
  if (a < D && b < D && (b / CT) > (a / CT)) C[(int64_t)a * D + b] = C[(int64_t)b * D + a];
}

// Cyclic Jacobi eigenvalue iteration on the symmetric D x D matrix A (global memory, L2 resident), V accumulates the
// rotations (V = I on entry): on exit diag(A) = eigenvalues, columns of V = eigenvectors.  One CTA of 1024 threads.
// Round-robin ordering (circle method) over m = D rounded up to even: round r pairs (m - 1, r) and
// ((r + k) mod (m - 1), (r - k) mod (m - 1)), k = 1 .. m/2 - 1 -- disjoint pairs, so the m/2 rotations of a round are
// applied together: rows, then columns, then V.  A sweep = m - 1 rounds; stops when a sweep rotates nothing.
constexpr int JT = 1024;
__global__ void __launch_bounds__(JT, 1) pca_jacobi_kernel(double* __restrict__ A, double* __restrict__ V, int D, int max_sweeps,
                                                         int* __restrict__ sweeps_out) {
  extern __shared__ double jsm[];
  const int m = (D + 1) & ~1, half = m / 2;
  double* cs = jsm;                                   // half x 2 : (c, s) of the round's rotations
  int* pq = reinterpret_cast<int*>(jsm + 2 * half);   // half x 2 : (p, q), p < q; p = -1: skip
  __shared__ int rotated;
  const int tid = threadIdx.x;
  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    if (tid == 0) rotated = 0;
    __syncthreads();
    for (int r = 0; r < m - 1; ++r) {
      // (1) rotation angles
      for (int k = tid; k < half; k += JT) {
        int i = (k == 0) ? m - 1 : (r + k) % (m - 1);
        int j = (k == 0) ? r : (r - k + (m - 1)) % (m - 1);
        int p = min(i, j), q = max(i, j);
        double c = 1.0, s = 0.0;
        if (q >= D) {
          p = -1;   // the dummy index of an odd D sits out
        } else {
          const double app = A[(int64_t)p * D + p], aqq = A[(int64_t)q * D + q], apq = A[(int64_t)p * D + q];
          // negligible against the diagonal (relative criterion; an exactly zero diagonal rotates whenever apq != 0)
          if (fabs(apq) <= 1.1102230246251565e-16 * sqrt(fabs(app) * fabs(aqq))) {
            p = -1;
          } else {
            const double tau = (aqq - app) / (2.0 * apq);
            const double t = ((tau >= 0.0) ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + t * t);
            s = t * c;
            rotated = 1;
          }
        }
        cs[2 * k] = c; cs[2 * k + 1] = s;
        pq[2 * k] = p; pq[2 * k + 1] = q;
      }
      __syncthreads();
      // (2) rows: A <- J^T A
      for (int e = tid; e < half * D; e += JT) {
        const int k = e / D, col = e - k * D;
        const int p = pq[2 * k], q = pq[2 * k + 1];
        if (p < 0) continue;
        const double c = cs[2 * k], s = cs[2 * k + 1];
        const double x = A[(int64_t)p * D + col], y = A[(int64_t)q * D + col];
        A[(int64_t)p * D + col] = c * x - s * y;
        A[(int64_t)q * D + col] = s * x + c * y;
      }
      __syncthreads();
      // (3) columns: A <- A J, V <- V J
      for (int e = tid; e < half * D; e += JT) {
        const int row = e / half, k = e - row * half;
        const int p = pq[2 * k], q = pq[2 * k + 1];
        if (p < 0) continue;
        const double c = cs[2 * k], s = cs[2 * k + 1];
        double x = A[(int64_t)row * D + p], y = A[(int64_t)row * D + q];
        A[(int64_t)row * D + p] = c * x - s * y;
        A[(int64_t)row * D + q] = s * x + c * y;
        x = V[(int64_t)row * D + p]; y = V[(int64_t)row * D + q];
        V[(int64_t)row * D + p] = c * x - s * y;
        V[(int64_t)row * D + q] = s * x + c * y;
      }
      __syncthreads();
      // (4) the rotated entries are zero by construction
      for (int k = tid; k < half; k += JT) {
        const int p = pq[2 * k], q = pq[2 * k + 1];
        if (p >= 0) { A[(int64_t)p * D + q] = 0.0; A[(int64_t)q * D + p] = 0.0; }
      }
      __syncthreads();
    }
    if (!rotated) break;
    __syncthreads();
  }
  if (tid == 0) *sweeps_out = sweep;
}

__global__ void pca_identity_kernel(double* __restrict__ V, int D) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < (int64_t)D * D) V[e] = ((e / D) == (e % D)) ? 1.0 : 0.0;
}

// out[n][k] = sum_j X[n][j] E[j][k] (+ bias[k]):  X is n x J (ld ldx), E is J x K row-major.  32 x 32 output tiles, J in
// chunks of 32 through shared memory.
__global__ void __launch_bounds__(256) pca_project_kernel(const double* __restrict__ X, int64_t ldx, int64_t n_obs, int J,
                                                          const double* __restrict__ E, int K, const double* __restrict__ bias,
                                                          double* __restrict__ out, int64_t ldo) {
  __shared__ double Xs[32][33];
  __shared__ double Es[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t n0 = (int64_t)blockIdx.x * 32;
  const int k0 = blockIdx.y * 32;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};   // rows ty, ty + 8, ty + 16, ty + 24; column tx
  for (int j0 = 0; j0 < J; j0 += 32) {
    for (int r = ty; r < 32; r += 8) {
      const int64_t n = n0 + r;
      Xs[r][tx] = (n < n_obs && j0 + tx < J) ? X[n * ldx + j0 + tx] : 0.0;
      Es[r][tx] = (j0 + r < J && k0 + tx < K) ? E[(int64_t)(j0 + r) * K + k0 + tx] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < 32; ++j) {
      const double ev = Es[j][tx];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fma(Xs[ty + 8 * i][j], ev, acc[i]);
    }
    __syncthreads();
  }
  if (k0 + tx < K) {
    const double b = bias ? bias[k0 + tx] : 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t n = n0 + ty + 8 * i;
      if (n < n_obs) out[n * ldo + k0 + tx] = acc[i] + b;
    }
  }
}

}  // namespace
}  // namespace gplvm

using namespace gplvm;

extern "C" int gplvm_pca(const double* X, int64_t N, int64_t D, int64_t d_keep, double* mean_out, double* cov_out, double* eigvals_out,
                         double* eigvecs_out, double* final_out, double* restored_out, int* sweeps_out) {
  if (!X) return fail(GPLVM_ERR_ARG, "gplvm_pca: null input");
  if (N < 2 || D < 1) return fail(GPLVM_ERR_ARG, "gplvm_pca: need N >= 2 observations (covariance divides by N - 1) and D >= 1");
  if (D > 4096) return fail(GPLVM_ERR_ARG, "gplvm_pca: D > 4096 is not supported by the one-CTA eigensolver");
  if (d_keep < 1) return fail(GPLVM_ERR_ARG, "gplvm_pca: desired dimension must be positive");
  const int64_t d = std::min<int64_t>(d_keep, D);   // PCA.hs:49: desiredDimensions >= y keeps every eigenvector
  GP_TRY(ensure_context());
  Context& c = ctx();
  std::lock_guard<std::recursive_mutex> lk(c.mu);
  CUDA_TRY(cudaSetDevice(c.shards[0].device));
  cudaStream_t st = c.shards[0].stream;
  const int Di = (int)D;
  Buf dX, dpart, dmean, dC, dA, dV, dE, dEt, dout, dsw;
  CUDA_TRY(cudaMalloc(&dX.p, sizeof(double) * (size_t)N * D));
  // the observations are rows already (PCA.hs:42 meanCovS works on rows): plain chunked upload
  {
    const int64_t chunk = std::max<int64_t>(1, ((int64_t)64 << 20) / D);
    for (int64_t n0 = 0; n0 < N; n0 += chunk) {
      const int64_t cn = std::min(chunk, N - n0);
      CUDA_TRY(cudaMemcpyAsync(dX.as<double>() + n0 * D, X + n0 * D, sizeof(double) * (size_t)cn * D, cudaMemcpyHostToDevice, st));
    }
  }
  const int sms = 148;
  const int ft = (int)ceil_div(D, 32);
  int nsplit = (int)std::min<int64_t>(std::max<int64_t>(1, 4 * sms / ft), ceil_div(N, 256));
  int64_t per = ceil_div(N, nsplit);
  nsplit = (int)ceil_div(N, per);
  const int tiles = (int)ceil_div(D, CT);
  int ksplit = (int)std::min<int64_t>(std::max<int64_t>(1, 4 * sms / std::max(1, tiles * (tiles + 1) / 2)), ceil_div(N, 1024));
  int64_t kper = round_up(ceil_div(N, ksplit), CK);
  ksplit = (int)ceil_div(N, kper);
  CUDA_TRY(cudaMalloc(&dpart.p, sizeof(double) * std::max<size_t>((size_t)nsplit * D, (size_t)ksplit * D * D)));
  CUDA_TRY(cudaMalloc(&dmean.p, sizeof(double) * D));
  CUDA_TRY(cudaMalloc(&dC.p, sizeof(double) * (size_t)D * D));
  CUDA_TRY(cudaMalloc(&dA.p, sizeof(double) * (size_t)D * D));
  CUDA_TRY(cudaMalloc(&dV.p, sizeof(double) * (size_t)D * D));
  CUDA_TRY(cudaMalloc(&dsw.p, sizeof(int)));
  // mean over the observations (Util-free: hmatrix meanCov), fixed-order sums
  GP_LAUNCH(pca_colsum_kernel, dim3(ft, nsplit), 256, 0, st, dX.as<double>(), D, N, Di, dpart.as<double>(), per);
  GP_LAUNCH(pca_reduce_kernel, (unsigned)ceil_div(D, 256), 256, 0, st, dpart.as<double>(), nsplit, D, 1.0 / (double)N, dmean.as<double>());
  GP_LAUNCH(pca_centre_kernel, sms * 8, 256, 0, st, dX.as<double>(), D, N, Di, dmean.as<double>());
  // covariance = Xc^T Xc / (N - 1)
  CUDA_TRY(cudaMemsetAsync(dpart.p, 0, sizeof(double) * (size_t)ksplit * D * D, st));
  GP_LAUNCH(pca_syrk_kernel, dim3(tiles, tiles, ksplit), 256, 0, st, dX.as<double>(), D, N, Di, dpart.as<double>(), kper);
  GP_LAUNCH(pca_reduce_kernel, (unsigned)ceil_div(D * D, 256), 256, 0, st, dpart.as<double>(), ksplit, D * D, 1.0 / (double)(N - 1),
            dC.as<double>());
  GP_LAUNCH(pca_symmetrise_kernel, dim3((unsigned)ceil_div(D, 32), (unsigned)ceil_div(D, 8)), dim3(32, 8), 0, st, dC.as<double>(), Di);
  // eigSH
  CUDA_TRY(cudaMemcpyAsync(dA.p, dC.p, sizeof(double) * (size_t)D * D, cudaMemcpyDeviceToDevice, st));
  GP_LAUNCH(pca_identity_kernel, (unsigned)ceil_div(D * D, 256), 256, 0, st, dV.as<double>(), Di);
  const int half = (Di + 1) / 2;
  const size_t jsmem = sizeof(double) * 2 * half + sizeof(int) * 2 * half;
  GP_LAUNCH(pca_jacobi_kernel, 1, JT, jsmem, st, dA.as<double>(), dV.as<double>(), Di, 60, dsw.as<int>());
  std::vector<double> hA((size_t)D * D), hV((size_t)D * D), hmean(D);
  int sweeps = 0;
  CUDA_TRY(cudaMemcpyAsync(hA.data(), dA.p, sizeof(double) * hA.size(), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(hV.data(), dV.p, sizeof(double) * hV.size(), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(hmean.data(), dmean.p, sizeof(double) * D, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(&sweeps, dsw.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  if (cov_out) CUDA_TRY(cudaMemcpyAsync(cov_out, dC.p, sizeof(double) * (size_t)D * D, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  if (sweeps_out) *sweeps_out = sweeps;
  if (sweeps >= 60) return fail(GPLVM_ERR_NUMERIC, "gplvm_pca: the Jacobi eigensolver did not converge in 60 sweeps");
  // order: eigenvalues descending (hmatrix eigSH); ties keep the lower index first.  Sign: largest-magnitude entry positive.
  std::vector<int> order(D);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return hA[(size_t)a * D + a] > hA[(size_t)b * D + b]; });
  std::vector<double> E((size_t)D * d), Et((size_t)d * D);
  for (int64_t k = 0; k < D; ++k) {
    const int src = order[k];
    if (eigvals_out) eigvals_out[k] = hA[(size_t)src * D + src];
    if (k >= d) continue;
    int64_t big = 0;
    for (int64_t j = 1; j < D; ++j)
      if (std::fabs(hV[(size_t)j * D + src]) > std::fabs(hV[(size_t)big * D + src])) big = j;
    const double sgn = (hV[(size_t)big * D + src] < 0.0) ? -1.0 : 1.0;
    for (int64_t j = 0; j < D; ++j) {
      const double v = sgn * hV[(size_t)j * D + src];
      E[(size_t)j * d + k] = v;
      Et[(size_t)k * D + j] = v;
    }
  }
  if (mean_out) memcpy(mean_out, hmean.data(), sizeof(double) * D);
  if (eigvecs_out) memcpy(eigvecs_out, E.data(), sizeof(double) * E.size());
  if (final_out || restored_out) {
    CUDA_TRY(cudaMalloc(&dE.p, sizeof(double) * E.size()));
    CUDA_TRY(cudaMalloc(&dEt.p, sizeof(double) * Et.size()));
    CUDA_TRY(cudaMemcpyAsync(dE.p, E.data(), sizeof(double) * E.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dEt.p, Et.data(), sizeof(double) * Et.size(), cudaMemcpyHostToDevice, st));
    // final (N x d) and restored (N x D) in chunks of observations through one staging buffer
    const int64_t chunk = std::max<int64_t>(1024, ((int64_t)32 << 20) / D);
    Buf dfin;
    CUDA_TRY(cudaMalloc(&dfin.p, sizeof(double) * (size_t)std::min(chunk, N) * d));
    if (restored_out) CUDA_TRY(cudaMalloc(&dout.p, sizeof(double) * (size_t)std::min(chunk, N) * D));
    for (int64_t n0 = 0; n0 < N; n0 += chunk) {
      const int64_t cn = std::min(chunk, N - n0);
      GP_LAUNCH(pca_project_kernel, dim3((unsigned)ceil_div(cn, 32), (unsigned)ceil_div(d, 32)), 256, 0, st, dX.as<double>() + n0 * D, D, cn, Di,
                dE.as<double>(), (int)d, (const double*)nullptr, dfin.as<double>(), d);
      if (final_out)
        CUDA_TRY(cudaMemcpyAsync(final_out + n0 * d, dfin.p, sizeof(double) * (size_t)cn * d, cudaMemcpyDeviceToHost, st));
      if (restored_out) {
        GP_LAUNCH(pca_project_kernel, dim3((unsigned)ceil_div(cn, 32), (unsigned)ceil_div(D, 32)), 256, 0, st, dfin.as<double>(), d, cn, (int)d,
                  dEt.as<double>(), Di, dmean.as<double>(), dout.as<double>(), D);
        CUDA_TRY(cudaMemcpyAsync(restored_out + n0 * D, dout.p, sizeof(double) * (size_t)cn * D, cudaMemcpyDeviceToHost, st));
      }
      CUDA_TRY(cudaStreamSynchronize(st));
    }
  }
  CUDA_TRY(cudaGetLastError());
  return GPLVM_OK;
}
