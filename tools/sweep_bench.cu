// sweep_bench.cu -- the register-resident 16 x 16 symmetric-sweep inverse of the masked E-step solver in isolation
// (masked_sweeps.cuh), to compare distributions of the matrix over lanes at the residency of the product kernel
// (8 warps per SM, <= 232 registers):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gplvmhaskell_b200/csrc -I include
//                                           -o tools/sweep_bench tools/sweep_bench.cu
//   V0: shipped -- 4 lanes per observation, lane j owns columns j, j+4, j+8, j+12, one pivot per step (16 steps, look-ahead)
//   V1: 2 x 2 block pivots -- lane j owns columns 2j, 2j+1, 8+2j, 9+2j (a pivot pair lives in one lane), 8 steps
// Every variant loads M (SPD) from global memory in its own layout, inverts, and writes vech(M^-1); the host checks against
// a double-precision Gauss-Jordan and reports ns per observation.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "masked_sweeps.cuh"

using namespace gplvm;
constexpr unsigned FULLM = 0xffffffffu;

// ---------------------------------------------------------------------------------------------------------------- V0
__global__ void __launch_bounds__(256, 1) v0_kernel(const double* __restrict__ Min, double* __restrict__ out, int64_t n, int reps) {
  const int lane = threadIdx.x & 31, rw = lane >> 2, j = lane & 3;
  const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
  for (int64_t base = warp * 8; base < n; base += nwarps * 8) {
    const int64_t i = min(base + rw, n - 1);
    double m[4][16];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int a = 0; a < 16; ++a) m[q][a] = Min[i * 256 + a * 16 + j + 4 * q];
    bool ok = true;
    double pprod = 1.0;
    for (int r = 0; r < reps; ++r) sweep_invert16<4, false>(m, j, ok, pprod);   // reps odd: inverse; even: back to M
    if (base + rw < n) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int a = 0; a < 16; ++a) out[i * 256 + a * 16 + j + 4 * q] = ok ? m[q][a] : 0.0;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------- V1
// lane j owns columns col(q) = 2 j + (q & 1) + 8 (q >> 1), q = 0..3; pivot pair P = {2 s, 2 s + 1} lives in lane s & 3,
// registers q0 = 2 (s >> 2), q0 + 1.  m[q][a] = M[a][col(q)].
// Step s: K = [[p11, p12], [p12, p22]] = M[P][P], Kinv = K^-1;  for i, l not in P:
//   B[i][l] = M[i][l] - [M[i][k] M[i][k+1]] Kinv [M[k][l]; M[k+1][l]];  B[i][P] = M[i][P] Kinv;  B[P][P] = -Kinv.
// The owner keeps its two pivot columns UNMIXED (U = M[:, P]; true B[:, P] = U Kinv, a linear map that commutes with
// the later row updates) and applies Kinv once at the end -- the 2 x 2 analogue of the deferred column scale of V0.
__device__ __forceinline__ void sweep_invert16_pair(double (&m)[4][16], int j, bool& ok) {
  double kinv[2][3];   // per owned pivot pair (q0 = 0, 2): i11, i12, i22
#pragma unroll
  for (int h = 0; h < 2; ++h) kinv[h][0] = kinv[h][2] = 1.0, kinv[h][1] = 0.0;
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const int js = s & 3, q0 = 2 * (s >> 2), k = 2 * s;
    // pivot block from its owner
    const double p11 = __shfl_sync(FULLM, m[q0][k], js, 4);
    const double p12 = __shfl_sync(FULLM, m[q0 + 1][k], js, 4);
    const double p22 = __shfl_sync(FULLM, m[q0 + 1][k + 1], js, 4);
    const double det = fma(p11, p22, -p12 * p12);
    ok = ok && (p11 > 0.0) && (det > 0.0);
    const double rd = rcp_nr(det);
    const double i11 = p22 * rd, i12 = -p12 * rd, i22 = p11 * rd;
    // rr (row-pair of the own columns through Kinv); zero for the owner's two pivot columns (they stay unmixed)
    double ra[4], rb[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const bool own = (j == js) && (q == q0 || q == q0 + 1);
      const double u = m[q][k], v = m[q][k + 1];
      ra[q] = own ? 0.0 : fma(u, i11, v * i12);
      rb[q] = own ? 0.0 : fma(u, i12, v * i22);
    }
#pragma unroll
    for (int a = 0; a < 16; ++a) {
      if (a == k || a == k + 1) continue;
      const double ca = __shfl_sync(FULLM, m[q0][a], js, 4);
      const double cb = __shfl_sync(FULLM, m[q0 + 1][a], js, 4);
#pragma unroll
      for (int q = 0; q < 4; ++q) m[q][a] = fma(-ca, ra[q], fma(-cb, rb[q], m[q][a]));
    }
    // rows k, k+1 of every column: B[P][l] = Kinv M[P][l]  (owner's pivot columns: -I, mixed into -Kinv at the end)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const bool own = (j == js) && (q == q0 || q == q0 + 1);
      m[q][k] = own ? ((q == q0) ? -1.0 : 0.0) : ra[q];
      m[q][k + 1] = own ? ((q == q0) ? 0.0 : -1.0) : rb[q];
    }
    if (j == js) {
      kinv[s >> 2][0] = i11;
      kinv[s >> 2][1] = i12;
      kinv[s >> 2][2] = i22;
    }
  }
  // deferred mix of the pivot column pairs, and the sign: M^-1 = -B
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int a = 0; a < 16; ++a) {
      const double u = m[2 * h][a], v = m[2 * h + 1][a];
      m[2 * h][a] = -fma(u, kinv[h][0], v * kinv[h][1]);
      m[2 * h + 1][a] = -fma(u, kinv[h][1], v * kinv[h][2]);
    }
}

__global__ void __launch_bounds__(256, 1) v1_kernel(const double* __restrict__ Min, double* __restrict__ out, int64_t n, int reps) {
  const int lane = threadIdx.x & 31, rw = lane >> 2, j = lane & 3;
  const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
  for (int64_t base = warp * 8; base < n; base += nwarps * 8) {
    const int64_t i = min(base + rw, n - 1);
    double m[4][16];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int a = 0; a < 16; ++a) m[q][a] = Min[i * 256 + a * 16 + 2 * j + (q & 1) + 8 * (q >> 1)];
    bool ok = true;
    for (int r = 0; r < reps; ++r) sweep_invert16_pair(m, j, ok);
    if (base + rw < n) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int a = 0; a < 16; ++a) out[i * 256 + a * 16 + 2 * j + (q & 1) + 8 * (q >> 1)] = ok ? m[q][a] : 0.0;
    }
  }
}

static void host_inverse(const double* M, double* inv) {
  double a[16][32];
  for (int r = 0; r < 16; ++r)
    for (int c = 0; c < 16; ++c) a[r][c] = M[r * 16 + c], a[r][16 + c] = (r == c);
  for (int k = 0; k < 16; ++k) {
    const double p = a[k][k];
    for (int c = 0; c < 32; ++c) a[k][c] /= p;
    for (int r = 0; r < 16; ++r)
      if (r != k) {
        const double f = a[r][k];
        for (int c = 0; c < 32; ++c) a[r][c] -= f * a[k][c];
      }
  }
  for (int r = 0; r < 16; ++r)
    for (int c = 0; c < 16; ++c) inv[r * 16 + c] = a[r][16 + c];
}

int main(int argc, char** argv) {
  const int64_t n = (argc > 1) ? atoll(argv[1]) : (1 << 20);
  std::vector<double> M((size_t)n * 256);
  srand(1);
  std::vector<double> V(40 * 16);
  for (int64_t i = 0; i < n; ++i) {
    if (i < 64) {   // distinct SPD matrices for the check; the rest repeat them (timing only)
      for (auto& v : V) v = rand() / (double)RAND_MAX - 0.5;
      for (int r = 0; r < 16; ++r)
        for (int c = 0; c < 16; ++c) {
          double s = (r == c) ? 1.0 : 0.0;
          for (int k = 0; k < 40; ++k) s += V[k * 16 + r] * V[k * 16 + c];
          M[i * 256 + r * 16 + c] = s;
        }
    } else {
      for (int e = 0; e < 256; ++e) M[i * 256 + e] = M[(i & 63) * 256 + e];
    }
  }
  double *dM, *dO;
  cudaMalloc(&dM, sizeof(double) * n * 256);
  cudaMalloc(&dO, sizeof(double) * n * 256);
  cudaMemcpy(dM, M.data(), sizeof(double) * n * 256, cudaMemcpyHostToDevice);
  std::vector<double> O((size_t)64 * 256), ref(256);
  typedef void (*kern_t)(const double*, double*, int64_t, int);
  struct { const char* name; kern_t k; } variants[] = {{"V0 one pivot per step (shipped)", v0_kernel}, {"V1 2x2 block pivots", v1_kernel}};
  for (auto& v : variants) {
    v.k<<<148, 256>>>(dM, dO, n, 1);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(O.data(), dO, sizeof(double) * 64 * 256, cudaMemcpyDeviceToHost);
    double worst = 0.0;
    for (int i = 0; i < 64; ++i) {
      host_inverse(&M[(size_t)i * 256], ref.data());
      double mx = 0.0, df = 0.0;
      for (int e2 = 0; e2 < 256; ++e2) mx = fmax(mx, fabs(ref[e2])), df = fmax(df, fabs(ref[e2] - O[(size_t)i * 256 + e2]));
      worst = fmax(worst, df / mx);
    }
    // timing: reps = 1 and reps = 9 in the same kernel isolate 8 inversions from the load / store
    float t1 = 0, t9 = 0;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int reps : {1, 9}) {
      v.k<<<148, 256>>>(dM, dO, n, reps);
      cudaEventRecord(a);
      for (int it = 0; it < 3; ++it) v.k<<<148, 256>>>(dM, dO, n, reps);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      (reps == 1 ? t1 : t9) = ms / 3;
    }
    printf("%-34s max rel err %.2e | kernel %.3f ms (1 inversion + load/store), %.3f ms (9) -> %.2f ns per inversion per SM-batch, "
           "%.3f ms per 2^20 inversions  %s\n",
           v.name, worst, t1, t9, (t9 - t1) / 8 * 1e6 / (double)n * 148, (t9 - t1) / 8 * (1048576.0 / n), cudaGetErrorString(e));
  }
  return 0;
}
