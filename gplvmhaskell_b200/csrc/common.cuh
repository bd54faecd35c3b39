// common.cuh -- shared host/device plumbing of libgplvm_b200 (error reporting, device shards, NCCL
// loaded at run time, small block-level linear algebra used by the replicated parameter updates).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/gplvm_b200.h"

namespace gplvm {

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

#define CUDA_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return ::gplvm::fail(GPLVM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                           __FILE__, __LINE__);                                                \
  } while (0)

#define GP_TRY(expr)            \
  do {                          \
    int _rc = (expr);           \
    if (_rc != GPLVM_OK) return _rc; \
  } while (0)

extern std::atomic<int64_t> g_kernel_launches;

// every kernel launch of the library goes through this macro so that bench.py can report launches
#define GP_LAUNCH(kernel, grid, block, smem, stream, ...)                 \
  do {                                                                    \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);           \
    ::gplvm::g_kernel_launches.fetch_add(1, std::memory_order_relaxed);   \
  } while (0)

// ------------------------------------------------------------------------------------------------
// devices / shards / collectives
// ------------------------------------------------------------------------------------------------
struct Shard {
  int device = 0;
  cudaStream_t stream = nullptr;
  void* comm = nullptr;  // ncclComm_t when world > 1
};

struct Context {
  std::recursive_mutex mu;
  std::vector<Shard> shards;  // GPUs driven by this process
  int requested = 1;          // gplvm_set_devices
  bool multiproc = false;     // gplvm_dist_init was called (one shard, rank = process rank)
  int rank = 0;               // rank of shards[0] in the world
  int world = 1;              // total number of shards over all processes
  bool ready = false;
};

Context& ctx();
int ensure_context();  // create streams (and the single-process NCCL clique) lazily
int shutdown_context();

// In-place sum over every shard in the world of `count` doubles; bufs[i] lives on shards[i].device.
// world == 1: no-op.  Enqueued on the shard streams (no host sync).
int allreduce_sum(const std::vector<double*>& bufs, size_t count);

// recv_host[world][bytes] <- every process's `send` (multi-process mode only; goes through the NCCL communicator)
int allgather_bytes(const void* send, void* recv_host, size_t bytes);
int nccl_unique_id(void* id128);
int nccl_init_rank(int rank, int world, int device, const void* id128);

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 1 / sqrt(x) and 1 / x for a positive normal x: hardware seed + 2 Newton steps (relative error ~ 1 ulp)
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
__device__ __forceinline__ double rcp_nr(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
  }
  return y;
}

// In-place inverse of a symmetric positive definite 16 x 16 matrix held column-per-lane by 16 consecutive lanes
// (lane c = lane & 15 owns m[a] = M[a][c], full symmetric storage), by 16 symmetric sweeps:
//   step k (pivot p = M[k][k], d = 1/p):  B[i][l] = M[i][l] - M[i][k] M[k][l] d,  B[i][k] = M[i][k] d,  B[k][k] = -d
// after all 16 steps the registers hold -M^-1; the function returns +M^-1 in m.  Lane c reads M[k][c] from its own
// column by symmetry, only column k travels (width-16 shuffles).  Lane k keeps column k unscaled and remembers d
// (applied at the end), so the update formula is identical in every lane.  pprod = product of the pivots = det M;
// ok = every pivot positive.  All 32 lanes of the warp must call it (two independent matrices per warp).
__device__ __forceinline__ void sweep_inverse16(double (&m)[16], int c, bool& ok, double& pprod) {
  double scale = 1.0;
  ok = true;
  pprod = 1.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const double pk = __shfl_sync(0xffffffffu, m[k], k, 16);
    ok = ok && (pk > 0.0);
    pprod *= pk;
    const double d = rcp_nr(pk);
    const double r = (c == k) ? 0.0 : m[k] * d;
#pragma unroll
    for (int a = 0; a < 16; ++a) {
      if (a == k) continue;
      const double ca = __shfl_sync(0xffffffffu, m[a], k, 16);
      m[a] = fma(-ca, r, m[a]);
    }
    m[k] = (c == k) ? -1.0 : r;
    if (c == k) scale = d;
  }
#pragma unroll
  for (int a = 0; a < 16; ++a) m[a] *= -scale;
}

// Sum over the whole block (any blockDim multiple of 32, <= 1024); result valid in every thread.
// `red` is a shared array of >= 33 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = (lane < nw) ? red[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}
__device__ __forceinline__ double block_max(double v, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = (lane < nw) ? red[lane] : 0.0;
    t = warp_max(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// In-place lower Cholesky of the n x n SPD matrix `a` (row-major, leading dimension ld) in shared
// memory by the whole block.  The strict upper triangle is left untouched.  Returns false (in every
// thread) when a pivot is not positive.
__device__ inline bool block_cholesky(double* a, int n, int ld) {
  const int tid = threadIdx.x, nt = blockDim.x;
  bool ok = true;
  for (int k = 0; k < n; ++k) {
    __syncthreads();
    const double akk = a[k * ld + k];
    if (!(akk > 0.0)) ok = false;
    const double lkk = sqrt(akk);
    __syncthreads();
    for (int i = k + tid; i < n; i += nt) a[i * ld + k] = (i == k) ? lkk : a[i * ld + k] / lkk;
    __syncthreads();
    const int r = n - k - 1;
    for (int idx = tid; idx < r * r; idx += nt) {
      const int i = k + 1 + idx / r, j = k + 1 + idx % r;
      if (j <= i) a[i * ld + j] -= a[i * ld + k] * a[j * ld + k];
    }
  }
  __syncthreads();
  return ok;
}

// inv (n x n, ld) = (L L^T)^-1 from the lower factor L (ld); one thread per column.
__device__ inline void block_chol_inverse(const double* L, double* inv, int n, int ld) {
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    // forward: L z = e_j   (z_i = 0 for i < j)
    for (int i = 0; i < n; ++i) {
      double s = (i == j) ? 1.0 : 0.0;
      for (int k = j; k < i; ++k) s -= L[i * ld + k] * inv[k * ld + j];
      inv[i * ld + j] = (i < j) ? 0.0 : s / L[i * ld + i];
    }
    // backward: L^T x = z
    for (int i = n - 1; i >= 0; --i) {
      double s = inv[i * ld + j];
      for (int k = i + 1; k < n; ++k) s -= L[k * ld + i] * inv[k * ld + j];
      inv[i * ld + j] = s / L[i * ld + i];
    }
  }
  __syncthreads();
}

// x <- (L L^T)^-1 x for one right-hand side held by ONE thread (x has n entries, stride 1).
__device__ __forceinline__ void chol_solve_vec(const double* L, int n, int ld, double* x) {
  for (int i = 0; i < n; ++i) {
    double s = x[i];
    for (int k = 0; k < i; ++k) s -= L[i * ld + k] * x[k];
    x[i] = s / L[i * ld + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = x[i];
    for (int k = i + 1; k < n; ++k) s -= L[k * ld + i] * x[k];
    x[i] = s / L[i * ld + i];
  }
}

// Philox-4x32-10 counter based generator (Salmon et al. 2011), used for the synthetic workloads.
struct Philox {
  static __host__ __device__ __forceinline__ void round(uint32_t (&c)[4], uint32_t (&k)[2]) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    c[0] = hi1 ^ c[1] ^ k[0];
    c[1] = lo1;
    c[2] = hi0 ^ c[3] ^ k[1];
    c[3] = lo0;
    k[0] += 0x9E3779B9u;
    k[1] += 0xBB67AE85u;
  }
  // 4 x 32 random bits for (seed, a, b)
  static __host__ __device__ __forceinline__ void gen(uint64_t seed, uint64_t a, uint64_t b, uint32_t (&out)[4]) {
    uint32_t c[4] = {(uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32)};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
    for (int i = 0; i < 10; ++i) round(c, k);
    for (int i = 0; i < 4; ++i) out[i] = c[i];
  }
  // one standard normal and one uniform in [0,1) for (seed, a, b)
  static __host__ __device__ __forceinline__ void normal_uniform(uint64_t seed, uint64_t a, uint64_t b, double& z,
                                                                 double& u) {
    uint32_t r[4];
    gen(seed, a, b, r);
    const double u1 = ((double)r[0] + 0.5) * (1.0 / 4294967296.0);
    const double u2 = ((double)r[1] + 0.5) * (1.0 / 4294967296.0);
    z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    u = ((double)r[2] + (double)r[3] * (1.0 / 4294967296.0)) * (1.0 / 4294967296.0);
  }
};

#endif  // __CUDACC__

}  // namespace gplvm
