// tma.cuh -- sm_100a building blocks shared by the fused kernels: mbarrier + TMA (cp.async.bulk[.tensor]) wrappers,
// the FP64 tensor-core MMA (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4) and the tensor-map encoder lookup.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace gplvm {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: every try_wait suspends the thread for a hardware time slice, so the limit below is seconds of wall
// clock -- far beyond any healthy TMA latency.  A copy that never completes (faulted tensor map, bad address) ends the
// kernel with a trap instead of hanging the GPU; the host sees the launch failure as GPLVM_ERR_CUDA at the next sync.
constexpr uint32_t MBAR_SPIN_LIMIT = 1u << 26;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < MBAR_SPIN_LIMIT; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
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


// 1-D bulk copy global -> shared (contiguous bytes; multiple of 16, 16-byte aligned), completes on an mbarrier
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void lds128(uint32_t addr, double& a, double& b) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline int get_encode_fn(EncodeTiledFn* fn) {
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


// 2-D FP64 tensor map over a row-major (rows x cols, leading dimension ld) matrix with 16-double x box_rows boxes,
// SWIZZLE_128B (each box row is one 128-byte line), out-of-bounds rows read as zero.
inline int encode_rows_map(CUtensorMap* map, const double* base, int64_t cols, int64_t rows, int64_t ld, int box_rows) {
  EncodeTiledFn enc = nullptr;
  GP_TRY(get_encode_fn(&enc));
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(double)};
  const cuuint32_t box[2] = {16u, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(GPLVM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return GPLVM_OK;
}

inline int sm_count(int dev) {
  static int cache[64] = {};
  if (!cache[dev & 63]) cudaDeviceGetAttribute(&cache[dev & 63], cudaDevAttrMultiProcessorCount, dev);
  return cache[dev & 63] > 0 ? cache[dev & 63] : 148;
}

}  // namespace gplvm
