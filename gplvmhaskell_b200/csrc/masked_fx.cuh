// masked_fx.cuh -- fixed-point conventions of the INT8 complement-sum kernel (masked_utc.cu, tcgen05 kind::i8).
//
//   miss (D x 152) = Miss^T (D x n) . [vech A | ez] (n x 152),   Miss[i][d] = 1 when observation i misses feature d
// (reference: the per-feature sums over the observations that miss / have the feature, emStepsMissed.stepTwo,
// src/GPLVM/PPCA.hs:119-147; complement form).
//
// One factor is exactly 0 / 1, so only the other needs splitting: every [vech A | ez] entry v is rounded ONCE to the
// fixed-point grid of its column class,  x = rint(v 2^(51 - E)),  |v| < 2^E  (E from the largest |entry| of the class --
// vech A or ez -- which the solver kernel tracks with an integer max), and  u = x + 2^51  in [0, 2^52] is cut into its
// 7 bytes.  Byte planes are u8 B-operands, the mask bits expand to s8 A-operands, the products accumulate exactly in
// s32, and a constant-1 column counts the offsets:  sum_i x_i = sum_s 256^s S_s - 2^51 cnt.  All sums are integers =>
// independent of the summation order and of the split of the observations over CTAs: bitwise reproducible.  The one
// rounding per entry is <= 2^-52 of the class maximum, i.e. no worse than the rounding of a single FP64 addition of that
// entry into a running sum; the accumulated sum itself is exact and is rounded to FP64 once at the end.
// Precision guard: the cutting threads count (sampled) the entries more than 2^20 below their class maximum; above 1 %
// the finalize kernel hands the step to the FP64 product of masked_fused.cu on the device.
// oracle/fixedpoint.py restates the arithmetic on the CPU (tests only).
#pragma once
#include "ppca.cuh"

namespace gplvm {
namespace fx {

constexpr int NSL = 7;              // byte planes of the 52-bit fixed-point value
constexpr int FXP = 51;             // |x| < 2^FXP
constexpr int IM_CHUNK = 128;       // observations per shared-memory tile
constexpr int IM_EAS = 152;
constexpr int IM_NV = 136;

// biased exponent of the class maximum, clamped so that both 2^(51 - E) and 2^(E - 51) are normal numbers
__device__ __forceinline__ int class_exponent(uint32_t hmax) {
  int ex = (int)((hmax >> 20) & 0x7ffu);
  return min(max(ex, 60), 2040);
}
__device__ __forceinline__ double scale_up(uint32_t hmax) {      // 2^(FXP - E),  E = ex - 1022
  return __hiloint2double((1023 + FXP + 1022 - class_exponent(hmax)) << 20, 0);
}
__device__ __forceinline__ double scale_down(uint32_t hmax) {    // 2^(E - FXP)
  return __hiloint2double((1023 - FXP - 1022 + class_exponent(hmax)) << 20, 0);
}

// eamax block (256 B, device): [0] / [1] high words of max |vech A| / max |ez| (solver), then two 64-bit counters of
// entries that the fixed-point grid resolves poorly, then S_COUNT doubles used as the `scal` block of the FP64 fallback
// kernels (slot S_DONE = 1.0: skip, the INT8 result stands).
constexpr int IM_GUARD_BITS = 20;
__host__ __device__ inline unsigned long long* guard_counts(uint32_t* eamax) { return reinterpret_cast<unsigned long long*>(eamax + 2); }
__host__ __device__ inline double* guard_scal(uint32_t* eamax) { return reinterpret_cast<double*>(eamax + 16); }
// more than 1 % of the entries of a class more than 2^20 below its maximum (an outlier sets the scale): the grid would
// leave typical entries fewer than 32 bits, so this step's sums are redone in FP64 (masked_miss_dmma_kernel)
__device__ __forceinline__ bool guard_fallback(const uint32_t* eamax, int64_t n_obs) {
  const unsigned long long* cnt = reinterpret_cast<const unsigned long long*>(eamax + 2);
  // the counters sample one observation in four
  return cnt[0] * 400ull > (unsigned long long)n_obs * IM_NV || cnt[1] * 400ull > (unsigned long long)n_obs * (IM_EAS - IM_NV);
}

// exact integer T (|T| < 2^(51 + 24)) -> FP64 with one rounding: the upper part is exact, the low 24 bits are fused in
__device__ __forceinline__ double int128_to_double(__int128 T) {
  const long long thi = (long long)(T >> 24), tlo = (long long)(T & 0xffffff);
  return fma((double)thi, 16777216.0, (double)tlo);
}

}  // namespace fx
}  // namespace gplvm
