// masked_sweeps.cuh -- the register-resident symmetric-sweep inverse shared by the per-observation solvers
// (masked_fused.cu: Gram matrices from FP64 DMMA; masked_gram_tc.cu: Gram matrices from tcgen05 fixed-point products).
// Reference: invMP of emStepsMissed.stepOne, src/GPLVM/PPCA.hs:106-117.
#pragma once
#include "common.cuh"

namespace gplvm {

// A group of LPR lanes owns one observation; lane j of the group holds columns j, j + LPR, ... of the symmetric
// 16 x 16 matrix: m[q][a] = M[a][j + LPR q].  On return m holds the same columns of M^-1.
// Step k (pivot p = M[k][k], d = 1/p):  B[i][l] = M[i][l] - M[i][k] M[k][l] d,  B[i][k] = M[i][k] d,  B[k][k] = -d;
// after k = 0..15 the matrix is -M^-1.  A lane reads M[k][c] from its own columns by symmetry; only column k travels
// (width-LPR shuffles from its owner).  The owner keeps column k UNSCALED and remembers d in `scale` (applied once at
// the end), which keeps the update formula identical in every lane (r = 0 makes it a no-op for the owner's column).
// ok: every pivot (Schur complement) was positive;  pprod (LL only): product of the pivots = det M.
template <int LPR, bool LL>
__device__ __forceinline__ void sweep_invert16(double (&m)[16 / LPR][16], int j, bool& ok, double& pprod) {
  constexpr int CPL = 16 / LPR;
  constexpr unsigned FULL = 0xffffffffu;
  double scale[CPL];
#pragma unroll
  for (int q = 0; q < CPL; ++q) scale[q] = 1.0;
  // The pivot of step k + 1 is ready as soon as row k + 1 has been updated in step k: that row goes first, and the
  // reciprocal chain of the next pivot (shuffle, rcp.approx, two Newton steps) overlaps the other 14 row updates.
  double pk = __shfl_sync(FULL, m[0][0], 0, LPR);
  ok = ok && (pk > 0.0);
  if (LL) pprod *= pk;
  double d = rcp_nr(pk);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int qk = k / LPR, jk = k % LPR;
    double rr[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) rr[q] = (q == qk && j == jk) ? 0.0 : m[q][k] * d;
    double dn = 0.0;
    if (k + 1 < 16) {
      const double ca = __shfl_sync(FULL, m[qk][k + 1], jk, LPR);
#pragma unroll
      for (int q = 0; q < CPL; ++q) m[q][k + 1] = fma(-ca, rr[q], m[q][k + 1]);
      pk = __shfl_sync(FULL, m[(k + 1) / LPR][k + 1], (k + 1) % LPR, LPR);
      ok = ok && (pk > 0.0);
      if (LL) pprod *= pk;
      dn = rcp_nr(pk);
    }
#pragma unroll
    for (int a = 0; a < 16; ++a) {
      if (a == k || a == k + 1) continue;
      const double ca = __shfl_sync(FULL, m[qk][a], jk, LPR);
#pragma unroll
      for (int q = 0; q < CPL; ++q) m[q][a] = fma(-ca, rr[q], m[q][a]);
    }
#pragma unroll
    for (int q = 0; q < CPL; ++q) m[q][k] = (q == qk && j == jk) ? -1.0 : rr[q];
    if (j == jk) scale[qk] = d;
    d = dn;
  }
#pragma unroll
  for (int q = 0; q < CPL; ++q)
#pragma unroll
    for (int a = 0; a < 16; ++a) m[q][a] *= -scale[q];      // now column j + LPR q of M^-1
}

}  // namespace gplvm
