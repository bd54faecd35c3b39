// masked_fused.cu -- batched per-observation solver of the NaN-masked E-step for M = 16 (north-star item 2):
// every observation applies its own NaN mask, builds and factors its own 16 x 16 system in REGISTERS with
// warp shuffles, and emits E[z_i] and A_i = E[z_i] E[z_i]^T + sigma^2 M_i^-1.
//
// Reference: emStepsMissed.stepOne, src/GPLVM/PPCA.hs:106-117 (M_i = sigma^2 I + W_i^T W_i, invMP, expXi) and the
// per-observation log-likelihood terms of PPCA.hs:149-159 (through the determinant lemma).
//
// Mapping: one warp per group of 8 observations (details at the kernel).  This is the FP64 DMMA solver: it serves the likelihood
// pass, D = 64 / 128, and any E-step whose tensor-core Gram path (masked_gram_tc.cu, the default at D = 256) has its precision
// gate closed.
//   build    : M_i = C0 - W_miss^T W_miss (complement form: work ~ #missing).  The missing-feature indices are compacted
//              from the packed mask bits; the Gram matrix is a K-compacted mma.sync.m8n8k4.f64 product whose operands
//              are gathered rows of W in shared memory (W is D x 16, pitch 18 doubles), one observation at a time.
//   then one QUAD of lanes per observation, each lane owning 4 columns of the symmetric 16 x 16 matrix in registers:
//   invert   : 16 symmetric sweeps (masked_sweeps.cuh).
//   outputs  : ez = M_i^-1 b, A_i = ez ez^T + sigma^2 M_i^-1, rows [vech A | ez] staged and stored as one coalesced block.
// b_i (and |y_c|^2 for the likelihood) come from the TMA + DMMA pass in dense_dmma.cu (MODE_MASKED_B).
#include "masked_fx.cuh"
#include "masked_sweeps.cuh"
#include "ppca.cuh"
#include "tma.cuh"

namespace gplvm {
namespace {

constexpr int MF = 16;
constexpr int NVF = MF * (MF + 1) / 2;   // 136
constexpr int EASF = NVF + MF;           // 152
// Pitch (doubles) of the W rows in shared memory: spreads the gathered rows over the banks.  (A residue-bucketed gather --
// lane t only gathers features d = t mod 4 from rows of pitch 20, which makes every gather bank-conflict free -- was built
// and measured in round 2: shared-memory load wavefronts -37 %, conflicts -96 %, but +21 % k-steps and +30 % instructions:
// 16.2 ms against 12.6 ms at N = 2^23.  The kernel is bound by issue latency at 8 warps per SM, not by the data pipe.)
constexpr int WP = 18;
constexpr int EAP = 156;                 // pitch (doubles) of the staged [vech A | ez] rows: 8 observations x 4 lanes bank-disjoint

// LL = false: writes EA[n] = [vech A_i | ez_i].   LL = true: accumulates the per-observation likelihood terms
//   cnt log 2pi + (cnt - M) log s2 + log det M_i + (yy_i - b_i . ez_i) / s2      into partials[blockIdx.x].
//
// One warp handles ROWS = 32 / LPR observations at a time.  During the sweeps a group of LPR lanes owns one
// observation: lane j of the group holds columns j, j + LPR, ... of the symmetric 16 x 16 matrix (CPL x 16 registers),
// so one 64-bit shuffle moves a column entry for ROWS observations at once -- the shared-memory/shuffle data pipe is what
// bounds this kernel: 256 column entries travel per observation, i.e. 512 / ROWS shuffle wavefronts.
//   LPR = 8: 4 observations per warp, 32 matrix registers per lane, 2 CTAs per SM;
//   LPR = 4: 8 observations per warp, 64 matrix registers per lane, 1 CTA per SM, half the shuffle wavefronts.
// The Gram matrices are built one observation at a time (own k-loop bound: no padding to the longest list of the warp)
// and staged through shared memory into the sweep layout; [vech A | ez] rows leave through the same staging area as
// coalesced 128-bit stores (the ROWS rows of a warp are contiguous in HBM).
constexpr int SOLVE_WARPS = 8;                  // warps per CTA
constexpr int MS_STRIDE = MF * 17 + 4;          // doubles per staged matrix; +4 spreads the lane groups over the banks

template <int LPR>
struct SolveCfg {
  static constexpr int WPS = WP;                 // pitch of the W rows
  static constexpr int NZ = 1;                   // zero rows appended to W (list padding)
  static constexpr int CPL = MF / LPR;           // matrix columns per lane
  static constexpr int ROWS = 32 / LPR;          // observations per warp
  static constexpr int MINB = (LPR == 8) ? 2 : 1;
  // LPR = 4: the index lists get their own region (a matrix is staged while later lists are still needed);
  // LPR = 8: they alias the staging area (two CTAs must fit in one SM) and the Gram fragments wait in registers
  static constexpr bool LIST_SEP = ROWS > 4;
  static __host__ __device__ constexpr int warp_doubles(int DD) { return ROWS * MS_STRIDE + (LIST_SEP ? ROWS * DD / 4 : 0); }
  static __host__ __device__ constexpr int smem_bytes(int DD) { return ((DD + NZ) * WPS + SOLVE_WARPS * warp_doubles(DD)) * (int)sizeof(double); }
};

template <int DD, bool LL, int LPR>
__global__ void __launch_bounds__(SOLVE_WARPS * 32, SolveCfg<LPR>::MINB)
masked_solve_kernel(const double* __restrict__ bsrc, const double* __restrict__ yy, const uint32_t* __restrict__ mbits, int64_t n_obs,
                    Par p, double* __restrict__ EA, double* __restrict__ partials, uint32_t* __restrict__ eamax,
                    const double* __restrict__ gate) {
  using SC = SolveCfg<LPR>;
  constexpr int CPL = SC::CPL, ROWS = SC::ROWS, WPS = SC::WPS, NZ = SC::NZ;
  constexpr int WPRT = DD / 32;
  static_assert(WPRT <= 8, "mask words per observation");
  static_assert(SC::LIST_SEP || ROWS * MS_STRIDE * 8 >= ROWS * DD * 2, "index lists must fit in the staging area");
  static_assert(MS_STRIDE >= EAP && EAP >= EASF, "[vech A | ez] rows must fit in the staging area");
  static_assert(DD % 4 == 0, "list region size");
  if (!LL && p.scal[S_DONE] != 0.0) return;
  if (!LL && gate != nullptr && *gate == 1.0) return;   // the tensor-core Gram solver (masked_gram_tc.cu) took this step
  extern __shared__ __align__(16) double sm[];
  double* Ws = sm;                               // (DD + NZ) x WPS : rows of W; rows DD .. = 0 (list / bucket padding)
  double* Msm = Ws + (DD + NZ) * WPS;            // per warp: ROWS staged matrices [+ index lists]
  __shared__ double red[40];
  for (int e = threadIdx.x; e < (DD + NZ) * MF; e += blockDim.x) {
    const int d = e >> 4, a = e & 15;
    Ws[d * WPS + a] = (d < DD) ? p.W[e] : 0.0;
  }
  __syncthreads();
  const double s2 = p.scal[S_SIGMA2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;         // MMA fragment coordinates
  const int rw = lane / LPR, j = lane % LPR;     // sweep coordinates: observation within the warp, lane within its group
  constexpr int WARP_DOUBLES = SC::warp_doubles(DD);
  double* Mw = Msm + warp * WARP_DOUBLES;
  // C0 = s2 I + W^T W (kept in p.Minv by the masked prep kernel) at this lane's C-fragment positions (tiles 00, 01, 11):
  // the staged matrix is C0 - Gram
  double c00[2], c01[2], c11[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    c00[q] = p.Minv[g * MF + 2 * t + q];
    c01[q] = p.Minv[g * MF + 8 + 2 * t + q];
    c11[q] = p.Minv[(8 + g) * MF + 8 + 2 * t + q];
  }
  uint16_t* lst = reinterpret_cast<uint16_t*>(SC::LIST_SEP ? Mw + ROWS * MS_STRIDE : Mw);   // ROWS x DD uint16
  constexpr unsigned FULL = 0xffffffffu;
  // The trip count depends on blockIdx only, so the compiler can see that every warp stays converged around the
  // shuffles (a warp whose observations lie past the end recomputes the last one and stores nothing).
  const int64_t cta0 = (int64_t)blockIdx.x * (SOLVE_WARPS * ROWS);
  const int64_t stride = (int64_t)gridDim.x * (SOLVE_WARPS * ROWS);
  const int64_t iters = (n_obs > cta0) ? (n_obs - cta0 + stride - 1) / stride : 0;
  double llacc = 0.0;
  uint32_t mxa = 0u, mxe = 0u;   // high words of the largest |vech A| and |ez| entries written (fixed-point scale, masked_fx.cuh)
  const double log2pi = 1.8378770664093453, logs2 = LL ? log(s2) : 0.0;

  for (int64_t it = 0; it < iters; ++it) {
    const int64_t nb = cta0 + it * stride + warp * ROWS;   // first observation of this warp
    // ---- (a) compact the missing-feature indices of the ROWS observations ----------------------------------------
    uint32_t mwa = 0u, mwb = 0u;   // lane l: word (l & 7) of observation nb + (l >> 3) [+ 4]
    if ((lane & 7) < WPRT) {
      const int64_t na = min(nb + (lane >> 3), n_obs - 1);
      mwa = mbits[na * WPRT + (lane & 7)];
      if (ROWS > 4) {
        const int64_t nc = min(nb + 4 + (lane >> 3), n_obs - 1);
        mwb = mbits[nc * WPRT + (lane & 7)];
      }
    }
    int cnt[ROWS];    // missing features of observation r
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      int cr = 0;
#pragma unroll
      for (int wd = 0; wd < WPRT; ++wd) {
        const uint32_t bits = __shfl_sync(FULL, (r < 4) ? mwa : mwb, (r & 3) * 8 + wd);
        if ((bits >> lane) & 1u) lst[r * DD + cr + __popc(bits & lt)] = (uint16_t)(wd * 32 + lane);
        cr += __popc(bits);
      }
      cnt[r] = cr;
    }
    __syncwarp();
    // ---- (b) K-compacted Gram matrices W_miss^T W_miss on the tensor cores (tiles 00, 01, 11), one observation at a
    //          time with its own loop bound; the last k-step is completed with the zero row of Ws ------------------------
    double stash[SC::LIST_SEP ? 1 : ROWS][6];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      double g00[2] = {0.0, 0.0}, g01[2] = {0.0, 0.0}, g11[2] = {0.0, 0.0};
      const int cr = cnt[r];
      const int nfull = cr >> 2;
      const uint16_t* lr = lst + r * DD;
#pragma unroll 4
      for (int ks = 0; ks < nfull; ++ks) {
        const int d = lr[ks * 4 + t];
        const double f0 = Ws[d * WPS + g], f1 = Ws[d * WPS + 8 + g];
        dmma(g00, f0, f0);
        dmma(g01, f0, f1);
        dmma(g11, f1, f1);
      }
      if (cr & 3) {   // warp-uniform: the last k-step is completed with the zero row of Ws
        const int e = nfull * 4 + t;
        const int d = (e < cr) ? (int)lr[e] : DD;
        const double f0 = Ws[d * WPS + g], f1 = Ws[d * WPS + 8 + g];
        dmma(g00, f0, f0);
        dmma(g01, f0, f1);
        dmma(g11, f1, f1);
      }
      if constexpr (SC::LIST_SEP) {
        double* Mh = Mw + r * MS_STRIDE;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          Mh[g * 17 + 2 * t + q] = c00[q] - g00[q];
          Mh[g * 17 + 8 + 2 * t + q] = c01[q] - g01[q];
          Mh[(8 + 2 * t + q) * 17 + g] = c01[q] - g01[q];
          Mh[(8 + g) * 17 + 8 + 2 * t + q] = c11[q] - g11[q];
        }
      } else {
        stash[r][0] = c00[0] - g00[0]; stash[r][1] = c00[1] - g00[1]; stash[r][2] = c01[0] - g01[0];
        stash[r][3] = c01[1] - g01[1]; stash[r][4] = c11[0] - g11[0]; stash[r][5] = c11[1] - g11[1];
      }
    }
    if constexpr (!SC::LIST_SEP) {
      __syncwarp();   // the index lists (aliased by the staging area) are dead from here on
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        double* Mh = Mw + r * MS_STRIDE;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          Mh[g * 17 + 2 * t + q] = stash[r][q];
          Mh[g * 17 + 8 + 2 * t + q] = stash[r][2 + q];
          Mh[(8 + 2 * t + q) * 17 + g] = stash[r][2 + q];
          Mh[(8 + g) * 17 + 8 + 2 * t + q] = stash[r][4 + q];
        }
      }
    }
    __syncwarp();
    // ---- (c) sweep layout: m[q][a] = M[a][j + LPR q] (staged as C0 - Gram) ------------------------------------------------------
    const bool valid = nb + rw < n_obs;
    const int64_t n = valid ? nb + rw : n_obs - 1;
    int nmiss = 0;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) nmiss = (rw == r) ? cnt[r] : nmiss;
    double m[CPL][MF];
    {
      const double* Mh = Mw + rw * MS_STRIDE;
#pragma unroll
      for (int q = 0; q < CPL; ++q)
#pragma unroll
        for (int a = 0; a < MF; ++a) m[q][a] = Mh[a * 17 + j + LPR * q];
    }
    __syncwarp();   // the staging area is free again (reused for the output rows below)
    // every feature missing: unsupported by the reference (Util.hs:81); flagged, outputs zeroed below
    const bool empty = nmiss >= DD;
    if (empty && j == 0) set_err(p.scal, 2.0);
    // ---- (d) in-place inverse by symmetric sweeps (masked_sweeps.cuh): m becomes the lane's columns of M_i^-1 ---------
    bool ok = true;
    double pprod = 1.0;
    sweep_invert16<LPR, LL>(m, j, ok, pprod);
    if (!ok && j == 0) set_err(p.scal, 1.0);
    // ---- (e) ez = M_i^-1 b ;  A_i = ez ez^T + s2 M_i^-1 ---------------------------------------------------------------
    double ez[CPL], bown[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) ez[q] = bown[q] = 0.0;
    {
      const double2* bp = reinterpret_cast<const double2*>(bsrc + n * MF);
#pragma unroll
      for (int h = 0; h < MF / 2; ++h) {
        const double2 bv = bp[h];
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
          ez[q] = fma(m[q][2 * h], bv.x, ez[q]);
          ez[q] = fma(m[q][2 * h + 1], bv.y, ez[q]);
        }
        // b[j + LPR q] for the likelihood dot product: element a belongs to this lane when a % LPR == j
        if (((2 * h) % LPR) == j) bown[(2 * h) / LPR] = bv.x;
        if (((2 * h + 1) % LPR) == j) bown[(2 * h + 1) / LPR] = bv.y;
      }
    }
    if (!LL) {
      // rows [vech A | ez] of the warp's observations into the staging area, then out as ONE contiguous block
      double* srow = Mw + rw * EAP;
#pragma unroll
      for (int a = 0; a < MF; ++a) {
        const double ea = __shfl_sync(FULL, ez[a / LPR], a % LPR, LPR);
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
          if (a < LPR * q) continue;   // compile time: columns j + LPR q > a lie above the diagonal for every lane
          const int c = j + LPR * q;
          const double v = empty ? 0.0 : fma(ea, ez[q], s2 * m[q][a]);
          if (a >= c) srow[a * (a + 1) / 2 + c] = v;
          // A_i is positive semi-definite: its largest |entry| is on the diagonal
          if (a == c && valid) mxa = max(mxa, (uint32_t)__double2hiint(v) & 0x7fffffffu);
        }
      }
#pragma unroll
      for (int q = 0; q < CPL; ++q) {
        const double v = empty ? 0.0 : ez[q];
        srow[NVF + j + LPR * q] = v;
        if (valid) mxe = max(mxe, (uint32_t)__double2hiint(v) & 0x7fffffffu);
      }
      __syncwarp();
      const int nrows = (int)min((int64_t)ROWS, n_obs - nb);   // <= 0 for a warp past the end
      if (nrows > 0) {
        const double2* src = reinterpret_cast<const double2*>(Mw);
        double2* dst = reinterpret_cast<double2*>(EA + nb * EASF);   // 1216-byte rows: 16-byte aligned
        for (int e = lane; e < nrows * (EASF / 2); e += 32) dst[e] = src[(e / (EASF / 2)) * (EAP / 2) + e % (EASF / 2)];
      }
      __syncwarp();   // the next iteration's lists / matrices reuse the area
    } else {
      double dot = 0.0;
#pragma unroll
      for (int q = 0; q < CPL; ++q) dot = fma(bown[q], ez[q], dot);
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(FULL, dot, o, LPR);
      if (j == 0 && valid && !empty) {
        const double cntd = (double)(DD - nmiss);
        const double logdetM = log(pprod);
        llacc += cntd * log2pi + ((cntd - (double)MF) * logs2 + logdetM) + (yy[n] - dot) / s2;
      }
    }
  }
  if (LL) {
    const double tsum = block_sum(llacc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = tsum;
  } else if (eamax != nullptr) {
    mxa = __reduce_max_sync(FULL, mxa);
    mxe = __reduce_max_sync(FULL, mxe);
    if (lane == 0) {
      atomicMax(&eamax[0], mxa);   // max is order-independent: deterministic
      atomicMax(&eamax[1], mxe);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Complement sums on the FP64 tensor cores:  miss (D x 152) = Miss^T (D x n) . [vech A | ez] (n x 152), where
// Miss[i][d] = 1.0 when observation i misses feature d.  MMA axes: M = features, N = the 152 statistics, K =
// observations.  A fragments are decoded from the packed mask bits (exact 0.0 / 1.0), B fragments come straight
// from the TMA-staged [vech A | ez] tile (row pitch 1216 B => the 4-row x 8-column fragment reads are
// bank-conflict free).  8 warps x 16 features per CTA; grid = (row blocks, D / 128).  Warp 0 of feature group 0
// also accumulates the column totals.  Fixed accumulation order => deterministic.
// partials: [row block][(D + 1) x 152]  (row D = totals)
constexpr int KB_TR = 32;       // observations per stage
constexpr int KB_STAGES = 3;
constexpr int KB_NT = EASF / 8; // 19 n-tiles

template <int DD>
__global__ void __launch_bounds__(256, 1) masked_miss_dmma_kernel(const double* __restrict__ EA, const uint32_t* __restrict__ mbits,
                                                                 int64_t n_obs, int64_t rows_per_cta, double* __restrict__ partials,
                                                                 int64_t Spart, const double* __restrict__ scal) {
  constexpr int WPRT = DD / 32;
  constexpr int EA_BYTES = KB_TR * EASF * 8;
  constexpr int MK_BYTES = KB_TR * WPRT * 4;
  constexpr int STAGE_BYTES = EA_BYTES + MK_BYTES;
  if (scal[S_DONE] != 0.0) return;
  extern __shared__ __align__(128) uint8_t kb_smem[];
  const uint32_t sm_u = smem_u32(kb_smem);
  const uint32_t bar_u = sm_u + KB_STAGES * STAGE_BYTES;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int f0 = ((int)blockIdx.y * 8 + w) * 16;     // this warp's 16 features (2 m-tiles)
  const bool totals = (blockIdx.y == 0);   // every warp of feature group 0 sums 1/8 of the rows (load balance)
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = min(n_obs, r0 + rows_per_cta);
  const int64_t ntiles = (r1 > r0) ? (r1 - r0 + KB_TR - 1) / KB_TR : 0;

  // stale shared memory must not hold NaN patterns: rows past the end of a partial tile are multiplied by 0.0
  for (int e = tid; e < KB_STAGES * STAGE_BYTES / 8; e += 256) reinterpret_cast<double*>(kb_smem)[e] = 0.0;
  if (tid == 0) {
    for (int s = 0; s < KB_STAGES; ++s) mbar_init(bar_u + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  auto issue = [&](int64_t i) {
    const int s = (int)(i % KB_STAGES);
    const int64_t ra = r0 + i * KB_TR;
    const int rows = (int)min((int64_t)KB_TR, r1 - ra);
    const uint32_t dst = sm_u + s * STAGE_BYTES;
    mbar_expect_tx(bar_u + 8 * s, (uint32_t)(rows * EASF * 8 + MK_BYTES));
    bulk_load_1d(dst, EA + ra * EASF, (uint32_t)(rows * EASF * 8), bar_u + 8 * s);
    bulk_load_1d(dst + EA_BYTES, mbits + ra * WPRT, (uint32_t)MK_BYTES, bar_u + 8 * s);  // mbits is padded by >= KB_TR rows
  };
  if (tid == 0)
    for (int64_t i = 0; i < KB_STAGES && i < ntiles; ++i) issue(i);

  double acc[2][KB_NT][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < KB_NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
  double tot[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  const int word = f0 >> 5, shift = (f0 & 31) + g;   // bit of feature f0 + g ; + 8 for the second m-tile

  for (int64_t i = 0; i < ntiles; ++i) {
    const int s = (int)(i % KB_STAGES);
    const int rows = (int)min((int64_t)KB_TR, r1 - (r0 + i * KB_TR));
    mbar_wait(bar_u + 8 * s, (uint32_t)((i / KB_STAGES) & 1));
    const uint32_t tile_u = sm_u + s * STAGE_BYTES;
    const uint32_t* mk = reinterpret_cast<const uint32_t*>(kb_smem + s * STAGE_BYTES + EA_BYTES);
    if (totals) {
      const double* tile = reinterpret_cast<const double*>(kb_smem + s * STAGE_BYTES);
      for (int r = w; r < rows; r += 8) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const int e = lane + 32 * k;
          if (e < EASF) tot[k] += tile[r * EASF + e];
        }
      }
    }
#pragma unroll 2
    for (int ks = 0; ks < KB_TR / 4; ++ks) {
      const int r = ks * 4 + t;
      const uint32_t bits = (r < rows) ? (mk[r * WPRT + word] >> shift) : 0u;
      const double a0 = (bits & 1u) ? 1.0 : 0.0;
      const double a1 = (bits & 0x100u) ? 1.0 : 0.0;
      const uint32_t brow = tile_u + (uint32_t)((r * EASF + g) * 8);
#pragma unroll
      for (int nt = 0; nt < KB_NT; ++nt) {
        const double b = lds64(brow + nt * 64);
        dmma(acc[0][nt], a0, b);
        dmma(acc[1][nt], a1, b);
      }
    }
    __syncthreads();  // every warp is done with this stage
    if (tid == 0 && i + KB_STAGES < ntiles) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(i + KB_STAGES);
    }
  }
  double* out = partials + (int64_t)blockIdx.x * Spart;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < KB_NT; ++nt)
      *reinterpret_cast<double2*>(&out[(int64_t)(f0 + mt * 8 + g) * EASF + nt * 8 + 2 * t]) =
          make_double2(acc[mt][nt][0], acc[mt][nt][1]);
  if (totals) {  // the 8 per-warp partial totals are combined in fixed order through shared memory
    __syncthreads();
    double* red = reinterpret_cast<double*>(kb_smem);
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int e = lane + 32 * k;
      if (e < EASF) red[w * EASF + e] = tot[k];
    }
    __syncthreads();
    for (int e = tid; e < EASF; e += 256) {
      double sacc = 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) sacc += red[q * EASF + e];
      out[(int64_t)DD * EASF + e] = sacc;
    }
  }
}

template <int DD>
int launch_miss_dmma(ShardState& sh, int rb, int64_t rows_per_cta, int64_t Spart, const double* scal) {
  const int smem = KB_STAGES * (KB_TR * EASF * 8 + KB_TR * (DD / 32) * 4) + 64;
  static bool attr_done[64] = {};
  if (!attr_done[sh.dev & 63]) {
    CUDA_TRY(cudaFuncSetAttribute((masked_miss_dmma_kernel<DD>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done[sh.dev & 63] = true;
  }
  dim3 grid(rb, DD / 128 > 0 ? DD / 128 : 1);
  GP_LAUNCH((masked_miss_dmma_kernel<DD>), grid, 256, smem, sh.st, sh.EA, sh.mbits, sh.n, rows_per_cta, sh.partials, Spart, scal);
  return GPLVM_OK;
}

template <int DD, bool LL, int LPR>
int launch_solve(ShardState& sh, const double* bsrc, const double* yy, int grid, const double* gate) {
  constexpr int smem = SolveCfg<LPR>::smem_bytes(DD);
  static bool attr_done[64] = {};
  if (!attr_done[sh.dev & 63]) {
    CUDA_TRY(cudaFuncSetAttribute((masked_solve_kernel<DD, LL, LPR>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done[sh.dev & 63] = true;
  }
  // class maxima + precision-guard counters (gated: the tensor-core solver cleared them and may have written the maxima)
  if (!LL && sh.eamax && !gate) CUDA_TRY(cudaMemsetAsync(sh.eamax, 0, 24, sh.st));
  GP_LAUNCH((masked_solve_kernel<DD, LL, LPR>), grid, SOLVE_WARPS * 32, smem, sh.st, bsrc, yy, sh.mbits, sh.n, sh.par, sh.EA, sh.partials,
            LL ? nullptr : sh.eamax, gate);
  return GPLVM_OK;
}

// 4 lanes per observation (8 observations per warp, 1 CTA per SM): 12.6 ms at N = 2^23 against 13.6 ms for 8 lanes per
// observation at 2 CTAs per SM (round-2 measurement, profiles/r2_masked_summary.md)
template <int DD>
int dispatch_solve(ShardState& sh, const double* bsrc, const double* yy, bool ll, int* grid_out, const double* gate) {
  constexpr int LPR = 4;
  int grid = (int)std::min<int64_t>(SolveCfg<LPR>::MINB * sm_count(sh.dev), ceil_div(sh.n, SOLVE_WARPS * SolveCfg<LPR>::ROWS));
  if (grid > sh.nparts) grid = sh.nparts;
  if (grid_out) *grid_out = grid;
  return ll ? launch_solve<DD, true, LPR>(sh, bsrc, yy, grid, nullptr) : launch_solve<DD, false, LPR>(sh, bsrc, yy, grid, gate);
}

}  // namespace

// E-step solve (ll = false) or likelihood terms (ll = true; per-CTA sums land in sh.partials[0 .. *grid_out))
int masked_fused_solve(gplvm_ppca_session* s, ShardState& sh, const double* bsrc, const double* yy, bool ll, int* grid_out) {
  // E-step at D = 256: Gram matrices from the tcgen05 fixed-point product (masked_gram_tc.cu); the FP64 DMMA solver below is
  // launched behind it and runs only when the precision gate of that path is closed.  GPLVM_GRAM_DMMA=1: FP64 DMMA only.
  static const bool force_dmma = []() {
    const char* e = getenv("GPLVM_GRAM_DMMA");
    return e ? (atoi(e) != 0) : false;
  }();
  const double* gate = nullptr;
  if (!ll && !force_dmma && s->D == 256 && s->M == 16 && sh.gramB && sh.gramT) {
    GP_TRY(masked_gram_tc_solve(s, sh, bsrc));
    gate = sh.gramT + masked_gram_tc_gate_index();
  }
  switch (s->D) {
    case 64: return dispatch_solve<64>(sh, bsrc, yy, ll, grid_out, gate);
    case 128: return dispatch_solve<128>(sh, bsrc, yy, ll, grid_out, gate);
    case 256: return dispatch_solve<256>(sh, bsrc, yy, ll, grid_out, gate);
    default: return fail(GPLVM_ERR_STATE, "fused masked solver: unsupported D");
  }
}


// complement sums [G^miss | S1^miss] (D x 152) + totals (152) into dst (length (D + 1) x 152)
int masked_fused_miss_sums(gplvm_ppca_session* s, ShardState& sh, double* dst) {
  // default: exact fixed-point sums on the INT8 tensor cores (tcgen05, masked_utc.cu), followed by the FP64 DMMA
  // product as a device-side fallback that runs only when the precision guard of the INT8 path fired (its `scal` block is the guard's:
  // S_DONE = 1.0 makes both kernels return at once).  GPLVM_MISS_DMMA=1 selects the FP64 product unconditionally.
  static const bool force_dmma = []() {
    const char* e = getenv("GPLVM_MISS_DMMA");
    return e ? (atoi(e) != 0) : false;
  }();
  const bool imma = !force_dmma && sh.mbitsT;
  if (imma) GP_TRY(masked_utc_miss_sums(s, sh, dst));
  double* scal = imma ? fx::guard_scal(sh.eamax) : sh.par.scal;
  const int64_t Spart = (int64_t)(s->D + 1) * EASF;
  const int fgroups = (int)std::max<int64_t>(1, s->D / 128);
  int rb = (int)std::min<int64_t>(std::max<int64_t>(1, sm_count(sh.dev) / fgroups), sh.nparts);
  const int64_t rows_per_cta = round_up(std::max<int64_t>(ceil_div(sh.n, rb), 4), 4);
  rb = (int)ceil_div(sh.n, rows_per_cta);
  int rc;
  switch (s->D) {
    case 128: rc = launch_miss_dmma<128>(sh, rb, rows_per_cta, Spart, scal); break;
    case 256: rc = launch_miss_dmma<256>(sh, rb, rows_per_cta, Spart, scal); break;
    default: return fail(GPLVM_ERR_STATE, "fused masked statistics: unsupported D");
  }
  GP_TRY(rc);
  return launch_reduce_partials(sh, Spart, rb, dst, scal);
}

}  // namespace gplvm
