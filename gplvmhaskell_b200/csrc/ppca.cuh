// ppca.cuh -- PPCA session state shared by dense.cu / dense_dmma.cu / masked.cu / session.cu.
#pragma once
#include <functional>

#include "common.cuh"

namespace gplvm {

// scalar slots of the device parameter block
enum Scal {
  S_SIGMA2 = 0,   // current sigma^2
  S_TOTAL = 1,    // dense: |Xc|_F^2 (constant of the data, PPCA.hs:77)
  S_LOGDETM = 2,  // log det (sigma^2 I + W^T W) over the first M rows
  S_DVAR = 3,     // |sigma2' - sigma2| of the last step (PPCA.hs:82,170)
  S_MAXDW = 4,    // max |W - W'| of the last step (PPCA.hs:81,171)
  S_DONE = 5,     // 1.0 once the Right-tol rule has fired (further steps are no-ops)
  S_STEPS = 6,    // EM steps executed
  S_ERR = 7,      // != 0: numeric failure (non positive pivot)
  S_TOL = 8,      // tolerance of the Right rule (used when S_USETOL != 0)
  S_USETOL = 9,
  S_LL = 10,      // log-likelihood written by the finalize kernels
  S_KNOWN = 11,   // masked: number of known entries (knownVars, PPCA.hs:146)
  S_ZBAR0 = 12,   // unused
  S_COUNT = 16
};

// Device pointers into one shard's parameter block (passed to kernels by value).
struct Par {
  double* W;     // D x MP   current W (padded columns are zero)
  double* mu;    // D        dense: feature means (constant); masked: EM mean
  double* A;     // D x MP   dense: W * Minv
  double* Minv;  // MP x MP  dense: (sigma2 I + W^T W)^-1 ; masked: C0 = sigma2 I + W^T W
  double* c;     // MP       masked: mean of E[z_i] over the observations (restore)
  double* nd;    // D        masked: number of observations knowing feature d (global)
  double* sy;    // D        masked: sum of known y_d
  double* syy;   // D        masked: sum of known y_d^2
  double* Wprev; // D x MP   masked: scratch for the update (new W before commit)
  double* scal;  // S_COUNT scalars
  int D, M, MP, NV;
  long long N;   // global number of observations
};

#ifdef __CUDACC__
// numeric-failure flag: 1 = not positive definite, 2 = an observation with every feature missing (sticky)
__device__ __forceinline__ void set_err(double* scal, double code) {
  if (code == 2.0 || scal[S_ERR] == 0.0) scal[S_ERR] = code;
}
#endif

// peer-exchange arguments of the fused all-reduce + update kernel
struct PeerArgs {
  double* const* peers;      // device array of `world` peer-mapped buffers, nullptr: statistics are local / already reduced
  int world, rank, parity;
  unsigned long long epoch;
  long long S;
};

struct ShardState {
  int dev = 0;
  cudaStream_t st = nullptr;
  int64_t n0 = 0, n = 0;  // global begin / local count
  int64_t ldx = 0;
  double* X = nullptr;         // n x ldx, observation-major; NaN = missing
  double* Ez = nullptr;        // n x MP (dense generic path / masked path)
  double* EA = nullptr;        // masked: n x EAS, row i = [vech(A_i) (NV) | ez_i (M)], A_i = ez ez^T + sigma2 Minv_i
  double* yy = nullptr;        // masked fused path: |(y_i - mu)_observed|^2 per observation
  uint32_t* mbits = nullptr;   // masked: n x WPR words, bit d set = feature d of the observation is missing
  uint32_t* mbitsT = nullptr;  // masked fused path: D x nwT words, bit j of word w = observation 32 w + j misses the feature
  int64_t nwT = 0;             //   4 x ceil(n / 128) words per feature
  uint32_t* eamax = nullptr;   // masked fused path: 256 B block: class maxima of the last solve, precision-guard counters, guard flags (masked_fx.cuh)
  uint8_t* gramB = nullptr;    // masked tensor-core Gram solver (masked_gram_tc.cu): byte-plane image of W (x) W, rebuilt every step
  double* gramT = nullptr;     //   its scale / C0 tables and the precision gate
  double* parblock = nullptr;  // parameter block
  double* stats = nullptr;     // packed statistics (all-reduced), `S` doubles
  double* partials = nullptr;  // nparts x S
  double* scratch = nullptr;   // per-feature scratch (masked update), D doubles x 4
  void* tmap = nullptr;        // host copy of the CUtensorMap for X (dense DMMA path)
  // peer exchange of the per-step statistics over NVLink (dense fast path, world > 1): every rank keeps two
  // publication buffers (step parity) + flags in its own HBM; the update kernel reads all ranks' buffers directly
  double* xchg = nullptr;          // 2 x S doubles, then 2 x ceil(S / 32) uint64 epoch flags (per parity, per block)
  double** peers_dev = nullptr;    // device array [world] of peer-mapped `xchg` pointers (own entry included)
  std::vector<void*> ipc_opened;   // peer mappings opened with cudaIpcOpenMemHandle (multi-process mode)
  PeerArgs peer_args{nullptr, 1, 0, 0, 0ull, 0};  // non-null peers: the next reduction is also the all-reduce
  int nparts = 0;
  Par par;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk0 = nullptr, evk1 = nullptr;
  // per-kernel timing of one EM step (gplvm_session_set_kernel_timing): mark i is recorded before segment i starts
  static constexpr int MAX_MARKS = 10;
  cudaEvent_t evseg[MAX_MARKS] = {};
  int nmarks = 0;   // marks recorded by the last step
};

// record segment mark `i` on the shard's stream when per-kernel timing is on (no-op otherwise)
inline void seg_mark(ShardState& sh, int i) {
  if (sh.evk0 && i < ShardState::MAX_MARKS && sh.evseg[i]) {
    cudaEventRecord(sh.evseg[i], sh.st);
    if (i + 1 > sh.nmarks) sh.nmarks = i + 1;
  }
}

}  // namespace gplvm

struct gplvm_ppca_session {
  int64_t D = 0, N = 0, M = 0;
  int MP = 0, NV = 0;
  int EAS = 0, WPR = 0;    // masked: row stride of EA (NV + M), mask words per observation
  bool missing = false;
  bool has_data = false, initialised = false;
  bool centred = false;    // dense: the resident observations have had their feature means subtracted (PPCA.hs:61)
  bool use_dmma = false;   // dense: fused DMMA kernel eligible
  int64_t S = 0;           // packed statistics length
  int64_t steps = 0;       // steps enqueued so far
  bool p2p = false;        // statistics are exchanged by the update kernel over peer memory (no NCCL call per step)
  uint64_t xchg_epoch = 0; // number of exchanges published so far (monotonic; selects the buffer parity)
  std::vector<gplvm::ShardState> sh;
};

namespace gplvm {

// dense.cu
int dense_init_constants(gplvm_ppca_session* s);
int dense_enqueue_step(gplvm_ppca_session* s);
int dense_loglik(gplvm_ppca_session* s, double* ll);
// dense_dmma.cu
bool dense_dmma_eligible(const gplvm_ppca_session* s);
int dense_dmma_prepare(gplvm_ppca_session* s);
int dense_dmma_stats(gplvm_ppca_session* s, ShardState& sh, const double* proj);
int dense_dmma_release(gplvm_ppca_session* s);
int masked_dmma_b(gplvm_ppca_session* s, ShardState& sh, const double* W, const double* mu, double* bout, double* yyout);
int masked_fused_solve(gplvm_ppca_session* s, ShardState& sh, const double* bsrc, const double* yy, bool ll, int* grid_out);
int masked_fused_miss_sums(gplvm_ppca_session* s, ShardState& sh, double* dst);
int masked_utc_prepare(gplvm_ppca_session* s, ShardState& sh);
int masked_utc_miss_sums(gplvm_ppca_session* s, ShardState& sh, double* dst);
// masked_gram_tc.cu: E-step solve with tcgen05 fixed-point Gram matrices (D = 256, M = 16); gate: gramT + masked_gram_tc_gate_index()
int masked_gram_tc_solve(gplvm_ppca_session* s, ShardState& sh, const double* bsrc);
size_t masked_gram_tc_image_bytes();
size_t masked_gram_tc_table_bytes();
int masked_gram_tc_gate_index();
int masked_dmma_s2(gplvm_ppca_session* s, ShardState& sh, const double* ez, int64_t ldez, double* dst);
int dense_dmma_nparts(const gplvm_ppca_session* s, const ShardState& sh);
// masked.cu
int masked_init_constants(gplvm_ppca_session* s);
int masked_enqueue_step(gplvm_ppca_session* s);
int masked_loglik(gplvm_ppca_session* s, double* ll);
int masked_restore(gplvm_ppca_session* s, double* restored, int64_t ld);
// session.cu
int session_allreduce_stats(gplvm_ppca_session* s, size_t count);
int session_setup_p2p(gplvm_ppca_session* s);
int session_peer_allreduce_stats(gplvm_ppca_session* s);   // dense.cu: in-place all-reduce of sh.stats over peer memory
void session_release_p2p(gplvm_ppca_session* s);
// Drain a feature-major device producer into a host matrix (D rows, row stride ld doubles, this shard's n columns starting at
// dst).  produce(c0, cn, stage, pitch) enqueues on sh.st the kernels that fill stage[d * pitch + i], i < cn, for the
// observations [c0, c0 + cn).  Two stage buffers and a copy stream overlap the kernels with the D2H copies; pageable caller
// memory goes through pinned bounce buffers drained by several host threads.
int session_d2h_rows(ShardState& sh, int64_t D, int64_t n, double* dst, int64_t ld, int nshards,
                     const std::function<int(int64_t, int64_t, double*, int64_t)>& produce);
int launch_reduce_partials(ShardState& sh, int64_t S, int nparts, double* dst = nullptr, double* scal = nullptr);  // scal: S_DONE skip flag block, default the session's

// Generic split-K statistics GEMM shared by dense.cu and masked.cu (defined in dense.cu).
// out block 1: rows [0,R1) x Q1 columns, block 2: rows [R1,R) x Q2 columns (Q2 <= Q1).
// The partial results go through sh.partials (stride Sout per split) and are summed in fixed order into dst.
enum RowKind { ROWS_DENSE = 0, ROWS_MASKED_VALUES = 1 };
int launch_stats_gemm(ShardState& sh, RowKind kind, int D, int R1, int R, int Q1, int Q2, const double* Cmat,
                      int ldc, int Cw1, const double* Cmat2, int ldc2, int64_t Sout, double* dst);

}  // namespace gplvm
