// session.cu -- PPCA sessions: observation shards resident in HBM, host<->device movement in the
// reference layout (Y is D x N feature-major; the device keeps observation-major rows so that one
// observation is one contiguous, TMA-loadable row), the synthetic workload generator and the EM
// driver implementing the reference stop rule (src/GPLVM/PPCA.hs:90-92, :173-181).
#include <thread>

#include "ppca.cuh"

namespace gplvm {
int moments_pass(gplvm_ppca_session* s, int mode);
int dense_centre(gplvm_ppca_session* s, double sign);

namespace {

// out (cols x ld_out) <- in^T, in is (rows x ld_in); 32x32 tiles through shared memory.  rows_on_x: the row tiles of `in`
// ride on gridDim.x (the observation axis always does: gridDim.y is limited to 65535 tiles)
__global__ void __launch_bounds__(256) transpose_kernel(const double* __restrict__ in, int64_t ld_in, double* __restrict__ out,
                                                        int64_t ld_out, int64_t rows, int64_t cols, int rows_on_x) {
  __shared__ double t[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c0 = (int64_t)(rows_on_x ? blockIdx.y : blockIdx.x) * 32, r0 = (int64_t)(rows_on_x ? blockIdx.x : blockIdx.y) * 32;
  for (int j = ty; j < 32; j += 8) {
    const int64_t r = r0 + j, c = c0 + tx;
    t[j][tx] = (r < rows && c < cols) ? in[r * ld_in + c] : 0.0;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int64_t c = c0 + j, r = r0 + tx;
    if (c < cols && r < rows) out[c * ld_out + r] = t[tx][j];
  }
}

// synthetic model of SURVEY.md 8(d): x_n = W_t z_n + mu_t + eps_n, all N(0,1); keyed by global n
__global__ void __launch_bounds__(256) synth_kernel(double* __restrict__ X, int64_t ldx, int64_t n0, int64_t n_obs, int D,
                                                    uint64_t seed, double nan_prob) {
  constexpr int MT = 16;  // true latent dimension of the generator
  __shared__ double zs[16][MT];
  const int64_t base = (int64_t)blockIdx.x * 16;
  {
    const int o = threadIdx.x / MT, m = threadIdx.x % MT;
    double z, u;
    Philox::normal_uniform(seed ^ 0x5851F42D4C957F2DULL, (uint64_t)(n0 + base + o), (uint64_t)m, z, u);
    zs[o][m] = z;
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    double wt[MT], mu, u;
#pragma unroll
    for (int m = 0; m < MT; ++m) Philox::normal_uniform(seed ^ 0x14057B7EF767814FULL, (uint64_t)d, (uint64_t)m, wt[m], u);
    Philox::normal_uniform(seed ^ 0x14057B7EF767814FULL, (uint64_t)d, (uint64_t)MT, mu, u);
    for (int o = 0; o < 16; ++o) {
      const int64_t n = base + o;
      if (n >= n_obs) break;
      double eps;
      Philox::normal_uniform(seed, (uint64_t)(n0 + n), (uint64_t)d, eps, u);
      double v = mu + eps;
#pragma unroll
      for (int m = 0; m < MT; ++m) v = fma(wt[m], zs[o][m], v);
      if (nan_prob > 0.0 && d != 0 && u < nan_prob) v = nan("");
      X[n * ldx + d] = v;
    }
  }
}

__global__ void set_params_kernel(Par p, const double* __restrict__ W0, double sigma2_0) {
  // W0 arrives as D x M (reference layout); pad to D x MP with zeros
  for (int e = threadIdx.x + blockIdx.x * blockDim.x; e < p.D * p.MP; e += blockDim.x * gridDim.x) {
    const int d = e / p.MP, a = e % p.MP;
    p.W[e] = (a < p.M) ? W0[(int64_t)d * p.M + a] : 0.0;
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    for (int i = 0; i < S_COUNT; ++i)
      if (i != S_TOTAL && i != S_KNOWN) p.scal[i] = 0.0;
    p.scal[S_SIGMA2] = sigma2_0;
  }
}

__global__ void set_scal_kernel(double* scal, int idx0, double v0, int idx1, double v1) {
  scal[idx0] = v0;
  if (idx1 >= 0) scal[idx1] = v1;
}

// pinned bounce buffers of the pageable-memory load path, two per device, kept until gplvm_dist_finalize / a new context
struct Bounce { void* p = nullptr; size_t bytes = 0; };
Bounce g_bounce[64][2];
std::mutex g_bounce_mu;
double* bounce_get(int dev, int slot, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_bounce_mu);
  Bounce& b = g_bounce[dev & 63][slot & 1];
  if (b.p && b.bytes < bytes) { cudaFreeHost(b.p); b = Bounce(); }
  if (!b.p) {
    if (cudaHostAlloc(&b.p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); b = Bounce(); return nullptr; }
    b.bytes = bytes;
  }
  return static_cast<double*>(b.p);
}

int64_t chunk_obs(int64_t D) {
  int64_t c = ((int64_t)32 << 20) / (D > 0 ? D : 1);  // <= 256 MB staging
  c = (c / 32) * 32;
  if (c < 1024) c = 1024;
  return c;
}

void free_shard(ShardState& sh) {
  cudaSetDevice(sh.dev);
  cudaFree(sh.X); cudaFree(sh.Ez); cudaFree(sh.EA); cudaFree(sh.mbits); cudaFree(sh.mbitsT); cudaFree(sh.eamax); cudaFree(sh.gramB); cudaFree(sh.gramT); cudaFree(sh.yy); cudaFree(sh.parblock); cudaFree(sh.stats);
  cudaFree(sh.partials); cudaFree(sh.scratch);
  if (sh.tmap) free(sh.tmap);
  if (sh.ev0) cudaEventDestroy(sh.ev0);
  if (sh.ev1) cudaEventDestroy(sh.ev1);
  if (sh.evk0) cudaEventDestroy(sh.evk0);
  if (sh.evk1) cudaEventDestroy(sh.evk1);
  for (cudaEvent_t e : sh.evseg)
    if (e) cudaEventDestroy(e);
  sh = ShardState();
}

}  // namespace

static bool host_pageable(const void* p) {
  cudaPointerAttributes attr{};
  const bool pg = (cudaPointerGetAttributes(&attr, p) != cudaSuccess) || attr.type == cudaMemoryTypeUnregistered;
  cudaGetLastError();
  return pg;
}
static int host_threads(size_t nshards) {
  return std::max(1, std::min(16, (int)std::thread::hardware_concurrency() / (int)std::max<size_t>(1, nshards)));
}
// rows d = 0 .. D-1: copy `bytes` from src + d * src_ld to dst + d * dst_ld with HT threads (rows interleaved over the threads)
static void threaded_rows_copy(double* dst, int64_t dst_ld, const double* src, int64_t src_ld, int64_t D, size_t bytes, int HT) {
  auto work = [&](int t) {
    for (int64_t d = t; d < D; d += HT) memcpy(dst + d * dst_ld, src + d * src_ld, bytes);
  };
  if (HT <= 1) { work(0); return; }
  std::vector<std::thread> pool;
  for (int t = 1; t < HT; ++t) pool.emplace_back(work, t);
  work(0);
  for (std::thread& th : pool) th.join();
}

int session_d2h_rows(ShardState& sh, int64_t D, int64_t n, double* dst, int64_t ld, int nshards,
                     const std::function<int(int64_t, int64_t, double*, int64_t)>& produce) {
  CUDA_TRY(cudaSetDevice(sh.dev));
  const bool pageable = host_pageable(dst);
  const int HT = host_threads((size_t)nshards);
  int64_t ch = pageable ? ((int64_t)8 << 20) / D : ((int64_t)32 << 20) / D;   // 64 MB bounce buffers / 256 MB direct copies
  ch = std::min(std::max<int64_t>(1024, (ch / 32) * 32), n);
  double* stage[2] = {nullptr, nullptr};
  double* bounce[2] = {nullptr, nullptr};
  cudaEvent_t prod[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
  cudaStream_t cs = nullptr;
  cudaError_t e = cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
    e = cudaMalloc(&stage[i], sizeof(double) * (size_t)D * ch);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&prod[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming);
    if (e == cudaSuccess && pageable) {
      bounce[i] = bounce_get(sh.dev, i, sizeof(double) * (size_t)D * ch);
      if (!bounce[i]) e = cudaErrorMemoryAllocation;
    }
  }
  int rc = GPLVM_OK;
  int64_t chunk = 0, prev_c0 = 0, prev_cn = 0;
  for (int64_t c0 = 0; e == cudaSuccess && rc == GPLVM_OK && c0 < n; c0 += ch, ++chunk) {
    const int b = (int)(chunk & 1);
    const int64_t cn = std::min(ch, n - c0);
    if (chunk >= 2) e = cudaStreamWaitEvent(sh.st, copied[b], 0);   // stage[b] has been copied out
    if (e != cudaSuccess) break;
    rc = produce(c0, cn, stage[b], ch);
    if (rc != GPLVM_OK) break;
    e = cudaEventRecord(prod[b], sh.st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(cs, prod[b], 0);
    if (e != cudaSuccess) break;
    if (!pageable) {
      e = cudaMemcpy2DAsync(dst + c0, sizeof(double) * ld, stage[b], sizeof(double) * ch, sizeof(double) * cn, D, cudaMemcpyDeviceToHost, cs);
    } else {
      e = cudaMemcpyAsync(bounce[b], stage[b], sizeof(double) * (size_t)D * ch, cudaMemcpyDeviceToHost, cs);
    }
    if (e == cudaSuccess) e = cudaEventRecord(copied[b], cs);
    if (e == cudaSuccess && pageable && chunk >= 1) {   // drain the previous chunk on the host while this one is on the wire
      e = cudaEventSynchronize(copied[b ^ 1]);
      if (e == cudaSuccess) threaded_rows_copy(dst + prev_c0, ld, bounce[b ^ 1], ch, D, sizeof(double) * (size_t)prev_cn, HT);
    }
    prev_c0 = c0; prev_cn = cn;
  }
  if (e == cudaSuccess && rc == GPLVM_OK) e = cudaStreamSynchronize(cs);
  if (e == cudaSuccess && rc == GPLVM_OK && pageable && chunk >= 1)
    threaded_rows_copy(dst + prev_c0, ld, bounce[(chunk - 1) & 1], ch, D, sizeof(double) * (size_t)prev_cn, HT);
  cudaStreamSynchronize(sh.st);
  for (int i = 0; i < 2; ++i) {
    if (stage[i]) cudaFree(stage[i]);
    if (prod[i]) cudaEventDestroy(prod[i]);
    if (copied[i]) cudaEventDestroy(copied[i]);
  }
  if (cs) cudaStreamDestroy(cs);
  if (rc != GPLVM_OK) return rc;
  if (e != cudaSuccess) return fail(GPLVM_ERR_CUDA, "D2H copy failed: %s", cudaGetErrorString(e));
  return GPLVM_OK;
}

void session_release_bounce() {
  std::lock_guard<std::mutex> lk(g_bounce_mu);
  for (auto& dev : g_bounce)
    for (Bounce& b : dev) {
      if (b.p) cudaFreeHost(b.p);
      b = Bounce();
    }
}

int session_allreduce_stats(gplvm_ppca_session* s, size_t count) {
  std::vector<double*> bufs;
  for (ShardState& sh : s->sh) bufs.push_back(sh.stats);
  return allreduce_sum(bufs, count);
}

// Peer exchange of the dense per-step statistics (PeerArgs): every rank allocates a publication buffer and maps
// everyone else's -- directly (cudaDeviceEnablePeerAccess) when one process drives several GPUs, through CUDA IPC
// handles gathered over the NCCL communicator when there is one process per GPU.  Any failure leaves p2p = false and the
// step falls back to ncclAllReduce.  GPLVM_NO_P2P=1 disables it.
int session_setup_p2p(gplvm_ppca_session* s) {
  Context& c = ctx();
  s->p2p = false;
  const char* env = getenv("GPLVM_NO_P2P");
  const bool dense_ok = !s->missing && s->use_dmma && s->MP == 16 && s->D <= 1024;   // the fused dense step
  const bool masked_ok = s->missing && s->S <= (1 << 20);                               // packed masked statistics (345 KB at D = 256)
  if (c.world < 2 || !(dense_ok || masked_ok) || (env && atoi(env) != 0)) return GPLVM_OK;
  // [2 x S doubles | 2 x blocks uint64 flags | magic].  Rounded up to whole 2 MB granules: cudaIpcGetMemHandle names the
  // underlying ALLOCATION, and a small cudaMalloc may be sub-allocated at a non-zero offset inside one (the mapping a
  // peer opens would then point somewhere else).  Every mapping is additionally verified with a per-rank magic word.
  const size_t flag_words = 2 * (size_t)ceil_div(s->S, 32);
  const size_t magic_off = sizeof(double) * (2 * (size_t)s->S) + sizeof(uint64_t) * flag_words;   // byte offset
  const size_t bytes = (size_t)round_up((int64_t)(magic_off + 64), (int64_t)2 << 20);
  for (ShardState& sh : s->sh) {
    if (cudaSetDevice(sh.dev) != cudaSuccess || cudaMalloc(&sh.xchg, bytes) != cudaSuccess ||
        cudaMemset(sh.xchg, 0, bytes) != cudaSuccess || cudaMalloc(&sh.peers_dev, sizeof(double*) * c.world) != cudaSuccess) {
      cudaGetLastError();
      session_release_p2p(s);
      return GPLVM_OK;
    }
  }
  std::vector<double*> ptrs(c.world, nullptr);
  if (!c.multiproc) {
    for (size_t i = 0; i < s->sh.size(); ++i) {
      for (size_t j = 0; j < s->sh.size(); ++j) {
        if (i == j) continue;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, s->sh[i].dev, s->sh[j].dev);
        if (!can) { session_release_p2p(s); return GPLVM_OK; }
        cudaSetDevice(s->sh[i].dev);
        cudaError_t e = cudaDeviceEnablePeerAccess(s->sh[j].dev, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); session_release_p2p(s); return GPLVM_OK; }
        cudaGetLastError();
      }
    }
    for (size_t i = 0; i < s->sh.size(); ++i) ptrs[i] = s->sh[i].xchg;
    for (ShardState& sh : s->sh) {
      cudaSetDevice(sh.dev);
      if (cudaMemcpy(sh.peers_dev, ptrs.data(), sizeof(double*) * c.world, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError(); session_release_p2p(s); return GPLVM_OK;
      }
    }
  } else {
    ShardState& sh = s->sh[0];
    cudaSetDevice(sh.dev);
    struct Slot { cudaIpcMemHandle_t h; int ok; int pad[3]; };
    Slot mine{};
    const unsigned long long my_magic = 0x5eedULL * 1000003ULL + (unsigned long long)c.rank;
    const bool wrote = cudaMemcpy(reinterpret_cast<char*>(sh.xchg) + magic_off, &my_magic, sizeof my_magic, cudaMemcpyHostToDevice) == cudaSuccess;
    mine.ok = (wrote && cudaIpcGetMemHandle(&mine.h, sh.xchg) == cudaSuccess) ? 1 : 0;
    cudaGetLastError();
    std::vector<Slot> all(c.world);
    if (allgather_bytes(&mine, all.data(), sizeof(Slot)) != GPLVM_OK) { session_release_p2p(s); return GPLVM_OK; }
    bool ok = true;
    for (int r = 0; r < c.world; ++r) ok = ok && all[r].ok;
    int opened_all = 1;
    if (ok) {
      for (int r = 0; r < c.world; ++r) {
        if (r == c.rank) { ptrs[r] = sh.xchg; continue; }
        void* q = nullptr;
        if (cudaIpcOpenMemHandle(&q, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); opened_all = 0; break; }
        sh.ipc_opened.push_back(q);
        ptrs[r] = static_cast<double*>(q);
        unsigned long long seen = 0;   // the mapping must show rank r's magic word
        if (cudaMemcpy(&seen, static_cast<char*>(q) + magic_off, sizeof seen, cudaMemcpyDeviceToHost) != cudaSuccess ||
            seen != 0x5eedULL * 1000003ULL + (unsigned long long)r) {
          cudaGetLastError();
          opened_all = 0;
          break;
        }
      }
    } else {
      opened_all = 0;
    }
    // every rank must take the same decision: gather the outcome
    std::vector<int> outcome(c.world, 0);
    if (allgather_bytes(&opened_all, outcome.data(), sizeof(int)) != GPLVM_OK) { session_release_p2p(s); return GPLVM_OK; }
    for (int r = 0; r < c.world; ++r) opened_all = opened_all && outcome[r];
    if (!opened_all || cudaMemcpy(sh.peers_dev, ptrs.data(), sizeof(double*) * c.world, cudaMemcpyHostToDevice) != cudaSuccess) {
      cudaGetLastError(); session_release_p2p(s); return GPLVM_OK;
    }
  }
  s->p2p = true;
  s->xchg_epoch = 0;
  return GPLVM_OK;
}

void session_release_p2p(gplvm_ppca_session* s) {
  for (ShardState& sh : s->sh) {
    cudaSetDevice(sh.dev);
    for (void* q : sh.ipc_opened) cudaIpcCloseMemHandle(q);
    sh.ipc_opened.clear();
    if (sh.xchg) cudaFree(sh.xchg);
    if (sh.peers_dev) cudaFree(sh.peers_dev);
    sh.xchg = nullptr;
    sh.peers_dev = nullptr;
  }
  s->p2p = false;
}

// Synchronise every shard and surface numeric failures.  collective: in multi-process mode the ranks exchange their
// error flags (S_ERR is rank-local: an all-NaN observation or a non positive definite M_i lives in one shard), so that every
// rank leaves a collective entry point (init / steps+sync / run) with the same verdict instead of one rank returning an
// error while its peers wait in the next all-reduce or return statistics poisoned by the failed shard.
static int check_async(gplvm_ppca_session* s, double* scal_out, bool collective = false) {
  double scal[S_COUNT], first[S_COUNT];
  double worst = 0.0;
  for (ShardState& sh : s->sh) {
    CUDA_TRY(cudaSetDevice(sh.dev));
    CUDA_TRY(cudaMemcpyAsync(scal, sh.par.scal, sizeof scal, cudaMemcpyDeviceToHost, sh.st));
    CUDA_TRY(cudaStreamSynchronize(sh.st));
    if (scal[S_ERR] > worst) worst = scal[S_ERR];   // 3 (expired wait) > 2 (all-NaN observation) > 1 (not positive definite)
    if (&sh == &s->sh[0]) memcpy(first, scal, sizeof scal);
  }
  if (collective && ctx().multiproc && ctx().world > 1) {
    std::vector<double> all(ctx().world, 0.0);
    GP_TRY(allgather_bytes(&worst, all.data(), sizeof(double)));
    for (double v : all)
      if (v > worst) worst = v;
  }
  if (worst == 2.0)
    return fail(GPLVM_ERR_UNSUPPORTED, "missing-data PPCA: an observation has every feature missing (unsupported by the reference, Util.hs:81)");
  if (worst == 3.0)
    return fail(GPLVM_ERR_NCCL, "a bounded device-side wait expired (peer statistics exchange: a rank did not publish its statistics; or the tcgen05 producer / MMA handshake of the masked complement sums)");
  if (worst != 0.0)
    return fail(GPLVM_ERR_NUMERIC, "PPCA: a matrix that must be positive definite is not (step %lld)", (long long)first[S_STEPS]);
  if (scal_out) memcpy(scal_out, first, sizeof first);
  return GPLVM_OK;
}

}  // namespace gplvm

using namespace gplvm;

extern "C" {

int gplvm_session_create(int64_t D, int64_t N_global, int64_t M, int missing, int64_t obs_begin, int64_t obs_count,
                         gplvm_ppca_session** out) {
  if (!out) return fail(GPLVM_ERR_ARG, "session_create: null out");
  *out = nullptr;
  if (D < 1 || N_global < 1 || M < 1) return fail(GPLVM_ERR_ARG, "session_create: D, N, M must be positive");
  if (M > GPLVM_MAX_M) return fail(GPLVM_ERR_ARG, "session_create: M = %lld exceeds GPLVM_MAX_M = %d", (long long)M, GPLVM_MAX_M);
  if (D > (1 << 20)) return fail(GPLVM_ERR_ARG, "session_create: D too large");
  if (obs_begin < 0 || obs_count < 1 || obs_begin + obs_count > N_global)
    return fail(GPLVM_ERR_ARG, "session_create: observation block [%lld, +%lld) outside [0, %lld)", (long long)obs_begin,
                (long long)obs_count, (long long)N_global);
  GP_TRY(ensure_context());
  Context& c = ctx();
  std::lock_guard<std::recursive_mutex> lk(c.mu);
  if (!c.multiproc && (obs_begin != 0 || obs_count != N_global))
    return fail(GPLVM_ERR_ARG, "session_create: partial observation blocks need gplvm_dist_init (multi-process mode)");
  if (c.multiproc && c.world > 1) {
    // the rank blocks must partition [0, N) in rank order and describe the same problem: overlapping blocks would be
    // summed `world` times by the all-reduce while N stays N (silently wrong W and sigma2)
    struct Desc { int64_t D, N, M, missing, begin, count; } mine{D, N_global, M, missing != 0, obs_begin, obs_count};
    std::vector<Desc> all(c.world);
    GP_TRY(allgather_bytes(&mine, all.data(), sizeof(Desc)));
    int64_t next = 0;
    for (int r = 0; r < c.world; ++r) {
      if (all[r].D != D || all[r].N != N_global || all[r].M != M || all[r].missing != (missing != 0))
        return fail(GPLVM_ERR_ARG, "session_create: rank %d describes a different problem (D, N, M, missing)", r);
      if (all[r].begin != next)
        return fail(GPLVM_ERR_ARG, "session_create: the rank blocks do not partition [0, %lld): rank %d starts at %lld, expected %lld",
                    (long long)N_global, r, (long long)all[r].begin, (long long)next);
      next += all[r].count;
    }
    if (next != N_global)
      return fail(GPLVM_ERR_ARG, "session_create: the rank blocks cover %lld of %lld observations", (long long)next, (long long)N_global);
  }
  auto* s = new gplvm_ppca_session();
  s->D = D; s->N = N_global; s->M = M; s->missing = missing != 0;
  s->use_dmma = dense_dmma_eligible(s);
  s->MP = s->missing ? (int)M : (s->use_dmma ? 16 : (int)round_up(M, 8));
  if (s->MP > GPLVM_MAX_M) s->MP = GPLVM_MAX_M;
  s->NV = s->MP * (s->MP + 1) / 2;
  s->EAS = s->NV + s->MP;
  s->WPR = (int)ceil_div(D, 32);
  if (s->missing)  // complement sums [G^miss | S1^miss] (D x EAS) ; totals (EAS) ; S2 (D x M)
    s->S = (D + 1) * (int64_t)s->EAS + D * s->MP;
  else             // XEz (D x MP) | EE (MP x MP)
    s->S = (D + s->MP) * s->MP;
  const int nsh = (int)c.shards.size();
  int used = nsh;
  if (obs_count < used) used = (int)obs_count;
  // balanced split: the first obs_count % used shards take one observation more, so no shard is ever empty
  const int64_t base = obs_count / used, extra = obs_count % used;
  s->sh.resize(used);
  for (int i = 0; i < used; ++i) {
    ShardState& sh = s->sh[i];
    sh.dev = c.shards[i].device;
    sh.st = c.shards[i].stream;
    sh.n0 = obs_begin + base * i + std::min<int64_t>(i, extra);
    sh.n = base + (i < extra ? 1 : 0);
  }
  if (used != nsh) {
    delete s;
    return fail(GPLVM_ERR_ARG, "session_create: fewer observations than GPUs");
  }
  const int MP = s->MP;
  for (ShardState& sh : s->sh) {
    cudaError_t e = cudaSetDevice(sh.dev);
    sh.ldx = round_up(D, 4);
    if (s->use_dmma) sh.ldx = D;
    sh.nparts = 592;
    if (s->use_dmma) sh.nparts = std::max(sh.nparts, dense_dmma_nparts(s, sh));
    const size_t parlen = (size_t)D * MP * 3 + (size_t)MP * MP + MP + (size_t)D * 4 + S_COUNT + 64;
    if (e == cudaSuccess) e = cudaMalloc(&sh.X, sizeof(double) * (size_t)sh.n * sh.ldx);
    if (e == cudaSuccess && (s->missing == s->use_dmma)) e = cudaMalloc(&sh.Ez, sizeof(double) * (size_t)sh.n * MP);  // generic dense Ez / fused masked b
    if (e == cudaSuccess && s->missing && s->use_dmma) e = cudaMalloc(&sh.yy, sizeof(double) * (size_t)sh.n);
    if (e == cudaSuccess && s->missing) e = cudaMalloc(&sh.EA, sizeof(double) * ((size_t)round_up(sh.n, 128) * s->EAS + 8));  // whole 128-observation chunks (masked_utc.cu)
    if (e == cudaSuccess && s->missing) e = cudaMalloc(&sh.mbits, sizeof(uint32_t) * (size_t)(sh.n + 64) * s->WPR + 64);  // padded: whole-tile bulk copies
    if (e == cudaSuccess && s->missing && s->use_dmma && (D == 128 || D == 256)) {
      sh.nwT = 4 * ceil_div(sh.n, (int64_t)128);
      e = cudaMalloc(&sh.mbitsT, sizeof(uint32_t) * (size_t)D * sh.nwT);
      if (e == cudaSuccess) e = cudaMalloc(&sh.eamax, 256);
      if (e == cudaSuccess && D == 256 && s->M == 16) {
        e = cudaMalloc(&sh.gramB, masked_gram_tc_image_bytes());
        if (e == cudaSuccess) e = cudaMemset(sh.gramB, 0, masked_gram_tc_image_bytes());   // padding rows of the last chunk stay zero
        if (e == cudaSuccess) e = cudaMalloc(&sh.gramT, masked_gram_tc_table_bytes());
        if (e == cudaSuccess) e = cudaMemset(sh.gramT, 0, masked_gram_tc_table_bytes());
      }
    }
    if (e == cudaSuccess) e = cudaMalloc(&sh.parblock, sizeof(double) * parlen);
    if (e == cudaSuccess) e = cudaMemset(sh.parblock, 0, sizeof(double) * parlen);
    if (e == cudaSuccess) e = cudaMalloc(&sh.stats, sizeof(double) * (size_t)std::max<int64_t>(s->S, 3 * D));
    if (e == cudaSuccess) e = cudaMalloc(&sh.partials, sizeof(double) * (size_t)std::max<int64_t>(s->S, 3 * D) * sh.nparts);
    // per-feature scratch (8 D) + per-block W'^T W' partials of the masked update (ceil(D / 8) blocks x MP x MP)
    if (e == cudaSuccess) e = cudaMalloc(&sh.scratch, sizeof(double) * ((size_t)D * 8 + (size_t)ceil_div(D, 8) * MP * MP));
    if (e == cudaSuccess) e = cudaEventCreate(&sh.ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&sh.ev1);
    if (e != cudaSuccess) {
      for (ShardState& t : s->sh) free_shard(t);
      delete s;
      return fail(GPLVM_ERR_CUDA, "session_create: allocation failed: %s", cudaGetErrorString(e));
    }
    double* q = sh.parblock;
    Par& p = sh.par;
    p.W = q; q += (size_t)D * MP;
    p.A = q; q += (size_t)D * MP;
    p.Wprev = q; q += (size_t)D * MP;
    p.Minv = q; q += (size_t)MP * MP;
    p.c = q; q += MP;
    p.mu = q; q += D;
    p.nd = q; q += D;
    p.sy = q; q += D;
    p.syy = q; q += D;
    p.scal = q; q += S_COUNT;
    p.D = (int)D; p.M = (int)M; p.MP = MP; p.NV = s->NV; p.N = N_global;
  }
  if (s->use_dmma) {
    int rc = dense_dmma_prepare(s);
    if (rc != GPLVM_OK) {
      for (ShardState& t : s->sh) free_shard(t);
      delete s;
      return rc;
    }
  }
  session_setup_p2p(s);
  *out = s;
  return GPLVM_OK;
}

int gplvm_session_destroy(gplvm_ppca_session* s) {
  if (!s) return GPLVM_OK;
  std::lock_guard<std::recursive_mutex> lk(ctx().mu);
  for (ShardState& sh : s->sh) {
    cudaSetDevice(sh.dev);
    cudaStreamSynchronize(sh.st);
  }
  dense_dmma_release(s);
  if (s->p2p && ctx().multiproc && ctx().world > 1) {
    // a peer's last reduction may still be reading this rank's publication buffer: rendezvous before unmapping / freeing
    int token = 1;
    std::vector<int> all(ctx().world, 0);
    allgather_bytes(&token, all.data(), sizeof token);
  }
  session_release_p2p(s);
  for (ShardState& sh : s->sh) free_shard(sh);
  delete s;
  return GPLVM_OK;
}

int gplvm_session_load_host(gplvm_ppca_session* s, const double* Y, int64_t ld) {
  if (!s || !Y) return fail(GPLVM_ERR_ARG, "session_load_host: null argument");
  std::lock_guard<std::recursive_mutex> lk(ctx().mu);
  const int64_t D = s->D;
  const int64_t first = s->sh[0].n0;
  // One shard: this thread.  Several GPUs driven by one process (gplvm_set_devices): one host thread per shard, so that
  // every GPU's PCIe link is busy at once (and, for pageable caller memory, the driver's staging copies run in parallel).
  // Pageable caller memory (what a Haskell ForeignPtr or a std::vector is): cudaMemcpy2D would stage it through the driver's
  // single-threaded bounce copy at ~10 GB/s.  Instead `HT` host threads copy the strided rows of a chunk into one of two
  // pinned bounce buffers while the previous chunk is on the wire.
  const bool pageable = host_pageable(Y);
  const int HT = host_threads(s->sh.size());
  auto load_shard = [&](ShardState& sh, std::string& err) -> int {
    if (cudaSetDevice(sh.dev) != cudaSuccess) { err = "cudaSetDevice failed"; return GPLVM_ERR_CUDA; }
    if (pageable) {
      const int64_t cb = std::min<int64_t>(std::max<int64_t>(1024, ((int64_t)8 << 20) / D), sh.n);   // <= 64 MB per bounce buffer
      double* bounce[2] = {nullptr, nullptr};
      double* stage[2] = {nullptr, nullptr};
      cudaEvent_t done[2] = {nullptr, nullptr};
      cudaError_t e = cudaSuccess;
      for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        bounce[i] = bounce_get(sh.dev, i, sizeof(double) * (size_t)D * cb);   // cached across calls: pinning costs ~0.3 ms per MB
        if (!bounce[i]) e = cudaErrorMemoryAllocation;
        if (e == cudaSuccess) e = cudaMalloc(&stage[i], sizeof(double) * (size_t)D * cb);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming);
      }
      int64_t chunk = 0;
      for (int64_t c0 = 0; e == cudaSuccess && c0 < sh.n; c0 += cb, ++chunk) {
        const int b = (int)(chunk & 1);
        const int64_t cn = std::min(cb, sh.n - c0);
        if (chunk >= 2) e = cudaEventSynchronize(done[b]);   // the H2D copy that last used this bounce buffer has finished
        if (e != cudaSuccess) break;
        threaded_rows_copy(bounce[b], cb, Y + (sh.n0 - first) + c0, ld, D, sizeof(double) * (size_t)cn, HT);
        e = cudaMemcpyAsync(stage[b], bounce[b], sizeof(double) * (size_t)D * cb, cudaMemcpyHostToDevice, sh.st);
        if (e == cudaSuccess) e = cudaEventRecord(done[b], sh.st);
        if (e != cudaSuccess) break;
        dim3 grid((unsigned)ceil_div(cn, 32), (unsigned)ceil_div(D, 32));
        GP_LAUNCH(transpose_kernel, grid, 256, 0, sh.st, stage[b], cb, sh.X + c0 * sh.ldx, sh.ldx, D, cn, 0);
      }
      if (e == cudaSuccess) e = cudaStreamSynchronize(sh.st);
      for (int i = 0; i < 2; ++i) {
        if (stage[i]) cudaFree(stage[i]);
        if (done[i]) cudaEventDestroy(done[i]);
      }
      if (e != cudaSuccess) { err = std::string("H2D copy (pageable path) failed: ") + cudaGetErrorString(e); return GPLVM_ERR_CUDA; }
      return GPLVM_OK;
    }
    const int64_t ch = std::min(chunk_obs(D), sh.n);
    double* stage = nullptr;
    cudaError_t e = cudaMalloc(&stage, sizeof(double) * (size_t)D * ch);
    for (int64_t c0 = 0; e == cudaSuccess && c0 < sh.n; c0 += ch) {
      const int64_t cn = std::min(ch, sh.n - c0);
      const double* src = Y + (sh.n0 - first) + c0;  // (feature 0, observation n0 + c0)
      e = cudaMemcpy2DAsync(stage, sizeof(double) * ch, src, sizeof(double) * ld, sizeof(double) * cn, D, cudaMemcpyHostToDevice, sh.st);
      if (e != cudaSuccess) break;
      dim3 grid((unsigned)ceil_div(cn, 32), (unsigned)ceil_div(D, 32));
      GP_LAUNCH(transpose_kernel, grid, 256, 0, sh.st, stage, ch, sh.X + c0 * sh.ldx, sh.ldx, D, cn, 0);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(sh.st);
    if (stage) cudaFree(stage);
    if (e != cudaSuccess) { err = std::string("H2D copy failed: ") + cudaGetErrorString(e); return GPLVM_ERR_CUDA; }
    return GPLVM_OK;
  };
  std::vector<int> rcs(s->sh.size(), GPLVM_OK);
  std::vector<std::string> errs(s->sh.size());
  if (s->sh.size() == 1) {
    rcs[0] = load_shard(s->sh[0], errs[0]);
  } else {
    std::vector<std::thread> pool;
    for (size_t i = 0; i < s->sh.size(); ++i) pool.emplace_back([&, i]() { rcs[i] = load_shard(s->sh[i], errs[i]); });
    for (std::thread& t : pool) t.join();
    cudaSetDevice(s->sh[0].dev);
  }
  for (size_t i = 0; i < rcs.size(); ++i)
    if (rcs[i] != GPLVM_OK) return fail(rcs[i], "session_load_host: %s", errs[i].c_str());
  s->has_data = true;
  s->initialised = false;
  s->centred = false;
  return GPLVM_OK;
}

int gplvm_session_read_data(gplvm_ppca_session* s, double* Y, int64_t ld) {
  if (!s || !Y) return fail(GPLVM_ERR_ARG, "session_read_data: null argument");
  if (!s->has_data) return fail(GPLVM_ERR_STATE, "session_read_data: no data loaded");
  std::lock_guard<std::recursive_mutex> lk(ctx().mu);
  const int64_t D = s->D, first = s->sh[0].n0;
  // a dense session keeps Xc = X - mean resident; hand back X (equal to the input up to one rounding)
  if (s->centred) GP_TRY(dense_centre(s, +1.0));
  for (ShardState& sh : s->sh) {
    GP_TRY(session_d2h_rows(sh, D, sh.n, Y + (sh.n0 - first), ld, (int)s->sh.size(),
                            [&](int64_t c0, int64_t cn, double* stage, int64_t pitch) -> int {
                              dim3 grid((unsigned)ceil_div(cn, 32), (unsigned)ceil_div(D, 32));
                              GP_LAUNCH(transpose_kernel, grid, 256, 0, sh.st, sh.X + c0 * sh.ldx, sh.ldx, stage, pitch, cn, D, 1);
                              return GPLVM_OK;
                            }));
  }
  if (s->centred) GP_TRY(dense_centre(s, -1.0));
  return GPLVM_OK;
}

int gplvm_session_synth(gplvm_ppca_session* s, uint64_t seed, double nan_prob) {
  if (!s) return fail(GPLVM_ERR_ARG, "session_synth: null session");
  if (nan_prob > 0.0 && !s->missing) return fail(GPLVM_ERR_ARG, "session_synth: nan_prob > 0 needs a missing-data session");
  std::lock_guard<std::recursive_mutex> lk(ctx().mu);
  for (ShardState& sh : s->sh) {
    CUDA_TRY(cudaSetDevice(sh.dev));
    GP_LAUNCH(synth_kernel, (unsigned)ceil_div(sh.n, 16), 256, 0, sh.st, sh.X, sh.ldx, sh.n0, sh.n, (int)s->D, seed, nan_prob);
  }
  for (ShardState& sh : s->sh) {
    CUDA_TRY(cudaSetDevice(sh.dev));
    CUDA_TRY(cudaStreamSynchronize(sh.st));
  }
  s->has_data = true;
  s->initialised = false;
  s->centred = false;
  return GPLVM_OK;
}

int gplvm_session_init(gplvm_ppca_session* s, const double* W0, double sigma2_0) {
  if (!s || !W0) return fail(GPLVM_ERR_ARG, "session_init: null argument");
  if (!s->has_data) return fail(GPLVM_ERR_STATE, "session_init: load or synthesise data first");
  if (!(sigma2_0 > 0.0)) return fail(GPLVM_ERR_ARG, "session_init: sigma2_0 must be positive");
  std::lock_guard<std::recursive_mutex> lk(ctx().mu);
  for (ShardState& sh : s->sh) {
    CUDA_TRY(cudaSetDevice(sh.dev));
    double* w0d = sh.par.Wprev;  // staging for the D x M host matrix
    CUDA_TRY(cudaMemcpyAsync(w0d, W0, sizeof(double) * s->D * s->M, cudaMemcpyHostToDevice, sh.st));
    GP_LAUNCH(set_params_kernel, 8, 256, 0, sh.st, sh.par, w0d, sigma2_0);
  }
  s->steps = 0;
  GP_TRY(s->missing ? masked_init_constants(s) : dense_init_constants(s));
  GP_TRY(check_async(s, nullptr, true));
  s->initialised = true;
  return GPLVM_OK;
}

int gplvm_session_steps(gplvm_ppca_session* s, int64_t steps) {
  if (!s) return fail(GPLVM_ERR_ARG, "session_steps: null session");
  if (!s->initialised) return fail(GPLVM_ERR_STATE, "session_steps: call gplvm_session_init first");
  std::lock_guard<std::recursive_mutex> lk(ctx().mu);
  for (ShardState& sh : s->sh) {
    CUDA_TRY(cudaSetDevice(sh.dev));
    CUDA_TRY(cudaEventRecord(sh.ev0, sh.st));
  }
  for (int64_t i = 0; i < steps; ++i) GP_TRY(s->missing ? masked_enqueue_step(s) : dense_enqueue_step(s));
  for (ShardState& sh : s->sh) {
    CUDA_TRY(cudaSetDevice(sh.dev));
    CUDA_TRY(cudaEventRecord(sh.ev1, sh.st));
  }
  s->steps += steps;
  return GPLVM_OK;
}

int gplvm_session_sync(gplvm_ppca_session* s) {
  if (!s) return fail(GPLVM_ERR_ARG, "session_sync: null session");
  std::lock_guard<std::recursive_mutex> lk(ctx().mu);
  return check_async(s, nullptr, true);
}

int gplvm_session_run(gplvm_ppca_session* s, int stop_kind, int64_t k, double tol, int64_t* steps_done) {
  if (!s) return fail(GPLVM_ERR_ARG, "session_run: null session");
  if (!s->initialised) return fail(GPLVM_ERR_STATE, "session_run: call gplvm_session_init first");
  std::lock_guard<std::recursive_mutex> lk(ctx().mu);
  double scal[S_COUNT];
  if (stop_kind == GPLVM_STOP_ITERATIONS) {
    // Left k: continue while k > iteration, iteration starting at 0  =>  max(k,0)+1 steps (PPCA.hs:91)
    const int64_t total = (k > 0 ? k : 0) + 1;
    GP_TRY(gplvm_session_steps(s, total));
    GP_TRY(check_async(s, scal, true));
  } else if (stop_kind == GPLVM_STOP_TOLERANCE) {
    // Right tol: the update kernel raises DONE once max(|ds2|, max|dW|) <= tol; later steps are no-ops.
    for (ShardState& sh : s->sh) {
      CUDA_TRY(cudaSetDevice(sh.dev));
      GP_LAUNCH(set_scal_kernel, 1, 1, 0, sh.st, sh.par.scal, (int)S_TOL, tol, (int)S_USETOL, 1.0);
    }
    const int64_t cap = (int64_t)1 << 40;
    int64_t batch = 4;
    for (int64_t done = 0; done < cap;) {
      GP_TRY(gplvm_session_steps(s, batch));
      done += batch;
      GP_TRY(check_async(s, scal, true));
      if (scal[S_DONE] != 0.0) break;
      if (batch < 64) batch *= 2;
    }
    for (ShardState& sh : s->sh) {
      CUDA_TRY(cudaSetDevice(sh.dev));
      GP_LAUNCH(set_scal_kernel, 1, 1, 0, sh.st, sh.par.scal, (int)S_USETOL, 0.0, (int)S_DONE, 0.0);
    }
    GP_TRY(check_async(s, nullptr, true));
  } else {
    return fail(GPLVM_ERR_ARG, "session_run: unknown stop_kind %d", stop_kind);
  }
  s->steps = (int64_t)scal[S_STEPS];
  if (steps_done) *steps_done = (int64_t)scal[S_STEPS];
  return GPLVM_OK;
}

int gplvm_session_params(gplvm_ppca_session* s, double* W, double* sigma2, double* mu, double* dsigma2, double* max_dw,
                         int64_t* steps_done) {
  if (!s) return fail(GPLVM_ERR_ARG, "session_params: null session");
  if (!s->initialised) return fail(GPLVM_ERR_STATE, "session_params: not initialised");
  std::lock_guard<std::recursive_mutex> lk(ctx().mu);
  double scal[S_COUNT];
  GP_TRY(check_async(s, scal));
  ShardState& sh = s->sh[0];
  CUDA_TRY(cudaSetDevice(sh.dev));
  if (W) {
    std::vector<double> tmp((size_t)s->D * s->MP);
    CUDA_TRY(cudaMemcpy(tmp.data(), sh.par.W, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost));
    for (int64_t d = 0; d < s->D; ++d)
      for (int64_t a = 0; a < s->M; ++a) W[d * s->M + a] = tmp[d * s->MP + a];
  }
  if (mu) CUDA_TRY(cudaMemcpy(mu, sh.par.mu, sizeof(double) * s->D, cudaMemcpyDeviceToHost));
  if (sigma2) *sigma2 = scal[S_SIGMA2];
  if (dsigma2) *dsigma2 = scal[S_DVAR];
  if (max_dw) *max_dw = scal[S_MAXDW];
  if (steps_done) *steps_done = (int64_t)scal[S_STEPS];
  return GPLVM_OK;
}

int gplvm_session_loglik(gplvm_ppca_session* s, double* loglik_out) {
  if (!s || !loglik_out) return fail(GPLVM_ERR_ARG, "session_loglik: null argument");
  if (!s->initialised) return fail(GPLVM_ERR_STATE, "session_loglik: not initialised");
  std::lock_guard<std::recursive_mutex> lk(ctx().mu);
  return s->missing ? masked_loglik(s, loglik_out) : dense_loglik(s, loglik_out);
}

int gplvm_session_restored(gplvm_ppca_session* s, double* restored, int64_t ld) {
  if (!s || !restored) return fail(GPLVM_ERR_ARG, "session_restored: null argument");
  if (!s->missing) return fail(GPLVM_ERR_STATE, "session_restored: dense sessions have no restored matrix (Nothing, PPCA.hs:91)");
  if (!s->initialised || s->steps < 1) return fail(GPLVM_ERR_STATE, "session_restored: run at least one EM step first");
  std::lock_guard<std::recursive_mutex> lk(ctx().mu);
  return masked_restore(s, restored, ld);
}

int gplvm_session_collective(gplvm_ppca_session* s) {
  if (!s) return -1;
  if (ctx().world < 2) return 0;
  return s->p2p ? 2 : 1;
}

int gplvm_session_set_kernel_timing(gplvm_ppca_session* s, int on) {
  if (!s) return fail(GPLVM_ERR_ARG, "null session");
  std::lock_guard<std::recursive_mutex> lk(ctx().mu);
  for (ShardState& sh : s->sh) {
    CUDA_TRY(cudaSetDevice(sh.dev));
    if (on && !sh.evk0) {
      CUDA_TRY(cudaEventCreate(&sh.evk0));
      CUDA_TRY(cudaEventCreate(&sh.evk1));
      for (cudaEvent_t& e : sh.evseg) CUDA_TRY(cudaEventCreate(&e));
      sh.nmarks = 0;
    } else if (!on && sh.evk0) {
      cudaEventDestroy(sh.evk0); cudaEventDestroy(sh.evk1);
      sh.evk0 = sh.evk1 = nullptr;
      for (cudaEvent_t& e : sh.evseg) {
        if (e) cudaEventDestroy(e);
        e = nullptr;
      }
      sh.nmarks = 0;
    }
  }
  return GPLVM_OK;
}

int gplvm_session_last_steps_ms(gplvm_ppca_session* s, double* total_ms, double* kernel_ms) {
  if (!s) return fail(GPLVM_ERR_ARG, "null session");
  std::lock_guard<std::recursive_mutex> lk(ctx().mu);
  double tmax = 0.0, kmax = 0.0;
  for (ShardState& sh : s->sh) {
    CUDA_TRY(cudaSetDevice(sh.dev));
    CUDA_TRY(cudaEventSynchronize(sh.ev1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, sh.ev0, sh.ev1));
    if (ms > tmax) tmax = ms;
    if (sh.evk0) {
      // events bracket the statistics pass of the LAST step of the batch
      CUDA_TRY(cudaEventElapsedTime(&ms, sh.evk0, sh.evk1));
      if (ms > kmax) kmax = ms;
    }
  }
  if (total_ms) *total_ms = tmax;
  if (kernel_ms) *kernel_ms = kmax;
  return GPLVM_OK;
}

// Per-kernel device times of the LAST EM step enqueued while kernel timing was on (max over the local GPUs).
static const char* const kDenseSegments[] = {"statistics kernel (X.A, X^T.Ez, Ez^T.Ez)", "partial reduction + all-reduce", "replicated update"};
static const char* const kMaskedSegments[] = {"b pass (TMA + DMMA)", "per-observation solver", "complement sums (tcgen05 INT8)",
                                              "S2 pass (TMA + DMMA)", "all-reduce", "replicated update + finalize"};

int gplvm_session_last_segments_ms(gplvm_ppca_session* s, double* ms_out, int capacity, int* n_out) {
  if (!s || !n_out) return fail(GPLVM_ERR_ARG, "null argument");
  std::lock_guard<std::recursive_mutex> lk(ctx().mu);
  *n_out = 0;
  int nseg = 0;
  for (ShardState& sh : s->sh) {
    if (!sh.evk0 || sh.nmarks < 2) continue;
    CUDA_TRY(cudaSetDevice(sh.dev));
    CUDA_TRY(cudaEventSynchronize(sh.evseg[sh.nmarks - 1]));
    nseg = sh.nmarks - 1;
    for (int i = 0; i < nseg && i < capacity; ++i) {
      float ms = 0.f;
      CUDA_TRY(cudaEventElapsedTime(&ms, sh.evseg[i], sh.evseg[i + 1]));
      if (&sh == &s->sh[0] || ms > ms_out[i]) ms_out[i] = ms;
    }
  }
  *n_out = nseg < capacity ? nseg : capacity;
  return GPLVM_OK;
}

int gplvm_session_masked_solver_path(gplvm_ppca_session* s, int* path_out) {
  if (!s || !path_out) return fail(GPLVM_ERR_ARG, "null argument");
  std::lock_guard<std::recursive_mutex> lk(ctx().mu);
  *path_out = -1;
  if (!s->missing || s->steps < 1) return GPLVM_OK;
  if (!s->use_dmma) { *path_out = 0; return GPLVM_OK; }
  *path_out = 1;
  static const bool force_dmma = []() {
    const char* e = getenv("GPLVM_GRAM_DMMA");
    return e ? (atoi(e) != 0) : false;
  }();
  ShardState& sh = s->sh[0];
  if (sh.gramT && !force_dmma) {
    double gate = 0.0;
    CUDA_TRY(cudaSetDevice(sh.dev));
    CUDA_TRY(cudaStreamSynchronize(sh.st));
    CUDA_TRY(cudaMemcpy(&gate, sh.gramT + masked_gram_tc_gate_index(), sizeof gate, cudaMemcpyDeviceToHost));
    if (gate == 1.0) *path_out = 2;
  }
  return GPLVM_OK;
}

const char* gplvm_session_segment_name(gplvm_ppca_session* s, int i) {
  if (!s || i < 0) return "";
  if (s->missing) return i < 6 ? kMaskedSegments[i] : "";
  return i < 3 ? kDenseSegments[i] : "";
}

}  // extern "C"
