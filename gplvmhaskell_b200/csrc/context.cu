// context.cu -- error reporting, device shards and the NCCL binding (loaded with dlopen so that the
// library has no link-time NCCL dependency and shares the copy torch already mapped, if any).
#include <dlfcn.h>

#include "common.cuh"

namespace gplvm {

static thread_local std::string t_last_error;
std::atomic<int64_t> g_kernel_launches{0};

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  t_last_error = buf;
}

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  t_last_error = buf;
  return code;
}

const char* last_error_cstr() { return t_last_error.c_str(); }

Context& ctx() {
  static Context c;
  return c;
}

// ------------------------------------------------------------------------------------------------
// NCCL through dlopen
// ------------------------------------------------------------------------------------------------
namespace {
typedef struct { char internal[128]; } NcclUniqueId;
typedef int (*fn_GetUniqueId)(NcclUniqueId*);
typedef int (*fn_CommInitRank)(void**, int, NcclUniqueId, int);
typedef int (*fn_CommInitAll)(void**, int, const int*);
typedef int (*fn_CommDestroy)(void*);
typedef int (*fn_AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_AllGather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*fn_Group)(void);
typedef const char* (*fn_GetErrorString)(int);

struct NcclApi {
  void* handle = nullptr;
  fn_GetUniqueId GetUniqueId = nullptr;
  fn_CommInitRank CommInitRank = nullptr;
  fn_CommInitAll CommInitAll = nullptr;
  fn_CommDestroy CommDestroy = nullptr;
  fn_AllReduce AllReduce = nullptr;
  fn_AllGather AllGather = nullptr;
  fn_Group GroupStart = nullptr, GroupEnd = nullptr;
  fn_GetErrorString GetErrorString = nullptr;
} g_nccl;

constexpr int kNcclFloat64 = 8;  // ncclDouble
constexpr int kNcclSum = 0;      // ncclSum

int load_nccl() {
  if (g_nccl.handle) return GPLVM_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return fail(GPLVM_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define LOAD(sym)                                                                  \
  g_nccl.sym = (decltype(g_nccl.sym))dlsym(h, "nccl" #sym);                        \
  if (!g_nccl.sym) return fail(GPLVM_ERR_NCCL, "libnccl lacks symbol nccl" #sym);
  LOAD(GetUniqueId) LOAD(CommInitRank) LOAD(CommInitAll) LOAD(CommDestroy) LOAD(AllReduce) LOAD(AllGather)
  LOAD(GroupStart) LOAD(GroupEnd) LOAD(GetErrorString)
#undef LOAD
  g_nccl.handle = h;
  return GPLVM_OK;
}

#define NCCL_TRY(expr)                                                                         \
  do {                                                                                         \
    int _r = (expr);                                                                           \
    if (_r != 0) return fail(GPLVM_ERR_NCCL, "%s failed: %s", #expr, g_nccl.GetErrorString(_r)); \
  } while (0)
}  // namespace

int nccl_unique_id(void* id128) {
  GP_TRY(load_nccl());
  NcclUniqueId id;
  NCCL_TRY(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof id);
  return GPLVM_OK;
}

int nccl_init_rank(int rank, int world, int device, const void* id128) {
  Context& c = ctx();
  std::lock_guard<std::recursive_mutex> lk(c.mu);
  if (c.ready) GP_TRY(shutdown_context());
  if (world < 1 || rank < 0 || rank >= world) return fail(GPLVM_ERR_ARG, "bad rank/world %d/%d", rank, world);
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(GPLVM_ERR_CUDA, "device %d not present (%d visible)", device, ndev);
  Shard s;
  s.device = device;
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
  if (world > 1) {
    GP_TRY(load_nccl());
    if (!id128) return fail(GPLVM_ERR_ARG, "unique id required for world > 1");
    NcclUniqueId id;
    memcpy(&id, id128, sizeof id);
    NCCL_TRY(g_nccl.CommInitRank(&s.comm, world, id, rank));
  }
  c.shards.assign(1, s);
  c.multiproc = true;
  c.rank = rank;
  c.world = world;
  c.requested = 1;
  c.ready = true;
  return GPLVM_OK;
}

int ensure_context() {
  Context& c = ctx();
  std::lock_guard<std::recursive_mutex> lk(c.mu);
  if (c.ready) return GPLVM_OK;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(GPLVM_ERR_CUDA, "no usable CUDA device (%s); libgplvm_b200 has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  const int n = c.requested;
  if (n > ndev) return fail(GPLVM_ERR_CUDA, "gplvm_set_devices(%d) but only %d GPUs visible", n, ndev);
  c.shards.clear();
  for (int i = 0; i < n; ++i) {
    Shard s;
    s.device = i;
    CUDA_TRY(cudaSetDevice(i));
    CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    c.shards.push_back(s);
  }
  if (n > 1) {
    GP_TRY(load_nccl());
    std::vector<void*> comms(n);
    std::vector<int> devs(n);
    for (int i = 0; i < n; ++i) devs[i] = i;
    NCCL_TRY(g_nccl.CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; ++i) c.shards[i].comm = comms[i];
  }
  CUDA_TRY(cudaSetDevice(c.shards[0].device));
  c.multiproc = false;
  c.rank = 0;
  c.world = n;
  c.ready = true;
  return GPLVM_OK;
}

void gp_release_workspace();
void session_release_bounce();

int shutdown_context() {
  Context& c = ctx();
  std::lock_guard<std::recursive_mutex> lk(c.mu);
  gp_release_workspace();
  session_release_bounce();
  for (Shard& s : c.shards) {
    cudaSetDevice(s.device);
    if (s.stream) cudaStreamSynchronize(s.stream);
    if (s.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(s.comm);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  c.shards.clear();
  c.ready = false;
  c.multiproc = false;
  c.world = 1;
  c.rank = 0;
  return GPLVM_OK;
}

int allgather_bytes(const void* send, void* recv_host, size_t bytes) {
  Context& c = ctx();
  if (!c.multiproc || c.world == 1) {
    memcpy(recv_host, send, bytes);
    return GPLVM_OK;
  }
  Shard& sh = c.shards[0];
  CUDA_TRY(cudaSetDevice(sh.device));
  char *dsend = nullptr, *drecv = nullptr;
  CUDA_TRY(cudaMalloc(&dsend, bytes));
  CUDA_TRY(cudaMalloc(&drecv, bytes * c.world));
  CUDA_TRY(cudaMemcpyAsync(dsend, send, bytes, cudaMemcpyHostToDevice, sh.stream));
  int r = g_nccl.AllGather(dsend, drecv, bytes, 0 /* ncclInt8 / ncclChar */, sh.comm, sh.stream);
  cudaError_t e = cudaMemcpyAsync(recv_host, drecv, bytes * c.world, cudaMemcpyDeviceToHost, sh.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(sh.stream);
  cudaFree(dsend);
  cudaFree(drecv);
  if (r != 0) return fail(GPLVM_ERR_NCCL, "ncclAllGather failed: %s", g_nccl.GetErrorString(r));
  if (e != cudaSuccess) return fail(GPLVM_ERR_CUDA, "allgather copy failed: %s", cudaGetErrorString(e));
  return GPLVM_OK;
}

int allreduce_sum(const std::vector<double*>& bufs, size_t count) {
  Context& c = ctx();
  if (c.world == 1) return GPLVM_OK;
  if (bufs.size() != c.shards.size()) return fail(GPLVM_ERR_STATE, "allreduce: shard count mismatch");
  NCCL_TRY(g_nccl.GroupStart());
  for (size_t i = 0; i < bufs.size(); ++i) {
    int r = g_nccl.AllReduce(bufs[i], bufs[i], count, kNcclFloat64, kNcclSum, c.shards[i].comm, c.shards[i].stream);
    if (r != 0) {
      g_nccl.GroupEnd();
      return fail(GPLVM_ERR_NCCL, "ncclAllReduce failed: %s", g_nccl.GetErrorString(r));
    }
  }
  NCCL_TRY(g_nccl.GroupEnd());
  return GPLVM_OK;
}

}  // namespace gplvm

// ------------------------------------------------------------------------------------------------
// C ABI: context
// ------------------------------------------------------------------------------------------------
namespace gplvm { const char* last_error_cstr(); }

extern "C" {

const char* gplvm_last_error(void) { return gplvm::last_error_cstr(); }
const char* gplvm_version(void) { return "gplvm_b200 0.2 (sm_100a)"; }

int gplvm_set_devices(int n_gpus) {
  using namespace gplvm;
  Context& c = ctx();
  std::lock_guard<std::recursive_mutex> lk(c.mu);
  if (n_gpus < 1) return fail(GPLVM_ERR_ARG, "gplvm_set_devices: n_gpus must be >= 1");
  if (c.multiproc) return fail(GPLVM_ERR_STATE, "gplvm_set_devices: process is in multi-process mode (gplvm_dist_init)");
  if (c.ready && n_gpus != c.requested) GP_TRY(shutdown_context());
  c.requested = n_gpus;
  return ensure_context();
}

int gplvm_get_devices(void) { return (int)gplvm::ctx().shards.size(); }

int gplvm_dist_unique_id(void* id128) {
  if (!id128) return gplvm::fail(GPLVM_ERR_ARG, "null id buffer");
  return gplvm::nccl_unique_id(id128);
}
int gplvm_dist_init(int rank, int world, int device, const void* id128) {
  return gplvm::nccl_init_rank(rank, world, device, id128);
}
int gplvm_dist_finalize(void) { return gplvm::shutdown_context(); }
int gplvm_dist_rank(void) { return gplvm::ctx().rank; }
int gplvm_dist_world(void) { return gplvm::ctx().world; }
int64_t gplvm_kernel_launches(void) { return gplvm::g_kernel_launches.load(); }

}  // extern "C"
