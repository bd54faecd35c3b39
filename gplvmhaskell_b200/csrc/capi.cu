// capi.cu -- one-shot C ABI entry points (host buffers in, host buffers out) built on the session API.
// They are what the reference's Haskell bodies call: emStepsFast / emStepsMissed of
// src/GPLVM/PPCA.hs:54-182 and their typed twins in src/GPLVM/TypeSafe/PPCA.hs:91-419.
#include <thread>

#include "ppca.cuh"

using namespace gplvm;

static int run_ppca(bool missing, const double* Y, int64_t D, int64_t N, int64_t M, const double* W0, double sigma2_0,
                    int stop_kind, int64_t k, double tol, double* W_out, double* sigma2_out, double* loglik_out,
                    double* restored_out, int64_t* steps_done) {
  if (!Y || !W0) return fail(GPLVM_ERR_ARG, "gplvm_ppca: null input");
  if (stop_kind != GPLVM_STOP_ITERATIONS && stop_kind != GPLVM_STOP_TOLERANCE)
    return fail(GPLVM_ERR_ARG, "gplvm_ppca: unknown stop_kind %d", stop_kind);
  // One-shot calls own the whole matrix.  After gplvm_dist_init(world > 1) every rank would build a session over all N
  // observations and the per-step all-reduce would add `world` copies of the statistics: refuse (use the session API
  // with this rank's block, or gplvm_set_devices in a single process).
  if (ctx().multiproc && ctx().world > 1)
    return fail(GPLVM_ERR_STATE, "gplvm_ppca*: one-shot calls are single-process (gplvm_set_devices); this process is rank %d of %d "
                "(gplvm_dist_init) -- use the session API with this rank's observation block", ctx().rank, ctx().world);
  gplvm_ppca_session* s = nullptr;
  GP_TRY(gplvm_session_create(D, N, M, missing ? 1 : 0, 0, N, &s));
  int rc = gplvm_session_load_host(s, Y, N);
  if (rc == GPLVM_OK) rc = gplvm_session_init(s, W0, sigma2_0);
  if (rc == GPLVM_OK) rc = gplvm_session_run(s, stop_kind, k, tol, steps_done);
  if (rc == GPLVM_OK && (W_out || sigma2_out)) rc = gplvm_session_params(s, W_out, sigma2_out, nullptr, nullptr, nullptr, nullptr);
  if (rc == GPLVM_OK && loglik_out) rc = gplvm_session_loglik(s, loglik_out);
  if (rc == GPLVM_OK && missing && restored_out) rc = gplvm_session_restored(s, restored_out, N);
  gplvm_session_destroy(s);
  return rc;
}

extern "C" {

int gplvm_ppca_dense(const double* Y, int64_t D, int64_t N, int64_t M, const double* W0, double sigma2_0, int stop_kind,
                     int64_t k, double tol, double* W_out, double* sigma2_out, double* loglik_out, int64_t* steps_done) {
  return run_ppca(false, Y, D, N, M, W0, sigma2_0, stop_kind, k, tol, W_out, sigma2_out, loglik_out, nullptr, steps_done);
}

int gplvm_ppca_missing(const double* Y, int64_t D, int64_t N, int64_t M, const double* W0, double sigma2_0, int stop_kind,
                       int64_t k, double tol, double* W_out, double* sigma2_out, double* loglik_out, double* restored_out,
                       int64_t* steps_done) {
  return run_ppca(true, Y, D, N, M, W0, sigma2_0, stop_kind, k, tol, W_out, sigma2_out, loglik_out, restored_out, steps_done);
}

int gplvm_ppca(const double* Y, int64_t D, int64_t N, int64_t M, const double* W0, double sigma2_0, int stop_kind, int64_t k,
               double tol, double* W_out, double* sigma2_out, double* loglik_out, double* restored_out, int* no_missed_out,
               int64_t* steps_done) {
  if (!Y || D < 1 || N < 1) return fail(GPLVM_ERR_ARG, "gplvm_ppca: bad input");
  // _noMissedData = not (isNaN (sumAll Y))  (PPCA.hs:44): a NaN anywhere poisons the sum.  The scan is host side
  // because Y is a host buffer at this boundary; it is a flag, not arithmetic.
  const int64_t total = D * N;
  std::atomic<bool> found{false};
  {
    const int nthreads = (int)std::max<int64_t>(1, std::min<int64_t>(16, total / (1 << 22)));
    std::vector<std::thread> pool;
    const int64_t per = (total + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t)
      pool.emplace_back([&, t]() {
        const int64_t b = t * per, e = std::min<int64_t>(total, b + per);
        for (int64_t i = b; i < e; i += 4096) {
          if (found.load(std::memory_order_relaxed)) return;
          const int64_t ee = std::min<int64_t>(e, i + 4096);
          bool any = false;
          for (int64_t q = i; q < ee; ++q) any |= (Y[q] != Y[q]);
          if (any) { found.store(true); return; }
        }
      });
    for (auto& th : pool) th.join();
  }
  const bool has_nan = found.load();
  if (no_missed_out) *no_missed_out = has_nan ? 0 : 1;
  return run_ppca(has_nan, Y, D, N, M, W0, sigma2_0, stop_kind, k, tol, W_out, sigma2_out, loglik_out, restored_out, steps_done);
}

}  // extern "C"
