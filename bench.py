#!/usr/bin/env python
"""bench.py -- PPCA EM iterations/s (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload all|dense|missing|dense1m|gp] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

A "step" is one EM iteration over the whole observation matrix (E-step statistics, all-reduce, M-step).

Headline (the JSON line's metric / value / e2e / roofline / cpu_baseline): synthetic dense PPCA, N_obs = 2^23, D = 256,
M = 16, FP64 (BASELINE.json "N=8M, D=256, M=16"), generated on the device by the library's counter-based generator
(SURVEY.md 8d).  With N > 1 GPUs the SAME 2^23 observations are row-sharded over the ranks (strong scaling); the
partial statistics (35 KB) are combined once per step by an all-reduce fused into the library's reduction kernel
over NVLink peer memory.

"secondary" (same JSON line; BASELINE.json configs 3, 4, 5), each with value / ms_per_step / roofline / clocks / e2e /
cpu_baseline:
  missing_8m  : NaN-masked PPCA, N = 2^23, 20 % NaN, row-sharded over the same ranks (every N)
  dense_1m    : dense PPCA N = 2^20, `Left 100` = 101 EM steps, 1 GPU (N = 1 runs only)
  gp_32768    : RBF kernel build fused with the Cholesky + GP log marginal likelihood, N = 32768, Q = 2 (N = 1 runs only)

value   : EM iterations/s with the observations resident in HBM; device time of exactly K steps (CUDA events on the
          library's stream), max over ranks, after W warm-up steps.  Inputs (17.2 GB at N=1) are far larger than the
          126 MB L2, so no flush is needed between steps.
parity_check : sigma2, log-likelihood and a projector checksum of W after EXACTLY W + K steps from the fixed init (read
          before any extra clock-sampling step), so the lines of the 1/2/4/8-GPU runs can be compared with each other;
          plus "oracle_prefix": a fresh session over the first 2^16 observations, sharded over the same ranks and
          through the same collective, against 2 steps of the CPU oracle on rank 0.
e2e     : the same metric through the reference-facing C ABI with HOST buffers.  N = 1: one gplvm_ppca_dense call (one
          H2D of Y in reference layout, K EM steps, the log-likelihood pass, D2H of W/sigma2/LL), wall clock around the
          call.  N > 1: one process per GPU, each rank drives the session API on its host shard (create, load_host,
          init, K steps, params, loglik); NOT a single one-shot call.
roofline: the dominant kernel (dense statistics pass), timed per launch with CUDA events on the launching stream, against
          the FP64 tensor (DMMA) peak measured in-run by the library's register-resident DMMA probe (MEASURED_PEAKS.json
          has no FP64 entry) and against MEASURED_PEAKS.json's HBM copy bandwidth.  `traffic` is the DRAM byte count of
          the committed ncu capture of that kernel (profiles/), scaled to this rank's observation count -- it is NOT
          measured in this run and `traffic_source` says so.
cpu_baseline : the reference cannot be built (no GHC); the CPU arm is the NumPy/OpenBLAS oracle that restates
          src/GPLVM/PPCA.hs:54-94 in the reference's literal operation order ("port") on all host cores, on a bounded
          sample (stated).  `--impl reference` runs the same oracle on the FULL 2^23 observations when the host has the
          memory (>= 48 GB available), else on a sample with "sampled": true in `config`.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D, M = 256, 16
N_FULL = 1 << 23
N_1M = 1 << 20
N_PREFIX = 1 << 16
SEED = 1234
W0_SEED = 4321
CPU_SAMPLE = 1 << 18
FLOP_OBS_MASKED = 52394.0   # complement form, 20 % NaN

# DRAM bytes per observation of the dominant kernels from the committed `ncu --set full` captures (dram__bytes_read.sum +
# dram__bytes_write.sum); NOT measured in the bench run itself.
TRAFFIC_DENSE = ((17.1803136e9 + 5.25e6) / (1 << 23), "profiles/r2_dense_dmma_ncu_raw.csv (ncu --set full capture of "
                 "dense_stats_dmma_kernel<256,0>, 2^23 observations: 17.180 GB read + 5.2 MB written); not measured in this run")
TRAFFIC_MASKED = (7.555e9 / (1 << 20), "ncu --set full captures of the four masked statistics kernels, per 2^20 observations: b pass 2.28 + "
                  "S2 pass 2.35 GB (profiles/r2_masked_kernels_ncu_raw.csv), tensor-core solver 1.42 GB (profiles/r2_solve_tc_ncu_raw.csv, "
                  "2^21 observations), complement sums 1.51 GB (profiles/r2_utc_ncu_raw.csv, 2^21 observations); not measured in this run")


def flops_per_obs_dense(d=D, m=M):
    return 4 * d * m + 2 * m * m   # X.A, X^T.Ez (2DM each) + Ez^T Ez (M^2 multiply-adds)


def bytes_per_obs(d=D):
    return 8 * d


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "MEASURED_PEAKS.json"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def mem_available_gb():
    try:
        with open("/proc/meminfo") as f:
            for ln in f:
                if ln.startswith("MemAvailable:"):
                    return int(ln.split()[1]) / (1 << 20)
    except OSError:
        pass
    return 0.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def blas_threads_all_cores():
    """Let the CPU arm use every host core (torchrun exports OMP_NUM_THREADS=1); returns the BLAS thread count in use."""
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=os.cpu_count() or 1)
        return max([p.get("num_threads", 1) for p in threadpoolctl.threadpool_info()] or [1])
    except Exception:
        return 1


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm (oracle/ is test infrastructure: executed here only as the reported baseline / the reference arm)
# ----------------------------------------------------------------------------------------------------------------------

def synth_host_dense(n: int, nthreads: int):
    """x = W_t z + mu_t + eps on the host (same model as the device generator, NumPy streams), D x n, centred in place
    exactly like oracle.center_dense (PPCA.hs:61).  Chunks are filled by a thread pool (NumPy releases the GIL)."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    rng = np.random.default_rng(SEED)
    Wt = rng.standard_normal((D, M))
    mu = rng.standard_normal((D, 1))
    Y = np.empty((D, n))
    chunk = 1 << 16
    seeds = np.random.SeedSequence(SEED).spawn((n + chunk - 1) // chunk)

    def fill(i):
        c0, c1 = i * chunk, min(n, (i + 1) * chunk)
        r = np.random.default_rng(seeds[i])
        blk = r.standard_normal((D, c1 - c0))
        blk += Wt @ r.standard_normal((M, c1 - c0))
        blk += mu
        Y[:, c0:c1] = blk

    with ThreadPoolExecutor(max_workers=max(1, nthreads)) as ex:
        list(ex.map(fill, range(len(seeds))))
    mean = Y.sum(axis=1) / float(n)          # Util.hs:159-164
    Y -= mean[:, None]                        # Util.hs:175-178 (same subtraction as center_dense, without the second copy)
    return Y


def cpu_port_rate(steps: int, warmup: int, sample: int, nthreads: int):
    """Oracle dense EM step (literal reference op order) on `sample` observations; returns (s/step, it/s scaled to N_FULL)."""
    import numpy as np
    from oracle import ppca as oppca
    Xc = synth_host_dense(sample, nthreads)
    W = np.random.default_rng(W0_SEED).standard_normal((D, M))
    s2 = 1.0
    for _ in range(warmup):
        W, s2, _, _ = oppca.dense_step(Xc, W, s2)
    t0 = time.perf_counter()
    for _ in range(steps):
        W, s2, _, _ = oppca.dense_step(Xc, W, s2)
    dt = (time.perf_counter() - t0) / max(1, steps)
    return dt, (1.0 / dt) * sample / N_FULL, s2


def cpu_port_rate_missing(steps: int, sample: int = 1 << 13):
    """Oracle masked EM step (sufficient-statistics restatement of PPCA.hs:106-147, validated against the literal
    per-observation closures in tests/test_oracle.py) on a bounded sample; returns (s/step, it/s at N_FULL)."""
    import numpy as np
    from oracle import ppca as oppca
    rng = np.random.default_rng(SEED)
    Y = rng.standard_normal((D, M)) @ rng.standard_normal((M, sample)) + rng.standard_normal((D, 1)) + rng.standard_normal((D, sample))
    mask = rng.random((D, sample)) < 0.2
    mask[0] = False
    Yn = np.where(mask, np.nan, Y)
    W = np.random.default_rng(W0_SEED).standard_normal((D, M))
    mu, s2 = np.zeros(D), 1.0
    t0 = time.perf_counter()
    for _ in range(steps):
        mu, W, s2, _, _, _ = oppca.missed_step_vectorised(Yn, mu, W, s2)
    dt = (time.perf_counter() - t0) / max(1, steps)
    return dt, (1.0 / dt) * sample / N_FULL


def run_reference(args):
    """Reference arm: the oracle's literal dense EM step on the host cores, FULL 2^23 observations when memory allows."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy  # noqa: F401  (loads the BLAS whose thread pool is resized below)
    cores = blas_threads_all_cores()
    avail = mem_available_gb()
    full = avail >= 48.0 and not args.nobs
    n = args.nobs or (N_FULL if full else CPU_SAMPLE)
    sampled = n != N_FULL
    warm = max(1, min(args.warmup, 1))            # one untimed step (first-touch of the temporaries); stated in config
    t_gen = time.perf_counter()
    dt, rate, s2 = cpu_port_rate(args.steps, warm, n, cores)
    t_all = time.perf_counter() - t_gen
    what = ("dense_step of oracle/ppca.py (NumPy/OpenBLAS restatement of PPCA.hs:63-82, literal op order: E[z] materialised, "
            "3 passes over X + |Xc|^2 per step) on %s; the GHC reference cannot be built (no GHC)"
            % ("ALL 2^23 observations, %d timed steps" % args.steps if not sampled else
               "%d of the 2^23 observations per step, scaled linearly to 2^23 (%s)"
               % (n, "--nobs override" if args.nobs else "host MemAvailable %.0f GB < 48 GB" % avail)))
    line = {"impl": "reference", "metric": "PPCA EM iterations/sec (N=8M,D=256,M=16,FP64)", "value": rate,
            "unit": "iterations/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 / rate, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "dense PPCA EM N=%d D=%d M=%d FP64 (CPU oracle port)" % (N_FULL, D, M), "sampled": sampled,
                       "sample_obs": n, "cpu_warmup_steps": warm, "host_mem_available_gb": round(avail, 1),
                       "wall_s_total": round(t_all, 1), "sigma2": s2},
            "cpu_baseline": {"value": rate, "unit": "iterations/s", "cores": cores, "kind": "port", "sample": what},
            "e2e": {"value": rate, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------

class Env:
    """torch.distributed plumbing + the library handle."""

    def __init__(self):
        import torch
        import ctypes as C
        import gplvmhaskell_b200 as g
        from gplvmhaskell_b200 import _lib
        self.torch, self.C, self.g, self._lib = torch, C, g, _lib
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.lib = g.load_library()
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if self.rank == 0:
                raw = (C.c_ubyte * 128)()
                _lib.check(self.lib.gplvm_dist_unique_id(raw))
                idbuf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
            dist.broadcast(idbuf, 0)
            raw = (C.c_ubyte * 128)(*idbuf.cpu().tolist())
            _lib.check(self.lib.gplvm_dist_init(self.rank, self.world, self.local_rank, raw))
            self.dist = dist
        else:
            _lib.check(self.lib.gplvm_dist_init(0, 1, self.local_rank, None))

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def block(self, n_obs: int):
        per = (n_obs + self.world - 1) // self.world
        begin = min(per * self.rank, n_obs)
        return begin, max(0, min(per, n_obs - begin))

    def gather_columns(self, Y):
        """Concatenate the ranks' D x count blocks on rank 0 (None elsewhere)."""
        if self.dist is None:
            return Y
        t = self.torch.from_numpy(Y).cuda()
        parts = [self.torch.empty_like(t) for _ in range(self.world)] if self.rank == 0 else None
        self.dist.gather(t, parts, dst=0)
        if self.rank != 0:
            return None
        import numpy as np
        return np.concatenate([p.cpu().numpy() for p in parts], axis=1)

    def finalize(self):
        if self.dist is not None:
            self.lib.gplvm_dist_finalize()
            self.dist.destroy_process_group()


def w_checksum(W):
    """Rotation / sign invariant summary of W: weighted diagonal of the projector W (W^T W)^-1 W^T and |W W^T|_F."""
    import numpy as np
    P = W @ np.linalg.solve(W.T @ W, W.T)
    return float(np.dot(np.arange(1, W.shape[0] + 1), np.diag(P))), float(np.linalg.norm(W @ W.T))


def oracle_prefix_check(env: Env, missing: bool):
    """A fresh session over the first n observations of the bench data (the generator is keyed by the global observation
    index), sharded over the same ranks and through the same collective as the timed run, against the CPU oracle."""
    import numpy as np
    g = env.g
    n = N_PREFIX if not missing else (1 << 13)
    steps = 2
    assert n % env.world == 0, "prefix check needs a world size dividing %d" % n
    per = n // env.world
    begin, count = per * env.rank, per
    W0 = np.random.default_rng(W0_SEED).standard_normal((D, M))
    with g.Session(D, n, M, missing=missing, obs_begin=begin, obs_count=count) as s:
        s.synth(SEED, 0.2 if missing else 0.0)
        Yfull = env.gather_columns(s.read_data())
        s.init(W0, 1.0)
        s.steps(steps)
        s.sync()
        p = s.params()
        ll = s.loglik()
        coll = env.lib.gplvm_session_collective(s.h)
    if env.rank != 0:
        return None
    from oracle import ppca as oppca
    if missing:
        mu, W, s2 = np.zeros(D), W0.copy(), 1.0
        for _ in range(steps):
            mu, W, s2, _, _, _ = oppca.missed_step_vectorised(Yfull, mu, W, s2)
        llr = oppca.missed_loglik_vectorised(Yfull, mu, W, s2)
        rel_mu = float(np.max(np.abs(p["mu"] - mu)) / np.max(np.abs(mu)))
    else:
        Xc = oppca.center_dense(Yfull)
        W, s2 = W0.copy(), 1.0
        for _ in range(steps):
            W, s2, _, _ = oppca.dense_step(Xc, W, s2)
        llr = oppca.dense_loglik(Xc, W, s2, literal_det=False)
        rel_mu = None
    out = {"n_obs": n, "steps": steps, "ranks": env.world, "collective": int(coll),
           "rel_W": float(np.max(np.abs(p["W"] - W)) / np.max(np.abs(W))), "rel_sigma2": abs(p["sigma2"] - s2) / s2,
           "rel_loglik": abs(ll - llr) / abs(llr), "rel_mu": rel_mu, "tol": 1e-9}
    out["ok"] = bool(out["rel_W"] < 1e-9 and out["rel_sigma2"] < 1e-9 and out["rel_loglik"] < 1e-9 and (rel_mu is None or rel_mu < 1e-9))
    return out


def run_ppca(env: Env, args, n_obs: int, missing: bool, steps: int, warmup: int, host_buf=None, want_e2e=True, want_cpu=True,
             left_k_run=False):
    """One PPCA workload: resident-data throughput, parity read-out, clocks, per-kernel times, roofline, e2e, CPU baseline."""
    import numpy as np
    g, lib, _lib, C, torch = env.g, env.lib, env._lib, env.C, env.torch
    begin, count = env.block(n_obs)
    W0 = np.random.default_rng(W0_SEED).standard_normal((D, M))
    sess = g.Session(D, n_obs, M, missing=missing, obs_begin=begin, obs_count=count)
    sess.synth(SEED, 0.2 if missing else 0.0)
    sess.init(W0, 1.0)
    collective = {0: "none (1 GPU)", 1: "one packed ncclAllReduce per step",
                  2: "all-reduce fused into the reduction kernel over NVLink peer memory (no NCCL call per step)"}.get(
        lib.gplvm_session_collective(sess.h), "?")
    sess.steps(warmup)
    sess.sync()
    env.barrier()
    sampler = ClockSampler(env.local_rank).start()
    launches0 = lib.gplvm_kernel_launches()
    env.barrier()
    if left_k_run:
        done = sess.run(g.Left(steps - 1))        # the reference's `Left k` driver: k + 1 steps, one blocking call
        assert done == warmup + steps, (done, warmup, steps)
    else:
        sess.steps(steps)
        sess.sync()
    env.barrier()
    launches = lib.gplvm_kernel_launches() - launches0
    total_ms, _ = sess.last_steps_ms()
    # parity read-out after EXACTLY warmup + steps EM steps (the log-likelihood pass does not change the state)
    p = sess.params()
    ll = sess.loglik()
    chk, fro = w_checksum(p["W"])
    parity = {"em_steps": int(p["steps"]), "sigma2": p["sigma2"], "loglik": ll, "proj_checksum": chk, "wwt_fro": fro,
              "mu_sum": float(np.sum(p["mu"]))}
    # nvidia-smi samples every 100 ms; when the timed region is shorter than a few samples keep the SAME load running
    # (untimed) until the clock record has at least 3 samples under load.  The ranks must agree on the number of extra
    # steps (every step carries a collective), so the decision is a MAX over ranks.
    t_extra = time.perf_counter()
    while env.max_over_ranks(1.0 if (len(sampler.lines) < 3 and time.perf_counter() - t_extra < 2.0) else 0.0) > 0.0:
        sess.steps(steps)
        sess.sync()
    clocks = sampler.stop()
    total_ms = env.max_over_ranks(total_ms)
    ms_per_step = total_ms / steps
    value = 1e3 / ms_per_step

    # per-kernel device times: CUDA events on the launching stream around every kernel of single steps
    sess.set_kernel_timing(True)
    segs = {}
    reps = 5
    for _ in range(reps):
        sess.steps(1)
        for name, ms in sess.last_segments_ms():
            segs[name] = segs.get(name, 0.0) + ms / reps
    sess.set_kernel_timing(False)
    segs = {k: env.max_over_ranks(v) for k, v in segs.items()}
    names = list(segs)
    stat_names = names[:4] if missing else names[:1]     # masked: b pass, solver, complement sums, S2 pass
    k_ms = sum(segs[nm] for nm in stat_names)

    peaks, peak_src = measured_peaks()
    dmma = C.c_double()
    _lib.check(lib.gplvm_probe_dmma_tflops(C.byref(dmma)))
    flop_obs = FLOP_OBS_MASKED if missing else float(flops_per_obs_dense())
    ach_tf = flop_obs * count / (k_ms * 1e-3) * 1e-12
    ach_gbs = bytes_per_obs() * count / (k_ms * 1e-3) * 1e-9
    tr_per_obs, tr_src = TRAFFIC_MASKED if missing else TRAFFIC_DENSE
    roofline = {"bound": "tensor", "achieved": ach_tf, "peak": dmma.value, "unit": "TFLOP/s", "frac": ach_tf / dmma.value,
                "traffic": tr_per_obs * count, "traffic_source": tr_src,
                "kernel": ("masked E-step + statistics kernels (b pass, per-observation solver with exact fixed-point Gram "
                           "matrices on tcgen05, exact INT8 complement sums on tcgen05, S2 pass) per rank; FP64-equivalent "
                           "algorithmic flops against the DMMA peak"
                           if missing else "dense statistics kernel (X.A, X^T.Ez, Ez^T.Ez) per rank"),
                "kernel_ms": k_ms, "segments_ms": segs,
                **({"masked_solver_path": {2: "tcgen05 fixed-point Gram (masked_gram_tc.cu)", 1: "FP64 DMMA Gram (masked_fused.cu)",
                                           0: "generic"}.get(sess.masked_solver_path(), "n/a")} if missing else {}),
                "peak_source": "in-run DMMA.8x8x4 register-chain probe (no FP64 entry in MEASURED_PEAKS.json); cublasDgemm "
                               "8192^3 measured 35.8 TFLOP/s (profiles/r1_fp64_peaks.log)",
                "hbm_achieved_gbs": ach_gbs, "hbm_peak_gbs": peaks.get("hbm_gbs"), "hbm_frac": ach_gbs / peaks.get("hbm_gbs"),
                "hbm_peak_source": peak_src, "flop_per_obs": flop_obs, "bytes_per_obs": bytes_per_obs()}

    # ---- e2e: host buffers through the C ABI -----------------------------------------------------------------------
    e2e = None
    if want_e2e and host_buf is not None:
        Yh = host_buf.numpy()
        assert Yh.shape == (D, count)
        _lib.check(lib.gplvm_session_read_data(sess.h, _lib.dptr(Yh), count))   # same synthetic data, now on the host
        sess.close()
        sess = None
        Wout = np.empty((D, M)); s2o, llo, st = C.c_double(), C.c_double(), C.c_int64()
        if env.world == 1:
            small = np.ascontiguousarray(Yh[:, :4096])
            fn = lib.gplvm_ppca_missing if missing else lib.gplvm_ppca_dense
            extra = (None,) if missing else ()
            _lib.check(fn(_lib.dptr(small), D, 4096, M, _lib.dptr(W0), 1.0, 0, 1, 0.0, _lib.dptr(Wout), C.byref(s2o), C.byref(llo),
                          *extra, C.byref(st)))
            env.barrier()
            t0 = time.perf_counter()
            if missing:   # restored matrix written over the (already uploaded) input buffer: one pinned 17 GB buffer
                _lib.check(lib.gplvm_ppca_missing(_lib.dptr(Yh), D, count, M, _lib.dptr(W0), 1.0, 0, steps - 1, 0.0, _lib.dptr(Wout),
                                                  C.byref(s2o), C.byref(llo), _lib.dptr(Yh), C.byref(st)))
            else:
                _lib.check(lib.gplvm_ppca_dense(_lib.dptr(Yh), D, count, M, _lib.dptr(W0), 1.0, 0, steps - 1, 0.0, _lib.dptr(Wout),
                                                C.byref(s2o), C.byref(llo), C.byref(st)))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            assert st.value == steps
            d2h = 8 * (D * M + 2) + (8 * D * count if missing else 0)
            what = ("one C-ABI call (gplvm_ppca_%s) on host buffers: H2D of Y (reference D x N layout, pinned) once + %d EM steps "
                    "+ log-likelihood pass + D2H of W, sigma2, LL%s" % ("missing" if missing else "dense", steps,
                                                                      " + restored matrix (D x N)" if missing else ""))
            e2e_parity = {"sigma2": s2o.value, "loglik": llo.value, "em_steps": int(st.value)}
        else:
            env.barrier()
            t0 = time.perf_counter()
            ph = {}
            s2 = g.Session(D, n_obs, M, missing=missing, obs_begin=begin, obs_count=count)
            ph["create"] = time.perf_counter() - t0
            s2.load_host(Yh)
            ph["load_host"] = time.perf_counter() - t0
            s2.init(W0, 1.0)
            ph["init"] = time.perf_counter() - t0
            s2.steps(steps)
            pp = s2.params()
            ph["steps+params"] = time.perf_counter() - t0
            llv = s2.loglik()
            ph["loglik"] = time.perf_counter() - t0
            d2h = 8 * (D * M + 2)
            if missing:
                _lib.check(lib.gplvm_session_restored(s2.h, _lib.dptr(Yh), count))
                d2h += 8 * D * count
                ph["restored"] = time.perf_counter() - t0
            env.barrier()
            dt = time.perf_counter() - t0
            s2.close()
            e2e_phases = {k: round(v, 4) for k, v in ph.items()}   # cumulative seconds on rank 0
            what = ("one process per GPU, each rank drives the C-ABI session API on its pinned host shard: create + load_host "
                    "(H2D, reference layout) + init + %d EM steps + params + log-likelihood%s; NOT a single one-shot call"
                    % (steps, " + restored matrix D2H" if missing else ""))
            e2e_parity = {"sigma2": pp["sigma2"], "loglik": llv, "em_steps": int(pp["steps"]), "cumulative_s_rank0": e2e_phases}
        dt = env.max_over_ranks(dt)
        e2e = {"value": steps / dt, "unit": "iterations/s", "h2d_bytes_per_step": 8 * D * n_obs / steps,
               "d2h_bytes_per_step": d2h * (env.world if missing else 1) / steps, "seconds": dt, "what": what, "result": e2e_parity}
    if sess is not None:
        sess.close()

    cpu_baseline = None
    if env.rank == 0 and env.world == 1 and want_cpu:
        ncores = blas_threads_all_cores()
        if missing:
            dt_step, rate = cpu_port_rate_missing(steps=8)
            cpu_baseline = {"value": rate, "unit": "iterations/s", "cores": ncores, "kind": "port",
                            "sample": "oracle/ppca.py missed_step_vectorised (NumPy restatement of PPCA.hs:106-147 in sufficient-"
                                      "statistics form; the literal closures are O(N D) inversions), 8 steps on 2^13 of the 2^23 "
                                      "observations (%.2f s/step), scaled linearly to 2^23" % dt_step}
        else:
            dt_step, rate, _ = cpu_port_rate(steps=40, warmup=1, sample=CPU_SAMPLE, nthreads=ncores)
            rate = rate * N_FULL / n_obs
            cpu_baseline = {"value": rate, "unit": "iterations/s", "cores": ncores, "kind": "port",
                            "sample": "oracle/ppca.py dense_step (NumPy/OpenBLAS restatement of PPCA.hs:63-82, literal op order), "
                                      "40 steps on 2^18 observations (%.3f s/step), scaled linearly to N=%d" % (dt_step, n_obs)}
    return {"value": value, "unit": "iterations/s", "ms_per_step": ms_per_step, "steps": steps, "warmup": warmup,
            "config": {"workload": "%s PPCA EM N=%d D=%d M=%d FP64%s" % ("missing" if missing else "dense", n_obs, D, M,
                                                                         " 20% NaN" if missing else ""),
                       "sharding": "observations row-sharded over %d rank(s)" % env.world, "collective": collective,
                       "l2": "inputs (%.1f GB per rank) exceed the 126 MB L2; no flush" % (8 * D * count / 1e9),
                       "init": "W0 ~ N(0,1) seed %d, sigma2_0 = 1" % W0_SEED,
                       "driver": "gplvm_session_run(Left %d)" % (steps - 1) if left_k_run else "gplvm_session_steps",
                       **({"arithmetic": "FP64 throughout, except two sums whose one factor is an exact 0/1 mask (per-observation Gram "
                                         "matrices, per-feature complement sums): every FP64 term is rounded once to a fixed-point grid "
                                         "(<= 2^-52 of its scale) and the sums are exact integers on the INT8 tensor cores; on-device gates "
                                         "hand badly scaled steps to the FP64 DMMA kernels (DESIGN.md 4.4)"} if missing else {})},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "parity_check": parity}


def run_gp(env: Env, args):
    """BASELINE config 5: RBF kernel build + Cholesky + GP log marginal likelihood, N = 32768, Q = 2, FP64.
    Replicas only: a step = one full build + factorisation + LML call on HOST inputs (X is 0.5 MB, so
    e2e == value)."""
    import numpy as np
    g, lib, _lib, C, torch = env.g, env.lib, env._lib, env.C, env.torch
    n = args.nobs or 32768
    rng = np.random.default_rng(2468)
    X = rng.uniform(0, 60.0, (n, 2))
    y = np.sin(X[:, 0]) + np.cos(X[:, 1]) + 0.1 * rng.standard_normal(n)
    il2 = np.full(2, 10.0)
    steps, warm = max(1, min(args.steps, 5)), 3
    for _ in range(warm):
        g.gp_chol_lml(X, y, il2, 1.0, 5e-5, want_factor=False)
    sampler = ClockSampler(env.local_rank).start()
    l0 = lib.gplvm_kernel_launches()
    torch.cuda.synchronize()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        _, lml, logdet = g.gp_chol_lml(X, y, il2, 1.0, 5e-5, want_factor=False)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    # The call is ~1000 launches on two streams (look-ahead: a high-priority panel chain beside the trailing updates), timed by
    # wall clock.  The host enqueues the whole factorisation in ~2.5 ms, yet single calls come out at 1.1-2.4x the steady time in
    # some runs: GPU-side scheduling of the two streams (the steady value repeats to 0.1 %).  value = median of the K calls; the
    # mean and every call are reported next to it.
    dt_mean = sum(times) / steps
    dt = sorted(times)[len(times) // 2]
    launches = lib.gplvm_kernel_launches() - l0
    clocks = sampler.stop()
    dmma = C.c_double()
    _lib.check(lib.gplvm_probe_dmma_tflops(C.byref(dmma)))
    tf = n ** 3 / 3.0 / dt * 1e-12
    cpu = None
    chk = None
    if env.rank == 0 and not args.no_cpu:
        from oracle import gp as ogp
        ns = 4096
        t1 = time.perf_counter()
        _, lml_s = ogp.gp_chol_lml(X[:ns], y[:ns], il2, 1.0, 5e-5)
        dc = time.perf_counter() - t1
        _, lml_g, _ = g.gp_chol_lml(X[:ns], y[:ns], il2, 1.0, 5e-5, want_factor=False)
        chk = {"n": ns, "rel_lml_vs_oracle": abs(lml_g - lml_s) / abs(lml_s), "tol": 1e-9, "ok": bool(abs(lml_g - lml_s) < 1e-9 * abs(lml_s))}
        cpu = {"value": 1.0 / (dc * (n / ns) ** 3), "unit": "factorisations/s", "cores": blas_threads_all_cores(), "kind": "port",
               "sample": "oracle/gp.py (NumPy kernel build + LAPACK dpotrf + triangular solve) at N=%d in %.2f s, scaled by (N/%d)^3" % (ns, dc, ns)}
    return {"metric": "GP kernel build + Cholesky + LML factorisations/sec (N=%d, Q=2, FP64)" % n, "value": 1.0 / dt,
            "unit": "factorisations/s", "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "ms_per_step_all": [t * 1e3 for t in times],
            "ms_per_step_mean": dt_mean * 1e3, "statistic": "median of the timed calls (single calls are occasionally 1.1-2.4x slower: two-stream scheduling on the GPU; mean and all calls alongside)",
            "scaling": "replicas only",
            "config": {"workload": "RBF/ARD kernel build fused with blocked Cholesky + GP LML, N=%d Q=2, jitter 5e-5" % n},
            "clocks": clocks, "e2e": {"value": 1.0 / dt, "unit": "factorisations/s", "h2d_bytes_per_step": 24 * n, "d2h_bytes_per_step": 16,
                                      "what": "gplvm_gp_chol_lml on host X, y (wall clock around the call)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": tf, "peak": dmma.value, "unit": "TFLOP/s", "frac": tf / dmma.value,
                         "traffic": None, "kernel": "whole call (trailing DMMA update dominates), N^3/3 flop",
                         "hbm_note": "the kernel matrix is generated inside the factorisation (never written and re-read): "
                                     "8 N^2 bytes of build traffic avoided"},
            "cpu_baseline": cpu, "parity_check": {"lml": lml, "logdet": logdet, "oracle_n4096": chk}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "dense", "missing", "dense1m", "gp"])
    ap.add_argument("--nobs", type=int, default=None, help="override the number of observations (debug)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device; libgplvm_b200 has no CPU fallback"}))
        return 2
    if int(os.environ.get("WORLD_SIZE", "1")) == 1 and args.gpus > 1:
        print(json.dumps({"error": "launch with torch.distributed.run for --gpus > 1"}))
        return 2
    env = Env()
    world, rank = env.world, env.rank
    n_obs = args.nobs or N_FULL
    _, count = env.block(n_obs)
    line = None
    secondary = {}
    gp_result = None
    if args.workload in ("all", "gp") and world == 1:
        gp_result = run_gp(env, args)
    host_buf = None
    if not args.no_e2e and args.workload != "gp":
        host_buf = torch.empty((D, count), dtype=torch.float64, pin_memory=True)   # shared by the dense and masked e2e legs
    if args.workload in ("all", "dense"):
        r = run_ppca(env, args, n_obs, False, args.steps, args.warmup, host_buf, want_e2e=not args.no_e2e, want_cpu=not args.no_cpu)
        r["parity_check"]["oracle_prefix"] = oracle_prefix_check(env, False)
        line = {"metric": "PPCA EM iterations/sec (N=8M,D=256,M=16,FP64)", "value": r["value"], "unit": "iterations/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": r["config"], "clocks": r["clocks"], "e2e": r["e2e"], "gpu_launches": r["gpu_launches"],
                "roofline": r["roofline"], "cpu_baseline": r["cpu_baseline"], "parity_check": r["parity_check"]}
    if args.workload in ("all", "missing"):
        ms = max(3, min(args.steps, 10))
        r = run_ppca(env, args, n_obs, True, ms, max(3, min(args.warmup, 3)), host_buf, want_e2e=not args.no_e2e, want_cpu=not args.no_cpu)
        r["parity_check"]["oracle_prefix"] = oracle_prefix_check(env, True)
        r["metric"] = "missing-data PPCA EM iterations/sec (N=8M,D=256,M=16,20% NaN,FP64)"
        r["scaling"] = "strong"
        secondary["missing_8m"] = r
    if args.workload in ("all", "dense1m") and world == 1:
        r = run_ppca(env, args, N_1M, False, 101, 3, None, want_e2e=False, want_cpu=not args.no_cpu, left_k_run=True)
        r["metric"] = "dense PPCA EM iterations/sec (N=1M,D=256,M=16,FP64; Left 100 = 101 steps, 1 GPU)"
        secondary["dense_1m"] = r
    if gp_result is not None:
        secondary["gp_32768"] = gp_result

    if rank == 0:
        if line is None:   # a single secondary workload was requested: promote it
            key = next(iter(secondary))
            line = dict(secondary.pop(key))
            line.setdefault("metric", key)
            line.update({"n_gpus": world, "higher_is_better": True, "vs_baseline": None, "dtype": "f64", "data": "synthetic"})
            line.setdefault("scaling", "strong")
        if secondary:
            line["secondary"] = secondary
        print(json.dumps(line), flush=True)
    env.finalize()
    return 0


if __name__ == "__main__":
    sys.exit(main())
