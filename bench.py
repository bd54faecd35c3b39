#!/usr/bin/env python
"""bench.py -- PPCA EM iterations/s (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload dense|missing|gp] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

A "step" is one EM iteration over the whole observation matrix (E-step statistics, all-reduce, M-step).
Workload at every N: synthetic dense PPCA, N_obs = 2^23, D = 256, M = 16, FP64 (BASELINE.json config
"N=8M, D=256, M=16"), generated on the device by the library's counter-based generator;
with N > 1 GPUs the SAME 2^23 observations are row-sharded over the ranks (strong scaling); the partial statistics
(35 KB dense) are combined once per step, by an all-reduce fused into the library's reduction kernel over NVLink peer
memory (dense) or one packed ncclAllReduce (masked).

value   : EM iterations/s with the observations resident in HBM; device time of exactly K steps (CUDA events
          on the library's stream), max over ranks, after W warm-up steps.  Inputs (17.2 GB at N=1) are far
          larger than the 126 MB L2, so no flush is needed between steps.
e2e     : the same metric through the reference-facing C-ABI call with HOST buffers: gplvm_ppca_dense
          (one H2D of Y in reference layout, K EM steps, the log-likelihood pass, D2H of W/sigma2/LL), wall
          clock around the call.  N > 1: the session API per rank on that rank's host shard.
roofline: the dominant kernel (dense statistics pass) against the FP64 tensor (DMMA) peak measured in-run by
          the library's register-resident DMMA probe (MEASURED_PEAKS.json has no FP64 entry) and against
          MEASURED_PEAKS.json's HBM copy bandwidth.
cpu_baseline / --impl reference: the reference cannot be built (no GHC); the CPU arm is the NumPy/OpenBLAS
          oracle that restates src/GPLVM/PPCA.hs:54-94 in the reference's literal operation order ("port"),
          on all host cores, each step on a bounded sample of 2^18 observations, scaled linearly to 2^23.
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
SEED = 1234
CPU_SAMPLE = 1 << 18


def flops_per_obs_dense(d=D, m=M):
    return 4 * d * m + 2 * m * m   # X.A, X^T.Ez (2DM each) + Ez^T Ez (M^2 multiply-adds)


def bytes_per_obs(d=D):
    return 8 * d


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


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


def cpu_port_rate(steps: int, warmup: int, sample: int = CPU_SAMPLE):
    """Oracle dense EM step (literal reference op order) on a bounded sample; returns (it/s at N_FULL, info)."""
    import numpy as np
    from oracle import ppca as oppca
    rng = np.random.default_rng(SEED)
    Wt = rng.standard_normal((D, M))
    Y = Wt @ rng.standard_normal((M, sample)) + rng.standard_normal((D, 1)) + rng.standard_normal((D, sample))
    Xc = oppca.center_dense(Y)
    W = np.random.default_rng(4321).standard_normal((D, M))
    s2 = 1.0
    for _ in range(warmup):
        W, s2, _, _ = oppca.dense_step(Xc, W, s2)
    t0 = time.perf_counter()
    for _ in range(steps):
        W, s2, _, _ = oppca.dense_step(Xc, W, s2)
    dt = (time.perf_counter() - t0) / max(1, steps)
    rate_sample = 1.0 / dt
    rate_full = rate_sample * sample / N_FULL
    return rate_full, dt


def cpu_port_rate_missing(steps: int, sample: int = 1 << 13):
    """Oracle masked EM step (sufficient-statistics restatement of PPCA.hs:106-147, validated against the literal
    per-observation closures in tests/test_oracle.py) on a bounded sample; returns (it/s at N_FULL, s/step)."""
    import numpy as np
    from oracle import ppca as oppca
    rng = np.random.default_rng(SEED)
    Y = rng.standard_normal((D, M)) @ rng.standard_normal((M, sample)) + rng.standard_normal((D, 1)) + rng.standard_normal((D, sample))
    mask = rng.random((D, sample)) < 0.2
    mask[0] = False
    Yn = np.where(mask, np.nan, Y)
    W = np.random.default_rng(4321).standard_normal((D, M))
    mu, s2 = np.zeros(D), 1.0
    t0 = time.perf_counter()
    for _ in range(steps):
        mu, W, s2, _, _, _ = oppca.missed_step_vectorised(Yn, mu, W, s2)
    dt = (time.perf_counter() - t0) / max(1, steps)
    return (1.0 / dt) * sample / N_FULL, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy  # noqa: F401  (loads the BLAS whose thread pool is resized below)
    cores = blas_threads_all_cores()
    rate, dt = cpu_port_rate(args.steps, max(1, min(args.warmup, 2)))
    sample = ("dense_step of oracle/ppca.py (NumPy/OpenBLAS restatement of PPCA.hs:63-82, literal op order) on 2^18 "
              "of the 2^23 observations per step, scaled linearly to 2^23; the GHC reference cannot be built (no GHC)")
    line = {"impl": "reference", "metric": "PPCA EM iterations/sec (N=8M,D=256,M=16,FP64)", "value": rate,
            "unit": "iterations/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 / rate, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": "dense PPCA EM N=2^23 D=256 M=16 FP64 (CPU oracle, sampled)"},
            "cpu_baseline": {"value": rate, "unit": "iterations/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def run_gp(args, g, lib, rank, world):
    """BASELINE config 5: RBF kernel build + Cholesky + GP log marginal likelihood, N = 32768, Q = 2, FP64.
    Replicas only: every rank factorises the same problem; a step = one full build+factor+LML call
    on HOST inputs (X is 0.5 MB, so e2e == value)."""
    import ctypes as C
    import numpy as np
    import torch
    from gplvmhaskell_b200 import _lib
    n = args.nobs or 32768
    rng = np.random.default_rng(2468)
    X = rng.uniform(0, 60.0, (n, 2))
    y = np.sin(X[:, 0]) + np.cos(X[:, 1]) + 0.1 * rng.standard_normal(n)
    il2 = np.full(2, 10.0)
    steps, warm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 3))
    for _ in range(warm):
        g.gp_chol_lml(X, y, il2, 1.0, 5e-5, want_factor=False)
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    l0 = lib.gplvm_kernel_launches()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        _, lml, logdet = g.gp_chol_lml(X, y, il2, 1.0, 5e-5, want_factor=False)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    launches = lib.gplvm_kernel_launches() - l0
    clocks = sampler.stop()
    dmma = C.c_double()
    _lib.check(lib.gplvm_probe_dmma_tflops(C.byref(dmma)))
    tf = n ** 3 / 3.0 / dt * 1e-12
    cpu = None
    if rank == 0 and not args.no_cpu:
        from oracle import gp as ogp
        ns = 4096
        t1 = time.perf_counter()
        ogp.gp_chol_lml(X[:ns], y[:ns], il2, 1.0, 5e-5)
        dc = time.perf_counter() - t1
        cpu = {"value": 1.0 / (dc * (n / ns) ** 3), "unit": "factorisations/s", "cores": os.cpu_count(), "kind": "port",
               "sample": "oracle/gp.py (NumPy kernel build + LAPACK dpotrf + triangular solve) at N=%d in %.2f s, scaled by (N/%d)^3" % (ns, dc, ns)}
    if rank == 0:
        print(json.dumps({"metric": "GP kernel build + Cholesky + LML factorisations/sec (N=%d, Q=2, FP64)" % n, "value": 1.0 / dt,
                          "unit": "factorisations/s", "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3,
                          "higher_is_better": True, "scaling": "replicas only", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": "RBF/ARD kernel build fused with blocked Cholesky + GP LML, N=%d Q=2" % n},
                          "clocks": clocks, "e2e": {"value": 1.0 / dt, "unit": "factorisations/s", "h2d_bytes_per_step": 24 * n,
                                                    "d2h_bytes_per_step": 64}, "gpu_launches": int(launches),
                          "roofline": {"bound": "tensor", "achieved": tf, "peak": dmma.value, "unit": "TFLOP/s", "frac": tf / dmma.value,
                                       "traffic": None, "kernel": "whole call (trailing DMMA update dominates), N^3/3 flop"},
                          "cpu_baseline": cpu, "lml": lml, "logdet": logdet}), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dense", choices=["dense", "missing", "gp"])
    ap.add_argument("--nobs", type=int, default=None, help="override the number of observations (debug)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import ctypes as C
    import gplvmhaskell_b200 as g
    from gplvmhaskell_b200 import _lib

    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device; libgplvm_b200 has no CPU fallback"}))
        return 2
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        print(json.dumps({"error": "launch with torch.distributed.run for --gpus > 1"}))
        return 2
    torch.cuda.set_device(local_rank)
    lib = g.load_library()
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = (C.c_ubyte * 128)()
            _lib.check(lib.gplvm_dist_unique_id(raw))
            idbuf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
        dist.broadcast(idbuf, 0)
        raw = (C.c_ubyte * 128)(*idbuf.cpu().tolist())
        _lib.check(lib.gplvm_dist_init(rank, world, local_rank, raw))
    else:
        _lib.check(lib.gplvm_dist_init(0, 1, local_rank, None))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.workload == "gp":
        return run_gp(args, g, lib, rank, world)
    missing = args.workload == "missing"
    n_obs = args.nobs or N_FULL
    per = (n_obs + world - 1) // world
    begin = per * rank
    count = min(per, n_obs - begin)
    W0 = np.random.default_rng(4321).standard_normal((D, M))

    sess = g.Session(D, n_obs, M, missing=missing, obs_begin=begin, obs_count=count)
    sess.synth(SEED, 0.2 if missing else 0.0)
    sess.init(W0, 1.0)
    collective = {0: "none (1 GPU)", 1: "one packed ncclAllReduce per step",
                  2: "all-reduce fused into the reduction kernel over NVLink peer memory (no NCCL call per step)"}.get(
        lib.gplvm_session_collective(sess.h), "?")
    sess.steps(args.warmup)
    sess.sync()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.gplvm_kernel_launches()
    barrier()
    sess.steps(args.steps)
    sess.sync()
    barrier()
    launches = lib.gplvm_kernel_launches() - launches0
    total_ms, _ = sess.last_steps_ms()
    # nvidia-smi samples every 100 ms; when the timed region is shorter than a few samples keep the SAME load running
    # (untimed) until the clock record has at least 3 samples under load
    # (the ranks must agree on the number of extra steps -- every step carries a collective -- so the decision is a MAX
    # over ranks; the extra steps are reported as part of "em_steps_total", which is what sigma2 / loglik correspond to)
    t_extra = time.perf_counter()
    extra_steps = 0
    while max_over_ranks(1.0 if (len(sampler.lines) < 3 and time.perf_counter() - t_extra < 2.0) else 0.0) > 0.0:
        sess.steps(args.steps)
        sess.sync()
        extra_steps += args.steps
    clocks = sampler.stop()
    total_ms = max_over_ranks(total_ms)
    ms_per_step = total_ms / args.steps
    value = 1e3 / ms_per_step

    # dominant kernel: statistics pass, timed per launch with CUDA events on the launching stream
    sess.set_kernel_timing(True)
    kms = []
    for _ in range(min(args.steps, 10)):
        sess.steps(1)
        kms.append(sess.last_steps_ms()[1])
    sess.set_kernel_timing(False)
    k_ms = max_over_ranks(sum(kms) / len(kms))
    p = sess.params()
    ll = sess.loglik()

    peaks, peak_kind = measured_peaks()
    dmma = C.c_double()
    _lib.check(lib.gplvm_probe_dmma_tflops(C.byref(dmma)))
    if missing:
        # ~52 394 flop / observation (complement form, 20 % NaN); FP64-compute bound
        flop_obs = 52394.0
    else:
        flop_obs = float(flops_per_obs_dense())
    ach_tf = flop_obs * count / (k_ms * 1e-3) * 1e-12
    ach_gbs = bytes_per_obs() * count / (k_ms * 1e-3) * 1e-9
    # DRAM traffic per launch from the committed `ncu --set full` captures (profiles/r1_dense_dmma_ncu_raw.csv: 17.180 GB
    # read + 4.4 MB written for 2^23 observations = the algorithmic 2048 B/obs; profiles/r1_masked_kernels_ncu_raw.csv:
    # 7.59 GB over the four masked kernels for 2^20 observations), scaled to this rank's observation count
    traffic = (7.59e9 / (1 << 20) if missing else (17.180e9 + 4.4e6) / (1 << 23)) * count
    roofline = {"bound": "tensor", "achieved": ach_tf, "peak": dmma.value, "unit": "TFLOP/s", "frac": ach_tf / dmma.value,
                "traffic": traffic, "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum (profiles/), per launch", "kernel": ("masked E-step + statistics (TMA/DMMA b pass, register solver, exact INT8 fixed-point complement sums (tcgen05 kind::i8), DMMA S2 pass) per rank; FP64-equivalent algorithmic flops against the DMMA peak"
                           if missing else "dense statistics pass (X.A, X^T.Ez, Ez^T.Ez) per rank"),
                "kernel_ms": k_ms, "peak_source": "in-run DMMA.8x8x4 register-chain probe (no FP64 entry in MEASURED_PEAKS.json); "
                "cublasDgemm 8192^3 measured 35.8 TFLOP/s (profiles/r1_fp64_peaks.log)",
                "hbm_achieved_gbs": ach_gbs, "hbm_peak_gbs": peaks.get("hbm_gbs"), "hbm_frac": ach_gbs / peaks.get("hbm_gbs"),
                "hbm_peak_source": peak_kind, "flop_per_obs": flop_obs, "bytes_per_obs": bytes_per_obs()}

    # ---- e2e: host buffers through the C ABI -----------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        host = torch.empty((D, count), dtype=torch.float64, pin_memory=True)
        Yh = host.numpy()
        _lib.check(lib.gplvm_session_read_data(sess.h, _lib.dptr(Yh), count))   # same synthetic data, now on the host
        sess.close()
        Wout = np.empty((D, M)); s2o, llo, st = C.c_double(), C.c_double(), C.c_int64()
        if world == 1 and not missing:
            small = np.ascontiguousarray(Yh[:, :4096])
            _lib.check(lib.gplvm_ppca_dense(_lib.dptr(small), D, 4096, M, _lib.dptr(W0), 1.0, 0, 1, 0.0, _lib.dptr(Wout),
                                            C.byref(s2o), C.byref(llo), C.byref(st)))
            barrier()
            t0 = time.perf_counter()
            _lib.check(lib.gplvm_ppca_dense(_lib.dptr(Yh), D, count, M, _lib.dptr(W0), 1.0, 0, args.steps - 1, 0.0,
                                            _lib.dptr(Wout), C.byref(s2o), C.byref(llo), C.byref(st)))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            assert st.value == args.steps
            d2h = 8 * (D * M + 2)
        else:
            R = torch.empty((D, count), dtype=torch.float64, pin_memory=True) if missing else None
            barrier()
            t0 = time.perf_counter()
            s2 = g.Session(D, n_obs, M, missing=missing, obs_begin=begin, obs_count=count)
            s2.load_host(Yh)
            s2.init(W0, 1.0)
            s2.steps(args.steps)
            pp = s2.params()
            llv = s2.loglik()
            d2h = 8 * (D * M + 2)
            if missing:
                _lib.check(lib.gplvm_session_restored(s2.h, _lib.dptr(R.numpy()), count))
                d2h += 8 * D * count
            barrier()
            dt = time.perf_counter() - t0
            s2.close()
        dt = max_over_ranks(dt)
        e2e = {"value": args.steps / dt, "unit": "iterations/s", "h2d_bytes_per_step": 8 * D * count * world / args.steps,
               "d2h_bytes_per_step": d2h / args.steps, "seconds": dt,
               "what": "one C-ABI call on host buffers: H2D of Y (reference D x N layout, pinned) once + %d EM steps + "
                       "log-likelihood pass + D2H of W, sigma2, LL%s" % (args.steps, " + restored matrix" if missing else "")}
    else:
        sess.close()

    cpu_baseline = None
    ncores = blas_threads_all_cores() if (rank == 0 and world == 1 and not args.no_cpu) else 1
    if rank == 0 and world == 1 and not args.no_cpu and missing:
        rate, dt_step = cpu_port_rate_missing(steps=8)
        cpu_baseline = {"value": rate, "unit": "iterations/s", "cores": ncores, "kind": "port",
                        "sample": "oracle/ppca.py missed_step_vectorised (NumPy restatement of PPCA.hs:106-147 in sufficient-"
                                  "statistics form; the literal closures are O(N D) inversions), 8 steps on 2^13 of the 2^23 "
                                  "observations (%.2f s/step), scaled linearly to 2^23" % dt_step}
    elif rank == 0 and world == 1 and not args.no_cpu:
        rate, dt_step = cpu_port_rate(steps=40, warmup=1)
        cpu_baseline = {"value": rate, "unit": "iterations/s", "cores": ncores, "kind": "port",
                        "sample": "oracle/ppca.py dense_step (NumPy/OpenBLAS restatement of PPCA.hs:63-82, literal op order), "
                                  "40 steps on 2^18 of the 2^23 observations (%.2f s/step), scaled linearly to 2^23" % dt_step}

    if rank == 0:
        line = {"metric": "PPCA EM iterations/sec (N=8M,D=256,M=16,FP64)", "value": value, "unit": "iterations/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "%s PPCA EM N=%d D=%d M=%d FP64%s" % (args.workload, n_obs, D, M,
                                                                              " 20% NaN" if missing else ""),
                           "sharding": "observations row-sharded over %d rank(s)" % world, "collective": collective,
                           "l2": "inputs (%.1f GB per rank) exceed the 126 MB L2; no flush" % (8 * D * count / 1e9),
                           "init": "W0 ~ N(0,1) seed 4321, sigma2_0 = 1"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu_baseline, "sigma2": p["sigma2"], "loglik": ll,
                "em_steps_total": args.warmup + args.steps + extra_steps + min(args.steps, 10)}
        print(json.dumps(line), flush=True)
    if dist is not None:
        lib.gplvm_dist_finalize()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
