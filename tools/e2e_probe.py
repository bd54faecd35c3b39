import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import torch
import gplvmhaskell_b200 as g
from gplvmhaskell_b200 import _lib
lib = g.load_library()
D, M, N = 256, 16, 1 << 23
W0 = np.random.default_rng(4321).standard_normal((D, M))
host = torch.empty((D, N), dtype=torch.float64, pin_memory=True)
Y = host.numpy()
mode = sys.argv[1] if len(sys.argv) > 1 else "set"
if mode == "dist":
    _lib.check(lib.gplvm_dist_init(0, 1, 0, None))
else:
    _lib.check(lib.gplvm_set_devices(1))
with g.Session(D, N, M, missing=False) as s:
    s.synth(1234)
    _lib.check(lib.gplvm_session_read_data(s.h, _lib.dptr(Y), N))
for rep in range(3):
    t0 = time.perf_counter(); ph = {}
    s2 = g.Session(D, N, M, missing=False); ph["create"] = time.perf_counter() - t0
    s2.load_host(Y); ph["load"] = time.perf_counter() - t0
    s2.init(W0, 1.0); ph["init"] = time.perf_counter() - t0
    s2.steps(20); p = s2.params(); ph["steps"] = time.perf_counter() - t0
    ll = s2.loglik(); ph["ll"] = time.perf_counter() - t0
    s2.close(); ph["close"] = time.perf_counter() - t0
    print(mode, rep, {k: round(v, 4) for k, v in ph.items()}, flush=True)
Wout = np.empty((D, M)); s2v, llv, st = C.c_double(), C.c_double(), C.c_int64()
for rep in range(2):
    t0 = time.perf_counter()
    _lib.check(lib.gplvm_ppca_dense(_lib.dptr(Y), D, N, M, _lib.dptr(W0), 1.0, 0, 19, 0.0, _lib.dptr(Wout), C.byref(s2v), C.byref(llv), C.byref(st)))
    print(mode, "oneshot", rep, round(time.perf_counter() - t0, 4), flush=True)
