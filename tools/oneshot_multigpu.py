"""One process, n GPUs: the mode a Haskell / CLI caller uses -- gplvm_set_devices(n) + ONE gplvm_ppca_dense call on a host
matrix in the reference layout (D x N).  Prints the end-to-end rate for pinned and for pageable caller memory.

    python tools/oneshot_multigpu.py [n_gpus] [log2 N] [steps]
"""
import ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gplvmhaskell_b200 as g
from gplvmhaskell_b200 import _lib

ngpu = int(sys.argv[1]) if len(sys.argv) > 1 else 2
N = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 23)
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
D, M = 256, 16
lib = g.load_library()
W0 = np.random.default_rng(4321).standard_normal((D, M))
# the bench data, generated on one GPU in slices and read back into a pinned host matrix
host = torch.empty((D, N), dtype=torch.float64, pin_memory=True)
Y = host.numpy()
_lib.check(lib.gplvm_set_devices(1))
with g.Session(D, N, M, missing=False) as s:   # 17 GB at N = 2^23: fits one B200
    s.synth(1234)
    _lib.check(lib.gplvm_session_read_data(s.h, _lib.dptr(Y), N))
out = {"n_gpus": ngpu, "N": N, "steps": steps}
for kind in ("pinned", "pageable"):
    Yk = Y if kind == "pinned" else np.array(Y, copy=True)
    res = {}
    for n in (1, ngpu):
        _lib.check(lib.gplvm_set_devices(n))
        Wout = np.empty((D, M)); s2, ll, st = C.c_double(), C.c_double(), C.c_int64()
        small = np.ascontiguousarray(Yk[:, :8192])
        _lib.check(lib.gplvm_ppca_dense(_lib.dptr(small), D, 8192, M, _lib.dptr(W0), 1.0, 0, 1, 0.0, _lib.dptr(Wout), C.byref(s2), C.byref(ll), C.byref(st)))
        t0 = time.perf_counter()
        _lib.check(lib.gplvm_ppca_dense(_lib.dptr(Yk), D, N, M, _lib.dptr(W0), 1.0, 0, steps - 1, 0.0, _lib.dptr(Wout), C.byref(s2), C.byref(ll), C.byref(st)))
        dt = time.perf_counter() - t0
        res[n] = {"seconds": round(dt, 4), "it_per_s": round(steps / dt, 2), "sigma2": s2.value, "loglik": ll.value}
    out[kind] = res
    out[kind + "_agree"] = abs(res[1]["sigma2"] - res[ngpu]["sigma2"]) / res[1]["sigma2"] < 1e-12 and abs(res[1]["loglik"] - res[ngpu]["loglik"]) / abs(res[1]["loglik"]) < 1e-12
    del Yk
_lib.check(lib.gplvm_set_devices(1))
print(json.dumps(out))
