"""Single-process multi-GPU check: gplvm_set_devices(n) row-shards one session over n GPUs of this process
(ncclCommInitAll + one packed all-reduce per EM step).  Results must agree with the 1-GPU run."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gplvmhaskell_b200 as g
from gplvmhaskell_b200 import _lib

def run(ngpu, missing):
    lib = g.load_library()
    _lib.check(lib.gplvm_set_devices(ngpu))
    D, N, M = 256, 200_003, 16
    W0 = np.random.default_rng(1).standard_normal((D, M))
    with g.Session(D, N, M, missing=missing) as s:
        s.synth(77, 0.2 if missing else 0.0)
        s.init(W0, 1.0)
        s.steps(5)
        p = s.params()
        ll = s.loglik()
        R = s.restored()[:, ::997] if missing else None
    return p["W"], p["sigma2"], ll, R

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    out = {}
    for missing in (False, True):
        a = run(1, missing)
        b = run(n, missing)
        out["missing" if missing else "dense"] = dict(
            dW=float(np.max(np.abs(a[0] - b[0])) / np.max(np.abs(a[0]))), ds2=abs(a[1] - b[1]) / a[1],
            dll=abs(a[2] - b[2]) / abs(a[2]), dR=(float(np.max(np.abs(a[3] - b[3]))) if missing else None))
    print(json.dumps(out))
    ok = all(v["dW"] < 1e-11 and v["ds2"] < 1e-12 and v["dll"] < 1e-12 for v in out.values())
    sys.exit(0 if ok else 1)
