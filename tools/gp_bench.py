"""Time the fused kernel-build + Cholesky + LML path (BASELINE config 5: N = 32768, Q = 2, FP64)."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gplvmhaskell_b200 as g

def run(n, reps=2):
    rng = np.random.default_rng(2468)
    X = rng.uniform(0, 60.0, (n, 2))
    y = np.sin(X[:, 0]) + np.cos(X[:, 1]) + 0.1 * rng.standard_normal(n)
    il2 = np.full(2, 10.0)
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        _, lml, logdet = g.gp_chol_lml(X, y, il2, 1.0, 5e-5, want_factor=False)
        best = min(best, time.perf_counter() - t0)
    return dict(n=n, seconds=best, tflops=n ** 3 / 3 / best * 1e-12, lml=lml, logdet=logdet)

if __name__ == "__main__":
    g.gp_chol_lml(np.random.rand(256, 2), np.random.rand(256), np.full(2, 10.0), want_factor=False)
    for n in [int(a) for a in sys.argv[1:]] or [8192, 16384, 32768]:
        print(json.dumps(run(n)), flush=True)
