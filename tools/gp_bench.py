#!/usr/bin/env python
"""GP factorisation timing helper: python tools/gp_bench.py [N] [reps]  (wall clock around gplvm_gp_chol_lml, no factor copy)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gplvmhaskell_b200 as g
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rng = np.random.default_rng(2468)
X = rng.uniform(0, 60.0, (n, 2))
y = np.sin(X[:, 0]) + np.cos(X[:, 1]) + 0.1 * rng.standard_normal(n)
il2 = np.full(2, 10.0)
for i in range(reps + 1):
    t0 = time.perf_counter()
    _, lml, _ = g.gp_chol_lml(X, y, il2, 1.0, 5e-5, want_factor=False)
    dt = time.perf_counter() - t0
    print("call %d: %.1f ms  %.2f TFLOP/s  lml %.10g" % (i, dt * 1e3, n ** 3 / 3 / dt * 1e-12, lml), flush=True)
