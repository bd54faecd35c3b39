"""Small instances of every fused kernel path, for `compute-sanitizer --tool memcheck python tools/sanitize_smoke.py`."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gplvmhaskell_b200 as g

rng = np.random.default_rng(0)
for d, n, m in ((256, 1000, 16), (128, 333, 16), (64, 100, 5), (7, 50, 2)):
    Y = rng.standard_normal((d, m)) @ rng.standard_normal((m, n)) + rng.standard_normal((d, n))
    W0 = rng.standard_normal((d, m))
    W, s2, ll, _ = g.em_steps_fast(Y, W0, 1.0, g.Left(2))
    print("dense", d, n, m, s2, ll)
    mask = rng.random((d, n)) < 0.2
    mask[0] = False
    Wm, s2m, llm, R = g.em_steps_missed(np.where(mask, np.nan, Y), W0, 1.0, g.Left(1))
    print("masked", d, n, m, s2m, llm, float(np.abs(R).max()))
X = rng.uniform(0, 15, (700, 2))
y = np.sin(X[:, 0])
L, lml, logdet = g.gp_chol_lml(X, y, np.full(2, 10.0), 1.0, 5e-5)
print("gp", lml, logdet)
mean, post, samp = g.gp_to_posterior_sample(np.linspace(-5, 5, 130), np.linspace(-4, 4, 40), np.sin(np.linspace(-4, 4, 40)),
                                            rng.standard_normal((130, 2)))
print("posterior", float(mean[0]), float(samp[0, 0]))
print("SANITIZE_SMOKE_DONE")
