"""A/B of the two per-observation solvers of the masked E-step (D = 256, M = 16): tcgen05 fixed-point Gram (default) against
the FP64 DMMA Gram (GPLVM_GRAM_DMMA=1).  Runs each in a child process (the switch is read once), prints segment times and the
difference of the parameters after 3 steps.

    python tools/masked_solver_ab.py [log2 N]
"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import json, sys, numpy as np
sys.path.insert(0, %r)
import gplvmhaskell_b200 as g
D, M, N = 256, 16, 1 << int(sys.argv[1])
W0 = np.random.default_rng(4321).standard_normal((D, M))
with g.Session(D, N, M, missing=True) as s:
    s.synth(1234, 0.2)
    s.init(W0, 1.0)
    s.steps(3); s.sync()
    p = s.params()
    s.set_kernel_timing(True)
    s.steps(3); s.sync()
    seg = s.last_segments_ms()
    ms = s.last_steps_ms()
    ll = s.loglik()
np.save(sys.argv[2], np.concatenate([p["W"].ravel(), p["mu"].ravel(), [p["sigma2"]]]))
print(json.dumps({"segments_ms": seg, "step_ms": ms, "sigma2": p["sigma2"], "loglik": ll}))
''' % ROOT

if __name__ == "__main__":
    lg = sys.argv[1] if len(sys.argv) > 1 else "22"
    import numpy as np
    out = {}
    for name, env in (("tcgen05", {}), ("dmma", {"GPLVM_GRAM_DMMA": "1"})):
        r = subprocess.run([sys.executable, "-c", CHILD, lg, f"/tmp/ab_{name}.npy"], env={**os.environ, **env}, capture_output=True, text=True)
        if r.returncode != 0:
            print(name, "FAILED", r.stderr[-2000:])
            sys.exit(1)
        out[name] = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        print(name, json.dumps(out[name]))
    a, b = np.load("/tmp/ab_tcgen05.npy"), np.load("/tmp/ab_dmma.npy")
    print("max |diff| / max |.| after 3 steps:", float(np.max(np.abs(a - b)) / np.max(np.abs(b))))
