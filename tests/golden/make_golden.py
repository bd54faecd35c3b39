"""Generate the committed golden fixtures under tests/golden/.  Run in the build container only:

    python tests/golden/make_golden.py

It reads the reference's two input fixtures (/root/reference/test.pca.txt, test.pca.nan.txt -- the
only test data the reference ships, README.md:66-75) and the GP example data of
src/GPLVM/GPExample.hs:14-118, runs the CPU oracle (oracle/ppca.py, oracle/gp.py) from a recorded
(W0, sigma2_0), and stores inputs + outputs as .npz.  The GPU box has no /root/reference, so the tests
read these files instead.  The reference itself cannot be run (no GHC) => these are ORACLE goldens,
"parity unpinned" (see oracle/ppca.py header).
"""
import os, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ppca, gp  # noqa: E402

REF = "/root/reference"


def parse_cli_text(txt: str) -> np.ndarray:
    """app/Main.hs:43-44: rows of space separated doubles or NaN; returns D x N (transpose)."""
    rows = [[float(tok) for tok in line.split()] for line in txt.replace("\r\n", "\n").split("\n") if line.strip()]
    return np.array(rows, dtype=np.float64).T


def main():
    Y = parse_cli_text(open(os.path.join(REF, "test.pca.txt")).read())
    Yn = parse_cli_text(open(os.path.join(REF, "test.pca.nan.txt")).read())
    assert Y.shape == (4, 150) and Yn.shape == (4, 150) and int(np.isnan(Yn).sum()) == 124
    rng = np.random.default_rng(4321)
    W0 = rng.standard_normal((4, 3))
    out = dict(Y=Y, Yn=Yn, W0=W0)
    # dense: benign sigma2_0 and the StdGen-like huge one (PPCA.hs:43), 1 step and 1001 steps
    for tag, s20 in (("s1", 1.0), ("sbig", 846930886.0)):
        for k in (0, 1000):
            r = ppca.em_steps_fast(Y, W0, s20, ("left", k))
            out[f"dense_{tag}_k{k}_W"] = r.W
            out[f"dense_{tag}_k{k}_var"] = r.variance
            out[f"dense_{tag}_k{k}_ll"] = r.final_exp_likelihood
    r = ppca.em_steps_fast(Y, W0, 1.0, ("right", 1e-6))
    out["dense_s1_tol_W"], out["dense_s1_tol_var"], out["dense_s1_tol_ll"], out["dense_s1_tol_steps"] = r.W, r.variance, r.final_exp_likelihood, r.steps_done
    # masked: literal per-observation algorithm
    for tag, s20 in (("s1", 1.0), ("sbig", 846930886.0)):
        for k in (0, 1, 1000):
            if tag == "sbig" and k == 1:
                continue
            r = ppca.em_steps_missed(Yn, W0, s20, ("left", k))
            out[f"miss_{tag}_k{k}_W"] = r.W
            out[f"miss_{tag}_k{k}_var"] = r.variance
            out[f"miss_{tag}_k{k}_ll"] = r.final_exp_likelihood
            out[f"miss_{tag}_k{k}_restored"] = r.restored_matrix
    r = ppca.em_steps_missed(Yn, W0, 1.0, ("right", 1e-5))
    out["miss_s1_tol_W"], out["miss_s1_tol_var"], out["miss_s1_tol_ll"], out["miss_s1_tol_steps"], out["miss_s1_tol_restored"] = r.W, r.variance, r.final_exp_likelihood, r.steps_done, r.restored_matrix
    np.savez_compressed(os.path.join(HERE, "iris_ppca.npz"), **out)

    # GP example data (GPExample.hs:14-68 is linspace(-5,5,50) rounded to 8 decimals; :102-118)
    x_test = np.round(np.linspace(-5.0, 5.0, 50), 8)
    x_train = np.array([-4.0, -3.0, -2.0, -1.0, 1.0])
    y_train = np.sin(x_train)
    K = gp.kernel_function(x_train, x_train)
    Ks = gp.kernel_function(x_test, x_train)
    Kss = gp.kernel_function(x_test, x_test)
    L, lml = gp.gp_chol_lml(x_train[:, None], y_train, np.array([10.0]), 1.0, 0.00005)
    mean, post = gp.gp_posterior_pieces(x_test, x_train, y_train)
    np.savez_compressed(os.path.join(HERE, "gp_example.npz"), x_test=x_test, x_train=x_train, y_train=y_train,
                        K=K, Ks=Ks, Kss=Kss, L=L, lml=lml, mean=mean, post=post)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
