#!/usr/bin/env python
"""Generate haskell/patches/*.patch: unified diffs against the reference checkout that swap the BODIES of the hot-path
functions for calls into GPLVM.B200 (haskell/GPLVM/B200.hs) and link the library.  Signatures, records, error texts and
the RNG draws of the reference stay untouched.

    python haskell/make_patches.py [/root/reference]        # rewrites haskell/patches/

Apply with:  cd <reference checkout> && for p in .../haskell/patches/*.patch; do patch -p1 < $p; done
             cp .../haskell/GPLVM/B200.hs src/GPLVM/B200.hs
The patches are UNVERIFIED Haskell (no GHC in the build image); tests/test_haskell_shim.py checks that they apply.
"""
from __future__ import annotations

import difflib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def read(ref, rel):
    """Lines without their terminator, plus the file's line ending (some reference files are CRLF)."""
    with open(os.path.join(ref, rel), newline="") as f:
        text = f.read()
    eol = "\r\n" if "\r\n" in text else "\n"
    lines = text.split(eol)
    assert lines[-1] == "", rel + ": no newline at end of file (the unified diff would need a marker)"
    return lines[:-1], eol


def find(lines, prefix, start=0):
    for i in range(start, len(lines)):
        if lines[i].startswith(prefix):
            return i
    raise KeyError(prefix)


def replace_block(lines, first_prefix, last_prefix, new_block):
    a = find(lines, first_prefix)
    b = find(lines, last_prefix, a)
    return lines[:a] + new_block + lines[b + 1:]


def insert_after(lines, prefix, new_lines):
    i = find(lines, prefix)
    return lines[:i + 1] + new_lines + lines[i + 1:]


def patch_ppca(lines):
    lines = insert_after(lines, "import GPLVM.Util", ["import GPLVM.B200 (ppcaDense, ppcaMissing)"])
    lines = replace_block(lines, "emStepsFast learnMatrix@", "  in stepOne initMatrix initVariance 0", [
        "emStepsFast learnMatrix initMatrix initVariance stopParam =",
        "  -- E-step, M-step and log-likelihood of the whole EM loop run in libgplvm_b200 (gplvm_ppca_dense);",
        "  -- the stop rule (Left k = k + 1 steps, Right tol) is applied on the device.",
        "  let (w, var, ll) = ppcaDense (computeS learnMatrix) (computeS initMatrix) initVariance stopParam",
        "  in (delay w, var, ll, Nothing)",
    ])
    lines = replace_block(lines, "emStepsMissed learnMatrix@", "  in stepOne initMu initMatrix initVariance 0", [
        "emStepsMissed learnMatrix initMatrix initVariance stopParam =",
        "  -- NaN-masked EM loop, log-likelihood and restored matrix run in libgplvm_b200 (gplvm_ppca_missing)",
        "  let (w, var, ll, restored) = ppcaMissing (computeS learnMatrix) (computeS initMatrix) initVariance stopParam",
        "  in (delay w, var, ll, Just (delay restored))",
    ])
    return lines


def patch_typesafe_ppca(lines):
    lines = insert_after(lines, "import GPLVM.Util", ["import GPLVM.B200 (ppcaDense, ppcaMissing)"])
    lines = replace_block(lines, "emStepsFast learnMatrix initMatrix", "   in stepOne initMatrix initVariance 0", [
        "emStepsFast learnMatrix initMatrix initVariance stopParam =",
        "  -- the dimension proofs of makePPCATypeSafe have already run; the arithmetic is the untyped one (gplvm_ppca_dense)",
        "  let (w, var, ll) = ppcaDense (R.computeS $ getInternal learnMatrix) (R.computeS $ getInternal initMatrix)",
        "                               initVariance stopParam",
        "  in (DimMatrix (R.delay w), var, ll, Nothing)",
    ])
    lines = replace_block(lines, "emStepsMissed learnMatrix initMatrix", "  in stepOne initMu initMatrix initVariance 0", [
        "emStepsMissed learnMatrix initMatrix initVariance stopParam =",
        "  let (w, var, ll, restored) = ppcaMissing (R.computeS $ getInternal learnMatrix) (R.computeS $ getInternal initMatrix)",
        "                                           initVariance stopParam",
        "  in (DimMatrix (R.delay w), var, ll, Just (DimMatrix (R.delay restored)))",
    ])
    return lines


def patch_gaussian_process(lines):
    lines = insert_after(lines, "import System.Random", [
        "",
        "import Data.Array.Repa.Repr.ForeignPtr (F)",
        "import Data.Type.Equality ((:~:)(..))",
        "import Data.Typeable (Typeable, eqT)",
        "import qualified GPLVM.B200 as B200",
    ])
    lines = replace_block(lines, "  , Eq a", "  , Eq a", ["  , Eq a", "  , Typeable a"])
    i = find(lines, "-- | Main GP function")
    helpers = [
        "-- | cholSH on the GPU when a ~ Double (gplvm_cholesky; LOWER factor, K = L L^T), hmatrix for any other field",
        "cholGP :: forall a. GPConstraint a => Matrix D a -> Matrix D a",
        "cholGP m = case eqT :: Maybe (a :~: Double) of",
        "  Just Refl -> maybe (error \"cholGP: matrix is not positive definite\") delay",
        "                     (B200.cholesky (computeS m :: Matrix F Double))",
        "  Nothing   -> cholSH m",
        "",
        "-- | linearSolveS with the Cholesky factor on the GPU when a ~ Double (gplvm_tri_solve_lower: forward substitution)",
        "solveGP :: forall a. GPConstraint a => Matrix D a -> Matrix D a -> Maybe (Matrix D a)",
        "solveGP l b = case eqT :: Maybe (a :~: Double) of",
        "  Just Refl -> delay <$> B200.triSolveLower (computeS l :: Matrix F Double) (computeS b :: Matrix F Double)",
        "  Nothing   -> delay <$> linearSolveS l b",
        "",
    ]
    lines = lines[:i] + helpers + lines[i:]
    out = []
    for ln in lines:
        ln = ln.replace("let cholK = cholSH $ trainingKernel", "let cholK = cholGP $ trainingKernel")
        ln = ln.replace("let postF' = cholSH $", "let postF' = cholGP $")
        ln = ln.replace("cholKSolve <- delay <$> linearSolveS cholK testPointMean", "cholKSolve <- solveGP cholK testPointMean")
        ln = ln.replace("cholKSolveOut <- delay <$> linearSolveS cholK (", "cholKSolveOut <- solveGP cholK (")
        out.append(ln)
    return out


def patch_gp_example(lines):
    lines = insert_after(lines, "import GPLVM.Util", [
        "",
        "import Data.Array.Repa.Repr.ForeignPtr (F)",
        "import GPLVM.B200 (rbfKernel)",
    ])
    lines = replace_block(lines, "kernelFunction matrix1 matrix2 =", "        fun = exp", [
        "kernelFunction matrix1 matrix2 =",
        "    -- gplvm_rbf_kernel: exp (-0.5 * (1/0.1) * sqDist), len(vec2) x len(vec1) like sqDist above",
        "    delay $ rbfKernel (column matrix1) (column matrix2) [1/0.1] 1.0",
        "    where",
        "        column v = computeS (toMatrix' v) :: Matrix F Double",
    ])
    return lines


def patch_pca(lines):
    lines = insert_after(lines, "import GPLVM.Types", ["import qualified GPLVM.B200 as B200"])
    a = find(lines, "makePCA desiredDimensions input =")
    b = find(lines, "  in PCA{..}", a)
    body = [
        "makePCA desiredDimensions input =",
        "  -- meanCovS, eigSH, the projection and the pseudo-inverse restore run in libgplvm_b200 (gplvm_pca)",
        "  let _inputData = input",
        "      (Z :. yInp :. _) = extent input",
        "      (meanVec, cov, evals, evecs, final, restored) = B200.pca (computeS input) desiredDimensions",
        "      _meanMatrix = extend (Z :. yInp :. All) (delay meanVec)",
        "      _covariance = trustSym cov",
        "      _eigenValues = evals",
        "      _eigenVectors = delay evecs     -- columns are eigenvectors",
        "      _finalData = delay final        -- data in the eigenvector basis",
        "      _restoredData = delay restored  -- back in the original space, mean added",
        "  in PCA{..}",
    ]
    return lines[:a] + body + lines[b + 1:]


def patch_cabal(lines):
    i = find(lines, "      GPLVM.GPExample")
    lines = lines[:i] + ["      GPLVM.B200"] + lines[i:]
    j = find(lines, "  build-tool-depends:")
    lines = lines[:j] + ["  extra-libraries:",
                         "      gplvm_b200"] + lines[j:]
    return lines


TARGETS = [
    ("PPCA.patch", "src/GPLVM/PPCA.hs", patch_ppca),
    ("TypeSafe-PPCA.patch", "src/GPLVM/TypeSafe/PPCA.hs", patch_typesafe_ppca),
    ("PCA.patch", "src/GPLVM/PCA.hs", patch_pca),
    ("GaussianProcess.patch", "src/GPLVM/GaussianProcess.hs", patch_gaussian_process),
    ("GPExample.patch", "src/GPLVM/GPExample.hs", patch_gp_example),
    ("cabal.patch", "GPLVMHaskell.cabal", patch_cabal),
]


def generate(ref):
    out = {}
    for name, rel, fn in TARGETS:
        old, eol = read(ref, rel)
        new = fn(list(old))
        cr = "\r" if eol == "\r\n" else ""
        diff = difflib.unified_diff([ln + cr for ln in old], [ln + cr for ln in new], "a/" + rel, "b/" + rel, lineterm="", n=3)
        out[name] = "\n".join(diff) + "\n"
    return out


if __name__ == "__main__":
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    os.makedirs(os.path.join(HERE, "patches"), exist_ok=True)
    for name, text in generate(ref).items():
        with open(os.path.join(HERE, "patches", name), "w", newline="") as f:
            f.write(text)
        print(name, text.count("\n"), "lines")
