"""CPU oracle for PCA by eigendecomposition -- TEST INFRASTRUCTURE ONLY (see oracle/ppca.py).

PARITY UNPINNED: the reference has no tests or golden vectors for this path and cannot be built here (no GHC).
Pinned instead by known answers (tests/test_oracle.py): numpy.cov / scikit-learn PCA on the iris fixture.

Reference lines restated (relative to /root/reference): src/GPLVM/PCA.hs:35-62 (makePCA); the typed twin
src/GPLVM/TypeSafe/PCA.hs:76-92 is arithmetically identical.  Third-party semantics assumed (hmatrix-0.20.0.0,
sources not vendored): meanCov = mean of the ROWS and (Xc^T Xc) / (N - 1); eigSH = symmetric eigendecomposition with
eigenvalues in DESCENDING order and eigenvectors in the columns; pinv of a matrix with orthonormal rows = transpose.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class PCAResult:
    covariance: np.ndarray     # D x D
    eigen_vectors: np.ndarray  # D x d (columns)
    eigen_values: np.ndarray   # D, descending
    final_data: np.ndarray     # N x d
    restored_data: np.ndarray  # N x D
    mean: np.ndarray           # D


def make_pca(desired_dimensions: int, X: np.ndarray) -> PCAResult:
    X = np.asarray(X, dtype=np.float64)
    n, dfeat = X.shape
    mean = X.sum(axis=0) / float(n)                                   # meanCovS, PCA.hs:42
    xc = X - mean                                                     # adjustInput', PCA.hs:44
    cov = (xc.T @ xc) / float(n - 1)
    w, v = np.linalg.eigh(cov)                                        # eigSH, PCA.hs:45 (ascending -> descending)
    order = np.argsort(-w, kind="stable")
    w, v = w[order], v[:, order]
    d = dfeat if desired_dimensions >= dfeat else desired_dimensions  # PCA.hs:48-51
    e = v[:, :d]                                                      # _eigenVectors (D x d), PCA.hs:53
    big = np.argmax(np.abs(e), axis=0)                                # sign convention of the library (LAPACK's is arbitrary)
    e = e * np.where(e[big, np.arange(d)] < 0, -1.0, 1.0)
    final = (e.T @ xc.T).T                                            # finalData', PCA.hs:55-56
    restored = (np.linalg.pinv(e.T) @ final.T).T + mean               # PCA.hs:58-60
    return PCAResult(cov, e, w, final, restored, mean)
