"""Host-side mirror of the reference's PCA interface over the C ABI (gplvm_pca).

Reference surface mirrored (paths relative to the reference checkout):
  src/GPLVM/PCA.hs:22-31            record PCA     -> class PCA (same field names)
  src/GPLVM/PCA.hs:35-62            makePCA        -> make_pca / makePCA
  src/GPLVM/TypeSafe/PCA.hs:56-74   makePCA (typed: d < x is proved first, `error "t"` otherwise) -> make_pca_type_safe

The reference's PCA takes the observations as ROWS (N x D): meanCovS works on rows (PCA.hs:42).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib


@dataclass
class PCA:
    """The reference record (PCA.hs:22-31)."""
    _inputData: np.ndarray      # N x D
    _covariance: np.ndarray     # D x D
    _eigenVectors: np.ndarray   # D x d, columns are eigenvectors
    _eigenValues: np.ndarray    # D, descending
    _finalData: np.ndarray      # N x d
    _restoredData: np.ndarray   # N x D
    _meanMatrix: np.ndarray     # N x D (the mean row repeated, PCA.hs:43)
    sweeps: int = 0


def make_pca(desired_dimensions: int, input_matrix) -> PCA:
    """makePCA (PCA.hs:35-62).  ``desired_dimensions >= D`` keeps every eigenvector (PCA.hs:49)."""
    lib = _lib.load()
    X = np.ascontiguousarray(input_matrix, dtype=np.float64)
    if X.ndim != 2:
        raise ValueError("input must be an N x D matrix (observations as rows)")
    N, D = X.shape
    d = min(int(desired_dimensions), D)
    if d < 1:
        raise ValueError("desired dimension must be positive")
    mean = np.empty(D)
    cov = np.empty((D, D))
    evals = np.empty(D)
    evecs = np.empty((D, d))
    final = np.empty((N, d))
    restored = np.empty((N, D))
    sweeps = C.c_int()
    _lib.check(lib.gplvm_pca(_lib.dptr(X), N, D, int(desired_dimensions), _lib.dptr(mean), _lib.dptr(cov), _lib.dptr(evals),
                             _lib.dptr(evecs), _lib.dptr(final), _lib.dptr(restored), C.byref(sweeps)))
    return PCA(X, cov, evecs, evals, final, restored, np.broadcast_to(mean, (N, D)), sweeps.value)


def make_pca_type_safe(desired_dimensions: int, input_matrix) -> PCA:
    """Typed makePCA (TypeSafe/PCA.hs:56-74): the proof d :<: x must hold, otherwise the reference calls ``error "t"``."""
    X = np.asarray(input_matrix)
    if X.ndim != 2:
        raise ValueError("input must be an N x D matrix (observations as rows)")
    if not int(desired_dimensions) < X.shape[1]:
        raise ValueError("t")
    return make_pca(desired_dimensions, X)


makePCA = make_pca
