"""Host-side mirror of the reference's GP kernel / Cholesky interface over the C ABI.

Reference surface mirrored (paths relative to the reference checkout):
  src/GPLVM/GPExample.hs:79-91      sqDist          (inside kernel_function)
  src/GPLVM/GPExample.hs:93-100     kernelFunction  -> kernel_function(v1, v2): len(v2) x len(v1)
  src/GPLVM/GaussianProcess.hs:67-68  cholSH (K + 5e-5 I) -> gp_chol_lml(..., jitter=5e-5) / cholesky
  src/GPLVM/GaussianProcess.hs:74,77  linearSolveS cholK -> tri_solve_lower
Extension named by the north star (not in the reference): ARD inputs and the log marginal likelihood.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib


def ard_kernel(X1, X2, inv_lengthscale2, variance: float = 1.0) -> np.ndarray:
    """K[j, i] = variance * exp(-0.5 * sum_q inv_l2[q] (X1[i,q]^2 + X2[j,q]^2 - 2 X1[i,q] X2[j,q])); n2 x n1."""
    lib = _lib.load()
    X1 = np.ascontiguousarray(np.asarray(X1, dtype=np.float64).reshape(len(X1), -1))
    X2 = np.ascontiguousarray(np.asarray(X2, dtype=np.float64).reshape(len(X2), -1))
    il2 = np.ascontiguousarray(np.asarray(inv_lengthscale2, dtype=np.float64).reshape(-1))
    if X1.shape[1] != X2.shape[1] or il2.shape[0] != X1.shape[1]:
        raise ValueError("X1, X2 and inv_lengthscale2 must agree on the input dimension Q")
    K = np.empty((X2.shape[0], X1.shape[0]), dtype=np.float64)
    _lib.check(lib.gplvm_rbf_kernel(_lib.dptr(X1), X1.shape[0], _lib.dptr(X2), X2.shape[0], X1.shape[1], _lib.dptr(il2),
                                    float(variance), _lib.dptr(K)))
    return K


def kernel_function(v1, v2) -> np.ndarray:
    """kernelFunction (GPExample.hs:93-100): RBF with l^2 = 0.1, unit variance, 1-D inputs."""
    return ard_kernel(np.asarray(v1, dtype=np.float64)[:, None], np.asarray(v2, dtype=np.float64)[:, None],
                      np.array([1.0 / 0.1]), 1.0)


kernelFunction = kernel_function


def gp_chol_lml(X, y, inv_lengthscale2, variance: float = 1.0, jitter: float = 0.00005, want_factor: bool = True):
    """Kernel build fused with the Cholesky and the log marginal likelihood.  Returns (L or None, lml, logdet)."""
    lib = _lib.load()
    X = np.ascontiguousarray(np.asarray(X, dtype=np.float64).reshape(len(X), -1))
    N, Q = X.shape
    il2 = np.ascontiguousarray(np.asarray(inv_lengthscale2, dtype=np.float64).reshape(-1))
    yv = None if y is None else np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1))
    if il2.shape[0] != Q:
        raise ValueError("inv_lengthscale2 must have Q = %d entries" % Q)
    if yv is not None and yv.shape[0] != N:
        raise ValueError("y must have N = %d entries" % N)
    L = np.empty((N, N), dtype=np.float64) if want_factor else None
    lml, logdet = C.c_double(), C.c_double()
    _lib.check(lib.gplvm_gp_chol_lml(_lib.dptr(X), N, Q, _lib.dptr(yv), _lib.dptr(il2), float(variance), float(jitter),
                                     _lib.dptr(L), C.byref(lml) if yv is not None else None, C.byref(logdet)))
    return L, (lml.value if yv is not None else None), logdet.value


def cholesky(A) -> np.ndarray:
    """Lower Cholesky factor (cholSH with the factor taken as lower)."""
    lib = _lib.load()
    A = np.ascontiguousarray(A, dtype=np.float64)
    if A.ndim != 2 or A.shape[0] != A.shape[1]:
        raise ValueError("A must be square")
    N = A.shape[0]
    L = np.empty((N, N), dtype=np.float64)
    _lib.check(lib.gplvm_cholesky(_lib.dptr(A), N, _lib.dptr(L)))
    return L


def tri_solve_lower(L, B) -> np.ndarray:
    """Solve L X = B for lower-triangular L (linearSolveS cholK ..., GaussianProcess.hs:74,77)."""
    lib = _lib.load()
    L = np.ascontiguousarray(L, dtype=np.float64)
    if L.ndim != 2 or L.shape[0] != L.shape[1]:
        raise ValueError("L must be square")
    if np.asarray(B).shape[0] != L.shape[0]:
        raise ValueError("B must have N = %d rows" % L.shape[0])
    B2 = np.ascontiguousarray(np.asarray(B, dtype=np.float64).reshape(L.shape[0], -1))
    X = np.empty_like(B2)
    _lib.check(lib.gplvm_tri_solve_lower(_lib.dptr(L), L.shape[0], _lib.dptr(B2), B2.shape[1], _lib.dptr(X)))
    return X.reshape(np.asarray(B).shape)


def gp_to_posterior_sample(observe, x_train, y_train, normals=None, inv_lengthscale2=(1.0 / 0.1,), variance: float = 1.0,
                           jitter_train: float = 0.00005, jitter_post: float = 1.0e-6, factor: str = "lower"):
    """gpToPosteriorSample (GaussianProcess.hs:51-98) with caller-supplied standard normals (n_obs x n_samples; the
    reference draws them from mkStdGen (-4)).  Returns (mean, posterior_factor, samples or None); raises GplvmError
    where the reference returns Nothing / cholSH throws.  ``factor="upper"`` reads the reference literally with hmatrix's
    upper Cholesky factor (gplvm_gp_posterior_upper; see include/gplvm_b200.h): different mean / factor / samples."""
    if factor not in ("lower", "upper"):
        raise ValueError("factor must be 'lower' or 'upper'")
    lib = _lib.load()
    Xo = np.ascontiguousarray(np.asarray(observe, dtype=np.float64).reshape(len(observe), -1))
    Xt = np.ascontiguousarray(np.asarray(x_train, dtype=np.float64).reshape(len(x_train), -1))
    yt = np.ascontiguousarray(np.asarray(y_train, dtype=np.float64).reshape(-1))
    il2 = np.ascontiguousarray(np.asarray(inv_lengthscale2, dtype=np.float64).reshape(-1))
    no, nt, q = Xo.shape[0], Xt.shape[0], Xo.shape[1]
    if Xt.shape[1] != q or il2.shape[0] != q:
        raise ValueError("observe, x_train and inv_lengthscale2 must agree on the input dimension Q")
    if yt.shape[0] != nt:
        raise ValueError("y_train must have n_train = %d entries" % nt)
    mean = np.empty(no)
    post = np.empty((no, no))
    samples = None
    nrm = None
    ns = 0
    if normals is not None:
        nrm = np.ascontiguousarray(np.asarray(normals, dtype=np.float64).reshape(no, -1))
        ns = nrm.shape[1]
        samples = np.empty((no, ns))
    fn = lib.gplvm_gp_posterior if factor == "lower" else lib.gplvm_gp_posterior_upper
    _lib.check(fn(_lib.dptr(Xo), no, _lib.dptr(Xt), _lib.dptr(yt), nt, q, _lib.dptr(il2), float(variance),
                                      float(jitter_train), float(jitter_post), _lib.dptr(nrm), ns, _lib.dptr(mean), _lib.dptr(post),
                                      _lib.dptr(samples)))
    return mean, post, samples


gpToPosteriorSample = gp_to_posterior_sample
