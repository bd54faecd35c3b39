"""CPU oracle for the GP kernel-matrix / Cholesky path -- TEST INFRASTRUCTURE ONLY (see oracle/ppca.py).

PARITY UNPINNED: the reference has no tests or golden vectors for this path and cannot be built here.

Reference lines restated (relative to /root/reference):
  src/GPLVM/GPExample.hs:79-91   sqDist          -> sq_dist   (expanded form a^2 + b^2 - 2ab; result is
                                                    len(v2) x len(v1))
  src/GPLVM/GPExample.hs:93-100  kernelFunction  -> kernel_function (exp(-0.5 * (1/0.1) * sqDist))
  src/GPLVM/GaussianProcess.hs:59-98  gpToPosteriorSample -> gp_posterior_pieces (without the RNG draw)
Extension named by BASELINE.json north_star / config 5 (NOT in the reference):
  ard_kernel, gp_chol_lml: Q-dimensional ARD RBF kernel, jittered Cholesky and log marginal likelihood;
  they reduce to sq_dist / kernel_function at Q = 1, inv_lengthscale2 = 1/0.1, variance = 1.
"""
from __future__ import annotations

import math
import numpy as np
import scipy.linalg as sla


def sq_dist(v1: np.ndarray, v2: np.ndarray) -> np.ndarray:
    """GPExample.hs:79-91.  out[j, i] = v1_i^2 + (v2_j^2 - 2 v1_i v2_j)  (note the association order)."""
    v1 = np.asarray(v1, dtype=np.float64); v2 = np.asarray(v2, dtype=np.float64)
    matrix2p = (v2 ** 2)[None, :] - 2.0 * np.outer(v1, v2)     # len(v1) x len(v2), :85-87
    return (v1 ** 2)[None, :] + matrix2p.T                      # :83, len(v2) x len(v1)


def kernel_function(v1: np.ndarray, v2: np.ndarray) -> np.ndarray:
    """GPExample.hs:93-100: exp(-0.5 * ((1/0.1) * d2))."""
    return np.exp(-0.5 * ((1.0 / 0.1) * sq_dist(v1, v2)))


def ard_kernel(X1: np.ndarray, X2: np.ndarray, inv_l2: np.ndarray, variance: float) -> np.ndarray:
    """out[j, i] = variance * exp(-0.5 * sum_q inv_l2[q] * (X1[i,q]^2 + (X2[j,q]^2 - 2 X1[i,q] X2[j,q]))).

    Same orientation (n2 x n1) and the same expanded-square form as sq_dist so that Q = 1 reproduces
    kernel_function bit for bit in exact arithmetic order."""
    X1 = np.asarray(X1, dtype=np.float64).reshape(len(X1), -1)
    X2 = np.asarray(X2, dtype=np.float64).reshape(len(X2), -1)
    acc = np.zeros((X2.shape[0], X1.shape[0]))
    for q in range(X1.shape[1]):
        acc += inv_l2[q] * sq_dist(X1[:, q], X2[:, q])
    return variance * np.exp(-0.5 * acc)


def gp_chol_lml(X: np.ndarray, y: np.ndarray, inv_l2: np.ndarray, variance: float, jitter: float):
    """K~ = K(X,X) + jitter I  (jitter default 5e-5, GaussianProcess.hs:67-68); L = chol(K~) lower;
    LML = -0.5 ||L^-1 y||^2 - sum log L_ii - (N/2) log 2 pi.   Returns (L, lml)."""
    K = ard_kernel(X, X, inv_l2, variance)
    n = K.shape[0]
    K[np.arange(n), np.arange(n)] += jitter
    L = np.linalg.cholesky(K)
    alpha = sla.solve_triangular(L, np.asarray(y, dtype=np.float64), lower=True)
    lml = -0.5 * float(alpha @ alpha) - float(np.sum(np.log(np.diag(L)))) - 0.5 * n * math.log(2 * math.pi)
    return L, lml


def gp_posterior_pieces_upper(observe: np.ndarray, x_train: np.ndarray, y_train: np.ndarray):
    """gpToPosteriorSample up to the random draw, GaussianProcess.hs:59-86, read LITERALLY with hmatrix's `chol`
    convention: cholSH returns the UPPER factor U (K = U^T U) and linearSolveS solves with it as is.  Returns
    (mean, posterior_cov_factor_upper) where sample = mean + factor @ normals (:89)."""
    k_ss = kernel_function(observe, observe)                                   # :61
    k = kernel_function(x_train, x_train)                                      # :64
    n = len(x_train)
    U = np.linalg.cholesky(k + 0.00005 * np.eye(n)).T                          # :67-68 (upper)
    k_s = kernel_function(observe, x_train)                                    # :71
    z = np.linalg.solve(U, k_s)                                                # :74  linearSolveS cholK K_s
    a = np.linalg.solve(U, np.asarray(y_train, dtype=np.float64)[:, None])     # :77
    mean = (a.T @ z)[0]                                                        # :80
    post = np.linalg.cholesky(k_ss + 1.0e-6 * np.eye(len(observe)) - z.T @ z).T  # :83-86 (upper)
    return mean, post


def gp_posterior_pieces(observe: np.ndarray, x_train: np.ndarray, y_train: np.ndarray):
    """gpToPosteriorSample up to (not including) the random draw, GaussianProcess.hs:59-86, with the
    factor taken as the LOWER Cholesky factor (SURVEY.md 8a ambiguity note).  Returns
    (mean, posterior_cov_factor) where sample = mean + factor @ normals."""
    k_ss = kernel_function(observe, observe)                                   # :61
    k = kernel_function(x_train, x_train)                                      # :64
    n = len(x_train)
    L = np.linalg.cholesky(k + 0.00005 * np.eye(n))                            # :67-68
    k_s = kernel_function(observe, x_train)                                    # :71  (n_train x n_obs)
    lk = np.linalg.solve(L, k_s)                                               # :74
    ly = np.linalg.solve(L, np.asarray(y_train, dtype=np.float64)[:, None])    # :77 (one column is enough)
    mean = (ly.T @ lk)[0]                                                      # :80
    post = np.linalg.cholesky(k_ss + 1.0e-6 * np.eye(len(observe)) - lk.T @ lk)  # :83-86
    return mean, post
