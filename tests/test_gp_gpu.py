"""GPU parity tests of the GP kernel-matrix build / Cholesky / log-marginal-likelihood path against the
oracle (oracle/gp.py, a restatement of src/GPLVM/GPExample.hs:79-100 and GaussianProcess.hs:59-98)."""
import math

import numpy as np
import pytest

import gplvmhaskell_b200 as g
from gplvmhaskell_b200 import _lib
from oracle import gp as ogp
from conftest import rel

pytestmark = pytest.mark.gpu


def test_kernel_function_reference_example(gp_golden):
    """kernelFunction on the reference's example data (GPExample.hs:14-118); orientation len(v2) x len(v1)."""
    gd = gp_golden
    K = g.kernel_function(gd["x_train"], gd["x_train"])
    Ks = g.kernel_function(gd["x_test"], gd["x_train"])
    Kss = g.kernel_function(gd["x_test"], gd["x_test"])
    assert Ks.shape == (5, 50)
    for a, b in ((K, gd["K"]), (Ks, gd["Ks"]), (Kss, gd["Kss"])):
        assert np.max(np.abs(a - b)) < 1e-14


@pytest.mark.parametrize("n1,n2,q", [(1, 1, 1), (33, 17, 2), (513, 1030, 3), (1000, 7, 5)])
def test_ard_kernel_vs_oracle(n1, n2, q):
    rng = np.random.default_rng(n1 + n2)
    X1, X2 = rng.uniform(0, 3, (n1, q)), rng.uniform(0, 3, (n2, q))
    il2 = rng.uniform(0.5, 4.0, q)
    K = g.ard_kernel(X1, X2, il2, 1.7)
    ref = ogp.ard_kernel(X1, X2, il2, 1.7)
    assert K.shape == (n2, n1)
    assert np.max(np.abs(K - ref)) < 1e-13


def test_chol_lml_reference_example(gp_golden):
    gd = gp_golden
    L, lml, logdet = g.gp_chol_lml(gd["x_train"][:, None], gd["y_train"], np.array([10.0]), 1.0, 0.00005)
    assert rel(L, gd["L"]) < 1e-12
    assert abs(lml - float(gd["lml"])) < 1e-9 * abs(float(gd["lml"]))
    assert abs(logdet - 2 * np.sum(np.log(np.diag(gd["L"])))) < 1e-9


@pytest.mark.parametrize("n,q", [(64, 1), (128, 2), (129, 2), (500, 2), (1000, 3), (2049, 2)])
def test_chol_lml_vs_oracle(n, q):
    rng = np.random.default_rng(n)
    X = rng.uniform(0, math.sqrt(n) * 0.6, (n, q))
    y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(n)
    il2 = np.full(q, 10.0)
    Lr, lmlr = ogp.gp_chol_lml(X, y, il2, 1.0, 5e-5)
    L, lml, logdet = g.gp_chol_lml(X, y, il2, 1.0, 5e-5)
    assert np.all(np.triu(L, 1) == 0.0)
    assert rel(L, Lr) < 1e-9
    assert abs(lml - lmlr) < 1e-9 * abs(lmlr)
    assert abs(logdet - 2 * np.sum(np.log(np.diag(Lr)))) < 1e-9 * max(1.0, abs(logdet))


def test_cholesky_round_trip_and_solve():
    rng = np.random.default_rng(0)
    n = 777
    B = rng.standard_normal((n, n))
    A = B @ B.T + n * np.eye(n)
    L = g.cholesky(A)
    assert rel(L @ L.T, A) < 1e-12
    assert rel(L, np.linalg.cholesky(A)) < 1e-11
    rhs = rng.standard_normal((n, 5))
    X = g.tri_solve_lower(L, rhs)
    assert rel(L @ X, rhs) < 1e-11


def test_not_positive_definite_is_reported():
    A = np.eye(200)
    A[150, 150] = -1.0
    with pytest.raises(g.GplvmError) as ei:
        g.cholesky(A)
    assert ei.value.code == _lib.ERR_NUMERIC


def test_lml_large_property():
    """N = 8192: no oracle factor comparison (kept small on the CPU side); check L L^T row samples and that the
    fused path (no y) returns the same log-determinant as the y path."""
    n = 8192
    rng = np.random.default_rng(3)
    X = rng.uniform(0, 60.0, (n, 2))
    y = np.sin(X[:, 0]) + np.cos(X[:, 1])
    il2 = np.full(2, 10.0)
    L, lml, logdet = g.gp_chol_lml(X, y, il2, 1.0, 5e-5)
    _, _, logdet2 = g.gp_chol_lml(X, None, il2, 1.0, 5e-5, want_factor=False)
    assert logdet == logdet2
    rows = rng.integers(0, n, 16)
    Kr = ogp.ard_kernel(X, X[rows], il2, 1.0)          # 16 x n
    Kr[np.arange(16), rows] += 5e-5
    assert np.max(np.abs(L[rows] @ L.T - Kr)) < 1e-11
    sign, ld = np.linalg.slogdet(ogp.ard_kernel(X, X, il2, 1.0) + 5e-5 * np.eye(n))
    assert abs(logdet - ld) < 1e-8 * abs(ld)


def test_posterior_reference_example(gp_golden):
    """gpToPosteriorSample on the reference's example (GPExample.hs:14-118): 50 test points, 5 training points."""
    gd = gp_golden
    rng = np.random.default_rng(4)
    normals = rng.standard_normal((50, 3))
    mean, post, samples = g.gp_to_posterior_sample(gd["x_test"], gd["x_train"], gd["y_train"], normals)
    assert np.max(np.abs(mean - gd["mean"])) < 1e-9
    # the posterior covariance is numerically rank deficient (1e-6 jitter): compare P P^T, not P
    assert np.max(np.abs(post @ post.T - gd["post"] @ gd["post"].T)) < 1e-9
    assert np.max(np.abs(samples - (mean[:, None] + post @ normals))) < 1e-12


def test_posterior_vs_oracle_larger():
    rng = np.random.default_rng(9)
    xt = np.sort(rng.uniform(-6, 6, 300))
    yt = np.sin(xt) + 0.05 * rng.standard_normal(300)
    xo = np.linspace(-5, 5, 411)
    mean_r, post_r = ogp.gp_posterior_pieces(xo, xt, yt)
    mean, post, _ = g.gp_to_posterior_sample(xo, xt, yt)
    assert np.max(np.abs(mean - mean_r)) < 1e-6          # K + 5e-5 I has condition number ~1e7
    assert np.max(np.abs(post @ post.T - post_r @ post_r.T)) < 1e-6


def test_posterior_upper_factor_reading_reference_example(gp_golden):
    """The reference read literally with hmatrix's UPPER Cholesky factor (gplvm_gp_posterior_upper): on the reference's own
    example (5 well separated training points, K ~ I) both readings are defined and differ by ~2e-7 in the mean."""
    gd = gp_golden
    rng = np.random.default_rng(4)
    normals = rng.standard_normal((50, 3))
    mean_r, post_r = ogp.gp_posterior_pieces_upper(gd["x_test"], gd["x_train"], gd["y_train"])
    mean, post, samples = g.gp_to_posterior_sample(gd["x_test"], gd["x_train"], gd["y_train"], normals, factor="upper")
    assert np.max(np.abs(mean - mean_r)) < 1e-9
    assert not np.any(np.tril(post, -1))                                             # upper factor
    assert np.max(np.abs(post.T @ post - post_r.T @ post_r)) < 1e-9                  # U^T U (rank deficient: compare the product)
    assert np.max(np.abs(samples - (mean[:, None] + post @ normals))) < 1e-12       # mean + U normals (GaussianProcess.hs:89)
    mean_l, _, _ = g.gp_to_posterior_sample(gd["x_test"], gd["x_train"], gd["y_train"])
    assert 1e-8 < np.max(np.abs(mean - mean_l)) < 1e-3                               # the two readings are different functions


def test_posterior_upper_factor_reading_can_fail_where_lower_works():
    """With correlated training points K_ss - Z^T Z (Z = U^-1 K_s) is not positive definite: the reference's cholSH would
    throw; the lower reading (the textbook posterior) is fine."""
    rng = np.random.default_rng(9)
    xt = np.sort(rng.uniform(-6, 6, 40))
    yt = np.sin(xt)
    xo = np.linspace(-5, 5, 61)
    g.gp_to_posterior_sample(xo, xt, yt)
    with pytest.raises(np.linalg.LinAlgError):
        ogp.gp_posterior_pieces_upper(xo, xt, yt)
    with pytest.raises(g.GplvmError) as ei:
        g.gp_to_posterior_sample(xo, xt, yt, factor="upper")
    assert ei.value.code == _lib.ERR_NUMERIC
