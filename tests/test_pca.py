"""PCA by eigendecomposition (makePCA, src/GPLVM/PCA.hs:35-62): the oracle against independent known answers (CPU),
and the CUDA path against the oracle (GPU).  Eigenvector signs are a convention, so parity goes through sign-invariant
quantities (eigenvalues, projector, restored data) plus the library's own sign rule."""
import os

import numpy as np
import pytest

from oracle import pca as opca
from conftest import rel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def iris_rows():
    """The reference's fixture as the reference's PCA wants it: observations as rows (150 x 4)."""
    return np.ascontiguousarray(dict(np.load(os.path.join(ROOT, "tests", "golden", "iris_ppca.npz")))["Y"].T)


def test_oracle_against_numpy_and_sklearn():
    X = iris_rows()
    assert X.shape == (150, 4)
    r = opca.make_pca(2, X)
    assert np.allclose(r.covariance, np.cov(X.T), rtol=1e-13, atol=0)                  # (N - 1) normalisation of meanCov
    assert np.all(np.diff(r.eigen_values) <= 0)                                          # eigSH: descending
    assert np.allclose(r.eigen_values, np.sort(np.linalg.eigvalsh(np.cov(X.T)))[::-1], rtol=1e-12)
    from sklearn.decomposition import PCA as SkPCA
    sk = SkPCA(n_components=2).fit(X)
    assert np.allclose(r.eigen_values[:2], sk.explained_variance_, rtol=1e-10)
    # same subspace, same reconstruction
    P = r.eigen_vectors @ r.eigen_vectors.T
    Psk = sk.components_.T @ sk.components_
    assert np.max(np.abs(P - Psk)) < 1e-10
    assert np.max(np.abs(r.restored_data - sk.inverse_transform(sk.transform(X)))) < 1e-10
    assert np.allclose(np.abs(r.final_data), np.abs(sk.transform(X)), atol=1e-10)
    assert r.final_data.shape == (150, 2) and r.restored_data.shape == (150, 4)
    # iris known answer: variance explained by the first two components
    assert abs(r.eigen_values[:2].sum() / r.eigen_values.sum() - 0.9776852063187949) < 1e-9


def test_oracle_keeps_everything_when_d_exceeds_the_dimension():
    X = np.random.default_rng(0).standard_normal((40, 5)) @ np.diag([3, 2, 1, 0.5, 0.1])
    r = opca.make_pca(9, X)                                                             # PCA.hs:49
    assert r.eigen_vectors.shape == (5, 5) and r.final_data.shape == (40, 5)
    assert np.max(np.abs(r.restored_data - X)) < 1e-12                                  # all components: exact restore


@pytest.mark.gpu
@pytest.mark.parametrize("n,d,keep", [(150, 4, 2), (1000, 7, 7), (5000, 64, 5), (3001, 33, 40), (20000, 256, 16), (300, 129, 3)])
def test_gpu_pca_vs_oracle(n, d, keep):
    import gplvmhaskell_b200 as g
    if (n, d) == (150, 4):
        X = iris_rows()
    else:
        rng = np.random.default_rng(n + d)
        X = rng.standard_normal((n, min(d, 12))) @ rng.standard_normal((min(d, 12), d)) * 2.0 + 0.3 * rng.standard_normal((n, d)) + rng.standard_normal(d)
    ref = opca.make_pca(keep, X)
    r = g.make_pca(keep, X)
    k = min(keep, d)
    assert r._eigenVectors.shape == (d, k) and r._finalData.shape == (n, k) and r._restoredData.shape == (n, d)
    assert rel(r._covariance, ref.covariance) < 1e-12
    assert rel(r._meanMatrix[0], ref.mean) < 1e-13 and r._meanMatrix.shape == (n, d)
    assert rel(r._eigenValues, ref.eigen_values) < 1e-11                                # relative to the largest eigenvalue
    assert np.max(np.abs(r._eigenVectors.T @ r._eigenVectors - np.eye(k))) < 1e-12     # orthonormal columns
    P, Pr = r._eigenVectors @ r._eigenVectors.T, ref.eigen_vectors @ ref.eigen_vectors.T
    gap = ref.eigen_values[k - 1] - (ref.eigen_values[k] if k < d else 0.0)
    if gap > 1e-6 * ref.eigen_values[0]:                                               # the kept subspace is well separated
        assert np.max(np.abs(P - Pr)) < 1e-9
        assert rel(r._restoredData, ref.restored_data) < 1e-9
        sep = np.min(-np.diff(ref.eigen_values[:k + 1 if k < d else k])) if k > 1 else gap
        if sep > 1e-6 * ref.eigen_values[0]:                                           # individual eigenvectors are too
            assert rel(r._eigenVectors, ref.eigen_vectors) < 1e-8 and rel(r._finalData, ref.final_data) < 1e-8
    assert 1 <= r.sweeps < 30


@pytest.mark.gpu
def test_gpu_pca_typed_variant_and_errors():
    import gplvmhaskell_b200 as g
    X = iris_rows()
    a = g.make_pca(2, X)
    b = g.make_pca_type_safe(2, X)
    assert np.array_equal(a._eigenVectors, b._eigenVectors) and np.array_equal(a._restoredData, b._restoredData)
    with pytest.raises(ValueError, match="^t$"):
        g.make_pca_type_safe(4, X)                                                       # d :<: x fails: error "t" (TypeSafe/PCA.hs:65)
    with pytest.raises(g.GplvmError):
        g.make_pca(2, X[:1])                                                             # N - 1 = 0


@pytest.mark.gpu
def test_gpu_pca_rank_deficient_input_converges():
    import gplvmhaskell_b200 as g
    rng = np.random.default_rng(5)
    X = rng.standard_normal((20, 3)) @ rng.standard_normal((3, 48))                      # rank 3 in 48 dimensions, N < D
    ref = opca.make_pca(3, X)
    r = g.make_pca(3, X)
    assert rel(r._eigenValues[:3], ref.eigen_values[:3]) < 1e-11
    assert np.max(np.abs(r._eigenValues[3:])) < 1e-12 * ref.eigen_values[0]
    assert rel(r._restoredData, X) < 1e-10
