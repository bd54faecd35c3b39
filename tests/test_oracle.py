"""CPU tests of the oracle (test infrastructure): golden fixtures, closed-form known answers and the
agreement between its literal and sufficient-statistics formulations.  No GPU needed."""
import math

import numpy as np
import pytest

from oracle import gp as ogp
from oracle import ppca as oppca
from conftest import rel


def test_golden_dense_regenerates(iris):
    r = oppca.em_steps_fast(iris["Y"], iris["W0"], 1.0, ("left", 0))
    assert rel(r.W, iris["dense_s1_k0_W"]) < 1e-13
    assert abs(r.variance - float(iris["dense_s1_k0_var"])) < 1e-13 * abs(r.variance)
    assert abs(r.final_exp_likelihood - float(iris["dense_s1_k0_ll"])) < 1e-12 * abs(r.final_exp_likelihood)
    assert r.steps_done == 1


def test_golden_missing_regenerates(iris):
    r = oppca.em_steps_missed(iris["Yn"], iris["W0"], 1.0, ("left", 1))
    assert r.steps_done == 2
    assert rel(r.W, iris["miss_s1_k1_W"]) < 1e-12
    assert rel(r.restored_matrix, iris["miss_s1_k1_restored"]) < 1e-12
    assert abs(r.variance - float(iris["miss_s1_k1_var"])) < 1e-12 * r.variance


def test_dense_fixed_point_is_tipping_bishop(iris):
    """Known answer independent of the oracle: at the EM fixed point sigma2 is the mean of the discarded
    eigenvalues of the (1/N) covariance and span(W) the leading eigenvectors (Tipping & Bishop 1999)."""
    Y = iris["Y"]
    d, n = Y.shape
    Xc = Y - Y.mean(axis=1, keepdims=True)
    S = Xc @ Xc.T / n
    ev, U = np.linalg.eigh(S)
    r = oppca.em_steps_fast(Y, iris["W0"], 1.0, ("left", 4000))
    assert abs(r.variance - ev[0]) < 1e-8 * ev[0]
    P = oppca.projector(r.W)
    Pref = U[:, 1:] @ U[:, 1:].T
    assert np.max(np.abs(P - Pref)) < 1e-6
    # SURVEY.md 8c cross-check value
    assert abs(ev[0] - 0.023676192353627) < 1e-12


def test_golden_1001_steps_known_values(iris):
    assert abs(float(iris["dense_sbig_k1000_ll"]) - (-381.92805336)) < 1e-6
    assert abs(float(iris["miss_s1_k1000_var"]) - 0.0199236485) < 1e-7  # SURVEY 8c probe value (other init)
    assert abs(float(iris["miss_s1_k1000_ll"]) - (-344.89248114)) < 1e-5


def test_left_runs_k_plus_one_steps(iris):
    for k in (0, 1, 5):
        assert oppca.em_steps_fast(iris["Y"], iris["W0"], 1.0, ("left", k)).steps_done == k + 1
    assert oppca.em_steps_fast(iris["Y"], iris["W0"], 1.0, ("left", -3)).steps_done == 1


def test_right_tolerance_rule(iris):
    r = oppca.em_steps_fast(iris["Y"], iris["W0"], 1.0, ("right", 1e-6))
    assert r.steps_done == int(iris["dense_s1_tol_steps"])
    assert rel(r.W, iris["dense_s1_tol_W"]) < 1e-12


def test_masked_vectorised_equals_literal(iris):
    Yn, W0 = iris["Yn"], iris["W0"]
    mu = np.zeros(4)
    W, s2 = W0.copy(), 1.0
    for _ in range(3):
        a = oppca.missed_step_literal(Yn, mu, W, s2)
        b = oppca.missed_step_vectorised(Yn, mu, W, s2)
        assert rel(b[0], a[0]) < 1e-12 and rel(b[1], a[1]) < 1e-12
        assert abs(b[2] - a[2]) < 1e-12 * a[2]
        assert rel(b[5]["EX"], a[5]["EX"]) < 1e-12
        mu, W, s2 = a[0], a[1], a[2]
    ll_a = oppca.missed_loglik_literal(Yn, mu, W, s2, a[5]["known"])
    ll_b = oppca.missed_loglik_vectorised(Yn, mu, W, s2)
    assert abs(ll_a - ll_b) < 1e-12 * abs(ll_a)


def test_masked_on_dense_data_without_nan_is_consistent():
    """Property: with no NaN the masked step's E-step equals the dense E-step with mu_0 = 0 (un-centred)."""
    rng = np.random.default_rng(0)
    Y = rng.standard_normal((6, 40))
    W0 = rng.standard_normal((6, 2))
    mu, W, s2, *_ = oppca.missed_step_literal(Y, np.zeros(6), W0, 0.7)
    Minv = np.linalg.inv(W0.T @ W0 + 0.7 * np.eye(2))
    EX = Minv @ W0.T @ Y
    assert rel(mu, (Y - W0 @ EX).mean(axis=1)) < 1e-12


def test_dispatch_on_nan_sum(iris):
    assert oppca.make_ppca(iris["Y"], iris["W0"], 1.0, ("left", 0)).no_missed_data
    r = oppca.make_ppca(iris["Yn"], iris["W0"], 1.0, ("left", 0))
    assert not r.no_missed_data and r.restored_matrix.shape == (4, 150)


def test_typesafe_dimension_checks():
    oppca.check_typesafe_dims((4, 150), (4, 3), 3)
    with pytest.raises(ValueError, match="x2 and d1"):
        oppca.check_typesafe_dims((4, 150), (4, 2), 3)
    with pytest.raises(ValueError, match="y1 and y2"):
        oppca.check_typesafe_dims((4, 150), (5, 3), 3)
    with pytest.raises(ValueError, match="at least one column"):
        oppca.check_typesafe_dims((4, 0), (4, 3), 3)


def test_gp_oracle_golden(gp_golden):
    g = gp_golden
    K = ogp.kernel_function(g["x_train"], g["x_train"])
    assert np.array_equal(K, g["K"])
    Ks = ogp.kernel_function(g["x_test"], g["x_train"])
    assert Ks.shape == (5, 50) and np.array_equal(Ks, g["Ks"])     # len(v2) x len(v1), GPExample.hs:83
    # closed form: k(a, b) = exp(-5 (a - b)^2)
    a, b = g["x_test"][7], g["x_train"][2]
    assert abs(Ks[2, 7] - math.exp(-5.0 * (a - b) ** 2)) < 1e-12
    L, lml = ogp.gp_chol_lml(g["x_train"][:, None], g["y_train"], np.array([10.0]), 1.0, 0.00005)
    assert np.allclose(L @ L.T, K + 5e-5 * np.eye(5), atol=1e-14)
    # LML against the textbook formula with an explicit inverse
    Kj = K + 5e-5 * np.eye(5)
    ref = -0.5 * g["y_train"] @ np.linalg.solve(Kj, g["y_train"]) - 0.5 * np.linalg.slogdet(Kj)[1] - 2.5 * math.log(2 * math.pi)
    assert abs(lml - ref) < 1e-10 * abs(ref)


def test_ard_reduces_to_reference_kernel(gp_golden):
    g = gp_golden
    K1 = ogp.ard_kernel(g["x_test"][:, None], g["x_train"][:, None], np.array([1.0 / 0.1]), 1.0)
    assert np.array_equal(K1, g["Ks"])


def test_dense_fixed_point_matches_sklearn_ppca(iris):
    """Third-party known answer: scikit-learn's probabilistic PCA (Tipping & Bishop ML solution from an SVD).  sklearn
    normalises the covariance by N - 1, the EM of the reference by N (PPCA.hs:80), hence the factor."""
    sk = pytest.importorskip("sklearn.decomposition")
    Y = iris["Y"]
    d, n = Y.shape
    pca = sk.PCA(n_components=3, svd_solver="full").fit(Y.T)
    r = oppca.em_steps_fast(Y, iris["W0"], 1.0, ("left", 4000))
    assert abs(r.variance * n / (n - 1) - pca.noise_variance_) < 1e-7 * pca.noise_variance_
    P_sk = pca.components_.T @ pca.components_
    assert np.max(np.abs(oppca.projector(r.W) - P_sk)) < 1e-6
    # model covariance C = W W^T + sigma2 I against sklearn's get_covariance() (same normalisation caveat)
    C = (r.W @ r.W.T + r.variance * np.eye(d)) * n / (n - 1)
    assert np.max(np.abs(C - pca.get_covariance())) < 1e-6


def test_masked_em_without_nan_reaches_the_dense_fixed_point(iris):
    """Property tying the two reference algorithms together: on complete data emStepsMissed (which also estimates
    the mean) converges to the same maximum-likelihood sigma2 and subspace as emStepsFast."""
    Y = iris["Y"]
    a = oppca.em_steps_fast(Y, iris["W0"], 1.0, ("left", 3000))
    b = oppca.em_steps_missed_vectorised(Y, iris["W0"], 1.0, ("left", 3000))
    assert abs(a.variance - b.variance) < 1e-7 * a.variance
    assert np.max(np.abs(oppca.projector(a.W) - oppca.projector(b.W))) < 1e-6
    assert np.max(np.abs(b.restored_matrix - Y)) < 1.0   # the restored matrix is the rank-3 reconstruction (+ mean)


# ---------------------------------------------------------------------------------------------------
# exact fixed-point arithmetic of the INT8 complement sums (oracle/fixedpoint.py restates the device arithmetic)
# ---------------------------------------------------------------------------------------------------

def _fx_case(seed, n=4000, outlier=None):
    import oracle.fixedpoint as fx
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(n) * rng.choice([1e-3, 1.0, 37.5], n)
    if outlier:
        v[17] = outlier
    mask = rng.random(n) < 0.2
    hmax = max(fx.hi_word_abs(x) for x in v)
    return fx, v, mask, hmax


def test_fixed_point_sum_is_the_exact_sum_of_the_rounded_entries():
    from fractions import Fraction
    fx, v, mask, hmax = _fx_case(1)
    u = fx.to_fixed(v, hmax)
    assert int(u.max()) <= 2 ** 52 and int(u.min()) >= 0
    x = [int(a) - 2 ** 51 for a in u]                                   # the rounded fixed-point entries
    # each entry is within half a grid step of v * scale
    sc = Fraction(fx.scale_up(hmax))
    assert all(abs(Fraction(float(a)) * sc - b) <= Fraction(1, 2) for a, b in zip(v, x))
    exact = sum(b for b, m in zip(x, mask) if m)
    got = fx.masked_sum(v, mask, hmax)
    assert got == float(Fraction(exact) * Fraction(fx.scale_down(hmax)))   # one rounding of the exact integer sum


def test_fixed_point_sum_is_independent_of_the_split_and_close_to_fsum():
    import math
    fx, v, mask, hmax = _fx_case(2)
    ref = math.fsum(v[mask])
    sums = {fx.masked_sum(v, mask, hmax, splits=s) for s in (1, 2, 7, 18, 37)}
    assert len(sums) == 1                                                # bitwise identical for every split
    got = sums.pop()
    bound = mask.sum() * 2.0 ** (fx.class_exponent(hmax) - 1022 - 52)    # count x half a grid step
    assert abs(got - ref) <= bound + abs(ref) * 2.0 ** -52


def test_fixed_point_grid_is_set_by_the_class_maximum():
    """An outlier coarsens the grid of every other entry: this is what the device-side precision guard watches."""
    import math
    fx, v, mask, hmax = _fx_case(3, outlier=1e9)
    mask[17] = False
    ref = math.fsum(v[mask])
    got = fx.masked_sum(v, mask, hmax)
    step = 2.0 ** (fx.class_exponent(hmax) - 1022 - 51)
    assert step > 1e-7                                                   # the grid is coarse ...
    assert abs(got - ref) <= mask.sum() * step / 2                       # ... but the sum of the rounded entries is still exact


def test_mask_bit_layouts_expand_to_mask_bytes():
    import oracle.fixedpoint as fx
    rng = np.random.default_rng(4)
    std = [int(x) for x in rng.integers(0, 2 ** 32, 4, dtype=np.uint64)]   # bit i of word w = observation 32 w + i
    bits = [(std[o >> 5] >> (o & 31)) & 1 for o in range(128)]
    # tcgen05 kernel: (W >> s) & 0x01010101 = mask bytes of observations 32 wq + 4 s .. + 3
    for wq in range(4):
        W = fx.expand_order_word(std[wq])
        for s in range(8):
            word = (W >> s) & 0x01010101
            assert [(word >> (8 * b)) & 0xFF for b in range(4)] == bits[32 * wq + 4 * s: 32 * wq + 4 * s + 4]
    # mma.sync kernel: (W_t >> (2 ks + h)) & 0x01010101 = A-fragment register (k = 16 h + 4 t .. + 3 of k32 step ks)
    Wt = fx.fragment_order_words(std)
    for t in range(4):
        for ks in range(4):
            for h in range(2):
                word = (Wt[t] >> (2 * ks + h)) & 0x01010101
                k0 = 32 * ks + 16 * h + 4 * t
                assert [(word >> (8 * b)) & 0xFF for b in range(4)] == bits[k0: k0 + 4]


def test_dense_statistics_in_byte_planes_track_fp64_over_em_steps():
    """Numerical design of DESIGN.md section 7 (both factors in byte planes, truncated plane pairs): the statistics agree
    with FP64 to ~1e-15 and the EM trajectory to ~1e-13 -- four orders inside the 1e-9 parity bound."""
    import oracle.fixedpoint as fx
    rng = np.random.default_rng(5)
    d, n, m = 48, 900, 6
    Wt = rng.standard_normal((d, m)) * rng.uniform(0.01, 3.0, size=(1, m))           # latent scales over 2.5 decades
    Y = Wt @ rng.standard_normal((m, n)) + rng.standard_normal((d, 1)) + 0.3 * rng.standard_normal((d, n)) * rng.uniform(0.01, 2.0, size=(d, 1))
    Xc = oppca.center_dense(Y)
    tot = float(np.sum(Xc * Xc))
    W0 = rng.standard_normal((d, m))
    ref0 = oppca.dense_stats(Xc, W0, 1.0)
    cache = {}
    got0 = fx.dense_stats_i8(Xc, W0, 1.0, cache)
    assert np.max(np.abs(got0 - ref0)) < 1e-14 * np.max(np.abs(ref0))
    Wa, s2a, Wb, s2b = W0.copy(), 1.0, W0.copy(), 1.0
    for _ in range(8):
        Wa, s2a, _, _ = oppca.dense_update_from_stats(oppca.dense_stats(Xc, Wa, s2a), tot, n, Wa, s2a)
        Wb, s2b, _, _ = oppca.dense_update_from_stats(fx.dense_stats_i8(Xc, Wb, s2b, cache), tot, n, Wb, s2b)
    assert abs(s2a - s2b) < 1e-12 * s2a
    assert np.max(np.abs(Wa - Wb)) < 1e-12 * np.max(np.abs(Wa))


def test_fixed_point_gram_matches_fp64_gram():
    """The tensor-core Gram path of the masked E-step (masked_gram_tc.cu; stepOne, src/GPLVM/PPCA.hs:106-117): the exact
    integer sum of the rounded products is within one rounding per term (2^-52 of the column-pair scale) of the exactly
    rounded sum, for well-scaled W and for latent columns spread over 2^-12 .. 2^12."""
    import math
    import oracle.fixedpoint as fx
    rng = np.random.default_rng(42)
    D, M = 256, 16
    for spread in (0, 12):
        W = rng.standard_normal((D, M)) * np.exp2(np.linspace(-spread, spread, M))[None, :]
        assert fx.gram_gate_open(W)
        e = fx.gram_column_exponents(W)
        assert np.all(np.max(np.abs(W), axis=0) < np.exp2(e)) and np.all(np.max(np.abs(W), axis=0) >= np.exp2(e - 1))
        for frac in (0.0, 0.2, 0.9, 1.0):
            miss = rng.random(D) < frac
            G = fx.gram_fixed_point(W, miss)
            Wm = W[miss]
            for a in range(M):
                for c in range(a + 1):
                    exact = math.fsum((Wm[:, a] * Wm[:, c]).tolist())       # exactly rounded sum of the FP64 products
                    scale = 2.0 ** (int(e[a]) + int(e[c]))
                    assert abs(G[a, c] - exact) <= (miss.sum() + 1) * 2.0 ** -52 * scale, (spread, frac, a, c)
            assert np.array_equal(G, G.T)


def test_fixed_point_gram_gate():
    import oracle.fixedpoint as fx
    rng = np.random.default_rng(43)
    W = rng.standard_normal((256, 16))
    assert fx.gram_gate_open(W)
    W[:, 3] *= 2.0 ** -30
    W[17, 3] = 1.5            # 255 of 4096 non-zero entries lie 2^30 below their column maximum: > 1 %
    assert not fx.gram_gate_open(W)
    W2 = rng.standard_normal((256, 16)) * 1e-200     # exponents outside +-400: the power-of-two scales would leave the normal range
    assert not fx.gram_gate_open(W2)
