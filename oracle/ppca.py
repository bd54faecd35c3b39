"""CPU oracle for the PPCA EM hot path of serokell/GPLVMHaskell -- TEST INFRASTRUCTURE ONLY.

This file is a NumPy/LAPACK FP64 restatement of the reference's algorithm, written so that every
function follows the reference's *operation order*.  It is the checker the CUDA path is compared
against; it is never shipped, never called by ``gplvm_b200`` and never measured as the product.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it.

PARITY UNPINNED: the reference ships no tests, no golden vectors and cannot be built here (no GHC,
see SURVEY.md section 8c), so this oracle is pinned only against (i) the closed-form Tipping & Bishop
fixed point of dense PPCA and scikit-learn's PPCA (tests/test_oracle.py), (ii) agreement between its literal and its
sufficient-statistics formulations and between the dense and masked fixed points on complete data.

All matrices use the reference layout: ``Y`` is D x N (feature-major: rows are features, columns are
observations; app/Main.hs:24, src/GPLVM/PPCA.hs:41,60), ``W`` is D x M.

Reference lines restated (all relative to /root/reference):
  src/GPLVM/PPCA.hs:32-49    makePPCA            -> make_ppca
  src/GPLVM/PPCA.hs:54-94    emStepsFast         -> em_steps_fast
  src/GPLVM/PPCA.hs:96-182   emStepsMissed       -> em_steps_missed (literal per-observation closures)
  src/GPLVM/Util.hs:159-178  meanColumn / meanColumnWithNan / substractMean
  src/GPLVM/Util.hs:68-114   deleteRows / getRows / deleteColumns / getColumns / diag (index deletion)
  src/GPLVM/Util.hs:193-198  mapDiagonal
  src/GPLVM/Util.hs:54-58    trace2S (trace over the min(r,c) diagonal)
  src/GPLVM/TypeSafe/PPCA.hs:42-79  makePPCATypeSafe dimension checks -> check_typesafe_dims
Third-party semantics assumed (hmatrix-0.20.0.0 via repa-linear-algebra@eb6b72d5, stack.yaml:13-19,
sources not vendored): inv/invS = LU inverse, detS = LU determinant, chol = dpotrf upper factor,
solveS = least-squares solve (exact for the SPD systems here), mul* = dgemm.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple, Union

import numpy as np

Stop = Tuple[str, Union[int, float]]  # ("left", k)  or ("right", tol) -- Either Int Double, PPCA.hs:36-37


@dataclass
class PPCAResult:
    """Mirror of the reference record ``PPCA`` (src/GPLVM/PPCA.hs:21-30) minus the echoed inputs."""
    no_missed_data: bool
    W: np.ndarray                     # D x M
    variance: float
    final_exp_likelihood: float
    restored_matrix: Optional[np.ndarray]  # D x N or None
    steps_done: int                   # number of EM steps executed (k+1 for ("left", k))


# ----------------------------------------------------------------------------------------------
# Util.hs restatements
# ----------------------------------------------------------------------------------------------

def mean_column(mat: np.ndarray) -> np.ndarray:
    """Util.hs:159-164: per-column mean of an n x c matrix (sum / n)."""
    n = mat.shape[0]
    return mat.sum(axis=0) / float(n)


def mean_column_with_nan(mat: np.ndarray) -> np.ndarray:
    """Util.hs:166-173: folds over the innermost dimension skipping NaN; sum / count per row."""
    known = ~np.isnan(mat)
    s = np.where(known, mat, 0.0).sum(axis=1)
    c = known.sum(axis=1).astype(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        return s / c


def map_diagonal_add(mat: np.ndarray, value: float) -> np.ndarray:
    """Util.hs:193-198 with f = (+ value): add to entries whose indices are all equal."""
    out = np.array(mat, dtype=np.float64, copy=True)
    k = min(out.shape)
    out[np.arange(k), np.arange(k)] += value
    return out


def trace2(mat: np.ndarray) -> float:
    """Util.hs:54-58."""
    k = min(mat.shape)
    return float(mat[np.arange(k), np.arange(k)].sum())


def _chol_upper(a: np.ndarray) -> np.ndarray:
    """hmatrix ``chol``: upper factor U with a = U^T U (PPCA.hs:75)."""
    return np.linalg.cholesky(a).T


# ----------------------------------------------------------------------------------------------
# Dense EM  (PPCA.hs:54-94)
# ----------------------------------------------------------------------------------------------

def center_dense(Y: np.ndarray) -> np.ndarray:
    """PPCA.hs:61: subtract the per-feature mean (Util.hs:175-178 on the transposed matrix)."""
    return (Y.T - mean_column(Y.T)).T


def dense_step(Xc: np.ndarray, W: np.ndarray, sigma2: float):
    """One stepOne+stepTwo of emStepsFast (PPCA.hs:63-82), literal order of operations.

    Returns (newW, newVariance, maxDiffNewOldW, diffVariance)."""
    d, n = Xc.shape
    # stepOne, PPCA.hs:65-69
    m = map_diagonal_add(W.T @ W, sigma2)
    inv_m = np.linalg.inv(m)
    exp_ez = inv_m @ (W.T @ Xc)                       # M x N, materialised like the reference
    exp_xez = Xc @ exp_ez.T                           # D x M
    exp_zz = (n * sigma2) * inv_m + exp_ez @ exp_ez.T  # M x M
    # stepTwo, PPCA.hs:74-82
    new_w = exp_xez @ np.linalg.inv(exp_zz)
    u = _chol_upper(exp_zz)
    wr = new_w @ u.T
    total_sum = float(np.sum(Xc ** 2))
    second = 2.0 * float(np.sum(exp_ez * (new_w.T @ Xc)))
    third = float(np.sum(wr ** 2))
    new_var = (total_sum - second + third) / float(n * d)
    max_diff_w = float(np.max(np.abs(W - new_w))) if W.size else 0.0
    max_diff_w = max(0.0, max_diff_w)                 # foldAllS max 0.0
    diff_var = abs(new_var - sigma2)
    return new_w, new_var, max_diff_w, diff_var


def dense_loglik(Xc: np.ndarray, W: np.ndarray, sigma2: float, literal_det: bool = True) -> float:
    """PPCA.hs:83-89.  Keeps the reference's /(N-1) inside and -N/2 outside.

    ``literal_det=True`` uses log(det C) exactly as the reference; if that over/underflows (large D)
    the slogdet value is used instead and the divergence from the reference is by construction."""
    d, n = Xc.shape
    new_c = map_diagonal_add(W @ W.T, sigma2)
    new_inv_m = np.linalg.inv(map_diagonal_add(W.T @ W, sigma2))
    new_inv_c = map_diagonal_add(-((W @ new_inv_m) @ W.T) / sigma2, 1.0 / sigma2)
    if literal_det:
        det = np.linalg.det(new_c)
        logdet = math.log(det) if (det > 0 and math.isfinite(det)) else np.linalg.slogdet(new_c)[1]
    else:
        logdet = np.linalg.slogdet(new_c)[1]
    tr = trace2((new_inv_c @ (Xc @ Xc.T)) / float(n - 1))
    return -1.0 * (n / 2.0) * (d * math.log(2 * math.pi) + logdet + tr)


def dense_stats(Xc: np.ndarray, W: np.ndarray, sigma2: float) -> np.ndarray:
    """Per-shard statistics of one dense E-step (appendix A.1; additive over row shards of the CENTRED data):
    packed [XEz (D x M) | Ez Ez^T (M x M)]."""
    inv_m = np.linalg.inv(map_diagonal_add(W.T @ W, sigma2))
    ez = inv_m @ (W.T @ Xc)
    return np.concatenate([(Xc @ ez.T).ravel(), (ez @ ez.T).ravel()])


def dense_update_from_stats(packed: np.ndarray, total_sum: float, n_total: int, W: np.ndarray, sigma2: float):
    """Replicated M-step of appendix A.1 from the all-reduced statistics; total_sum = |Xc|_F^2 (all shards)."""
    d, m = W.shape
    xez = packed[:d * m].reshape(d, m)
    ee = packed[d * m:].reshape(m, m)
    inv_m = np.linalg.inv(map_diagonal_add(W.T @ W, sigma2))
    ezz = (n_total * sigma2) * inv_m + ee
    new_w = xez @ np.linalg.inv(ezz)
    second = 2.0 * float(np.sum(new_w * xez))
    third = float(np.sum((new_w @ _chol_upper(ezz).T) ** 2))
    new_var = (total_sum - second + third) / float(n_total * d)
    return new_w, new_var, max(0.0, float(np.max(np.abs(W - new_w)))), abs(new_var - sigma2)


def _continue(stop: Stop, iteration: int, diff_var: float, max_diff_w: float) -> bool:
    """PPCA.hs:90-92 / :173-181."""
    kind, val = stop
    if kind == "left":
        return int(val) > iteration
    return max(diff_var, max_diff_w) > float(val)


def em_steps_fast(Y: np.ndarray, W0: np.ndarray, sigma2_0: float, stop: Stop) -> PPCAResult:
    """emStepsFast, PPCA.hs:54-94.  ("left", k) runs iterations 0..k inclusive = k+1 EM steps."""
    Xc = center_dense(np.asarray(Y, dtype=np.float64))
    W = np.array(W0, dtype=np.float64, copy=True)
    s2 = float(sigma2_0)
    it = 0
    while True:
        W_new, s2_new, mdw, dv = dense_step(Xc, W, s2)
        if _continue(stop, it, dv, mdw):
            W, s2, it = W_new, s2_new, it + 1
            continue
        ll = dense_loglik(Xc, W_new, s2_new)
        return PPCAResult(True, W_new, s2_new, ll, None, it + 1)


# ----------------------------------------------------------------------------------------------
# Missing-data EM  (PPCA.hs:96-182) -- literal, one closure per observation like the reference
# ----------------------------------------------------------------------------------------------

def _missed_estep_literal(Y, mu, W, sigma2):
    """stepOne, PPCA.hs:106-117: per observation i the unknown indices, W_i, M_i^-1, E[x_i]."""
    d, n = Y.shape
    unknown, known, exp_x, inv_mp = [], [], [], []
    for i in range(n):
        yi = Y[:, i]
        unk = np.flatnonzero(np.isnan(yi))            # unknownIndices, :109-111
        kn = np.setdiff1d(np.arange(d), unk)          # deleteRows = getRows ([0..] \\ idx), Util.hs:73-74
        w_p = W[kn, :]                                # oldWP, :114
        m_p = map_diagonal_add(w_p.T @ w_p, sigma2)   # mP, :115
        inv_m = np.linalg.inv(m_p)                    # invMP, :116
        ex = (inv_m @ w_p.T) @ (yi[kn] - mu[kn])      # expXi, :117
        unknown.append(unk); known.append(kn); exp_x.append(ex); inv_mp.append(inv_m)
    return unknown, known, exp_x, inv_mp


def missed_step_literal(Y: np.ndarray, mu: np.ndarray, W: np.ndarray, sigma2: float):
    """One stepOne+stepTwo of emStepsMissed (PPCA.hs:106-147,170-171), literal.

    Returns (newMu, newW, newVariance, maxDiffNewOldW, diffVariance, estep) where ``estep`` carries
    the per-observation E-step quantities (needed by the lazy LL / restore of the final step)."""
    d, n = Y.shape
    m_lat = W.shape[1]
    unknown, known, exp_x, inv_mp = _missed_estep_literal(Y, mu, W, sigma2)
    EX = np.stack(exp_x, axis=1)                                           # expX, :122  (M x N)
    new_mu = mean_column_with_nan(Y - W @ EX)                               # newMu, :123 (old W)
    # newW, :125-138: per feature j over the observations that know feature j
    new_w = np.zeros((d, m_lat))
    for j in range(d):
        kn_cols = np.flatnonzero(~np.isnan(Y[j, :]))                        # knownColumns, :131
        ex_pj = EX[:, kn_cols]                                              # expXPj, :132
        sum_inv = np.zeros((m_lat, m_lat))
        for i in kn_cols:                                                   # sumInvMP, :133 (foldl1 +^)
            sum_inv = sum_inv + inv_mp[i]
        gj = ex_pj @ ex_pj.T + sigma2 * sum_inv                             # gj, :134
        rhs = ex_pj @ (Y[j, kn_cols] - new_mu[j])                           # expXtrXj, :136
        new_w[j, :] = np.linalg.solve(gj, rhs)                              # solveS, :137
    # newVariance, :140-147 -- new W, new mu with OLD E[x_i], M_i^-1 and sigma2
    acc = 0.0
    for i in range(n):
        kn = known[i]
        w_p = new_w[kn, :]
        resid = Y[kn, i] - w_p @ exp_x[i] - new_mu[kn]
        acc += float(np.sum(resid ** 2) + sigma2 * np.sum(np.diag((w_p @ inv_mp[i]) @ w_p.T)))
    known_vars = float(np.sum(~np.isnan(Y)))
    new_var = acc / known_vars
    max_diff_w = max(0.0, float(np.max(np.abs(W - new_w))))
    diff_var = abs(new_var - sigma2)
    estep = dict(unknown=unknown, known=known, EX=EX)
    return new_mu, new_w, new_var, max_diff_w, diff_var, estep


def missed_loglik_literal(Y, new_mu, new_w, new_var, known) -> float:
    """expLogLikelihood, PPCA.hs:149-159 (C_i is #O_i x #O_i; the reference calls it invMY)."""
    n = Y.shape[1]
    acc = 0.0
    for i in range(n):
        kn = known[i]
        yc = (Y[kn, i] - new_mu[kn])[:, None]
        c_i = map_diagonal_add(new_w[kn, :] @ new_w[kn, :].T, new_var)
        sign, logdet = np.linalg.slogdet(c_i)
        det = np.linalg.det(c_i)
        ld = math.log(det) if (det > 0 and math.isfinite(det)) else logdet
        acc += len(kn) * math.log(2 * math.pi) + ld + trace2(np.linalg.solve(c_i, yc @ yc.T))
    return -1.0 * acc / 2.0


def missed_restore_literal(new_mu, new_w, new_var, EX) -> np.ndarray:
    """restoredData, PPCA.hs:161-168: uses EX of the *last E-step executed* (stale parameters)."""
    mu_x = mean_column(EX.T)                                                # :162
    final_mu = new_mu + new_w @ mu_x                                        # :163
    wtw = new_w.T @ new_w                                                   # :164
    factor1 = new_w @ np.linalg.inv(wtw)                                    # :165
    factor2 = map_diagonal_add(wtw, new_var)                                # :166
    restored_centered = factor1 @ factor2 @ EX                              # :167
    return restored_centered + final_mu[:, None]                           # :168


def em_steps_missed(Y: np.ndarray, W0: np.ndarray, sigma2_0: float, stop: Stop,
                    step_fn=None) -> PPCAResult:
    """emStepsMissed, PPCA.hs:96-182; mu_0 = 0 (:103)."""
    Y = np.asarray(Y, dtype=np.float64)
    d, n = Y.shape
    step_fn = step_fn or missed_step_literal
    mu = np.zeros(d)
    W = np.array(W0, dtype=np.float64, copy=True)
    s2 = float(sigma2_0)
    it = 0
    while True:
        mu_new, W_new, s2_new, mdw, dv, est = step_fn(Y, mu, W, s2)
        if _continue(stop, it, dv, mdw):
            mu, W, s2, it = mu_new, W_new, s2_new, it + 1
            continue
        ll = missed_loglik_literal(Y, mu_new, W_new, s2_new, est["known"])
        restored = missed_restore_literal(mu_new, W_new, s2_new, est["EX"])
        return PPCAResult(False, W_new, s2_new, ll, restored, it + 1)


# ----------------------------------------------------------------------------------------------
# Sufficient-statistics restatement of the masked step -- vectorised so
# that parity tests can run at sizes where the literal closures take minutes.  Validated against
# missed_step_literal in tests/test_oracle.py; it is still an oracle (CPU, NumPy), not the product.
# ----------------------------------------------------------------------------------------------

def missed_constants(Y: np.ndarray):
    """Data constants of appendix A.3 (sums over observations => additive over row shards):
    packed [n_d | sy_d | syy_d]."""
    obs = ~np.isnan(Y)
    y0 = np.where(obs, Y, 0.0)
    return np.concatenate([obs.sum(axis=1).astype(np.float64), y0.sum(axis=1), (y0 * y0).sum(axis=1)])


def missed_stats(Y: np.ndarray, mu: np.ndarray, W: np.ndarray, sigma2: float, chunk: int = 8192):
    """Per-shard sufficient statistics of one masked E-step (additive over row shards), packed as
    [S1 (D x M) | S2 (D x M) | G (D x M x M) | sum_i ez_i (M)], plus the shard's EX (M x n)."""
    d, n = Y.shape
    m = W.shape[1]
    obs = ~np.isnan(Y)
    y0 = np.where(obs, Y, 0.0)
    EX = np.empty((m, n))
    S1 = np.zeros((d, m)); S2 = np.zeros((d, m)); G = np.zeros((d, m, m))
    ww = W[:, :, None] * W[:, None, :]
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        o = obs[:, c0:c1].astype(np.float64)
        Mi = np.einsum("dn,dab->nab", o, ww) + sigma2 * np.eye(m)
        b = np.einsum("dn,da->na", o * (y0[:, c0:c1] - mu[:, None]), W)
        Minv = np.linalg.inv(Mi)
        ez = np.einsum("nab,nb->na", Minv, b)
        A = ez[:, :, None] * ez[:, None, :] + sigma2 * Minv
        EX[:, c0:c1] = ez.T
        S1 += o @ ez
        S2 += (o * y0[:, c0:c1]) @ ez
        G += np.einsum("dn,nab->dab", o, A)
    packed = np.concatenate([S1.ravel(), S2.ravel(), G.ravel(), EX.sum(axis=1)])
    return packed, EX


def missed_update_from_stats(packed: np.ndarray, consts: np.ndarray, W: np.ndarray, sigma2: float):
    """Replicated M-step of appendix A.3 from the (all-reduced) statistics and data constants."""
    d, m = W.shape
    S1 = packed[:d * m].reshape(d, m)
    S2 = packed[d * m:2 * d * m].reshape(d, m)
    G = packed[2 * d * m:2 * d * m + d * m * m].reshape(d, m, m)
    n_d, sy, syy = consts[:d], consts[d:2 * d], consts[2 * d:3 * d]
    kk = float(n_d.sum())
    new_mu = (sy - np.einsum("da,da->d", W, S1)) / n_d
    r = S2 - new_mu[:, None] * S1
    new_w = np.linalg.solve(G, r[:, :, None])[:, :, 0]
    c = syy - 2 * new_mu * sy + new_mu ** 2 * n_d
    quad = np.einsum("da,dab,db->d", new_w, G, new_w)
    new_var = float(np.sum(c - 2 * np.einsum("da,da->d", new_w, r) + quad) / kk)
    max_diff_w = max(0.0, float(np.max(np.abs(W - new_w))))
    diff_var = abs(new_var - sigma2)
    return new_mu, new_w, new_var, max_diff_w, diff_var


def missed_step_vectorised(Y: np.ndarray, mu: np.ndarray, W: np.ndarray, sigma2: float,
                           chunk: int = 8192):
    packed, EX = missed_stats(Y, mu, W, sigma2, chunk)
    new_mu, new_w, new_var, max_diff_w, diff_var = missed_update_from_stats(packed, missed_constants(Y), W, sigma2)
    return new_mu, new_w, new_var, max_diff_w, diff_var, dict(EX=EX, known=None)


def missed_loglik_vectorised(Y, new_mu, new_w, new_var, chunk: int = 8192) -> float:
    """Appendix A.4: the #O_i x #O_i determinant/solve reduced to M x M (matrix determinant lemma
    and Woodbury).  Equal to missed_loglik_literal to ~1e-15 relative (tests/test_oracle.py)."""
    d, n = Y.shape
    m = new_w.shape[1]
    obs = ~np.isnan(Y)
    y0 = np.where(obs, Y - new_mu[:, None], 0.0)
    ww = new_w[:, :, None] * new_w[:, None, :]
    acc = 0.0
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        o = obs[:, c0:c1].astype(np.float64)
        cnt = o.sum(axis=0)
        Mi = np.einsum("dn,dab->nab", o, ww) + new_var * np.eye(m)
        L = np.linalg.cholesky(Mi)
        logdet_m = 2.0 * np.log(np.einsum("naa->na", L)).sum(axis=1)
        bvec = np.einsum("dn,da->na", y0[:, c0:c1], new_w)
        t = np.linalg.solve(L, bvec[:, :, None])[:, :, 0]
        yy = (y0[:, c0:c1] ** 2).sum(axis=0)
        quad = (yy - (t * t).sum(axis=1)) / new_var
        logdet_c = (cnt - m) * math.log(new_var) + logdet_m
        acc += float(np.sum(cnt * math.log(2 * math.pi) + logdet_c + quad))
    return -0.5 * acc


def em_steps_missed_vectorised(Y, W0, sigma2_0, stop: Stop) -> PPCAResult:
    Y = np.asarray(Y, dtype=np.float64)
    d, n = Y.shape
    mu = np.zeros(d); W = np.array(W0, dtype=np.float64, copy=True); s2 = float(sigma2_0); it = 0
    while True:
        mu_new, W_new, s2_new, mdw, dv, est = missed_step_vectorised(Y, mu, W, s2)
        if _continue(stop, it, dv, mdw):
            mu, W, s2, it = mu_new, W_new, s2_new, it + 1
            continue
        ll = missed_loglik_vectorised(Y, mu_new, W_new, s2_new)
        restored = missed_restore_literal(mu_new, W_new, s2_new, est["EX"])
        return PPCAResult(False, W_new, s2_new, ll, restored, it + 1)


# ----------------------------------------------------------------------------------------------
# Entry points  (PPCA.hs:32-49, TypeSafe/PPCA.hs:42-79)
# ----------------------------------------------------------------------------------------------

def make_ppca(Y, W0, sigma2_0, stop: Stop, vectorised_missing: bool = False) -> PPCAResult:
    """makePPCA with the RNG draw replaced by explicit (W0, sigma2_0) (PPCA.hs:42-43 stay caller
    side).  Dispatch on ``isNaN (sumAll Y)`` exactly like PPCA.hs:44-48."""
    Y = np.asarray(Y, dtype=np.float64)
    no_missed = not math.isnan(float(np.sum(Y)))
    if no_missed:
        return em_steps_fast(Y, W0, sigma2_0, stop)
    if vectorised_missing:
        return em_steps_missed_vectorised(Y, W0, sigma2_0, stop)
    return em_steps_missed(Y, W0, sigma2_0, stop)


def check_typesafe_dims(Y_shape, W0_shape, desired_dim: int) -> None:
    """The three runtime proofs of makePPCATypeSafe (TypeSafe/PPCA.hs:58-66,72-79), same texts."""
    y1, x1 = Y_shape
    y2, x2 = W0_shape
    if x2 != desired_dim:
        raise ValueError("dimensions x2 and d1 should be equal, but they are not")
    if y1 != y2:
        raise ValueError("dimensions y1 and y2 should be equal, but they are not")
    if x1 < 1:   # One :<: x1 is One <= x1 (TypeSafe/Types.hs:63-76)
        raise ValueError("Input matrix should have at least one column")


def projector(W: np.ndarray) -> np.ndarray:
    """Rotation/sign invariant summary of W used for parity (north_star): W (W^T W)^-1 W^T."""
    return W @ np.linalg.solve(W.T @ W, W.T)
