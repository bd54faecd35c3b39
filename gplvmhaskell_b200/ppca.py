"""Host-side mirror of the reference's PPCA interface over the C ABI (include/gplvm_b200.h).

Reference surface mirrored (paths relative to the reference checkout):
  src/GPLVM/PPCA.hs:21-30            record PPCA            -> class PPCA (same field names)
  src/GPLVM/PPCA.hs:32-49            makePPCA               -> make_ppca / makePPCA
  src/GPLVM/TypeSafe/PPCA.hs:42-79   makePPCATypeSafe       -> make_ppca_type_safe / makePPCATypeSafe
  src/GPLVM/TypeSafe/Types.hs:40-49  record PPCA (typed)    -> same class, field ``desiredDimensions``

The reference draws (W0, sigma2_0) from its ``gen`` argument (PPCA.hs:42-43: normals for W0 and a raw
StdGen Int for the variance).  Haskell's StdGen cannot be reproduced here, so ``gen`` is either a
``(W0, sigma2_0)`` pair (what the Haskell shim passes) or a numpy Generator / int seed from
which a W0 ~ N(0,1) and a StdGen-like sigma2_0 in [1, 2147483562] are drawn.

All arithmetic runs in libgplvm_b200.so on the GPU; this module only validates, marshals and dispatches.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple, Union

import numpy as np

from . import _lib

Stop = Union[Tuple[str, Union[int, float]], int, float]


@dataclass
class PPCA:
    """The reference record (PPCA.hs:21-30 / TypeSafe/Types.hs:40-49)."""
    _noMissedData: bool
    _learningData: np.ndarray
    desiredDimension: int
    stopParameter: Tuple[str, Union[int, float]]
    _variance: float
    _W: np.ndarray
    _finalExpLikelihood: float
    _restoredMatrix: Optional[np.ndarray]
    steps_done: int = 0

    @property
    def desiredDimensions(self) -> int:  # spelling of the typed record
        return self.desiredDimension


def Left(k: int):
    """Either Int Double, Left: number of iterations (runs k+1 EM steps, PPCA.hs:91)."""
    return ("left", int(k))


def Right(tol: float):
    """Either Int Double, Right: stop once max(|dsigma2|, max|dW|) <= tol (PPCA.hs:92)."""
    return ("right", float(tol))


def _normalise_stop(stop: Stop):
    if isinstance(stop, tuple):
        kind, val = stop
        kind = kind.lower()
        if kind not in ("left", "right"):
            raise ValueError("stop parameter must be Left k or Right tol")
        return (kind, int(val)) if kind == "left" else (kind, float(val))
    if isinstance(stop, (int, np.integer)):
        return ("left", int(stop))
    return ("right", float(stop))


def _draw_init(gen, d: int, m: int):
    if isinstance(gen, tuple) and len(gen) == 2:
        w0 = np.ascontiguousarray(np.asarray(gen[0], dtype=np.float64))
        return w0, float(gen[1])
    rng = gen if isinstance(gen, np.random.Generator) else np.random.default_rng(gen)
    w0 = rng.standard_normal((d, m))                       # randomMatrixD, Util.hs:139-151
    s2 = float(rng.integers(1, 2147483563))                 # fst (next gen), PPCA.hs:43
    return w0, s2


def _run(Y: np.ndarray, M: int, stop, W0: np.ndarray, s20: float, which: str):
    lib = _lib.load()
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    if Y.ndim != 2:
        raise ValueError("learning vectors must be a D x N matrix")
    D, N = Y.shape
    W0 = np.ascontiguousarray(W0, dtype=np.float64)
    # the C ABI copies D * M doubles from W0: a smaller buffer would be read out of bounds
    if W0.shape != (D, int(M)):
        raise ValueError("W0 must be D x M = %d x %d, got %s" % (D, int(M), "x".join(map(str, W0.shape))))
    kind, val = stop
    stop_kind = _lib.STOP_ITERATIONS if kind == "left" else _lib.STOP_TOLERANCE
    k = int(val) if kind == "left" else 0
    tol = float(val) if kind == "right" else 0.0
    W = np.empty((D, M), dtype=np.float64)
    s2 = C.c_double()
    ll = C.c_double()
    steps = C.c_int64()
    restored = None
    if which == "dense":
        rc = lib.gplvm_ppca_dense(_lib.dptr(Y), D, N, M, _lib.dptr(W0), s20, stop_kind, k, tol, _lib.dptr(W),
                                  C.byref(s2), C.byref(ll), C.byref(steps))
        no_missed = True
    elif which == "missing":
        restored = np.empty((D, N), dtype=np.float64)
        rc = lib.gplvm_ppca_missing(_lib.dptr(Y), D, N, M, _lib.dptr(W0), s20, stop_kind, k, tol, _lib.dptr(W),
                                    C.byref(s2), C.byref(ll), _lib.dptr(restored), C.byref(steps))
        no_missed = False
    else:
        restored = np.empty((D, N), dtype=np.float64)
        nm = C.c_int()
        rc = lib.gplvm_ppca(_lib.dptr(Y), D, N, M, _lib.dptr(W0), s20, stop_kind, k, tol, _lib.dptr(W), C.byref(s2),
                            C.byref(ll), _lib.dptr(restored), C.byref(nm), C.byref(steps))
        no_missed = bool(nm.value)
        if no_missed:
            restored = None
    _lib.check(rc)
    return no_missed, W, s2.value, ll.value, restored, steps.value


def em_steps_fast(Y, W0, sigma2_0, stop: Stop):
    """emStepsFast (PPCA.hs:54-94): returns (W, variance, expLogLikelihood, None)."""
    W0 = np.asarray(W0, dtype=np.float64)
    if W0.ndim != 2:
        raise ValueError("W0 must be a D x M matrix")
    _, W, s2, ll, _, steps = _run(Y, W0.shape[1], _normalise_stop(stop), W0, sigma2_0, "dense")
    return W, s2, ll, None


def em_steps_missed(Y, W0, sigma2_0, stop: Stop):
    """emStepsMissed (PPCA.hs:96-182): returns (W, variance, expLogLikelihood, restored)."""
    W0 = np.asarray(W0, dtype=np.float64)
    if W0.ndim != 2:
        raise ValueError("W0 must be a D x M matrix")
    _, W, s2, ll, restored, steps = _run(Y, W0.shape[1], _normalise_stop(stop), W0, sigma2_0, "missing")
    return W, s2, ll, restored


def make_ppca(learning_vectors, desired_dimension: int, stop_parameter: Stop, gen) -> PPCA:
    """makePPCA (PPCA.hs:32-49).  ``learning_vectors`` is D x N (rows = features), as in the reference."""
    Y = np.asarray(learning_vectors, dtype=np.float64)
    if Y.ndim != 2:
        raise ValueError("learning vectors must be a D x N matrix")
    stop = _normalise_stop(stop_parameter)
    W0, s20 = _draw_init(gen, Y.shape[0], int(desired_dimension))
    no_missed, W, s2, ll, restored, steps = _run(Y, int(desired_dimension), stop, W0, s20, "auto")
    return PPCA(no_missed, Y, int(desired_dimension), stop, s2, W, ll, restored, steps)


def check_ppca_input_matrices_prop(y_shape, w0_shape, desired_dimensions: int) -> None:
    """checkPPCAInputMatricesProp + the three error texts of makePPCATypeSafe (TypeSafe/PPCA.hs:58-66,72-79).
    Raised BEFORE any FFI call, like the Haskell side does."""
    y1, x1 = y_shape
    y2, x2 = w0_shape
    if x2 != desired_dimensions:
        raise ValueError("dimensions x2 and d1 should be equal, but they are not")
    if y1 != y2:
        raise ValueError("dimensions y1 and y2 should be equal, but they are not")
    if x1 < 1:
        raise ValueError("Input matrix should have at least one column")


def with_list_of_indexes(y: int, indexes, f):
    """withListOfIndexes (TypeSafe/PPCA.hs:81-89): run ``f`` on the index list unless an index exceeds the type-level
    bound ``y`` (the reference compares with ``<``, so ``y`` itself passes); same error text."""
    indexes = list(indexes)
    if not indexes:
        return f([])
    if y < max(indexes):
        raise IndexError("index is out of range")
    return f(indexes)


withListOfIndexes = with_list_of_indexes


def make_ppca_type_safe(learning_vectors, desired_dimensions: int, stop_parameter: Stop, gen) -> PPCA:
    """makePPCATypeSafe (TypeSafe/PPCA.hs:42-70): same arithmetic, dimension proofs first."""
    Y = np.asarray(learning_vectors, dtype=np.float64)
    if Y.ndim != 2:
        raise ValueError("learning vectors must be a D x N matrix")
    stop = _normalise_stop(stop_parameter)
    W0, s20 = _draw_init(gen, Y.shape[0], int(desired_dimensions))
    check_ppca_input_matrices_prop(Y.shape, W0.shape, int(desired_dimensions))
    no_missed, W, s2, ll, restored, steps = _run(Y, int(desired_dimensions), stop, W0, s20, "auto")
    return PPCA(no_missed, Y, int(desired_dimensions), stop, s2, W, ll, restored, steps)


# reference spellings
makePPCA = make_ppca
makePPCATypeSafe = make_ppca_type_safe
emStepsFast = em_steps_fast
emStepsMissed = em_steps_missed


class Session:
    """Observation shards resident in HBM (gplvm_session_* of include/gplvm_b200.h).  Used by bench.py and
    by multi-process launches; the one-shot calls above are built on the same C functions."""

    def __init__(self, D: int, N: int, M: int, missing: bool, obs_begin: int = 0, obs_count: Optional[int] = None):
        self.lib = _lib.load()
        self.D, self.N, self.M, self.missing = int(D), int(N), int(M), bool(missing)
        self.obs_begin = int(obs_begin)
        self.obs_count = int(N if obs_count is None else obs_count)
        h = C.c_void_p()
        _lib.check(self.lib.gplvm_session_create(self.D, self.N, self.M, int(self.missing), self.obs_begin,
                                                 self.obs_count, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.lib.gplvm_session_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_host(self, Y: np.ndarray, ld: Optional[int] = None):
        """Y: D x obs_count block (or a view into a D x ld matrix starting at this rank's first observation)."""
        Y = np.asarray(Y, dtype=np.float64)
        if Y.ndim != 2 or Y.shape[0] != self.D:
            raise ValueError("Y must have D = %d rows" % self.D)
        if ld is None:
            Y = np.ascontiguousarray(Y)
            ld = Y.shape[1]
            if ld != self.obs_count:
                raise ValueError("Y must be D x obs_count = %d x %d" % (self.D, self.obs_count))
        elif ld < self.obs_count or Y.strides != (8 * ld, 8) or Y.shape[1] < self.obs_count:
            raise ValueError("Y must be a row-major view with row stride ld >= obs_count covering this rank's block")
        _lib.check(self.lib.gplvm_session_load_host(self.h, _lib.dptr(Y), int(ld)))

    def synth(self, seed: int, nan_prob: float = 0.0):
        _lib.check(self.lib.gplvm_session_synth(self.h, int(seed), float(nan_prob)))

    def read_data(self) -> np.ndarray:
        Y = np.empty((self.D, self.obs_count), dtype=np.float64)
        _lib.check(self.lib.gplvm_session_read_data(self.h, _lib.dptr(Y), self.obs_count))
        return Y

    def init(self, W0: np.ndarray, sigma2_0: float):
        W0 = np.ascontiguousarray(W0, dtype=np.float64)
        if W0.shape != (self.D, self.M):
            raise ValueError("W0 must be D x M")
        _lib.check(self.lib.gplvm_session_init(self.h, _lib.dptr(W0), float(sigma2_0)))

    def steps(self, n: int):
        _lib.check(self.lib.gplvm_session_steps(self.h, int(n)))

    def run(self, stop: Stop) -> int:
        kind, val = _normalise_stop(stop)
        done = C.c_int64()
        _lib.check(self.lib.gplvm_session_run(self.h, _lib.STOP_ITERATIONS if kind == "left" else _lib.STOP_TOLERANCE,
                                              int(val) if kind == "left" else 0, float(val) if kind == "right" else 0.0,
                                              C.byref(done)))
        return done.value

    def sync(self):
        _lib.check(self.lib.gplvm_session_sync(self.h))

    def params(self):
        W = np.empty((self.D, self.M), dtype=np.float64)
        mu = np.empty(self.D, dtype=np.float64)
        s2, dv, mdw = C.c_double(), C.c_double(), C.c_double()
        steps = C.c_int64()
        _lib.check(self.lib.gplvm_session_params(self.h, _lib.dptr(W), C.byref(s2), _lib.dptr(mu), C.byref(dv),
                                                 C.byref(mdw), C.byref(steps)))
        return dict(W=W, sigma2=s2.value, mu=mu, dsigma2=dv.value, max_dw=mdw.value, steps=steps.value)

    def loglik(self) -> float:
        ll = C.c_double()
        _lib.check(self.lib.gplvm_session_loglik(self.h, C.byref(ll)))
        return ll.value

    def restored(self) -> np.ndarray:
        R = np.empty((self.D, self.obs_count), dtype=np.float64)
        _lib.check(self.lib.gplvm_session_restored(self.h, _lib.dptr(R), self.obs_count))
        return R

    def set_kernel_timing(self, on: bool):
        _lib.check(self.lib.gplvm_session_set_kernel_timing(self.h, int(on)))

    def last_segments_ms(self):
        """[(segment name, ms)] of the last EM step enqueued with kernel timing on."""
        buf = (C.c_double * 16)()
        n = C.c_int()
        _lib.check(self.lib.gplvm_session_last_segments_ms(self.h, buf, 16, C.byref(n)))
        return [(self.lib.gplvm_session_segment_name(self.h, i).decode(), buf[i]) for i in range(n.value)]

    def masked_solver_path(self) -> int:
        """2: tcgen05 fixed-point Gram solver, 1: FP64 DMMA Gram solver, 0: generic path, -1: no masked step yet."""
        p = C.c_int()
        _lib.check(self.lib.gplvm_session_masked_solver_path(self.h, C.byref(p)))
        return p.value

    def last_steps_ms(self):
        t, k = C.c_double(), C.c_double()
        _lib.check(self.lib.gplvm_session_last_steps_ms(self.h, C.byref(t), C.byref(k)))
        return t.value, k.value
