"""oracle/fixedpoint.py -- TEST INFRASTRUCTURE (never imported by the product).

CPU restatement (NumPy + Python integers) of the exact fixed-point arithmetic that the INT8 complement-sum kernels
(gplvmhaskell_b200/csrc/masked_utc.cu, masked_fx.cuh) use for

    miss (D x 152) = Miss^T (D x n) . [vech A | ez] (n x 152)          (reference: the per-feature sums of
                                                                        emStepsMissed.stepTwo, src/GPLVM/PPCA.hs:119-147)

One factor is exactly 0 / 1; every entry v of the other is rounded once to x = rint(v 2^(51 - E)) with |v| < 2^E for the
whole column class, u = x + 2^51 is cut into 7 bytes, byte planes are summed exactly in integers (int32 per CTA on the
device, int64 across CTAs), a constant-1 column counts the offsets, and the 128-bit integer sum is rounded to FP64 once.
PARITY UNPINNED in the sense of DESIGN.md section 2 (there is no reference implementation of this representation; the
reference sums in FP64): these functions pin the DEVICE arithmetic, the parity tests pin the results against the oracle.
"""
from __future__ import annotations

import struct

import numpy as np

NSL = 7      # byte planes
FXP = 51     # |x| < 2^FXP
MAGIC = 6755399441055744.0   # 2^52 + 2^51


def hi_word_abs(v: float) -> int:
    """High 32 bits of |v| (what the solver kernel feeds to atomicMax)."""
    return (struct.unpack("<Q", struct.pack("<d", abs(float(v))))[0] >> 32) & 0x7FFFFFFF


def class_exponent(hmax: int) -> int:
    """Biased exponent of the class maximum, clamped as in masked_fx.cuh."""
    return min(max((hmax >> 20) & 0x7FF, 60), 2040)


def scale_up(hmax: int) -> float:
    return float(2.0 ** (FXP + 1022 - class_exponent(hmax)))


def scale_down(hmax: int) -> float:
    return float(2.0 ** (class_exponent(hmax) - 1022 - FXP))


def to_fixed(v: np.ndarray, hmax: int) -> np.ndarray:
    """u = rint(v * 2^(51 - E)) + 2^51 through the mantissa of v * scale + (2^52 + 2^51), as the kernels do."""
    y = np.asarray(v, dtype=np.float64) * scale_up(hmax) + MAGIC        # one rounding (the device uses an FMA; v * scale is exact)
    bits = y.view(np.uint64)
    return bits - np.uint64(0x4330000000000000)                          # exact also when y == 2^53


def planes(u: np.ndarray) -> np.ndarray:
    """Byte planes s = 0..6 of u (shape (..., 7), uint8)."""
    return np.stack([((u >> np.uint64(8 * s)) & np.uint64(0xFF)).astype(np.uint8) for s in range(NSL)], axis=-1)


def masked_sum(v: np.ndarray, mask: np.ndarray, hmax: int, splits: int = 1) -> float:
    """sum_i mask_i v_i the device way: integer plane sums per split, int64 across splits, one rounding at the end."""
    u = to_fixed(v, hmax)
    pl = planes(u).astype(np.int64)
    m = np.asarray(mask).astype(np.int64)
    S = [0] * NSL
    cnt = 0
    for part, mpart in zip(np.array_split(pl, splits), np.array_split(m, splits)):
        for s in range(NSL):
            S[s] += int((part[:, s] * mpart).sum())
        cnt += int(mpart.sum())
    T = sum(S[s] << (8 * s) for s in range(NSL)) - (cnt << FXP)
    return assemble(T, hmax)


def assemble(T: int, hmax: int) -> float:
    """128-bit integer -> FP64 with one rounding: fma(float(T >> 24), 2^24, float(T & 0xffffff)) * 2^(E - 51)."""
    thi, tlo = T >> 24, T & 0xFFFFFF
    assert abs(thi) < 2 ** 53
    # float(thi) * 2^24 and float(tlo) are exact; their sum is rounded once (the device fuses it in an FMA)
    from fractions import Fraction
    exact = Fraction(thi) * 2 ** 24 + tlo
    return float(exact) * scale_down(hmax)


def expand_order_word(std_word: int) -> int:
    """masked_utc.cu: stored bit 8 b + s  <=>  observation 4 s + b of the 32 covered by the word."""
    out = 0
    for s in range(8):
        for b in range(4):
            out |= ((std_word >> (4 * s + b)) & 1) << (8 * b + s)
    return out


def fragment_order_words(std_words: list[int]) -> list[int]:
    """round-1 mma.sync IMMA kernel (removed in round 2, kept for the layout tests): word t, bit 8 b + 2 ks + h  <=>  observation 32 ks + 16 h + 4 t + b of the 128-observation chunk."""
    out = []
    for t in range(4):
        o = 0
        for ks in range(4):
            for h in range(2):
                for b in range(4):
                    o |= ((std_words[ks] >> (16 * h + 4 * t + b)) & 1) << (8 * b + 2 * ks + h)
        out.append(o)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# Numerical design of the NEXT step: FP64-faithful dense statistics with BOTH factors in byte planes
# (Ozaki splitting) -- emulated exactly with integers; tools/dense_i8_proto.cu measures step 1 of it on tcgen05.
# ---------------------------------------------------------------------------------------------------------------------

def balanced_planes(xi: np.ndarray, n: int) -> list[np.ndarray]:
    """Balanced base-256 digits (each in [-128, 127]) of an int64 array, least significant first."""
    out, x = [], xi.astype(np.int64).copy()
    for _ in range(n):
        lo = ((x & 0xFF) ^ 0x80) - 0x80          # sign-extended low byte
        out.append(lo)
        x = (x - lo) >> 8
    assert np.all(x == 0), "digits overflow"
    return out


def pair_product(Ap: list[np.ndarray], Bp: list[np.ndarray], top_keep: int) -> np.ndarray:
    """sum over plane pairs of 256^(p+q) Ap[p] @ Bp[q], keeping the `top_keep` most significant weight classes p + q
    (what accumulates in one group of s32 TMEM columns per class).  Exact: returns an object array of Python integers."""
    P, Q = len(Ap), len(Bp)
    top = P + Q - 2
    acc = None
    for w in range(top - top_keep + 1, top + 1):
        cls = np.zeros((Ap[0].shape[0], Bp[0].shape[1]), dtype=np.int64)
        for p in range(P):
            q = w - p
            if 0 <= q < Q:
                cls += Ap[p] @ Bp[q]
        assert np.max(np.abs(cls)) < 2 ** 62
        term = cls.astype(object) * (256 ** w)
        acc = term if acc is None else acc + term
    return acc


def _to_float(obj: np.ndarray) -> np.ndarray:
    return np.array([[float(v) for v in row] for row in obj], dtype=np.float64)   # one rounding per entry


def dense_stats_i8(Xc: np.ndarray, W: np.ndarray, sigma2: float, cache: dict) -> np.ndarray:
    """Same packed statistics as oracle.ppca.dense_stats ([XEz | Ez Ez^T]; Xc is D x N), formed the planned device way:
    X' = Xc / c_f (feature maxima folded into the projector) cut ONCE into 7 planes; projector A' = c_f W M^-1 cut per
    step with one scale per latent column; Ez cut into 8 planes with the a-priori bound max_i |x'_i| |A'_a|; every
    product keeps its 7 most significant weight classes."""
    from . import ppca as op
    if "Xp" not in cache:
        c = np.max(np.abs(Xc), axis=1)
        c[c == 0] = 1.0
        xi = np.rint(Xc / c[:, None] * 2.0 ** 54).astype(np.int64)
        cache.update(c=c, Xp=balanced_planes(xi, 7), rown=float(np.sqrt(np.max(np.sum((Xc / c[:, None]) ** 2, axis=0)))))
    c, Xp = cache["c"], cache["Xp"]
    inv_m = np.linalg.inv(op.map_diagonal_add(W.T @ W, sigma2))
    A = (W @ inv_m) * c[:, None]
    sa = np.max(np.abs(A), axis=0)
    Apl = balanced_planes(np.rint(A / sa * 2.0 ** 54).astype(np.int64), 7)
    ez = _to_float(pair_product([x.T for x in Xp], Apl, 7)) * 2.0 ** -108 * sa            # N x M
    se = cache["rown"] * np.sqrt(np.sum(A ** 2, axis=0))                                # |Ez_a| <= |x'| |A'_a|
    Epl = balanced_planes(np.rint(ez / se * 2.0 ** 62).astype(np.int64), 8)
    xez = _to_float(pair_product(Xp, Epl, 7)) * 2.0 ** -(54 + 62) * c[:, None] * se[None, :]
    ee = _to_float(pair_product([e.T for e in Epl], Epl, 7)) * 2.0 ** -124 * se[:, None] * se[None, :]
    return np.concatenate([xez.ravel(), ee.ravel()])


# ---------------------------------------------------------------------------------------------------------------------
# Gram matrices of the masked E-step on the INT8 tensor cores (gplvmhaskell_b200/csrc/masked_gram_tc.cu)
#   M_i = C0 - G_i,  G_i[a][c] = sum_{d missing in observation i} W[d][a] W[d][c]      (emStepsMissed.stepOne, PPCA.hs:106-117)
# ---------------------------------------------------------------------------------------------------------------------
def gram_column_exponents(W: np.ndarray) -> np.ndarray:
    """e_a with 2^e_a > max_d |W[d][a]| (ilogb + 1), 0 for an all-zero column: what gram_planes_kernel computes."""
    cmax = np.max(np.abs(W), axis=0)
    e = np.zeros(W.shape[1], dtype=np.int64)
    nz = cmax > 0
    e[nz] = np.floor(np.log2(cmax[nz])).astype(np.int64) + 1
    # log2 of a float just below a power of two may round up: ilogb semantics are restored by one comparison
    e[nz & (np.ldexp(1.0, (e - 1).astype(np.int32)) > cmax)] -= 1
    return e


def gram_gate_open(W: np.ndarray, guard_bits: int = 20) -> bool:
    """Precision gate of the tensor-core Gram path: at most 1 % of the non-zero entries of W more than 2^guard_bits below their
    column maximum, and column exponents within +-400."""
    e = gram_column_exponents(W)
    a = np.abs(W)
    nz = a > 0
    small = nz & (a < np.ldexp(1.0, (e - 1 - guard_bits).astype(np.int32))[None, :])
    return bool(np.all(np.abs(e) < 400) and small.sum() * 100 <= nz.sum())


def gram_fixed_point(W: np.ndarray, miss: np.ndarray) -> np.ndarray:
    """G[a][c] for ONE observation (miss: boolean over the D features) the device way: every product W[d][a] W[d][c] is
    computed in FP64 (one rounding), scaled by the exact power of two 2^(51 - e_a - e_c), rounded to an integer x, offset to
    u = x + 2^51 and cut into 7 byte planes; the planes of the missing features are summed exactly, the integer
    sum_s 256^s S_s - 2^51 cnt is converted to FP64 with one rounding and scaled back by 2^(e_a + e_c - 51)."""
    D, M = W.shape
    e = gram_column_exponents(W)
    G = np.zeros((M, M))
    cnt = int(np.count_nonzero(miss))
    Wm = W[np.asarray(miss, dtype=bool)]
    for a in range(M):
        for c in range(a + 1):
            prod = Wm[:, a] * Wm[:, c]                                         # FP64 product, as on the device
            y = prod * float(np.ldexp(1.0, FXP - int(e[a]) - int(e[c]))) + MAGIC   # device: one FMA; the scaling is exact
            u = y.view(np.uint64) - np.uint64(0x4330000000000000)
            S = planes(u).astype(np.int64).sum(axis=0)                       # exact plane sums (s32 accumulators in TMEM)
            T = sum(int(S[s]) << (8 * s) for s in range(NSL)) - (cnt << FXP)   # exact 64-bit integer
            G[a, c] = G[c, a] = float(T) * float(np.ldexp(1.0, int(e[a]) + int(e[c]) - FXP))
    return G
