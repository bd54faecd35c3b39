"""oracle/fixedpoint.py -- TEST INFRASTRUCTURE (never imported by the product).

CPU restatement (NumPy + Python integers) of the exact fixed-point arithmetic that the INT8 complement-sum kernels
(gplvmhaskell_b200/csrc/masked_imma.cu, masked_utc.cu, masked_fx.cuh) use for

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
    """masked_imma.cu: word t, bit 8 b + 2 ks + h  <=>  observation 32 ks + 16 h + 4 t + b of the 128-observation chunk."""
    out = []
    for t in range(4):
        o = 0
        for ks in range(4):
            for h in range(2):
                for b in range(4):
                    o |= ((std_words[ks] >> (16 * h + 4 * t + b)) & 1) << (8 * b + 2 * ks + h)
        out.append(o)
    return out
