"""Dim / Side / exact scalar helpers  (reference: src/utils.jl:7-72)."""
from __future__ import annotations

from fractions import Fraction

AXES = ("x", "y", "z")


class Dim(int):
    """Dim(D) with the reference's 1-based D (src/utils.jl:7-9)."""


class Side(int):
    """Side(1) = left, Side(2) = right (src/utils.jl:16-20)."""


Left, Right = Side(1), Side(2)


def fma(a: float, b: float, c: float) -> float:
    """Correctly rounded a*b+c (Julia `muladd` on FMA hardware); exact rational arithmetic, one rounding.
    Host-side scalars only (coordinates, sub-grid origins) -- never on the device path."""
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def fma_t(a, b, c, dtype):
    """muladd in the element type `dtype` (numpy float64 | float32): the exact a*b+c rounded ONCE to that type."""
    import numpy as np
    dtype = np.dtype(dtype)
    x = Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c))
    d = float(x)                                   # nearest binary64
    if dtype == np.float64:
        return np.float64(d)
    f = np.float32(d)
    if Fraction(d) == x:                           # exact in binary64: one (ties-to-even) rounding to binary32
        return f
    # d was itself rounded: pick the binary32 neighbour nearest to the exact value (no double rounding)
    cands = (np.nextafter(f, np.float32(-np.inf)), f, np.nextafter(f, np.float32(np.inf)))
    return min(cands, key=lambda t: abs(Fraction(float(t)) - x))


def remove_dim(D: int, A):
    """remove_dim(Dim(D), A) (src/utils.jl:27-32); D is 1-based."""
    return tuple(a for i, a in enumerate(A, start=1) if i != D)


def insert_dim(D: int, A, v):
    """insert_dim(Dim(D), A, v) (src/utils.jl:47-51); D is 1-based."""
    A = list(A)
    A.insert(D - 1, v)
    return tuple(A)


class DoubleBuffer:
    """DoubleBuffer{T} (src/DoubleBuffers.jl:5-16): two handles and a swap.  (The fused sweeps ping-pong INSIDE the
    library, behind one Field handle; this is the user-level helper the reference exports.)"""

    def __init__(self, front, back):
        self.front, self.back = front, back


def swap_(db: DoubleBuffer):
    db.front, db.back = db.back, db.front
    return db.front, db.back


def front(db: DoubleBuffer):
    return db.front


def back(db: DoubleBuffer):
    return db.back
