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


def remove_dim(D: int, A):
    """remove_dim(Dim(D), A) (src/utils.jl:27-32); D is 1-based."""
    return tuple(a for i, a in enumerate(A, start=1) if i != D)


def insert_dim(D: int, A, v):
    """insert_dim(Dim(D), A, v) (src/utils.jl:47-51); D is 1-based."""
    A = list(A)
    A.insert(D - 1, v)
    return tuple(A)
