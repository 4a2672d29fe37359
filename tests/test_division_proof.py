"""
CPU-only: the decision procedure behind the two-operation exact division (chmy.jl_b200/csrc/fast_common.cuh: div2_exact).

The sequence  q = RN(x * rc + RN(x * rl))  with rc = RN(1/c), rl = RN(1/c - rc)  is used by the fused sweeps only for divisors
for which a search over the operands whose quotient lies next to a rounding midpoint finds no failure.  Two things are checked
here, with exact rational arithmetic:
  * SOUNDNESS of the method, exhaustively, in a toy binary format of p bits: whenever the search (restated below for any p)
    says "exact for every operand", brute force over all 2^(p-1) significands agrees -- for every divisor of the format;
  * the C function is that search at p = 53: it agrees with the restatement on random divisors.
"""
import random
from fractions import Fraction

import pytest


def rn(v: Fraction, p: int) -> Fraction:
    """round to nearest, ties to even, to p significant bits (unbounded exponent)"""
    if v == 0:
        return v
    s, a = (-1 if v < 0 else 1), abs(v)
    e = a.numerator.bit_length() - a.denominator.bit_length()
    if Fraction(2) ** e > a:
        e -= 1
    if Fraction(2) ** (e + 1) <= a:
        e += 1
    ulp = Fraction(2) ** (e - p + 1)
    k = a / ulp                              # in [2^(p-1), 2^p)
    n = k.numerator // k.denominator
    r = k - n
    if r > Fraction(1, 2) or (r == Fraction(1, 2) and n % 2 == 1):
        n += 1
    return s * n * ulp


def two_op(x: Fraction, c: Fraction, p: int) -> Fraction:
    rc = rn(1 / c, p)
    rl = rn(1 / c - rc, p)
    return rn(x * rc + rn(x * rl, p), p)


def search_says_exact(c_sig: int, p: int) -> bool:
    """the search of div2_exact for a divisor with integer significand c_sig (p bits), any p >= 6"""
    M = c_sig
    while M % 2 == 0:
        M //= 2
    if M == 1:
        return True
    K = M >> (p - 3)                         # |N| <= M 2^-(p-3): twice the sequence's error bound 2^-(p-1) (1 + 2^-(p-2))
    if K == 0:
        return True
    L = M.bit_length()
    c = Fraction(c_sig)
    lo, hi = 1 << (p - 1), 1 << p
    for s in (L - 1, L):
        inv = pow(pow(2, s + 1, M), -1, M)
        for N in range(1, K + 1, 2):
            for r in (N % M, (-N) % M):
                X = (r * inv) % M
                if X < lo:
                    X += (lo - X + M - 1) // M * M
                while X < hi:
                    if two_op(Fraction(X), c, p) != rn(Fraction(X) / c, p):
                        return False
                    X += M
    return True


@pytest.mark.parametrize("p", [7, 8, 9, 10])
def test_the_search_is_sound_in_a_toy_format(p):
    lo, hi = 1 << (p - 1), 1 << p
    proven = refused = wrongly_refused = 0
    for c_sig in range(lo, hi):
        if c_sig == hi - 1:
            continue                          # the all-ones significand Markstein's conditions (and markstein_ok) exclude
        c = Fraction(c_sig)
        brute = all(two_op(Fraction(X), c, p) == rn(Fraction(X) / c, p) for X in range(lo, hi))
        said = search_says_exact(c_sig, p)
        if said:
            assert brute, (p, c_sig)          # soundness: "exact" is never claimed for a divisor with a failing operand
            proven += 1
        else:
            refused += 1
            wrongly_refused += brute          # the search only refuses when it has SEEN a failing operand
    assert wrongly_refused == 0 and proven > 0.6 * (hi - lo), (proven, refused, wrongly_refused)


def test_the_library_function_is_that_search_at_53_bits():
    import chmy_b200
    rng = random.Random(12)
    n_refused = 0
    for _ in range(1500):
        sig = rng.getrandbits(52) | (1 << 52)
        if sig == (1 << 53) - 1:
            continue
        c = float(sig) * 2.0 ** rng.randint(-60, 8)
        want = search_says_exact(sig, 53)
        assert chmy_b200.division_two_op_exact(c) == want, (c, want)
        n_refused += not want
    assert n_refused > 0                      # ~1.3 % of divisors have a failing operand: the sample holds some


@pytest.mark.parametrize("p", [7, 8, 9, 10])
def test_the_four_operation_sequence_is_exact_for_every_divisor_markstein_admits(p):
    """q = RN(x rc + RN(x rl)); r = x - c q (exact in an fma); RN(q + r rc) == RN(x / c) -- exhaustively in a toy format, for
    every divisor but the all-ones significand (markstein_ok) and every operand"""
    lo, hi = 1 << (p - 1), 1 << p
    for c_sig in range(lo, hi - 1):
        c = Fraction(c_sig)
        rc = rn(1 / c, p)
        rl = rn(1 / c - rc, p)
        for X in range(lo, hi):
            x = Fraction(X)
            q = rn(x * rc + rn(x * rl, p), p)
            r = x - c * q
            assert rn(r, p) == r, (p, c_sig, X)            # the residual of a faithful quotient is representable
            assert rn(q + r * rc, p) == rn(x / c, p), (p, c_sig, X)
