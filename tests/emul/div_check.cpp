// div_check.cpp -- TEST INFRASTRUCTURE.  The exact-division sequence of the tuned / fused kernels (div_u in
// chmy.jl_b200/csrc/fused_sv.cuh: the host twin of fast_common.cuh's device function) against IEEE division on the host, over
// the operand families the device self-test uses (chmy_selftest_division): random significands over 120 binades, exact
// multiples of the divisor and their 1-ulp neighbours, and operands whose quotient lies next to a rounding midpoint.  Returns the number of operands whose quotient differs in any bit.
#include <cstdint>
#include <cstring>

#include "../../chmy.jl_b200/csrc/fused_sv.cuh"

static inline double bits_to_double(uint64_t b) {
    double x;
    memcpy(&x, &b, sizeof(x));
    return x;
}
static inline uint64_t double_to_bits(double x) {
    uint64_t b;
    memcpy(&b, &x, sizeof(b));
    return b;
}

extern "C" long long div_check(double c, long long n, unsigned long long seed, int mode) {
    const DivC d = divc_of(c);
    long long bad = 0;
    for (long long t = 0; t < n; ++t) {
        uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(t + 1);   // splitmix64
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        double x;
        if (mode == 0) {
            const uint64_t e = 1023ull - 60ull + (z >> 52) % 121ull;
            x = bits_to_double((z & 0x800FFFFFFFFFFFFFull) | (e << 52));
        } else if (mode == 1) {
            const double m = (double)(long long)(z >> 12);
            x = m * c;
            if (z & 1) x = bits_to_double(double_to_bits(x) + ((z >> 1) & 3) - 1);
        } else {           // quotients next to the midpoint of two neighbouring doubles: x ~ c * (q + ulp(q) / 2), and its neighbours
            const uint64_t e = 1023ull - 30ull + (z >> 52) % 61ull;
            const double q = bits_to_double((z & 0x800FFFFFFFFFFFFFull) | (e << 52));
            const double h = bits_to_double(((e - 53ull) << 52));                  // half an ulp of q
            const long double xl = (long double)c * ((long double)q + (long double)(q < 0 ? -h : h));
            x = (double)xl;
            x = bits_to_double(double_to_bits(x) + ((z >> 3) % 5) - 2);
        }
        const double a = x / c, b = div_u<false>(x, d);
        if (double_to_bits(a) != double_to_bits(b) && !(a != a && b != b)) ++bad;
    }
    return bad;
}
