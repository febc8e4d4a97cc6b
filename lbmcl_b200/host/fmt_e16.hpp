// "%.16e" without printf: the VTI writer's number format (reference lbmcl.hpp:289-333 prints every value with
// std::scientific << std::setprecision(16), i.e. exactly what printf("%.16e") prints).
//
// ASCII output dominates every `-e N` run (~23 characters per value, 1.5 GB per file at
// 256^3), and glibc's printf spends ~0.3 us per value in its arbitrary-precision path.  fmt_e16() produces
// the SAME bytes -- the correctly rounded (ties to even) 17-significant-digit decimal expansion -- with one
// 53 x 256-bit integer product: for v = m * 2^-s the digits are round(m * 10^k / 2^s) with k chosen so that
// the quotient has 17 digits; the remainder decides the rounding exactly.  Values outside the fast range
// (|v| < 1e-45, |v| >= 2^53, subnormals, inf, nan) go through snprintf, so the output is byte-identical for
// every double (tests/test_host_format.py compares millions of values with snprintf).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>

namespace lbm_fmt {

__extension__ typedef unsigned __int128 u128;  // GCC / Clang extension (64 x 64 -> 128-bit products)

struct U256 {
    uint64_t w[4];
};

struct Pow10Table {
    static constexpr int KMAX = 61;  // 53 + ceil(61 * log2(10)) = 256 bits
    static constexpr int DMIN = -46, DMAX = 17;
    U256 p[KMAX + 1];
    int limbs[KMAX + 1];             // non-zero 64-bit words of p[k]
    double dec[DMAX - DMIN + 1];     // 10^d as a double (nearest), d = DMIN .. DMAX: only steers the first guess
    Pow10Table()
    {
        U256 v = {{1, 0, 0, 0}};
        for (int k = 0; k <= KMAX; ++k) {
            p[k] = v;
            limbs[k] = v.w[3] ? 4 : v.w[2] ? 3 : v.w[1] ? 2 : 1;
            u128 carry = 0;
            for (int i = 0; i < 4; ++i) {
                const u128 t = (u128)v.w[i] * 10u + carry;
                v.w[i] = (uint64_t)t;
                carry = t >> 64;
            }
        }
        for (int d = DMIN; d <= DMAX; ++d) dec[d - DMIN] = std::pow(10.0, d);
    }
};

inline const Pow10Table &pow10_table()
{
    static const Pow10Table t;
    return t;
}

// two decimal digits per table entry
inline const char *digit_pairs()
{
    static const char d[] =
        "0001020304050607080910111213141516171819202122232425262728293031323334353637383940414243444546474849"
        "5051525354555657585960616263646566676869707172737475767778798081828384858687888990919293949596979899";
    return d;
}

// Writes printf("%.16e", v) to `out` (no terminator); returns the number of characters (at most 24 for
// finite values; callers should provide 32 bytes).
inline int fmt_e16(double v, char *out)
{
    uint64_t bits;
    std::memcpy(&bits, &v, sizeof bits);
    const bool neg = (bits >> 63) != 0;
    const int bexp = (int)((bits >> 52) & 0x7ff);
    const uint64_t frac = bits & ((1ull << 52) - 1);
    char *p = out;
    if (bexp == 0 && frac == 0) {  // +-0
        if (neg) *p++ = '-';
        std::memcpy(p, "0.0000000000000000e+00", 22);
        return (int)(p - out) + 22;
    }
    const int s = 1075 - bexp;  // v = m * 2^-s
    if (bexp == 0 || bexp == 0x7ff || s < 1 || s > 255) return std::snprintf(out, 32, "%.16e", v);
    const uint64_t m = frac | (1ull << 52);
    // decimal exponent estimate from the binary one: floor(log10(m * 2^-s)), possibly one too small
    // (m in [2^52, 2^53) => log10(v) in [(52 - s) * log10(2), (53 - s) * log10(2))); a comparison with 10^(d+1) as a
    // double settles it except within an ulp of a power of ten, where the range check below corrects the guess
    const Pow10Table &tab = pow10_table();
    int d = ((52 - s) * 78913) >> 18;  // floor((52 - s) * log10(2)), exact for |52 - s| < 1650 (arithmetic shift)
    if (d + 1 >= Pow10Table::DMIN && d + 1 <= Pow10Table::DMAX && std::fabs(v) >= tab.dec[d + 1 - Pow10Table::DMIN]) ++d;
    uint64_t q = 0;
    bool round_bit = false, sticky = false;
    for (int attempt = 0;; ++attempt) {
        const int k = 16 - d;
        if (k < 0 || k > Pow10Table::KMAX || attempt > 2) return std::snprintf(out, 32, "%.16e", v);
        const U256 &pw = tab.p[k];
        // N = m * 10^k  (fits 256 bits for k <= KMAX); only the non-zero words of 10^k are multiplied
        uint64_t n[4] = {0, 0, 0, 0};
        {
            const int nl = tab.limbs[k];
            u128 carry = 0;
            int i = 0;
            for (; i < nl; ++i) {
                const u128 t = (u128)pw.w[i] * m + carry;
                n[i] = (uint64_t)t;
                carry = t >> 64;
            }
            if (i < 4) n[i] = (uint64_t)carry;
        }
        // q = N >> s  (17 digits => fits 64 bits when d is right; detect overflow of that assumption)
        const int li = s >> 6, off = s & 63;
        q = n[li] >> off;
        if (off && li + 1 < 4) q |= n[li + 1] << (64 - off);
        bool high = false;  // bits of N >> s above the 64 we kept
        if (off) {
            if (li + 1 < 4 && (n[li + 1] >> off) != 0) high = true;
            for (int i = li + 2; i < 4; ++i) high = high || n[i] != 0;
        } else {
            for (int i = li + 1; i < 4; ++i) high = high || n[i] != 0;
        }
        if (high || q >= 100000000000000000ull) {  // >= 10^17: d was too small
            ++d;
            continue;
        }
        if (q < 10000000000000000ull) {  // < 10^16: d was too large
            --d;
            continue;
        }
        // remainder: bit s-1 is the rounding bit, everything below is sticky
        const int rb = s - 1, rli = rb >> 6, roff = rb & 63;
        round_bit = ((n[rli] >> roff) & 1u) != 0;
        sticky = roff ? (n[rli] & ((1ull << roff) - 1)) != 0 : false;
        for (int i = 0; i < rli; ++i) sticky = sticky || n[i] != 0;
        break;
    }
    if (round_bit && (sticky || (q & 1u))) {  // round to nearest, ties to even (glibc in the default mode)
        ++q;
        if (q == 100000000000000000ull) {
            q = 10000000000000000ull;
            ++d;
        }
    }
    // 17 digits: D.DDDDDDDDDDDDDDDD
    if (neg) *p++ = '-';
    // q = h dddddddd dddddddd: the two 8-digit halves are converted independently, four digit pairs each
    const char *pairs = digit_pairs();
    const uint64_t top = q / 10000000000000000ull;                // leading digit
    const uint64_t rest = q - top * 10000000000000000ull;         // 16 digits
    const uint32_t a = (uint32_t)(rest / 100000000ull), b = (uint32_t)(rest % 100000000ull);
    const uint32_t a1 = a / 10000, a0 = a % 10000, b1 = b / 10000, b0 = b % 10000;
    p[0] = (char)('0' + top);
    p[1] = '.';
    std::memcpy(p + 2, pairs + 2 * (a1 / 100), 2);
    std::memcpy(p + 4, pairs + 2 * (a1 % 100), 2);
    std::memcpy(p + 6, pairs + 2 * (a0 / 100), 2);
    std::memcpy(p + 8, pairs + 2 * (a0 % 100), 2);
    std::memcpy(p + 10, pairs + 2 * (b1 / 100), 2);
    std::memcpy(p + 12, pairs + 2 * (b1 % 100), 2);
    std::memcpy(p + 14, pairs + 2 * (b0 / 100), 2);
    std::memcpy(p + 16, pairs + 2 * (b0 % 100), 2);
    p += 18;
    *p++ = 'e';
    int ad = d;
    if (ad < 0) {
        *p++ = '-';
        ad = -ad;
    } else {
        *p++ = '+';
    }
    if (ad >= 100) {
        *p++ = (char)('0' + ad / 100);
        ad %= 100;
    }
    *p++ = pairs[2 * ad];
    *p++ = pairs[2 * ad + 1];
    return (int)(p - out);
}

}  // namespace lbm_fmt
