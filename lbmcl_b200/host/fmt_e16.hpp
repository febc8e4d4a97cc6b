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
    U256 p[KMAX + 1];
    Pow10Table()
    {
        U256 v = {{1, 0, 0, 0}};
        for (int k = 0; k <= KMAX; ++k) {
            p[k] = v;
            u128 carry = 0;
            for (int i = 0; i < 4; ++i) {
                const u128 t = (u128)v.w[i] * 10u + carry;
                v.w[i] = (uint64_t)t;
                carry = t >> 64;
            }
        }
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
    // (m in [2^52, 2^53) => log10(v) in [(52 - s) * log10(2), (53 - s) * log10(2)))
    int d = (int)std::floor((52 - s) * 0.30102999566398120);
    uint64_t q = 0;
    bool round_bit = false, sticky = false;
    for (int attempt = 0;; ++attempt) {
        const int k = 16 - d;
        if (k < 0 || k > Pow10Table::KMAX || attempt > 2) return std::snprintf(out, 32, "%.16e", v);
        const U256 &pw = pow10_table().p[k];
        // N = m * 10^k  (fits 256 bits for k <= KMAX)
        uint64_t n[4];
        u128 carry = 0;
        for (int i = 0; i < 4; ++i) {
            const u128 t = (u128)pw.w[i] * m + carry;
            n[i] = (uint64_t)t;
            carry = t >> 64;
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
    char dig[17];
    const char *pairs = digit_pairs();
    uint64_t hi = q / 100000000ull;           // 9 digits
    uint32_t lo = (uint32_t)(q % 100000000ull);  // 8 digits
    for (int i = 15; i >= 9; i -= 2) {
        const uint32_t r = lo % 100;
        lo /= 100;
        dig[i] = pairs[2 * r];
        dig[i + 1] = pairs[2 * r + 1];
    }
    uint32_t h = (uint32_t)hi;
    for (int i = 7; i >= 1; i -= 2) {
        const uint32_t r = h % 100;
        h /= 100;
        dig[i] = pairs[2 * r];
        dig[i + 1] = pairs[2 * r + 1];
    }
    dig[0] = (char)('0' + h);
    *p++ = dig[0];
    *p++ = '.';
    std::memcpy(p, dig + 1, 16);
    p += 16;
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
