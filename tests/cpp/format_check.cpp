// fmt_e16() against printf("%.16e") -- byte for byte -- over random and adversarial doubles.
// Usage: format_check [n_random]   (exit status 0 = identical everywhere; prints a throughput line)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../lbmcl_b200/host/fmt_e16.hpp"

static long g_bad = 0, g_n = 0;

static void check(double v)
{
    char a[64], b[64];
    const int na = lbm_fmt::fmt_e16(v, a);
    const int nb = std::snprintf(b, sizeof b, "%.16e", v);
    ++g_n;
    if (na != nb || std::memcmp(a, b, (size_t)nb) != 0) {
        if (g_bad < 20) {
            a[na] = 0;
            std::printf("MISMATCH %a: got '%s' want '%s'\n", v, a, b);
        }
        ++g_bad;
    }
}

int main(int argc, char **argv)
{
    const long n = argc > 1 ? std::atol(argv[1]) : 2000000;
    std::mt19937_64 rng(12345);
    // every bit pattern class: random 64-bit patterns (all exponents, nan, inf, subnormals)
    for (long i = 0; i < n / 4; ++i) {
        uint64_t u = rng();
        double v;
        std::memcpy(&v, &u, sizeof v);
        check(v);
    }
    // the values a cavity run prints: densities near 1, velocities down to 1e-30, floats widened to double
    std::uniform_real_distribution<double> mant(1.0, 10.0);
    std::uniform_int_distribution<int> ex(-50, 20);
    for (long i = 0; i < n / 4; ++i) {
        const double v = mant(rng) * std::pow(10.0, ex(rng));
        check(v);
        check(-v);
        check((double)(float)v);
        check(1.0 + (mant(rng) - 5.0) * 0.02);
    }
    for (long i = 0; i < n / 4; ++i) {
        uint32_t u = (uint32_t)rng();
        float f;
        std::memcpy(&f, &u, sizeof f);
        check((double)f);
    }
    // ties and near-ties: dyadic rationals with few bits, powers of ten and their neighbours, integers
    for (int e = -80; e <= 60; ++e)
        for (int k = 1; k < 4096; k += 1) check(std::ldexp((double)k, e));
    for (int e = -60; e <= 25; ++e) {
        const double p = std::pow(10.0, e);
        check(p);
        check(std::nextafter(p, 0.0));
        check(std::nextafter(p, 1e300));
        check(9.9999999999999995 * p);
        check(9.99999999999999995 * p);
    }
    for (long i = 0; i < 100000; ++i) check((double)i);
    const double specials[] = {0.0, -0.0, 1.0, -1.0, 0.05, 0.0500000007450580597, 5e-324, 1.7976931348623157e308,
                               2.2250738585072014e-308, 1e-45, 9.9e-46, 9007199254740992.0, 4503599627370496.0,
                               4503599627370495.5, 0.1, 0.2, 0.3, 1.0 / 3.0, 123456789.125, 1e16, 1e17, 99999999999999999.0};
    for (double v : specials) {
        check(v);
        check(-v);
    }
    check(std::nan(""));
    check(-std::nan(""));
    check(1.0 / 0.0 * 1.0);
    // throughput on cavity-like values
    std::vector<double> vals(1 << 20);
    for (double &v : vals) v = (double)(float)(1.0 + (mant(rng) - 5.0) * 0.02);
    char buf[64];
    long sink = 0;
    auto t0 = std::chrono::steady_clock::now();
    for (double v : vals) sink += lbm_fmt::fmt_e16(v, buf);
    auto t1 = std::chrono::steady_clock::now();
    for (double v : vals) sink += std::snprintf(buf, sizeof buf, "%.16e", v);
    auto t2 = std::chrono::steady_clock::now();
    const double a = std::chrono::duration<double, std::nano>(t1 - t0).count() / vals.size();
    const double b = std::chrono::duration<double, std::nano>(t2 - t1).count() / vals.size();
    std::printf("checked %ld values, %ld mismatches; fmt_e16 %.1f ns/value, snprintf %.1f ns/value (%.1fx) [%ld]\n", g_n, g_bad, a, b,
                b / a, sink);
    return g_bad == 0 ? 0 : 1;
}
