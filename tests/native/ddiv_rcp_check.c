/* CPU check of the float64 division used by decode_one (flashe_kernels.cu, ddiv_rn_known_rcp):
 * with y = RN(1/d), two FMA residual corrections of q0 = RN(n*y) must equal the IEEE quotient n/d.
 * Domain = what decode computes (jzf_quantize.py:102-107): n = v * two_an with v an integer below 2^32
 * (or any double, second half of the cases), d = (2^e - 1) * n_clients, plus unstructured doubles with
 * exponents in [-400, 400].  Test infrastructure only.
 * usage: ddiv_rcp_check <cases_log2> <seed>   -> prints the number of mismatches, exit 1 if any. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t s[2];
static uint64_t rnd(void) {  /* xorshift128+ */
    uint64_t a = s[0], b = s[1];
    s[0] = b; a ^= a << 23; s[1] = a ^ b ^ (a >> 17) ^ (b >> 26);
    return s[1] + b;
}
static double from_bits(uint64_t u) { double f; memcpy(&f, &u, 8); return f; }
static double rnd_double(int emin, int emax, uint64_t style) {
    uint64_t m = rnd() & 0xfffffffffffffull;
    if (style == 1) m = 0xfffffffffffffull - (m & 0xffff);   /* mantissa near all ones */
    if (style == 2) m &= 0xffff;                             /* near a power of two */
    uint64_t e = (uint64_t)(1023 + emin + (int)(rnd() % (uint64_t)(emax - emin + 1)));
    return from_bits((e << 52) | m);
}

static double div_alg(double n, double d, double y) {
    double q = n * y;
    double r = fma(-d, q, n);
    q = fma(r, y, q);
    r = fma(-d, q, n);
    return fma(r, y, q);
}

int main(int argc, char** argv) {
    int lg = argc > 1 ? atoi(argv[1]) : 24;
    s[0] = 0x9E3779B97F4A7C15ull ^ (argc > 2 ? strtoull(argv[2], 0, 10) : 1); s[1] = 0xD1B54A32D192ED03ull;
    uint64_t bad = 0, cases = 1ull << lg;
    for (uint64_t i = 0; i < cases; ++i) {
        double n, d;
        const uint64_t w = rnd();
        if (i & 1) {                                   /* decode-shaped: integer * alpha scale over (2^e - 1) * clients */
            const int e = 1 + (int)(w % 24);
            const uint64_t nc = 1 + ((w >> 8) % ((i & 2) ? 4096 : 64));
            d = (double)(((1ull << e) - 1) * nc);
            uint64_t v = (w >> 24) & 0xffffffffull;
            if ((i & 12) == 4) v &= (1ull << (e + 6)) - 1;    /* sums of e-bit values */
            if ((i & 12) == 8) v = 0xffffffffull - (v & 0xff);
            const double two_an = rnd_double(-60, 60, (w >> 60) & 3);
            volatile double prod = (double)v * two_an;        /* __dmul_rn */
            n = prod;
        } else {
            n = rnd_double(-400, 400, (w >> 3) & 3);
            d = rnd_double(-400, 400, (w >> 5) & 3);
            if (w & 1) n = -n;
        }
        volatile double yv = 1.0 / d;                  /* RN(1/d): IEEE division */
        const double want = n / d, got = div_alg(n, d, yv);
        if (memcmp(&want, &got, 8) != 0) {
            if (bad < 5) printf("mismatch n=%a d=%a want=%a got=%a\n", n, d, want, got);
            ++bad;
        }
    }
    printf("%llu cases, %llu mismatches\n", (unsigned long long)cases, (unsigned long long)bad);
    return bad ? 1 : 0;
}
