/* Exhaustive-ish CPU check of the division algorithm used by encode_one (flashe_kernels.cu,
 * div_rn_known_rcp): with y = RN(1/d), two FMA residual corrections of q0 = RN(v*y) must equal the
 * IEEE quotient v/d for every float pair in the guarded domain.  Test infrastructure only.
 * usage: div_rcp_check <pairs_log2> <seed>   -> prints the number of mismatches, exit 1 if any. */
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
static float from_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static float div_alg(float v, float d, float y) {
    float q = v * y;
    float r = fmaf(-d, q, v);
    q = fmaf(r, y, q);
    r = fmaf(-d, q, v);
    return fmaf(r, y, q);
}

int main(int argc, char** argv) {
    int lg = argc > 1 ? atoi(argv[1]) : 26;
    s[0] = 0x9E3779B97F4A7C15ull ^ (argc > 2 ? strtoull(argv[2], 0, 10) : 1); s[1] = 0xD1B54A32D192ED03ull;
    uint64_t bad = 0, n = 1ull << lg;
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t w = rnd();
        /* divisor: exponent in [-60, 60]; numerator: exponent in [-60, 60] as well (covers v/d over 2^±120) */
        uint32_t md = (uint32_t)w & 0x7fffff, mv = (uint32_t)(w >> 23) & 0x7fffff;
        uint32_t ed = 127 - 60 + (uint32_t)((w >> 46) % 121), ev = 127 - 60 + (uint32_t)((w >> 54) % 121);
        /* bias a share of the cases towards the hard mantissas (all ones / tiny / near powers of two) */
        if ((i & 7) == 1) md = 0x7fffff - (md & 0xff);
        if ((i & 7) == 2) md = md & 0xff;
        if ((i & 7) == 3) mv = 0x7fffff - (mv & 0xff);
        if ((i & 7) == 4) mv = mv & 0xff;
        float d = from_bits((ed << 23) | md), v = from_bits((ev << 23) | mv);
        float y = (float)(1.0 / (double)d);
        float want = v / d, got = div_alg(v, d, y);
        if (memcmp(&want, &got, 4) != 0) { if (bad < 5) printf("mismatch v=%a d=%a want=%a got=%a\n", v, d, want, got); ++bad; }
    }
    /* the encode domain proper: d = 2*alpha, v = (clip+alpha)*65535 for many alphas, v sweeping a dense grid */
    for (int k = 0; k < 64; ++k) {
        float alpha = (float)(5.938345 * pow(10.0, -4.0 + 0.1 * k));
        float d = (float)(2.0 * (double)alpha), y = (float)(1.0 / (double)d);
        uint32_t top; float vmax = (alpha + alpha) * 65535.0f; memcpy(&top, &vmax, 4);
        for (uint32_t u = 0x20000000u; u <= top; u += 977u) {
            float v = from_bits(u), want = v / d, got = div_alg(v, d, y);
            if (memcmp(&want, &got, 4) != 0) { if (bad < 10) printf("mismatch(enc) v=%a d=%a\n", v, d); ++bad; }
        }
    }
    printf("%llu mismatches\n", (unsigned long long)bad);
    return bad ? 1 : 0;
}
