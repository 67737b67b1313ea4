/* Host proof-by-exhaustion helper for the CMVN variance rounding trick of csrc/kernels.cu (cmvn_chains::term).
 *
 * The reference accumulates  S = (float)((double)S + pow((double)(x - mean), 2))   (numpy.hpp:819-825 of the pinned SDK copy).
 * The kernel keeps S in a double register and rounds it to float precision with two FMAs:
 *     t  = fma(d, d, S)                                   RN53(S + d^2), d*d exact
 *     m1 = { hi: max(hi(t), 0x38100000), lo: lo(d) }      t's binade, even mantissa, floor at 2^-126
 *     g  = fma(m1,  2^29, t)                              rounds t at float granularity (ties to even)
 *     S' = fma(m1, -2^29, g)                              exact
 * This program checks S' == (double)(float)t over random operands in several regimes (wide exponents, operands with few
 * significant bits -- these produce exact ties by the million --, the float-denormal range, all-ones mantissas).
 * Exit status 0 = every case identical.  Usage: check_round_trick [iterations]
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int32_t hi32(double x) { uint64_t u; memcpy(&u, &x, 8); return (int32_t)(u >> 32); }
static uint32_t lo32(double x) { uint64_t u; memcpy(&u, &x, 8); return (uint32_t)u; }
static double mk(int32_t h, uint32_t l) { uint64_t u = ((uint64_t)(uint32_t)h << 32) | l; double x; memcpy(&x, &u, 8); return x; }
static uint64_t st = 88172645463325252ull;
static uint64_t rnd(void) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; }

static double trick(double t, double d) {
    int32_t h = hi32(t);
    double m1 = mk(h > (897 << 20) ? h : (897 << 20), lo32(d));
    double g = fma(m1, 536870912.0, t);
    return fma(m1, -536870912.0, g);
}

int main(int argc, char **argv) {
    long iters = argc > 1 ? atol(argv[1]) : 20000000L, bad = 0, n = 0, ties = 0, denorm = 0;
    for (long it = 0; it < iters; it++) {
        uint32_t sb = (uint32_t)rnd(), db = (uint32_t)rnd();
        float S, df;
        switch (it & 7) {
            case 0: case 1: case 2:  /* any finite operands */
                sb &= 0x7fffffffu; memcpy(&S, &sb, 4); memcpy(&df, &db, 4);
                if (!(S < 1e30f) || !(fabsf(df) < 1e15f)) continue;
                break;
            case 3:  /* sums whose mantissa is all ones in the high word: the top of the binade */
                sb = (sb & 0x7f800000u) | 0x007ffff8u | (sb & 7u); memcpy(&S, &sb, 4); memcpy(&df, &db, 4);
                if (!(S < 1e30f) || !(fabsf(df) < 1e15f)) continue;
                break;
            case 4: case 5: {  /* few significant bits: exact ties */
                S = (float)((rnd() % 100000) / 64.0);
                int sh = (int)(rnd() % 20);
                df = ldexpf((float)((int32_t)(rnd() % 4096) - 2048), -sh);
                break;
            }
            default:  /* around and below the smallest normal float */
                S = ldexpf((float)(rnd() % (1 << 24)), -149 - (int)(rnd() % 4) + (int)(rnd() % 30));
                df = ldexpf((float)(rnd() % (1 << 24)), -90 + (int)(rnd() % 30));
        }
        const double d = (double)df, t = fma(d, d, (double)S);
        const float want = (float)t;
        const double got = trick(t, d);
        n++;
        if ((lo32(t) & 0x1fffffffu) == 0x10000000u) ties++;
        if (t != 0.0 && t < 1.1754943508222875e-38) denorm++;
        if (got != (double)want) {
            if (bad < 10) printf("MISMATCH S=%a d=%a t=%a want=%a got=%a\n", S, df, t, (double)want, got);
            bad++;
        }
    }
    printf("checked=%ld mismatches=%ld exact_ties=%ld float_denormal_sums=%ld\n", n, bad, ties, denorm);
    return bad != 0;
}
