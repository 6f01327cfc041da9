/* Host-side check of exp2_tab() (modulatedgps_b200/csrc/stream_kernels.cu): the same table, constants and operation
 * order in plain C (fma() is exact on the host as on the device), against expl() in long double.
 *     gcc -O2 -o /tmp/exp_tab_check tools/exp_tab_check.c -lm && /tmp/exp_tab_check [samples]
 * Prints the worst relative error in units of 2^-53 (ulp/2 of a double in [1, 2)); exits 1 above 2.2 (= 1.1 ulp).
 * Used by tests/test_host_logic.py::test_exp_tab_is_accurate_to_an_ulp. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../modulatedgps_b200/csrc/exp_tab.h"

static const double tab[64] = {EXP_TAB64_VALUES};

/* exp2_tab(y) = exp(y ln2 / 64): the argument arrives in units of ln2/64 (the kernels' z.x contraction is pre-scaled) */
static double exp2_tab(double y) {
    const double MAGIC = 6755399441055744.0;
    const double tt = y + MAGIC;
    int64_t bits;
    memcpy(&bits, &tt, 8);
    const int k = (int)(uint32_t)bits;
    const double rr = y - (tt - MAGIC);
    double p = fma(rr, EXP2_C5, EXP2_C4);
    p = fma(p, rr, EXP2_C3);
    p = fma(p, rr, EXP2_C2);
    p = fma(p, rr, EXP2_C1);
    p *= rr;
    const double tj = tab[k & 63];
    double res = fma(tj, p, tj);
    int64_t rb, yb;
    memcpy(&rb, &res, 8);
    rb += ((int64_t)(k >> 6)) << 52;
    memcpy(&res, &rb, 8);
    memcpy(&yb, &y, 8);
    return (uint32_t)((uint64_t)yb >> 32) > 0xC0EF8F17u ? 0.0 : res;   /* y < -700 * 64 / ln 2, on the integer pipe */
}

int main(int argc, char** argv) {
    const long n = argc > 1 ? atol(argv[1]) : 20000000L;
    double worst = 0.0, worst_x = 0.0;
    uint64_t st = 88172645463325252ULL;
    for (long i = 0; i < n; ++i) {
        st ^= st << 13; st ^= st >> 7; st ^= st << 17;   /* xorshift64 */
        const double u = (double)(st >> 11) / 9007199254740992.0;
        /* a third of the samples where Kuf lives (exponents of a few tens), the rest over the whole range */
        const double x = (i % 3 == 0) ? -40.0 * u : -700.0 + 1400.0 * u;
        if (x > 709.0) continue;
        /* the kernels never form x: they see y.  Reference = exp of the real number y represents. */
        const double y = x * EXP_TAB_L;
        const double a = exp2_tab(y);
        /* 2^(y/64) = 2^(k/64) e^(rr ln2/64) with k = round(y): both factors are accurate to a long-double ulp, whereas
         * expl(y ln2/64) would carry |y ln2/64| 2^-64 of argument rounding — a third of the unit used below at x = 700 */
        const long double kk = roundl((long double)y);
        const long double b = exp2l(kk / 64.0L) * expl(((long double)y - kk) * (0.693147180559945309417232121458176568L / 64.0L));
        const double e = fabs((double)(((long double)a - b) / b)) / 1.1102230246251565e-16;
        if (e > worst) { worst = e; worst_x = x; }
    }
    const double lo = exp2_tab(-700.5 * EXP_TAB_L), edge = exp2_tab(-699.5 * EXP_TAB_L);
    printf("worst relative error %.3f x 2^-53 at x = %.17g; exp2_tab(0) = %.17g, exp2_tab(64/ln2) = %.17g, below -700: %g, above: %g\n",
           worst, worst_x, exp2_tab(0.0), exp2_tab(EXP_TAB_L), lo, edge);
    return (worst <= 2.2 && exp2_tab(0.0) == 1.0 && lo == 0.0 && edge > 0.0) ? 0 : 1;
}
