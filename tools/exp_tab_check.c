/* Host-side check of exp_tab() (modulatedgps_b200/csrc/stream_kernels.cu): the same table, constants and operation
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

static double exp_tab(double x) {
    const double MAGIC = 6755399441055744.0;
    const double tt = fma(x, EXP_TAB_L, MAGIC);
    int64_t bits;
    memcpy(&bits, &tt, 8);
    const int k = (int)(uint32_t)bits;
    const double kf = tt - MAGIC;
    double r = fma(kf, -EXP_TAB_C_HI, x);
    r = fma(kf, -EXP_TAB_C_LO, r);
    double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p *= r;
    const double tj = tab[k & 63];
    double res = fma(tj, p, tj);
    int64_t rb;
    memcpy(&rb, &res, 8);
    rb += ((int64_t)(k >> 6)) << 52;
    memcpy(&res, &rb, 8);
    return x < -700.0 ? 0.0 : res;
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
        const double a = exp_tab(x);
        const long double b = expl((long double)x);
        const double e = fabs((double)(((long double)a - b) / b)) / 1.1102230246251565e-16;
        if (e > worst) { worst = e; worst_x = x; }
    }
    printf("worst relative error %.3f x 2^-53 at x = %.17g; exp_tab(0) = %.17g, exp_tab(1) = %.17g\n", worst, worst_x,
           exp_tab(0.0), exp_tab(1.0));
    return (worst <= 2.2 && exp_tab(0.0) == 1.0) ? 0 : 1;
}
