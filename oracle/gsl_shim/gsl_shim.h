/*
 * TEST INFRASTRUCTURE ONLY -- not product code.
 *
 * gsl_cdf_chisq_P(x, nu): the chi-square cumulative distribution function, the one GSL entry point the reference's
 * nd/_change.pyx uses (:147-148, through CythonGSL).  GSL is not installed here and the reference pins no GSL
 * version (setup.py:61-69 only probes `gsl-config`), so this restates the PUBLISHED function, not GSL's source:
 *
 *     P(x; nu) = P(a, y) = gamma(a, y) / Gamma(a),   a = nu / 2,  y = x / 2   (regularised lower incomplete gamma;
 *                                                    Abramowitz & Stegun 26.4.19, DLMF 8.2.4)
 *
 * evaluated by the two classical expansions (A&S 6.5.29 / 6.5.31, DLMF 8.11.4 / 8.9.2):
 *     y <  a + 1 :  P = e^{-y} y^a / Gamma(a+1) * sum_{n>=0} y^n / ((a+1)(a+2)...(a+n))
 *     y >= a + 1 :  Q = 1 - P = e^{-y} y^a / Gamma(a) * 1 / (y+1-a- 1(1-a)/(y+3-a- 2(2-a)/(y+5-a- ...)))  (modified Lentz)
 * both to double precision (absolute error < 1e-13 for the degrees of freedom of this test, nu <= 120).  GSL's own gsl_cdf_gamma_P documents the same accuracy class, so
 * the two agree to a few ulps; tests/test_oracle_change.py checks this file against published chi-square table
 * values and scipy.special.gammainc.
 */
#ifndef ND_B200_GSL_SHIM_H
#define ND_B200_GSL_SHIM_H
#include <math.h>

static inline double ndshim_gamma_p(double a, double y) {
    if (!(y > 0.0)) return (y == 0.0 || y < 0.0) ? 0.0 : y;   /* NaN propagates */
    if (isinf(y)) return 1.0;
    const double lg = lgamma(a);
    if (y < a + 1.0) {
        double ap = a, sum = 1.0 / a, del = sum;
        for (int n = 0; n < 100000; ++n) {
            ap += 1.0;
            del *= y / ap;
            sum += del;
            if (fabs(del) < fabs(sum) * 1e-17) break;
        }
        return sum * exp(-y + a * log(y) - lg);
    }
    const double tiny = 1e-300;
    double b = y + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
    for (int i = 1; i < 100000; ++i) {
        const double an = -(double)i * ((double)i - a);
        b += 2.0;
        d = an * d + b;
        if (fabs(d) < tiny) d = tiny;
        c = b + an / c;
        if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < 1e-16) break;
    }
    return 1.0 - exp(-y + a * log(y) - lg) * h;
}

static inline double gsl_cdf_chisq_P(double x, double nu) { return ndshim_gamma_p(0.5 * nu, 0.5 * x); }

#endif
