"""
TEST INFRASTRUCTURE ONLY.  NumPy restatement of the reference's omnibus change detection
(nd/_change.pyx: `_f` :19-23, `_rho` :26-30, `_omega2` :33-39, `_z` :45-79, `single_pixel_omnibus` :139-160,
`single_pixel_change_detection` :224-260, `change_detection` :263-287), statement by statement, with the
reference's typing: `floating` locals take the dtype of the data (float32 or float64), `prod_of_dets` / `logQ`
and the constants are double.

PARITY PINNED (round 2): the reference's own nd/_change.pyx is compiled UNMODIFIED into oracle/_ref/_change by
oracle/build_ref.py; its only third-party call, `cython_gsl.gsl_cdf_chisq_P` (:147-148; GSL is not installed and
the reference pins no version), is supplied by the stand-in oracle/gsl_shim (the published regularised incomplete
gamma function, checked against chi-square tables and scipy).  tests/test_oracle_change.py compares this
restatement with outputs of that binary (tests/golden/change_golden.npz): probabilities to 1e-13, change maps
exactly.  Here the CDF is `scipy.special.gammainc(f/2, z/2)` -- what the reference's own deprecated `array_omnibus`
uses (`scipy.stats.chi2.cdf`, :124-125); shim, scipy and GSL agree to ~1e-15, so a decision `p > alpha` could differ
only for a probability within that distance of alpha.
"""
import numpy as np
from scipy.special import gammainc


def _f(p, k, n):
    return (k - 1) * p ** 2


def _rho(p, k, n):
    return 1 - (2 * p ** 2 - 1) / (6 * (k - 1) * p) * (k / n - 1 / (n * k))


def _omega2(p, k, n, rho):
    return p ** 2 * (p ** 2 - 1) / (24 * rho ** 2) * (k / (n ** 2) - 1 / ((n * k) ** 2)) \
        - p ** 2 * (k - 1) / 4 * (1 - 1 / rho) ** 2


def _z(ts, n):
    """nd/_change.pyx:45-79; ts is (k, 4) = [C11, C12.real, C12.imag, C22]."""
    T = ts.dtype.type
    k = ts.shape[0]
    p = T(2)
    c11sum = c22sum = c12rsum = c12isum = T(0)
    prod_of_dets = np.float64(1.0)
    for i in range(k):
        det = T(T(ts[i, 0] * ts[i, 3]) - T(T(ts[i, 1] * ts[i, 1]) + T(ts[i, 2] * ts[i, 2])))
        prod_of_dets = np.float64(prod_of_dets * np.float64(det))
        c11sum = T(c11sum + ts[i, 0])
        c12rsum = T(c12rsum + ts[i, 1])
        c12isum = T(c12isum + ts[i, 2])
        c22sum = T(c22sum + ts[i, 3])
    det_of_sum = T(T(c11sum * c22sum) - T(T(c12rsum * c12rsum) + T(c12isum * c12isum)))
    with np.errstate(all='ignore'):
        # p * k is evaluated in `floating`, everything after it in double
        logQ = np.float64(n) * (np.float64(T(p * T(k))) * np.log(np.float64(k)) + np.log(prod_of_dets)
                                - np.float64(k) * np.log(np.float64(det_of_sum)))
    rho = T(_rho(2.0, float(k), float(n)))
    z = T(np.float64(T(-2) * rho) * logQ)
    return z


def single_pixel_omnibus(ts, n):
    """nd/_change.pyx:139-160"""
    T = ts.dtype.type
    k = float(ts.shape[0])
    f = _f(2.0, k, float(n))
    rho = _rho(2.0, k, float(n))
    omega2 = _omega2(2.0, k, float(n), rho)
    z = _z(ts, n)
    with np.errstate(all='ignore'):
        P1 = T(chisq_P(np.float64(z), f))
        P2 = T(chisq_P(np.float64(z), f + 4))
        return T(np.float64(P1) + omega2 * np.float64(T(P2 - P1)))


def chisq_P(x, nu):
    """gsl_cdf_chisq_P(x, nu): 0 for x <= 0, else the regularised lower incomplete gamma P(nu/2, x/2)."""
    if np.isnan(x):
        return np.float64(np.nan)
    if x <= 0:
        return np.float64(0.0)
    return np.float64(gammainc(nu / 2.0, x / 2.0))


def single_pixel_change_detection(ts, alpha, n):
    """nd/_change.pyx:224-260; returns the uint8 change vector of one pixel."""
    k = ts.shape[0]
    result = np.zeros(k, np.uint8)
    l = 0
    r = 0
    while True:
        p_H0_l = single_pixel_omnibus(ts[l:], n)
        if not (p_H0_l > alpha):
            break
        for j in range(2, k - l + 1):
            p_H0_lj = single_pixel_omnibus(ts[l:l + j], n)
            r = j - 1
            if p_H0_lj > alpha:
                result[l + r] = 1
                break
        l = l + r
        if l >= k - 1:
            break
    return result


def change_detection(values, alpha, n=1):
    """nd/_change.pyx:263-287; values is (rows, cols, k, 4)."""
    rows, cols, k = values.shape[:3]
    out = np.zeros((rows, cols, k), np.uint8)
    for i in range(rows):
        for j in range(cols):
            out[i, j] = single_pixel_change_detection(values[i, j], alpha, n)
    return out
