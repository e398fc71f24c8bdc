"""
Pins the oracle of the omnibus change detection:

  1. the GSL stand-in oracle/gsl_shim/gsl_shim.h (chi-square CDF) against PUBLISHED chi-square table values and
     scipy.special.gammainc;
  2. the NumPy restatement oracle/omnibus_oracle.py against the reference's OWN nd/_change.pyx compiled unmodified
     into oracle/_ref/_change (when /root/reference or a prebuilt oracle/_ref is present) and against the committed
     golden file generated from it (always).
"""
import ctypes
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "change_golden.npz")


@pytest.fixture(scope="module")
def shim():
    d = tempfile.mkdtemp()
    src = os.path.join(d, "shim.c")
    with open(src, "w") as fh:
        fh.write('#include "gsl_shim.h"\ndouble chisq_P(double x, double nu) { return gsl_cdf_chisq_P(x, nu); }\n')
    so = os.path.join(d, "shim.so")
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "oracle", "gsl_shim"), src, "-o", so, "-lm"])
    L = ctypes.CDLL(so)
    L.chisq_P.restype = ctypes.c_double
    L.chisq_P.argtypes = [ctypes.c_double, ctypes.c_double]
    return L.chisq_P


# Upper percentage points of the chi-square distribution as printed in every statistics handbook
# (e.g. Abramowitz & Stegun table 26.8): P(chi2_nu <= x) = p, x given to 3 decimals.
TABLE = [(1, 3.841, 0.95), (1, 6.635, 0.99), (2, 5.991, 0.95), (2, 9.210, 0.99), (4, 9.488, 0.95), (4, 13.277, 0.99),
         (4, 0.711, 0.05), (8, 15.507, 0.95), (10, 18.307, 0.95), (10, 23.209, 0.99), (10, 3.940, 0.05),
         (20, 31.410, 0.95), (30, 43.773, 0.95), (40, 55.758, 0.95), (60, 79.082, 0.95), (100, 124.342, 0.95)]


def test_shim_matches_published_chi_square_table(shim):
    for nu, x, p in TABLE:
        assert abs(shim(x, nu) - p) < 2e-4, (nu, x, p, shim(x, nu))       # the table is rounded to 3 decimals in x
    assert shim(2 * np.log(2), 2) == pytest.approx(0.5, abs=1e-15)       # chi2_2 is Exp(1/2): median 2 ln 2
    assert shim(0.0, 4) == 0.0 and shim(-1.0, 4) == 0.0 and shim(np.inf, 4) == 1.0 and np.isnan(shim(np.nan, 4))


def test_shim_matches_scipy_to_double_precision(shim):
    from scipy.special import gammainc
    rng = np.random.default_rng(0)
    worst = 0.0
    for nu in [4, 8, 12, 36, 40, 76, 116, 120]:                          # f = 4 (k - 1) and f + 4 of the dual-pol test
        for x in np.concatenate([rng.uniform(0, 3 * nu, 200), [1e-8, 1e-3, 10.0 * nu]]):
            worst = max(worst, abs(shim(float(x), nu) - gammainc(nu / 2.0, x / 2.0)))
    assert worst < 1e-13, worst        # the prefactor exp(-y + a ln y - lgamma(a)) loses ~a ulps for a up to 60


def test_numpy_oracle_matches_golden_from_the_compiled_reference():
    from oracle import omnibus_oracle as oo
    z = np.load(GOLDEN)
    meta = json.loads(str(z["__meta__"]))
    for name, m in meta.items():
        v = z[name + "__in"]
        prob = np.array([[oo.single_pixel_omnibus(v[i, j], m["n"]) for j in range(v.shape[1])] for i in range(v.shape[0])])
        tol = 1e-13 if v.dtype == np.float64 else 1e-7          # float32: the result is rounded to float32 (1 ulp = 6e-8)
        assert np.abs(prob.astype(np.float64) - z[name + "__prob"].astype(np.float64)).max() <= tol, name
        if name.startswith(("step", "jumps", "pair")):                         # the python loops are slow: a subset
            assert np.array_equal(oo.change_detection(v, m["alpha"], m["n"]), z[name + "__change"]), name


def test_compiled_reference_reproduces_golden_when_present():
    from oracle import build_ref, ref_change
    if not build_ref.build_change():
        pytest.skip("oracle/_ref/_change is not built and /root/reference is absent")
    z = np.load(GOLDEN)
    meta = json.loads(str(z["__meta__"]))
    for name, m in meta.items():
        v = z[name + "__in"]
        assert np.array_equal(ref_change.change_detection(v, m["alpha"], m["n"]), z[name + "__change"]), name
        assert np.array_equal(ref_change.omnibus_probability(v, m["n"]), z[name + "__prob"]), name
