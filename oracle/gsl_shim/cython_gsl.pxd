# TEST INFRASTRUCTURE ONLY.  Minimal stand-in for the `cython_gsl` package (CythonGSL), which the reference's
# nd/_change.pyx cimports (:9) for exactly ONE function, `gsl_cdf_chisq_P` (:147-148).  GSL and CythonGSL are not
# installed in this image; this .pxd lets oracle/build_ref.py compile the reference's own _change.pyx UNMODIFIED,
# with the chi-square CDF supplied by oracle/gsl_shim/gsl_shim.h.
cdef extern from "gsl_shim.h" nogil:
    double gsl_cdf_chisq_P(double x, double nu)
