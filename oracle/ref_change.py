"""
TEST INFRASTRUCTURE ONLY.  Thin wrappers around the reference's own omnibus change detection, compiled UNMODIFIED
from /root/reference/nd/_change.pyx into oracle/_ref/_change by oracle/build_ref.py (with the chi-square CDF of
oracle/gsl_shim standing in for GSL, which is not installed).  Pins oracle/omnibus_oracle.py and, through the golden
file tests/golden/change_golden.npz, the CUDA kernels of include/ndchg.h.
"""
import importlib
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def available():
    from . import build_ref
    return build_ref.change_built()


def _mod():
    ref_dir = os.path.join(_HERE, "_ref")
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    return importlib.import_module("_change")


def change_detection(values, alpha, n=1, njobs=1):
    """nd/_change.pyx:263-287: (rows, cols, k, 4) float32/float64 -> (rows, cols, k) uint8."""
    v = np.ascontiguousarray(values)
    return np.asarray(_mod().change_detection(v, float(alpha), int(n), int(njobs))).copy()


def single_pixel_omnibus(ts, n):
    """nd/_change.pyx:139-160: probability of change for one (k, 4) series, in the dtype of the data."""
    return _mod().single_pixel_omnibus(np.ascontiguousarray(ts), int(n))


def omnibus_probability(values, n):
    v = np.ascontiguousarray(values)
    out = np.empty(v.shape[:2], dtype=v.dtype)
    for i in range(v.shape[0]):
        for j in range(v.shape[1]):
            out[i, j] = single_pixel_omnibus(v[i, j], n)
    return out
