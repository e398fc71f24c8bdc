"""
TEST INFRASTRUCTURE ONLY.  ctypes wrapper of the plain-C restatement `oracle/nlm_oracle.c`
(reference nd/_filters.pyx:317-420).  `build()` compiles it with gcc into oracle/libnlm_oracle.so.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "nlm_oracle.c")
LIB = os.path.join(_HERE, "libnlm_oracle.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O3", "-fopenmp", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"])
    return LIB


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        for name in ("nlm_oracle_f32", "nlm_oracle_f64"):
            fn = getattr(_lib, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64),
                           ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32),
                           ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.POINTER(ctypes.c_int64)]
    return _lib


def nlmeans(arr, r, f, sigma, h, n_eff=-1, semantics="as_written", threads=None, roi=None):
    """Run the C restatement on a (N0,N1,N2,V) float32/float64 array (any strides).
    roi = ((a0, a1), (b0, b1), (c0, c1)): compute only the output voxels of that box (the rest stays 0)."""
    L = _load()
    arr = np.asarray(arr)
    if arr.dtype not in (np.float32, np.float64):
        raise TypeError("No matching signature found")
    out = np.zeros(arr.shape, dtype=arr.dtype)
    I64 = ctypes.c_int64 * 4
    U32 = ctypes.c_uint32 * 3
    it = arr.itemsize
    if threads is not None:
        os.environ["OMP_NUM_THREADS"] = str(int(threads))
    fn = L.nlm_oracle_f32 if arr.dtype == np.float32 else L.nlm_oracle_f64
    rc = fn(arr.ctypes.data, out.ctypes.data, I64(*arr.shape), I64(*[s // it for s in arr.strides]),
            I64(*[s // it for s in out.strides]), U32(*[int(x) for x in r]), U32(*[int(x) for x in f]),
            float(sigma), float(h), float(n_eff), 1 if semantics == "reference_compiled" else 0,
            (ctypes.c_int64 * 6)(*[int(v) for ab in roi for v in ab]) if roi is not None else None)
    if rc:
        raise ValueError("No solution")
    return out
