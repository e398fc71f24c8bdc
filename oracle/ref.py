"""
TEST INFRASTRUCTURE ONLY.  Thin wrappers around the reference's own compiled kernel
(`oracle/_ref/`, built by `oracle/build_ref.py` from `/root/reference/nd/_filters.pyx`).

Two semantics:

* `reference_compiled(...)` -- the unmodified kernel called directly.  On LP64 the patch
  loops `range(-f[i], f[i]+1)` (nd/_filters.pyx:373-375) never execute when f[i] > 0
  because `f` is `unsigned int` (nd/_filters.pyx:323): d^2 == 0, every weight == 1.
* `as_written(...)` -- what the .pyx text / docs describe.  Obtained from the SAME
  unmodified binary by the pad+augment+crop construction: reflect-pad by
  r+f, stack the (2f+1)^3 shifted copies as extra variables, run with f=(0,0,0), crop.
  `as_written_patched(...)` is the 3-cast copy, a faster cross-check.
"""
import importlib
import itertools
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def available():
    from . import build_ref
    return build_ref.built()


def _mod(name):
    ref_dir = os.path.join(_HERE, "_ref")
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    return importlib.import_module(name)


def _u32(x):
    return np.asarray(x, dtype=np.uint32)


def reference_compiled(arr, r, f, sigma, h, n_eff=-1):
    """Direct call of the unmodified `_pixelwise_nlmeans_3d` (nd/_filters.pyx:320)."""
    out = np.empty_like(arr)
    _mod("_filters")._pixelwise_nlmeans_3d(arr, out, _u32(r), _u32(f), float(sigma), float(h), float(n_eff))
    return out


def as_written_patched(arr, r, f, sigma, h, n_eff=-1):
    """The .pyx with `<SIZE_TYPE>` casts on the three patch loops."""
    out = np.empty_like(arr)
    _mod("_filters_aw")._pixelwise_nlmeans_3d(arr, out, _u32(r), _u32(f), float(sigma), float(h), float(n_eff))
    return out


def as_written(arr, r, f, sigma, h, n_eff=-1):
    """As-written semantics from the UNMODIFIED binary."""
    r = [int(x) for x in r]
    f = [int(x) for x in f]
    if not any(f):
        return reference_compiled(arr, r, f, sigma, h, n_eff)
    N, V = arr.shape[:3], arr.shape[3]
    P = np.pad(arr, [(r[i] + f[i],) * 2 for i in range(3)] + [(0, 0)], mode="reflect")
    M = tuple(N[i] + 2 * r[i] for i in range(3))
    shifts = [(0, 0, 0)] + [d for d in itertools.product(*[range(-k, k + 1) for k in f])
                            if d != (0, 0, 0)]
    aug = np.concatenate([P[f[0] + d[0]:f[0] + d[0] + M[0],
                            f[1] + d[1]:f[1] + d[1] + M[1],
                            f[2] + d[2]:f[2] + d[2] + M[2]] for d in shifts], axis=-1)
    aug = np.ascontiguousarray(aug)
    out = np.empty_like(aug)
    _mod("_filters")._pixelwise_nlmeans_3d(aug, out, _u32(r), _u32((0, 0, 0)), float(sigma), float(h), float(n_eff))
    return np.ascontiguousarray(out[r[0]:r[0] + N[0], r[1]:r[1] + N[1], r[2]:r[2] + N[2], :V])


def find_weight(weight_sum, sq_weight_sum, n):
    """nd/_filters.pyx:297-314 (raises ValueError('No solution'))."""
    return _mod("_filters").find_weight(float(weight_sum), float(sq_weight_sum), float(n))
