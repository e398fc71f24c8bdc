"""
Change detection, mirroring reference nd/change.py:13-119: `ChangeDetection`,
`OmnibusTest(ml, n, alpha).apply(ds)` and the function form `omnibus`.  The tutorial pipeline is
`ds.filter.nlmeans(...)` followed by `ds_nlm.nd.change_omnibus(n=50, alpha=1e-4)` (examples/tutorial_s1.ipynb:133).

Below `_omnibus_change_detection` the reference calls its Cython / GSL extension `_change.change_detection`
(nd/_change.pyx:263-287); here that call goes to the CUDA kernel behind include/ndchg.h (one GPU thread per pixel).
There is no CPU fallback.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .algorithm import Algorithm, wrap_algorithm
from .dataset import DataArray
from .filters import BoxcarFilter, disassemble_complex

__all__ = ['ChangeDetection', 'OmnibusTest', 'omnibus']

_VARS = ['C11', 'C12__re', 'C12__im', 'C22']


def _check(rc):
    if rc == 0:
        return
    msg = _lib.lib().ndchg_last_error().decode('utf-8', 'replace')
    if rc == _lib.EDTYPE:
        raise TypeError(msg)
    if rc == _lib.EINVAL:
        raise ValueError(msg)
    raise RuntimeError('ndchg: ' + msg)


def _prepare(values):
    values = np.asarray(values)
    if values.ndim != 4 or values.shape[3] != 4:
        raise ValueError('Buffer has wrong number of dimensions (expected (rows, cols, time, 4))')
    if values.dtype not in (np.float32, np.float64):
        raise TypeError('No matching signature found')
    if not torch.cuda.is_available():
        raise RuntimeError('nd_b200 change detection needs a CUDA device; there is no CPU fallback')
    return values, torch.from_numpy(np.ascontiguousarray(values)).cuda()


def change_detection(values, alpha, n=1, njobs=1):
    """`_change.change_detection(values, alpha, n, njobs)` (nd/_change.pyx:263-287) on the GPU: `values` is
    (rows, cols, time, 4) = [C11, Re C12, Im C12, C22], already multilooked with `n` looks; returns the uint8
    array (rows, cols, time) of detected changes.  `njobs` (OpenMP threads in the reference) is ignored."""
    values, t = _prepare(values)
    res = torch.empty(values.shape[:3], dtype=torch.uint8, device=t.device)
    _check(_lib.lib().ndchg_change_detection(ctypes.c_void_p(t.data_ptr()), _lib.i64(values.shape[:3]), _lib.i64(t.stride()),
                                             0 if values.dtype == np.float32 else 1, ctypes.c_void_p(res.data_ptr()),
                                             float(alpha), int(n), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return res.cpu().numpy()


def omnibus_probability(values, n=1):
    """`single_pixel_omnibus` (nd/_change.pyx:139-160) of every pixel over its whole series."""
    values, t = _prepare(values)
    prob = torch.empty(values.shape[:2], dtype=t.dtype, device=t.device)
    _check(_lib.lib().ndchg_omnibus_probability(ctypes.c_void_p(t.data_ptr()), _lib.i64(values.shape[:3]), _lib.i64(t.stride()),
                                                0 if values.dtype == np.float32 else 1, ctypes.c_void_p(prob.data_ptr()),
                                                int(n), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return prob.cpu().numpy()


class ChangeDetection(Algorithm):
    njobs = 1

    def __init__(self, njobs=1):
        self.njobs = njobs


def _omnibus_change_detection(ds, alpha=0.01, ml=None, n=1, njobs=1):
    """Conradsen et al. (2015) omnibus change detection (reference nd/change.py:32-78)."""
    ds_m = ds.copy(deep=True)
    disassemble_complex(ds_m)
    if ml is not None:                                   # multilooking
        ds_m = BoxcarFilter(w=ml).apply(ds_m)
        n = ml ** 2
    dims = ('y', 'x', 'time')
    stack = []
    for v in _VARS:
        var = ds_m[v]
        stack.append(np.transpose(var.values, [list(var.dims).index(d) for d in dims]))
    values = np.stack(stack, axis=-1)
    change = change_detection(values, alpha=alpha, n=n, njobs=njobs)
    coords = {k: c for k, c in ds.coords.items()}
    return DataArray(np.asarray(change, dtype=bool), dims, coords=coords, attrs=ds.attrs, name='change')


class OmnibusTest(ChangeDetection):
    """
    OmnibusTest (reference nd/change.py:81-116).

    Parameters
    ----------
    ml : int, optional
        Multilooking window size (default: the dataset is already multilooked).
    n : int, optional
        The number of looks in `ds`; ignored if `ml` is given (default: 1).
    alpha : float (0. ... 1.), optional
        The significance level (default: 0.01).
    """

    def __init__(self, ml=None, n=1, alpha=0.01, *args, **kwargs):
        self.ml = ml
        self.n = n
        self.alpha = alpha
        super().__init__(*args, **kwargs)

    def apply(self, ds):
        return _omnibus_change_detection(ds, alpha=self.alpha, ml=self.ml, n=self.n, njobs=self.njobs)


omnibus = wrap_algorithm(OmnibusTest, 'omnibus')
