"""
Omnibus change detection: the CUDA kernels behind include/ndchg.h against the NumPy
restatement of nd/_change.pyx in oracle/omnibus_oracle.py.

PARITY PINNED: tests/golden/change_golden.npz holds outputs of the reference's OWN nd/_change.pyx, compiled
unmodified against a stand-in for its only third-party call, `gsl_cdf_chisq_P` (oracle/gsl_shim, checked against
published chi-square tables and scipy in tests/test_oracle_change.py); the NumPy restatement is pinned to the same
file there.  Bars: probabilities within 1e-12 (float64 data) / 2e-6 (float32 data: the statistic itself is rounded
to float32 like in the reference); change maps identical.
"""
import inspect
import os
import re

import numpy as np
import pytest

from nd_b200 import _lib, change
from nd_b200.dataset import DataArray, concat, generate_test_dataset

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def wishart_series(rows, cols, k, looks, scales, seed=0, dtype=np.float64):
    """(rows, cols, k, 4) dual-pol covariance samples [C11, Re C12, Im C12, C22] with `looks` looks; the scene
    power follows `scales[t]`."""
    rng = np.random.default_rng(seed)
    z = (rng.normal(size=(rows, cols, k, looks, 2)) + 1j * rng.normal(size=(rows, cols, k, looks, 2))) / np.sqrt(2)
    z = z * np.sqrt(np.asarray(scales, dtype=np.float64))[None, None, :, None, None]
    c11 = (np.abs(z[..., 0]) ** 2).mean(-1)
    c22 = (np.abs(z[..., 1]) ** 2).mean(-1)
    c12 = (z[..., 0] * np.conj(z[..., 1])).mean(-1)
    return np.stack([c11, c12.real, c12.imag, c22], axis=-1).astype(dtype)


# ---- CPU: boundary and oracle sanity ------------------------------------------------------------------
def test_library_exports_every_symbol_of_ndchg_h():
    header = open(os.path.join(ROOT, "include", "ndchg.h")).read()
    declared = set(re.findall(r"\b(ndchg_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.CHG_SYMBOLS)
    L = _lib.lib()
    for sym in declared:
        assert hasattr(L, sym), sym


def test_interface_mirrors_the_reference():
    sig = inspect.signature(change.OmnibusTest.__init__).parameters
    assert [sig[k].default for k in ('ml', 'n', 'alpha')] == [None, 1, 0.01]           # nd/change.py:104
    assert issubclass(change.OmnibusTest, change.ChangeDetection) and change.OmnibusTest().njobs == 1
    assert list(inspect.signature(change.change_detection).parameters) == ['values', 'alpha', 'n', 'njobs']
    assert callable(change.omnibus)


def test_oracle_detects_a_step_change_and_nothing_else():
    from oracle import omnibus_oracle as oo
    v = wishart_series(2, 3, 10, 9, [1.0] * 5 + [10.0] * 5, seed=1)
    res = oo.change_detection(v, alpha=0.9999, n=9)                    # `alpha` is the confidence the test must exceed
    assert res[:, :, 5].all() and (res.sum(-1) == 1).all()
    flat = wishart_series(2, 3, 10, 9, [1.0] * 10, seed=2)
    assert oo.change_detection(flat, alpha=0.9999, n=9).sum() == 0


def test_oracle_chisq_matches_scipy_stats():
    from scipy.stats import chi2
    from oracle import omnibus_oracle as oo
    for x, nu in [(0.5, 4), (3.0, 4), (20.0, 12), (7.7, 36), (-1.0, 4), (0.0, 8)]:
        assert abs(oo.chisq_P(x, nu) - (chi2.cdf(x, nu) if x > 0 else 0.0)) < 1e-15


def test_compute_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        change.change_detection(wishart_series(2, 2, 4, 4, [1, 1, 1, 1]), 0.01, 4)
    with pytest.raises(TypeError):
        change.change_detection(np.zeros((2, 2, 4, 4), np.int32), 0.01, 4)
    with pytest.raises(ValueError):
        change.change_detection(np.zeros((2, 2, 4), np.float32), 0.01, 4)


# ---- GPU ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 2e-6)])
@pytest.mark.parametrize("k,looks,amp", [(2, 4, 0.3), (5, 9, 0.3), (12, 16, 0.3), (30, 50, 0.3), (30, 50, 0.05)])
def test_omnibus_probability_matches_oracle(dtype, tol, k, looks, amp):
    from oracle import omnibus_oracle as oo
    scales = 1.0 + amp * np.sin(np.arange(k))          # mild changes: probabilities spread over (0, 1)
    v = wishart_series(6, 7, k, looks, scales, seed=k, dtype=dtype)
    got = change.omnibus_probability(v, n=looks)
    ref = np.array([[oo.single_pixel_omnibus(v[i, j], looks) for j in range(7)] for i in range(6)])
    assert got.dtype == dtype and np.abs(got.astype(np.float64) - ref.astype(np.float64)).max() < tol
    assert amp == 0.3 and k == 30 or 0.05 < np.median(ref) < 0.999     # not saturated (except the deliberately saturated case)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_change_maps_equal_oracle(dtype):
    from oracle import omnibus_oracle as oo
    rng = np.random.default_rng(5)
    for k, looks, alpha in [(10, 9, 0.9), (10, 9, 0.99), (16, 4, 0.999), (7, 25, 0.5)]:
        scales = np.exp(np.cumsum(rng.choice([0.0, 0.0, 0.0, 1.2, -1.0], size=k)))      # a few jumps per series
        v = wishart_series(8, 9, k, looks, scales, seed=k + looks, dtype=dtype)
        got = change.change_detection(v, alpha=alpha, n=looks)
        ref = oo.change_detection(v, alpha=alpha, n=looks)
        assert got.dtype == np.uint8 and got.shape == (8, 9, k)
        assert np.array_equal(got, ref), (k, looks, alpha, int((got != ref).sum()))
        assert ref.sum() > 0


@pytest.mark.gpu
def test_golden_from_the_compiled_reference():
    """CUDA kernels against outputs of the reference's own compiled nd/_change.pyx (tests/golden/make_change_golden.py)."""
    import json
    z = np.load(os.path.join(ROOT, "tests", "golden", "change_golden.npz"))
    meta = json.loads(str(z["__meta__"]))
    for name, m in meta.items():
        v = z[name + "__in"]
        got = change.change_detection(v, alpha=m["alpha"], n=m["n"])
        assert np.array_equal(got, z[name + "__change"]), (name, int((got != z[name + "__change"]).sum()))
        prob = change.omnibus_probability(v, n=m["n"])
        tol = 1e-12 if v.dtype == np.float64 else 2e-6
        assert np.abs(prob.astype(np.float64) - z[name + "__prob"].astype(np.float64)).max() < tol, name
        assert m["changes"] > 0


@pytest.mark.gpu
def test_strided_input_and_nan_pixels():
    from oracle import omnibus_oracle as oo
    v = wishart_series(5, 6, 8, 9, [1, 1, 1, 6, 6, 6, 6, 6], seed=3)
    vm = np.ascontiguousarray(np.moveaxis(v, -1, 0))                     # variable-major, like `to_array()`
    view = np.moveaxis(vm, 0, -1)
    assert not view.flags['C_CONTIGUOUS']
    assert np.array_equal(change.change_detection(view, 0.99, 9), oo.change_detection(v, 0.99, 9))
    v[1, 2, 3, 0] = np.nan                                               # NaN: comparisons are false, no change
    got = change.change_detection(v, 0.99, 9)
    assert np.array_equal(got, oo.change_detection(v, 0.99, 9)) and got[1, 2].sum() == 0


@pytest.mark.gpu
def test_change_reference_test():
    """nd/tests/test_change_omnibus.py:7-20 and test_change_common.py:22-33."""
    ds1 = generate_test_dataset(dims={'y': 5, 'x': 5, 'time': 10}, mean=[1, 0, 0, 1], sigma=0.1).isel(time=slice(None, 5))
    ds2 = generate_test_dataset(dims={'y': 5, 'x': 5, 'time': 10}, mean=[10, 0, 0, 10], sigma=0.1).isel(time=slice(5, None))
    ds = concat([ds1, ds2], 'time')
    changes = change.OmnibusTest(n=9, alpha=0.9).apply(ds)
    assert isinstance(changes, DataArray) and changes.name == 'change' and changes.dims == ('y', 'x', 'time')
    assert changes.dtype == bool and changes.isel(time=5).all()
    assert (changes.sum(dim='time') == 1).all()
    assert list(changes.coords) == list(ds.coords) and dict(changes.attrs) == dict(ds.attrs)
    same = change.omnibus(ds, n=9, alpha=0.9)
    assert np.array_equal(same.values, changes.values)
    ml = change.OmnibusTest(ml=3, alpha=0.9).apply(ds)                     # multilooking through the GPU BoxcarFilter
    assert ml.values.shape == changes.values.shape and ml.isel(time=5).all()
