"""
Sibling filters: ConvolutionFilter / BoxcarFilter / GaussianFilter.

The reference implements them as thin wrappers over scipy.ndimage (nd/filters.py:205-381) and its own
tests compare with scipy using assert_equal (nd/tests/test_convolution_filter.py, test_gaussian_filter.py),
so scipy.ndimage -- installed in this image, the arithmetic's real home -- is the oracle here and the bar is
BIT-EXACT agreement for float32 and float64 (the CUDA kernels restate scipy's C loops operation by operation).

CPU part (-m "not gpu"): header/symbols, host logic (kernel construction, argument handling, class
interface).  GPU part (-m gpu): parity through the C ABI.
"""
import inspect
import os
import re

import numpy as np
import pytest
import scipy.ndimage as sn

from nd_b200 import _lib, _ndimage
from nd_b200.dataset import Dataset, generate_test_dataset
from nd_b200.filters import (BoxcarFilter, ConvolutionFilter, Filter, GaussianFilter, NLMeansFilter, _expand_kernel,
                             boxcar, convolution, gaussian)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILTER_CLASSES = [ConvolutionFilter, BoxcarFilter, GaussianFilter, NLMeansFilter]


# ---- CPU: boundary and host logic ------------------------------------------------------------------
def test_library_exports_every_symbol_of_ndflt_h():
    header = open(os.path.join(ROOT, "include", "ndflt.h")).read()
    declared = set(re.findall(r"\b(ndflt_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.FLT_SYMBOLS)
    L = _lib.lib()
    for sym in declared:
        assert hasattr(L, sym), sym


@pytest.mark.parametrize("sigma,order,radius", [(1.0, 0, 4), (2.5, 0, 10), (0.7, 1, 3), (1.3, 2, 5), (3.0, 3, 12)])
def test_gaussian_kernel_equals_scipy(sigma, order, radius):
    from scipy.ndimage import _filters as sf
    assert np.array_equal(_ndimage._gaussian_kernel1d(sigma, order, radius), sf._gaussian_kernel1d(sigma, order, radius))


def test_gaussian_plan_skips_zero_sigma_and_reverses_kernel():
    plan = _ndimage.gaussian_plan(3, [1.0, 0, 2.0])
    assert [p[0] for p in plan] == [0, 2]
    assert len(plan[0][1]) == 9 and len(plan[1][1]) == 17            # radius = int(4 sigma + 0.5)
    plan = _ndimage.gaussian_plan(2, 1.0, order=[0, 1], radius=[2, None])
    assert len(plan[0][1]) == 5
    from scipy.ndimage import _filters as sf
    assert np.array_equal(plan[1][1], sf._gaussian_kernel1d(1.0, 1, 4)[::-1])
    with pytest.raises(RuntimeError):
        _ndimage.gaussian_plan(3, [1.0, 2.0])


def test_expand_kernel():
    new_kernel = _expand_kernel(np.ones((2, 3)), ('x', 'y'), ('x', 'a', 'y', 's'))      # nd/tests/test_convolution_filter.py:17-24
    assert new_kernel.shape == (2, 1, 3, 1)
    with pytest.raises(ValueError):
        _expand_kernel(np.ones((2, 3)), ('x', 'y'), ('x', 'a'))
    with pytest.raises(ValueError):
        _expand_kernel(np.ones((2, 3)), ('x',), ('x', 'a'))


@pytest.mark.parametrize("f", FILTER_CLASSES)
def test_filter_signature(f):
    """nd/tests/test_filters_common.py:37-41"""
    assert list(inspect.signature(f._filter).parameters.keys()) == ['self', 'arr', 'axes', 'output']
    assert issubclass(f, Filter)


def test_buffers_and_parallel_dimension():
    ds = generate_test_dataset(dims={'y': 20, 'x': 30, 'time': 10})
    c = ConvolutionFilter(('y', 'x'), np.ones((5, 3)))
    assert (c._buffer('y'), c._buffer('x'), c._buffer('time')) == (2, 1, 0)
    assert c._parallel_dimension(ds) == 'time'
    assert ConvolutionFilter(('y', 'x', 'time'), np.ones((3, 3, 3)))._parallel_dimension(ds) == 'x'
    b = BoxcarFilter(('y', 'x'), w=5)
    assert b.kernel.shape == (5, 5) and b.kernel[0, 0] == 1.0 / 25 and b._buffer('y') == 2
    g = GaussianFilter(('y', 'x'), sigma=[1, 2.2])
    assert (g._buffer('y'), g._buffer('x'), g._buffer('time')) == (4, 9, 0)
    assert g._parallel_dimension(ds) == 'time'
    assert ConvolutionFilter().kernel.shape == (1, 1)
    for fn in (convolution, boxcar, gaussian):
        assert callable(fn)


def test_argument_errors_match_scipy_without_touching_the_gpu():
    a = np.zeros((6, 7))
    with pytest.raises(RuntimeError):
        _ndimage.convolve(a, np.ones(3))                              # weights rank
    with pytest.raises(ValueError):
        _ndimage.convolve(a, np.ones((3, 3)), origin=2)               # invalid origin
    with pytest.raises(TypeError):
        _ndimage.convolve(a + 1j, np.ones((3, 3)))                    # complex
    with pytest.raises(RuntimeError):
        _ndimage.correlate1d(a, np.ones((2, 2)))
    with pytest.raises(ValueError):
        _ndimage.correlate1d(a, np.ones(3), axis=5)


def test_compute_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        _ndimage.convolve(np.zeros((6, 7)), np.ones((3, 3)))
    with pytest.raises(RuntimeError):
        GaussianFilter(('y', 'x')).apply(generate_test_dataset(dims={'y': 8, 'x': 8, 'time': 2}))


# ---- CPU: the restated algorithm is pinned against scipy itself, bit for bit --------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_numpy_restatement_of_ni_filters_equals_scipy(dtype):
    """oracle/ndimage_numpy.py is the statement of scipy's NI_Correlate / NI_Correlate1D / NI_ExtendLine the CUDA
    kernels were written from; it must reproduce scipy exactly (summation order, separate multiply and add,
    boundary extension, origins, kernel flip)."""
    from oracle import ndimage_numpy as on
    rng = np.random.default_rng(0)
    a = rng.normal(size=(7, 9, 5)).astype(dtype)
    k = rng.random((3, 4, 1))
    k[1, 2, 0] = 0.0
    for mode in ("reflect", "constant", "nearest", "mirror", "wrap"):
        for origin in (0, (1, -1, 0)):
            assert np.array_equal(on.correlate(a, k, mode, 0.7, origin), sn.correlate(a, k, mode=mode, cval=0.7, origin=origin))
            assert np.array_equal(on.convolve(a, k, mode, 0.7, origin), sn.convolve(a, k, mode=mode, cval=0.7, origin=origin))
        w_long = rng.random(13)                                       # longer than two of the axes
        for axis in range(3):
            for w in (w_long, np.array([1.0, 2.0, 3.0, 2.0, 1.0]), np.array([-1.0, 0.0, 1.0]), np.array([0.2, 0.5, 0.1, 0.9])):
                for origin in (0, 1, -1):
                    assert np.array_equal(on.correlate1d(a, w, axis, mode, -0.3, origin),
                                          sn.correlate1d(a, w, axis=axis, mode=mode, cval=-0.3, origin=origin)), (mode, axis, len(w), origin)
    for sigma in (1, [1.5, 0, 0.8], 2.0):
        assert np.array_equal(on.gaussian_filter(a, sigma), sn.gaussian_filter(a, sigma))


# ---- GPU: bit-exact parity with scipy ------------------------------------------------------------------
def _rand(shape, dtype, seed=0):
    return np.random.default_rng(seed).normal(size=shape).astype(dtype)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape,kshape", [((40,), (5,)), ((23, 31), (3, 5)), ((9, 14, 11), (3, 1, 5)), ((5, 6, 7, 8), (3, 3, 1, 3)),
                                          ((17, 19), (4, 2)), ((6, 5), (9, 13))])
def test_convolve_and_correlate_equal_scipy(dtype, shape, kshape):
    a = _rand(shape, dtype, seed=len(shape))
    k = np.random.default_rng(1).random(kshape)
    assert np.array_equal(_ndimage.convolve(a, k), sn.convolve(a, k))
    assert np.array_equal(_ndimage.correlate(a, k), sn.correlate(a, k))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["reflect", "constant", "nearest", "mirror", "wrap", "grid-wrap", "grid-constant", "grid-mirror"])
def test_boundary_modes_and_origins_equal_scipy(mode):
    a = _rand((7, 12), np.float64, seed=3)
    k = np.random.default_rng(2).random((5, 4))
    k[1, 2] = 0.0                                                     # a hole in the footprint
    for origin in (0, (1, -1), (-2, 1)):
        assert np.array_equal(_ndimage.convolve(a, k, mode=mode, cval=1.5, origin=origin),
                              sn.convolve(a, k, mode=mode, cval=1.5, origin=origin)), (mode, origin)
        assert np.array_equal(_ndimage.correlate(a, k, mode=mode, cval=-2.0, origin=origin),
                              sn.correlate(a, k, mode=mode, cval=-2.0, origin=origin)), (mode, origin)
    b = _rand((9, 7, 5), np.float32, seed=4)                          # dense 3x3 / 3x3x3 register-window kernels, short axes
    k3 = np.random.default_rng(6).random((3, 3, 3))
    for origin in (0, (1, -1, 1), (-1, 0, -1)):
        assert np.array_equal(_ndimage.correlate(b, k3, mode=mode, cval=0.5, origin=origin),
                              sn.correlate(b, k3, mode=mode, cval=0.5, origin=origin)), (mode, origin)
        o2 = origin if origin == 0 else (origin[0], origin[1], 0)
        assert np.array_equal(_ndimage.convolve(b, k3[:, :, :1], mode=mode, cval=0.5, origin=o2),
                              sn.convolve(b, k3[:, :, :1], mode=mode, cval=0.5, origin=o2)), (mode, o2)
    w = np.random.default_rng(5).random(15)                           # kernel longer than the axis: repeated extension
    for axis in (0, 1):
        for origin in (0, 3, -7):
            assert np.array_equal(_ndimage.correlate1d(a, w, axis=axis, mode=mode, cval=0.25, origin=origin),
                                  sn.correlate1d(a, w, axis=axis, mode=mode, cval=0.25, origin=origin)), (mode, axis, origin)


@pytest.mark.gpu
def test_correlate1d_symmetric_antisymmetric_general_equal_scipy():
    a = _rand((33, 18, 5), np.float64, seed=7)
    sym = np.array([1.0, 4.0, 6.0, 4.0, 1.0]) / 16
    anti = np.array([-1.0, -2.0, 0.0, 2.0, 1.0])
    gen = np.array([0.1, 0.7, 0.2, 0.5])
    for w in (sym, anti, gen):
        for axis in range(3):
            assert np.array_equal(_ndimage.correlate1d(a, w, axis=axis), sn.correlate1d(a, w, axis=axis))
    out = np.empty_like(a)
    assert _ndimage.correlate1d(a, sym, axis=-1, output=out) is out
    assert np.array_equal(out, sn.correlate1d(a, sym, axis=-1))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_gaussian_filter_equals_scipy(dtype):
    a = _rand((24, 37, 9), dtype, seed=11)
    for kw in (dict(sigma=1), dict(sigma=[1.5, 0, 0.8]), dict(sigma=2.0, order=[0, 1, 0]), dict(sigma=1.2, truncate=2.5),
               dict(sigma=[1, 2, 0], mode='nearest'), dict(sigma=1.0, mode=['reflect', 'wrap', 'constant'], cval=3.0),
               dict(sigma=[0, 0, 0]), dict(sigma=1.0, radius=[2, 3, 1]), dict(sigma=1.0, axes=(0, 2))):
        assert np.array_equal(_ndimage.gaussian_filter(a, **kw), sn.gaussian_filter(a, **kw)), kw
    assert np.array_equal(_ndimage.gaussian_filter1d(a, 1.7, axis=1), sn.gaussian_filter1d(a, 1.7, axis=1))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.int32, np.uint16])
def test_integer_rasters_equal_scipy(dtype):
    """scipy accepts integer rasters (double arithmetic, output stored in the input dtype with a C cast -- for the
    Gaussian after EVERY 1-D pass); ADVICE r1 asked for the same instead of a TypeError."""
    rng = np.random.default_rng(7)
    hi = 200 if dtype == np.uint8 else 3000
    a = rng.integers(0, hi, size=(40, 33, 5)).astype(dtype)
    k = rng.uniform(0, 1, size=(3, 5, 1))
    k /= k.sum()
    assert np.array_equal(_ndimage.convolve(a, k), sn.convolve(a, k))
    assert np.array_equal(_ndimage.correlate1d(a, [0.25, 0.5, 0.25], axis=1), sn.correlate1d(a, [0.25, 0.5, 0.25], axis=1))
    got, want = _ndimage.gaussian_filter(a, [1.0, 1.5, 0]), sn.gaussian_filter(a, [1.0, 1.5, 0])
    assert got.dtype == dtype and np.array_equal(got, want)
    out = np.empty_like(a)
    _ndimage.gaussian_filter(a, 1.2, output=out, mode='nearest')
    assert np.array_equal(out, sn.gaussian_filter(a, 1.2, mode='nearest'))


@pytest.mark.gpu
def test_strided_views_and_output_argument():
    base = _rand((12, 20, 6), np.float64, seed=13)
    a = base.transpose(2, 0, 1)[:, ::2, 1:]                           # non-contiguous view
    k = np.random.default_rng(3).random((3, 3, 3))
    out = np.zeros((6, 6, 40))[:, :, ::2][:, :, :19]                  # strided output
    assert out.shape == a.shape
    _ndimage.convolve(a, k, output=out)
    assert np.array_equal(out, sn.convolve(a, k))
    with pytest.raises(RuntimeError):
        _ndimage.convolve(a, k, output=np.zeros((2, 2, 2)))


@pytest.mark.gpu
def test_device_tensors_with_arbitrary_strides():
    import torch
    base = torch.from_numpy(_rand((10, 16, 12), np.float32, seed=17)).cuda()
    t_in = base.permute(2, 0, 1)                                      # variable-major style view
    t_out = torch.empty((16, 12, 10), dtype=torch.float32, device='cuda').permute(1, 2, 0)
    k = np.random.default_rng(4).random((3, 5, 1))
    _ndimage.correlate_device(t_in, t_out, k, [0, 0, 0])
    assert np.array_equal(t_out.cpu().numpy(), sn.correlate(t_in.cpu().numpy(), k))
    g_out = torch.empty_like(t_out)
    _ndimage.gaussian_filter_device(t_in, g_out, [1.0, 0.0, 2.0])
    assert np.array_equal(g_out.cpu().numpy(), sn.gaussian_filter(t_in.cpu().numpy(), [1.0, 0.0, 2.0]))
    with pytest.raises(ValueError):
        _ndimage.correlate_device(t_in, t_in, k, [0, 0, 0])           # in place is not supported


# ---- GPU: the reference's own Dataset-level tests ----------------------------------------------------
identity_kernel = np.zeros((3, 3))
identity_kernel[1, 1] = 1


@pytest.mark.gpu
def test_convolve_dataset_identity_and_kernel():
    """nd/tests/test_convolution_filter.py:34-48"""
    ds = generate_test_dataset()
    assert ConvolutionFilter(('y', 'x'), identity_kernel).apply(ds).equals(ds)
    kernel = np.random.default_rng(42).random((5, 5))
    nd_kernel = _expand_kernel(kernel, ('y', 'x'), ds['C11'].dims)
    assert np.array_equal(ConvolutionFilter(('y', 'x'), kernel).apply(ds)['C11'].values, sn.convolve(ds['C11'].values, nd_kernel))
    assert np.array_equal(convolution(ds, dims=('y', 'x'), kernel=kernel)['C11'].values, sn.convolve(ds['C11'].values, nd_kernel))
    assert np.array_equal(ConvolutionFilter(('y', 'x'), kernel, mode='wrap').apply(ds)['C22'].values,
                          sn.convolve(ds['C22'].values, nd_kernel, mode='wrap'))


@pytest.mark.gpu
def test_convolve_complex():
    """nd/tests/test_convolution_filter.py:51-57: complex variables go through real and imaginary parts."""
    from nd_b200.filters import assemble_complex
    ds = generate_test_dataset()
    assemble_complex(ds)
    assert np.iscomplexobj(ds['C12'].values)
    assert ConvolutionFilter(('y', 'x'), identity_kernel).apply(ds).equals(ds)
    kernel = np.random.default_rng(1).random((3, 3))
    res = ConvolutionFilter(('y', 'x'), kernel).apply(ds)
    nd_kernel = _expand_kernel(kernel, ('y', 'x'), ds['C12'].dims)
    assert np.array_equal(res['C12'].values.real, sn.convolve(ds['C12'].values.real, nd_kernel))
    assert np.array_equal(res['C12'].values.imag, sn.convolve(ds['C12'].values.imag, nd_kernel))


@pytest.mark.gpu
def test_boxcar():
    """nd/tests/test_convolution_filter.py:60-67"""
    ds = generate_test_dataset()
    w = 5
    assert BoxcarFilter(('y', 'x'), w).apply(ds).equals(ConvolutionFilter(('y', 'x'), np.ones((w, w)) / w**2).apply(ds))
    assert np.array_equal(boxcar(ds, dims=('y', 'x', 'time'), w=3)['C11'].values,
                          sn.convolve(ds['C11'].values, np.ones((3, 3, 3)) / 27))


@pytest.mark.gpu
def test_gaussian_dataset():
    """nd/tests/test_gaussian_filter.py:10-27"""
    ds = generate_test_dataset()
    res = GaussianFilter(dims=('y', 'x', 'time'), sigma=1).apply(ds)
    assert np.array_equal(res['C11'].values, sn.gaussian_filter(ds['C11'].values, sigma=1))
    res2 = GaussianFilter(dims=('y', 'x'), sigma=1).apply(ds)
    t = list(ds['C11'].dims).index('time')
    sl = [slice(None)] * 3
    sl[t] = 0
    assert np.array_equal(res2['C11'].values[tuple(sl)], sn.gaussian_filter(ds['C11'].values[tuple(sl)], sigma=1))


@pytest.mark.gpu
@pytest.mark.parametrize("f", [ConvolutionFilter, BoxcarFilter, GaussianFilter])
def test_filters_common(f):
    """nd/tests/test_filters_common.py:20-60: output type / dims / shape, dims-order invariance, njobs invariance."""
    ds = generate_test_dataset(dims={'y': 20, 'x': 30, 'time': 10})
    result = f(dims=('y', 'x')).apply(ds)
    assert isinstance(result, Dataset)
    for v in ds.data_vars:
        assert ds[v].dims == result[v].dims and ds[v].shape == result[v].shape
    swapped = f(dims=('x', 'y')).apply(ds)
    for v in ds.data_vars:
        assert np.allclose(result[v].values, swapped[v].values, rtol=1e-5)
    for dims in [('x', 'y'), ('x', 'y', 'time')]:
        a, b = f(dims=dims).apply(ds), f(dims=dims).apply(ds, njobs=2)
        for v in ds.data_vars:
            assert np.allclose(a[v].values, b[v].values, rtol=1e-5)
