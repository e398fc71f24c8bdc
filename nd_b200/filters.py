"""
`Filter` / `NLMeansFilter`, mirroring reference nd/filters.py:82-198 and :388-466, and the sibling filters
`ConvolutionFilter` / `BoxcarFilter` / `GaussianFilter` (nd/filters.py:205-381), whose
scipy.ndimage calls are replaced by the CUDA kernels behind `nd_b200._ndimage`.

Same public surface: `NLMeansFilter(dims, r, sigma, h, f, n_eff).apply(ds, inplace=False, njobs=1)`
takes and returns a Dataset with the same dims, coords and attrs; the filter plugs in through
`Filter._filter(self, arr, axes, output)` (signature pinned by the reference's
nd/tests/test_filters_common.py:37-41).  Below that boundary the reference calls its Cython kernel;
here `_filter` calls `nd_b200._filters._pixelwise_nlmeans_3d`, which runs on the GPU.

Datasets: anything with the small xarray subset used here (`data_vars`, `ds[name].dims/.values`,
`copy(deep=True)`, item assignment / deletion): real `xarray.Dataset`s when xarray is installed,
`nd_b200.dataset.Dataset` otherwise (xarray is absent in this image).
"""
import re
from abc import abstractmethod

import numpy as np

from .algorithm import Algorithm, parallelize, wrap_algorithm
from ._filters import _pixelwise_nlmeans_3d, nlmeans_variables
from . import _ndimage as snf          # GPU stand-in for `scipy.ndimage.filters` (same call signatures)


# ---- helpers restated from nd/utils.py:450-524 and nd/io.py:26-123 ---------------------------
def _is_dataarray(ds):
    return hasattr(ds, 'values') and hasattr(ds, 'dims') and not hasattr(ds, 'data_vars')


def get_vars_for_dims(ds, dims, invert=False):
    """All variables of `ds` whose dims contain `dims` (reference nd/utils.py:450-469)."""
    return [v for v in ds.data_vars if set(ds[v].dims).issuperset(set(dims)) != invert]


def is_complex(ds):
    """reference nd/utils.py:502-524"""
    if _is_dataarray(ds):
        return np.iscomplexobj(ds.values)
    return bool(np.any([np.iscomplexobj(ds[v].values) for v in ds.data_vars]))


def disassemble_complex(ds):
    """Split complex variables into `__re` / `__im` in place (reference nd/io.py:26-69)."""
    for vn in list(ds.data_vars):
        var = ds[vn]
        if not np.iscomplexobj(var.values):
            continue
        ds[vn + '__re'] = (var.dims, np.ascontiguousarray(var.values.real))
        ds[vn + '__im'] = (var.dims, np.ascontiguousarray(var.values.imag))
        del ds[vn]


def assemble_complex(ds):
    """Inverse of `disassemble_complex`, in place (reference nd/io.py:72-123)."""
    names = list(ds.data_vars)
    stems = {}
    for vn in names:
        m = re.match(r'(?P<stem>.*)(?:_real|__re)$', vn)
        if m:
            stems.setdefault(m.group('stem'), {})['re'] = vn
        m = re.match(r'(?P<stem>.*)(?:_imag|__im)$', vn)
        if m:
            stems.setdefault(m.group('stem'), {})['im'] = vn
    for stem, parts in stems.items():
        if 're' in parts and 'im' in parts:
            re_v, im_v = ds[parts['re']], ds[parts['im']]
            ds[stem] = (re_v.dims, re_v.values + im_v.values * 1j)
            del ds[parts['re']]
            del ds[parts['im']]


def _expand_kernel(kernel, kernel_dims, new_dims):
    """
    Reshape a kernel spanning some dimensions to cover a superset of dimensions
    (reference nd/filters.py:36-75).
    """
    if not set(new_dims).issuperset(set(kernel_dims)):
        raise ValueError('`new_dims` must be a superset of `kernel_dims`.')
    if kernel.ndim != len(kernel_dims):
        raise ValueError('The length of `kernel_dims` must match the dimension of `kernel`.')
    new_kernel_shape = np.ones(len(new_dims), dtype=int)
    new_kernel_shape[[new_dims.index(_) for _ in kernel_dims]] = kernel.shape
    return kernel.reshape(new_kernel_shape)


def _dim_sizes(ds):
    sizes = {}
    if _is_dataarray(ds):
        return dict(zip(ds.dims, ds.values.shape))
    for v in ds.data_vars:
        for d, n in zip(ds[v].dims, ds[v].values.shape):
            sizes[d] = n
    return sizes


def _largest_extra_dimension(ds, dims):
    """Largest dimension that is not filtered, else the largest dimension (nd/filters.py:232-243, :345-356)."""
    sizes = _dim_sizes(ds)
    extra = [d for d in sizes if d not in dims]
    pool = extra if extra else list(sizes)
    return sorted(pool, key=lambda d: sizes[d], reverse=True)[0]


class Filter(Algorithm):
    """
    The base class for a generic filter (reference nd/filters.py:82-198).

    Parameters
    ----------
    dims : tuple of str
        The dimensions along which the filter is applied.
    """

    # If per_variable is True, the filter is applied independently for each variable.
    per_variable = True
    # If supports_complex is False, complex variables are split into two reals before filtering.
    supports_complex = False
    dims = ()

    @abstractmethod
    def __init__(self, *args, **kwargs):
        return

    @parallelize
    def apply(self, ds, inplace=False):
        """
        Apply the filter to the input dataset.

        Parameters
        ----------
        ds : xarray.Dataset
            The input dataset
        inplace : bool, optional
            If True, overwrite the input data inplace (default: False).

        Returns
        -------
        xarray.Dataset
            The filtered dataset
        """
        if inplace:
            raise NotImplementedError('Inplace filtering is not currently implemented.')

        convert_complex = is_complex(ds) and not self.supports_complex
        if convert_complex:
            disassemble_complex(ds)

        if _is_dataarray(ds):
            # DataArray: the raw values go to the filter, whose last axis then plays "variables"
            # (quirk of the reference preserved, nd/filters.py:139-145 + :447-463).
            result = ds.copy(deep=True)
            axes = tuple(list(result.dims).index(d) for d in self.dims)
            self._filter(ds.values, axes, output=result.values)
        else:
            variables = get_vars_for_dims(ds, self.dims)
            result = None
            if not self.per_variable and variables:
                result = self._apply_joint_streamed(ds, variables)
            if result is not None:
                pass
            elif self.per_variable:
                result = ds.copy(deep=True)
                for v in variables:
                    vdims = result[v].dims
                    axes = tuple(list(vdims).index(d) for d in self.dims)
                    self._filter(ds[v].values, axes, output=result[v].values)
            else:
                result = ds.copy(deep=True)
                if variables:
                    self._apply_joint(ds, result, variables)

        if convert_complex:
            assemble_complex(ds)
        return result

    def _apply_joint_streamed(self, ds, variables):
        """per_variable=False without the gather / scatter of `_apply_joint`: when every filtered variable already
        has the layout the kernel wants (dims == filter dims followed by the others, C order, one float dtype) and
        the filter can take the variables one array each (`_filter_variables`), they are streamed to the GPU straight
        from the Dataset's own arrays and the result is written straight into the new Dataset's arrays -- no
        `to_array()` block, no deep copy of what is about to be overwritten, no `xr.merge` (nd/filters.py:164-185).
        Returns the result Dataset, or None when the fast path does not apply (then `_apply_joint` runs)."""
        fv = getattr(self, '_filter_variables', None)
        if fv is None:
            return None
        vdims = tuple(ds[variables[0]].dims)
        ordered_dims = tuple(self.dims) + tuple(d for d in vdims if d not in self.dims)
        srcs = [ds[v].values for v in variables]
        if (vdims != ordered_dims or any(tuple(ds[v].dims) != vdims for v in variables)
                or any(not isinstance(a, np.ndarray) or a.dtype != srcs[0].dtype or a.shape != srcs[0].shape
                       or not a.flags['C_CONTIGUOUS'] for a in srcs)
                or srcs[0].dtype not in (np.float32, np.float64)):
            return None
        outs = [np.empty_like(a) for a in srcs]
        axes = tuple(ordered_dims.index(d) for d in self.dims)
        if not fv(srcs, axes, outs):
            return None
        result = ds.copy(deep=False)
        for v in list(ds.data_vars):
            if v in variables:
                result[v] = ds[v].copy(deep=False, data=outs[variables.index(v)])
            else:
                result[v] = ds[v].copy(deep=True)
        return result

    def _apply_joint(self, ds, result, variables):
        """per_variable=False: the variables are an extra trailing axis (nd/filters.py:164-185)."""
        vdims = tuple(ds[variables[0]].dims)
        for v in variables[1:]:
            if set(ds[v].dims) != set(vdims):
                raise ValueError('all filtered variables must share the same dimensions '
                                 '(%r has %r, %r has %r)' % (variables[0], vdims, v, tuple(ds[v].dims)))
        ordered_dims = tuple(self.dims) + tuple(d for d in vdims if d not in self.dims)
        dtype = np.result_type(*[ds[v].values.dtype for v in variables])
        shape = tuple(ds[variables[0]].values.shape[vdims.index(d)] for d in ordered_dims)
        # variable-major block like `to_array()`; the (dims..., variable) view handed to `_filter`
        # is a transposed view of it (nd/filters.py:170)
        block = np.empty((len(variables),) + shape, dtype=dtype)
        for i, v in enumerate(variables):
            src = ds[v].values
            perm = [list(ds[v].dims).index(d) for d in ordered_dims]
            block[i] = np.transpose(src, perm)
        out_block = block.copy()
        arr = np.moveaxis(block, 0, -1)
        out = np.moveaxis(out_block, 0, -1)
        axes = tuple(ordered_dims.index(d) for d in self.dims)
        self._filter(arr, axes, output=out)
        for i, v in enumerate(variables):
            dst_dims = list(result[v].dims)
            perm = [ordered_dims.index(d) for d in dst_dims]
            result[v].values[...] = np.transpose(out_block[i], perm)

    @abstractmethod
    def _filter(self, arr, axes, output=None):
        """This method must be implemented by all derived classes."""
        return


# ------------------------------------------------------------------------------------------------
# Sibling filters (reference nd/filters.py:205-381).  Same classes, attributes and `_filter` bodies; `snf`
# is nd_b200._ndimage, whose `convolve` / `gaussian_filter` run on the GPU and match scipy bit for bit.
# ------------------------------------------------------------------------------------------------
class ConvolutionFilter(Filter):
    """
    Kernel-convolution of a Dataset (reference nd/filters.py:205-268).

    Parameters
    ----------
    dims : tuple, optional
        The dataset dimensions corresponding to the kernel axes (default: ('y', 'x')).
    kernel : ndarray
        The convolution kernel.
    kwargs : dict, optional
        Extra keyword arguments with the meaning of ``scipy.ndimage.convolve`` (mode, cval, origin).
    """

    per_variable = True
    supports_complex = True
    kwargs = {}

    def __init__(self, dims=('y', 'x'), kernel=None, **kwargs):
        if kernel is None:
            kernel = np.ones([1] * len(dims))
        self.dims = tuple(dims)
        self.kernel = kernel
        self.kwargs = kwargs

    def _parallel_dimension(self, ds):
        return _largest_extra_dimension(ds, self.dims)

    def _buffer(self, dim):
        if dim not in self.dims:
            return 0
        axis = self.dims.index(dim)
        return self.kernel.shape[axis] // 2

    def _filter(self, arr, axes, output):
        # Reshape kernel to match dimension of input array (a reshape, not a transpose: nd/filters.py:256-259).
        new_kernel_shape = np.ones(arr.ndim, dtype=int)
        new_kernel_shape[list(axes)] = self.kernel.shape
        nd_kernel = self.kernel.reshape(new_kernel_shape)
        if np.iscomplexobj(arr):
            snf.convolve(np.real(arr), nd_kernel, output=np.real(output), **self.kwargs)
            snf.convolve(np.imag(arr), nd_kernel, output=np.imag(output), **self.kwargs)
        else:
            snf.convolve(arr, nd_kernel, output=output, **self.kwargs)


convolution = wrap_algorithm(ConvolutionFilter, 'convolution')


class BoxcarFilter(ConvolutionFilter):
    """
    A boxcar filter of odd width `w` along `dims` (reference nd/filters.py:277-301).
    """

    def __init__(self, dims=('y', 'x'), w=3, **kwargs):
        N = len(dims)
        self.dims = tuple(dims)
        self.kernel = np.ones((w,) * N, dtype=np.float64) / w**N
        self.kwargs = kwargs


boxcar = wrap_algorithm(BoxcarFilter, 'boxcar')


class GaussianFilter(Filter):
    """
    A Gaussian filter (reference nd/filters.py:310-378).

    Parameters
    ----------
    dims : tuple of str, optional
        The dimensions along which to apply the Gaussian filtering (default: ('y', 'x')).
    sigma : float or sequence of float
        The standard deviation for the Gaussian kernel, per dimension if a sequence.
    kwargs : dict, optional
        Extra keyword arguments with the meaning of ``scipy.ndimage.gaussian_filter``.
    """

    def __init__(self, dims=('y', 'x'), sigma=1, **kwargs):
        if isinstance(sigma, (int, float)):
            sigma = [sigma] * len(dims)
        self.dims = tuple(dims)
        self.sigma = sigma
        self.kwargs = kwargs

    def _parallel_dimension(self, ds):
        return _largest_extra_dimension(ds, self.dims)

    def _buffer(self, dim):
        if dim not in self.dims:
            return 0
        # the kernel is truncated after 4 sigma by default (nd/filters.py:362-368)
        axis = self.dims.index(dim)
        sigma = self.sigma[axis]
        truncate = 4.0
        return int(truncate * sigma + 0.5)

    def _filter(self, arr, axes, output):
        # Generate n-dimensional sigma
        ndsigma = [0] * arr.ndim
        for ax, s in zip(axes, self.sigma):
            ndsigma[ax] = s
        if np.iscomplexobj(arr):
            # (unreachable through `apply`: supports_complex is False; kept as in the reference, including
            #  its write of the imaginary part into np.real(output), nd/filters.py:373-376)
            snf.gaussian_filter(np.real(arr), sigma=ndsigma, output=np.real(output), **self.kwargs)
            snf.gaussian_filter(np.imag(arr), sigma=ndsigma, output=np.real(output), **self.kwargs)
        else:
            snf.gaussian_filter(arr, sigma=ndsigma, output=output, **self.kwargs)


gaussian = wrap_algorithm(GaussianFilter, 'gaussian')


class NLMeansFilter(Filter):
    """
    Non-Local Means (Buades2011), reference nd/filters.py:388-466.

    Parameters
    ----------
    dims : tuple of str
        The dataset dimensions along which to filter.
    r : {int, sequence}
        The radius
    sigma : float
        The standard deviation of the noise present in the data.
    h : float
    f : int
    n_eff : float, optional
        The desired effective sample size (-1: none, default).
    semantics : {'as_written', 'reference_compiled'}, keyword-only, optional
        Patch-distance semantics.  Default 'as_written' (or ND_NLM_SEMANTICS).
    """

    per_variable = False
    _supports_njobs = True      # njobs = GPUs: y-shards with an r+f halo (nd_b200/shard.py, nd_b200/_filters.py)

    def __init__(self, dims=('y', 'x'), r=1, sigma=1, h=1, f=1, n_eff=-1, *, semantics=None, kernel='auto'):
        if isinstance(r, (int, float)):
            r = [r] * len(dims)
        self.dims = tuple(dims)
        self.r = np.array(r, dtype=np.uint32)
        self.f = np.array([f if _ > 0 else 0 for _ in self.r], dtype=np.uint32)
        self.sigma = sigma
        self.h = h
        self.n_eff = n_eff
        self.semantics = semantics
        self.kernel = kernel
        self._njobs = 1

    def _parallel_dimension(self, ds):
        """Largest non-filter dimension, else the largest dimension (nd/filters.py:424-435)."""
        sizes = {}
        for v in ds.data_vars:
            for d, n in zip(ds[v].dims, ds[v].values.shape):
                sizes[d] = n
        extra = [d for d in sizes if d not in self.dims]
        pool = extra if extra else list(sizes)
        return sorted(pool, key=lambda d: sizes[d], reverse=True)[0]

    def _buffer(self, dim):
        """Halo needed when sharding over `dim`: r + f on that axis (nd/filters.py:437-445)."""
        if dim not in self.dims:
            return 0
        axis = self.dims.index(dim)
        return int(self.r[axis] + self.f[axis])

    def _filter(self, arr, axes, output):
        # Pad r and f to three dimensions (nd/filters.py:451-454); `axes` is ignored as in the reference.
        pad_before = np.zeros(4 - arr.ndim, dtype=self.r.dtype)
        pad_after = np.zeros(arr.ndim - len(self.r) - 1, dtype=self.r.dtype)
        r = np.concatenate([pad_before, self.r, pad_after])
        f = np.concatenate([pad_before, self.f, pad_after])
        # Pad input and output to four dimensions (views, never copies) (nd/filters.py:459-460).
        values = arr.reshape((1,) * (4 - arr.ndim) + arr.shape)
        _out = output.reshape((1,) * (4 - output.ndim) + output.shape)
        if not np.shares_memory(_out, output):
            raise ValueError('output must be viewable as a 4-D array without copying')
        if len(self.r) == 0 or not np.any(r):
            # r == 0 everywhere: no neighbours, self weight 1 -> exact identity (nd/_filters.pyx:406-420;
            # pinned exactly by nd/tests/test_nlmeans_filter.py:17-25).  Nothing to compute.
            if self.n_eff < 0:
                _out[...] = values
                return
        njobs, shard_axis = self._njobs_and_shard_axis(len(pad_before))
        _pixelwise_nlmeans_3d(values, _out, r, f, self.sigma, self.h, self.n_eff,
                              semantics=self.semantics, kernel=self.kernel, njobs=njobs, shard_axis=shard_axis)

    def _njobs_and_shard_axis(self, leading):
        njobs = getattr(self, '_njobs', 1)
        shard_axis = None
        shard_dim = getattr(self, '_shard_dim', None)
        if njobs != 1 and shard_dim in self.dims:
            # `_parallel_dimension` picked a filtered dimension (there is no other one): its position among the
            # kernel axes.  A non-filtered dimension is found by the same rule one level down (largest free axis).
            shard_axis = leading + self.dims.index(shard_dim)
        return njobs, shard_axis

    def _filter_variables(self, arrays, axes, outputs):
        """`_filter` for V separate per-variable arrays (same shape / dtype, C order, dims == filter dims followed by
        the non-filter dims).  Returns False when this layout cannot be streamed (the caller then gathers the
        variables into one block and calls `_filter`)."""
        ndim = arrays[0].ndim
        if ndim > 3 or ndim < len(self.r) or not np.any(self.r) or len(self.r) == 0:
            return False
        # Missing axes are singletons.  `_filter` puts them in front (nd/filters.py:451-460); here they go BEHIND the
        # data axes, which is the same filter (a singleton axis carries no neighbours) but leaves the first data
        # axis -- 'y' of a 2-D image -- as axis 0, the axis the slab pipeline and the GPU shards cut along.
        pad_after = np.zeros(3 - len(self.r), dtype=self.r.dtype)
        r = np.concatenate([self.r, pad_after])
        f = np.concatenate([self.f, pad_after])
        njobs, shard_axis = self._njobs_and_shard_axis(0)
        if shard_axis not in (None, 0):
            return False
        a3 = [a.reshape(a.shape + (1,) * (3 - ndim)) for a in arrays]
        o3 = [o.reshape(o.shape + (1,) * (3 - ndim)) for o in outputs]
        if r[0] == 0 and f[0] == 0 and njobs > 1 and a3[0].shape[0] < njobs:
            return False                      # a short free leading axis: shard along a filtered one (block path)
        return nlmeans_variables(a3, o3, r, f, self.sigma, self.h, self.n_eff, semantics=self.semantics,
                                 kernel=self.kernel, njobs=njobs)


nlmeans = wrap_algorithm(NLMeansFilter, 'nlmeans')
