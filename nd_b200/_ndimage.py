"""
GPU stand-ins for the scipy.ndimage calls the reference's sibling filters make:

    snf.convolve(arr, nd_kernel, output=output, **kwargs)                 nd/filters.py:260-268
    snf.gaussian_filter(arr, sigma=ndsigma, output=output, **kwargs)      nd/filters.py:370-378

Same names, argument meaning and error behaviour as scipy.ndimage (`convolve`, `correlate`, `correlate1d`,
`gaussian_filter1d`, `gaussian_filter`); the Python-level argument handling (kernel flip and origin
negation for convolutions, Gaussian kernel construction, per-axis sequencing) restates scipy's
`ndimage/_filters.py`, the arithmetic runs in the CUDA kernels behind include/ndflt.h, which reproduce
scipy's C loops operation by operation: float64 results are bit-identical to scipy's.

Host arrays go H2D -> kernels -> D2H; CUDA tensors are filtered in place on the device
(`*_device` functions).  There is no CPU fallback: without the CUDA library or a GPU these raise.
"""
import ctypes
import numbers

import numpy as np
import torch

from . import _lib

MODES = {'reflect': 0, 'grid-mirror': 0, 'constant': 1, 'grid-constant': 1, 'nearest': 2, 'mirror': 3,
         'wrap': 4, 'grid-wrap': 4}
_DT = {torch.float32: 0, torch.float64: 1}


def _mode(mode):
    if mode not in MODES:
        raise RuntimeError('boundary mode not supported')           # scipy _ni_support._extend_mode_to_code
    return MODES[mode]


def _normalize_sequence(value, rank):
    """scipy _ni_support._normalize_sequence"""
    if hasattr(value, '__iter__') and not isinstance(value, str):
        seq = list(value)
        if len(seq) != rank:
            raise RuntimeError('sequence argument must have length equal to input rank')
        return seq
    return [value] * rank


def _pad4(t):
    """(tensor viewed with 4 axes: leading axes of extent 1, element strides)."""
    if t.dim() > 4:
        raise RuntimeError('nd_b200 filters support arrays with at most 4 dimensions (got %d)' % t.dim())
    lead = 4 - t.dim()
    return [1] * lead + list(t.shape), [0] * lead + list(t.stride()), lead


def _check_pair(t_in, t_out):
    if not (t_in.is_cuda and t_out.is_cuda):
        raise ValueError('expected CUDA tensors')
    if t_in.dtype not in _DT:
        raise TypeError('nd_b200 filters support float32 / float64 data only (got %s)' % t_in.dtype)
    if t_out.dtype != t_in.dtype:
        raise TypeError('output dtype %s differs from the input dtype %s' % (t_out.dtype, t_in.dtype))
    if tuple(t_in.shape) != tuple(t_out.shape):
        raise RuntimeError('output shape not correct')
    if t_in.numel() and t_in.untyped_storage().data_ptr() == t_out.untyped_storage().data_ptr():
        raise ValueError('input and output must not share memory')


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _check(rc):
    if rc == 0:
        return
    msg = _lib.lib().ndflt_last_error().decode('utf-8', 'replace')
    if rc == _lib.EDTYPE:
        raise TypeError(msg)
    if rc == _lib.EINVAL:
        raise ValueError(msg)
    raise RuntimeError('ndflt: ' + msg)


# ---- device level --------------------------------------------------------------------------------
def correlate_device(t_in, t_out, weights, origins, mode='reflect', cval=0.0):
    """N-D correlation of CUDA tensor `t_in` into `t_out` (ndflt_correlate).  `weights` has t_in.dim() axes."""
    _check_pair(t_in, t_out)
    weights = np.ascontiguousarray(weights, dtype=np.float64)
    if weights.ndim != t_in.dim():
        raise RuntimeError('filter weights array has incorrect shape.')
    shape, istr, lead = _pad4(t_in)
    _, ostr, _ = _pad4(t_out)
    kshape = [1] * lead + list(weights.shape)
    org = [0] * lead + [int(o) for o in origins]
    if t_in.numel() == 0:
        return t_out
    _check(_lib.lib().ndflt_correlate(ctypes.c_void_p(t_in.data_ptr()), ctypes.c_void_p(t_out.data_ptr()), _lib.i64(shape),
                                      _lib.i64(istr), _lib.i64(ostr), _DT[t_in.dtype],
                                      weights.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), _lib.i64(kshape),
                                      _lib.i64(org), _mode(mode), float(cval), _stream(t_out)))
    return t_out


def correlate1d_device(t_in, t_out, weights, axis, origin=0, mode='reflect', cval=0.0):
    """1-D correlation along `axis` of CUDA tensor `t_in` into `t_out` (ndflt_correlate1d)."""
    _check_pair(t_in, t_out)
    weights = np.ascontiguousarray(weights, dtype=np.float64)
    if weights.ndim != 1 or weights.shape[0] < 1:
        raise RuntimeError('no filter weights given')
    shape, istr, lead = _pad4(t_in)
    _, ostr, _ = _pad4(t_out)
    if t_in.numel() == 0:
        return t_out
    _check(_lib.lib().ndflt_correlate1d(ctypes.c_void_p(t_in.data_ptr()), ctypes.c_void_p(t_out.data_ptr()), _lib.i64(shape),
                                        _lib.i64(istr), _lib.i64(ostr), _DT[t_in.dtype], int(axis) + lead,
                                        weights.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), int(weights.shape[0]),
                                        int(origin), _mode(mode), float(cval), _stream(t_out)))
    return t_out


def _gaussian_kernel1d(sigma, order, radius):
    """The 1-D Gaussian (derivative) kernel of scipy.ndimage (`_filters._gaussian_kernel1d`)."""
    if order < 0:
        raise ValueError('order must be non-negative')
    exponent_range = np.arange(order + 1)
    sigma2 = sigma * sigma
    x = np.arange(-radius, radius + 1)
    phi_x = np.exp(-0.5 / sigma2 * x ** 2)
    phi_x = phi_x / phi_x.sum()
    if order == 0:
        return phi_x
    # f(x) = q(x) * phi(x), q a polynomial; f'(x) = (q'(x) + q(x) * p'(x)) * phi(x), p'(x) = -x / sigma^2
    q = np.zeros(order + 1)
    q[0] = 1
    D = np.diag(exponent_range[1:], 1)
    P = np.diag(np.ones(order) / -sigma2, -1)
    Q_deriv = D + P
    for _ in range(order):
        q = Q_deriv.dot(q)
    q = (x[:, None] ** exponent_range).dot(q)
    return q * phi_x


def gaussian_plan(ndim, sigma, order=0, mode='reflect', truncate=4.0, radius=None, axes=None):
    """[(axis, reversed weights, mode)] in execution order -- scipy `gaussian_filter` / `gaussian_filter1d`."""
    if axes is None:
        axes = list(range(ndim))
    elif isinstance(axes, numbers.Integral):
        axes = [int(axes)]
    axes = [a + ndim if a < 0 else a for a in axes]
    if any(a < 0 or a >= ndim for a in axes) or len(set(axes)) != len(axes):
        raise ValueError('invalid axes')
    n = len(axes)
    orders, sigmas = _normalize_sequence(order, n), _normalize_sequence(sigma, n)
    modes, radiuses = _normalize_sequence(mode, n), _normalize_sequence(radius, n)
    plan = []
    for axis, sg, od, md, rd in zip(axes, sigmas, orders, modes, radiuses):
        if not sg > 1e-15:
            continue
        sd = float(sg)
        lw = int(float(truncate) * sd + 0.5)              # make the radius of the filter equal to truncate std devs
        if rd is not None:
            lw = rd
        if not isinstance(lw, numbers.Integral) or lw < 0:
            raise ValueError('Radius must be a nonnegative integer.')
        weights = _gaussian_kernel1d(sd, int(od), int(lw))[::-1]   # correlation, not convolution: revert the kernel
        plan.append((axis, np.ascontiguousarray(weights), md))
    return plan


def gaussian_filter_device(t_in, t_out, sigma, order=0, mode='reflect', cval=0.0, truncate=4.0, radius=None, axes=None,
                           integer=False):
    """Sequence of 1-D correlations like scipy: every pass rounds to the array dtype; pass k reads the output
    of pass k-1.  Needs one scratch tensor because the kernels are out of place.  `integer`: the caller's array
    has an integer dtype and is carried as float64 here -- scipy stores every pass into the integer output with a
    C cast, i.e. truncation toward zero, which is applied between the passes."""
    _check_pair(t_in, t_out)
    plan = gaussian_plan(t_in.dim(), sigma, order, mode, truncate, radius, axes)
    if not plan:
        t_out.copy_(t_in)
        return t_out
    scratch = torch.empty_like(t_out) if len(plan) > 1 else None
    src = t_in
    for k, (axis, weights, md) in enumerate(plan):
        # the last pass must land in t_out; alternate between t_out and the scratch tensor before it
        dst = t_out if (len(plan) - 1 - k) % 2 == 0 else scratch
        correlate1d_device(src, dst, weights, axis, 0, md, cval)
        if integer:
            dst.trunc_()
        src = dst
    return t_out


# ---- host level (scipy.ndimage signatures) -------------------------------------------------------------
def _is_integer(a):
    return np.issubdtype(np.asarray(a).dtype, np.integer)


def _to_device(a):
    """float32 / float64 arrays go to the device as they are.  Integer rasters (uint8 / int16 ... as scipy accepts
    them, nd/filters.py:260-268) are carried as float64 -- scipy computes every output value in double and stores
    it into an array of the input dtype with a C cast (`_finish` does the same)."""
    a = np.asarray(a)
    if _is_integer(a):
        a = a.astype(np.float64)
    if a.dtype not in (np.float32, np.float64):
        raise TypeError('nd_b200 filters support float32 / float64 and integer data only (got %s)' % a.dtype)
    if not torch.cuda.is_available():
        raise RuntimeError('nd_b200 filters need a CUDA device; there is no CPU fallback')
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _finish(result, input, output):
    res = result.cpu().numpy()
    if _is_integer(input):
        with np.errstate(invalid='ignore'):
            res = np.trunc(res).astype(np.asarray(input).dtype)      # the C cast of scipy's output store
    if output is None:
        return res
    if not isinstance(output, np.ndarray):
        raise TypeError('output must be a numpy array (or None)')
    if output.shape != np.asarray(input).shape:
        raise RuntimeError('output shape not correct')
    output[...] = res
    return output


def _correlate_or_convolve(input, weights, output, mode, cval, origin, convolution):
    """scipy `_filters._correlate_or_convolve`."""
    input = np.asarray(input)
    if np.iscomplexobj(input):
        raise TypeError('Complex type not supported')
    origins = _normalize_sequence(origin, input.ndim)
    weights = np.asarray(weights, dtype=np.float64)
    wshape = [ii for ii in weights.shape if ii > 0]
    if len(wshape) != input.ndim:
        raise RuntimeError('filter weights array has incorrect shape.')
    if convolution:
        weights = weights[tuple([slice(None, None, -1)] * weights.ndim)]
        for ii in range(len(origins)):
            origins[ii] = -origins[ii]
            if not weights.shape[ii] & 1:
                origins[ii] -= 1
    for org, lenw in zip(origins, wshape):
        if (lenw // 2 + org < 0) or (lenw // 2 + org >= lenw):
            raise ValueError('Invalid origin; origin must satisfy -(weights.shape[k] // 2) <= origin[k] <= '
                             '(weights.shape[k]-1) // 2')
    t_in = _to_device(input)
    t_out = torch.empty_like(t_in)
    correlate_device(t_in, t_out, weights, origins, mode, cval)
    return _finish(t_out, input, output)


def correlate(input, weights, output=None, mode='reflect', cval=0.0, origin=0):
    return _correlate_or_convolve(input, weights, output, mode, cval, origin, False)


def convolve(input, weights, output=None, mode='reflect', cval=0.0, origin=0):
    return _correlate_or_convolve(input, weights, output, mode, cval, origin, True)


def correlate1d(input, weights, axis=-1, output=None, mode='reflect', cval=0.0, origin=0):
    input = np.asarray(input)
    if np.iscomplexobj(input):
        raise TypeError('Complex type not supported')
    weights = np.asarray(weights, dtype=np.float64)
    if weights.ndim != 1 or weights.shape[0] < 1:
        raise RuntimeError('no filter weights given')
    axis = axis + input.ndim if axis < 0 else axis
    if axis < 0 or axis >= input.ndim:
        raise ValueError('invalid axis')
    if (len(weights) // 2 + origin < 0) or (len(weights) // 2 + origin >= len(weights)):
        raise ValueError('Invalid origin; origin must satisfy -(len(weights) // 2) <= origin <= (len(weights)-1) // 2')
    t_in = _to_device(input)
    t_out = torch.empty_like(t_in)
    correlate1d_device(t_in, t_out, weights, axis, origin, mode, cval)
    return _finish(t_out, input, output)


def gaussian_filter1d(input, sigma, axis=-1, order=0, output=None, mode='reflect', cval=0.0, truncate=4.0, *,
                      radius=None):
    input = np.asarray(input)
    t_in = _to_device(input)
    t_out = torch.empty_like(t_in)
    gaussian_filter_device(t_in, t_out, sigma, order, mode, cval, truncate, radius, axes=axis, integer=_is_integer(input))
    return _finish(t_out, input, output)


def gaussian_filter(input, sigma, order=0, output=None, mode='reflect', cval=0.0, truncate=4.0, *, radius=None,
                    axes=None):
    input = np.asarray(input)
    if np.iscomplexobj(input):
        raise TypeError('Complex type not supported')
    t_in = _to_device(input)
    t_out = torch.empty_like(t_in)
    gaussian_filter_device(t_in, t_out, sigma, order, mode, cval, truncate, radius, axes, integer=_is_integer(input))
    return _finish(t_out, input, output)
