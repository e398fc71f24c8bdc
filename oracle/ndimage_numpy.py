"""
TEST INFRASTRUCTURE ONLY.  NumPy restatement of the scipy.ndimage arithmetic that the reference's sibling filters
run on (nd/filters.py:260-268 `snf.convolve`, :370-378 `snf.gaussian_filter`) -- the algorithm the CUDA kernels
behind include/ndflt.h were written from, pinned here against scipy itself (tests/test_sibling_filters.py compares
every function below with scipy.ndimage BIT FOR BIT on the CPU).

scipy `ndimage/src/ni_filters.c`, restated:
  NI_Correlate    out[p] = sum over the footprint taps (|w| > DBL_EPSILON) in C order of in[ext(p + k - size//2 - origin)] * w_k,
                  double accumulation from 0.0, one multiply and one add per tap (no fused multiply-add);
  NI_Correlate1D  odd symmetric kernels:  x[0] w[0] + sum_{j=-s1}^{-1} (x[j] + x[-j]) w[j]   (antisymmetric: x[j] - x[-j]),
                  other kernels:          x[s2] w[s2] + sum_{j=-s1}^{s2-1} x[j] w[j];
  NI_ExtendLine   reflect (d c b a | a b c d | d c b a), mirror, wrap, nearest, constant -- repeated for long kernels.
float32 data is computed in double and rounded once per pass.
"""
import numpy as np

EPS = np.finfo(np.float64).eps


def extend_index(i, n, mode):
    """Index of the element that position i stands for; -1 = the constant."""
    if 0 <= i < n:
        return i
    if mode in ('reflect', 'grid-mirror'):
        if n <= 1:
            return 0
        p = 2 * n
        m = i % p
        return m if m < n else p - 1 - m
    if mode == 'mirror':
        if n <= 1:
            return 0
        p = 2 * n - 2
        m = i % p
        return m if m < n else p - m
    if mode in ('wrap', 'grid-wrap'):
        return i % n if n > 1 else 0
    if mode == 'nearest':
        return 0 if i < 0 else n - 1
    return -1


def _shifted(a, offsets, mode, cval):
    """a[ext(p + offsets)] for every p, as a float64 array."""
    idx = []
    const = np.zeros(a.shape, dtype=bool)
    for ax, (n, off) in enumerate(zip(a.shape, offsets)):
        q = np.array([extend_index(i + off, n, mode) for i in range(n)])
        shape = [1] * a.ndim
        shape[ax] = n
        const |= (q < 0).reshape(shape)
        idx.append(np.where(q < 0, 0, q).reshape(shape))
    out = a[tuple(np.broadcast_arrays(*idx))].astype(np.float64)
    out[np.broadcast_to(const, a.shape)] = cval
    return out


def correlate(a, weights, mode='reflect', cval=0.0, origin=0):
    a = np.asarray(a)
    weights = np.asarray(weights, dtype=np.float64)
    origins = [origin] * a.ndim if np.isscalar(origin) else list(origin)
    tmp = np.zeros(a.shape, dtype=np.float64)
    for k in np.ndindex(*weights.shape):
        w = weights[k]
        if not abs(w) > EPS:
            continue
        offs = [k[d] - weights.shape[d] // 2 - origins[d] for d in range(a.ndim)]
        tmp = tmp + _shifted(a, offs, mode, cval) * w
    return tmp.astype(a.dtype)


def convolve(a, weights, mode='reflect', cval=0.0, origin=0):
    """scipy `_correlate_or_convolve(convolution=True)`: flip the kernel, negate the origins (minus one more for
    even sizes), then correlate."""
    a = np.asarray(a)
    weights = np.asarray(weights, dtype=np.float64)
    origins = [origin] * a.ndim if np.isscalar(origin) else list(origin)
    weights = weights[tuple([slice(None, None, -1)] * weights.ndim)]
    origins = [-o - (0 if weights.shape[d] & 1 else 1) for d, o in enumerate(origins)]
    return correlate(a, weights, mode, cval, origins)


def correlate1d(a, weights, axis=-1, mode='reflect', cval=0.0, origin=0):
    a = np.asarray(a)
    w = np.asarray(weights, dtype=np.float64)
    axis = axis % a.ndim
    nw = len(w)
    s1, s2 = nw // 2, nw - nw // 2 - 1
    sym = 0
    if nw & 1:
        sym = 1
        if any(abs(w[i + s1] - w[s1 - i]) > EPS for i in range(1, s1 + 1)):
            sym = -1
            if any(abs(w[s1 + i] + w[s1 - i]) > EPS for i in range(1, s1 + 1)):
                sym = 0

    def x(j):
        offs = [0] * a.ndim
        offs[axis] = j - origin
        return _shifted(a, offs, mode, cval)

    if sym:
        tmp = x(0) * w[s1]
        for j in range(-s1, 0):
            tmp = tmp + ((x(j) + x(-j)) if sym > 0 else (x(j) - x(-j))) * w[s1 + j]
    else:
        tmp = x(s2) * w[s1 + s2]
        for j in range(-s1, s2):
            tmp = tmp + x(j) * w[s1 + j]
    return tmp.astype(a.dtype)


def gaussian_kernel1d(sigma, radius):
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum()


def gaussian_filter(a, sigma, mode='reflect', cval=0.0, truncate=4.0):
    """Order-0 Gaussian: one correlate1d per axis with sigma > 1e-15, every pass rounded to the array dtype."""
    a = np.asarray(a)
    sigmas = [sigma] * a.ndim if np.isscalar(sigma) else list(sigma)
    out = a
    for axis, sg in enumerate(sigmas):
        if sg > 1e-15:
            out = correlate1d(out, gaussian_kernel1d(float(sg), int(truncate * float(sg) + 0.5))[::-1], axis, mode, cval, 0)
    return out if out is not a else a.copy()
