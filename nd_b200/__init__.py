"""
nd_b200 -- a B200-native (sm_100a) replacement for ONE hot path of jnhansen/nd:
the non-local-means denoiser `nd.filters.NLMeansFilter` (reference nd/filters.py:388-466,
nd/_filters.pyx:317-420).  The Python API mirrors the reference; the arithmetic runs in
hand-written CUDA kernels behind a C ABI (include/ndnlm.h, nd_b200/libndnlm.so).
There is no CPU fallback.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401
