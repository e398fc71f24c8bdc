#!/usr/bin/env python
"""Development timing (GPU) of the sibling-filter kernels: achieved HBM GB/s per pass (algorithmic bytes = one read +
one write of the array per pass)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nd_b200 import _ndimage

def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

for dtype in (torch.float32, torch.float64):
    shape = (2048, 4096, 32) if dtype == torch.float32 else (1024, 4096, 32)
    a = torch.randn(shape, dtype=dtype, device="cuda"); b = torch.empty_like(a)
    nbytes = a.numel() * a.element_size()
    for name, fn, passes in [
        ("gaussian sigma=1 (y,x,t): 3 passes of 9 taps", lambda: _ndimage.gaussian_filter_device(a, b, 1.0), 3),
        ("gaussian sigma=2 (y,x): 2 passes of 17 taps", lambda: _ndimage.gaussian_filter_device(a, b, [2.0, 2.0, 0]), 2),
        ("boxcar 3x3 (y,x)", lambda: _ndimage.correlate_device(a, b, np.ones((3, 3, 1)) / 9, [0, 0, 0]), 1),
        ("boxcar 5x5 (y,x)", lambda: _ndimage.correlate_device(a, b, np.ones((5, 5, 1)) / 25, [0, 0, 0]), 1),
        ("boxcar 3x3x3", lambda: _ndimage.correlate_device(a, b, np.ones((3, 3, 3)) / 27, [0, 0, 0]), 1),
        ("copy (torch)", lambda: b.copy_(a), 1),
    ]:
        ms = timeit(fn)
        print("%-8s %-48s %8.3f ms  %7.1f GB/s per pass (frac of 6550: %.3f)" % (
            str(dtype)[6:], name, ms, passes * 2 * nbytes / ms / 1e6, passes * 2 * nbytes / ms / 1e6 / 6550.1), flush=True)
