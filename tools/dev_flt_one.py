#!/usr/bin/env python
"""One gaussian (2 plane passes, 17 taps) + one 3x3 boxcar on a float32 cube: the launch set profiled with ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nd_b200 import _ndimage
a = torch.randn((1024, 4096, 32), dtype=torch.float32, device="cuda"); b = torch.empty_like(a)
for _ in range(2):
    _ndimage.gaussian_filter_device(a, b, [2.0, 2.0, 0.0])
    _ndimage.correlate_device(a, b, np.ones((3, 3, 1)) / 9, [0, 0, 0])
torch.cuda.synchronize()
