#!/usr/bin/env python
"""Development timing (GPU): float64 data on the float64 tiled kernel, the generic kernel, and the fp32 tiled kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nd_b200 import device
CFG3 = ((5, 5, 2), (1, 1, 1))
CFG4 = ((7, 7, 2), (2, 2, 2))
for dtype, kernel, shape, (r, f) in ((np.float64, "tiled64", (240, 2048, 32, 4), CFG3), (np.float64, "generic", (48, 1024, 32, 4), CFG3),
                                     (np.float32, "generic", (48, 1024, 32, 4), CFG3), (np.float64, "tiled", (240, 2048, 32, 4), CFG3),
                                     (np.float32, "tiled", (240, 2048, 32, 4), CFG3),
                                     (np.float64, "tiled64", (96, 1024, 64, 4), CFG4), (np.float64, "generic", (16, 256, 64, 4), CFG4)):
    cube = device.synth_cube(*shape).to(torch.float64 if dtype == np.float64 else torch.float32)
    plan = device.Plan(shape, r, f, 0.25, 0.5, -1, dtype=dtype, kernel=kernel)
    padded = plan.new_padded("cuda"); internal = plan.new_internal_out("cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    plan.stage(cube, padded)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for it in range(3):
        e0.record(); plan.run(padded, internal, flag); e1.record(); torch.cuda.synchronize()
        if it: best = min(best, e0.elapsed_time(e1))
    vox = shape[0] * shape[1] * shape[2]
    print("%-8s r=%s f=%s %-78s %9.3f ms  %8.1f Mvoxel/s" % (np.dtype(dtype).name, r, f[0], plan.kernel_name, best, vox / best / 1e3), flush=True)
