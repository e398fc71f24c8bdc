#!/usr/bin/env python
"""Development timing (GPU): float64 2-D instantiations, forced by index.  Usage: dev_f64_2d.py IDX [IDX ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nd_b200 import device
for var in sys.argv[1:]:
    os.environ["NDNLM_TILED_VARIANT"] = var
    shape = (1, 4096, 4096, 4)
    cube = device.synth_cube(*shape).to(torch.float64)
    plan = device.Plan(shape, (0, 3, 3), (0, 1, 1), 0.25, 0.5, -1, dtype=np.float64, kernel="tiled64")
    padded = plan.new_padded("cuda"); internal = plan.new_internal_out("cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    plan.stage(cube, padded)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for it in range(4):
        e0.record(); plan.run(padded, internal, flag); e1.record(); torch.cuda.synchronize()
        if it: best = min(best, e0.elapsed_time(e1))
    print(var, plan.kernel_name, "%.3f ms %.1f Mvoxel/s" % (best, shape[1] * shape[2] / best / 1e3), flush=True)
