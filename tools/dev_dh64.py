#!/usr/bin/env python
"""Development timing (GPU): float64 data on the float64 tiled instantiations with and without double-duty halo
warps (NDNLM_DH is read at plan creation), cfg3 and cfg4 parameters; results are compared bitwise."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nd_b200 import device
for shape, r, f in (((210, 2048, 32, 4), (5, 5, 2), (1, 1, 1)), ((96, 1024, 64, 4), (7, 7, 2), (2, 2, 2))):
    cube = device.synth_cube(*shape).to(torch.float64)
    outs = []
    for dh in ("0", "1"):
        os.environ["NDNLM_DH"] = dh
        plan = device.Plan(shape, r, f, 0.25, 0.5, -1, dtype=np.float64, kernel="tiled64")
        padded = plan.new_padded("cuda"); internal = plan.new_internal_out("cuda")
        flag = torch.zeros(1, dtype=torch.int32, device="cuda")
        plan.stage(cube, padded)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for it in range(3):
            e0.record(); plan.run(padded, internal, flag); e1.record(); torch.cuda.synchronize()
            if it: best = min(best, e0.elapsed_time(e1))
        out = torch.empty_like(cube); plan.unstage(internal, out); outs.append(out)
        vox = shape[0] * shape[1] * shape[2]
        print("float64 DH=%s r=%s f=%s %-82s %9.3f ms  %8.1f Mvoxel/s" % (dh, r, f[0], plan.kernel_name, best, vox / best / 1e3), flush=True)
    print("   bitwise equal:", bool(torch.equal(outs[0], outs[1])), flush=True)
