#!/usr/bin/env python
"""Development timing (GPU): stage / run / unstage of one plan with CUDA events.
Usage: dev_bench.py NY NX NT [V] [ry rx rt] [f] [--variant I] [--steps K]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nd_b200 import device

def main():
    args, opts, i = [], {}, 1
    while i < len(sys.argv):
        if sys.argv[i].startswith("--"):
            opts[sys.argv[i]] = sys.argv[i + 1]; i += 2
        else:
            args.append(sys.argv[i]); i += 1
    ny, nx, nt = (int(x) for x in args[:3])
    V = int(args[3]) if len(args) > 3 else 4
    r = tuple(int(x) for x in args[4:7]) if len(args) > 6 else (5, 5, 2)
    f = int(args[7]) if len(args) > 7 else 1
    steps = int(opts.get("--steps", 3))
    if "--variant" in opts:
        os.environ["NDNLM_TILED_VARIANT"] = opts["--variant"]
    fv = tuple(f if x > 0 else 0 for x in r)
    cube = device.synth_cube(ny, nx, nt, V)
    plan = device.Plan(cube.shape, r, fv, 0.25, 0.5, -1)
    padded = plan.new_padded("cuda"); internal = plan.new_internal_out("cuda")
    out = torch.empty_like(cube); flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    best = None
    for it in range(steps + 1):
        ev[0].record(); plan.stage(cube, padded)
        ev[1].record(); plan.run(padded, internal, flag)
        ev[2].record(); plan.unstage(internal, out)
        ev[3].record(); torch.cuda.synchronize()
        t = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
        if it > 0 and (best is None or t[1] < best[1]): best = t
    vox = ny * nx * nt
    mvox = vox / best[1] / 1e3
    peak = 148 * 128 * 2 * 1.965e9
    res = {"shape": [ny, nx, nt, V], "r": r, "f": fv, "kernel": plan.kernel_name, "desc": plan.describe(),
           "ms_stage": best[0], "ms_run": best[1], "ms_unstage": best[2], "Mvoxel_s_run": mvox,
           "roofline_frac_fp32": plan.flops_per_voxel * vox / (best[1] * 1e-3) / peak,
           "flag": int(flag.item()), "out_mean": float(out.mean().item()), "in_mean": float(cube.mean().item())}
    print("BENCH " + json.dumps(res), flush=True)

if __name__ == "__main__":
    main()
