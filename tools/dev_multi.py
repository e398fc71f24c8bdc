#!/usr/bin/env python
"""Development sweep (GPU), ONE process: for each forced tiled-kernel instantiation, a parity check against the
C oracle on a small cube and a CUDA-event timing of the kernel on a larger one.
Usage: dev_multi.py V0,V1,...  [--shape NY,NX,NT,V] [--r a,b,c] [--f F] [--pshape ...] [--steps K] [--neff N]
       variant 'auto' = the library's own choice."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nd_b200 import device
from oracle import c_port


def opts():
    o = {"--shape": "296,4096,32,4", "--r": "5,5,2", "--f": "1", "--pshape": "33,70,13,4", "--steps": "3", "--neff": "-1"}
    a = sys.argv[2:]
    for i in range(0, len(a), 2):
        o[a[i]] = a[i + 1]
    return o


def make_data(shape, seed=0):
    rng = np.random.default_rng(seed)
    a = rng.gamma(4.0, 0.25, size=shape)
    a[..., 1::2] = rng.normal(0, 0.3, size=a[..., 1::2].shape)
    a *= (1.0 + (np.arange(shape[1]) // 16 % 3)[None, :, None, None] * 0.5)
    return a.astype(np.float32)


def main():
    o = opts()
    shape = tuple(int(x) for x in o["--shape"].split(","))
    pshape = tuple(int(x) for x in o["--pshape"].split(","))
    r = tuple(int(x) for x in o["--r"].split(","))
    f = tuple(int(o["--f"]) if x > 0 else 0 for x in r)
    n_eff = float(o["--neff"])
    steps = int(o["--steps"])
    sigma, h = 0.3, 0.6
    a = make_data(pshape)
    ref = c_port.nlmeans(a, r, f, sigma, h, n_eff)
    t_small = torch.from_numpy(a).cuda()
    cube = device.synth_cube(*shape)
    peak = 148 * 128 * 2 * 1.965e9
    for v in sys.argv[1].split(","):
        if v == "auto":
            os.environ.pop("NDNLM_TILED_VARIANT", None)
        else:
            os.environ["NDNLM_TILED_VARIANT"] = v
        try:
            plan = device.Plan(pshape, r, f, sigma, h, n_eff, kernel="tiled")
            out = plan.apply(t_small).cpu().numpy()
            err = max(float(np.abs(out[..., k].astype(np.float64) - ref[..., k]).max() / np.abs(ref[..., k]).max())
                      for k in range(ref.shape[-1]))
            plan = device.Plan(cube.shape, r, f, 0.25, 0.5, n_eff, kernel="tiled")
            padded = plan.new_padded("cuda"); internal = plan.new_internal_out("cuda")
            flag = torch.zeros(1, dtype=torch.int32, device="cuda")
            plan.stage(cube, padded)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e30
            for it in range(steps + 1):
                e0.record(); plan.run(padded, internal, flag); e1.record(); torch.cuda.synchronize()
                if it > 0:
                    best = min(best, e0.elapsed_time(e1))
            vox = shape[0] * shape[1] * shape[2]
            print("V %-4s %s grid=%d smem=%d  err=%.2e  ms=%.3f  Mvox/s=%.1f  frac=%.4f" % (
                v, plan.kernel_name, plan.info.grid, plan.info.smem_bytes, err, best, vox / best / 1e3,
                plan.flops_per_voxel * vox / (best * 1e-3) / peak), flush=True)
            del padded, internal
        except Exception as e:
            print("V %-4s FAILED %r" % (v, e), flush=True)


if __name__ == "__main__":
    main()
