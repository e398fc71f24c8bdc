#!/usr/bin/env python
"""Development timing (GPU) of the omnibus change-detection kernel: Mpixel/s for a (rows, cols, k, 4) cube."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nd_b200 import _lib

def run(rows, cols, k, dtype, looks=50, alpha=0.9999, jumps=True):
    g = torch.Generator(device="cuda").manual_seed(0)
    z = torch.randn((rows, cols, k, looks, 2, 2), device="cuda", generator=g, dtype=torch.float32) / 2 ** 0.5
    if jumps:
        z[:, :, k // 2:] *= 1.5
    zc = torch.complex(z[..., 0], z[..., 1])
    c11 = (zc[..., 0].abs() ** 2).mean(-1); c22 = (zc[..., 1].abs() ** 2).mean(-1)
    c12 = (zc[..., 0] * zc[..., 1].conj()).mean(-1)
    v = torch.stack([c11, c12.real, c12.imag, c22], dim=-1).to(dtype).contiguous()
    del z, zc
    res = torch.empty((rows, cols, k), dtype=torch.uint8, device="cuda")
    L = _lib.lib()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    def call():
        rc = L.ndchg_change_detection(ctypes.c_void_p(v.data_ptr()), _lib.i64(v.shape[:3]), _lib.i64(v.stride()),
                                      0 if dtype == torch.float32 else 1, ctypes.c_void_p(res.data_ptr()), alpha, looks, st)
        assert rc == 0
    call(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(3):
        e0.record(); call(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    nbytes = v.numel() * v.element_size() + res.numel()
    print("%s %dx%dx%d looks=%d: %.3f ms  %.1f Mpixel/s  %.1f GB/s of algorithmic bytes  changes/pixel=%.3f" % (
        str(dtype)[6:], rows, cols, k, looks, best, rows * cols / best / 1e3, nbytes / best / 1e6, res.float().sum().item() / (rows * cols)), flush=True)

run(1024, 1024, 24, torch.float32)
run(1024, 1024, 24, torch.float64)
run(1024, 1024, 24, torch.float32, jumps=False)
