#!/usr/bin/env python
"""
Generates tests/golden/nlm_golden.npz from the reference's OWN compiled kernel
(`oracle/_ref`, built from /root/reference/nd/_filters.pyx by oracle/build_ref.py).
Run in the build container (needs /root/reference); the .npz is committed so that the GPU box,
which has no /root/reference, can still check parity against reference outputs.

For every case it stores the input, the parameters and three reference outputs:
  out_compiled    unmodified kernel, direct call            (semantics 'reference_compiled')
  out_as_written  unmodified kernel via pad+augment+crop    (semantics 'as_written')
  out_patched     3-cast copy of the .pyx                   (cross-check of as_written; equal to 1 ulp)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import build_ref, ref  # noqa: E402


def sar_like(shape, seed, dtype):
    rng = np.random.default_rng(seed)
    a = rng.gamma(4.0, 0.25, size=shape)
    a[..., 1::2] = rng.normal(0, 0.3, size=a[..., 1::2].shape)
    a *= (1.0 + (np.arange(shape[1]) // 8 % 3)[None, :, None, None] * 0.5)
    return a.astype(dtype)


def slc_crop():
    """cfg1 input: a crop of the reference's data/slc.data/*.img (ENVI BSQ, big-endian float32, 206x500)."""
    base = "/root/reference/data/slc.data"
    planes = [np.fromfile(os.path.join(base, n + ".img"), dtype=">f4").reshape(206, 500)
              for n in ("C11", "C12_real", "C12_imag", "C22")]
    a = np.stack(planes, axis=-1).astype(np.float32)[80:112, 340:388]
    return np.ascontiguousarray(a[None])          # (1, 32, 48, 4): kernel axes (-, y, x)


CASES = [
    # name, array builder, r, f, sigma, h, n_eff
    ("3d_f1_f32", lambda: sar_like((9, 14, 7, 4), 1, np.float32), (2, 2, 1), (1, 1, 1), 0.3, 0.6, -1),
    ("3d_f1_f64", lambda: sar_like((8, 10, 6, 4), 2, np.float64), (1, 2, 1), (1, 1, 1), 0.3, 0.6, -1),
    ("3d_f1_neff", lambda: sar_like((8, 10, 6, 4), 3, np.float32), (2, 2, 1), (1, 1, 1), 0.3, 1.5, 6.0),
    ("2d_f1_slc", slc_crop, (0, 3, 3), (0, 1, 1), 0.001, 0.003, -1),
    ("2d_f2_f32", lambda: sar_like((1, 16, 20, 3), 4, np.float32), (0, 2, 3), (0, 2, 2), 0.3, 0.6, -1),
    ("3d_f0_f32", lambda: sar_like((6, 9, 8, 4), 5, np.float32), (1, 2, 2), (0, 0, 0), 0.3, 0.6, -1),
    ("3d_batch_t", lambda: sar_like((10, 12, 4, 4), 6, np.float32), (2, 2, 0), (1, 1, 0), 0.3, 0.6, -1),
    ("1d_f1_V1", lambda: sar_like((3, 4, 24, 1), 7, np.float32), (0, 0, 4), (0, 0, 1), 0.3, 0.6, -1),
    ("3d_f1_V6", lambda: sar_like((7, 9, 6, 6), 8, np.float32), (1, 1, 1), (1, 1, 1), 0.3, 0.6, -1),
]


def main():
    assert build_ref.build(), "oracle/_ref could not be built (needs /root/reference)"
    out = {}
    meta = {}
    for name, make, r, f, sigma, h, n_eff in CASES:
        a = make()
        out[name + "__in"] = a
        out[name + "__out_compiled"] = ref.reference_compiled(a, r, f, sigma, h, n_eff)
        out[name + "__out_as_written"] = ref.as_written(a, r, f, sigma, h, n_eff)
        out[name + "__out_patched"] = ref.as_written_patched(a, r, f, sigma, h, n_eff)
        meta[name] = {"r": r, "f": f, "sigma": sigma, "h": h, "n_eff": n_eff, "dtype": str(a.dtype), "shape": a.shape}
        d = np.abs(out[name + "__out_as_written"].astype(np.float64) - out[name + "__out_patched"]).max()
        print("%-12s shape=%s  |F3 - patched| = %.3g   |compiled - as_written| = %.3g" % (
            name, a.shape, d, np.abs(out[name + "__out_compiled"].astype(np.float64) - out[name + "__out_as_written"]).max()))
    out["__meta__"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(HERE, "nlm_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
