#!/usr/bin/env python
"""
Generates tests/golden/change_golden.npz from the reference's OWN omnibus change detection (`oracle/_ref/_change`,
built unmodified from /root/reference/nd/_change.pyx by oracle/build_ref.py against the GSL stand-in oracle/gsl_shim).
Run in the build container (needs /root/reference); the .npz is committed so that the GPU box can check the CUDA
kernels against reference outputs.

Per case: the (rows, cols, k, 4) input, `prob` = single_pixel_omnibus of every pixel (nd/_change.pyx:139-160) and
`change` = change_detection(values, alpha, n) (:263-287).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import build_ref, ref_change  # noqa: E402


def wishart_series(rows, cols, k, looks, scales, seed=0, dtype=np.float64):
    rng = np.random.default_rng(seed)
    z = (rng.normal(size=(rows, cols, k, looks, 2)) + 1j * rng.normal(size=(rows, cols, k, looks, 2))) / np.sqrt(2)
    z = z * np.sqrt(np.asarray(scales, dtype=np.float64))[None, None, :, None, None]
    c11 = (np.abs(z[..., 0]) ** 2).mean(-1)
    c22 = (np.abs(z[..., 1]) ** 2).mean(-1)
    c12 = (z[..., 0] * np.conj(z[..., 1])).mean(-1)
    return np.stack([c11, c12.real, c12.imag, c22], axis=-1).astype(dtype)


def main():
    assert build_ref.build_change(), "oracle/_ref/_change is not built"
    rng = np.random.default_rng(11)
    out, meta = {}, {}
    cases = [("step_f64", np.float64, 10, 9, 0.99, None), ("step_f32", np.float32, 10, 9, 0.99, None),
             ("jumps_f64", np.float64, 16, 4, 0.999, "jumps"), ("jumps_f32", np.float32, 16, 4, 0.999, "jumps"),
             ("mild_f64", np.float64, 12, 16, 0.9, "mild"), ("mild_f32", np.float32, 12, 16, 0.9, "mild"),
             ("long_f64", np.float64, 30, 50, 0.5, "mild"), ("pair_f32", np.float32, 2, 4, 0.9, "mild")]
    for name, dtype, k, looks, alpha, kind in cases:
        if kind is None:
            scales = [1.0] * (k // 2) + [8.0] * (k - k // 2)
        elif kind == "jumps":
            scales = np.exp(np.cumsum(rng.choice([0.0, 0.0, 0.0, 1.2, -1.0], size=k)))
        else:
            scales = 1.0 + 0.3 * np.sin(np.arange(k))
        v = wishart_series(9, 8, k, looks, scales, seed=k + looks, dtype=dtype)
        out[name + "__in"] = v
        out[name + "__prob"] = ref_change.omnibus_probability(v, looks)
        out[name + "__change"] = ref_change.change_detection(v, alpha, looks)
        meta[name] = {"n": looks, "alpha": alpha, "changes": int(out[name + "__change"].sum())}
    out["__meta__"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "change_golden.npz"), **out)
    print(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
