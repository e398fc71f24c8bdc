"""Shared test helpers (CPU side)."""
import numpy as np


def sar_like(shape, seed=0, dtype=np.float32):
    rng = np.random.default_rng(seed)
    a = rng.gamma(4.0, 0.25, size=shape)
    a[..., 1::2] = rng.normal(0, 0.3, size=a[..., 1::2].shape)
    a *= (1.0 + (np.arange(shape[1]) // 8 % 3)[None, :, None, None] * 0.5)
    return a.astype(dtype)


def scaled_err(out, ref):
    """Per-variable max |out-ref| / max |ref_v|  (SURVEY.md 8(d) parity metric); returns the worst variable."""
    worst = 0.0
    for v in range(ref.shape[-1]):
        d = np.abs(out[..., v].astype(np.float64) - ref[..., v].astype(np.float64))
        worst = max(worst, float(np.nanmax(d) / max(np.nanmax(np.abs(ref[..., v])), 1e-300)))
    return worst


def np_stage(slab, pads, axis=None, lo_edge="reflect", hi_edge="reflect"):
    """NumPy statement of what `ndnlm_stage` produces (in USER axis order): reflect-pad every axis by
    pads[a]; on `axis`, 'halo' edges are left as NaN for the neighbour exchange to fill."""
    out = np.pad(slab, [(p, p) for p in pads] + [(0, 0)], mode="reflect")
    if axis is not None and pads[axis] > 0:
        idx = [slice(None)] * 4
        if lo_edge == "halo":
            idx[axis] = slice(0, pads[axis])
            out[tuple(idx)] = np.nan
        if hi_edge == "halo":
            idx[axis] = slice(out.shape[axis] - pads[axis], None)
            out[tuple(idx)] = np.nan
    return out
