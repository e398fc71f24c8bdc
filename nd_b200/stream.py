"""
Host <-> device slab pipeline for arrays that live in HOST memory.

`_pixelwise_nlmeans_3d(arr, output, ...)` on host arrays is PCIe-bound once the kernel runs at
~1 Gvoxel/s (32 B per voxel each way).  This driver cuts the cube into slabs along axis 0 -- the same
1-D split with an `r+f` buffer the reference uses for its worker pool (`xr_split` / `xr_merge`,
nd/utils.py:288-340; also the `buffer` of nd/tiling.py:18-106) -- and overlaps, on three CUDA streams
with double-buffered device slabs,

    H2D of slab i+1   |   stage + nlm kernel + unstage of slab i   |   D2H of slab i-1.

Every slab is copied with its buffer rows, but only its interior is filtered: the buffer rows serve as the
halo of the staged cube (NDNLM_EDGE_SOURCE), so no voxel is computed twice and the result is bitwise the
unsliced call (a voxel's arithmetic does not depend on where its slab starts).
"""
import itertools
import math
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import _lib
from . import device as dev
from .shard import ShardPlan


_POOL = None


def _pool():
    global _POOL
    if _POOL is None:
        _POOL = ThreadPoolExecutor(max_workers=max(2, min(16, os.cpu_count() or 2)), thread_name_prefix='ndb200-copy')
    return _POOL


def par_copy(dst, src):
    """dst.copy_(src) for two equally shaped CPU tensors, split over a few threads when it is large: one thread
    moves ~5 GB/s between pageable and pinned memory, far below what PCIe and the kernels sustain."""
    nbytes = src.numel() * src.element_size()
    if nbytes < (64 << 20) or src.dim() == 0 or src.shape[0] < 2:
        dst.copy_(src)
        return
    parts = int(min(_pool()._max_workers, max(1, nbytes // (32 << 20)), src.shape[0]))
    step = int(math.ceil(src.shape[0] / parts))
    futs = [_pool().submit(dst[a:a + step].copy_, src[a:a + step]) for a in range(0, src.shape[0], step)]
    for f in futs:
        f.result()


class _Shape:
    """shape / dtype / itemsize of the logical (N0, N1, N2, V) array when the variables are separate arrays."""

    def __init__(self, shape, dtype):
        self.shape, self.dtype, self.itemsize, self.ndim = shape, np.dtype(dtype), np.dtype(dtype).itemsize, 4


class _VarList:
    """V separate (N0, N1, N2) host arrays seen as the chunks [v] of a variable-major block."""

    def __init__(self, arrays):
        self.t = [torch.from_numpy(a) for a in arrays]

    def __getitem__(self, ix):
        return self.t[ix[0]]

    def is_pinned(self):
        return all(t.is_pinned() for t in self.t)


def dense_axis_order(a):
    """Axes of `a` from slowest to fastest if `a` is a dense, non-overlapping (possibly transposed)
    block of memory with positive strides; else None."""
    if any(s <= 0 for s in a.strides) or a.size == 0:
        return None
    order = sorted(range(a.ndim), key=lambda k: (-a.strides[k], k))
    expect = a.itemsize
    for k in reversed(order):
        if a.shape[k] != 1 and a.strides[k] != expect:
            return None
        expect *= a.shape[k]
    return order


def can_pipeline(arr, output, min_rows=64):
    return (arr.ndim == 4 and arr.shape == output.shape and arr.shape[0] >= 2 * min_rows
            and dense_axis_order(arr) is not None and dense_axis_order(arr) == dense_axis_order(output))


def apply_host_pipelined(arr, output, r3, f3, sigma, h, n_eff=-1, semantics=None, kernel='auto', slab_rows=None,
                         copy_only=False, row_range=None):
    """Filter host array `arr` into host array `output` through the slab pipeline.  Raises
    ValueError('No solution') like the reference when find_weight fails anywhere.

    `arr` / `output` may also be two LISTS of V dense C-ordered (N0, N1, N2) arrays, one per variable (what a Dataset
    holds): they are streamed as the chunks of a variable-major block without ever being gathered into one on the host.
    row_range=(lo, hi) restricts the work to rows [lo, hi) of axis 0 (the rows outside are still read as the
    buffer of the first / last slab): one GPU's share of a multi-GPU apply over a host array.
    copy_only=True runs the same slabs, streams and copies WITHOUT the kernels (the filtered rows are a device
    copy of the input): the host-memory / PCIe ceiling of this pipeline, reported by bench.py beside `e2e`."""
    per_variable = isinstance(arr, (list, tuple))
    if per_variable:
        if len(arr) != len(output) or not arr:
            raise ValueError('need one output array per input variable')
        for a, o in zip(arr, output):
            if (a.ndim != 3 or a.shape != arr[0].shape or o.shape != a.shape or a.dtype != arr[0].dtype or o.dtype != a.dtype
                    or not (a.flags['C_CONTIGUOUS'] and o.flags['C_CONTIGUOUS'])):
                raise ValueError('per-variable arrays must be C-contiguous (N0, N1, N2) arrays of one shape and dtype')
        order = [3, 0, 1, 2]
        hv_in, hv_out = _VarList(arr), _VarList(output)
        arr = _Shape(tuple(arr[0].shape) + (len(arr),), arr[0].dtype)
    else:
        order = dense_axis_order(arr)
        if order is None or order != dense_axis_order(output):
            raise ValueError('apply_host_pipelined needs dense arrays with identical memory layout')
        hv_in = torch.from_numpy(arr.transpose(order))
        hv_out = torch.from_numpy(output.transpose(order))
    n0 = arr.shape[0]
    halo = int(r3[0]) + int(f3[0])
    device = torch.device('cuda', torch.cuda.current_device())
    tdtype = torch.float32 if arr.dtype == np.float32 else torch.float64
    if slab_rows is None:
        # ~256 MiB per slab: small enough that filling / draining the pipeline (first H2D, last D2H) is a few
        # milliseconds, a whole number of kernel tiles along axis 0 so that no tile row is partly empty
        row_bytes = arr.itemsize * int(np.prod(arr.shape[1:]))
        slab_rows = max(4 * halo + 16, min(n0, (1 << 28) // max(row_bytes, 1)))
        probe = dev.Plan((min(n0, max(slab_rows, halo + 1)),) + tuple(arr.shape[1:]), r3, f3, sigma, h, n_eff,
                         semantics=semantics, dtype=arr.dtype, kernel=kernel)
        if probe.is_tiled:
            tile0 = int(probe.info.tile[list(probe.info.role_axis).index(0)])
            if tile0 > 0:
                slab_rows = max(tile0, slab_rows // tile0 * tile0)
    r_lo, r_hi = (0, n0) if row_range is None else (int(row_range[0]), int(row_range[1]))
    if not (0 <= r_lo < r_hi <= n0):
        raise ValueError('row_range must lie inside the array')
    sp = ShardPlan.from_rows(r_hi - r_lo, slab_rows, halo)
    sp.ranges = [(lo + r_lo, hi + r_lo) for lo, hi in sp.ranges]
    sp.n = n0            # buffered_range clips against the whole array, not the row range

    # host views in memory order: hv[outer..., rows, inner...] with every [outer][lo:hi] block contiguous
    k0 = order.index(0)
    inv = [order.index(a) for a in range(4)]
    max_buf = max(sp.buffered_range(i)[1] - sp.buffered_range(i)[0] for i in range(sp.nshards))
    max_int = max(hi - lo for lo, hi in sp.ranges)
    in_shape = [arr.shape[a] for a in order]
    out_shape = list(in_shape)
    in_shape[k0] = max_buf
    out_shape[k0] = max_int
    d_in = [torch.empty(in_shape, dtype=tdtype, device=device) for _ in range(2)]
    d_out = [torch.empty(out_shape, dtype=tdtype, device=device) for _ in range(2)]

    # One plan per distinct INTERIOR height.  A slab is staged from its buffered rows -- the `buffer` rows
    # of xr_split are read in place as the halo (NDNLM_EDGE_SOURCE) -- and only its interior is filtered, so
    # nothing is computed twice (the reference's workers filter the buffer rows too and xr_merge drops them).
    plans = {}
    pad_bytes = out_bytes = 0
    for lo, hi in sp.ranges:
        shape = (hi - lo,) + tuple(arr.shape[1:])
        if shape not in plans:
            plans[shape] = dev.Plan(shape, r3, f3, sigma, h, n_eff, semantics=semantics, dtype=arr.dtype, kernel=kernel)
            pad_bytes = max(pad_bytes, plans[shape].padded_bytes)
            out_bytes = max(out_bytes, plans[shape].out_bytes)
    padded = [torch.empty(pad_bytes, dtype=torch.uint8, device=device) for _ in range(2)]
    internal = [None, None]        # internal output buffers: only for output slabs the kernels cannot write in place
    flag = torch.zeros(1, dtype=torch.int32, device=device)

    cur = torch.cuda.current_stream()
    s_h2d, s_comp, s_d2h = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    for s in (s_h2d, s_comp, s_d2h):
        s.wait_stream(cur)
    ev_in_free = [None, None]     # compute on buffer b finished -> d_in[b] may be overwritten
    ev_out_free = [None, None]    # D2H from buffer b finished   -> d_out[b] may be overwritten

    def block(t, lo, hi):
        idx = [slice(None)] * 4
        idx[k0] = slice(lo, hi)
        return t[tuple(idx)]

    # Host <-> device copies are issued per CONTIGUOUS chunk ([outer index][rows] of the memory-order views: one
    # chunk for a C-ordered (N0, N1, N2, V) array, V chunks for the variable-major block `Filter.apply` hands
    # over), so every copy is a plain cudaMemcpyAsync.  Pinned host arrays are copied in place; pageable ones
    # (NumPy arrays of a Dataset) go through double-buffered pinned staging slabs filled / drained by this thread,
    # which keeps the DMA asynchronous and overlapped with the kernels.
    outer = list(itertools.product(*[range(n) for n in in_shape[:k0]])) or [()]
    pinned_in, pinned_out = hv_in.is_pinned(), hv_out.is_pinned()
    st_in = None if pinned_in else [torch.empty(in_shape, dtype=tdtype, pin_memory=True) for _ in range(2)]
    st_out = None if pinned_out else [torch.empty(out_shape, dtype=tdtype, pin_memory=True) for _ in range(2)]
    ev_st_in_free = [None, None]  # H2D out of staging slab b finished -> it may be refilled
    pending_out = [None, None]    # (event, lo, hi): staging slab b holds filtered rows still to be copied out

    def drain(b):
        if pending_out[b] is not None:
            ev, lo_, hi_ = pending_out[b]
            ev.synchronize()
            for ix in outer:
                par_copy(hv_out[ix][lo_:hi_], st_out[b][ix][:hi_ - lo_])
            pending_out[b] = None

    for i in range(sp.nshards):
        b = i & 1
        lo, hi = sp.ranges[i]
        blo, bhi = sp.buffered_range(i)
        rows = bhi - blo
        if not pinned_in:
            if ev_st_in_free[b] is not None:
                ev_st_in_free[b].synchronize()
            for ix in outer:
                par_copy(st_in[b][ix][:rows], hv_in[ix][blo:bhi])
        with torch.cuda.stream(s_h2d):
            if ev_in_free[b] is not None:
                s_h2d.wait_event(ev_in_free[b])
            for ix in outer:
                src = hv_in[ix][blo:bhi] if pinned_in else st_in[b][ix][:rows]
                d_in[b][ix][:rows].copy_(src, non_blocking=True)
            ev_h2d = torch.cuda.Event()
            ev_h2d.record(s_h2d)
            ev_st_in_free[b] = ev_h2d
        with torch.cuda.stream(s_comp):
            s_comp.wait_event(ev_h2d)
            if ev_out_free[b] is not None:
                s_comp.wait_event(ev_out_free[b])
            plan = plans[(hi - lo,) + tuple(arr.shape[1:])]
            a_in = block(d_in[b], lo - blo, hi - blo).permute(inv)     # logical (rows, N1, N2, V) view of the interior
            a_out = block(d_out[b], 0, hi - lo).permute(inv)
            if copy_only:
                a_out.copy_(a_in)
            else:
                plan.stage(a_in, padded[b], 0, 'reflect' if lo == 0 else 'source', 'reflect' if hi == n0 else 'source')
                if internal[b] is None and not plan.output_is_native(a_out):
                    internal[b] = torch.empty(out_bytes, dtype=torch.uint8, device=device)
                plan.run_into(padded[b], a_out, flag, internal[b])
            ev_comp = torch.cuda.Event()
            ev_comp.record(s_comp)
            ev_in_free[b] = ev_comp
        if not pinned_out:
            drain(b)                                                   # staging slab b is about to be overwritten
        with torch.cuda.stream(s_d2h):
            s_d2h.wait_event(ev_comp)
            for ix in outer:
                dst = hv_out[ix][lo:hi] if pinned_out else st_out[b][ix][:hi - lo]
                dst.copy_(d_out[b][ix][:hi - lo], non_blocking=True)
            ev_d2h = torch.cuda.Event()
            ev_d2h.record(s_d2h)
            ev_out_free[b] = ev_d2h
            if not pinned_out:
                pending_out[b] = (ev_d2h, lo, hi)
    if not pinned_out:
        drain(0)
        drain(1)
    for s in (s_h2d, s_comp, s_d2h):
        cur.wait_stream(s)
    cur.synchronize()
    _lib.check_flag(flag.item())
    return sp.nshards


def device_slab_rows(n0, inner_shape, r3, f3, sigma, h, n_eff=-1, semantics=None, kernel='auto', slab_rows=None,
                     dtype=np.float32):
    """Rows per slab `apply_device_streamed` uses for these arguments: the requested (or ~4 GiB) height rounded down
    to a whole number of kernel tiles along axis 0, so that no tile row of a slab is partly empty."""
    n0 = int(n0)
    inner_shape = tuple(int(x) for x in inner_shape)
    halo = int(r3[0]) + int(f3[0])
    if slab_rows is None:
        row_bytes = np.dtype(dtype).itemsize * int(np.prod(inner_shape))
        slab_rows = max(4 * halo + 16, min(n0, (4 << 30) // max(row_bytes, 1)))
    slab_rows = min(int(slab_rows), n0)
    probe = dev.Plan((max(slab_rows, halo + 1),) + inner_shape, r3, f3, sigma, h, n_eff, semantics=semantics,
                     dtype=dtype, kernel=kernel)
    if probe.is_tiled and slab_rows < n0:
        tile0 = int(probe.info.tile[list(probe.info.role_axis).index(0)])
        if tile0 > 0:
            slab_rows = max(tile0, slab_rows // tile0 * tile0)
    return slab_rows


def apply_device_streamed(source, sink, n0, inner_shape, r3, f3, sigma, h, n_eff=-1, semantics=None, kernel='auto',
                          slab_rows=None, lo_rows=None, hi_rows=None, dtype=np.float32, on_kernel=None):
    """Filter a cube of `n0` rows that is PRODUCED and CONSUMED slab by slab on the device -- cubes (or y-shards
    of cubes) whose input + staged copy + result do not fit in HBM at once (BASELINE configs[3] on 2 / 4 GPUs,
    configs[4] anywhere).  Same 1-D split with an `r+f` buffer as the host pipeline above (the reference's
    `xr_split` rule, nd/utils.py:288-340); every slab is staged from its buffered rows (NDNLM_EDGE_SOURCE), so
    nothing is computed twice and the result is bitwise the unsliced call.

    source(lo, hi, out)    fills the contiguous device tensor `out` (hi-lo, N1, N2, V) with rows [lo, hi) of the cube
    sink(lo, hi, result)   consumes the filtered rows [lo, hi) (device tensor, valid until the next slab)
    lo_rows / hi_rows      (r0+f0, N1, N2, V) device tensors: the rows just below row 0 / just above row n0-1 when the
                           cube is a y-shard of a larger one (exchanged with the neighbouring GPUs beforehand);
                           None = the cube really ends there and is reflected (reference `_idx`, nd/_filters.pyx:34-40)
    on_kernel(a, b)        optional hook called around every kernel launch (bench.py records CUDA events there)
    Returns the number of slabs.  Raises ValueError('No solution') like the reference."""
    n0 = int(n0)
    inner_shape = tuple(int(x) for x in inner_shape)
    halo = int(r3[0]) + int(f3[0])
    device = torch.device('cuda', torch.cuda.current_device())
    tdtype = torch.float32 if np.dtype(dtype) == np.float32 else torch.float64
    for nb in (lo_rows, hi_rows):
        if nb is not None and tuple(nb.shape) != (halo,) + inner_shape:
            raise ValueError('neighbour rows must have shape %s' % ((halo,) + inner_shape,))
    slab_rows = device_slab_rows(n0, inner_shape, r3, f3, sigma, h, n_eff, semantics=semantics, kernel=kernel,
                                 slab_rows=slab_rows, dtype=dtype)
    sp = ShardPlan.from_rows(n0, slab_rows, halo)
    max_int = max(hi - lo for lo, hi in sp.ranges)
    d_in = torch.empty((max_int + 2 * halo,) + inner_shape, dtype=tdtype, device=device)
    d_out = torch.empty((max_int,) + inner_shape, dtype=tdtype, device=device)
    plans = {}
    for lo, hi in sp.ranges:
        if hi - lo not in plans:
            plans[hi - lo] = dev.Plan((hi - lo,) + inner_shape, r3, f3, sigma, h, n_eff, semantics=semantics,
                                      dtype=dtype, kernel=kernel)
    padded = torch.empty(max(p.padded_bytes for p in plans.values()), dtype=torch.uint8, device=device)
    internal = None                # internal output buffer: only when the kernels cannot write the output slab in place
    flag = torch.zeros(1, dtype=torch.int32, device=device)
    for lo, hi in sp.ranges:
        # rows below: the neighbour shard's rows at the shard edge, rows of this cube otherwise (none at a true edge)
        below = halo if (lo > 0 or lo_rows is not None) else 0
        above = halo if (hi < n0 or hi_rows is not None) else 0
        own_lo, own_hi = max(lo - below, 0), min(hi + above, n0)
        off = below - (lo - own_lo)              # rows taken from lo_rows
        if off > 0:
            d_in[:off].copy_(lo_rows[halo - off:])
        source(own_lo, own_hi, d_in[off:off + own_hi - own_lo])
        top = above - (own_hi - hi)              # rows taken from hi_rows
        if top > 0:
            d_in[off + own_hi - own_lo:off + own_hi - own_lo + top].copy_(hi_rows[:top])
        plan = plans[hi - lo]
        a_in = d_in[below:below + hi - lo]
        a_out = d_out[:hi - lo]
        plan.stage(a_in, padded, 0, 'source' if below else 'reflect', 'source' if above else 'reflect')
        if on_kernel is not None:
            on_kernel(True, plan)
        direct = plan.output_is_native(a_out)          # C-ordered slabs of 4 variables: the kernels write them in place
        if not direct and internal is None:
            internal = torch.empty(max(p.out_bytes for p in plans.values()), dtype=torch.uint8, device=device)
        plan.run(padded, a_out if direct else internal, flag)
        if on_kernel is not None:
            on_kernel(False, plan)
        if not direct:
            plan.unstage(internal, a_out)
        sink(lo, hi, a_out)
    _lib.check_flag(flag.item())
    return sp.nshards
