"""
Drop-in for `nd._filters` (reference nd/_filters.pyx): the same entry point, same argument
meaning and error behaviour, on host NumPy arrays -- but the arithmetic runs on the GPU through
libndnlm.so (include/ndnlm.h).  There is no CPU fallback: without a CUDA device or without the
built library this raises.
"""
import logging
import math

import numpy as np

from . import _lib


def find_weight(weight_sum, sq_weight_sum, n):
    """Closed form of reference nd/_filters.pyx:297-314 (host scalar helper; the kernels
    evaluate the same expression per voxel on the device)."""
    if n - 1 > weight_sum ** 2 / sq_weight_sum:
        raise ValueError('No solution')
    rt = math.sqrt(n * weight_sum * weight_sum - n * n * sq_weight_sum + n * sq_weight_sum)
    return (weight_sum + rt) / (n - 1)


def _as_u32_3(x, name):
    a = np.asarray(x)
    if isinstance(x, np.ndarray) and a.dtype != np.uint32:
        # the reference's typed memoryview `unsigned int[:]` rejects other dtypes (nd/_filters.pyx:322-323)
        raise ValueError("Buffer dtype mismatch, expected 'unsigned int' but got '%s' for %s" % (a.dtype, name))
    if a.ndim != 1 or a.shape[0] != 3:
        raise ValueError('%s must have exactly 3 entries (one per array axis)' % name)
    if np.any(np.asarray(a, dtype=np.int64) < 0):
        raise ValueError('%s must be non-negative' % name)
    return tuple(int(v) for v in a)


def _pixelwise_nlmeans_3d(arr, output, r, f, sigma, h, n_eff=-1, *, semantics=None, kernel='auto',
                          njobs=1, shard_axis=None, devices=None, pipeline=None, slab_rows=None):
    """
    GPU replacement of `nd._filters._pixelwise_nlmeans_3d` (reference nd/_filters.pyx:320-420).

    arr, output : (N0, N1, N2, V) float32 or float64 NumPy arrays (any strides); `output` is written
                  in place.  Other dtypes raise TypeError like the fused `floating` signature.
    r, f        : 3 search / patch radii (uint32 arrays as in the reference, or sequences).
    sigma, h, n_eff : as in the reference; n_eff < 0 means self weight = max weight.
    semantics   : 'as_written' (default) or 'reference_compiled' (SURVEY.md D1); also ND_NLM_SEMANTICS.
    njobs       : number of GPUs to shard over along `shard_axis` (default: the largest axis that is
                  not filtered, else the largest axis -- reference nd/filters.py:424-435).
    pipeline    : stream host arrays through the GPU in slabs with overlapped copies (default: arrays >= 256 MiB;
                  with njobs > 1: one slab pipeline per GPU whenever the shards run along axis 0 -- False keeps the
                  shards resident on the devices and exchanges their halo rows over NVLink peer copies).
    devices     : optional explicit CUDA device index per shard (e.g. [0, 0] exercises the shard /
                  halo-exchange layer on a single GPU).
    """
    import torch
    from . import device as dev

    arr = np.asarray(arr) if not isinstance(arr, np.ndarray) else arr
    if not isinstance(output, np.ndarray):
        raise TypeError('output must be a NumPy array (it is written in place)')
    if arr.dtype not in (np.float32, np.float64) or output.dtype != arr.dtype:
        raise TypeError('No matching signature found')          # fused-type dispatch failure in the reference
    if arr.ndim != 4 or output.ndim != 4:
        raise ValueError('Buffer has wrong number of dimensions (expected 4, got %d)' % arr.ndim)
    if output.shape != arr.shape:
        raise ValueError('output shape %s does not match input shape %s' % (output.shape, arr.shape))
    r3 = _as_u32_3(r, 'r')
    f3 = _as_u32_3(f, 'f')
    _lib.lib()                                                  # fail loudly if the library is missing
    if not torch.cuda.is_available():
        raise RuntimeError('nd_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    if arr.size == 0:
        return

    if any(s < 0 for s in arr.strides):
        arr = np.ascontiguousarray(arr)
    njobs = int(njobs)
    if njobs == -1:
        njobs = torch.cuda.device_count()
    if devices is not None:
        njobs = len(devices)
    if njobs > 1 or devices is not None:
        _sharded(arr, output, r3, f3, sigma, h, n_eff, semantics, kernel, njobs, shard_axis, devices, pipeline)
        return

    # large host arrays: slab pipeline (H2D / kernels / D2H overlapped), nd_b200/stream.py
    from . import stream as _stream
    if pipeline is None:
        pipeline = arr.nbytes >= (256 << 20)
    if pipeline and _stream.can_pipeline(arr, output) and (r3[0] + f3[0]) * 8 <= arr.shape[0]:
        _stream.apply_host_pipelined(arr, output, r3, f3, sigma, h, n_eff, semantics=semantics, kernel=kernel,
                                     slab_rows=slab_rows)
        return
    if pipeline:
        # not an error, but worth knowing when a large array takes the slow road
        logging.getLogger('nd_b200').info(
            'host array of shape %s (strides %s) is copied to the GPU in one piece: the slab pipeline needs a dense '
            'layout shared by input and output and at least %d rows along axis 0', arr.shape, arr.strides,
            max(128, 8 * (r3[0] + f3[0])))

    plan = dev.Plan(arr.shape, r3, f3, sigma, h, n_eff, semantics=semantics, dtype=arr.dtype, kernel=kernel)
    d_in = torch.from_numpy(arr).cuda()
    d_out = torch.empty_like(d_in)
    plan.apply(d_in, d_out)
    _copy_back(output, d_out)


def nlmeans_variables(arrays, outputs, r, f, sigma, h, n_eff=-1, *, semantics=None, kernel='auto', njobs=1,
                      devices=None):
    """The same filter on V SEPARATE host arrays (one (N0, N1, N2) C-ordered array per variable, as a Dataset holds
    them) written into V separate outputs -- what `Filter.apply` would otherwise gather into one
    (N0, N1, N2, V) block, filter, and scatter again (reference nd/filters.py:164-185: `to_array`, deep copy,
    `xr.merge`): three host-side passes over the data that cost more than the filter once it runs on a GPU.
    The arrays are streamed straight from / to the Dataset's own memory by the slab pipeline (nd_b200/stream.py),
    over `njobs` GPUs when asked.  Returns False (nothing done) when the arrays do not qualify for streaming."""
    import threading
    import torch
    from . import stream as _stream
    from .shard import ShardPlan

    r3 = _as_u32_3(r, 'r')
    f3 = _as_u32_3(f, 'f')
    arrays, outputs = list(arrays), list(outputs)
    if not arrays or any(not isinstance(a, np.ndarray) for a in arrays + outputs):
        return False
    a0 = arrays[0]
    halo = r3[0] + f3[0]
    ok = (a0.dtype in (np.float32, np.float64) and a0.ndim == 3 and a0.size > 0 and a0.shape[0] >= max(8 * halo, 2)
          and all(a.shape == a0.shape and a.dtype == a0.dtype and a.flags['C_CONTIGUOUS'] for a in arrays)
          and all(o.shape == a0.shape and o.dtype == a0.dtype and o.flags['C_CONTIGUOUS'] and o.flags['WRITEABLE']
                  for o in outputs) and len(outputs) == len(arrays))
    if not ok:
        return False
    _lib.lib()
    if not torch.cuda.is_available():
        raise RuntimeError('nd_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    njobs = int(njobs)
    if njobs == -1:
        njobs = torch.cuda.device_count()
    if devices is None:
        if njobs > torch.cuda.device_count():
            raise ValueError('njobs=%d but only %d CUDA devices are visible' % (njobs, torch.cuda.device_count()))
        devices = list(range(njobs)) if njobs > 1 else [torch.cuda.current_device()]
    sp = ShardPlan(a0.shape[0], len(devices), halo)
    if sp.nshards > 1 and min(hi - lo for lo, hi in sp.ranges) < 8 * max(halo, 1):
        sp = ShardPlan(a0.shape[0], 1, halo)
    errors = []

    def worker(i):
        try:
            torch.cuda.set_device(int(devices[i]))
            _stream.apply_host_pipelined(arrays, outputs, r3, f3, sigma, h, n_eff, semantics=semantics, kernel=kernel,
                                         row_range=sp.ranges[i])
        except BaseException as e:
            errors.append(e)

    if sp.nshards == 1:
        worker(0)
    else:
        threads = [threading.Thread(target=worker, args=(i,)) for i in range(sp.nshards)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    if errors:
        raise errors[0]
    return True


def _copy_back(output, d_out):
    import torch
    if all(s >= 0 for s in output.strides) and output.flags.writeable:
        torch.from_numpy(output).copy_(d_out)
    else:
        output[...] = d_out.cpu().numpy()


def choose_shard_axis(shape, r3, f3, njobs):
    """The reference's rule (`NLMeansFilter._parallel_dimension`, nd/filters.py:424-435): the largest axis that is
    not filtered, else the largest axis -- but an unfiltered axis too short to give every job a row (the usual 2-D
    case: a leading axis of extent 1) is no use, so only unfiltered axes with at least `njobs` rows qualify."""
    free = [a for a in range(3) if r3[a] == 0 and f3[a] == 0 and shape[a] >= njobs]
    cand = free if free else [a for a in range(3) if r3[a] > 0 or f3[a] > 0] or [0, 1, 2]
    return max(cand, key=lambda a: shape[a])


def _sharded(arr, output, r3, f3, sigma, h, n_eff, semantics, kernel, njobs, shard_axis, devices=None, pipeline=None):
    """Single-process multi-GPU apply over a HOST array (`njobs` = GPUs).

    Shards along axis 0 of a dense array are streamed: one slab pipeline per GPU (nd_b200/stream.py), each driven by
    its own thread, reading the `r+f` buffer rows of its neighbours straight from the host array -- the reference's
    own scheme (`xr_split` hands every worker its chunk plus a buffer, nd/utils.py:288-340), so no GPU-to-GPU
    exchange is needed.  Other shard axes / layouts keep their shards resident on the devices and exchange halo rows
    over NVLink peer copies."""
    import threading
    import warnings
    import torch
    from . import device as dev
    from . import stream as _stream
    from .shard import ShardPlan, exchange_halos_p2p

    ndev = torch.cuda.device_count()
    if devices is None:
        if njobs > ndev:
            raise ValueError('njobs=%d but only %d CUDA devices are visible' % (njobs, ndev))
        devices = list(range(njobs))
    devices = [int(d) for d in devices]
    if shard_axis is None:
        shard_axis = choose_shard_axis(arr.shape, r3, f3, njobs)
    halo = r3[shard_axis] + f3[shard_axis]
    sp = ShardPlan(arr.shape[shard_axis], njobs, halo)
    if sp.nshards < njobs:
        warnings.warn('axis %d (%d rows, halo %d) carries only %d of the %d requested shards'
                      % (shard_axis, arr.shape[shard_axis], halo, sp.nshards, njobs))
    devices = devices[:sp.nshards]

    if (pipeline is not False and shard_axis == 0 and _stream.can_pipeline(arr, output, min_rows=1)
            and min(hi - lo for lo, hi in sp.ranges) >= 8 * max(halo, 1)):
        errors = []

        def worker(i):
            try:
                torch.cuda.set_device(devices[i])
                _stream.apply_host_pipelined(arr, output, r3, f3, sigma, h, n_eff, semantics=semantics, kernel=kernel,
                                             row_range=sp.ranges[i])
            except BaseException as e:          # re-raised in the caller's thread
                errors.append(e)

        threads = [threading.Thread(target=worker, args=(i,)) for i in range(sp.nshards)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return

    # every shard uses the staged layout (role assignment, kernel) of the whole array: the halo messages of
    # neighbouring shards must mean the same rows
    whole = dev.Plan(arr.shape, r3, f3, sigma, h, n_eff, semantics=semantics, dtype=arr.dtype, kernel=kernel)
    plans, paddeds, outs, flags, slabs = [], [], [], [], []
    for i, (lo, hi) in enumerate(sp.ranges):
        idx = [slice(None)] * 4
        idx[shard_axis] = slice(lo, hi)
        slab = arr[tuple(idx)]
        with torch.cuda.device(devices[i]):
            plan = dev.Plan(slab.shape, r3, f3, sigma, h, n_eff, semantics=semantics, dtype=arr.dtype,
                            kernel=whole.kernel_request, roles=whole.roles)
            if plan.roles != whole.roles or plan.is_tiled != whole.is_tiled:
                raise RuntimeError('shard %d would use another staged layout than the whole array' % i)
            d_in = torch.from_numpy(slab).to('cuda:%d' % devices[i], non_blocking=True)
            padded = plan.new_padded(d_in.device)
            lo_e, hi_e = sp.edges(i)
            plan.stage(d_in, padded, shard_axis, lo_e, hi_e)
        plans.append(plan); paddeds.append(padded); slabs.append((idx, d_in))
    exchange_halos_p2p(plans, paddeds, shard_axis)
    for i, plan in enumerate(plans):
        with torch.cuda.device(devices[i]):
            internal = plan.new_internal_out(paddeds[i].device)
            flag = torch.zeros(1, dtype=torch.int32, device=paddeds[i].device)
            plan.run(paddeds[i], internal, flag)
            d_out = torch.empty_like(slabs[i][1])
            plan.unstage(internal, d_out)
        outs.append(d_out); flags.append(flag)
    bad = 0
    for i, d_out in enumerate(outs):
        torch.cuda.synchronize(devices[i])
        bad |= int(flags[i].item())
        output[tuple(slabs[i][0])] = d_out.cpu().numpy()
    _lib.check_flag(bad)
