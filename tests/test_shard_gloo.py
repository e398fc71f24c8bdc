"""
World-size-2 test of the N>1 host path on CPU (gloo): each rank stages only its own y-rows
(reflect at the global edge, 'halo' on the interior edge), the halo rows travel with
`shard.exchange_halos_dist` exactly as on NCCL, and filtering the staged slab must reproduce the
unsharded result bit for bit -- the analogue of the reference's `test_parallelized_filter`
(nd/tests/test_filters_common.py:54-60).  The filter itself is played by the C oracle here (this is
a test of the shard / exchange logic, not of the CUDA kernels; the GPU version of this test lives in
tests/test_gpu_parity.py).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import np_stage, sar_like
from nd_b200.shard import ShardPlan, exchange_halos_dist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, shape, r, f, axis, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import c_port
        a = sar_like(shape, seed=21, dtype=np.float32)          # every rank derives the same global cube
        pads = [r[i] + f[i] for i in range(3)]
        sp = ShardPlan(shape[axis], world, pads[axis])
        lo, hi = sp.ranges[rank]
        idx = [slice(None)] * 4
        idx[axis] = slice(lo, hi)
        slab = a[tuple(idx)]
        lo_e, hi_e = sp.edges(rank)
        staged = np_stage(slab, pads, axis, lo_e, hi_e)          # what ndnlm_stage leaves on this rank
        p = pads[axis]
        n = hi - lo

        def rows(start, stop):
            ix = [slice(None)] * 4
            ix[axis] = slice(start, stop)
            return tuple(ix)

        # ndnlm_halo_pack: my first / last `p` INTERIOR rows (padded coordinates [p, 2p) and [n, n+p))
        send_lo = torch.from_numpy(np.ascontiguousarray(staged[rows(p, 2 * p)])) if rank > 0 else None
        send_hi = torch.from_numpy(np.ascontiguousarray(staged[rows(n, n + p)])) if rank < world - 1 else None
        recv_lo = torch.empty_like(send_lo) if rank > 0 else None
        recv_hi = torch.empty_like(send_hi) if rank < world - 1 else None
        nmsg = exchange_halos_dist(send_lo, send_hi, recv_lo, recv_hi, rank, world)
        # ndnlm_halo_unpack: lower pad rows [0, p), upper pad rows [p+n, 2p+n)
        if recv_lo is not None:
            staged[rows(0, p)] = recv_lo.numpy()
        if recv_hi is not None:
            staged[rows(p + n, 2 * p + n)] = recv_hi.numpy()
        assert not np.isnan(staged).any()
        # filtering the staged slab and cropping its interior == filtering my rows of the global cube
        out = c_port.nlmeans(staged, r, f, 0.3, 0.6)
        crop = tuple(slice(pads[i], pads[i] + slab.shape[i]) for i in range(3)) + (slice(None),)
        q.put((rank, lo, hi, nmsg, out[crop]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape,r,f,axis", [((14, 10, 5, 3), (2, 1, 1), (1, 1, 1), 0),
                                            ((6, 16, 4, 2), (1, 2, 0), (1, 1, 0), 1)])
def test_sharded_equals_unsharded_gloo(shape, r, f, axis):
    from oracle import c_port
    c_port.build()
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(rk, world, port, shape, r, f, axis, q)) for rk in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = c_port.nlmeans(sar_like(shape, seed=21, dtype=np.float32), r, f, 0.3, 0.6)
    for rank, lo, hi, nmsg, out in results:
        idx = [slice(None)] * 4
        idx[axis] = slice(lo, hi)
        assert nmsg == 2                                    # one send + one recv per interior edge
        assert np.array_equal(out, full[tuple(idx)]), "rank %d differs from the unsharded result" % rank
