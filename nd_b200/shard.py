"""
Sharding / halo layer -- the GPU analogue of the reference's only data-parallel strategy,
`xr_split` + `Pool.map` + `xr_merge` (nd/utils.py:288-401), and of the `buffer` overlap of
nd/tiling.py:18-106.

The cube is split into contiguous ranges along ONE axis (normally 'y'); every shard needs
`halo = r + f` rows of its neighbours (`NLMeansFilter._buffer`, nd/filters.py:437-445).
Instead of re-reading overlapping chunks from a pickled Dataset, each GPU stages only its own
rows (global edges are reflect-padded locally) and the halo rows travel GPU-to-GPU once per
apply: peer copies over NVLink in the single-process driver, NCCL send/recv in the
one-process-per-GPU driver.  No exchange is needed after the filter (outputs are disjoint).
"""
import math

import torch


class ShardPlan:
    """Contiguous 1-D decomposition; same chunking rule as `xr_split` (nd/utils.py:305-310)."""

    def __init__(self, n, chunks, halo):
        n, chunks, halo = int(n), int(chunks), int(halo)
        if chunks < 1:
            raise ValueError('chunks must be >= 1')
        self.n = n
        self.halo = halo
        # Every shard must hold at least halo + 1 rows: its neighbours take `halo` rows from it and a global edge
        # reflects once (ndnlm_plan_create: r + f <= N - 1).  Where the reference's `xr_split` would hand out
        # thinner chunks, use fewer shards (njobs larger than the cube can carry degrades, it does not fail).
        while True:
            chunksize = int(math.ceil(n / chunks))
            self.ranges = [(i * chunksize, min((i + 1) * chunksize, n)) for i in range(chunks) if i * chunksize < n]
            if len(self.ranges) <= 1 or min(hi - lo for lo, hi in self.ranges) >= halo + 1:
                break
            chunks = len(self.ranges) - 1
        self.requested = chunks

    @classmethod
    def from_rows(cls, n, rows, halo):
        """Slabs of `rows` rows each (the last one takes the remainder; a remainder too small to carry a halo
        is merged into the slab before it)."""
        n, rows, halo = int(n), max(int(rows), int(halo) + 1, 1), int(halo)
        self = cls(n, 1, halo)
        starts = list(range(0, n, rows))
        if len(starts) > 1 and n - starts[-1] < halo + 1:
            starts.pop()
        self.ranges = [(lo, starts[k + 1] if k + 1 < len(starts) else n) for k, lo in enumerate(starts)]
        return self

    @property
    def nshards(self):
        return len(self.ranges)

    def edges(self, i):
        """('reflect'|'halo', 'reflect'|'halo') for the low / high end of shard i."""
        return ('reflect' if i == 0 else 'halo', 'reflect' if i == self.nshards - 1 else 'halo')

    def buffered_range(self, i):
        """The range `xr_split` would hand to worker i (interior +- buffer, clipped)."""
        lo, hi = self.ranges[i]
        return max(lo - self.halo, 0), min(hi + self.halo, self.n)


def exchange_halos_p2p(plans, paddeds, axis):
    """Single-process multi-GPU halo exchange with peer copies (NVLink P2P).
    plans[i] / paddeds[i] live on GPU i; shard i+1 is the upper neighbour of shard i."""
    n = len(plans)
    pending = []
    for i in range(n - 1):
        lo_dev, hi_dev = paddeds[i].device, paddeds[i + 1].device
        nbytes = plans[i].halo_bytes(axis)
        if nbytes == 0:
            continue
        # last rows of shard i -> lower pad of shard i+1
        with torch.cuda.device(lo_dev):
            up = torch.empty(nbytes, dtype=torch.uint8, device=lo_dev)
            plans[i].halo_pack(paddeds[i], axis, 1, up)
        # first rows of shard i+1 -> upper pad of shard i
        with torch.cuda.device(hi_dev):
            down = torch.empty(nbytes, dtype=torch.uint8, device=hi_dev)
            plans[i + 1].halo_pack(paddeds[i + 1], axis, 0, down)
        pending.append((i, up, down))
    for dev in {p.device for p in paddeds}:
        torch.cuda.synchronize(dev)
    for i, up, down in pending:
        lo_dev, hi_dev = paddeds[i].device, paddeds[i + 1].device
        with torch.cuda.device(hi_dev):
            plans[i + 1].halo_unpack(paddeds[i + 1], axis, 0, up.to(hi_dev, non_blocking=True))
        with torch.cuda.device(lo_dev):
            plans[i].halo_unpack(paddeds[i], axis, 1, down.to(lo_dev, non_blocking=True))
    return 2 * len(pending)


def exchange_halos_dist(send_lo, send_hi, recv_lo, recv_hi, rank, world, group=None):
    """One-process-per-GPU halo exchange (torch.distributed: NCCL on GPUs, gloo in CPU tests).

    send_lo: my first rows (for rank-1), send_hi: my last rows (for rank+1);
    recv_lo: rows from rank-1 (my lower pad), recv_hi: rows from rank+1 (my upper pad).
    Edge ranks pass None for the missing side.  All messages are posted as one batch."""
    import torch.distributed as dist
    ops = []
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, send_lo, rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, recv_lo, rank - 1, group))
    if rank < world - 1:
        ops.append(dist.P2POp(dist.isend, send_hi, rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, recv_hi, rank + 1, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return len(ops)


class DistributedShard:
    """One rank's share of a y-sharded apply (one process per GPU).

    Usage per apply:  stage(slab) -> exchange() -> run([out_slab]) -> unstage(out_slab)."""

    def __init__(self, plan, axis, rank, world, group=None):
        self.plan, self.axis, self.rank, self.world, self.group = plan, axis, rank, world, group
        self.lo_edge = 'reflect' if rank == 0 else 'halo'
        self.hi_edge = 'reflect' if rank == world - 1 else 'halo'
        dev = torch.device('cuda', torch.cuda.current_device())
        self.padded = plan.new_padded(dev)
        self.internal = None           # allocated on first use: output slabs with the kernels' own layout never need it
        self._direct = False
        self.flag = torch.zeros(1, dtype=torch.int32, device=dev)
        nbytes = plan.halo_bytes(axis)
        mk = lambda: torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self.send_lo, self.recv_lo = (mk(), mk()) if rank > 0 else (None, None)
        self.send_hi, self.recv_hi = (mk(), mk()) if rank < world - 1 else (None, None)

    def stage(self, slab):
        self.plan.stage(slab, self.padded, self.axis, self.lo_edge, self.hi_edge)

    def exchange(self):
        if self.world == 1 or self.plan.halo_bytes(self.axis) == 0:
            return 0
        if self.send_lo is not None:
            self.plan.halo_pack(self.padded, self.axis, 0, self.send_lo)
        if self.send_hi is not None:
            self.plan.halo_pack(self.padded, self.axis, 1, self.send_hi)
        n = exchange_halos_dist(self.send_lo, self.send_hi, self.recv_lo, self.recv_hi, self.rank, self.world, self.group)
        if self.recv_lo is not None:
            self.plan.halo_unpack(self.padded, self.axis, 0, self.recv_lo)
        if self.recv_hi is not None:
            self.plan.halo_unpack(self.padded, self.axis, 1, self.recv_hi)
        return n

    def run(self, out_slab=None):
        """Run the kernels.  With `out_slab` given and laid out like the internal output (Plan.output_is_native) they
        write it directly and `unstage` becomes a no-op; otherwise the result waits in the internal buffer."""
        self._direct = out_slab is not None and self.plan.output_is_native(out_slab)
        if self._direct:
            self.plan.run(self.padded, out_slab, self.flag)
            return
        if self.internal is None:
            self.internal = self.plan.new_internal_out(self.padded.device)
        self.plan.run(self.padded, self.internal, self.flag)

    def unstage(self, out_slab):
        if not self._direct:
            self.plan.unstage(self.internal, out_slab)
