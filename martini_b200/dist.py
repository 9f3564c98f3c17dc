"""Multi-GPU decomposition of the projection: one process per GPU, the cube split into
contiguous x-slabs.

The path shards naturally.  Voxels are independent outputs and a particle only reaches the
pixels of its candidate box, so rank r computes rows ``[x_lo, x_hi)`` of the cube from the
particles whose box touches those rows (every rank filters the full particle list on the
device with ``mtn_plan``; particles straddling a boundary are simply processed by both
neighbours -- halo replication -- and each rank writes only its own rows).  There is no
reduction and no atomics, so the assembled cube is the single-GPU cube.  x is the slowest
array axis of the C-ordered ``(nx, ny, C)`` cube, hence a slab is one contiguous byte range
and assembly is a pure concatenation: the one exchange step is a gather of slabs (NCCL over
NVLink; gloo in the CPU tests).
"""

from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

TILE = 8  # slab boundaries are kept on brick boundaries (csrc/common.cuh: MTN_TILE)


def row_work(px, sm_range, nx, ny_weight=None):
    """Per-row work estimate: how many particle boxes cover each cube row (optionally
    weighted, e.g. by box height x live channels).  ``px``/``sm_range`` are host arrays."""
    lo = np.clip(np.ceil(px - sm_range), 0, nx).astype(np.int64)
    hi = np.clip(np.floor(px + sm_range) + 1, 0, nx).astype(np.int64)
    ok = hi > lo
    w = np.ones(px.shape) if ny_weight is None else np.asarray(ny_weight, dtype=np.float64)
    diff = np.zeros(nx + 1)
    np.add.at(diff, lo[ok], w[ok])
    np.add.at(diff, hi[ok], -w[ok])
    return np.cumsum(diff)[:nx]


def slab_bounds(nx, world, work=None, align=TILE):
    """``world + 1`` row boundaries 0 = b_0 <= ... <= b_world = nx.

    Without ``work`` the rows are split evenly; with a per-row work estimate the boundaries
    equalise the work prefix sum instead of the area.  Interior boundaries are multiples of
    ``align`` so every slab starts on a brick boundary."""
    if world < 1 or nx < 1:
        raise ValueError("need world >= 1 and nx >= 1")
    if work is None:
        work = np.ones(nx)
    work = np.asarray(work, dtype=np.float64) + 1e-12 * (np.sum(work) / nx + 1.0)  # strictly increasing prefix
    prefix = np.concatenate(([0.0], np.cumsum(work)))
    targets = prefix[-1] * np.arange(1, world) / world
    cuts = np.searchsorted(prefix, targets)
    cuts = (np.round(cuts / align) * align).astype(int)
    b = np.concatenate(([0], np.clip(cuts, 0, nx), [nx]))
    return np.maximum.accumulate(b).tolist()


def gather_slabs(slab: torch.Tensor, bounds, full: torch.Tensor | None = None, dst: int = 0):
    """Assemble the cube on rank ``dst`` from every rank's slab (rows bounds[r]:bounds[r+1]).

    ``full`` (only needed on ``dst``) is the preallocated (nx, ny, C) result.  Slabs may have
    different heights (work-balanced partition): rank ``dst`` posts one receive per peer
    directly into the destination rows of ``full`` -- slabs are contiguous byte ranges, so no
    staging copy is needed -- and copies its own slab locally."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        if full is not None and full.data_ptr() != slab.data_ptr():
            full.copy_(slab)
        return slab if full is None else full
    rank, world = dist.get_rank(), dist.get_world_size()
    if rank == dst:
        assert full is not None and full.shape[0] == bounds[-1]
        reqs = []
        for r in range(world):
            rows = full[bounds[r]:bounds[r + 1]]
            if r == dst:
                rows.copy_(slab)
            elif rows.numel():
                reqs.append(dist.irecv(rows, src=r))
        for q in reqs:
            q.wait()
        return full
    if slab.numel():
        dist.send(slab, dst=dst)
    return None


def allgather_slabs(slab: torch.Tensor, bounds, full: torch.Tensor):
    """Every rank ends up with the whole cube (for callers that continue on all GPUs)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        full.copy_(slab)
        return full
    world = dist.get_world_size()
    for r in range(world):
        rows = full[bounds[r]:bounds[r + 1]]
        if r == dist.get_rank():
            rows.copy_(slab)
        if rows.numel():
            dist.broadcast(rows, src=r)
    return full


class PeerCube:
    """Fused assembly: the projection kernel's voxel stores go straight into the destination
    rank's cube over NVLink, so no gather pass follows the compute.

    The cube is allocated in symmetric memory (``torch.distributed._symmetric_memory``: every
    rank allocates the buffer, the CUDA driver maps all of them into every process); rank r
    gets ``rows`` = a tensor aliasing rows [bounds[r], bounds[r+1]) of rank ``dst``'s buffer and
    hands it to ``mtn_project`` as its slab.  The C ABI needs nothing special: a slab is just a
    device pointer, and every voxel is stored once, as 16-byte vector stores (512 contiguous
    bytes per warp) -- the access pattern NVLink peer stores like.  Use with MTN_CUBE_ZEROED
    only (accumulate mode would read the cube back over the link).

    Per insertion: ``begin()`` (the destination rank zeroes its cube locally -- 0.1 ms for
    537 MB at HBM rate; every rank zeroing its own rows through the peer mapping instead puts
    the whole cube through the destination's NVLink ingress, 0.6 ms at 8 ranks -- then a
    device-side barrier on the allocation's signal pads so no store can overtake the memset),
    the ranks project, ``end()`` (barrier: all stores have landed)."""

    def __init__(self, shape, bounds, device, dst=0, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group or dist.group.WORLD
        self.rank, self.dst = dist.get_rank(), dst
        nx, ny, nc = shape
        self.buf = symm_mem.empty((nx, ny, nc), dtype=torch.float64, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        lo, hi = bounds[self.rank], bounds[self.rank + 1]
        self.rows = self.hdl.get_buffer(dst, (hi - lo, ny, nc), torch.float64, storage_offset=lo * ny * nc)

    def _barrier(self):
        try:  # device-side barrier on the signal pads of the symmetric allocation (stream-ordered)
            self.hdl.barrier(channel=0)
        except Exception:  # noqa: BLE001 -- older handle API
            torch.cuda.current_stream().synchronize()
            dist.barrier(group=self.group)

    def begin(self):
        if self.rank == self.dst:
            self.buf.zero_()
        self._barrier()

    def end(self):
        self._barrier()
        return self.buf if self.rank == self.dst else None


class HostCube:
    """The assembled cube in host memory, for callers that want the result on the host: every
    rank copies its slab device -> host over its own PCIe link into its rows of one array
    shared by the ranks of the node (POSIX shared memory, page-locked with cudaHostRegister so
    the copies are asynchronous DMA).  No rank ever holds the whole cube on a device and no
    single link carries all of it (measured on this pool's 8-GPU boxes: 29 ms instead of 46 ms
    for the 537 MB weak-scaled cube; the boxes' host side saturates at about 43 GB/s aggregate).
    ``array`` is the numpy view."""

    def __init__(self, shape, bounds, pin=True):
        from multiprocessing import resource_tracker, shared_memory

        self.rank = dist.get_rank() if dist.is_initialized() else 0
        world = dist.get_world_size() if dist.is_initialized() else 1
        nbytes = int(np.prod(shape)) * 8
        name = [None]
        if self.rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=nbytes)
            name[0] = self.shm.name
        if world > 1:
            dist.broadcast_object_list(name, src=0)
        if self.rank != 0:
            self.shm = shared_memory.SharedMemory(name=name[0])
            # the creator unlinks it; keep this process's resource tracker out of it
            resource_tracker.unregister(self.shm._name, "shared_memory")
        self.array = np.ndarray(shape, dtype=np.float64, buffer=self.shm.buf)
        self.tensor = torch.from_numpy(self.array)
        self.pinned = False
        if pin and torch.cuda.is_available():
            self.pinned = torch.cuda.cudart().cudaHostRegister(self.tensor.data_ptr(), nbytes, 0) in (0, None) or \
                self.tensor.is_pinned()
        self.rows = self.tensor[bounds[self.rank]:bounds[self.rank + 1]]

    def store(self, slab: torch.Tensor):
        """Enqueue the copy of this rank's slab into its rows (asynchronous if page-locked)."""
        self.rows.copy_(slab, non_blocking=True)

    def close(self):
        if self.pinned:
            torch.cuda.cudart().cudaHostUnregister(self.tensor.data_ptr())
            self.pinned = False
        del self.rows, self.tensor, self.array
        self.shm.close()
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.barrier()
        if self.rank == 0:
            self.shm.unlink()


def insert_sharded(engine, case, dev=None, ctx=None, bounds=None, gather=True, full=None):
    """Run the hot path for this rank's slab and (optionally) gather the cube on rank 0.

    Returns (result dict of run_hot_path, assembled cube or None)."""
    from . import pipeline

    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    ctx = ctx or pipeline.prepare(case)
    nx, ny, nc = ctx.shape
    bounds = bounds or slab_bounds(nx, world)
    x_lo, x_hi = bounds[rank], bounds[rank + 1]
    if dev is None:
        dev = pipeline.upload(engine, case)
    slab = torch.zeros((x_hi - x_lo, ny, nc), dtype=torch.float64, device=engine.device)
    # slab_bounds keeps the cuts on brick boundaries, so a narrow cube can leave a rank without
    # rows: it projects nothing but still takes part in the gather
    out = None
    if x_hi > x_lo:
        out = pipeline.run_hot_path(engine, case, dev=dev, cube=slab, x_lo=x_lo, x_hi=x_hi, zeroed=True, ctx=ctx)
    cube = None
    if gather:
        if rank == 0 and full is None:
            full = torch.empty((nx, ny, nc), dtype=torch.float64, device=engine.device)
        cube = gather_slabs(slab, bounds, full, dst=0)
    return out, cube


# --------------------------------------------------------------------------------------------
# Particle routing: the one exchange step of the input side.
#
# Every rank starts with a contiguous 1 / world share of the particle list (uploaded from its
# own host buffers over its own PCIe link).  A particle is needed by every rank whose slab its
# candidate box [px - r, px + r] reaches, so each rank buckets its share by destination slab
# (halo particles go to both neighbours) and the buckets are exchanged with one all-to-all over
# NVLink.  Received buckets are concatenated in source-rank order and every bucket keeps its
# particles in index order, so a rank sees its particles in ascending global index -- the
# summation order of the single-GPU run.  The destination test is conservative (one pixel of
# slack on both sides); the exact candidate-box predicate is applied afterwards by mtn_plan on
# the receiving rank, exactly as in the single-GPU path.
# --------------------------------------------------------------------------------------------
ROUTED_KEYS = ("px", "py", "pz", "sm_length", "v", "mHI", "D", "sigma")


def route_particles(dev, sm_range, bounds, group=None):
    """``dev``: dict of this rank's per-particle device tensors (any subset of ROUTED_KEYS that
    are tensors; scalars are passed through), ``sm_range`` their candidate-box half widths in
    pixels (K0's output).  Returns the dict of the particles whose box may reach this rank's
    slab ``[bounds[rank], bounds[rank + 1])``."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    keys = [k for k in ROUTED_KEYS if isinstance(dev.get(k), torch.Tensor)]
    out = {k: v for k, v in dev.items() if k not in keys}
    if world == 1:
        out.update({k: dev[k] for k in keys})
        return out
    px = dev["px"]
    device = px.device
    b = torch.as_tensor(bounds, dtype=torch.float64, device=device)
    r = torch.nan_to_num(sm_range, nan=0.0, posinf=float(bounds[-1]) + 2.0)
    lo = torch.floor(px - r) - 1.0   # first / last cube row the box may reach (one pixel of slack)
    hi = torch.ceil(px + r) + 1.0
    ok = ~torch.isnan(px)
    # rank d is a destination iff its slab [b[d], b[d+1]) intersects [lo, hi] and is not empty
    x_lo, x_hi = b[:-1].unsqueeze(1), b[1:].unsqueeze(1)
    want = (lo.unsqueeze(0) < x_hi) & (hi.unsqueeze(0) >= x_lo) & (x_hi > x_lo) & ok.unsqueeze(0)  # (world, n)
    dest, idx = torch.nonzero(want, as_tuple=True)  # sorted by destination, then particle index
    send_counts = torch.bincount(dest, minlength=world)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    sc, rc = send_counts.tolist(), recv_counts.tolist()
    n_recv = sum(rc)
    # one gather + one all-to-all per quantity: the receive side is then already the contiguous
    # array the C ABI wants (no packing into records, no transpose afterwards)
    for k in keys:
        send = dev[k].index_select(0, idx)
        recv = torch.empty(n_recv, dtype=send.dtype, device=device)
        dist.all_to_all_single(recv, send, output_split_sizes=rc, input_split_sizes=sc, group=group)
        out[k] = recv
    return out


def chunk_of(n, rank, world):
    """The contiguous share [a, b) of an n-particle list that rank ``rank`` uploads."""
    return (n * rank) // world, (n * (rank + 1)) // world


class PeerRouter:
    """The input exchange fused with its bucketing: every rank's routed particles are stored
    straight into the destination ranks' inboxes through peer-mapped pointers (symmetric
    memory over NVLink) by ``mtn_route_scatter`` -- no send buffers, no all-to-all.  Host
    involvement per step: one 8-byte read-back (how many particles arrived) and two barriers.

    ``capacity``: particles an inbox can hold (per quantity); ``route`` raises if a rank would
    receive more."""

    def __init__(self, n_fields, capacity, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group or dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.capacity, self.n_fields = int(capacity), int(n_fields)
        self.inbox = symm_mem.empty((n_fields, self.capacity), dtype=torch.float64, device=device)
        self.hdl = symm_mem.rendezvous(self.inbox, self.group)
        self._peers = [self.hdl.get_buffer(r, (n_fields, self.capacity), torch.float64) for r in range(self.world)]
        self.ptrs = [t.data_ptr() for t in self._peers]

    def _barrier(self):
        """All ranks' work enqueued so far is complete.  Device-side (symmetric-memory signal
        pads, stream-ordered) where the handle offers it; NCCL barrier after a stream sync
        otherwise."""
        try:
            self.hdl.barrier(channel=0)
        except Exception:  # noqa: BLE001 -- older handle API
            torch.cuda.current_stream().synchronize()
            dist.barrier(group=self.group)

    def route(self, engine, dev, sm_range, bounds, inbox_free=False):
        """``inbox_free``: the caller guarantees that every rank has finished reading its inbox
        of the previous step (e.g. a barrier closed that step): saves one barrier."""
        keys = [k for k in ROUTED_KEYS if isinstance(dev.get(k), torch.Tensor)]
        assert len(keys) <= self.n_fields
        out = {k: v for k, v in dev.items() if k not in keys}
        totals, scratch = engine.route_count(dev["px"], sm_range, bounds)
        counts = torch.empty((self.world, self.world), dtype=torch.int64, device=totals.device)  # [src][dst]
        dist.all_gather_into_tensor(counts, totals, group=self.group)
        src_offsets = counts[:self.rank].sum(dim=0)
        arriving = counts.sum(dim=0).tolist()        # the one read-back (also orders the host after the gather)
        n_recv = arriving[self.rank]
        if max(arriving) > self.capacity:            # (same verdict on every rank)
            raise RuntimeError(f"PeerRouter: {max(arriving)} particles for one rank exceed the inbox "
                               f"capacity {self.capacity}")
        if not inbox_free:
            self._barrier()                          # every rank is done reading its inbox of the last step
        engine.route_scatter(dev["px"], sm_range, bounds, [dev[k] for k in keys], self.ptrs, self.capacity,
                             src_offsets, scratch)
        self._barrier()                              # all stores have landed
        out.update({k: self.inbox[i, :n_recv] for i, k in enumerate(keys)})
        return out
