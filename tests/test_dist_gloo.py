"""CPU tests of the multi-GPU host logic: slab partition and the slab gather over a
world-size-2 ``gloo`` process group (the GPU run uses the same code over NCCL)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from martini_b200 import dist as mdist


def test_slab_bounds_even_and_aligned():
    assert mdist.slab_bounds(256, 1) == [0, 256]
    assert mdist.slab_bounds(256, 4) == [0, 64, 128, 192, 256]
    b = mdist.slab_bounds(250, 3)
    assert b[0] == 0 and b[-1] == 250 and all(x % 8 == 0 for x in b[1:-1]) and b == sorted(b)
    b = mdist.slab_bounds(5, 8)  # more ranks than tiles: some slabs are empty, none negative
    assert b[0] == 0 and b[-1] == 5 and b == sorted(b) and len(b) == 9


def test_slab_bounds_balance_work_not_area():
    rng = np.random.Generator(np.random.PCG64(3))
    nx = 512
    px = np.r_[rng.normal(120.0, 40.0, 90000), rng.uniform(0, nx, 10000)]  # mass piled up near row 120
    r = np.full(px.size, 4.0)
    work = mdist.row_work(px, r, nx)
    assert work.shape == (nx,) and np.isclose(work.sum(), np.sum(np.clip(np.floor(px + r) + 1, 0, nx)
                                                             - np.clip(np.ceil(px - r), 0, nx)))
    b = mdist.slab_bounds(nx, 4, work)
    per = [work[b[i]:b[i + 1]].sum() for i in range(4)]
    assert max(per) / (sum(per) / 4) < 1.35          # work-balanced ...
    assert (b[1] - b[0]) < nx // 4 and (b[4] - b[3]) > nx // 4  # ... by unequal areas


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, bounds, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ny, nc = 3, 4
        rows = bounds[rank + 1] - bounds[rank]
        # each rank's slab holds its global row index, so the assembled cube is checkable
        slab = (torch.arange(bounds[rank], bounds[rank + 1], dtype=torch.float64)[:, None, None]
                .expand(rows, ny, nc).contiguous())
        full = torch.full((bounds[-1], ny, nc), -1.0, dtype=torch.float64) if rank == 0 else None
        res = mdist.gather_slabs(slab, bounds, full, dst=0)
        everyone = torch.full((bounds[-1], ny, nc), -1.0, dtype=torch.float64)
        mdist.allgather_slabs(slab, bounds, everyone)
        expect = torch.arange(bounds[-1], dtype=torch.float64)[:, None, None].expand(bounds[-1], ny, nc)
        ok = bool(torch.equal(everyone, expect))
        if rank == 0:
            ok = ok and bool(torch.equal(res, expect))
        else:
            ok = ok and res is None
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and float(t) == world
        out_q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("bounds", ([0, 16, 40], [0, 40, 40], [0, 8, 24]))
def test_gather_slabs_world2_gloo(bounds):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, bounds, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results == {0: True, 1: True}


def test_single_process_gather_is_identity():
    slab = torch.ones(4, 2, 2, dtype=torch.float64)
    full = torch.zeros(4, 2, 2, dtype=torch.float64)
    assert torch.equal(mdist.gather_slabs(slab, [0, 4], full), slab)


def _host_cube_worker(rank, world, port, q):
    import torch.distributed as dist

    from martini_b200 import dist as mdist

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    shape, bounds = (12, 5, 4), [0, 8, 12]
    hc = mdist.HostCube(shape, bounds, pin=False)
    lo, hi = bounds[rank], bounds[rank + 1]
    slab = torch.arange(lo * 20, hi * 20, dtype=torch.float64).reshape(hi - lo, 5, 4)
    hc.store(slab)
    dist.barrier()
    if rank == 0:  # rank 0 sees the rows written by the other process
        q.put(bool(np.array_equal(hc.array, np.arange(12 * 20, dtype=np.float64).reshape(shape))))
    dist.barrier()
    hc.close()
    dist.destroy_process_group()


def test_host_cube_is_shared_between_ranks():
    """dist.HostCube: every rank writes its slab rows into one host array (world size 2, gloo)."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_host_cube_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok


def _sharded_worker(rank, world, port, nx, q):
    """dist.insert_sharded end to end on two ranks: the kernels run under the SIMT emulator
    (tests/emu, CPU tensors), the gather over gloo."""
    import torch.distributed as dist

    from martini_b200 import dist as mdist
    from martini_b200 import synthetic
    from martini_b200.pipeline import run_hot_path
    from tests.emu import EmuEngine

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        eng = EmuEngine()
        case = synthetic.make_case("cfg2", n=1500, nx=nx, ny=24, nc=32, seed=11)
        bounds = mdist.slab_bounds(nx, world)
        out, cube = mdist.insert_sharded(eng, case, bounds=bounds)
        if rank == 0:
            want = run_hot_path(eng, case)["cube"]
            ok = bool((cube - want).abs().max() <= 1e-13 * want.abs().max())
            q.put((ok, bounds))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nx", (40, 8))
def test_insert_sharded_world2_gloo(nx):
    """Two ranks, one cube: each projects its x-slab, rank 0 gathers.  nx = 8 leaves rank 0 or 1
    without rows (cuts stay on brick boundaries): it must still take part in the gather."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, nx, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, bounds = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok
    if nx == 8:
        assert 0 in [b - a for a, b in zip(bounds[:-1], bounds[1:])]


def _route_worker(rank, world, port, q):
    """dist.route_particles over gloo: every rank starts with a contiguous share of the
    particles and ends with exactly those whose box can reach its slab, in global index order;
    the slab cubes computed from the routed sets concatenate to the single-engine cube."""
    import torch.distributed as dist

    from martini_b200 import dist as mdist
    from martini_b200 import pipeline, synthetic
    from tests.emu import EmuEngine

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        eng = EmuEngine()
        case = synthetic.make_case("cfg3", n=3000, nx=48, ny=24, nc=32, seed=5)
        case["px"][7] = np.nan  # a NaN coordinate goes nowhere
        n = case["px"].size
        ctx = pipeline.prepare(case)
        bounds = [0, 16, 48]
        a, b = mdist.chunk_of(n, rank, world)
        share = dict(case, **{k: case[k][a:b] for k in pipeline.particle_keys(case)})
        dev = pipeline.upload(eng, share)
        _, _, sm_range, _ = eng.smoothing_setup(dev["sm_length"], ctx.table)
        routed = mdist.route_particles(dev, sm_range, bounds)
        # expected set: conservative superset test evaluated on the whole list
        _, _, r_all, _ = eng.smoothing_setup(eng.to_device(case["sm_length"]), ctx.table)
        r_all, px = r_all.numpy(), case["px"]
        must = (np.floor(px + r_all) >= bounds[rank]) & (np.ceil(px - r_all) <= bounds[rank + 1] - 1) & ~np.isnan(px)
        got_px = routed["px"].numpy()
        full_px = px[~np.isnan(px)]
        pos = np.searchsorted(np.sort(full_px), got_px)  # every routed particle is a real one
        ok = bool(np.all(np.isin(px[must], got_px))) and bool(np.all(np.isin(got_px, full_px)))
        # ascending global index: positions of the routed particles in the original list increase
        order = np.array([np.flatnonzero(px == x)[0] for x in got_px])
        ok = ok and bool(np.all(np.diff(order) > 0)) and pos.size == got_px.size
        x_lo, x_hi = bounds[rank], bounds[rank + 1]
        slab = torch.zeros((x_hi - x_lo, 24, 32), dtype=torch.float64)
        rcase = dict(case, **{k: routed[k].numpy() for k in pipeline.particle_keys(case)})
        pipeline.run_hot_path(eng, rcase, dev=routed, cube=slab, x_lo=x_lo, x_hi=x_hi, zeroed=True, ctx=ctx)
        full = torch.empty((48, 24, 32), dtype=torch.float64) if rank == 0 else None
        cube = mdist.gather_slabs(slab, bounds, full, dst=0)
        if rank == 0:
            want = pipeline.run_hot_path(eng, case)["cube"]
            ok = ok and bool((cube - want).abs().max() <= 1e-13 * want.abs().max())
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_route_particles_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_route_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert results == {0: True, 1: True}
