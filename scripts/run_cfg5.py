#!/usr/bin/env python
"""BASELINE config 5: one large source sharded over the GPUs of a box.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 \
        scripts/run_cfg5.py [--particles 100000000 --nx 2048 --nc 512 --out profiles/r1_cfg5_8gpu.json]

Every rank generates the same particle set on its device (torch generator, fixed seed), the
cube is split into work-balanced x-slabs (martini_b200.dist.slab_bounds on a device-built
row-work histogram), each rank runs K0 -> K1 -> plan -> project on its slab (halo particles
processed by both neighbours) and the slabs are gathered on rank 0 over NCCL.  Rank 0 then
checks seeded pixel columns against the CPU oracle (reference-structured loop over all
particles) and writes a JSON record.  Strong scaling: the problem is fixed, N varies.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from martini_b200 import dist as mdist  # noqa: E402
from martini_b200 import pipeline  # noqa: E402
from martini_b200.engine import Engine  # noqa: E402
from martini_b200.synthetic import channel_edges  # noqa: E402


def generate(n, nx, ny, nc, device, seed=20260005):
    from martini_b200.synthetic import make_case_device

    return make_case_device("cfg5", device, n=n, nx=nx, nc=nc, seed=seed)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=100_000_000)
    ap.add_argument("--nx", type=int, default=2048)
    ap.add_argument("--nc", type=int, default=512)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--check-pixels", type=int, default=12)
    ap.add_argument("--out", default=None)
    ap.add_argument("--fused", action="store_true",
                    help="store the slabs straight into rank 0's cube over NVLink (no gather pass)")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    eng = Engine(f"cuda:{local}")
    nx = ny = a.nx
    t0 = time.perf_counter()
    case, dev = generate(a.particles, nx, ny, a.nc, eng.device)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    ctx = pipeline.prepare(case)

    # work-balanced slabs from a device-built per-row histogram of box coverage
    r_est = torch.ceil(dev["sm_length"] * 1.6)
    lo = torch.clamp(torch.ceil(dev["px"] - r_est), 0, nx).long()
    hi = torch.clamp(torch.floor(dev["px"] + r_est) + 1, 0, nx).long()
    diff = torch.zeros(nx + 1, dtype=torch.float64, device=eng.device)
    w = (2 * r_est + 1)
    ok = hi > lo
    diff.index_add_(0, lo[ok], w[ok])
    diff.index_add_(0, hi[ok], -w[ok])
    work = torch.cumsum(diff, 0)[:nx].cpu().numpy()
    del r_est, lo, hi, diff, w, ok
    bounds = mdist.slab_bounds(nx, world, work)
    x_lo, x_hi = bounds[rank], bounds[rank + 1]
    fused = a.fused and world > 1
    peer = mdist.PeerCube((nx, ny, a.nc), bounds, eng.device) if fused else None
    slab = peer.rows if fused else torch.zeros((x_hi - x_lo, ny, a.nc), dtype=torch.float64, device=eng.device)
    full = torch.empty((nx, ny, a.nc), dtype=torch.float64, device=eng.device) if (rank == 0 and world > 1 and not fused) else None

    def step():
        slab.zero_()
        out = pipeline.run_hot_path(eng, case, dev=dev, cube=slab, x_lo=x_lo, x_hi=x_hi, zeroed=True, ctx=ctx)
        return out, mdist.gather_slabs(slab, bounds, full, dst=0)

    times, t_comp = [], []
    out = cube = None
    for s in range(a.steps + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        if fused:
            peer.begin()
        else:
            slab.zero_()
        out = pipeline.run_hot_path(eng, case, dev=dev, cube=slab, x_lo=x_lo, x_hi=x_hi, zeroed=True, ctx=ctx)
        e1.record()
        cube = peer.end() if fused else mdist.gather_slabs(slab, bounds, full, dst=0)
        e2.record()
        torch.cuda.synchronize()
        if s > 0:  # first pass sizes the workspaces
            times.append(e0.elapsed_time(e2))
            t_comp.append(e0.elapsed_time(e1))
    t = torch.tensor([float(np.mean(times)), float(np.mean(t_comp))], dtype=torch.float64, device=eng.device)
    u = torch.tensor([float(out["plan"].updates_dense), float(out["plan"].n_pairs), float(out["plan"].n_kept)],
                     dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        per_rank = [torch.zeros_like(u) for _ in range(world)]
        dist.all_gather(per_rank, u)
        dist.all_reduce(u)
    else:
        per_rank = [u.clone()]
    if rank == 0:
        rec = {
            "workload": f"config 5: {a.particles} particles, {nx}x{ny}x{a.nc} cube, WendlandC2Kernel + "
                        "GaussianSpectrum(7 km/s), 64 discs + 10% background", "n_gpus": world,
            "scaling": "strong", "slab_bounds": bounds,
            "assembly": "fused peer stores over NVLink (symmetric memory)" if fused else "NCCL gather",
            "ms_per_insertion": float(t[0]), "ms_compute_max_rank": float(t[1]),
            "updates_dense": float(u[0]), "updates_per_s": float(u[0]) / (float(t[0]) * 1e-3),
            "pairs_per_rank": [float(p[1]) for p in per_rank], "kept_per_rank": [float(p[2]) for p in per_rank],
            "generate_s": t_gen, "cube_bytes": int(nx * ny * a.nc * 8),
        }
        # parity: seeded pixel columns against the reference-structured oracle over ALL particles
        if a.check_pixels > 0:
            from tests.parity import oracle_pixels

            host = {k: dev[k].cpu().numpy() for k in ("px", "py", "pz", "sm_length", "v", "mHI", "D")}
            hcase = dict(case, **host)
            rng = np.random.Generator(np.random.PCG64(55))
            # pixels near disc centres (bright) and anywhere
            pix = [(int(rng.integers(0, nx)), int(rng.integers(0, ny))) for _ in range(a.check_pixels // 2)]
            cx = [int((i + 0.5) * nx / 8) for i in range(8)]
            pix += [(int(np.clip(cx[int(rng.integers(0, 8))] + rng.integers(-40, 40), 0, nx - 1)),
                     int(np.clip(cx[int(rng.integers(0, 8))] + rng.integers(-40, 40), 0, ny - 1)))
                    for _ in range(a.check_pixels - len(pix))]
            t1 = time.perf_counter()
            ref = oracle_pixels(hcase, pix)
            got = np.array([cube[i, j].cpu().numpy() for i, j in pix])
            peak = float(cube.abs().max())
            rec["parity"] = {"pixels": len(pix), "max_abs_diff_over_peak": float(np.abs(got - ref).max() / peak),
                             "ref_max_over_peak": float(np.abs(ref).max() / peak), "oracle_s": time.perf_counter() - t1,
                             "tolerance": 1e-6}
            assert rec["parity"]["max_abs_diff_over_peak"] <= 1e-6
        print(json.dumps(rec))
        if a.out:
            with open(a.out, "w") as f:
                json.dump(rec, f, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
