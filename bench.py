#!/usr/bin/env python
"""Benchmark of the particle->datacube projection (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A *step* is one pass of the whole hot path over one synthetic source: smoothing set-up (K0),
prune (K1), plan, project -- everything `Martini.__init__` + `insert_source_in_cube` do on
the hot path -- into a freshly zeroed cube.  Workload at N = 1: BASELINE config 2 (1e6
particles, 256 x 256 x 128 cube, WendlandC2Kernel + GaussianSpectrum(7 km/s)).  At N > 1
(torchrun, one rank per GPU) the workload is scaled weakly: N such discs side by side in a
(256 N) x 256 x 128 cube, each rank owning one 256-row x-slab (halo particles replicated, every
rank filters the full particle list by footprint), slabs gathered to rank 0 over NCCL.

metric  = particle-channel updates / s, with one update = one (particle, pixel of its
          candidate box, channel) term of the reference's sum (martini.py:279-281):
          U_dense = C * sum_p n_x(p) n_y(p), counted on the device by mtn_plan.
value   = U_dense * steps / device time, inputs resident in HBM.
e2e     = same, but every step also copies the particle arrays from pinned host memory and
          reads the finished cube back to the host (N > 1: every rank reads its slab back over
          its own PCIe link into one host cube shared by the ranks, martini_b200.dist.HostCube).

`--impl reference` times the CPU restatement of the reference algorithm (oracle/, the
reference itself needs astropy, which cannot be installed here) with all host threads on a
bounded sample of the same workload.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle_channel_updates_per_s"
UNIT = "updates/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=("b200", "reference"))
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--particles", type=int, default=None, help="override particles per GPU")
    ap.add_argument("--sample-pixels", type=int, default=512, help="CPU baseline pixel sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-slabs", type=int, default=1,
                    help="x-sub-slabs of the end-to-end leg (each one's read-back overlaps the next one's "
                         "projection); measured on config 2: 7.16 / 7.51 / 7.70 ms for 1 / 2 / 3 -- the extra "
                         "planning passes cost more than the overlap hides, so the default is 1")
    return ap.parse_args()


def make_workload(name, n_gpus, particles=None):
    """Weak-scaled workload: n_gpus copies of the named config stacked along x."""
    from martini_b200 import synthetic

    base = synthetic.make_case(name, n=particles)
    if n_gpus == 1:
        return base
    nx, ny, nc = base["shape"]
    parts = [base] + [synthetic.make_case(name, n=particles, seed=20260002 + 1000 * r)
                      for r in range(1, n_gpus)]
    case = dict(base)
    for k in ("px", "py", "pz", "sm_length", "v", "mHI", "D"):
        case[k] = np.concatenate([p[k] + (nx * r if k == "px" else 0.0) for r, p in enumerate(parts)])
    if np.ndim(base["sigma"]) > 0:
        case["sigma"] = np.concatenate([p["sigma"] for p in parts])
    case["shape"] = (nx * n_gpus, ny, nc)
    return case


def workload_config(case, name, n_gpus):
    nx, ny, nc = case["shape"]
    return {
        "workload": f"BASELINE config 2: synthetic SPHSource, {case['px'].size // n_gpus:d} "
                    f"particles and a {nx // n_gpus}x{ny}x{nc} cube per GPU, "
                    f"{case['kernel'][0]} + {case['spectrum']} spectrum" if name == "cfg2" else name,
        "particles": int(case["px"].size), "cube": [int(nx), int(ny), int(nc)],
        "kernel": case["kernel"][0], "spectrum": case["spectrum"],
        "partition": "1 GPU" if n_gpus == 1 else f"{n_gpus} x-slabs of {nx // n_gpus} rows, halo "
                     "particles replicated, slabs stored into rank 0's cube over NVLink (peer "
                     "stores from the projection kernel; NCCL gather as fallback)",
        "l2": "L2 flushed between timed steps by writing a 512 MiB buffer",
    }


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.proc = None
        self.idx = gpu_index
        self.path = f"/tmp/mtn_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "nvidia-smi unavailable"}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons, power = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
                power.append(float(p[3]))
            except ValueError:
                continue
            for nm, val in zip(names, p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "no samples"}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)),
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": float(np.max(power))}


# ------------------------------------------------------------------------------ CPU arm
def cpu_reference_arm(case, n_sample, steps, warmup, ncpu=None):
    """Time the reference algorithm restated on the CPU (oracle/martini_oracle.py, numpy +
    scipy, ThreadPool over pixels like martini.py:345-362) on a bounded sample.

    Full-cube time is extrapolated as T = T_spectra + (X*Y / n_sample) * T_sample: the
    reference's pixel loop costs the same O(N) mask scan at every pixel.  Returns
    (updates_per_s, seconds_per_full_insertion, description, cores).
    """
    from oracle import martini_oracle as O
    from tests.parity import SPEC, oracle_prepare

    ncpu = ncpu or (os.cpu_count() or 1)
    nx, ny, nc = case["shape"]
    t0 = time.perf_counter()
    k, kind, pix, acc = oracle_prepare(case)
    k.apply_mask(acc)
    p = pix[:, acc]
    sig = case["sigma"]
    sig = sig[acc] if np.ndim(sig) > 0 else sig
    t_prune = time.perf_counter() - t0
    t0 = time.perf_counter()
    spectra = O.init_spectra(kind, case["edges"], case["v"][acc], sig, case["mHI"][acc], case["D"][acc])
    t_spec = time.perf_counter() - t0
    u_dense = O.count_updates(p, k.sm_ranges, nx, ny, nc)
    rng = np.random.Generator(np.random.PCG64(12345))
    times = []
    for s in range(warmup + steps):
        pixels = [(int(i), int(j)) for i, j in zip(rng.integers(0, nx, n_sample), rng.integers(0, ny, n_sample))]
        t0 = time.perf_counter()
        O.insert_pixels(pixels, p, k, spectra, ncpu=ncpu)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    t_sample = float(np.mean(times))
    t_full = t_prune + t_spec + t_sample * (nx * ny / n_sample)
    desc = (f"oracle port of the reference loop: prune + init_spectra for all "
            f"{int(acc.sum())} kept particles ({t_prune + t_spec:.1f} s, once) + {n_sample} seeded "
            f"pixels per step through ThreadPool({ncpu}), extrapolated x{nx * ny / n_sample:.0f} "
            f"to the {nx}x{ny} pixel loop")
    return u_dense / t_full, t_full, desc, ncpu, t_sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    case = make_workload(args.workload, 1, args.particles)
    val, t_full, desc, ncpu, t_sample = cpu_reference_arm(case, args.sample_pixels, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_sample * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(case, args.workload, 1),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": ncpu, "kind": "port", "sample": desc,
                         "full_insertion_s": t_full},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference itself cannot be imported (astropy missing, no network); this is "
                "the astropy-free restatement in oracle/, a lower bound on the reference's time",
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------ GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    from martini_b200 import dist as mdist
    from martini_b200 import pipeline
    from martini_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    n_gpus = world
    if args.gpus != world and rank == 0 and world > 1:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}", file=sys.stderr)
    eng = Engine(f"cuda:{local}")
    dev_t = eng.device

    case = make_workload(args.workload, n_gpus, args.particles)
    ctx = pipeline.prepare(case)
    nx, ny, nc = ctx.shape
    bounds = mdist.slab_bounds(nx, n_gpus)  # the weak-scaled discs are identical: equal rows
    x_lo, x_hi = bounds[rank], bounds[rank + 1]
    pinned = pipeline.pin_case(case)
    dev = pipeline.upload(eng, case, pinned)
    # N > 1: the slabs are stored straight into rank 0's cube over NVLink (fused assembly,
    # martini_b200.dist.PeerCube); NCCL gather if symmetric memory is unavailable
    peer = None
    if world > 1:
        try:
            peer = mdist.PeerCube((nx, ny, nc), bounds, dev_t)
        except Exception as exc:  # noqa: BLE001
            if rank == 0:
                print(f"symmetric memory unavailable ({exc}); falling back to the NCCL gather", file=sys.stderr)
        ok = torch.tensor([1.0 if peer is not None else 0.0], device=dev_t)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok) == 0.0:
            peer = None
    slab = peer.rows if peer is not None else torch.zeros((x_hi - x_lo, ny, nc), dtype=torch.float64, device=dev_t)
    full = torch.empty((nx, ny, nc), dtype=torch.float64, device=dev_t) if (world > 1 and rank == 0 and peer is None) else None
    # end-to-end leg: the result is read back into one host cube shared by the ranks; every
    # rank copies its own slab over its own PCIe link (martini_b200.dist.HostCube)
    host_cube = mdist.HostCube((nx, ny, nc), bounds)
    slab_e2e = slab if peer is None else torch.zeros((x_hi - x_lo, ny, nc), dtype=torch.float64, device=dev_t)
    copy_stream = torch.cuda.Stream(device=dev_t)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev_t)

    def step(e2e=False):
        if e2e:
            # host buffers in, host cube out: upload, project into the local slab, read it back
            pipeline.upload(eng, case, pinned, out=dev)
            return pipeline.run_hot_path_to_host(eng, case, host_cube.rows, dev, ctx, slab_e2e, x_lo=x_lo,
                                                 x_hi=x_hi, n_slabs=args.e2e_slabs, copy_stream=copy_stream)
        if peer is not None:
            peer.begin()  # rank 0 zeroes the cube, barrier
        else:
            slab.zero_()
        out = pipeline.run_hot_path(eng, case, dev=dev, cube=slab, x_lo=x_lo, x_hi=x_hi, zeroed=True, ctx=ctx)
        if peer is not None:
            peer.end()  # barrier: every rank's stores have landed in rank 0's cube
        elif world > 1:
            mdist.gather_slabs(slab, bounds, full, dst=0)
        return out

    def timed(n_steps, e2e=False, stage_times=None):
        tot = 0.0
        out = None
        for _ in range(n_steps):
            flush.fill_(1)  # evict L2
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = step(e2e)
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
            if stage_times is not None:
                for k, v in eng.last_timing_ms().items():
                    stage_times.setdefault(k, []).append(v)
        t = torch.tensor([tot], dtype=torch.float64, device=dev_t)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t) / n_steps, out

    # warm-up (also sizes the workspaces), then the timed region with clocks sampled
    timed(max(args.warmup, 3))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.set_timing(True)
    stage = {}
    ms_step, out = timed(args.steps, stage_times=stage)
    eng.set_timing(False)
    ms_e2e, _ = timed(args.steps, e2e=True)
    clocks = sampler.stop() if rank == 0 else None

    # whole-job units: every rank's slab updates (device-counted), summed
    u = torch.tensor([out["plan"].updates_dense], dtype=torch.float64, device=dev_t)
    pairs = torch.tensor([out["plan"].n_pairs], dtype=torch.float64, device=dev_t)
    if world > 1:
        dist.all_reduce(u)
        dist.all_reduce(pairs)
    u_dense = float(u)

    # executed algorithmic work of the projection kernel (diagnostic pass, untimed)
    eng.set_count_exec(True)
    step()
    torch.cuda.synchronize()
    ex = eng.last_exec_counts()
    eng.set_count_exec(False)

    if rank == 0:
        p64 = eng.fp64_peak_tflops()
        t_proj = float(np.mean(stage["project"])) * 1e-3
        flops_fma = 2.0 * ex["updates"]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        plan = out["plan"]
        bytes_alg = 88.0 * plan.n_pairs + 8.0 * (x_hi - x_lo) * ny * nc
        # DRAM bytes of one launch from the committed ncu --set full capture of this workload
        # (profiles/r1_final_project_kernel.md); only meaningful for the default config at N = 1
        traffic = None
        if n_gpus == 1 and args.workload == "cfg2" and args.particles is None:
            try:
                import re

                txt = open(os.path.join(ROOT, "profiles", "r1_final_project_kernel.md")).read()
                rd = re.search(r"dram__bytes_read\.sum` = ([0-9.]+) Mbyte", txt)
                wr = re.search(r"dram__bytes_write\.sum` = ([0-9.]+) Mbyte", txt)
                if rd and wr:
                    traffic = (float(rd.group(1)) + float(wr.group(1))) * 1e6
            except OSError:
                pass
        roofline = {
            "kernel": "project_kernel", "bound": "fp64",
            "achieved": flops_fma / t_proj / 1e12, "peak": p64, "unit": "TFLOP/s",
            "frac": flops_fma / t_proj / 1e12 / p64, "traffic": traffic,
            "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, bytes, ncu capture "
                            "profiles/r1_final_project_kernel.md; algorithmic bytes are in hbm.achieved",
            "peak_source": "FP64 FMA microbenchmark run on this GPU in this process "
                           "(mtn_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
            "algorithmic": {"fma_updates": ex["updates"], "kernel_integrals": ex["weights"],
                            "edge_erfs": ex["erfs"], "note": "rank 0 slab, one launch"},
            "kernel_ms": t_proj * 1e3,
            "hbm": {"achieved": bytes_alg / t_proj / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": bytes_alg / t_proj / 1e9 / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s"},
            "stage_ms": {k: float(np.mean(v)) for k, v in stage.items()},
        }
        h2d = pipeline.h2d_bytes(case)
        d2h = int(nx * ny * nc * 8)
        line = {
            "metric": METRIC, "value": u_dense / (ms_step * 1e-3), "unit": UNIT, "n_gpus": n_gpus,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(case, args.workload, n_gpus),
            "updates_per_step": u_dense, "insertion_wall_ms": ms_step,
            "e2e": {"value": u_dense / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "result": ("host cube shared by the ranks (POSIX shared memory, page-locked), "
                               "each rank copies its own slab" if n_gpus > 1 else "pinned host cube")
                              + f"; pipeline.run_hot_path_to_host with {args.e2e_slabs} x-sub-slab(s) per rank"},
            "gpu_launches": int(out["launches"]) * args.steps,
            "gpu_launches_per_step": int(out["launches"]),
            "clocks": clocks, "roofline": roofline,
        }
        if n_gpus == 1:
            # for information: the same insertion through the MARTINI-compatible classes
            # (Martini.__init__ prunes the host objects with numpy, insert_source_in_cube copies
            # the cube back into DataCube._array) -- not part of any timed region above
            from martini_b200 import DataCube, Martini, PixelSource
            from martini_b200.spectral_models import DiracDeltaSpectrum, GaussianSpectrum

            t0 = time.perf_counter()
            dc = DataCube(n_px_x=nx, n_px_y=ny, n_channels=nc, px_size=case["px_size"],
                          channel_width=float(abs(case["edges"][1] - case["edges"][0])))
            spec = GaussianSpectrum(sigma=case["sigma"]) if case["spectrum"] == "gaussian" else DiracDeltaSpectrum()
            if case["spectrum"] == "gaussian" and np.ndim(case["sigma"]) > 0:
                spec.half_width = lambda source, _s=case["sigma"]: _s  # per-particle widths as given
            m = Martini(source=PixelSource.from_case(case), datacube=dc, spectral_model=spec,
                        sph_kernel=pipeline.kernel_from_spec(case["kernel"]), quiet=True, engine=eng)
            m.insert_source_in_cube(skip_validation=True)
            line["martini_class_wall_ms"] = (time.perf_counter() - t0) * 1e3
        if n_gpus == 1 and not args.no_cpu_baseline:
            val, t_full, desc, ncpu, _ = cpu_reference_arm(
                make_workload(args.workload, 1, args.particles), args.sample_pixels, 3, 1)
            line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": ncpu, "kind": "port",
                                    "sample": desc, "full_insertion_s": t_full}
        print(json.dumps(line))
    host_cube.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
