#!/usr/bin/env python
"""Benchmark of the particle->datacube projection (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A *step* is one pass of the whole hot path over one synthetic source: smoothing set-up (K0),
prune (K1), plan, project -- everything `Martini.__init__` + `insert_source_in_cube` do on
the hot path -- into a freshly zeroed cube.

N = 1   headline = BASELINE config 3 at full size (1e7 particles, 512 x 512 x 256 cube,
        adaptive CubicSplineKernel + GaussianSpectrum(sigma="thermal")): the largest
        configuration BASELINE.json lists for one GPU.  Configs 2 and 4 (full size) are
        measured in the same run and reported under ``other_configs`` with their own
        ms_per_step / roofline / e2e / cpu_baseline.
N > 1   STRONG scaling of the same config-3 cube (torchrun, one rank per GPU): the cube is
        cut into work-balanced x-slabs, every rank projects its slab (halo particles are
        processed by both neighbours) and stores it straight into rank 0's cube over NVLink
        (martini_b200.dist.PeerCube; NCCL gather as fallback).  At N = 8 the north-star
        target (config 5: 1e8 particles -> 2048 x 2048 x 512) runs as well, under
        ``extra.cfg5`` (time, updates/s, per-rank balance, oracle pixel check).

metric  = particle-channel updates / s, one update = one (particle, pixel of its candidate
          box, channel) term of the reference's sum (martini.py:279-281):
          U_dense = C * sum_p n_x(p) n_y(p), counted on the device by mtn_plan.
value   = U_dense * steps / device time, inputs resident in HBM.
e2e     = the same with HOST buffers: every step copies the particle arrays from pinned host
          memory and reads the finished cube back to the host (N > 1: every rank reads its
          slab back over its own PCIe link into one host cube shared by the ranks).

`--impl reference` times the CPU restatement of the reference algorithm (oracle/, the
reference itself needs astropy, which cannot be installed here) with all host threads on a
bounded pixel sample of the SAME workload (config 3 at every N; rank 0 only).
"""

from __future__ import annotations

import argparse
import hashlib
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
HEADLINE = "cfg3"

DESCR = {
    "cfg2": "BASELINE config 2: synthetic SPHSource, 1e6 particles, 256x256x128 cube, "
            "WendlandC2Kernel + GaussianSpectrum(7 km/s)",
    "cfg3": "BASELINE config 3: synthetic TNG-like source, 1e7 particles, 512x512x256 cube, adaptive "
            "CubicSplineKernel + GaussianSpectrum(sigma='thermal')",
    "cfg4": "BASELINE config 4: 1e7 particles with large smoothing lengths (8-40 px), 512x512x256 cube, "
            "GaussianKernel(truncate=3) + DiracDeltaSpectrum",
    "cfg5": "BASELINE config 5: 1e8 particles, 2048x2048x512 cube, WendlandC2Kernel + "
            "GaussianSpectrum(7 km/s), 64 discs + 10 % background",
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=("b200", "reference"))
    ap.add_argument("--workload", default=HEADLINE)
    ap.add_argument("--others", default="cfg2,cfg4",
                    help="N = 1: further configs measured into other_configs ('none' to skip)")
    ap.add_argument("--other-steps", type=int, default=3)
    ap.add_argument("--particles", type=int, default=None, help="override the particle count")
    ap.add_argument("--sample-pixels", type=int, default=256, help="CPU baseline pixel sample per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-full", action="store_true",
                    help="also run config 2 on the CPU in full, once (minutes) and record it under profiles/")
    ap.add_argument("--no-cfg5", action="store_true", help="N = 8: skip the config-5 target run")
    ap.add_argument("--cfg5-particles", type=int, default=100_000_000)
    ap.add_argument("--force-cfg5", action="store_true", help="run the config-5 block at any N > 1 (testing)")
    ap.add_argument("--no-class", action="store_true", help="skip the Martini-class wall time")
    return ap.parse_args()


def workload_config(name, case=None, particles=None):
    """The `config` object: identical for both arms and for every N (strong scaling)."""
    cfg = {"workload": DESCR.get(name, name) + (f" [particles overridden: {particles}]" if particles else ""),
           "l2": "inputs + cube exceed the 126 MB L2 and a 512 MiB buffer is written between timed steps"}
    if case is not None:
        nx, ny, nc = case["shape"]
        cfg.update(particles=int(np.size(case["px"])) if "px" in case else None,
                   cube=[int(nx), int(ny), int(nc)], kernel=case["kernel"][0], spectrum=case["spectrum"])
    return cfg


def csrc_hash():
    """Hash of the CUDA sources + header: ties a bench line to the ncu summaries under profiles/."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "martini_b200", "csrc")
    for f in sorted(os.listdir(d)) + ["../../include/martini_b200.h"]:
        with open(os.path.join(d, f), "rb") as fh:
            h.update(f.encode())
            h.update(fh.read())
    return h.hexdigest()[:16]


def profiled_traffic(name, kernel):
    """DRAM bytes of one launch of `kernel` from the committed ncu --set full summary of the SAME
    source hash and workload (profiles/ncu_summaries.json, written by
    profiles/make_ncu_summary.py); (None, reason) if there is none."""
    try:
        db = json.load(open(os.path.join(ROOT, "profiles", "ncu_summaries.json")))
    except (OSError, ValueError):
        return None, "profiles/ncu_summaries.json missing"
    sha = csrc_hash()
    for e in db.get("captures", []):
        if e.get("csrc_hash") == sha and e.get("workload") == name and str(e.get("kernel", "")).find(kernel) >= 0:
            return float(e["dram_bytes_read"]) + float(e["dram_bytes_write"]), f"profiles/{e.get('file', '')}"
    return None, f"no ncu capture of {kernel} on {name} for csrc hash {sha} under profiles/"


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


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
def cpu_reference_arm(case, n_sample, steps, warmup, ncpu=None, ncpu1_pixels=8, budget_s=4.0):
    """Time the reference algorithm restated on the CPU (oracle/martini_oracle.py: numpy + scipy,
    ThreadPool over pixels like martini.py:345-362) on a bounded pixel sample.

    Per pixel the reference scans ALL particles for the candidate mask (martini.py:272-274),
    evaluates the kernel weights of the masked ones and sums their spectra.  When the N x C
    spectra array fits in 4 GB it is materialised first, as the reference does
    (spectral_models.py:63-147); otherwise (configs 3 and 4: 20 GB) the spectra are evaluated
    for the masked particles of each sampled pixel only -- less work than the reference does.
    The full-cube time is extrapolated as T = T_once + (X*Y / n_sample) * T_sample.
    """
    from oracle import martini_oracle as O
    from tests.parity import oracle_prepare

    ncpu = ncpu or (os.cpu_count() or 1)
    nx, ny, nc = case["shape"]
    t0 = time.perf_counter()
    k, kind, pix, acc = oracle_prepare(case)
    k.apply_mask(acc)
    p = np.ascontiguousarray(pix[:, acc])
    sig = case["sigma"]
    sig = sig[acc] if np.ndim(sig) > 0 else sig
    v, mHI, D = case["v"][acc], case["mHI"][acc], case["D"][acc]
    t_once = time.perf_counter() - t0
    materialise = p.shape[1] * nc * 8 <= (4 << 30)
    if materialise:
        t0 = time.perf_counter()
        spectra = O.init_spectra(kind, case["edges"], v, sig, mHI, D)
        t_once += time.perf_counter() - t0

        def pixel(ij):
            return O.evaluate_pixel_spectrum(ij, p, k, spectra)
    else:
        def pixel(ij):
            ijc = np.array(ij)[..., np.newaxis]
            mask = (np.abs(ijc - p[:2]) <= k.sm_ranges).all(axis=0)
            sel = np.flatnonzero(mask)
            w = k.px_weight(p[:2, sel] - ijc, mask=sel)
            sp = O.init_spectra(kind, case["edges"], v[sel], sig if np.ndim(sig) == 0 else sig[sel],
                                mHI[sel], D[sel])
            np.multiply(sp, w[:, np.newaxis], out=sp)
            return np.sum(sp, axis=-2)

    u_dense = O.count_updates(p, k.sm_ranges, nx, ny, nc)
    rng = np.random.Generator(np.random.PCG64(12345))

    def sample(n):
        return [(int(i), int(j)) for i, j in zip(rng.integers(0, nx, n), rng.integers(0, ny, n))]

    # single-thread leg first (the reference's default ncpu=1); it also sizes the threaded sample
    # so that one step stays near `budget_s` seconds whatever the config costs per pixel
    pixels = sample(ncpu1_pixels)
    t0 = time.perf_counter()
    for ij in pixels:
        pixel(ij)
    t1 = (time.perf_counter() - t0) / ncpu1_pixels
    t_full_1 = t_once + t1 * nx * ny
    n_sample = int(min(n_sample, max(ncpu, budget_s * ncpu * 0.5 / max(t1, 1e-6))))

    from multiprocess.pool import ThreadPool

    times = []
    with ThreadPool(processes=ncpu) as pool:
        for s in range(warmup + steps):
            pixels = sample(n_sample)
            t0 = time.perf_counter()
            list(pool.imap(pixel, pixels))
            dt = time.perf_counter() - t0
            if s >= warmup:
                times.append(dt)
    t_sample = float(np.mean(times))
    t_full = t_once + t_sample * (nx * ny / n_sample)
    desc = (f"oracle port of the reference loop on {int(acc.sum())} kept particles: set-up"
            f"{' + init_spectra (N x C materialised)' if materialise else ''} {t_once:.1f} s once, then "
            f"{n_sample} seeded pixels per step through ThreadPool({ncpu})"
            f"{'' if materialise else ', spectra evaluated for the masked particles of each pixel only'}, "
            f"extrapolated x{nx * ny / n_sample:.0f} to the {nx}x{ny} pixel loop")
    return {"value": u_dense / t_full, "unit": UNIT, "cores": ncpu, "kind": "port", "sample": desc,
            "extrapolated_s": t_full, "ms_per_sample_step": t_sample * 1e3,
            "ncpu1_updates_per_s": u_dense / t_full_1, "ncpu1_extrapolated_s": t_full_1,
            "ncpu1_sample": f"{ncpu1_pixels} seeded pixels, one thread", "cpu_model": cpu_model(),
            "updates_dense": u_dense}


def recorded_full_run(name):
    """A full (not extrapolated) CPU run recorded by `bench.py --cpu-full` on a GPU box."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", f"cpu_full_{name}.json")))
        return {"full_run_s": rec["full_run_s"], "full_run_cores": rec["cores"],
                "full_run_cpu_model": rec["cpu_model"], "full_run_source": f"profiles/cpu_full_{name}.json (recorded run)"}
    except (OSError, ValueError, KeyError):
        return {}


def cpu_full_run(name="cfg2"):
    """The whole config-2 insertion on the CPU, reference-structured (every pixel scans all
    particles), ThreadPool over all cores.  Minutes; run once, recorded under profiles/."""
    from martini_b200 import synthetic
    from oracle import martini_oracle as O
    from tests.parity import oracle_prepare

    case = synthetic.make_case(name)
    ncpu = os.cpu_count() or 1
    nx, ny, nc = case["shape"]
    t0 = time.perf_counter()
    k, kind, pix, acc = oracle_prepare(case)
    k.apply_mask(acc)
    p = pix[:, acc]
    sig = case["sigma"]
    spectra = O.init_spectra(kind, case["edges"], case["v"][acc], sig[acc] if np.ndim(sig) > 0 else sig,
                             case["mHI"][acc], case["D"][acc])
    cube = O.insert_source_in_cube(np.zeros((nx, ny, nc)), p, k, spectra, case["px_size"], ncpu=ncpu,
                                   skip_validation=True)
    t = time.perf_counter() - t0
    rec = {"workload": DESCR[name], "full_run_s": t, "cores": ncpu, "cpu_model": cpu_model(),
           "updates_dense": O.count_updates(p, k.sm_ranges, nx, ny, nc), "cube_sum": float(cube.sum()),
           "cube_peak": float(cube.max())}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for d in ("gpurun_out", "profiles"):
        with open(os.path.join(ROOT, d, f"cpu_full_{name}.json"), "w") as f:
            json.dump(rec, f, indent=1)
    return rec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from martini_b200 import synthetic

    case = synthetic.make_case(args.workload, n=args.particles)
    cb = cpu_reference_arm(case, args.sample_pixels, args.steps, args.warmup)
    cb.update(recorded_full_run(args.workload))
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_sample_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args.workload, case, args.particles),
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference itself cannot be imported (astropy missing, no network); this is the "
                "astropy-free restatement in oracle/, a lower bound on the reference's time; ms_per_step "
                "is one bounded pixel-sample step, value is U_dense / the extrapolated full insertion",
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------ GPU arm
class Timer:
    """CUDA-event timing of `fn` on torch's current stream, max over ranks, L2 flushed between
    steps."""

    def __init__(self, dev_t, world):
        import torch

        self.torch, self.dev_t, self.world = torch, dev_t, world
        self.flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev_t)

    def run(self, fn, n_steps, after=None):
        torch = self.torch
        import torch.distributed as dist

        tot, out = 0.0, None
        for _ in range(n_steps):
            self.flush.fill_(1)  # evict L2
            if self.world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
            if after is not None:
                after()
        t = torch.tensor([tot], dtype=torch.float64, device=self.dev_t)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t) / n_steps, out


def roofline_block(eng, name, plan, ex, stage, slab_voxels, n_gpus):
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    p64 = eng.fp64_peak_tflops()
    # the kernels that compute the sum: the brick kernel and, if the insertion has one, the
    # column / splat kernel of its second stream (executed-work counts cover both)
    t_brick, t_stream2 = float(np.mean(stage["project"])) * 1e-3, float(np.mean(stage.get("stream2", [0.0]))) * 1e-3
    t_proj = t_brick + t_stream2
    route2 = {0: None, 1: "column_kernel", 2: "splat_kernel"}[int(getattr(plan, "route2", 0))]
    dominant = "project_kernel" if t_brick >= t_stream2 else route2
    flops_fma = 2.0 * ex["updates"]
    # F_alg = 2 U_exec + K_w P_exec + K_s N (C_live + 1) with the flop counts of the
    # evaluations as implemented here: one kernel integral = a degree-9 Horner (18) + R^2 and
    # the 1/h^2 scale (6) = 24 flops (the Gaussian kernel's five erfs are counted as five edge
    # evaluations); one edge erf = a degree-9 Horner (18) + argument (3) = 21 flops
    kw, ks = 24.0, 21.0
    flops_full = flops_fma + kw * ex["weights"] + ks * ex["erfs"]
    bytes_alg = 88.0 * (plan.n_pairs + plan.n_pairs2) + 8.0 * slab_voxels
    traffic, src = profiled_traffic(name, dominant) if n_gpus == 1 else (None, "N > 1: not profiled per rank")
    return {
        "kernel": dominant, "kernels_ms": {"project_kernel": t_brick * 1e3, **({route2: t_stream2 * 1e3} if route2 else {})},
        "bound": "fp64",
        "achieved": flops_fma / t_proj / 1e12, "peak": p64, "unit": "TFLOP/s",
        "frac": flops_fma / t_proj / 1e12 / p64,
        "achieved_full": flops_full / t_proj / 1e12, "frac_full": flops_full / t_proj / 1e12 / p64,
        "flops_note": "achieved/frac count 2 flops per executed (non-zero weight x non-zero spectrum) "
                      "update only; *_full adds 24 flops per kernel integral and 21 per edge erf (SURVEY 8d F_alg)",
        "traffic": traffic, "traffic_source": src,
        "peak_source": "FP64 FMA microbenchmark run on this GPU in this process (mtn_fp64_peak); "
                       "MEASURED_PEAKS.json has no FP64 entry",
        "algorithmic": {"fma_updates": ex["updates"], "kernel_integrals": ex["weights"],
                        "edge_erfs": ex["erfs"], "pairs": int(plan.n_pairs), "pairs2": int(plan.n_pairs2),
                        "kept": int(plan.n_kept),
                        "note": "rank 0 slab, one launch"},
        "kernel_ms": t_proj * 1e3,
        "hbm": {"achieved": bytes_alg / t_proj / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": bytes_alg / t_proj / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s"},
        "stage_ms": {k: float(np.mean(v)) for k, v in stage.items()},
        "csrc_hash": csrc_hash(),
    }


def class_path(eng, case):
    """The same insertion through the MARTINI-compatible classes -- the call a user makes:
    Martini(source=, datacube=, sph_kernel=, spectral_model=) + insert_source_in_cube() +
    datacube._array on the host.  The source's arrays are page-locked once (input preparation,
    like the pinned buffers of the array-level leg); every step builds the source / cube / kernel
    / Martini objects afresh, uploads, prunes, inserts and reads the cube back.
    Returns (step function, bytes in, bytes out)."""
    import torch

    from martini_b200 import DataCube, Martini, PixelSource, pipeline
    from martini_b200.spectral_models import DiracDeltaSpectrum, GaussianSpectrum

    nx, ny, nc = case["shape"]
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    pix = pin(np.vstack((case["px"], case["py"], case["pz"])))
    arr = {k: pin(case[k]) for k in ("sm_length", "v", "D", "mHI")}
    sigma = pin(case["sigma"]) if np.ndim(case["sigma"]) > 0 else case["sigma"]
    kernel_spec = case["kernel"]
    h2d = pix.nbytes + sum(a.nbytes for a in arr.values()) + (sigma.nbytes if np.ndim(sigma) > 0 else 0)
    state = {}

    def step():
        src = PixelSource(pixcoords=pix, sm_lengths=arr["sm_length"], radial_velocity=arr["v"],
                          distance_p=arr["D"], mHI_g=arr["mHI"])
        dc = DataCube(n_px_x=nx, n_px_y=ny, n_channels=nc, px_size=case["px_size"],
                      channel_width=float(abs(case["edges"][1] - case["edges"][0])))
        spec = GaussianSpectrum(sigma=7.0) if case["spectrum"] == "gaussian" else DiracDeltaSpectrum()
        if case["spectrum"] == "gaussian":
            spec.half_width = lambda source, _s=sigma: _s  # the case's line widths as given (scalar or per particle)
        m = Martini(source=src, datacube=dc, spectral_model=spec, sph_kernel=pipeline.kernel_from_spec(kernel_spec),
                    quiet=True, engine=eng)
        m.insert_source_in_cube(skip_validation=True)
        state["t_device_done"] = time.perf_counter()
        host = m.datacube._array
        assert host.shape[:3] == (nx, ny, nc)
        state["plan"] = m.last_plan
        return m

    return step, h2d, nx * ny * nc * 8, state


def measure_single(eng, name, args, steps, warmup, timer, clocks_for=None):
    """One GPU, one config at full size: device-resident value, e2e with host buffers, executed
    work, roofline, CPU baseline.  Returns the block that goes into the JSON line."""
    import torch

    from martini_b200 import pipeline, synthetic

    case = synthetic.make_case(name, n=args.particles)
    ctx = pipeline.prepare(case)
    nx, ny, nc = ctx.shape
    pinned = pipeline.pin_case(case)
    dev = pipeline.upload(eng, case, pinned)
    slab = torch.zeros((nx, ny, nc), dtype=torch.float64, device=eng.device)
    host = torch.empty((nx, ny, nc), dtype=torch.float64).pin_memory()
    copy_stream = torch.cuda.Stream(device=eng.device)

    def step():
        slab.zero_()
        return pipeline.run_hot_path(eng, case, dev=dev, cube=slab, zeroed=True, ctx=ctx)

    def step_e2e():
        pipeline.upload(eng, case, pinned, out=dev)
        return pipeline.run_hot_path_to_host(eng, case, host, dev, ctx, slab, copy_stream=copy_stream)

    timer.run(step, max(warmup, 3))
    sampler = None
    if clocks_for is not None:
        sampler = ClockSampler(clocks_for)
        sampler.start()
    eng.set_timing(True)
    stage = {}

    def grab():
        for k, v in eng.last_timing_ms().items():
            stage.setdefault(k, []).append(v)

    ms_step, out = timer.run(step, steps, after=grab)
    eng.set_timing(False)
    timer.run(step_e2e, 1)
    ms_e2e, _ = timer.run(step_e2e, steps)
    clocks = sampler.stop() if sampler is not None else None
    plan = out["plan"]
    u_dense = float(plan.updates_dense)
    eng.set_count_exec(True)
    step()
    torch.cuda.synchronize()
    ex = eng.last_exec_counts()
    eng.set_count_exec(False)
    blk = {
        "config": workload_config(name, case, args.particles),
        "value": u_dense / (ms_step * 1e-3), "unit": UNIT, "ms_per_step": ms_step, "steps": steps,
        "updates_per_step": u_dense, "insertion_wall_ms": ms_step,
        "e2e_pipeline": {"value": u_dense / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                         "h2d_bytes_per_step": pipeline.h2d_bytes(case), "d2h_bytes_per_step": int(nx * ny * nc * 8),
                         "result": "array level: pinned host arrays -> pipeline.run_hot_path_to_host (upload, K0, "
                                   "K1, plan, project, read-back into a pinned host cube)"},
        "gpu_launches_per_step": int(out["launches"]),
        "roofline": roofline_block(eng, name, plan, ex, stage, nx * ny * nc, 1),
    }
    blk["e2e"] = blk["e2e_pipeline"]
    if clocks is not None:
        blk["clocks"] = clocks
    if not args.no_class:
        # the headline end-to-end figure: the same insertion through the public classes
        del slab, host
        torch.cuda.empty_cache()
        try:
            cstep, h2d, d2h, state = class_path(eng, case)
            timer.run(cstep, 2)
            ms_class, _ = timer.run(cstep, steps)
            assert state["plan"].updates_dense == plan.updates_dense
            blk["e2e"] = {"value": u_dense / (ms_class * 1e-3), "unit": UNIT, "ms_per_step": ms_class,
                          "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                          "result": "through Martini.insert_source_in_cube: Martini(source=PixelSource(page-locked "
                                    "host arrays), datacube=DataCube(...), sph_kernel=, spectral_model=) + "
                                    "insert_source_in_cube() + datacube._array on the host (pooled page-locked buffer)"}
            blk["martini_class_wall_ms"] = ms_class
        except Exception as exc:  # noqa: BLE001 -- the array-level figure stands in, and says so
            blk["martini_class_error"] = repr(exc)[:300]
        slab = host = None
    del slab, host, dev, pinned
    torch.cuda.empty_cache()
    if not args.no_cpu_baseline:
        cb = cpu_reference_arm(case, args.sample_pixels, 2, 1)
        cb.update(recorded_full_run(name))
        if "full_run_s" in cb:
            cb["extrapolated_over_full"] = cb["extrapolated_s"] / cb["full_run_s"]
        blk["cpu_baseline"] = cb
    return blk


def balanced_bounds(eng, dev, ctx, world):
    """Work-balanced slab boundaries from a per-row histogram (box height x rows covered) built
    on the device from every rank's share of the particles and summed over the ranks: identical
    on all of them."""
    import torch
    import torch.distributed as dist

    from martini_b200 import dist as mdist

    nx = ctx.shape[0]
    kid, valid, sm_range, h_eff = eng.smoothing_setup(dev["sm_length"], ctx.table)
    r = torch.nan_to_num(sm_range, posinf=float(nx))
    lo = torch.clamp(torch.ceil(dev["px"] - r), 0, nx).long()
    hi = torch.clamp(torch.floor(dev["px"] + r) + 1, 0, nx).long()
    ok = hi > lo
    w = (2 * r + 1 + 8.0)[ok]  # box height + a per-particle constant (set-up, spectrum)
    diff = torch.zeros(nx + 1, dtype=torch.float64, device=eng.device)
    diff.index_add_(0, lo[ok], w)
    diff.index_add_(0, hi[ok], -w)
    if world > 1:
        dist.all_reduce(diff)
    work = torch.cumsum(diff, 0)[:nx].cpu().numpy()
    return mdist.slab_bounds(nx, world, work)


def measure_strong(eng, name, args, timer, world, rank, local, case=None, dev=None, steps=None, warmup=None,
                   want_e2e=True):
    """N GPUs, one cube.  Every rank holds a contiguous 1/N share of the particle list; a step
    routes the particles to the ranks whose x-slab they reach (K0 on the share, one all-to-all
    over NVLink, dist.route_particles), then K0 / K1 / plan / project on what arrived, the
    projection kernels storing straight into rank 0's cube (dist.PeerCube)."""
    import torch
    import torch.distributed as dist

    from martini_b200 import dist as mdist
    from martini_b200 import pipeline, synthetic

    steps = steps or args.steps
    warmup = warmup or max(args.warmup, 3)
    dev_t = eng.device
    pinned = None
    if case is None:
        case = synthetic.make_case(name, n=args.particles)
        n_all = case["px"].size
        a, b = mdist.chunk_of(n_all, rank, world)
        share = dict(case, **{k: case[k][a:b] for k in pipeline.particle_keys(case)})
        pinned = pipeline.pin_case(share)
        dev = pipeline.upload(eng, share, pinned)
    else:  # device-generated case: keep this rank's share only
        n_all = dev["px"].numel()
        a, b = mdist.chunk_of(n_all, rank, world)
        dev = {k: (v[a:b].clone() if isinstance(v, torch.Tensor) and v.ndim == 1 and v.numel() == n_all else v)
               for k, v in dev.items()}
        share = case
        torch.cuda.empty_cache()
    ctx = pipeline.prepare(case)
    nx, ny, nc = ctx.shape
    bounds = balanced_bounds(eng, dev, ctx, world)
    x_lo, x_hi = bounds[rank], bounds[rank + 1]
    peer = None
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
    full = torch.empty((nx, ny, nc), dtype=torch.float64, device=dev_t) if (rank == 0 and peer is None) else None
    route_ms = []
    # the input exchange: fused bucket + peer stores into the ranks' inboxes (dist.PeerRouter);
    # all-to-all over NCCL if symmetric memory is unavailable
    router = None
    if peer is not None:
        try:
            n_f = sum(isinstance(dev.get(k), torch.Tensor) for k in mdist.ROUTED_KEYS)
            router = mdist.PeerRouter(n_f, n_all, dev_t)
        except Exception as exc:  # noqa: BLE001
            if rank == 0:
                print(f"PeerRouter unavailable ({exc}); routing with all-to-all", file=sys.stderr)
        ok = torch.tensor([1.0 if router is not None else 0.0], device=dev_t)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok) == 0.0:
            router = None

    def routed_inputs(inbox_free=False):
        _, _, sm_range, _ = eng.smoothing_setup(dev["sm_length"], ctx.table)
        if router is not None:
            return router.route(eng, dev, sm_range, bounds, inbox_free=inbox_free)
        return mdist.route_particles(dev, sm_range, bounds)

    def step():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mine = routed_inputs(inbox_free=peer is not None)  # (the last step ended with PeerCube.end's barrier)
        e1.record()
        if peer is not None:
            peer.begin()
        else:
            slab.zero_()
        out = None
        if x_hi > x_lo:
            out = pipeline.run_hot_path(eng, share, dev=mine, cube=slab, x_lo=x_lo, x_hi=x_hi, zeroed=True, ctx=ctx)
        if peer is not None:
            peer.end()
        else:
            mdist.gather_slabs(slab, bounds, full, dst=0)
        torch.cuda.synchronize()
        route_ms.append(e0.elapsed_time(e1))
        return out

    timer.run(step, warmup)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    eng.set_timing(True)
    stage = {}
    route_ms.clear()

    def grab():
        for k, v in eng.last_timing_ms().items():
            stage.setdefault(k, []).append(v)

    ms_step, out = timer.run(step, steps, after=grab)
    eng.set_timing(False)
    route = float(np.mean(route_ms))
    ms_e2e = None
    if want_e2e and pinned is not None:
        host_cube = mdist.HostCube((nx, ny, nc), bounds)
        slab_e2e = torch.zeros((x_hi - x_lo, ny, nc), dtype=torch.float64, device=dev_t)
        copy_stream = torch.cuda.Stream(device=dev_t)

        def step_e2e():
            pipeline.upload(eng, share, pinned, out=dev)  # this rank's 1/N of the arrays, its own PCIe link
            mine = routed_inputs()
            if x_hi > x_lo:
                pipeline.run_hot_path_to_host(eng, share, host_cube.rows, mine, ctx, slab_e2e, x_lo=x_lo,
                                              x_hi=x_hi, copy_stream=copy_stream)

        timer.run(step_e2e, 1)
        ms_e2e, _ = timer.run(step_e2e, steps)
        host_cube.close()
        del slab_e2e
    clocks = sampler.stop() if sampler else None
    if out is None:  # a rank without rows
        from types import SimpleNamespace

        out = {"plan": SimpleNamespace(updates_dense=0, n_pairs=0, n_pairs2=0, n_kept=0, route2=0), "launches": 0}
        stage = {k: [0.0] for k in eng.STAGES}
    plan = out["plan"]
    stats = torch.tensor([float(plan.updates_dense), float(plan.n_pairs + plan.n_pairs2), float(plan.n_kept),
                          float(np.mean(stage["project"]) + np.mean(stage["stream2"])), route], dtype=torch.float64,
                         device=dev_t)
    per_rank = [torch.zeros_like(stats) for _ in range(world)]
    dist.all_gather(per_rank, stats)
    u_dense = float(sum(p[0] for p in per_rank))
    eng.set_count_exec(True)
    step()
    torch.cuda.synchronize()
    ex = eng.last_exec_counts()
    eng.set_count_exec(False)
    blk = {
        "value": u_dense / (ms_step * 1e-3), "unit": UNIT, "ms_per_step": ms_step, "steps": steps,
        "updates_per_step": u_dense, "insertion_wall_ms": ms_step, "slab_bounds": bounds,
        "partition": f"{world} work-balanced x-slabs of one cube; every rank holds 1/{world} of the particle list "
                     "and routes it by slab (halo particles go to both neighbours): " + ("bucketing kernel that stores straight into the destination ranks' inboxes over NVLink (dist.PeerRouter), " if router is not None else "one all-to-all per quantity over NCCL, ")
                     + ("slabs stored into rank 0's cube over NVLink by the projection kernels' own stores "
                        "(symmetric memory)" if peer is not None else "NCCL gather to rank 0"),
        "per_rank": {"pairs": [float(p[1]) for p in per_rank], "kept": [float(p[2]) for p in per_rank],
                     "kernels_ms": [float(p[3]) for p in per_rank], "route_ms": [float(p[4]) for p in per_rank]},
        "gpu_launches_per_step": int(out["launches"]),
        "roofline": roofline_block(eng, name, plan, ex, stage, (x_hi - x_lo) * ny * nc, world),
    }
    if ms_e2e is not None:
        blk["e2e"] = {"value": u_dense / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                      "h2d_bytes_per_step": pipeline.h2d_bytes(case),
                      "d2h_bytes_per_step": int(nx * ny * nc * 8),
                      "result": "host cube shared by the ranks (POSIX shared memory, page-locked): each rank "
                                f"uploads its 1/{world} of the particle arrays from pinned host memory and copies "
                                "its own slab back; byte counts are the aggregates over the ranks"}
    if clocks is not None:
        blk["clocks"] = clocks
    cube = peer.buf if (peer is not None and rank == 0) else full
    del router
    return blk, cube, case, dev, peer


def cfg5_block(eng, args, timer, world, rank, local):
    """North-star target: config 5 on all GPUs of the box, strong-scaled slabs, fused assembly;
    rank 0 checks seeded pixel columns against the reference-structured oracle."""
    import torch

    from martini_b200 import synthetic

    t0 = time.perf_counter()
    case, dev = synthetic.make_case_device("cfg5", eng.device, n=args.cfg5_particles)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    blk, cube, case, dev, peer = measure_strong(eng, "cfg5", args, timer, world, rank, local, case=case, dev=dev,
                                                steps=3, warmup=2, want_e2e=False)
    blk["config"] = workload_config("cfg5")
    blk["generate_s"] = t_gen
    del dev
    torch.cuda.empty_cache()
    if rank == 0:
        from tests.parity import PixelOracle

        nx, ny, nc = case["shape"]
        rng = np.random.Generator(np.random.PCG64(55))
        n_pix = 64
        pix = [(int(rng.integers(0, nx)), int(rng.integers(0, ny))) for _ in range(n_pix // 2)]
        cx = [int((i + 0.5) * nx / 8) for i in range(8)]
        pix += [(int(np.clip(cx[int(rng.integers(0, 8))] + rng.integers(-40, 40), 0, nx - 1)),
                 int(np.clip(cx[int(rng.integers(0, 8))] + rng.integers(-40, 40), 0, ny - 1)))
                for _ in range(n_pix - len(pix))]
        t1 = time.perf_counter()
        # the oracle only needs the particles that can reach a sampled column: select them on the
        # device (same seed: same particles; an ascending index subset keeps the summation order)
        _, dev_all = synthetic.make_case_device("cfg5", eng.device, n=args.cfg5_particles)
        from martini_b200 import pipeline

        _, _, sm_range, _ = eng.smoothing_setup(dev_all["sm_length"], pipeline.prepare(case).table)
        near = torch.zeros_like(dev_all["px"], dtype=torch.bool)
        for i, j in pix:
            near |= ((dev_all["px"] - i).abs() <= sm_range + 1) & ((dev_all["py"] - j).abs() <= sm_range + 1)
        idx = torch.nonzero(near).flatten()
        host = {k: dev_all[k][idx].cpu().numpy() for k in ("px", "py", "pz", "sm_length", "v", "mHI", "D")}
        del dev_all, near, sm_range
        torch.cuda.empty_cache()
        hcase = dict(case, **host)
        ref = PixelOracle(hcase).pixels(pix)
        got = np.array([cube[i, j].cpu().numpy() for i, j in pix])
        peak = float(cube.abs().max())
        blk["parity"] = {"pixels": len(pix), "max_abs_diff_over_peak": float(np.abs(got - ref).max() / peak),
                         "ref_max_over_peak": float(np.abs(ref).max() / peak),
                         "oracle_s": time.perf_counter() - t1, "tolerance": 1e-6,
                         "ok": bool(np.abs(got - ref).max() <= 1e-6 * peak)}
    return blk


def run_b200(args):
    import torch
    import torch.distributed as dist

    from martini_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    if args.gpus != world and rank == 0 and world > 1:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}", file=sys.stderr)
    eng = Engine(f"cuda:{local}")
    timer = Timer(eng.device, world)
    common = {"metric": METRIC, "unit": UNIT, "n_gpus": world, "warmup": max(args.warmup, 3),
              "higher_is_better": True, "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
    if world == 1:
        blk = measure_single(eng, args.workload, args, args.steps, args.warmup, timer, clocks_for=local)
        line = dict(common, scaling="strong", **blk)  # (N > 1 runs the same cube: strong scaling)
        line["gpu_launches"] = blk["gpu_launches_per_step"] * args.steps
        others = [w for w in args.others.split(",") if w and w != "none" and w != args.workload]
        if args.particles is None and others:
            line["other_configs"] = {}
            for w in others:
                line["other_configs"][w] = measure_single(eng, w, args, args.other_steps, 3, timer)
        if args.cpu_full:
            line["cpu_full_cfg2"] = cpu_full_run("cfg2")
        # key order: the contract's keys first
        print(json.dumps(line))
        return
    blk, cube, case, dev, peer = measure_strong(eng, args.workload, args, timer, world, rank, local)
    line = dict(common, scaling="strong", config=workload_config(args.workload, case, args.particles), **blk)
    line["gpu_launches"] = blk["gpu_launches_per_step"] * args.steps
    del cube, dev, peer
    torch.cuda.empty_cache()
    if (world == 8 or args.force_cfg5) and not args.no_cfg5 and args.particles is None:
        try:
            line["extra"] = {"cfg5": cfg5_block(eng, args, timer, world, rank, local)}
        except Exception as exc:  # noqa: BLE001 -- the headline line must survive
            line["extra"] = {"cfg5": {"error": repr(exc)[:300]}}
    if rank == 0:
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
