#!/usr/bin/env python
"""Time mtn_convolve_beam on BASELINE-sized cubes: CUDA events, L2 flushed between
repetitions, FP64 FMA fraction against the peak measured in the same process.

    python scripts/bench_convolve.py [--out gpurun_out/r2_convolve.json]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from martini_b200 import DataCube, GaussianBeam  # noqa: E402
from martini_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
a = ap.parse_args()
eng = Engine("cuda:0")
p64 = eng.fp64_peak_tflops()
flush = torch.empty(512 << 20, dtype=torch.uint8, device=eng.device)
res = []
for (nx, ny, nc), bmaj, px in (((154, 154, 32), 30.0, 10.0), ((538, 538, 256), 30.0, 10.0), ((2074, 2074, 512), 30.0, 10.0),
                               ((538, 538, 256), 60.0, 5.0)):
    dc = DataCube(n_px_x=8, n_px_y=8, n_channels=4, px_size=px, channel_width=4.0)
    beam = GaussianBeam(bmaj=bmaj, bmin=bmaj, bpa=0.0, truncate=4.0)
    beam.init_kernel(dc)
    k = eng.to_device(np.ascontiguousarray(beam.kernel))
    cube = torch.rand((nx, ny, nc), dtype=torch.float64, device=eng.device)
    times = []
    for _ in range(4):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = eng.convolve_beam(cube, k, scale=beam.area)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
        del out
    ms = float(np.min(times[1:]))
    taps = int(beam.kernel.size)
    fma = float(nx) * ny * nc * taps  # (interior count: border voxels see fewer taps)
    res.append({"cube": [nx, ny, nc], "beam_taps": list(beam.kernel.shape), "ms": ms,
                "tflops": 2 * fma / (ms * 1e-3) / 1e12, "fp64_frac": 2 * fma / (ms * 1e-3) / 1e12 / p64,
                "voxels_per_s": nx * ny * nc / (ms * 1e-3)})
    del cube
    torch.cuda.empty_cache()
rec = {"kernel": "convolve_beam_kernel", "fp64_peak_tflops": p64, "cases": res}
print(json.dumps(rec))
if a.out:
    json.dump(rec, open(a.out, "w"), indent=1)
