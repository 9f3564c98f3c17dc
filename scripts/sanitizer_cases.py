#!/usr/bin/env python
"""Three small insertions (brick, column and splat kernels; several sort tiles) for
compute-sanitizer:   compute-sanitizer --tool racecheck python scripts/sanitizer_cases.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from martini_b200 import synthetic
from martini_b200.engine import Engine
from martini_b200.pipeline import run_hot_path

eng = Engine("cuda:0")
cases = [synthetic.make_case("cfg2", n=12000, nx=48, ny=40, nc=96),
         synthetic.make_case("cfg3", n=20000, nx=48, ny=48, nc=64),
         synthetic.make_case("cfg4", n=300, nx=48, ny=48, nc=32)]
for c in cases:
    out = run_hot_path(eng, c)
    torch.cuda.synchronize()
    cube = out["cube"]
    print(c["name"], "pairs", out["plan"].n_pairs, out["plan"].n_pairs2, "sum %.6e" % float(cube.sum()),
          "finite", bool(torch.isfinite(cube).all()))
