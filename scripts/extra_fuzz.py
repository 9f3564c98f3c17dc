"""Offline fuzz sweep beyond the seeds of tests/test_emu_fuzz.py (CPU, SIMT emulator against the
oracle):  python scripts/extra_fuzz.py 280 1080"""
import sys, numpy as np
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from tests.fuzz_cases import random_case
from tests.emu import EmuEngine
from tests.parity import oracle_hot_path
from martini_b200.pipeline import run_hot_path
eng = EmuEngine(); eng.set_schedule("forward")
lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = 0; worst = 0.0
for seed in range(lo, hi):
    case, extras = random_case(seed)
    nx = case["shape"][0]
    ref = oracle_hot_path(case, cube0=extras["prefill"])
    x_lo, x_hi = extras["slab"] or (0, nx)
    x_lo, x_hi = int(x_lo), int(x_hi)
    cube0 = None
    if extras["prefill"] is not None:
        cube0 = eng.to_device(np.ascontiguousarray(extras["prefill"][x_lo:x_hi]))
    try:
        out = run_hot_path(eng, case, cube=cube0, x_lo=x_lo, x_hi=x_hi)
    except Exception as e:
        print("seed", seed, "EXC", repr(e)[:200]); bad += 1; continue
    ok = np.array_equal(out["accept"].numpy().astype(bool), ref["accept"])
    got, want = out["cube"].numpy(), ref["cube"][x_lo:x_hi]
    peak = np.abs(ref["cube"]).max()
    err = np.abs(got - want).max() / peak if peak > 0 else np.abs(got - want).max()
    tight = 1e-10 if "WendlandC6" in case["kernel"][0] else 1e-11
    if (x_lo, x_hi) == (0, nx): ok &= out["plan"].updates_dense == ref["updates"]
    if not ok or not (err <= tight):
        print("seed", seed, "FAIL", ok, err, case["kernel"], case["spectrum"], case["shape"]); bad += 1
    worst = max(worst, float(err))
print("range", lo, hi, "bad", bad, "worst rel err", worst, "violations", eng.violations())
