"""Seeded random cases for the fuzz parity tests (tests/test_emu_fuzz.py on the emulator,
tests/test_gpu_fuzz.py on the GPU): every kernel (primitive and adaptive), both spectra, scalar
and per-particle line widths, either channel direction, odd cube shapes, slabs, pre-filled
cubes, particles on pixel centres / pixel edges / channel edges, NaN coordinates, zero masses,
sub-pixel and cube-sized smoothing lengths.  Test infrastructure."""

import numpy as np

KERNELS = (
    ("_WendlandC2Kernel", {}), ("_WendlandC6Kernel", {}), ("_CubicSplineKernel", {}),
    ("_QuarticSplineKernel", {}), ("_GaussianKernel", {"truncate": 2.5}),
    ("_GaussianKernel", {"truncate": 4.0}), ("_GaussianKernel", {"truncate": 6.0}),
    ("DiracDeltaKernel", {}),
    ("WendlandC2Kernel", {}), ("WendlandC6Kernel", {}), ("CubicSplineKernel", {}),
    ("QuarticSplineKernel", {}), ("GaussianKernel", {"truncate": 3.0}), ("GaussianKernel", {"truncate": 5.0}),
)


def random_case(seed):
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    nx, ny = int(rng.integers(1, 41)), int(rng.integers(1, 41))
    nc = int(rng.choice([1, 2, 7, 31, 32, 33, 63, 64, 65, 100, 129, 200]))
    n = int(rng.integers(1, 700))
    kernel = KERNELS[seed % len(KERNELS)]
    dv = float(rng.choice([0.5, 1.0, 4.0, 10.0, 20.0]))
    edges = dv * (nc / 2.0 - np.arange(nc + 1)) + float(rng.uniform(-50, 50))
    increasing = bool(rng.integers(0, 2))
    if increasing:
        edges = edges[::-1].copy()
    px = rng.uniform(-4.0, nx + 4.0, n)
    py = rng.uniform(-4.0, ny + 4.0, n)
    # a share of the particles exactly on pixel centres, pixel edges and far outside
    snap = rng.random(n)
    px = np.where(snap < 0.15, np.round(px), np.where(snap < 0.3, np.round(px) + 0.5, px))
    py = np.where(snap < 0.15, np.round(py), np.where((snap >= 0.2) & (snap < 0.35), np.round(py) - 0.5, py))
    px[rng.random(n) < 0.03] += 1000.0
    med = float(rng.choice([0.2, 0.6, 1.2, 2.5, 6.0]))
    sm = np.clip(rng.lognormal(np.log(med), 0.6, n), 0.05, 60.0)
    if kernel[0] == "DiracDeltaKernel":
        sm = sm * 0.1
    lo, hi = edges.min(), edges.max()
    v = rng.uniform(lo - 3 * dv, hi + 3 * dv, n)
    on_edge = rng.random(n) < 0.1
    v[on_edge] = rng.choice(edges, int(on_edge.sum()))
    spectrum = "diracdelta" if seed % 3 == 2 else "gaussian"
    if spectrum == "gaussian":
        sigma = float(rng.uniform(0.3, 30.0)) if seed % 2 else rng.uniform(0.3, 30.0, n)
    else:
        sigma = 0.0
    mHI = rng.uniform(0.5, 2.0, n) * 1.0e5
    mHI[rng.random(n) < 0.05] = 0.0
    case = {
        "name": f"fuzz{seed}", "px": px, "py": py, "sm_length": sm, "v": v, "sigma": sigma, "mHI": mHI,
        "D": rng.uniform(1.0, 30.0, n), "edges": edges, "shape": (nx, ny, nc),
        "px_size": float(rng.uniform(0.5, 20.0)), "kernel": kernel, "spectrum": spectrum,
    }
    # channel pixel coordinate of v (what pruning sees), as the front-end computes it
    case["pz"] = (v - edges[0]) / (edges[1] - edges[0]) - 0.5
    bad = rng.random(n) < 0.02
    case["px"] = np.where(bad, np.nan, case["px"])
    for k in ("px", "py", "pz", "sm_length", "v", "mHI", "D"):
        case[k] = np.ascontiguousarray(case[k], dtype=np.float64)
    extras = {
        "slab": None if rng.random() < 0.5 or nx < 2 else tuple(sorted(rng.choice(nx + 1, 2, replace=False))),
        "prefill": rng.normal(0.0, 1e-9, (nx, ny, nc)) if rng.random() < 0.4 else None,
    }
    return case, extras
