"""tests/parity.PixelOracle (the pre-filtered, blocked oracle used for full-size configs) against
the plain reference-structured oracle: the SAME bits, pixel by pixel."""

import numpy as np
import pytest

from martini_b200 import synthetic
from tests.parity import PixelOracle, oracle_hot_path, oracle_pixels

CASES = {
    "cfg2": dict(n=20000, nx=64, ny=64, nc=32),
    "cfg3": dict(n=30000, nx=64, ny=64, nc=32),
    "cfg4": dict(n=3000, nx=64, ny=64, nc=16),
    "demo": dict(),
}


def make(name):
    case = synthetic.make_case(name, **CASES[name])
    if name == "cfg4":
        case["sm_length"] = case["sm_length"] * 0.3
    return case


@pytest.mark.parametrize("name", sorted(CASES))
def test_prefiltered_selection_is_bit_identical(name):
    case = make(name)
    nx, ny, _ = case["shape"]
    po = PixelOracle(case)
    rng = np.random.default_rng(1)
    pix = [(int(a), int(b)) for a, b in zip(rng.integers(0, nx, 48), rng.integers(0, ny, 48))]
    pix += [(0, 0), (nx - 1, ny - 1), (nx // 2, ny // 2)]
    want = oracle_pixels(case, pix)
    assert np.array_equal(po.pixels(pix, threads=1), want)
    assert np.array_equal(po.pixels(pix, threads=4), want)
    # blocked evaluation with the running sum carried in as row 0: same sequential sum
    assert np.array_equal(np.array([po.pixel(ij, block=7) for ij in pix]), want)
    assert np.array_equal(np.array([po.pixel(ij, block=1) for ij in pix[:8]]), want[:8])


@pytest.mark.parametrize("name", sorted(CASES))
def test_total_flux_identity(name):
    case = make(name)
    ref = oracle_hot_path(case)["cube"]
    po = PixelOracle(case)
    for chunk, threads in ((5000, 1), (700, 3)):
        tf = po.total_flux(chunk=chunk, threads=threads)
        assert abs(tf - ref.sum()) <= 1e-12 * abs(ref.sum())


def test_infinite_range_particles_are_candidates_everywhere():
    """sm_range = inf (GlobalProfile's DiracDeltaKernel(size_in_fwhm=inf)) bypasses the hash."""
    case = synthetic.make_case("cfg2", n=500, nx=4, ny=4, nc=16)
    case["kernel"] = ("DiracDeltaKernel", {"size_in_fwhm": np.inf})
    pix = [(i, j) for i in range(4) for j in range(4)]
    assert np.array_equal(PixelOracle(case).pixels(pix, threads=1), oracle_pixels(case, pix))
