"""Corner cases of the kernels rewritten in the second half of round 2, under the SIMT emulator
(CPU; the GPU twins are in tests/test_gpu_round2b.py):

* column kernel -- the live edges of a batch's particles are enumerated as one run: crowded
  pixels (several batches and several work items per pixel), windows of one to a few edges
  (many particles share a 31-lane step and add to the same channels, in index order), windows
  longer than a step, more than one channel superblock;
* radix sort -- one-pass keys, 9- and 10-bit digits (the widest), pair counts that are not a
  multiple of the block tile;
* plan_count -- candidate boxes above 2^16 pixels (the three-limb warp reduction of U_dense).
"""

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from martini_b200 import synthetic  # noqa: E402
from martini_b200.pipeline import run_hot_path  # noqa: E402
from tests.emu import EmuEngine  # noqa: E402
from tests.parity import check_cube, oracle_hot_path  # noqa: E402


@pytest.fixture(scope="module")
def eng():
    e = EmuEngine()
    e.set_schedule("forward")
    yield e
    assert e.violations() == 0


def run_and_check(eng, case):
    out = run_hot_path(eng, case)
    ref = oracle_hot_path(case)
    assert np.array_equal(out["accept"].numpy().astype(bool), ref["accept"])
    assert out["plan"].updates_dense == ref["updates"]
    assert np.abs(ref["cube"]).max() > 0
    check_cube(out["cube"].numpy(), ref["cube"])
    return out


def column_case(n, nx, ny, nc, seed, sigma_lo, sigma_hi, n_pixels=None, dv=4.0):
    """cfg3-shaped case whose particles are all smaller than a pixel (DiracDelta fallback of the
    adaptive kernel: the column route), optionally crowded into `n_pixels` pixels, with line
    widths drawn from [sigma_lo, sigma_hi] km/s."""
    case = synthetic.make_case("cfg3", n=n, nx=nx, ny=ny, nc=nc, seed=seed)
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    case["sm_length"] = np.full(n, 0.2)
    if n_pixels is not None:
        cx = rng.integers(2, nx - 2, n_pixels)
        cy = rng.integers(2, ny - 2, n_pixels)
        k = rng.integers(0, n_pixels, n)
        case["px"] = cx[k] + rng.uniform(-0.45, 0.45, n)
        case["py"] = cy[k] + rng.uniform(-0.45, 0.45, n)
    case["sigma"] = rng.uniform(sigma_lo, sigma_hi, n)
    edges = case["edges"]
    case["v"] = rng.uniform(min(edges[0], edges[-1]) - 10.0, max(edges[0], edges[-1]) + 10.0, n)
    case.pop("T", None)
    set_pz(case)
    return case


def set_pz(case):
    """Channel pixel coordinate of the line centre (what _prune_particles tests), as
    synthetic._finish computes it."""
    e = case["edges"]
    case["pz"] = np.ascontiguousarray((e[0] - case["v"]) / (e[0] - e[1]) - 0.5)


COLUMN_CASES = {
    # ~100 particles per pixel: four batches of 32 per key, windows of ~40 edges (two steps each)
    "crowded_pixels": lambda: column_case(2400, 12, 12, 96, 31, 6.0, 9.0, n_pixels=24),
    # windows of one to three edges: ten and more particles in one 31-lane step
    "narrow_lines": lambda: column_case(3000, 10, 10, 48, 32, 0.05, 0.4, n_pixels=12),
    # a mix, lines cut by both ends of the band
    "mixed_widths": lambda: column_case(3000, 16, 14, 40, 33, 0.1, 30.0, n_pixels=40),
    # more than one channel superblock (1024 channels each), windows that straddle the boundary
    "two_superblocks": lambda: column_case(600, 6, 6, 1100, 34, 2.0, 60.0, n_pixels=8, dv=1.0),
}


@pytest.mark.parametrize("name", sorted(COLUMN_CASES))
def test_column_stream(eng, name):
    case = COLUMN_CASES[name]()
    out = run_and_check(eng, case)
    assert out["plan"].n_pairs2 > 0.9 * out["plan"].n_kept  # it was the column kernel that ran


def test_column_stream_is_deterministic_under_any_schedule(eng):
    case = COLUMN_CASES["narrow_lines"]()
    a = run_hot_path(eng, case)["cube"].clone()
    e2 = EmuEngine()
    e2.set_schedule("reverse")
    b = run_hot_path(e2, case)["cube"]
    assert torch.equal(a, b)
    assert e2.violations() == 0


def test_sort_with_nine_bit_digits(eng):
    """160 000 pixel keys = 18 bits: two passes of 9 bits; 20 000 pairs: 2.4 block tiles."""
    case = column_case(20000, 400, 400, 8, 41, 2.0, 6.0)
    out = run_and_check(eng, case)
    assert out["plan"].n_pairs2 > 16384


def test_sort_with_ten_bit_digits(eng):
    """(tile, channel) keys of the splat stream: 32 x 32 tiles x 600 channels = 20 bits, two
    passes of 10 bits."""
    case = synthetic.make_case("cfg4", n=150, nx=256, ny=256, nc=600, seed=42)
    case["sm_length"] = case["sm_length"] * 0.25  # (2 - 10 px: a few tiles per particle)
    out = run_and_check(eng, case)
    assert out["plan"].n_pairs2 > 0


def test_sort_single_pass(eng):
    """A cube of two tiles and one channel block: 6 brick keys, one 3-bit pass."""
    case = synthetic.make_case("cfg2", n=1500, nx=8, ny=16, nc=20, seed=43)
    run_and_check(eng, case)


def test_updates_dense_with_boxes_above_2_to_16_pixels(eng):
    """U_dense is reduced per warp in three 16-bit limbs of the box area: boxes of 300 x 300
    pixels carry into the second limb, sums of 32 of them into the third."""
    case = synthetic.make_case("cfg2", n=96, nx=300, ny=300, nc=4, seed=44)
    case["sm_length"] = np.full(96, 90.0)
    case["px"] = np.full(96, 150.0) + np.linspace(-3, 3, 96)
    case["py"] = np.full(96, 150.0) - np.linspace(-3, 3, 96)
    case["v"] = np.full(96, float(np.mean(case["edges"])))
    set_pz(case)
    out = run_hot_path(eng, case)
    ref = oracle_hot_path(case)
    assert ref["updates"] > 96 * 65536 * 4
    assert out["plan"].updates_dense == ref["updates"]
    check_cube(out["cube"].numpy(), ref["cube"])
