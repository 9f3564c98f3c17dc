"""Kernel variants (compile-time switches of csrc/: MTN_NCW = 4 gives 8-warp CTAs with 16
accumulators per thread instead of 4-warp CTAs with 32, measurements in profiles/README.md)
under the SIMT emulator: each must reproduce the oracle and -- where the variant only moves
work between threads -- the default kernel's cube bit for bit, under every thread schedule.
Test infrastructure; see tests/emu/__init__.py.
"""

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from martini_b200.pipeline import run_hot_path  # noqa: E402
from tests import test_emu_parity as T  # noqa: E402
from tests.emu import EmuEngine  # noqa: E402

#: name -> (switches, bit-identical to the default kernel?)
VARIANTS = {
    # (not bit-identical: the work-item chunk length follows the resident CTA count, so crowded
    # bricks are cut -- and their partial sums associated -- differently)
    "ncw4": (("MTN_NCW=4",), False),
}
FAST_CASES = ("cfg2_odd_shape", "cfg2_one_channel_block_partial", "cfg3_thermal", "cfg4_wide_dirac",
              "adaptive_gauss", "increasing_edges", "dirac_edges", "crowded_bricks")


@pytest.fixture(scope="module")
def base():
    return EmuEngine()


@pytest.fixture(scope="module", params=sorted(VARIANTS))
def variant(request):
    defines, exact = VARIANTS[request.param]
    eng = EmuEngine(defines)
    yield eng, exact
    assert eng.violations() == 0


@pytest.mark.parametrize("name", FAST_CASES)
def test_variant_vs_oracle_and_default(base, variant, name):
    eng, exact = variant
    case = T.CASES[name]
    out = T.run_and_check(eng, case)
    ref = run_hot_path(base, case)
    if exact:
        assert torch.equal(out["cube"], ref["cube"])
    assert out["plan"].n_pairs == ref["plan"].n_pairs


@pytest.mark.parametrize("mode", ("reverse", "shuffle"))
def test_variant_schedule_independent(variant, mode):
    eng, _ = variant
    for name in ("cfg2_odd_shape", "crowded_bricks"):
        case = T.CASES[name]
        eng.set_schedule("forward")
        want = run_hot_path(eng, case)["cube"]
        try:
            eng.set_schedule(mode, seed=7)
            got = run_hot_path(eng, case)["cube"]
        finally:
            eng.set_schedule("forward")
        assert torch.equal(got, want)


def test_variant_accumulate_and_slabs(variant):
    eng, _ = variant
    case = T.CASES["cfg2_odd_shape"]
    rng = np.random.Generator(np.random.PCG64(99))
    cube0 = rng.normal(0.0, 1e-6, case["shape"])
    out = run_hot_path(eng, case, cube=eng.to_device(cube0.copy()))
    T.check_cube(out["cube"].numpy(), T.oracle_hot_path(case, cube0=cube0)["cube"])
    full = run_hot_path(eng, case)["cube"]
    nx = case["shape"][0]
    parts = [run_hot_path(eng, case, x_lo=a, x_hi=b)["cube"] for a, b in ((0, 11), (11, nx))]
    T.G.assert_same_cube(torch.cat(parts, dim=0), full)
