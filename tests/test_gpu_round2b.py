"""GPU twins of tests/test_emu_round2b.py: the corner cases of the column kernel (edges of a
batch as one run), the radix sort (one pass, 9- and 10-bit digits) and plan_count's U_dense
reduction, through the C ABI on a B200 against the CPU oracle."""

import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu

from martini_b200 import synthetic  # noqa: E402
from martini_b200.engine import Engine  # noqa: E402
from martini_b200.pipeline import run_hot_path  # noqa: E402
from tests.parity import check_cube, oracle_hot_path  # noqa: E402
from tests.test_emu_round2b import COLUMN_CASES, column_case, set_pz  # noqa: E402


@pytest.fixture(scope="module")
def eng():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return Engine("cuda:0")


def run_and_check(eng, case):
    out = run_hot_path(eng, case)
    ref = oracle_hot_path(case)
    assert np.array_equal(out["accept"].cpu().numpy().astype(bool), ref["accept"])
    assert out["plan"].updates_dense == ref["updates"]
    assert np.abs(ref["cube"]).max() > 0
    check_cube(out["cube"].cpu().numpy(), ref["cube"])
    return out


@pytest.mark.parametrize("name", sorted(COLUMN_CASES))
def test_column_stream(eng, name):
    out = run_and_check(eng, COLUMN_CASES[name]())
    assert out["plan"].n_pairs2 > 0.9 * out["plan"].n_kept


def test_column_stream_is_bit_reproducible(eng):
    case = COLUMN_CASES["narrow_lines"]()
    a = run_hot_path(eng, case)["cube"].clone()
    for _ in range(3):
        assert torch.equal(a, run_hot_path(eng, case)["cube"])


def test_sort_with_nine_bit_digits(eng):
    out = run_and_check(eng, column_case(20000, 400, 400, 8, 41, 2.0, 6.0))
    assert out["plan"].n_pairs2 > 16384


def test_sort_with_ten_bit_digits(eng):
    case = synthetic.make_case("cfg4", n=150, nx=256, ny=256, nc=600, seed=42)
    case["sm_length"] = case["sm_length"] * 0.25
    run_and_check(eng, case)


def test_sort_single_pass(eng):
    run_and_check(eng, synthetic.make_case("cfg2", n=1500, nx=8, ny=16, nc=20, seed=43))


def test_updates_dense_with_boxes_above_2_to_16_pixels(eng):
    case = synthetic.make_case("cfg2", n=96, nx=300, ny=300, nc=4, seed=44)
    case["sm_length"] = np.full(96, 90.0)
    case["px"] = np.full(96, 150.0) + np.linspace(-3, 3, 96)
    case["py"] = np.full(96, 150.0) - np.linspace(-3, 3, 96)
    case["v"] = np.full(96, float(np.mean(case["edges"])))
    set_pz(case)
    out = run_hot_path(eng, case)
    ref = oracle_hot_path(case)
    assert out["plan"].updates_dense == ref["updates"]
    check_cube(out["cube"].cpu().numpy(), ref["cube"])
