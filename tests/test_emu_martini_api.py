"""The drop-in ``Martini`` class over the SIMT emulator, on the CPU: the class-level tests of
tests/test_gpu_martini_api.py (the reference's own behavioural tests for the hot path) re-run
with ``EmuEngine`` in place of ``Engine`` -- same host classes, same C ABI, the csrc/ sources
compiled as host code.  A logic check for where there is no GPU; the ``-m gpu`` suite remains
the proof, and the product never loads the emulated library (tests/emu/__init__.py).
"""

import pytest

torch = pytest.importorskip("torch")

import martini_b200.martini as M  # noqa: E402
from tests import test_gpu_martini_api as A  # noqa: E402
from tests.emu import EmuEngine  # noqa: E402

_ENGINE = None


def _emu_engine(device=None):
    global _ENGINE
    if _ENGINE is None:
        _ENGINE = EmuEngine()
    return _ENGINE


@pytest.fixture(autouse=True)
def emulated_engine(monkeypatch):
    monkeypatch.setattr(M, "Engine", _emu_engine)
    yield
    assert _ENGINE is None or _ENGINE.violations() == 0


def _rerun(fn):
    """The test function without the module-level ``gpu`` mark (parametrisation is kept)."""
    marks = [m for m in getattr(fn, "pytestmark", []) if m.name != "gpu"]
    clone = type(fn)(fn.__code__, fn.__globals__, fn.__name__, fn.__defaults__, fn.__closure__)
    clone.__dict__.update({k: v for k, v in fn.__dict__.items() if k != "pytestmark"})
    clone.__kwdefaults__ = fn.__kwdefaults__
    clone.__doc__ = fn.__doc__
    if marks:
        clone.pytestmark = marks
    return clone


#: asserts CUDA residency of the cube -- meaningless on CPU tensors
_GPU_ONLY = {"test_cube_stays_on_the_device_between_steps"}

for _name in dir(A):
    if _name.startswith("test_") and _name not in _GPU_ONLY:
        globals()[_name] = _rerun(getattr(A, _name))
