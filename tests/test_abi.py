"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/martini_b200.h declares (no compute calls: there is no GPU here)."""

import ctypes
import os
import re

import pytest

import __graft_entry__ as entry
from martini_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    entry.build()
    return L.load()


def header_functions():
    src = open(os.path.join(ROOT, "include", "martini_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mtn_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(lib):
    names = header_functions()
    assert len(names) >= 14
    assert sorted(L.SYMBOLS) == names
    for n in names:
        assert hasattr(lib, n), n


def test_version_and_error_string(lib):
    assert lib.mtn_version() == 200
    assert isinstance(lib.mtn_last_error(), bytes)


def test_struct_layouts_match_header(tmp_path):
    """sizeof / offsetof of every struct as gcc lays out include/martini_b200.h against the
    ctypes mirrors in martini_b200/_lib.py."""
    import shutil
    import subprocess

    structs = {"MtnKernelEntry": L.MtnKernelEntry, "MtnKernelTable": L.MtnKernelTable,
               "MtnParticles": L.MtnParticles, "MtnCube": L.MtnCube, "MtnPlan": L.MtnPlan,
               "MtnFrontEnd": L.MtnFrontEnd}
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "martini_b200.h"', "int main(void) {"]
    for name, cls in structs.items():
        lines.append(f'printf("{name} %zu\\n", sizeof({name}));')
        for field, _ in cls._fields_:
            lines.append(f'printf("{name}.{field} %zu\\n", offsetof({name}, {field}));')
    lines += ["return 0; }"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(ln.split() for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for name, cls in structs.items():
        assert int(out[name]) == ctypes.sizeof(cls), name
        for field, _ in cls._fields_:
            assert int(out[f"{name}.{field}"]) == getattr(cls, field).offset, (name, field)


def test_invalid_arguments_are_reported_not_crashed(lib):
    # argument validation happens before any CUDA call, so this is safe without a GPU
    t = L.MtnKernelTable()
    t.n = 0
    rc = lib.mtn_smoothing_setup(0, None, ctypes.byref(t), None, None, None, None, None)
    assert rc == -1 and b"kernel table" in lib.mtn_last_error()
    with pytest.raises(L.MartiniB200Error, match="kernel table"):
        L.check(rc, "mtn_smoothing_setup")


def test_no_cpu_fallback_without_cuda():
    import torch

    from martini_b200.engine import Engine

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(L.MartiniB200Error, match="no CPU fallback"):
        Engine()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(L.MartiniB200Error, match="has not been built"):
        L.load()


def test_product_never_imports_oracle_or_emulator():
    """The product package must not reference oracle/ or the CPU test suite's SIMT emulator
    (tests/emu/): a CPU fallback would void parity.  The only trace of the emulator in the
    product sources is the MTN_HOST_EMU switch of the launch macro, defined by tests/emu alone."""
    pkg = os.path.join(ROOT, "martini_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+(oracle|tests)\b", text, flags=re.M), f
                assert "libmartini_emu" not in text and "cuda_emu" not in text.replace("tests/emu/cuda_emu.h", ""), f
                assert not re.search(r"#\s*define\s+MTN_HOST_EMU", text), f
    # and the build recipe of the product library never defines the switch
    entry = open(os.path.join(ROOT, "__graft_entry__.py")).read()
    assert "MTN_HOST_EMU" not in entry


def test_unsupported_plugins_raise():
    from martini_b200 import spectral_models as S
    from martini_b200 import sph_kernels as K

    class MyKernel(K._WendlandC2Kernel):
        pass

    class MySpectrum(S.GaussianSpectrum):
        pass

    with pytest.raises(NotImplementedError, match="MyKernel"):
        K.kernel_table(MyKernel())
    with pytest.raises(NotImplementedError, match="MyKernel"):
        K.kernel_table(K._AdaptiveKernel((MyKernel(), K.DiracDeltaKernel())))
    with pytest.raises(NotImplementedError, match="MySpectrum"):
        S.spectrum_kind(MySpectrum())
    assert K.kernel_table(K.GaussianKernel(truncate=4.0)).adaptive
