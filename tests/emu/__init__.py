"""SIMT emulator harness for the CPU test suite -- TEST INFRASTRUCTURE, never part of the product.

``build()`` compiles martini_b200/csrc/api.cu -- the very sources nvcc turns into
``libmartini_b200.so`` -- as plain C++ with g++ against ``cuda_emu.h`` / ``cuda_emu.cpp`` (CUDA
threads as fibers, blocks one after the other, host memory as device memory) into
``tests/emu/libmartini_emu.so``.  ``EmuEngine`` is ``martini_b200.engine.Engine`` bound to that
library with CPU tensors as "device" buffers, so the parity tests can drive every kernel of
the hot path through the same C ABI and host code without a GPU.

What this checks: kernel logic (indexing, predicates, barrier placement, arithmetic).  What it
cannot: performance, hardware memory ordering, the PTX helpers (mbarrier / cp.async.bulk have
emulated stand-ins) and the opt-in warp-specialised kernel.  The GPU parity tests
(``-m gpu``) remain the proof; the product never loads this library
(``tests/test_abi.py::test_product_never_imports_oracle_or_emulator``).
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import torch

from martini_b200 import _lib as L
from martini_b200.engine import Engine

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "martini_b200", "csrc")
LIB = os.path.join(HERE, "libmartini_emu.so")


#: MTN_EMU_SANITIZE=address (or "address,undefined"): build the emulated library with the
#: compiler's sanitizers, so that a kernel reading or writing outside a caller's buffer (inputs,
#: cube, accept mask ...) or outside a static __shared__ array aborts the test run.  Run as
#:   LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 \
#:   MTN_EMU_SANITIZE=address python -m pytest tests/test_emu_parity.py tests/test_emu_fuzz.py
SANITIZE = os.environ.get("MTN_EMU_SANITIZE", "")


def lib_path(defines=()) -> str:
    """One emulated library per set of -D switches (experimental kernel variants)."""
    tag = "_".join(d.replace("=", "").replace("MTN_", "").lower() for d in sorted(defines))
    if SANITIZE:
        tag = "_".join(t for t in (tag, "san" + SANITIZE.replace(",", "")) if t)
    if not tag:
        return LIB
    return os.path.join(HERE, f"libmartini_emu_{tag}.so")

CUDA_INCLUDE = os.environ.get("CUDA_INCLUDE", "/usr/local/cuda/include")

_libs = {}


def _stale(lib) -> bool:
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    srcs += [os.path.join(HERE, f) for f in ("cuda_emu.h", "cuda_emu.cpp")]
    srcs.append(os.path.join(ROOT, "include", "martini_b200.h"))
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force: bool = False, defines=()) -> str:
    """g++ build of the emulated library; -ffp-contract=off keeps the explicitly rounded
    predicates (``__dsub_rn`` ...) honest."""
    lib = lib_path(defines)
    if force or _stale(lib):
        cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-ffp-contract=off",
               "-Wl,-Bsymbolic",  # our cuda* stand-ins, not the libcudart torch has loaded
               *[f"-D{d}" for d in defines],
               *([f"-fsanitize={SANITIZE}", "-fno-omit-frame-pointer"] if SANITIZE else []),
               "-include", os.path.join(HERE, "cuda_emu.h"), f"-I{CUDA_INCLUDE}", f"-I{HERE}",
               "-x", "c++", os.path.join(CSRC, "api.cu"), os.path.join(HERE, "cuda_emu.cpp"),
               "-o", lib]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("emulator build failed:\n" + res.stdout + res.stderr)
    return lib


def load(defines=()):
    key = tuple(sorted(defines))
    if key not in _libs:
        lib = C.CDLL(build(defines=key))
        for name, (restype, argtypes) in L.SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        lib.mtn_emu_violations.restype = C.c_int
        lib.mtn_emu_set_schedule.argtypes = [C.c_int, C.c_ulonglong]
        lib.mtn_emu_set_schedule.restype = None
        _libs[key] = lib
    return _libs[key]


class EmuEngine(Engine):
    """``Engine`` over the emulated library: same host code, CPU tensors, no stream."""

    def __init__(self, defines=()):  # (deliberately does not call Engine.__init__: no CUDA here)
        self.lib = load(defines)
        self.device = torch.device("cpu")
        self._scratch = None
        self._workspace = None
        self.last_plan = None
        self.last_launches = 0

    def _stream(self):
        return None

    SCHEDULES = {"forward": 0, "reverse": 1, "shuffle": 2}

    def set_schedule(self, mode: str = "forward", seed: int = 0):
        """Order in which a block's runnable threads take their turns; results must not
        depend on it."""
        self.lib.mtn_emu_set_schedule(self.SCHEDULES[mode], seed)

    def violations(self) -> int:
        """Collectives that were entered with already-exited lanes in their mask."""
        return int(self.lib.mtn_emu_violations())
