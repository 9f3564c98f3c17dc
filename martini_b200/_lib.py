"""ctypes binding of ``libmartini_b200.so`` (the C ABI declared in include/martini_b200.h).

The product path has no CPU fallback: if the CUDA library has not been built this module
raises at import of the first symbol, and every call that fails raises ``MartiniB200Error``
with the library's own message.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MTN_B200_LIB: developer switch -- another build of the same C ABI (an A/B variant from
# scripts/build_variants.sh); it must export every symbol, or load() raises
LIB_PATH = os.environ.get("MTN_B200_LIB") or os.path.join(_HERE, "libmartini_b200.so")

MTN_MAX_KERNELS = 8

# kernel / spectrum / flag codes (include/martini_b200.h)
KERNEL_WENDLANDC2 = 0
KERNEL_WENDLANDC6 = 1
KERNEL_CUBICSPLINE = 2
KERNEL_GAUSSIAN = 3
KERNEL_DIRACDELTA = 4
KERNEL_QUARTICSPLINE = 5
SPECTRUM_GAUSSIAN = 0
SPECTRUM_DIRACDELTA = 1
PRUNE_SPATIAL, PRUNE_SPECTRAL, PRUNE_MASS = 1, 2, 4
CUBE_ACCUMULATE, CUBE_ZEROED = 0, 1
ERR_INVALID, ERR_CUDA, ERR_WORKSPACE, ERR_LIMIT = -1, -2, -3, -4


class MartiniB200Error(RuntimeError):
    """A call into libmartini_b200.so failed."""


class MtnKernelEntry(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("valid_is_max", C.c_int32),
        ("rescale", C.c_double),
        ("size_in_fwhm", C.c_double),
        ("valid_size", C.c_double),
        ("truncate", C.c_double),
        ("norm", C.c_double),
    ]


class MtnKernelTable(C.Structure):
    _fields_ = [
        ("n", C.c_int32),
        ("adaptive", C.c_int32),
        ("k", MtnKernelEntry * MTN_MAX_KERNELS),
    ]


class MtnParticles(C.Structure):
    _fields_ = [
        ("n", C.c_int64),
        ("px", C.c_void_p),
        ("py", C.c_void_p),
        ("h_eff", C.c_void_p),
        ("sm_range", C.c_void_p),
        ("kernel_id", C.c_void_p),
        ("v", C.c_void_p),
        ("sigma", C.c_void_p),
        ("sigma_scalar", C.c_double),
        ("mHI", C.c_void_p),
        ("mHI_scalar", C.c_double),
        ("D", C.c_void_p),
        ("D_scalar", C.c_double),
        ("accept", C.c_void_p),
    ]


class MtnCube(C.Structure):
    _fields_ = [
        ("nx", C.c_int32),
        ("ny", C.c_int32),
        ("n_channels", C.c_int32),
        ("x_lo", C.c_int32),
        ("x_hi", C.c_int32),
        ("spectrum", C.c_int32),
        ("flags", C.c_int32),
        ("px_size_arcsec", C.c_double),
        ("edges", C.c_void_p),
        ("slab", C.c_void_p),
        ("edges_direction", C.c_int32),
        ("reserved", C.c_int32),
    ]


class MtnFrontEnd(C.Structure):
    _fields_ = [
        ("rotation", C.c_double * 9),
        ("direction", C.c_double * 3),
        ("distance_mpc", C.c_double),
        ("vpeculiar", C.c_double),
        ("hubble", C.c_double),
        ("ra0_rad", C.c_double),
        ("dec0_rad", C.c_double),
        ("px_size_arcsec", C.c_double),
        ("crpix", C.c_double * 3),
        ("spectral_centre", C.c_double),
        ("channel_width", C.c_double),
        ("freq_mode", C.c_int32),
        ("reserved", C.c_int32),
    ]


class MtnPlan(C.Structure):
    _fields_ = [
        ("n_kept", C.c_int64),
        ("n_pairs", C.c_int64),
        ("n_bricks", C.c_int64),
        ("updates_dense", C.c_int64),
        ("chunk", C.c_int64),
        ("n_pairs2", C.c_int64),
        ("chunk2", C.c_int64),
        ("edges_increasing", C.c_int32),
        ("route2", C.c_int32),
        ("workspace_bytes", C.c_size_t),
    ]


#: every symbol include/martini_b200.h declares, with (restype, argtypes)
SYMBOLS = {
    "mtn_version": (C.c_int, []),
    "mtn_last_error": (C.c_char_p, []),
    "mtn_last_launch_count": (C.c_int, []),
    "mtn_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "mtn_smoothing_setup": (
        C.c_int,
        [C.c_int64, C.c_void_p, C.POINTER(MtnKernelTable), C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_void_p],
    ),
    "mtn_prune": (
        C.c_int,
        [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
         C.c_void_p, C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
         C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "mtn_plan_scratch_bytes": (C.c_size_t, [C.c_int64, C.POINTER(MtnCube)]),
    "mtn_plan": (
        C.c_int,
        [C.POINTER(MtnParticles), C.POINTER(MtnKernelTable), C.POINTER(MtnCube), C.c_void_p, C.c_size_t,
         C.POINTER(MtnPlan), C.c_void_p],
    ),
    "mtn_project": (
        C.c_int,
        [C.POINTER(MtnParticles), C.POINTER(MtnKernelTable), C.POINTER(MtnCube),
         C.POINTER(MtnPlan), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p],
    ),
    "mtn_sky_to_pix": (
        C.c_int,
        [C.POINTER(MtnFrontEnd), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p,
         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "mtn_route_scratch_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "mtn_route_count": (
        C.c_int,
        [C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_void_p, C.c_size_t,
         C.c_void_p, C.c_void_p],
    ),
    "mtn_route_scatter": (
        C.c_int,
        [C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_int32,
         C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "mtn_set_timing": (C.c_int, [C.c_int]),
    "mtn_last_timing": (C.c_int, [C.POINTER(C.c_float), C.c_int]),
    "mtn_set_count_exec": (C.c_int, [C.c_int]),
    "mtn_last_exec_counts": (C.c_int, [C.POINTER(C.c_int64)]),
    "mtn_fp64_peak": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p]),
    "mtn_convolve_beam": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
         C.c_double, C.c_void_p],
    ),
    "mtn_table_error": (C.c_int, [C.c_int32, C.POINTER(C.c_double)]),
    "mtn_probe_kernel_integral": (
        C.c_int,
        [C.POINTER(MtnKernelEntry), C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_void_p],
    ),
    "mtn_probe_spectra": (
        C.c_int,
        [C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_int32,
         C.c_void_p, C.c_void_p, C.c_void_p],
    ),
}

_lib = None


def load():
    """Load the shared library (once) and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MartiniB200Error(
            f"{LIB_PATH} is missing: the CUDA library has not been built. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` from the repository root. "
            "martini_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int, what: str, lib=None) -> None:
    """Raise with the message of the library that made the call (``lib``; default: the one
    ``load()`` returns)."""
    if rc != 0:
        msg = (lib or load()).mtn_last_error().decode(errors="replace")
        raise MartiniB200Error(f"{what} failed (code {rc}): {msg}")
