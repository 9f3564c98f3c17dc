"""The oracle reproduces, bit for bit, what the reference's own code produced.

Fixtures under tests/golden/ were written by tests/golden/make_golden.py, which executes
the unmodified reference modules (sph_kernels.py, spectral_models.py, martini.py hot loop)
under a scale-1 units shim.  These tests never touch /root/reference.
"""

import glob
import os

import numpy as np
import pytest

from oracle import martini_oracle as O

PRIMS = {
    "_WendlandC2Kernel": ("_WendlandC2Kernel", {}),
    "_WendlandC6Kernel": ("_WendlandC6Kernel", {}),
    "_CubicSplineKernel": ("_CubicSplineKernel", {}),
    "_GaussianKernel_t3p0": ("_GaussianKernel", {"truncate": 3.0}),
    "_GaussianKernel_t6p0": ("_GaussianKernel", {"truncate": 6.0}),
    "_GaussianKernel_t2p5": ("_GaussianKernel", {"truncate": 2.5}),
    "DiracDeltaKernel": ("DiracDeltaKernel", {}),
    "_QuarticSplineKernel": ("_QuarticSplineKernel", {}),
}
ADAPTIVE = {
    "WendlandC2Kernel": ("WendlandC2Kernel", {}),
    "WendlandC6Kernel": ("WendlandC6Kernel", {}),
    "CubicSplineKernel": ("CubicSplineKernel", {}),
    "GaussianKernel_t3p0": ("GaussianKernel", {"truncate": 3.0}),
    "GaussianKernel_t4p0": ("GaussianKernel", {"truncate": 4.0}),
    "QuarticSplineKernel": ("QuarticSplineKernel", {}),
}


@pytest.mark.parametrize("tag", sorted(PRIMS))
def test_kernel_integrals_bit_exact(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "kernels.npz"))
    name, kw = PRIMS[tag]
    k = O.make_kernel(name, **kw)
    assert k._rescale == g[f"rescale_{tag}"]
    assert k.size_in_fwhm == g[f"size_in_fwhm_{tag}"]
    assert np.float64(k.norm) == g[f"norm_{tag}"]
    k.sm_lengths = g["h"]
    w = k.px_weight(np.vstack((g["dx"], g["dy"])))
    assert np.array_equal(w, g[f"w_{tag}"])
    assert (w > 0).sum() > 30  # the comparison is not vacuous
    if name != "DiracDeltaKernel":
        assert np.array_equal(k.eval_kernel(g["evalk_r"], 1.0), g[f"evalk_{tag}"])


@pytest.mark.parametrize("tag", sorted(ADAPTIVE))
def test_adaptive_selection_bit_exact(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "adaptive.npz"))
    name, kw = ADAPTIVE[tag]
    k = O.make_kernel(name, **kw)
    k.init_sm(g["sm_lengths"])
    assert np.array_equal(k.kernel_indices, g[f"kidx_{tag}"])
    assert np.array_equal(k.size_in_fwhm, g[f"size_in_fwhm_{tag}"])
    assert np.array_equal(k._rescale, g[f"rescale_{tag}"])
    assert np.array_equal(k.sm_ranges, g[f"sm_ranges_{tag}"])
    assert set(np.unique(k.kernel_indices)) >= {0, 1, 2}


@pytest.mark.parametrize("edir", ("dec", "inc"))
@pytest.mark.parametrize(
    "sname", ("gauss7", "gaussP", "gaussP_ncpu3", "gauss7_f32", "dirac")
)
def test_spectra_bit_exact(golden_dir, sname, edir):
    g = np.load(os.path.join(golden_dir, "spectra.npz"))
    kind = O.SPEC_DIRACDELTA if sname == "dirac" else O.SPEC_GAUSSIAN
    sigma = g["sigma"] if "gaussP" in sname else 7.0
    dtype = np.float32 if sname.endswith("f32") else np.float64
    sp = O.init_spectra(kind, g[f"edges_{edir}"], g["v"], sigma, g["mHI"], g["D"], dtype=dtype)
    ref = g[f"spectra_{sname}_{edir}"]
    assert sp.dtype == ref.dtype
    assert np.array_equal(sp, ref)
    assert (ref > 0).any()


@pytest.mark.parametrize("sname", ("gauss3", "gaussP", "dirac"))
@pytest.mark.parametrize("flags", range(1, 8))
def test_prune_mask_bit_exact(golden_dir, sname, flags):
    g = np.load(os.path.join(golden_dir, "prune.npz"))
    nx, ny, nc, pad = g["shape"]
    hw = {"gauss3": 3.0, "gaussP": g["sigma"], "dirac": 0.0}[sname]
    k = O.make_kernel("_CubicSplineKernel")
    k.init_sm(g["sm_lengths"])
    assert np.array_equal(k.sm_ranges, g["sm_ranges"])
    acc = O.prune_mask(
        np.vstack((g["px"], g["py"], g["pz"])), k.sm_ranges, g["mHI"],
        nx + 2 * pad, ny + 2 * pad, nc, hw, np.max(np.abs(np.diff(g["edges"]))),
        spatial=bool(flags & 1), spectral=bool(flags & 2), mass=bool(flags & 4),
    )
    ref = g[f"accept_{sname}_{flags}"]
    assert np.array_equal(acc, ref)
    assert 0 < ref.sum() < ref.size


INSERT_FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "insert_*.npz")))


def run_oracle_case(g, fast=False, dtype=np.float64):
    """Prune + insert a golden case with the oracle; returns the final cube."""
    nx, ny, nc, pad = g["shape"]
    name, trunc, sname = str(g["kernel"]), float(g["truncate"]), str(g["spectrum"])
    k = O.make_kernel(name, **({"truncate": trunc} if trunc else {}))
    k.init_sm(g["sm_lengths"])
    kind = O.SPEC_DIRACDELTA if sname == "dirac" else O.SPEC_GAUSSIAN
    sigma = g["sigma"]
    hw = 0.0 if sname == "dirac" else sigma
    X, Y = nx + 2 * pad, ny + 2 * pad
    pix = np.vstack((g["px"], g["py"], g["pz"]))
    acc = O.prune_mask(pix, k.sm_ranges, g["mHI"], X, Y, nc, hw, np.max(np.abs(np.diff(g["edges"]))))
    k.apply_mask(acc)
    sig = sigma[acc] if sigma.ndim else sigma
    cube0 = g["initial"] if g["initial"].size else np.zeros((X, Y, nc))
    if fast:
        cube = O.insert_fast(cube0, pix[:, acc], k, kind, g["edges"], g["v"][acc], sig,
                             g["mHI"][acc], g["D"][acc], float(g["px_size"]))
    else:
        spectra = O.init_spectra(kind, g["edges"], g["v"][acc], sig, g["mHI"][acc], g["D"][acc], dtype=dtype)
        cube = O.insert_source_in_cube(cube0, pix[:, acc], k, spectra, float(g["px_size"]),
                                       skip_validation=True)
    return acc, cube


@pytest.mark.parametrize("path", INSERT_FILES, ids=lambda p: os.path.basename(p)[7:-4])
def test_insert_bit_exact(path):
    g = np.load(path)
    acc, cube = run_oracle_case(g)
    assert np.array_equal(acc, g["accept"])
    assert np.array_equal(cube, g["cube"])
    assert np.abs(g["cube"]).max() > 0
    # the candidate-list variant used at larger sizes gives the very same bits
    _, cube_fast = run_oracle_case(g, fast=True)
    assert np.array_equal(cube_fast, g["cube"])


F32_FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "insertf32_*.npz")))


@pytest.mark.parametrize("path", F32_FILES, ids=lambda p: os.path.basename(p)[10:-4])
def test_insert_float32_mode_bit_exact(path):
    """The reference's spec_dtype=float32 mode (spectral_models.py:43-61, 297-300)."""
    g = np.load(path)
    acc, cube = run_oracle_case(g, dtype=np.float32)
    assert np.array_equal(acc, g["accept"])
    assert np.array_equal(cube, g["cube"])
    # ... and it is the float64 mode to float32 rounding: the gap the GPU path is allowed
    g64 = np.load(path.replace("insertf32_", "insert_"))
    assert 0 < np.abs(g["cube"] - g64["cube"]).max() <= 1e-6 * np.abs(g64["cube"]).max()


def test_insert_fixture_count():
    assert len(F32_FILES) == 2
    assert len(INSERT_FILES) == 42  # (8 primitive + 6 adaptive) kernels x 3 spectra


# ------------------------------------------------------------------------------------------
# seam.npz: the reference's _init_sm_lengths (sph_kernels.py:235-255, and through
# _AdaptiveKernel._init_sm_lengths :1241-1274) and GaussianSpectrum.half_width("thermal")
# (spectral_models.py:465-485), run unmodified under the scaled-unit stand-in
# ------------------------------------------------------------------------------------------
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
SEAM_PX = (("px10", 10.0), ("px3", 3.0), ("px0p7", 0.7))


@pytest.mark.parametrize("tag,px_size", SEAM_PX)
def test_sm_lengths_px_vs_reference(tag, px_size):
    g = np.load(os.path.join(GOLDEN, "seam.npz"))
    got = O.sm_lengths_px(g["hsm_kpc"], g["distance_Mpc"], px_size)
    ref = g[f"sm_lengths_{tag}"]
    assert np.array_equal(got == 0, ref == 0)
    assert np.allclose(got, ref, rtol=4.5e-16, atol=0)  # the converter factor's last ulp is astropy's
    # adaptive selection and ranges from the oracle's own lengths: same choices as the reference
    for name in ("WendlandC2Kernel", "CubicSplineKernel", "GaussianKernel"):
        k = O.make_kernel(name)
        k.init_sm(got)
        assert np.array_equal(k.kernel_indices, g[f"kidx_{name}_{tag}"])
        assert np.array_equal(k.sm_ranges, g[f"sm_ranges_{name}_{tag}"])


def test_thermal_half_width_vs_reference():
    g = np.load(os.path.join(GOLDEN, "seam.npz"))
    assert np.array_equal(O.thermal_sigma(g["T_K"]), g["half_width_thermal_kms"])


def test_product_host_seam_functions_vs_reference():
    """The product's host mirrors of the same two functions (martini_b200/sources.py,
    spectral_models.py) against the reference-made fixture."""
    from types import SimpleNamespace

    from martini_b200.sources import SPHSource
    from martini_b200.spectral_models import GaussianSpectrum

    g = np.load(os.path.join(GOLDEN, "seam.npz"))
    hw = GaussianSpectrum(sigma="thermal").half_width(SimpleNamespace(T_g=g["T_K"]))
    assert np.array_equal(hw, g["half_width_thermal_kms"])
    n = g["hsm_kpc"].size
    for tag, px_size in SEAM_PX:
        src = SPHSource.__new__(SPHSource)
        src.npart, src.hsm_g, src.distance_p = n, g["hsm_kpc"], g["distance_Mpc"]
        got = src.sm_lengths_px(SimpleNamespace(px_size=px_size))
        assert np.allclose(got, g[f"sm_lengths_{tag}"], rtol=4.5e-16, atol=0)
