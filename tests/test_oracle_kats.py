"""Replay, astropy-free, the reference's own known-answer tests for the hot path.

Each test names the reference test it replays (paths relative to /root/reference/tests).
These pin the oracle to analytic identities independent of the golden fixtures.
"""

import numpy as np
import pytest

from oracle import martini_oracle as O

FWHM_KERNELS = (
    ("_WendlandC2Kernel", {}),
    ("_WendlandC6Kernel", {}),
    ("_CubicSplineKernel", {}),
    ("_GaussianKernel", {}),
    ("_QuarticSplineKernel", {}),
)


def total_kernel_weight(k, h, ngrid=50):
    """test_sph_kernels.py:44-72."""
    r = np.arange(0, ngrid)
    xgrid, ygrid = np.meshgrid(np.r_[r[1:][::-1], r], np.r_[r[1:][::-1], r])
    dij = np.vstack((xgrid.flatten(), ygrid.flatten())).astype(np.float64)
    k.sm_lengths = np.ones(dij.shape[1]) * h
    return np.sum(k.px_weight(dij))


@pytest.mark.parametrize(("name", "kw"), FWHM_KERNELS)
def test_fwhm_is_one(name, kw):
    """test_sph_kernels.py:78-83."""
    k = O.make_kernel(name, **kw)
    assert np.isclose(k.eval_kernel(0.5, 1), k.eval_kernel(0, 1) / 2)


@pytest.mark.parametrize(("name", "kw"), FWHM_KERNELS)
def test_extent(name, kw):
    """test_sph_kernels.py:85-91."""
    k = O.make_kernel(name, **kw)
    assert k.eval_kernel(k.size_in_fwhm + 1.0e-5, 1) == 0
    assert k.eval_kernel(k.size_in_fwhm - 1.0e-5, 1) > 0


def test_rescale_constants():
    """the fsolve'd FWHM rescales."""
    assert np.isclose(O.make_kernel("_WendlandC2Kernel")._rescale, 1.5933199337482518, rtol=1e-12)
    assert np.isclose(O.make_kernel("_WendlandC6Kernel")._rescale, 1.9803392771718, rtol=1e-12)
    assert np.isclose(O.make_kernel("_CubicSplineKernel")._rescale, 1.3843671526381416, rtol=1e-12)
    assert np.isclose(O.make_kernel("_QuarticSplineKernel")._rescale, 1.5679974586080268, rtol=1e-12)


@pytest.mark.parametrize(
    "name", ("_WendlandC2Kernel", "_WendlandC6Kernel", "_CubicSplineKernel", "_QuarticSplineKernel")
)
def test_kernel_validation_minsize(name):
    """test_sph_kernels.py:136-160."""
    k = O.make_kernel(name)
    assert np.isclose(total_kernel_weight(k, 20), 1.0, rtol=1.0e-3)
    assert np.isclose(total_kernel_weight(k, k.min_valid_size * 1.001), 1.0, rtol=1.0e-2)
    assert not np.isclose(total_kernel_weight(k, k.min_valid_size * 0.9), 1.0, rtol=1.0e-2)


def test_kernel_validation_maxsize_diracdelta():
    """test_sph_kernels.py:162-182."""
    k = O.make_kernel("DiracDeltaKernel")
    assert np.isclose(total_kernel_weight(k, 20), 1.0, rtol=1.0e-3)
    assert np.isclose(total_kernel_weight(k, k.max_valid_size * 0.999), 1.0, rtol=1.0e-2)


@pytest.mark.parametrize("truncate", (1.0, 2.0, 3.0, 4.0, 5.0, 6.0))
def test_kernel_validation_gaussian(truncate):
    """test_sph_kernels.py:184-207."""
    if truncate < 2.0:
        with pytest.raises(RuntimeError, match="with truncation <2sigma"):
            O.make_kernel("_GaussianKernel", truncate=truncate)
        return
    k = O.make_kernel("_GaussianKernel", truncate=truncate)
    assert np.isclose(total_kernel_weight(k, 20), 1.0, rtol=3.0e-3)
    assert np.isclose(total_kernel_weight(k, k.min_valid_size * 1.1), 1.0, rtol=1.0e-2)


@pytest.mark.parametrize(("name", "kw"), FWHM_KERNELS)
def test_2D_integral(name, kw):
    """test_sph_kernels.py:93-134, on a coarser radial sampling to stay quick."""
    k = O.make_kernel(name, **kw)
    vmax, h = 50, 25
    r = np.arange(0, vmax)
    xg, yg = np.meshgrid(np.r_[r[1:][::-1], r], np.r_[r[1:][::-1], r])
    rg = np.sqrt(xg**2 + yg**2)
    xg3, yg3, zg3 = np.meshgrid(*(np.r_[r[1:][::-1], r],) * 3)
    rg3 = np.sqrt(xg3**2 + yg3**2 + zg3**2)
    Rg3 = np.sqrt(xg3**2 + yg3**2)
    W3 = k.eval_kernel(rg3.ravel().astype(float), h).reshape(rg3.shape)
    y2, y3 = [], []
    for ri in (0.5 * (r[1:] + r[:-1]))[4::5]:
        sel = rg <= ri
        k.sm_lengths = h * np.ones(sel.sum())
        y2.append(np.sum(k.px_weight(np.vstack((xg[sel], yg[sel])).astype(float))))
        y3.append(np.sum(W3[Rg3 <= ri]))
    assert np.isclose(np.sum(W3), 1.0, rtol=1.0e-2)
    assert np.allclose(y2, y3, rtol=2.0e-2)


@pytest.mark.parametrize(
    "name", ("WendlandC2Kernel", "WendlandC6Kernel", "CubicSplineKernel", "QuarticSplineKernel")
)
def test_kernel_selection(name):
    """test_sph_kernels.py:284-315: hsm (3, 1, 0.55, 0.1) kpc on 1 kpc pixels."""
    k = O.make_kernel(name)
    k.init_sm(np.array([3.0, 1.0, 0.55, 0.1]))
    assert all(k.kernel_indices == np.array([0, 0, 2, 1]))


def test_kernel_selection_gaussian():
    """test_sph_kernels.py:317-333."""
    k = O.make_kernel("GaussianKernel", truncate=3.0)
    k.init_sm(np.array([3.0, 1.0, 0.55, 0.1]))
    assert all(k.kernel_indices == np.array([0, 2, 2, 1]))


@pytest.mark.parametrize("threshold_factor", (0.9, 1.1))
@pytest.mark.parametrize(
    "name", ("_WendlandC2Kernel", "_WendlandC6Kernel", "_CubicSplineKernel", "_QuarticSplineKernel")
)
def test_confirm_validation(name, threshold_factor):
    """test_sph_kernels.py:209-278 (raise below threshold, 'use this with care')."""
    k = O.make_kernel(name)
    k.init_sm(np.array([k.min_valid_size * threshold_factor]))
    if threshold_factor < 1:
        with pytest.raises(RuntimeError, match="use this with care"):
            k.confirm_validation()
        assert not k.confirm_validation(noraise=True).any()
    else:
        assert k.confirm_validation().all()


# ------------------------------------------------------------------ spectral models
EDGES64 = (32 - np.arange(65)) * 4.0 + 1000.0  # 64 channels of 4 km/s, decreasing


def test_thermal_sigma_kat():
    """test_spectral_models.py:43-48: sigma(1e4 K) = 9.0853727258 km/s."""
    assert np.isclose(O.thermal_sigma(1.0e4), 9.0853727258, rtol=1e-9)


@pytest.mark.parametrize("sigma", ("thermal", 7.0))
def test_init_spectra_flux(sigma):
    """test_spectral_models.py:15-34: sum_c S_c dv = mHI / 2.36e5 at D = 1 Mpc."""
    sg = O.thermal_sigma(np.array([1.0e4])) if sigma == "thermal" else sigma
    mHI = np.array([1.0e4])
    sp = O.init_spectra(O.SPEC_GAUSSIAN, EDGES64, np.array([1000.0]), sg, mHI, np.array([1.0]))
    assert np.isclose(sp[0].sum() * 4.0, mHI[0] / 2.36e5, rtol=1.0e-5)


def test_init_spectra_flux_diracdelta():
    """test_spectral_models.py:101-119."""
    mHI = np.array([1.0e4])
    sp = O.init_spectra(O.SPEC_DIRACDELTA, EDGES64, np.array([1001.0]), 0.0, mHI, np.array([1.0]))
    assert np.isclose(sp[0].sum() * 4.0, mHI[0] / 2.36e5, rtol=1.0e-5)


@pytest.mark.parametrize("sigma", (9.0853727258, 7.0))
def test_spectral_function_normalised(sigma):
    """test_spectral_models.py:50-72."""
    s = O.gaussian_spectral_function(
        EDGES64[np.newaxis, 1:], EDGES64[np.newaxis, :-1], np.array([[1000.0]]), sigma
    )
    assert np.isclose(s.sum(), 1.0, rtol=1.0e-4)
    s = O.diracdelta_spectral_function(
        EDGES64[np.newaxis, 1:], EDGES64[np.newaxis, :-1], np.array([[1001.0]])
    )
    assert s.sum() == 1.0


def test_nonmonotonic_edges_raise():
    """spectral_models.py:187."""
    with pytest.raises(ValueError, match="Channel edges are not monotonic sequence."):
        O.init_spectra(O.SPEC_GAUSSIAN, np.array([0.0, 2.0, 1.0]), np.zeros(1), 1.0, np.ones(1), np.ones(1))


def test_numpy_axis_sum_is_sequential():
    """martini.py:281 sums over particles with np.sum(axis=-2): a sequential add in
    particle order for a C-ordered (n, C) array -- the order the oracle documents."""
    rng = np.random.Generator(np.random.PCG64(7))
    a = rng.normal(size=(1000, 16)) * 10.0 ** rng.uniform(-8, 8, size=(1000, 1))
    seq = np.zeros(16)
    for row in a:
        seq = seq + row
    assert np.array_equal(np.sum(a, axis=-2), seq)


# ------------------------------------------------------------------ prune truth table
@pytest.mark.parametrize(("ra_off", "ra_in"), ((0, True), (3, True), (9, False), (-3, True), (-9, False)))
@pytest.mark.parametrize(("dec_off", "dec_in"), ((0, True), (3, True), (9, False), (-3, True), (-9, False)))
@pytest.mark.parametrize(("v_off", "v_in"), ((0, True), (3, True), (7, False), (-3, True), (-7, False)))
@pytest.mark.parametrize(("mass_off", "mass_in"), ((0, False), (1, True)))
@pytest.mark.parametrize("flags", range(8))
def test_prune_truth_table(ra_off, ra_in, dec_off, dec_in, v_off, v_in, mass_off, mass_in, flags):
    """test_martini.py:330-443: 2x2x2 cube, pad 5, 1 arcsec px, 1 km/s channels,
    _CubicSplineKernel (hsm 1 kpc = 1 px -> sm_range 2), sigma = 1 km/s."""
    spatial, spectral, mass = bool(flags & 1), bool(flags & 2), bool(flags & 4)
    expect = all(
        ([ra_in, dec_in] if spatial else []) + ([v_in] if spectral else []) + ([mass_in] if mass else [])
    )
    # pixel coordinates of the offset particle: cube centre is pixel 0.5 + pad (crpix shifted by pad)
    pad = 5
    pix = np.array([[0.5 + pad - ra_off], [0.5 + pad + dec_off], [1.0 - v_off]], dtype=float)
    k = O.make_kernel("_CubicSplineKernel")
    k.init_sm(np.array([1.0]))
    assert k.sm_ranges[0] == 2.0
    acc = O.prune_mask(pix, k.sm_ranges, np.array([mass_off * 1.0e4]), 2 + 2 * pad, 2 + 2 * pad, 2,
                       1.0, 1.0, spatial=spatial, spectral=spectral, mass=mass)
    assert bool(acc[0]) == expect


def test_prune_nan():
    """test_martini.py:445-493."""
    k = O.make_kernel("_CubicSplineKernel")
    k.init_sm(np.array([1.0, 1.0]))
    pix = np.array([[np.nan, 5.0], [5.0, 5.0], [1.0, 1.0]])
    acc = O.prune_mask(pix, k.sm_ranges, np.ones(2), 12, 12, 2, 1.0, 1.0)
    assert acc.tolist() == [False, True]


# ------------------------------------------------------------------ mass conservation
@pytest.mark.parametrize(
    ("name", "kw"),
    FWHM_KERNELS + (("DiracDeltaKernel", {}),),
)
@pytest.mark.parametrize("spec", (O.SPEC_GAUSSIAN, O.SPEC_DIRACDELTA))
def test_mass_accuracy(name, kw, spec):
    """test_martini.py:205-241: mass recovered from the cube within 1 %."""
    rng = np.random.Generator(np.random.PCG64(11))
    n, X, Y, C, D, px_size = 40, 24, 24, 16, 3.0, 10.0
    k = O.make_kernel(name, **kw)
    sm = np.full(n, 0.4 if name == "DiracDeltaKernel" else 2.6)
    if name == "_GaussianKernel":
        sm[:] = 2.4
    k.init_sm(sm)
    pix = np.vstack((rng.uniform(8, 16, n), rng.uniform(8, 16, n), rng.uniform(6, 10, n)))
    edges = 32.0 - 4.0 * np.arange(C + 1)
    v = 32.0 - 4.0 * pix[2]
    mHI = rng.uniform(1, 2, n) * 1e6
    k.confirm_validation()
    spectra = O.init_spectra(spec, edges, v, 3.0, mHI, np.full(n, D))
    cube = O.insert_source_in_cube(np.zeros((X, Y, C)), pix, k, spectra, px_size)
    mass = 2.36e5 * D**2 * np.sum((cube * px_size**2).sum((0, 1)) * np.abs(np.diff(edges)))
    assert np.isclose(mass, mHI.sum(), rtol=1.0e-2)


def test_parallel_equals_serial():
    """test_martini.py:821-859: ncpu=2 gives the serial cube."""
    rng = np.random.Generator(np.random.PCG64(12))
    n, X, Y, C = 100, 12, 12, 8
    k = O.make_kernel("_GaussianKernel")
    k.init_sm(rng.uniform(2.4, 4.0, n))
    pix = np.vstack((rng.uniform(0, X, n), rng.uniform(0, Y, n), rng.uniform(0, C, n)))
    edges = 16.0 - 4.0 * np.arange(C + 1)
    spectra = O.init_spectra(O.SPEC_GAUSSIAN, edges, 16.0 - 4.0 * pix[2], 7.0, np.ones(n) * 1e5, np.full(n, 3.0))
    c1 = O.insert_source_in_cube(np.zeros((X, Y, C)), pix, k, spectra, 10.0, ncpu=1)
    c2 = O.insert_source_in_cube(np.zeros((X, Y, C)), pix, k, spectra, 10.0, ncpu=2)
    assert np.array_equal(c1, c2)


def test_count_updates():
    pix = np.array([[5.0, 0.2, 100.0], [5.0, 0.2, 5.0]])
    r = np.array([2.0, 1.0, 3.0])
    # 5x5 box; box [-0.8,1.2] -> pixels 0,1 -> 2x2; far outside in x -> 0
    assert O.count_updates(pix, r, 10, 10, 4) == (25 + 4 + 0) * 4
