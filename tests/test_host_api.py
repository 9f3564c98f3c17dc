"""CPU tests of the host-side mirror of MARTINI's data model (no GPU needed)."""

import numpy as np
import pytest

from martini_b200 import DataCube, L_coords, SPHSource, demo_source
from martini_b200 import spectral_models as S
from martini_b200 import sph_kernels as K


def test_init_pixcoords_kat():
    """reference tests/test_sources.py:335-380: 6 particles, 1 kpc = 1 arcsec, h = 0."""
    distance = 1.0e-3 / np.deg2rad(1.0 / 3600.0)  # Mpc such that 1 kpc subtends 1 arcsec
    lin = np.linspace(-2.5, 2.5, 6)
    source = SPHSource(distance=distance, h=0.0, T_g=np.ones(6) * 1e4, mHI_g=np.ones(6) * 1e4,
                       xyz_g=np.vstack((np.zeros(6), lin, lin)).T,
                       vxyz_g=np.vstack((lin, np.zeros(6), np.zeros(6))).T, hsm_g=np.ones(6))
    dc = DataCube(n_px_x=6, n_px_y=6, n_channels=6, px_size=1.0, channel_width=1.0)
    source._init_skycoords()
    source._init_pixcoords(dc)
    expected = np.vstack((np.arange(6)[::-1], np.arange(6), np.arange(6)[::-1]))
    assert np.allclose(source.pixcoords, expected, atol=1e-4)
    assert np.allclose(source.sm_lengths_px(dc), 1.0, atol=1e-6)


@pytest.mark.parametrize("ra", (0.0, 30.0, -30.0))
@pytest.mark.parametrize("dec", (0.0, 30.0, -30.0))
def test_source_centre_lands_on_cube_centre(ra, dec):
    s = SPHSource(distance=3.0, ra=ra, dec=dec, mHI_g=np.ones(1), xyz_g=np.zeros((1, 3)),
                  vxyz_g=np.zeros((1, 3)), hsm_g=np.ones(1))
    dc = DataCube(n_px_x=8, n_px_y=8, n_channels=4, px_size=10.0, channel_width=5.0,
                  spectral_centre=s.vsys, ra=ra, dec=dec)
    s._init_skycoords()
    s._init_pixcoords(dc)
    assert np.allclose(s.pixcoords[:, 0], (3.5, 3.5, 1.5), atol=1e-9)
    assert np.isclose(s.radial_velocity[0], 210.0) and np.isclose(s.distance_p[0], 3.0)


def test_datacube_layout_pad_and_edges():
    dc = DataCube(n_px_x=4, n_px_y=3, n_channels=6, px_size=10.0, channel_width=-4.0, spectral_centre=100.0)
    assert dc._array.shape == (4, 3, 6) and dc._array.dtype == np.float64
    e = dc.velocity_channel_edges
    assert e.shape == (7,) and np.all(np.diff(e) == -4.0) and e[0] == 112.0 and e[-1] == 88.0
    dc._array[:] = 1.0
    dc.add_pad((2, 1))
    assert dc._array.shape == (8, 5, 6) and dc._array.sum() == 72.0 and dc._array[0].sum() == 0
    with pytest.raises(RuntimeError, match="already padded"):
        dc.add_pad((1, 1))
    dc.drop_pad()
    assert dc._array.shape == (4, 3, 6) and (dc._array == 1.0).all()
    assert DataCube(n_px_x=2, n_px_y=2, n_channels=2, px_size=1.0, channel_width=1.0,
                    stokes_axis=True)._array.shape == (2, 2, 2, 1)


def test_apply_mask_semantics():
    s = demo_source(N=200)
    dc = DataCube(n_px_x=16, n_px_y=16, n_channels=8, px_size=30.0, channel_width=20.0, spectral_centre=s.vsys)
    s._init_skycoords()
    s._init_pixcoords(dc)
    mask = np.arange(200) % 2 == 0
    s.apply_mask(mask)
    assert s.npart == 100 and s.pixcoords.shape == (3, 100) and s.mHI_g.shape == (100,)
    with pytest.raises(ValueError, match="same length"):
        s.apply_mask(np.ones(3, dtype=bool))
    with pytest.raises(RuntimeError, match="No non-zero mHI source particles in target region."):
        s.apply_mask(np.zeros(100, dtype=bool))


def test_demo_source_matches_reference_recipe():
    """Same legacy-seeded sequence as martini/_demo.py: total mass and disc scale."""
    s = demo_source()
    assert s.npart == 500 and np.isclose(s.mHI_g.sum(), 5.0e9)
    r = np.sqrt((s.xyz_g**2).sum(axis=1))
    assert 2.5 < np.median(r) < 3.6
    assert np.isclose(s.vsys, 210.0)


def test_L_coords_inclination():
    s = demo_source(N=300)  # the disc's angular momentum is inclined 60 deg to the line of sight (x)
    L = np.sum(np.cross(s.xyz_g, s.mHI_g[:, None] * s.vxyz_g), axis=0)
    assert np.isclose(np.degrees(np.arccos(L[0] / np.linalg.norm(L))), 60.0, atol=3.0)
    assert L_coords().pa == 270.0


def test_thermal_sigma_kat():
    """reference tests/test_spectral_models.py:43-48."""
    src = type("S", (), {"T_g": np.array([1.0e4])})()
    assert np.isclose(S.GaussianSpectrum(sigma="thermal").half_width(src)[0], 9.0853727258, rtol=1e-9)
    assert S.GaussianSpectrum(sigma=7.0).half_width(src) == 7.0
    assert S.DiracDeltaSpectrum().half_width(src) == 0.0
    assert S.GaussianSpectrum(spec_dtype=np.float32).spec_dtype == np.float32  # accepted, see class
    with pytest.raises(NotImplementedError, match="float64"):
        S.GaussianSpectrum(spec_dtype=np.float16)


def test_kernel_constants_and_fwhm():
    """reference tests/test_sph_kernels.py:78-91 on the host-side kernel descriptions."""
    for cls in (K._WendlandC2Kernel, K._WendlandC6Kernel, K._CubicSplineKernel, K._GaussianKernel,
                K._QuarticSplineKernel):
        k = cls()
        assert np.isclose(k.eval_kernel(0.5, 1), k.eval_kernel(0, 1) / 2)
        assert k.eval_kernel(k.size_in_fwhm + 1e-5, 1) == 0 and k.eval_kernel(k.size_in_fwhm - 1e-5, 1) > 0
    with pytest.raises(RuntimeError, match="with truncation <2sigma"):
        K._GaussianKernel(truncate=1.0)
    with pytest.raises(NotImplementedError):
        K.AdaptiveKernel()
    t = K.kernel_table(K.WendlandC2Kernel())
    assert [e["kind"] for e in t.entries] == [0, 4, 3] and t.adaptive


# ------------------------------------------------------------------ noise
def _noise_setup(seed=0, rms=1.0):
    from martini_b200 import GaussianBeam, GaussianNoise

    dc = DataCube(n_px_x=64, n_px_y=48, n_channels=8, px_size=15.0, channel_width=4.0)
    beam = GaussianBeam()
    beam.init_kernel(dc)
    dc.add_pad(beam.needs_pad())
    return dc, beam, GaussianNoise(rms=rms, seed=seed)


def test_noise_shape_seed_noseed_reset():
    """reference tests/test_noise.py:13-94 on the mirror class."""
    dc, beam, gen = _noise_setup()
    n1 = gen.generate(dc, beam)
    assert n1.shape == dc._array.shape
    _, _, gen2 = _noise_setup()
    assert np.array_equal(n1, gen2.generate(dc, beam))  # same seed, same stream
    _, _, a = _noise_setup(seed=None)
    _, _, b = _noise_setup(seed=None)
    assert not np.allclose(a.generate(dc, beam), b.generate(dc, beam))
    gen.reset_rng()
    assert gen.seed is not None and np.array_equal(gen.generate(dc, beam), n1)


def test_noise_is_the_reference_formula():
    """noise.py:140-155 restated: default_rng(seed).normal(scale=rms*2.19568*sqrt(pi s_maj s_min))."""
    dc, beam, gen = _noise_setup(seed=7, rms=3.0e-3)
    s_maj = beam.bmaj / 2 / np.sqrt(2 * np.log(2)) / dc.px_size
    s_min = beam.bmin / 2 / np.sqrt(2 * np.log(2)) / dc.px_size
    scale = 3.0e-3 * 2.19568 * np.sqrt(np.pi * s_maj * s_min)
    want = np.random.default_rng(seed=7).normal(scale=scale, size=dc._array.shape)
    assert np.array_equal(gen.generate(dc, beam), want)


def test_datacube_array_residency():
    """The host array is materialised on access and a host access drops the device copy."""
    import torch

    class FakeEngine:
        def to_device(self, a):
            return torch.from_numpy(np.array(a, dtype=np.float64))

    dc = DataCube(n_px_x=4, n_px_y=3, n_channels=2, px_size=10.0, channel_width=4.0, stokes_axis=True)
    assert dc._array.shape == (4, 3, 2, 1) and dc._array_is_zero
    dc._array[1, 2, 0, 0] = 5.0  # in-place edits of the host array are seen by the next device step
    dev = dc._device_array(FakeEngine())
    assert dev.shape == (4, 3, 2) and float(dev[1, 2, 0]) == 5.0 and not dc._array_is_zero
    dev *= 2.0
    dc._set_device_array(dev)
    assert dc._host is None and not dc._array_is_zero
    assert dc._array.shape == (4, 3, 2, 1) and dc._array[1, 2, 0, 0] == 10.0
    assert dc._dev is None  # the host copy is authoritative again
    dc.add_pad((1, 2))
    assert dc._array.shape == (6, 7, 2, 1) and dc._array[2, 4, 0, 0] == 10.0
