"""GPU tests of the drop-in ``Martini`` class: the reference's own behavioural tests for the
hot path (tests/test_martini.py of the reference), replayed against martini_b200."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from martini_b200 import DataCube, GlobalProfile, Martini, PixelSource, SPHSource, demo, demo_source  # noqa: E402
from martini_b200 import synthetic  # noqa: E402
from martini_b200.martini import _BaseMartini, _PadOnlyBeam  # noqa: E402
from martini_b200.spectral_models import DiracDeltaSpectrum, GaussianSpectrum  # noqa: E402
from martini_b200.sph_kernels import (CubicSplineKernel, DiracDeltaKernel, GaussianKernel,  # noqa: E402
                                      WendlandC2Kernel, _CubicSplineKernel, _GaussianKernel,
                                      _QuarticSplineKernel, _WendlandC2Kernel, _WendlandC6Kernel)
from tests.parity import check_cube, oracle_hot_path  # noqa: E402

KPC_PER_ARCSEC_DISTANCE = 1.0e-3 / np.deg2rad(1.0 / 3600.0)  # Mpc: 1 kpc subtends 1 arcsec


def single_particle_source(ra_off=0.0, dec_off=0.0, v_off=0.0, mHI=1.0e4, hsm=1.0, distance=None):
    """One particle offset by arcsec on the sky (1 kpc = 1 arcsec) and km/s in velocity."""
    d = KPC_PER_ARCSEC_DISTANCE if distance is None else distance
    return SPHSource(distance=d, h=0.7, T_g=np.ones(1) * 1e4, mHI_g=np.ones(1) * mHI,
                     xyz_g=np.array([[0.0, ra_off, dec_off]]), vxyz_g=np.array([[v_off, 0.0, 0.0]]),
                     hsm_g=np.ones(1) * hsm)


def mass_in_cube(m):
    dc = m.datacube
    a = dc._array
    if dc.padx:
        a = a[dc.padx:-dc.padx, dc.pady:-dc.pady]
    dv = np.abs(np.diff(dc.velocity_channel_edges))
    return 2.36e5 * m.source.distance**2 * np.sum((a * dc.px_size**2).sum((0, 1)).squeeze() * dv)


@pytest.mark.parametrize("ra_off,ra_in", ((0, True), (3, True), (9, False), (-3, True), (-9, False)))
@pytest.mark.parametrize("dec_off,dec_in", ((0, True), (9, False), (-3, True)))
@pytest.mark.parametrize("v_off,v_in", ((0, True), (3, True), (7, False), (-7, False)))
@pytest.mark.parametrize("mass_off,mass_in", ((0, False), (1, True)))
@pytest.mark.parametrize("flags", ((1, 1, 1), (1, 0, 1), (0, 1, 0), (0, 0, 0)))
def test_prune_particles(ra_off, ra_in, dec_off, dec_in, v_off, v_in, mass_off, mass_in, flags):
    """reference test_martini.py:330-443: 2x2x2 cube, pad 5, _CubicSplineKernel, sigma 1 km/s."""
    spatial, spectral, mass = (bool(f) for f in flags)
    expect = all(([ra_in, dec_in] if spatial else []) + ([v_in] if spectral else [])
                 + ([mass_in] if mass else []))
    source = single_particle_source(ra_off, dec_off, v_off, mHI=mass_off * 1.0e4)
    datacube = DataCube(n_px_x=2, n_px_y=2, n_channels=2, spectral_centre=source.vsys, px_size=1.0,
                        channel_width=1.0)
    kwargs = dict(source=source, datacube=datacube, beam=_PadOnlyBeam(5), noise=None,
                  sph_kernel=_CubicSplineKernel(), spectral_model=GaussianSpectrum(sigma=1.0),
                  quiet=True, _prune_kwargs={"spatial": spatial, "spectral": spectral, "mass": mass})
    if not expect:
        with pytest.raises(RuntimeError, match="No non-zero mHI source particles in target region."):
            _BaseMartini(**kwargs)
    else:
        assert _BaseMartini(**kwargs).source.npart == 1


@pytest.mark.parametrize("kernel,hsm", ((_WendlandC2Kernel, 2.0), (_WendlandC6Kernel, 2.0),
                                        (_CubicSplineKernel, 2.0), (_GaussianKernel, 2.5),
                                        (_QuarticSplineKernel, 2.0), (DiracDeltaKernel, 0.3),
                                        (WendlandC2Kernel, 1.0), (GaussianKernel, 0.7)))
@pytest.mark.parametrize("spectrum", (GaussianSpectrum, DiracDeltaSpectrum))
def test_mass_accuracy(kernel, hsm, spectrum):
    """reference test_martini.py:205-241: mass recovered from the cube within 1 %."""
    # 1 kpc pixels at 3 Mpc; positions kept off the pixel corners (the Dirac-delta kernel is
    # strict: a particle exactly on a pixel edge lands nowhere, sph_kernels.py:1165)
    source = SPHSource(distance=3.0, mHI_g=np.ones(4) * 1e6, T_g=np.ones(4) * 1e4,
                       xyz_g=np.array([[0, 0.3, 0.2], [0, 1.3, 0.1], [0, 0.2, 1.4], [0, -1.3, -1.1]]),
                       vxyz_g=np.array([[0, 0, 0], [5, 0, 0], [-5, 0, 0], [10.0, 0, 0]]),
                       hsm_g=np.ones(4) * hsm)
    datacube = DataCube(n_px_x=32, n_px_y=32, n_channels=32, px_size=68.75493542 / 1.0,
                        channel_width=4.0, spectral_centre=source.vsys)  # 1 kpc pixels at 3 Mpc
    m = Martini(source=source, datacube=datacube, sph_kernel=kernel(), spectral_model=spectrum(),
                quiet=True)
    m.insert_source_in_cube()
    assert datacube.array_unit == "Jy/arcsec2"
    assert np.isclose(mass_in_cube(m), 4e6, rtol=1e-2)


@pytest.mark.parametrize("kernel,hsm", ((_WendlandC2Kernel, 2.0), (DiracDeltaKernel, 0.3), (GaussianKernel, 0.7)))
@pytest.mark.parametrize("spectrum", (GaussianSpectrum, DiracDeltaSpectrum))
def test_mass_accuracy_frequency_channels(kernel, hsm, spectrum):
    """The reference's mass-accuracy matrix also runs frequency-mode cubes (test_martini.py:205-241
    with dc_zeros in Hz, datacube.py:196-207): channels of equal frequency width, the spectral
    centre given as a velocity.  Mass within 1 %, and the cube equals the velocity-mode cube of
    the same channels (v = c (1 - f / f_HI) is linear, so the edges agree to rounding)."""
    from martini_b200.datacube import C_KMS, HI_FREQ_HZ

    def build(**channels):
        source = SPHSource(distance=3.0, mHI_g=np.ones(4) * 1e6, T_g=np.ones(4) * 1e4,
                           xyz_g=np.array([[0, 0.3, 0.2], [0, 1.3, 0.1], [0, 0.2, 1.4], [0, -1.3, -1.1]]),
                           vxyz_g=np.array([[0, 0, 0], [5, 0, 0], [-5, 0, 0], [10.0, 0, 0]]),
                           hsm_g=np.ones(4) * hsm)
        dc = DataCube(n_px_x=32, n_px_y=32, n_channels=32, px_size=68.75493542, spectral_centre=source.vsys,
                      **channels)
        m = Martini(source=source, datacube=dc, sph_kernel=kernel(), spectral_model=spectrum(), quiet=True)
        m.insert_source_in_cube()
        return m

    mf = build(channel_width=HI_FREQ_HZ * 4.0 / C_KMS, channel_unit="Hz")
    dc = mf.datacube
    assert dc._freq_channel_mode and dc.channel_unit == "Hz"
    assert np.all(np.diff(dc.channel_edges) > 0) and np.all(np.diff(dc.velocity_channel_edges) < 0)
    assert np.isclose(mass_in_cube(mf), 4e6, rtol=1e-2)
    mv = build(channel_width=4.0)
    assert np.allclose(dc.velocity_channel_edges, mv.datacube.velocity_channel_edges, rtol=0, atol=1e-9)
    a, b = mf.datacube._array, mv.datacube._array
    if spectrum is GaussianSpectrum:
        assert np.abs(a - b).max() <= 1e-9 * np.abs(b).max()
    else:  # a Dirac line may flip channel when an edge moves by an ulp: compare the moment-0 maps
        assert np.abs(a.sum(2) - b.sum(2)).max() <= 1e-9 * np.abs(b.sum(2)).max()
    mf.reset()  # reset keeps the channel mode
    assert mf.datacube._freq_channel_mode and np.allclose(mf.datacube.channel_edges, dc.channel_edges)


def _engine_for_tests():
    import martini_b200.martini as M

    return M.Engine("cuda:0")  # (the emulated twin of this module swaps M.Engine)


def test_device_front_end_kat():
    """SURVEY row f1 on the device (mtn_sky_to_pix): the reference's pixel-coordinate KAT
    (tests/test_sources.py:335-380: 6 particles, 1 kpc = 1 arcsec, h = 0)."""
    distance = 1.0e-3 / np.deg2rad(1.0 / 3600.0)
    lin = np.linspace(-2.5, 2.5, 6)
    source = SPHSource(distance=distance, h=0.0, T_g=np.ones(6) * 1e4, mHI_g=np.ones(6) * 1e4,
                       xyz_g=np.vstack((np.zeros(6), lin, lin)).T,
                       vxyz_g=np.vstack((lin, np.zeros(6), np.zeros(6))).T, hsm_g=np.ones(6))
    dc = DataCube(n_px_x=6, n_px_y=6, n_channels=6, px_size=1.0, channel_width=1.0)
    dev = source._init_on_device(_engine_for_tests(), dc)
    expected = np.vstack((np.arange(6)[::-1], np.arange(6), np.arange(6)[::-1]))
    got = np.vstack([dev[k].cpu().numpy() for k in ("px", "py", "pz")])
    assert np.allclose(got, expected, atol=1e-4)
    assert np.allclose(source.pixcoords, expected, atol=1e-4)  # the lazily fetched host mirror
    assert np.allclose(dev["sm_length"].cpu().numpy(), 1.0, atol=1e-6)


@pytest.mark.parametrize("freq", (False, True))
@pytest.mark.parametrize("ra,dec", ((0.0, 0.0), (148.3, -31.7), (271.0, 64.2)))
def test_device_front_end_matches_host_mirror(ra, dec, freq):
    """mtn_sky_to_pix against the host numpy restatement of sph_source.py:265-362 +
    sph_kernels.py:250-253 on a random source: rotation, translation, peculiar velocity, Hubble
    flow, TAN projection with pad, both channel modes; 1e-9 pixel / 1e-12 relative."""
    from martini_b200.datacube import C_KMS, HI_FREQ_HZ

    rng = np.random.Generator(np.random.PCG64(17))
    n = 5000
    kw = dict(distance=12.5, vpeculiar=83.0, ra=ra, dec=dec, h=0.7, mHI_g=np.ones(n),
              xyz_g=rng.normal(0, 12.0, (n, 3)), vxyz_g=rng.normal(0, 150.0, (n, 3)),
              hsm_g=rng.lognormal(0.0, 0.7, n), L_coords=None)
    host, devs = SPHSource(**kw), SPHSource(**kw)
    ch = dict(channel_width=HI_FREQ_HZ * 5.0 / C_KMS, channel_unit="Hz") if freq else dict(channel_width=5.0)

    def cube():
        dc = DataCube(n_px_x=64, n_px_y=48, n_channels=40, px_size=6.0, spectral_centre=host.vsys,
                      ra=ra + 0.01, dec=dec - 0.005, **ch)
        dc.add_pad((7, 5))
        return dc

    host._init_skycoords()
    host._init_pixcoords(cube())
    dev = devs._init_on_device(_engine_for_tests(), cube())
    for k, want in (("px", host.pixcoords[0]), ("py", host.pixcoords[1]), ("pz", host.pixcoords[2])):
        assert np.abs(dev[k].cpu().numpy() - want).max() < 1e-9, k
    assert np.allclose(dev["v"].cpu().numpy(), host.radial_velocity, rtol=1e-12, atol=1e-10)
    assert np.allclose(dev["D"].cpu().numpy(), host.distance_p, rtol=1e-13)
    assert np.allclose(dev["sm_length"].cpu().numpy(), host.sm_lengths_px(cube()), rtol=1e-12)
    assert np.allclose(devs.distance_p, host.distance_p, rtol=1e-13) and devs.pixcoords.shape == (3, n)


def test_kernel_validation_raises_unless_skipped():
    """reference test_sph_kernels.py:209-278: 'use this with care'."""
    def build():
        source = SPHSource(distance=3.0, mHI_g=np.ones(2) * 1e6, xyz_g=np.zeros((2, 3)),
                           vxyz_g=np.zeros((2, 3)), hsm_g=np.array([0.5, 3.0]))  # 0.5 px < 1.51
        dc = DataCube(n_px_x=16, n_px_y=16, n_channels=16, px_size=68.75493542, channel_width=4.0,
                      spectral_centre=source.vsys)
        return Martini(source=source, datacube=dc, sph_kernel=_WendlandC2Kernel(),
                       spectral_model=GaussianSpectrum(), quiet=True)
    with pytest.raises(RuntimeError, match="use this with care"):
        build().insert_source_in_cube()
    m = build()
    m.insert_source_in_cube(skip_validation=True)
    assert m.datacube._array.max() > 0


def test_reset_and_reinsert():
    """reference test_martini.py:495-507."""
    m = demo(quiet=True)
    first = m.datacube._array.copy()
    assert first.shape == (154, 154, 32) and first.sum() > 0
    with pytest.raises(RuntimeError, match="reset"):
        m.insert_source_in_cube()
    m.reset()
    assert m.datacube._array.sum() == 0 and m.datacube._array.shape == (154, 154, 32)
    m.insert_source_in_cube()
    assert np.array_equal(m.datacube._array, first)


def test_reset_after_convolve_repads_for_the_beam():
    """insert -> convolve_beam -> reset -> insert -> convolve_beam equals a fresh instance:
    reset() re-pads with beam.needs_pad() (reference martini.py:409-425), not with the pad
    convolve_beam() has already dropped (which shifted the re-inserted source by 13 px)."""
    m = demo(quiet=True, convolve=True)
    first = m.datacube._array.copy()
    assert first.shape[:2] == (128, 128) and m.datacube.padx == 0
    m.reset()
    assert (m.datacube.padx, m.datacube.pady) == (13, 13)
    assert m.datacube._array.shape[:2] == (154, 154) and m.datacube._array.sum() == 0
    m.insert_source_in_cube()
    m.convolve_beam()
    assert np.array_equal(m.datacube._array, first)


def test_demo_mass_and_spectra_attribute():
    """BASELINE config 1: demo source into the demo cube; reference test_martini.py:523-531."""
    m = demo(quiet=True)
    assert m.spectral_model.spectra is None  # never materialised by insertion
    assert np.isclose(mass_in_cube(m), m.source.input_mass, rtol=2e-2)
    m.init_spectra()
    sp = m.spectral_model.spectra
    assert sp.shape == (m.source.npart, 32) and sp.sum() > 0


def test_noise_before_insertion_is_kept():
    """martini.py:916-927: out = (in + inserted) / px^2."""
    src = single_particle_source(hsm=3.0, mHI=1e6)
    dc = DataCube(n_px_x=16, n_px_y=16, n_channels=8, px_size=1.0, channel_width=4.0, spectral_centre=src.vsys)
    rng = np.random.Generator(np.random.PCG64(5))
    noise = rng.normal(0, 1e-9, dc._array.shape)
    dc._array += noise
    m = Martini(source=src, datacube=dc, sph_kernel=_CubicSplineKernel(), spectral_model=GaussianSpectrum(), quiet=True)
    m.insert_source_in_cube()
    src2 = single_particle_source(hsm=3.0, mHI=1e6)
    dc2 = DataCube(n_px_x=16, n_px_y=16, n_channels=8, px_size=1.0, channel_width=4.0, spectral_centre=src2.vsys)
    m2 = Martini(source=src2, datacube=dc2, sph_kernel=_CubicSplineKernel(), spectral_model=GaussianSpectrum(), quiet=True)
    m2.insert_source_in_cube()
    assert np.allclose(m.datacube._array, m2.datacube._array + noise, rtol=1e-12, atol=1e-25)


def test_martini_class_vs_oracle():
    """The Martini class on a synthetic pixel-space source equals the oracle's cube."""
    case = synthetic.make_case("cfg2", n=20000, nx=48, ny=40, nc=64, seed=21)
    nx, ny, nc = case["shape"]
    dc = DataCube(n_px_x=nx, n_px_y=ny, n_channels=nc, px_size=case["px_size"], channel_width=4.0)
    assert np.allclose(dc.velocity_channel_edges, case["edges"])
    m = Martini(source=PixelSource.from_case(case), datacube=dc, sph_kernel=WendlandC2Kernel(),
                spectral_model=GaussianSpectrum(sigma=7.0), quiet=True)
    ref = oracle_hot_path(case)
    assert m.source.npart == int(ref["accept"].sum())
    assert np.array_equal(m.sph_kernel.kernel_indices, ref["kernel"].kernel_indices)
    assert np.array_equal(m.sph_kernel.sm_ranges, ref["kernel"].sm_ranges)
    m.insert_source_in_cube()
    check_cube(m.datacube._array, ref["cube"])


def test_global_profile():
    """reference test_martini.py:862-912: spectrum integrates to the source mass within 1 %,
    sky position is ignored by the pruning."""
    s = demo_source(N=400)
    gp = GlobalProfile(source=s, spectral_model=GaussianSpectrum(sigma=7.0), n_channels=64,
                       channel_width=10.0, spectral_centre=s.vsys, quiet=True)
    spec = gp.spectrum
    assert spec.shape == (64,)
    mass = 2.36e5 * 3.0**2 * np.sum(spec * 10.0)
    assert np.isclose(mass, 5.0e9, rtol=1e-2)


def test_unsupported_plugins_raise_before_any_work():
    class MyKernel(_CubicSplineKernel):
        pass

    src = single_particle_source()
    dc = DataCube(n_px_x=4, n_px_y=4, n_channels=4, px_size=1.0, channel_width=1.0)
    with pytest.raises(NotImplementedError, match="MyKernel"):
        Martini(source=src, datacube=dc, sph_kernel=MyKernel(), spectral_model=GaussianSpectrum())


@pytest.mark.parametrize("bmaj,bmin,bpa", ((30.0, 30.0, 0.0), (40.0, 20.0, 30.0)))
def test_convolve_beam_vs_fftconvolve(bmaj, bmin, bpa):
    """Martini.convolve_beam (martini.py:863-901) on the GPU against the
    reference's per-channel scipy fftconvolve, incl. pad drop and the Jy/beam conversion."""
    from martini_b200.beams import GaussianBeam
    from oracle import martini_oracle as O

    s = demo_source(N=300)
    dc = DataCube(n_px_x=48, n_px_y=40, n_channels=16, px_size=10.0, channel_width=20.0, spectral_centre=s.vsys)
    beam = GaussianBeam(bmaj=bmaj, bmin=bmin, bpa=bpa, truncate=4.0)
    m = Martini(source=s, datacube=dc, beam=beam, spectral_model=GaussianSpectrum(sigma=7.0),
                sph_kernel=CubicSplineKernel(), quiet=True)
    pad = beam.needs_pad()
    assert pad == (int(np.ceil(bmaj * 4 / 10 + 1)),) * 2 and dc._array.shape == (48 + 2 * pad[0], 40 + 2 * pad[1], 16)
    assert np.isclose(beam.kernel.sum(), 1.0, rtol=2e-3)  # truncated at 4 FWHM
    m.insert_source_in_cube()
    before = m.datacube._array.copy()
    ref = O.convolve_beam(before, beam.kernel, beam.area, *pad)
    m.convolve_beam()
    got = m.datacube._array
    assert got.shape == (48, 40, 16) and m.datacube.padx == 0 and m.datacube.array_unit == "Jy/beam"
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
    assert np.abs(ref).max() > 0


def test_convolve_beam_needs_pad_and_beam():
    s = demo_source(N=50)
    dc = DataCube(n_px_x=16, n_px_y=16, n_channels=4, px_size=10.0, channel_width=50.0, spectral_centre=s.vsys)
    m = Martini(source=s, datacube=dc, spectral_model=GaussianSpectrum(), sph_kernel=CubicSplineKernel(), quiet=True)
    m.insert_source_in_cube()
    with pytest.warns(UserWarning, match="no beam object"):
        m.convolve_beam()


# ------------------------------------------------------------------ SURVEY row f4
def _noisy_martini(seed=0, rms=1.0e-6, insert=True, N=300):
    from martini_b200 import GaussianBeam, GaussianNoise

    s = demo_source(N=N)
    dc = DataCube(n_px_x=64, n_px_y=64, n_channels=16, px_size=10.0, channel_width=20.0, spectral_centre=s.vsys)
    m = Martini(source=s, datacube=dc, beam=GaussianBeam(bmaj=30.0, bmin=30.0), noise=GaussianNoise(rms=rms, seed=seed),
                spectral_model=GaussianSpectrum(sigma=7.0), sph_kernel=CubicSplineKernel(), quiet=True)
    if insert:
        m.insert_source_in_cube()
    return m


def test_noise_amplitude():
    """reference tests/test_noise.py:28-36: zero the cube, add noise, convolve, measure the rms."""
    m = _noisy_martini()
    m.datacube._array[...] = 0.0
    m.add_noise()
    m.convolve_beam()
    a = m.datacube._array
    assert a.shape == (64, 64, 16) and m.datacube.array_unit == "Jy/beam"
    assert np.isclose(np.sqrt(np.mean(a**2)), m.noise.target_rms, rtol=0.1)


def test_add_noise_is_the_reference_arithmetic_and_order_independent():
    """martini.py:916-928: the noise realisation (Jy/beam) / beam area, in the cube's unit,
    added to the array -- before or after the source insertion."""
    after = _noisy_martini(seed=4)
    clean = after.datacube._array.copy()
    after.add_noise()
    gen = _noisy_martini(seed=4, insert=False).noise
    want = clean + gen.generate(after.datacube, after.beam) / after.beam.area
    assert np.abs(after.datacube._array - want).max() <= 1e-15 * np.abs(want).max()
    before = _noisy_martini(seed=4, insert=False)
    before.add_noise()  # cube still in Jy/pix^2
    before.insert_source_in_cube()
    assert np.abs(before.datacube._array - want).max() <= 1e-12 * np.abs(want).max()


def test_add_noise_needs_noise_and_beam():
    s = demo_source(N=50)
    dc = DataCube(n_px_x=16, n_px_y=16, n_channels=4, px_size=10.0, channel_width=50.0, spectral_centre=s.vsys)
    m = Martini(source=s, datacube=dc, spectral_model=GaussianSpectrum(), sph_kernel=CubicSplineKernel(), quiet=True)
    with pytest.warns(UserWarning, match="no noise object"):
        m.add_noise()
    from martini_b200 import GaussianNoise

    m.noise = GaussianNoise()
    with pytest.warns(UserWarning, match="no beam object"):
        m.add_noise()


def test_cube_stays_on_the_device_between_steps():
    m = _noisy_martini()
    dc = m.datacube
    assert dc._host is None and dc._dev is not None and dc._dev.is_cuda  # no download after insertion
    m.add_noise()
    m.convolve_beam()
    assert dc._host is None and dc._dev.shape == (64, 64, 16)
    a = dc._array  # first host access materialises the array ...
    assert a.shape == (64, 64, 16) and dc._dev is None  # ... and makes the host copy authoritative


def test_float32_spec_dtype_is_accepted():
    """The reference's memory-saving mode; here spectra never exist as an array, the answer is
    the float64 one (tests/test_gpu_parity.py pins the gap to the reference's float32 cube)."""
    def run(dtype):
        s = demo_source(N=200)
        dc = DataCube(n_px_x=32, n_px_y=32, n_channels=16, px_size=10.0, channel_width=20.0, spectral_centre=s.vsys)
        m = Martini(source=s, datacube=dc, spectral_model=GaussianSpectrum(sigma=7.0, spec_dtype=dtype),
                    sph_kernel=CubicSplineKernel(), quiet=True)
        m.insert_source_in_cube()
        return m
    m32, m64 = run(np.float32), run(np.float64)
    assert np.array_equal(m32.datacube._array, m64.datacube._array)
    m32.init_spectra()
    assert m32.spectral_model.spectra.dtype == np.float32
