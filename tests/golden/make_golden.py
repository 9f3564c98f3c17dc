"""Generate the golden fixtures in this directory by running the reference's OWN code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It loads ``martini/sph_kernels.py``, ``martini/spectral_models.py`` and
``martini/martini.py`` unmodified from ``/root/reference`` under the scale-1 units shim
``oracle/refshim.py`` (astropy is not installable here) and records, as ``.npz`` files:

* ``kernels.npz``   -- ``_px_weight`` of every primitive kernel on a set of (dx, dy, h),
                       their ``_rescale`` / ``size_in_fwhm`` / ``norm`` constants, and
                       ``eval_kernel`` KAT points;
* ``adaptive.npz``  -- ``kernel_indices`` / ``size_in_fwhm`` / ``_rescale`` /
                       ``sm_ranges`` chosen by the five public adaptive kernels;
* ``spectra.npz``   -- ``init_spectra`` output of GaussianSpectrum (scalar and
                       per-particle sigma, float64 and float32) and DiracDeltaSpectrum;
* ``prune.npz``     -- ``_BaseMartini._prune_particles`` accept masks;
* ``insert_*.npz``  -- full ``_prune_particles`` + ``_insert_source_in_cube`` runs
                       (inputs and the final Jy/arcsec^2 cube) for every kernel x spectrum.

Inputs are drawn from numpy's PCG64 with fixed seeds and stored next to the outputs, so the
tests never need the reference tree.
"""

from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refshim  # noqa: E402

K, S, M, U = refshim.load_reference()
Q = refshim.Quantity


# ---------------------------------------------------------------------------------------
# fake source / datacube objects carrying exactly the attributes the hot path reads
# ---------------------------------------------------------------------------------------


class FakeSource:
    def __init__(self, pixcoords, mHI, v, D, sigma=None):
        self.pixcoords = Q(pixcoords)
        self.mHI_g = Q(mHI)
        self.skycoords = SimpleNamespace(radial_velocity=Q(v), distance=Q(D))
        self.distance = Q(float(np.mean(D)))
        self.input_mass = Q(float(np.sum(mHI)))
        self._sigma = None if sigma is None else Q(sigma)
        self.npart = self.mHI_g.size

    def apply_mask(self, mask):
        # mirrors sources/sph_source.py:364-393
        if mask.size != self.pixcoords.shape[-1]:
            raise ValueError("Mask must have same length as particle arrays.")
        mask_sum = np.sum(mask)
        if mask_sum == 0:
            raise RuntimeError("No non-zero mHI source particles in target region.")
        self.npart = mask_sum
        self.mHI_g = self.mHI_g[mask]
        self.pixcoords = self.pixcoords[:, mask]
        self.skycoords = SimpleNamespace(
            radial_velocity=self.skycoords.radial_velocity[mask],
            distance=self.skycoords.distance[mask],
        )
        if self._sigma is not None and self._sigma.ndim > 0:
            self._sigma = self._sigma[mask]


class FakeDataCube:
    def __init__(self, nx, ny, nc, pad, edges, px_size, initial=None):
        self.n_px_x, self.n_px_y, self.n_channels = nx, ny, nc
        self.padx = self.pady = pad
        self.px_size = Q(px_size)
        self.velocity_channel_edges = Q(edges)
        shape = (nx + 2 * pad, ny + 2 * pad, nc)
        self._array = Q(np.zeros(shape) if initial is None else initial.copy())
        # datacube.py:189-194
        self.arcsec2_to_pix = (
            U.Jy * U.pix**-2,
            U.Jy * U.arcsec**-2,
            lambda x: x / self.px_size.to_value(U.arcsec) ** 2,
            lambda x: x * self.px_size.to_value(U.arcsec) ** 2,
        )


def set_sm(kernel, sm_lengths):
    """Give a reference kernel its pixel-unit smoothing lengths, then let the reference's
    own code derive everything else (adaptive selection, sm_ranges)."""
    sm = Q(sm_lengths)
    if isinstance(kernel, K._AdaptiveKernel):
        # _AdaptiveKernel._init_sm_lengths calls super()._init_sm_lengths (the WCS-bound
        # arctan(hsm/D) step, upstream of the hot path) and then does the selection.
        orig = K._BaseSPHKernel._init_sm_lengths
        K._BaseSPHKernel._init_sm_lengths = lambda self, source, datacube: setattr(
            self, "sm_lengths", sm
        )
        try:
            kernel._init_sm_lengths(
                source=SimpleNamespace(mHI_g=np.zeros(sm.shape)), datacube=None
            )
        finally:
            K._BaseSPHKernel._init_sm_lengths = orig
    else:
        kernel.sm_lengths = sm
    kernel._init_sm_ranges()


def ref_kernel(name, **kw):
    return getattr(K, name)(**kw)


PRIMS = (
    ("_WendlandC2Kernel", {}),
    ("_WendlandC6Kernel", {}),
    ("_CubicSplineKernel", {}),
    ("_GaussianKernel", {"truncate": 3.0}),
    ("_GaussianKernel", {"truncate": 6.0}),
    ("_GaussianKernel", {"truncate": 2.5}),
    ("DiracDeltaKernel", {}),
    ("_QuarticSplineKernel", {}),
)
ADAPTIVE = (
    ("WendlandC2Kernel", {}),
    ("WendlandC6Kernel", {}),
    ("CubicSplineKernel", {}),
    ("GaussianKernel", {"truncate": 3.0}),
    ("GaussianKernel", {"truncate": 4.0}),
    ("QuarticSplineKernel", {}),
)


def tag(name, kw):
    return name + ("" if not kw else "_t" + str(kw["truncate"]).replace(".", "p"))


# ---------------------------------------------------------------------------------------
def gen_kernels():
    rng = np.random.Generator(np.random.PCG64(101))
    n = 4000
    h = np.r_[rng.uniform(0.3, 12.0, n), [1.0, 2.0, 2.5, 4.0, 0.4]]
    # offsets spanning well past every kernel's support, plus exact specials
    dx = np.r_[rng.uniform(-1.0, 1.0, n) * h[:n] * 2.2, [0.0, 0.0, 2.5, 0.5, -0.5]]
    dy = np.r_[rng.uniform(-1.0, 1.0, n) * h[:n] * 2.2, [0.0, 1.0, 0.0, 0.5, 0.25]]
    out = {"dx": dx, "dy": dy, "h": h}
    for name, kw in PRIMS:
        k = ref_kernel(name, **kw)
        k.sm_lengths = Q(h)
        w = k._px_weight(Q(np.vstack((dx, dy))))
        t = tag(name, kw)
        out[f"w_{t}"] = np.asarray(w)
        out[f"rescale_{t}"] = np.float64(k._rescale)
        out[f"size_in_fwhm_{t}"] = np.float64(k.size_in_fwhm)
        out[f"norm_{t}"] = np.float64(getattr(k, "norm", 1.0))
        if name != "DiracDeltaKernel":
            rr = np.linspace(0.0, 2.5, 26)
            out[f"evalk_{t}"] = np.asarray(k.eval_kernel(rr, 1.0))
    out["evalk_r"] = np.linspace(0.0, 2.5, 26)
    np.savez_compressed(os.path.join(HERE, "kernels.npz"), **out)


def gen_adaptive():
    rng = np.random.Generator(np.random.PCG64(102))
    sm = np.r_[
        rng.lognormal(np.log(1.0), 0.9, 500),
        [0.1, 0.336, 0.3, 0.45, 0.5, 0.55, 0.9477, 1.0, 3.0, 1.16 / 1.3843671526381416],
    ]
    out = {"sm_lengths": sm}
    for name, kw in ADAPTIVE:
        k = ref_kernel(name, **kw)
        set_sm(k, sm)
        t = tag(name, kw)
        out[f"kidx_{t}"] = np.asarray(k.kernel_indices)
        out[f"size_in_fwhm_{t}"] = np.asarray(k.size_in_fwhm)
        out[f"rescale_{t}"] = np.asarray(k._rescale)
        out[f"sm_ranges_{t}"] = np.asarray(k.sm_ranges)
    np.savez_compressed(os.path.join(HERE, "adaptive.npz"), **out)


class _PerParticleSigma(S.GaussianSpectrum):
    """GaussianSpectrum whose per-particle widths are handed over already in km/s (the
    reference's ``"thermal"`` mode computes them with an m/s -> km/s unit conversion the
    scale-1 shim cannot express; the formula itself is a KAT in test_oracle_kats.py)."""

    def half_width(self, source):
        return source._sigma


def make_spectrum(kind, sigma, dtype=np.float64, ncpu=None):
    if kind == "dirac":
        return S.DiracDeltaSpectrum(spec_dtype=dtype, ncpu=ncpu)
    if np.ndim(sigma) == 0:
        return S.GaussianSpectrum(sigma=Q(float(sigma)), spec_dtype=dtype, ncpu=ncpu)
    return _PerParticleSigma(sigma="per-particle", spec_dtype=dtype, ncpu=ncpu)


def gen_spectra():
    rng = np.random.Generator(np.random.PCG64(103))
    n, C = 300, 24
    edges_dec = 100.0 - 4.0 * np.arange(C + 1)  # decreasing, as DataCube produces
    edges_inc = edges_dec[::-1].copy()
    v = np.r_[rng.uniform(-20.0, 120.0, n - 4), [100.0, 96.0, 4.0, 50.0]]  # some exactly on edges
    sig = rng.uniform(2.0, 15.0, n)
    mHI = rng.uniform(0.5, 2.0, n) * 1.0e5
    D = rng.uniform(3.0, 4.0, n)
    out = {"edges_dec": edges_dec, "edges_inc": edges_inc, "v": v, "sigma": sig, "mHI": mHI, "D": D}
    for ename, edges in (("dec", edges_dec), ("inc", edges_inc)):
        dc = SimpleNamespace(velocity_channel_edges=Q(edges))
        for sname, kind, sigma, dtype, ncpu in (
            ("gauss7", "gauss", 7.0, np.float64, None),
            ("gaussP", "gauss", sig, np.float64, None),
            ("gaussP_ncpu3", "gauss", sig, np.float64, 3),
            ("gauss7_f32", "gauss", 7.0, np.float32, None),
            ("dirac", "dirac", None, np.float64, None),
        ):
            src = FakeSource(np.zeros((3, n)), mHI, v, D, sigma=sigma if np.ndim(sigma) else None)
            sm = make_spectrum(kind, sigma, dtype, ncpu)
            sm.init_spectra(src, dc)
            out[f"spectra_{sname}_{ename}"] = np.asarray(sm.spectra)
    np.savez_compressed(os.path.join(HERE, "spectra.npz"), **out)


def make_martini(source, datacube, kernel, spectrum):
    m = object.__new__(M._BaseMartini)
    m.quiet = True
    m.source = source
    m._datacube = datacube
    m.beam = None
    m.noise = None
    m.sph_kernel = kernel
    m.spectral_model = spectrum
    return m


def gen_prune():
    rng = np.random.Generator(np.random.PCG64(104))
    n = 2000
    nx, ny, nc, pad = 6, 5, 8, 3
    edges = 40.0 - 5.0 * np.arange(nc + 1)
    X, Y = nx + 2 * pad, ny + 2 * pad
    px = rng.uniform(-8.0, X + 8.0, n)
    py = rng.uniform(-8.0, Y + 8.0, n)
    pz = rng.uniform(-6.0, nc + 6.0, n)
    # exact-threshold and NaN cases
    px[:6] = [-2.0, X + 2.0, np.nan, 3.0, 3.0, 3.0]
    py[:6] = [3.0, 3.0, 3.0, -2.0, Y + 2.0, np.nan]
    sm = rng.uniform(0.4, 1.6, n)
    sm[:6] = 1.4  # sm_range = ceil(1.4*1.384) = 2 for the cubic spline
    mHI = np.where(rng.uniform(size=n) < 0.1, 0.0, 1.0e4)
    sig = rng.uniform(1.0, 9.0, n)
    out = {"px": px, "py": py, "pz": pz, "sm_lengths": sm, "mHI": mHI, "sigma": sig,
           "edges": edges, "shape": np.array([nx, ny, nc, pad])}
    for sname, kind, sigma in (("gauss3", "gauss", 3.0), ("gaussP", "gauss", sig), ("dirac", "dirac", None)):
        # flags == 0 (nothing pruned) makes the reference hand a 0-d mask to apply_mask,
        # which raises for N > 1 (sph_source.py:374-375); nothing to record there
        for flags in range(1, 8):
            spatial, spectral, mass = bool(flags & 1), bool(flags & 2), bool(flags & 4)
            src = FakeSource(np.vstack((px, py, pz)), mHI, np.zeros(n), np.ones(n),
                             sigma=sigma if np.ndim(sigma) else None)
            k = ref_kernel("_CubicSplineKernel")
            set_sm(k, sm)
            out.setdefault("sm_ranges", np.asarray(k.sm_ranges))
            m = make_martini(src, FakeDataCube(nx, ny, nc, pad, edges, 1.0), k, make_spectrum(kind, sigma))
            # recover the accept mask through a tracer array pruned alongside the rest
            tracer = {}
            orig_apply = src.apply_mask
            src.apply_mask = lambda mask, _t=tracer, _o=orig_apply: (_t.setdefault("mask", np.asarray(mask).copy()), _o(mask))
            m._prune_particles(spatial=spatial, spectral=spectral, mass=mass)
            out[f"accept_{sname}_{flags}"] = tracer["mask"].astype(np.bool_)
    np.savez_compressed(os.path.join(HERE, "prune.npz"), **out)


def gen_insert(dtype=np.float64, prefix="insert", only=None):
    """``dtype`` is the spectral model's ``spec_dtype``; the float32 fixtures (prefix
    "insertf32", two cases) pin what the reference produces in its float32 mode."""
    nx, ny, nc, pad = 9, 7, 12, 2
    X, Y = nx + 2 * pad, ny + 2 * pad
    px_size = 3.0
    n = 90
    cases = []
    for name, kw in PRIMS + ADAPTIVE:
        for sname in ("gauss7", "gaussP", "dirac"):
            cases.append((name, kw, sname))
    for ic, (name, kw, sname) in enumerate(cases):
        if only is not None and (name, sname) not in only:
            continue
        rng = np.random.Generator(np.random.PCG64(1000 + ic))
        edges = 30.0 - 5.0 * np.arange(nc + 1)
        if ic % 4 == 3:
            edges = edges[::-1].copy()  # increasing edges
        px = rng.uniform(-3.0, X + 3.0, n)
        py = rng.uniform(-3.0, Y + 3.0, n)
        v = rng.uniform(-50.0, 50.0, n)
        # pixel-centred / pixel-edge / channel-edge specials
        px[:4] = [4.0, 4.5, 6.0, 2.0]
        py[:4] = [5.0, 5.0, 3.5, 2.0]
        v[:4] = [0.0, 5.0, -10.0, 2.5]
        pz = (30.0 - v) / 5.0 if edges[0] > edges[-1] else (v + 30.0) / 5.0
        sm = rng.lognormal(np.log(1.3), 0.7, n)
        if name == "DiracDeltaKernel":
            sm = np.minimum(sm, 3.0)
        mHI = rng.uniform(0.5, 2.0, n) * 1.0e6
        mHI[5] = 0.0
        D = rng.uniform(3.0, 3.2, n)
        sig = rng.uniform(2.0, 9.0, n)
        sigma = {"gauss7": 7.0, "gaussP": sig, "dirac": None}[sname]
        kind = "dirac" if sname == "dirac" else "gauss"
        initial = rng.normal(0.0, 1.0e-3, (X, Y, nc)) if ic % 5 == 0 else None
        src = FakeSource(np.vstack((px, py, pz)), mHI, v, D, sigma=sigma if np.ndim(sigma) else None)
        dc = FakeDataCube(nx, ny, nc, pad, edges, px_size, initial=initial)
        k = ref_kernel(name, **kw)
        set_sm(k, sm)
        sm_ranges0 = np.asarray(k.sm_ranges).copy()
        m = make_martini(src, dc, k, make_spectrum(kind, sigma, dtype=dtype))
        tracer = {}
        orig_apply = src.apply_mask
        src.apply_mask = lambda mask, _t=tracer, _o=orig_apply: (_t.setdefault("mask", np.asarray(mask).copy()), _o(mask))
        m._prune_particles()
        ncpu = 2 if ic % 7 == 0 else 1
        m._insert_source_in_cube(skip_validation=True, progressbar=False, ncpu=ncpu)
        out = {
            "px": px, "py": py, "pz": pz, "v": v, "sm_lengths": sm, "mHI": mHI, "D": D,
            "sigma": np.asarray(7.0 if sigma is None else sigma), "edges": edges,
            "shape": np.array([nx, ny, nc, pad]), "px_size": np.float64(px_size),
            "accept": tracer["mask"].astype(np.bool_), "sm_ranges": sm_ranges0,
            "cube": np.asarray(m._datacube._array),
            "initial": np.zeros(0) if initial is None else initial,
            "kernel": np.array(name), "truncate": np.float64(kw.get("truncate", 0.0)),
            "spectrum": np.array(sname),
        }
        np.savez_compressed(os.path.join(HERE, f"{prefix}_{tag(name, kw)}_{sname}.npz"), **out)


def gen_seam():
    """The two seam functions whose arithmetic IS a unit conversion, run unmodified under the
    scaled-unit stand-in (oracle/refshim.py: load_reference_scaled):
    _BaseSPHKernel._init_sm_lengths (sph_kernels.py:235-255) -- arctan(hsm / D) in pixels,
    through _AdaptiveKernel._init_sm_lengths (:1241-1274) as well, so that the adaptive
    selection is pinned end to end from kpc / Mpc / arcsec inputs -- and
    GaussianSpectrum.half_width with sigma="thermal" (spectral_models.py:465-485)."""
    from oracle.refshim import load_reference_scaled

    Ks, Ss, Us = load_reference_scaled()
    Qs = Us.Quantity
    rng = np.random.Generator(np.random.PCG64(20260107))
    n = 4000
    hsm = np.r_[10 ** rng.uniform(-2.5, 1.5, n - 4), 0.0, 1e-6, 1.0, 100.0]  # kpc
    dist = np.r_[rng.uniform(0.5, 80.0, n - 4), 3.0, 3.0, 3.0, 0.05]  # Mpc, per particle
    out = {"hsm_kpc": hsm, "distance_Mpc": dist}
    for px_size in (10.0, 3.0, 0.7):
        src = SimpleNamespace(
            hsm_g=Qs(hsm, Us.kpc), mHI_g=Qs(np.ones(n), Us.Msun),
            skycoords=SimpleNamespace(transform_to=lambda frame: SimpleNamespace(distance=Qs(dist, Us.Mpc))))
        dc = SimpleNamespace(px_size=Qs(px_size, Us.arcsec), coordinate_frame=None)
        t = f"px{px_size:g}".replace(".", "p")
        k = Ks._WendlandC2Kernel()
        k._init_sm_lengths(src, dc)
        assert k.sm_lengths.unit == Us.pix
        out[f"sm_lengths_{t}"] = np.asarray(k.sm_lengths.value)
        for name in ("WendlandC2Kernel", "CubicSplineKernel", "GaussianKernel"):
            ka = getattr(Ks, name)()
            ka._init_sm_lengths(source=src, datacube=dc)
            ka._init_sm_ranges()
            assert np.array_equal(np.asarray(ka.sm_lengths.value), out[f"sm_lengths_{t}"])
            out[f"kidx_{name}_{t}"] = np.asarray(ka.kernel_indices)
            out[f"sm_ranges_{name}_{t}"] = np.asarray(ka.sm_ranges.value)
    T = np.r_[10 ** rng.uniform(1.0, 7.0, n - 2), 1.0e4, 8.0e3]
    g = Ss.GaussianSpectrum(sigma="thermal")
    hw = g.half_width(SimpleNamespace(T_g=Qs(T, Us.K)))
    assert hw.unit == Us.km / Us.s
    out["T_K"] = T
    out["half_width_thermal_kms"] = np.asarray(hw.value)
    np.savez_compressed(os.path.join(HERE, "seam.npz"), **out)


if __name__ == "__main__":
    if "--seam-only" in sys.argv:
        gen_seam()
        sys.exit(0)
    gen_kernels()
    gen_adaptive()
    gen_spectra()
    gen_prune()
    gen_seam()
    gen_insert()
    gen_insert(dtype=np.float32, prefix="insertf32",
               only={("WendlandC2Kernel", "gauss7"), ("CubicSplineKernel", "gaussP")})
    tot = sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith(".npz"))
    print("golden fixtures written:", tot // 1024, "KiB")
