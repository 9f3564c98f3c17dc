"""``Martini`` / ``GlobalProfile`` with the reference's constructor and method signatures
(martini/martini.py), running the particle->datacube projection on the GPU.

What changes relative to the reference is *where* four things run:

* ``_prune_particles``            (martini.py:168-241)  -> ``mtn_prune``            (K1)
* ``sph_kernel._init_sm_ranges`` / adaptive selection  -> ``mtn_smoothing_setup``  (K0)
* ``spectral_model.init_spectra`` (spectral_models.py:63-147), ``sph_kernel._px_weight``
  (sph_kernels.py:85-119) and the pixel loop of ``_insert_source_in_cube``
  (martini.py:285-366)                                  -> ``mtn_plan`` + ``mtn_project``

Everything else keeps the reference's semantics: the constructor pads the cube for the beam,
initialises coordinates, prunes (raising ``RuntimeError("No non-zero mHI source particles in
target region.")`` when nothing is left) and applies the mask to the host objects;
``insert_source_in_cube`` validates the kernel (``RuntimeError`` ... "use this with care"
unless ``skip_validation``), *adds* to whatever the cube holds and converts Jy/pix^2 to
Jy/arcsec^2.  ``spectral_model.spectra`` stays ``None`` unless ``init_spectra()`` is called:
the N x C spectra array is never needed.  Kernels or spectral models that are not the
built-in classes raise ``NotImplementedError`` -- there is no CPU fallback.
"""

from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from .datacube import DataCube
from .engine import Engine
from .spectral_models import check_monotonic, spectrum_kind
from .sph_kernels import DiracDeltaKernel, kernel_table


class _BaseMartini:
    """martini.py:47-624 (hot-path part)."""

    def __init__(self, *, source, datacube, beam=None, noise=None, sph_kernel, spectral_model,
                 quiet=False, _prune_kwargs=None, device="cuda:0", engine=None):
        self.quiet = quiet
        self.source = source
        self._datacube = datacube
        self.beam = beam
        self.noise = noise
        self.sph_kernel = sph_kernel
        self.spectral_model = spectral_model
        # unsupported plug-ins fail here, before any work (no CPU fallback)
        self._table = kernel_table(sph_kernel)
        self._spectrum = spectrum_kind(spectral_model)
        self.engine = engine or Engine(device)
        sph_kernel._engine = self.engine

        if self.beam is not None:
            self.beam.init_kernel(self._datacube)
            self._datacube.add_pad(self.beam.needs_pad())

        # coordinates: one fused kernel on the device where the source offers it (SPHSource:
        # mtn_sky_to_pix), the host mirror of the reference's steps otherwise
        front = getattr(self.source, "_init_on_device", None)
        dev_coords = front(self.engine, self._datacube) if front is not None else None  # after datacube is padded
        if dev_coords is None:
            self.source._init_skycoords()
            self.source._init_pixcoords(self._datacube)
        self._init_device_particles(dev_coords)
        self._prune_particles(**(_prune_kwargs or {}))

    # ------------------------------------------------------------------ device state
    def _init_device_particles(self, dev_coords=None):
        """Upload the seam arrays (or take them from the device front-end) and run K0
        (sph_kernels.py:235-262, 1241-1274).  Everything the projection needs stays on the device
        from here on."""
        eng, src = self.engine, self.source
        scalar_or_dev = lambda x: eng.to_device(x) if np.ndim(x) > 0 else float(x)  # noqa: E731
        if dev_coords is not None:
            self._dev = dict(dev_coords)
        else:
            self._dev = {
                "px": eng.to_device(src.pixcoords[0]), "py": eng.to_device(src.pixcoords[1]),
                "pz": eng.to_device(src.pixcoords[2]),
                "sm_length": eng.to_device(src.sm_lengths_px(self._datacube)),
                "v": eng.to_device(src.radial_velocity), "D": scalar_or_dev(src.distance_p),
            }
        self._dev["mHI"] = scalar_or_dev(src.mHI_g)
        self._dev["sigma"] = scalar_or_dev(self.spectral_model.half_width(src))
        kid, valid, sm_range, h_eff = eng.smoothing_setup(self._dev["sm_length"], self._table)
        self._dev.update(kernel_id=kid, valid=valid, sm_range=sm_range, h_eff=h_eff)
        # host copies of the kernel state are fetched when somebody reads them (sph_kernels.py)
        self.sph_kernel._set_device_state(self._dev["sm_length"], sm_range, kid, valid)

    def _mass_on_device(self, accept=None):
        m = self._dev["mHI"]
        n = self._dev["px"].numel()
        if not isinstance(m, torch.Tensor):
            return m * (n if accept is None else int(accept.sum()))
        return float(m.sum() if accept is None else (m * accept).sum())

    def _prune_particles(self, spatial=True, spectral=True, mass=True, obj_type_str="data cube"):
        """martini.py:168-241; the accept mask is computed on the GPU and stays there for the
        projection.  The host objects are pruned exactly like the reference's (:233-234), but
        lazily: ``source`` and ``sph_kernel`` apply the mask when their arrays are first read."""
        if not self.quiet:
            print(f"Source module contained {self.source.npart} particles with total HI mass of "
                  f"{self._mass_on_device():.2e} Msun.")
        dc, d = self._datacube, self._dev
        edges = dc.velocity_channel_edges
        check_monotonic(edges)
        nx_tot, ny_tot = dc.n_px_x + 2 * dc.padx, dc.n_px_y + 2 * dc.pady
        accept, n_accept = self.engine.prune(d["px"], d["py"], d["pz"], d["sm_range"], d["mHI"], d["sigma"],
                                             float(np.max(np.abs(np.diff(edges)))), nx_tot, ny_tot,
                                             dc.n_channels, spatial, spectral, mass)
        self._dev["accept"] = accept
        n_kept = int(n_accept)  # the one scalar that comes back; raises below if nothing is left
        self.source._defer_mask(lambda: accept.cpu().numpy().astype(bool), n_kept)
        self.sph_kernel._apply_mask(accept)
        if not self.quiet:
            print(f"Pruned particles that will not contribute to {obj_type_str}, "
                  f"{n_kept} particles remaining with total HI mass of "
                  f"{self._mass_on_device(accept):.2e} Msun.")

    # ------------------------------------------------------------------ public methods
    def init_spectra(self):
        """martini.py:153-166.  Optional here: insertion never needs the N x C array."""
        if not self.quiet:
            print("Initializing spectra...")
        self.spectral_model.init_spectra(self.source, self._datacube, engine=self.engine)
        if not self.quiet:
            print("Spectra initialized.")

    def _insert_source_in_cube(self, skip_validation=False, progressbar=None, ncpu=1, quiet=None):
        """martini.py:285-407.  ``progressbar`` and ``ncpu`` are accepted and ignored."""
        dc = self._datacube
        if dc.array_unit != "Jy/pix2":
            raise RuntimeError("The data cube already holds an inserted source; call reset() first.")
        self.sph_kernel._confirm_validation(noraise=skip_validation, quiet=self.quiet)
        eng, d = self.engine, self._dev
        zero = dc._array_is_zero
        if zero:
            nx, ny = dc.n_px_x + 2 * dc.padx, dc.n_px_y + 2 * dc.pady
            cube = torch.zeros((nx, ny, dc.n_channels), dtype=torch.float64, device=eng.device)
        else:
            cube = dc._device_array(eng)
        edges_host = np.asarray(dc.velocity_channel_edges)
        edges = eng.to_device(edges_host)
        gauss = self._spectrum == L.SPECTRUM_GAUSSIAN
        plan = eng.insert(px=d["px"], py=d["py"], h_eff=d["h_eff"], sm_range=d["sm_range"], v=d["v"],
                          kernel_id=d["kernel_id"], sigma=d["sigma"] if gauss else 1.0,
                          mHI=d["mHI"], D=d["D"], accept=d["accept"], table=self._table,
                          spectrum=self._spectrum, edges=edges, cube=cube,
                          px_size_arcsec=dc.px_size, zeroed=zero,
                          edges_increasing=bool(edges_host[1] > edges_host[0]))
        dc._set_device_array(cube, eng)  # stays on the GPU until someone reads datacube._array
        dc.array_unit = "Jy/arcsec2"
        self.last_plan = plan
        if (quiet is None and not self.quiet) or (quiet is not None and not quiet):
            self._print_summary()

    def _print_summary(self):
        """martini.py:367-406, evaluated where the cube lives (no device -> host copy)."""
        dc = self._datacube
        a = dc._device_array(self.engine)
        full = a
        if dc.padx > 0 and dc.pady > 0:
            a = a[dc.padx:-dc.padx, dc.pady:-dc.pady]
        flux = float(a.sum()) * dc.px_size**2
        dv = torch.as_tensor(np.abs(np.diff(dc.velocity_channel_edges)), device=a.device)
        mass = 2.36e5 * self.source.distance**2 * float((a.sum(dim=(0, 1)) * dc.px_size**2 * dv).sum())
        nz = full[full > 0]
        print("Source inserted.",
              f"  Flux density in cube: {flux:.2e} Jy",
              f"  Mass in cube (assuming distance {self.source.distance:.2f} Mpc and a spatially"
              f" resolved source): {mass:.2e} Msun",
              f"    [{mass / self.source.input_mass * 100:.0f}% of initial source mass]",
              f"  Maximum pixel: {float(full.max()):.2e} Jy / arcsec2",
              f"  Median non-zero pixel: {float(nz.median()) if nz.numel() else 0.0:.2e} Jy / arcsec2",
              sep="\n")

    def reset(self):
        """martini.py:409-425."""
        dc = self._datacube
        new = DataCube(n_px_x=dc.n_px_x, n_px_y=dc.n_px_y, n_channels=dc.n_channels,
                       px_size=dc.px_size, channel_width=dc.channel_width, channel_unit=dc.channel_unit,
                       spectral_centre=dc.spectral_centre, spectral_centre_unit=dc.channel_unit,
                       ra=dc.ra, dec=dc.dec, stokes_axis=dc.stokes_axis)
        if self.beam is not None:
            # the reference re-pads with the beam's own requirement (martini.py:423-424), not
            # with the current pad: convolve_beam() has dropped that one
            new.add_pad(self.beam.needs_pad())
        self._datacube = new


class Martini(_BaseMartini):
    """martini.py:627-861."""

    @property
    def datacube(self):
        return self._datacube

    def insert_source_in_cube(self, skip_validation=False, progressbar=None, ncpu=1):
        """Populate the DataCube with flux from the source particles (martini.py:827-861)."""
        self._insert_source_in_cube(skip_validation=skip_validation, progressbar=progressbar, ncpu=ncpu)


    def add_noise(self):
        """Insert noise into the data cube (martini.py:903-937).  The realisation is the
        reference's (``noise.generate``: numpy Generator, same seed -> same numbers); it is
        converted from Jy/beam to the cube's current unit and added on the device."""
        from warnings import warn

        if self.noise is None:
            warn("Skipping noise, no noise object provided to Martini.")
            return
        if self.beam is None:
            warn("Skipping noise, no beam object (required to estimate post-convolution"
                 " noise level) provided to Martini.")
            return
        dc, eng = self._datacube, self.engine
        if dc.array_unit not in ("Jy/pix2", "Jy/arcsec2"):
            raise RuntimeError("add_noise expects a cube in Jy/pix2 or Jy/arcsec2 (before convolve_beam).")
        noise = self.noise.generate(dc, self.beam)  # Jy/beam, shape of datacube._array
        # Jy/beam -> Jy/arcsec2 (beam_angular_area) -> the cube's unit (arcsec2_to_pix)
        factor = 1.0 / self.beam.area * (dc.px_size**2 if dc.array_unit == "Jy/pix2" else 1.0)
        cube = dc._device_array(eng)
        noise_dev = eng.to_device(np.ascontiguousarray(noise.reshape(noise.shape[:3])))
        cube.add_(noise_dev, alpha=factor)
        dc._set_device_array(cube)
        if not self.quiet:
            print("Noise added.",
                  f"  Noise cube RMS: {float(noise_dev.std(correction=0)) * factor:.2e} (before beam convolution).",
                  "  Data cube RMS after noise addition (before beam convolution): "
                  f"{float(cube.std(correction=0)):.2e}", sep="\n")

    def convolve_beam(self):
        """Convolve the cube with the beam, drop the pad, convert to Jy/beam
        (martini.py:863-901), on the GPU (``mtn_convolve_beam``)."""
        from warnings import warn

        if self.beam is None:
            warn("Skipping beam convolution, no beam object provided to Martini.")
            return
        dc = self._datacube
        need = self.beam.needs_pad()
        if dc.padx < need[0] or dc.pady < need[1]:
            raise ValueError(
                "datacube padding insufficient for beam convolution (perhaps you loaded a"
                " datacube state with datacube.load_state that was previously initialized"
                " by martini with a smaller beam?)")
        if dc.array_unit != "Jy/arcsec2":
            raise RuntimeError("convolve_beam expects a cube in Jy/arcsec2: insert the source first.")
        eng = self.engine
        out = eng.convolve_beam(dc._device_array(eng), self.beam.kernel, scale=self.beam.area)  # x area: -> Jy/beam
        # drop_pad (martini.py:899, datacube.py:710-729) on the device
        out = out[dc.padx:dc.padx + dc.n_px_x, dc.pady:dc.pady + dc.n_px_y, :].contiguous()
        dc._set_device_array(out)
        dc.padx = dc.pady = 0
        dc.array_unit = "Jy/beam"
        if not self.quiet:
            nz = dc._array[dc._array > 0]
            print("Beam convolved.",
                  f"  Data cube RMS after beam convolution: {np.std(dc._array):.2e} Jy / beam",
                  f"  Maximum pixel: {dc._array.max():.2e} Jy / beam",
                  f"  Median non-zero pixel: {np.median(nz) if nz.size else 0.0:.2e} Jy / beam", sep="\n")


class GlobalProfile(_BaseMartini):
    """Spatially integrated spectrum (martini.py:1369-1832): a 1 x 1 pixel cube, all particles
    at pixel (0, 0), ``DiracDeltaKernel(size_in_fwhm=inf)``; pruning by velocity only."""

    def __init__(self, *, source, spectral_model, n_channels=64, channel_width=4.0,
                 spectral_centre=0.0, quiet=False, device="cuda:0", engine=None, channel_unit=None,
                 spectral_centre_unit=None):
        self._source_in = source
        dc = DataCube(n_px_x=1, n_px_y=1, n_channels=n_channels, px_size=1.0,
                      channel_width=channel_width, spectral_centre=spectral_centre, channel_unit=channel_unit,
                      spectral_centre_unit=spectral_centre_unit,
                      ra=source.ra if hasattr(source, "ra") else 0.0,
                      dec=source.dec if hasattr(source, "dec") else 0.0)
        self._inserted = False
        super().__init__(source=_AtOrigin(source), datacube=dc, beam=None, noise=None,
                         sph_kernel=DiracDeltaKernel(size_in_fwhm=np.inf), spectral_model=spectral_model,
                         quiet=quiet, _prune_kwargs={"spatial": False, "obj_type_str": "spectrum"},
                         device=device, engine=engine)

    def insert_source_in_spectrum(self):
        """martini.py:1551-1596."""
        self._insert_source_in_cube(skip_validation=True, quiet=True)
        self._inserted = True

    @property
    def spectrum(self):
        """Jy per channel (martini.py:1598-1614)."""
        if not self._inserted:
            self.insert_source_in_spectrum()
        dc = self._datacube
        return dc._array.reshape(-1)[:dc.n_channels] * dc.px_size**2

    @property
    def channel_mids(self):
        return self._datacube.velocity_channel_mids

    @property
    def channel_edges(self):
        return self._datacube.velocity_channel_edges


class _AtOrigin:
    """Wrap a source so that every particle sits at pixel (0, 0) (martini.py:1541-1549)."""

    _init_on_device = None  # (host front-end: the pixel coordinates are overwritten below)

    def __init__(self, source):
        self._s = source

    def __getattr__(self, name):
        return getattr(self._s, name)

    def _init_skycoords(self):
        self._s._init_skycoords()

    def _init_pixcoords(self, datacube):
        self._s._init_pixcoords(datacube)
        self._s.pixcoords[:2] = 0.0

    def sm_lengths_px(self, datacube):
        # any positive length: sm_range = ceil(length * inf) = inf reaches the single pixel
        # (martini.py:1537); the Dirac-delta weight itself ignores the smoothing length
        return np.ones(self._s.npart)

    def apply_mask(self, mask):
        self._s.apply_mask(mask)


class _PadOnlyBeam:
    """Stand-in for the (out of scope) beam classes: only the pad size matters to the hot path
    (beams.py:88-101, 277-295)."""

    def __init__(self, pad):
        self._pad = int(pad)

    def init_kernel(self, datacube):
        pass

    def needs_pad(self):
        return (self._pad, self._pad)


def demo(quiet=False, device="cuda:0", convolve=False):
    """The hot-path slice of the reference's ``demo()`` (martini/_demo.py:95-161): the demo
    source into the demo cube with CubicSplineKernel + GaussianSpectrum(7 km/s).  The 30 arcsec
    Gaussian beam truncated at 4 sigma pads the cube by ceil(30*4/10 + 1) = 13 pixels; noise,
    and FITS output are outside this package's scope; ``convolve=True`` also runs the beam
    convolution with the demo's 30 arcsec Gaussian beam."""
    from .sources import demo_source
    from .spectral_models import GaussianSpectrum
    from .sph_kernels import CubicSplineKernel

    from .beams import GaussianBeam

    source = demo_source()
    datacube = DataCube(n_px_x=128, n_px_y=128, n_channels=32, px_size=10.0, channel_width=10.0,
                        spectral_centre=source.vsys)
    beam = GaussianBeam(bmaj=30.0, bmin=30.0, bpa=0.0, truncate=4.0) if convolve else _PadOnlyBeam(13)
    m = Martini(source=source, datacube=datacube, beam=beam, noise=None,
                spectral_model=GaussianSpectrum(sigma=7.0), sph_kernel=CubicSplineKernel(),
                quiet=quiet, device=device)
    m.insert_source_in_cube()
    if convolve:
        m.convolve_beam()
    return m
