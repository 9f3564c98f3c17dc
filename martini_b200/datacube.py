"""``DataCube`` with the reference's constructor and array/pad/channel semantics
(martini/datacube.py), astropy-free.

Only what the projection path needs is mirrored: the float64 array
``(n_px_x + 2 padx, n_px_y + 2 pady, n_channels[, 1])`` in Jy/pix^2 (channel fastest,
datacube.py:183-186), ``add_pad`` / ``drop_pad`` (:669-729), the velocity channel edges
(:551-562) and the simple TAN / VRAD world coordinate system (:426-486) in closed form.
FITS/HDF5 I/O and ``from_wcs`` are out of scope.  Both channel modes exist: ``channel_width`` in
km/s (velocity channels) or, with ``channel_unit="Hz"`` or an astropy frequency, in Hz.

Units are fixed: ``px_size`` arcsec, ``channel_width`` and ``spectral_centre`` km/s, ``ra`` /
``dec`` degrees.  astropy Quantities are accepted and converted if astropy is installed.
"""

from __future__ import annotations

import numpy as np


def _value(x, unit):
    """Strip an astropy Quantity to ``unit`` (a string), or pass plain numbers through."""
    if hasattr(x, "to_value"):
        import astropy.units as U  # only reachable when the caller already uses astropy

        return x.to_value(U.Unit(unit))
    return x


HI_FREQ_HZ = 1.420405751e9   # datacube.py:21
C_KMS = 299792.458           # astropy.constants.c


def _is_frequency(x, unit):
    """Whether a channel width is a frequency: an astropy Quantity says so itself, a plain
    number goes by ``unit`` ("km/s" or "Hz")."""
    if hasattr(x, "unit"):
        import astropy.units as U

        return U.get_physical_type(x) == "frequency"
    if unit not in (None, "km/s", "Hz"):
        raise ValueError("Channel width must have frequency or velocity units.")
    return unit == "Hz"


class DataCube:
    def __init__(self, *, n_px_x, n_px_y, n_channels, px_size, channel_width,
                 spectral_centre=0.0, ra=0.0, dec=0.0, stokes_axis=False,
                 coordinate_frame=None, specsys="icrs", channel_unit=None, spectral_centre_unit=None):
        if coordinate_frame is not None or specsys.lower() != "icrs":
            raise NotImplementedError("martini_b200.DataCube supports the ICRS frame / specsys only")
        self.stokes_axis = stokes_axis
        self.n_px_x, self.n_px_y, self.n_channels = int(n_px_x), int(n_px_y), int(n_channels)
        self.px_size = float(_value(px_size, "arcsec"))
        # channel mode (datacube.py:196-207): a channel width in Hz gives a FREQ axis (channels
        # of equal frequency width, frequency increasing with index), in km/s a VRAD axis
        # (velocity decreasing with index); the spectral centre is converted to the channel
        # unit with the radio Doppler convention at the HI rest frequency
        self._freq_channel_mode = _is_frequency(channel_width, channel_unit)
        unit = "Hz" if self._freq_channel_mode else "km/s"
        self.channel_unit = unit
        self.channel_width = abs(float(_value(channel_width, unit)))
        centre_is_freq = _is_frequency(spectral_centre, spectral_centre_unit) if (
            hasattr(spectral_centre, "unit") or spectral_centre_unit is not None) else False
        centre = float(_value(spectral_centre, "Hz" if centre_is_freq else "km/s"))
        if centre_is_freq and not self._freq_channel_mode:
            centre = C_KMS * (1.0 - centre / HI_FREQ_HZ)
        elif self._freq_channel_mode and not centre_is_freq:
            centre = HI_FREQ_HZ * (1.0 - centre / C_KMS)
        self.spectral_centre = centre
        self.ra = float(_value(ra, "deg"))
        self.dec = float(_value(dec, "deg"))
        self.padx = self.pady = 0
        self._dev = None  # device copy (torch, 3-D); authoritative while _host is None
        self._host = None  # host array; a cube that is known to be all zeros has neither copy
        self._known_zero = True
        #: "Jy/pix2" until insert_source_in_cube converts to "Jy/arcsec2" (martini.py:364-366)
        self.array_unit = "Jy/pix2"

    # ------------------------------------------------------------------ array residency
    # Between insert_source_in_cube, add_noise and convolve_beam the cube stays on the GPU; the
    # host array of the reference (`DataCube._array`) is materialised on first access.  A
    # host access also drops the device copy, because the caller may modify the array in
    # place (the reference's own tests do).
    def _shape(self):
        return (self.n_px_x + 2 * self.padx, self.n_px_y + 2 * self.pady, self.n_channels) + (
            (1,) if self.stokes_axis else ())

    @property
    def _array(self):
        if self._host is None:
            if self._dev is None:
                self._host = np.zeros(self._shape())
            else:
                eng = self.__dict__.get("_engine")
                a = eng.to_host(self._dev) if eng is not None else self._dev.cpu().numpy()
                self._host = a[..., np.newaxis] if self.stokes_axis else a
        self._dev = None
        self._known_zero = False  # the caller holds a writable reference from here on
        return self._host

    @_array.setter
    def _array(self, value):
        self._host = value
        self._dev = None
        self._known_zero = False

    def _device_array(self, engine):
        """The cube as a 3-D device tensor (uploaded from the host copy if needed)."""
        if self._dev is None:
            if self._host is None:
                import torch

                self._dev = torch.zeros(self._shape()[:3], dtype=torch.float64, device=engine.device)
            else:
                h = self._host
                self._dev = engine.to_device(np.ascontiguousarray(h.reshape(h.shape[:3])))
        return self._dev

    def _set_device_array(self, tensor, engine=None):
        """Make ``tensor`` (3-D, on the device) the cube's contents; the host copy is stale.
        ``engine``: whose page-locked buffer pool a later host read uses."""
        if engine is not None:
            self._engine = engine
        self._dev = tensor
        self._host = None
        self._known_zero = False

    @property
    def _array_is_zero(self):
        """True if the cube holds only zeros (checked where the data lives; a cube nobody has
        touched since construction / reset is known to be zero without looking)."""
        if self._known_zero:
            return True
        if self._host is None:
            return not bool(self._dev.any())
        return not self._host.any()

    # ------------------------------------------------------------------ channels
    @property
    def channel_edges(self):
        """(C+1,) channel edges in the cube's own spectral unit (datacube.py:513-532): the
        spectral axis has crpix = C/2 + 0.5 and cdelt = +|channel_width| (FREQ) or
        -|channel_width| (VRAD) (:469-481), edges at pixel k - 1/2."""
        k = np.arange(self.n_channels + 1)
        sign = 1.0 if self._freq_channel_mode else -1.0
        return self.spectral_centre + sign * self.channel_width * (k - self.n_channels / 2.0)

    @property
    def channel_mids(self):
        e = self.channel_edges
        return 0.5 * (e[1:] + e[:-1])

    @property
    def velocity_channel_edges(self):
        """(C+1,) channel edges in km/s (datacube.py:551-562); decreasing with channel index in
        both modes (radio convention: v = c (1 - f / f_HI))."""
        e = self.channel_edges
        return C_KMS * (1.0 - e / HI_FREQ_HZ) if self._freq_channel_mode else e

    @property
    def velocity_channel_mids(self):
        m = self.channel_mids
        return C_KMS * (1.0 - m / HI_FREQ_HZ) if self._freq_channel_mode else m

    @property
    def frequency_channel_edges(self):
        e = self.channel_edges
        return e if self._freq_channel_mode else HI_FREQ_HZ * (1.0 - e / C_KMS)

    @property
    def frequency_channel_mids(self):
        m = self.channel_mids
        return m if self._freq_channel_mode else HI_FREQ_HZ * (1.0 - m / C_KMS)

    # ------------------------------------------------------------------ pad
    def add_pad(self, pad):
        """datacube.py:669-708."""
        if self.padx > 0 or self.pady > 0:
            raise RuntimeError("Tried to add padding to already padded datacube array.")
        px, py = int(pad[0]), int(pad[1])
        if self._known_zero:  # nothing to copy: the padded cube is all zeros too
            self._host = self._dev = None
            self.padx, self.pady = px, py
            return
        shape = (self.n_px_x + 2 * px, self.n_px_y + 2 * py, self.n_channels)
        new = np.zeros(shape + ((1,) if self.stokes_axis else ()))
        new[px:px + self.n_px_x, py:py + self.n_px_y, ...] = self._array
        self._array = new
        self.padx, self.pady = px, py

    def drop_pad(self):
        """datacube.py:710-729."""
        if self.padx == 0 and self.pady == 0:
            return
        self._array = self._array[self.padx:self.padx + self.n_px_x, self.pady:self.pady + self.n_px_y, ...]
        self.padx = self.pady = 0

    # ------------------------------------------------------------------ world -> pixel
    def world2pix(self, ra_deg, dec_deg, v_kms):
        """0-indexed pixel coordinates (x, y, channel) of sky positions and radial velocities:
        the RA---TAN / DEC--TAN / VRAD system of datacube.py:426-486 with
        crpix = n/2 + 0.5 + pad, cdelt = (-px_size, +px_size, -channel_width),
        crval = (ra, dec, spectral_centre), evaluated like wcs_world2pix(..., origin=0)."""
        a, d = np.deg2rad(ra_deg), np.deg2rad(dec_deg)
        a0, d0 = np.deg2rad(self.ra), np.deg2rad(self.dec)
        cosc = np.sin(d0) * np.sin(d) + np.cos(d0) * np.cos(d) * np.cos(a - a0)
        xi = np.cos(d) * np.sin(a - a0) / cosc          # gnomonic projection-plane coordinates
        eta = (np.cos(d0) * np.sin(d) - np.sin(d0) * np.cos(d) * np.cos(a - a0)) / cosc
        scale = np.rad2deg(1.0) * 3600.0 / self.px_size  # radians -> pixels
        px = -xi * scale + (self.n_px_x / 2.0 + 0.5 + self.padx) - 1.0
        py = eta * scale + (self.n_px_y / 2.0 + 0.5 + self.pady) - 1.0
        if self._freq_channel_mode:  # FREQ axis: the particle's line frequency, radio convention
            f = HI_FREQ_HZ * (1.0 - np.asarray(v_kms) / C_KMS)
            pz = (f - self.spectral_centre) / self.channel_width + (self.n_channels / 2.0 + 0.5) - 1.0
        else:
            pz = -(np.asarray(v_kms) - self.spectral_centre) / self.channel_width + (
                self.n_channels / 2.0 + 0.5) - 1.0
        return px, py, pz

    def __repr__(self):
        return (f"DataCube({self.n_px_x}x{self.n_px_y}x{self.n_channels}, pad=({self.padx},"
                f"{self.pady}), px_size={self.px_size} arcsec, channel_width={self.channel_width} {self.channel_unit})")
