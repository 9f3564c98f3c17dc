"""``DataCube`` with the reference's constructor and array/pad/channel semantics
(martini/datacube.py), astropy-free.

Only what the projection path needs is mirrored: the float64 array
``(n_px_x + 2 padx, n_px_y + 2 pady, n_channels[, 1])`` in Jy/pix^2 (channel fastest,
datacube.py:183-186), ``add_pad`` / ``drop_pad`` (:669-729), the velocity channel edges
(:551-562) and the simple TAN / VRAD world coordinate system (:426-486) in closed form.
FITS/HDF5 I/O, ``from_wcs`` and frequency-mode channels are out of scope.

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


class DataCube:
    def __init__(self, *, n_px_x, n_px_y, n_channels, px_size, channel_width,
                 spectral_centre=0.0, ra=0.0, dec=0.0, stokes_axis=False,
                 coordinate_frame=None, specsys="icrs"):
        if coordinate_frame is not None or specsys.lower() != "icrs":
            raise NotImplementedError("martini_b200.DataCube supports the ICRS frame / specsys only")
        self.stokes_axis = stokes_axis
        self.n_px_x, self.n_px_y, self.n_channels = int(n_px_x), int(n_px_y), int(n_channels)
        self.px_size = float(_value(px_size, "arcsec"))
        self.channel_width = abs(float(_value(channel_width, "km/s")))
        self.spectral_centre = float(_value(spectral_centre, "km/s"))
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
                a = self._dev.cpu().numpy()
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

    def _set_device_array(self, tensor):
        """Make ``tensor`` (3-D, on the device) the cube's contents; the host copy is stale."""
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
    def velocity_channel_edges(self):
        """(C+1,) channel edges in km/s, decreasing with channel index: the VRAD axis has
        cdelt = -|channel_width| and crpix = C/2 + 0.5 (datacube.py:469-481)."""
        k = np.arange(self.n_channels + 1)
        return self.spectral_centre + self.channel_width * (self.n_channels / 2.0 - k)

    @property
    def velocity_channel_mids(self):
        e = self.velocity_channel_edges
        return 0.5 * (e[1:] + e[:-1])

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
        pz = -(np.asarray(v_kms) - self.spectral_centre) / self.channel_width + (
            self.n_channels / 2.0 + 0.5) - 1.0
        return px, py, pz

    def __repr__(self):
        return (f"DataCube({self.n_px_x}x{self.n_px_y}x{self.n_channels}, pad=({self.padx},"
                f"{self.pady}), px_size={self.px_size} arcsec, channel_width={self.channel_width} km/s)")
