"""``GaussianBeam`` with the reference's constructor and attributes (martini/beams.py).

The beam image is O(pad^2) set-up work done once on the host, with the same scipy calls as the
reference (a bicubic spline of the finely sampled Gaussian integrated over each pixel,
beams.py:103-159); the convolution of the cube with it runs on the GPU
(``mtn_convolve_beam``).  Units: arcsec and degrees (astropy Quantities accepted).
"""

from __future__ import annotations

import numpy as np
import scipy.interpolate

from .datacube import _value


class _BaseBeam:
    """beams.py:14-159."""

    def __init__(self, bmaj=15.0, bmin=None, bpa=0.0):
        self.bmaj = float(_value(bmaj, "arcsec"))
        self.bmin = float(_value(bmin, "arcsec")) if bmin is not None else self.bmaj
        self.bpa = float(_value(bpa, "deg"))
        self.px_size = None
        self.kernel = None
        self.area = (np.pi * self.bmaj * self.bmin) / 4 / np.log(2)  # arcsec^2, beams.py:85

    def needs_pad(self):
        """beams.py:88-101."""
        if self.kernel is None:
            raise RuntimeError("Beam kernel not initialized.")
        return self.kernel.shape[0] // 2, self.kernel.shape[1] // 2

    def init_kernel(self, datacube):
        """Beam image on the cube's pixel grid, beams.py:103-159 (same sampling, same spline,
        same -- transposing -- default meshgrid indexing for the pixel-edge grids)."""
        self.px_size = datacube.px_size
        npx_x, npx_y = self.kernel_size_px()
        px_edges_x = np.arange(-npx_x - 0.5, npx_x + 0.50001, 1) * self.px_size
        px_edges_y = np.arange(-npx_y - 0.5, npx_y + 0.50001, 1) * self.px_size
        fine_x = np.arange(-npx_x - 0.5, npx_x + 0.501, 0.1) * self.px_size
        fine_y = np.arange(-npx_y - 0.5, npx_y + 0.501, 0.1) * self.px_size
        rbs = scipy.interpolate.RectBivariateSpline(
            fine_x, fine_y, self.f_kernel()(*np.meshgrid(fine_x, fine_y, indexing="ij")), kx=3, ky=3)
        xgrid, ygrid = np.meshgrid(px_edges_x, px_edges_y)
        self.kernel = np.vectorize(rbs.integral)(xgrid[1:, :-1], xgrid[1:, 1:], ygrid[:-1, 1:], ygrid[1:, 1:])


class GaussianBeam(_BaseBeam):
    """Elliptical Gaussian beam (beams.py:198-299)."""

    def __init__(self, bmaj=15.0, bmin=None, bpa=0.0, truncate=4.0):
        self.truncate = truncate
        super().__init__(bmaj=bmaj, bmin=bmin, bpa=bpa)

    def f_kernel(self):
        """beams.py:231-275."""
        to_sigma = 1.0 / (2.0 * np.sqrt(2.0 * np.log(2.0)))
        smaj, smin, pa = self.bmaj * to_sigma, self.bmin * to_sigma, np.deg2rad(self.bpa)
        a = np.power(np.cos(pa), 2) / (2.0 * smin**2) + np.power(np.sin(pa), 2) / (2.0 * smaj**2)
        b = -np.sin(2.0 * pa) / (4 * smin**2) + np.sin(2.0 * pa) / (4 * smaj**2)
        c = np.power(np.sin(pa), 2) / (2.0 * smin**2) + np.power(np.cos(pa), 2) / (2.0 * smaj**2)
        A = 1.0 / (2.0 * np.pi * smin * smaj)  # arcsec^-2
        return lambda x, y: A * np.exp(-a * np.power(x, 2) - 2.0 * b * x * y - c * np.power(y, 2))

    def kernel_size_px(self):
        """beams.py:277-295."""
        size = int(np.ceil(self.bmaj * self.truncate / self.px_size + 1))
        return size, size
