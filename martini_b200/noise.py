"""``GaussianNoise`` with the reference's constructor, attributes and random stream
(martini/noise.py): the realisation comes from ``numpy.random.default_rng(seed)`` exactly as in
the reference, so a given seed gives the same noise cube; ``Martini.add_noise`` converts it to
the cube's unit and adds it on the device.  Units: ``rms`` in Jy/beam (astropy Quantities
accepted).
"""

from __future__ import annotations

import numpy as np

from .datacube import _value


class _BaseNoise:
    """noise.py:10-76."""

    def __init__(self, seed=0):
        self.seed = seed
        self.rng = np.random.default_rng(seed=seed)

    def generate(self, datacube, beam):
        raise NotImplementedError

    def reset_rng(self):
        """Reset the random number generator to its initial state (noise.py:73-76)."""
        self.rng = np.random.default_rng(seed=self.seed)


class GaussianNoise(_BaseNoise):
    """Gaussian noise whose rms after beam convolution is about ``rms`` (noise.py:79-155)."""

    def __init__(self, rms=1.0, seed=0):
        self.target_rms = float(_value(rms, "Jy/beam"))
        super().__init__(seed=seed)

    def generate(self, datacube, beam):
        """Noise cube in Jy/beam with the shape of ``datacube._array`` (padded), noise.py:110-155:
        the pre-convolution rms is target * 2.19568 * sqrt(pi * sigma_maj * sigma_min), sigmas in
        pixels."""
        sig_maj = beam.bmaj / 2 / np.sqrt(2 * np.log(2)) / datacube.px_size  # same operation order
        sig_min = beam.bmin / 2 / np.sqrt(2 * np.log(2)) / datacube.px_size
        rms = self.target_rms * 2.19568 * np.sqrt(np.pi * sig_maj * sig_min)
        shape = (datacube.n_px_x + 2 * datacube.padx, datacube.n_px_y + 2 * datacube.pady,
                 datacube.n_channels) + ((1,) if datacube.stokes_axis else ())
        return self.rng.normal(scale=rms, size=shape)
