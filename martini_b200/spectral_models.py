"""Spectral model classes with the reference's names (martini/spectral_models.py).

Host-side descriptions only: the per-particle line spectra are evaluated inside the CUDA
projection kernel (the N x C ``spectra`` array of the reference is never materialised), or
by the device probe when ``init_spectra`` is called explicitly.
"""

from __future__ import annotations

import numpy as np

from . import _lib as L

#: CODATA 2018 k_B [J/K] and m_p [kg] as astropy.constants ships them (spectral_models.py:483)
K_B = 1.380649e-23
M_P = 1.67262192369e-27


def _kms(x):
    """Strip an astropy velocity Quantity to km/s, or pass floats/arrays through."""
    if hasattr(x, "to_value"):
        import astropy.units as U  # only reachable when the caller already uses astropy

        return x.to_value(U.km / U.s)
    return x


class _BaseSpectrum:
    """spectral_models.py:13-301."""

    _kind = None

    def __init__(self, ncpu=None, spec_dtype=np.float64):
        self.ncpu = ncpu if ncpu is not None else 1  # accepted for API compatibility
        self.spectra = None
        # The projection kernel evaluates spectra in float64 in registers whatever spec_dtype
        # says: float32 (the reference's memory-saving mode, spectral_models.py:43-61) is
        # accepted and gives the float64 answer, which lies within the reference's own float32
        # rounding (~2e-7 of the cube peak) of what the reference computes in that mode.
        if np.dtype(spec_dtype) not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise NotImplementedError(
                "martini_b200 evaluates spectra in float64 inside the projection kernel; "
                "spec_dtype must be float64 or float32"
            )
        self.spec_dtype = spec_dtype

    def half_width(self, source):  # pragma: no cover - abstract
        raise NotImplementedError

    def init_spectra(self, source, datacube, engine=None):
        """Materialise the (N, C) spectra [Jy] on the GPU (spectral_models.py:63-147).

        Only for callers that want to look at ``.spectra``; ``insert_source_in_cube`` does
        not need it.  The result is a host numpy array.
        """
        from .engine import Engine

        eng = engine or Engine()
        edges = np.asarray(datacube.velocity_channel_edges, dtype=np.float64)
        check_monotonic(edges)
        amp = np.broadcast_to(
            np.asarray(source.mHI_g, dtype=float) * np.power(np.asarray(source.distance_p, dtype=float), -2)
            / 2.36e5, source.radial_velocity.shape)
        self.spectra = eng.probe_spectra(self._kind, source.radial_velocity, self.half_width(source),
                                         np.ascontiguousarray(amp), edges).cpu().numpy().astype(
                                             self.spec_dtype, copy=False)


def check_monotonic(edges):
    d = np.diff(edges)
    if not (np.all(d > 0) or np.all(d < 0)):
        raise ValueError("Channel edges are not monotonic sequence.")  # spectral_models.py:187


class GaussianSpectrum(_BaseSpectrum):
    """Gaussian line of fixed or thermal width (spectral_models.py:303-485).

    ``sigma`` is in km/s (float or astropy Quantity), or the string ``"thermal"`` for
    sqrt(k_B T / m_p) from the particle temperatures.
    """

    _kind = L.SPECTRUM_GAUSSIAN

    def __init__(self, sigma=7.0, ncpu=None, spec_dtype=np.float64):
        self.sigma_mode = sigma if isinstance(sigma, str) else _kms(sigma)
        super().__init__(ncpu=ncpu, spec_dtype=spec_dtype)

    def half_width(self, source):
        if isinstance(self.sigma_mode, str):
            if self.sigma_mode != "thermal":
                raise ValueError("sigma must be a velocity or 'thermal'")
            return np.sqrt(K_B * np.asarray(source.T_g, dtype=np.float64) / M_P) * 1.0e-3
        return self.sigma_mode


class DiracDeltaSpectrum(_BaseSpectrum):
    """All flux in the channel containing the particle velocity (spectral_models.py:487-587)."""

    _kind = L.SPECTRUM_DIRACDELTA

    def half_width(self, source):
        return 0.0


def spectrum_kind(spectral_model) -> int:
    """Device code for a spectral model, or ``NotImplementedError`` for user subclasses."""
    t = type(spectral_model)
    if t is GaussianSpectrum or t is DiracDeltaSpectrum:
        return t._kind
    raise NotImplementedError(
        f"spectral model class {t.__name__} is not supported by martini_b200 (only "
        "GaussianSpectrum and DiracDeltaSpectrum run on the GPU, and there is no CPU fallback)"
    )
