"""Binding of ``libmartini_b200.so`` for MARTINI's OWN classes (kyleaoman/martini).

This is the stub of INTEGRATION.md as a shipped module: given a reference ``Martini`` object
``m`` -- real ``martini.sph_kernels`` / ``martini.spectral_models`` instances, an ``SPHSource``
and a ``DataCube`` holding astropy ``Quantity`` arrays -- the two functions below do what
``_BaseMartini._prune_particles`` (martini/martini.py:168-241) and
``_BaseMartini._insert_source_in_cube`` (martini/martini.py:285-407) do, on the GPU through
the C ABI, and leave ``m`` in the state the reference's methods would::

    from martini_b200.reference_adapter import patch
    patch(martini.martini._BaseMartini)      # or call the two functions yourself

Only attribute reads documented in SURVEY.md section 8(b) are used: ``source.pixcoords``,
``source.skycoords.radial_velocity`` / ``.distance``, ``source.mHI_g``; ``sph_kernel.sm_lengths``,
``.kernels``, ``._rescale``, ``.size_in_fwhm``, ``.min_valid_size`` / ``.max_valid_size``,
``.truncate``, ``.norm``; ``spectral_model.half_width(source)``; ``datacube._array``,
``.velocity_channel_edges``, ``.px_size``, ``.n_px_x/y``, ``.padx/y``, ``.n_channels``.
User-subclassed kernels or spectral models raise ``NotImplementedError``: there is no CPU
fallback.

The module itself does not import astropy: ``units`` defaults to ``astropy.units`` when the
functions are called (the caller of a reference ``Martini`` has it), and the CPU test suite
passes the stand-in the reference-made golden fixtures were generated under
(oracle/refshim.py), so the very objects of ``/root/reference`` run through this code in
tests/test_reference_adapter.py.
"""

from __future__ import annotations

import numpy as np

from . import _lib as L
from .engine import Engine, KernelTable

#: exact class name -> MTN_KERNEL_* (a subclass may override _kernel_integral: refused)
_KIND = {"_WendlandC2Kernel": L.KERNEL_WENDLANDC2, "_WendlandC6Kernel": L.KERNEL_WENDLANDC6,
         "_CubicSplineKernel": L.KERNEL_CUBICSPLINE, "_GaussianKernel": L.KERNEL_GAUSSIAN,
         "DiracDeltaKernel": L.KERNEL_DIRACDELTA, "_QuarticSplineKernel": L.KERNEL_QUARTICSPLINE}
_ADAPTIVE = ("_AdaptiveKernel", "WendlandC2Kernel", "WendlandC6Kernel", "CubicSplineKernel", "GaussianKernel",
             "QuarticSplineKernel")
_SPECTRA = {"GaussianSpectrum": L.SPECTRUM_GAUSSIAN, "DiracDeltaSpectrum": L.SPECTRUM_DIRACDELTA}


def _units(units):
    if units is not None:
        return units
    import astropy.units as U  # the caller of a reference Martini object has astropy

    return U


def _entry(k):
    """One MtnKernelEntry from a reference primitive kernel object."""
    name = type(k).__name__
    if name not in _KIND:
        raise NotImplementedError(f"SPH kernel {name} cannot run on the GPU (martini_b200 has no CPU fallback)")
    is_max = getattr(k, "max_valid_size", None) is not None
    return dict(kind=_KIND[name], valid_is_max=int(is_max), rescale=float(k._rescale),
                size_in_fwhm=float(k.size_in_fwhm),
                valid_size=float(k.max_valid_size if is_max else k.min_valid_size),
                truncate=float(getattr(k, "truncate", 0.0)), norm=float(getattr(k, "norm", 1.0)))


def kernel_table(sph_kernel) -> KernelTable:
    """Simple kernel: one entry; the adaptive kernels: one per member of ``.kernels``, in order
    (sph_kernels.py:1207-1240)."""
    name = type(sph_kernel).__name__
    if name in _ADAPTIVE:
        return KernelTable([_entry(k) for k in sph_kernel.kernels], adaptive=True)
    return KernelTable([_entry(sph_kernel)], adaptive=False)


def spectrum_kind(spectral_model) -> int:
    kind = _SPECTRA.get(type(spectral_model).__name__)
    if kind is None:
        raise NotImplementedError(f"spectral model {type(spectral_model).__name__} cannot run on the GPU "
                                  "(martini_b200 has no CPU fallback)")
    return kind


def _seam(m, eng, U):
    """The device copies of everything the hot path reads from ``m`` (units stripped with
    to_value), with K0's outputs; cached on ``m`` until the particle count changes."""
    src, sk = m.source, m.sph_kernel
    n = int(np.shape(src.pixcoords)[-1])
    cached = getattr(m, "_b200_seam", None)
    if cached is not None and cached["n"] == n:
        return cached
    val = lambda q, unit: q.to_value(unit) if hasattr(q, "to_value") else q  # noqa: E731 (plain arrays: already in `unit`)
    f = lambda q, unit: eng.to_device(np.ascontiguousarray(val(q, unit), dtype=np.float64))  # noqa: E731
    table = kernel_table(sk)
    hw = m.spectral_model.half_width(src)
    hw = hw.to_value(U.km / U.s) if hasattr(hw, "to_value") else hw
    mHI = np.asarray(src.mHI_g.to_value(U.Msun), dtype=np.float64)
    d = {
        "n": n, "table": table,
        "px": f(src.pixcoords[0], U.pix), "py": f(src.pixcoords[1], U.pix), "pz": f(src.pixcoords[2], U.pix),
        "sm_length": f(sk.sm_lengths, U.pix),                                  # sph_kernels.py:250-253
        "v": f(src.skycoords.radial_velocity, U.km / U.s),                       # spectral_models.py:92
        "D": f(np.broadcast_to(src.skycoords.distance, (n,)) if np.ndim(src.skycoords.distance) == 0
               else src.skycoords.distance, U.Mpc) if hasattr(src.skycoords.distance, "to_value") else None,
        "mHI": eng.to_device(mHI) if mHI.ndim > 0 else float(mHI),
        "sigma": eng.to_device(np.asarray(hw, dtype=np.float64)) if np.ndim(hw) > 0 else float(hw),
    }
    kid, valid, sm_range, h_eff = eng.smoothing_setup(d["sm_length"], table)     # mtn_smoothing_setup
    d.update(kernel_id=kid, valid=valid, sm_range=sm_range, h_eff=h_eff)
    m._b200_seam = d
    return d


def prune_particles_b200(m, spatial=True, spectral=True, mass=True, obj_type_str="data cube", engine=None,
                         units=None):
    """Drop-in body for ``_BaseMartini._prune_particles`` (martini/martini.py:168-241): the
    accept mask comes from ``mtn_prune`` (bit-exact) and is applied to the host objects with
    the reference's own ``source.apply_mask`` / ``sph_kernel._apply_mask`` (:233-234)."""
    U = _units(units)
    eng = engine or getattr(m, "_b200_engine", None) or Engine("cuda:0")
    m._b200_engine = eng
    d, dc = _seam(m, eng, U), m._datacube
    edges = np.asarray(dc.velocity_channel_edges.to_value(U.km / U.s), dtype=np.float64)
    accept, _ = eng.prune(d["px"], d["py"], d["pz"], d["sm_range"], d["mHI"], d["sigma"],
                          float(np.max(np.abs(np.diff(edges)))), dc.n_px_x + 2 * dc.padx,
                          dc.n_px_y + 2 * dc.pady, dc.n_channels, spatial, spectral, mass)
    mask = accept.cpu().numpy().astype(bool)
    m.source.apply_mask(mask)          # raises RuntimeError if nothing is left, like the reference
    m.sph_kernel._apply_mask(mask)
    m._b200_seam = None                # the host objects changed: re-read them on the next call
    return mask


def insert_source_in_cube_b200(m, skip_validation=False, progressbar=None, ncpu=1, engine=None, units=None):
    """Drop-in body for ``_BaseMartini._insert_source_in_cube`` (martini/martini.py:285-407):
    ``datacube._array`` receives (array + inserted source) converted to Jy/arcsec^2 (:338,
    :364-366).  ``progressbar`` and ``ncpu`` are accepted and ignored."""
    import torch

    U = _units(units)
    eng = engine or getattr(m, "_b200_engine", None) or Engine("cuda:0")
    m._b200_engine = eng
    dc, sk = m._datacube, m.sph_kernel
    spectrum = spectrum_kind(m.spectral_model)
    sk._confirm_validation(noraise=skip_validation, quiet=getattr(m, "quiet", True))  # unchanged host check
    d = _seam(m, eng, U)
    edges_host = np.asarray(dc.velocity_channel_edges.to_value(U.km / U.s), dtype=np.float64)
    shape = tuple(dc._array.shape)
    host = np.ascontiguousarray(np.asarray(dc._array.to_value(U.Jy / U.pix**2), dtype=np.float64).reshape(shape[:3]))
    zero = not host.any()
    cube = (torch.zeros(shape[:3], dtype=torch.float64, device=eng.device) if zero else eng.to_device(host))
    gauss = spectrum == L.SPECTRUM_GAUSSIAN
    plan = eng.insert(px=d["px"], py=d["py"], h_eff=d["h_eff"], sm_range=d["sm_range"], v=d["v"],
                      kernel_id=d["kernel_id"], sigma=d["sigma"] if gauss else 1.0, mHI=d["mHI"],
                      D=d["D"], table=d["table"], spectrum=spectrum, edges=eng.to_device(edges_host),
                      cube=cube, px_size_arcsec=float(dc.px_size.to_value(U.arcsec)), zeroed=zero,
                      edges_increasing=bool(edges_host[1] > edges_host[0]))   # mtn_plan + mtn_project
    dc._array = cube.cpu().numpy().reshape(shape) * U.Jy / U.arcsec**2          # martini.py:364-366
    m._b200_plan = plan
    return plan


def patch(base_martini_cls, engine=None, units=None):
    """Replace the two hot-path methods of the reference's ``_BaseMartini`` (and thereby of
    ``Martini`` / ``GlobalProfile``) by the GPU versions."""
    def _prune(self, spatial=True, spectral=True, mass=True, obj_type_str="data cube"):
        prune_particles_b200(self, spatial, spectral, mass, obj_type_str, engine=engine, units=units)

    def _insert(self, skip_validation=False, progressbar=None, ncpu=1, quiet=None):
        insert_source_in_cube_b200(self, skip_validation, progressbar, ncpu, engine=engine, units=units)

    base_martini_cls._prune_particles = _prune
    base_martini_cls._insert_source_in_cube = _insert
    return base_martini_cls
