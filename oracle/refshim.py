"""Run the reference's own hot-path source files without astropy (golden generation).

TEST INFRASTRUCTURE ONLY, and only usable in the build container: it reads
``/root/reference`` (which does not exist on the GPU box).  Used solely by
``tests/golden/make_golden.py`` to produce the committed fixtures under
``tests/golden/``.

The reference is pure Python but imports ``astropy`` (not installed, no network).  On the
hot path astropy only wraps numpy arrays in ``Quantity`` objects; the arithmetic is
numpy + ``scipy.special.erf``.  This module registers a stand-in ``astropy`` in which
*every unit has scale 1* and ``Quantity`` is a thin ``numpy.ndarray`` subclass, then loads
``martini/sph_kernels.py``, ``martini/spectral_models.py`` and ``martini/martini.py``
UNMODIFIED from the reference tree (the modules that are not on the hot path --
``martini.datacube``, ``martini.sources``, ``martini.beams``, ``martini.noise`` -- are
replaced by empty stubs so their astropy.wcs/coordinates imports never run).

Faithfulness: with inputs supplied already in (pix, km/s, Mpc, Msun, arcsec) no unit
conversion with a scale != 1 occurs on the path, except the final Jy/pix^2 -> Jy/arcsec^2
step, which goes through the caller-supplied equivalency function exactly as astropy
would apply it (``Quantity.to`` below).  What the shim cannot reproduce is the few-ulp
effect of astropy's own conversions when the real package mixes m/s and km/s.
"""

from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"


class Unit:
    """A unit of scale 1.  Products/powers of units are units; x * unit is a Quantity."""

    __array_ufunc__ = None  # make ndarray.__mul__ defer to Unit.__rmul__
    __array_priority__ = 1.0e6

    def __mul__(self, other):
        return UNIT if isinstance(other, Unit) else Quantity(other)

    __rmul__ = __mul__

    def __truediv__(self, other):
        return UNIT if isinstance(other, Unit) else Quantity(1 / np.asarray(other))

    def __rtruediv__(self, other):
        return UNIT if isinstance(other, Unit) else Quantity(other)

    def __pow__(self, p):
        return UNIT

    def __eq__(self, other):
        return isinstance(other, Unit)

    def __hash__(self):
        return 1

    def to(self, other, *a, **k):
        return 1.0

    def __repr__(self):
        return "<shim unit>"


UNIT = Unit()


class Quantity(np.ndarray):
    """ndarray that answers the handful of Quantity methods the hot path calls."""

    def __new__(cls, value, unit=None, dtype=None, copy=True):
        arr = np.array(value, dtype=dtype, subok=False)
        if arr.dtype.kind in "iub" and dtype is None:
            arr = arr.astype(np.float64)  # astropy converts integer input to float
        return arr.view(cls)

    def __class_getitem__(cls, item):
        return cls  # ``U.Quantity[U.pix]`` annotations

    @property
    def unit(self):
        return UNIT

    @property
    def value(self):
        return self.view(np.ndarray)

    @property
    def isscalar(self):
        return self.ndim == 0

    def to_value(self, unit=None, equivalencies=()):
        v = self.view(np.ndarray)
        return v[()] if v.ndim == 0 else v

    def to(self, unit=None, equivalencies=()):
        if equivalencies:
            # astropy applies the (from, to, forward, backward) tuple's forward function
            forward = equivalencies[0][2]
            return Quantity(forward(self.view(np.ndarray)))
        return self.copy()

    def __lshift__(self, other):
        return self.view(Quantity)

    def __ilshift__(self, other):
        return self

    def __array_function__(self, func, types, args, kwargs):
        # astropy's Quantity survives np.vstack & co.; keep the subclass here too
        out = super().__array_function__(func, types, args, kwargs)
        if isinstance(out, np.ndarray) and not isinstance(out, Quantity):
            out = out.view(Quantity)
        return out

    # ``q *= unit`` / ``q /= unit`` only relabel the unit in astropy; values are untouched
    def __imul__(self, other):
        return self if isinstance(other, Unit) else super().__imul__(other)

    def __itruediv__(self, other):
        return self if isinstance(other, Unit) else super().__itruediv__(other)


def _units_module():
    m = types.ModuleType("astropy.units")
    for name in (
        "pix", "Jy", "arcsec", "km", "s", "m", "Msun", "Mpc", "kpc", "K", "deg", "rad",
        "Hz", "beam", "dimensionless_unscaled", "one",
    ):
        setattr(m, name, UNIT)
    m.Quantity = Quantity
    m.Unit = Unit
    m.allclose = lambda a, b, **k: np.allclose(np.asarray(a), np.asarray(b), **k)
    m.isclose = lambda a, b, **k: np.isclose(np.asarray(a), np.asarray(b), **k)
    return m


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def load_reference(root: str = REFERENCE_ROOT):
    """Return (sph_kernels, spectral_models, martini) modules loaded from ``root``.

    The modules are registered under a private package name so they never shadow a real
    ``martini`` / ``astropy`` installation.
    """
    if not os.path.isdir(os.path.join(root, "martini")):
        raise FileNotFoundError(f"reference tree not found at {root}")
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k.split(".")[0] in ("astropy", "martini")}
    units = _units_module()
    fakes = {
        "astropy": _stub("astropy", units=units, __version__="shim", __path__=[]),
        "astropy.units": units,
        "astropy.constants": _stub("astropy.constants"),
        "astropy.io": _stub("astropy.io", fits=_stub("astropy.io.fits"), __path__=[]),
        "astropy.io.fits": _stub("astropy.io.fits"),
        "astropy.time": _stub("astropy.time", Time=object),
        "astropy.coordinates": _stub("astropy.coordinates", Angle=object),
        "martini": _stub("martini", __path__=[os.path.join(root, "martini")]),
        "martini.datacube": _stub("martini.datacube", DataCube=object, _GlobalProfileDataCube=object),
        "martini.sources": _stub("martini.sources", SPHSource=object),
        "martini.beams": _stub("martini.beams", _BaseBeam=object),
        "martini.noise": _stub("martini.noise", _BaseNoise=object),
        # __version__.py asks importlib.metadata for the installed distribution
        "martini.__version__": _stub("martini.__version__", __version__="2.1.18"),
    }
    fakes["astropy"].constants = fakes["astropy.constants"]
    sys.modules.update(fakes)
    try:
        mods = []
        for sub in ("sph_kernels", "spectral_models", "martini"):
            name = f"martini.{sub}"
            spec = importlib.util.spec_from_file_location(
                name, os.path.join(root, "martini", f"{sub}.py")
            )
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
            mods.append(mod)
    finally:
        for k in list(sys.modules):
            if k.split(".")[0] in ("astropy", "martini"):
                del sys.modules[k]
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
    return mods[0], mods[1], mods[2], units
