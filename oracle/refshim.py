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


# --------------------------------------------------------------------------------------------
# Units WITH scale factors, for the two seam functions whose arithmetic is a unit conversion:
# _BaseSPHKernel._init_sm_lengths (sph_kernels.py:235-255: arctan(hsm / D) -> pixels) and
# GaussianSpectrum.half_width, sigma="thermal" (spectral_models.py:465-485: sqrt(k_B T / m_p)
# -> km/s).  Under the scale-1 shim above both would be meaningless, so make_golden.py runs
# them -- unmodified -- under this second stand-in, in which a unit is (exact rational scale,
# dimension vector), a Quantity carries its unit through * / sqrt / arctan, and ``.to`` multiplies
# the value by ONE correctly rounded factor from_scale / to_scale (what astropy's converters do
# up to the order of their own floating-point operations, which cannot be known without
# astropy: the fixtures pin the formulas and constants, not astropy's last ulp).
# --------------------------------------------------------------------------------------------
from fractions import Fraction  # noqa: E402

_DIMS = ("length", "time", "mass", "temperature", "angle", "pix")


class SUnit:
    __array_ufunc__ = None
    __array_priority__ = 1.0e6

    def __init__(self, scale, dims):
        self.scale = Fraction(scale)
        self.dims = tuple(Fraction(d) for d in dims)

    @classmethod
    def base(cls, name, scale=1):
        return cls(scale, [1 if d == name else 0 for d in _DIMS])

    def __mul__(self, other):
        if isinstance(other, SUnit):
            return SUnit(self.scale * other.scale, [a + b for a, b in zip(self.dims, other.dims)])
        if isinstance(other, SQuantity):
            return SQuantity(other.view(np.ndarray), other.unit * self)
        return SQuantity(other, self)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, SUnit):
            return SUnit(self.scale / other.scale, [a - b for a, b in zip(self.dims, other.dims)])
        return SQuantity(1.0 / np.asarray(other, dtype=np.float64), self)

    def __rtruediv__(self, other):
        if isinstance(other, SQuantity):
            return SQuantity(other.view(np.ndarray), other.unit / self)
        return SQuantity(other, SUnit(1, [0] * len(_DIMS)) / self)

    def __pow__(self, p):
        p = Fraction(p).limit_denominator(16)
        if p.denominator == 1:
            scale = self.scale ** int(p)
        else:  # half-integer powers only arise on units whose scale is 1 here (SI: J / kg)
            assert self.scale == 1, "fractional power of a scaled unit"
            scale = Fraction(1)
        return SUnit(scale, [d * p for d in self.dims])

    def __eq__(self, other):
        return isinstance(other, SUnit) and self.scale == other.scale and self.dims == other.dims

    def __hash__(self):
        return hash((self.scale, self.dims))

    def factor_to(self, other):
        """Multiply a value in this unit by the result to express it in `other`."""
        assert self.dims == other.dims, f"incompatible units {self.dims} -> {other.dims}"
        return float(self.scale / other.scale)

    def to(self, other, *a, **k):
        return self.factor_to(other)

    def __repr__(self):
        return f"<SUnit {float(self.scale):g} {dict((n, str(d)) for n, d in zip(_DIMS, self.dims) if d)}>"


_ONE = SUnit(1, [0] * len(_DIMS))


class SQuantity(np.ndarray):
    def __new__(cls, value, unit=None, dtype=None, copy=True):
        if isinstance(value, SQuantity) and unit is None:
            unit = value.unit
        arr = np.array(value, dtype=dtype, subok=False)
        if arr.dtype.kind in "iub" and dtype is None:
            arr = arr.astype(np.float64)
        out = arr.view(cls)
        out._unit = unit if unit is not None else _ONE
        return out

    def __array_finalize__(self, obj):
        self._unit = getattr(obj, "_unit", _ONE)

    def __class_getitem__(cls, item):
        return cls

    @property
    def unit(self):
        return self._unit

    @property
    def value(self):
        return self.view(np.ndarray)

    @property
    def isscalar(self):
        return self.ndim == 0

    def _converted(self, unit, equivalencies=()):
        if unit.dims == self._unit.dims:
            return self.view(np.ndarray) * self._unit.factor_to(unit)
        for funit, tunit in equivalencies:  # pixel_scale: (pix, physical unit)
            if self._unit.dims == tunit.dims and unit.dims == funit.dims:
                return self.view(np.ndarray) * float(self._unit.scale / tunit.scale * funit.scale / unit.scale)
            if self._unit.dims == funit.dims and unit.dims == tunit.dims:
                return self.view(np.ndarray) * float(self._unit.scale / funit.scale * tunit.scale / unit.scale)
        raise ValueError(f"cannot convert {self._unit} to {unit}")

    def to(self, unit, equivalencies=()):
        return SQuantity(self._converted(unit, equivalencies), unit)

    def to_value(self, unit=None, equivalencies=()):
        v = self.view(np.ndarray) if unit is None else self._converted(unit, equivalencies)
        return v[()] if np.ndim(v) == 0 else v

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        units = [i.unit if isinstance(i, SQuantity) else _ONE for i in inputs]
        raw = [i.view(np.ndarray) if isinstance(i, SQuantity) else i for i in inputs]
        if method != "__call__":
            return NotImplemented
        if ufunc is np.multiply:
            return SQuantity(ufunc(*raw, **kwargs), units[0] * units[1])
        if ufunc in (np.divide, np.true_divide):
            return SQuantity(ufunc(*raw, **kwargs), units[0] / units[1])
        if ufunc is np.sqrt:
            return SQuantity(ufunc(*raw, **kwargs), units[0] ** Fraction(1, 2))
        if ufunc is np.arctan:  # argument to dimensionless first, result in radians
            x = raw[0] * units[0].factor_to(_ONE)
            return SQuantity(ufunc(x, **kwargs), SUNITS["rad"])
        if ufunc in (np.add, np.subtract):
            other = raw[1] * units[1].factor_to(units[0]) if units[1] != units[0] else raw[1]
            return SQuantity(ufunc(raw[0], other, **kwargs), units[0])
        if ufunc in (np.greater, np.greater_equal, np.less, np.less_equal, np.equal, np.not_equal):
            other = raw[1] * units[1].factor_to(units[0]) if units[1] != units[0] else raw[1]
            return ufunc(raw[0], other, **kwargs)
        if ufunc in (np.ceil, np.floor, np.absolute, np.negative):
            return SQuantity(ufunc(*raw, **kwargs), units[0])
        return NotImplemented

    def __lshift__(self, unit):
        return self.to(unit)


_PC = Fraction(30856775814913673)  # metres per parsec (IAU 2015, as astropy rounds it)
SUNITS = {
    "m": SUnit.base("length"), "km": SUnit.base("length", 1000), "pc": SUnit.base("length", _PC),
    "kpc": SUnit.base("length", 1000 * _PC), "Mpc": SUnit.base("length", 10**6 * _PC),
    "s": SUnit.base("time"), "kg": SUnit.base("mass"), "K": SUnit.base("temperature"),
    "rad": SUnit.base("angle"), "pix": SUnit.base("pix"),
    "dimensionless_unscaled": _ONE, "one": _ONE,
}
# pi / 648000 rad per arcsec, pi / 180 per degree: one correctly rounded float each
SUNITS["arcsec"] = SUnit.base("angle", Fraction(np.pi) / 648000)
SUNITS["deg"] = SUnit.base("angle", Fraction(np.pi) / 180)
SUNITS["Msun"] = SUnit.base("mass", Fraction(1.988409870698051e30))
SUNITS["J"] = SUNITS["kg"] * SUNITS["m"] ** 2 / SUNITS["s"] ** 2
SUNITS["Hz"] = _ONE / SUNITS["s"]
SUNITS["Jy"] = SUnit(Fraction(1, 10**26), (SUNITS["J"] / SUNITS["m"] ** 2).dims)  # W m^-2 Hz^-1
SUNITS["beam"] = _ONE


def _scaled_units_module():
    m = types.ModuleType("astropy.units")
    for name, u in SUNITS.items():
        setattr(m, name, u)
    m.Quantity = SQuantity
    m.Unit = SUnit
    # astropy.units.pixel_scale(pixscale): pix <-> the physical unit one pixel spans
    m.pixel_scale = lambda q: [(SUNITS["pix"], SUnit(q.unit.scale * Fraction(float(q.value)), q.unit.dims)
                                * SUNITS["pix"])]
    return m


def load_reference_scaled(root: str = REFERENCE_ROOT):
    """(sph_kernels, spectral_models, units) of the reference loaded under the scaled-unit
    stand-in; CODATA 2018 k_B and m_p as astropy.constants ships them."""
    if not os.path.isdir(os.path.join(root, "martini")):
        raise FileNotFoundError(f"reference tree not found at {root}")
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k.split(".")[0] in ("astropy", "martini")}
    units = _scaled_units_module()
    consts = _stub("astropy.constants",
                   k_B=SQuantity(1.380649e-23, SUNITS["J"] / SUNITS["K"]),
                   m_p=SQuantity(1.67262192369e-27, SUNITS["kg"]))
    fakes = {
        "astropy": _stub("astropy", units=units, constants=consts, __version__="shim", __path__=[]),
        "astropy.units": units,
        "astropy.constants": consts,
        "martini": _stub("martini", __path__=[os.path.join(root, "martini")]),
        "martini.datacube": _stub("martini.datacube", DataCube=object, _GlobalProfileDataCube=object),
        "martini.sources": _stub("martini.sources", SPHSource=object),
        "martini.sources.sph_source": _stub("martini.sources.sph_source", SPHSource=object),
    }
    sys.modules.update(fakes)
    try:
        mods = []
        for sub in ("sph_kernels", "spectral_models"):
            name = f"martini.{sub}"
            spec = importlib.util.spec_from_file_location(name, os.path.join(root, "martini", f"{sub}.py"))
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
    return mods[0], mods[1], units
