"""SPH kernel classes with the reference's names and attributes (martini/sph_kernels.py).

These are host-side *descriptions*: they carry the constants the CUDA library needs
(``_rescale``, ``size_in_fwhm``, validity thresholds, Gaussian ``truncate``/``norm``) and
the per-particle arrays the reference exposes (``sm_lengths``, ``sm_ranges``,
``kernel_indices``).  The pixel integrals themselves run on the GPU
(``csrc/kernel_integrals.cuh``); ``_px_weight`` here calls the device probe, it does not
compute on the CPU.  A user subclass that overrides ``kernel``/``_kernel_integral`` cannot be
expressed on the device and is rejected by :func:`kernel_table` (no CPU fallback).

Only ``kernel(q)`` -- the 3-D kernel *value* used to find the FWHM rescale at construction
(sph_kernels.py:14-48) and by ``eval_kernel`` -- is evaluated in numpy: it is O(1) set-up
work outside the hot path.
"""

from __future__ import annotations

import numpy as np
from scipy.optimize import fsolve
from scipy.special import erf

from . import _lib as L
from .engine import KernelTable


def find_fwhm(f):
    """FWHM of ``f`` (maximum at 0, symmetric); same root find as sph_kernels.py:14-48."""
    return 2 * fsolve(lambda q: f(q) - f(np.zeros(1)) / 2, 0.5)[0]


def _to_host(x):
    """Device tensor or array-like -> host numpy array."""
    return x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)


class _LazyValid:
    """What ``_validate`` returns: the per-particle validity array of the reference, fetched
    from the device only if somebody looks at it."""

    def __init__(self, kernel):
        self._k = kernel

    def __array__(self, dtype=None, copy=None):
        a = np.asarray(self._k._valid, dtype=bool)
        return a if dtype is None else a.astype(dtype)

    def all(self):
        return self._k._all_valid()

    def __getattr__(self, name):
        return getattr(np.asarray(self), name)

    def __getitem__(self, i):
        return np.asarray(self)[i]

    def __len__(self):
        return len(np.asarray(self))


class _BaseSPHKernel:
    """Attributes shared by all kernels (sph_kernels.py:51-337)."""

    _kind = None
    min_valid_size = None
    max_valid_size = None

    def __init__(self):
        self._rescale = 1.0
        self.size_in_fwhm = None
        self._engine = None  # set by Martini; used by _px_weight

    # -- 3-D kernel value -------------------------------------------------------------
    def kernel(self, q):  # pragma: no cover - abstract
        raise NotImplementedError

    def eval_kernel(self, r, h):
        """Kernel value at ``r`` for FWHM ``h`` (sph_kernels.py:192-219)."""
        q = np.array(np.asarray(r, dtype=float) / h / self._rescale)
        W = self.kernel(q) / np.power(h * self._rescale, 3)
        return W.item() if q.ndim == 0 else W

    def _fwhm_rescale(self):
        fwhm = find_fwhm(lambda r: self.eval_kernel(r, 1))
        self.size_in_fwhm = 1 / fwhm
        self._rescale /= fwhm

    # -- table for the device ----------------------------------------------------------
    def _entry(self):
        return {
            "kind": self._kind,
            "valid_is_max": int(self.max_valid_size is not None),
            "rescale": float(self._rescale),
            "size_in_fwhm": float(self.size_in_fwhm),
            "valid_size": float(
                self.max_valid_size if self.max_valid_size is not None else self.min_valid_size
            ),
            "truncate": float(getattr(self, "truncate", 0.0)),
            "norm": float(getattr(self, "norm", 1.0)),
        }

    def _table(self) -> KernelTable:
        return KernelTable([self._entry()], adaptive=False)

    # -- per-particle state --------------------------------------------------------------
    # Martini keeps the per-particle kernel state (smoothing lengths and ranges in pixels,
    # kernel choice, validity -- the outputs of mtn_smoothing_setup) and the prune mask on the
    # device; the host copies the reference exposes as attributes (sm_lengths, sm_ranges,
    # kernel_indices ...) are made, and compacted by the mask, on first access.
    _state = None     # dict of device tensors / host arrays: sm_lengths, sm_ranges, kernel_ids, valid
    _mask = None      # pending prune mask (device tensor or host array), applied on access
    _host = None      # materialised host arrays

    def _set_device_state(self, sm_lengths, sm_ranges, kernel_ids, valid):
        """Called by Martini after mtn_smoothing_setup (device tensors or host arrays)."""
        self._state = {"sm_lengths": sm_lengths, "sm_ranges": sm_ranges, "kernel_ids": kernel_ids, "valid": valid}
        self._mask = None
        self._host = None

    def _apply_mask(self, mask):
        if self._host is not None:
            m = _to_host(mask).astype(bool)
            self._host = {k: v[m] for k, v in self._host.items()}
        elif self._mask is None:
            self._mask = mask
        else:  # a second mask refers to the already compacted arrays
            self._materialise()
            self._apply_mask(mask)

    def _materialise(self):
        if self._host is None and self._state is not None:
            h = {k: _to_host(v) for k, v in self._state.items()}
            if self._mask is not None:
                m = _to_host(self._mask).astype(bool)
                h = {k: v[m] for k, v in h.items()}
            self._host, self._mask = self._derive(h), None
        return self._host

    def _derive(self, h):
        return h

    def _get(self, key, default=None):
        h = self._materialise()
        return default if h is None else h.get(key, default)

    sm_lengths = property(lambda self: self._get("sm_lengths"), lambda self, v: self._set("sm_lengths", v))
    sm_ranges = property(lambda self: self._get("sm_ranges"), lambda self, v: self._set("sm_ranges", v))
    _valid = property(lambda self: self._get("valid"))

    def _set(self, key, value):
        if self._materialise() is None:
            self._host = {}
        self._host[key] = value

    def _all_valid(self):
        """Whether every kept particle passes the kernel's accuracy check -- on the device if
        the state still lives there (one scalar comes back)."""
        if self._host is None and self._state is not None and hasattr(self._state["valid"], "device"):
            ok = self._state["valid"].bool()
            if self._mask is not None and hasattr(self._mask, "device"):
                ok = ok | ~self._mask.bool()
                return bool(ok.all())
        return bool(np.asarray(self._valid, dtype=bool).all())

    def _validate(self, sm_lengths=None, noraise=False, quiet=False):
        if not self._all_valid() and not noraise:
            raise RuntimeError(self._validation_message())
        return _LazyValid(self)

    def _confirm_validation(self, noraise=False, quiet=False):
        return self._validate(None, noraise=noraise, quiet=quiet)

    def _validation_message(self):
        name = type(self).__name__
        if self.max_valid_size is not None:
            cond = f"provided smoothing scale (FWHM) must be <= {self.max_valid_size:f} px"
        else:
            cond = f"SPH smoothing lengths must be >= {self.min_valid_size:f} px"
        return (
            f"martini.sph_kernels.{name}._validate:\n{cond} in size for the {name} kernel "
            "integral approximation accuracy within 1%.\nThis check may be disabled by calling "
            "martini.martini.Martini.insert_source_in_cube with 'skip_validation=True', but "
            "use this with care."
        )

    def _px_weight(self, dij, mask=Ellipsis):
        """Pixel-integrated weights [pix^-2] for offsets ``dij`` (2, n), on the GPU."""
        from .engine import Engine

        eng = self._engine or Engine()
        h = np.asarray(self.sm_lengths, dtype=float)[mask]
        resc = self._rescale if np.ndim(self._rescale) == 0 else np.asarray(self._rescale)[mask]
        return self._device_integral(eng, np.asarray(dij, dtype=float), h * resc, mask)

    def _device_integral(self, eng, dij, h_eff, mask):
        return eng.probe_kernel_integral(self._entry(), dij[0], dij[1], h_eff).cpu().numpy()


class _WendlandC2Kernel(_BaseSPHKernel):
    """Wendland C2, (21/2pi)(1-q)^4(4q+1) (sph_kernels.py:340-478)."""

    _kind = L.KERNEL_WENDLANDC2
    min_valid_size = 1.51

    def __init__(self):
        super().__init__()
        self._fwhm_rescale()

    def kernel(self, q):
        return np.where(q < 1, (1 - q) ** 4 * (4 * q + 1), 0.0) * (21 / 2 / np.pi)


class _WendlandC6Kernel(_BaseSPHKernel):
    """Wendland C6 (sph_kernels.py:481-721)."""

    _kind = L.KERNEL_WENDLANDC6
    min_valid_size = 1.29

    def __init__(self):
        super().__init__()
        self._fwhm_rescale()

    def kernel(self, q):
        poly = 1 + 8 * q + 25 * q**2 + 32 * q**3
        return np.where(q < 1, (1 - q) ** 8 * poly, 0.0) * (1365 / 64 / np.pi)


class _CubicSplineKernel(_BaseSPHKernel):
    """Cubic spline M4 (sph_kernels.py:724-894)."""

    _kind = L.KERNEL_CUBICSPLINE
    min_valid_size = 1.16

    def __init__(self):
        super().__init__()
        self._fwhm_rescale()

    def kernel(self, q):
        W = np.where(q < 0.5, 1 - 6 * q**2 + 6 * q**3, 2 * (1 - q) ** 3)
        return np.where(q > 1, 0.0, W) * (8 / np.pi)


class _GaussianKernel(_BaseSPHKernel):
    """Truncated Gaussian with FWHM 1 (sph_kernels.py:897-1080)."""

    _kind = L.KERNEL_GAUSSIAN

    def __init__(self, truncate=3.0):
        self.truncate = truncate
        if truncate < 2:
            raise RuntimeError(
                "GaussianKernel with truncation <2sigma will cause large errors in total mass."
            )
        # validity thresholds by truncation radius, sph_kernels.py:938-947
        for lim, size in ((3, 3.7), (4, 2.3357), (5, 1.1288), (6, 0.45), (np.inf, 0.336)):
            if truncate < lim:
                self.min_valid_size = size
                break
        self.norm = erf(truncate / np.sqrt(2)) - 2 * truncate / np.sqrt(2 * np.pi) * np.exp(
            -np.power(truncate, 2) / 2
        )
        super().__init__()
        self.size_in_fwhm = truncate / (2 * np.sqrt(2 * np.log(2)))

    def kernel(self, q):
        sig = 1 / (2 * np.sqrt(2 * np.log(2)))
        g = np.power(sig * np.sqrt(2 * np.pi), -3) * np.exp(-np.power(q / sig, 2) / 2)
        return np.where(q < self.truncate * sig, g, 0.0) / self.norm


class DiracDeltaKernel(_BaseSPHKernel):
    """Point-like particles (sph_kernels.py:1083-1204)."""

    _kind = L.KERNEL_DIRACDELTA
    max_valid_size = 0.5

    def __init__(self, size_in_fwhm=1.0):
        super().__init__()
        self.size_in_fwhm = size_in_fwhm
        self._rescale = 1.0

    def kernel(self, q):
        return np.where(q, 0, np.inf)


class _QuarticSplineKernel(_BaseSPHKernel):
    """Quartic spline M5 (sph_kernels.py:1402-1600)."""

    _kind = L.KERNEL_QUARTICSPLINE
    min_valid_size = 1.2385

    def __init__(self):
        super().__init__()
        self._fwhm_rescale()

    def kernel(self, q):
        q = np.asarray(q, dtype=float)
        W = np.where(q < 1, (1 - q) ** 4, 0.0)
        W = W - np.where(q < 0.6, 5 * (0.6 - q) ** 4, 0.0)
        W = W + np.where(q < 0.2, 10 * (0.2 - q) ** 4, 0.0)
        return W * (15625 / 512 / np.pi)


PRIMITIVE_KERNELS = (
    _WendlandC2Kernel,
    _WendlandC6Kernel,
    _CubicSplineKernel,
    _GaussianKernel,
    DiracDeltaKernel,
    _QuarticSplineKernel,
)


class _AdaptiveKernel(_BaseSPHKernel):
    """First-valid-kernel-per-particle meta kernel (sph_kernels.py:1207-1399)."""

    def __init__(self, kernels):
        self.kernels = tuple(kernels)
        super().__init__()

    def kernel(self, q):
        return self.kernels[0].kernel(q)

    def eval_kernel(self, r, h):
        return self.kernels[0].eval_kernel(r, h)

    def _table(self) -> KernelTable:
        return KernelTable([k._entry() for k in self.kernels], adaptive=True)

    def _derive(self, h):
        # the reference keeps -1 for "no kernel validated" (:1254) and maps it to entry 0
        kid = h["kernel_ids"].astype(int)
        h["kernel_indices"] = np.where(h["valid"].astype(bool), kid, -1)
        h["size_in_fwhm"] = np.array([k.size_in_fwhm for k in self.kernels])[kid]
        h["rescale"] = np.array([k._rescale for k in self.kernels])[kid]
        return h

    kernel_indices = property(lambda self: self._get("kernel_indices"), lambda self, v: self._set("kernel_indices", v))
    size_in_fwhm = property(lambda self: self._get("size_in_fwhm"), lambda self, v: self._set("size_in_fwhm", v))
    _rescale = property(lambda self: self._get("rescale", 1.0), lambda self, v: self._set("rescale", v))

    def _validation_message(self):
        return (
            "martini.sph_kernels._AdaptiveKernel._validate:\nSome particles have no kernel "
            "candidate for which accuracy passes validation.\nThis check may be disabled by "
            "calling martini.martini.Martini.insert_source_in_cube with 'skip_validation=True', "
            "but use this with care.\n"
        )

    def _device_integral(self, eng, dij, h_eff, mask):
        out = np.zeros(h_eff.shape)
        kidx = np.maximum(np.asarray(self.kernel_indices)[mask], 0)
        for ik in np.unique(kidx):
            sel = kidx == ik
            out[sel] = (
                eng.probe_kernel_integral(self.kernels[ik]._entry(), dij[0, sel], dij[1, sel],
                                          h_eff[sel]).cpu().numpy()
            )
        return out


def _fallbacks():
    return DiracDeltaKernel(), _GaussianKernel(truncate=6.0)


class WendlandC2Kernel(_AdaptiveKernel):
    """sph_kernels.py:1603-1660."""

    def __init__(self):
        super().__init__((_WendlandC2Kernel(),) + _fallbacks())


class WendlandC6Kernel(_AdaptiveKernel):
    """sph_kernels.py:1663-1720."""

    def __init__(self):
        super().__init__((_WendlandC6Kernel(),) + _fallbacks())


class CubicSplineKernel(_AdaptiveKernel):
    """sph_kernels.py:1723-1786."""

    def __init__(self):
        super().__init__((_CubicSplineKernel(),) + _fallbacks())


class GaussianKernel(_AdaptiveKernel):
    """sph_kernels.py:1789-1863."""

    def __init__(self, truncate=3.0):
        super().__init__((_GaussianKernel(truncate=truncate),) + _fallbacks())


class QuarticSplineKernel(_AdaptiveKernel):
    """sph_kernels.py:1866-1933."""

    def __init__(self):
        super().__init__((_QuarticSplineKernel(),) + _fallbacks())


class AdaptiveKernel:
    """Removed upstream in v2.0.3 (sph_kernels.py:1936-1959); kept to raise the same error."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            "Pre-configured adaptive kernels have been implemented in v2.0.3. You most likely "
            "want to use WendlandC2Kernel(), WendlandC6Kernel(), CubicSplineKernel() or "
            "QuarticSplineKernel() where you previously used AdaptiveKernel(...)."
        )


_ADAPTIVE_PUBLIC = (
    WendlandC2Kernel,
    WendlandC6Kernel,
    CubicSplineKernel,
    GaussianKernel,
    QuarticSplineKernel,
)


def kernel_table(sph_kernel) -> KernelTable:
    """Device table for a kernel object, or ``NotImplementedError`` for anything the CUDA
    library cannot express (user subclasses may override the integral: no CPU fallback)."""
    t = type(sph_kernel)
    if t in PRIMITIVE_KERNELS:
        return sph_kernel._table()
    if t is _AdaptiveKernel or t in _ADAPTIVE_PUBLIC:
        for k in sph_kernel.kernels:
            if type(k) not in PRIMITIVE_KERNELS:
                raise NotImplementedError(
                    f"adaptive kernel member {type(k).__name__} is not a built-in kernel; "
                    "martini_b200 has no CPU fallback for user-defined kernels"
                )
        return sph_kernel._table()
    raise NotImplementedError(
        f"SPH kernel class {t.__name__} is not supported by martini_b200 (only the built-in "
        "kernels run on the GPU, and there is no CPU fallback)"
    )
