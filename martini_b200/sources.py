"""``SPHSource`` with the reference's constructor and attributes (martini/sources/
sph_source.py), astropy-free, plus an astropy-free coordinate front-end for the common case
(ICRS frame and spectral system, gnomonic projection).

The front-end -- rotate, translate to (ra, dec, distance), add peculiar velocity and Hubble
flow, then (RA, Dec, v_radial, D) -> pixel coordinates -- is O(N) numpy set-up work upstream of
the hot path; it feeds the arrays the CUDA path consumes:
``pixcoords`` (3, N), ``radial_velocity``, ``distance_p``, ``mHI_g``, ``hsm_g``, ``T_g``.

Units are fixed: kpc, km/s, Msun, K, Mpc, degrees.  astropy Quantities are accepted and
converted when astropy is installed.
"""

from __future__ import annotations

from typing import NamedTuple

import numpy as np

from .datacube import _value


class L_coords(NamedTuple):
    """Orientation by inclination / azimuthal rotation / position angle in degrees
    (martini/L_coords.py)."""

    incl: float = 0.0
    az_rot: float = 0.0
    pa: float = 270.0


def _rot(axis, angle_rad):
    """Right-handed rotation matrix about a coordinate axis (scipy's from_euler convention)."""
    c, s = np.cos(angle_rad), np.sin(angle_rad)
    if axis == "x":
        return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
    if axis == "y":
        return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def L_align(xyz, vxyz, m, frac=0.3, Laxis="x"):
    """Rotation matrix aligning the angular momentum of the central mass fraction ``frac``
    with ``Laxis`` (martini/sources/_L_align.py:13-120).  xyz, vxyz: (N, 3)."""
    xyz, vxyz = xyz.T, vxyz.T
    rsort = np.argsort(np.sum(np.power(xyz, 2), axis=0), kind="quicksort")
    L = np.cross(xyz, m[np.newaxis] * vxyz, axis=0)[:, rsort]
    ms = m[rsort]
    mcumul = np.cumsum(ms) / np.sum(ms)
    nfrac = min(max(int(np.argmin(np.abs(mcumul - frac))), 100), len(ms))
    Lsum = np.sum(L[:, :nfrac], axis=1)
    zhat = Lsum / np.sqrt(np.sum(np.power(Lsum, 2)))
    xaxis = np.array([1.0, 1.0, 1.0]) / np.sqrt(3)
    if (zhat == xaxis).all():
        raise RuntimeError("Angular momentum exactly aligned with arbitrarily chosen vector, L_align failed.")
    xhat = xaxis - xaxis.dot(zhat) * zhat
    xhat = xhat / np.sqrt(np.sum(np.power(xhat, 2)))
    yhat = np.cross(zhat, xhat)
    rotmat = np.vstack((xhat, yhat, zhat))
    shift = {"z": 0, "y": 2, "x": 1}[Laxis]
    return np.roll(rotmat, shift, axis=0)  # rows = new basis vectors: x' = rotmat . x


def _lazy(name):
    """A per-particle attribute that a deferred prune mask (``_defer_mask``) is applied to on
    first access."""
    key = "_lz_" + name

    def get(self):
        if self.__dict__.get("_pending_mask") is not None:
            self._flush_mask()
        value = self.__dict__.get(key)
        if callable(value):  # a host mirror of a device array, fetched on first access
            value = self.__dict__[key] = value()
        return value

    def set(self, value):
        if self.__dict__.get("_pending_mask") is not None:
            self._flush_mask()
        self.__dict__[key] = value

    return property(get, set)


class SPHSource:
    """Generic particle source (sph_source.py:171-263).

    Pruning (``apply_mask``, sph_source.py:364-393) can be deferred: ``Martini`` computes the
    accept mask on the GPU and the projection uses it there, so the host copies of the
    per-particle arrays are only compacted when somebody reads them (``_defer_mask``);
    ``npart`` is known at once from the device-side count."""

    _pending_mask = None

    def __init__(self, *, distance, vpeculiar=0.0, rotation=None, L_coords=None, ra=0.0, dec=0.0,
                 h=0.7, T_g=None, mHI_g, xyz_g, vxyz_g, hsm_g=None, coordinate_axis=None,
                 coordinate_frame=None):
        if isinstance(rotation, dict):
            raise ValueError("The method to specify rotations in martini has been updated; pass a "
                             "rotation matrix / scipy Rotation, or L_coords=L_coords(...).")
        if coordinate_frame is not None:
            raise NotImplementedError("martini_b200.SPHSource supports the ICRS frame only")
        xyz_g = np.asarray(_value(xyz_g, "kpc"), dtype=np.float64)
        vxyz_g = np.asarray(_value(vxyz_g, "km/s"), dtype=np.float64)
        if coordinate_axis is None:
            if xyz_g.ndim != 2 or xyz_g.shape == (3, 3):
                raise RuntimeError("martini.sources.SPHSource: cannot guess coordinate_axis with shape "
                                   "(3, 3), provide explicitly." if xyz_g.shape == (3, 3) else
                                   "martini.sources.SPHSource: incorrect coordinate shape (not (3, N) or (N, 3)).")
            if xyz_g.shape[0] == 3 and xyz_g.shape[1] != 3:
                coordinate_axis = 0
            elif xyz_g.shape[1] == 3:
                coordinate_axis = 1
            else:
                raise RuntimeError("martini.sources.SPHSource: incorrect coordinate shape (not (3, N) or (N, 3)).")
        if xyz_g.shape != vxyz_g.shape:
            raise ValueError("martini.sources.SPHSource: xyz_g and vxyz_g must have matching shapes.")
        if coordinate_axis == 0:
            xyz_g, vxyz_g = xyz_g.T, vxyz_g.T
        self.h = h
        self.T_g = None if T_g is None else np.asarray(_value(T_g, "K"), dtype=np.float64)
        self.mHI_g = np.asarray(_value(mHI_g, "Msun"), dtype=np.float64)
        self.input_mass = self.mHI_g.sum()
        self.xyz_g = np.ascontiguousarray(xyz_g)      # (N, 3) kpc
        self.vxyz_g = np.ascontiguousarray(vxyz_g)    # (N, 3) km/s
        self.hsm_g = None if hsm_g is None else np.asarray(_value(hsm_g, "kpc"), dtype=np.float64)
        self.npart = self.xyz_g.shape[0]
        self.ra = float(_value(ra, "deg"))
        self.dec = float(_value(dec, "deg"))
        self.distance = float(_value(distance, "Mpc"))
        self.vpeculiar = float(_value(vpeculiar, "km/s"))
        self.vhubble = self.h * 100.0 * self.distance
        self.vsys = self.vhubble + self.vpeculiar
        self.rotate(rotation=rotation, L_coords=L_coords)
        self.skycoords = None
        self.pixcoords = None
        self.radial_velocity = None
        self.distance_p = None

    # ------------------------------------------------------------------ transforms
    def rotate(self, rotation=None, *, L_coords=None):
        """sph_source.py:442-530.  ``rotation``: 3x3 matrix or scipy Rotation."""
        if rotation is None and L_coords is None:
            return np.eye(3)
        if rotation is not None and L_coords is not None:
            raise ValueError("Multiple rotations in a single call not allowed.")
        if rotation is not None:
            do_rot = rotation.as_matrix() if hasattr(rotation, "as_matrix") else np.asarray(rotation, float)
        else:
            incl, az, pa = (float(_value(x, "deg")) for x in L_coords)
            do_rot = L_align(self.xyz_g, self.vxyz_g, np.broadcast_to(self.mHI_g, (self.npart,)),
                             frac=0.3, Laxis="x")
            do_rot = _rot("x", np.deg2rad(az)).dot(do_rot)
            do_rot = _rot("y", np.deg2rad(incl)).dot(do_rot)
            do_rot = _rot("x", np.deg2rad(pa - 90.0 if incl >= 0 else pa - 270.0)).dot(do_rot)
        self.xyz_g = self.xyz_g.dot(do_rot.T)
        self.vxyz_g = self.vxyz_g.dot(do_rot.T)
        return do_rot

    def translate(self, translation_vector):
        self.xyz_g = self.xyz_g + np.asarray(_value(translation_vector, "kpc"), dtype=np.float64)

    def boost(self, boost_vector):
        self.vxyz_g = self.vxyz_g + np.asarray(_value(boost_vector, "km/s"), dtype=np.float64)

    # ------------------------------------------------------------------ sky / pixel coordinates
    def _init_skycoords(self):
        """RA, Dec, distance and radial velocity of every particle (sph_source.py:265-326):
        rotate to (ra, dec), translate by the distance, add the peculiar velocity along the
        line of sight and the Hubble flow of every particle."""
        a0, d0 = np.deg2rad(self.ra), np.deg2rad(self.dec)
        unit = np.array([np.cos(d0) * np.cos(a0), np.cos(d0) * np.sin(a0), np.sin(d0)])
        R = _rot("z", a0).dot(_rot("y", -d0))
        xyz = self.xyz_g.dot(R.T) + unit * (self.distance * 1.0e3)           # kpc
        vxyz = self.vxyz_g.dot(R.T) + unit * self.vpeculiar
        vxyz = vxyz + (self.h * 100.0) * (xyz * 1.0e-3)                        # Hubble flow
        r = np.sqrt(np.sum(xyz * xyz, axis=1))
        ra = np.rad2deg(np.arctan2(xyz[:, 1], xyz[:, 0]))
        dec = np.rad2deg(np.arcsin(xyz[:, 2] / r))
        self.skycoords = {"ra": ra, "dec": dec}
        self.distance_p = r * 1.0e-3                                            # Mpc
        self.radial_velocity = np.sum(xyz * vxyz, axis=1) / r                   # km/s

    def _init_pixcoords(self, datacube):
        """(3, N) pixel coordinates, 0-indexed, pad included (sph_source.py:328-362)."""
        assert self.skycoords is not None, "Initialize source.skycoords before calling _init_pixcoords."
        px, py, pz = datacube.world2pix(self.skycoords["ra"], self.skycoords["dec"], self.radial_velocity)
        self.pixcoords = np.vstack((px, py, pz))

    def _init_on_device(self, engine, datacube):
        """The coordinate front-end on the GPU (``mtn_sky_to_pix``: what ``_init_skycoords`` +
        ``_init_pixcoords`` + ``sm_lengths_px`` compute on the host, fused into one pass over the
        particles).  Returns the device tensors the projection reads; the host attributes
        (``pixcoords``, ``radial_velocity``, ``distance_p``, ``skycoords``) are fetched from them
        when somebody reads them."""
        from . import _lib as L

        a0, d0 = np.deg2rad(self.ra), np.deg2rad(self.dec)
        fe = L.MtnFrontEnd()
        R = _rot("z", a0).dot(_rot("y", -d0))
        fe.rotation[:] = list(R.ravel())
        fe.direction[:] = [np.cos(d0) * np.cos(a0), np.cos(d0) * np.sin(a0), np.sin(d0)]
        fe.distance_mpc, fe.vpeculiar, fe.hubble = self.distance, self.vpeculiar, self.h * 100.0
        fe.ra0_rad, fe.dec0_rad = np.deg2rad(datacube.ra), np.deg2rad(datacube.dec)
        fe.px_size_arcsec = datacube.px_size
        fe.crpix[:] = [datacube.n_px_x / 2.0 + 0.5 + datacube.padx, datacube.n_px_y / 2.0 + 0.5 + datacube.pady,
                       datacube.n_channels / 2.0 + 0.5]
        fe.spectral_centre, fe.channel_width = datacube.spectral_centre, datacube.channel_width
        fe.freq_mode = int(datacube._freq_channel_mode)
        hsm = self.hsm_g if self.hsm_g is not None else 0.0
        px, py, pz, v, D, sm = engine.sky_to_pix(fe, engine.to_device(self.xyz_g), engine.to_device(self.vxyz_g),
                                                 engine.to_device(hsm) if np.ndim(hsm) > 0 else float(hsm))
        host = lambda t: t.cpu().numpy()  # noqa: E731
        self.__dict__["_lz_pixcoords"] = lambda: np.vstack((host(px), host(py), host(pz)))
        self.__dict__["_lz_radial_velocity"] = lambda: host(v)
        self.__dict__["_lz_distance_p"] = lambda: host(D)
        self.__dict__["_lz_skycoords"] = self._skycoords_from_host
        self.__dict__["_lz__sm_lengths"] = lambda: host(sm)
        return {"px": px, "py": py, "pz": pz, "v": v, "D": D, "sm_length": sm}

    def _skycoords_from_host(self):
        a0, d0 = np.deg2rad(self.ra), np.deg2rad(self.dec)
        unit = np.array([np.cos(d0) * np.cos(a0), np.cos(d0) * np.sin(a0), np.sin(d0)])
        xyz = self.xyz_g.dot(_rot("z", a0).dot(_rot("y", -d0)).T) + unit * (self.distance * 1.0e3)
        r = np.sqrt(np.sum(xyz * xyz, axis=1))
        return {"ra": np.rad2deg(np.arctan2(xyz[:, 1], xyz[:, 0])), "dec": np.rad2deg(np.arcsin(xyz[:, 2] / r))}

    def sm_lengths_px(self, datacube):
        """Smoothing lengths in pixels: arctan(hsm / D) / px_size (sph_kernels.py:250-253)."""
        if self.__dict__.get("_lz__sm_lengths") is not None:  # computed by the device front-end
            return self._sm_lengths
        hsm = np.broadcast_to(self.hsm_g, (self.npart,)) if self.hsm_g is not None else np.zeros(self.npart)
        # kpc / Mpc -> dimensionless, radians -> pixels: the operation order astropy's unit
        # converters give the reference (pinned by tests/golden/seam.npz)
        return np.arctan(hsm / self.distance_p * 1.0e-3) * (1.0 / (datacube.px_size * (np.pi / 648000.0)))

    # ------------------------------------------------------------------ pruning
    @property
    def npart(self):
        return self._pending_n if self.__dict__.get("_pending_mask") is not None else self._npart

    @npart.setter
    def npart(self, value):
        self._npart = value

    def _defer_mask(self, fetch_mask, n_kept):
        """Like ``apply_mask(fetch_mask())``, but the host arrays are compacted on first access.
        ``n_kept`` is the number of True entries (raises like apply_mask if it is 0)."""
        if self.__dict__.get("_pending_mask") is not None:
            self._flush_mask()
        if int(n_kept) == 0:
            raise RuntimeError("No non-zero mHI source particles in target region.")
        self._pending_n = int(n_kept)
        self._pending_mask = fetch_mask

    def _flush_mask(self):
        fetch, self._pending_mask = self._pending_mask, None
        if fetch is not None:
            self.apply_mask(fetch())

    def pin_memory(self):
        """Move the arrays the projection uploads into page-locked host memory, so that
        ``Martini``'s host -> device copies run at full PCIe rate and asynchronously (optional:
        pageable arrays work, the driver then stages them itself)."""
        import torch

        for key in ("pixcoords", "radial_velocity", "distance_p", "mHI_g", "T_g", "hsm_g", "_sm_lengths"):
            a = getattr(self, key, None)
            if isinstance(a, np.ndarray) and a.ndim > 0 and a.size:
                t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
                setattr(self, key, t.numpy())  # (the view keeps the pinned tensor alive)
        return self

    def apply_mask(self, mask):
        """sph_source.py:364-393."""
        if self.__dict__.get("_pending_mask") is not None:
            self._flush_mask()
        mask = np.asarray(mask, dtype=bool)
        if mask.size != self.npart:
            raise ValueError("Mask must have same length as particle arrays.")
        if mask.sum() == 0:
            raise RuntimeError("No non-zero mHI source particles in target region.")
        for name in ("skycoords", "pixcoords", "radial_velocity", "distance_p", "_sm_lengths"):
            getattr(self, name, None)  # host mirrors of device arrays are fetched before anything is compacted
        self.npart = int(mask.sum())
        if self.T_g is not None and self.T_g.ndim > 0:
            self.T_g = self.T_g[mask]
        if self.mHI_g.ndim > 0:
            self.mHI_g = self.mHI_g[mask]
        if self.xyz_g.shape[0] == mask.size:  # (PixelSource carries no Cartesian coordinates)
            self.xyz_g, self.vxyz_g = self.xyz_g[mask], self.vxyz_g[mask]
        if self.skycoords is not None:
            self.skycoords = {k: v[mask] for k, v in self.skycoords.items()}
            self.radial_velocity = self.radial_velocity[mask]
            self.distance_p = self.distance_p[mask]
        if self.pixcoords is not None:
            self.pixcoords = self.pixcoords[:, mask]
        if self.hsm_g is not None and self.hsm_g.ndim > 0:
            self.hsm_g = self.hsm_g[mask]
        if type(self) is SPHSource and self._sm_lengths is not None:
            self._sm_lengths = self._sm_lengths[mask]


class PixelSource(SPHSource):
    """A source given directly at the hot path's seam: pixel coordinates, smoothing lengths in
    pixels, radial velocities, distances.  Used for the synthetic benchmark workloads, where
    the coordinate front-end is not the subject (see martini_b200/synthetic.py)."""

    def __init__(self, *, pixcoords, sm_lengths, radial_velocity, distance_p, mHI_g, T_g=None,
                 sigma=None):
        # (arrays that are already float64 and contiguous are kept as they are -- no copies, so
        # page-locked inputs stay page-locked)
        self.pixcoords = np.ascontiguousarray(pixcoords, dtype=np.float64)
        self.npart = self.pixcoords.shape[1]
        self._sm_lengths = np.ascontiguousarray(sm_lengths, dtype=np.float64)
        self.radial_velocity = np.ascontiguousarray(radial_velocity, dtype=np.float64)
        d = np.asarray(distance_p, dtype=np.float64)
        self.distance_p = np.ascontiguousarray(d) if d.shape == (self.npart,) else np.full(self.npart, float(d))
        self.mHI_g = np.asarray(mHI_g, dtype=np.float64)
        self.T_g = None if T_g is None else np.asarray(T_g, dtype=np.float64)
        self.hsm_g = None
        self.skycoords = {}
        self.xyz_g = self.vxyz_g = np.zeros((0, 3))
        self._mHI0, self._D0 = (self.mHI_g, self.npart), self.distance_p

    # (summary quantities of the UNPRUNED source, as the constructor would have computed them,
    # evaluated when somebody asks: the original arrays are kept by reference until then)
    @property
    def input_mass(self):
        if "_input_mass" not in self.__dict__:
            m0, n0 = self.__dict__.pop("_mHI0")
            self._input_mass = np.broadcast_to(m0, (n0,)).sum()
        return self._input_mass

    @property
    def distance(self):
        if "_distance" not in self.__dict__:
            self._distance = float(np.mean(self.__dict__.pop("_D0")))
        return self._distance

    _init_on_device = None  # (already at the seam: nothing to transform)

    def _init_skycoords(self):
        pass

    def _init_pixcoords(self, datacube):
        pass

    def sm_lengths_px(self, datacube):
        return self._sm_lengths

    def apply_mask(self, mask):
        mask = np.asarray(mask, dtype=bool)
        super().apply_mask(mask)
        self._sm_lengths = self._sm_lengths[mask]

    @classmethod
    def from_case(cls, case):
        """Build from a synthetic case dict (martini_b200/synthetic.py)."""
        return cls(pixcoords=np.vstack((case["px"], case["py"], case["pz"])),
                   sm_lengths=case["sm_length"], radial_velocity=case["v"], distance_p=case["D"],
                   mHI_g=case["mHI"], T_g=case.get("T"))


for _name in ("T_g", "mHI_g", "xyz_g", "vxyz_g", "skycoords", "radial_velocity", "distance_p", "pixcoords",
              "hsm_g", "_sm_lengths"):
    setattr(SPHSource, _name, _lazy(_name))


def demo_source(N=500):
    """The toy galaxy of martini/_demo.py:20-92 (same legacy-seeded random sequence)."""
    from scipy.optimize import fsolve

    np.random.seed(0)
    phi = np.random.rand(N) * 2 * np.pi
    r = np.empty(N, dtype=float)
    for i, Lr in enumerate(np.random.rand(N)):
        r[i] = fsolve(lambda x: Lr - 0.5 * (2 - np.exp(-x) * (np.power(x, 2) + 2 * x + 2)), 1.0)[0]
    r *= 3 / np.sort(r)[N // 2]
    z = -np.log(np.random.rand(N))
    z *= 0.5 / np.sort(z)[N // 2] * np.sign(np.random.rand(N) - 0.5)
    xyz = np.vstack((r * np.cos(phi), r * np.sin(phi), z))
    vphi = 50 * np.arctan(r)
    vxyz = np.vstack((-vphi * np.sin(phi), vphi * np.cos(phi), (np.random.rand(N) * 2.0 - 1.0) * 5))
    mHI = np.ones(N) + 0.01 * (np.random.rand(N) - 0.5)
    mHI = mHI / mHI.sum() * 5.0e9
    hsm = 40 / np.sqrt(N) * (np.ones(N) + 1.8 * (np.random.rand(N) - 0.5))
    return SPHSource(distance=3.0, L_coords=L_coords(incl=60.0, az_rot=0.0, pa=270.0), ra=0.0, dec=0.0,
                     h=0.7, T_g=np.ones(N) * 8e3, mHI_g=mHI, xyz_g=xyz, vxyz_g=vxyz, hsm_g=hsm)
