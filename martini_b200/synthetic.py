"""Synthetic SPH sources generated directly in pixel / channel space (float64).

These are the workloads of BASELINE.json / SURVEY.md section 8(d).  A *case* is a plain
dict of host numpy arrays at the seam between MARTINI's coordinate machinery and the hot
path: what ``source.pixcoords``, ``sph_kernel.sm_lengths``, ``skycoords.radial_velocity``,
``skycoords.distance``, ``mHI_g``, ``T_g`` and ``datacube.velocity_channel_edges`` would hold
after ``Martini.__init__`` has run its (astropy-bound) set-up, before pruning.

    px, py, pz   pixel coordinates (pad included; pz = channel coordinate, used by pruning)
    sm_length    smoothing FWHM in pixels
    v            radial velocity [km/s]       sigma   line width [km/s], scalar or (N,)
    mHI          HI mass [Msun]               D       distance [Mpc]
    edges        (C+1,) channel edges [km/s]  shape   (nx, ny, C) padded cube shape
    px_size      pixel size [arcsec]
    kernel       (class name in martini.sph_kernels, kwargs)
    spectrum     "gaussian" | "diracdelta"
"""

from __future__ import annotations

import numpy as np


def channel_edges(nc, dv, v_centre=0.0):
    """Edges of a velocity-mode DataCube: decreasing with channel index
    (datacube.py:469-475: VRAD axis, cdelt = -|channel_width|, crpix = C/2 + 0.5)."""
    return v_centre + dv * (nc / 2.0 - np.arange(nc + 1))


def _finish(case, rng):
    nx, ny, nc = case["shape"]
    e = case["edges"]
    dv = e[0] - e[1]
    case["pz"] = (e[0] - case["v"]) / dv - 0.5  # channel pixel coordinate of v
    case.setdefault("px_size", 10.0)
    for k in ("px", "py", "pz", "sm_length", "v", "mHI", "D"):
        case[k] = np.ascontiguousarray(case[k], dtype=np.float64)
    return case


def disc(rng, n, r_d, incl_deg, centre):
    """Inclined exponential disc: R ~ Gamma(2, r_d), phi ~ U(0, 2pi)."""
    R = rng.gamma(2.0, r_d, n)
    phi = rng.uniform(0.0, 2.0 * np.pi, n)
    ci = np.cos(np.deg2rad(incl_deg))
    return centre[0] + R * np.cos(phi), centre[1] + R * np.sin(phi) * ci, R, phi


def make_case(name, n=None, nx=None, ny=None, nc=None, seed=None):
    """Build a named workload (optionally rescaled through n / nx / ny / nc)."""
    if name in ("cfg2", "smoke", "cfg2c6", "cfg2q"):
        # config 2: 1e6 particles, 256x256x128, WendlandC2Kernel + GaussianSpectrum(7 km/s)
        # (cfg2c6 / cfg2q: the same source with WendlandC6Kernel / QuarticSplineKernel, for
        # timing the kernels BASELINE.json names no config for)
        n = n or 1_000_000
        nx, ny, nc = nx or 256, ny or 256, nc or 128
        seed = 20260002 if seed is None else seed
        rng = np.random.Generator(np.random.PCG64(seed))
        scale = nx / 256.0
        cx, cy = (nx - 1) / 2.0, (ny - 1) / 2.0
        x, y, R, phi = disc(rng, n, 20.0 * scale, 60.0, (cx, cy))
        h = np.clip(rng.lognormal(np.log(2.5), 0.5, n), 0.3, 12.0)
        v = 200.0 * (2 / np.pi) * np.arctan(R / (10.0 * scale)) * np.sin(np.deg2rad(60.0)) * np.cos(
            phi) + rng.normal(0.0, 8.0, n)
        case = {
            "name": name, "px": x, "py": y, "sm_length": h, "v": v, "sigma": 7.0,
            "mHI": (1.0e9 / n) * (1.0 + 0.01 * rng.uniform(-0.5, 0.5, n)),
            "D": np.full(n, 10.0), "edges": channel_edges(nc, 4.0), "shape": (nx, ny, nc),
            "kernel": ({"cfg2c6": "WendlandC6Kernel", "cfg2q": "QuarticSplineKernel"}.get(name, "WendlandC2Kernel"), {}),
            "spectrum": "gaussian",
        }
        return _finish(case, rng)
    if name == "cfg3":
        # config 3: TNG-like, 1e7 particles, 512x512x256, CubicSplineKernel + thermal sigma
        n = n or 10_000_000
        nx, ny, nc = nx or 512, ny or 512, nc or 256
        seed = 20260003 if seed is None else seed
        rng = np.random.Generator(np.random.PCG64(seed))
        scale = nx / 512.0
        cx, cy = (nx - 1) / 2.0, (ny - 1) / 2.0
        nd = int(0.7 * n)
        xd, yd, Rd, phid = disc(rng, nd, 40.0 * scale, 60.0, (cx, cy))
        xh = rng.normal(cx, 120.0 * scale, n - nd)
        yh = rng.normal(cy, 120.0 * scale, n - nd)
        x, y = np.r_[xd, xh], np.r_[yd, yh]
        R = np.hypot(x - cx, (y - cy) / np.cos(np.deg2rad(60.0)))
        phi = np.arctan2((y - cy) / np.cos(np.deg2rad(60.0)), x - cx)
        # h ~ (local surface density)^(-1/3) (tng_source.py:433-438), 0.2 .. 15 px
        sd = 0.7 * np.exp(-R / (40.0 * scale)) / (40.0 * scale) ** 2 + 0.3 * np.exp(
            -0.5 * (R / (120.0 * scale)) ** 2) / (2 * np.pi * (120.0 * scale) ** 2)
        h = np.clip(1.806 * 0.35 * (n * sd) ** (-1.0 / 3.0) * rng.lognormal(0.0, 0.2, n), 0.2, 15.0)
        v = 250.0 * (2 / np.pi) * np.arctan(R / (10.0 * scale)) * np.sin(np.deg2rad(60.0)) * np.cos(
            phi) + rng.normal(0.0, 8.0, n)
        T = 10.0 ** rng.normal(4.0, 0.3, n)
        case = {
            "name": name, "px": x, "py": y, "sm_length": h, "v": v,
            "sigma": np.sqrt(1.380649e-23 * T / 1.67262192369e-27) / 1.0e3, "T": T,
            "mHI": (1.0e9 / n) * (1.0 + 0.01 * rng.uniform(-0.5, 0.5, n)),
            "D": np.full(n, 10.0), "edges": channel_edges(nc, 4.0), "shape": (nx, ny, nc),
            "kernel": ("CubicSplineKernel", {}), "spectrum": "gaussian",
        }
        return _finish(case, rng)
    if name == "cfg4":
        # config 4: wide footprints, GaussianKernel(truncate=3) + DiracDeltaSpectrum
        n = n or 10_000_000
        nx, ny, nc = nx or 512, ny or 512, nc or 256
        seed = 20260004 if seed is None else seed
        rng = np.random.Generator(np.random.PCG64(seed))
        edges = channel_edges(nc, 4.0)
        case = {
            "name": name,
            "px": rng.uniform(-0.1 * nx, 1.1 * nx, n), "py": rng.uniform(-0.1 * ny, 1.1 * ny, n),
            "sm_length": rng.uniform(8.0, 40.0, n) * (nx / 512.0),
            "v": rng.uniform(edges[-1] - 20.0, edges[0] + 20.0, n), "sigma": 0.0,
            "mHI": (1.0e9 / n) * (1.0 + 0.01 * rng.uniform(-0.5, 0.5, n)),
            "D": np.full(n, 10.0), "edges": edges, "shape": (nx, ny, nc),
            "kernel": ("GaussianKernel", {"truncate": 3.0}), "spectrum": "diracdelta",
        }
        return _finish(case, rng)
    if name == "cfg5":
        # config 5: 64 discs on a jittered 8x8 grid + 10 % background, 2048x2048x512
        n = n or 100_000_000
        nx, ny, nc = nx or 2048, ny or 2048, nc or 512
        seed = 20260005 if seed is None else seed
        rng = np.random.Generator(np.random.PCG64(seed))
        scale = nx / 2048.0
        nb = n // 10
        nd = n - nb
        which = rng.integers(0, 64, nd)
        gx = (which % 8 + 0.5 + rng.uniform(-0.2, 0.2, 64)[which]) * nx / 8.0
        gy = (which // 8 + 0.5 + rng.uniform(-0.2, 0.2, 64)[which]) * ny / 8.0
        vsys = rng.uniform(-600.0, 600.0, 64)
        R = rng.gamma(2.0, 30.0 * scale, nd)
        phi = rng.uniform(0.0, 2.0 * np.pi, nd)
        x = np.r_[gx + R * np.cos(phi), rng.uniform(0, nx, nb)]
        y = np.r_[gy + R * np.sin(phi) * 0.5, rng.uniform(0, ny, nb)]
        v = np.r_[
            vsys[which] + 200.0 * (2 / np.pi) * np.arctan(R / (10.0 * scale)) * 0.866 * np.cos(phi)
            + rng.normal(0.0, 8.0, nd),
            rng.uniform(-900.0, 900.0, nb),
        ]
        case = {
            "name": name, "px": x, "py": y,
            "sm_length": np.clip(rng.lognormal(np.log(3.0), 0.6, n), 0.2, 20.0), "v": v,
            "sigma": 7.0, "mHI": (1.0e9 / n) * (1.0 + 0.01 * rng.uniform(-0.5, 0.5, n)),
            "D": np.full(n, 10.0), "edges": channel_edges(nc, 4.0), "shape": (nx, ny, nc),
            "kernel": ("WendlandC2Kernel", {}), "spectrum": "gaussian",
        }
        return _finish(case, rng)
    if name == "demo":
        # config 1 emulation: the demo() source shape (500 particles, 154x154x32 padded cube,
        # CubicSplineKernel + GaussianSpectrum(7 km/s)), SURVEY.md section 8(d)
        n = n or 500
        nx, ny, nc = nx or 154, ny or 154, nc or 32
        seed = 0 if seed is None else seed
        rng = np.random.Generator(np.random.PCG64(seed))
        x, y, R, phi = disc(rng, n, 20.6, 60.0, (76.5, 76.5))
        case = {
            "name": name, "px": x, "py": y,
            "sm_length": (40.0 / np.sqrt(n)) * (1.0 + 0.9 * rng.uniform(-1, 1, n)) * 6.875,
            "v": 50.0 * np.arctan(R / 6.875) * np.sin(np.deg2rad(60.0)) * np.cos(phi),
            "sigma": 7.0, "mHI": np.full(n, 5.0e9 / n), "D": np.full(n, 3.0),
            "edges": channel_edges(nc, 10.0), "shape": (nx, ny, nc),
            "kernel": ("CubicSplineKernel", {}), "spectrum": "gaussian",
        }
        return _finish(case, rng)
    raise KeyError(name)


def make_case_device(name, device, n=None, nx=None, nc=None, seed=20260005):
    """Config 5 generated on the device (1e8 particles are 6.4 GB of float64: too slow to draw
    with numpy and to push through PCIe for a benchmark).  Same recipe as ``make_case("cfg5")``
    (64 discs on a jittered 8 x 8 grid + 10 % uniform background) with torch's generator, so
    every rank that calls it with the same seed holds the same particles.  Returns
    (case dict without the per-particle arrays, dict of device tensors)."""
    import torch

    if name != "cfg5":
        raise KeyError(name)
    n = n or 100_000_000
    nx = ny = nx or 2048
    nc = nc or 512
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    f64 = dict(dtype=torch.float64, device=device)
    u = lambda m: torch.rand(m, generator=g, **f64)  # noqa: E731
    scale = nx / 2048.0
    nb = n // 10
    nd = n - nb
    which = torch.randint(0, 64, (nd,), generator=g, device=device)
    jx, jy = u(64) * 0.4 - 0.2, u(64) * 0.4 - 0.2
    vsys = u(64) * 1200.0 - 600.0
    gx = ((which % 8).double() + 0.5 + jx[which]) * nx / 8.0
    gy = ((which // 8).double() + 0.5 + jy[which]) * ny / 8.0
    R = -(30.0 * scale) * (torch.log(u(nd)) + torch.log(u(nd)))  # Gamma(2, 30 px)
    phi = u(nd) * (2 * np.pi)
    px = torch.cat((gx + R * torch.cos(phi), u(nb) * nx))
    py = torch.cat((gy + R * torch.sin(phi) * 0.5, u(nb) * ny))
    v = torch.cat((vsys[which] + 200.0 * (2 / np.pi) * torch.atan(R / (10.0 * scale)) * 0.866 * torch.cos(phi)
                   + torch.randn(nd, generator=g, **f64) * 8.0, u(nb) * 1800.0 - 900.0))
    del gx, gy, R, phi, which
    sm = torch.clamp(torch.exp(np.log(3.0) + 0.6 * torch.randn(n, generator=g, **f64)), 0.2, 20.0)
    mHI = (1.0e9 / n) * (1.0 + 0.01 * (u(n) - 0.5))
    edges = channel_edges(nc, 4.0)
    pz = (edges[0] - v) / 4.0 - 0.5
    dev = {"px": px, "py": py, "pz": pz, "sm_length": sm, "v": v, "mHI": mHI,
           "D": torch.full((n,), 10.0, **f64), "sigma": 7.0,
           "edges": torch.from_numpy(edges).to(device)}
    case = {"name": "cfg5", "shape": (nx, ny, nc), "edges": edges, "px_size": 10.0, "sigma": 7.0,
            "kernel": ("WendlandC2Kernel", {}), "spectrum": "gaussian"}
    return case, dev
