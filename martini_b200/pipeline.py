"""The whole hot path on one device, array level: smoothing setup -> prune -> plan -> project.

``upload`` moves a synthetic *case* (see synthetic.py) to the device; ``run_hot_path`` runs
the four C-ABI stages on device-resident inputs and returns device tensors.  Used by
``smoke()``, the parity tests and ``bench.py``; the ``Martini`` class runs the same stages
split between its constructor (setup + prune) and ``insert_source_in_cube``.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _lib as L
from . import sph_kernels as K
from .engine import KernelTable

SPECTRA = {"gaussian": L.SPECTRUM_GAUSSIAN, "diracdelta": L.SPECTRUM_DIRACDELTA}

PARTICLE_KEYS = ("px", "py", "pz", "sm_length", "v", "mHI", "D")


def kernel_from_spec(spec):
    name, kw = spec
    return getattr(K, name)(**kw)


@dataclass
class CaseContext:
    """Per-case constants that do not depend on the particle data (built once)."""

    table: KernelTable
    spectrum: int
    max_abs_dv: float
    shape: tuple
    px_size: float
    edges_increasing: bool = False


def prepare(case) -> CaseContext:
    kernel = kernel_from_spec(case["kernel"])  # runs the FWHM root find: keep out of the loop
    return CaseContext(
        table=K.kernel_table(kernel),
        spectrum=SPECTRA[case["spectrum"]],
        max_abs_dv=float(np.max(np.abs(np.diff(case["edges"])))),
        shape=tuple(int(s) for s in case["shape"]),
        px_size=float(case["px_size"]),
        edges_increasing=bool(case["edges"][1] > case["edges"][0]),
    )


def particle_keys(case):
    return PARTICLE_KEYS + (("sigma",) if np.ndim(case["sigma"]) > 0 else ())


def pin_case(case):
    """Copy the per-particle arrays of a case into pinned host memory (for timed H2D)."""
    return {k: torch.from_numpy(np.ascontiguousarray(case[k])).pin_memory() for k in particle_keys(case)}


def upload(engine, case, pinned=None, out=None):
    """Host case -> dict of device tensors.  ``pinned`` (from :func:`pin_case`) makes the
    copies asynchronous; ``out`` reuses previously allocated device tensors."""
    dev = {} if out is None else out
    for k in particle_keys(case):
        src = pinned[k] if pinned is not None else torch.from_numpy(np.ascontiguousarray(case[k]))
        if k in dev and isinstance(dev[k], torch.Tensor) and dev[k].shape == src.shape:
            dev[k].copy_(src, non_blocking=True)
        else:
            dev[k] = src.to(engine.device, non_blocking=True)
    if np.ndim(case["sigma"]) == 0:
        dev["sigma"] = float(case["sigma"])
    if "edges" not in dev:
        dev["edges"] = torch.from_numpy(np.ascontiguousarray(case["edges"])).to(engine.device)
    return dev


def h2d_bytes(case):
    return int(sum(np.asarray(case[k]).nbytes for k in particle_keys(case)))


def run_hot_path(engine, case, dev=None, cube=None, x_lo=0, x_hi=None, prune=(True, True, True),
                 zeroed=None, ctx: CaseContext | None = None):
    """Run K0 -> K1 -> plan -> project for ``case``.

    ``dev``  device tensors from :func:`upload` (uploaded here if None);
    ``cube`` float64 device tensor (x_hi-x_lo, ny, C) to accumulate into; a zero cube is
             allocated if None.  Returns dict(cube, accept, n_accept, plan, kernel_id, ...).
    """
    ctx = ctx or prepare(case)
    nx, ny, nc = ctx.shape
    if dev is None:
        dev = upload(engine, case)
    x_hi = nx if x_hi is None else x_hi
    gauss = ctx.spectrum == L.SPECTRUM_GAUSSIAN
    kid, valid, sm_range, h_eff = engine.smoothing_setup(dev["sm_length"], ctx.table)
    accept, n_accept = engine.prune(dev["px"], dev["py"], dev["pz"], sm_range, dev["mHI"],
                                    dev["sigma"] if gauss else 0.0, ctx.max_abs_dv, nx, ny, nc,
                                    *prune)
    if cube is None:
        cube = torch.zeros((x_hi - x_lo, ny, nc), dtype=torch.float64, device=engine.device)
        zeroed = True if zeroed is None else zeroed
    plan = engine.insert(
        px=dev["px"], py=dev["py"], h_eff=h_eff, sm_range=sm_range, v=dev["v"], kernel_id=kid,
        sigma=dev["sigma"] if gauss else 1.0, mHI=dev["mHI"], D=dev["D"], accept=accept,
        table=ctx.table, spectrum=ctx.spectrum, edges=dev["edges"], cube=cube,
        px_size_arcsec=ctx.px_size, x_lo=x_lo, x_hi=x_hi, nx_full=nx, zeroed=bool(zeroed),
        edges_increasing=ctx.edges_increasing,
    )
    return {"cube": cube, "accept": accept, "n_accept": n_accept, "plan": plan, "kernel_id": kid,
            "valid": valid, "sm_range": sm_range, "h_eff": h_eff,
            "launches": engine.last_launches + 2}  # + smoothing_setup + prune


def run_hot_path_to_host(engine, case, host_rows, dev, ctx: CaseContext, slab, x_lo=0, x_hi=None,
                         n_slabs=1, copy_stream=None, prune=(True, True, True)):
    """The hot path with the result delivered to host memory: K0 -> K1 once, then plan + project
    in ``n_slabs`` x-sub-slabs of rows [x_lo, x_hi); each finished sub-slab is copied device ->
    host on ``copy_stream`` while the next one is being projected.  (On BASELINE config 2 the
    extra planning passes cost more than the overlap hides -- 7.2 ms with one sub-slab, 7.5 with
    two -- so the default is one; larger cubes per GPU shift the balance.)

    ``slab``       float64 device tensor (x_hi - x_lo, ny, C), used as the staging cube;
    ``host_rows``  page-locked host tensor of the same shape (e.g. ``dist.HostCube.rows``).
    Returns dict(plans, accept, n_accept, launches); on return the current stream waits for the
    copies, so an event recorded next covers them.
    """
    from . import dist as mdist

    nx, ny, nc = ctx.shape
    x_hi = nx if x_hi is None else x_hi
    gauss = ctx.spectrum == L.SPECTRUM_GAUSSIAN
    kid, valid, sm_range, h_eff = engine.smoothing_setup(dev["sm_length"], ctx.table)
    accept, n_accept = engine.prune(dev["px"], dev["py"], dev["pz"], sm_range, dev["mHI"],
                                    dev["sigma"] if gauss else 0.0, ctx.max_abs_dv, nx, ny, nc, *prune)
    copy_stream = copy_stream or torch.cuda.Stream(device=engine.device)
    main = torch.cuda.current_stream(engine.device)
    bounds = mdist.slab_bounds(x_hi - x_lo, n_slabs)
    plans, launches = [], 2
    slab.zero_()
    for k in range(n_slabs):
        a, b = bounds[k], bounds[k + 1]
        if b == a:
            continue
        part = slab[a:b]
        plans.append(engine.insert(
            px=dev["px"], py=dev["py"], h_eff=h_eff, sm_range=sm_range, v=dev["v"], kernel_id=kid,
            sigma=dev["sigma"] if gauss else 1.0, mHI=dev["mHI"], D=dev["D"], accept=accept,
            table=ctx.table, spectrum=ctx.spectrum, edges=dev["edges"], cube=part,
            px_size_arcsec=ctx.px_size, x_lo=x_lo + a, x_hi=x_lo + b, nx_full=nx, zeroed=True,
            edges_increasing=ctx.edges_increasing))
        launches += engine.last_launches
        done = torch.cuda.Event()
        done.record(main)
        copy_stream.wait_event(done)
        with torch.cuda.stream(copy_stream):
            host_rows[a:b].copy_(part, non_blocking=True)
    main.wait_stream(copy_stream)
    return {"plans": plans, "accept": accept, "n_accept": n_accept, "launches": launches}
