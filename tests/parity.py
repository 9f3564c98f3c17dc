"""Oracle side of the parity tests: run a synthetic *case* through oracle/martini_oracle.py.

Test infrastructure (imports ``oracle``); never imported by the product package.
"""

from __future__ import annotations

import numpy as np

from oracle import martini_oracle as O

SPEC = {"gaussian": O.SPEC_GAUSSIAN, "diracdelta": O.SPEC_DIRACDELTA}


def oracle_prepare(case, prune=(True, True, True)):
    """Kernel set-up + prune mask, the part of Martini.__init__ on the hot path."""
    name, kw = case["kernel"]
    k = O.make_kernel(name, **kw)
    k.init_sm(case["sm_length"])
    nx, ny, nc = case["shape"]
    kind = SPEC[case["spectrum"]]
    hw = case["sigma"] if kind == O.SPEC_GAUSSIAN else 0.0
    pix = np.vstack((case["px"], case["py"], case["pz"]))
    acc = O.prune_mask(pix, k.sm_ranges, case["mHI"], nx, ny, nc, hw,
                       np.max(np.abs(np.diff(case["edges"]))), *prune)
    return k, kind, pix, acc


def oracle_hot_path(case, cube0=None, prune=(True, True, True)):
    """Full oracle run (candidate-list variant: bit-identical to the reference-structured
    loop, see tests/test_oracle_golden.py).  Returns dict(cube, accept, kernel, updates)."""
    k, kind, pix, acc = oracle_prepare(case, prune)
    sm_ranges0 = k.sm_ranges.copy()
    kidx0 = getattr(k, "kernel_indices", None)
    kidx0 = None if kidx0 is None else kidx0.copy()
    k.apply_mask(acc)
    nx, ny, nc = case["shape"]
    sig = case["sigma"]
    sig = sig[acc] if np.ndim(sig) > 0 else sig
    if cube0 is None:
        cube0 = np.zeros((nx, ny, nc))
    cube = O.insert_fast(cube0, pix[:, acc], k, kind, case["edges"], case["v"][acc], sig,
                         case["mHI"][acc], case["D"][acc], case["px_size"])
    upd = O.count_updates(pix[:, acc], k.sm_ranges, nx, ny, nc)
    return {"cube": cube, "accept": acc, "kernel": k, "updates": upd, "sm_ranges": sm_ranges0,
            "kernel_indices": kidx0}


def oracle_pixels(case, pixels, prune=(True, True, True)):
    """Reference-structured spectra [Jy/arcsec^2] of selected pixels only (for big cases)."""
    k, kind, pix, acc = oracle_prepare(case, prune)
    k.apply_mask(acc)
    sig = case["sigma"]
    sig = sig[acc] if np.ndim(sig) > 0 else sig
    p = pix[:, acc]
    out = []
    for ij in pixels:
        ijc = np.array(ij)[..., np.newaxis]
        mask = (np.abs(ijc - p[:2]) <= k.sm_ranges).all(axis=0)
        sel = np.flatnonzero(mask)
        w = k.px_weight(p[:2, sel] - ijc, mask=sel)
        sp = O.init_spectra(kind, case["edges"], case["v"][acc][sel],
                            sig if np.ndim(sig) == 0 else sig[sel], case["mHI"][acc][sel],
                            case["D"][acc][sel])
        np.multiply(sp, w[:, np.newaxis], out=sp)
        out.append(np.sum(sp, axis=-2) / case["px_size"] ** 2)
    return np.array(out)


def check_cube(cube, ref, rtol_voxel=1e-6, rtol_flux=1e-9):
    """The north-star tolerance: per voxel |d| <= 1e-6 x cube peak, total flux to 1e-9."""
    peak = np.abs(ref).max()
    err = np.abs(cube - ref).max()
    assert err <= rtol_voxel * peak, f"max|d| = {err:.3e} > {rtol_voxel:g} x peak {peak:.3e}"
    s = ref.sum()
    assert abs(cube.sum() - s) <= rtol_flux * abs(s), (cube.sum(), s)
    return err / peak if peak > 0 else 0.0
