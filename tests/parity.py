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


class PixelOracle:
    """Reference-structured per-pixel sums for big cases (1e7 particles), affordable because
    the candidate scan is pre-filtered.

    The reference selects a pixel's particles with a boolean mask over all N
    (martini.py:272-274).  Here the particles are bucketed once by a spatial hash (classes of
    sm_range, cells at least as wide as the class's largest range, so a pixel's candidates
    live in the 3 x 3 cells around it); the exact predicate of the reference is then applied
    to those candidates only, and the survivors are put in ascending particle order -- an
    ascending index array selects the same elements in the same order as the boolean mask,
    so every later operation (``px_weight``, ``init_spectra``, the sequential
    ``np.sum(axis=-2)``) sees bit-identical operands.  ``tests/test_oracle_golden.py`` checks
    this class against :func:`oracle_pixels` (full mask) on small cases.
    """

    def __init__(self, case, prune=(True, True, True)):
        k, kind, pix, acc = oracle_prepare(case, prune)
        k.apply_mask(acc)
        self.case, self.k, self.kind, self.acc = case, k, kind, acc
        self.p = pix[:, acc]
        sig = case["sigma"]
        self.sig = sig[acc] if np.ndim(sig) > 0 else sig
        self.v, self.mHI, self.D = case["v"][acc], case["mHI"][acc], case["D"][acc]
        r = np.asarray(k.sm_ranges, dtype=np.float64)
        px, py = self.p[0], self.p[1]
        self.always = np.flatnonzero(~np.isfinite(r))  # r = inf: candidate of every pixel
        fin = np.isfinite(r)
        cls = np.zeros(r.shape, dtype=np.int64)
        cls[fin] = np.ceil(np.log2(np.maximum(r[fin], 1.0))).astype(np.int64)  # cell = 2^cls >= r
        self.classes = []
        for c in np.unique(cls[fin]):
            idx = np.flatnonzero(fin & (cls == c))
            s = float(2 ** int(c))
            cx = np.floor(px[idx] / s).astype(np.int64)
            cy = np.floor(py[idx] / s).astype(np.int64)
            cx0, cy0 = cx.min() - 1, cy.min() - 1
            ncy = int(cy.max() - cy0 + 3)
            key = (cx - cx0) * ncy + (cy - cy0)
            order = np.argsort(key, kind="stable")  # ascending particle index inside a cell
            self.classes.append((s, cx0, cy0, ncy, key[order], idx[order]))

    def candidates(self, i, j):
        out = [self.always]
        for s, cx0, cy0, ncy, keys, idx in self.classes:
            ci, cj = int(np.floor(i / s)) - cx0, int(np.floor(j / s)) - cy0
            for a in (ci - 1, ci, ci + 1):
                lo = np.searchsorted(keys, a * ncy + max(cj - 1, 0), side="left")
                hi = np.searchsorted(keys, a * ncy + cj + 1, side="right")
                if hi > lo:
                    out.append(idx[lo:hi])
        return np.concatenate(out) if out else np.zeros(0, dtype=np.int64)

    def select(self, i, j):
        """Ascending indices of the particles the reference's mask selects for pixel (i, j)."""
        cand = self.candidates(i, j)
        ijc = np.array((i, j))[..., np.newaxis]
        m = (np.abs(ijc - self.p[:2, cand]) <= self.k.sm_ranges[cand]).all(axis=0)
        return np.sort(cand[m])

    #: rows per block of the per-pixel spectra array (keeps the temporaries in cache for pixels
    #: with 1e5 contributors).  np.sum(axis=-2) adds the rows one after the other, so carrying
    #: the running sum in as row 0 of the next block reproduces the one-shot sum bit for bit
    #: (tests/test_pixel_oracle.py).
    BLOCK = 4096

    def pixel(self, ij, block=None):
        i, j = int(ij[0]), int(ij[1])
        sel = self.select(i, j)
        ijc = np.array((i, j))[..., np.newaxis]
        w = self.k.px_weight(self.p[:2, sel] - ijc, mask=sel)
        block = block or self.BLOCK
        run = None
        for a in range(0, max(sel.size, 1), block):
            s = sel[a:a + block]
            sp = O.init_spectra(self.kind, self.case["edges"], self.v[s],
                                self.sig if np.ndim(self.sig) == 0 else self.sig[s],
                                self.mHI[s], self.D[s])
            np.multiply(sp, w[a:a + block, np.newaxis], out=sp)
            if run is not None:
                sp = np.concatenate((run[np.newaxis], sp), axis=0)
            run = np.sum(sp, axis=-2)
        return run / self.case["px_size"] ** 2

    def pixels(self, pixels, threads=None):
        import os
        from concurrent.futures import ThreadPoolExecutor

        threads = threads or min(32, os.cpu_count() or 1)
        if threads == 1 or len(pixels) < 4:
            return np.array([self.pixel(ij) for ij in pixels])
        with ThreadPoolExecutor(threads) as pool:
            return np.array(list(pool.map(self.pixel, pixels, chunksize=8)))

    def total_flux(self, chunk=100_000, threads=None):
        """Sum of the whole cube [Jy/arcsec^2 summed over voxels] without building it:
        sum_p (sum over the pixels of p's clipped candidate box of W_p) x (sum_c S_p(c)) /
        px_size^2, every factor from the same oracle functions the per-pixel path uses;
        accumulated with math.fsum.  (A re-association of the reference's sum: good to ~1e-13
        relative, the check it serves is the north-star's 1e-9.)"""
        import math
        import os
        from concurrent.futures import ThreadPoolExecutor

        X, Y, C = self.case["shape"]
        px, py = self.p[0], self.p[1]
        r = self.k.sm_ranges

        def part(a):
            b = min(px.size, a + chunk)
            sl = np.arange(a, b)
            ilo = np.maximum(0, np.ceil(px[sl] - r[sl])).astype(np.int64)
            ihi = np.minimum(X - 1, np.floor(px[sl] + r[sl])).astype(np.int64)
            jlo = np.maximum(0, np.ceil(py[sl] - r[sl])).astype(np.int64)
            jhi = np.minimum(Y - 1, np.floor(py[sl] + r[sl])).astype(np.int64)
            nxp, nyp = np.maximum(0, ihi - ilo + 1), np.maximum(0, jhi - jlo + 1)
            cnt = nxp * nyp
            tot = int(cnt.sum())
            if tot == 0:
                return np.zeros(0)
            rep = np.repeat(np.arange(sl.size), cnt)
            kk = np.arange(tot) - np.repeat(np.cumsum(cnt) - cnt, cnt)
            nyr = nyp[rep]
            ii, jj = ilo[rep] + kk // nyr, jlo[rep] + kk % nyr
            pid = sl[rep]
            w = self.k.px_weight(self.p[:2, pid] - np.vstack((ii, jj)), mask=pid)
            wsum = np.bincount(rep, weights=w, minlength=sl.size)
            ssum = np.zeros(sl.size)
            for c in range(0, sl.size, 8192):  # spectra in cache-sized row blocks
                s2 = sl[c:c + 8192]
                sp = O.init_spectra(self.kind, self.case["edges"], self.v[s2],
                                    self.sig if np.ndim(self.sig) == 0 else self.sig[s2],
                                    self.mHI[s2], self.D[s2])
                ssum[c:c + 8192] = sp.sum(axis=1)
            return wsum * ssum

        starts = list(range(0, px.size, chunk))
        threads = threads or min(32, os.cpu_count() or 1)
        if threads > 1 and len(starts) > 1:
            with ThreadPoolExecutor(threads) as pool:
                terms = list(pool.map(part, starts))
        else:
            terms = [part(a) for a in starts]
        if not terms:
            return 0.0
        return math.fsum(np.concatenate(terms).tolist()) / self.case["px_size"] ** 2
