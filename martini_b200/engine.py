"""Array-level host API over the C ABI: device buffers are torch tensors, the work is CUDA.

This is the layer the MARTINI-compatible classes (``martini_b200.martini.Martini``) and
``bench.py`` call.  All arrays are float64 in pixel / km/s / Mpc / Msun units (see
include/martini_b200.h).  Nothing here computes on the CPU: every method enqueues kernels of
``libmartini_b200.so`` on torch's current CUDA stream.
"""

from __future__ import annotations

import contextlib
import ctypes as C
import functools
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib as L


@dataclass
class KernelTable:
    """Host description of the SPH kernel(s) in use (-> ``MtnKernelTable``)."""

    entries: list = field(default_factory=list)  # dicts: kind, valid_is_max, rescale, ...
    adaptive: bool = False

    def to_c(self) -> L.MtnKernelTable:
        if not 1 <= len(self.entries) <= L.MTN_MAX_KERNELS:
            raise ValueError("kernel table needs 1..8 entries")
        t = L.MtnKernelTable()
        t.n = len(self.entries)
        t.adaptive = 1 if self.adaptive else 0
        for i, e in enumerate(self.entries):
            t.k[i].kind = int(e["kind"])
            t.k[i].valid_is_max = int(e.get("valid_is_max", 0))
            t.k[i].rescale = float(e.get("rescale", 1.0))
            t.k[i].size_in_fwhm = float(e["size_in_fwhm"])
            t.k[i].valid_size = float(e.get("valid_size", 0.0))
            t.k[i].truncate = float(e.get("truncate", 0.0))
            t.k[i].norm = float(e.get("norm", 1.0))
        return t


class MergedPlan:
    """The plans of the x-sub-slabs an insertion was split into, summed (``n_kept`` counts a
    particle once per sub-slab it reaches)."""

    def __init__(self, parts):
        self.parts = parts
        for k in ("n_kept", "n_pairs", "n_pairs2", "updates_dense", "workspace_bytes"):
            setattr(self, k, sum(getattr(p, k) for p in parts))
        self.route2 = parts[0].route2
        self.n_bricks = sum(p.n_bricks for p in parts)
        self.edges_increasing = parts[0].edges_increasing


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _scalar_or_tensor(x, device):
    """Return (tensor_or_None, scalar) for an input that may be a scalar or an array."""
    if x is None:
        return None, 0.0
    if isinstance(x, torch.Tensor):
        if x.ndim == 0:
            return None, float(x)
        return x.to(device=device, dtype=torch.float64).contiguous(), 0.0
    a = np.asarray(x, dtype=np.float64)
    if a.ndim == 0:
        return None, float(a)
    return torch.from_numpy(np.ascontiguousarray(a)).to(device), 0.0


def _on_device(method):
    """Run an Engine method with the engine's GPU as the current CUDA device: the C ABI
    launches on the current device and keeps per-device state (tables, function attributes),
    so two engines on different GPUs may share a host thread."""

    @functools.wraps(method)
    def wrapped(self, *args, **kwargs):
        with self._device_ctx():
            return method(self, *args, **kwargs)

    return wrapped


class Engine:
    """One CUDA device's projection engine (owns grow-only scratch tensors)."""

    def __init__(self, device="cuda:0"):
        self.lib = L.load()
        if not torch.cuda.is_available():
            raise L.MartiniB200Error("martini_b200 needs a CUDA device; there is no CPU fallback")
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self._scratch = None
        self._workspace = None
        self.last_plan = None
        self.last_launches = 0

    # ------------------------------------------------------------------ helpers
    def _check(self, rc, what):
        L.check(rc, what, self.lib)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _device_ctx(self):
        return torch.cuda.device(self.device) if self.device.type == "cuda" else contextlib.nullcontext()

    def to_device(self, x, dtype=torch.float64):
        if isinstance(x, torch.Tensor):
            return x.to(device=self.device, dtype=dtype).contiguous()
        a = np.ascontiguousarray(np.asarray(x))
        if not a.flags.writeable:  # (broadcast views: torch refuses to alias read-only memory)
            a = a.copy()
        t = torch.from_numpy(a)
        if a.dtype != np.dtype("uint8") or dtype != torch.uint8:
            t = t.to(dtype)
        return t.to(self.device, non_blocking=True)

    # ------------------------------------------------------------------ page-locked host buffers
    # A finished cube is read back into page-locked memory (full PCIe rate, no page faults on a
    # fresh allocation).  cudaHostAlloc is slow (~0.3 ms per MB), so buffers are pooled by size:
    # the numpy array handed to the caller owns its buffer until the array (and every view of
    # it) is garbage-collected, then the buffer returns to the pool for the next cube.
    def to_host(self, tensor):
        """Device tensor -> numpy array in pooled page-locked host memory."""
        import weakref

        pool = self.__dict__.setdefault("_pinned_pool", {})
        key = (tuple(tensor.shape), tensor.dtype)
        free = pool.setdefault(key, [])
        buf = free.pop() if free else torch.empty(tensor.shape, dtype=tensor.dtype, pin_memory=self.device.type == "cuda")
        buf.copy_(tensor, non_blocking=True)
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()
        arr = buf.numpy()
        weakref.finalize(arr, free.append, buf)
        return arr

    def _grow(self, name, nbytes):
        buf = getattr(self, name)
        if buf is None or buf.numel() < nbytes:
            buf = None
            setattr(self, name, None)
            buf = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=self.device)
            setattr(self, name, buf)
        return buf

    @_on_device
    def device_info(self):
        sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
        self._check(self.lib.mtn_device_info(C.byref(sm), C.byref(ma), C.byref(mi)), "mtn_device_info")
        return {"sm_count": sm.value, "cc": (ma.value, mi.value)}

    # ------------------------------------------------------------------ K0 / K1
    @_on_device
    def smoothing_setup(self, sm_length: torch.Tensor, table: KernelTable):
        """-> (kernel_id u8, valid u8, sm_range f64, h_eff f64); see mtn_smoothing_setup."""
        n = sm_length.numel()
        kid = torch.empty(n, dtype=torch.uint8, device=self.device)
        valid = torch.empty(n, dtype=torch.uint8, device=self.device)
        rng = torch.empty(n, dtype=torch.float64, device=self.device)
        heff = torch.empty(n, dtype=torch.float64, device=self.device)
        tc = table.to_c()
        self._check(
            self.lib.mtn_smoothing_setup(n, _ptr(sm_length), C.byref(tc), _ptr(kid), _ptr(valid),
                                         _ptr(rng), _ptr(heff), self._stream()),
            "mtn_smoothing_setup",
        )
        return kid, valid, rng, heff

    @_on_device
    def prune(self, px, py, pz, sm_range, mHI, half_width, max_abs_dv, nx_tot, ny_tot, n_channels,
              spatial=True, spectral=True, mass=True):
        """-> (accept u8 tensor, n_accept 0-d int64 tensor); see mtn_prune."""
        n = px.numel()
        mt, ms = _scalar_or_tensor(mHI, self.device)
        ht, hs = _scalar_or_tensor(half_width, self.device)
        accept = torch.empty(n, dtype=torch.uint8, device=self.device)
        count = torch.zeros((), dtype=torch.int64, device=self.device)
        flags = (L.PRUNE_SPATIAL if spatial else 0) | (L.PRUNE_SPECTRAL if spectral else 0) | (
            L.PRUNE_MASS if mass else 0)
        self._check(
            self.lib.mtn_prune(n, _ptr(px), _ptr(py), _ptr(pz), _ptr(sm_range), _ptr(mt), ms,
                               _ptr(ht), hs, float(max_abs_dv), int(nx_tot), int(ny_tot),
                               int(n_channels), flags, _ptr(accept), _ptr(count), self._stream()),
            "mtn_prune",
        )
        return accept, count

    # ------------------------------------------------------------------ plan + project
    #: test hook: treat a slab with more (particle, key) pairs than this like one that exceeds the
    #: library's 32-bit pair index, i.e. split it (None: only the library's own limit)
    pair_limit = None

    def insert(self, *, cube, x_lo=0, x_hi=None, nx_full=None, **kw):
        """Project particles into ``cube`` (rows [x_lo, x_hi) of the full cube); see
        :meth:`_insert_slab`.  A slab whose sorted pair list would exceed the 32-bit index of the
        radix sort (2^32 - 1 pairs per stream: wide footprints on a big cube) is cut into
        x-sub-slabs, each planned and projected on its own -- every sub-slab writes only its own
        rows, halo particles are simply planned for both, so the result is the unsplit one."""
        if x_hi is None:
            x_hi = x_lo + cube.shape[0]
        nx_full = int(nx_full if nx_full is not None else x_hi)
        plan = self._insert_slab(cube=cube, x_lo=x_lo, x_hi=x_hi, nx_full=nx_full, **kw)
        if plan is not None:
            return plan
        rows = x_hi - x_lo
        if rows <= 8:
            raise L.MartiniB200Error("insert: a slab of 8 rows still exceeds the 32-bit pair index")
        mid = x_lo + max(8, (rows // 2 + 7) // 8 * 8)  # cuts stay on brick boundaries
        parts = [self.insert(cube=cube[a - x_lo:b - x_lo], x_lo=a, x_hi=b, nx_full=nx_full, **kw)
                 for a, b in ((x_lo, mid), (mid, x_hi))]
        merged = MergedPlan(parts)
        self.last_plan = merged
        return merged

    @_on_device
    def _insert_slab(self, *, px, py, h_eff, sm_range, v, kernel_id=None, sigma=None, mHI=None, D=None,
               accept=None, table: KernelTable, spectrum: int, edges: torch.Tensor,
               cube: torch.Tensor, px_size_arcsec: float, x_lo: int = 0, x_hi: int | None = None,
               nx_full: int | None = None, zeroed: bool = False, edges_increasing: bool | None = None):
        """Project particles into ``cube`` (a (x_hi-x_lo, ny, C) float64 device tensor holding
        rows [x_lo, x_hi) of the full cube), in place:  cube = (cube + inserted) / px_size^2.

        Returns the ``MtnPlan`` (n_kept, n_pairs, updates_dense, ...), or None if the slab has
        more pairs than the sort can index (``insert`` then splits it).
        """
        assert cube.is_contiguous() and cube.dtype == torch.float64 and cube.ndim == 3
        n = px.numel()
        st, ss = _scalar_or_tensor(sigma, self.device)
        mt, ms = _scalar_or_tensor(mHI, self.device)
        dt, ds = _scalar_or_tensor(D, self.device)
        keep = (st, mt, dt)  # keep temporaries alive until the kernels are enqueued
        p = L.MtnParticles()
        p.n = n
        p.px, p.py, p.h_eff, p.sm_range = _ptr(px), _ptr(py), _ptr(h_eff), _ptr(sm_range)
        p.kernel_id = _ptr(kernel_id)
        p.v = _ptr(v)
        p.sigma, p.sigma_scalar = _ptr(st), ss
        p.mHI, p.mHI_scalar = _ptr(mt), ms
        p.D, p.D_scalar = _ptr(dt), ds
        p.accept = _ptr(accept)
        c = L.MtnCube()
        if x_hi is None:
            x_hi = x_lo + cube.shape[0]
        c.nx = int(nx_full if nx_full is not None else x_hi)
        c.ny, c.n_channels = int(cube.shape[1]), int(cube.shape[2])
        c.x_lo, c.x_hi = int(x_lo), int(x_hi)
        assert c.x_hi - c.x_lo == cube.shape[0] and edges.numel() == c.n_channels + 1
        c.spectrum = int(spectrum)
        c.flags = L.CUBE_ZEROED if zeroed else L.CUBE_ACCUMULATE
        c.px_size_arcsec = float(px_size_arcsec)
        c.edges, c.slab = _ptr(edges), _ptr(cube)
        # the caller usually knows the direction of its (host-made) edges: saves a read-back + sync
        c.edges_direction = 0 if edges_increasing is None else (1 if edges_increasing else -1)
        tc = table.to_c()
        stream = self._stream()

        sbytes = self.lib.mtn_plan_scratch_bytes(n, C.byref(c))
        scratch = self._grow("_scratch", sbytes)
        plan = L.MtnPlan()
        rc = self.lib.mtn_plan(C.byref(p), C.byref(tc), C.byref(c), _ptr(scratch), scratch.numel(),
                               C.byref(plan), stream)
        too_many = rc == L.ERR_LIMIT and b"pairs exceed" in self.lib.mtn_last_error()
        if too_many or (rc == 0 and self.pair_limit is not None and max(plan.n_pairs, plan.n_pairs2) > self.pair_limit
                        and c.x_hi - c.x_lo > 8):
            return None  # the caller splits the slab
        self._check(rc, "mtn_plan")
        ws = self._grow("_workspace", plan.workspace_bytes)
        self._check(self.lib.mtn_project(C.byref(p), C.byref(tc), C.byref(c), C.byref(plan),
                                     _ptr(scratch), scratch.numel(), _ptr(ws), ws.numel(), stream),
                "mtn_project")
        self.last_plan = plan
        self.last_launches = self.lib.mtn_last_launch_count() + 4  # + mtn_plan's four kernels
        del keep
        return plan

    # ------------------------------------------------------------------ coordinate front-end
    @_on_device
    def sky_to_pix(self, fe: "L.MtnFrontEnd", xyz, vxyz, hsm):
        """-> (px, py, pz, v, D, sm_length) device tensors; see mtn_sky_to_pix.  ``xyz`` / ``vxyz``:
        (n, 3) float64 device tensors; ``hsm``: (n,) tensor or a scalar."""
        n = xyz.shape[0]
        out = [torch.empty(n, dtype=torch.float64, device=self.device) for _ in range(6)]
        ht, hs = _scalar_or_tensor(hsm, self.device)
        self._check(self.lib.mtn_sky_to_pix(C.byref(fe), n, _ptr(xyz), _ptr(vxyz), _ptr(ht), hs,
                                           *[_ptr(t) for t in out], self._stream()), "mtn_sky_to_pix")
        return tuple(out)

    # ------------------------------------------------------------------ multi-GPU routing
    @_on_device
    def route_count(self, px, sm_range, bounds):
        """-> (totals int64[world] on the device, scratch to hand to route_scatter)."""
        world = len(bounds) - 1
        n = px.numel()
        b = (C.c_int32 * (world + 1))(*[int(x) for x in bounds])
        nbytes = self.lib.mtn_route_scratch_bytes(n, world)
        scratch = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=self.device)
        totals = torch.zeros(world, dtype=torch.int64, device=self.device)
        self._check(self.lib.mtn_route_count(n, _ptr(px), _ptr(sm_range), world, b, _ptr(scratch), scratch.numel(),
                                            _ptr(totals), self._stream()), "mtn_route_count")
        return totals, scratch

    @_on_device
    def route_scatter(self, px, sm_range, bounds, fields, inbox_ptrs, capacity, src_offsets, scratch):
        """Store the routed particles' ``fields`` (device tensors) into the ranks' inboxes
        (``inbox_ptrs``: one device pointer per rank, peer-mapped)."""
        world = len(bounds) - 1
        b = (C.c_int32 * (world + 1))(*[int(x) for x in bounds])
        f = (C.c_void_p * len(fields))(*[t.data_ptr() for t in fields])
        d = (C.c_void_p * world)(*[int(p) for p in inbox_ptrs])
        self._check(self.lib.mtn_route_scatter(px.numel(), _ptr(px), _ptr(sm_range), world, b, len(fields), f, d,
                                              int(capacity), _ptr(src_offsets), _ptr(scratch), self._stream()),
                    "mtn_route_scatter")

    # ------------------------------------------------------------------ beam convolution
    @_on_device
    def convolve_beam(self, cube: torch.Tensor, kernel: torch.Tensor, scale: float = 1.0):
        """out[x,y,c] = scale * (cube[:, :, c] (*) kernel)[x, y], 'same' size; see mtn_convolve_beam."""
        assert cube.ndim == 3 and cube.dtype == torch.float64 and cube.is_contiguous()
        kernel = self.to_device(kernel)
        out = torch.empty_like(cube)
        self._check(self.lib.mtn_convolve_beam(_ptr(cube), _ptr(out), cube.shape[0], cube.shape[1],
                                           cube.shape[2], _ptr(kernel), kernel.shape[0], kernel.shape[1],
                                           float(scale), self._stream()), "mtn_convolve_beam")
        return out

    # ------------------------------------------------------------------ diagnostics
    STAGES = ("emit", "sort", "items", "project", "reduce", "finalize", "stream2")

    def set_timing(self, enable: bool):
        self._check(self.lib.mtn_set_timing(int(enable)), "mtn_set_timing")

    @_on_device
    def last_timing_ms(self):
        """Stage durations [ms] of the last insert() (needs set_timing(True)); synchronises."""
        arr = (C.c_float * 7)()
        self._check(self.lib.mtn_last_timing(arr, 7), "mtn_last_timing")
        return dict(zip(self.STAGES, (float(x) for x in arr)))

    def set_count_exec(self, enable: bool):
        self._check(self.lib.mtn_set_count_exec(int(enable)), "mtn_set_count_exec")

    def last_exec_counts(self):
        arr = (C.c_int64 * 3)()
        self._check(self.lib.mtn_last_exec_counts(arr), "mtn_last_exec_counts")
        return {"updates": int(arr[0]), "weights": int(arr[1]), "erfs": int(arr[2])}

    @_on_device
    def fp64_peak_tflops(self):
        t, ms = C.c_double(), C.c_double()
        self._check(self.lib.mtn_fp64_peak(C.byref(t), C.byref(ms), self._stream()), "mtn_fp64_peak")
        return t.value

    @_on_device
    def probe_kernel_integral(self, entry: dict, dx, dy, h, closed_form=False):
        e = KernelTable([entry]).to_c().k[0]
        dx, dy, h = (self.to_device(a) for a in (dx, dy, h))
        out = torch.empty_like(dx)
        self._check(self.lib.mtn_probe_kernel_integral(C.byref(e), int(closed_form), dx.numel(), _ptr(dx),
                                                   _ptr(dy), _ptr(h), _ptr(out), self._stream()),
                "mtn_probe_kernel_integral")
        return out

    @_on_device
    def table_error(self, kind: int) -> float:
        err = C.c_double()
        self._check(self.lib.mtn_table_error(int(kind), C.byref(err)), "mtn_table_error")
        return err.value

    @_on_device
    def probe_spectra(self, spectrum, v, sigma, amp, edges):
        v, amp, edges = (self.to_device(a) for a in (v, amp, edges))
        st, ss = _scalar_or_tensor(sigma, self.device)
        nchan = edges.numel() - 1
        out = torch.empty((v.numel(), nchan), dtype=torch.float64, device=self.device)
        self._check(self.lib.mtn_probe_spectra(int(spectrum), v.numel(), _ptr(v), _ptr(st), ss,
                                           _ptr(amp), nchan, _ptr(edges), _ptr(out), self._stream()),
                "mtn_probe_spectra")
        return out
