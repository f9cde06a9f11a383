"""Host-side mirror of the reference's hot path over the C ABI (include/tofrt.h).

`TofEngine` = one tofrt context (one GPU).  `setup()` plays the role of the
per-iteration setup of main_rt.py:469-479 plus the module globals
(main_rt.py:449-462); `tof_table()` is the sweep body main_rt.py:464-501 for a
batch of geometries; `shoot_rays()` returns the dict of main_rt.py:432-441.
"""
import ctypes

import numpy as np

from . import _capi
from . import reference_setup as rs

RECORD_KEYS = ("pipe_x", "pipe_z", "lens2_x", "lens2_z", "target_x", "idx", "legs")


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def pinned_empty(shape, dtype=np.float64):
    """NumPy array in page-locked host memory (tofrt_host_alloc): pass such arrays as `out=` to
    TofEngine.tof_table so that the result copies overlap the kernels.  Freed with the array."""
    import weakref
    lib = _capi.load()
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    ptr = lib.tofrt_host_alloc(max(nbytes, 1))
    if not ptr:
        raise MemoryError("tofrt_host_alloc failed")
    buf = (ctypes.c_char * max(nbytes, 1)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    weakref.finalize(buf, lib.tofrt_host_free, ptr)
    return arr


class TofTable:
    """Element x geometry table, geometry-major: tof[g, e], k_hit[g, e]."""

    def __init__(self, tof, k_hit, records=None):
        self.tof = tof
        self.k_hit = k_hit
        self.records = records or {}

    @property
    def hit(self):
        return self.k_hit >= 0


class TofEngine:
    def __init__(self, device=0):
        self._lib = _capi.load()
        if self._lib.tofrt_device_count() <= device:
            raise RuntimeError(f"CUDA device {device} not available (the engine has no CPU fallback)")
        self._ctx = ctypes.c_void_p()
        _capi.check(self._lib.tofrt_create(ctypes.byref(self._ctx), int(device)))
        self.device = int(device)
        self.K = self.E = 0
        self.S, self.multi = 1, False

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.tofrt_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- main_rt.py:449-479 -------------------------------------------------
    def setup(self, alpha, z_f, x_shooter, z_shooter, x_rx, c1=rs.C1, c2=rs.C2, l0=rs.L0, h0=rs.H0,
              d=None, rtol=1e-5, atol=1e-6):
        """x_shooter / z_shooter: scalars (the reference: one shot, main_rt.py:481-482) or equal-length
        arrays (full-matrix capture: every listed element shoots; tables gain a leading S axis)."""
        alpha, z_f, x_rx = _f64(alpha), _f64(z_f), _f64(x_rx)
        if alpha.ndim != 1 or z_f.shape != alpha.shape:
            raise ValueError("Coordinate arrays must have the same length.")
        l0, h0 = np.float64(l0), np.float64(h0)
        d = np.float64(l0 + h0) if d is None else np.float64(d)
        self.params = _capi.Params(float(c1), float(c2), float(l0), float(h0), float(d))
        self.multi = np.ndim(x_shooter) > 0
        xs, zs = _f64(np.atleast_1d(x_shooter)), _f64(np.atleast_1d(z_shooter))
        if xs.shape != zs.shape or xs.ndim != 1:
            raise ValueError("x_shooter and z_shooter must have the same length")
        self.x_shooter, self.z_shooter = xs, zs
        self.alpha, self.z_f, self.x_rx = alpha, z_f, x_rx
        _capi.check(self._lib.tofrt_setup_shooters(
            self._ctx, ctypes.byref(self.params), alpha.ctypes.data, z_f.ctypes.data, len(alpha),
            xs.ctypes.data, zs.ctypes.data, len(xs), x_rx.ctypes.data, len(x_rx), float(rtol), float(atol)))
        self.K, self.E, self.S = len(alpha), len(x_rx), len(xs)
        return self

    def _shape(self, G, last):
        return (self.S, G, last) if self.multi else (G, last)

    def _check_out(self, t, shape, dtype, dev, what):
        """A caller-supplied output tensor must be exactly what the kernel writes: anything else would be an
        out-of-bounds or mis-typed device write."""
        if (tuple(t.shape) != tuple(shape) or t.dtype != dtype or not t.is_cuda or t.device != dev or
                not t.is_contiguous()):
            raise ValueError(f"{what} must be a contiguous {dtype} CUDA tensor of shape {tuple(shape)} on {dev}")
        return t

    @staticmethod
    def _check_records(records):
        for key in records:
            if key not in RECORD_KEYS:
                raise ValueError(f"unknown record {key!r}")

    def set_surface(self, kind):
        """'lens' (the reference's constant-TOF lens, default) or 'flat' (plane z = h0; extension).
        Call before setup()."""
        _capi.check(self._lib.tofrt_set_option(self._ctx, 2, {"lens": 0, "flat": 1}[kind]))
        return self

    def set_math(self, mode):
        """'angle' (the reference's angle bookkeeping, in its operation order) or 'vector' (the same interface
        laws from direction cosines, csrc/tofrt_raymath.h: faster, same decisions); see TOFRT_OPT_MATH."""
        _capi.check(self._lib.tofrt_set_option(self._ctx, 3, {"angle": 0, "vector": 1}[mode]))
        return self

    def set_small_g(self, on):
        """Small-G latency mode (TOFRT_OPT_SMALL_G, default on): few geometries run a CTA per geometry."""
        _capi.check(self._lib.tofrt_set_option(self._ctx, 5, 1 if on else 0))
        return self

    @property
    def math_mode(self):
        return "vector" if self._lib.tofrt_get_info(self._ctx, 3) == 1 else "angle"

    def set_search(self, mode):
        """'guided' (default) or 'scan' (the reference's literal O(N) search); same results."""
        value = {"scan": 0, "guided": 1}[mode]
        _capi.check(self._lib.tofrt_set_option(self._ctx, 1, value))
        return self

    @property
    def lens_shape(self):
        """+1 concave, -1 convex, 0 neither (guided search falls back to the scan)."""
        return self._lib.tofrt_get_info(self._ctx, 1)

    STAT_KEYS = ("rays", "bracket_rays", "probes", "exact_diffs", "lens2_rays", "landing_rays", "receiver_tests",
                 "winner_rays", "fallback_rays", "scan_rays", "on_sample_rays")

    def set_stats(self, on):
        """Run the instrumented instantiation of the table kernel (same results, a little slower) so that
        last_stats() can report the executed work; see TOFRT_OPT_STATS."""
        _capi.check(self._lib.tofrt_set_option(self._ctx, 4, 1 if on else 0))
        return self

    def last_stats(self):
        """Executed-work counters of the last table call made with set_stats(True) (see tofrt_last_stats)."""
        arr = (ctypes.c_ulonglong * len(self.STAT_KEYS))()
        _capi.check(self._lib.tofrt_last_stats(self._ctx, arr, len(self.STAT_KEYS)))
        return dict(zip(self.STAT_KEYS, (int(x) for x in arr)))

    def debug_search(self, m, b, with_points=False):
        """Test hook: find_line_curve_intersection of the current setup (current search mode) applied
        to arbitrary lines y = m x + b -> idx (and the intersection points)."""
        m, b = _f64(m).ravel(), _f64(b).ravel()
        idx = np.empty(len(m), dtype=np.int32)
        xi = np.empty(len(m)) if with_points else None
        zi = np.empty(len(m)) if with_points else None
        _capi.check(self._lib.tofrt_debug_search_host(self._ctx, m.ctypes.data, b.ctypes.data, len(m), idx.ctypes.data,
                                                      xi.ctypes.data if with_points else None,
                                                      zi.ctypes.data if with_points else None))
        return (idx, xi, zi) if with_points else idx

    def lens(self):
        """lens_1_x, lens_1_z (main_rt.py:338)."""
        x, z = np.empty(self.K), np.empty(self.K)
        _capi.check(self._lib.tofrt_lens_host(self._ctx, x.ctypes.data, z.ctypes.data))
        return x, z

    def kernel_info(self):
        v = [ctypes.c_int() for _ in range(4)]
        _capi.check(self._lib.tofrt_kernel_info(self._ctx, *[ctypes.byref(x) for x in v]))
        return dict(zip(("grid", "block", "smem_bytes", "regs"), (x.value for x in v)))

    def measure_fp64_peak(self):
        """FP64 thread-instructions/s of this GPU: (DFMA stream, un-fused DMUL+DADD stream)."""
        a, b = ctypes.c_double(), ctypes.c_double()
        _capi.check(self._lib.tofrt_measure_fp64_peak(self._ctx, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def measure_fp32_peak(self):
        """FP32 thread-instructions/s of this GPU: (FFMA stream, un-fused FMUL+FADD stream)."""
        a, b = ctypes.c_double(), ctypes.c_double()
        _capi.check(self._lib.tofrt_measure_fp32_peak(self._ctx, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    # ---- main_rt.py:464-501, host buffers -------------------------------------
    def tof_table(self, radii, offsets, records=(), out=None):
        """out = (tof, k_hit): preallocated C-contiguous arrays (float64 / int32), e.g. from
        `pinned_empty`, reused across calls."""
        radii, offsets = _f64(radii).ravel(), _f64(offsets).ravel()
        if radii.shape != offsets.shape:
            raise ValueError("radii and offsets must have the same length")
        G = len(radii)
        if out is None:
            tof = np.empty(self._shape(G, self.E))
            k_hit = np.empty(self._shape(G, self.E), dtype=np.int32)
        else:
            tof, k_hit = out
            if (tof.shape != self._shape(G, self.E) or k_hit.shape != tof.shape or tof.dtype != np.float64 or
                    k_hit.dtype != np.int32 or not tof.flags.c_contiguous or not k_hit.flags.c_contiguous):
                raise ValueError("out must be C-contiguous (float64, int32) arrays of the table's shape")
        rec, out = None, {}
        if records:
            rec = _capi.RayRecords()
            for key in records:
                if key not in RECORD_KEYS:
                    raise ValueError(f"unknown record {key!r}")
                shape = self._shape(G, 4 * self.K if key == "legs" else self.K)
                if key == "legs":
                    shape = shape[:-1] + (4, self.K)
                out[key] = np.empty(shape, dtype=np.int32 if key == "idx" else np.float64)
                setattr(rec, key, out[key].ctypes.data)
        _capi.check(self._lib.tofrt_table_host(self._ctx, radii.ctypes.data, offsets.ctypes.data, G,
                                               tof.ctypes.data, k_hit.ctypes.data,
                                               ctypes.byref(rec) if rec is not None else None))
        return TofTable(tof, k_hit, out)

    # ---- sparse form: only the hits, ordered by (row, receiver) ---------------------------
    def tof_hits(self, radii, offsets, capacity=None, buffers=None):
        """(row, e, k, tof) arrays of the table entries with a hit; row = geometry index
        (shooter * G + geometry with several shooters).  `buffers` = preallocated (pinned) arrays."""
        radii, offsets = _f64(radii).ravel(), _f64(offsets).ravel()
        if radii.shape != offsets.shape:
            raise ValueError("radii and offsets must have the same length")
        G = len(radii)
        n = ctypes.c_longlong(0)
        while True:
            if buffers is None:
                cap = int(capacity) if capacity else max(1024, G * self.S * self.E // 16)
                bufs = (np.empty(cap, np.int32), np.empty(cap, np.int32), np.empty(cap, np.int32), np.empty(cap))
            else:
                bufs, cap = buffers, len(buffers[0])
            rc = self._lib.tofrt_hits_host(self._ctx, radii.ctypes.data, offsets.ctypes.data, G, cap,
                                           bufs[0].ctypes.data, bufs[1].ctypes.data, bufs[2].ctypes.data,
                                           bufs[3].ctypes.data, ctypes.byref(n))
            if rc == _capi.E_TOOLARGE and n.value > cap and buffers is None:
                capacity = n.value  # retry once with the exact size
                continue
            _capi.check(rc)
            return tuple(b[: n.value] for b in bufs)

    # ---- same, device-resident torch tensors, async on the current stream -------
    def tof_table_cuda(self, radii, offsets, out_tof=None, out_k_hit=None, records=()):
        import torch
        if not (radii.is_cuda and offsets.is_cuda and radii.dtype == torch.float64 and
                offsets.dtype == torch.float64 and radii.is_contiguous() and offsets.is_contiguous()):
            raise ValueError("radii/offsets must be contiguous float64 CUDA tensors")
        if radii.device.index != self.device or radii.shape != offsets.shape:
            raise ValueError("radii/offsets must live on the engine's device and have the same shape")
        G = radii.numel()
        dev = radii.device
        shape = self._shape(G, self.E)
        tof = (self._check_out(out_tof, shape, torch.float64, dev, "out_tof") if out_tof is not None
               else torch.empty(shape, dtype=torch.float64, device=dev))
        k_hit = (self._check_out(out_k_hit, shape, torch.int32, dev, "out_k_hit") if out_k_hit is not None
                 else torch.empty(shape, dtype=torch.int32, device=dev))
        self._check_records(records)
        rec, out = None, {}
        if records:
            rec = _capi.RayRecords()
            for key in records:
                shape = self._shape(G, self.K)
                if key == "legs":
                    shape = shape[:-1] + (4, self.K)
                out[key] = torch.empty(shape, dtype=torch.int32 if key == "idx" else torch.float64, device=dev)
                setattr(rec, key, out[key].data_ptr())
        stream = torch.cuda.current_stream(dev).cuda_stream
        _capi.check(self._lib.tofrt_table_dev(self._ctx, radii.data_ptr(), offsets.data_ptr(), G,
                                              tof.data_ptr(), k_hit.data_ptr(),
                                              ctypes.byref(rec) if rec is not None else None,
                                              ctypes.c_void_p(stream)))
        return TofTable(tof, k_hit, out)

    # ---- FP32 mode (reported separately: no bit-parity claim) ----------------------
    def tof_table_f32_cuda(self, radii, offsets, out_tof=None, out_k_hit=None):
        """Same table with the per-ray arithmetic in float32; radii/offsets are the float64 CUDA tensors."""
        import torch
        if not (radii.is_cuda and offsets.is_cuda and radii.dtype == torch.float64 and
                offsets.dtype == torch.float64 and radii.is_contiguous() and offsets.is_contiguous()):
            raise ValueError("radii/offsets must be contiguous float64 CUDA tensors")
        if radii.device.index != self.device or radii.shape != offsets.shape:
            raise ValueError("radii/offsets must live on the engine's device and have the same shape")
        G, dev = radii.numel(), radii.device
        shape = self._shape(G, self.E)
        tof = (self._check_out(out_tof, shape, torch.float32, dev, "out_tof") if out_tof is not None
               else torch.empty(shape, dtype=torch.float32, device=dev))
        k_hit = (self._check_out(out_k_hit, shape, torch.int32, dev, "out_k_hit") if out_k_hit is not None
                 else torch.empty(shape, dtype=torch.int32, device=dev))
        stream = torch.cuda.current_stream(dev).cuda_stream
        _capi.check(self._lib.tofrt_table_f32_dev(self._ctx, radii.data_ptr(), offsets.data_ptr(), G,
                                                  tof.data_ptr(), k_hit.data_ptr(), ctypes.c_void_p(stream)))
        return TofTable(tof, k_hit)

    # ---- two-point Fermat solve (extension; parity unpinned) ---------------
    def fermat_tof(self, fx, fz, with_alpha=False):
        """One-way travel time element -> lens -> focal point for every (receiver, focal point) pair:
        returns (tof[E, F], kmin[E, F][, alpha[E, F]]), host arrays."""
        fx, fz = _f64(fx).ravel(), _f64(fz).ravel()
        if fx.shape != fz.shape:
            raise ValueError("fx and fz must have the same length")
        F = len(fx)
        tof = np.empty((self.E, F))
        kmin = np.empty((self.E, F), dtype=np.int32)
        alpha = np.empty((self.E, F)) if with_alpha else None
        _capi.check(self._lib.tofrt_fermat_host(self._ctx, fx.ctypes.data, fz.ctypes.data, F, tof.ctypes.data,
                                                kmin.ctypes.data, alpha.ctypes.data if with_alpha else None))
        return (tof, kmin, alpha) if with_alpha else (tof, kmin)

    def fermat_tof_cuda(self, fx, fz, with_alpha=False):
        """Same, device-resident torch tensors, asynchronous on the current stream."""
        import torch
        if not (fx.is_cuda and fz.is_cuda and fx.dtype == torch.float64 and fz.dtype == torch.float64 and
                fx.is_contiguous() and fz.is_contiguous() and fx.shape == fz.shape):
            raise ValueError("fx/fz must be contiguous float64 CUDA tensors of the same shape")
        F, dev = fx.numel(), fx.device
        tof = torch.empty((self.E, F), dtype=torch.float64, device=dev)
        kmin = torch.empty((self.E, F), dtype=torch.int32, device=dev)
        alpha = torch.empty((self.E, F), dtype=torch.float64, device=dev) if with_alpha else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        _capi.check(self._lib.tofrt_fermat_dev(self._ctx, fx.data_ptr(), fz.data_ptr(), F, tof.data_ptr(),
                                               kmin.data_ptr(), alpha.data_ptr() if with_alpha else None,
                                               ctypes.c_void_p(stream)))
        return (tof, kmin, alpha) if with_alpha else (tof, kmin)

    # ---- 3-D two-point solve (extension, SURVEY row f4 / BASELINE config 5; parity unpinned) ----
    def setup_fermat3d(self, ex, ey, n_side=9, ang=None):
        """Matrix-array elements (ex[e], ey[e], d) and the n_side x n_side grid of surface samples
        (default: np.linspace over the launch-angle range of setup()) for fermat3d_tof."""
        ex, ey = _f64(ex).ravel(), _f64(ey).ravel()
        if ex.shape != ey.shape:
            raise ValueError("ex and ey must have the same length")
        if ang is None:
            ang = np.linspace(self.alpha[0], self.alpha[-1], int(n_side))
        ang = _f64(ang).ravel()
        _capi.check(self._lib.tofrt_fermat3d_setup(self._ctx, ang.ctypes.data, len(ang), ex.ctypes.data,
                                                   ey.ctypes.data, len(ex)))
        self.f3_ang, self.f3_E = ang, len(ex)
        return self

    def fermat3d_tof(self, fx, fy, fz, with_kmin=True, with_angles=False):
        """One-way travel time element -> lens -> voxel for every (element, voxel) pair:
        (tof[E3, F][, kmin[E3, F]][, ab[E3, F, 2]]), host arrays.  Pairs in the plane y = 0 are solved by the 2-D
        solver (fermat_tof) when both search the same aperture: kmin = -1 - j, ab = (alpha, 0); see include/tofrt.h."""
        fx, fy, fz = _f64(fx).ravel(), _f64(fy).ravel(), _f64(fz).ravel()
        if not (fx.shape == fy.shape == fz.shape):
            raise ValueError("fx, fy and fz must have the same length")
        F, E = len(fx), self.f3_E
        tof = np.empty((E, F))
        kmin = np.empty((E, F), dtype=np.int32) if with_kmin else None
        ab = np.empty((E, F, 2)) if with_angles else None
        _capi.check(self._lib.tofrt_fermat3d_host(self._ctx, fx.ctypes.data, fy.ctypes.data, fz.ctypes.data, F,
                                                  tof.ctypes.data, kmin.ctypes.data if with_kmin else None,
                                                  ab.ctypes.data if with_angles else None))
        return tuple(x for x in (tof, kmin, ab) if x is not None) if (with_kmin or with_angles) else tof

    def fermat3d_tof_cuda(self, fx, fy, fz, with_kmin=False, with_angles=False, out_tof=None):
        """Same, device-resident torch tensors, asynchronous on the current stream."""
        import torch
        for v in (fx, fy, fz):
            if not (v.is_cuda and v.dtype == torch.float64 and v.is_contiguous() and v.shape == fx.shape):
                raise ValueError("fx/fy/fz must be contiguous float64 CUDA tensors of the same shape")
        F, dev, E = fx.numel(), fx.device, self.f3_E
        if fx.device.index != self.device:
            raise ValueError("fx/fy/fz must live on the engine's device")
        tof = (self._check_out(out_tof, (E, F), torch.float64, dev, "out_tof") if out_tof is not None
               else torch.empty((E, F), dtype=torch.float64, device=dev))
        kmin = torch.empty((E, F), dtype=torch.int32, device=dev) if with_kmin else None
        ab = torch.empty((E, F, 2), dtype=torch.float64, device=dev) if with_angles else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        _capi.check(self._lib.tofrt_fermat3d_dev(self._ctx, fx.data_ptr(), fy.data_ptr(), fz.data_ptr(), F,
                                                 tof.data_ptr(), kmin.data_ptr() if with_kmin else None,
                                                 ab.data_ptr() if with_angles else None, ctypes.c_void_p(stream)))
        res = tuple(x for x in (tof, kmin, ab) if x is not None)
        return res if len(res) > 1 else tof

    # ---- main_rt.py:337-441 -----------------------------------------------------
    def shoot_rays(self, r_outer, pipe_offset, with_idx=False, with_legs=False):
        if self.multi:
            raise ValueError("shoot_rays is the reference's single-shot call (main_rt.py:481): set up one shooter "
                             "(scalar x_shooter), or use tof_table(records=...) for full-matrix capture")
        recs = ["pipe_x", "pipe_z", "lens2_x", "lens2_z", "target_x"]
        if with_idx:
            recs.append("idx")
        if with_legs:
            recs.append("legs")
        t = self.tof_table([r_outer], [pipe_offset], records=recs)
        lx, lz = self.lens()
        res = {
            "lens_1_x": lx, "lens_1_z": lz,
            "pipe_x": t.records["pipe_x"][0], "pipe_z": t.records["pipe_z"][0],
            "lens_2_x": t.records["lens2_x"][0], "lens_2_z": t.records["lens2_z"][0],
            "target_x": t.records["target_x"][0], "target_z": self.z_f.copy(),
        }
        if with_idx:
            res["idx"] = t.records["idx"][0]
        if with_legs:
            res["legs"] = t.records["legs"][0]
        return res
