"""ctypes front-end of the CPU oracle (oracle/tofrt_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package never
imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int_p = ctypes.POINTER(ctypes.c_int)


class Params(ctypes.Structure):
    """The module globals the reference's functions read (main_rt.py:449-455)."""
    _fields_ = [("c1", ctypes.c_double), ("c2", ctypes.c_double), ("l0", ctypes.c_double),
                ("h0", ctypes.c_double), ("d", ctypes.c_double)]


def build(force=False):
    """Compile liboracle.so / liboracle_dm.so with oracle/Makefile."""
    args = ["make", "-C", _HERE] + (["-B"] if force else [])
    subprocess.run(args, check=True, capture_output=True)


def _dp(a):
    return a.ctypes.data_as(_c_double_p)


def _ip(a):
    return a.ctypes.data_as(_c_int_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Oracle:
    """One loaded oracle library.  device_math=True loads the bit-exact twin."""

    def __init__(self, device_math=False):
        name = "liboracle_dm.so" if device_math else "liboracle.so"
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        self.lib = ctypes.CDLL(path)
        self.device_math = bool(self.lib.tofo_uses_device_math())
        assert self.device_math == device_math
        self.lib.tofo_max_threads.restype = ctypes.c_int

    @staticmethod
    def params(c1=6400.0, c2=1483.0, l0=0.12156646438729327, h0=0.08843353561270673):
        l0, h0 = np.float64(l0), np.float64(h0)
        return Params(float(c1), float(c2), float(l0), float(h0), float(l0 + h0))

    def set_surface(self, flat):
        """Extension: flat interface z = h0 instead of the reference's lens (process-wide switch)."""
        self.lib.tofo_set_surface(ctypes.c_int(1 if flat else 0))

    def set_math(self, vector):
        """Twin build: the product's vector form of the return-path interface laws (process-wide switch)."""
        self.lib.tofo_set_math(ctypes.c_int(1 if vector else 0))

    def max_threads(self):
        return int(self.lib.tofo_max_threads())

    def lens(self, p, alpha):
        alpha = _f64(alpha)
        x, z = np.empty_like(alpha), np.empty_like(alpha)
        rc = self.lib.tofo_lens(ctypes.byref(p), _dp(alpha), ctypes.c_int(len(alpha)), _dp(x), _dp(z))
        assert rc == 0
        return x, z

    def shoot_rays(self, p, r_outer, pipe_offset, x_a, z_a, z_f, alpha):
        """main_rt.py:337 shoot_rays -> dict of the eight arrays + idx, a_l, b_l."""
        alpha, z_f = _f64(alpha), _f64(z_f)
        n = len(alpha)
        if n < 2 or len(z_f) != n:
            raise ValueError("Curve needs at least two points.")
        out = {k: np.empty(n) for k in ("lens_1_x", "lens_1_z", "pipe_x", "pipe_z", "lens_2_x",
                                        "lens_2_z", "target_x", "target_z", "a_l", "b_l")}
        idx = np.empty(n, dtype=np.int32)
        margin = np.empty(n)
        rc = self.lib.tofo_shoot_rays(
            ctypes.byref(p), ctypes.c_double(r_outer), ctypes.c_double(pipe_offset),
            ctypes.c_double(x_a), ctypes.c_double(z_a), _dp(z_f), _dp(alpha), ctypes.c_int(n),
            _dp(out["lens_1_x"]), _dp(out["lens_1_z"]), ctypes.c_int(0),
            _dp(out["pipe_x"]), _dp(out["pipe_z"]), _dp(out["lens_2_x"]), _dp(out["lens_2_z"]),
            _dp(out["target_x"]), _dp(out["target_z"]), _ip(idx), _dp(out["a_l"]), _dp(out["b_l"]), _dp(margin))
        assert rc == 0, rc
        out["idx"] = idx
        out["margin"] = margin
        return out

    def ray_legs(self, p, x_a, z_a, res):
        n = len(res["target_x"])
        legs = np.empty((4, n))
        rc = self.lib.tofo_ray_legs(
            ctypes.byref(p), ctypes.c_double(x_a), ctypes.c_double(z_a), ctypes.c_int(n),
            _dp(_f64(res["lens_1_x"])), _dp(_f64(res["lens_1_z"])), _dp(_f64(res["pipe_x"])),
            _dp(_f64(res["pipe_z"])), _dp(_f64(res["lens_2_x"])), _dp(_f64(res["lens_2_z"])),
            _dp(_f64(res["target_x"])), _dp(_f64(res["target_z"])), _dp(legs))
        assert rc == 0
        return legs

    def table(self, p, radius, offset, x_a, z_a, z_f, alpha, x_rx, rtol=1e-5, atol=1e-6,
              want_idx=False, nthreads=0, want_margin=False):
        """main_rt.py:464-501 -> (tof[E][G], k_hit[E][G], idx[G][n] or None[, margin[G][n]]).
        margin: how far each ray's return-path decision is from rounding noise (min |diff| at the first sign
        change, or over all samples when there is none; NaN for NaN lines)."""
        radius, offset, z_f, alpha, x_rx = map(_f64, (radius, offset, z_f, alpha, x_rx))
        G, n, E = len(radius), len(alpha), len(x_rx)
        assert len(offset) == G and len(z_f) == n
        tof = np.empty((E, G))
        k_hit = np.empty((E, G), dtype=np.int32)
        idx = np.empty((G, n), dtype=np.int32) if want_idx else None
        margin = np.empty((G, n)) if want_margin else None
        rc = self.lib.tofo_table(
            ctypes.byref(p), _dp(radius), _dp(offset), ctypes.c_int(G), ctypes.c_double(x_a),
            ctypes.c_double(z_a), _dp(z_f), _dp(alpha), ctypes.c_int(n), _dp(x_rx), ctypes.c_int(E),
            ctypes.c_double(rtol), ctypes.c_double(atol), _dp(tof), _ip(k_hit), _ip(idx),
            _dp(margin) if want_margin else None, ctypes.c_int(nthreads))
        assert rc == 0, rc
        return (tof, k_hit, idx, margin) if want_margin else (tof, k_hit, idx)

    # ---- 3-D two-point Fermat solve: CPU twin of the CUDA kernel (device-math build only) ----
    def fermat2d(self, p, alpha, x_el, z_el, fx, fz):
        """2-D two-point solve, CPU twin of the CUDA kernel -> tof[E][F], kmin[E][F], alpha[E][F]."""
        assert self.device_math, "the 2-D two-point twin exists in the device-math build only"
        alpha, x_el, fx, fz = map(_f64, (alpha, x_el, fx, fz))
        E, F = len(x_el), len(fx)
        tof, kmin, al = np.empty((E, F)), np.empty((E, F), dtype=np.int32), np.empty((E, F))
        rc = self.lib.tofo_fermat2d(ctypes.byref(p), _dp(alpha), ctypes.c_int(len(alpha)), _dp(x_el), ctypes.c_int(E),
                                    ctypes.c_double(z_el), _dp(fx), _dp(fz), ctypes.c_int(F), _dp(tof), _ip(kmin), _dp(al))
        assert rc == 0, rc
        return tof, kmin, al

    def fermat3d(self, p, ang, ex, ey, z_el, fx, fy, fz, alpha2d=None):
        """-> tof[E][F], kmin[E][F], ab[E][F][2], evals[E][F], gap[E][F].  alpha2d: the 2-D setup's lens angles --
        pairs in the plane y = 0 are then solved by the 2-D solver, as the product does (kmin = -1 - j)."""
        assert self.device_math, "the 3-D twin exists in the device-math build only"
        ang, ex, ey, fx, fy, fz = map(_f64, (ang, ex, ey, fx, fy, fz))
        E, F = len(ex), len(fx)
        tof, kmin = np.empty((E, F)), np.empty((E, F), dtype=np.int32)
        ab, evals, gap = np.empty((E, F, 2)), np.empty((E, F), dtype=np.int32), np.empty((E, F))
        rc = self.lib.tofo_fermat3d(ctypes.byref(p), _dp(ang), ctypes.c_int(len(ang)), _dp(ex), _dp(ey),
                                    ctypes.c_int(E), ctypes.c_double(z_el), _dp(fx), _dp(fy), _dp(fz),
                                    ctypes.c_int(F), _dp(tof), _ip(kmin), _dp(ab), _ip(evals), _dp(gap),
                                    _dp(_f64(alpha2d)) if alpha2d is not None else None,
                                    ctypes.c_int(0 if alpha2d is None else len(alpha2d)))
        assert rc == 0, rc
        return tof, kmin, ab, evals, gap

    def fermat3d_from(self, p, ang, A, F, k):
        """Refinement started from grid sample k -> (tof, a, b)."""
        ang = _f64(ang)
        t, ab = ctypes.c_double(), np.empty(2)
        rc = self.lib.tofo_fermat3d_from(ctypes.byref(p), _dp(ang), ctypes.c_int(len(ang)),
                                         *[ctypes.c_double(float(v)) for v in (*A, *F)], ctypes.c_int(int(k)),
                                         ctypes.byref(t), _dp(ab))
        assert rc == 0, rc
        return t.value, ab[0], ab[1]

    def fermat3d_eval(self, p, ang, a, b, A, F):
        """t, gradient (2), Hessian (aa, ab, bb) of the travel time at (a, b)."""
        ang, A, F = _f64(ang), _f64(A), _f64(F)
        t, g, H = ctypes.c_double(), np.empty(2), np.empty(3)
        rc = self.lib.tofo_fermat3d_eval(ctypes.byref(p), _dp(ang), ctypes.c_int(len(ang)), ctypes.c_double(a),
                                         ctypes.c_double(b), _dp(A), _dp(F), ctypes.byref(t), _dp(g), _dp(H))
        assert rc == 0
        return t.value, g, H
