"""NumPy/SciPy oracle of the two-point Fermat solve (test infrastructure).

Built from the reference's own surface primitives, restated in NumPy in the reference's
operation order: h_from_alpha (main_rt.py:180-189), dh_from_alpha (:192-214),
x_z_from_alpha (:217-223), dz_dx_from_alpha (:226-234).  No function of the reference computes
a two-point travel time, so this oracle is pinned only by analytic known answers
(tests/test_gpu_fermat.py): parity unpinned."""
import numpy as np
from scipy.optimize import brentq


class Lens:
    def __init__(self, c1=6400.0, c2=1483.0, l0=0.12156646438729327, h0=0.08843353561270673, flat=False):
        self.flat = flat  # extension: plane z = h0 instead of the reference's lens
        self.c1, self.c2 = np.float64(c1), np.float64(c2)
        self.l0, self.h0 = np.float64(l0), np.float64(h0)
        self.d = self.l0 + self.h0

    def h(self, alpha):
        if self.flat:
            return self.h0 / np.cos(alpha)
        c1, c2, d = self.c1, self.c2, self.d
        T = self.l0 / c1 + self.h0 / c2
        A = (np.square(c1) / np.square(c2)) - 1
        B = 2 * d * np.cos(alpha) - 2 * T * np.square(c1) / c2
        C = (np.square(c1) * np.square(T)) - np.square(d)
        return (-B - np.sqrt(np.square(B) - 4 * A * C)) / (2 * A)

    def dh(self, alpha):
        if self.flat:
            return self.h0 * np.sin(alpha) / (np.cos(alpha) * np.cos(alpha))
        c1, c2, d = self.c1, self.c2, self.d
        T = self.l0 / c1 + self.h0 / c2
        A = (np.square(c1) / np.square(c2)) - 1
        C = (np.square(c1) * np.square(T)) - np.square(d)
        phi_1, phi_2, phi_3 = -1 / (2 * A), (-2 * T * np.square(c1)) / c2, 2 * d
        B = phi_2 + phi_3 * np.cos(alpha)
        dB = -phi_3 * np.sin(alpha)
        return phi_1 * (dB + (1 / (2 * np.sqrt(np.square(B) - 4 * A * C))) * (2 * B * dB))

    def point(self, alpha):
        h = self.h(alpha)
        return h * np.sin(alpha), h * np.cos(alpha)

    def tangent(self, alpha):
        h, dh = self.h(alpha), self.dh(alpha)
        return dh * np.sin(alpha) + h * np.cos(alpha), dh * np.cos(alpha) - h * np.sin(alpha)  # dx, dz

    def tof(self, alpha, ax, az, fx, fz):
        px, pz = self.point(alpha)
        return np.hypot(px - ax, pz - az) / self.c1 + np.hypot(px - fx, pz - fz) / self.c2

    def dtof(self, alpha, ax, az, fx, fz):
        px, pz = self.point(alpha)
        dx, dz = self.tangent(alpha)
        r1, r2 = np.hypot(px - ax, pz - az), np.hypot(px - fx, pz - fz)
        return ((px - ax) * dx + (pz - az) * dz) / (self.c1 * r1) + ((px - fx) * dx + (pz - fz) * dz) / (self.c2 * r2)


def fermat_table(lens, alpha, x_el, z_el, fx, fz):
    """tof[E][F], kmin[E][F], alpha*[E][F] and the runner-up gap of the discrete minimum."""
    px, pz = lens.point(alpha)
    E, F, N = len(x_el), len(fx), len(alpha)
    tof, kmin = np.empty((E, F)), np.empty((E, F), dtype=np.int32)
    astar, gap = np.empty((E, F)), np.empty((E, F))
    for e in range(E):
        d1 = np.hypot(px - x_el[e], pz - z_el) / lens.c1
        for f in range(F):
            t = d1 + np.hypot(px - fx[f], pz - fz[f]) / lens.c2
            j = int(np.argmin(t))
            kmin[e, f] = j
            srt = np.partition(t, 1)[:2]
            gap[e, f] = srt[1] - srt[0]
            a, b = alpha[max(j - 1, 0)], alpha[min(j + 1, N - 1)]
            args = (x_el[e], z_el, fx[f], fz[f])
            if lens.dtof(a, *args) < 0 < lens.dtof(b, *args):
                al = brentq(lens.dtof, a, b, args=args, xtol=1e-16, rtol=1e-15, maxiter=200)
            else:
                al = alpha[j]
            astar[e, f] = al
            tof[e, f] = lens.tof(al, *args)
    return tof, kmin, astar, gap
