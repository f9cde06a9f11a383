"""NumPy/SciPy oracle of the 3-D two-point Fermat solve (test infrastructure; SURVEY.md row f4,
BASELINE config 5).

The 3-D lens is the surface of revolution of the reference's lens profile about the z axis:
a point of the surface in direction n (unit vector, polar angle alpha from +z) is
P = h(alpha) n with the reference's h_from_alpha (main_rt.py:180-189), restated here through
tests/fermat_oracle.Lens.  Directions are parametrised by two angles that are regular on the
whole lens cap (the poles sit on the y axis, 90 degrees away from the lens):

    n(a, b) = (sin a cos b, sin b, cos a cos b),     cos(alpha) = cos a cos b

so that b = 0 is the reference's 2-D x-z plane with a = alpha.  For an element A = (ax, ay, d)
of a matrix array on the plane z = d and a focal point F = (fx, fy, fz) in the water, the oracle
minimises t(a, b) = |A - P|/c1 + |P - F|/c2: discrete argmin over an n_side x n_side grid of
(a, b) samples over [-alpha_max, alpha_max]^2, then a bounded quasi-Newton refinement inside the
neighbouring cells (SciPy L-BFGS-B / TNC polish with numerical gradients of this file's own t --
deliberately independent of the analytic derivatives the CUDA kernel uses).

The reference has no 3-D code at all: parity is UNPINNED; the pins are analytic known answers
(tests/test_gpu_fermat3d.py) and the consistency with the 2-D solver in the plane y = 0."""
import numpy as np
from scipy.optimize import minimize

from fermat_oracle import Lens


def direction(a, b):
    return np.sin(a) * np.cos(b), np.sin(b), np.cos(a) * np.cos(b)


def surface_point(lens, a, b):
    nx, ny, nz = direction(a, b)
    h = lens.h(np.arccos(np.clip(nz, -1.0, 1.0)))
    return h * nx, h * ny, h * nz


def tof3d(lens, a, b, A, F):
    px, py, pz = surface_point(lens, a, b)
    r1 = np.sqrt((px - A[0]) ** 2 + (py - A[1]) ** 2 + (pz - A[2]) ** 2)
    r2 = np.sqrt((px - F[0]) ** 2 + (py - F[1]) ** 2 + (pz - F[2]) ** 2)
    return r1 / lens.c1 + r2 / lens.c2


def grid_angles(alpha_lo, alpha_hi, n_side):
    return np.linspace(alpha_lo, alpha_hi, n_side)


def solve_pair(lens, ang, A, F):
    """-> (tof, kmin, a*, b*, gap): kmin = row-major index (ia * n + ib) of the best grid sample,
    gap = runner-up minus minimum over the grid."""
    n = len(ang)
    aa, bb = np.meshgrid(ang, ang, indexing="ij")
    t = tof3d(lens, aa, bb, A, F)
    k = int(np.argmin(t))
    srt = np.partition(t.ravel(), 1)[:2]
    ia, ib = divmod(k, n)
    lo = (ang[max(ia - 1, 0)], ang[max(ib - 1, 0)])
    hi = (ang[min(ia + 1, n - 1)], ang[min(ib + 1, n - 1)])
    scale = 1e4  # seconds -> ~1: keeps the optimiser's absolute tolerances meaningful
    fun = lambda v: scale * tof3d(lens, v[0], v[1], A, F)
    best = (t.ravel()[k], ang[ia], ang[ib])
    x0 = np.array([ang[ia], ang[ib]])
    for method, opts in (("L-BFGS-B", dict(ftol=1e-30, gtol=1e-14, maxiter=200)),
                         ("TNC", dict(ftol=1e-30, gtol=1e-14, xtol=1e-14, maxfun=500))):
        res = minimize(fun, x0, method=method, bounds=list(zip(lo, hi)), options=opts)
        tv = tof3d(lens, res.x[0], res.x[1], A, F)
        if tv < best[0]:
            best = (tv, res.x[0], res.x[1])
        x0 = np.array(best[1:])
    return best[0], k, best[1], best[2], srt[1] - srt[0]


def fermat3d_table(lens, ang, ex, ey, z_el, fx, fy, fz):
    E, Fn = len(ex), len(fx)
    tof, kmin = np.empty((E, Fn)), np.empty((E, Fn), dtype=np.int32)
    a_s, b_s, gap = np.empty((E, Fn)), np.empty((E, Fn)), np.empty((E, Fn))
    for e in range(E):
        for f in range(Fn):
            tof[e, f], kmin[e, f], a_s[e, f], b_s[e, f], gap[e, f] = solve_pair(
                lens, ang, (ex[e], ey[e], z_el), (fx[f], fy[f], fz[f]))
    return tof, kmin, a_s, b_s, gap
