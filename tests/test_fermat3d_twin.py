"""CPU: the 3-D two-point solve's shared arithmetic (csrc/tofrt_fermat3d.h, compiled by gcc into
the oracle twin) against the independent NumPy/SciPy oracle and finite differences.  The reference
is 2-D: parity UNPINNED (BASELINE config 5)."""
import numpy as np

from ray_tracing_ultrasound_b200 import reference_setup as rs
from fermat_oracle import Lens
import fermat3d_oracle as f3


def _ang(n):
    return np.linspace(-rs.ALPHA_MAX, rs.ALPHA_MAX, n)


def test_gradient_and_hessian_match_finite_differences(oracle_dm):
    p, lens, ang = oracle_dm.params(), Lens(), _ang(17)
    rng = np.random.default_rng(0)
    for _ in range(20):
        A = np.array([rng.uniform(-0.01, 0.01), rng.uniform(-0.01, 0.01), float(rs.D)])
        F = np.array([rng.uniform(-0.03, 0.03), rng.uniform(-0.03, 0.03), rng.uniform(0.0, 0.05)])
        a, b = rng.uniform(-0.8, 0.8, 2)
        t, g, H = oracle_dm.fermat3d_eval(p, ang, a, b, A, F)
        assert abs(t - f3.tof3d(lens, a, b, A, F)) < 1e-19
        eps = 1e-6
        num_g = [(f3.tof3d(lens, a + eps, b, A, F) - f3.tof3d(lens, a - eps, b, A, F)) / (2 * eps),
                 (f3.tof3d(lens, a, b + eps, A, F) - f3.tof3d(lens, a, b - eps, A, F)) / (2 * eps)]
        assert np.allclose(g, num_g, rtol=0, atol=1e-13)
        ga = lambda aa, bb: oracle_dm.fermat3d_eval(p, ang, aa, bb, A, F)[1]
        da, db = (ga(a + eps, b) - ga(a - eps, b)) / (2 * eps), (ga(a, b + eps) - ga(a, b - eps)) / (2 * eps)
        assert np.allclose(H, [da[0], 0.5 * (da[1] + db[0]), db[1]], rtol=1e-6, atol=1e-12)


def test_twin_matches_scipy_oracle(oracle_dm):
    p, lens = oracle_dm.params(), Lens()
    rng = np.random.default_rng(1)
    E, Fn = 5, 24
    ex, ey = rng.uniform(-0.0095, 0.0095, E), rng.uniform(-0.0095, 0.0095, E)
    fx, fy, fz = rng.uniform(-0.03, 0.03, Fn), rng.uniform(-0.03, 0.03, Fn), rng.uniform(0.0, 0.05, Fn)
    for n in (9, 17):
        ang = _ang(n)
        tof, kmin, ab, evals, gap = oracle_dm.fermat3d(p, ang, ex, ey, float(rs.D), fx, fy, fz)
        o_t, o_k, o_a, o_b, o_gap = f3.fermat3d_table(lens, ang, ex, ey, float(rs.D), fx, fy, fz)
        assert np.array_equal(kmin[gap > 1e-16], o_k[gap > 1e-16])
        assert (tof - o_t).max() < 1e-18          # never worse than SciPy's minimum ...
        assert np.abs(tof - o_t).max() < 1e-15    # ... and the same minimum (times are ~1e-4 s)
        assert evals.max() <= 8
        # the refined time is never above the best grid sample's
        aa, bb = np.meshgrid(ang, ang, indexing="ij")
        for e in range(E):
            for f in range(0, Fn, 5):
                tg = f3.tof3d(lens, aa, bb, (ex[e], ey[e], float(rs.D)), (fx[f], fy[f], fz[f]))
                assert tof[e, f] <= tg.min() + 1e-19


def test_twin_plane_y0_is_the_2d_problem(oracle_dm):
    """b = 0 is the reference's x-z plane: the 3-D solve of in-plane pairs equals a 1-D minimisation of the
    2-D oracle's travel time over alpha (fermat_oracle.Lens.tof, built from main_rt.py:180-223)."""
    from scipy.optimize import minimize_scalar
    p, lens, ang = oracle_dm.params(), Lens(), _ang(17)
    rng = np.random.default_rng(2)
    ex, fx, fz = rng.uniform(-0.0095, 0.0095, 6), rng.uniform(-0.02, 0.02, 10), rng.uniform(0.0, 0.04, 10)
    tof, kmin, ab, _, _ = oracle_dm.fermat3d(p, ang, ex, np.zeros(6), float(rs.D), fx, np.zeros(10), fz)
    assert np.abs(ab[..., 1]).max() < 1e-12
    for e in range(6):
        for f in range(10):
            r = minimize_scalar(lambda al: lens.tof(al, ex[e], float(rs.D), fx[f], fz[f]) * 1e4,
                                bounds=(-rs.ALPHA_MAX, rs.ALPHA_MAX), method="bounded", options=dict(xatol=1e-12))
            assert abs(tof[e, f] - r.fun * 1e-4) < 1e-17


def test_twin_2d_solver_matches_numpy_oracle_and_the_in_plane_rule(oracle_dm):
    """The 2-D two-point twin (tofo_fermat2d) against the NumPy/SciPy oracle built from the reference's primitives, and
    the in-plane rule of the 3-D twin: with the 2-D lens angles given, pairs at y = 0 return the 2-D solver's bits."""
    from fermat_oracle import fermat_table
    p, lens = oracle_dm.params(), Lens()
    alpha = rs.launch_angles()
    x_a, z_a = rs.transducer()
    rng = np.random.default_rng(4)
    fx, fz = rng.uniform(-0.04, 0.04, 25), rng.uniform(-0.01, 0.045, 25)
    xe = x_a[::8]
    tof, kmin, al = oracle_dm.fermat2d(p, alpha, xe, float(rs.D), fx, fz)
    o_tof, o_k, o_al, gap = fermat_table(lens, alpha, xe, float(rs.D), fx, fz)
    assert np.abs(tof - o_tof).max() < 1e-15 and np.abs(al - o_al).max() < 1e-8
    assert np.array_equal(kmin[gap > 1e-18], o_k[gap > 1e-18])
    ang = np.linspace(alpha[0], alpha[-1], 17)
    ey = np.zeros(len(xe)); ey[1::2] = 3e-4
    t3, k3, ab, ev, _ = oracle_dm.fermat3d(p, ang, xe, ey, float(rs.D), fx, np.zeros(25), fz, alpha2d=alpha)
    on = ey == 0.0
    assert np.array_equal(t3[on], tof[on]) and np.array_equal(k3[on], -1 - kmin[on]) and np.array_equal(ab[on][..., 0], al[on])
    assert (k3[~on] >= 0).all() and (ev[~on] > 0).all()
    free = oracle_dm.fermat3d(p, ang, xe, ey, float(rs.D), fx, np.zeros(25), fz)[0]
    assert np.abs(free - t3).max() < 1e-17 and np.array_equal(free[~on], t3[~on])


def test_twin_flat_interface_known_answer(oracle_dm):
    """Flat interface z = h0 in 3-D: for a voxel straight below an element the entry point is straight above it and
    the time is (d - h0)/c1 + (h0 - z)/c2; off-axis pairs obey Snell's law (gradient zero at the entry point)."""
    p = oracle_dm.params()
    ang = _ang(17)
    rng = np.random.default_rng(4)
    ex, ey = rng.uniform(-0.009, 0.009, 8), rng.uniform(-0.009, 0.009, 8)
    fz = rng.uniform(0.0, 0.05, 8)
    try:
        oracle_dm.set_surface(True)
        tof, kmin, ab, evals, gap = oracle_dm.fermat3d(p, ang, ex, ey, float(rs.D), ex, ey, fz)
        straight = (float(rs.D) - float(rs.H0)) / float(rs.C1) + (float(rs.H0) - fz) / float(rs.C2)
        assert np.abs(np.diag(tof) - straight).max() < 1e-18
        for e in range(8):
            for f in range(8):
                _, g, H = oracle_dm.fermat3d_eval(p, ang, ab[e, f, 0], ab[e, f, 1], (ex[e], ey[e], float(rs.D)),
                                                  (ex[f], ey[f], fz[f]))
                assert np.abs(g).max() < 1e-15 and H[0] > 0 and H[0] * H[2] - H[1] ** 2 > 0
    finally:
        oracle_dm.set_surface(False)
