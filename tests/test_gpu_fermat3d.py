"""GPU: the 3-D two-point solve (extension: SURVEY row f4 / BASELINE config 5 -- matrix array, 3-D
volume, curved surface) against the oracle's bit-exact twin, the NumPy/SciPy oracle, the 2-D solver
in the plane y = 0 and analytic known answers.  The reference is 2-D: parity UNPINNED."""
import numpy as np
import pytest

from ray_tracing_ultrasound_b200 import reference_setup as rs
from ray_tracing_ultrasound_b200._capi import TofrtError
from fermat_oracle import Lens
import fermat3d_oracle as f3

pytestmark = pytest.mark.gpu
D = float(rs.D)


def _setup(engine, n_side=17, E=None, seed=0, n_alpha=905):
    x_a, z_a = rs.transducer()
    alpha, zf = rs.launch_angles(n_alpha), rs.focus_line(n_alpha)
    engine.setup(alpha, zf, x_a[32], z_a[32], x_a)
    if E is None:   # 32 x 32 matrix array, pitch 0.6 mm (the 2-D array's pitch, main_rt.py:460)
        g = (np.arange(32) - 15.5) * float(rs.PITCH)
        ex, ey = (v.ravel() for v in np.meshgrid(g, g, indexing="ij"))
    else:
        rng = np.random.default_rng(seed)
        ex, ey = rng.uniform(-0.0095, 0.0095, E), rng.uniform(-0.0095, 0.0095, E)
    engine.setup_fermat3d(ex, ey, n_side=n_side)
    return ex, ey


def _voxels(rng, F, lat=0.03):
    return rng.uniform(-lat, lat, F), rng.uniform(-lat, lat, F), rng.uniform(0.0, 0.05, F)


@pytest.mark.parametrize("n_side", [9, 17, 33])
def test_fermat3d_bit_exact_vs_twin(engine, oracle_dm, n_side):
    ex, ey = _setup(engine, n_side, E=24, seed=n_side)
    fx, fy, fz = _voxels(np.random.default_rng(5), 3000)
    tof, kmin, ab = engine.fermat3d_tof(fx, fy, fz, with_angles=True)
    p = oracle_dm.params()
    o_t, o_k, o_ab, evals, gap = oracle_dm.fermat3d(p, engine.f3_ang, ex, ey, D, fx, fy, fz)
    same = kmin == o_k
    # the FP32 search may pick the other sample of a near-tie only (float error of a ~1e-4 s time: ~1e-11 s)
    assert same.mean() > 0.98 and gap[~same].max(initial=0.0) < 2e-10
    assert np.array_equal(tof[same], o_t[same]) and np.array_equal(ab[same], o_ab[same])
    for e, f in np.argwhere(~same)[:50]:
        t, a, b = oracle_dm.fermat3d_from(p, engine.f3_ang, (ex[e], ey[e], D), (fx[f], fy[f], fz[f]), kmin[e, f])
        assert t == tof[e, f] and a == ab[e, f, 0] and b == ab[e, f, 1]
        assert abs(t - o_t[e, f]) < 1e-17    # either start refines to the same minimum
    assert evals.max() <= 8


def test_fermat3d_matches_scipy_oracle(engine):
    ex, ey = _setup(engine, 17, E=6, seed=3)
    fx, fy, fz = _voxels(np.random.default_rng(6), 30)
    tof, kmin, ab = engine.fermat3d_tof(fx, fy, fz, with_angles=True)
    lens = Lens()
    o_t, o_k, o_a, o_b, gap = f3.fermat3d_table(lens, engine.f3_ang, ex, ey, D, fx, fy, fz)
    assert np.abs(tof - o_t).max() < 1e-15 and (tof - o_t).max() < 1e-18
    assert np.array_equal(kmin[gap > 1e-9], o_k[gap > 1e-9])
    assert np.abs(ab[..., 0] - o_a).max() < 1e-6 and np.abs(ab[..., 1] - o_b).max() < 1e-6


def test_fermat3d_plane_y0_is_the_2d_solver_bit_for_bit(engine, oracle_dm):
    """Consistency pin of SURVEY 8(d) config 5: restricted to the plane y = 0 the 3-D engine reproduces the 2-D engine
    bit for bit -- pairs with element and voxel at y = 0 ARE the 2-D problem and run the 2-D solver's arithmetic
    (kmin = -1 - j names the 2-D lens sample); pairs off the plane are untouched; the CPU twin follows the same rule;
    and the genuinely 3-D solve of the same in-plane pairs (twin without the rule) agrees to 1e-17 s."""
    x_a, z_a = rs.transducer()
    alpha, zf = rs.launch_angles(), rs.focus_line()
    engine.setup(alpha, zf, x_a[32], z_a[32], x_a)
    rng = np.random.default_rng(8)
    fx, fz = rng.uniform(-0.02, 0.02, 200), rng.uniform(0.0, 0.04, 200)
    t2, k2, al2 = engine.fermat_tof(fx, fz, with_alpha=True)
    engine.setup_fermat3d(x_a, np.zeros_like(x_a), n_side=17)
    t3, k3, ab = engine.fermat3d_tof(fx, np.zeros_like(fx), fz, with_angles=True)
    assert np.array_equal(t3, t2) and np.array_equal(k3, -1 - k2)
    assert np.array_equal(ab[..., 0], al2) and not ab[..., 1].any()
    # a matrix array with one row on y = 0, voxels on and off the plane: mixed table against the twin
    ex = np.concatenate([x_a[::8], x_a[::8]]); ey = np.concatenate([np.zeros(len(x_a[::8])), np.full(len(x_a[::8]), 6e-4)])
    gy = np.where(np.arange(200) % 2 == 0, 0.0, rng.uniform(-0.02, 0.02, 200))
    engine.setup_fermat3d(ex, ey, n_side=17)
    t, k, ab = engine.fermat3d_tof(fx, gy, fz, with_angles=True)
    p = oracle_dm.params()
    o_t, o_k, o_ab, _, gap = oracle_dm.fermat3d(p, engine.f3_ang, ex, ey, D, fx, gy, fz, alpha2d=alpha)
    inplane = (ey[:, None] == 0.0) & (gy[None, :] == 0.0)
    assert inplane.sum() == len(x_a[::8]) * 100 and (k[inplane] < 0).all() and (k[~inplane] >= 0).all()
    same = k == o_k
    assert same[inplane].all() and same.mean() > 0.98
    assert np.array_equal(t[same], o_t[same]) and np.array_equal(ab[same], o_ab[same])
    assert np.array_equal(t[: len(x_a[::8]), ::2], t2[::8, ::2])              # the y = 0 row is the linear-array table
    free3d = oracle_dm.fermat3d(p, engine.f3_ang, ex, ey, D, fx, gy, fz)[0]    # no rule: 3-D Newton everywhere
    assert np.abs(free3d - t)[inplane].max() < 1e-17
    # a different aperture for the 3-D search: no dispatch (the two solvers would not search the same set)
    engine.setup_fermat3d(x_a, np.zeros_like(x_a), ang=np.linspace(-0.5, 0.5, 17))
    assert (engine.fermat3d_tof(fx, np.zeros_like(fx), fz)[1] >= 0).all()


def test_fermat3d_known_answers(engine, oracle_dm):
    ex, ey = _setup(engine, 17)     # 32 x 32 matrix array
    assert len(ex) == 1024
    lens = Lens()
    # (i) constant-TOF lens (main_rt.py:181-187): array centre -> any entry point -> origin takes T
    engine.setup_fermat3d([0.0], [0.0], n_side=17)
    tof = engine.fermat3d_tof([0.0], [0.0], [0.0], with_kmin=False)
    assert abs(tof[0, 0] - (lens.l0 / lens.c1 + lens.h0 / lens.c2)) < 1e-18
    # (ii) rotation about the lens axis leaves the time unchanged (entry points inside the square
    #      (a, b) aperture in both orientations: the aperture itself is not rotation invariant)
    rng = np.random.default_rng(9)
    ex, ey = rng.uniform(-0.006, 0.006, 40), rng.uniform(-0.006, 0.006, 40)
    fx, fy, fz = _voxels(rng, 300, lat=0.015)
    engine.setup_fermat3d(ex, ey, n_side=17)
    t0, _, ab0 = engine.fermat3d_tof(fx, fy, fz, with_angles=True)
    inside = lambda ab: (np.abs(ab) < rs.ALPHA_MAX - 1e-6).all(axis=-1)
    for th in (0.3, np.pi / 2, 2.2):
        c, s = np.cos(th), np.sin(th)
        engine.setup_fermat3d(c * ex - s * ey, s * ex + c * ey, n_side=17)
        t1, _, ab1 = engine.fermat3d_tof(c * fx - s * fy, s * fx + c * fy, fz, with_angles=True)
        ok = inside(ab0) & inside(ab1)
        assert ok.mean() > 0.8 and np.abs(t1 - t0)[ok].max() < 1e-17
    # (iii) Snell's law in 3-D at the refined entry point: the tangential part of
    #       (P - A)/(c1 |P - A|) + (P - F)/(c2 |P - F|) vanishes
    engine.setup_fermat3d(ex, ey, n_side=17)
    tof, kmin, ab = engine.fermat3d_tof(fx, fy, fz, with_angles=True)
    eps = 1e-6
    for e in (0, 7, 39):
        a, b = ab[e, :, 0], ab[e, :, 1]
        P = np.array(f3.surface_point(lens, a, b))
        Ta = (np.array(f3.surface_point(lens, a + eps, b)) - np.array(f3.surface_point(lens, a - eps, b))) / (2 * eps)
        Tb = (np.array(f3.surface_point(lens, a, b + eps)) - np.array(f3.surface_point(lens, a, b - eps))) / (2 * eps)
        u1 = P - np.array([ex[e], ey[e], D])[:, None]
        u2 = P - np.array([fx, fy, fz])
        v = u1 / (lens.c1 * np.linalg.norm(u1, axis=0)) + u2 / (lens.c2 * np.linalg.norm(u2, axis=0))
        interior = (np.abs(a) < rs.ALPHA_MAX - 1e-9) & (np.abs(b) < rs.ALPHA_MAX - 1e-9)
        assert interior.mean() > 0.8
        for T in (Ta, Tb):   # finite-difference tangents: ~1e-10 relative
            assert np.abs((v * T).sum(axis=0) / np.linalg.norm(T, axis=0))[interior].max() < 2e-9 / lens.c2
        # the same statement with the analytic gradient (twin): dt/da = dt/db = 0 at the entry point
        for f in np.flatnonzero(interior)[:25]:
            _, g, H = oracle_dm.fermat3d_eval(oracle_dm.params(), engine.f3_ang, a[f], b[f],
                                              (ex[e], ey[e], D), (fx[f], fy[f], fz[f]))
            assert np.abs(g).max() < 1e-15      # s/rad; |H| ~ 3e-5 s/rad^2, so the angles are good to ~1e-10 rad
    # (iv) mirror symmetry y -> -y
    engine.setup_fermat3d(ex, -ey, n_side=17)
    assert np.abs(engine.fermat3d_tof(fx, -fy, fz, with_kmin=False) - tof).max() < 1e-17


def test_fermat3d_device_api_chunks_and_errors(engine):
    import torch
    ex, ey = _setup(engine, 9, E=16, seed=1, n_alpha=300)
    rng = np.random.default_rng(10)
    F = 200_037                    # not a multiple of the 256-voxel tile; several pipelined host chunks
    fx, fy, fz = _voxels(rng, F)
    t_h, k_h, ab_h = engine.fermat3d_tof(fx, fy, fz, with_angles=True)
    g = [torch.from_numpy(v).cuda() for v in (fx, fy, fz)]
    t_d, k_d, ab_d = engine.fermat3d_tof_cuda(*g, with_kmin=True, with_angles=True)
    torch.cuda.synchronize()
    assert np.array_equal(t_d.cpu().numpy(), t_h) and np.array_equal(k_d.cpu().numpy(), k_h)
    assert np.array_equal(ab_d.cpu().numpy(), ab_h)
    assert np.isfinite(t_h).all() and (k_h >= 0).all() and (k_h < 81).all()
    t_only = engine.fermat3d_tof(fx[:1000], fy[:1000], fz[:1000], with_kmin=False)
    assert np.array_equal(t_only, t_h[:, :1000])
    assert engine.fermat3d_tof([], [], [], with_kmin=False).shape == (16, 0)
    with pytest.raises(ValueError):
        engine.setup_fermat3d(ex, ey, n_side=1)
    with pytest.raises(ValueError):
        engine.setup_fermat3d(ex, ey[:-1])
    # a new tofrt_setup invalidates the 3-D tables (the lens may have changed)
    x_a, z_a = rs.transducer()
    engine.setup(rs.launch_angles(), rs.focus_line(), x_a[32], z_a[32], x_a)
    with pytest.raises(TofrtError):
        engine.fermat3d_tof(fx[:10], fy[:10], fz[:10])


def test_fermat3d_flat_interface(engine, oracle_dm):
    """Flat-interface option in 3-D (plane z = h0): vertical path known answer and the twin."""
    engine.set_surface("flat")
    oracle_dm.set_surface(True)
    try:
        ex, ey = _setup(engine, 17, E=20, seed=4)
        rng = np.random.default_rng(11)
        fz = rng.uniform(0.0, 0.05, 20)
        tof, kmin, ab = engine.fermat3d_tof(ex, ey, fz, with_angles=True)    # voxel e straight below element e
        lens = Lens(flat=True)
        straight = (D - lens.h0) / lens.c1 + (lens.h0 - fz) / lens.c2
        assert np.abs(np.diag(tof) - straight).max() < 1e-18
        fx, fy, fz = _voxels(rng, 500)
        tof, kmin, ab = engine.fermat3d_tof(fx, fy, fz, with_angles=True)
        o_t, o_k, o_ab, evals, gap = oracle_dm.fermat3d(oracle_dm.params(), engine.f3_ang, ex, ey, D, fx, fy, fz)
        same = kmin == o_k
        assert same.mean() > 0.98 and np.array_equal(tof[same], o_t[same])
        assert np.abs(tof - o_t).max() < 1e-17
    finally:
        engine.set_surface("lens")
        oracle_dm.set_surface(False)
