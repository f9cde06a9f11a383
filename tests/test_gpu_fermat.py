"""GPU: the two-point Fermat solve (extension) against its NumPy/SciPy oracle and
analytic known answers.  Parity is UNPINNED by the reference (it has no such function)."""
import numpy as np
import pytest

from ray_tracing_ultrasound_b200 import reference_setup as rs
from fermat_oracle import Lens, fermat_table

pytestmark = pytest.mark.gpu


def _setup(engine, n_alpha=905):
    x_a, z_a = rs.transducer()
    alpha, zf = rs.launch_angles(n_alpha), rs.focus_line(n_alpha)
    engine.setup(alpha, zf, x_a[32], z_a[32], x_a)
    return x_a, z_a, alpha


def test_fermat_matches_numpy_oracle(engine):
    x_a, z_a, alpha = _setup(engine)
    rng = np.random.default_rng(7)
    F = 60
    fx, fz = rng.uniform(-0.04, 0.04, F), rng.uniform(-0.01, 0.045, F)   # focal points in the water
    tof, kmin, al = engine.fermat_tof(fx, fz, with_alpha=True)
    assert tof.shape == (65, F) and kmin.dtype == np.int32
    lens = Lens()
    elems = np.arange(0, 65, 4)
    o_tof, o_k, o_al, gap = fermat_table(lens, alpha, x_a[elems], z_a[0], fx, fz)
    assert np.abs(tof[elems] - o_tof).max() < 1e-15          # seconds (times are ~1e-4 s)
    clear = gap > 1e-18                                      # discrete argmin: exact unless two samples tie
    assert np.array_equal(kmin[elems][clear], o_k[clear]) and clear.mean() > 0.99
    assert np.abs(al[elems] - o_al).max() < 1e-8             # t is stationary there: alpha is loosely determined
    # the refined time is never above the best sample's time
    px, pz = lens.point(alpha)
    for i, e in enumerate(elems):
        t_samples = np.hypot(px - x_a[e], pz - z_a[0])[:, None] / lens.c1 + \
            np.hypot(px[:, None] - fx[None, :], pz[:, None] - fz[None, :]) / lens.c2
        assert (tof[e] <= t_samples.min(axis=0) + 1e-18).all()


def test_fermat_bit_exact_vs_twin(engine, oracle_dm):
    """Every (element, focal point) pair against the oracle's CPU twin of the kernel (tofo_fermat2d: the same
    operations compiled by gcc): times, discrete minimisers and entry angles identical -- the kernel's FP32-screened
    argmin is exactly the plain FP64 first minimum."""
    x_a, z_a, alpha = _setup(engine)
    rng = np.random.default_rng(17)
    fx, fz = rng.uniform(-0.04, 0.04, 400), rng.uniform(-0.01, 0.045, 400)
    tof, kmin, al = engine.fermat_tof(fx, fz, with_alpha=True)
    o_t, o_k, o_al = oracle_dm.fermat2d(oracle_dm.params(), alpha, x_a, float(z_a[0]), fx, fz)
    assert np.array_equal(kmin, o_k) and np.array_equal(tof, o_t) and np.array_equal(al, o_al)


def test_fermat_known_answers(engine):
    x_a, z_a, alpha = _setup(engine)
    lens = Lens()
    # (i) the lens is the constant-TOF surface between the array centre and the origin
    #     (main_rt.py:181-187): every entry point gives T = l0/c1 + h0/c2.
    tof, kmin = engine.fermat_tof(np.array([0.0]), np.array([0.0]))
    T = lens.l0 / lens.c1 + lens.h0 / lens.c2
    assert abs(tof[32, 0] - T) < 1e-18
    # (ii) Snell's law at the refined entry point: sin(th1)/c1 == sin(th2)/c2 w.r.t. the surface normal
    rng = np.random.default_rng(3)
    fx, fz = rng.uniform(-0.03, 0.03, 40), rng.uniform(0.0, 0.04, 40)
    tof, kmin, al = engine.fermat_tof(fx, fz, with_alpha=True)
    for e in (0, 17, 32, 50, 64):
        px, pz = lens.point(al[e])
        dx, dz = lens.tangent(al[e])
        tx, tz = dx / np.hypot(dx, dz), dz / np.hypot(dx, dz)
        u1 = np.array([px - x_a[e], pz - z_a[0]]); u1 /= np.hypot(*u1)
        u2 = np.array([fx - px, fz - pz]); u2 /= np.hypot(*u2)
        s1, s2 = u1[0] * tx + u1[1] * tz, u2[0] * tx + u2[1] * tz   # tangential components = sines
        interior = (kmin[e] > 0) & (kmin[e] < len(alpha) - 1)
        assert np.abs(s1 / lens.c1 - s2 / lens.c2)[interior].max() < 1e-12
    # (iii) mirror symmetry x -> -x
    tof_m, _ = engine.fermat_tof(-fx, fz)
    assert np.abs(tof_m[::-1] - tof).max() < 1e-17


def test_fermat_device_api_and_sizes(engine):
    import torch
    x_a, z_a, alpha = _setup(engine, 300)
    gx, gz = np.meshgrid(np.linspace(-0.03, 0.03, 40), np.linspace(0.0, 0.04, 33))
    fx, fz = gx.ravel(), gz.ravel()            # 1 320 focal points: not a multiple of the 256-point tile
    t_h, k_h = engine.fermat_tof(fx, fz)
    t_d, k_d = engine.fermat_tof_cuda(torch.from_numpy(fx).cuda(), torch.from_numpy(fz).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(t_d.cpu().numpy(), t_h) and np.array_equal(k_d.cpu().numpy(), k_h)
    assert np.isfinite(t_h).all() and (k_h >= 0).all() and (k_h < 300).all()
    t0, k0 = engine.fermat_tof([], [])
    assert t0.shape == (65, 0)
