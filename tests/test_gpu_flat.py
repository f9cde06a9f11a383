"""GPU: the flat-interface surface option (extension: BASELINE configs 1-2 name a "flat interface"
that the reference does not have -- parity UNPINNED).  Checked against the oracle twin compiled with
the same switch, analytic known answers, and (two-point solve) the NumPy/SciPy oracle."""
import numpy as np
import pytest

from ray_tracing_ultrasound_b200 import reference_setup as rs
from fermat_oracle import Lens, fermat_table

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("math_mode")]


@pytest.fixture()
def flat_engine(engine, oracle_dm):
    engine.set_surface("flat")
    oracle_dm.set_surface(True)
    yield engine
    engine.set_surface("lens")
    oracle_dm.set_surface(False)


def test_flat_interface_pulse_echo_matches_twin(flat_engine, oracle_dm):
    engine = flat_engine
    x_a, z_a = rs.transducer()
    alpha, zf = rs.launch_angles(), rs.focus_line()
    engine.setup(alpha, zf, x_a[32], z_a[32], x_a)
    assert engine.lens_shape == 0                      # a plane is neither strictly concave nor convex: literal scan
    lx, lz = engine.lens()
    assert np.abs(lz - rs.H0).max() < 1e-16 and np.all(np.diff(lx) > 0)
    rng = np.random.default_rng(4)
    radii, offsets = rng.uniform(0.01, 0.08, 200), rng.uniform(-0.01, 0.01, 200)
    t = engine.tof_table(radii, offsets, records=("idx", "target_x"))
    p = oracle_dm.params()
    tof, k_hit, idx = oracle_dm.table(p, radii, offsets, x_a[32], z_a[32], zf, alpha, x_a, want_idx=True)
    assert np.array_equal(t.records["idx"], idx)
    assert np.array_equal(t.k_hit, k_hit.T) and np.array_equal(t.tof, tof.T)
    assert t.hit.sum() > 100
    # normal-incidence echo on a centred pipe: 2 (l0/c1 + (h0 - r)/c2)
    for r in (0.01, 0.03, 0.06):
        tt = engine.tof_table([r], [0.0])
        assert tt.k_hit[0, 32] >= 0
        assert abs(tt.tof[0, 32] - 2 * (rs.L0 / rs.C1 + (rs.H0 - r) / rs.C2)) < 1e-17


def test_flat_interface_two_point_solve_water_steel(flat_engine):
    """The textbook immersion case: array in water above a flat steel surface (c1 < c2)."""
    engine = flat_engine
    kw = dict(c1=1483.0, c2=5600.0, l0=0.03, h0=0.05)   # 30 mm water path; interface at z = 50 mm; array at z = 80 mm
    d = float(np.float64(kw["l0"]) + np.float64(kw["h0"]))
    x_a, _ = rs.transducer(d=d)
    alpha = np.linspace(-0.6, 0.6, 1201)
    engine.setup(alpha, np.ones(1201) * d, x_a[32], d, x_a, **kw)
    rng = np.random.default_rng(9)
    fx, fz = rng.uniform(-0.02, 0.02, 50), rng.uniform(0.0, 0.045, 50)   # focal points inside the steel
    tof, kmin, al = engine.fermat_tof(fx, fz, with_alpha=True)
    lens = Lens(flat=True, **kw)
    elems = np.arange(0, 65, 8)
    o_tof, o_k, o_al, gap = fermat_table(lens, alpha, x_a[elems], d, fx, fz)
    assert np.abs(tof[elems] - o_tof).max() < 1e-15
    assert np.array_equal(kmin[elems][gap > 1e-18], o_k[gap > 1e-18])
    # Snell at the plane: sin(th1)/c1 == sin(th2)/c2 with the angles to the z axis
    px, pz = lens.point(al)
    s1 = (px - x_a[:, None]) / np.hypot(px - x_a[:, None], pz - d)
    s2 = (fx[None, :] - px) / np.hypot(fx[None, :] - px, fz[None, :] - pz)
    interior = (kmin > 0) & (kmin < len(alpha) - 1)
    assert np.abs(s1 / kw["c1"] - s2 / kw["c2"])[interior].max() < 1e-12
    # straight down: element above the focal point -> no refraction
    t0, _ = engine.fermat_tof(np.array([x_a[40]]), np.array([0.02]))
    assert abs(t0[40, 0] - (kw["l0"] / kw["c1"] + (kw["h0"] - 0.02) / kw["c2"])) < 1e-17
