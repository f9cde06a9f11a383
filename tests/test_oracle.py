"""CPU: pin the oracle (oracle/tofrt_oracle.c) against the reference's own
fixtures and against outputs of the live reference (tests/golden/*.npz)."""
import numpy as np
import pytest

NOISE_FLOOR = 1e-13  # metres; SURVEY.md H1: below this the reference's own idx is polyfit noise


def _setup(golden, key="sweep"):
    g = golden[key]
    return g["x_a"], g["alpha"]


@pytest.mark.parametrize("which", ["oracle_libm", "oracle_dm"])
def test_table_matches_database_2_csv_and_live_reference(which, golden, request):
    o = request.getfixturevalue(which)
    sw, db = golden["sweep"], golden["database_2"]
    p = o.params()
    x_a, alpha = _setup(golden)
    zf = np.ones(len(alpha)) * p.d
    tof, k_hit, idx = o.table(p, sw["radius"], sw["offset"], x_a[32], p.d, zf, alpha, x_a, want_idx=True)
    # live reference (main_rt.py:487-501 replicated on its outputs)
    assert np.array_equal(k_hit.T, sw["k_hit"])
    assert np.abs(tof.T - sw["tof"]).max() < 1e-15
    # the reference's committed fixture, row order r-major / offset / elem_idx
    assert np.array_equal(db["elem_idx"].reshape(210, 65), np.tile(np.arange(65), (210, 1)))
    assert np.array_equal(db["radius"].reshape(210, 65)[:, 0], sw["radius"])
    assert np.array_equal(db["offset"].reshape(210, 65)[:, 0], sw["offset"])
    assert np.array_equal((k_hit.T >= 0), db["hitted"].reshape(210, 65))
    assert np.abs(tof.T - db["tof_total"].reshape(210, 65)).max() < 1e-15
    assert (k_hit >= 0).sum() == 498
    # discrete return-path index: identical wherever the reference's own decision is not noise
    ridx = sw["idx"].astype(np.int32)
    oi = np.where(idx < -1, -1, idx)
    line_ok = ridx != -2
    differs = (oi != ridx) & line_ok
    assert not (differs & (sw["margin"] > NOISE_FLOOR)).any()
    assert differs.sum() < 3000  # all in the 12 degenerate geometries (2 324 with NumPy libm)
    assert (idx[~line_ok] == 0).all()  # NaN line: every sign "changes", idx = 0 -> NaN intersection


@pytest.mark.parametrize("which", ["oracle_libm", "oracle_dm"])
def test_ray_records_match_live_reference(which, golden, request):
    o = request.getfixturevalue(which)
    rays = golden["rays"]
    p = o.params()
    x_a, alpha = _setup(golden)
    zf = np.ones(len(alpha)) * p.d
    for i, (r, off) in enumerate(zip(rays["radius"], rays["offset"])):
        res = o.shoot_rays(p, r, off, x_a[32], p.d, zf, alpha)
        for key, tol in (("pipe_x", 1e-15), ("pipe_z", 1e-15), ("lens_2_x", 1e-14), ("lens_2_z", 1e-14),
                         ("target_x", 1e-11)):
            a, b = res[key], rays[key][i]
            assert np.array_equal(np.isnan(a), np.isnan(b)), (key, r, off)
            ok = ~np.isnan(a)
            # target_x = (z_f - b)/a amplifies an ulp of the slope by |target_x|
            scale = np.maximum(1.0, np.abs(b[ok])) if key == "target_x" else 1.0
            assert (np.abs(a[ok] - b[ok]) <= tol * scale).all(), (key, r, off)
        assert np.array_equal(res["lens_1_x"], golden["sweep"]["lens_x"]) or \
            np.abs(res["lens_1_x"] - golden["sweep"]["lens_x"]).max() < 1e-16


@pytest.mark.parametrize("which", ["oracle_libm", "oracle_dm"])
def test_compare_case_matches_compare_csv(which, golden, request):
    """main_compare.py:463-524: K = N = 1810, r = 0.037, off = 0.0038, four legs per ray."""
    o = request.getfixturevalue(which)
    cl, cc = golden["compare_live"], golden["compare"]
    p = o.params()
    x_a, alpha = cl["x_a"], cl["alpha"]
    assert np.array_equal(alpha, cc["alpha"])
    zf = np.ones(len(alpha)) * p.d
    res = o.shoot_rays(p, 0.037, 0.0038, x_a[32], p.d, zf, alpha)
    legs = o.ray_legs(p, x_a[32], p.d, res)
    for ref_legs in (cl["legs"], cc["tof"]):
        assert np.array_equal(np.isnan(legs), np.isnan(ref_legs))
        ok = ~np.isnan(legs)
        assert np.abs(legs[ok] - ref_legs[ok]).max() < 1e-17
    # hitted flag of main_compare.py:518-521 (atol = 1e-4, any element)
    tx = res["target_x"]
    hitted = np.array([np.isclose(t, x_a, atol=1e-4).any() for t in tx])
    assert np.array_equal(hitted, cc["hitted"]) and hitted.sum() == 393
    ridx = cl["idx"]
    oi = np.where(res["idx"] < -1, -1, res["idx"])
    bad = (oi != ridx) & (ridx != -2)
    assert not (bad & (cl["margin"] > NOISE_FLOOR)).any()


def test_off_centre_shooters_match_live_reference(golden, oracle_libm):
    """No fixture exists for other shooters; the live reference is the pin."""
    sh = golden["shooters"]
    o, p = oracle_libm, oracle_libm.params()
    x_a, alpha = sh["x_a"], sh["alpha"]
    zf = np.ones(len(alpha)) * p.d
    for i, s in enumerate(sh["shooter"]):
        tof, k_hit, idx = o.table(p, [sh["radius"][i]], [sh["offset"][i]], x_a[s], p.d, zf, alpha, x_a,
                                  want_idx=True)
        assert np.array_equal(k_hit[:, 0], sh["k_hit"][i])
        assert np.abs(tof[:, 0] - sh["tof"][i]).max() < 1e-15
        oi = np.where(idx[0] < -1, -1, idx[0])
        bad = (oi != sh["idx"][i]) & (sh["idx"][i] != -2)
        assert not (bad & (sh["margin"][i] > NOISE_FLOOR)).any()


def test_known_answers(oracle_libm):
    """Analytic KATs of SURVEY.md section 4."""
    o, p = oracle_libm, oracle_libm.params()
    from ray_tracing_ultrasound_b200 import reference_setup as rs
    x_a, z_a = rs.transducer()
    alpha, zf = rs.launch_angles(), rs.focus_line()
    lx, lz = o.lens(p, alpha)
    # constant-TOF lens (main_rt.py:181-187)
    T = p.l0 / p.c1 + p.h0 / p.c2
    tofs = np.hypot(lx - 0.0, lz - p.d) / p.c1 + np.hypot(lx, lz) / p.c2
    assert np.abs(tofs - T).max() < 1e-18
    # h(0) = h0
    x0, z0 = o.lens(p, np.array([0.0, 0.0]))
    assert abs(z0[0] - p.h0) < 1e-16 and x0[0] == 0.0
    # normal-incidence echo: centre ray, centred pipe, r < h0
    for r in (0.01, 0.02, 0.03, 0.05):
        tof, k_hit, _ = o.table(p, [r], [0.0], x_a[32], z_a[32], zf, alpha, x_a)
        assert k_hit[32, 0] >= 0
        assert abs(tof[32, 0] - 2 * (p.l0 / p.c1 + (p.h0 - r) / p.c2)) < 1e-18


def test_twin_and_libm_oracles_agree(golden, oracle_libm, oracle_dm):
    """The device-math twin differs from the libm build only by elementary-function ulps."""
    rng = np.random.default_rng(0)
    from ray_tracing_ultrasound_b200 import reference_setup as rs
    x_a, z_a = rs.transducer()
    alpha, zf = rs.launch_angles(), rs.focus_line()
    rad, off = rng.uniform(0.01, 0.1, 128), rng.uniform(-0.01, 0.01, 128)
    p = oracle_libm.params()
    t1, k1, i1 = oracle_libm.table(p, rad, off, x_a[32], z_a[32], zf, alpha, x_a, want_idx=True)
    t2, k2, i2 = oracle_dm.table(p, rad, off, x_a[32], z_a[32], zf, alpha, x_a, want_idx=True)
    assert np.array_equal(k1, k2)
    assert np.abs(t1 - t2).max() < 1e-15
    assert (i1 != i2).mean() < 1e-4


def test_argument_errors(oracle_libm):
    p = oracle_libm.params()
    with pytest.raises(ValueError):
        oracle_libm.shoot_rays(p, 0.05, 0.0, 0.0, p.d, np.array([p.d]), np.array([0.0]))
