"""CPU: the vector form of the return-path interface laws (csrc/tofrt_raymath.h, TOFRT_MATH_VECTOR),
evaluated through the oracle's device-math twin, against the angle form (the reference's operation order)
and against the independent libm oracle: same hit decisions, same first-sign-change indices wherever the
decision is not rounding noise, travel times far inside 1e-12 s, same NaN pattern."""
import numpy as np

from ray_tracing_ultrasound_b200 import reference_setup as rs


def _tables(oracle_dm, oracle_libm, radii, offsets, n_alpha=905):
    x_a, z_a = rs.transducer()
    alpha, zf = rs.launch_angles(n_alpha), rs.focus_line(n_alpha)
    p = oracle_dm.params()
    args = (p, radii, offsets, x_a[32], z_a[32], zf, alpha, x_a)
    try:
        oracle_dm.set_math(True)
        vec = oracle_dm.table(*args, want_idx=True)
    finally:
        oracle_dm.set_math(False)
    return vec, oracle_dm.table(*args, want_idx=True), oracle_libm.table(*args, want_idx=True)


def test_vector_form_on_the_fixture_sweep(oracle_dm, oracle_libm, golden):
    radii, offsets = rs.sweep_grid()
    (t1, k1, i1), (t0, k0, i0), (tl, kl, il) = _tables(oracle_dm, oracle_libm, radii, offsets)
    assert np.array_equal(k1, k0) and np.array_equal(k1, kl)
    hit = k1 >= 0
    # one entry sits at 5e-17 s: geometry (r, off) = (0.01, -0.01) grazes the pipe at its rightmost point (discriminant ~ 0,
    # so x_q carries sqrt(eps)-amplified rounding in every formulation); everything else agrees to < 2e-18 s
    d0, dl = np.abs(t1 - t0)[hit], np.abs(t1 - tl)[hit]
    assert d0.max() < 2e-16 and dl.max() < 2e-16
    assert np.sort(d0)[-2] < 1e-17 and np.sort(dl)[-2] < 1e-17
    # the reference's own fixture
    db = golden["database_2"]
    assert np.array_equal(hit.T, db["hitted"].reshape(210, 65))
    assert np.abs(t1.T - db["tof_total"].reshape(210, 65)).max() < 1e-12
    # the index differs only in the 12 degenerate geometries (the ray retraces its path through a lens sample)
    degenerate = (offsets == 0) | ((radii == 0.01) & (np.abs(offsets) == 0.01))
    assert degenerate.sum() == 12
    assert np.array_equal(i1[~degenerate], i0[~degenerate]) and np.array_equal(i1[~degenerate], il[~degenerate])
    sw = golden["sweep"]     # the unmodified reference: identical wherever its own decision exceeds the noise floor
    ridx = sw["idx"].astype(np.int32)
    gi = np.where(i1 < -1, -1, i1)
    assert not ((gi != ridx) & (ridx != -2) & (sw["margin"] > 1e-13)).any()


def test_vector_form_on_random_and_grazing_geometries(oracle_dm, oracle_libm):
    rng = np.random.default_rng(3)
    radii = np.concatenate([rng.uniform(0.003, 0.15, 400), rng.uniform(0.07, 0.12, 200)])
    offsets = np.concatenate([rng.uniform(-0.03, 0.03, 400), rng.uniform(-0.05, 0.05, 200)])
    (t1, k1, i1), (t0, k0, i0), (tl, kl, il) = _tables(oracle_dm, oracle_libm, radii, offsets)
    assert np.array_equal(k1, k0) and np.array_equal(k1, kl)
    assert np.array_equal(i1, i0) and np.array_equal(i1, il)
    hit = k1 >= 0
    assert hit.sum() > 500 and np.abs(t1 - tl)[hit].max() < 1e-16


def test_vector_form_per_ray_nan_pattern(oracle_dm):
    x_a, z_a = rs.transducer()
    alpha, zf = rs.launch_angles(), rs.focus_line()
    p = oracle_dm.params()
    for r, off in ((0.037, 0.0038), (0.1, -0.01), (0.09, 0.03), (0.005, 0.02)):
        a = oracle_dm.shoot_rays(p, r, off, x_a[32], z_a[32], zf, alpha)
        try:
            oracle_dm.set_math(True)
            b = oracle_dm.shoot_rays(p, r, off, x_a[32], z_a[32], zf, alpha)
        finally:
            oracle_dm.set_math(False)
        for key in ("a_l", "lens_2_x", "target_x"):
            assert np.array_equal(np.isnan(a[key]), np.isnan(b[key])), key
        ok = ~np.isnan(a["target_x"]) & (np.abs(a["a_l"]) < 1e6)      # near-vertical return lines are ill-conditioned in both forms
        if ok.any():   # landing points agree to ~1e-13 of their size (1 mm floor): nothing near the 1e-6 m hit tolerance
            err = np.abs(a["target_x"] - b["target_x"])[ok] / np.maximum(np.abs(a["target_x"][ok]), 1e-3)
            assert err.max() < 2e-12


def test_vector_form_other_materials_and_lenses(oracle_dm, oracle_libm):
    """Random wave speeds, lens dimensions, ray counts, arrays, shooters and tolerances (the parameter space of
    tools/fuzz_parity.py): the vector form keeps the hit decisions of the libm oracle (reference operation order)."""
    n_hits = 0
    for i in range(8):
        r4 = np.random.default_rng(7000 + i)
        kw = dict(c1=r4.uniform(3000, 7000), c2=r4.uniform(900, 2500), l0=r4.uniform(0.05, 0.2), h0=r4.uniform(0.04, 0.12))
        K = int(r4.integers(64, 700))
        nel = int(r4.choice([16, 64, 128]))
        d = float(np.float64(kw["l0"]) + np.float64(kw["h0"]))
        xa, _ = rs.transducer(num_elements=nel, pitch=np.float64(r4.uniform(2e-4, 8e-4)), d=d)
        al = np.linspace(-r4.uniform(0.3, 0.85), r4.uniform(0.3, 0.85), K)
        zf = np.ones(K) * d
        sh = int(r4.integers(0, len(xa)))
        rtol, atol = 10.0 ** r4.uniform(-7, -4), 10.0 ** r4.uniform(-6, -4)
        radii, offsets = r4.uniform(0.05, 1.1, 250) * kw["h0"], r4.uniform(-0.3, 0.3, 250) * kw["h0"]
        p = oracle_dm.params(**kw)
        args = (p, radii, offsets, xa[sh], d, zf, al, xa)
        try:
            oracle_dm.set_math(True)
            t1, k1, i1 = oracle_dm.table(*args, rtol=rtol, atol=atol, want_idx=True)
        finally:
            oracle_dm.set_math(False)
        tl, kl, il, margin = oracle_libm.table(*args, rtol=rtol, atol=atol, want_idx=True, want_margin=True)
        assert np.array_equal(k1, kl), i
        hit = k1 >= 0
        n_hits += int(hit.sum())
        if hit.any():
            assert np.abs(t1 - tl)[hit].max() < 1e-12
        # the index may only differ where the libm oracle's own decision is rounding noise: ray by ray, its two
        # deciding diffs must both sit below the 1e-13 m floor; and such rays are rare
        differs = i1 != il
        assert not (differs & (margin > 1e-13)).any(), (i, int((differs & (margin > 1e-13)).sum()))
        assert differs.mean() < 2e-3
    assert n_hits > 1000


def test_oracles_against_the_live_reference_with_a_128_element_array(oracle_dm, oracle_libm, golden_array128):
    """BASELINE configs 3/4 (E = 129): the libm oracle and both forms of the twin against the UNMODIFIED reference
    run with num_elements = 128 on 96 geometries of the 1024 x 1024 grid (tests/golden/array128.npz)."""
    gd = golden_array128
    x_a, z_a = rs.transducer(num_elements=128)
    alpha, zf = rs.launch_angles(), rs.focus_line()
    assert np.array_equal(x_a, gd["x_a"]) and np.array_equal(alpha, gd["alpha"])
    radii, offsets = rs.synthetic_grid(1024, 1024)
    flat = gd["flat_index"]
    assert np.array_equal(radii[flat], gd["radius"]) and np.array_equal(offsets[flat], gd["offset"])
    p = oracle_libm.params()
    args = (p, radii[flat], offsets[flat], x_a[64], z_a[64], zf, alpha, x_a)
    ridx = gd["idx"].astype(np.int32)                     # reference encoding: -1 no sign change, -2 NaN line
    tables = [oracle_libm.table(*args, want_idx=True), oracle_dm.table(*args, want_idx=True)]
    try:
        oracle_dm.set_math(True)
        tables.append(oracle_dm.table(*args, want_idx=True))
    finally:
        oracle_dm.set_math(False)
    for tof, k_hit, idx in tables:
        assert np.array_equal(k_hit.T, gd["k_hit"])
        assert np.abs(tof.T - gd["tof"]).max() < 1e-12
        gi = np.where(idx < -1, -1, idx)
        differs = (gi != ridx) & (ridx != -2)
        assert not (differs & (gd["margin"] > 1e-13)).any()
        assert differs.sum() <= 2e-3 * ridx.size
