"""GPU: parity of the CUDA path (through the C ABI) with the oracle and the
reference's fixtures.

Bars (BASELINE.json north_star): discrete indices (idx, k_hit) bit-exact;
travel times within 1e-12 s of the reference.  Against the oracle's device-math
twin (same elementary functions, gcc-compiled) every output is bit-exact.
"""
import numpy as np
import pytest

from ray_tracing_ultrasound_b200 import reference_setup as rs

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("math_mode")]

TOF_TOL = 1e-12  # seconds, north_star
NOISE_FLOOR = 1e-13


def _std(engine, n_alpha=rs.NUM_ALPHA_POINTS_RT, shooter=32, atol=1e-6):
    x_a, z_a = rs.transducer()
    alpha, zf = rs.launch_angles(n_alpha), rs.focus_line(n_alpha)
    engine.setup(alpha, zf, x_a[shooter], z_a[shooter], x_a, atol=atol)
    return x_a, z_a, alpha, zf


def _same(a, b):
    """bit-exact incl. NaN positions"""
    return np.array_equal(a, b, equal_nan=True)


def test_fixture_sweep_bit_exact_vs_twin_and_within_tol_vs_reference(engine, golden, oracle_dm, oracle_libm):
    """BASELINE config 1: the main_rt.py default sweep (210 geometries x 65 receivers)."""
    x_a, z_a, alpha, zf = _std(engine)
    radii, offsets = rs.sweep_grid()
    t = engine.tof_table(radii, offsets, records=("idx", "target_x", "pipe_x", "pipe_z", "lens2_x", "lens2_z"))
    p = oracle_dm.params()
    # bit-exact twin
    tof, k_hit, idx = oracle_dm.table(p, radii, offsets, x_a[32], z_a[32], zf, alpha, x_a, want_idx=True)
    assert np.array_equal(t.k_hit, k_hit.T)
    assert np.array_equal(t.records["idx"], idx)
    assert np.array_equal(t.tof, tof.T)
    # independent libm oracle: indices equal, times within tolerance
    tof2, k_hit2, idx2, margin2 = oracle_libm.table(p, radii, offsets, x_a[32], z_a[32], zf, alpha, x_a, want_idx=True,
                                                    want_margin=True)
    assert np.array_equal(t.k_hit, k_hit2.T)
    assert np.abs(t.tof - tof2.T).max() <= TOF_TOL
    # ... and idx, ray by ray: identical wherever the oracle's own decision is not rounding noise (its two deciding
    # diffs exceed 1e-13 m); what differs sits in the 12 degenerate geometries and is counted
    from test_gpu_configs import assert_idx_agrees
    n_noise = assert_idx_agrees(t.records["idx"], idx2, margin2, max_noise_frac=2e-2, what="fixture sweep")
    degenerate = (offsets == 0) | ((radii == 0.01) & (np.abs(offsets) == 0.01))
    assert np.array_equal(t.records["idx"][~degenerate], idx2[~degenerate])
    print(f"fixture sweep: {n_noise} of {idx2.size} rays differ from the libm oracle in idx, all below the noise floor")
    # the reference's fixture database_2.csv and the live reference
    db, sw = golden["database_2"], golden["sweep"]
    assert np.array_equal(t.hit, db["hitted"].reshape(210, 65))
    assert np.abs(t.tof - db["tof_total"].reshape(210, 65)).max() <= TOF_TOL
    assert np.array_equal(t.k_hit, sw["k_hit"])
    assert np.abs(t.tof - sw["tof"]).max() <= TOF_TOL
    ridx = sw["idx"].astype(np.int32)
    gi = np.where(t.records["idx"] < -1, -1, t.records["idx"])
    differs = (gi != ridx) & (ridx != -2)
    assert not (differs & (sw["margin"] > NOISE_FLOOR)).any()
    assert np.array_equal(np.isnan(t.records["target_x"]), np.unpackbits(sw["nan_target"], axis=1)[:, :905].astype(bool))
    # per-ray records vs the twin, ray by ray
    for g in (0, 10, 57, 209):
        res = oracle_dm.shoot_rays(p, radii[g], offsets[g], x_a[32], z_a[32], zf, alpha)
        for a, b in (("pipe_x", "pipe_x"), ("pipe_z", "pipe_z"), ("lens2_x", "lens_2_x"),
                     ("lens2_z", "lens_2_z"), ("target_x", "target_x")):
            assert _same(t.records[a][g], res[b]), (a, g)


def test_compare_case_vs_compare_csv(engine, golden, oracle_dm):
    """main_compare.py: K = N = 1810, r = 0.037, off = 0.0038, legs + hitted(atol 1e-4)."""
    x_a, z_a, alpha, zf = _std(engine, rs.NUM_ALPHA_POINTS_COMPARE, atol=1e-4)
    res = engine.shoot_rays(0.037, 0.0038, with_idx=True, with_legs=True)
    cc, cl = golden["compare"], golden["compare_live"]
    legs = res["legs"]
    for ref in (cc["tof"], cl["legs"]):
        assert np.array_equal(np.isnan(legs), np.isnan(ref))
        ok = ~np.isnan(legs)
        assert np.abs(legs[ok] - ref[ok]).max() <= TOF_TOL
    hitted = np.array([np.isclose(tx, x_a, atol=1e-4).any() for tx in res["target_x"]])
    assert np.array_equal(hitted, cc["hitted"])
    p = oracle_dm.params()
    tw = oracle_dm.shoot_rays(p, 0.037, 0.0038, x_a[32], z_a[32], zf, alpha)
    assert np.array_equal(res["idx"], tw["idx"])
    assert _same(res["target_x"], tw["target_x"])
    assert _same(legs, oracle_dm.ray_legs(p, x_a[32], z_a[32], tw))
    assert np.array_equal(res["lens_1_x"], tw["lens_1_x"]) and np.array_equal(res["lens_1_z"], tw["lens_1_z"])
    assert np.array_equal(res["target_z"], zf)


def test_random_geometries_bit_exact(engine, oracle_dm, oracle_libm):
    x_a, z_a, alpha, zf = _std(engine)
    rng = np.random.default_rng(0)
    G = 3000
    radii = rng.uniform(0.005, 0.12, G)
    offsets = rng.uniform(-0.02, 0.02, G)
    t = engine.tof_table(radii, offsets, records=("idx", "target_x"))
    p = oracle_dm.params()
    tof, k_hit, idx = oracle_dm.table(p, radii, offsets, x_a[32], z_a[32], zf, alpha, x_a, want_idx=True)
    assert np.array_equal(t.records["idx"], idx)
    assert np.array_equal(t.k_hit, k_hit.T)
    assert np.array_equal(t.tof, tof.T)
    tof2, k_hit2, idx2, margin2 = oracle_libm.table(p, radii, offsets, x_a[32], z_a[32], zf, alpha, x_a, want_idx=True,
                                                    want_margin=True)
    assert np.array_equal(t.k_hit, k_hit2.T)
    assert np.abs(t.tof - tof2.T).max() <= TOF_TOL
    from test_gpu_configs import assert_idx_agrees
    assert_idx_agrees(t.records["idx"], idx2, margin2, what="random geometries")


def test_off_centre_shooters(engine, golden, oracle_dm):
    sh = golden["shooters"]
    p = oracle_dm.params()
    for i, s in enumerate(sh["shooter"]):
        x_a, z_a, alpha, zf = _std(engine, shooter=int(s))
        t = engine.tof_table([sh["radius"][i]], [sh["offset"][i]], records=("idx",))
        assert np.array_equal(t.k_hit[0], sh["k_hit"][i])
        assert np.abs(t.tof[0] - sh["tof"][i]).max() <= TOF_TOL
        tof, k_hit, idx = oracle_dm.table(p, [sh["radius"][i]], [sh["offset"][i]], x_a[s], z_a[s], zf, alpha,
                                          x_a, want_idx=True)
        assert np.array_equal(t.records["idx"], idx) and np.array_equal(t.tof[0], tof[:, 0])


@pytest.mark.parametrize("K", [2, 3, 17, 32, 33, 100, 1810, 4096])
def test_ray_counts(engine, oracle_dm, K):
    """K not a multiple of the warp/chunk size, the minimum K = 2, and the largest staged K."""
    x_a, z_a = rs.transducer()
    alpha, zf = rs.launch_angles(K), rs.focus_line(K)
    engine.setup(alpha, zf, x_a[32], z_a[32], x_a)
    rng = np.random.default_rng(K)
    G = 40 if K <= 1810 else 12
    radii, offsets = rng.uniform(0.01, 0.1, G), rng.uniform(-0.01, 0.01, G)
    t = engine.tof_table(radii, offsets, records=("idx", "target_x"))
    p = oracle_dm.params()
    tof, k_hit, idx = oracle_dm.table(p, radii, offsets, x_a[32], z_a[32], zf, alpha, x_a, want_idx=True)
    assert np.array_equal(t.records["idx"], idx)
    assert np.array_equal(t.k_hit, k_hit.T) and np.array_equal(t.tof, tof.T)


def test_receiver_edge_cases(engine, oracle_dm):
    """Unsorted, duplicated, NaN and infinite receivers; one receiver; huge tolerance."""
    x_a, z_a = rs.transducer()
    alpha, zf = rs.launch_angles(), rs.focus_line()
    p = oracle_dm.params()
    rng = np.random.default_rng(5)
    radii, offsets = rng.uniform(0.01, 0.1, 64), rng.uniform(-0.01, 0.01, 64)
    weird = np.concatenate([x_a[::-1][:20], [np.nan, np.inf, -np.inf, 0.0, 0.0, 1e-7, -1e-7, 5.0, -5.0], x_a[10:14]])
    for x_rx, rtol, atol in ((weird, 1e-5, 1e-6), (np.array([0.0]), 1e-5, 1e-6), (x_a, 1e-5, 1e-3),
                             (x_a, 0.5, 10.0), (x_a, 0.0, 0.0)):
        engine.setup(alpha, zf, x_a[32], z_a[32], x_rx, rtol=rtol, atol=atol)
        t = engine.tof_table(radii, offsets)
        tof, k_hit, _ = oracle_dm.table(p, radii, offsets, x_a[32], z_a[32], zf, alpha, x_rx, rtol=rtol, atol=atol)
        assert np.array_equal(t.k_hit, k_hit.T), (len(x_rx), rtol, atol)
        assert np.array_equal(t.tof, tof.T), (len(x_rx), rtol, atol)


def test_degenerate_geometries(engine, oracle_dm):
    """Rays that miss the pipe (the reference crashes there: NaN-out is the documented
    extension), pipe swallowing the lens, zero / negative / NaN radius, huge offsets."""
    x_a, z_a, alpha, zf = _std(engine)
    radii = np.array([1e-4, 0.001, 0.01, 0.2, 0.5, 0.0, -0.01, np.nan, 0.05, 0.05, 0.05, 1e-9, 0.0884, 0.03])
    offsets = np.array([0.0, 0.05, 0.03, 0.0, 0.1, 0.0, 0.0, 0.0, np.nan, 1.0, -0.06, 0.0, 0.0, np.inf])
    t = engine.tof_table(radii, offsets, records=("idx", "target_x", "pipe_x", "lens2_x"))
    p = oracle_dm.params()
    tof, k_hit, idx = oracle_dm.table(p, radii, offsets, x_a[32], z_a[32], zf, alpha, x_a, want_idx=True)
    assert np.array_equal(t.records["idx"], idx)
    assert np.array_equal(t.k_hit, k_hit.T) and np.array_equal(t.tof, tof.T)
    for g in range(len(radii)):
        res = oracle_dm.shoot_rays(p, radii[g], offsets[g], x_a[32], z_a[32], zf, alpha)
        assert _same(t.records["target_x"][g], res["target_x"]), g
        assert _same(t.records["pipe_x"][g], res["pipe_x"]), g
        assert _same(t.records["lens2_x"][g], res["lens_2_x"]), g


def test_argument_errors_and_empty_input(engine):
    x_a, z_a = rs.transducer()
    with pytest.raises(ValueError):  # main_rt.py:24-29
        engine.setup(np.array([0.0]), np.array([0.21]), 0.0, 0.21, x_a)
    with pytest.raises(ValueError):
        engine.setup(np.zeros(4), np.zeros(5), 0.0, 0.21, x_a)
    _std(engine)
    t = engine.tof_table([], [])
    assert t.tof.shape == (0, 65) and t.k_hit.shape == (0, 65)
    with pytest.raises(ValueError):
        engine.tof_table([0.01, 0.02], [0.0])


def test_device_resident_api_matches_host_api(engine):
    import torch
    x_a, z_a, alpha, zf = _std(engine)
    radii, offsets = rs.synthetic_grid(48, 40)
    th = engine.tof_table(radii, offsets)
    r_d = torch.from_numpy(radii).cuda()
    o_d = torch.from_numpy(offsets).cuda()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        td = engine.tof_table_cuda(r_d, o_d, records=("idx",))
    s.synchronize()
    assert np.array_equal(td.tof.cpu().numpy(), th.tof)
    assert np.array_equal(td.k_hit.cpu().numpy(), th.k_hit)
    # determinism: same bytes on a second run
    td2 = engine.tof_table_cuda(r_d, o_d)
    torch.cuda.synchronize()
    assert torch.equal(td.tof, td2.tof) and torch.equal(td.k_hit, td2.k_hit)


def test_device_api_rejects_wrong_outputs_and_unknown_records(engine):
    """A mis-shaped / mis-typed / non-contiguous output tensor would be an out-of-bounds device write: refuse it."""
    import torch
    _std(engine)
    radii, offsets = rs.synthetic_grid(8, 8)
    r_d, o_d = torch.from_numpy(radii).cuda(), torch.from_numpy(offsets).cuda()
    ok_t, ok_k = torch.empty((64, 65), dtype=torch.float64, device="cuda"), torch.empty((64, 65), dtype=torch.int32, device="cuda")
    engine.tof_table_cuda(r_d, o_d, out_tof=ok_t, out_k_hit=ok_k)
    for bad_t, bad_k in ((torch.empty((63, 65), dtype=torch.float64, device="cuda"), ok_k),
                         (torch.empty((64, 65), dtype=torch.float32, device="cuda"), ok_k),
                         (ok_t, torch.empty((64, 65), dtype=torch.int64, device="cuda")),
                         (torch.empty((64, 130), dtype=torch.float64, device="cuda")[:, ::2], ok_k),
                         (torch.empty((64, 65), dtype=torch.float64), ok_k)):
        with pytest.raises(ValueError):
            engine.tof_table_cuda(r_d, o_d, out_tof=bad_t, out_k_hit=bad_k)
    with pytest.raises(ValueError):
        engine.tof_table_f32_cuda(r_d, o_d, out_tof=ok_t)          # float64 where the FP32 mode writes float32
    with pytest.raises(ValueError):
        engine.tof_table_cuda(r_d, o_d, records=("idx", "no_such_record"))
    x_a, z_a = rs.transducer()
    engine.setup(rs.launch_angles(), rs.focus_line(), x_a[:3], z_a[:3], x_a)
    with pytest.raises(ValueError):
        engine.shoot_rays(0.05, 0.0)                               # several shooters: not the single-shot call
    torch.cuda.synchronize()


def test_calls_in_flight_on_two_streams_do_not_share_a_work_counter(engine):
    """Every launch draws its own work-counter slot: two table calls queued back to back on different streams (and a
    host-buffer call on the library's own streams) must each produce complete, identical tables."""
    import torch
    _std(engine)
    radii, offsets = rs.synthetic_grid(256, 64)
    r_d, o_d = torch.from_numpy(radii).cuda(), torch.from_numpy(offsets).cuda()
    ref = engine.tof_table_cuda(r_d, o_d)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for rep in range(4):
        with torch.cuda.stream(s1):
            a = engine.tof_table_cuda(r_d, o_d)
        with torch.cuda.stream(s2):
            b = engine.tof_table_cuda(r_d, o_d)
        outs += [a, b]
    th = engine.tof_table(radii[:4096], offsets[:4096])            # runs on the context's own non-blocking streams
    torch.cuda.synchronize()
    for t in outs:
        assert torch.equal(t.tof, ref.tof) and torch.equal(t.k_hit, ref.k_hit)
    assert np.array_equal(th.tof, ref.tof[:4096].cpu().numpy())


def test_two_engines_with_different_sizes_on_one_gpu(engine):
    """The opt-in shared-memory limit is per kernel and per DEVICE: a second, smaller setup on another context must
    not lower it under a live, larger one (main_rt's K = 905 engine next to main_compare's K = 1810 one)."""
    from ray_tracing_ultrasound_b200 import TofEngine
    x_a, z_a = rs.transducer()
    big = TofEngine(0).setup(rs.launch_angles(4096), rs.focus_line(4096), x_a[32], z_a[32], x_a)
    small = TofEngine(0).setup(rs.launch_angles(64), rs.focus_line(64), x_a[32], z_a[32], x_a)
    radii, offsets = rs.synthetic_grid(6, 6)
    a = big.tof_table(radii, offsets)
    b = small.tof_table(radii, offsets)
    a2 = big.tof_table(radii, offsets)
    assert np.array_equal(a.tof, a2.tof) and np.array_equal(a.k_hit, a2.k_hit) and b.tof.shape == (36, 65)
    big.close(); small.close()


def test_small_g_latency_mode_equals_the_regular_kernel(engine, oracle_dm):
    """TOFRT_OPT_SMALL_G: a CTA per geometry with one chunk per warp (few geometries) gives the regular kernel's
    bytes -- table, records, hit list -- on the reference's 210-geometry sweep, a single geometry, several shooters."""
    x_a, z_a, alpha, zf = _std(engine)
    radii, offsets = rs.sweep_grid()
    keys = ("idx", "target_x", "pipe_x", "lens2_z", "legs")
    engine.set_small_g(False)
    a = engine.tof_table(radii, offsets, records=keys)
    one_a = engine.tof_table(radii[57:58], offsets[57:58])
    hits_a = engine.tof_hits(radii, offsets)
    engine.set_small_g(True)
    b = engine.tof_table(radii, offsets, records=keys)
    one_b = engine.tof_table(radii[57:58], offsets[57:58])
    hits_b = engine.tof_hits(radii, offsets)
    assert np.array_equal(a.tof, b.tof) and np.array_equal(a.k_hit, b.k_hit) and int(b.hit.sum()) == 498
    for k in keys:
        assert _same(a.records[k], b.records[k]), k
    assert np.array_equal(one_a.tof, one_b.tof) and np.array_equal(one_a.k_hit, one_b.k_hit)
    assert all(np.array_equal(u, v) for u, v in zip(hits_a, hits_b)) and len(hits_b[0]) == 498
    p = oracle_dm.params()
    tof, k_hit, idx = oracle_dm.table(p, radii, offsets, x_a[32], z_a[32], zf, alpha, x_a, want_idx=True)
    assert np.array_equal(b.k_hit, k_hit.T) and np.array_equal(b.tof, tof.T) and np.array_equal(b.records["idx"], idx)
    # several shooters (3 x 20 items), and the K = 1810 case of main_compare.py (57 chunks > 32 warps: regular kernel)
    engine.setup(alpha, zf, x_a[30:33], z_a[30:33], x_a)
    engine.set_small_g(False)
    m0 = engine.tof_table(radii[:20], offsets[:20])
    engine.set_small_g(True)
    m1 = engine.tof_table(radii[:20], offsets[:20])
    assert m1.tof.shape == (3, 20, 65) and np.array_equal(m0.tof, m1.tof) and np.array_equal(m0.k_hit, m1.k_hit)
    _std(engine, rs.NUM_ALPHA_POINTS_COMPARE)
    c1 = engine.tof_table(radii[:3], offsets[:3])
    engine.set_small_g(False)
    c0 = engine.tof_table(radii[:3], offsets[:3])
    engine.set_small_g(True)
    assert np.array_equal(c0.tof, c1.tof) and np.array_equal(c0.k_hit, c1.k_hit)


def test_full_size_grid_properties(engine, oracle_dm):
    """BASELINE config 3 at full size (129 receivers, 1024 x 1024 geometries): the table is
    checked through size-independent properties -- a random subsample against the oracle,
    sub-grid consistency (a slice recomputed alone gives the same bytes), and invariants."""
    import torch
    x_a, z_a = rs.transducer(num_elements=128)
    alpha, zf = rs.launch_angles(), rs.focus_line()
    engine.setup(alpha, zf, x_a[64], z_a[64], x_a)
    radii, offsets = rs.synthetic_grid(1024, 1024)
    r_d, o_d = torch.from_numpy(radii).cuda(), torch.from_numpy(offsets).cuda()
    t = engine.tof_table_cuda(r_d, o_d)
    torch.cuda.synchronize()
    G, E = len(radii), len(x_a)
    assert t.tof.shape == (G, E)
    k_hit, tof = t.k_hit, t.tof
    # invariants: tof == 0 exactly where not hit, plausible where hit, k_hit in range
    assert bool(((k_hit >= -1) & (k_hit < len(alpha))).all())
    assert bool(((tof == 0) == (k_hit < 0)).all())
    hit_t = tof[k_hit >= 0]
    assert float(hit_t.min()) > 3e-5 and float(hit_t.max()) < 3e-4
    assert bool((k_hit[:, 64] >= 0).any())
    # random subsample vs the bit-exact twin
    rng = np.random.default_rng(0)
    sel = np.sort(rng.choice(G, 512, replace=False))
    p = oracle_dm.params()
    tof_o, k_o, _ = oracle_dm.table(p, radii[sel], offsets[sel], x_a[64], z_a[64], zf, alpha, x_a)
    sel_d = torch.from_numpy(sel).cuda()
    assert np.array_equal(k_hit[sel_d].cpu().numpy(), k_o.T)
    assert np.array_equal(tof[sel_d].cpu().numpy(), tof_o.T)
    # a contiguous slice recomputed on its own gives identical bytes (what sharding relies on)
    lo, hi = 300_000, 300_000 + 4096
    t2 = engine.tof_table_cuda(r_d[lo:hi].contiguous(), o_d[lo:hi].contiguous())
    torch.cuda.synchronize()
    assert torch.equal(t2.tof, tof[lo:hi]) and torch.equal(t2.k_hit, k_hit[lo:hi])


def _table_all(engine, radii, offsets):
    t = engine.tof_table(radii, offsets, records=("idx", "target_x", "lens2_x", "lens2_z"))
    return t


@pytest.mark.parametrize("case", ["sweep", "random", "grazing", "compare", "convex"])
def test_guided_search_equals_literal_scan(engine, oracle_dm, case):
    """TOFRT_SEARCH_GUIDED must return exactly what the reference's O(N) scan returns."""
    x_a, z_a = rs.transducer()
    kw = {}
    n_alpha = rs.NUM_ALPHA_POINTS_RT
    rng = np.random.default_rng(11)
    if case == "sweep":
        radii, offsets = rs.sweep_grid()  # includes the 12 degenerate geometries (line through a sample)
    elif case == "random":
        radii, offsets = rng.uniform(0.003, 0.15, 2000), rng.uniform(-0.03, 0.03, 2000)
    elif case == "grazing":
        # pipes that reach into / touch the lens: tangent and multiply-crossing return paths
        radii = rng.uniform(0.07, 0.12, 1500)
        offsets = rng.uniform(-0.05, 0.05, 1500)
    elif case == "compare":
        n_alpha = rs.NUM_ALPHA_POINTS_COMPARE
        radii, offsets = np.array([0.037, 0.05, 0.02]), np.array([0.0038, 0.0, -0.009])
    else:  # a parameter set (unphysical: l0 < 0) whose sampled surface is a CONVEX polyline
        kw = dict(c1=1986.5, c2=3134.7, l0=-0.20058, h0=0.04446)
        radii, offsets = rng.uniform(0.01, 0.12, 500), rng.uniform(-0.01, 0.01, 500)
    alpha, zf = rs.launch_angles(n_alpha), rs.focus_line(n_alpha)
    z_sh = z_a[32]
    if case == "convex":
        alpha = np.linspace(-0.3326, 0.3326, 400)
        z_sh = float(np.float64(kw["l0"]) + np.float64(kw["h0"]))
        zf = np.ones(400) * z_sh
    engine.setup(alpha, zf, x_a[32], z_sh, x_a, **kw)
    assert engine.lens_shape == (-1 if case == "convex" else 1)
    engine.set_search("scan")
    a = _table_all(engine, radii, offsets)
    engine.set_search("guided")
    b = _table_all(engine, radii, offsets)
    assert np.array_equal(a.records["idx"], b.records["idx"])
    assert np.array_equal(a.k_hit, b.k_hit) and np.array_equal(a.tof, b.tof)
    for key in ("target_x", "lens2_x", "lens2_z"):
        assert _same(a.records[key], b.records[key]), key
    # and both equal the oracle twin
    p = oracle_dm.params(**kw)
    tof, k_hit, idx = oracle_dm.table(p, radii, offsets, x_a[32], z_sh, zf, alpha, x_a, want_idx=True)
    assert np.array_equal(b.records["idx"], idx)
    assert np.array_equal(b.k_hit, k_hit.T) and np.array_equal(b.tof, tof.T)
    assert (idx >= 0).mean() > 0.2  # the case really exercises the search


def test_non_monotone_lens_disables_guided_search(engine, oracle_dm):
    """A launch-angle list that is not increasing gives a lens polyline that is neither
    concave nor convex: the engine must notice and use the literal scan."""
    x_a, z_a = rs.transducer()
    alpha = rs.launch_angles(200).copy()
    rng = np.random.default_rng(3)
    rng.shuffle(alpha)
    zf = rs.focus_line(200)
    engine.setup(alpha, zf, x_a[32], z_a[32], x_a)
    assert engine.lens_shape == 0
    radii, offsets = rng.uniform(0.01, 0.1, 50), rng.uniform(-0.01, 0.01, 50)
    t = engine.tof_table(radii, offsets, records=("idx",))
    p = oracle_dm.params()
    tof, k_hit, idx = oracle_dm.table(p, radii, offsets, x_a[32], z_a[32], zf, alpha, x_a, want_idx=True)
    assert np.array_equal(t.records["idx"], idx)
    assert np.array_equal(t.k_hit, k_hit.T) and np.array_equal(t.tof, tof.T)


def test_fp32_mode_is_close_to_fp64(engine):
    """FP32 mode carries no bit-parity claim; its measured error is reported (bench.py "fp32").
    Here: the same rays give the same travel time to ~1e-9 s, and hit decisions differ only for
    landing points within float rounding (~5e-7 m) of the 1e-6 m tolerance band."""
    import torch
    x_a, z_a, alpha, zf = _std(engine)
    radii, offsets = rs.synthetic_grid(96, 96)
    r_d, o_d = torch.from_numpy(radii).cuda(), torch.from_numpy(offsets).cuda()
    t64 = engine.tof_table_cuda(r_d, o_d)
    t32 = engine.tof_table_f32_cuda(r_d, o_d)
    torch.cuda.synchronize()
    assert t32.tof.dtype == torch.float32 and t32.tof.shape == t64.tof.shape
    h64, h32 = t64.k_hit >= 0, t32.k_hit >= 0
    assert int(h64.sum()) > 1000
    assert float((h64 != h32).float().mean()) < 5e-3          # measured ~1.3e-3 of all entries
    both = h64 & h32
    same_ray = both & (t64.k_hit == t32.k_hit)
    assert float(same_ray.sum()) / float(both.sum()) > 0.95
    dt = (t32.tof.double() - t64.tof)[same_ray].abs()
    assert float(dt.max()) < 5e-9 and float(dt.median()) < 5e-11  # measured 1.2e-9 / 5e-12
    assert bool(((t32.tof == 0) == (t32.k_hit < 0)).all())


def test_full_matrix_capture_every_element_shoots(engine, golden, oracle_dm):
    """S shooters per geometry in one call; each shooter's block equals the
    single-shooter call and the oracle; the live reference pins the shooters it can run."""
    x_a, z_a = rs.transducer()
    alpha, zf = rs.launch_angles(), rs.focus_line()
    p = oracle_dm.params()
    rng = np.random.default_rng(21)
    radii, offsets = rng.uniform(0.04, 0.1, 24), rng.uniform(-0.008, 0.008, 24)
    engine.setup(alpha, zf, x_a, z_a, x_a)  # every element shoots: S = E = 65
    t = engine.tof_table(radii, offsets, records=("idx",))
    assert t.tof.shape == (65, 24, 65) and t.records["idx"].shape == (65, 24, 905)
    for s in range(65):
        tof, k_hit, idx = oracle_dm.table(p, radii, offsets, x_a[s], z_a[s], zf, alpha, x_a, want_idx=True)
        assert np.array_equal(t.records["idx"][s], idx), s
        assert np.array_equal(t.k_hit[s], k_hit.T) and np.array_equal(t.tof[s], tof.T), s
    # device-resident API, same bytes
    import torch
    td = engine.tof_table_cuda(torch.from_numpy(radii).cuda(), torch.from_numpy(offsets).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(td.tof.cpu().numpy(), t.tof) and np.array_equal(td.k_hit.cpu().numpy(), t.k_hit)
    # the live reference (tests/golden/shooters.npz)
    sh = golden["shooters"]
    t2 = engine.tof_table(sh["radius"], sh["offset"])
    for i, s in enumerate(sh["shooter"]):
        assert np.array_equal(t2.k_hit[s, i], sh["k_hit"][i])
        assert np.abs(t2.tof[s, i] - sh["tof"][i]).max() <= TOF_TOL
    # a single shooter given as a scalar keeps the 2-D layout
    engine.setup(alpha, zf, x_a[7], z_a[7], x_a)
    t1 = engine.tof_table(radii, offsets)
    assert t1.tof.shape == (24, 65) and np.array_equal(t1.tof, t.tof[7]) and np.array_equal(t1.k_hit, t.k_hit[7])


@pytest.mark.parametrize("n_elements", [1024, 4096])
def test_many_receivers(engine, oracle_dm, n_elements):
    """More receivers than the per-warp shared-memory first-hit state can hold (BASELINE config 5 has
    1 025): the state moves into the output rows; results stay bit-identical to the oracle."""
    x_a, z_a = rs.transducer(num_elements=n_elements, pitch=np.float64(0.04 / n_elements))
    alpha, zf = rs.launch_angles(), rs.focus_line()
    engine.setup(alpha, zf, x_a[n_elements // 2], z_a[0], x_a, atol=1e-5)
    rng = np.random.default_rng(n_elements)
    radii, offsets = rng.uniform(0.01, 0.1, 40), rng.uniform(-0.01, 0.01, 40)
    p = oracle_dm.params()
    tof, k_hit, _ = oracle_dm.table(p, radii, offsets, x_a[n_elements // 2], z_a[0], zf, alpha, x_a, atol=1e-5)
    for mode in ("guided", "scan"):
        engine.set_search(mode)
        t = engine.tof_table(radii, offsets)
        assert t.tof.shape == (40, n_elements + 1)
        assert int(t.hit.sum()) > 500
        assert np.array_equal(t.k_hit, k_hit.T) and np.array_equal(t.tof, tof.T)
    engine.set_search("guided")
    # the hit list of such a setup takes the dense route (no per-row slots with the state in global memory)
    g, e, k, th = engine.tof_hits(radii, offsets)
    gg, ee = np.nonzero(t.hit)
    assert np.array_equal(g, gg) and np.array_equal(e, ee) and np.array_equal(k, t.k_hit[gg, ee])
    assert np.array_equal(th, t.tof[gg, ee])
    from ray_tracing_ultrasound_b200 import TofrtError
    import torch
    if n_elements >= 4096:
        with pytest.raises(TofrtError):
            engine.tof_table_f32_cuda(torch.from_numpy(radii).cuda(), torch.from_numpy(offsets).cuda())


def test_sparse_hit_list_equals_dense_table(engine):
    """tofrt_hits_host returns exactly the dense table's hits, ordered by (geometry, receiver)."""
    x_a, z_a, alpha, zf = _std(engine)
    radii, offsets = rs.synthetic_grid(60, 50)
    t = engine.tof_table(radii, offsets)
    g, e, k, tof = engine.tof_hits(radii, offsets)
    gg, ee = np.nonzero(t.hit)                     # row-major order = (geometry, receiver)
    assert len(g) == len(gg) > 1000
    assert np.array_equal(g, gg) and np.array_equal(e, ee)
    assert np.array_equal(k, t.k_hit[gg, ee]) and np.array_equal(tof, t.tof[gg, ee])
    # capacity too small: first `capacity` hits, total reported, then an exact-size retry inside tof_hits
    g2, e2, k2, tof2 = engine.tof_hits(radii, offsets, capacity=100)
    assert np.array_equal(g2, gg) and np.array_equal(tof2, tof)
    # several shooters: row = shooter * G + geometry
    engine.setup(alpha, zf, x_a[30:34], z_a[30:34], x_a)
    t4 = engine.tof_table(radii[:200], offsets[:200])
    g4, e4, k4, tof4 = engine.tof_hits(radii[:200], offsets[:200])
    ss, gg4, ee4 = np.nonzero(t4.hit)
    assert np.array_equal(g4, ss * 200 + gg4) and np.array_equal(e4, ee4)
    assert np.array_equal(tof4, t4.tof[ss, gg4, ee4])
    # no geometry at all
    assert all(len(a) == 0 for a in engine.tof_hits([], []))


def test_sparse_hit_list_pipeline_many_chunks_and_slot_overflow(engine, monkeypatch):
    """tofrt_hits_host's chunked pipeline: 11 chunks through the ring of three buffers with a ragged last chunk,
    several shooters across chunk boundaries, a capacity smaller than the result -- and the fallback to the dense route
    when a row holds more hits than its 16-entry slot (a tolerance wide enough that a ray lands on many receivers)."""
    monkeypatch.setenv("TOFRT_HIT_CHUNK_ROWS", "1024")
    x_a, z_a, alpha, zf = _std(engine)
    radii, offsets = rs.synthetic_grid(131, 83)          # 10 873 rows = 10 chunks of 1 024 + 633

    def check(t, hits, n_geoms):
        g, e, k, tof = hits
        nz = np.nonzero(t.hit)
        rows = nz[0] if t.hit.ndim == 2 else nz[0] * n_geoms + nz[1]
        assert len(g) == len(rows) > 1000
        assert np.array_equal(g, rows) and np.array_equal(e, nz[-1])
        assert np.array_equal(k, t.k_hit[nz]) and np.array_equal(tof, t.tof[nz])

    t = engine.tof_table(radii, offsets)
    check(t, engine.tof_hits(radii, offsets), len(radii))
    check(t, engine.tof_hits(radii, offsets, capacity=777), len(radii))
    engine.setup(alpha, zf, x_a[31:34], z_a[31:34], x_a)  # rows = 3 x 4 000: shooters change inside a chunk
    check(engine.tof_table(radii[:4000], offsets[:4000]), engine.tof_hits(radii[:4000], offsets[:4000]), 4000)
    _std(engine, atol=3e-3)                               # +-3 mm: ~20 of the 65 receivers match one landing point
    t = engine.tof_table(radii[:3000], offsets[:3000])
    assert int(t.hit.sum(axis=1).max()) > 16
    check(t, engine.tof_hits(radii[:3000], offsets[:3000]), 3000)


def test_host_pipeline_many_chunks_with_records(engine):
    """The host path cuts the geometries into ~64 MB chunks (double-buffered); with per-ray records a
    chunk is only ~1 600 geometries, so 5 000 geometries exercise several chunks, the slot reuse and
    a ragged last chunk.  Must equal the single-launch device path byte for byte."""
    import torch
    x_a, z_a, alpha, zf = _std(engine)
    rng = np.random.default_rng(77)
    G = 5003
    radii, offsets = rng.uniform(0.01, 0.1, G), rng.uniform(-0.01, 0.01, G)
    keys = ("idx", "target_x", "legs", "pipe_z")
    th = engine.tof_table(radii, offsets, records=keys)
    td = engine.tof_table_cuda(torch.from_numpy(radii).cuda(), torch.from_numpy(offsets).cuda(), records=keys)
    torch.cuda.synchronize()
    assert np.array_equal(th.tof, td.tof.cpu().numpy()) and np.array_equal(th.k_hit, td.k_hit.cpu().numpy())
    for k in keys:
        assert _same(th.records[k], td.records[k].cpu().numpy()), k
    assert th.records["legs"].shape == (G, 4, 905)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_other_material_and_lens_parameters(engine, oracle_dm, seed):
    """Nothing is specialised to the reference's constants (main_rt.py:449-460): random wave speeds,
    lens dimensions, apertures, ray counts and tolerances stay bit-identical to the oracle."""
    rng = np.random.default_rng(1000 + seed)
    kw = dict(c1=rng.uniform(3000, 7000), c2=rng.uniform(900, 2500),
              l0=rng.uniform(0.05, 0.2), h0=rng.uniform(0.04, 0.12))
    K = int(rng.integers(100, 1500))
    n_el = int(rng.choice([16, 32, 96]))
    d = float(np.float64(kw["l0"]) + np.float64(kw["h0"]))
    x_a, _ = rs.transducer(num_elements=n_el, pitch=np.float64(rng.uniform(2e-4, 1e-3)), d=d)
    alpha = np.linspace(-rng.uniform(0.3, 0.8), rng.uniform(0.3, 0.8), K)
    zf = np.ones(K) * d + rng.uniform(-1e-3, 1e-3, K)          # a ragged focus line: z_f is an array (main_rt.py:404)
    shooter = int(rng.integers(0, len(x_a)))
    rtol, atol = 10.0 ** rng.uniform(-7, -4), 10.0 ** rng.uniform(-7, -4)
    engine.setup(alpha, zf, x_a[shooter], d, x_a, rtol=rtol, atol=atol, **kw)
    G = 300
    radii = rng.uniform(0.1, 0.9, G) * kw["h0"]
    offsets = rng.uniform(-0.15, 0.15, G) * kw["h0"]
    t = engine.tof_table(radii, offsets, records=("idx", "target_x"))
    p = oracle_dm.params(**kw)
    tof, k_hit, idx = oracle_dm.table(p, radii, offsets, x_a[shooter], d, zf, alpha, x_a, rtol=rtol, atol=atol,
                                      want_idx=True)
    assert np.array_equal(t.records["idx"], idx)
    assert np.array_equal(t.k_hit, k_hit.T) and np.array_equal(t.tof, tof.T)
    assert (idx >= 0).mean() > 0.3


def test_pinned_output_buffers(engine):
    """tof_table(out=...) with page-locked arrays from tofrt_host_alloc gives the same bytes."""
    from ray_tracing_ultrasound_b200 import pinned_empty
    x_a, z_a, alpha, zf = _std(engine)
    radii, offsets = rs.synthetic_grid(64, 64)
    t = engine.tof_table(radii, offsets)
    tof, kh = pinned_empty((len(radii), 65)), pinned_empty((len(radii), 65), np.int32)
    t2 = engine.tof_table(radii, offsets, out=(tof, kh))
    assert t2.tof is tof and np.array_equal(tof, t.tof) and np.array_equal(kh, t.k_hit)
    with pytest.raises(ValueError):
        engine.tof_table(radii, offsets, out=(tof[:10], kh[:10]))
    del tof, kh, t2
