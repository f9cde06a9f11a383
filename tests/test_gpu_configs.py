"""GPU: every BASELINE.json config with a 2-D analogue, at full size, against the
INDEPENDENT libm oracle (reference operation order, glibc elementary functions) and the live-reference goldens.

Contract (north_star): k_hit identical; |dt| <= 1e-12 s; the discrete return-path index idx identical for every ray
whose decision is not rounding noise -- min(|diffs[idx]|, |diffs[idx+1]|) > 1e-13 m in the oracle's own arithmetic --
with the number of excluded rays asserted small.  Both math modes (TOFRT_OPT_MATH) run every test.
"""
import numpy as np
import pytest

from ray_tracing_ultrasound_b200 import reference_setup as rs

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("math_mode")]

TOF_TOL = 1e-12
NOISE_FLOOR = 1e-13


def assert_idx_agrees(idx_gpu, idx_ref, margin_ref, max_noise_frac=2e-3, what=""):
    """idx equal wherever the oracle's own decision exceeds the noise floor; returns the number of differing rays."""
    differs = idx_gpu != idx_ref
    decided = margin_ref > NOISE_FLOOR          # NaN margins (NaN lines) compare False: both sides give idx 0 there
    bad = differs & decided
    assert not bad.any(), (what, int(bad.sum()), np.argwhere(bad)[:5].tolist())
    n = int(differs.sum())
    assert n <= max_noise_frac * idx_ref.size, (what, n, idx_ref.size)
    return n


def _setup(engine, n_elements, K=rs.NUM_ALPHA_POINTS_RT):
    x_a, z_a = rs.transducer(num_elements=n_elements)
    alpha, zf = rs.launch_angles(K), rs.focus_line(K)
    c = n_elements // 2
    engine.setup(alpha, zf, x_a[c], z_a[c], x_a)
    return x_a, z_a, alpha, zf, c


def _subsample_vs_libm(engine, oracle_libm, x_a, z_a, alpha, zf, c, radii, offsets, k_hit_d, tof_d, sel, what):
    """Rows `sel` of a device-resident table against the libm oracle + idx of the same geometries."""
    import torch
    p = oracle_libm.params()
    tof_o, k_o, idx_o, margin = oracle_libm.table(p, radii[sel], offsets[sel], x_a[c], z_a[c], zf, alpha, x_a,
                                                  want_idx=True, want_margin=True)
    sel_d = torch.from_numpy(sel).cuda()
    k_g, t_g = k_hit_d[sel_d].cpu().numpy(), tof_d[sel_d].cpu().numpy()
    assert np.array_equal(k_g, k_o.T), what
    assert np.abs(t_g - tof_o.T).max() <= TOF_TOL, what
    t = engine.tof_table(radii[sel], offsets[sel], records=("idx",))
    assert np.array_equal(t.k_hit, k_g) and np.array_equal(t.tof, t_g), what   # slice recomputed alone: same bytes
    return assert_idx_agrees(t.records["idx"], idx_o, margin, what=what)


def test_config1_fixture_sweep_idx_vs_libm_oracle(engine, oracle_libm):
    """BASELINE config 1 (main_rt.py default sweep): idx against the independent oracle, ray by ray."""
    x_a, z_a, alpha, zf, c = _setup(engine, 64)
    radii, offsets = rs.sweep_grid()
    t = engine.tof_table(radii, offsets, records=("idx",))
    p = oracle_libm.params()
    tof, k_hit, idx, margin = oracle_libm.table(p, radii, offsets, x_a[c], z_a[c], zf, alpha, x_a, want_idx=True,
                                                want_margin=True)
    assert np.array_equal(t.k_hit, k_hit.T) and np.abs(t.tof - tof.T).max() <= TOF_TOL
    degenerate = (offsets == 0) | ((radii == 0.01) & (np.abs(offsets) == 0.01))
    n = assert_idx_agrees(t.records["idx"], idx, margin, max_noise_frac=2e-2, what="config 1")
    # outside the 12 degenerate geometries (the ray retraces its own path through a lens sample) nothing differs at all
    assert np.array_equal(t.records["idx"][~degenerate], idx[~degenerate])
    print(f"config 1: {n} of {idx.size} rays differ from the libm oracle, all inside the noise floor")


def test_config2_grid_512_fp64_and_fp32(engine, oracle_libm):
    """BASELINE config 2: 64-element array (E = 65), 512 x 512 geometries, FP64 table and FP32 mode."""
    import torch
    x_a, z_a, alpha, zf, c = _setup(engine, 64)
    radii, offsets = rs.synthetic_grid(512, 512)
    r_d, o_d = torch.from_numpy(radii).cuda(), torch.from_numpy(offsets).cuda()
    t64 = engine.tof_table_cuda(r_d, o_d)
    torch.cuda.synchronize()
    G = len(radii)
    assert t64.tof.shape == (G, 65)
    assert bool(((t64.tof == 0) == (t64.k_hit < 0)).all())
    rng = np.random.default_rng(2)
    sel = np.sort(rng.choice(G, 2048, replace=False))
    n = _subsample_vs_libm(engine, oracle_libm, x_a, z_a, alpha, zf, c, radii, offsets, t64.k_hit, t64.tof, sel, "config 2")
    print(f"config 2: {n} of {2048 * 905} subsampled rays differ in idx (noise floor)")
    # FP32 mode: reported, not gated on parity -- bounds a little above the measured figures (bench.py "fp32")
    t32 = engine.tof_table_f32_cuda(r_d, o_d)
    torch.cuda.synchronize()
    h64, h32 = t64.k_hit >= 0, t32.k_hit >= 0
    both = h64 & h32
    same_ray = both & (t64.k_hit == t32.k_hit)
    assert float((h64 != h32).float().mean()) < 5e-3
    assert float(same_ray.sum()) / float(both.sum()) > 0.95
    dt = (t32.tof.double() - t64.tof)[same_ray].abs()
    assert float(dt.max()) < 5e-9 and float(dt.median()) < 5e-11


def test_config3_array128_vs_live_reference_golden(engine, golden_array128, oracle_libm):
    """BASELINE configs 3/4 use a 128-element array (E = 129).  tests/golden/array128.npz holds the UNMODIFIED
    reference (main_rt.py with num_elements = 128) on 96 geometries of the 1024 x 1024 grid."""
    gd = golden_array128
    x_a, z_a, alpha, zf, c = _setup(engine, 128)
    assert np.array_equal(x_a, gd["x_a"]) and np.array_equal(alpha, gd["alpha"]) and c == 64
    radii, offsets = rs.synthetic_grid(1024, 1024)
    flat = gd["flat_index"]
    assert np.array_equal(radii[flat], gd["radius"]) and np.array_equal(offsets[flat], gd["offset"])
    t = engine.tof_table(radii[flat], offsets[flat], records=("idx", "target_x"))
    assert np.array_equal(t.k_hit, gd["k_hit"]) and np.array_equal(t.hit, gd["hit"])
    assert np.abs(t.tof - gd["tof"]).max() <= TOF_TOL
    assert int(t.hit.sum()) > 200
    nan_ref = np.unpackbits(gd["nan_target"], axis=1)[:, :905].astype(bool)
    assert np.array_equal(np.isnan(t.records["target_x"]), nan_ref)
    ridx = gd["idx"].astype(np.int32)                     # reference encoding: -1 no sign change, -2 NaN line
    gi = np.where(t.records["idx"] < -1, -1, t.records["idx"])
    differs = (gi != ridx) & (ridx != -2)
    assert not (differs & (gd["margin"] > NOISE_FLOOR)).any()
    assert differs.sum() <= 2e-3 * ridx.size
    # and the independent oracle on the same geometries
    p = oracle_libm.params()
    tof, k_hit, idx, margin = oracle_libm.table(p, radii[flat], offsets[flat], x_a[c], z_a[c], zf, alpha, x_a,
                                                want_idx=True, want_margin=True)
    assert np.array_equal(t.k_hit, k_hit.T)
    assert_idx_agrees(t.records["idx"], idx, margin, what="config 3 golden geometries")


def test_config3_grid_1024_vs_libm_oracle(engine, oracle_libm):
    """BASELINE config 3 at full size: a 2 048-geometry random subsample against the independent oracle."""
    import torch
    x_a, z_a, alpha, zf, c = _setup(engine, 128)
    radii, offsets = rs.synthetic_grid(1024, 1024)
    r_d, o_d = torch.from_numpy(radii).cuda(), torch.from_numpy(offsets).cuda()
    t = engine.tof_table_cuda(r_d, o_d)
    torch.cuda.synchronize()
    rng = np.random.default_rng(3)
    sel = np.sort(rng.choice(len(radii), 2048, replace=False))
    n = _subsample_vs_libm(engine, oracle_libm, x_a, z_a, alpha, zf, c, radii, offsets, t.k_hit, t.tof, sel, "config 3")
    print(f"config 3: {n} of {2048 * 905} subsampled rays differ in idx (noise floor)")


def test_config4_grid_4096_on_one_gpu(engine, oracle_libm):
    """BASELINE config 4 (129 x 4096 x 4096, 2.16e9 table entries, 26 GB) on ONE GPU: invariants, a random subsample
    against the independent oracle, and shard equivalence -- every 8th-of-the-grid slice a rank would own under
    either sharding rule (contiguous or round-robin rows of the radius axis) recomputed alone gives the same bytes."""
    import torch
    from ray_tracing_ultrasound_b200 import sharding
    x_a, z_a, alpha, zf, c = _setup(engine, 128)
    radii, offsets = rs.synthetic_grid(4096, 4096)
    G, E = len(radii), len(x_a)
    r_d, o_d = torch.from_numpy(radii).cuda(), torch.from_numpy(offsets).cuda()
    tof = torch.empty((G, E), dtype=torch.float64, device="cuda")
    k_hit = torch.empty((G, E), dtype=torch.int32, device="cuda")
    engine.tof_table_cuda(r_d, o_d, out_tof=tof, out_k_hit=k_hit)
    torch.cuda.synchronize()
    for lo in range(0, G, G // 8):       # invariants, an eighth at a time (keeps the temporaries small)
        kh, tt = k_hit[lo:lo + G // 8], tof[lo:lo + G // 8]
        assert bool(((kh >= -1) & (kh < len(alpha))).all())
        assert bool(((tt == 0) == (kh < 0)).all())
        assert bool((kh[:, 64] >= 0).any())
    rng = np.random.default_rng(4)
    sel = np.sort(rng.choice(G, 2048, replace=False))
    n = _subsample_vs_libm(engine, oracle_libm, x_a, z_a, alpha, zf, c, radii, offsets, k_hit, tof, sel, "config 4")
    print(f"config 4: {n} of {2048 * 905} subsampled rays differ in idx (noise floor)")
    # shard equivalence on a 64 x 64 sub-grid: rank 5 of 8 under the interleaved rule
    rows = sharding.interleaved_rows(4096, 8, 5)[:64]                     # radius rows this rank owns
    cols = np.arange(0, 4096, 64)
    idx = (rows[:, None] * 4096 + cols[None, :]).ravel()
    idx_d = torch.from_numpy(idx).cuda()
    t2 = engine.tof_table_cuda(r_d[idx_d].contiguous(), o_d[idx_d].contiguous())
    torch.cuda.synchronize()
    assert torch.equal(t2.tof, tof[idx_d]) and torch.equal(t2.k_hit, k_hit[idx_d])
    del tof, k_hit
    torch.cuda.empty_cache()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_geometries_and_materials_idx_vs_libm_oracle(engine, oracle_libm, seed):
    """Random geometries on the reference's constants (seed 0) and random materials / lenses / arrays (seeds 1, 2):
    idx against the independent oracle with the per-ray margin rule."""
    rng = np.random.default_rng(500 + seed)
    if seed == 0:
        kw, K, n_el = {}, 905, 64
        d = float(rs.D)
        x_a, _ = rs.transducer()
        alpha = rs.launch_angles()
        radii, offsets = rng.uniform(0.005, 0.12, 1500), rng.uniform(-0.02, 0.02, 1500)
        rtol, atol = 1e-5, 1e-6
    else:
        kw = dict(c1=rng.uniform(3000, 7000), c2=rng.uniform(900, 2500), l0=rng.uniform(0.05, 0.2), h0=rng.uniform(0.04, 0.12))
        K = int(rng.integers(100, 1200))
        d = float(np.float64(kw["l0"]) + np.float64(kw["h0"]))
        x_a, _ = rs.transducer(num_elements=int(rng.choice([16, 64, 128])), pitch=np.float64(rng.uniform(2e-4, 8e-4)), d=d)
        alpha = np.linspace(-rng.uniform(0.3, 0.85), rng.uniform(0.3, 0.85), K)
        radii, offsets = rng.uniform(0.05, 1.0, 600) * kw["h0"], rng.uniform(-0.25, 0.25, 600) * kw["h0"]
        rtol, atol = 10.0 ** rng.uniform(-7, -4), 10.0 ** rng.uniform(-6, -4)
    zf = np.ones(len(alpha)) * d
    sh = int(rng.integers(0, len(x_a)))
    engine.setup(alpha, zf, x_a[sh], d, x_a, rtol=rtol, atol=atol, **kw)
    t = engine.tof_table(radii, offsets, records=("idx",))
    p = oracle_libm.params(**kw)
    tof, k_hit, idx, margin = oracle_libm.table(p, radii, offsets, x_a[sh], d, zf, alpha, x_a, rtol=rtol, atol=atol,
                                                want_idx=True, want_margin=True)
    assert np.array_equal(t.k_hit, k_hit.T)
    hit = t.k_hit >= 0
    if hit.any():
        assert np.abs(t.tof - tof.T)[hit].max() <= TOF_TOL
    n = assert_idx_agrees(t.records["idx"], idx, margin, what=f"seed {seed}")
    print(f"seed {seed}: {n} of {idx.size} rays differ in idx (noise floor)")
