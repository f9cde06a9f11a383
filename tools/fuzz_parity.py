"""Large differential runs (under gpurun), results summarised in profiles/fuzz_r1.json:
 (1) GPU (guided) vs the CPU twin oracle, bit for bit, on random geometries over wide ranges;
 (2) GPU guided vs GPU literal scan on millions of geometries (every output compared on device)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ray_tracing_ultrasound_b200 import TofEngine, reference_setup as rs
from oracle.oracle import Oracle

out = {}
x_a, z_a = rs.transducer(num_elements=128)
alpha, zf = rs.launch_angles(), rs.focus_line()
eng = TofEngine(0).setup(alpha, zf, x_a[64], z_a[64], x_a)
o = Oracle(device_math=True); p = o.params()
o.set_math(eng.math_mode == "vector")   # the twin follows the engine's TOFRT_OPT_MATH mode (default: vector)
rng = np.random.default_rng(2024)

# (1) vs the twin, including idx of every ray
n1 = int(os.environ.get("FUZZ_ORACLE_GEOMS", "60000"))
bad = 0; rays = 0; t0 = time.time()
for lo in range(0, n1, 10000):
    n = min(10000, n1 - lo)
    radii = rng.uniform(0.002, 0.15, n); offsets = rng.uniform(-0.04, 0.04, n)
    t = eng.tof_table(radii, offsets, records=("idx",))
    tof, k_hit, idx = o.table(p, radii, offsets, x_a[64], z_a[64], zf, alpha, x_a, want_idx=True)
    bad += int((t.records["idx"] != idx).sum()) + int((t.k_hit != k_hit.T).sum()) + int((t.tof != tof.T).sum())
    rays += n * len(alpha)
out["gpu_vs_twin_oracle"] = {"geometries": n1, "rays": rays, "table_entries": n1 * len(x_a), "mismatches": bad,
                             "ranges": "r in [0.002, 0.15], offset in [-0.04, 0.04], E=129, K=N=905",
                             "seconds": round(time.time() - t0, 1)}

# (2) guided vs scan on device
n2 = int(os.environ.get("FUZZ_MODE_GEOMS", "4000000"))
mism = 0; fallback = 0; t0 = time.time()
for lo in range(0, n2, 1000000):
    n = min(1000000, n2 - lo)
    r = torch.from_numpy(rng.uniform(0.002, 0.15, n)).cuda(); of = torch.from_numpy(rng.uniform(-0.04, 0.04, n)).cuda()
    eng.set_search("guided"); eng.set_stats(True); a = eng.tof_table_cuda(r, of); torch.cuda.synchronize(); fallback += eng.last_stats()["scan_rays"]; eng.set_stats(False)
    eng.set_search("scan"); b = eng.tof_table_cuda(r, of); torch.cuda.synchronize()
    mism += int((a.tof != b.tof).sum()) + int((a.k_hit != b.k_hit).sum())
    del a, b
out["guided_vs_scan"] = {"geometries": n2, "rays": n2 * len(alpha), "table_entries": n2 * len(x_a), "mismatches": mism,
                         "rays_that_fell_back_to_the_scan": int(fallback), "seconds": round(time.time() - t0, 1)}
# (2b) with per-ray idx on a smaller batch
n3 = 200000
r = torch.from_numpy(rng.uniform(0.002, 0.15, n3)).cuda(); of = torch.from_numpy(rng.uniform(-0.04, 0.04, n3)).cuda()
eng.set_search("guided"); a = eng.tof_table_cuda(r, of, records=("idx", "target_x"))
eng.set_search("scan"); b = eng.tof_table_cuda(r, of, records=("idx", "target_x"))
torch.cuda.synchronize()
ta, tb = a.records["target_x"], b.records["target_x"]
out["guided_vs_scan_per_ray"] = {"geometries": n3, "rays": n3 * len(alpha),
                                 "idx_mismatches": int((a.records["idx"] != b.records["idx"]).sum()),
                                 "target_x_mismatches": int(((ta != tb) & ~(ta.isnan() & tb.isnan())).sum())}
# (3) random setups: material speeds, lens dimensions, ray counts, arrays, shooters, tolerances
n4 = int(os.environ.get("FUZZ_SETUPS", "24"))
per = int(os.environ.get("FUZZ_SETUP_GEOMS", "20000"))
bad = 0; rays = 0; t0 = time.time()
for i in range(n4):
    r4 = np.random.default_rng(5000 + i)
    kw = dict(c1=r4.uniform(3000, 7000), c2=r4.uniform(900, 2500), l0=r4.uniform(0.05, 0.2), h0=r4.uniform(0.04, 0.12))
    K = int(r4.integers(64, 2000)); nel = int(r4.choice([16, 64, 128, 256]))
    d = float(np.float64(kw["l0"]) + np.float64(kw["h0"]))
    xa, _ = rs.transducer(num_elements=nel, pitch=np.float64(r4.uniform(2e-4, 8e-4)), d=d)
    al = np.linspace(-r4.uniform(0.3, 0.85), r4.uniform(0.3, 0.85), K)
    zf4 = np.ones(K) * d
    sh = int(r4.integers(0, len(xa)))
    rtol, atol = 10.0 ** r4.uniform(-7, -4), 10.0 ** r4.uniform(-7, -4)
    eng.setup(al, zf4, xa[sh], d, xa, rtol=rtol, atol=atol, **kw)
    eng.set_search("guided" if i % 3 else "scan")
    eng.set_math("vector" if i % 2 else "angle"); o.set_math(bool(i % 2))
    radii = r4.uniform(0.05, 1.1, per) * kw["h0"]; offsets = r4.uniform(-0.3, 0.3, per) * kw["h0"]
    t = eng.tof_table(radii, offsets, records=("idx",))
    p4 = o.params(**kw)
    tof, k_hit, idx = o.table(p4, radii, offsets, xa[sh], d, zf4, al, xa, rtol=rtol, atol=atol, want_idx=True)
    bad += int((t.records["idx"] != idx).sum()) + int((t.k_hit != k_hit.T).sum()) + int((t.tof != tof.T).sum())
    rays += per * K
out["random_setups_vs_twin_oracle"] = {"setups": n4, "geometries_per_setup": per, "rays": rays, "mismatches": bad,
                                       "varied": "c1, c2, l0, h0, K in [64, 2000), 16-256 elements, pitch, aperture, shooter, rtol/atol, search mode",
                                       "seconds": round(time.time() - t0, 1)}
eng.set_math("vector"); o.set_math(True)
# (3b) the two math modes against each other on device: hit decisions identical, times within 1e-15 s
n6 = int(os.environ.get("FUZZ_MATH_GEOMS", "4000000"))
eng.setup(alpha, zf, x_a[64], z_a[64], x_a); eng.set_search("guided")
kdiff = 0; dtmax = 0.0; t0 = time.time()
for lo in range(0, n6, 1000000):
    n = min(1000000, n6 - lo)
    r = torch.from_numpy(rng.uniform(0.002, 0.15, n)).cuda(); of = torch.from_numpy(rng.uniform(-0.04, 0.04, n)).cuda()
    eng.set_math("angle"); a = eng.tof_table_cuda(r, of); torch.cuda.synchronize()
    eng.set_math("vector"); b = eng.tof_table_cuda(r, of); torch.cuda.synchronize()
    kdiff += int((a.k_hit != b.k_hit).sum())
    both = (a.k_hit >= 0) & (b.k_hit >= 0)
    dtmax = max(dtmax, float((a.tof - b.tof)[both].abs().max()))
    del a, b
out["vector_vs_angle_math"] = {"geometries": n6, "table_entries": n6 * len(x_a), "k_hit_mismatches": kdiff,
                               "max_abs_dt_s": dtmax, "seconds": round(time.time() - t0, 1)}
# (4) 3-D two-point solve (extension) vs its CPU twin: bit-identical wherever the start sample agrees
x65, z65 = rs.transducer()
eng.setup(rs.launch_angles(), rs.focus_line(), x65[32], z65[32], x65)
n5 = int(os.environ.get("FUZZ_3D_VOXELS", "40000"))
r5 = np.random.default_rng(77)
gm = (np.arange(32) - 15.5) * float(rs.PITCH)
ex, ey = (v.ravel()[::4] for v in np.meshgrid(gm, gm, indexing="ij"))      # 256 of the 32 x 32 elements
eng.setup_fermat3d(ex, ey)
fx, fy, fz = r5.uniform(-0.03, 0.03, n5), r5.uniform(-0.03, 0.03, n5), r5.uniform(0.0, 0.05, n5)
t0 = time.time()
tg, kg, abg = eng.fermat3d_tof(fx, fy, fz, with_angles=True)
tt, kt, abt, ev, gap = o.fermat3d(o.params(), eng.f3_ang, ex, ey, float(rs.D), fx, fy, fz)
same = kg == kt
out["fermat3d_vs_twin"] = {"pairs": int(tg.size), "same_start_sample_frac": float(same.mean()),
                           "bit_identical_where_same_start": bool((tg[same] == tt[same]).all() and (abg[same] == abt[same]).all()),
                           "max_abs_dt_other_start_s": float(np.abs(tg - tt)[~same].max(initial=0.0)),
                           "max_grid_gap_other_start_s": float(gap[~same].max(initial=0.0)),
                           "max_newton_evaluations": int(ev.max()), "seconds": round(time.time() - t0, 1)}
# (5) the fused hit emission (tofrt_hits_host, pipelined chunks) against the dense table on random geometries
eng.setup(alpha, zf, x_a[64], z_a[64], x_a); eng.set_search("guided"); eng.set_math("vector")
n7 = int(os.environ.get("FUZZ_HITS_GEOMS", "1000000"))
r7, o7 = rng.uniform(0.002, 0.15, n7), rng.uniform(-0.04, 0.04, n7)
t0 = time.time()
dense = eng.tof_table_cuda(torch.from_numpy(r7).cuda(), torch.from_numpy(o7).cuda())
hg, he, hk, ht = eng.tof_hits(r7, o7)
gg, ee = torch.nonzero(dense.k_hit >= 0, as_tuple=True)
out["hit_list_vs_dense"] = {"geometries": n7, "hits": int(len(hg)),
                            "identical": bool(len(hg) == len(gg) and np.array_equal(hg, gg.cpu().numpy()) and
                                              np.array_equal(he, ee.cpu().numpy()) and
                                              np.array_equal(hk, dense.k_hit[gg, ee].cpu().numpy()) and
                                              np.array_equal(ht, dense.tof[gg, ee].cpu().numpy())),
                            "seconds": round(time.time() - t0, 1)}
del dense, gg, ee
# (6) small-G latency mode (CTA per geometry) against the regular kernel on many small random batches
n8 = int(os.environ.get("FUZZ_SMALLG_BATCHES", "400"))
bad = 0; tot = 0; t0 = time.time()
for i in range(n8):
    n = int(rng.integers(1, 400))
    r8 = torch.from_numpy(rng.uniform(0.002, 0.15, n)).cuda(); o8 = torch.from_numpy(rng.uniform(-0.04, 0.04, n)).cuda()
    eng.set_small_g(True); a = eng.tof_table_cuda(r8, o8, records=("idx", "target_x"))
    eng.set_small_g(False); b = eng.tof_table_cuda(r8, o8, records=("idx", "target_x"))
    torch.cuda.synchronize()
    ta, tb = a.records["target_x"], b.records["target_x"]
    bad += int((a.tof != b.tof).sum()) + int((a.k_hit != b.k_hit).sum()) + int((a.records["idx"] != b.records["idx"]).sum())
    bad += int(((ta != tb) & ~(ta.isnan() & tb.isnan())).sum())
    tot += n
eng.set_small_g(True)
out["small_g_vs_regular_kernel"] = {"batches": n8, "geometries": tot, "rays": tot * len(alpha), "mismatches": bad,
                                    "seconds": round(time.time() - t0, 1)}
# (7) random setups (as in (3)): small-G mode vs regular kernel and hit list vs dense table, random batch sizes,
#     one to three shooters
n9 = int(os.environ.get("FUZZ_MODE_SETUPS", "40"))
bad = 0; rows = 0; hits = 0; t0 = time.time()
for i in range(n9):
    r9 = np.random.default_rng(7000 + i)
    kw = dict(c1=r9.uniform(3000, 7000), c2=r9.uniform(900, 2500), l0=r9.uniform(0.05, 0.2), h0=r9.uniform(0.04, 0.12))
    K = int(r9.integers(64, 2000)); nel = int(r9.choice([16, 64, 128, 256]))
    d = float(np.float64(kw["l0"]) + np.float64(kw["h0"]))
    xa, _ = rs.transducer(num_elements=nel, pitch=np.float64(r9.uniform(2e-4, 8e-4)), d=d)
    al = np.linspace(-r9.uniform(0.3, 0.85), r9.uniform(0.3, 0.85), K)
    ns = int(r9.integers(1, 4)); sh = r9.integers(0, len(xa), ns)
    rtol, atol = 10.0 ** r9.uniform(-7, -4), 10.0 ** r9.uniform(-7, -3)
    eng.setup(al, np.ones(K) * d, xa[sh] if ns > 1 else xa[sh[0]], np.full(ns, d) if ns > 1 else d, xa, rtol=rtol, atol=atol, **kw)
    n = int(r9.integers(1, 3000)) if i % 2 else int(r9.integers(30000, 90000))
    rr, oo = r9.uniform(0.05, 1.1, n) * kw["h0"], r9.uniform(-0.3, 0.3, n) * kw["h0"]
    rd, od = torch.from_numpy(rr).cuda(), torch.from_numpy(oo).cuda()
    eng.set_small_g(True); a = eng.tof_table_cuda(rd, od)
    eng.set_small_g(False); b = eng.tof_table_cuda(rd, od)
    eng.set_small_g(True)
    torch.cuda.synchronize()
    bad += int((a.tof != b.tof).sum()) + int((a.k_hit != b.k_hit).sum())
    hg, he, hk, ht = eng.tof_hits(rr, oo)
    kh = a.k_hit.reshape(-1, len(xa)); tf = a.tof.reshape(-1, len(xa))
    gg, ee = torch.nonzero(kh >= 0, as_tuple=True)
    same = (len(hg) == len(gg) and np.array_equal(hg, gg.cpu().numpy()) and np.array_equal(he, ee.cpu().numpy()) and
            np.array_equal(hk, kh[gg, ee].cpu().numpy()) and np.array_equal(ht, tf[gg, ee].cpu().numpy()))
    bad += 0 if same else 1
    rows += n * ns; hits += len(hg)
    del a, b
out["random_setups_modes"] = {"setups": n9, "rows": rows, "hits": hits, "mismatches": bad,
                              "checked": "small-G mode = regular kernel; hit list = dense table; 1-3 shooters, batches of 1-3000 and 30000-90000 geometries",
                              "seconds": round(time.time() - t0, 1)}
eng.setup(alpha, zf, x_a[64], z_a[64], x_a)
print(json.dumps(out, indent=1))
