"""Run every BASELINE.json config that has a 2-D analogue (SURVEY.md section 8 d) on one GPU and
print one JSON object: ms per table, ray-solves/s, FP32 error.  `python tools/config_report.py [--big]`"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ray_tracing_ultrasound_b200 import TofEngine, reference_setup as rs


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / reps


def run(name, nel, K, radii, offsets, out, pitch=None, fp32=True):
    x_a, z_a = rs.transducer(num_elements=nel) if pitch is None else rs.transducer(num_elements=nel, pitch=np.float64(pitch))
    alpha, zf = rs.launch_angles(K), rs.focus_line(K)
    eng = TofEngine(0).setup(alpha, zf, x_a[nel // 2], z_a[nel // 2], x_a)
    r, o = torch.from_numpy(np.ascontiguousarray(radii)).cuda(), torch.from_numpy(np.ascontiguousarray(offsets)).cuda()
    G, E = len(radii), len(x_a)
    tof = torch.empty((G, E), dtype=torch.float64, device="cuda"); kh = torch.empty((G, E), dtype=torch.int32, device="cuda")
    res = {"E": E, "K": K, "G": G, "ray_solves": E * G}
    for mode in ("guided", "scan"):
        eng.set_search(mode)
        t, ms = timed(lambda: eng.tof_table_cuda(r, o, out_tof=tof, out_k_hit=kh))
        res[mode + "_ms"] = ms
        res[mode + "_ray_solves_per_s"] = E * G / (ms * 1e-3)
    res["hits"] = int((kh >= 0).sum())
    eng.set_search("guided")
    if not fp32:
        out[name] = res
        eng.close()
        return
    tof32 = torch.empty((G, E), dtype=torch.float32, device="cuda"); kh32 = torch.empty_like(kh)
    t32, ms32 = timed(lambda: eng.tof_table_f32_cuda(r, o, out_tof=tof32, out_k_hit=kh32))
    same = (kh >= 0) & (kh == kh32)
    res.update({"f32_ms": ms32, "f32_ray_solves_per_s": E * G / (ms32 * 1e-3),
                "f32_max_abs_dt_same_ray_s": float((tof32.double() - tof)[same].abs().max()),
                "f32_hit_flag_mismatch_frac": float(((kh >= 0) != (kh32 >= 0)).float().mean())})
    out[name] = res
    eng.close()
    del tof, kh, tof32, kh32
    torch.cuda.empty_cache()


out = {}
run("config1_main_rt_default_sweep", 64, 905, *rs.sweep_grid(), out)
run("config1b_main_compare", 64, 1810, np.array([0.037]), np.array([0.0038]), out)
run("config2_64el_512x512", 64, 905, *rs.synthetic_grid(512, 512), out)
run("config3_128el_1024x1024", 128, 905, *rs.synthetic_grid(1024, 1024), out)
run("sensitivity_K1810_128el_256x256", 128, 1810, *rs.synthetic_grid(256, 256), out)
run("sensitivity_K4096_128el_128x128", 128, 4096, *rs.synthetic_grid(128, 128), out)
# many receivers: the first-hit state no longer fits shared memory and lives in the output rows (HG instantiation)
run("large_E_1025_receivers_128x128_HG", 1024, 905, *rs.synthetic_grid(128, 128), out, pitch=0.04 / 1024, fp32=False)
# SURVEY row f2, full-matrix capture: every element shoots (S = E); stage-1 quantities per (shooter, ray), table [S][G][E]
def run_full_matrix(name, nel, K, radii, offsets, out):
    x_a, z_a = rs.transducer(num_elements=nel)
    alpha, zf = rs.launch_angles(K), rs.focus_line(K)
    eng = TofEngine(0).setup(alpha, zf, x_a, z_a, x_a)
    r, o = torch.from_numpy(np.ascontiguousarray(radii)).cuda(), torch.from_numpy(np.ascontiguousarray(offsets)).cuda()
    G, E = len(radii), len(x_a)
    tof = torch.empty((E, G, E), dtype=torch.float64, device="cuda"); kh = torch.empty((E, G, E), dtype=torch.int32, device="cuda")
    t, ms = timed(lambda: eng.tof_table_cuda(r, o, out_tof=tof, out_k_hit=kh))
    # the centre shooter's block against the single-shooter setup
    one = TofEngine(0).setup(alpha, zf, x_a[nel // 2], z_a[nel // 2], x_a).tof_table_cuda(r, o)
    out[name] = {"S": E, "E": E, "K": K, "G": G, "ray_solves": E * E * G, "guided_ms": ms,
                 "guided_ray_solves_per_s": E * E * G / (ms * 1e-3), "hits": int((kh >= 0).sum()),
                 "centre_shooter_block_equals_single_shooter_call": bool((tof[nel // 2] == one.tof).all() and (kh[nel // 2] == one.k_hit).all())}
    eng.close()


run_full_matrix("full_matrix_capture_129_shooters_64x64", 128, 905, *rs.synthetic_grid(64, 64), out)
if "--big" in sys.argv:
    run("config4_128el_4096x4096_one_gpu", 128, 905, *rs.synthetic_grid(4096, 4096), out)
print(json.dumps(out, indent=1))
