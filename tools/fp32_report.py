"""FP32 mode vs FP64: timing and error statistics (run under gpurun)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ray_tracing_ultrasound_b200 import TofEngine, reference_setup as rs

def stats(t64, t32):
    h64, h32 = t64.k_hit >= 0, t32.k_hit >= 0
    both = h64 & h32
    dt = (t32.tof.double() - t64.tof)[both].abs()
    same_k = (t64.k_hit == t32.k_hit)[both]
    dts = (t32.tof.double() - t64.tof)[both][same_k].abs()
    return {"entries": int(h64.numel()), "hits_f64": int(h64.sum()), "hits_f32": int(h32.sum()),
            "hit_flag_mismatch": int((h64 != h32).sum()), "k_hit_mismatch_among_common_hits": int((~same_k).sum()),
            "max_abs_dt_common_hits_s": float(dt.max()), "max_abs_dt_same_ray_s": float(dts.max()),
            "median_abs_dt_same_ray_s": float(dts.median())}

out = {}
for name, nel, grid in (("config1_sweep", 64, None), ("config2_512x512", 64, (512, 512)), ("config3_1024x1024", 128, (1024, 1024))):
    x_a, z_a = rs.transducer(num_elements=nel); alpha, zf = rs.launch_angles(), rs.focus_line()
    eng = TofEngine(0).setup(alpha, zf, x_a[nel // 2], z_a[nel // 2], x_a)
    radii, offsets = rs.sweep_grid() if grid is None else rs.synthetic_grid(*grid)
    r, o = torch.from_numpy(radii).cuda(), torch.from_numpy(offsets).cuda()
    res = {}
    for mode, fn in (("f64", eng.tof_table_cuda), ("f32", eng.tof_table_f32_cuda)):
        for _ in range(2): t = fn(r, o)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); t = fn(r, o); e1.record(); torch.cuda.synchronize()
        res[mode] = (t, e0.elapsed_time(e1))
    s = stats(res["f64"][0], res["f32"][0])
    s.update({"f64_ms": res["f64"][1], "f32_ms": res["f32"][1], "G": len(radii), "E": len(x_a)})
    out[name] = s
    eng.close()
print(json.dumps(out, indent=1))
