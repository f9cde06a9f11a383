"""Two-point Fermat TOF table timing (extension): E elements x n x n focal points."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ray_tracing_ultrasound_b200 import TofEngine, reference_setup as rs
out = {}
for nel, n, K in ((64, 512, 905), (128, 1024, 905)):
    x_a, z_a = rs.transducer(num_elements=nel)
    alpha, zf = rs.launch_angles(K), rs.focus_line(K)
    eng = TofEngine(0).setup(alpha, zf, x_a[nel // 2], z_a[nel // 2], x_a)
    gx, gz = np.meshgrid(np.linspace(-0.04, 0.04, n), np.linspace(-0.01, 0.045, n))
    fx, fz = torch.from_numpy(gx.ravel()).cuda(), torch.from_numpy(gz.ravel()).cuda()
    eng.fermat_tof_cuda(fx, fz); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); t, k = eng.fermat_tof_cuda(fx, fz); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    pairs = len(x_a) * n * n
    out[f"{nel}el_{n}x{n}"] = {"pairs": pairs, "ms": ms, "pairs_per_s": pairs / (ms * 1e-3), "tof_min": float(t.min()), "tof_max": float(t.max())}
    eng.close()
print(json.dumps(out, indent=1))
