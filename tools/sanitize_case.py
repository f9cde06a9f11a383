"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / initcheck under gpurun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ray_tracing_ultrasound_b200 import TofEngine, reference_setup as rs

x_a, z_a = rs.transducer()
for K in (905, 77):
    alpha, zf = rs.launch_angles(K), rs.focus_line(K)
    eng = TofEngine(0).setup(alpha, zf, x_a[32], z_a[32], x_a)
    radii, offsets = rs.sweep_grid()
    radii, offsets = np.append(radii[::7], [1e-4, np.nan, 0.2]), np.append(offsets[::7], [0.05, 0.0, 0.0])
    for mode in ("guided", "scan"):
        eng.set_search(mode)
        t = eng.tof_table(radii, offsets, records=("idx", "target_x", "legs"))
        print(K, mode, int(t.hit.sum()), int((t.records["idx"] >= 0).sum()))
    import torch
    r, o = torch.from_numpy(radii).cuda(), torch.from_numpy(offsets).cuda()
    t32 = eng.tof_table_f32_cuda(r, o)
    torch.cuda.synchronize()
    print(K, "f32", int(t32.hit.sum()))
    eng.close()
print("done")
