import sys; sys.path.insert(0,'.')
import numpy as np, torch, time
from ray_tracing_ultrasound_b200 import TofEngine, reference_setup as rs
x_a,z_a=rs.transducer(num_elements=128); alpha,zf=rs.launch_angles(),rs.focus_line()
eng=TofEngine(0).setup(alpha,zf,x_a[64],z_a[64],x_a)
radii,offsets=rs.synthetic_grid(128,1024)
r=torch.from_numpy(radii).cuda(); o=torch.from_numpy(offsets).cuda()
for mode in ('guided','scan'):
    eng.set_search(mode)
    eng.tof_table_cuda(r,o); torch.cuda.synchronize()
    t=time.time(); eng.tof_table_cuda(r,o); torch.cuda.synchronize(); print(mode, (time.time()-t)*1e3,'ms', eng.last_stats())
