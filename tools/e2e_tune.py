#!/usr/bin/env python
"""GPU box: e2e rate of tofrt_hits_host / tofrt_table_host on config 3 against the device-resident kernel time, for
several chunk sizes of the hit-list pipeline (TOFRT_HIT_CHUNK_ROWS), plus graph-replayed latencies of config 1."""
import ctypes, json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ray_tracing_ultrasound_b200 import TofEngine, _capi, reference_setup as rs

x_a, z_a = rs.transducer(num_elements=128)
alpha, zf = rs.launch_angles(), rs.focus_line()
radii, offsets = rs.synthetic_grid(1024, 1024)
G, E = len(radii), len(x_a)
eng = TofEngine(0).setup(alpha, zf, x_a[64], z_a[64], x_a)
dev = torch.device("cuda", 0)
r_d, o_d = torch.from_numpy(radii).to(dev), torch.from_numpy(offsets).to(dev)
tof_d = torch.empty((G, E), dtype=torch.float64, device=dev); kh_d = torch.empty((G, E), dtype=torch.int32, device=dev)
for _ in range(3): eng.tof_table_cuda(r_d, o_d, out_tof=tof_d, out_k_hit=kh_d)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): eng.tof_table_cuda(r_d, o_d, out_tof=tof_d, out_k_hit=kh_d)
e1.record(); torch.cuda.synchronize()
dev_ms = e0.elapsed_time(e1) / 10
out = {"device_ms": dev_ms}
r_h, o_h = torch.from_numpy(radii).pin_memory(), torch.from_numpy(offsets).pin_memory()
cap = int((kh_d >= 0).sum()) + 1024
hb = [torch.empty(cap, dtype=torch.int32).pin_memory() for _ in range(3)] + [torch.empty(cap, dtype=torch.float64).pin_memory()]
n_hits = ctypes.c_longlong(0)
def hits():
    _capi.check(eng._lib.tofrt_hits_host(eng._ctx, r_h.data_ptr(), o_h.data_ptr(), G, cap, hb[0].data_ptr(), hb[1].data_ptr(),
                                         hb[2].data_ptr(), hb[3].data_ptr(), ctypes.byref(n_hits)))
for rows in (32768, 65536, 131072, 262144, 524288, 1048576):
    os.environ["TOFRT_HIT_CHUNK_ROWS"] = str(rows)
    hits(); hits()
    t0 = time.perf_counter()
    for _ in range(10): hits()
    out[f"hits_ms_chunk_{rows}"] = (time.perf_counter() - t0) / 10 * 1e3
# config 1 latency through a CUDA graph (no Python / launch-call overhead between the calls)
xa1, za1 = rs.transducer()
e1g = TofEngine(0).setup(rs.launch_angles(), rs.focus_line(), xa1[32], za1[32], xa1)
rr, oo = rs.sweep_grid()
r1, o1 = torch.from_numpy(rr).to(dev), torch.from_numpy(oo).to(dev)
t1 = torch.empty((210, 65), dtype=torch.float64, device=dev); k1 = torch.empty((210, 65), dtype=torch.int32, device=dev)
for mode in (True, False):
    e1g.set_small_g(mode)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): e1g.tof_table_cuda(r1, o1, out_tof=t1, out_k_hit=k1)
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(20): e1g.tof_table_cuda(r1, o1, out_tof=t1, out_k_hit=k1)
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    out[f"config1_graph_us_small_g_{int(mode)}"] = a.elapsed_time(b) / 20 * 1e3
print(json.dumps(out, indent=1))
