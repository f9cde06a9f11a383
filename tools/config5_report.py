#!/usr/bin/env python
"""BASELINE config 5 (extension, parity unpinned): 32x32 matrix array x 256^3 voxels through the lens of
revolution, 3-D two-point solve (tofrt_fermat3d_*).  The volume is sharded over z-slabs: with W ranks
(torchrun) each rank solves 256/W slabs; on one GPU `--slabs` picks how many of the 256 slabs to run
(default 32 = one GPU's share of the 8-GPU job).  Prints one JSON line (rank 0).

    python tools/config5_report.py [--slabs 32] [--n-side 17] [--steps 3] [--cpu]
    torchrun --nproc-per-node 8 tools/config5_report.py --slabs 256
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--slabs", type=int, default=32, help="z-slabs (of 256) solved in total, split over the ranks")
    ap.add_argument("--n-side", type=int, default=9)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--cpu", action="store_true", help="time the CPU twin on a small sample (rank 0)")
    ap.add_argument("--e2e", action="store_true", help="also time the host-buffer path on 2 slabs")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    from ray_tracing_ultrasound_b200 import TofEngine, reference_setup as rs
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    x_a, z_a = rs.transducer()
    alpha, zf = rs.launch_angles(), rs.focus_line()
    eng = TofEngine(local).setup(alpha, zf, x_a[32], z_a[32], x_a)
    g = (np.arange(32) - 15.5) * float(rs.PITCH)          # 32 x 32 matrix array, 0.6 mm pitch
    ex, ey = (v.ravel() for v in np.meshgrid(g, g, indexing="ij"))
    eng.setup_fermat3d(ex, ey, n_side=args.n_side)
    E = len(ex)
    xs, ys, zs = np.linspace(-0.03, 0.03, 256), np.linspace(-0.03, 0.03, 256), np.linspace(0.0, 0.05, 256)
    per = args.slabs // world
    z_mine = zs[rank * per:(rank + 1) * per]
    vz, vy, vx = (v.ravel() for v in np.meshgrid(z_mine, ys, xs, indexing="ij"))   # x fastest
    F = len(vx)
    fx, fy, fz = (torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (vx, vy, vz))
    tof = torch.empty((E, F), dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng.fermat3d_tof_cuda(fx, fy, fz, out_tof=tof)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        eng.fermat3d_tof_cuda(fx, fy, fz, out_tof=tof)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    finite = bool(torch.isfinite(tof).all())
    out = {"config": "config5: 32x32 matrix array (E=1024) x 256x256x%d voxels, lens of revolution, n_side=%d"
                     % (args.slabs, args.n_side),
           "parity": "unpinned (no 3-D code in the reference)", "n_gpus": world, "E": E, "F_per_gpu": F,
           "pairs": E * F * world, "kernel_ms": ms, "pairs_per_s": E * F * world / (ms * 1e-3),
           "full_256_cubed_estimate_s_at_this_rate": 1024 * 256 ** 3 / (E * F * world / (ms * 1e-3)),
           "tof_write_gbs_per_gpu": E * F * 8 / (ms * 1e-3) / 1e9, "all_finite": finite,
           "tof_min_max_s": [float(tof.min()), float(tof.max())]}
    if args.e2e and rank == 0:
        n2 = 2 * 256 * 256
        hx, hy, hz = (np.ascontiguousarray(v[:n2]) for v in (vx, vy, vz))
        eng.fermat3d_tof(hx[:1024], hy[:1024], hz[:1024], with_kmin=False)
        t0 = time.perf_counter()
        th = eng.fermat3d_tof(hx, hy, hz, with_kmin=False)
        dt = time.perf_counter() - t0
        out["e2e_host_buffers"] = {"pairs_per_s": E * n2 / dt, "voxels": n2, "d2h_bytes": E * n2 * 8,
                                   "matches_device": bool(np.array_equal(th, tof[:, :n2].cpu().numpy())),
                                   "note": "pageable NumPy output array"}
    if args.cpu and rank == 0:
        from oracle.oracle import Oracle
        o = Oracle(device_math=True)
        sel = np.linspace(0, F - 1, 2000).astype(np.int64)
        t0 = time.perf_counter()
        ct, ck, cab, cev, cgap = o.fermat3d(o.params(), eng.f3_ang, ex[::16], ey[::16], float(rs.D), vx[sel], vy[sel], vz[sel])
        dt = time.perf_counter() - t0
        gt = tof[::16][:, torch.from_numpy(sel).to(dev)].cpu().numpy()
        out["cpu_twin"] = {"pairs_per_s": ct.size / dt, "cores": 1, "sample": "64 elements x 2000 voxels",
                           "max_abs_dt_vs_gpu_s": float(np.abs(ct - gt).max()),
                           "bit_identical_frac": float((ct == gt).mean()), "mean_evals": float(cev.mean())}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
