#!/usr/bin/env python
"""bench.py -- element x pixel ray-solves/s of the TOF table hot path.

    python bench.py --gpus N --steps K --warmup W          # product arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N ...          # CPU arm (oracle port on the host cores)

A "step" is one pass of the hot path (main_rt.py:464-501: trace the K rays of the shot and assemble the
receiver table) over one batch of synthetic geometries, 128-element array (+ centre element: E = 129),
K = N = 905 launch rays / lens samples:

  N = 1   BASELINE config 3: the 1024 x 1024 (r_outer, pipe_offset) grid on one GPU.
  N > 1   BASELINE config 4: the 4096 x 4096 grid SHARDED across the N GPUs (strong scaling: total work fixed).
          Radius rows are dealt round-robin (rank r owns rows r, r + N, ...: ray_tracing_ultrasound_b200/sharding.py),
          because the cost of a geometry depends on its radius -- contiguous radius slabs left the slowest rank 3.6 %
          behind at 8 GPUs.  No collective on the data path (geometries are independent); outside the timed region
          every run all-gathers a bounded slice over NCCL and checks it against a single-GPU recomputation.

`value`   : E * G_total / t with radius/offset resident in HBM (CUDA-event timed, max over ranks).
`e2e`     : the same pass through the C ABI with HOST (pinned) buffers, H2D of radius/offset and D2H of the result
            inside the timed region.  --e2e-form sparse (default): tofrt_hits_host, the list of the table's hits
            (every other entry is `False, 0`: the complete result); dense: tofrt_table_host, the full table.  The
            other form is reported beside it.
`roofline`: FP64 instruction issue.  `frac` = FP64-pipe instructions the kernel executed (counted by the instrumented
            instantiation of the same kernel, outside the timed region) / launch time / peak; `frac_executed` = the same
            with useful operations (division, square root = 1) instead of instructions; `frac_vs_reference_algorithm` =
            algorithmic FP64 ops of SURVEY section 8(d) (the reference's full N-sample scan per ray) / launch time / peak.  peak = FP64 instruction rate measured on this GPU by
            tofrt_measure_fp64_peak (MEASURED_PEAKS.json has no FP64 entry).  The hardware-side figure (FP64-pipe
            utilisation from ncu) is quoted from profiles/ only when the capture belongs to this very source.
`cpu_baseline`: the oracle port (C, all host threads) on a bounded sample.
"""
import argparse
import ctypes
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ray-solves/sec (element x pixel)"
UNIT = "ray-solves/s"
N_ELEM, K_RAYS, GRID_O = 128, 905, 1024
NCU_SUMMARY = os.path.join(ROOT, "profiles", "ncu_r2_table_kernel.json")

# FP64 operations per item of the vector-form kernel (+, -, *, /, sqrt = 1; fma = 2; comparisons not counted),
# read off csrc/tofrt_raymath.h and csrc/tofrt_table_kernel.cuh; DESIGN.md section 3.3 lists them line by line
OPS = {
    "ray": 16 + 12 + 3,        # tofrt_vec_pipe, tofrt_vec_reflect, m = sl / cl and b
    "fast_ray": 16,            # tol, the two end diffs, probe constants
    "probe": 2,                # one fma per bracket probe
    "exact_diff": 3,           # main_rt.py:78: mul, add, sub
    "lens2_ray": 9 + 33,       # chord intersection + bounds; tofrt_vec_return_head
    "landing_ray": 11,         # tofrt_vec_return_tail
    "receiver_test": 3,        # np.isclose
    "winner_ray": 24,          # three legs (dist / c) and the sum
}


# FP64-pipe INSTRUCTIONS per item of the same kernel (a division is 9 of them, a square root 8, a comparison 1), read
# off the per-line SASS table of the profiled build (tools/line_breakdown.py on profiles' guided capture).  With the
# counters of the instrumented launch they reproduce ncu's sm__inst_executed_pipe_fp64 of that build to within 4 %.
FP64_INST = {
    "ray": 51,                 # pipe 19, reflect 18, m 9, b 1, cell index 4
    "fast_ray": 31,            # tol, end diffs, certainty tests, probe constants
    "probe": 3,                # DFMA + 2 DSETP
    "exact_diff": 6,           # DMUL, DADD, DADD + sign tests
    "lens2_ray": 69,           # chord intersection 28, tofrt_vec_return_head 41
    "landing_ray": 26,         # tofrt_vec_return_tail
    "receiver_test": 5,
    "winner_ray": 54,
}


def source_hash():
    """sha256 over the CUDA sources + the C header: stamps ncu captures (tools/ncu_summary.py) to the build they profile."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "ray_tracing_ultrasound_b200", "csrc")
    for name in sorted(os.listdir(d)) + ["../../include/tofrt.h"]:
        if name.endswith((".cu", ".cuh", ".h")):
            h.update(open(os.path.join(d, name), "rb").read())
    return h.hexdigest()[:16]


def workload(rank, world, grid_r=1024):
    """-> x_a, z_a, alpha, zf, radii, offsets (this rank's), description"""
    from ray_tracing_ultrasound_b200 import reference_setup as rs, sharding
    x_a, z_a = rs.transducer(num_elements=N_ELEM)
    alpha, zf = rs.launch_angles(K_RAYS), rs.focus_line(K_RAYS)
    if world == 1:
        radii, offsets = rs.synthetic_grid(grid_r, GRID_O)
        return x_a, z_a, alpha, zf, radii, offsets, {
            "name": f"config 3: {N_ELEM}-element array (E={len(x_a)}), K=N={K_RAYS} rays/lens samples, "
                    f"{grid_r}x{GRID_O} (r_outer, pipe_offset) grid on 1 GPU", "grid": [grid_r, GRID_O]}
    n = 4096
    radii, offsets = rs.synthetic_grid(n, n)
    r_s, o_s = sharding.shard_geometries_interleaved(radii, offsets, world, rank, block=n)
    return x_a, z_a, alpha, zf, r_s, o_s, {
        "name": f"config 4: {N_ELEM}-element array (E={len(x_a)}), K=N={K_RAYS}, {n}x{n} grid sharded across {world} GPUs "
                f"(radius rows dealt round-robin)", "grid": [n, n]}


def algorithmic_ops(G, K, N, E):
    """W_g = K*(155 + 3N) + 2*E*K FP64 ops per geometry."""
    return float(G) * (K * (155.0 + 3.0 * N) + 2.0 * E * K)


def executed_ops(st, per_item=None):
    """FP64 operations (per_item = OPS) or FP64-pipe instructions (per_item = FP64_INST) the table kernel executed,
    from the instrumented launch's counters (TofEngine.last_stats)."""
    w = per_item or OPS
    return float(w["ray"] * st["rays"] + w["fast_ray"] * (st["bracket_rays"] + st["scan_rays"]) +
                 w["probe"] * st["probes"] + w["exact_diff"] * st["exact_diffs"] + w["lens2_ray"] * st["lens2_rays"] +
                 w["landing_ray"] * st["landing_rays"] + w["receiver_test"] * st["receiver_tests"] +
                 w["winner_ray"] * st["winner_rays"])


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region (the profiling recipe's clocks line).
    Reads NVML in-process (what nvidia-smi prints) every 20 ms; falls back to nvidia-smi."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz, self.source = None, "nvml"

    def _run_nvml(self):
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
            bits = int(get_reasons(h))
            for bit, name in self.REASONS.items():
                if bits & bit:
                    self.reasons.add(name)
            time.sleep(0.02)

    def _run_smi(self):
        self.source = "nvidia-smi"
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.05)

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples),
                "source": self.source}


def cpu_port_rate(x_a, z_a, alpha, zf, radii, offsets, target_s=8.0):
    """Oracle port (C, pthreads over geometries) on a bounded, evenly strided sample."""
    from oracle.oracle import Oracle
    o = Oracle(device_math=False)
    p = o.params()
    threads = o.max_threads()
    n0 = min(len(radii), 64 * threads)
    sel = np.linspace(0, len(radii) - 1, n0).astype(np.int64)
    t0 = time.perf_counter()
    o.table(p, radii[sel], offsets[sel], x_a[len(x_a) // 2], z_a[0], zf, alpha, x_a)
    dt = time.perf_counter() - t0
    n = int(min(len(radii), max(n0, n0 * target_s / max(dt, 1e-3))))
    sel = np.linspace(0, len(radii) - 1, n).astype(np.int64)
    t0 = time.perf_counter()
    o.table(p, radii[sel], offsets[sel], x_a[len(x_a) // 2], z_a[0], zf, alpha, x_a)
    dt = time.perf_counter() - t0
    return len(x_a) * n / dt, threads, n, dt


def run_reference(args, rank, world):
    """CPU arm: the reference is pure Python and cannot travel to the GPU box, so this arm
    times the oracle port of the same path on all host threads (kind = "port")."""
    if rank != 0:
        return
    x_a, z_a, alpha, zf, radii, offsets, desc = workload(0, 1)
    if world > 1:
        from ray_tracing_ultrasound_b200 import reference_setup as rs
        radii, offsets = rs.synthetic_grid(4096, 4096)
        desc = {"name": f"config 4: {N_ELEM}-element array (E={len(x_a)}), K=N={K_RAYS}, 4096x4096 grid"}
    from oracle.oracle import Oracle
    o = Oracle(device_math=False)
    p = o.params()
    threads = o.max_threads()
    n = 256 * threads  # geometries per step: a bounded, evenly strided sample of the workload
    sel = np.linspace(0, len(radii) - 1, n).astype(np.int64)
    r_s, o_s = radii[sel], offsets[sel]
    for _ in range(args.warmup):
        o.table(p, r_s[: n // 8], o_s[: n // 8], x_a[64], z_a[64], zf, alpha, x_a)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.table(p, r_s, o_s, x_a[64], z_a[64], zf, alpha, x_a)
    dt = time.perf_counter() - t0
    value = len(x_a) * n * args.steps / dt
    sample = f"{n} of {len(radii)} geometries per step, evenly strided, E={len(x_a)}, K=N={K_RAYS}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": desc["name"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=1024, help="radii of the N=1 grid (default 1024; offsets stay 1024)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-3d", action="store_true", help="skip the 3-D two-point solve report (config 5 slice)")
    ap.add_argument("--no-extras", action="store_true", help="timed steps, e2e and roofline only (profiling runs)")
    ap.add_argument("--e2e-form", default="sparse", choices=["sparse", "dense"],
                    help="which host-buffer form `e2e.value` reports (the other one is reported beside it)")
    ap.add_argument("--search", default="guided", choices=["guided", "scan"],
                    help="return-path search of the timed steps (identical results; see include/tofrt.h)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # `python bench.py --gpus N` typed by hand: become the launch line the driver uses (one rank per GPU)
        import socket
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        os.execv(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                  f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                                  "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:])
    # stdout carries exactly ONE JSON line: anything a library prints meanwhile (NCCL's version banner at the first
    # communicator, ...) goes to stderr -- file descriptor 1 is pointed at stderr until the line is written
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from ray_tracing_ultrasound_b200 import TofEngine, _build, _capi, sharding, reference_setup as rs
    if _build.is_stale():
        _build.build()
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    x_a, z_a, alpha, zf, radii, offsets, desc = workload(rank, world, args.grid)
    E, K, G = len(x_a), len(alpha), len(radii)
    eng = TofEngine(local).setup(alpha, zf, x_a[N_ELEM // 2], z_a[N_ELEM // 2], x_a)
    eng.set_search(args.search)
    dev = torch.device("cuda", local)
    r_d, o_d = torch.from_numpy(radii).to(dev), torch.from_numpy(offsets).to(dev)
    tof_d = torch.empty((G, E), dtype=torch.float64, device=dev)
    kh_d = torch.empty((G, E), dtype=torch.int32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step():
        eng.tof_table_cuda(r_d, o_d, out_tof=tof_d, out_k_hit=kh_d)

    def timed(fn, n):
        """n calls of fn bracketed by CUDA events on the current stream -> ms per call (this rank)"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    barrier()
    sampler.stop_flag = True
    my_ms = ev[0].elapsed_time(ev[-1])
    launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = allmax(my_ms)
    G_total = G
    rank_ms = [my_ms / args.steps]
    if world > 1:
        gt = torch.tensor([G], dtype=torch.int64, device=dev)
        dist.all_reduce(gt)
        G_total = int(gt.item())
        allms = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(allms, torch.tensor([my_ms / args.steps], dtype=torch.float64, device=dev))
        rank_ms = [float(t.item()) for t in allms]
    value = E * G_total * args.steps / (total_ms * 1e-3)

    # ---- e2e: C ABI with host (pinned) buffers, copies inside the timed region ----
    r_h = torch.from_numpy(radii).pin_memory()
    o_h = torch.from_numpy(offsets).pin_memory()
    n_hits_dense = int((kh_d >= 0).sum())
    cap = n_hits_dense + 1024
    hb = (torch.empty(cap, dtype=torch.int32).pin_memory(), torch.empty(cap, dtype=torch.int32).pin_memory(),
          torch.empty(cap, dtype=torch.int32).pin_memory(), torch.empty(cap, dtype=torch.float64).pin_memory())
    n_hits = ctypes.c_longlong(0)

    def hits_step():
        _capi.check(eng._lib.tofrt_hits_host(eng._ctx, r_h.data_ptr(), o_h.data_ptr(), G, cap, hb[0].data_ptr(),
                                             hb[1].data_ptr(), hb[2].data_ptr(), hb[3].data_ptr(), ctypes.byref(n_hits)))

    Gd = min(G, 2 * 1024 * 1024)   # the dense form is measured on at most 2 Mi geometries per rank (26 GB would be pinned otherwise)
    tof_h = torch.empty((Gd, E), dtype=torch.float64).pin_memory()
    kh_h = torch.empty((Gd, E), dtype=torch.int32).pin_memory()

    def dense_step():
        _capi.check(eng._lib.tofrt_table_host(eng._ctx, r_h.data_ptr(), o_h.data_ptr(), Gd, tof_h.data_ptr(),
                                              kh_h.data_ptr(), None))

    def wall(fn, n):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        barrier()
        return allmax(time.perf_counter() - t0) / n

    main_steps, side_steps = max(10, args.steps), 3
    sparse_first = args.e2e_form == "sparse"
    hits_s = wall(hits_step, main_steps if sparse_first else side_steps)
    dense_s = wall(dense_step, side_steps if sparse_first else main_steps)
    nh = int(n_hits.value)
    hits_ok = bool(nh == n_hits_dense and torch.equal(
        hb[3][:nh], tof_d.cpu()[hb[0][:nh].long(), hb[1][:nh].long()] if G <= 2 * 1024 * 1024 else
        tof_d[hb[0][:nh].long().to(dev), hb[1][:nh].long().to(dev)].cpu()))
    dense_ok = bool(torch.equal(tof_h, tof_d[:Gd].cpu()) and torch.equal(kh_h, kh_d[:Gd].cpu()))
    nh_total = int(allmax(0)) if False else nh
    if world > 1:
        t = torch.tensor([nh], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        nh_total = int(t.item())
    sparse = {"value": E * G_total / hits_s, "unit": UNIT, "h2d_bytes_per_step": int(G_total * 16),
              "d2h_bytes_per_step": int(nh_total * 20), "hits_per_step": nh_total, "matches_device_path": hits_ok,
              "steps": main_steps if sparse_first else side_steps,
              "api": "tofrt_hits_host: hits emitted by the table kernel itself, chunks pipelined with their D2H"}
    dense = {"value": E * Gd * world / dense_s, "unit": UNIT, "h2d_bytes_per_step": int(Gd * 16 * world),
             "d2h_bytes_per_step": int(Gd * E * 12 * world), "matches_device_path": dense_ok,
             "steps": side_steps if sparse_first else main_steps, "geometries_per_rank": Gd,
             "api": "tofrt_table_host: full [G][E] table, 64 MB chunks pipelined with their D2H"}
    best = sparse if sparse_first else dense
    e2e = {"value": best["value"], "unit": UNIT, "form": "sparse_hits" if sparse_first else "dense",
           "h2d_bytes_per_step": best["h2d_bytes_per_step"], "d2h_bytes_per_step": best["d2h_bytes_per_step"],
           "steps": best["steps"], "vs_device_resident": best["value"] / value, "sparse_hits": sparse, "dense": dense,
           "note": "pinned host buffers; H2D of radius/offset and D2H of the result inside the timed region; the form is "
                   "chosen by --e2e-form (default sparse: the hit list IS the complete result, misses are `False, 0` rows)"}
    del tof_h, kh_h

    # ---- executed work: one launch of the instrumented instantiation, outside the timed region ----
    eng.set_stats(True)
    step()
    torch.cuda.synchronize()
    stats = eng.last_stats()
    eng.set_stats(False)

    extras = {}
    if not args.no_extras:
        # ---- the other search mode, device-resident, for the roofline discussion ----
        other = "scan" if args.search == "guided" else "guided"
        eng.set_search(other)
        tof_o, kh_o = torch.empty_like(tof_d), torch.empty_like(kh_d)
        eng.tof_table_cuda(r_d, o_d, out_tof=tof_o, out_k_hit=kh_o)
        barrier()
        other_ms = timed(lambda: eng.tof_table_cuda(r_d, o_d, out_tof=tof_o, out_k_hit=kh_o), 2)
        extras["other_ms"], extras["other"] = other_ms, other
        extras["other_same"] = bool(torch.equal(tof_o, tof_d) and torch.equal(kh_o, kh_d))
        eng.set_stats(True)
        eng.tof_table_cuda(r_d, o_d, out_tof=tof_o, out_k_hit=kh_o)
        torch.cuda.synchronize()
        extras["other_stats"] = eng.last_stats()
        eng.set_stats(False)
        eng.set_search(args.search)

        # ---- the reference-order math mode (TOFRT_MATH_ANGLE) beside the default vector form ----
        eng.set_math("angle")
        eng.tof_table_cuda(r_d, o_d, out_tof=tof_o, out_k_hit=kh_o)
        barrier()
        angle_ms = timed(lambda: eng.tof_table_cuda(r_d, o_d, out_tof=tof_o, out_k_hit=kh_o), 2)
        both_a = (kh_o >= 0) & (kh_d >= 0)
        math_info = {"default": "vector", "angle_kernel_ms": angle_ms,
                     "k_hit_mismatches_vs_angle": int((kh_o != kh_d).sum()),
                     "max_abs_dt_vs_angle_s": float((tof_o - tof_d)[both_a].abs().max())}
        del both_a
        # discrete return-path index, ray by ray, on a bounded evenly strided subsample: vector form vs the
        # reference-order angle form (device), and vs the independent libm oracle (CPU) with its noise margin
        sel = np.linspace(0, G - 1, min(G, 16384)).astype(np.int64)
        rs_d, os_d = r_d[torch.from_numpy(sel).to(dev)].contiguous(), o_d[torch.from_numpy(sel).to(dev)].contiguous()
        idx_a = eng.tof_table_cuda(rs_d, os_d, records=("idx",)).records["idx"]
        eng.set_math("vector")
        idx_v = eng.tof_table_cuda(rs_d, os_d, records=("idx",)).records["idx"]
        torch.cuda.synchronize()
        math_info["idx_rays_compared"] = int(idx_v.numel())
        math_info["idx_differs_vs_angle"] = int((idx_v != idx_a).sum())
        if rank == 0:
            from oracle.oracle import Oracle
            ol = Oracle(device_math=False)
            sub = sel[:: max(1, len(sel) // 1024)][:1024]
            _, _, idx_l, margin = ol.table(ol.params(), radii[sub], offsets[sub], x_a[64], z_a[64], zf, alpha, x_a,
                                           want_idx=True, want_margin=True)
            idx_g = idx_v.cpu().numpy()[:: max(1, len(sel) // 1024)][:1024]
            diff = idx_g != idx_l
            math_info["idx_vs_libm_oracle"] = {"rays": int(idx_l.size), "differ": int(diff.sum()),
                                               "differ_above_noise_floor_1e-13m": int((diff & (margin > 1e-13)).sum())}
        math_info["note"] = ("TOFRT_OPT_MATH: 'vector' (default) evaluates the return path from direction cosines "
                             "(csrc/tofrt_raymath.h: 3 divisions and 5 square roots per ray), 'angle' keeps the reference's "
                             "angle bookkeeping and operation order; both are bit-identical to their CPU twin and pass the "
                             "same parity tests; idx may differ from the reference order only inside its rounding noise")
        extras["math"] = math_info
        del tof_o, kh_o, idx_a, idx_v

        # ---- FP32 mode, reported separately (north_star): time, roofline and measured error vs FP64 ----
        tof32 = torch.empty((G, E), dtype=torch.float32, device=dev)
        kh32 = torch.empty((G, E), dtype=torch.int32, device=dev)
        eng.tof_table_f32_cuda(r_d, o_d, out_tof=tof32, out_k_hit=kh32)
        barrier()
        f32_ms = timed(lambda: eng.tof_table_f32_cuda(r_d, o_d, out_tof=tof32, out_k_hit=kh32), 2)
        h64, h32 = kh_d >= 0, kh32 >= 0
        both = h64 & h32
        same_ray = both & (kh_d == kh32)
        dts = (tof32.double() - tof_d)[same_ray].abs()
        extras["fp32"] = {
            "value": E * G / (f32_ms * 1e-3), "unit": UNIT + " per GPU", "kernel_ms": f32_ms,
            "max_abs_dt_same_ray_s": float(dts.max()), "median_abs_dt_same_ray_s": float(dts.median()),
            "max_abs_dt_common_hits_s": float((tof32.double() - tof_d)[both].abs().max()),
            "hit_flag_mismatch_frac": float((h64 != h32).float().mean()),
            "first_ray_mismatch_frac_of_common_hits": float(1.0 - same_ray.sum().double() / both.sum().double()),
            "note": "per-ray arithmetic in float32 (tables rounded from FP64); landing points move by ~5e-7 m, so hit "
                    "decisions at atol = 1e-6 m differ; no parity claim for this mode",
        }
        del tof32, kh32, h64, h32, both, same_ray, dts

        # ---- BASELINE config 1 (the reference's own 210-geometry sweep) and 1b (main_compare.py's single K = 1810
        #      shot): latency of one table call, device-resident, small-G mode on (default) and off ----
        if rank == 0:
            small = {}
            xa1, za1 = rs.transducer()
            for name, kk, (rr, oo) in (("config1_210_geometries_K905", 905, rs.sweep_grid()),
                                       ("config1b_1_geometry_K1810", 1810, (np.array([0.037]), np.array([0.0038])))):
                e1 = TofEngine(local).setup(rs.launch_angles(kk), rs.focus_line(kk), xa1[32], za1[32], xa1)
                r1, o1 = torch.from_numpy(np.ascontiguousarray(rr)).to(dev), torch.from_numpy(np.ascontiguousarray(oo)).to(dev)
                t1 = torch.empty((len(rr), 65), dtype=torch.float64, device=dev)
                k1 = torch.empty((len(rr), 65), dtype=torch.int32, device=dev)
                res = {}
                for mode in (True, False):
                    e1.set_small_g(mode)
                    call = lambda: e1.tof_table_cuda(r1, o1, out_tof=t1, out_k_hit=k1)
                    for _ in range(5):
                        call()
                    torch.cuda.synchronize()
                    res["small_g_on_us" if mode else "small_g_off_us"] = timed(call, 50) * 1e3
                res["ray_solves_per_s"] = 65 * len(rr) / (res["small_g_on_us"] * 1e-6)
                small[name] = res
                e1.close()
            small["note"] = ("one tofrt_table_dev call incl. launch, back to back on one stream; config 1b has 57 chunks per "
                             "shot (> 32 warps), so it always takes the regular kernel")
            extras["small_configs"] = small

        # ---- 3-D two-point solve (BASELINE config 5, extension -- parity unpinned), reported separately ----
        if rank == 0 and not args.no_3d:
            gm = (np.arange(32) - 15.5) * float(rs.PITCH)                   # 32 x 32 matrix array
            ex, ey = (v.ravel() for v in np.meshgrid(gm, gm, indexing="ij"))
            eng.setup_fermat3d(ex, ey)                                       # default 9 x 9 surface grid
            vz, vy, vx = (v.ravel() for v in np.meshgrid(np.linspace(0.0, 0.05, 256)[:4], np.linspace(-0.03, 0.03, 256),
                                                         np.linspace(-0.03, 0.03, 256), indexing="ij"))
            f3 = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (vx, vy, vz)]
            t3 = torch.empty((len(ex), len(vx)), dtype=torch.float64, device=dev)
            eng.fermat3d_tof_cuda(*f3, out_tof=t3)
            torch.cuda.synchronize()
            ms3 = timed(lambda: eng.fermat3d_tof_cuda(*f3, out_tof=t3), 3)
            extras["fermat3d"] = {"value": len(ex) * len(vx) / (ms3 * 1e-3), "unit": "(element, voxel) pairs/s per GPU",
                                  "kernel_ms": ms3, "workload": "config 5 slice: 32x32 matrix array x 256x256x4 voxels of the "
                                  "256^3 volume, lens of revolution, 9x9 surface grid + projected Newton",
                                  "all_finite": bool(torch.isfinite(t3).all()),
                                  "note": "no 3-D code in the reference: parity unpinned (oracles: NumPy/SciPy + bit-exact twin, tests/)"}
            del t3, f3

    # ---- multi-GPU: NCCL all-gather of a bounded slice + shard equivalence, outside the timed region ----
    gather_info = None
    if world > 1:
        # (1) a 64 x 64 sub-grid of config 4, sharded by the same round-robin rule, gathered and compared with a
        #     single-GPU recomputation of the whole sub-grid on every rank
        sub_r, sub_o = rs.synthetic_grid(4096, 4096)
        pick = (np.arange(0, 4096, 64)[:, None] * 4096 + np.arange(0, 4096, 64)[None, :]).ravel()
        sub_r, sub_o = np.ascontiguousarray(sub_r[pick]), np.ascontiguousarray(sub_o[pick])
        mine_r, mine_o = sharding.shard_geometries_interleaved(sub_r, sub_o, world, rank, block=64)
        tm = eng.tof_table_cuda(torch.from_numpy(mine_r).to(dev), torch.from_numpy(mine_o).to(dev))
        full_t, full_k = sharding.gather_table_interleaved(tm.tof, tm.k_hit, len(sub_r), 64)
        one = eng.tof_table_cuda(torch.from_numpy(sub_r).to(dev), torch.from_numpy(sub_o).to(dev))
        torch.cuda.synchronize()
        equiv = bool(torch.equal(full_t, one.tof) and torch.equal(full_k, one.k_hit))
        # (2) bandwidth-sized gather of this run's own shards (65 536 rows per rank), checked against the local rows
        rows = min(G, 65536)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        tof_all, kh_all = sharding.gather_table(tof_d[:rows], kh_d[:rows], rows * world)
        g1.record()
        torch.cuda.synchronize()
        ok = bool(equiv and torch.equal(tof_all[rank * rows:(rank + 1) * rows], tof_d[:rows]) and
                  torch.equal(kh_all[rank * rows:(rank + 1) * rows], kh_d[:rows]) and tof_all.shape[0] == rows * world)
        okt = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        gather_info = {"ok": bool(okt.item()), "shard_equivalence_64x64": equiv, "rows_per_rank": rows,
                       "bytes_gathered_per_rank": int(rows * world * E * 12), "ms": g0.elapsed_time(g1),
                       "gbs_per_rank": rows * world * E * 12 / (g0.elapsed_time(g1) * 1e-3) / 1e9}
        del tof_all, kh_all, full_t, full_k

    if rank == 0:
        peak_fma, peak_issue = eng.measure_fp64_peak()
        peak32_fma, peak32_issue = eng.measure_fp32_peak()
        info = eng.kernel_info()
        kern_s = float(np.mean(launch_ms)) * 1e-3
        ops_ref = algorithmic_ops(G, K, K, E)
        ops_exec = executed_ops(stats)
        inst_exec = executed_ops(stats, FP64_INST)
        roofline = {
            "bound": "fp64-issue", "kernel": "tofrt_table_kernel", "kernel_ms": kern_s * 1e3, "search": args.search,
            "peak": peak_issue / 1e12, "unit": "TFLOP/s",
            "peak_source": "tofrt_measure_fp64_peak on this GPU (un-fused DMUL+DADD thread-instructions/s); "
                           "MEASURED_PEAKS.json holds no FP64 figure",
            "peak_dfma_tflops": 2 * peak_fma / 1e12,
            # executed FP64-pipe instructions / time / measured FP64 instruction peak: the pipe's utilisation (guided mode)
            "achieved": inst_exec / kern_s / 1e12, "frac": inst_exec / kern_s / peak_issue,
            "frac_fp64_instructions": inst_exec / kern_s / peak_issue, "executed_fp64_instructions_per_launch": inst_exec,
            "fp64_instructions_per_item": FP64_INST,
            # executed useful operations (a division or square root = 1): lower bound of the above
            "frac_executed": ops_exec / kern_s / peak_issue, "executed_fp64_ops_per_launch": ops_exec,
            "executed_counters": stats, "ops_per_item": OPS,
            # the reference's algorithm (full N-sample scan per ray) done in this time
            "frac_vs_reference_algorithm": ops_ref / kern_s / peak_issue, "algorithmic_ops_per_launch": ops_ref,
            "traffic": None, "table_write_gbs": G * E * 12 / kern_s / 1e9,
            "note": "frac = frac_fp64_instructions: FP64-pipe instructions the timed kernel executed (counters of the instrumented "
                    "instantiation of the same kernel, run outside the timed region, x the per-item instruction counts read off "
                    "the SASS: fp64_instructions_per_item) / launch time / measured FP64 instruction peak -- the in-run twin of "
                    "ncu's sm__inst_executed_pipe_fp64 (roofline.ncu, quoted when the capture belongs to this source).  "
                    "frac_executed counts useful operations instead (ops_per_item; a division or square root = ONE operation "
                    "although it issues ~9 FP64 instructions): a lower bound.  "
                    "frac_vs_reference_algorithm uses SURVEY 8(d)'s count G*(K*(155+3N)+2EK), i.e. the reference's full "
                    "N-sample scan per ray, which the guided search provably skips: it exceeds 1 and is a speed-up over "
                    "the reference algorithm at the roofline, not a utilisation.",
        }
        try:   # the hardware-side figures: quoted only from a capture of THIS source whose time agrees within 3 %
            ncu = json.load(open(NCU_SUMMARY))
            cap_ms = ncu["kernel_ms"]
            if ncu.get("source_hash") == source_hash() and abs(cap_ms - kern_s * 1e3) <= 0.03 * kern_s * 1e3:
                roofline["ncu"] = {k: ncu[k] for k in ("fp64_pipe_pct", "issue_slots_busy_pct", "kernel_ms", "dram_bytes",
                                                       "warp_instructions", "source_hash", "file") if k in ncu}
                roofline["traffic"] = ncu.get("dram_bytes")
            else:
                roofline["ncu"] = {"stale": True, "why": "profiles/ncu_r2_table_kernel.json was captured on another source "
                                   f"or its kernel time ({cap_ms:.2f} ms) is not within 3 % of this run's"}
        except Exception:
            roofline["ncu"] = None
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            roofline["hbm_peak_gbs_measured"] = peaks.get("hbm_gbs")
        except Exception:
            roofline["hbm_peak_gbs_measured"] = None
        if "other_ms" in extras:
            o_s = extras["other_ms"] * 1e-3
            roofline["roofline_" + extras["other"]] = {
                "kernel_ms": extras["other_ms"], "frac_vs_reference_algorithm": ops_ref / o_s / peak_issue,
                "frac_executed": executed_ops(extras["other_stats"]) / o_s / peak_issue, "modes_identical": extras["other_same"],
                "note": "the other search mode, same outputs bit for bit; for the literal scan every exact diff is 3 FP64 "
                        "operations and the SURVEY 8(d) count is the like-for-like one"}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc["name"], "G_total": G_total, "G_per_gpu": G, "E": E, "K": K,
                       "parallelism": f"geometry-sharded x{world}" + ("" if world == 1 else ", radius rows round-robin"),
                       "l2": "outputs (1.6 GB per 1 Mi geometries) exceed L2; inputs are 16 B per geometry + 60 KB of tables",
                       "kernel": info, "source_hash": source_hash()},
            "clocks": sampler.summary(),
            "rank_ms": {"min": min(rank_ms), "max": max(rank_ms), "max_over_min": max(rank_ms) / min(rank_ms)},
            "e2e": e2e,
            "gpu_launches": args.steps * 1,
            "roofline": roofline,
        }
        if "math" in extras:
            out["math_modes"] = extras["math"]
        if "fp32" in extras:
            f = extras["fp32"]
            # FP32 mode: the same SURVEY 8(d) operation count against the measured FP32 issue rate
            f["roofline"] = {"bound": "fp32-issue", "peak": peak32_issue / 1e12, "peak_ffma_tflops": 2 * peak32_fma / 1e12,
                             "frac_vs_reference_algorithm": ops_ref / (f["kernel_ms"] * 1e-3) / peak32_issue,
                             "peak_source": "tofrt_measure_fp32_peak on this GPU (un-fused FMUL+FADD thread-instructions/s)"}
            out["fp32"] = f
        if "small_configs" in extras:
            out["small_configs"] = extras["small_configs"]
        if "fermat3d" in extras:
            out["fermat3d"] = extras["fermat3d"]
        if gather_info:
            out["all_gather"] = gather_info
        if not args.no_cpu:
            v, threads, n, dt = cpu_port_rate(x_a, z_a, alpha, zf, radii, offsets)
            numpy_leg = None
            try:   # the reference's NumPy CPU path: its expressions vectorised over the rays of a shot (oracle/oracle_numpy.py)
                if world > 1:
                    raise RuntimeError("measured on single-GPU runs only")
                from oracle import oracle_numpy
                procs = max(1, min(threads, os.cpu_count() or 1))
                ngeo = 48 * procs
                sel = np.linspace(0, G - 1, ngeo).astype(np.int64)
                nv, ndt = oracle_numpy.timed_rate(x_a, z_a, zf, alpha, N_ELEM // 2, radii[sel], offsets[sel], procs)
                numpy_leg = {"value": nv, "unit": UNIT, "cores": procs, "kind": "port",
                             "sample": f"{ngeo} of {G} geometries, evenly strided, {ndt:.1f} s, {procs} processes",
                             "note": "main_rt.py's own NumPy expressions in its operation order, vectorised over the K rays "
                                     "of a shot (K x N diffs matrix, K x E np.isclose) instead of Python loops; the verbatim "
                                     "script does 35.6 ray-solves/s per core (BASELINE.md) and cannot travel to the GPU box"}
            except Exception as exc:   # the baseline must never take the bench line down
                numpy_leg = {"unavailable": repr(exc)}
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "numpy": numpy_leg,
                                   "note": "C restatement of main_rt.py (oracle/), pthreads over geometries; the reference's "
                                           "own NumPy path measured 35.6 ray-solves/s on one core of the build container "
                                           "(BASELINE.md section 2) and cannot travel to the GPU box",
                                   "sample": f"{n} of {G} geometries, evenly strided, {dt:.1f} s, oracle C port "
                                             f"(glibc libm), E={E}, K=N={K}"}
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps(out), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
