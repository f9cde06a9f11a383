"""GPU, two or more devices on the box (skipped otherwise): one process per GPU over NCCL, as bench.py --gpus N runs --
block-cyclic shards of a geometry grid traced by the CUDA path on each rank, all-gathered over NCCL (NVLink), and compared
on every rank with the whole grid recomputed on that rank alone; plus the contiguous rule and the two-point table gather.
Run it with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, os.environ["TOFRT_ROOT"])
    from ray_tracing_ultrasound_b200 import TofEngine, sharding, reference_setup as rs
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    x_a, z_a = rs.transducer(num_elements=128)
    alpha, zf = rs.launch_angles(), rs.focus_line()
    eng = TofEngine(rank).setup(alpha, zf, x_a[64], z_a[64], x_a)
    radii, offsets = rs.synthetic_grid(96, 80)           # 7 680 geometries, rows of 80
    G = len(radii)
    full = eng.tof_table_cuda(torch.from_numpy(radii).to(dev), torch.from_numpy(offsets).to(dev))
    ok = True
    # block-cyclic shards (one radius row per block), what bench.py uses for config 4
    r_s, o_s = sharding.shard_geometries_interleaved(radii, offsets, world, rank, block=80)
    mine = eng.tof_table_cuda(torch.from_numpy(r_s).to(dev), torch.from_numpy(o_s).to(dev))
    t_all, k_all = sharding.gather_table_interleaved(mine.tof, mine.k_hit, G, 80)
    ok &= bool(torch.equal(t_all, full.tof) and torch.equal(k_all, full.k_hit))
    # contiguous shards
    r_c, o_c = sharding.shard_geometries(radii, offsets, world, rank)
    mine = eng.tof_table_cuda(torch.from_numpy(r_c).to(dev), torch.from_numpy(o_c).to(dev))
    t_all, k_all = sharding.gather_table(mine.tof, mine.k_hit, G)
    ok &= bool(torch.equal(t_all, full.tof) and torch.equal(k_all, full.k_hit))
    ok &= int((full.k_hit >= 0).sum()) > 1000
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_OK" if int(flag.item()) == 1 else "MULTI_GPU_MISMATCH", world)
    dist.destroy_process_group()
""")


def test_sharded_tables_gathered_over_nccl_equal_the_single_gpu_table(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs on the box")
    world = 2 if n < 4 else 4
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, TOFRT_ROOT=ROOT)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stderr[-3000:]
    assert f"MULTI_GPU_OK {world}" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
