"""CPU, world_size 2, gloo: the host-side sharding logic (bounds, shard -> gather round trip).
The table rows here come from the oracle (the CUDA path needs a GPU); what is tested is that a
sharded table gathered back equals the unsharded one, bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ray_tracing_ultrasound_b200 import reference_setup as rs, sharding


def test_shard_bounds_cover_everything():
    for G in (0, 1, 5, 210, 1025):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(G, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == G
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)


def _worker(rank, world, port, G, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import Oracle
    o = Oracle(device_math=True)
    p = o.params()
    x_a, z_a = rs.transducer()
    alpha, zf = rs.launch_angles(200), rs.focus_line(200)
    radii, offsets = rs.synthetic_grid(G, 1)
    r_s, o_s = sharding.shard_geometries(radii, offsets, world, rank)
    tof, k_hit, _ = o.table(p, r_s, o_s, x_a[32], z_a[32], zf, alpha, x_a)
    tof_full, k_full = sharding.gather_table(torch.from_numpy(np.ascontiguousarray(tof.T)),
                                             torch.from_numpy(np.ascontiguousarray(k_hit.T)), G)
    if rank == 0:
        tof1, k1, _ = o.table(p, radii, offsets, x_a[32], z_a[32], zf, alpha, x_a)
        ok = np.array_equal(tof_full.numpy(), tof1.T) and np.array_equal(k_full.numpy(), k1.T)
        out.put(bool(ok) and tof_full.shape == (G, len(x_a)))
    dist.destroy_process_group()


@pytest.mark.parametrize("G", [8, 7])
def test_shard_gather_roundtrip_gloo_world2(G):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, G, out)) for r in range(2)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(120)
        assert pr.exitcode == 0
    assert out.get() is True


def test_interleaved_index_partitions_everything():
    for G, block in ((0, 4), (1, 4), (210, 21), (1000, 7), (4096 * 8, 4096)):
        for world in (1, 2, 3, 8):
            parts = [sharding.interleaved_index(G, world, r, block) for r in range(world)]
            allidx = np.concatenate(parts) if parts else np.empty(0, np.int64)
            assert np.array_equal(np.sort(allidx), np.arange(G))
            if G >= block * world:
                assert max(map(len, parts)) - min(map(len, parts)) <= block
            for r, ix in enumerate(parts):      # whole blocks, dealt round-robin
                assert np.array_equal((ix // block) % world, np.full(len(ix), r))
    assert np.array_equal(sharding.interleaved_rows(10, 4, 1), [1, 5, 9])


def _worker_interleaved(rank, world, port, G, block, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import Oracle
    o = Oracle(device_math=True)
    p = o.params()
    x_a, z_a = rs.transducer()
    alpha, zf = rs.launch_angles(200), rs.focus_line(200)
    radii, offsets = rs.synthetic_grid(G // 3 + 1, 3)
    radii, offsets = radii[:G], offsets[:G]
    r_s, o_s = sharding.shard_geometries_interleaved(radii, offsets, world, rank, block)
    tof, k_hit, _ = o.table(p, r_s, o_s, x_a[32], z_a[32], zf, alpha, x_a)
    tof_full, k_full = sharding.gather_table_interleaved(torch.from_numpy(np.ascontiguousarray(tof.T)),
                                                         torch.from_numpy(np.ascontiguousarray(k_hit.T)), G, block)
    if rank == 0:
        tof1, k1, _ = o.table(p, radii, offsets, x_a[32], z_a[32], zf, alpha, x_a)
        ok = np.array_equal(tof_full.numpy(), tof1.T) and np.array_equal(k_full.numpy(), k1.T)
        out.put(bool(ok) and tof_full.shape == (G, len(x_a)))
    dist.destroy_process_group()


@pytest.mark.parametrize("G,block", [(12, 3), (11, 2), (5, 4)])
def test_interleaved_shard_gather_roundtrip_gloo_world2(G, block):
    """Block-cyclic shards (what bench.py --gpus N uses for config 4) gathered back equal the unsharded table."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker_interleaved, args=(r, 2, port, G, block, out)) for r in range(2)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(120)
        assert pr.exitcode == 0
    assert out.get() is True


def _worker3d(rank, world, port, F, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import Oracle
    o = Oracle(device_math=True)
    p = o.params()
    rng = np.random.default_rng(5)
    ang = np.linspace(-rs.ALPHA_MAX, rs.ALPHA_MAX, 9)
    ex, ey = rng.uniform(-0.009, 0.009, 6), rng.uniform(-0.009, 0.009, 6)
    pts = (rng.uniform(-0.03, 0.03, F), rng.uniform(-0.03, 0.03, F), rng.uniform(0.0, 0.05, F))
    mine = sharding.shard_points(pts, world, rank)
    tof = o.fermat3d(p, ang, ex, ey, float(rs.D), *mine)[0]
    full = sharding.gather_point_table(torch.from_numpy(tof), F)
    if rank == 0:
        ref = o.fermat3d(p, ang, ex, ey, float(rs.D), *pts)[0]
        out.put(bool(np.array_equal(full.numpy(), ref)) and full.shape == (6, F))
    dist.destroy_process_group()


@pytest.mark.parametrize("F", [10, 9])
def test_voxel_shards_gather_roundtrip_gloo_world2(F):
    """The two-point tables are element-major [E][F]: voxel shards interleave on gathering."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker3d, args=(r, 2, port, F, out)) for r in range(2)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(120)
        assert pr.exitcode == 0
    assert out.get() is True
