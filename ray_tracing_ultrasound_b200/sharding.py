"""Geometry sharding across ranks (SURVEY.md section 8 e).

Every geometry (r_outer, pipe_offset) is independent, so the G axis is cut into contiguous
slices, one per rank, with no collective on the data path.  The only collective is the optional
all-gather of the table shards when a consumer wants the whole table on every rank; the table is
geometry-major, so a shard is one contiguous block and the gather is a plain concatenation.
Works with the NCCL backend (CUDA tensors, the product) and with gloo (CPU tensors, the tests).
"""
import numpy as np


def shard_bounds(G, world_size, rank):
    """Contiguous, balanced split: the first G % world_size ranks get one extra geometry."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(int(G), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_geometries(radii, offsets, world_size, rank):
    lo, hi = shard_bounds(len(radii), world_size, rank)
    return np.ascontiguousarray(radii[lo:hi]), np.ascontiguousarray(offsets[lo:hi])


def gather_table(tof_shard, k_hit_shard, G, group=None):
    """All-gather the [G_r][E] shards into the full [G][E] table on every rank.

    Shards may differ by one row (G not divisible by the world size): they are padded to the
    largest shard for `all_gather_into_tensor` and the padding is dropped afterwards."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    E = tof_shard.shape[1]
    rows = [shard_bounds(G, world, r) for r in range(world)]
    max_rows = max(hi - lo for lo, hi in rows)

    def gather(shard):
        pad = torch.zeros((max_rows, E), dtype=shard.dtype, device=shard.device)
        pad[: shard.shape[0]] = shard
        out = torch.empty((world * max_rows, E), dtype=shard.dtype, device=shard.device)
        dist.all_gather_into_tensor(out, pad, group=group)
        return torch.cat([out[r * max_rows: r * max_rows + (hi - lo)] for r, (lo, hi) in enumerate(rows)])

    return gather(tof_shard), gather(k_hit_shard)


# ---- the two-point solves: element-major tables [E][F], focal points / voxels sharded ----
def shard_points(points, world_size, rank):
    """Contiguous slice of each coordinate array (fx, fz) or (fx, fy, fz) for this rank."""
    lo, hi = shard_bounds(len(points[0]), world_size, rank)
    return tuple(np.ascontiguousarray(p[lo:hi]) for p in points)


def gather_point_table(shard, F, group=None):
    """All-gather the [E][F_r] shards of a two-point table (tofrt_fermat*_dev output) into [E][F] on
    every rank.  The table is element-major, so the shards interleave: gather the transposed,
    point-major blocks (contiguous per rank) and transpose back."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    E = shard.shape[0]
    cols = [shard_bounds(F, world, r) for r in range(world)]
    max_cols = max(hi - lo for lo, hi in cols)
    pad = torch.zeros((max_cols, E), dtype=shard.dtype, device=shard.device)
    pad[: shard.shape[1]] = shard.t()
    out = torch.empty((world * max_cols, E), dtype=shard.dtype, device=shard.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    full = torch.cat([out[r * max_cols: r * max_cols + (hi - lo)] for r, (lo, hi) in enumerate(cols)])
    return full.t().contiguous()
