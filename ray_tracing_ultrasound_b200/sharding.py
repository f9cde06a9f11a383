"""Geometry sharding across ranks (SURVEY.md section 8 e).

Every geometry (r_outer, pipe_offset) is independent, so the G axis is split across ranks with no
collective on the data path.  Two rules:

* contiguous slices (`shard_bounds`, `shard_geometries`, `gather_table`): a shard is one block of the
  geometry-major table and the gather is a plain concatenation;
* block-cyclic / interleaved (`interleaved_index`, `shard_geometries_interleaved`,
  `gather_table_interleaved`): blocks of `block` consecutive geometries (one radius row of a grid, say) are
  dealt round-robin, rank r taking blocks r, r + W, r + 2W, ...  The cost of a geometry depends on its radius
  (how many rays reach the lens again), so contiguous radius slabs leave the ranks unevenly loaded (measured:
  3.6 % at 8 GPUs); dealing rows round-robin gives every rank the same mix.  Each rank's output stays one
  contiguous [G_r][E] block; the gather scatters the blocks back to grid order.

The only collective is the optional all-gather of the table shards when a consumer wants the whole table on
every rank.  Works with the NCCL backend (CUDA tensors, the product) and with gloo (CPU tensors, the tests).
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


# ---- block-cyclic (interleaved) sharding ----------------------------------------------------------
def interleaved_rows(n_rows, world_size, rank):
    """Rows (blocks) rank, rank + W, rank + 2W, ... of n_rows."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return np.arange(int(rank), int(n_rows), int(world_size), dtype=np.int64)


def interleaved_index(G, world_size, rank, block):
    """Flat geometry indices of this rank under the block-cyclic rule, in the order the rank stores them."""
    G, block = int(G), int(block)
    if block < 1:
        raise ValueError("block must be positive")
    rows = interleaved_rows((G + block - 1) // block, world_size, rank)
    idx = (rows[:, None] * block + np.arange(block, dtype=np.int64)[None, :]).ravel()
    return idx[idx < G]


def shard_geometries_interleaved(radii, offsets, world_size, rank, block):
    idx = interleaved_index(len(radii), world_size, rank, block)
    return np.ascontiguousarray(radii[idx]), np.ascontiguousarray(offsets[idx])


def gather_table_interleaved(tof_shard, k_hit_shard, G, block, group=None):
    """All-gather block-cyclic [G_r][E] shards into the full [G][E] table (grid order) on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    E = tof_shard.shape[1]
    index = [interleaved_index(G, world, r, block) for r in range(world)]
    max_rows = max(len(ix) for ix in index)

    def gather(shard):
        pad = torch.zeros((max_rows, E), dtype=shard.dtype, device=shard.device)
        pad[: shard.shape[0]] = shard
        out = torch.empty((world * max_rows, E), dtype=shard.dtype, device=shard.device)
        dist.all_gather_into_tensor(out, pad, group=group)
        full = torch.empty((G, E), dtype=shard.dtype, device=shard.device)
        for r, ix in enumerate(index):
            full[torch.from_numpy(ix).to(shard.device)] = out[r * max_rows: r * max_rows + len(ix)]
        return full

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
