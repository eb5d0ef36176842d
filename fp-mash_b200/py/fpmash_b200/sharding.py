"""Multi-GPU partitioning of the hot path: one process per GPU.

* sketching shards by input file / sequence: no data-path collective (`assign_by_size`);
* dist shards the QUERY rows over ranks; the reference panel is replicated with ONE exchange
  step, an all-gather of each rank's panel shard (`all_gather_rows`, NCCL over NVLink on GPUs,
  gloo in the CPU tests); results stay sharded by query row.

torch.distributed is plumbing here; it is imported lazily so the single-GPU path needs no torch.
"""


def shard_range(n, rank, world):
    """Contiguous [lo, hi) of n items owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_bounds(n, world):
    return [shard_range(n, r, world) for r in range(world)]


def assign_by_size(sizes, world):
    """Size-balanced assignment of files to ranks (largest first onto the lightest rank).
    Returns per-rank lists of file indices, each ascending so per-rank output order is input
    order (the reference's ThreadPool returns sketches in submission order)."""
    loads = [0] * world
    owner = [0] * len(sizes)
    for i in sorted(range(len(sizes)), key=lambda j: (-sizes[j], j)):
        r = min(range(world), key=lambda q: (loads[q], q))
        owner[i] = r
        loads[r] += sizes[i]
    return [[i for i in range(len(sizes)) if owner[i] == r] for r in range(world)]


def all_gather_rows(local_rows, n_total, group=None):
    """All-gather a row-sharded 2-D tensor whose shards follow shard_range(n_total, rank, world).
    Uneven shards are padded to the largest shard for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    bounds = shard_bounds(n_total, world)
    width = max(hi - lo for lo, hi in bounds)
    if all(hi - lo == width for lo, hi in bounds):
        out = torch.empty((n_total,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
        dist.all_gather_into_tensor(out, local_rows.contiguous(), group=group)
        return out
    pad = torch.zeros((width,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
    pad[:local_rows.shape[0]] = local_rows
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:hi - lo] for p, (lo, hi) in zip(parts, bounds)], dim=0)
