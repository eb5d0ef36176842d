"""Host-side helpers for the one-process-per-GPU deployment of the hot path.

The partitioning itself lives in the library (csrc/dist_multi.cu: fpm_shard_range, fpm_dist_grid_shape, fpm_dist_block,
fpm_dist_sharded_dev); this module only does what the HOST program has to do around it:

* sketching shards by input file / sequence, no data-path collective (`assign_by_size`);
* `init_comm`: distribute the NCCL unique id of rank 0 over whatever process group the host already has
  (torch.distributed here; MPI or a file work the same) and create the library's communicator;
* `assemble_blocks`: put the ranks' result blocks back into the reference's query-major matrix
  (CommandDistance.cpp:355-359) -- used by the tests and by rank 0 when it wants the whole table.

torch.distributed is plumbing; it is imported lazily so the single-GPU path needs no torch.
"""
import numpy as np


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


def shard_range(n, part, parts):
    """Rows [n*part/parts, n*(part+1)/parts): the row shard `part` of `parts` (same rule as fpm_shard_range)."""
    return n * part // parts, n * (part + 1) // parts


def grid_shape(world, n_qry, n_ref):
    """q_parts x r_parts = world minimising n_qry/q_parts + n_ref/r_parts (same rule as fpm_dist_grid_shape)."""
    best, shape = None, (1, world)
    for q in range(1, world + 1):
        if world % q:
            continue
        cost = n_qry / q + n_ref / (world // q)
        if best is None or cost < best * (1 - 1e-12):
            best, shape = cost, (q, world // q)
    return shape


def block_of(rank, world, n_qry, n_ref):
    """(q_begin, q_end, r_begin, r_end) of the block rank `rank` compares (same rule as fpm_dist_block)."""
    qp, rp = grid_shape(world, n_qry, n_ref)
    qi, rj = divmod(rank, rp)
    return (shard_range(n_qry, qi * rp, world)[0], shard_range(n_qry, (qi + 1) * rp, world)[0],
            shard_range(n_ref, rj * qp, world)[0], shard_range(n_ref, (rj + 1) * qp, world)[0])


def senders_and_receivers(rank, world, n_qry, n_ref):
    """The exchange step as fpm_dist_sharded_dev performs it.  Returns {"q_send": [...], "r_send": [...], "q_recv": [...],
    "r_recv": [...]}: ranks this rank sends its query / reference row shard to, and ranks whose shards make up its blocks."""
    qp, rp = grid_shape(world, n_qry, n_ref)
    qi, rj = divmod(rank, rp)
    return {"q_send": [(rank // rp) * rp + d for d in range(rp)], "r_send": [d * rp + rank // qp for d in range(qp)],
            "q_recv": list(range(qi * rp, (qi + 1) * rp)), "r_recv": list(range(rj * qp, (rj + 1) * qp))}


def assemble_blocks(blocks, n_qry, n_ref, dtype):
    """blocks: iterable of ((q0, q1, r0, r1), array[q1-q0][r1-r0]) -> the full [n_qry][n_ref] query-major matrix."""
    out = np.zeros((n_qry, n_ref), dtype=dtype)
    seen = np.zeros((n_qry, n_ref), dtype=bool)
    for (q0, q1, r0, r1), a in blocks:
        assert not seen[q0:q1, r0:r1].any(), "blocks overlap"
        out[q0:q1, r0:r1] = np.asarray(a).reshape(q1 - q0, r1 - r0)
        seen[q0:q1, r0:r1] = True
    assert seen.all(), "blocks do not cover the pair space"
    return out


def init_comm(ctx, group=None):
    """Create ctx's NCCL communicator over the ranks of an initialised torch.distributed group: rank 0's unique id is
    broadcast as 128 bytes, then every rank calls fpm_comm_init_rank."""
    import torch
    import torch.distributed as dist
    import fpmash_b200 as fpm

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    on_gpu = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    t = torch.zeros(fpm.COMM_ID_BYTES, dtype=torch.uint8, device=dev)
    if rank == 0:
        t = torch.frombuffer(bytearray(fpm.comm_unique_id()), dtype=torch.uint8).to(dev)
    dist.broadcast(t, src=0, group=group)
    ctx.comm_init(bytes(t.cpu().numpy().tobytes()), rank, world)
    return rank, world
