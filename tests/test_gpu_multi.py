"""GPU parity of the several-GPU paths (csrc/dist_multi.cu) against the single-GPU path, byte for byte.

* fpm_multi_* (one process, one host thread per GPU) runs on any box: with one GPU the same device is listed several
  times, which exercises the grid, the block uploads, the strided result copies and the hit merge just the same.
* fpm_dist_sharded_dev (one process per GPU, NCCL) needs two devices: NCCL refuses two ranks on one GPU.  Run it with
  `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`.
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from util import dirty_dna, mutate, random_dna, sorted_sketch_panel

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _panels(seed, n_ref, n_qry, s):
    rng = np.random.default_rng(seed)
    rh, rs = sorted_sketch_panel(rng, n_ref, s, n_clusters=6)
    qh, qs = sorted_sketch_panel(rng, n_qry, s, n_clusters=6)
    qh[:5] = rh[:5]; qs[:5] = rs[:5]
    rs[7] = 0; qs[9] = 0; qs[n_qry - 1] = 0; rs[n_ref - 1] = 0            # empty sketches, also in the last block
    rl = rng.integers(1000, 6_000_000, size=n_ref).astype(np.uint64)
    ql = rng.integers(1000, 6_000_000, size=n_qry).astype(np.uint64)
    return (rh, rs, rl), (qh, qs, ql)


@pytest.mark.parametrize("n_gpu", [2, 3, 4, 8])
@pytest.mark.parametrize("n_ref,n_qry,s", [(700, 450, 200), (97, 1301, 64)])
def test_multi_dist_tile_and_hits_equal_single_gpu(ctx, fpm, n_gpu, n_ref, n_qry, s):
    ref, qry = _panels(n_ref + n_gpu, n_ref, n_qry, s)
    have = fpm.device_count()
    multi = fpm.Multi(devices=[i % have for i in range(n_gpu)])
    try:
        assert multi.size() == n_gpu
        want, _ = ctx.dist_tile(ref, qry, s, 21, 4.0 ** 21, raw=True)
        got, _ = multi.dist_tile(ref, qry, s, 21, 4.0 ** 21, raw=True)
        assert want.tobytes() == got.tobytes()
        for max_d, max_p in ((0.3, 1.0), (1.0, 1e-10), (1.0, 1.0)):
            wh = ctx.dist_hits(ref, qry, s, 21, 4.0 ** 21, max_distance=max_d, max_pvalue=max_p, raw=True)
            gh = multi.dist_hits(ref, qry, s, 21, 4.0 ** 21, max_distance=max_d, max_pvalue=max_p, raw=True, capacity=64)   # forces the retry protocol
            assert len(wh) == len(gh) and wh.tobytes() == gh.tobytes()
    finally:
        multi.close()


def test_multi_sketch_batch_equals_single_gpu(ctx, fpm):
    rng = np.random.default_rng(31)
    base = random_dna(rng, 120000)
    groups = [[base], [mutate(rng, base, 0.02)], [dirty_dna(rng, 50000)], [random_dna(rng, 20)], [random_dna(rng, 300000)],
              [random_dna(rng, 7000), random_dna(rng, 9000)], [b"ACGT"], [random_dna(rng, 150000)], [random_dna(rng, 40000)]]
    seq = b"".join(b"".join(r + b"\0" for r in g) for g in groups)
    offs = np.cumsum([0] + [sum(len(r) + 1 for r in g) for g in groups]).astype(np.uint64)
    p = fpm.make_sketch_params(k=21, s=1000, want_counts=True)
    want = ctx.sketch_batch(np.frombuffer(seq, dtype=np.uint8), offs, p)
    have = fpm.device_count()
    for n_gpu in (2, 4, 16):
        multi = fpm.Multi(devices=[i % have for i in range(n_gpu)])
        try:
            got = multi.sketch_batch(np.frombuffer(seq, dtype=np.uint8), offs, p)
            assert np.array_equal(got["n"], want["n"])
            for g, n in enumerate(want["n"]):          # (slots beyond a sketch's size are unspecified)
                assert np.array_equal(got["hashes"][g, :n], want["hashes"][g, :n]) and np.array_equal(got["counts"][g, :n], want["counts"][g, :n]), g
        finally:
            multi.close()


WORKER = r"""
import os, sys
root = %(root)r
for p in (os.path.join(root, "fp-mash_b200", "py"), os.path.join(root, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch, torch.distributed as dist
import fpmash_b200 as fpm
from fpmash_b200 import sharding as sh
from util import sorted_sketch_panel
rank, world = int(sys.argv[1]), %(world)d
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=world, device_id=dev)
ctx = fpm.Context(rank)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
sh.init_comm(ctx)
for case, (n_r, n_q, s, mode) in enumerate([(1500, 1500, 200, {}), (333, 2047, 96, {}), (1500, 1500, 200, {"no_prune": True}), (5, 3, 50, {}),
                                            (1500, 1500, 200, {"p2p": 1}), (333, 2047, 96, {"p2p": 1}), (1500, 1500, 200, {"group": True})]):
    # the exchange step: all-gather + keep the block (default for panels of this size) or grouped send/receive of just the block's shards
    os.environ["FPMASH_EXCHANGE"] = "p2p" if mode.pop("p2p", 0) else "allgather"
    rng = np.random.default_rng(100 + case %% 4)             # same panels on every rank; each keeps only its row shards
    rh, rs = sorted_sketch_panel(rng, n_r, s, n_clusters=8)
    qh, qs = sorted_sketch_panel(rng, n_q, s, n_clusters=8)
    rs[n_r - 1] = 0; qs[0] = 0
    rl = rng.integers(1000, 6_000_000, size=n_r).astype(np.uint64); ql = rng.integers(1000, 6_000_000, size=n_q).astype(np.uint64)
    T = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a.view(np.int32)).to(dev)
    r0, r1 = sh.shard_range(n_r, rank, world); q0, q1 = sh.shard_range(n_q, rank, world)
    d = {k: T(np.ascontiguousarray(v)) for k, v in dict(rh=rh[r0:r1], rs=rs[r0:r1], rl=rl[r0:r1], qh=qh[q0:q1], qs=qs[q0:q1], ql=ql[q0:q1]).items()}
    blk = sh.block_of(rank, world, n_q, n_r)
    cap = (blk[1] - blk[0]) * (blk[3] - blk[2])
    out = torch.zeros(max(cap, 1) * 24, dtype=torch.uint8, device=dev)
    ctx.set_dist_mode(**mode)
    got_blk = ctx.dist_sharded_dev((d["rh"].data_ptr(), d["rs"].data_ptr(), d["rl"].data_ptr(), r1 - r0, s), n_r,
                                   (d["qh"].data_ptr(), d["qs"].data_ptr(), d["ql"].data_ptr(), q1 - q0, s), n_q, s, 21, 4.0 ** 21, out.data_ptr(), cap)
    torch.cuda.synchronize()
    assert got_blk == blk, (got_blk, blk)
    mine = out.cpu().numpy()[:cap * 24].view(fpm.PAIR_DTYPE).reshape(blk[1] - blk[0], blk[3] - blk[2])
    # hits mode over the same call
    hcap = max(cap, 1)
    hout = torch.zeros(hcap * 32, dtype=torch.uint8, device=dev)
    nh, hb = ctx.dist_hits_sharded_dev((d["rh"].data_ptr(), d["rs"].data_ptr(), d["rl"].data_ptr(), r1 - r0, s), n_r,
                                       (d["qh"].data_ptr(), d["qs"].data_ptr(), d["ql"].data_ptr(), q1 - q0, s), n_q, s, 21, 4.0 ** 21, hout.data_ptr(), hcap,
                                       max_distance=0.2)
    hits = hout.cpu().numpy()[:nh * 32].view(fpm.HIT_DTYPE)
    gathered = [None] * world
    dist.all_gather_object(gathered, (blk, mine, hits))
    if rank == 0:
        full = sh.assemble_blocks([(b, m) for b, m, _ in gathered], n_q, n_r, fpm.PAIR_DTYPE)
        want, _ = ctx.dist_tile((rh, rs, rl), (qh, qs, ql), s, 21, 4.0 ** 21, raw=True)      # the single-GPU result
        assert full.tobytes() == want.tobytes(), "case %%d: sharded matrix differs from the single-GPU matrix" %% case
        wh = ctx.dist_hits((rh, rs, rl), (qh, qs, ql), s, 21, 4.0 ** 21, max_distance=0.2, raw=True)
        allh = np.concatenate([h for _, _, h in gathered])
        allh = allh[np.lexsort((allh["ref"], allh["query"]))]
        assert allh.tobytes() == wh.tobytes(), "case %%d: sharded hits differ" %% case
    ctx.set_dist_mode()
    dist.barrier()
ctx.comm_destroy()
dist.destroy_process_group()
print("ok", rank)
"""


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_dist_over_nccl_equals_single_gpu(fpm, tmp_path, world):
    if fpm.device_count() < world:
        pytest.skip("needs %d GPUs (NCCL refuses two ranks on one device); run under gpurun --gpus %d" % (world, world))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "w.py"
    script.write_text(WORKER % {"root": ROOT, "port": port, "world": world})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(world)]
    for p in procs:
        out, err = p.communicate(timeout=600)
        assert p.returncode == 0 and "ok" in out, err[-3000:]


READS_WORKER = r"""
import os, sys
root = %(root)r
for p in (os.path.join(root, "fp-mash_b200", "py"), os.path.join(root, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch, torch.distributed as dist
import fpmash_b200 as fpm
from fpmash_b200 import sharding as sh
from util import mutate, random_dna
rank, world = int(sys.argv[1]), %(world)d
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=world, device_id=dev)
ctx = fpm.Context(rank)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
sh.init_comm(ctx)
rng = np.random.default_rng(12)                      # the same read set on every rank; each keeps a contiguous part
genome = random_dna(rng, 60000)
reads = []
for _ in range(14000):                               # ~35x of 150 bp reads with errors, both strands, a few N
    p = int(rng.integers(0, len(genome) - 150))
    r = mutate(rng, genome[p:p + 150], 0.01)
    if rng.random() < 0.5:
        r = r[::-1].translate(bytes.maketrans(b"ACGT", b"TGCA"))
    if rng.random() < 0.02:
        r = r[:70] + b"N" + r[71:]
    reads.append(r)
tiny = [random_dna(rng, 400)]                        # fewer windows than sketch slots: the accept-all path
for case, (recs, k, s, m) in enumerate([(reads, 21, 1000, 2), (reads, 21, 100, 2), (reads, 21, 100, 1), (reads, 21, 300, 3), (reads, 16, 200, 2), (tiny, 21, 1000, 1),
                                        (reads[:3], 21, 1000, 2)]):
    cut = [len(recs) * r // world for r in range(world + 1)]
    if case == 6:
        cut = [0] + [len(recs)] * world                # everything on rank 0, the other ranks hold nothing
    mine = b"".join(x + b"\0" for x in recs[cut[rank]:cut[rank + 1]])
    buf = torch.zeros(max(len(mine), 16) + 64, dtype=torch.uint8, device=dev)
    if mine:
        buf[:len(mine)] = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(dev)
    prm = fpm.make_sketch_params(k=k, s=s, min_cov=m, want_counts=True)
    oh = torch.zeros(s, dtype=torch.int64, device=dev); oc = torch.zeros(s, dtype=torch.int32, device=dev)
    on = torch.zeros(1, dtype=torch.int32, device=dev); ok = torch.zeros(1, dtype=torch.int64, device=dev)
    ctx.sketch_reads_sharded_dev(buf.data_ptr(), len(mine), prm, oh.data_ptr(), oc.data_ptr(), on.data_ptr(), ok.data_ptr())
    torch.cuda.synchronize()
    n = int(on.item())
    want = ctx.sketch_records([recs], k=k, s=s, min_cov=m, want_counts=True, want_kmers=True)[0]      # the whole read set on this GPU alone
    assert n == len(want["hashes"]), (case, rank, n, len(want["hashes"]))
    assert np.array_equal(oh.cpu().numpy().view(np.uint64)[:n], want["hashes"]), (case, rank, "hashes")
    assert np.array_equal(oc.cpu().numpy().view(np.uint32)[:n], want["counts"]), (case, rank, "counts")
    assert int(ok.item()) == want["kmers"], (case, rank, "windows")
    dist.barrier()
ctx.comm_destroy()
dist.destroy_process_group()
print("ok", rank)
"""


@pytest.mark.parametrize("world", [2, 4])
def test_read_set_over_several_gpus_equals_single_gpu(fpm, tmp_path, world):
    """One read set (mash sketch -r, -m 1..3, counts with the reference's order-dependent top count) spread over the ranks:
    hashes, multiplicities and window count identical to the single-GPU sketch of the whole stream."""
    if fpm.device_count() < world:
        pytest.skip("needs %d GPUs; run under gpurun --gpus %d" % (world, world))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "w.py"
    script.write_text(READS_WORKER % {"root": ROOT, "port": port, "world": world})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(world)]
    for p in procs:
        out, err = p.communicate(timeout=600)
        assert p.returncode == 0 and "ok" in out, err[-3000:]
