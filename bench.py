#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native fp-mash hot path.

Metric (BASELINE.json): Gk-mers/sec sketched; sketch-pairs/sec dist.
Workload at N GPUs: configs[1] "mash sketch k=21 s=1000 canonical over 1,000 synthetic 5 Mbp
bacterial-size genomes" PER GPU (sketching shards by input file, no collective: weak scaling),
plus -- reported in the same JSON line under "dist" and, compactly, as the LAST key "dist_summary" --
configs[2] "all-vs-all mash dist of 20,000 sketches (k=21 s=1000)" cut into query x reference blocks over
the ranks by the library itself (fpm_dist_sharded_dev: grouped ncclSend/ncclRecv of the row shards, then
every rank compares its block; strong scaling: the 4e8 pairs are fixed), and under "configs" the other
BASELINE configs at their named sizes (C1 fp mode, C4 read set with -m 2, C5 k=32 s=10000).
Parity is checked inside the run: CPU-reference sketches == GPU sketches, CPU-reference dist rows == the
device-path output, pruned == merge-all, and several GPUs == one GPU (digests + rank 0's block byte for byte).

A "step" = one pass of the hot path over one batch: fpm_sketch_batch_dev over all genomes of the
rank (value, inputs resident in HBM) / fpm_sketch_batch with pinned HOST buffers (e2e, H2D and
D2H inside the timed region).  Inputs (5 GB per GPU) are far larger than the 126 MB L2, so no
explicit L2 flush is needed between iterations.

`--impl reference` times the reference's own CPU implementation (oracle/_ref: its hash.cpp,
MurmurHash3.cpp, MinHashHeap.cpp ... compiled unmodified, driven by the restated sketchFile
loop with a -p = all-cores thread pool) on a bounded sample of the same workload.

PyTorch is plumbing here (device buffers, RNG for synthetic genomes, torch.distributed); every
kernel in the timed regions is this repo's own, reached through the C ABI.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "fp-mash_b200", "py"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

K = 21
S = 1000
W_INT32_OPS = {21: 74, 32: 96, 16: 64}   # int32-op equivalents of Murmur per k-mer
NCU_SKETCH_TRAFFIC_RATIO = (526.58 + 23.66) / 500.0005   # DRAM bytes per algorithmic byte, ncu --set full capture of sketch_hash_kernel_v2<21,true> (profiles/r02_sketch_hash_v4.txt: 100 x 5 Mbp launch)
NCU_DIST_TRAFFIC_RATIO = (27.22 + 193.05) / (3200 * 3200 * 24 / 1e6 + 2 * 3200 * 1001 * 4 / 1e6)   # same for dist_tile32_kernel (profiles/r01_dist_tile32_v4.txt)
IMAD_WIDE_RATE = 8.99 / 18.45         # IMAD.WIDE issue rate relative to IMAD, measured (profiles/ubench/int_mix.cu: 8.99 vs 18.45 T/s)
SMEM_BYTES_PER_CLK_PER_SM = 128        # one 32-lane x 4-byte wavefront per clock (B300_MICROARCH.md / measured LSU pipe limit)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genomes", type=int, default=1000, help="genomes per GPU (config: 1000)")
    ap.add_argument("--genome-len", type=int, default=5_000_000, help="bases per genome (config: 5 Mbp)")
    ap.add_argument("--dist-sketches", type=int, default=20000, help="all-vs-all dist panel (config: 20000); 0 skips dist")
    ap.add_argument("--cpu-genomes", type=int, default=512, help="genomes in the CPU-baseline sample (512 x 5 Mbp = ~2.2 s on 16 cores per pass)")
    ap.add_argument("--cpu-dist-queries", type=int, default=256, help="query rows in the CPU dist sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the compact C1 / C4 / C5 measurements")
    ap.add_argument("--c5-queries", type=int, default=100000, help="C5 dist: query sketches (config: 100000); 0 skips C5 dist")
    ap.add_argument("--c5-refs", type=int, default=10000, help="C5 dist: reference sketches (config: 10000)")
    ap.add_argument("--c5-merge-all", type=int, default=1, help="C5 dist: also time the same call with every pair merged (1e13 merge steps, a few seconds)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------------------------
def gen_genomes(torch, n, length, device, rank):
    """n genomes of `length` bases: 20 iid-uniform ancestors; genome g = ancestor g%20 with iid
    substitutions at rate 0.001*(1+g//20).  Layout = the C-ABI batch layout: records back to back,
    each followed by one 0x00 byte; one record (= one FASTA file) per sketch."""
    stride = length + 1
    buf = torch.zeros(n * stride, dtype=torch.uint8, device=device)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    gen = torch.Generator(device=device)
    anc = []
    for a in range(min(20, n)):
        gen.manual_seed(7_000_000 + 1000 * rank + a)
        anc.append(torch.randint(0, 4, (length,), generator=gen, device=device, dtype=torch.uint8))
    for g in range(n):
        gen.manual_seed(1_000_000 * (rank + 1) + g)
        rate = min(0.05, 0.001 * (1 + g // 20))
        mask = torch.rand(length, generator=gen, device=device) < rate
        sub = torch.randint(0, 4, (length,), generator=gen, device=device, dtype=torch.uint8)
        codes = torch.where(mask, sub, anc[g % len(anc)])
        buf[g * stride:g * stride + length] = lut[codes.long()]
    import numpy as np
    offsets = (np.arange(n + 1, dtype=np.uint64) * np.uint64(stride))
    return buf, offsets


def gen_sketch_panel(torch, n, s, device, seed, n_clusters=20, shared=0.6):
    """n sorted duplicate-free u64 sketches: bottom-s of (cluster core subset U private hashes),
    values uniform below 2^64*s/4999980 so magnitudes match real bottom-s of 5 Mbp genomes."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    hi = int((1 << 64) * s / 4_999_980)
    big = (1 << 62)
    cores = torch.randint(0, hi, (n_clusters, 2 * s), generator=gen, device=device, dtype=torch.int64)
    out = torch.empty((n, s), dtype=torch.int64, device=device)
    chunk = 2000
    for i0 in range(0, n, chunk):
        i1 = min(n, i0 + chunk)
        m = i1 - i0
        cl = (torch.arange(i0, i1, device=device) % n_clusters)
        core = cores[cl]                                                       # [m][2s]
        keep = torch.rand((m, 2 * s), generator=gen, device=device) < shared
        core = torch.where(keep, core, torch.full_like(core, big))
        priv = torch.randint(0, hi, (m, s), generator=gen, device=device, dtype=torch.int64)
        cand, _ = torch.sort(torch.cat([core, priv], dim=1), dim=1)
        dup = torch.zeros_like(cand, dtype=torch.bool)
        dup[:, 1:] = cand[:, 1:] == cand[:, :-1]
        cand = torch.where(dup, torch.full_like(cand, big), cand)
        cand, _ = torch.sort(cand, dim=1)
        out[i0:i1] = cand[:, :s]
    assert bool((out < big).all()), "panel generation produced a short sketch"
    return out


# ----------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        busy = [x for x in sm if mx and x > 0.3 * mx] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU code on a bounded sample
# ----------------------------------------------------------------------------------------------
def cpu_sketch_baseline(host_seq, length, n_genomes, threads):
    """oracle/_ref (reference TUs) sketching n_genomes genomes with `threads` pool threads."""
    import numpy as np
    from oracle_py import RefLib
    if not RefLib.available():
        from oracle_py import build
        build(ref=True)
    if not RefLib.available():
        return None
    ref = RefLib()
    stride = length + 1
    sample = np.array(host_seq[:n_genomes * stride], copy=True)          # the reference upper-cases in place
    offsets = np.arange(n_genomes + 1, dtype=np.uint64) * np.uint64(stride)
    t0 = time.perf_counter()
    ref.sketch_batch(sample, offsets, K, S, seed=42, use64=True, threads=threads)
    dt = time.perf_counter() - t0
    windows = n_genomes * (length - K + 1)
    return windows / dt / 1e9, dt


def run_reference(args):
    import numpy as np
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    n = min(args.cpu_genomes, args.genomes)
    rng = np.random.default_rng(12345)
    stride = args.genome_len + 1
    host = np.zeros(n * stride, dtype=np.uint8)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    for g in range(n):
        host[g * stride:g * stride + args.genome_len] = lut[rng.integers(0, 4, size=args.genome_len, dtype=np.uint8)]
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_sketch_baseline(host, args.genome_len, n, threads)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libmashref.so is not built"}))
            return 0
        if i >= args.warmup:
            vals.append(r)
    total_t = sum(dt for _, dt in vals)
    value = n * (args.genome_len - K + 1) * len(vals) / total_t / 1e9
    sample = "%d of %d genomes x %d bp per step, reference TUs (hash.cpp/MinHashHeap.cpp) + restated sketchFile loop, -p %d" % (
        n, args.genomes, args.genome_len, threads)
    line = {
        "impl": "reference", "metric": "Gk-mers/sec sketched (mash sketch k=21 s=1000 canonical)", "value": value,
        "unit": "Gk-mers/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total_t / len(vals), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        # the same config object as our arm prints (the bounded sample is described in cpu_baseline.sample)
        "config": {"workload": "mash sketch -k 21 -s 1000 canonical over %d synthetic %.1f Mbp genomes per GPU (BASELINE configs[1])" % (args.genomes, args.genome_len / 1e6),
                   "genomes_per_gpu": args.genomes, "genome_len": args.genome_len, "k": K, "sketch_size": S, "seed": 42,
                   "l2": "inputs (%.1f GB per GPU) exceed the 126 MB L2; no flush needed" % (args.genomes * (args.genome_len + 1) / 1e9),
                   "parallelism": "files sharded over %d GPU(s), no collective" % int(os.environ.get("WORLD_SIZE", "1"))},
        "cpu_baseline": {"value": value, "unit": "Gk-mers/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "Gk-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def gen_panel_rows(torch, row0, row1, s, device, seed, n_clusters, shared=0.6, chunk=500, core_seed=None):
    """Rows [row0, row1) of a clustered panel of sorted duplicate-free u64 sketches (same construction as gen_sketch_panel),
    generated chunk by chunk with per-chunk seeds so that every rank can produce exactly its own rows."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed if core_seed is None else core_seed)     # panels built on the same cores are related cluster by cluster
    hi = int((1 << 64) * s / 4_999_980)
    big = 1 << 62
    cores = torch.randint(0, hi, (n_clusters, 2 * s), generator=gen, device=device, dtype=torch.int64)
    out = torch.empty((row1 - row0, s), dtype=torch.int64, device=device)
    for c in range(row0 // chunk, (row1 + chunk - 1) // chunk):
        i0, i1 = c * chunk, (c + 1) * chunk
        gen.manual_seed(seed * 1_000_003 + c)
        m = i1 - i0
        core = cores[torch.arange(i0, i1, device=device) % n_clusters]
        keep = torch.rand((m, 2 * s), generator=gen, device=device) < shared
        core = torch.where(keep, core, torch.full_like(core, big))
        cand, _ = torch.sort(torch.cat([core, torch.randint(0, hi, (m, s), generator=gen, device=device, dtype=torch.int64)], dim=1), dim=1)
        dup = torch.zeros_like(cand, dtype=torch.bool)
        dup[:, 1:] = cand[:, 1:] == cand[:, :-1]
        cand, _ = torch.sort(torch.where(dup, torch.full_like(cand, big), cand), dim=1)
        lo, hi_r = max(i0, row0), min(i1, row1)
        out[lo - row0:hi_r - row0] = cand[lo - i0:hi_r - i0, :s]
        del core, keep, cand, dup
    return out


def pair_digest(torch, out_u8, n_pairs):
    """Order-independent digest of fpm_pair records in HBM: wrap-around int64 sums of numer, denom (without the pass flag),
    the distance bit patterns and the p-value bit patterns.  Sums over blocks add up to the sum over the whole matrix."""
    if n_pairs == 0:
        return torch.zeros(4, dtype=torch.int64, device=out_u8.device)
    w = out_u8[:n_pairs * 24].view(torch.int32).view(n_pairs, 6)
    d = out_u8[:n_pairs * 24].view(torch.int64).view(n_pairs, 3)
    return torch.stack([w[:, 0].sum(dtype=torch.int64), (w[:, 1] & 0x7fffffff).sum(dtype=torch.int64), d[:, 1].sum(), d[:, 2].sum()])


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import fpmash_b200 as fpm
    from fpmash_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the fp-mash B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    # stdout carries ONE JSON line: libraries (NCCL prints its version banner to fd 1) get stderr for the rest of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(t):
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    ctx = fpm.Context(local_rank)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    if world > 1:
        sharding.init_comm(ctx)            # the library's own NCCL communicator (unique id of rank 0 broadcast over torch.distributed)
    params = fpm.make_sketch_params(k=K, s=S, seed=42)
    threads = os.cpu_count() or 1

    # ---- sketch: value (HBM resident) --------------------------------------------------------
    n, L = args.genomes, args.genome_len
    seq, offsets = gen_genomes(torch, n, L, device, rank)
    windows_per_step = n * (L - K + 1)
    out_h = torch.zeros((n, S), dtype=torch.int64, device=device)
    out_n = torch.zeros(n, dtype=torch.int32, device=device)

    def step_resident():
        ctx.sketch_batch_dev(seq.data_ptr(), seq.numel(), offsets, params, out_h.data_ptr(), None, out_n.data_ptr())

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    barrier()
    ctx.set_timing(True)
    launches0 = ctx.launch_count()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launch_count() - launches0
    hash_ms, hash_n = ctx.get_timing(fpm.KERNEL_SKETCH_HASH)
    sel_ms, sel_n = ctx.get_timing(fpm.KERNEL_SKETCH_SELECT)
    ctx.set_timing(False)
    ms_per_step = ms_total / args.steps
    value = world * windows_per_step / (ms_per_step * 1e-3) / 1e9
    assert int(out_n.min().item()) == S, "sketches are not full"

    # ---- sketch: e2e (pinned host buffers through the host-pointer C-ABI call) ----------------
    pinned_seq = fpm.PinnedBuffer(seq.numel())      # page-locked, on the NUMA node of this rank's GPU
    host_seq = pinned_seq.array
    torch.from_numpy(host_seq).copy_(seq)
    torch.cuda.synchronize()
    e2e_res = None
    for _ in range(max(1, args.warmup)):
        e2e_res = ctx.sketch_batch(host_seq, offsets, params)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_res = ctx.sketch_batch(host_seq, offsets, params)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * windows_per_step * args.steps / e2e_s / 1e9
    clocks = sampler.stop() if rank == 0 else None   # sampled across the resident and the end-to-end timed regions
    gpu_sketches = out_h.cpu().numpy().view(np.uint64)
    assert np.array_equal(e2e_res["hashes"], gpu_sketches), "e2e and resident sketches differ"
    h2d = int(seq.numel() + offsets.nbytes)
    d2h = int(n * S * 8 + n * 4)

    # ---- rooflines of the dominant kernel ------------------------------------------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    hash_ms_avg = hash_ms / max(hash_n, 1)
    alg_bytes = float(seq.numel())                                    # 1 B per base read once
    achieved_gbs = alg_bytes / (hash_ms_avg * 1e-3) / 1e9
    int_peak = ctx.int32_peak()                       # max of the three below
    int_peaks = ctx.int32_peaks()                     # ALU pipe only, FMA pipe (IMAD) only, alternating
    int_ops = windows_per_step * W_INT32_OPS[K]
    int_achieved = int_ops / (hash_ms_avg * 1e-3)
    n_sm = torch.cuda.get_device_properties(device).multi_processor_count

    # ---- CPU baseline (rank 0, N=1 only): the reference's own TUs on a bounded sample, outputs compared with the GPU's ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle_py import RefLib
        cg = min(args.cpu_genomes, n)
        if not RefLib.available():
            from oracle_py import build
            build(ref=True)
        if RefLib.available():
            stride = L + 1
            sample = np.array(host_seq[:cg * stride], copy=True)          # the reference upper-cases in place
            soff = np.arange(cg + 1, dtype=np.uint64) * np.uint64(stride)
            t0 = time.perf_counter()
            ch, _, cn = RefLib().sketch_batch(sample, soff, K, S, seed=42, use64=True, threads=threads)
            dt = time.perf_counter() - t0
            cpu = {"value": cg * (L - K + 1) / dt / 1e9, "unit": "Gk-mers/s", "cores": threads, "kind": "reference",
                   "sample": "%d of %d genomes (%d Mbp), reference TUs hash.cpp/MurmurHash3.cpp/MinHashHeap.cpp + restated sketchFile loop, -p %d, %.1f s" % (cg, n, cg * L // 1_000_000, threads, dt),
                   "matches_gpu": bool(np.array_equal(ch, gpu_sketches[:cg]) and (cn == S).all())}
            del sample
    del pinned_seq, host_seq, e2e_res

    # ---- dist (configs[2]): all-vs-all of nd sketches, query x reference blocks over the ranks (library call) -------------
    dist_obj, dist_summary = None, None
    kspace = 4.0 ** K
    if args.dist_sketches > 0:
        nd = args.dist_sketches
        panel = gen_sketch_panel(torch, nd, S, device, seed=3)
        sizes = torch.full((nd,), S, dtype=torch.int32, device=device)
        lengths = torch.full((nd,), 5_000_000, dtype=torch.int64, device=device)
        s0, s1 = sharding.shard_range(nd, rank, world)                  # this rank's row shard of the panel (queries = references)
        blk = sharding.block_of(rank, world, nd, nd)
        blk_pairs = (blk[1] - blk[0]) * (blk[3] - blk[2])
        out_pairs = torch.empty((max(blk_pairs, 1) * 24,), dtype=torch.uint8, device=device)
        steps_ctr = torch.zeros(1, dtype=torch.int64, device=device)
        shard = (panel[s0:s1].data_ptr(), sizes[s0:s1].data_ptr(), lengths[s0:s1].data_ptr(), s1 - s0, S)
        full_ptrs = (panel.data_ptr(), sizes.data_ptr(), lengths.data_ptr(), nd, S)

        def step_dist():
            if world > 1:   # exchange step (row shards -> blocks, grouped ncclSend/ncclRecv) + this rank's block, one library call
                ctx.dist_sharded_dev(shard, nd, shard, nd, S, K, kspace, out_pairs.data_ptr(), blk_pairs, steps_ctr.data_ptr())
            else:
                ctx.dist_tile_dev(full_ptrs, full_ptrs, S, K, kspace, out_pairs.data_ptr(), steps_ctr.data_ptr())

        dsteps = max(3, min(args.steps, 10))
        pairs = nd * nd

        def timed_dist(**mode):
            """dsteps timed passes of the all-vs-all dist in the given kernel mode."""
            ctx.set_dist_mode(**mode)
            try:
                for _ in range(2):
                    step_dist()
                barrier()
                steps_ctr.zero_()
                ctx.set_timing(True)
                l0 = ctx.launch_count()
                e0.record()
                for _ in range(dsteps):
                    step_dist()
                e1.record()
                barrier()
                ms = max_over_ranks(e0.elapsed_time(e1)) / dsteps
                tile_ms, tile_n = ctx.get_timing(fpm.KERNEL_DIST_TILE)
                lit_ms, lit_n = ctx.get_timing(fpm.KERNEL_DIST_LITERAL)
                pack_ms, pack_n = ctx.get_timing(fpm.KERNEL_DIST_PACK)
                ex_ms, ex_n = ctx.get_timing(fpm.KERNEL_DIST_EXCHANGE)
                ctx.set_timing(False)
                st = sum_over_ranks(steps_ctr.clone())
                return {"ms": ms, "tile": max_over_ranks(tile_ms / dsteps), "tile_n": tile_n, "lit_n": lit_n, "pack": max_over_ranks(pack_ms / dsteps),
                        "exchange": max_over_ranks(ex_ms / dsteps), "merge_steps": int(st.item()) / dsteps, "launches": ctx.launch_count() - l0,
                        "digest": [int(x) for x in sum_over_ranks(pair_digest(torch, out_pairs, blk_pairs)).tolist()]}
            finally:
                ctx.set_dist_mode()

        full = timed_dist(no_prune=True)        # every pair merged: the merge kernel's own throughput and rooflines
        run = timed_dist()                      # the product path: pairs without a shared hash are answered without a merge
        # parity inside the bench: (1) pruned == merge-all, (2) several GPUs == one GPU (rank 0 recomputes the whole matrix alone)
        parity = {"pruned_equals_merge_all": run["digest"] == full["digest"]}
        if world > 1:
            ok = torch.ones(1, dtype=torch.int64, device=device)
            if rank == 0:
                alone = torch.empty((pairs * 24,), dtype=torch.uint8, device=device)
                ctx.dist_tile_dev(full_ptrs, full_ptrs, S, K, kspace, alone.data_ptr(), None)
                torch.cuda.synchronize()
                ok[0] = int([int(x) for x in pair_digest(torch, alone, pairs).tolist()] == run["digest"])
                # and record for record on this rank's own block
                mine = alone.view(pairs, 24).view(nd, nd, 24)[blk[0]:blk[1], blk[2]:blk[3]].reshape(-1)
                ok[0] &= int(torch.equal(mine, out_pairs[:blk_pairs * 24]))
                del alone, mine
            dist.broadcast(ok, src=0)
            parity["sharded_equals_single_gpu"] = bool(ok.item())
        d_ms = run["ms"]
        smem_peak = n_sm * SMEM_BYTES_PER_CLK_PER_SM * 1.965e9
        qp, rp = sharding.grid_shape(world, nd, nd)
        dist_obj = {
            "metric": "sketch-pairs/sec dist (all-vs-all, k=21 s=1000)", "value": pairs / (d_ms * 1e-3), "unit": "pairs/s", "ms_per_step": d_ms, "steps": dsteps,
            "scaling": "strong", "workload": "all-vs-all mash dist of %d sketches (BASELINE configs[2])" % nd,
            "sharding": "%d x %d blocks of query x reference rows over the ranks (fpm_dist_sharded_dev)" % (qp, rp) if world > 1 else "single GPU",
            "gpu_launches": run["launches"], "literal_launches": run["lit_n"], "merge_steps_per_pair": run["merge_steps"] / pairs,
            "kernel_ms": {"exchange": run["exchange"], "prepass": run["pack"], "fill+tiles": run["tile"]},
            "merge_all_pairs": {"value": pairs / (full["ms"] * 1e-3), "ms_per_step": full["ms"], "merge_steps_per_pair": full["merge_steps"] / pairs,
                                "kernel_ms": {"exchange": full["exchange"], "prepass": full["pack"], "tiles": full["tile"]},
                                "roofline_smem": {"achieved": full["merge_steps"] / world * 8 / (full["tile"] * 1e-3) / 1e12, "peak": smem_peak / 1e12, "unit": "TB/s",
                                                  "frac": full["merge_steps"] / world * 8 / (full["tile"] * 1e-3) / smem_peak},
                                "roofline_int": {"frac": full["merge_steps"] / world * 3 / (full["tile"] * 1e-3) / int_peak}},
            "parity": parity, "digest": run["digest"],
        }
        # e2e on a stated sample of query rows (host panels in, 24-byte records out), and the CPU reference on a sample of the DEVICE output
        if rank == 0 and world == 1:
            qs_n = min(nd, 8192)
            hp = panel.cpu().pin_memory().numpy().view(np.uint64)            # pinned host panels (the contract's "inputs from pinned host memory")
            hs = np.full(nd, S, dtype=np.uint32)
            hl = np.full(nd, 5_000_000, dtype=np.uint64)
            pinned = torch.empty(qs_n * nd * 24, dtype=torch.uint8, pin_memory=True).numpy().view(fpm.PAIR_DTYPE).reshape(qs_n, nd)
            ctx.dist_tile((hp, hs, hl), (hp[:qs_n], hs[:qs_n], hl[:qs_n]), S, K, kspace, out=pinned, raw=True)
            t0 = time.perf_counter()
            got, _ = ctx.dist_tile((hp, hs, hl), (hp[:qs_n], hs[:qs_n], hl[:qs_n]), S, K, kspace, out=pinned, raw=True)
            dt = time.perf_counter() - t0
            dist_obj["e2e"] = {"value": qs_n * nd / dt, "unit": "pairs/s", "sample": "%d query rows x %d refs" % (qs_n, nd),
                               "h2d_bytes_per_step": int((nd + qs_n) * (S * 8 + 12)), "d2h_bytes_per_step": int(qs_n * nd * 24)}
            cap = 64 << 20
            hit_buf = torch.empty(cap * 32, dtype=torch.uint8, pin_memory=True).numpy().view(fpm.HIT_DTYPE)
            try:
                php = hp
                ctx.dist_hits((php, hs, hl), (php, hs, hl), S, K, kspace, max_distance=0.25, out=hit_buf, raw=True)
                t0 = time.perf_counter()
                hits = ctx.dist_hits((php, hs, hl), (php, hs, hl), S, K, kspace, max_distance=0.25, out=hit_buf, raw=True)
                dt = time.perf_counter() - t0
                sub = hits[hits["query"] < qs_n]
                same = bool(np.array_equal(np.nonzero(got["distance"] <= 0.25)[1], sub["ref"]) and np.array_equal(got["numer"][got["distance"] <= 0.25], sub["numer"]))
                dist_obj["e2e_filtered"] = {"value": nd * nd / dt, "unit": "pairs/s", "filter": "-d 0.25", "hits": int(len(hits)), "seconds": dt,
                                            "d2h_bytes_per_step": int(len(hits) * 32), "agrees_with_matrix_rows": same}
            except fpm.FpmError as e:
                dist_obj["e2e_filtered"] = {"error": str(e)}
            del hit_buf
            if not args.no_cpu:
                from oracle_py import RefLib
                if RefLib.available():
                    cq = min(args.cpu_dist_queries, nd)
                    rows = np.linspace(0, nd - 1, cq).astype(np.int64)          # spread over the clusters and the grouped tile order
                    t0 = time.perf_counter()
                    cn, cd, cdist = RefLib().dist_batch(hp, hs, hp[rows], hs[rows], S, K, threads=threads)
                    dt = time.perf_counter() - t0
                    # the DEVICE path's output (grouped panels, prefilled records, tile list): rows of out_pairs left by the last timed step
                    devrows = out_pairs.view(nd, nd * 24)[torch.from_numpy(rows).to(device)].cpu().numpy().view(fpm.PAIR_DTYPE).reshape(cq, nd)
                    ok = (np.array_equal(cn.reshape(cq, nd), devrows["numer"]) and np.array_equal(cd.reshape(cq, nd), devrows["denom"] & 0x7fffffff)
                          and np.allclose(cdist.reshape(cq, nd), devrows["distance"], rtol=1e-12, atol=0))
                    dist_obj["cpu_baseline"] = {"value": cq * nd / dt, "unit": "pairs/s", "cores": threads, "kind": "reference",
                                                "sample": "%d query rows x %d refs (compareSketches over the reference's HashList, <=4096-pair chunks, no p-value), compared with the device-path output" % (cq, nd),
                                                "matches_gpu": bool(ok)}
        dist_summary = {"pairs_per_s": dist_obj["value"], "ms": d_ms, "exchange_ms": run["exchange"], "prepass_ms": run["pack"], "fill_tiles_ms": run["tile"],
                        "merge_all_pairs_per_s": pairs / (full["ms"] * 1e-3), "merge_all_ms": full["ms"], "roofline_smem": dist_obj["merge_all_pairs"]["roofline_smem"]["frac"],
                        "grid": "%dx%d" % (qp, rp), "parity": parity}
        del panel, out_pairs

    # ---- the other BASELINE configs, compact (C1 fp mode, C4 read set -m 2, C5 k=32 s=10000) --------------------------------
    configs = run_other_configs(args, torch, np, fpm, sharding, ctx, device, rank, world, seq, offsets, n, L, barrier, max_over_ranks, sum_over_ranks, threads) if not args.no_configs else None

    if rank == 0:
        line = {
            "metric": "Gk-mers/sec sketched (mash sketch k=21 s=1000 canonical)", "value": value, "unit": "Gk-mers/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "mash sketch -k 21 -s 1000 canonical over %d synthetic %.1f Mbp genomes per GPU (BASELINE configs[1])" % (n, L / 1e6),
                       "genomes_per_gpu": n, "genome_len": L, "k": K, "sketch_size": S, "seed": 42,
                       "l2": "inputs (%.1f GB per GPU) exceed the 126 MB L2; no flush needed" % (seq.numel() / 1e9),
                       "parallelism": "files sharded over %d GPU(s), no collective" % world},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Gk-mers/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "roofline": {"bound": "int32-alu", "achieved": int_achieved / 1e12, "peak": int_peak / 1e12, "unit": "Tint32-op/s", "frac": int_achieved / int_peak,
                         "traffic": alg_bytes * NCU_SKETCH_TRAFFIC_RATIO, "kernel": "sketch_hash_kernel_v2<21,true>", "launch_ms": hash_ms_avg,
                         "share_of_step": hash_ms / ms_total if ms_total else None,
                         "ops_per_kmer": W_INT32_OPS[K],
                         "peak_source": "measured live: alternating IMAD/LOP3, 16 chains (fpm_measure_int32_peak); SURVEY 8d: the sketch kernel is integer-ALU bound",
                         "peaks": {"alu_lop3": int_peaks[0] / 1e12, "fma_imad": int_peaks[1] / 1e12, "alternating": int_peaks[2] / 1e12},
                         "hbm": {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak, "peak_source": hbm_src,
                                 "note": "1 B per base read once; dram traffic 1.10x (ncu profiles/r02_sketch_hash_v4.txt)"}},
            "kernel_ms": {"sketch_hash": hash_ms_avg, "sketch_select": sel_ms / max(sel_n, 1)},
            "cpu_baseline": cpu,
            "dist": dist_obj,
            "configs": configs,
            "dist_summary": dist_summary,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        ctx.comm_destroy()
        dist.destroy_process_group()
    return 0


def run_other_configs(args, torch, np, fpm, sharding, ctx, device, rank, world, seq, offsets, n, L, barrier, max_over_ranks, sum_over_ranks, threads):
    """BASELINE.json configs[0], [3], [4] at their named sizes (compact numbers; the headline stays configs[1])."""
    out = {}
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, reps):
        fn()
        barrier()
        ev0.record()
        for _ in range(reps):
            fn()
        ev1.record()
        barrier()
        return max_over_ranks(ev0.elapsed_time(ev1)) / reps

    # -- C5 sketch shape: k=32 s=10000 over this rank's genomes (the same 1000 x 5 Mbp buffer; a rate, so the sample size does not matter)
    g5 = min(n, 200)
    p5 = fpm.make_sketch_params(k=32, s=10000)
    oh5 = torch.zeros((g5, 10000), dtype=torch.int64, device=device)
    on5 = torch.zeros(g5, dtype=torch.int32, device=device)
    ms = timed(lambda: ctx.sketch_batch_dev(seq.data_ptr(), int(offsets[g5]), offsets[:g5 + 1], p5, oh5.data_ptr(), None, on5.data_ptr()), 3)
    out["C5_sketch"] = {"Gkmers_per_s": world * g5 * (L - 31) / (ms * 1e-3) / 1e9, "ms": ms, "k": 32, "s": 10000, "genomes_per_gpu": g5, "full": bool((on5 == 10000).all().item())}
    del oh5, on5

    # -- C4: 1,000,000 reads x 150 bp (30x of a 5 Mbp genome), 1 % substitutions, 0.1 % N, both strands, -m 2 with counts: ONE sketch, one GPU
    if True:
        gen = torch.Generator(device=device)
        gen.manual_seed(4)
        GL, n_reads, rl = 5_000_000, 1_000_000, 150
        genome = torch.randint(0, 4, (GL,), generator=gen, device=device, dtype=torch.uint8)
        start = torch.randint(0, GL - rl, (n_reads,), generator=gen, device=device)
        codes = genome[start[:, None] + torch.arange(rl, device=device)[None, :]]
        rev = torch.rand(n_reads, generator=gen, device=device) < 0.5
        codes = torch.where(rev[:, None], 3 - codes.flip(1), codes)
        err = torch.rand((n_reads, rl), generator=gen, device=device) < 0.01
        codes = torch.where(err, torch.randint(0, 4, (n_reads, rl), generator=gen, device=device, dtype=torch.uint8), codes)
        reads = lut[codes.long()]
        reads[torch.rand((n_reads, rl), generator=gen, device=device) < 0.001] = ord("N")
        buf = torch.zeros((n_reads, rl + 1), dtype=torch.uint8, device=device)
        buf[:, :rl] = reads
        buf = buf.reshape(-1)
        del reads, codes, err, rev, start
        offs = np.array([0, buf.numel()], dtype=np.uint64)
        p4 = fpm.make_sketch_params(k=21, s=1000, min_cov=2, want_counts=True)
        oh = torch.zeros((1, 1000), dtype=torch.int64, device=device)
        oc = torch.zeros((1, 1000), dtype=torch.int32, device=device)
        on = torch.zeros(1, dtype=torch.int32, device=device)
        ok = torch.zeros(1, dtype=torch.int64, device=device)
        # one read set = ONE sketch: on several GPUs every rank takes a contiguous range of reads and the library merges the
        # per-rank candidate tables (fpm_sketch_reads_sharded_dev); every rank ends up with the complete sketch
        rd0, rd1 = sharding.shard_range(n_reads, rank, world)
        mine = buf[rd0 * (rl + 1):rd1 * (rl + 1)].clone()

        def c4_step(kmers_ptr=None):
            if world > 1:
                ctx.sketch_reads_sharded_dev(mine.data_ptr(), mine.numel(), p4, oh.data_ptr(), oc.data_ptr(), on.data_ptr(), kmers_ptr)
            else:
                ctx.sketch_batch_dev(buf.data_ptr(), buf.numel(), offs, p4, oh.data_ptr(), oc.data_ptr(), on.data_ptr(), kmers_ptr)

        c4_step(ok.data_ptr())
        windows = int(ok.item())
        barrier()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            c4_step()
        torch.cuda.synchronize()
        ms = max_over_ranks((time.perf_counter() - t0) / reps * 1e3)
        c4 = {"Gkmers_per_s": windows / (ms * 1e-3) / 1e9, "ms": ms, "reads": n_reads, "valid_windows": windows, "min_cov": 2, "gpus": world, "full": int(on.item()) == 1000}
        if world > 1:    # several GPUs == one GPU: rank 0 sketches the whole read set alone and compares hashes and multiplicities
            same = torch.ones(1, dtype=torch.int64, device=device)
            if rank == 0:
                oh1 = torch.zeros_like(oh); oc1 = torch.zeros_like(oc); on1 = torch.zeros_like(on)
                ctx.sketch_batch_dev(buf.data_ptr(), buf.numel(), offs, p4, oh1.data_ptr(), oc1.data_ptr(), on1.data_ptr())
                torch.cuda.synchronize()
                same[0] = int(torch.equal(oh1, oh) and torch.equal(oc1, oc) and torch.equal(on1, on))
            dist_ = __import__("torch.distributed").distributed
            dist_.broadcast(same, src=0)
            c4["sharded_equals_single_gpu"] = bool(same.item())
        if world == 1 and not args.no_cpu:
            from oracle_py import RefLib
            if RefLib.available():
                host = buf.cpu().numpy()
                hoff = np.array([0, host.size], dtype=np.uint64)
                t0 = time.perf_counter()
                ch, cc, cn = RefLib().sketch_batch(host, hoff, 21, 1000, seed=42, use64=True, min_cov=2, threads=1, want_counts=True)
                dt = time.perf_counter() - t0
                c4["cpu_reference_Gkmers_per_s"] = windows / dt / 1e9           # read mode is single-threaded in the reference (Sketch.cpp:203-210)
                c4["matches_cpu_reference"] = bool(np.array_equal(ch[0], oh.cpu().numpy().view(np.uint64)[0]) and np.array_equal(cc[0], oc.cpu().numpy().view(np.uint32)[0]))
        out["C4_reads"] = c4
        del buf, mine

    if rank == 0:

        # -- C1: fp mode, 5 ids x 2000 fingerprint lines (1..20 tokens each), 32-bit hashes, literal comparison of the unsorted lists
        rng = np.random.default_rng(1)
        lines = [list(rng.integers(1, 101, size=int(rng.integers(1, 21)))) for _ in range(5 * 2000)]
        t0 = time.perf_counter()
        fh = ctx.fp_hash_batch(lines, seed=42, use64=False)
        hp = np.ascontiguousarray(fh.reshape(5, 2000))
        got, _ = ctx.dist_tile((hp, np.full(5, 2000, np.uint32), np.full(5, 10000, np.uint64)), (hp, np.full(5, 2000, np.uint32), np.full(5, 10000, np.uint64)),
                               2000, 1, 10.0, sorted_unique=False)
        c1 = {"ms": (time.perf_counter() - t0) * 1e3, "lines": len(lines), "pairs": 25, "self_pairs_distance_0": bool((np.diag(got["distance"]) == 0).all())}
        out["C1_fp"] = c1

        # -- ingestion of a gzip'ed collection (every input of the reference goes through zlib, Sketch.cpp:1340-1346): the headline
        # genomes as 1000 .fna.gz files, inflated by fpm_gunzip_batch (one warp per file) and parsed + sketched without the inflated bytes
        # visiting the host, next to zlib on the host threads.  (the first 32 genomes, compressed here, repeated: a stream's rate is what it is.)
        try:
            import ctypes as C
            import gzip
            import zlib
            from concurrent.futures import ThreadPoolExecutor
            n_gz = 1000
            m_gz = min(32, n)
            host_seq = seq[:int(offsets[m_gz])].cpu().numpy()
            plain = []
            for i in range(m_gz):
                g = host_seq[int(offsets[i]):int(offsets[i + 1]) - 1].tobytes()
                plain.append(b">genome%d\n" % i + b"\n".join(g[j:j + 80] for j in range(0, len(g), 80)) + b"\n")
            with ThreadPoolExecutor(threads) as ex:
                gz_m = list(ex.map(lambda b: gzip.compress(b, 6), plain))
            gz = [gz_m[i % m_gz] for i in range(n_gz)]
            raw_bytes = sum(len(plain[i % m_gz]) for i in range(n_gz))
            t0 = time.perf_counter()
            with ThreadPoolExecutor(threads) as ex:
                inflated = sum(ex.map(lambda b: len(zlib.decompress(b, 31)), gz[:256]))
            host_gbs = inflated / (time.perf_counter() - t0) / 1e9
            goff = np.cumsum(np.array([0] + [len(b) for b in gz], dtype=np.uint64)).astype(np.uint64)
            blob = torch.frombuffer(bytearray(b"".join(gz)), dtype=torch.uint8).pin_memory()
            ends = np.zeros(n_gz, dtype=np.uint64)
            total, status = C.c_uint64(0), C.c_int(0)
            ctx.set_timing(True)
            t0 = time.perf_counter()
            fpm._check(fpm.lib.fpm_gunzip_batch(ctx._h, blob.data_ptr(), goff.ctypes.data, n_gz, ends.ctypes.data, C.byref(total), C.byref(status)))
            t_inflate = time.perf_counter() - t0
            k_ms, k_n = ctx.get_timing(6)
            ctx.set_timing(False)
            parsed = ctx.fasta_parse_resident(total.value, fetch_sequence=False, fetch_headers=True) if status.value == 0 else None
            same = False
            t_all = None
            if parsed is not None:
                recs, lengths, _, headers = parsed
                goffs = np.append(recs["seq_begin"], np.uint64(int(recs["seq_begin"][-1]) + int(lengths[-1]) + 1)).astype(np.uint64)
                res = ctx.sketch_parsed(goffs, fpm.make_sketch_params(k=K, s=S))
                t_all = time.perf_counter() - t0
                # the same genomes sketched from the resident buffer of the headline run
                oh = torch.zeros((m_gz, S), dtype=torch.int64, device=device)
                on = torch.zeros(m_gz, dtype=torch.int32, device=device)
                ctx.sketch_batch_dev(seq.data_ptr(), int(offsets[m_gz]), offsets[:m_gz + 1], fpm.make_sketch_params(k=K, s=S), oh.data_ptr(), None, on.data_ptr())
                want = oh.cpu().numpy().view(np.uint64)
                same = bool(len(headers) == n_gz and headers[m_gz + 1] == b">genome1" and all(np.array_equal(res["hashes"][i], want[i % m_gz]) for i in range(0, n_gz, 37)))
            out["gz_ingest"] = {"files": n_gz, "compressed_GB": float(goff[-1]) / 1e9, "inflated_GB": raw_bytes / 1e9, "status": status.value,
                                "gunzip_kernel_ms": k_ms, "kernel_launches": k_n, "inflate_GBps": raw_bytes / (k_ms * 1e-3) / 1e9 if k_ms else None,
                                "upload_plus_inflate_s": t_inflate, "upload_inflate_parse_sketch_s": t_all,
                                "Gkmers_per_s_from_gz_files": (n_gz * (L - K + 1) / t_all / 1e9) if t_all else None,
                                "host_zlib_GBps": host_gbs, "host_threads": threads, "sketches_equal_plain_input": same}
            del blob
        except Exception as e:                                        # a side measurement: never takes the bench line down
            out["gz_ingest"] = {"error": repr(e)[:200]}

    # -- C5 dist at full size: 100,000 queries x 10,000 references, s=10000, k=32 (1e9 pairs), blocks over the ranks
    if args.c5_queries > 0:
        nq, nr, s5 = args.c5_queries, args.c5_refs, 10000
        q0, q1 = sharding.shard_range(nq, rank, world)
        r0, r1 = sharding.shard_range(nr, rank, world)
        qp_ = gen_panel_rows(torch, q0, q1, s5, device, 52, 100, core_seed=5)     # 100 clusters shared by both panels: a query is
        rp_ = gen_panel_rows(torch, r0, r1, s5, device, 51, 100, core_seed=5)     # related to the 1 % of references of its cluster
        qs = torch.full((q1 - q0,), s5, dtype=torch.int32, device=device); ql = torch.full((q1 - q0,), 5_000_000, dtype=torch.int64, device=device)
        rs = torch.full((r1 - r0,), s5, dtype=torch.int32, device=device); rlen = torch.full((r1 - r0,), 5_000_000, dtype=torch.int64, device=device)
        blk = sharding.block_of(rank, world, nq, nr)
        bp = (blk[1] - blk[0]) * (blk[3] - blk[2])
        o5 = torch.empty((max(bp, 1) * 24,), dtype=torch.uint8, device=device)
        st5 = torch.zeros(1, dtype=torch.int64, device=device)
        rptr = (rp_.data_ptr(), rs.data_ptr(), rlen.data_ptr(), r1 - r0, s5)
        qptr = (qp_.data_ptr(), qs.data_ptr(), ql.data_ptr(), q1 - q0, s5)

        def c5_step():
            if world > 1:
                ctx.dist_sharded_dev(rptr, nr, qptr, nq, s5, 32, 4.0 ** 32, o5.data_ptr(), bp, st5.data_ptr())
            else:
                ctx.dist_tile_dev(rptr, qptr, s5, 32, 4.0 ** 32, o5.data_ptr(), st5.data_ptr())

        def c5_timed(**mode):
            ctx.set_dist_mode(**mode)
            try:
                c5_step()
                barrier()
                st5.zero_()
                ev0.record()
                c5_step()
                ev1.record()
                barrier()
                sec = max_over_ranks(ev0.elapsed_time(ev1)) * 1e-3
                return sec, int(sum_over_ranks(st5.clone()).item()) / (nq * nr), [int(x) for x in sum_over_ranks(pair_digest(torch, o5, bp)).tolist()]
            finally:
                ctx.set_dist_mode()

        sec, spp, dg = c5_timed()
        out["C5_dist"] = {"pairs_per_s": nq * nr / sec, "seconds": sec, "queries": nq, "refs": nr, "s": s5, "k": 32, "merge_steps_per_pair": spp, "digest": dg}
        if args.c5_merge_all:
            sec_a, spp_a, dg_a = c5_timed(no_prune=True)
            out["C5_dist"].update({"merge_all_seconds": sec_a, "merge_all_steps_per_s": spp_a * nq * nr / sec_a, "merge_all_equals_pruned": dg_a == dg})
    return out


if __name__ == "__main__":
    sys.exit(main())
