#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native fp-mash hot path.

Metric (BASELINE.json): Gk-mers/sec sketched; sketch-pairs/sec dist.
Workload at N GPUs: configs[1] "mash sketch k=21 s=1000 canonical over 1,000 synthetic 5 Mbp
bacterial-size genomes" PER GPU (sketching shards by input file, no collective: weak scaling),
plus -- reported in the same JSON line under "dist" -- configs[2] "all-vs-all mash dist of
20,000 sketches (k=21 s=1000)", queries row-sharded over the ranks with an NCCL all-gather of
the reference panel (strong scaling: the 4e8 pairs are fixed).

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
NCU_SKETCH_TRAFFIC_RATIO = (315.19 + 16.98) / 300.0   # DRAM bytes per algorithmic byte, ncu capture of sketch_hash_kernel_v2 (profiles/r01_sketch_hash_v4.txt)
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
        "config": {"workload": "mash sketch -k 21 -s 1000 canonical over %d synthetic %.1f Mbp genomes per GPU (BASELINE configs[1])" % (args.genomes, args.genome_len / 1e6),
                   "genomes_per_gpu": args.genomes, "genome_len": args.genome_len, "k": K, "sketch_size": S, "seed": 42,
                   "cpu_sample_genomes_per_step": n},
        "cpu_baseline": {"value": value, "unit": "Gk-mers/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "Gk-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import fpmash_b200 as fpm

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

    ctx = fpm.Context(local_rank)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    params = fpm.make_sketch_params(k=K, s=S, seed=42)

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
    assert np.array_equal(e2e_res["hashes"], out_h.cpu().numpy().view(np.uint64)), "e2e and resident sketches differ"
    h2d = int(seq.numel() + offsets.nbytes)
    d2h = int(n * S * 8 + n * 4)

    # ---- roofline of the dominant kernel -----------------------------------------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    hash_ms_avg = hash_ms / max(hash_n, 1)
    alg_bytes = float(seq.numel())                                    # 1 B per base read once
    achieved_gbs = alg_bytes / (hash_ms_avg * 1e-3) / 1e9
    int_peak = ctx.int32_peak()                       # max of the three below
    int_peaks = ctx.int32_peaks()                     # ALU pipe only, FMA pipe (IMAD) only, alternating
    int_ops = windows_per_step * W_INT32_OPS[K]
    int_achieved = int_ops / (hash_ms_avg * 1e-3)

    # ---- dist (configs[2]) -------------------------------------------------------------------
    dist_obj = None
    if args.dist_sketches > 0:
        nd = args.dist_sketches
        panel = gen_sketch_panel(torch, nd, S, device, seed=3)
        sizes = torch.full((nd,), S, dtype=torch.int32, device=device)
        lengths = torch.full((nd,), 5_000_000, dtype=torch.int64, device=device)
        q0, q1 = rank * nd // world, (rank + 1) * nd // world
        r0, r1 = q0, q1                                            # this rank's shard of the reference panel
        ref_full = torch.empty_like(panel) if world > 1 else panel
        out_pairs = torch.empty(((q1 - q0) * nd * 24,), dtype=torch.uint8, device=device)
        steps_ctr = torch.zeros(1, dtype=torch.int64, device=device)
        kspace = 4.0 ** K

        def step_dist():
            if world > 1:   # the one exchange step of the path: all-gather the reference panel over NVLink
                dist.all_gather_into_tensor(ref_full, panel[r0:r1].contiguous())
            ctx.dist_tile_dev((ref_full.data_ptr(), sizes.data_ptr(), lengths.data_ptr(), nd, S),
                              (panel[q0:q1].data_ptr(), sizes[q0:q1].data_ptr(), lengths[q0:q1].data_ptr(), q1 - q0, S),
                              S, K, kspace, out_pairs.data_ptr(), steps_ctr.data_ptr())

        dsteps = max(2, min(args.steps, 3))
        pairs = nd * nd
        n_sm = torch.cuda.get_device_properties(device).multi_processor_count

        def timed_dist(**mode):
            """dsteps timed passes of the sharded all-vs-all dist in the given kernel mode."""
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
                ctx.set_timing(False)
                return {"ms": ms, "tile_avg": tile_ms / max(tile_n, 1), "tile_n": tile_n, "lit_n": lit_n, "pack_avg": pack_ms / max(pack_n, 1),
                        "merge_steps": int(steps_ctr.item()) / dsteps, "launches": ctx.launch_count() - l0}
            finally:
                ctx.set_dist_mode()

        full = timed_dist(no_prune=True)        # every pair merged: the merge kernel's own throughput and rooflines
        run = timed_dist()                      # the product path: pairs without a shared hash are answered without a merge
        d_ms, tile_avg, merge_steps = run["ms"], run["tile_avg"], run["merge_steps"]
        out_bytes = (q1 - q0) * nd * 24 + 2 * nd * (S + 1) * 4
        dist_obj = {
            "metric": "sketch-pairs/sec dist (all-vs-all, k=21 s=1000)", "value": pairs / (d_ms * 1e-3), "unit": "pairs/s",
            "ms_per_step": d_ms, "steps": dsteps, "scaling": "strong",
            "config": {"workload": "all-vs-all mash dist of %d sketches (BASELINE configs[2])" % nd, "sketches": nd,
                       "sketch_size": S, "k": K, "sharding": "query rows over ranks, NCCL all-gather of the reference panel" if world > 1 else "single GPU"},
            "gpu_launches": run["launches"], "fast_path_launches": run["tile_n"], "literal_launches": run["lit_n"],
            "merge_steps_per_pair": merge_steps / ((q1 - q0) * nd),
            "kernel_ms": {"dist_tile32": tile_avg, "rank_compress_index_mark": run["pack_avg"]},
            "pruning": "exact: the rank pre-pass sorts every hash of both panels, which is also an inverted index; a pair whose sketches share no hash has common = 0 and denom = min(s, |A|+|B|) and is answered without a merge; related sketches are grouped into the same tiles; merge_steps_per_pair counts executed steps only; merge_all_pairs below is the same call with every pair merged",
            "roofline": {"bound": "hbm", "achieved": out_bytes / (tile_avg * 1e-3) / 1e9 if run["tile_n"] else None, "peak": hbm_peak, "unit": "GB/s",
                         "frac": out_bytes / (tile_avg * 1e-3) / 1e9 / hbm_peak if run["tile_n"] else None, "traffic": None,
                         "note": "dist_fill_unshared_kernel + dist_tile32_kernel of the pruned run (both inside kernel_ms.dist_tile32): 24 B per pair streamed in matrix order for the pairs without a shared hash, merged pairs overwritten at their original positions, rank panels read once"},
            "merge_all_pairs": {
                "value": pairs / (full["ms"] * 1e-3), "unit": "pairs/s", "ms_per_step": full["ms"], "merge_steps_per_pair": full["merge_steps"] / ((q1 - q0) * nd),
                "kernel_ms": {"dist_tile32": full["tile_avg"], "rank_compress": full["pack_avg"]},
                "roofline_int": {"bound": "int32-alu", "achieved": full["merge_steps"] * 3 / (full["tile_avg"] * 1e-3) / 1e12, "peak": int_peak / 1e12, "unit": "Tint32-op/s",
                                 "frac": (full["merge_steps"] * 3 / (full["tile_avg"] * 1e-3)) / int_peak,
                                 "note": "algorithmic ops = merge steps x 3 (SURVEY.md 8d) / dist_tile32_kernel time; peak measured by fpm_measure_int32_peak"},
                "roofline_smem": {"bound": "shared-memory bandwidth", "achieved": full["merge_steps"] * 8 / (full["tile_avg"] * 1e-3) / 1e12,
                                  "peak": n_sm * SMEM_BYTES_PER_CLK_PER_SM * 1.965e9 / 1e12, "unit": "TB/s",
                                  "frac": (full["merge_steps"] * 8 / (full["tile_avg"] * 1e-3)) / (n_sm * SMEM_BYTES_PER_CLK_PER_SM * 1.965e9),
                                  "note": "what bounds the merge: every step is two 4-byte shared-memory loads per pair (one LDS.32 wavefront per list per warp); peak = SMs x 128 B/clk x 1965 MHz; ncu: shared-memory pipe 87 % busy incl. staging"},
                "roofline": {"bound": "hbm", "achieved": out_bytes / (full["tile_avg"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": out_bytes / (full["tile_avg"] * 1e-3) / 1e9 / hbm_peak,
                             "traffic": out_bytes * NCU_DIST_TRAFFIC_RATIO,
                             "traffic_note": "dram bytes of one ncu --set full capture of dist_tile32_kernel (profiles/r01_dist_tile32_v4.txt) scaled to this launch; below the algorithmic bytes because part of the output was still in L2 when the capture ended"},
            },
        }
        # e2e on a stated sample of query rows (host panels in, 24-byte records out)
        if rank == 0:
            qs = min(nd, 8192)
            hp = panel.cpu().numpy().view(np.uint64)
            hs = np.full(nd, S, dtype=np.uint32)
            hl = np.full(nd, 5_000_000, dtype=np.uint64)
            pinned = torch.empty(qs * nd * 24, dtype=torch.uint8, pin_memory=True).numpy().view(fpm.PAIR_DTYPE).reshape(qs, nd)
            ctx.dist_tile((hp, hs, hl), (hp[:qs], hs[:qs], hl[:qs]), S, K, kspace, out=pinned, raw=True)
            t0 = time.perf_counter()
            got, _ = ctx.dist_tile((hp, hs, hl), (hp[:qs], hs[:qs], hl[:qs]), S, K, kspace, out=pinned, raw=True)
            dt = time.perf_counter() - t0
            dist_obj["e2e"] = {"value": qs * nd / dt, "unit": "pairs/s", "sample": "%d query rows x %d refs, one GPU" % (qs, nd),
                               "h2d_bytes_per_step": int((nd + qs) * (S * 8 + 12)), "d2h_bytes_per_step": int(qs * nd * 24)}
            # the same through `mash dist -d 0.25` semantics: every pair of the full all-vs-all is decided on the GPU, only the
            # passing ones come back (fpm_dist_hits) -- host panels in, sorted hit records out
            cap = 64 << 20
            hit_buf = torch.empty(cap * 32, dtype=torch.uint8, pin_memory=True).numpy().view(fpm.HIT_DTYPE)
            try:
                php = panel.cpu().pin_memory().numpy().view(np.uint64)     # pinned host panels, as a caller that cares would hold them
                ctx.dist_hits((php, hs, hl), (php, hs, hl), S, K, kspace, max_distance=0.25, out=hit_buf, raw=True)
                ctx.set_timing(True)
                t0 = time.perf_counter()
                hits = ctx.dist_hits((php, hs, hl), (php, hs, hl), S, K, kspace, max_distance=0.25, out=hit_buf, raw=True)
                dt = time.perf_counter() - t0
                f_tile, f_pack = ctx.get_timing(fpm.KERNEL_DIST_TILE)[0], ctx.get_timing(fpm.KERNEL_DIST_PACK)[0]
                ctx.set_timing(False)
                sub = hits[hits["query"] < qs]
                same = bool(np.array_equal(np.nonzero(got["distance"] <= 0.25)[1], sub["ref"]) and
                            np.array_equal(got["numer"][got["distance"] <= 0.25], sub["numer"]))
                dist_obj["e2e_filtered"] = {"value": nd * nd / dt, "unit": "pairs/s", "filter": "-d 0.25", "hits": int(len(hits)), "seconds": dt,
                                            "kernel_ms": {"dist_tile32": f_tile, "rank_compress_index_mark": f_pack},
                                            "sample": "all %d x %d pairs, one GPU, one call" % (nd, nd), "h2d_bytes_per_step": int(2 * nd * (S * 8 + 12)),
                                            "d2h_bytes_per_step": int(len(hits) * 32), "agrees_with_matrix_rows": same}
            except fpm.FpmError as e:
                dist_obj["e2e_filtered"] = {"error": str(e)}
            del hit_buf
            if not args.no_cpu and world == 1:
                from oracle_py import RefLib
                if RefLib.available():
                    cq = min(args.cpu_dist_queries, nd)
                    threads = os.cpu_count() or 1
                    t0 = time.perf_counter()
                    cn, cd, cdist = RefLib().dist_batch(hp, hs, hp[:cq], hs[:cq], S, K, threads=threads)
                    dt = time.perf_counter() - t0
                    ok = (np.array_equal(cn.reshape(cq, nd), got["numer"][:cq]) and np.array_equal(cd.reshape(cq, nd), got["denom"][:cq] & 0x7fffffff)
                          and np.allclose(cdist.reshape(cq, nd), got["distance"][:cq], rtol=1e-12, atol=0))
                    dist_obj["cpu_baseline"] = {"value": cq * nd / dt, "unit": "pairs/s", "cores": threads, "kind": "reference",
                                                "sample": "%d query rows x %d refs (compareSketches loop over the reference's HashList, <=4096-pair chunks, no p-value)" % (cq, nd),
                                                "matches_gpu": bool(ok)}

    # ---- CPU baseline (rank 0, N=1 only) -----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        cg = min(args.cpu_genomes, n)
        r = cpu_sketch_baseline(host_seq, L, cg, threads)
        if r is not None:
            cpu = {"value": r[0], "unit": "Gk-mers/s", "cores": threads, "kind": "reference",
                   "sample": "%d of %d genomes (%d Mbp), reference TUs hash.cpp/MurmurHash3.cpp/MinHashHeap.cpp + restated sketchFile loop, -p %d, %.1f s" % (
                       cg, n, cg * L // 1_000_000, threads, r[1])}

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
            "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                         "traffic": alg_bytes * NCU_SKETCH_TRAFFIC_RATIO, "kernel": "sketch_hash_kernel_v2<21,true>", "launch_ms": hash_ms_avg,
                         "traffic_note": "dram__bytes_read+write of one ncu --set full capture (profiles/r01_sketch_hash_v4.txt: 332.2 MB for a 300.0 MB launch) scaled to this launch size",
                         "share_of_step": hash_ms / ms_total if ms_total else None,
                         "peak_source": hbm_src,
                         "note": "algorithmic bytes = 1 B per base read once; this kernel is integer-ALU bound, see roofline_int"},
            "roofline_int": {"bound": "int32-alu", "achieved": int_achieved / 1e12, "peak": int_peak / 1e12, "unit": "Tint32-op/s",
                             "frac": int_achieved / int_peak,
                             "note": "algorithmic ops = %d int32-op equivalents per k-mer (Murmur only, SURVEY.md 8d); peak = best of three inline-PTX microbenchmarks run live (strictly alternating IMAD/LOP3 with 16 independent chains: both integer pipes busy, ~0.94 warp instructions per clock and SM sub-partition)" % W_INT32_OPS[K],
                             "pipe_bound": {"note": "the same roofline per pipe, k=21: the ALU pipe carries 44 of the 74 algorithmic ops; the FMA pipe carries the ten 64-bit multiplies and two x*5+c = 12 wide multiplies (IMAD.WIDE: 0.487 x the IMAD rate, profiles/ubench/int_mix.cu) + 24 IMADs; each pipe at its measured single-pipe peak",
                                            "alu_pipe_ceiling_gkmers": int_peaks[0] / 44 / 1e9,
                                            "fma_pipe_ceiling_gkmers": 1.0 / (12 / (IMAD_WIDE_RATE * int_peaks[1]) + 24 / int_peaks[1]) / 1e9,
                                            "frac_of_tighter_ceiling": (windows_per_step / (hash_ms_avg * 1e-3) / 1e9) / min(int_peaks[0] / 44 / 1e9, 1.0 / (12 / (IMAD_WIDE_RATE * int_peaks[1]) + 24 / int_peaks[1]) / 1e9)} if K == 21 else None,
                             "peaks_measured": {"alu_pipe_lop3": int_peaks[0] / 1e12, "fma_pipe_imad": int_peaks[1] / 1e12, "alternating": int_peaks[2] / 1e12}},
            "kernel_ms": {"sketch_hash": hash_ms_avg, "sketch_select": sel_ms / max(sel_n, 1)},
            "cpu_baseline": cpu,
            "dist": dist_obj,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
