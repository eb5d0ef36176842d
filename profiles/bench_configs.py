#!/usr/bin/env python
"""Timings of the BASELINE.json configs that are parity-test shapes rather than the bench line:
  C4  read-mode sketch of synthetic 30x FASTQ-like reads (5 Mbp genome, 150 bp reads, -m 2, counts)
  C5  s=10000, k=32 sketches; dist of Q queries x 10000 references at s=10000 (Q scaled to fit one GPU-minute)
Inputs resident in HBM, CUDA events via the library's per-kernel timers plus wall clock around the C-ABI call.
usage: python profiles/bench_configs.py [c5_queries]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fp-mash_b200", "py"))
import numpy as np
import torch
import fpmash_b200 as fpm

dev = torch.device("cuda", 0)
ctx = fpm.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
out = {}


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


# ---- C4: 1,000,000 reads x 150 bp of one 5 Mbp genome, 1 % substitutions, 0.1 % N, both strands ----------------
gen = torch.Generator(device=dev)
gen.manual_seed(4)
L, n_reads, rl = 5_000_000, 1_000_000, 150
genome = torch.randint(0, 4, (L,), generator=gen, device=dev, dtype=torch.uint8)
start = torch.randint(0, L - rl, (n_reads,), generator=gen, device=dev)
idx = start[:, None] + torch.arange(rl, device=dev)[None, :]
codes = genome[idx]
rev = torch.rand(n_reads, generator=gen, device=dev) < 0.5
codes = torch.where(rev[:, None], 3 - codes.flip(1), codes)
err = torch.rand((n_reads, rl), generator=gen, device=dev) < 0.01
codes = torch.where(err, torch.randint(0, 4, (n_reads, rl), generator=gen, device=dev, dtype=torch.uint8), codes)
lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
reads = lut[codes.long()]
reads[torch.rand((n_reads, rl), generator=gen, device=dev) < 0.001] = ord("N")
buf = torch.zeros((n_reads, rl + 1), dtype=torch.uint8, device=dev)
buf[:, :rl] = reads
buf = buf.reshape(-1)
offs = np.array([0, buf.numel()], dtype=np.uint64)
p = fpm.make_sketch_params(k=21, s=1000, min_cov=2, want_counts=True)
oh = torch.zeros((1, 1000), dtype=torch.int64, device=dev)
oc = torch.zeros((1, 1000), dtype=torch.int32, device=dev)
on = torch.zeros(1, dtype=torch.int32, device=dev)
ok = torch.zeros(1, dtype=torch.int64, device=dev)
ctx.sketch_batch_dev(buf.data_ptr(), buf.numel(), offs, p, oh.data_ptr(), oc.data_ptr(), on.data_ptr(), ok.data_ptr())
windows = int(ok.item())
ctx.set_timing(True)
dt = timed(lambda: ctx.sketch_batch_dev(buf.data_ptr(), buf.numel(), offs, p, oh.data_ptr(), oc.data_ptr(), on.data_ptr()))
hm, hn = ctx.get_timing(fpm.KERNEL_SKETCH_HASH)
ctx.set_timing(False)
out["C4_read_sketch_m2"] = {"reads": n_reads, "read_len": rl, "valid_windows": windows, "sketch_full": int(on.item()) == 1000,
                            "coverage_estimate": float(oc.sum().item()) / 1000.0, "ms": dt * 1e3, "Gk-mers/s": windows / dt / 1e9,
                            "hash_passes_per_call": hn / 4, "hash_kernel_ms_per_pass": hm / max(hn, 1),
                            "note": "one threshold pass (coverage guess 32x for -m) + trace pass for the order-dependent top count"}
del buf, reads, codes, idx

# ---- C5: k=32, s=10000 sketches of 5 Mbp genomes -------------------------------------------------------------------
ng = 200
seq = torch.zeros(ng * (L + 1), dtype=torch.uint8, device=dev)
for g in range(ng):
    gen.manual_seed(500 + g)
    seq[g * (L + 1):g * (L + 1) + L] = lut[torch.randint(0, 4, (L,), generator=gen, device=dev, dtype=torch.uint8).long()]
offs = np.arange(ng + 1, dtype=np.uint64) * np.uint64(L + 1)
p5 = fpm.make_sketch_params(k=32, s=10000)
oh5 = torch.zeros((ng, 10000), dtype=torch.int64, device=dev)
on5 = torch.zeros(ng, dtype=torch.int32, device=dev)
dt = timed(lambda: ctx.sketch_batch_dev(seq.data_ptr(), seq.numel(), offs, p5, oh5.data_ptr(), None, on5.data_ptr()))
out["C5_sketch_k32_s10000"] = {"genomes": ng, "ms": dt * 1e3, "Gk-mers/s": ng * (L - 31) / dt / 1e9, "sketches_full": bool((on5 == 10000).all().item())}
del seq

# ---- C5 dist: Q queries x 10000 references, s = 10000 (clustered synthetic sketches) -------------------------------
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
s5, nref = 10000, 10000
gen.manual_seed(5)
hi = int((1 << 64) * s5 / 4_999_969)
big = 1 << 62


def panel(n, seed):
    gen.manual_seed(seed)
    cores = torch.randint(0, hi, (100, 2 * s5), generator=gen, device=dev, dtype=torch.int64)
    res = torch.empty((n, s5), dtype=torch.int64, device=dev)
    for i0 in range(0, n, 500):
        i1 = min(n, i0 + 500)
        m = i1 - i0
        core = cores[torch.arange(i0, i1, device=dev) % 100]
        keep = torch.rand((m, 2 * s5), generator=gen, device=dev) < 0.6
        core = torch.where(keep, core, torch.full_like(core, big))
        cand, _ = torch.sort(torch.cat([core, torch.randint(0, hi, (m, s5), generator=gen, device=dev, dtype=torch.int64)], dim=1), dim=1)
        dup = torch.zeros_like(cand, dtype=torch.bool)
        dup[:, 1:] = cand[:, 1:] == cand[:, :-1]
        cand, _ = torch.sort(torch.where(dup, torch.full_like(cand, big), cand), dim=1)
        res[i0:i1] = cand[:, :s5]
    return res


ref, qry = panel(nref, 51), panel(Q, 52)
rs = torch.full((nref,), s5, dtype=torch.int32, device=dev)
qs = torch.full((Q,), s5, dtype=torch.int32, device=dev)
rlens = torch.full((nref,), 5_000_000, dtype=torch.int64, device=dev)
qlens = torch.full((Q,), 5_000_000, dtype=torch.int64, device=dev)
pairs = torch.empty(Q * nref * 24, dtype=torch.uint8, device=dev)
steps = torch.zeros(1, dtype=torch.int64, device=dev)
call = lambda: ctx.dist_tile_dev((ref.data_ptr(), rs.data_ptr(), rlens.data_ptr(), nref, s5), (qry.data_ptr(), qs.data_ptr(), qlens.data_ptr(), Q, s5),
                                 s5, 32, 4.0 ** 32, pairs.data_ptr(), steps.data_ptr())
call()
torch.cuda.synchronize()
steps.zero_()
t0 = time.perf_counter()
call()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
out["C5_dist_s10000"] = {"queries": Q, "references": nref, "pairs": Q * nref, "ms": dt * 1e3, "pairs/s": Q * nref / dt,
                         "merge_steps/s": float(steps.item()) / dt, "steps_per_pair": float(steps.item()) / (Q * nref),
                         "full_config_estimate_s_1gpu": 1e9 / (Q * nref / dt)}
print(json.dumps(out, indent=1))
