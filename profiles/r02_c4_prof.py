"""configs[3] (1,000,000 reads x 150 bp, -m 2, counts) on one GPU: time of the call and, under ncu, its launch list."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g._paths()
import fpmash_b200 as fpm
dev = torch.device("cuda", 0)
ctx = fpm.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
gen = torch.Generator(device=dev); gen.manual_seed(4)
GL, n_reads, rl = 5_000_000, 1_000_000, 150
genome = torch.randint(0, 4, (GL,), generator=gen, device=dev, dtype=torch.uint8)
start = torch.randint(0, GL - rl, (n_reads,), generator=gen, device=dev)
codes = genome[start[:, None] + torch.arange(rl, device=dev)[None, :]]
rev = torch.rand(n_reads, generator=gen, device=dev) < 0.5
codes = torch.where(rev[:, None], 3 - codes.flip(1), codes)
err = torch.rand((n_reads, rl), generator=gen, device=dev) < 0.01
codes = torch.where(err, torch.randint(0, 4, (n_reads, rl), generator=gen, device=dev, dtype=torch.uint8), codes)
lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
reads = lut[codes.long()]
reads[torch.rand((n_reads, rl), generator=gen, device=dev) < 0.001] = ord("N")
buf = torch.zeros((n_reads, rl + 1), dtype=torch.uint8, device=dev); buf[:, :rl] = reads; buf = buf.reshape(-1)
offs = np.array([0, buf.numel()], dtype=np.uint64)
p4 = fpm.make_sketch_params(k=21, s=1000, min_cov=2, want_counts=True)
oh = torch.zeros((1, 1000), dtype=torch.int64, device=dev); oc = torch.zeros((1, 1000), dtype=torch.int32, device=dev)
on = torch.zeros(1, dtype=torch.int32, device=dev); ok = torch.zeros(1, dtype=torch.int64, device=dev)
step = lambda k=None: ctx.sketch_batch_dev(buf.data_ptr(), buf.numel(), offs, p4, oh.data_ptr(), oc.data_ptr(), on.data_ptr(), k)
step(ok.data_ptr()); windows = int(ok.item())
for _ in range(3): step()
torch.cuda.synchronize()
ctx.set_timing(True)
ts = []
for _ in range(10):
    t0 = time.perf_counter(); step(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
hm, hn = ctx.get_timing(fpm.KERNEL_SKETCH_HASH); sm, sn = ctx.get_timing(fpm.KERNEL_SKETCH_SELECT)
ctx.set_timing(False)
print("C4: call %.3f ms (min of 10) = %.1f Gk-mers/s; hash kernel %.3f ms x %d per call, select %.3f ms; digest %x %d" % (
    min(ts), windows / min(ts) / 1e6, hm / hn, hn // 10, sm / sn, int(oh.sum().item()) & 0xffffffffffff, int(oc.sum().item())))
torch.cuda.profiler.start(); step(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
