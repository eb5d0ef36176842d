import sys, time
sys.path.insert(0, "fp-mash_b200/py")
import numpy as np, torch, fpmash_b200 as fpm
dev = torch.device("cuda", 0)
ctx = fpm.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
L, ng = 5_000_000, 100
gen = torch.Generator(device=dev); lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
seq = torch.zeros(ng * (L + 1), dtype=torch.uint8, device=dev)
for g in range(ng):
    gen.manual_seed(500 + g)
    seq[g * (L + 1):g * (L + 1) + L] = lut[torch.randint(0, 4, (L,), generator=gen, device=dev, dtype=torch.uint8).long()]
offs = np.arange(ng + 1, dtype=np.uint64) * np.uint64(L + 1)
for k, s in ((21, 1000), (32, 1000), (21, 10000), (32, 10000), (21, 4000), (21, 5000)):
    p = fpm.make_sketch_params(k=k, s=s)
    oh = torch.zeros((ng, s), dtype=torch.int64, device=dev); on = torch.zeros(ng, dtype=torch.int32, device=dev)
    f = lambda: ctx.sketch_batch_dev(seq.data_ptr(), seq.numel(), offs, p, oh.data_ptr(), None, on.data_ptr())
    f(); torch.cuda.synchronize()
    ctx.set_timing(True); l0 = ctx.launch_count()
    t0 = time.perf_counter(); f(); f(); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 2
    hm, hn = ctx.get_timing(fpm.KERNEL_SKETCH_HASH); sm, sn = ctx.get_timing(fpm.KERNEL_SKETCH_SELECT)
    ctx.set_timing(False)
    print("k=%d s=%d: %.2f ms/call  %.1f Gk/s | hash %.2f ms x%d  select %.2f ms x%d  launches/call %d" % (k, s, dt * 1e3, ng * (L - k + 1) / dt / 1e9, hm / max(hn, 1), hn // 2, sm / max(sn, 1), sn // 2, (ctx.launch_count() - l0) // 2))

# -i style: many short records, each its own sketch (accept-all thresholds: every window is a survivor)
nrec, rl = 200000, 1500
buf = torch.zeros((nrec, rl + 1), dtype=torch.uint8, device=dev)
gen.manual_seed(9)
buf[:, :rl] = lut[torch.randint(0, 4, (nrec, rl), generator=gen, device=dev, dtype=torch.uint8).long()]
buf = buf.reshape(-1)
offs = np.arange(nrec + 1, dtype=np.uint64) * np.uint64(rl + 1)
p = fpm.make_sketch_params(k=21, s=1000)
oh = torch.zeros((nrec, 1000), dtype=torch.int64, device=dev); on = torch.zeros(nrec, dtype=torch.int32, device=dev)
f = lambda: ctx.sketch_batch_dev(buf.data_ptr(), buf.numel(), offs, p, oh.data_ptr(), None, on.data_ptr())
f(); torch.cuda.synchronize()
t0 = time.perf_counter(); f(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("-i mode, %d records x %d bp (accept-all): %.1f ms  %.2f Gk/s" % (nrec, rl, dt * 1e3, nrec * (rl - 20) / dt / 1e9))
