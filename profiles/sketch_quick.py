"""Quick k=21 s=1000 sketch-kernel timing (100 x 5 Mbp resident in HBM); prints hash-kernel ms and Gk-mers/s.
A parity check against a reference result file keeps variants honest: FPM_REF=write|check."""
import os, sys, time
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
for k, s in ((21, 1000),) + (((32, 1000), (16, 1000)) if "--all" in sys.argv else ()) + (((32, 10000), (21, 10000)) if "--large" in sys.argv else ()):
    p = fpm.make_sketch_params(k=k, s=s)
    oh = torch.zeros((ng, s), dtype=torch.int64, device=dev); on = torch.zeros(ng, dtype=torch.int32, device=dev)
    f = lambda: ctx.sketch_batch_dev(seq.data_ptr(), seq.numel(), offs, p, oh.data_ptr(), None, on.data_ptr())
    for _ in range(3): f()
    torch.cuda.synchronize()
    ctx.set_timing(True)
    for _ in range(10): f()
    torch.cuda.synchronize()
    hm, hn = ctx.get_timing(fpm.KERNEL_SKETCH_HASH)
    ctx.set_timing(False)
    ms = hm / hn
    digest = int(oh.sum().item()) & 0xffffffffffff
    print("k=%d s=%d: hash %.3f ms  %.1f Gk-mers/s (kernel only)  digest %012x" % (k, s, ms, ng * (L - k + 1) / ms / 1e6, digest))
