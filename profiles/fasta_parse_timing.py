"""Timing of the GPU FASTA parser + sketch on one 256 MB batch of 70-column FASTA (first call vs steady state)."""
import sys, time
sys.path.insert(0, "fp-mash_b200/py")
import numpy as np, fpmash_b200 as fpm, ctypes as C
rng = np.random.default_rng(0)
lut = np.frombuffer(b"ACGT", dtype=np.uint8)
L = 5_000_000 // 70 * 70
rows = np.full((L // 70, 71), ord("\n"), dtype=np.uint8); rows[:, :70] = lut[rng.integers(0, 4, size=L, dtype=np.uint8)].reshape(-1, 70)
one = b">genome synthetic\n" + rows.tobytes()
nfiles = 50
raw_np = np.frombuffer(b"".join(one + b"\0" for _ in range(nfiles)), dtype=np.uint8)
ctx = fpm.Context(0)
pin = fpm.PinnedBuffer(raw_np.size); pin.array[:] = raw_np
lib = fpm.lib
p = fpm.make_sketch_params(k=21, s=1000)
for it in range(3):
    nrec, nseq, status = C.c_uint64(0), C.c_uint64(0), C.c_int(0)
    t0 = time.perf_counter()
    fpm._check(lib.fpm_fasta_parse(ctx._h, pin.array.ctypes.data, pin.array.size, C.byref(nrec), C.byref(nseq), C.byref(status)))
    recs = np.zeros(nrec.value, dtype=fpm.FASTA_RECORD_DTYPE)
    fpm._check(lib.fpm_fasta_records(ctx._h, recs.ctypes.data))
    t1 = time.perf_counter()
    goff = np.append(recs["seq_begin"], np.uint64(nseq.value)).astype(np.uint64)
    out = ctx.sketch_parsed(goff, p)
    t2 = time.perf_counter()
    print("call %d: %d records, %.1f MB raw -> parse %.2f ms (%.1f GB/s incl. H2D), sketch %.2f ms" % (it, nrec.value, raw_np.size / 1e6, (t1 - t0) * 1e3, raw_np.size / (t1 - t0) / 1e9, (t2 - t1) * 1e3))
