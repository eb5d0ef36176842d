"""gzip'ed genome collection: inflate on the GPU (fpm_gunzip_batch, one warp per file) against zlib on the host cores.
usage: python profiles/r02_gunzip_perf.py [n_files] [genome_len]   (40 distinct genomes, repeated to n_files)"""
import gzip, os, sys, time, zlib
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import __graft_entry__ as g
g._paths()
import fpmash_b200 as fpm
import ctypes as C

n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
glen = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
rng = np.random.default_rng(7)
lut = np.frombuffer(b"ACGT", dtype=np.uint8)
def genome(i):
    s = lut[rng.integers(0, 4, size=glen)].tobytes()
    return b">genome%d\n" % i + b"\n".join(s[j:j + 80] for j in range(0, glen, 80)) + b"\n"
plain = [genome(i) for i in range(40)]
t0 = time.time()
with ThreadPoolExecutor(16) as ex:
    gz40 = list(ex.map(lambda p: gzip.compress(p, 6), plain))
print("compressed 40 genomes in %.1f s, ratio %.2f" % (time.time() - t0, sum(map(len, plain)) / sum(map(len, gz40))), flush=True)
gz = [gz40[i % 40] for i in range(n_files)]
raw_bytes = sum(len(plain[i % 40]) for i in range(n_files))
comp_bytes = sum(map(len, gz))
# host: zlib on 16 threads (releases the GIL)
t0 = time.time()
with ThreadPoolExecutor(16) as ex:
    outs = list(ex.map(lambda b: len(zlib.decompress(b, 31)), gz))
th = time.time() - t0
print("host zlib, 16 threads: %.3f s = %.2f GB/s inflated" % (th, raw_bytes / th / 1e9), flush=True)
ctx = fpm.Context(0)
ctx.set_timing(True)
sizes = np.array([0] + [len(b) for b in gz], dtype=np.uint64)
off = np.cumsum(sizes).astype(np.uint64)
blob = np.frombuffer(b"".join(gz), dtype=np.uint8)
hp = fpm.host_alloc(blob.size) if hasattr(fpm, "host_alloc") else None
ends = np.zeros(n_files, dtype=np.uint64)
total, status = C.c_uint64(0), C.c_int(0)
for it in range(3):
    t0 = time.time()
    fpm._check(fpm.lib.fpm_gunzip_batch(ctx._h, blob.ctypes.data, off.ctypes.data, n_files, ends.ctypes.data, C.byref(total), C.byref(status)))
    t = time.time() - t0
    ms, n = ctx.get_timing(6) if hasattr(ctx, "get_timing") else (0, 0)
    print("gpu gunzip %d files: status %d, %.3f s wall (pageable H2D of %.2f GB included), kernel %.1f ms (%d launches) = %.1f GB/s inflated, %.1f MB/s per stream"
          % (n_files, status.value, t, comp_bytes / 1e9, ms, n, raw_bytes / (ms / 1e3) / 1e9 if ms else 0, raw_bytes / n_files / (ms / 1e3) / 1e6 if ms else 0), flush=True)
    ctx.set_timing(True)
assert total.value == raw_bytes + n_files
if n_files > 1000:
    sys.exit(0)
chk = np.zeros(total.value, dtype=np.uint8)
fpm._check(fpm.lib.fpm_gunzip_output(ctx._h, chk.ctypes.data))
o = 0
for i in range(min(n_files, 80)):
    p = plain[i % 40]
    assert chk[o:o + len(p)].tobytes() == p and chk[o + len(p)] == 0, i
    o += len(p) + 1
print("output checked against the plain genomes")
