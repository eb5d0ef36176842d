"""`mash sketch` over a directory of .fna.gz genomes: GPU inflate route against the host zlib route (same .msh), with the
CLI's own timing trace.  usage: python profiles/r02_cli_gz.py [n_files] [genome_len]"""
import gzip, os, subprocess, sys, tempfile, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MASH = os.path.join(ROOT, "fp-mash_b200", "bin", "mash")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
glen = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
d = tempfile.mkdtemp(prefix="gzcli_")
lut = np.frombuffer(b"ACGT", dtype=np.uint8)
def make(i):
    rng = np.random.default_rng(i)
    s = lut[rng.integers(0, 4, size=glen)].tobytes()
    text = b">genome%d contig1\n" % i + b"\n".join(s[j:j + 80] for j in range(0, glen, 80)) + b"\n"
    open(os.path.join(d, "g%04d.fna.gz" % i), "wb").write(gzip.compress(text, 6))
with ThreadPoolExecutor(16) as ex:
    list(ex.map(make, range(n)))
open(os.path.join(d, "list.txt"), "w").write("".join("g%04d.fna.gz\n" % i for i in range(n)))
print("%d files, %.2f GB compressed" % (n, sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d)) / 1e9), flush=True)
for tag, env in (("gpu", {}), ("host", {"FPMASH_GPU_GUNZIP": "0"})):
    for rep in range(2):
        t0 = time.time()
        r = subprocess.run([MASH, "sketch", "-p", "16", "-l", "list.txt", "-o", tag], cwd=d, capture_output=True, text=True,
                           env=dict(os.environ, FPMASH_TIMING="1", **env))
        dt = time.time() - t0
        assert r.returncode == 0, r.stderr[-2000:]
    trace = [l for l in r.stderr.splitlines() if l.startswith("[timing]")][-14:]
    print("%s route: %.2f s wall (second run)" % (tag, dt))
    for l in trace:
        print("   ", l)
a, b = open(os.path.join(d, "gpu.msh"), "rb").read(), open(os.path.join(d, "host.msh"), "rb").read()
print("identical .msh:", a == b, len(a))
