#!/usr/bin/env python
"""End-to-end timing of the C++ CLI on synthetic FASTA files (wall clock, includes file parsing).
usage: python profiles/cli_e2e.py [n_genomes] [genome_len] [threads...]"""
import os, subprocess, sys, tempfile, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MASH = os.path.join(ROOT, "fp-mash_b200", "bin", "mash")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
L = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
threads = [int(x) for x in sys.argv[3:]] or [1, 16]
d = tempfile.mkdtemp(prefix="fpm_cli_")
rng = np.random.default_rng(0)
lut = np.frombuffer(b"ACGT", dtype=np.uint8)
names = []
t0 = time.time()
for g in range(n):
    L = (L // 70) * 70
    rows = np.full((L // 70, 71), ord("\n"), dtype=np.uint8)
    rows[:, :70] = lut[rng.integers(0, 4, size=L, dtype=np.uint8)].reshape(-1, 70)
    p = os.path.join(d, "g%04d.fna" % g)
    with open(p, "wb") as f:
        f.write(b">genome%d synthetic\n" % g + rows.tobytes())
    names.append(p)
open(os.path.join(d, "list.txt"), "w").write("\n".join(names) + "\n")
print("generated %d x %d bp in %.1f s" % (n, L, time.time() - t0))
for p in threads:
    out = os.path.join(d, "out_p%d" % p)
    t0 = time.time()
    subprocess.check_call([MASH, "sketch", "-l", "-p", str(p), "-o", out, os.path.join(d, "list.txt")])
    dt = time.time() - t0
    print("mash sketch -p %2d: %.2f s  -> %.2f Gk-mers/s end to end from FASTA files (%.1f GB)" % (p, dt, n * (L - 20) / dt / 1e9, n * L / 1e9))
t0 = time.time()
subprocess.check_call([MASH, "dist", out + ".msh", out + ".msh"], stdout=subprocess.DEVNULL)
print("mash dist %d x %d: %.2f s" % (n, n, time.time() - t0))
