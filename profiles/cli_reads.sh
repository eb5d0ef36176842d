# read-mode CLI (configs[3] shape: 1M reads x 150 bp FASTQ, -m 2): GPU FASTQ route vs host reader, FPMASH_TIMING milestones
M=fp-mash_b200/bin/mash
D=$(mktemp -d)
python - "$D" <<'PY'
import sys, numpy as np, os
d = sys.argv[1]; rng = np.random.default_rng(4); lut = np.frombuffer(b"ACGT", dtype=np.uint8)
L, n, rl = 5_000_000, 1_000_000, 150
g = lut[rng.integers(0, 4, size=L, dtype=np.uint8)]
start = rng.integers(0, L - rl, size=n)
reads = g[start[:, None] + np.arange(rl)[None, :]]
rows = np.empty((n, 4, rl + 1), dtype=np.uint8)
with open(os.path.join(d, "reads.fastq"), "wb") as f:
    for i0 in range(0, n, 100000):
        blk = reads[i0:i0 + 100000]
        out = bytearray()
        for i, r in enumerate(blk):
            out += b"@read%d\n" % (i0 + i) + r.tobytes() + b"\n+\n" + b"I" * rl + b"\n"
        f.write(out)
print("fastq bytes", os.path.getsize(os.path.join(d, "reads.fastq")))
PY
export FPMASH_TIMING=1
echo "== GPU FASTQ route"; $M sketch -r -m 2 -o $D/gpu $D/reads.fastq 2>&1 | grep -E "timing|Estimated"
echo "== host reader"; FPMASH_GPU_PARSE=0 $M sketch -r -m 2 -o $D/host $D/reads.fastq 2>&1 | grep -E "timing|Estimated"
cmp $D/gpu.msh $D/host.msh && echo "identical .msh"
rm -rf $D
