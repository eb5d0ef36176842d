# FPMASH_TIMING milestones of the CLI: tiny input (fixed costs), 100 x 5 Mbp FASTA with the GPU parser and with the host reader, dist
M=fp-mash_b200/bin/mash
D=$(mktemp -d)
python - "$D" <<'PY'
import sys, numpy as np, os
d = sys.argv[1]; rng = np.random.default_rng(0); lut = np.frombuffer(b"ACGT", dtype=np.uint8)
def write(p, L):
    rows = np.full((L // 70, 71), ord("\n"), dtype=np.uint8); rows[:, :70] = lut[rng.integers(0, 4, size=(L // 70) * 70, dtype=np.uint8)].reshape(-1, 70)
    open(p, "wb").write(b">g synthetic\n" + rows.tobytes())
write(os.path.join(d, "tiny.fna"), 7000)
names = []
for g in range(100):
    p = os.path.join(d, "g%03d.fna" % g); write(p, 5_000_000); names.append(p)
open(os.path.join(d, "list.txt"), "w").write("\n".join(names) + "\n")
PY
export FPMASH_TIMING=1
echo "== tiny"; $M sketch -o $D/tiny $D/tiny.fna 2>&1 | grep timing
echo "== 100 x 5 Mbp -p 16, GPU FASTA parser"; $M sketch -l -p 16 -o $D/all $D/list.txt 2>&1 | grep timing
echo "== 100 x 5 Mbp -p 16, host reader (FPMASH_GPU_PARSE=0)"; FPMASH_GPU_PARSE=0 $M sketch -l -p 16 -o $D/all0 $D/list.txt 2>&1 | grep timing
cmp $D/all.msh $D/all0.msh && echo "identical .msh"
echo "== dist 100x100"; $M dist $D/all.msh $D/all.msh 2>&1 | grep timing
rm -rf $D
