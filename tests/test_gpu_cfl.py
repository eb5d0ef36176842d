"""GPU: lyn2vec's CFL fingerprints.  Pinned by the reference's own DNA1-CFL.txt, which
lyn2vec produced from DNA1.fasta (README.md:34-52), and by a plain Duval restatement in Python
(lyn2vec/factorizations.py:102-126)."""
import gzip
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
MASH = os.path.join(ROOT, "fp-mash_b200", "bin", "mash")


def duval(word):
    """CFL (factorizations.py:102-126), 1-based indices as in the reference."""
    out, k, n = [], 0, len(word)
    while k < n:
        i, j = k + 1, k + 2
        while True:
            if j == n + 1 or word[j - 1] < word[i - 1]:
                while k < i:
                    out.append(j - i)
                    k = k + j - i
                break
            i = k + 1 if word[j - 1] > word[i - 1] else i + 1
            j += 1
    return out


def shifts(s, size=100):
    """shift_string (fingerprint_utils.py:95-110)."""
    if len(s) < size:
        return [s]
    return [(s[i:i + size] + s[:max(0, i + size - len(s))]) for i in range(len(s))]


def test_fingerprint_cli_reproduces_lyn2vec_file(tmp_path):
    shutil.copy(os.path.join(GOLDEN, "DNA1.fasta"), tmp_path / "DNA1.fasta")
    r = subprocess.run([MASH, "fingerprint", "-o", "DNA1-CFL.txt", "DNA1.fasta"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    want = gzip.open(os.path.join(GOLDEN, "DNA1-CFL.txt.gz"), "rb").read()
    assert open(tmp_path / "DNA1-CFL.txt", "rb").read() == want
    # and the whole README chain: fingerprints -> mash sketch -fp -> the reference's .msh, byte for byte
    r = subprocess.run([MASH, "sketch", "-fp", "DNA1-CFL.txt", "-o", "DNA1-sketch.msh"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(tmp_path / "DNA1-sketch.msh", "rb").read() == open(os.path.join(GOLDEN, "DNA1-sketch.msh"), "rb").read()


def test_cfl_rows_and_fused_hashes(ctx, oracle):
    rng = np.random.default_rng(5)
    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
    recs = [alpha[rng.integers(0, 4, size=n)].tobytes() for n in (350, 100, 99, 1, 7, 180)]
    recs.append(b"A" * 150)
    recs.append(b"TGCA" * 40)
    recs.append(bytes(rng.integers(33, 127, size=130, dtype=np.uint8)))
    rows, hashes, woff = ctx.cfl_fingerprint_batch(recs, window=100)
    w = 0
    for r, rec in enumerate(recs):
        assert int(woff[r]) == w
        for word in shifts(rec):
            assert rows[w] == duval(word), (r, w)
            assert int(hashes[w]) == oracle.fp_hash(rows[w], 42, False)
            w += 1
    assert w == int(woff[-1]) == len(rows)
    rows16, h16, _ = ctx.cfl_fingerprint_batch(recs[:2], window=16, use64=True)
    assert rows16[5] == duval(shifts(recs[0], 16)[5]) and int(h16[5]) == oracle.fp_hash(rows16[5], 42, True)
