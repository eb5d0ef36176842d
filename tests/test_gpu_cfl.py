"""GPU: lyn2vec's CFL / ICFL / CFL_ICFL fingerprints.  Pinned by the reference's own
DNA1-CFL.txt, which lyn2vec produced from DNA1.fasta (README.md:34-52), by the ICFL and CFL_ICFL-30 files the
reference's lyn2vec writes for the same FASTA (tests/golden/make_golden.py), and by plain restatements of Duval and
of the recursive inverse Lyndon factorisation in Python (lyn2vec/factorizations.py:102-126,143-248,265-300)."""
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


def icfl(word):
    """ICFL_recursive (factorizations.py:143-248), restated: lengths of the inverse Lyndon factors of `word`."""
    n = len(word)
    if n == 0:
        return []
    # find_pre: the longest prefix along which word[j] <= word[i] keeps holding, plus the character that breaks it
    i, j = 0, 1
    while j < n and word[j] <= word[i]:
        i = 0 if word[j] < word[i] else i + 1
        j += 1
    if n == 1 or j == n:
        return [n]
    # find_bre: shortest border b of word[:j] (walking the failure-function chain) with word[b] < word[j]
    f = [0] * j
    k = 0
    for t in range(1, j):
        while k > 0 and word[k] != word[t]:
            k = f[k - 1]
        if word[k] == word[t]:
            k += 1
        f[t] = k
    t, last = j, f[j - 1]
    while t > 0:
        if word[f[t - 1]] < word[j]:
            last = f[t - 1]
        t = f[t - 1]
    rest = icfl(word[j - last:])
    if rest[0] > last:
        return [j - last] + rest
    return [j - last + rest[0]] + rest[1:]


def cfl_icfl(word, c):
    """CFL_icfl (factorizations.py:265-300) without the << >> markers (dropped by fingerprint_utils.py:459-463)."""
    out, pos = [], 0
    for flen in duval(word):
        out += [flen] if flen <= c else icfl(word[pos:pos + flen])
        pos += flen
    return out


@pytest.mark.parametrize("fact", ["ICFL", "CFL_ICFL-30", "CFL_COMB", "ICFL_COMB", "CFL_ICFL_COMB-10"])
def test_fingerprint_cli_icfl_reproduces_lyn2vec_file(tmp_path, fact):
    shutil.copy(os.path.join(GOLDEN, "DNA1.fasta"), tmp_path / "DNA1.fasta")
    r = subprocess.run([MASH, "fingerprint", "-t", fact, "DNA1.fasta"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    want = gzip.open(os.path.join(GOLDEN, "DNA1-%s.txt.gz" % fact), "rb").read()
    assert open(tmp_path / ("fingerprint_%s.txt" % fact), "rb").read() == want


def test_icfl_rows_and_fused_hashes(ctx, oracle):
    rng = np.random.default_rng(6)
    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
    recs = [alpha[rng.integers(0, 4, size=n)].tobytes() for n in (350, 100, 99, 1, 2, 7, 180)]
    recs += [b"A" * 150, b"TGCA" * 40, b"T" * 60 + b"A" * 60, b"ACGT" * 30 + b"T", bytes(range(126, 33, -1)), bytes(range(33, 127))]
    recs.append(bytes(rng.integers(33, 127, size=130, dtype=np.uint8)))
    recs.append(bytes(rng.integers(65, 67, size=256, dtype=np.uint8)))          # binary alphabet: long borders
    for fact, fn in (("ICFL", icfl), ("CFL_ICFL-10", lambda w: cfl_icfl(w, 10)), ("CFL_ICFL-30", lambda w: cfl_icfl(w, 30))):
        rows, hashes, woff = ctx.fingerprint_batch(recs, window=100, factorization=fact)
        w = 0
        for r, rec in enumerate(recs):
            assert int(woff[r]) == w
            for word in shifts(rec):
                assert rows[w] == fn(word), (fact, r, w, word)
                assert sum(rows[w]) == len(word)
                assert int(hashes[w]) == oracle.fp_hash(rows[w], 42, False)
                w += 1
        assert w == int(woff[-1]) == len(rows)
    rows, _, _ = ctx.fingerprint_batch(recs[-1:], window=256, factorization="ICFL")
    assert rows[3] == icfl(shifts(recs[-1], 256)[3])


def comb(word, alg, alg_rc):
    """d_duval_ (lyn2vec/factorizations_comb.py:213-245): the boundaries of alg(word) and the mirrored boundaries of
    alg_rc(reverse complement of word)."""
    cuts, pos = set(), 0
    for f in alg(word):
        pos += f
        cuts.add(pos)
    rc = bytes(word[::-1]).translate(bytes.maketrans(b"ACGT", b"TGCA"))
    pos = 0
    for f in alg_rc(rc):
        cuts.add(len(word) - pos)
        pos += f
    out, last = [], 0
    for p in sorted(cuts):
        out.append(p - last)
        last = p
    return out


def test_comb_rows_and_fused_hashes(ctx, oracle):
    """CFL_COMB / ICFL_COMB / CFL_ICFL_COMB-<C>, windows with N, short records; the reverse complement of CFL_ICFL_COMB-<C>
    is factorised with threshold 30 whatever C is, as the reference does (factorizations_comb.py:221)."""
    rng = np.random.default_rng(16)
    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
    alpha_n = np.frombuffer(b"ACGTN", dtype=np.uint8)
    recs = [alpha[rng.integers(0, 4, size=n)].tobytes() for n in (350, 100, 99, 1, 2, 7, 180)]
    recs += [alpha_n[rng.integers(0, 5, size=220)].tobytes(), b"A" * 150, b"TGCA" * 40, b"T" * 60 + b"A" * 60, b"ACGT" * 30 + b"T"]
    recs.append(bytes(rng.integers(65, 67, size=256, dtype=np.uint8)).replace(b"B", b"C"))    # two letters: long borders
    cases = (("CFL_COMB", duval, duval), ("ICFL_COMB", icfl, icfl), ("CFL_ICFL_COMB-10", lambda w: cfl_icfl(w, 10), lambda w: cfl_icfl(w, 30)),
             ("CFL_ICFL_COMB-30", lambda w: cfl_icfl(w, 30), lambda w: cfl_icfl(w, 30)))
    for fact, fn, fn_rc in cases:
        rows, hashes, woff = ctx.fingerprint_batch(recs, window=100, factorization=fact)
        w = 0
        for r, rec in enumerate(recs):
            for word in shifts(rec):
                assert rows[w] == comb(word, fn, fn_rc), (fact, r, w, word)
                assert sum(rows[w]) == len(word)
                assert int(hashes[w]) == oracle.fp_hash(rows[w], 42, False)
                w += 1
        assert w == int(woff[-1]) == len(rows)
