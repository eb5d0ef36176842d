"""GPU, end to end through the C++ host binary (`mash sketch`, `mash dist`): the reference's own
recipes (Makefile.in:95-115, README.md:34-98) reproduce its fixtures -- .msh files byte for
byte, dist rows character for character."""
import gzip
import os
import shutil
import subprocess

import numpy as np
import pytest

import mshpy
from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
MASH = os.path.join(ROOT, "fp-mash_b200", "bin", "mash")


def run(args, cwd=None):
    r = subprocess.run([MASH] + args, capture_output=True, text=True, cwd=cwd)
    assert r.returncode == 0, r.stderr
    return r


def gunzip_to(name, dst, crlf=False):
    data = gzip.open(os.path.join(GOLDEN, name), "rb").read()
    if crlf:
        data = data.replace(b"\n", b"\r\n")
    with open(dst, "wb") as f:
        f.write(data)


def same_bytes(a, b):
    return open(a, "rb").read() == open(b, "rb").read()


def test_sketch_test_sequence_byte_identical(tmp_path):
    os.makedirs(tmp_path / "new_data")
    shutil.copy(os.path.join(GOLDEN, "test_sequence.fasta"), tmp_path / "new_data" / "test_sequence.fasta")
    r = run(["sketch", "new_data/test_sequence.fasta"], cwd=tmp_path)
    assert "Sketching new_data/test_sequence.fasta..." in r.stderr and "Writing to new_data/test_sequence.fasta.msh..." in r.stderr
    assert same_bytes(tmp_path / "new_data" / "test_sequence.fasta.msh", os.path.join(GOLDEN, "test_sequence.msh"))


@pytest.mark.parametrize("n", [1, 2, 3])
def test_sketch_fingerprints_byte_identical(tmp_path, n):
    """README recipe: mash sketch -fp DNAn-CFL.txt -o DNAn-sketch.msh (4 capnp segments, 5 far pointers)."""
    gunzip_to("DNA%d-CFL.txt.gz" % n, tmp_path / ("DNA%d-CFL.txt" % n))
    r = run(["sketch", "-fp", "DNA%d-CFL.txt" % n, "-o", "DNA%d-sketch.msh" % n], cwd=tmp_path)
    assert "Initializing from fingerprints..." in r.stdout and "Initialization complete." in r.stdout
    assert same_bytes(tmp_path / ("DNA%d-sketch.msh" % n), os.path.join(GOLDEN, "DNA%d-sketch.msh" % n))


def test_sketch_reads_byte_identical(tmp_path):
    """Makefile.in:106-107: mash sketch -r -I reads reads1.fastq reads2.fastq -o reads.msh.  The
    fixture was made from CRLF files (its comment holds a '\\r'), so the inputs are converted."""
    gunzip_to("reads1.fastq.gz", tmp_path / "reads1.fastq", crlf=True)
    gunzip_to("reads2.fastq.gz", tmp_path / "reads2.fastq", crlf=True)
    r = run(["sketch", "-r", "-I", "reads", "reads1.fastq", "reads2.fastq", "-o", "reads.msh"], cwd=tmp_path)
    # cerr << double prints 6 significant digits (502359.6 -> 502360); the stored length truncates to 502359
    assert "Estimated genome size: 502360" in r.stderr and "Estimated coverage:    1.115" in r.stderr
    assert same_bytes(tmp_path / "reads.msh", os.path.join(GOLDEN, "reads.msh"))


def test_sketch_then_paste_byte_identical(tmp_path):
    """paste_example/read1_2.msh = the two read files sketched as genomes, then pasted."""
    os.makedirs(tmp_path / "test")
    gunzip_to("reads1.fastq.gz", tmp_path / "test" / "reads1.fastq", crlf=True)
    gunzip_to("reads2.fastq.gz", tmp_path / "test" / "reads2.fastq", crlf=True)
    run(["sketch", "./test/reads1.fastq"], cwd=tmp_path)
    run(["sketch", "./test/reads2.fastq"], cwd=tmp_path)
    run(["paste", "read1_2", "./test/reads1.fastq.msh", "./test/reads2.fastq.msh"], cwd=tmp_path)
    assert same_bytes(tmp_path / "read1_2.msh", os.path.join(GOLDEN, "read1_2.msh"))
    # one invocation over both files gives the same two sketches
    run(["sketch", "-o", "both", "./test/reads1.fastq", "./test/reads2.fastq"], cwd=tmp_path)
    assert same_bytes(tmp_path / "both.msh", os.path.join(GOLDEN, "read1_2.msh"))


def test_dist_reproduces_genomes_dist(tmp_path):
    """Makefile.in:109-111: mash dist genomes.msh reads.msh == test/ref/genomes.dist."""
    files = [os.path.join(GOLDEN, "genome%d.fna.msh" % i) for i in (1, 2, 3)]
    run(["paste", "genomes"] + files, cwd=tmp_path)
    r = run(["dist", "genomes.msh", os.path.join(GOLDEN, "reads.msh")], cwd=tmp_path)
    want = [l.split("\t") for l in open(os.path.join(GOLDEN, "genomes.dist")).read().splitlines()]
    got = [l.split("\t") for l in r.stdout.splitlines()]
    assert len(got) == 3
    for g, w in zip(got, want):
        assert g[0] == "data/" + w[0] and g[1:] == w[1:]
    # tutorials.rst:24,56-57
    r = run(["dist", files[0], files[1], files[2]], cwd=tmp_path)
    rows = [l.split("\t") for l in r.stdout.splitlines()]
    assert rows[0][2:] == ["0.0222766", "0", "456/1000"] and rows[1][2:] == ["0", "0", "1000/1000"]
    # table output and thresholds
    t = run(["dist", "-t", "genomes.msh", files[0]], cwd=tmp_path).stdout.splitlines()
    assert t[0].split("\t") == ["#query", "data/genome1.fna", "data/genome2.fna", "data/genome3.fna"]
    assert t[1].split("\t") == ["data/genome1.fna", "0", "0.0222766", "0"]
    d = run(["dist", "-d", "0.01", "genomes.msh", files[0]], cwd=tmp_path).stdout.splitlines()
    assert [l.split("\t")[0] for l in d] == ["data/genome1.fna", "data/genome3.fna"]


def test_dist_fingerprint_mode_matches_oracle(tmp_path, oracle):
    """mash dist -fp a.txt b.txt: unsorted untruncated 32-bit lists, k=1, kmerSpace=10 -- the literal
    loop defines the result (no golden output exists in the reference: SURVEY.md Appendix A.18)."""
    for n in (1, 2):
        gunzip_to("DNA%d-CFL.txt.gz" % n, tmp_path / ("DNA%d-CFL.txt" % n))
    r = run(["dist", "-fp", "DNA1-CFL.txt", "DNA2-CFL.txt"], cwd=tmp_path)
    rows = [l.split("\t") for l in r.stdout.splitlines() if "\t" in l]
    assert len(rows) == 25
    a = mshpy.load(os.path.join(GOLDEN, "DNA1-sketch.msh")).refs
    b = mshpy.load(os.path.join(GOLDEN, "DNA2-sketch.msh")).refs
    k = 0
    for q in b:
        for ref in a:
            w = oracle.compare(ref["hashes32"], q["hashes32"], ref["length"], q["length"], 1000, 1, 10.0)
            g = rows[k]
            k += 1
            assert g[0] == ref["name"] and g[1] == q["name"]
            assert g[2] == "%g" % w["distance"] and g[3] == "%g" % w["pvalue"] and g[4] == "%d/%d" % (w["numer"], w["denom"])
    # the same comparison from the .msh files truncates each list to its first 1000 hashes (Appendix A.9)
    run(["sketch", "-fp", "DNA1-CFL.txt", "-o", "a"], cwd=tmp_path)
    run(["sketch", "-fp", "DNA2-CFL.txt", "-o", "b"], cwd=tmp_path)
    r = run(["dist", "-fp", "a.msh", "b.msh"], cwd=tmp_path)
    rows = [l.split("\t") for l in r.stdout.splitlines() if "\t" in l]
    w = oracle.compare(a[0]["hashes32"][:1000], b[0]["hashes32"][:1000], a[0]["length"], b[0]["length"], 1000, 1, 10.0)
    assert rows[0][4] == "%d/%d" % (w["numer"], w["denom"]) and rows[0][2] == "%g" % w["distance"]


def test_sketch_individual_and_options(tmp_path, oracle):
    from util import dirty_dna
    rng = np.random.default_rng(8)
    recs = [dirty_dna(rng, n) for n in (5000, 12, 30000, 800)]
    with open(tmp_path / "multi.fa", "wb") as f:
        for i, r in enumerate(recs):
            f.write(b">rec%d some comment %d\n" % (i, i))
            for p in range(0, len(r), 70):
                f.write(r[p:p + 70] + b"\n")
    run(["sketch", "-i", "-k", "16", "-s", "400", "-S", "7", "-M", "multi.fa"], cwd=tmp_path)
    m = mshpy.load(tmp_path / "multi.fa.msh")
    assert (m.kmer_size, m.sketch_size, m.hash_seed, m.concatenated, m.used_old_list) == (16, 400, 7, False, False)
    keep = [(i, r) for i, r in enumerate(recs) if len(r) >= 16]
    assert [x["name"] for x in m.refs] == ["rec%d" % i for i, _ in keep]
    for x, (i, r) in zip(m.refs, keep):
        w = oracle.sketch([r], k=16, s=400, seed=7)
        assert x["hashes32"] == [int(v) for v in w["hashes"]] and x["counts32"] == [int(v) for v in w["counts"]]
        assert x["length"] == len(r) and x["comment"] == "some comment %d" % i
    # whole-file mode, gzipped input, noncanonical
    with open(tmp_path / "multi.fa", "rb") as f, gzip.open(tmp_path / "multi.fa.gz", "wb") as g:
        g.write(f.read())
    run(["sketch", "-n", "-o", "whole", "multi.fa.gz"], cwd=tmp_path)
    m = mshpy.load(tmp_path / "whole.msh")
    w = oracle.sketch(recs, k=21, s=1000, noncanonical=True)
    assert m.refs[0]["hashes64"] == [int(v) for v in w["hashes"]] and m.noncanonical
    assert m.refs[0]["length"] == sum(len(r) for _, r in [(i, r) for i, r in enumerate(recs) if len(r) >= 21])
    assert m.refs[0]["comment"] == "[3 seqs] rec0 some comment 0 [...]" and m.refs[0]["name"] == "multi.fa.gz"


def test_triangle_matches_dist_and_positional_fp(tmp_path, oracle):
    """mash triangle (CommandTriangle.cpp): Phylip lower triangle from the dist kernel; -fp uses the
    fork's positional compareFingerprints with a chi-square(1) p-value."""
    import math
    files = [os.path.join(GOLDEN, "genome%d.fna.msh" % i) for i in (1, 2, 3)]
    r = run(["triangle"] + files, cwd=tmp_path)
    lines = r.stdout.splitlines()
    assert lines[0] == "\t3" and lines[1] == "data/genome1.fna"
    assert lines[2].split("\t") == ["data/genome2.fna", "0.0222766"]
    assert lines[3].split("\t") == ["data/genome3.fna", "0", "0.0222766"]
    assert "Max p-value: 0" in r.stderr
    e = run(["triangle", "-E"] + files, cwd=tmp_path).stdout.splitlines()
    assert e[0].split("\t") == ["data/genome2.fna", "data/genome1.fna", "0.0222766", "0", "456/1000"]
    assert len(e) == 3 and e[2].split("\t")[4] == "456/1000"
    # fingerprints: 5 references of 2000 unsorted 32-bit hashes each
    gunzip_to("DNA1-CFL.txt.gz", tmp_path / "DNA1-CFL.txt")
    t = run(["triangle", "-fp", "-E", "DNA1-CFL.txt"], cwd=tmp_path).stdout.splitlines()
    rows = [l.split("\t") for l in t if l.count("\t") == 4]
    refs = mshpy.load(os.path.join(GOLDEN, "DNA1-sketch.msh")).refs
    want = []
    for i in range(1, 5):
        for j in range(i):
            a, b = refs[i]["hashes32"], refs[j]["hashes32"]
            m = min(len(a), len(b))
            matches = sum(1 for x, y in zip(a[:m], b[:m]) if x == y)
            p = math.erfc(math.sqrt(matches / 2.0))
            if 1.0 - matches / m <= 1.0 and p <= 1.0:
                want.append([refs[i]["name"], refs[j]["name"], "%g" % (1.0 - matches / m), "%g" % p, "%d/%d" % (matches, m)])
    assert rows == want


def test_sketch_protein_cli(tmp_path, oracle, fpm):
    rng = np.random.default_rng(4)
    aa = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYXB", dtype=np.uint8)
    seqs = [aa[rng.integers(0, len(aa), size=n)].tobytes() for n in (4000, 300)]
    with open(tmp_path / "p.faa", "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b">p%d\n" % i + s + b"\n")
    run(["sketch", "-a", "-i", "-s", "300", "p.faa"], cwd=tmp_path)
    m = mshpy.load(tmp_path / "p.faa.msh")
    assert (m.kmer_size, m.alphabet, m.noncanonical) == (9, "ACDEFGHIKLMNPQRSTVWY", True)
    table = fpm.nucleotide_alphabet("ACDEFGHIKLMNPQRSTVWY")
    for x, s in zip(m.refs, seqs):
        w = oracle.sketch([s], k=9, s=300, alphabet=table, noncanonical=True)
        assert x["hashes64"] == [int(v) for v in w["hashes"]]


def test_parallel_parsing_keeps_submission_order(tmp_path):
    """-p N parses files on N threads; the .msh must be byte-identical to the single-threaded run
    (the reference's ThreadPool returns outputs in submission order, ThreadPool.hxx:127-167)."""
    from util import random_dna
    rng = np.random.default_rng(21)
    names = []
    for i in range(23):
        nm = "g%02d.fa" % i
        with open(tmp_path / nm, "wb") as f:
            for r in range(1 + i % 3):
                f.write(b">g%d_%d c\n" % (i, r) + random_dna(rng, int(rng.integers(30, 60000))) + b"\n")
        names.append(nm)
    (tmp_path / "list.txt").write_text("\n".join(names) + "\n")
    run(["sketch", "-l", "-o", "one", "list.txt"], cwd=tmp_path)
    run(["sketch", "-l", "-p", "6", "-o", "six", "list.txt"], cwd=tmp_path)
    assert same_bytes(tmp_path / "one.msh", tmp_path / "six.msh")
    run(["sketch", "-i", "-p", "4", "-o", "ind4"] + names, cwd=tmp_path)
    run(["sketch", "-i", "-o", "ind1"] + names, cwd=tmp_path)
    assert same_bytes(tmp_path / "ind1.msh", tmp_path / "ind4.msh")
    assert len(mshpy.load(tmp_path / "ind4.msh").refs) == sum(1 + i % 3 for i in range(23))


def test_dist_filtered_output_is_the_filtered_full_output(tmp_path):
    """`mash dist -d D -v P` (list output) takes the hits path -- one GPU call, only passing pairs come back; its rows must
    be exactly the rows of the unfiltered run that satisfy the thresholds, in the same order (CommandDistance.cpp:303-333)."""
    rng = np.random.default_rng(77)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    roots = [lut[rng.integers(0, 4, size=30_000)] for _ in range(4)]
    names = []
    for i in range(36):
        g = roots[i % 4].copy()
        m = rng.random(g.size) < 0.002 * (1 + i // 4)
        g[m] = lut[rng.integers(0, 4, size=int(m.sum()))]
        names.append("g%02d.fa" % i)
        with open(tmp_path / names[-1], "wb") as f:
            f.write(b">g%d\n" % i + g.tobytes() + b"\n")
    run(["sketch", "-s", "300", "-o", "all"] + names, cwd=tmp_path)
    full = [l.split("\t") for l in run(["dist", "all.msh", "all.msh"], cwd=tmp_path).stdout.splitlines()]
    assert len(full) == 36 * 36
    for flags, keep in ((["-d", "0.02"], lambda r: float(r[2]) <= 0.02),
                        (["-v", "1e-30"], lambda r: float(r[3]) <= 1e-30),
                        (["-d", "0.05", "-v", "1e-10"], lambda r: float(r[2]) <= 0.05 and float(r[3]) <= 1e-10)):
        got = [l.split("\t") for l in run(["dist"] + flags + ["all.msh", "all.msh"], cwd=tmp_path).stdout.splitlines()]
        want = [r for r in full if keep(r)]
        assert 36 <= len(want) < 36 * 36 and got == want, flags


def test_thousand_reference_msh_layout(tmp_path, oracle):
    """configs[1] writes ONE .msh with 1000 references: the reference list leaves segment 0, so every name, comment and
    hash list sits behind a far pointer with its own landing pad (SURVEY.md Appendix D -- the layout the reference's
    capnp arena produces, unpinned by its fixtures).  Written by the product, read back by the independent decoder."""
    rng = np.random.default_rng(77)
    recs = [bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=3000)]) for _ in range(1000)]
    with open(tmp_path / "many.fa", "wb") as f:
        for i, r in enumerate(recs):
            f.write(b">seq%04d sample %d\n" % (i, i) + r + b"\n")
    run(["sketch", "-i", "-o", "many", "many.fa"], cwd=tmp_path)
    data = open(tmp_path / "many.msh", "rb").read()
    m = mshpy.Msh(data)
    assert len(m.refs) == 1000 and m.kmer_size == 21 and m.sketch_size == 1000
    # step 2 of the walk-through: segment 1 is exactly the landing pad + tag + 1000 x 9 words, and it is full
    assert m.segment_words[1] == 9002
    # step 3: later segments have capacities 10026, 20052, 40104, ... (nextSize doubling from 1024 + 9002); all but the last are
    # filled until the next object (plus its landing pad) no longer fits
    caps, nxt = [], 1024 + 9002
    for w in m.segment_words[2:]:
        caps.append(nxt)
        nxt += nxt
        assert w <= caps[-1]
    import struct
    order = []
    for i in range(1000):
        e_idx = 2 + 9 * i                                              # element i of the composite list in segment 1
        for slot, what in ((2, "name"), (3, "comment"), (5, "hashes64")):
            w = struct.unpack_from("<Q", m.segs[1], 8 * (e_idx + 2 + slot))[0]
            lo, hi = w & 0xffffffff, w >> 32
            assert lo & 3 == 2 and not (lo >> 2) & 1, "reference %d %s: not a single-far pointer" % (i, what)
            pad_seg, pad_idx = hi, lo >> 3
            assert pad_seg >= 2
            pad = struct.unpack_from("<Q", m.segs[pad_seg], 8 * pad_idx)[0]
            assert pad & 3 == 1 and ((pad & 0xffffffff) >> 2) == 0, "landing pad is a list pointer with offset 0: the object follows it"
            order.append((pad_seg, pad_idx))
        assert m.refs[i]["name"] == "seq%04d" % i and m.refs[i]["comment"] == "sample %d" % i and m.refs[i]["length"] == 3000
    assert order == sorted(order), "objects are allocated in call order name, comment, hashes per reference"
    # a segment is left only when the next object did not fit into what was left of its capacity
    for (sa, ia), (sb, ib) in zip(order, order[1:]):
        if sb != sa:
            assert sb == sa + 1 and ib == 0
    for seg_i, cap in zip(range(2, len(m.segment_words) - 1), caps):
        first_next = next(k for k, (s_, _) in enumerate(order) if s_ == seg_i + 1)
        kind = first_next % 3
        ref = m.refs[first_next // 3]
        size = 1 + ((len(ref["name"]) + 1 + 7) // 8 if kind == 0 else (len(ref["comment"]) + 1 + 7) // 8 if kind == 1 else len(ref["hashes64"]))
        assert m.segment_words[seg_i] + size > cap
    for i in (0, 1, 499, 999):
        want = oracle.sketch([recs[i]], k=21, s=1000)
        assert m.refs[i]["hashes64"] == [int(x) for x in want["hashes"]]
    # and the product's own reader agrees with what it wrote: paste reproduces the file
    run(["paste", "again", "many.msh"], cwd=tmp_path)
    assert same_bytes(tmp_path / "again.msh", tmp_path / "many.msh")


def test_cli_on_several_gpus_prints_the_same(tmp_path):
    """`mash sketch` / `mash dist` over all GPUs of the box (fpm_multi_*: query x reference blocks, whole sketches per GPU) must
    write the same files and print the same rows as on one GPU.  FPMASH_GPUS beyond the devices present repeats them."""
    rng = np.random.default_rng(5)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    base = lut[rng.integers(0, 4, size=120000)]
    names = []
    for i in range(14):
        g = base.copy() if i % 2 == 0 else lut[rng.integers(0, 4, size=90000)]
        idx = rng.random(g.size) < 0.01 * (1 + i)
        g[idx] = lut[rng.integers(0, 4, size=int(idx.sum()))]
        with open(tmp_path / ("g%02d.fa" % i), "wb") as f:
            f.write(b">g%d\n" % i + bytes(g) + b"\n")
        names.append("g%02d.fa" % i)

    def with_env(extra, args):
        env = dict(os.environ, **extra)
        r = subprocess.run([MASH] + args, capture_output=True, text=True, cwd=tmp_path, env=env)
        assert r.returncode == 0, r.stderr
        return r.stdout

    one = {"FPMASH_GPUS": "1"}
    many = {"FPMASH_GPUS": "3", "FPMASH_MULTI_MIN": "1", "FPMASH_MSH_STREAM": "0"}      # (sketch-file queries would otherwise be streamed on one GPU)
    loaded = {"FPMASH_GPUS": "1", "FPMASH_MSH_STREAM": "0"}
    with_env(one, ["sketch", "-o", "one"] + names)
    with_env(many, ["sketch", "-o", "many"] + names)
    assert same_bytes(tmp_path / "one.msh", tmp_path / "many.msh")
    for args in (["dist", "one.msh", "one.msh"], ["dist", "-d", "0.2", "one.msh", "one.msh"], ["dist", "-t", "one.msh", "one.msh"],
                 ["dist", "-v", "1e-20", "one.msh", "g03.fa", "g04.fa"]):
        want = with_env(one, args)                      # sketch-file queries streamed from the mapped file, resident reference panel
        assert want == with_env(many, args), args       # all GPUs, queries loaded
        assert want == with_env(loaded, args), args     # one GPU, queries loaded
        assert len(want) > 100


def test_streamed_queries_in_small_chunks(tmp_path):
    """`mash dist ref.msh q1.msh q2.msh`: the query files are mapped and streamed chunk by chunk against the resident reference
    panel.  Two query files, differing sketch sizes (the larger is reduced like loadCapnp does), names and comments kept."""
    rng = np.random.default_rng(8)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    base = lut[rng.integers(0, 4, size=60000)]
    for name, n in (("a", 40), ("b", 25), ("r", 12)):
        with open(tmp_path / (name + ".fa"), "wb") as f:
            for i in range(n):
                g = base.copy()
                idx = rng.random(g.size) < 0.002 * (1 + i % 7)
                g[idx] = lut[rng.integers(0, 4, size=int(idx.sum()))]
                f.write(b">%s%d the %dth of %s\n" % (name.encode(), i, i, name.encode()) + bytes(g) + b"\n")
    run(["sketch", "-i", "-s", "400", "-o", "r", "r.fa"], cwd=tmp_path)
    run(["sketch", "-i", "-s", "400", "-o", "a", "a.fa"], cwd=tmp_path)
    run(["sketch", "-i", "-s", "700", "-o", "b", "b.fa"], cwd=tmp_path)

    def out(env, args):
        r = subprocess.run([MASH] + args, capture_output=True, text=True, cwd=tmp_path, env=dict(os.environ, **env))
        assert r.returncode == 0, r.stderr
        return r.stdout, r.stderr

    for args in (["dist", "r.msh", "a.msh", "b.msh"], ["dist", "-C", "-d", "0.05", "r.msh", "a.msh", "b.msh"], ["dist", "-t", "r.msh", "b.msh", "a.msh"]):
        s_out, s_err = out({}, args)
        l_out, l_err = out({"FPMASH_MSH_STREAM": "0"}, args)
        assert s_out == l_out and len(s_out) > 1000, args
        assert ("will be reduced" in s_err) == ("will be reduced" in l_err)
