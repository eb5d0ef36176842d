"""CPU: pin the oracle (oracle/mash_oracle.cpp) against the reference's own fixtures
(tests/golden) and against the reference's own translation units
(oracle/_ref, when built).  An oracle that fails here is not allowed to judge the CUDA path."""
import gzip
import json
import os
import re

import mpmath as mp
import numpy as np
import pytest

import mshpy
from conftest import GOLDEN

mp.mp.dps = 60


def gz_lines(name):
    with gzip.open(os.path.join(GOLDEN, name), "rt") as f:
        return f.read().splitlines()


def fastq_records(name):
    lines = gz_lines(name)
    return [lines[i + 1].encode() for i in range(0, len(lines), 4)]


def exact_tail(x, n, r):
    """P[Binomial(n, r) >= x] summed exactly at 60 digits."""
    r = mp.mpf(r)
    lo, hi = (x, n) if x > n * r else (0, x - 1)
    s = mp.mpf(0)
    t = mp.binomial(n, lo) * r ** lo * (1 - r) ** (n - lo)
    for i in range(lo, hi + 1):
        s += t
        t = t * (n - i) / (i + 1) * r / (1 - r)
    return s if x > n * r else 1 - s


# ---- pocket vectors (computed with the reference's hash.cpp) ------------
POCKET = [(b"ATGCATGCATGCATGCATGCA", 14844149108877162497, 2207119361, 18188561536538430162),
          (b"CATGCATGCATGCATGCATGC", 10703850894209713636, 1176372708, 14181391542666385171),
          (b"GCATGCATGCATGCATGCATG", 17987483124073101136, 2109077328, 14682935879223208092),
          (b"TGCATGCATGCATGCATGCAT", 11471179132836535170, 4153957250, 784558921633152789),
          (b"ACGTACGTACGTACGT", 4706917051267373191, 2886031495, 16250028995740358070),
          (b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA", 7775287419336189913, 1880680409, 4661435869264867092),
          (b"ACGTTGCAACGTTGCAACGTTGCAACGTTGCA", 11427927621569422678, 3263505750, 3401833214145748567),
          (b"A", 16750156190880784680, 969040168, 243126998722523514)]


def test_pocket_vectors(oracle):
    for kmer, h64, h32, h64_seed0 in POCKET:
        assert oracle.get_hash(kmer, 42, True) == h64
        assert oracle.get_hash(kmer, 42, False) == h32
        assert oracle.get_hash(kmer, 0, True) == h64_seed0
    assert oracle.fp_hash([8, 34, 57, 1], 42, True) == 6706282462166398061
    assert oracle.fp_hash([8, 34, 57, 1], 42, False) == 819737709
    assert oracle.fp_hash([1, 1, 1, 1, 2, 34, 60], 42, False) == 2641509094


def test_reference_generated_vectors(oracle):
    """ref_vectors.json was produced by the reference's own hash.cpp / MinHashHeap.cpp."""
    with open(os.path.join(GOLDEN, "ref_vectors.json")) as f:
        vec = json.load(f)
    for v in vec["hash"]:
        assert oracle.get_hash(v["kmer"].encode(), v["seed"], True) == int(v["h64"])
        assert oracle.get_hash(v["kmer"].encode(), v["seed"], False) == v["h32"]
    for v in vec["fp"]:
        assert oracle.fp_hash(v["tokens"], v["seed"], False) == v["h32"]
        assert oracle.fp_hash(v["tokens"], v["seed"], True) == int(v["h64"])
    for v in vec["heap"]:
        h, c, est = oracle.heap_stream(v["stream"], v["s"], v["m"], v["use64"])
        assert [int(x) for x in h] == v["hashes"] and [int(x) for x in c] == v["counts"]
        assert est == v["set_size"]
    for v in vec["sketch"]:
        out = oracle.sketch([r.encode() for r in v["records"]], k=v["k"], s=v["s"], min_cov=v["min_cov"])
        assert [str(int(x)) for x in out["hashes"]] == v["hashes"]
        assert [int(x) for x in out["counts"]] == v["counts"]
        assert out["set_size"] == v["set_size"]


def test_reads_fixture(oracle):
    """mash sketch -r reads1.fastq reads2.fastq: test/ref/reads.json (hashes + length) and
    new_data/reads/reads.msh (counts); records alternate between the two files."""
    r1, r2 = fastq_records("reads1.fastq.gz"), fastq_records("reads2.fastq.gz")
    recs = [x for pair in zip(r1, r2) for x in pair]
    out = oracle.sketch(recs, k=21, s=1000)
    txt = open(os.path.join(GOLDEN, "reads.json")).read()
    gold = [int(x) for x in re.findall(r"^\s*(\d+),?\s*$", txt, re.M)]
    assert len(gold) == 1000 and [int(x) for x in out["hashes"]] == gold
    assert int(out["set_size"]) == 502359 == int(re.search(r'"length" : (\d+)', txt).group(1))
    m = mshpy.load(os.path.join(GOLDEN, "reads.msh"))
    assert m.refs[0]["hashes"] == gold
    assert m.refs[0]["counts32"] == [int(x) for x in out["counts"]] and sum(m.refs[0]["counts32"]) == 1115
    assert m.refs[0]["length"] == 502359


def test_test_sequence_fixture(oracle):
    m = mshpy.load(os.path.join(GOLDEN, "test_sequence.msh"))
    recs = [l.encode() for l in open(os.path.join(GOLDEN, "test_sequence.fasta")).read().split("\n") if l and l[0] != ">"]
    out = oracle.sketch(recs, k=21, s=1000)
    assert m.refs[0]["hashes"] == [int(x) for x in out["hashes"]] == [10703850894209713636, 14844149108877162497]
    assert m.refs[0]["length"] == out["length"] == 73
    assert m.refs[0]["comment"] == "[2 seqs] sequence1 taxid 1 [...]"


@pytest.mark.parametrize("n", [1, 2, 3])
def test_fingerprint_fixtures(oracle, n):
    """mash sketch -fp DNAn-CFL.txt: one 32-bit hash per line, unsorted, untruncated; length rule."""
    m = mshpy.load(os.path.join(GOLDEN, "DNA%d-sketch.msh" % n))
    assert m.kmer_size == 1 and m.alphabet == "0123456789" and m.noncanonical and len(m.refs) == 5
    lines = gz_lines("DNA%d-CFL.txt.gz" % n)
    by_id, order = {}, []
    for ln in lines:
        f = ln.split()
        if f[0] not in by_id:
            by_id[f[0]] = []
            order.append(f[0])
        by_id[f[0]].append([int(x) for x in f[1:]])
    assert [r["name"] for r in m.refs] == order
    for r in m.refs:
        rows = by_id[r["name"]]
        assert r["hashes32"] == [oracle.fp_hash(t, 42, False) for t in rows]
        assert r["length"] == len(rows[0]) + sum(len(t) for t in rows)       # Sketch.cpp:117,134
        assert r["comment"] == "FingerPrint : " + r["name"]


def test_dist_known_answers(oracle):
    """test/ref/genomes.dist (3 rows, 6 significant digits) and tutorials.rst:24,56-57."""
    g = [mshpy.load(os.path.join(GOLDEN, "genome%d.fna.msh" % i)).refs[0] for i in (1, 2, 3)]
    reads = mshpy.load(os.path.join(GOLDEN, "reads.msh")).refs[0]
    rows = [l.split("\t") for l in open(os.path.join(GOLDEN, "genomes.dist")).read().splitlines()]
    ks = 4.0 ** 21
    for gi, row in zip(g, rows):
        out = oracle.compare(gi["hashes"], reads["hashes"], gi["length"], reads["length"], 1000, 21, ks)
        assert "%d/%d" % (out["numer"], out["denom"]) == row[4]
        assert "%g" % out["distance"] == row[2]
        assert "%g" % out["pvalue"] == row[3]
    out = oracle.compare(g[0]["hashes"], g[1]["hashes"], g[0]["length"], g[1]["length"], 1000, 21, ks)
    assert (out["numer"], out["denom"], "%g" % out["distance"], "%g" % out["pvalue"]) == (456, 1000, "0.0222766", "0")
    out = oracle.compare(g[0]["hashes"], g[2]["hashes"], g[0]["length"], g[2]["length"], 1000, 21, ks)
    assert (out["numer"], out["denom"], out["distance"], out["pvalue"]) == (1000, 1000, 0.0, 0.0)


def test_pvalue_against_exact_tail(oracle):
    cases = [(41, 1000, 1e-6), (35, 1000, 2.3e-6), (1, 1000, 1e-9), (999, 1000, 0.998), (500, 1000, 0.998), (1000, 1000, 0.5),
             (3, 10000, 2e-5), (9000, 10000, 0.9), (5000, 10000, 0.4999), (100, 2000, 0.05), (101, 2000, 0.05), (99, 2000, 0.05),
             (7, 37, 0.2), (1, 1, 0.3), (10000, 10000, 0.9999), (2, 400, 1e-12)]
    for x, n, r in cases:
        want = exact_tail(x, n, r)
        got = oracle.binom_tail(x, n, r)
        if want < mp.mpf("1e-300"):
            assert got < 1e-299
        else:
            assert abs(mp.mpf(got) - want) <= mp.mpf("1e-13") * want, (x, n, r)


def test_compare_closed_form(oracle):
    """compareSketches on sorted duplicate-free lists equals the set formulation of SURVEY.md a9."""
    rng = np.random.default_rng(3)
    for _ in range(300):
        s = int(rng.integers(1, 40))
        a = np.unique(rng.integers(0, 60, size=int(rng.integers(0, 50)))).astype(np.uint64)[:s + 5]
        b = np.unique(rng.integers(0, 60, size=int(rng.integers(0, 50)))).astype(np.uint64)[:s + 5]
        out = oracle.compare(a, b, 1000, 1000, s, 21, 4.0 ** 21)
        u = np.union1d(a, b)
        denom = min(s, len(u))
        common = len(np.intersect1d(np.intersect1d(a, b), u[:denom]))
        assert (out["numer"], out["denom"]) == (common, denom)


def test_compare_without_a_shared_hash_is_closed_form(oracle):
    """What the GPU path's pruning relies on: when the two lists share no hash the reference loop yields
    common = 0, denom = min(s, |A| + |B|), distance 1 and p-value 1 -- except two empty lists: 0/0, distance 0."""
    rng = np.random.default_rng(17)
    for _ in range(300):
        s = int(rng.integers(1, 60))
        u = rng.permutation(200)[:int(rng.integers(0, 90))].astype(np.uint64)
        cut = int(rng.integers(0, len(u) + 1))
        a, b = np.sort(u[:cut]), np.sort(u[cut:])                    # disjoint by construction, any sizes incl. empty
        for md, mp_ in ((1.0, 1.0), (0.3, 1.0), (1.0, 1e-3)):
            out = oracle.compare(a, b, 5000, 7000, s, 21, 4.0 ** 21, md, mp_)
            assert (out["numer"], out["denom"]) == (0, min(s, len(a) + len(b)))
            if len(a) + len(b) == 0:
                assert out["distance"] == 0.0 and out["passed"] == (mp_ >= 1.0)
                if out["passed"]:
                    assert out["pvalue"] == 1.0
            else:
                assert out["distance"] == 1.0 and out["passed"] == (md >= 1.0 and mp_ >= 1.0)
                if md >= 1.0:
                    assert out["pvalue"] == 1.0


# ---- the oracle vs the reference's own code (only where oracle/_ref was built) ---------------
def test_oracle_matches_reference_heap(oracle, reflib):
    rng = np.random.default_rng(11)
    for t in range(300):
        s, m, use64 = int(rng.integers(1, 10)), int(rng.integers(1, 4)), bool(t & 1)
        stream = rng.integers(0, 30, size=int(rng.integers(0, 200))).astype(np.uint64)
        a = oracle.heap_stream(stream, s, m, use64)
        b = reflib.heap_stream(stream, s, m, use64)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]


def test_oracle_matches_reference_sketch(oracle, reflib):
    from util import dirty_dna
    rng = np.random.default_rng(12)
    for k, s, m, nc in [(21, 500, 1, False), (16, 300, 1, False), (32, 1000, 1, True), (21, 100, 2, False), (5, 2000, 1, False)]:
        recs = [dirty_dna(rng, int(rng.integers(5, 30000))) for _ in range(4)]
        if m > 1:
            recs += recs[:2]
        a = oracle.sketch(recs, k=k, s=s, min_cov=m, noncanonical=nc)
        b = reflib.sketch(recs, k=k, s=s, min_cov=m, noncanonical=nc)
        assert np.array_equal(a["hashes"], b["hashes"]) and np.array_equal(a["counts"], b["counts"])
        assert a["set_size"] == b["set_size"] and a["multiplicity"] == b["multiplicity"]
