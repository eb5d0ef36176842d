"""GPU parity: rolling canonical k-mer + MurmurHash3 + bottom-s selection vs the oracle.
Every call goes through the C ABI (fpmash_b200 ctypes binding).  Bit-exact."""
import numpy as np
import pytest

from util import dirty_dna, mutate, random_dna

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k", [1, 2, 3, 7, 8, 9, 15, 16, 17, 21, 24, 25, 31, 32])
@pytest.mark.parametrize("noncanonical", [False, True])
def test_kmer_hash_stream_matches_getHash(ctx, oracle, k, noncanonical):
    rng = np.random.default_rng(1000 + k)
    rec = dirty_dna(rng, 20011)
    want = oracle.sketch([rec], k=k, s=10, noncanonical=noncanonical, trace=True)["trace"]
    got = ctx.kmer_hashes(rec, k=k, s=10, noncanonical=noncanonical)
    assert len(got) == len(want)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("k", list(range(1, 33)))
def test_production_kernel_hashes_every_window_like_getHash(ctx, oracle, k):
    """The sketch kernel itself (table-driven Murmur, lazy finish, threshold filter) with a sketch larger than the
    number of windows: the sketch then IS the set of all window hashes with their multiplicities."""
    rng = np.random.default_rng(3000 + k)
    recs = [dirty_dna(rng, 6000, n_rate=0.002), random_dna(rng, 2500)]
    for noncanonical in (False, True):
        got = ctx.sketch_records([recs], k=k, s=20000, noncanonical=noncanonical, want_counts=True)[0]
        want = oracle.sketch(recs, k=k, s=20000, noncanonical=noncanonical)
        assert np.array_equal(got["hashes"], want["hashes"]) and np.array_equal(got["counts"], want["counts"])


def test_kmer_hash_preserve_case_and_seed(ctx, oracle):
    rng = np.random.default_rng(7)
    rec = dirty_dna(rng, 5000, lower_rate=0.3)
    for seed in (0, 42, 12345):
        want = oracle.sketch([rec], k=21, s=10, seed=seed, preserve_case=True, trace=True)["trace"]
        got = ctx.kmer_hashes(rec, k=21, s=10, seed=seed, preserve_case=True)
        assert np.array_equal(got, want)


def test_pocket_known_answers(ctx):
    # SURVEY.md Appendix C (computed with the reference's own hash.cpp)
    kat = {b"ATGCATGCATGCATGCATGCA": 14844149108877162497, b"CATGCATGCATGCATGCATGC": 10703850894209713636,
           b"GCATGCATGCATGCATGCATG": 17987483124073101136, b"TGCATGCATGCATGCATGCAT": 11471179132836535170}
    for kmer, h in kat.items():
        assert int(ctx.kmer_hashes(kmer, k=21, s=1, noncanonical=True)[0]) == h
    assert int(ctx.kmer_hashes(b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA", k=32, s=1, noncanonical=True)[0]) == 7775287419336189913
    assert int(ctx.kmer_hashes(b"ACGTACGTACGTACGT", k=16, s=1, noncanonical=True)[0]) == 2886031495
    assert int(ctx.kmer_hashes(b"A", k=1, s=1, noncanonical=True)[0]) == 969040168


def _check_groups(ctx, oracle, groups, **kw):
    got = ctx.sketch_records(groups, want_counts=True, want_kmers=True, **kw)
    okw = {k: v for k, v in kw.items()}
    for g, recs in enumerate(groups):
        want = oracle.sketch(recs, trace=True, **okw)
        assert np.array_equal(got[g]["hashes"], want["hashes"]), "group %d hashes" % g
        assert np.array_equal(got[g]["counts"], want["counts"]), "group %d counts" % g
        assert got[g]["kmers"] == len(want["trace"]), "group %d window count" % g


@pytest.mark.parametrize("k,s", [(21, 1000), (16, 1000), (32, 10000), (11, 500), (21, 50)])
def test_sketch_genomes(ctx, oracle, k, s):
    rng = np.random.default_rng(k * 100 + s)
    base = random_dna(rng, 300000)
    groups = [[base], [mutate(rng, base, 0.01)], [dirty_dna(rng, 150000)],
              [random_dna(rng, 70000), random_dna(rng, 5), random_dna(rng, 90000)]]
    _check_groups(ctx, oracle, groups, k=k, s=s)


def test_sketch_small_and_empty_groups(ctx, oracle):
    rng = np.random.default_rng(5)
    groups = [[b"ACGT"], [random_dna(rng, 21)], [random_dna(rng, 20)], [b"N" * 100], [random_dna(rng, 1500)],
              [b"ACGTACGTACGTACGTACGTACGTACGT" * 50], [b"A" * 3000], [random_dna(rng, 40000)]]
    _check_groups(ctx, oracle, groups, k=21, s=1000)


@pytest.mark.parametrize("min_cov", [1, 2, 3])
def test_sketch_reads_min_cov_and_counts(ctx, oracle, min_cov):
    rng = np.random.default_rng(40 + min_cov)
    genome = random_dna(rng, 20000)
    reads = []
    for _ in range(4000):   # ~30x of 150 bp reads with errors, both strands
        p = int(rng.integers(0, len(genome) - 150))
        r = mutate(rng, genome[p:p + 150], 0.01)
        if rng.random() < 0.5:
            r = r[::-1].translate(bytes.maketrans(b"ACGT", b"TGCA"))
        if rng.random() < 0.02:
            r = r[:70] + b"N" + r[71:]
        reads.append(r)
    _check_groups(ctx, oracle, [reads], k=21, s=1000, min_cov=min_cov)
    _check_groups(ctx, oracle, [reads], k=21, s=100, min_cov=min_cov)   # full sketch: top-count rule


def test_sketch_repeats_trigger_rerun(ctx, oracle):
    # few distinct k-mers: the first threshold admits < s distinct hashes, forcing re-runs
    rng = np.random.default_rng(9)
    unit = random_dna(rng, 3000)
    groups = [[unit * 100], [random_dna(rng, 200000)]]
    _check_groups(ctx, oracle, groups, k=21, s=1000)


def test_noncanonical_sketch(ctx, oracle):
    rng = np.random.default_rng(11)
    _check_groups(ctx, oracle, [[random_dna(rng, 120000)]], k=21, s=1000, noncanonical=True)


def test_fp_hash_batch(ctx, oracle):
    rng = np.random.default_rng(3)
    lines = [list(rng.integers(1, 200, size=int(rng.integers(0, 25)))) for _ in range(3000)]
    lines[0] = [8, 34, 57, 1]
    lines[1] = [1, 1, 1, 1, 2, 34, 60]
    got = ctx.fp_hash_batch(lines, seed=42, use64=False)
    assert int(got[0]) == 819737709 and int(got[1]) == 2641509094   # DNA1-sketch.json
    for use64 in (False, True):
        got = ctx.fp_hash_batch(lines, seed=42, use64=use64)
        want = np.array([oracle.fp_hash(t, 42, use64) for t in lines], dtype=np.uint64)
        assert np.array_equal(got, want)


PROTEIN = "ACDEFGHIKLMNPQRSTVWY"


@pytest.mark.parametrize("alphabet,k,s", [(PROTEIN, 9, 1000), (PROTEIN, 5, 400), ("ACGTN", 21, 500), ("01", 32, 300), ("acgu", 16, 200)])
def test_sketch_custom_alphabets(ctx, oracle, fpm, alphabet, k, s):
    """-a / -z: noncanonical, every window whose k bytes are all in the alphabet (case folded)."""
    rng = np.random.default_rng(len(alphabet) * 100 + k)
    letters = np.frombuffer((alphabet.upper() + alphabet.lower() + "XZ*-").encode(), dtype=np.uint8)
    p = np.r_[np.full(2 * len(alphabet), 0.97 / (2 * len(alphabet))), np.full(4, 0.03 / 4)]
    groups = [[letters[rng.choice(len(letters), size=n, p=p)].tobytes() for n in sizes] for sizes in ([60000], [3000, 7, 12000], [k], [k - 1])]
    table = fpm.nucleotide_alphabet(alphabet)
    got = ctx.sketch_records(groups, k=k, s=s, alphabet=alphabet, noncanonical=True, want_counts=True, want_kmers=True)
    for g, recs in enumerate(groups):
        want = oracle.sketch(recs, k=k, s=s, alphabet=table, noncanonical=True, trace=True)
        assert np.array_equal(got[g]["hashes"], want["hashes"]) and np.array_equal(got[g]["counts"], want["counts"]), g
        assert got[g]["kmers"] == len(want["trace"])
    # reads-style multiplicity filter on the generic path
    recs = groups[0] * 2 + groups[1]
    got = ctx.sketch_records([recs], k=k, s=50, alphabet=alphabet, noncanonical=True, min_cov=2, want_counts=True)[0]
    want = oracle.sketch(recs, k=k, s=50, alphabet=table, noncanonical=True, min_cov=2)
    assert np.array_equal(got["hashes"], want["hashes"]) and np.array_equal(got["counts"], want["counts"])


def test_custom_alphabet_requires_noncanonical(ctx, fpm):
    with pytest.raises(fpm.FpmError) as e:
        ctx.sketch_records([[b"ACDEFGHIK" * 10]], k=9, s=10, alphabet=PROTEIN)
    assert e.value.code == fpm.FPM_ERR_UNSUPPORTED


def test_full_size_genome_and_large_sketch(ctx, oracle):
    """One 5 Mbp genome at the headline parameters, and k=32 s=10000 (config 5's sketch shape)."""
    rng = np.random.default_rng(2026)
    g = random_dna(rng, 5_000_000)
    for k, s in ((21, 1000), (32, 10000)):
        got = ctx.sketch_records([[g]], k=k, s=s, want_counts=True)[0]
        want = oracle.sketch([g], k=k, s=s)
        assert np.array_equal(got["hashes"], want["hashes"]) and np.array_equal(got["counts"], want["counts"])


@pytest.mark.parametrize("s,n", [(20000, 400000), (50000, 90000), (30000, 2000000)])
def test_sketch_sizes_beyond_shared_memory(ctx, oracle, s, n):
    """-s above ~12900: qualifying hashes no longer fit the shared-memory sort; the global-memory path takes over."""
    rng = np.random.default_rng(s)
    g = dirty_dna(rng, n, n_rate=0.001)
    got = ctx.sketch_records([[g], [g[: n // 3]]], k=21, s=s, want_counts=True)
    for rec, out in zip(([g], [g[: n // 3]]), got):
        want = oracle.sketch(rec, k=21, s=s)
        assert np.array_equal(out["hashes"], want["hashes"]) and np.array_equal(out["counts"], want["counts"])


def test_streaming_entry_points_match_batch(ctx, oracle):
    """fpm_sketch_stream_*: pieces of arbitrary size (cutting records anywhere) give the batch result."""
    rng = np.random.default_rng(77)
    genome = random_dna(rng, 30000)
    reads = [mutate(rng, genome[p:p + 150], 0.01) for p in rng.integers(0, len(genome) - 150, size=5000)]
    groups = [reads, [dirty_dna(rng, 200000)], [b"ACGT"], []]
    for piece, double_buffered in ((1000, False), (1 << 16, False), (1 << 22, False), (4096, True), (1 << 18, True)):
        got = ctx.sketch_stream(groups, piece=piece, double_buffered=double_buffered, k=21, s=500, min_cov=2, want_counts=True)
        for g, recs in zip(got, groups):
            want = oracle.sketch(recs, k=21, s=500, min_cov=2)
            assert np.array_equal(g["hashes"], want["hashes"]) and np.array_equal(g["counts"], want["counts"])
