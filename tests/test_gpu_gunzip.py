"""gzip'ed FASTA inflated on the GPU (csrc/gunzip.cu; the reference reads through zlib's gzread, Sketch.cpp:1340-1346):
byte-identical with zlib for every block type / header field / multi-member layout, refusal of everything zlib refuses,
and `mash sketch` over .gz files = the same .msh through the GPU route and through the host's zlib route."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import gz_cases
from conftest import ROOT

pytestmark = pytest.mark.gpu
MASH = os.path.join(ROOT, "fp-mash_b200", "bin", "mash")


def test_gunzip_batch_equals_zlib(ctx):
    cases = gz_cases.good_cases(big=True)
    names = sorted(cases)
    # every case alone (claimed size right or wrong: multi-member and padded files take the second pass) ...
    for nm in names:
        gz, want = cases[nm]
        if b"\0" in want:
            continue
        status, ends, raw = ctx.gunzip_batch([gz])
        assert status == 0, nm
        assert bytes(raw) == want + b"\0", nm
        assert int(ends[0]) == len(want), nm
    # ... and all of them as one batch, files at odd offsets of the compressed buffer
    files = [cases[nm] for nm in names if b"\0" not in cases[nm][1]]
    status, ends, raw = ctx.gunzip_batch([g for g, _ in files])
    assert status == 0
    assert bytes(raw) == b"".join(w + b"\0" for _, w in files)
    assert [int(e) for e in ends] == list(np.cumsum([len(w) + 1 for _, w in files]) - 1)
    assert ctx.gunzip_batch([])[0] == 0
    # a stream that expands far beyond the first layout's guess (16 : 1), and one whose trailer is followed by padding that
    # reads as an absurd size claim: the measuring pass settles both
    runs = b">n\n" + b"N" * 3_000_000 + b"\n"
    padded = gzip.compress(cases["dna_l6"][1][:50000]) + b"\0" * 7 + b"\xff\xff\xff\x7f"
    status, ends, raw = ctx.gunzip_batch([gzip.compress(runs, 9), padded])
    assert status == 0 and bytes(raw) == runs + b"\0" + cases["dna_l6"][1][:50000] + b"\0"


def test_gunzip_batch_refuses_what_zlib_refuses(ctx):
    good = gz_cases.good_cases()["dna_l6"][0]
    for nm, gz in gz_cases.bad_cases().items():
        with pytest.raises(Exception):
            gzip.decompress(gz)
        assert ctx.gunzip_batch([gz])[0] == 1, nm
        assert ctx.gunzip_batch([good, gz, good])[0] == 1, nm          # one damaged file fails the batch: the host reader takes it
    # a decompressed 0x00 would read as a file boundary in the raw batch
    assert ctx.gunzip_batch([gzip.compress(b">a\nACGT\0ACGT\n")])[0] == 2
    # the context is usable afterwards
    assert ctx.gunzip_batch([good])[0] == 0


def test_gunzip_many_files_then_parse_and_sketch(ctx, fpm):
    """A collection: 300 .gz genomes of different sizes and compression levels in one batch -> records, headers and
    sketches identical to the same files given uncompressed."""
    rng = np.random.default_rng(12)
    plain = []
    for i in range(300):
        n_rec = int(rng.integers(1, 4))
        plain.append(b"".join(gz_cases.fasta_text(int(rng.integers(200, 60000)), 1000 * i + j, width=int(rng.integers(50, 90)), name="g%d_r%d" % (i, j))
                              for j in range(n_rec)))
    gz = [gzip.compress(p, int(rng.integers(1, 10))) for p in plain]
    status, ends, raw = ctx.gunzip_batch(gz)
    assert status == 0
    assert bytes(raw) == b"".join(p + b"\0" for p in plain)
    got = ctx.fasta_parse_resident(len(raw))
    want = ctx.fasta_parse(plain)
    assert got is not None and want is not None
    # (fasta_parse just replaced the resident batch by its own upload of the same bytes: compare, then redo the gz route for the sketches)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and bytes(got[2]) == bytes(want[2])
    whole = b"".join(p + b"\0" for p in plain)
    for r, h in zip(got[0], got[3]):
        assert h == whole[int(r["hdr_begin"]):int(r["hdr_end"])]
    assert got[3][0].startswith(b">g0_r0 some comment")
    status, ends, _ = ctx.gunzip_batch(gz, fetch=False)
    recs, lengths, _, _ = ctx.fasta_parse_resident(int(ends[-1]) + 1, fetch_sequence=False, fetch_headers=False)
    goff = np.append(recs["seq_begin"], np.uint64(int(recs["seq_begin"][-1]) + int(lengths[-1]) + 1)).astype(np.uint64)
    p = fpm.make_sketch_params(k=21, s=200)
    a = ctx.sketch_parsed(goff, p)
    ctx.fasta_parse(plain, fetch_sequence=False)
    b = ctx.sketch_parsed(goff, p)
    assert np.array_equal(a["hashes"], b["hashes"]) and np.array_equal(a["n"], b["n"])


@pytest.mark.parametrize("individual", [False, True])
def test_cli_gz_gpu_route_equals_zlib_route(tmp_path, individual):
    """`mash sketch` over .gz files: inflated on the GPU (FPMASH_GPU_GUNZIP=1 takes the route for any number of files)
    = read through zlib on the host (=0): same .msh bytes, incl. a multi-member file, a FASTQ .gz and a damaged file in
    the batch (which send it to the host reader)."""
    rng = np.random.default_rng(13)
    names = []
    for i in range(12):
        text = b"".join(gz_cases.fasta_text(int(rng.integers(100, 30000)), 50 * i + j, name="c%d_%d" % (i, j)) for j in range(int(rng.integers(1, 4))))
        nm = "g%02d.fa.gz" % i
        if i == 5:
            (tmp_path / nm).write_bytes(gzip.compress(text[:1000]) + gzip.compress(text[1000:]))           # two members
        else:
            (tmp_path / nm).write_bytes(gzip.compress(text, int(rng.integers(1, 10))))
        names.append(nm)
    (tmp_path / "plain_inside.fa").write_bytes(gz_cases.fasta_text(5000, 777))
    opts = ["-k", "16", "-s", "100"] + (["-i"] if individual else [])
    env1 = dict(os.environ, FPMASH_GPU_GUNZIP="1")
    env0 = dict(os.environ, FPMASH_GPU_GUNZIP="0")
    def both(files, tag):
        a = subprocess.run([MASH, "sketch"] + opts + ["-o", "gpu_" + tag] + files, cwd=tmp_path, capture_output=True, text=True, env=env1)
        b = subprocess.run([MASH, "sketch"] + opts + ["-o", "host_" + tag] + files, cwd=tmp_path, capture_output=True, text=True, env=env0)
        assert a.returncode == b.returncode, (a.stderr, b.stderr)
        if a.returncode == 0:
            assert (tmp_path / ("gpu_%s.msh" % tag)).read_bytes() == (tmp_path / ("host_%s.msh" % tag)).read_bytes(), tag
        else:
            # (the batch route announces the files it queued behind the damaged one before it fails)
            assert a.stderr.strip().splitlines()[-1] == b.stderr.strip().splitlines()[-1], tag
        return a.returncode
    assert both(names, "all") == 0
    assert both(names[:4] + ["plain_inside.fa"] + names[4:], "mixed") == 0          # gz batch, plain batch, gz batch: input order kept
    (tmp_path / "reads.fq.gz").write_bytes(gzip.compress(b"@r1\n" + b"ACGTTGCA" * 10 + b"\n+\n" + b"I" * 80 + b"\n"))
    assert both(names[:3] + ["reads.fq.gz"] + names[3:6], "with_fastq") == 0
    g = bytearray((tmp_path / names[2]).read_bytes()); g[-6] ^= 0x40
    (tmp_path / "damaged.fa.gz").write_bytes(bytes(g))                     # CRC mismatch: whatever zlib's gzread makes of it (the batch goes to the host reader)
    both(names[:2] + ["damaged.fa.gz"] + names[3:5], "with_damaged")
