"""GPU FASTA ingestion against the host reader that mirrors the reference's kseq.h
(`mash debug-parse`, itself pinned to kseq.h by tests/test_host_cpu.py): same records, same lengths, same bytes."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
MASH = os.path.join(ROOT, "fp-mash_b200", "bin", "mash")


def host_records(path):
    """(name+comment text is checked by the CLI tests) -> list of (length, fnv1a of the sequence bytes)."""
    out = subprocess.run([MASH, "debug-parse", path], capture_output=True, check=True).stdout.split(b"\n")
    while out and out[-1] == b"":
        out.pop()
    assert out[-1].startswith(b"END\t-1"), out[-1]
    recs = []
    for line in out[:-1]:
        f = line.split(b"\t")
        recs.append((int(f[-2]), int(f[-1])))
    return recs


def fnv(b):
    h = 1469598103934665603
    for c in bytes(b):
        h = ((h ^ c) * 1099511628211) & 0xffffffffffffffff
    return h


def fasta_cases(rng):
    lut = np.frombuffer(b"ACGTacgtNRYKM", dtype=np.uint8)
    def seq(n):
        return lut[rng.integers(0, len(lut), size=n)].tobytes()
    def wrap(s, w, eol=b"\n"):
        return eol.join(s[i:i + w] for i in range(0, len(s), w)) + eol
    cases = {
        "plain": b">r1 first record\n" + wrap(seq(1000), 70) + b">r2\n" + wrap(seq(333), 70),
        "crlf": b">r1 c\r\n" + wrap(seq(500), 60, b"\r\n") + b">r2\r\n" + wrap(seq(61), 60, b"\r\n"),
        "no_final_newline": b">a\n" + seq(200),
        "header_only_at_eof": b">a\n" + seq(50) + b"\n>b",
        "header_line_at_eof": b">a\n" + seq(50) + b"\n>b comment\n",
        "empty_records": b">a\n>b\n\n>c\n" + seq(10) + b"\n",
        "junk_before_first": b"junk line\n; comment\n>a x\n" + wrap(seq(100), 50),
        "blank_and_spaces": b">a\n" + seq(30) + b"\n\n  " + seq(20) + b" \t" + seq(7) + b"\n\n",
        "gt_mid_line": b">a\nACGT" + seq(40) + b">b glued header\n" + seq(25) + b"\n",
        "gt_inside_header": b">a has > inside > twice\n" + seq(33) + b"\n",
        "long_single_line": b">chr one line\n" + seq(300_000) + b"\n>tail\n" + seq(5000) + b"\n",
        "many_short": b"".join(b">r%d\n" % i + seq(int(rng.integers(0, 90))) + b"\n" for i in range(3000)),
        "control_bytes": b">a\n" + seq(20) + bytes([1, 2, 127, 128, 200, 255]) + seq(20) + b"\n",
        "empty_file": b"",
        "no_records": b"just text\nno headers here\n",
    }
    # boundaries of the 4 KB chunks and of the 128-byte slices: headers and newlines straddling them
    for pad in (4095, 4096, 4097, 127, 128, 129, 8191):
        cases["straddle_%d" % pad] = b">x\n" + seq(pad - 3) + b">boundary header that is long enough to cross\n" + seq(200) + b"\n"
    return cases


def test_fasta_parse_matches_host_reader(ctx, tmp_path):
    rng = np.random.default_rng(3)
    cases = fasta_cases(rng)
    names = sorted(cases)
    want = []
    for nm in names:
        p = tmp_path / (nm + ".fa")
        p.write_bytes(cases[nm])
        want.append(host_records(str(p)))
    # every file alone, then all of them in one call (0x00-separated)
    for group in [[nm] for nm in names] + [names]:
        got = ctx.fasta_parse([cases[nm] for nm in group])
        assert got is not None, group
        recs, lengths, seq = got
        exp = [r for nm in group for r in want[names.index(nm)]]
        assert len(recs) == len(exp), (group, len(recs), len(exp))
        for i, (l, h) in enumerate(exp):
            assert int(lengths[i]) == l, (group, i)
            b = int(recs["seq_begin"][i])
            assert fnv(seq[b:b + l]) == h, (group, i)
            assert seq[b + l] == 0
        assert int(seq.size) == sum(l + 1 for l, _ in exp)
    # headers: the table points at them in the raw bytes
    raw = b"".join(cases[nm] + b"\0" for nm in names)
    recs, _, _ = ctx.fasta_parse([cases[nm] for nm in names], fetch_sequence=False)
    for r in recs:
        assert raw[int(r["hdr_begin"])] == ord(">")
        assert raw[int(r["hdr_end"])] in (ord("\n"), 0)
        assert b"\n" not in raw[int(r["hdr_begin"]):int(r["hdr_end"])]


def test_fasta_parse_rejects_fastq(ctx):
    assert ctx.fasta_parse([b"@r1\nACGT\n+\nIIII\n"]) is None
    assert ctx.fasta_parse([b">r1\nACGT\n+\nIIII\n"]) is None
    assert ctx.fasta_parse([b">ok\nACGT\n", b">r\nAC@GT\n"]) is None
    assert ctx.fasta_parse([b">plus + and @ in the header are fine\nACGT\n"]) is not None


def test_sketch_parsed_equals_sketch_batch(ctx, fpm):
    rng = np.random.default_rng(4)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    files, batch, goff = [], bytearray(), [0]
    for f in range(6):
        recs = [lut[rng.integers(0, 4, size=int(rng.integers(500, 40000)))].tobytes() for _ in range(int(rng.integers(1, 4)))]
        text = b"".join(b">f%d_r%d\n" % (f, i) + b"\n".join(r[j:j + 80] for j in range(0, len(r), 80)) + b"\n" for i, r in enumerate(recs))
        files.append(text)
        for r in recs:
            batch += r + b"\0"
        goff.append(len(batch))
    recs, lengths, seq = ctx.fasta_parse(files)
    assert bytes(seq) == bytes(batch)
    p = fpm.make_sketch_params(k=21, s=400)
    a = ctx.sketch_parsed(np.array(goff, dtype=np.uint64), p)
    b = ctx.sketch_batch(np.frombuffer(bytes(batch), dtype=np.uint8), np.array(goff, dtype=np.uint64), p)
    assert np.array_equal(a["hashes"], b["hashes"]) and np.array_equal(a["n"], b["n"])


@pytest.mark.parametrize("individual", [False, True])
def test_cli_gpu_parser_route_equals_host_reader_route(tmp_path, individual):
    """`mash sketch` through the GPU parser and through the host reader (FPMASH_GPU_PARSE=0): same .msh bytes, or
    the same failure, for every tricky input -- names, comments, [N seqs], lengths, short-record rules, -i."""
    rng = np.random.default_rng(5)
    cases = fasta_cases(rng)
    cases["two_comments_then_none"] = b">a first comment\n" + b"ACGT" * 20 + b"\n>b\n" + b"GATTACA" * 12 + b"\n>c  spaced  comment \n" + b"TTGCA" * 15 + b"\n"
    cases["lone_gt_at_eof"] = b">a\n" + b"ACGT" * 20 + b"\n>"
    cases["name_then_eof"] = b">a\n" + b"ACGT" * 20 + b"\n>b "
    for nm, data in sorted(cases.items()):
        (tmp_path / (nm + ".fa")).write_bytes(data)
    opts = ["-k", "15", "-s", "50"] + (["-i"] if individual else [])
    env0 = dict(os.environ, FPMASH_GPU_PARSE="0")
    failing = ("empty_file", "no_records", "empty_records")     # the reference exits with an error on these (every CLI start costs
    for nm in failing:                                            # seconds of CUDA initialisation, so only they run one by one)
        a = subprocess.run([MASH, "sketch"] + opts + ["-o", "gpu_" + nm, nm + ".fa"], cwd=tmp_path, capture_output=True, text=True)
        b = subprocess.run([MASH, "sketch"] + opts + ["-o", "host_" + nm, nm + ".fa"], cwd=tmp_path, capture_output=True, text=True, env=env0)
        assert a.returncode == b.returncode, (nm, a.stderr, b.stderr)
        if a.returncode == 0:
            assert (tmp_path / ("gpu_%s.msh" % nm)).read_bytes() == (tmp_path / ("host_%s.msh" % nm)).read_bytes(), nm
        else:
            assert a.stderr.replace("gpu_", "") == b.stderr.replace("host_", ""), nm
    # all other files in one command (one raw batch, several files), then with a FASTQ file in the middle (the whole
    # batch falls back to the host reader)
    good = [nm + ".fa" for nm in sorted(cases) if nm not in failing]
    (tmp_path / "reads.fq").write_bytes(b"@r1\n" + b"ACGTTGCA" * 10 + b"\n+\n" + b"I" * 80 + b"\n")
    for extra in ([], ["reads.fq"]):
        files = good[:5] + extra + good[5:]
        a = subprocess.run([MASH, "sketch"] + opts + ["-o", "gpu_all"] + files, cwd=tmp_path, capture_output=True, text=True)
        b = subprocess.run([MASH, "sketch"] + opts + ["-o", "host_all"] + files, cwd=tmp_path, capture_output=True, text=True, env=env0)
        assert a.returncode == 0 and b.returncode == 0, (a.stderr, b.stderr)
        assert (tmp_path / "gpu_all.msh").read_bytes() == (tmp_path / "host_all.msh").read_bytes()


def _fastq(rng, n, lens=(150, 150), n_rate=0.002):
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    recs, text = [], []
    for i in range(n):
        l = int(rng.integers(lens[0], lens[1] + 1))
        r = bytearray(lut[rng.integers(0, 4, size=l)].tobytes())
        for p in np.nonzero(rng.random(l) < n_rate)[0]:
            r[p] = ord("N")
        q = bytes(rng.integers(33, 75, size=l, dtype=np.uint8))      # quality bytes incl. '@' '+' '>' at line starts
        if i % 7 == 0 and l:
            q = b"@" + q[1:]
        if i % 11 == 0 and l:
            q = b"+" + q[1:]
        recs.append(bytes(r))
        text.append(b"@read%d some comment %d\n" % (i, i) + bytes(r) + b"\n+" + (b"read%d" % i if i % 3 == 0 else b"") + b"\n" + q + b"\n")
    return recs, text


def test_fastq_gpu_stream_equals_host_records(ctx):
    """Clean four-line FASTQ parsed on the GPU = the same reads streamed from the host (sketch with -m 2 and counts, whose
    largest element's count depends on the order of the reads)."""
    rng = np.random.default_rng(8)
    genome_reads, text = _fastq(rng, 3000, lens=(10, 180))
    genome_reads = genome_reads + genome_reads[:1500]                 # repeated reads: counts >= 2
    text = text + text[:1500]
    kw = dict(k=21, s=300, min_cov=2, want_counts=True)
    want = ctx.sketch_stream([genome_reads], **kw)[0]
    # one piece, then uneven pieces cut at record boundaries
    for cuts in ([0, len(text)], [0, 1, 2, 700, 701, 2999, len(text)]):
        pieces = [b"".join(text[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
        got, infos = ctx.sketch_fastq_pieces(pieces, **kw)
        assert np.array_equal(got["hashes"], want["hashes"]) and np.array_equal(got["counts"], want["counts"])
        assert sum(i[0] for i in infos) == len(text)
        assert sum(i[1] for i in infos) == sum(len(r) >= 21 for r in genome_reads)
        a0 = cuts[0]
        first = next(i for i, r in enumerate(genome_reads[a0:cuts[1]]) if len(r) >= 21)
        assert infos[0][3] == first
    # the line table: header text of the first read of the last piece
    ends = ctx.fastq_line_ends(0, 4)
    piece = pieces[-1]
    assert piece[:int(ends[0])].startswith(b"@read") and piece[int(ends[1]) + 1:int(ends[1]) + 2] == b"+"


def test_fastq_gpu_rejects_what_the_general_reader_treats_differently(ctx):
    ok = b"@r1 c\nACGTACGTACGTACGTACGTACGTAC\n+\nIIIIIIIIIIIIIIIIIIIIIIIIII\n"
    assert ctx.sketch_fastq_pieces([ok], k=21, s=10) is not None
    # CRLF files: the reader drops the '\r' of sequence and quality lines; so does the device parser
    crlf = ctx.sketch_fastq_pieces([ok.replace(b"\n", b"\r\n") * 3], k=21, s=10)
    lf = ctx.sketch_fastq_pieces([ok * 3], k=21, s=10)
    assert crlf is not None and np.array_equal(crlf[0]["hashes"], lf[0]["hashes"]) and crlf[1][0][:2] == lf[1][0][:2] == (3, 3)
    short = b"@r1\r\nACGTACGTACGTACGTACGT\r\n+\r\nIIIIIIIIIIIIIIIIIIII\r\n"       # 20 bases + '\r': not a read of 21
    assert ctx.sketch_fastq_pieces([short], k=21, s=10)[1][0][:2] == (1, 0)
    bad = {
        "cr_inside_sequence": b"@r1\nAC\rGT\n+\nIIII\n",
        "cr_only_on_the_sequence_line": b"@r1\nACGT\r\n+\nIIIII\n",
        "fasta": b">r1\nACGT\n>r2\nACGT\n",
        "short_quality": b"@r1\nACGTACGT\n+\nIIII\n",
        "long_quality": b"@r1\nACGT\n+\nIIIIII\n",
        "multi_line_sequence": b"@r1\nACGT\nACGT\n+\nIIIIIIII\n",
        "no_plus": b"@r1\nACGT\n-\nIIII\n",
        "no_at": b"r1\nACGT\n+\nIIII\n",
        "blank_in_sequence": b"@r1\nAC GT\n+\nIIIII\n",
        "plus_in_sequence": b"@r1\nAC+GT\n+\nIIIII\n",
        "three_lines": b"@r1\nACGT\n+\n",
        "empty_header": b"\nACGT\n+\nIIII\n",
        "second_record_broken": ok + b"@r2\nACGT\n+\nII\n",
    }
    for nm, data in bad.items():
        assert ctx.sketch_fastq_pieces([data], k=21, s=10) is None, nm


def test_cli_read_mode_gpu_fastq_route_equals_host_reader_route(tmp_path):
    """`mash sketch -r -m 2` on one plain FASTQ file: GPU route (pieces parsed on the device) and host reader write the
    same .msh, byte for byte -- hashes, counts (order-dependent top count), comment with [N seqs], estimated length.
    Also CRLF FASTQ (GPU route as well) and FASTA reads (the FASTA parser's route)."""
    import gzip
    from conftest import GOLDEN
    data = gzip.open(os.path.join(GOLDEN, "reads1.fastq.gz"), "rb").read()
    (tmp_path / "reads1.fastq").write_bytes(data)
    (tmp_path / "no_final_newline.fastq").write_bytes(data.rstrip(b"\n"))
    (tmp_path / "crlf.fastq").write_bytes(data[:200000].rsplit(b"\n@", 1)[0].replace(b"\n", b"\r\n") + b"\r\n")
    rng = np.random.default_rng(9)
    recs, text = _fastq(rng, 400, lens=(5, 160))                      # some reads shorter than k, incl. the first ones
    (tmp_path / "short_first.fastq").write_bytes(b"@tiny first comment\nACGT\n+\nIIII\n@tiny2\nAC\n+\nII\n" + b"".join(text) * 2)   # every read twice: -m 2
    (tmp_path / "reads.fa").write_bytes(b"".join(b">r%d\n%s\n" % (i, r) for i, r in enumerate(recs)) * 2)
    env0 = dict(os.environ, FPMASH_GPU_PARSE="0")
    for nm in ("reads1.fastq", "no_final_newline.fastq", "crlf.fastq", "short_first.fastq", "reads.fa"):
        a = subprocess.run([MASH, "sketch", "-r", "-m", "2", "-k", "21", "-s", "500", "-o", "gpu_" + nm, nm], cwd=tmp_path, capture_output=True, text=True)
        b = subprocess.run([MASH, "sketch", "-r", "-m", "2", "-k", "21", "-s", "500", "-o", "host_" + nm, nm], cwd=tmp_path, capture_output=True, text=True, env=env0)
        assert a.returncode == b.returncode == 0, (nm, a.stderr, b.stderr)
        assert (tmp_path / ("gpu_%s.msh" % nm)).read_bytes() == (tmp_path / ("host_%s.msh" % nm)).read_bytes(), nm
        assert a.stderr.replace("gpu_", "") == b.stderr.replace("host_", ""), nm
    # the same file in 64 KB pieces: every piece boundary goes through the record-boundary search
    envp = dict(os.environ, FPMASH_FASTQ_PIECE="65536")
    c = subprocess.run([MASH, "sketch", "-r", "-m", "2", "-k", "21", "-s", "500", "-o", "pieces", "reads1.fastq"], cwd=tmp_path, capture_output=True, text=True, env=envp)
    assert c.returncode == 0, c.stderr
    assert (tmp_path / "pieces.msh").read_bytes() == (tmp_path / "host_reads1.fastq.msh").read_bytes().replace(b"host_reads1.fastq", b"reads1.fastq") or \
        (tmp_path / "pieces.msh").read_bytes() == (tmp_path / "gpu_reads1.fastq.msh").read_bytes()
    # and the routes are the ones claimed: the timing trace names them
    envt = dict(os.environ, FPMASH_TIMING="1")
    t = subprocess.run([MASH, "sketch", "-r", "-m", "2", "-k", "21", "-s", "500", "-o", "trace", "reads1.fastq"], cwd=tmp_path, capture_output=True, text=True, env=envt)
    assert "reads parsed on the GPU" in t.stderr
    t = subprocess.run([MASH, "sketch", "-r", "-m", "2", "-k", "21", "-s", "500", "-o", "trace", "crlf.fastq"], cwd=tmp_path, capture_output=True, text=True, env=envt)
    assert "reads parsed on the GPU" in t.stderr                       # CRLF FASTQ too
    (tmp_path / "wrapped.fastq").write_bytes(b"@r1\nACGTACGTACGTACGTACGTACGT\nACGTACGT\n+\nIIIIIIIIIIIIIIIIIIIIIIII\nIIIIIIII\n" * 4)
    t = subprocess.run([MASH, "sketch", "-r", "-k", "21", "-s", "500", "-o", "trace", "wrapped.fastq"], cwd=tmp_path, capture_output=True, text=True, env=envt)
    assert t.returncode == 0 and "reads parsed on the GPU" not in t.stderr      # wrapped lines: the host reader
    t = subprocess.run([MASH, "sketch", "-k", "21", "-s", "500", "-o", "trace2", "reads.fa"], cwd=tmp_path, capture_output=True, text=True, env=envt)
    assert "fasta parsed on the GPU" in t.stderr


def test_fasta_read_set_pieces_equal_streamed_reads(ctx):
    """FASTA reads parsed on the GPU piece by piece and appended device to device = the same reads streamed from the host."""
    rng = np.random.default_rng(21)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    reads = [lut[rng.integers(0, 4, size=int(rng.integers(5, 400)))].tobytes() for _ in range(3000)]
    reads = reads + reads[:1500]                                       # repeated reads: counts >= 2
    text = [b">r%d c%d\n" % (i, i) + b"\n".join(r[j:j + 70] for j in range(0, len(r), 70)) + b"\n" for i, r in enumerate(reads)]
    kw = dict(k=21, s=300, min_cov=2, want_counts=True)
    want = ctx.sketch_stream([reads], **kw)[0]
    for cuts in ([0, len(text)], [0, 1, 2, 900, 901, 4499, len(text)]):
        got = ctx.sketch_fasta_pieces([b"".join(text[a:b]) for a, b in zip(cuts[:-1], cuts[1:])], **kw)
        assert got is not None
        assert np.array_equal(got["hashes"], want["hashes"]) and np.array_equal(got["counts"], want["counts"])
    assert ctx.sketch_fasta_pieces([b">r\nACGT\n+\nIIII\n"], **kw) is None


def test_cli_read_mode_fasta_reads_gpu_route_equals_host_reader_route(tmp_path):
    """`mash sketch -r -m 2` on FASTA reads (wrapped lines, comments, short first reads, CRLF, no final newline, a lone '>' at
    the end): the GPU route in small pieces and the host reader write the same .msh; the timing trace names the route."""
    rng = np.random.default_rng(22)
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    reads = [lut[rng.integers(0, 5 if i % 50 == 0 else 4, size=int(rng.integers(3, 500)))].tobytes() for i in range(4000)]
    def fasta(rs, eol=b"\n", width=60):
        return b"".join(b">read%d some comment %d" % (i, i) * (i % 3 != 0) + b">r%d" % i * (i % 3 == 0) + eol +
                        eol.join(r[j:j + width] for j in range(0, len(r), width)) + eol for i, r in enumerate(rs))
    base = fasta(reads) * 2
    files = {
        "reads.fa": base,
        "short_first.fa": b">tiny first comment\nACGT\n>tiny2\nAC\n" + base,
        "no_final_newline.fa": base.rstrip(b"\n"),
        "lone_gt.fa": base + b">",
        "crlf.fa": fasta(reads[:800], eol=b"\r\n") * 2,
        "junk_first.fa": b"junk before the first record\n" + base,
        "all_short.fa": b">a c\nACGT\n>b\nACG\n",
    }
    for nm, data in files.items():
        (tmp_path / nm).write_bytes(data)
    env0 = dict(os.environ, FPMASH_GPU_PARSE="0")
    envp = dict(os.environ, FPMASH_FASTQ_PIECE="65536")
    for nm in files:
        opts = ["sketch", "-r", "-m", "2", "-k", "21", "-s", "400"]
        a = subprocess.run([MASH] + opts + ["-o", "gpu_" + nm, nm], cwd=tmp_path, capture_output=True, text=True)
        c = subprocess.run([MASH] + opts + ["-o", "pcs_" + nm, nm], cwd=tmp_path, capture_output=True, text=True, env=envp)
        b = subprocess.run([MASH] + opts + ["-o", "host_" + nm, nm], cwd=tmp_path, capture_output=True, text=True, env=env0)
        assert a.returncode == b.returncode == c.returncode, (nm, a.stderr, b.stderr, c.stderr)
        assert a.stderr.replace("gpu_", "") == b.stderr.replace("host_", "") == c.stderr.replace("pcs_", ""), nm
        if a.returncode == 0:
            want = (tmp_path / ("host_%s.msh" % nm)).read_bytes()
            assert (tmp_path / ("gpu_%s.msh" % nm)).read_bytes() == want, nm
            assert (tmp_path / ("pcs_%s.msh" % nm)).read_bytes() == want, nm
    envt = dict(os.environ, FPMASH_TIMING="1")
    t = subprocess.run([MASH, "sketch", "-r", "-k", "21", "-s", "400", "-o", "trace", "reads.fa"], cwd=tmp_path, capture_output=True, text=True, env=envt)
    assert "reads parsed on the GPU" in t.stderr
