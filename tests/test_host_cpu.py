"""CPU: host logic of the product -- .msh codec vs the reference's fixtures (byte level), the
FASTA/FASTQ reader vs kseq.h, the C-ABI surface, error behaviour without a GPU, scalar dist math,
and the multi-GPU sharding helpers under gloo with world_size 2.  No kernel runs here."""
import ctypes
import gzip
import os
import re
import shutil
import subprocess
import sys

import numpy as np
import pytest

import mshpy
from conftest import GOLDEN, ROOT

MASH = os.path.join(ROOT, "fp-mash_b200", "bin", "mash")


@pytest.fixture(scope="session", autouse=True)
def built():
    if not os.path.exists(MASH) or not os.path.exists(os.path.join(ROOT, "fp-mash_b200", "lib", "libfpmash_b200.so")):
        subprocess.check_call(["make", "-s", "-j", "8", "-C", os.path.join(ROOT, "fp-mash_b200"), "all"])


def run(args, **kw):
    return subprocess.run([MASH] + args, capture_output=True, text=True, **kw)


# ---- .msh container ---------------------------------------------------------------------------
ROUNDTRIP = ["test_sequence.msh", "reads.msh", "genome1.fna.msh", "genome2.fna.msh", "genome3.fna.msh", "read1_2.msh", "example1.msh"]


@pytest.mark.parametrize("name", ROUNDTRIP)
def test_msh_decode_reencode_is_byte_identical(tmp_path, name):
    """decode -> re-encode through the product's codec (mash paste) reproduces the reference's
    files byte for byte: single- and multi-segment layouts, far pointers, counts lists."""
    out = tmp_path / "out"
    r = run(["paste", str(out), os.path.join(GOLDEN, name)])
    assert r.returncode == 0, r.stderr
    assert open(str(out) + ".msh", "rb").read() == open(os.path.join(GOLDEN, name), "rb").read()


def test_msh_paste_truncates_to_sketch_size_like_loadCapnp(tmp_path):
    # fp sketches hold 2000 hashes but advertise s=1000: loadCapnp keeps the FIRST 1000 (Sketch.cpp:1133-1145)
    out = tmp_path / "fp"
    assert run(["paste", str(out), os.path.join(GOLDEN, "DNA1-sketch.msh")]).returncode == 0
    a, b = mshpy.load(str(out) + ".msh"), mshpy.load(os.path.join(GOLDEN, "DNA1-sketch.msh"))
    assert len(a.refs) == 5
    for x, y in zip(a.refs, b.refs):
        assert x["hashes32"] == y["hashes32"][:1000] and x["name"] == y["name"] and x["length"] == y["length"]


def test_msh_paste_concatenates(tmp_path):
    out = tmp_path / "all"
    files = [os.path.join(GOLDEN, "genome%d.fna.msh" % i) for i in (1, 2, 3)]
    assert run(["paste", str(out)] + files).returncode == 0
    m = mshpy.load(str(out) + ".msh")
    assert [r["name"] for r in m.refs] == ["data/genome1.fna", "data/genome2.fna", "data/genome3.fna"]
    for r, f in zip(m.refs, files):
        assert r["hashes64"] == mshpy.load(f).refs[0]["hashes64"]
    assert run(["paste", str(out)] + files).returncode == 1          # refuses to overwrite (CommandPaste)


def test_info_dump_matches_fixture_content():
    r = run(["info", "-d", os.path.join(GOLDEN, "reads.msh")])
    assert r.returncode == 0
    nums = [int(x) for x in re.findall(r"^\s+(\d+),?$", r.stdout, re.M)]
    m = mshpy.load(os.path.join(GOLDEN, "reads.msh"))
    assert nums == m.refs[0]["hashes"] + m.refs[0]["counts32"]
    assert '"length" : 502359' in r.stdout and '"hashBits" : 64' in r.stdout
    t = run(["info", "-t", os.path.join(GOLDEN, "genome1.fna.msh")])
    assert t.stdout.splitlines()[1].split("\t")[:3] == ["1000", "4639675", "data/genome1.fna"]
    h = run(["info", "-H", os.path.join(GOLDEN, "DNA1-sketch.msh")])
    assert "K-mer size:                    1 (32-bit hashes)" in h.stdout and "Sketches:                      5" in h.stdout


def test_many_reference_layout(tmp_path):
    """> 113 references: the reference list itself leaves segment 0 and every text gets a landing
    pad (unpinned by fixtures, checked here for self-consistency and for the
    predicted segment sizes)."""
    files = [os.path.join(GOLDEN, "genome%d.fna.msh" % (1 + i % 3)) for i in range(150)]
    lst = tmp_path / "list.txt"
    lst.write_text("\n".join(files) + "\n")
    out = tmp_path / "big"
    assert run(["paste", "-l", str(out), str(lst)]).returncode == 0
    m = mshpy.load(str(out) + ".msh")
    assert len(m.refs) == 150 and m.segment_words[1] == 1 + 1 + 9 * 150
    for i, r in enumerate(m.refs):
        assert r["hashes64"] == mshpy.load(files[i]).refs[0]["hashes64"]
    back = tmp_path / "back"
    assert run(["paste", str(back), str(out) + ".msh"]).returncode == 0
    assert open(str(back) + ".msh", "rb").read() == open(str(out) + ".msh", "rb").read()


# ---- FASTA/FASTQ reader vs kseq.h --------------------------------------------------------------
def fnv(b):
    h = 1469598103934665603
    for c in b:
        h = ((h ^ c) * 1099511628211) & 0xffffffffffffffff
    return h


NASTY = {
    "crlf.fa": b">s1 first comment\r\nACGT\r\nacgt\r\n>s2\r\nGGGG\r\n",
    "stale_comment.fa": b">short has_comment here\nAC\n>long\nACGTACGTACGTACGTACGTACGTACGT\n>third\tx y\nTTTT",
    "inner_markers.fa": b">a\nACGT>b desc\nGG+TT\nIIII\n@c\nAA@d\nCC\n",
    "blank_lines.fa": b"\n\n>a\n\nAC GT\n\n\n>b\n>c\nNNNN\n",
    "fastq.fq": b"@r1 c1\nACGT\n+\nIIII\n@r2\nGGCC\n+r2\n@@@@\n@r3\nAAAA\nTTTT\n+\nIIIIIIII\n",
    "fastq_trunc.fq": b"@r1\nACGTACGT\n+\nIII",
    "no_newline.fa": b">x\nACGTAC",
    "empty.fa": b"",
    "garbage_head.fa": b"junk\nmore junk\n>ok 1\nACGT\n",
    "long_lines.fa": b">a x\n" + b"ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTAC\n" * 40
                     + b">b\n" + b"ACGTNNNNACGT ACGTACGTACGTACGTAC\tGTACGTACGTACGTACG\r\nTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT>c inner\nAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA+\n",
    "fastq_tricky_quality.fq": b"@r1\nACGTACGTACGTACGTACGTACGTACGTACGTAC\n+\n@@>>++IIIIIIIIIIIIIIIIIIIIIIIIIIII\n"
                               b"@r2 c\r\nACGTACGTACGTACGTACGTAC\r\n+\r\nIIIIIIIIIIIIIIIIIIIIII\r\n"
                               b"@r3\nACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT\n+\nIIIIIIIIII IIIIIIIIII\nIIIIIIIIIIIIIIIIIIII\n@r4\nAC\n+\nII",
    "fastq_multiline.fq": b"@m1\nACGTACGTAC\nGGGGGTTTTT\n+m1\nIIIIIIIIII\nJJJJJJJJJJ\n@m2\nAAAA\n+\nIIII\n",
}


@pytest.mark.parametrize("name", sorted(NASTY) + ["reads1.fastq.gz", "test_sequence.fasta", "DNA1.fasta"])
def test_fastx_reader_matches_kseq(tmp_path, reflib, name):
    if name in NASTY:
        path = str(tmp_path / name)
        with open(path, "wb") as f:
            f.write(NASTY[name])
    else:
        path = os.path.join(GOLDEN, name)
    want, code = reflib.parse_file(path)
    r = subprocess.run([MASH, "debug-parse", path], capture_output=True)
    assert r.returncode == 0
    lines = r.stdout.decode("latin1").split("\n")
    lines = [l for l in lines if l != ""]
    assert lines[-1] == "END\t%d" % code
    got = [l.split("\t") for l in lines[:-1]]
    # kseq crashes on an empty FIRST record; everything else must agree record by record
    assert len(got) == len(want)
    for g, (nm, cm, seq, cstr) in zip(got, want):
        # a tab inside a comment shifts the columns: compare from both ends
        assert g[0] == nm and int(g[-2]) == len(seq) and int(g[-1]) == fnv(seq)
        assert "\t".join(g[1:-2]) == cm + "\t" + cstr


# ---- C ABI ------------------------------------------------------------------------------------
def test_c_abi_exports_every_declared_symbol(fpm):
    hdr = open(os.path.join(ROOT, "include", "fpmash_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(fpm_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(fpm.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), "declared in include/fpmash_b200.h but not exported: " + sym
    assert declared == set(fpm.EXPORTED)
    assert fpm.lib.fpm_abi_version() == 2


def test_no_gpu_fails_loudly(fpm):
    if fpm.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(fpm.FpmError) as e:
        fpm.Context(0)
    assert e.value.code == fpm.FPM_ERR_NO_DEVICE and "no CPU fallback" in str(e.value)
    r = run(["sketch", "-o", "/tmp/should_not_exist_fpm", os.path.join(GOLDEN, "test_sequence.fasta")])
    assert r.returncode == 1 and "ERROR" in r.stderr and not os.path.exists("/tmp/should_not_exist_fpm.msh")


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "fp-mash_b200")
    for dp, _, fs in os.walk(pkg):
        if os.sep + "build" in dp:
            continue
        for f in fs:
            if f.endswith((".cu", ".cuh", ".h", ".cpp", ".py")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle" not in txt.lower(), os.path.join(dp, f)


def test_option_surface_and_errors():
    assert run(["sketch", "-k", "40", "x.fa"]).returncode == 1
    r = run(["sketch", "-k", "40", "x.fa"])
    assert "must be an integer between 1 and 32 (40 given)" in r.stderr
    assert "Unrecognized option: -Q" in run(["dist", "-Q", "a", "b"]).stderr
    assert "-k requires an argument" in run(["sketch", "x.fa", "-k"]).stderr
    r = run(["sketch", "-i", "-r", "x.fa"])
    assert r.returncode == 1 and "cannot be used with" in r.stderr
    r = run(["dist", "-k", "16", os.path.join(GOLDEN, "genome1.fna.msh"), os.path.join(GOLDEN, "genome2.fna.msh")])
    assert r.returncode == 1 and "cannot be used when a sketch is provided" in r.stderr
    assert "Unrecognized unit" in run(["sketch", "-g", "abc", "x.fa"]).stderr
    assert "must be a whole number" in run(["sketch", "-g", "1.5k", "x.fa"]).stderr
    for cmd in ("sketch", "dist", "paste", "info"):
        assert run([cmd, "-h"]).returncode == 0


# ---- scalar dist math (the same header the kernels compile) ------------------------------------
def test_pvalue_and_distance_match_oracle(fpm, oracle):
    rng = np.random.default_rng(5)
    worst = 0.0
    for ks in (4.0 ** 21, 4.0 ** 32, 10.0, 4.0 ** 16, 4.0 ** 11):
        for _ in range(300):
            n = int(rng.choice([1000, 10000, 500, 37]))
            x = int(rng.integers(1, n + 1))
            lr = int(rng.choice([73, 2000, 10596, 502359, 4639675, 5000000, 3 * 10 ** 9]))
            lq = int(rng.choice([73, 2000, 10596, 502359, 4639675, 5000000]))
            a, b = fpm.pvalue(x, lr, lq, ks, n), oracle.pvalue(x, lr, lq, ks, n)
            if b > 1e-290:
                worst = max(worst, abs(a - b) / b)
            else:
                assert a < 1e-289
    assert worst < 1e-12
    assert fpm.pvalue(0, 1, 1, 10.0, 5) == 1.0
    for c, d in [(0, 10), (10, 10), (0, 0), (456, 1000), (1, 1000), (999, 1000)]:
        want = 0.0 if c == d else 1.0 if c == 0 else min(1.0, -np.log(2 * (c / d) / (1 + c / d)) / 21)
        assert fpm.distance(c, d, 21) == pytest.approx(want, rel=1e-15, abs=0)


# ---- multi-GPU plumbing on CPU: the grid rules of the library, and the exchange step under gloo -------------------------
def test_grid_and_shard_rules_agree_with_the_library(fpm):
    """fpm_shard_range / fpm_dist_grid_shape / fpm_dist_block (C, no GPU needed) == the host helpers; the blocks of a grid
    tile the pair space exactly once; every row shard reaches exactly the ranks whose block contains it."""
    from fpmash_b200 import sharding as sh
    parts = sh.assign_by_size([5, 1, 9, 3, 3, 7], 2)
    assert sorted(sum(parts, [])) == list(range(6)) and all(p == sorted(p) for p in parts)
    loads = [sum([5, 1, 9, 3, 3, 7][i] for i in p) for p in parts]
    assert abs(loads[0] - loads[1]) <= 2
    for world in (1, 2, 3, 4, 6, 8):
        for n_q, n_r in ((20000, 20000), (100000, 10000), (7, 5), (1, 64), (0, 9), (1000, 3)):
            assert fpm.dist_grid_shape(world, n_q, n_r) == sh.grid_shape(world, n_q, n_r)
            qp, rp = sh.grid_shape(world, n_q, n_r)
            assert qp * rp == world
            cover = np.zeros((n_q, n_r), dtype=np.int32) if n_q * n_r <= 10**6 else None
            for rank in range(world):
                assert fpm.shard_range(n_q, rank, world) == sh.shard_range(n_q, rank, world)
                blk = sh.block_of(rank, world, n_q, n_r)
                assert fpm.dist_block(rank, world, n_q, n_r) == blk
                if cover is not None:
                    cover[blk[0]:blk[1], blk[2]:blk[3]] += 1
                ex = sh.senders_and_receivers(rank, world, n_q, n_r)
                # my block's rows are exactly the shards I receive
                assert blk[0] == sh.shard_range(n_q, ex["q_recv"][0], world)[0] and blk[1] == sh.shard_range(n_q, ex["q_recv"][-1], world)[1]
                assert blk[2] == sh.shard_range(n_r, ex["r_recv"][0], world)[0] and blk[3] == sh.shard_range(n_r, ex["r_recv"][-1], world)[1]
                # and every rank I send to expects me
                for d in ex["q_send"]:
                    assert rank in sh.senders_and_receivers(d, world, n_q, n_r)["q_recv"]
                for d in ex["r_send"]:
                    assert rank in sh.senders_and_receivers(d, world, n_q, n_r)["r_recv"]
            if cover is not None:
                assert (cover == 1).all()
    # 8 GPUs, all-vs-all: 2 x 4 -- a GPU indexes 3/4 of a panel's rows instead of 9/8 with row sharding
    assert sh.grid_shape(8, 20000, 20000) in ((2, 4), (4, 2))


WORKER = r"""
import os, sys
sys.path.insert(0, os.path.join(%(root)r, "fp-mash_b200", "py"))
sys.path.insert(0, os.path.join(%(root)r, "oracle"))
import numpy as np, torch, torch.distributed as dist
from fpmash_b200 import sharding as sh
from oracle_py import Oracle
world = %(world)d
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=world)
rank = dist.get_rank()
orc = Oracle()
rng = np.random.default_rng(5)
n_q, n_r, s = 11, 9, 40
pool = rng.choice(1 << 20, size=400, replace=False).astype(np.uint64)
def panel(n):
    return np.stack([np.sort(rng.choice(pool, size=s, replace=False)) for _ in range(n)])
Q, R = panel(n_q), panel(n_r)                       # same seed on every rank: each rank then keeps only ITS row shard
q0, q1 = sh.shard_range(n_q, rank, world); r0, r1 = sh.shard_range(n_r, rank, world)
my_q, my_r = torch.from_numpy(Q[q0:q1].astype(np.int64)), torch.from_numpy(R[r0:r1].astype(np.int64))
# the exchange step exactly as fpm_dist_sharded_dev performs it (grouped send/recv; gloo instead of NCCL)
ex = sh.senders_and_receivers(rank, world, n_q, n_r)
blk = sh.block_of(rank, world, n_q, n_r)
bq = torch.zeros((blk[1] - blk[0], s), dtype=torch.int64); br = torch.zeros((blk[3] - blk[2], s), dtype=torch.int64)
ops = []
for role, mine, sends, recvs, out, base, n in (("q", my_q, ex["q_send"], ex["q_recv"], bq, blk[0], n_q), ("r", my_r, ex["r_send"], ex["r_recv"], br, blk[2], n_r)):
    for d in sends:
        if d != rank and mine.shape[0]:
            ops.append(dist.P2POp(dist.isend, mine.contiguous(), d))
    for src in recvs:
        a, b = sh.shard_range(n, src, world)
        if a == b:
            continue
        if src == rank:
            out[a - base:b - base] = mine
        else:
            ops.append(dist.P2POp(dist.irecv, out[a - base:b - base], src))
if ops:
    for w in dist.batch_isend_irecv(ops):
        w.wait()
assert np.array_equal(bq.numpy().astype(np.uint64), Q[blk[0]:blk[1]]) and np.array_equal(br.numpy().astype(np.uint64), R[blk[2]:blk[3]])
# each rank compares its block; gathered blocks == the reference's query-major table
mine = np.zeros((blk[1] - blk[0], blk[3] - blk[2]), dtype=np.int64)
for i in range(mine.shape[0]):
    for j in range(mine.shape[1]):
        c = orc.compare(br.numpy().astype(np.uint64)[j], bq.numpy().astype(np.uint64)[i], 1000, 1000, s, 21, 4.0 ** 21)
        mine[i, j] = c["numer"] * 1000 + c["denom"]
gathered = [None] * world
dist.all_gather_object(gathered, (blk, mine))
full = sh.assemble_blocks(gathered, n_q, n_r, np.int64)
for q in range(n_q):
    for r in range(n_r):
        c = orc.compare(R[r], Q[q], 1000, 1000, s, 21, 4.0 ** 21)
        assert full[q, r] == c["numer"] * 1000 + c["denom"], (q, r)
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


@pytest.mark.parametrize("world", [2, 4])
def test_block_exchange_gloo(tmp_path, world):
    """world_size 2 and 4 on CPU (gloo): row shards -> blocks by the library's send/receive lists, blocks compared by the
    oracle, gathered blocks == the single-process query-major table."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "w.py"
    script.write_text(WORKER % {"root": ROOT, "port": port, "world": world})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(world)]
    for p in procs:
        out, err = p.communicate(timeout=240)
        assert p.returncode == 0 and "ok" in out, err[-2000:]


def test_record_layouts_match_the_header(fpm, tmp_path):
    """include/fpmash_b200.h is plain C; the numpy record types of the binding must have the C compiler's layout."""
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\n'
                   'int main(void){printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu\\n", sizeof(fpm_pair), offsetof(fpm_pair, distance), offsetof(fpm_pair, pvalue),'
                   ' sizeof(fpm_hit), offsetof(fpm_hit, numer), offsetof(fpm_hit, distance), offsetof(fpm_hit, pvalue)); return 0;}\n'
                   % os.path.join(ROOT, "include", "fpmash_b200.h"))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-o", str(exe), str(src)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    p, h = fpm.PAIR_DTYPE, fpm.HIT_DTYPE
    assert got == [p.itemsize, p.fields["distance"][1], p.fields["pvalue"][1], h.itemsize, h.fields["numer"][1], h.fields["distance"][1], h.fields["pvalue"][1]]
    assert got[0] == 24 and got[3] == 32


def test_gunzip_core_on_the_cpu_equals_zlib(tmp_path):
    """The sequential half of the GPU inflate (csrc/gunzip_core.cuh: bit reader, Huffman codes, block and member headers,
    CRC arithmetic) is plain C++; tests/native/gunzip_sim.cpp runs it on the CPU the way lane 0 of the kernel does."""
    import gz_cases
    exe = str(tmp_path / "gunzip_sim")
    subprocess.run(["g++", "-O2", "-std=c++17", "-x", "c++", os.path.join(ROOT, "tests", "native", "gunzip_sim.cpp"), "-o", exe], check=True)
    for name, (gz, want) in gz_cases.good_cases(big=True).items():
        (tmp_path / "in.gz").write_bytes(gz)
        r = subprocess.run([exe, str(tmp_path / "in.gz"), str(tmp_path / "out")], timeout=120)
        assert r.returncode == 0, name
        assert (tmp_path / "out").read_bytes() == want, name
    for name, gz in gz_cases.bad_cases().items():
        (tmp_path / "in.gz").write_bytes(gz)
        r = subprocess.run([exe, str(tmp_path / "in.gz"), str(tmp_path / "out")], timeout=120)
        assert r.returncode == 2, name


def test_gunzip_core_random_streams(tmp_path):
    """Randomised: inputs from incompressible to highly repetitive, every zlib level / strategy / window size / memLevel
    (memLevel changes how often a block ends), through the CPU run of the decoder core."""
    import random
    import zlib
    exe = str(tmp_path / "gunzip_sim")
    subprocess.run(["g++", "-O2", "-std=c++17", "-x", "c++", os.path.join(ROOT, "tests", "native", "gunzip_sim.cpp"), "-o", exe], check=True)
    rnd = random.Random(20260)
    strategies = [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED]
    for case in range(120):
        n = rnd.choice([0, 1, 2, 31, 257, 4096, 65535, 65536, 70000, 300000])
        kind = case % 4
        if kind == 0:
            data = bytes(rnd.getrandbits(8) for _ in range(n))
        elif kind == 1:
            data = "".join(rnd.choice("ACGT") for _ in range(n)).encode()
        elif kind == 2:
            unit = bytes(rnd.getrandbits(8) for _ in range(rnd.choice([1, 2, 3, 7, 100, 5000])))
            data = (unit * (n // len(unit) + 1))[:n]
        else:
            words = [bytes(rnd.getrandbits(8) for _ in range(rnd.randint(1, 12))) for _ in range(50)]
            data = b"".join(rnd.choice(words) for _ in range(n // 6 + 1))[:n]
        co = zlib.compressobj(rnd.randint(0, 9), zlib.DEFLATED, 16 + rnd.randint(9, 15), rnd.randint(1, 9), rnd.choice(strategies))
        gz = co.compress(data) + co.flush()
        (tmp_path / "in.gz").write_bytes(gz)
        r = subprocess.run([exe, str(tmp_path / "in.gz"), str(tmp_path / "out")], timeout=120)
        assert r.returncode == 0, case
        assert (tmp_path / "out").read_bytes() == data, case
        # one flipped bit somewhere: refused (CRC) or, if the flip hit a don't-care header byte, still the same data
        if len(gz) > 30 and case % 3 == 0:
            g = bytearray(gz)
            g[rnd.randrange(10, len(g))] ^= 1 << rnd.randrange(8)
            (tmp_path / "in.gz").write_bytes(bytes(g))
            r = subprocess.run([exe, str(tmp_path / "in.gz"), str(tmp_path / "out")], timeout=120)
            assert r.returncode in (0, 2), case
            if r.returncode == 0:
                assert (tmp_path / "out").read_bytes() == data, case
