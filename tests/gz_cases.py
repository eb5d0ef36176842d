"""gzip streams for the inflate tests (CPU simulation of csrc/gunzip_core.cuh in test_host_cpu.py, the kernel in test_gpu_gunzip.py):
every DEFLATE block type, header field and multi-member layout zlib's gzread accepts, and the damaged streams it refuses."""
import gzip
import io
import random
import zlib


def fasta_text(n, seed, width=70, name=None):
    r = random.Random(seed)
    core = "".join(r.choice("ACGT") for _ in range(n))
    lines = [">%s some comment %d" % (name or "seq%d" % seed, seed)] + [core[i:i + width] for i in range(0, n, width)]
    return ("\n".join(lines) + "\n").encode()


def good_cases(big=False):
    """name -> (gzip bytes, expected output)"""
    rnd = random.Random(1)
    d = fasta_text(300000, 3)
    cases = {}
    for lvl in (1, 6, 9):
        cases["dna_l%d" % lvl] = (gzip.compress(d, lvl), d)
    cases["empty"] = (gzip.compress(b""), b"")
    cases["one_byte"] = (gzip.compress(b"A"), b"A")
    cases["stored"] = (gzip.compress(d[:200000], 0), d[:200000])
    co = zlib.compressobj(9, zlib.DEFLATED, 31, 9, zlib.Z_FIXED)
    cases["fixed"] = (co.compress(d[:50000]) + co.flush(), d[:50000])
    co = zlib.compressobj(6, zlib.DEFLATED, 31, 1, zlib.Z_HUFFMAN_ONLY)
    cases["huffman_only"] = (co.compress(d[:50000]) + co.flush(), d[:50000])
    x = b"A" * 100000 + d[:1000] + b"\n" * 5000
    co = zlib.compressobj(6, zlib.DEFLATED, 31, 9, zlib.Z_RLE)
    cases["rle_overlapping_copies"] = (co.compress(x) + co.flush(), x)
    rb = bytes(rnd.getrandbits(8) | 1 for _ in range(100000))             # no 0x00: the batch separator
    cases["random_bytes"] = (gzip.compress(rb, 6), rb)
    txt = (b"the quick brown fox jumps over the lazy dog; " * 3000)[:120001]
    cases["text_long_codes"] = (gzip.compress(txt + rb[:20000] + txt, 9), txt + rb[:20000] + txt)
    cases["multi_member"] = (gzip.compress(d[:1000]) + gzip.compress(d[1000:5000], 1) + gzip.compress(b"") + gzip.compress(d[5000:]), d)
    bio = io.BytesIO()
    with gzip.GzipFile(filename="a_file_name.fa", mode="wb", fileobj=bio, mtime=12345) as g:
        g.write(d[:10000])
    cases["fname"] = (bio.getvalue(), d[:10000])
    raw = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = raw.compress(d[:20000]) + raw.flush()
    tail = zlib.crc32(d[:20000]).to_bytes(4, "little") + (20000).to_bytes(4, "little")
    hdr = bytes([0x1f, 0x8b, 8, 4 | 8 | 16 | 2, 0, 0, 0, 0, 0, 3]) + (5).to_bytes(2, "little") + b"EXTRA" + b"name\0" + b"comment text\0"
    hdr += (zlib.crc32(hdr) & 0xffff).to_bytes(2, "little")
    cases["all_header_fields"] = (hdr + body + tail, d[:20000])
    cases["name_longer_than_the_ring"] = (bytes([0x1f, 0x8b, 8, 8, 0, 0, 0, 0, 0, 3]) + b"n" * 5000 + b"\0" + body + tail, d[:20000])
    cases["trailing_zeros"] = (gzip.compress(d[:3000]) + b"\0" * 100, d[:3000])
    cases["trailing_garbage"] = (gzip.compress(d[:3000]) + b"garbage garbage garbage garbage", d[:3000])
    co = zlib.compressobj(6, zlib.DEFLATED, 31)
    parts = []
    for i in range(0, 100000, 777):
        parts.append(co.compress(d[i:i + 777]))
        parts.append(co.flush(zlib.Z_FULL_FLUSH if i % 2 else zlib.Z_SYNC_FLUSH))
    parts.append(co.flush())
    gz = b"".join(parts)
    cases["many_flushes"] = (gz, gzip.decompress(gz))
    if big:
        b5 = fasta_text(5_000_000, 9)
        cases["genome_5mb"] = (gzip.compress(b5, 6), b5)
    return cases


def bad_cases():
    """name -> gzip bytes zlib refuses"""
    d = fasta_text(300000, 3)
    bad = {}
    g = bytearray(gzip.compress(d, 6)); g[len(g) // 2] ^= 0x10
    bad["flipped_bit"] = bytes(g)
    g = gzip.compress(d, 6)
    bad["truncated"] = g[:len(g) // 2]
    bad["truncated_trailer"] = g[:-3]
    g = bytearray(gzip.compress(d[:5000], 6)); g[-5] ^= 1
    bad["crc_mismatch"] = bytes(g)
    g = bytearray(gzip.compress(d[:5000], 6)); g[-1] ^= 1
    bad["isize_mismatch"] = bytes(g)
    bad["not_gzip"] = d[:5000]
    bad["second_member_bad_header"] = gzip.compress(d[:3000]) + bytes([0x1f, 0x8b, 8, 0xe0]) + b"x" * 40
    bad["reserved_block_type"] = bytes([0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 3, 0x07]) + b"\0" * 30
    return bad
