"""Independent pure-Python reader of .msh files (unpacked Cap'n Proto, capnp/MinHash.capnp) used
by the tests to look inside fixtures and inside files written by the product, without going
through the product's own codec.  Layout: SURVEY.md section 5.1."""
import struct


class Msh:
    def __init__(self, data: bytes):
        nseg = struct.unpack_from("<I", data, 0)[0] + 1
        sizes = struct.unpack_from("<%dI" % nseg, data, 4)
        pos = (4 + 4 * nseg + 7) & ~7
        self.segs = []
        for w in sizes:
            self.segs.append(data[pos:pos + 8 * w])
            pos += 8 * w
        self.segment_words = list(sizes)
        root = self._resolve(0, 0)
        rs = self._struct(root)
        w0, w1, w2 = (self._data(rs, i) for i in range(3))
        self.kmer_size = w0 & 0xffffffff
        self.window_size = w0 >> 32
        self.sketch_size = w1 & 0xffffffff
        self.concatenated = bool((w1 >> 32) & 1)
        self.noncanonical = bool((w1 >> 33) & 1)
        self.preserve_case = bool((w1 >> 34) & 1)
        self.hash_seed = ((w2 >> 32) & 0xffffffff) ^ 42
        self.alphabet = self._text(self._ptr(rs, 2))
        self.refs = []
        self.used_old_list = False
        for which in (3, 0):
            rl = self._struct(self._ptr(rs, which))
            lst = self._ptr(rl, 0) if rl else None
            if lst and lst[0] == 1:
                seg, idx, hi = lst[1], lst[2], lst[3]
                assert hi & 7 == 7
                tag = self._word(seg, idx)
                n = (tag & 0xffffffff) >> 2
                nd, np_ = (tag >> 32) & 0xffff, tag >> 48
                if n:
                    self.used_old_list = which == 0
                    for i in range(n):
                        e = (seg, idx + 1 + i * (nd + np_), nd, np_)
                        d0, d1 = self._data(e, 0), self._data(e, 1)
                        ref = {"name": self._text(self._ptr(e, 2)), "comment": self._text(self._ptr(e, 3)),
                               "length": d1 if d1 else d0 & 0xffffffff, "counts_sorted": bool((d0 >> 32) & 1),
                               "hashes32": self._list(self._ptr(e, 4), 4), "hashes64": self._list(self._ptr(e, 5), 8),
                               "counts32": self._list(self._ptr(e, 6), 4)}
                        ref["hashes"] = ref["hashes64"] if ref["hashes64"] is not None else (ref["hashes32"] or [])
                        self.refs.append(ref)
                    break

    def _word(self, seg, idx):
        return struct.unpack_from("<Q", self.segs[seg], 8 * idx)[0]

    def _resolve(self, seg, idx):
        w = self._word(seg, idx)
        if w == 0:
            return None
        lo, hi = w & 0xffffffff, w >> 32
        kind = lo & 3
        if kind == 2:
            assert not (lo >> 2) & 1, "double-far pointers do not occur in mash files"
            return self._resolve(hi, lo >> 3)
        off = lo >> 2
        if off & (1 << 29):
            off -= 1 << 30
        return (kind, seg, idx + 1 + off, hi)

    def _struct(self, p):
        if not p:
            return None
        assert p[0] == 0
        return (p[1], p[2], p[3] & 0xffff, p[3] >> 16)

    def _data(self, s, i):
        return self._word(s[0], s[1] + i) if s and i < s[2] else 0

    def _ptr(self, s, i):
        return self._resolve(s[0], s[1] + s[2] + i) if s and i < s[3] else None

    def _text(self, p):
        if not p:
            return None
        n = p[3] >> 3
        return self.segs[p[1]][8 * p[2]:8 * p[2] + n - 1].decode("latin1")

    def _list(self, p, width):
        if not p:
            return None
        n = p[3] >> 3
        assert p[3] & 7 == (4 if width == 4 else 5)
        return list(struct.unpack_from("<%d%s" % (n, "I" if width == 4 else "Q"), self.segs[p[1]], 8 * p[2]))


def load(path):
    with open(path, "rb") as f:
        return Msh(f.read())
