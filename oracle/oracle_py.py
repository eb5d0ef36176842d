"""ctypes front-end for the parity oracle -- TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, never by the product (fp-mash_b200/).

`Oracle` wraps oracle/liboracle.so (the CPU restatement, oracle/mash_oracle.cpp).
`RefLib` wraps oracle/_ref/libmashref.so (the reference's own translation units compiled
where they lie under /root/reference; see oracle/Makefile) when it has been built.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def build(ref=True):
    """Compile the oracle (and oracle/_ref when the reference tree is present)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"] + (["ref"] if ref else []))


def nucleotide_alphabet(chars="ACGT", preserve_case=False):
    """setAlphabetFromString (Sketch.cpp:1260-1289) -> 256-entry validity table."""
    a = np.zeros(256, dtype=np.uint8)
    for ch in chars:
        c = ord(ch)
        if not preserve_case and 96 < c < 123:
            c -= 32
        a[c] = 1
    return a


class OrcPair(C.Structure):
    _fields_ = [("numer", C.c_uint64), ("denom", C.c_uint64), ("distance", C.c_double),
                ("pvalue", C.c_double), ("pass_", C.c_int)]


class Oracle:
    def __init__(self, path=None):
        path = path or os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = self.L = C.CDLL(path)
        L.orc_get_hash.restype = C.c_uint64
        L.orc_get_hash.argtypes = [C.c_char_p, C.c_int, C.c_uint32, C.c_int]
        L.orc_murmur3_x64_128.argtypes = [C.c_char_p, C.c_int, C.c_uint32, u64p]
        L.orc_fp_hash.restype = C.c_uint64
        L.orc_fp_hash.argtypes = [u64p, C.c_int, C.c_uint32, C.c_int]
        L.orc_heap_new.restype = C.c_void_p
        L.orc_heap_new.argtypes = [C.c_int, C.c_uint64, C.c_uint64]
        L.orc_heap_free.argtypes = [C.c_void_p]
        L.orc_heap_offer.argtypes = [C.c_void_p, u64p, C.c_uint64]
        L.orc_heap_add_sequence.restype = C.c_uint64
        L.orc_heap_add_sequence.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_int, C.c_uint32,
                                            C.c_int, C.c_int, u8p, u64p, C.c_uint64]
        L.orc_heap_size.restype = C.c_uint64
        L.orc_heap_size.argtypes = [C.c_void_p]
        L.orc_heap_set_size.restype = C.c_double
        L.orc_heap_set_size.argtypes = [C.c_void_p]
        L.orc_heap_multiplicity.restype = C.c_double
        L.orc_heap_multiplicity.argtypes = [C.c_void_p]
        L.orc_heap_result.restype = C.c_uint64
        L.orc_heap_result.argtypes = [C.c_void_p, u64p, u32p]
        L.orc_pvalue.restype = C.c_double
        L.orc_pvalue.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_uint64]
        L.orc_binom_tail.restype = C.c_double
        L.orc_binom_tail.argtypes = [C.c_uint64, C.c_uint64, C.c_double]
        L.orc_compare_sketches.argtypes = [u64p, C.c_uint64, u64p, C.c_uint64, C.c_uint64, C.c_uint64,
                                           C.c_uint64, C.c_int, C.c_double, C.c_double, C.c_double,
                                           C.POINTER(OrcPair)]

    # -- hashing --------------------------------------------------------------------
    def get_hash(self, data: bytes, seed=42, use64=True):
        return self.L.orc_get_hash(data, len(data), seed, int(use64))

    def murmur128(self, data: bytes, seed=42):
        out = (C.c_uint64 * 2)()
        self.L.orc_murmur3_x64_128(data, len(data), seed, out)
        return out[0], out[1]

    def fp_hash(self, tokens, seed=42, use64=False):
        t = np.ascontiguousarray(tokens, dtype=np.uint64)
        return self.L.orc_fp_hash(_ptr(t, u64p), len(t), seed, int(use64))

    # -- sketching ------------------------------------------------------------------
    def sketch(self, records, k=21, s=1000, seed=42, noncanonical=False, preserve_case=False,
               alphabet=None, min_cov=1, use64=None, trace=False):
        """sketchFile's per-record loop (Sketch.cpp:1352-1422) over in-memory records:
        records shorter than k are skipped, the rest go through addMinHashes in order.
        Returns dict(hashes, counts, set_size, multiplicity, length, n_valid[, trace])."""
        if alphabet is None:
            alphabet = nucleotide_alphabet("ACGT", preserve_case)
        if use64 is None:
            use64 = float(int(alphabet.sum())) ** k > 2.0 ** 32   # Sketch.cpp:1288
        h = self.L.orc_heap_new(int(use64), s, min_cov)
        total_len = 0
        n_valid = 0
        tr = []
        try:
            for rec in records:
                if len(rec) < k:
                    continue
                n_valid += 1
                total_len += len(rec)
                buf = C.create_string_buffer(bytes(rec), len(rec))
                if trace:
                    t = np.empty(len(rec), dtype=np.uint64)
                    n = self.L.orc_heap_add_sequence(h, buf, len(rec), k, seed, int(noncanonical),
                                                     int(preserve_case), _ptr(alphabet, u8p),
                                                     _ptr(t, u64p), len(t))
                    tr.append(t[:n].copy())
                else:
                    self.L.orc_heap_add_sequence(h, buf, len(rec), k, seed, int(noncanonical),
                                                 int(preserve_case), _ptr(alphabet, u8p), None, 0)
            n = self.L.orc_heap_size(h)
            hashes = np.empty(n, dtype=np.uint64)
            counts = np.empty(n, dtype=np.uint32)
            self.L.orc_heap_result(h, _ptr(hashes, u64p), _ptr(counts, u32p))
            out = dict(hashes=hashes, counts=counts, set_size=self.L.orc_heap_set_size(h),
                       multiplicity=self.L.orc_heap_multiplicity(h), length=total_len,
                       n_valid=n_valid, use64=bool(use64))
            if trace:
                out["trace"] = np.concatenate(tr) if tr else np.empty(0, dtype=np.uint64)
            return out
        finally:
            self.L.orc_heap_free(h)

    def heap_stream(self, hashes, s, min_cov=1, use64=True):
        """Feed a raw hash stream through the literal MinHashHeap restatement."""
        hs = np.ascontiguousarray(hashes, dtype=np.uint64)
        h = self.L.orc_heap_new(int(use64), s, min_cov)
        try:
            self.L.orc_heap_offer(h, _ptr(hs, u64p), len(hs))
            n = self.L.orc_heap_size(h)
            out_h = np.empty(n, dtype=np.uint64)
            out_c = np.empty(n, dtype=np.uint32)
            self.L.orc_heap_result(h, _ptr(out_h, u64p), _ptr(out_c, u32p))
            return out_h, out_c, self.L.orc_heap_set_size(h)
        finally:
            self.L.orc_heap_free(h)

    # -- distance -------------------------------------------------------------------
    def pvalue(self, x, len_ref, len_qry, kmer_space, n):
        return self.L.orc_pvalue(x, len_ref, len_qry, kmer_space, n)

    def binom_tail(self, x, n, r):
        return self.L.orc_binom_tail(x, n, r)

    def compare(self, ref, qry, len_ref, len_qry, s, k, kmer_space, max_d=1.0, max_p=1.0):
        a = np.ascontiguousarray(ref, dtype=np.uint64)
        b = np.ascontiguousarray(qry, dtype=np.uint64)
        out = OrcPair()
        self.L.orc_compare_sketches(_ptr(a, u64p), len(a), _ptr(b, u64p), len(b), len_ref, len_qry,
                                    s, k, kmer_space, max_d, max_p, C.byref(out))
        return dict(numer=out.numer, denom=out.denom, distance=out.distance, pvalue=out.pvalue,
                    passed=bool(out.pass_))


REF_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_char_p, C.POINTER(C.c_char), C.c_uint64, C.POINTER(C.c_char), C.c_int)


class RefLib:
    """The reference's own hash.cpp / MinHashHeap.cpp / kseq.h behind oracle/ref_harness.cpp."""

    PATH = os.path.join(HERE, "_ref", "libmashref.so")

    @classmethod
    def available(cls):
        return os.path.exists(cls.PATH)

    def __init__(self):
        L = self.L = C.CDLL(self.PATH)
        L.ref_get_hash.restype = C.c_uint64
        L.ref_get_hash.argtypes = [C.c_char_p, C.c_int, C.c_uint32, C.c_int]
        L.ref_fp_hash.restype = C.c_uint64
        L.ref_fp_hash.argtypes = [u64p, C.c_int, C.c_uint32, C.c_int]
        L.ref_heap_new.restype = C.c_void_p
        L.ref_heap_new.argtypes = [C.c_int, C.c_uint64, C.c_uint64]
        L.ref_heap_free.argtypes = [C.c_void_p]
        L.ref_heap_offer.argtypes = [C.c_void_p, u64p, C.c_uint64, C.c_int]
        L.ref_heap_add_sequence.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_int, C.c_uint32, C.c_int,
                                            C.c_int, C.c_int, u8p]
        L.ref_heap_set_size.restype = C.c_double
        L.ref_heap_set_size.argtypes = [C.c_void_p]
        L.ref_heap_multiplicity.restype = C.c_double
        L.ref_heap_multiplicity.argtypes = [C.c_void_p]
        L.ref_heap_result.restype = C.c_uint64
        L.ref_heap_result.argtypes = [C.c_void_p, C.c_int, u64p, u32p, C.c_uint64]
        L.ref_parse_file.restype = C.c_int
        L.ref_parse_file.argtypes = [C.c_char_p, REF_CB, C.c_void_p]
        L.ref_sketch_batch.argtypes = [C.c_void_p, u64p, C.c_uint64, C.c_int, C.c_uint64, C.c_uint32, C.c_int,
                                       C.c_int, C.c_uint64, C.c_int, u64p, u32p, u64p]
        L.ref_dist_batch.argtypes = [u64p, u32p, C.c_uint64, u64p, u32p, C.c_uint64, C.c_uint64, C.c_int,
                                     C.c_int, u32p, u32p, f64p]

    def get_hash(self, data: bytes, seed=42, use64=True):
        return self.L.ref_get_hash(data, len(data), seed, int(use64))

    def fp_hash(self, tokens, seed=42, use64=False):
        t = np.ascontiguousarray(tokens, dtype=np.uint64)
        return self.L.ref_fp_hash(_ptr(t, u64p), len(t), seed, int(use64))

    def heap_stream(self, hashes, s, min_cov=1, use64=True):
        hs = np.ascontiguousarray(hashes, dtype=np.uint64)
        h = self.L.ref_heap_new(int(use64), s, min_cov)
        try:
            self.L.ref_heap_offer(h, _ptr(hs, u64p), len(hs), int(use64))
            out_h = np.empty(s + 1, dtype=np.uint64)
            out_c = np.empty(s + 1, dtype=np.uint32)
            n = self.L.ref_heap_result(h, int(use64), _ptr(out_h, u64p), _ptr(out_c, u32p), s + 1)
            return out_h[:n].copy(), out_c[:n].copy(), self.L.ref_heap_set_size(h)
        finally:
            self.L.ref_heap_free(h)

    def sketch(self, records, k=21, s=1000, seed=42, noncanonical=False, preserve_case=False,
               alphabet=None, min_cov=1, use64=None):
        if alphabet is None:
            alphabet = nucleotide_alphabet("ACGT", preserve_case)
        if use64 is None:
            use64 = float(int(alphabet.sum())) ** k > 2.0 ** 32
        h = self.L.ref_heap_new(int(use64), s, min_cov)
        try:
            total = 0
            for rec in records:
                if len(rec) < k:
                    continue
                total += len(rec)
                buf = C.create_string_buffer(bytes(rec), len(rec))
                self.L.ref_heap_add_sequence(h, buf, len(rec), k, seed, int(use64), int(noncanonical),
                                             int(preserve_case), _ptr(alphabet, u8p))
            out_h = np.empty(s + 1, dtype=np.uint64)
            out_c = np.empty(s + 1, dtype=np.uint32)
            n = self.L.ref_heap_result(h, int(use64), _ptr(out_h, u64p), _ptr(out_c, u32p), s + 1)
            return dict(hashes=out_h[:n].copy(), counts=out_c[:n].copy(),
                        set_size=self.L.ref_heap_set_size(h),
                        multiplicity=self.L.ref_heap_multiplicity(h), length=total)
        finally:
            self.L.ref_heap_free(h)

    def parse_file(self, path):
        recs = []

        def cb(_u, name, comment, clen, seq, l):
            # (name, comment by length, sequence, comment as the C string kseq exposes -- stale when absent)
            recs.append((name.decode("latin1"), C.string_at(comment, clen).decode("latin1"),
                         C.string_at(seq, l), C.string_at(comment).decode("latin1")))
        code = self.L.ref_parse_file(path.encode(), REF_CB(cb), None)
        return recs, code

    def sketch_batch(self, seqs: np.ndarray, offsets: np.ndarray, k, s, seed=42, use64=True,
                     noncanonical=False, min_cov=1, threads=1, want_counts=False):
        """CPU baseline: one genome per offsets interval, `threads` pool threads.
        `seqs` is modified in place (upper-casing) like the reference."""
        n = len(offsets) - 1
        out_h = np.zeros((n, s), dtype=np.uint64)
        out_c = np.zeros((n, s), dtype=np.uint32) if want_counts else None
        out_n = np.zeros(n, dtype=np.uint64)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        self.L.ref_sketch_batch(seqs.ctypes.data_as(C.c_void_p), _ptr(off, u64p), n, k, s, seed, int(use64),
                                int(noncanonical), min_cov, threads, _ptr(out_h, u64p), _ptr(out_c, u32p),
                                _ptr(out_n, u64p))
        return out_h, out_c, out_n

    def dist_batch(self, ref, ref_n, qry, qry_n, s, k, threads=1):
        ref = np.ascontiguousarray(ref, dtype=np.uint64)
        qry = np.ascontiguousarray(qry, dtype=np.uint64)
        rn = np.ascontiguousarray(ref_n, dtype=np.uint32)
        qn = np.ascontiguousarray(qry_n, dtype=np.uint32)
        npairs = ref.shape[0] * qry.shape[0]
        numer = np.empty(npairs, dtype=np.uint32)
        denom = np.empty(npairs, dtype=np.uint32)
        dist = np.empty(npairs, dtype=np.float64)
        self.L.ref_dist_batch(_ptr(ref, u64p), _ptr(rn, u32p), ref.shape[0], _ptr(qry, u64p), _ptr(qn, u32p),
                              qry.shape[0], s, k, threads, _ptr(numer, u32p), _ptr(denom, u32p), _ptr(dist, f64p))
        return numer, denom, dist
