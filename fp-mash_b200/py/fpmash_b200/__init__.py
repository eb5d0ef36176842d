"""fpmash_b200 -- ctypes binding of the C ABI in include/fpmash_b200.h.

This is plumbing for tests and bench.py: it loads fp-mash_b200/lib/libfpmash_b200.so (the
hand-written sm_100a kernels) and exposes the reference's hot-path operations with the
reference's own argument names.  There is NO CPU fallback: if the library is missing the
import fails, and if no CUDA device is present every compute call raises FpmError.

Reference interfaces mirrored (mash/src/mash/ of UmbertoDellaMonica/fp-mash):
    sketch_batch    <- sketchFile / sketchSequence / addMinHashes   (Sketch.cpp:1299-1517, 664-735)
    fp_hash_batch   <- getHashFingerPrint via initFromFingerprints   (hash.cpp:45-73, Sketch.cpp:132)
    dist_tile       <- compare / compareSketches / pValue            (CommandDistance.cpp:335-450)
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FPMASH_B200_LIB") or os.path.normpath(os.path.join(_HERE, "..", "..", "lib", "libfpmash_b200.so"))
HEADER_PATH = os.path.normpath(os.path.join(_HERE, "..", "..", "..", "include", "fpmash_b200.h"))

FPM_OK = 0
FPM_ERR_NO_DEVICE = -1
FPM_ERR_CUDA = -2
FPM_ERR_ARG = -3
FPM_ERR_UNSUPPORTED = -4
FPM_ERR_NOMEM = -5
FPM_ERR_COMM = -7
FPM_PAIR_PASS = 0x80000000

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)


class FpmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("fpmash_b200 error %d: %s" % (code, msg))
        self.code = code


class SketchParams(C.Structure):
    _fields_ = [("kmer_size", C.c_int32), ("sketch_size", C.c_uint32), ("seed", C.c_uint32),
                ("min_cov", C.c_uint32), ("noncanonical", C.c_uint8), ("preserve_case", C.c_uint8),
                ("use64", C.c_uint8), ("want_counts", C.c_uint8), ("alphabet", C.c_uint8 * 256)]


class Pair(C.Structure):
    _fields_ = [("numer", C.c_uint32), ("denom", C.c_uint32), ("distance", C.c_double), ("pvalue", C.c_double)]


PAIR_DTYPE = np.dtype([("numer", "<u4"), ("denom", "<u4"), ("distance", "<f8"), ("pvalue", "<f8")])


class DistParams(C.Structure):
    _fields_ = [("sketch_size", C.c_uint32), ("kmer_size", C.c_int32), ("kmer_space", C.c_double),
                ("max_distance", C.c_double), ("max_pvalue", C.c_double), ("sorted_unique", C.c_uint8)]


class Panel(C.Structure):
    _fields_ = [("hashes", C.c_void_p), ("sizes", C.c_void_p), ("lengths", C.c_void_p),
                ("n", C.c_uint64), ("stride", C.c_uint64)]


if not os.path.exists(LIB_PATH):
    raise ImportError("fpmash_b200: %s is missing -- build it with `make -C fp-mash_b200` "
                      "(or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)

lib = C.CDLL(LIB_PATH)

_VP = C.c_void_p
lib.fpm_abi_version.restype = C.c_int
lib.fpm_device_count.restype = C.c_int
lib.fpm_ctx_create.argtypes = [C.c_int, C.POINTER(_VP)]
lib.fpm_ctx_destroy.argtypes = [_VP]
lib.fpm_last_error.restype = C.c_char_p
lib.fpm_ctx_sync.argtypes = [_VP]
lib.fpm_ctx_stream.restype = _VP
lib.fpm_ctx_stream.argtypes = [_VP]
lib.fpm_ctx_set_stream.argtypes = [_VP, _VP]
lib.fpm_host_alloc.argtypes = [C.c_size_t, C.POINTER(_VP)]
lib.fpm_host_free.argtypes = [_VP]
lib.fpm_ctx_launch_count.restype = C.c_uint64
lib.fpm_ctx_launch_count.argtypes = [_VP]
lib.fpm_sketch_batch.argtypes = [_VP, C.POINTER(SketchParams), _VP, C.c_uint64, _VP, C.c_uint32, _VP, _VP, _VP, _VP]
lib.fpm_sketch_batch_dev.argtypes = [_VP, C.POINTER(SketchParams), _VP, C.c_uint64, _VP, C.c_uint32, _VP, _VP, _VP, _VP]
lib.fpm_sketch_stream_begin.argtypes = [_VP]
lib.fpm_sketch_stream_append.argtypes = [_VP, _VP, C.c_uint64]
lib.fpm_sketch_stream_end_group.argtypes = [_VP]
lib.fpm_sketch_stream_append_async.argtypes = [_VP, _VP, C.c_uint64, u64p]
lib.fpm_sketch_stream_wait.argtypes = [_VP, C.c_uint64]
lib.fpm_sketch_stream_finish.argtypes = [_VP, C.POINTER(SketchParams), _VP, _VP, _VP, _VP]
lib.fpm_kmer_hashes.argtypes = [_VP, C.POINTER(SketchParams), _VP, C.c_uint64, _VP, u64p]
lib.fpm_fp_hash_batch.argtypes = [_VP, _VP, _VP, C.c_uint64, C.c_uint32, C.c_int, _VP]
lib.fpm_cfl_fingerprint_batch.argtypes = [_VP, _VP, _VP, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, _VP, _VP, _VP, _VP]
lib.fpm_fingerprint_batch.argtypes = [_VP, _VP, _VP, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_int, _VP, _VP, _VP, _VP]
FACT_CFL, FACT_ICFL, FACT_CFL_ICFL, FACT_CFL_COMB, FACT_ICFL_COMB, FACT_CFL_ICFL_COMB = 0, 1, 2, 3, 4, 5
lib.fpm_dist_tile.argtypes = [_VP, C.POINTER(DistParams), C.POINTER(Panel), C.POINTER(Panel), _VP]
lib.fpm_fp_positional_tile.argtypes = [_VP, C.POINTER(DistParams), C.POINTER(Panel), C.POINTER(Panel), _VP]
lib.fpm_dist_tile_dev.argtypes = [_VP, C.POINTER(DistParams), C.POINTER(Panel), C.POINTER(Panel), _VP, _VP]
lib.fpm_pvalue.restype = C.c_double
lib.fpm_pvalue.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_uint64]
lib.fpm_distance.restype = C.c_double
lib.fpm_distance.argtypes = [C.c_uint64, C.c_uint64, C.c_int]
lib.fpm_measure_int32_peak.argtypes = [_VP, C.POINTER(C.c_double)]
lib.fpm_get_int32_peaks.argtypes = [_VP, C.POINTER(C.c_double)]
lib.fpm_fastq_stream_append.argtypes = [_VP, _VP, C.c_uint64, C.c_uint32, C.POINTER(C.c_int), u64p]
lib.fpm_fastq_line_ends.argtypes = [_VP, C.c_uint64, C.c_uint64, _VP]
lib.fpm_fasta_parse.argtypes = [_VP, _VP, C.c_uint64, u64p, u64p, C.POINTER(C.c_int)]
lib.fpm_fasta_records.argtypes = [_VP, _VP]
lib.fpm_fasta_sequence.argtypes = [_VP, _VP]
lib.fpm_gunzip_batch.argtypes = [_VP, _VP, _VP, C.c_uint32, _VP, u64p, C.POINTER(C.c_int)]
lib.fpm_gunzip_output.argtypes = [_VP, _VP]
lib.fpm_fasta_headers.argtypes = [_VP, _VP, _VP]
lib.fpm_sketch_stream_append_parsed.argtypes = [_VP]
lib.fpm_sketch_parsed.argtypes = [_VP, C.POINTER(SketchParams), _VP, C.c_uint32, _VP, _VP, _VP]
FASTA_RECORD_DTYPE = np.dtype([("hdr_begin", "<u8"), ("hdr_end", "<u8"), ("seq_begin", "<u8")])
lib.fpm_dist_hits.argtypes = [_VP, _VP, _VP, _VP, _VP, C.c_uint64, u64p]
lib.fpm_dist_hits_dev.argtypes = [_VP, _VP, _VP, _VP, _VP, C.c_uint64, u64p, _VP]
HIT_DTYPE = np.dtype([("query", "<u4"), ("ref", "<u4"), ("numer", "<u4"), ("denom", "<u4"), ("distance", "<f8"), ("pvalue", "<f8")])
FPM_ERR_CAPACITY = -6
lib.fpm_ctx_set_timing.argtypes = [_VP, C.c_int]
lib.fpm_ctx_set_dist_mode.argtypes = [_VP, C.c_int]
lib.fpm_ctx_get_timing.argtypes = [_VP, C.c_int, C.POINTER(C.c_double), u64p]

KERNEL_SKETCH_HASH, KERNEL_SKETCH_SELECT, KERNEL_DIST_TILE, KERNEL_DIST_LITERAL, KERNEL_DIST_PACK, KERNEL_DIST_EXCHANGE = range(6)


class Block(C.Structure):
    _fields_ = [("q_begin", C.c_uint64), ("q_end", C.c_uint64), ("r_begin", C.c_uint64), ("r_end", C.c_uint64)]


COMM_ID_BYTES = 128
lib.fpm_shard_range.restype = None
lib.fpm_shard_range.argtypes = [C.c_uint64, C.c_int, C.c_int, u64p, u64p]
lib.fpm_dist_grid_shape.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.POINTER(C.c_int), C.POINTER(C.c_int)]
lib.fpm_dist_block.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(Block)]
lib.fpm_comm_get_unique_id.argtypes = [_VP]
lib.fpm_comm_init_rank.argtypes = [_VP, _VP, C.c_int, C.c_int]
lib.fpm_comm_adopt.argtypes = [_VP, _VP, C.c_int, C.c_int]
lib.fpm_comm_destroy.argtypes = [_VP]
lib.fpm_comm_rank.argtypes = [_VP]
lib.fpm_comm_size.argtypes = [_VP]
lib.fpm_dist_sharded_dev.argtypes = [_VP, C.POINTER(DistParams), C.POINTER(Panel), C.c_uint64, C.POINTER(Panel), C.c_uint64, _VP, C.c_uint64,
                                     C.POINTER(Block), _VP]
lib.fpm_dist_hits_sharded_dev.argtypes = [_VP, C.POINTER(DistParams), C.POINTER(Panel), C.c_uint64, C.POINTER(Panel), C.c_uint64, _VP, C.c_uint64,
                                          u64p, C.POINTER(Block), _VP]
lib.fpm_sketch_reads_sharded_dev.argtypes = [_VP, C.POINTER(SketchParams), _VP, C.c_uint64, _VP, _VP, _VP, _VP]
lib.fpm_dist_set_reference.argtypes = [_VP, C.POINTER(Panel)]
lib.fpm_multi_create.argtypes = [_VP, C.c_int, C.POINTER(_VP)]
lib.fpm_multi_destroy.argtypes = [_VP]
lib.fpm_multi_destroy.restype = None
lib.fpm_multi_size.argtypes = [_VP]
lib.fpm_multi_ctx.argtypes = [_VP, C.c_int]
lib.fpm_multi_ctx.restype = _VP
lib.fpm_dist_tile_multi.argtypes = [_VP, C.POINTER(DistParams), C.POINTER(Panel), C.POINTER(Panel), _VP]
lib.fpm_dist_hits_multi.argtypes = [_VP, C.POINTER(DistParams), C.POINTER(Panel), C.POINTER(Panel), _VP, C.c_uint64, u64p]
lib.fpm_sketch_batch_multi.argtypes = [_VP, C.POINTER(SketchParams), _VP, C.c_uint64, _VP, C.c_uint32, _VP, _VP, _VP, _VP]


def shard_range(n, part, parts):
    """Rows [n*part/parts, n*(part+1)/parts) -- fpm_shard_range."""
    a, b = C.c_uint64(0), C.c_uint64(0)
    lib.fpm_shard_range(n, part, parts, C.byref(a), C.byref(b))
    return int(a.value), int(b.value)


def dist_grid_shape(world, n_qry, n_ref):
    qp, rp = C.c_int(0), C.c_int(0)
    _check(lib.fpm_dist_grid_shape(world, n_qry, n_ref, C.byref(qp), C.byref(rp)))
    return qp.value, rp.value


def dist_block(rank, world, n_qry, n_ref):
    """(q_begin, q_end, r_begin, r_end) of the block `rank` compares -- fpm_dist_block."""
    b = Block()
    _check(lib.fpm_dist_block(rank, world, n_qry, n_ref, C.byref(b)))
    return int(b.q_begin), int(b.q_end), int(b.r_begin), int(b.r_end)


def comm_unique_id():
    buf = (C.c_uint8 * COMM_ID_BYTES)()
    _check(lib.fpm_comm_get_unique_id(buf))
    return bytes(buf)

EXPORTED = ["fpm_abi_version", "fpm_device_count", "fpm_ctx_create", "fpm_ctx_destroy", "fpm_last_error",
            "fpm_ctx_sync", "fpm_ctx_stream", "fpm_ctx_set_stream", "fpm_host_alloc", "fpm_host_free",
            "fpm_ctx_launch_count", "fpm_sketch_batch", "fpm_sketch_batch_dev", "fpm_kmer_hashes",
            "fpm_sketch_stream_begin", "fpm_sketch_stream_append", "fpm_sketch_stream_end_group", "fpm_sketch_stream_finish",
            "fpm_fp_hash_batch", "fpm_cfl_fingerprint_batch", "fpm_fingerprint_batch", "fpm_dist_tile", "fpm_dist_tile_dev", "fpm_fp_positional_tile", "fpm_pvalue", "fpm_distance",
            "fpm_measure_int32_peak", "fpm_get_int32_peaks", "fpm_ctx_set_timing", "fpm_ctx_get_timing", "fpm_ctx_set_dist_mode",
            "fpm_fasta_parse", "fpm_fasta_records", "fpm_fasta_sequence", "fpm_sketch_parsed",
            "fpm_fastq_stream_append", "fpm_fastq_line_ends", "fpm_dist_hits", "fpm_dist_hits_dev",
            "fpm_shard_range", "fpm_dist_grid_shape", "fpm_dist_block", "fpm_comm_get_unique_id", "fpm_comm_init_rank", "fpm_comm_adopt",
            "fpm_comm_destroy", "fpm_comm_rank", "fpm_comm_size", "fpm_dist_sharded_dev", "fpm_dist_hits_sharded_dev",
            "fpm_multi_create", "fpm_multi_destroy", "fpm_multi_size", "fpm_multi_ctx", "fpm_dist_tile_multi", "fpm_dist_hits_multi",
            "fpm_sketch_batch_multi", "fpm_sketch_reads_sharded_dev", "fpm_dist_set_reference", "fpm_sketch_stream_append_async", "fpm_sketch_stream_wait",
            "fpm_gunzip_batch", "fpm_gunzip_output", "fpm_fasta_headers", "fpm_sketch_stream_append_parsed"]


def _check(rc):
    if rc != FPM_OK:
        raise FpmError(rc, lib.fpm_last_error().decode("utf-8", "replace"))


class PinnedBuffer:
    """Page-locked host memory from fpm_host_alloc (placed on the current device's NUMA node) as a numpy uint8 array."""

    def __init__(self, nbytes):
        p = _VP()
        _check(lib.fpm_host_alloc(int(nbytes), C.byref(p)))
        self._p = p
        self.array = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(int(nbytes),))

    def close(self):
        if self._p:
            self.array = None
            lib.fpm_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def device_count():
    return lib.fpm_device_count()


def nucleotide_alphabet(chars="ACGT", preserve_case=False):
    """setAlphabetFromString (Sketch.cpp:1260-1289)."""
    a = np.zeros(256, dtype=np.uint8)
    for ch in chars:
        c = ord(ch)
        if not preserve_case and 96 < c < 123:
            c -= 32
        a[c] = 1
    return a


def make_sketch_params(k=21, s=1000, seed=42, min_cov=1, noncanonical=False, preserve_case=False,
                       alphabet="ACGT", want_counts=False, use64=None):
    """Sketch::Parameters as resolved by sketchParameterSetup (sketchParameterSetup.cpp:9-106)."""
    p = SketchParams()
    table = nucleotide_alphabet(alphabet, preserve_case)
    p.kmer_size, p.sketch_size, p.seed, p.min_cov = k, s, seed, min_cov
    p.noncanonical, p.preserve_case, p.want_counts = int(noncanonical), int(preserve_case), int(want_counts)
    if use64 is None:
        use64 = float(int(table.sum())) ** k > 2.0 ** 32   # Sketch.cpp:1288
    p.use64 = int(use64)
    C.memmove(p.alphabet, table.ctypes.data, 256)
    return p


def pack_records(groups):
    """groups: list of sketches, each a list of records (bytes).  Returns (seq u8 array,
    group_offsets u64 array) in the layout fpm_sketch_batch wants: records back to back,
    each followed by one 0x00 byte."""
    total = sum(len(r) + 1 for g in groups for r in g)
    seq = np.zeros(total, dtype=np.uint8)
    off = np.zeros(len(groups) + 1, dtype=np.uint64)
    pos = 0
    for gi, g in enumerate(groups):
        for r in g:
            n = len(r)
            seq[pos:pos + n] = np.frombuffer(bytes(r), dtype=np.uint8)
            pos += n + 1
        off[gi + 1] = pos
    return seq, off


class Context:
    """One CUDA device + stream (fpm_ctx)."""

    def __init__(self, device=0):
        self._h = _VP()
        _check(lib.fpm_ctx_create(device, C.byref(self._h)))

    def close(self):
        if self._h:
            lib.fpm_ctx_destroy(self._h)
            self._h = _VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def sync(self):
        _check(lib.fpm_ctx_sync(self._h))

    def stream(self):
        return lib.fpm_ctx_stream(self._h)

    def set_stream(self, cuda_stream):
        _check(lib.fpm_ctx_set_stream(self._h, _VP(cuda_stream)))

    def launch_count(self):
        return int(lib.fpm_ctx_launch_count(self._h))

    def set_dist_mode(self, force64=False, no_prune=False, no_group=False, saturate=False, group=False):
        """force64=True: run the 64-bit tile kernel even where the 32-bit rank kernel applies; no_prune=True: the rank
        kernel merges every pair (no skipping of pairs that share no hash); saturate=True: always bound the marking walks
        by the sizes of the reference components (by default only when the panels are large enough for that to pay); group=True:
        always reorder the panels so that related sketches share tiles (by default only when the marked pairs are scattered)."""
        _check(lib.fpm_ctx_set_dist_mode(self._h, 1 if force64 else (2 if no_prune else (3 if no_group else (4 if saturate else (5 if group else 0))))))

    def set_timing(self, enable=True):
        _check(lib.fpm_ctx_set_timing(self._h, int(enable)))

    def get_timing(self, kernel_id):
        """(total ms, launches) of one kernel since set_timing(True)."""
        ms = C.c_double()
        n = C.c_uint64()
        _check(lib.fpm_ctx_get_timing(self._h, kernel_id, C.byref(ms), C.byref(n)))
        return ms.value, int(n.value)

    def int32_peak(self):
        v = C.c_double()
        _check(lib.fpm_measure_int32_peak(self._h, C.byref(v)))
        return v.value

    def int32_peaks(self):
        """(ALU-only, IMAD-only, alternating) thread-instructions/s from the last int32_peak()."""
        a = (C.c_double * 3)()
        _check(lib.fpm_get_int32_peaks(self._h, a))
        return tuple(a)

    # -- sketch ---------------------------------------------------------------------------
    def sketch_batch(self, seq, group_offsets, params, want_kmers=False):
        """Host buffers in, host arrays out (H2D + kernels + D2H)."""
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        off = np.ascontiguousarray(group_offsets, dtype=np.uint64)
        ng = len(off) - 1
        s = params.sketch_size
        hashes = np.zeros((ng, s), dtype=np.uint64)
        counts = np.zeros((ng, s), dtype=np.uint32) if params.want_counts else None
        n = np.zeros(ng, dtype=np.uint32)
        kmers = np.zeros(ng, dtype=np.uint64) if want_kmers else None
        _check(lib.fpm_sketch_batch(self._h, C.byref(params), seq.ctypes.data, seq.size, off.ctypes.data, ng,
                                    hashes.ctypes.data, counts.ctypes.data if counts is not None else None,
                                    n.ctypes.data, kmers.ctypes.data if kmers is not None else None))
        return dict(hashes=hashes, counts=counts, n=n, kmers=kmers)

    def sketch_batch_dev(self, d_seq_ptr, seq_bytes, group_offsets, params, d_hashes_ptr, d_counts_ptr, d_n_ptr,
                         d_kmers_ptr=None):
        """Device pointers (ints) for sequence and outputs; group offsets stay on the host."""
        off = np.ascontiguousarray(group_offsets, dtype=np.uint64)
        _check(lib.fpm_sketch_batch_dev(self._h, C.byref(params), _VP(d_seq_ptr), seq_bytes, off.ctypes.data,
                                        len(off) - 1, _VP(d_hashes_ptr), _VP(d_counts_ptr) if d_counts_ptr else None,
                                        _VP(d_n_ptr), _VP(d_kmers_ptr) if d_kmers_ptr else None))

    def sketch_records(self, groups, **kw):
        """Convenience: groups = list of lists of record bytes -> list of dict per sketch."""
        want_kmers = kw.pop("want_kmers", False)
        params = make_sketch_params(**kw)
        seq, off = pack_records(groups)
        r = self.sketch_batch(seq, off, params, want_kmers=want_kmers)
        out = []
        for g in range(len(groups)):
            n = int(r["n"][g])
            d = dict(hashes=r["hashes"][g, :n].copy())
            if r["counts"] is not None:
                d["counts"] = r["counts"][g, :n].copy()
            if r["kmers"] is not None:
                d["kmers"] = int(r["kmers"][g])
            out.append(d)
        return out

    def sketch_stream(self, groups, piece=1 << 16, double_buffered=False, **kw):
        """Like sketch_records but through the streaming entry points, in pieces of `piece` bytes.  double_buffered: the pieces go
        through two pinned staging buffers with fpm_sketch_stream_append_async / _wait, as the CLI's reader does."""
        params = make_sketch_params(**kw)
        _check(lib.fpm_sketch_stream_begin(self._h))
        stages = [PinnedBuffer(piece), PinnedBuffer(piece)] if double_buffered else None
        tickets, cur = [None, None], 0
        for g in groups:
            buf = np.frombuffer(b"".join(bytes(r) + b"\0" for r in g), dtype=np.uint8).copy() if g else np.zeros(0, dtype=np.uint8)
            for p0 in range(0, buf.size, piece):
                chunk = np.ascontiguousarray(buf[p0:p0 + piece])
                if not double_buffered:
                    _check(lib.fpm_sketch_stream_append(self._h, chunk.ctypes.data, chunk.size))
                    continue
                if tickets[cur] is not None:
                    _check(lib.fpm_sketch_stream_wait(self._h, tickets[cur]))
                stages[cur].array[:chunk.size] = chunk
                t = C.c_uint64(0)
                _check(lib.fpm_sketch_stream_append_async(self._h, stages[cur].array.ctypes.data, chunk.size, C.byref(t)))
                tickets[cur] = t.value
                cur ^= 1
            _check(lib.fpm_sketch_stream_end_group(self._h))
        if double_buffered:
            for t in tickets:
                if t is not None:
                    _check(lib.fpm_sketch_stream_wait(self._h, t))
        ng, s = len(groups), params.sketch_size
        hashes = np.zeros((ng, s), dtype=np.uint64)
        counts = np.zeros((ng, s), dtype=np.uint32)
        n = np.zeros(ng, dtype=np.uint32)
        _check(lib.fpm_sketch_stream_finish(self._h, C.byref(params), hashes.ctypes.data,
                                            counts.ctypes.data if params.want_counts else None, n.ctypes.data, None))
        return [dict(hashes=hashes[g, :n[g]].copy(), counts=counts[g, :n[g]].copy()) for g in range(ng)]

    def sketch_fastq_pieces(self, pieces, **kw):
        """One read set given as raw four-line FASTQ pieces (each starting at a record boundary and ending with a newline):
        parsed on the GPU, sketched as ONE group.  Returns None if a piece is not clean FASTQ, else
        (dict(hashes, counts), infos) with infos[i] = (reads, reads >= k, bytes appended, first read >= k) of piece i."""
        params = make_sketch_params(**kw)
        _check(lib.fpm_sketch_stream_begin(self._h))
        infos = []
        for pc in pieces:
            buf = np.frombuffer(bytes(pc), dtype=np.uint8).copy()
            status = C.c_int(0)
            info = (C.c_uint64 * 6)()
            _check(lib.fpm_fastq_stream_append(self._h, buf.ctypes.data, buf.size, params.kmer_size, C.byref(status), info))
            if status.value != 0:
                _check(lib.fpm_sketch_stream_end_group(self._h))
                return None
            infos.append(tuple(int(x) for x in info[:4]))
        _check(lib.fpm_sketch_stream_end_group(self._h))
        s = params.sketch_size
        hashes = np.zeros((1, s), dtype=np.uint64)
        counts = np.zeros((1, s), dtype=np.uint32)
        n = np.zeros(1, dtype=np.uint32)
        _check(lib.fpm_sketch_stream_finish(self._h, C.byref(params), hashes.ctypes.data,
                                            counts.ctypes.data if params.want_counts else None, n.ctypes.data, None))
        return dict(hashes=hashes[0, :n[0]].copy(), counts=counts[0, :n[0]].copy()), infos

    def sketch_fasta_pieces(self, pieces, **kw):
        """One read set given as raw FASTA pieces (each starting at a record boundary): every piece is parsed on the GPU
        (fpm_fasta_parse) and its compacted sequence appended to the read stream device to device; sketched as ONE group.
        Returns None if a piece is not plain FASTA, else dict(hashes, counts)."""
        params = make_sketch_params(**kw)
        _check(lib.fpm_sketch_stream_begin(self._h))
        for pc in pieces:
            raw = np.frombuffer(bytes(pc) + b"\0", dtype=np.uint8).copy()
            nrec, nseq, status = C.c_uint64(0), C.c_uint64(0), C.c_int(0)
            _check(lib.fpm_fasta_parse(self._h, raw.ctypes.data, raw.size, C.byref(nrec), C.byref(nseq), C.byref(status)))
            if status.value != 0:
                _check(lib.fpm_sketch_stream_end_group(self._h))
                return None
            _check(lib.fpm_sketch_stream_append_parsed(self._h))
        _check(lib.fpm_sketch_stream_end_group(self._h))
        s = params.sketch_size
        hashes = np.zeros((1, s), dtype=np.uint64)
        counts = np.zeros((1, s), dtype=np.uint32)
        n = np.zeros(1, dtype=np.uint32)
        _check(lib.fpm_sketch_stream_finish(self._h, C.byref(params), hashes.ctypes.data,
                                            counts.ctypes.data if params.want_counts else None, n.ctypes.data, None))
        return dict(hashes=hashes[0, :n[0]].copy(), counts=counts[0, :n[0]].copy())

    def fastq_line_ends(self, first_line, n_lines):
        out = np.zeros(max(n_lines, 1), dtype=np.uint64)
        _check(lib.fpm_fastq_line_ends(self._h, first_line, n_lines, out.ctypes.data))
        return out[:n_lines]

    def kmer_hashes(self, record: bytes, **kw):
        """Hash of every valid window of one record, in order (getHash parity)."""
        params = make_sketch_params(**kw)
        seq = np.frombuffer(bytes(record) + b"\0", dtype=np.uint8).copy()
        out = np.zeros(max(seq.size, 1), dtype=np.uint64)
        cnt = C.c_uint64()
        _check(lib.fpm_kmer_hashes(self._h, C.byref(params), seq.ctypes.data, seq.size, out.ctypes.data, C.byref(cnt)))
        return out[:cnt.value].copy()

    # -- fingerprints ---------------------------------------------------------------------
    def fp_hash_batch(self, lines, seed=42, use64=False):
        """lines: list of token lists (one fingerprint line each) -> u64 array of hashes."""
        off = np.zeros(len(lines) + 1, dtype=np.uint64)
        for i, t in enumerate(lines):
            off[i + 1] = off[i] + len(t)
        tok = np.zeros(max(int(off[-1]), 1), dtype=np.uint64)
        p = 0
        for t in lines:
            tok[p:p + len(t)] = np.asarray(t, dtype=np.uint64)
            p += len(t)
        out = np.zeros(len(lines), dtype=np.uint64)
        _check(lib.fpm_fp_hash_batch(self._h, tok.ctypes.data, off.ctypes.data, len(lines), seed, int(use64), out.ctypes.data))
        return out

    def fingerprint_batch(self, records, window=100, factorization="CFL", seed=42, use64=False):
        """lyn2vec basic/shift on the GPU: records = list of bytes; factorization = "CFL", "ICFL", "CFL_ICFL-<C>", "CFL_COMB", "ICFL_COMB" or "CFL_ICFL_COMB-<C>"
        (lyn2vec's --type_factorization).  Returns (rows, hashes, window_offsets): rows[w] = list of factor lengths
        of window w, hashes[w] = getHashFingerPrint(rows[w])."""
        if factorization == "CFL":
            mode, sub = FACT_CFL, 0
        elif factorization == "ICFL":
            mode, sub = FACT_ICFL, 0
        elif factorization.startswith("CFL_ICFL-"):
            mode, sub = FACT_CFL_ICFL, int(factorization.split("-")[1])
        elif factorization == "CFL_COMB":
            mode, sub = FACT_CFL_COMB, 0
        elif factorization == "ICFL_COMB":
            mode, sub = FACT_ICFL_COMB, 0
        elif factorization.startswith("CFL_ICFL_COMB-"):
            mode, sub = FACT_CFL_ICFL_COMB, int(factorization.split("-")[1])
        else:
            raise ValueError("unknown factorization " + factorization)
        off = np.zeros(len(records) + 1, dtype=np.uint64)
        for i, r in enumerate(records):
            off[i + 1] = off[i] + len(r)
        seq = np.frombuffer(b"".join(bytes(r) for r in records) + b"\0", dtype=np.uint8).copy()
        woff = np.zeros(len(records) + 1, dtype=np.uint64)
        _check(lib.fpm_fingerprint_batch(self._h, seq.ctypes.data, off.ctypes.data, len(records), window, mode, sub, seed, int(use64),
                                         None, None, None, woff.ctypes.data))
        nw = int(woff[-1])
        hashes = np.zeros(nw, dtype=np.uint64)
        tok = np.zeros((nw, window), dtype=np.uint16)
        ntok = np.zeros(nw, dtype=np.uint16)
        _check(lib.fpm_fingerprint_batch(self._h, seq.ctypes.data, off.ctypes.data, len(records), window, mode, sub, seed, int(use64),
                                         hashes.ctypes.data, tok.ctypes.data, ntok.ctypes.data, woff.ctypes.data))
        rows = [tok[w, :ntok[w]].tolist() for w in range(nw)]
        return rows, hashes, woff

    def cfl_fingerprint_batch(self, records, window=100, seed=42, use64=False):
        return self.fingerprint_batch(records, window, "CFL", seed, use64)

    # -- FASTA ingestion ------------------------------------------------------------------
    def fasta_parse(self, files, fetch_sequence=True):
        """files: list of bytes (raw FASTA file contents, without 0x00).  Returns None when the input is not plain
        FASTA, else (records, lengths, sequence): records = FASTA_RECORD_DTYPE array, lengths[i] = sequence length
        of record i, sequence = the compacted batch (records each followed by 0x00) or None."""
        raw = np.frombuffer(b"".join(bytes(f) + b"\0" for f in files), dtype=np.uint8).copy()
        nrec, nseq, status = C.c_uint64(0), C.c_uint64(0), C.c_int(0)
        _check(lib.fpm_fasta_parse(self._h, raw.ctypes.data, raw.size, C.byref(nrec), C.byref(nseq), C.byref(status)))
        if status.value != 0:
            return None
        recs = np.zeros(nrec.value, dtype=FASTA_RECORD_DTYPE)
        if nrec.value:
            _check(lib.fpm_fasta_records(self._h, recs.ctypes.data))
        nxt = np.append(recs["seq_begin"][1:], np.uint64(nseq.value)) if nrec.value else np.zeros(0, dtype=np.uint64)
        lengths = (nxt - recs["seq_begin"] - np.uint64(1)).astype(np.uint64) if nrec.value else nxt
        seq = None
        if fetch_sequence:
            seq = np.zeros(max(nseq.value, 1), dtype=np.uint8)
            if nseq.value:
                _check(lib.fpm_fasta_sequence(self._h, seq.ctypes.data))
            seq = seq[:nseq.value]
        return recs, lengths, seq

    def gunzip_batch(self, gz_files, fetch=True):
        """gz_files: list of bytes (gzip files).  Inflates them on the device into the raw batch of fasta_parse (each file
        followed by 0x00).  Returns (status, file_end, raw): status 0 = all clean; file_end[i] = offset of file i's 0x00;
        raw = the batch copied back (fetch=True) or None."""
        sizes = np.array([0] + [len(g) for g in gz_files], dtype=np.uint64)
        off = np.cumsum(sizes).astype(np.uint64)
        blob = np.frombuffer(b"".join(bytes(g) for g in gz_files) + b"\0", dtype=np.uint8).copy()
        ends = np.zeros(max(len(gz_files), 1), dtype=np.uint64)
        total, status = C.c_uint64(0), C.c_int(0)
        _check(lib.fpm_gunzip_batch(self._h, blob.ctypes.data, off.ctypes.data, len(gz_files), ends.ctypes.data, C.byref(total), C.byref(status)))
        if status.value != 0:
            return status.value, None, None
        raw = None
        if fetch:
            raw = np.zeros(max(total.value, 1), dtype=np.uint8)
            _check(lib.fpm_gunzip_output(self._h, raw.ctypes.data))
            raw = raw[:total.value]
        return 0, ends[:len(gz_files)], raw

    def fasta_parse_resident(self, n_bytes, fetch_sequence=True, fetch_headers=True):
        """fasta_parse over the batch gunzip_batch left on the device.  Returns None when it is not plain FASTA, else
        (records, lengths, sequence, headers): headers[i] = bytes [hdr_begin, hdr_end) of record i."""
        nrec, nseq, status = C.c_uint64(0), C.c_uint64(0), C.c_int(0)
        _check(lib.fpm_fasta_parse(self._h, None, n_bytes, C.byref(nrec), C.byref(nseq), C.byref(status)))
        if status.value != 0:
            return None
        recs = np.zeros(nrec.value, dtype=FASTA_RECORD_DTYPE)
        if nrec.value:
            _check(lib.fpm_fasta_records(self._h, recs.ctypes.data))
        nxt = np.append(recs["seq_begin"][1:], np.uint64(nseq.value)) if nrec.value else np.zeros(0, dtype=np.uint64)
        lengths = (nxt - recs["seq_begin"] - np.uint64(1)).astype(np.uint64) if nrec.value else nxt
        seq = None
        if fetch_sequence:
            seq = np.zeros(max(nseq.value, 1), dtype=np.uint8)
            if nseq.value:
                _check(lib.fpm_fasta_sequence(self._h, seq.ctypes.data))
            seq = seq[:nseq.value]
        headers = None
        if fetch_headers:
            hoff = np.zeros(nrec.value + 1, dtype=np.uint64)
            hoff[1:] = np.cumsum(recs["hdr_end"] - recs["hdr_begin"])
            hb = np.zeros(max(int(hoff[-1]), 1), dtype=np.uint8)
            if nrec.value:
                _check(lib.fpm_fasta_headers(self._h, hoff.ctypes.data, hb.ctypes.data))
            headers = [hb[int(hoff[i]):int(hoff[i + 1])].tobytes() for i in range(nrec.value)]
        return recs, lengths, seq, headers

    def sketch_parsed(self, group_offsets, params):
        """Sketch the sequence the last fasta_parse left on the device."""
        goff = np.ascontiguousarray(group_offsets, dtype=np.uint64)
        n = len(goff) - 1
        s = params.sketch_size
        hashes = np.zeros((n, s), dtype=np.uint64)
        counts = np.zeros((n, s), dtype=np.uint32)
        out_n = np.zeros(n, dtype=np.uint32)
        _check(lib.fpm_sketch_parsed(self._h, C.byref(params), goff.ctypes.data, n, hashes.ctypes.data,
                                     counts.ctypes.data if params.want_counts else None, out_n.ctypes.data))
        return {"hashes": hashes, "counts": counts, "n": out_n}

    # -- dist -----------------------------------------------------------------------------
    @staticmethod
    def _panel(hashes, sizes, lengths):
        h = np.ascontiguousarray(hashes, dtype=np.uint64)
        if h.ndim != 2:
            raise ValueError("panel hashes must be [n][stride]")
        sz = np.ascontiguousarray(sizes, dtype=np.uint32)
        ln = np.ascontiguousarray(lengths, dtype=np.uint64)
        p = Panel(h.ctypes.data, sz.ctypes.data, ln.ctypes.data, h.shape[0], h.shape[1])
        return p, (h, sz, ln)

    def dist_set_reference(self, ref):
        """Upload and keep a reference panel (hashes, sizes, lengths), or drop it with None; afterwards pass ref=None to
        dist_tile / dist_hits."""
        if ref is None:
            _check(lib.fpm_dist_set_reference(self._h, None))
            self._ref_n = None
            return
        pr, keep = self._panel(*ref)
        _check(lib.fpm_dist_set_reference(self._h, C.byref(pr)))
        self._ref_n = int(pr.n)

    def dist_tile(self, ref, qry, sketch_size, kmer_size, kmer_space, max_distance=1.0, max_pvalue=1.0,
                  sorted_unique=True, out=None, raw=False):
        """ref / qry: (hashes [n][stride], sizes [n], lengths [n]).  Returns a structured array
        [n_qry][n_ref] of (numer, denom, distance, pvalue) plus a bool `pass` matrix.  `out` may be a
        caller-owned (e.g. pinned) PAIR_DTYPE array of that shape.  ref=None: the resident reference panel."""
        if ref is None:
            pq, keep_q = self._panel(*qry)
            dp = DistParams(sketch_size, kmer_size, kmer_space, max_distance, max_pvalue, int(sorted_unique))
            if out is None:
                out = np.zeros((pq.n, self._ref_n), dtype=PAIR_DTYPE)
            _check(lib.fpm_dist_tile(self._h, C.byref(dp), None, C.byref(pq), out.ctypes.data))
            if raw:
                return out, None
            passed = (out["denom"] & FPM_PAIR_PASS) != 0
            out["denom"] &= 0x7fffffff
            return out, passed
        pr, keep_r = self._panel(*ref)
        pq, keep_q = self._panel(*qry)
        dp = DistParams(sketch_size, kmer_size, kmer_space, max_distance, max_pvalue, int(sorted_unique))
        if out is None:
            out = np.zeros((pq.n, pr.n), dtype=PAIR_DTYPE)
        assert out.dtype == PAIR_DTYPE and out.shape == (pq.n, pr.n) and out.flags["C_CONTIGUOUS"]
        _check(lib.fpm_dist_tile(self._h, C.byref(dp), C.byref(pr), C.byref(pq), out.ctypes.data))
        if raw:   # the records exactly as the C ABI wrote them (pass flag still in bit 31 of denom)
            return out, None
        passed = (out["denom"] & FPM_PAIR_PASS) != 0
        out["denom"] &= 0x7fffffff
        return out, passed

    def dist_hits(self, ref, qry, sketch_size, kmer_size, kmer_space, max_distance=1.0, max_pvalue=1.0, sorted_unique=True,
                  capacity=None, out=None, raw=False):
        """`mash dist -d D -v P`: only the pairs that pass the filters, as HIT_DTYPE records sorted by (query, ref) (denom
        without the pass flag).  The capacity is grown and the call repeated when more pairs pass than fit; `out` may be a
        caller-owned (e.g. pinned) HIT_DTYPE array, which then fixes the capacity (FpmError FPM_ERR_CAPACITY if too small)."""
        pr, keep_r = self._panel(*ref)
        pq, keep_q = self._panel(*qry)
        dp = DistParams(sketch_size, kmer_size, kmer_space, max_distance, max_pvalue, int(sorted_unique))
        n = C.c_uint64(0)
        if out is not None:
            assert out.dtype == HIT_DTYPE and out.ndim == 1 and out.flags["C_CONTIGUOUS"]
            _check(lib.fpm_dist_hits(self._h, C.byref(dp), C.byref(pr), C.byref(pq), out.ctypes.data, out.size, C.byref(n)))
        else:
            cap = int(capacity) if capacity is not None else max(1024, 8 * int(pq.n + pr.n))
            while True:
                out = np.zeros(cap, dtype=HIT_DTYPE)
                rc = lib.fpm_dist_hits(self._h, C.byref(dp), C.byref(pr), C.byref(pq), out.ctypes.data, cap, C.byref(n))
                if rc != FPM_ERR_CAPACITY:
                    _check(rc)
                    break
                cap = int(n.value)
        hits = out[:n.value]
        if not raw:   # raw: the records exactly as the C ABI wrote them (every hit carries the pass flag in bit 31 of denom)
            hits["denom"] &= 0x7fffffff
        return hits

    def dist_hits_dev(self, ref_ptrs, qry_ptrs, sketch_size, kmer_size, kmer_space, d_out_ptr, capacity, d_steps_ptr=None,
                      max_distance=1.0, max_pvalue=1.0, sorted_unique=True):
        """Device panels in, sorted HIT_DTYPE records at d_out_ptr (room for `capacity`); returns the number of hits."""
        pr = Panel(*ref_ptrs)
        pq = Panel(*qry_ptrs)
        dp = DistParams(sketch_size, kmer_size, kmer_space, max_distance, max_pvalue, int(sorted_unique))
        n = C.c_uint64(0)
        _check(lib.fpm_dist_hits_dev(self._h, C.byref(dp), C.byref(pr), C.byref(pq), _VP(d_out_ptr), int(capacity), C.byref(n),
                                     _VP(d_steps_ptr) if d_steps_ptr else None))
        return int(n.value)

    def fp_positional_tile(self, ref, qry, max_distance=1.0, max_pvalue=1.0):
        """compareFingerprints (mash triangle -fp): positional matches, chi-square p-value."""
        pr, keep_r = self._panel(*ref)
        pq, keep_q = self._panel(*qry)
        dp = DistParams(0, 1, 10.0, max_distance, max_pvalue, 0)
        out = np.zeros((pq.n, pr.n), dtype=PAIR_DTYPE)
        _check(lib.fpm_fp_positional_tile(self._h, C.byref(dp), C.byref(pr), C.byref(pq), out.ctypes.data))
        passed = (out["denom"] & FPM_PAIR_PASS) != 0
        out["denom"] &= 0x7fffffff
        return out, passed

    def dist_tile_dev(self, ref_ptrs, qry_ptrs, sketch_size, kmer_size, kmer_space, d_out_ptr, d_steps_ptr=None,
                      max_distance=1.0, max_pvalue=1.0, sorted_unique=True):
        """ref_ptrs / qry_ptrs: (hashes_ptr, sizes_ptr, lengths_ptr, n, stride) device pointers."""
        pr = Panel(*ref_ptrs)
        pq = Panel(*qry_ptrs)
        dp = DistParams(sketch_size, kmer_size, kmer_space, max_distance, max_pvalue, int(sorted_unique))
        _check(lib.fpm_dist_tile_dev(self._h, C.byref(dp), C.byref(pr), C.byref(pq), _VP(d_out_ptr),
                                     _VP(d_steps_ptr) if d_steps_ptr else None))


    # -- several GPUs, one process each -------------------------------------------------------
    def comm_init(self, unique_id, rank, world):
        """Collective: create this context's NCCL communicator from the id rank 0 got from comm_unique_id()."""
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(unique_id)
        _check(lib.fpm_comm_init_rank(self._h, buf, rank, world))

    def comm_destroy(self):
        _check(lib.fpm_comm_destroy(self._h))

    def sketch_reads_sharded_dev(self, d_seq_ptr, seq_bytes, params, d_hashes_ptr, d_counts_ptr, d_n_ptr, d_kmers_ptr=None):
        """Collective: this rank's contiguous part of ONE read set in, the complete sketch out on every rank."""
        _check(lib.fpm_sketch_reads_sharded_dev(self._h, C.byref(params), _VP(d_seq_ptr), int(seq_bytes), _VP(d_hashes_ptr),
                                                _VP(d_counts_ptr) if d_counts_ptr else None, _VP(d_n_ptr), _VP(d_kmers_ptr) if d_kmers_ptr else None))

    def dist_sharded_dev(self, ref_shard_ptrs, n_ref_total, qry_shard_ptrs, n_qry_total, sketch_size, kmer_size, kmer_space, d_out_ptr,
                         out_capacity, d_steps_ptr=None, max_distance=1.0, max_pvalue=1.0, sorted_unique=True):
        """Collective all-ranks comparison: this rank's row shards in, its block of the result matrix out (dense, at d_out_ptr).
        Returns the block (q_begin, q_end, r_begin, r_end)."""
        pr = Panel(*ref_shard_ptrs)
        pq = Panel(*qry_shard_ptrs)
        dp = DistParams(sketch_size, kmer_size, kmer_space, max_distance, max_pvalue, int(sorted_unique))
        b = Block()
        _check(lib.fpm_dist_sharded_dev(self._h, C.byref(dp), C.byref(pr), n_ref_total, C.byref(pq), n_qry_total, _VP(d_out_ptr), int(out_capacity),
                                        C.byref(b), _VP(d_steps_ptr) if d_steps_ptr else None))
        return int(b.q_begin), int(b.q_end), int(b.r_begin), int(b.r_end)

    def dist_hits_sharded_dev(self, ref_shard_ptrs, n_ref_total, qry_shard_ptrs, n_qry_total, sketch_size, kmer_size, kmer_space, d_out_ptr,
                              capacity, d_steps_ptr=None, max_distance=1.0, max_pvalue=1.0, sorted_unique=True):
        pr = Panel(*ref_shard_ptrs)
        pq = Panel(*qry_shard_ptrs)
        dp = DistParams(sketch_size, kmer_size, kmer_space, max_distance, max_pvalue, int(sorted_unique))
        b = Block()
        n = C.c_uint64(0)
        _check(lib.fpm_dist_hits_sharded_dev(self._h, C.byref(dp), C.byref(pr), n_ref_total, C.byref(pq), n_qry_total, _VP(d_out_ptr), int(capacity),
                                             C.byref(n), C.byref(b), _VP(d_steps_ptr) if d_steps_ptr else None))
        return int(n.value), (int(b.q_begin), int(b.q_end), int(b.r_begin), int(b.r_end))


class Multi:
    """All GPUs of the box driven from one process (fpm_multi_*): what `mash dist` / `mash sketch` use."""

    def __init__(self, devices=None, n_devices=0):
        h = _VP()
        if devices is not None:
            arr = (C.c_int * len(devices))(*devices)
            _check(lib.fpm_multi_create(arr, len(devices), C.byref(h)))
        else:
            _check(lib.fpm_multi_create(None, n_devices, C.byref(h)))
        self._h = h

    def close(self):
        if self._h:
            lib.fpm_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def size(self):
        return lib.fpm_multi_size(self._h)

    def dist_tile(self, ref, qry, sketch_size, kmer_size, kmer_space, max_distance=1.0, max_pvalue=1.0, sorted_unique=True, out=None, raw=False):
        pr, keep_r = Context._panel(*ref)
        pq, keep_q = Context._panel(*qry)
        dp = DistParams(sketch_size, kmer_size, kmer_space, max_distance, max_pvalue, int(sorted_unique))
        if out is None:
            out = np.zeros((pq.n, pr.n), dtype=PAIR_DTYPE)
        assert out.dtype == PAIR_DTYPE and out.shape == (pq.n, pr.n) and out.flags["C_CONTIGUOUS"]
        _check(lib.fpm_dist_tile_multi(self._h, C.byref(dp), C.byref(pr), C.byref(pq), out.ctypes.data))
        if raw:
            return out, None
        passed = (out["denom"] & FPM_PAIR_PASS) != 0
        out["denom"] &= 0x7fffffff
        return out, passed

    def dist_hits(self, ref, qry, sketch_size, kmer_size, kmer_space, max_distance=1.0, max_pvalue=1.0, sorted_unique=True, capacity=None, raw=False):
        pr, keep_r = Context._panel(*ref)
        pq, keep_q = Context._panel(*qry)
        dp = DistParams(sketch_size, kmer_size, kmer_space, max_distance, max_pvalue, int(sorted_unique))
        n = C.c_uint64(0)
        cap = int(capacity) if capacity is not None else max(1024, 8 * int(pq.n + pr.n))
        while True:
            out = np.zeros(cap, dtype=HIT_DTYPE)
            rc = lib.fpm_dist_hits_multi(self._h, C.byref(dp), C.byref(pr), C.byref(pq), out.ctypes.data, cap, C.byref(n))
            if rc != FPM_ERR_CAPACITY:
                _check(rc)
                break
            cap = int(n.value)
        hits = out[:n.value]
        if not raw:
            hits["denom"] &= 0x7fffffff
        return hits

    def sketch_batch(self, seq, group_offsets, params, want_kmers=False):
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        goff = np.ascontiguousarray(group_offsets, dtype=np.uint64)
        n = len(goff) - 1
        s = params.sketch_size
        hashes = np.zeros((n, s), dtype=np.uint64)
        counts = np.zeros((n, s), dtype=np.uint32) if params.want_counts else None
        out_n = np.zeros(n, dtype=np.uint32)
        kmers = np.zeros(n, dtype=np.uint64) if want_kmers else None
        _check(lib.fpm_sketch_batch_multi(self._h, C.byref(params), seq.ctypes.data, seq.size, goff.ctypes.data, n, hashes.ctypes.data,
                                          counts.ctypes.data if counts is not None else None, out_n.ctypes.data,
                                          kmers.ctypes.data if kmers is not None else None))
        return {"hashes": hashes, "counts": counts, "n": out_n, "kmers": kmers}


def pvalue(x, len_ref, len_qry, kmer_space, n):
    return lib.fpm_pvalue(x, len_ref, len_qry, kmer_space, n)


def distance(common, denom, kmer_size):
    return lib.fpm_distance(common, denom, kmer_size)
