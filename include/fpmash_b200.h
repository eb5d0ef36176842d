/* fpmash_b200.h -- C ABI of the B200-native fp-mash hot path (mash sketch / mash dist).
 *
 * The reference (UmbertoDellaMonica/fp-mash, Mash 2.3 fork) has no FFI layer; the seam this
 * library sits behind is the worker-function contract of its ThreadPool plus the free
 * functions of libmash.a.  Each entry point below names the
 * reference interface it replaces (paths relative to mash/src/mash/).
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns FPM_OK (0) or a negative
 *     fpm_status, and fpm_last_error() gives the message for the calling thread;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails
 *     with FPM_ERR_NO_DEVICE;
 *   - "_dev" variants take device pointers (inputs already resident in HBM, outputs left
 *     in HBM) and are asynchronous on the context's stream; the plain variants take host
 *     pointers and include the H2D/D2H copies;
 *   - a context is bound to one device and one stream; use one context per host thread
 *     (the reference calls its workers from -p pthreads: ThreadPool.hxx:182-230).
 */
#ifndef FPMASH_B200_H
#define FPMASH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FPM_ABI_VERSION 2

typedef enum fpm_status {
    FPM_OK = 0,
    FPM_ERR_NO_DEVICE = -1,   /* no CUDA device / driver: the product path refuses to run   */
    FPM_ERR_CUDA = -2,        /* a CUDA runtime call failed (message has the CUDA error)    */
    FPM_ERR_ARG = -3,         /* invalid argument                                            */
    FPM_ERR_UNSUPPORTED = -4, /* valid in the reference but outside the accelerated path     */
    FPM_ERR_NOMEM = -5,
    /* -6 is FPM_ERR_CAPACITY (see fpm_dist_hits) */
    FPM_ERR_COMM = -7         /* NCCL unavailable or a collective failed                      */
} fpm_status;

typedef struct fpm_ctx fpm_ctx;

/* ---- context ------------------------------------------------------------------------- */

int fpm_abi_version(void);
int fpm_device_count(void);                          /* 0 when no usable CUDA device        */
int fpm_ctx_create(int device, fpm_ctx** out);
void fpm_ctx_destroy(fpm_ctx* ctx);
const char* fpm_last_error(void);                    /* thread-local, never NULL            */
int fpm_ctx_sync(fpm_ctx* ctx);                      /* cudaStreamSynchronize(ctx stream)   */
void* fpm_ctx_stream(fpm_ctx* ctx);                  /* the cudaStream_t, for event timing  */
/* Adopt an external stream (e.g. torch's current stream) for all later launches.          */
int fpm_ctx_set_stream(fpm_ctx* ctx, void* cuda_stream);

/* Pinned host memory for zero-staging H2D/D2H (cudaHostAlloc / cudaFreeHost).             */
int fpm_host_alloc(size_t bytes, void** out);
void fpm_host_free(void* p);

/* Kernel launches issued by this context so far (all of them are this library's own).     */
uint64_t fpm_ctx_launch_count(const fpm_ctx* ctx);

/* ---- sketch: nucleotide k-mers -> bottom-s MinHash sketch ---------------------------- */

/* Mirrors the fields of Sketch::Parameters (Sketch.h:40-113) that reach the hot loop.     */
typedef struct fpm_sketch_params {
    int32_t kmer_size;        /* -k, 1..32                                                  */
    uint32_t sketch_size;     /* -s  (minHashesPerWindow)                                   */
    uint32_t seed;            /* -S                                                         */
    uint32_t min_cov;         /* -m  (1 unless read mode; MinHashHeap multiplicityMinimum)  */
    uint8_t noncanonical;     /* -n                                                         */
    uint8_t preserve_case;    /* -Z                                                         */
    uint8_t use64;            /* alphabetSize^k > 2^32 (Sketch.cpp:1288)                    */
    uint8_t want_counts;      /* fill out_counts (parameters.counts / read mode)            */
    uint8_t alphabet[256];    /* Sketch::Parameters::alphabet                               */
} fpm_sketch_params;

/* Replaces sketchFile / sketchSequence -> addMinHashes -> MinHashHeap::tryInsert ->
 * setMinHashesForReference (Sketch.cpp:1299-1517, 664-735, 1291-1297; MinHashHeap.cpp:68-146)
 * for a whole batch of sketches ("groups") at once.
 *
 * seq          : sequence bytes as kseq delivers them (any case, IUPAC codes allowed).
 *                Records are laid out back to back, EACH FOLLOWED BY ONE 0x00 BYTE, in the
 *                order the reference would feed them to addMinHashes; records shorter than k
 *                may be present (they produce no window, like Sketch.cpp:1374-1378).
 * group_offsets: n_groups+1 byte offsets into seq; group g = one output sketch = bytes
 *                [group_offsets[g], group_offsets[g+1]) (a whole number of records).
 * out_hashes   : [n_groups][sketch_size] ascending hashes (u64; 32-bit hashes zero-extended)
 * out_counts   : [n_groups][sketch_size] multiplicities, or NULL (HashSet::toHashList counts,
 *                including the reference's order-dependent count of the largest element)
 * out_n        : [n_groups] number of hashes in each sketch (<= sketch_size)
 * out_kmers    : [n_groups] number of valid k-mer windows hashed, or NULL
 */
int fpm_sketch_batch(fpm_ctx* ctx, const fpm_sketch_params* p,
                     const uint8_t* seq, uint64_t seq_bytes,
                     const uint64_t* group_offsets, uint32_t n_groups,
                     uint64_t* out_hashes, uint32_t* out_counts, uint32_t* out_n, uint64_t* out_kmers);

/* Same, device pointers.  d_seq must be 16-byte aligned.  Synchronises internally (the
 * exact bottom-s needs a host decision between passes) but leaves results in HBM.          */
int fpm_sketch_batch_dev(fpm_ctx* ctx, const fpm_sketch_params* p,
                         const uint8_t* d_seq, uint64_t seq_bytes,
                         const uint64_t* h_group_offsets, uint32_t n_groups,
                         uint64_t* d_out_hashes, uint32_t* d_out_counts, uint32_t* d_out_n, uint64_t* d_out_kmers);

/* Streaming variant for inputs larger than any host staging buffer (`mash sketch -r` turns a whole read set
 * into ONE sketch): begin, then append sequence pieces (same layout: records each followed by 0x00, a piece
 * may end anywhere between records) and close each sketch with end_group; the bytes accumulate in HBM and
 * finish() sketches all groups exactly like fpm_sketch_batch.  append() returns after its copy completed. */
int fpm_sketch_stream_begin(fpm_ctx* ctx);
int fpm_sketch_stream_append(fpm_ctx* ctx, const uint8_t* seq, uint64_t bytes);
/* append() without the wait: returns a ticket at once; the piece must not be modified before fpm_sketch_stream_wait(ticket)
 * returned.  Tickets complete in order, and waiting is only ever needed for the two most recent ones -- enough for the
 * intended use, two staging buffers filled in turn while the other one's copy is in flight.                              */
int fpm_sketch_stream_append_async(fpm_ctx* ctx, const uint8_t* seq, uint64_t bytes, uint64_t* ticket);
int fpm_sketch_stream_wait(fpm_ctx* ctx, uint64_t ticket);
int fpm_sketch_stream_end_group(fpm_ctx* ctx);
int fpm_sketch_stream_finish(fpm_ctx* ctx, const fpm_sketch_params* p, uint64_t* out_hashes, uint32_t* out_counts,
                             uint32_t* out_n, uint64_t* out_kmers);

/* Raw hash stream of one record (every valid window in order), for parity tests of the
 * rolling + MurmurHash3 kernel against getHash (hash.cpp:12-40).  out_hashes needs
 * seq_bytes entries; *out_count receives the number written.                               */
int fpm_kmer_hashes(fpm_ctx* ctx, const fpm_sketch_params* p, const uint8_t* seq, uint64_t seq_bytes,
                    uint64_t* out_hashes, uint64_t* out_count);

/* ---- sketch: -fp fingerprint mode ---------------------------------------------------- */

/* Replaces getHashFingerPrint (hash.cpp:45-73) as called per line by
 * Sketch::initFromFingerprints (Sketch.cpp:132): MurmurHash3_x64_128 over the raw
 * little-endian bytes of each line's uint64 tokens.  line_offsets has n_lines+1 entries
 * (token indices).  out_hashes[i] = (use64 ? h1 : low 32 bits of h1) of line i.            */
int fpm_fp_hash_batch(fpm_ctx* ctx, const uint64_t* tokens, const uint64_t* line_offsets, uint64_t n_lines,
                      uint32_t seed, int use64, uint64_t* out_hashes);

/* Upstream of -fp: lyn2vec's `--type basic --type_factorization CFL --shift shift` on the GPU.
 * Replaces shift_string + CFL (Duval) + the row writer (lyn2vec/fingerprint_utils.py:95-110,443-476,
 * lyn2vec/factorizations.py:102-126): for every record, every circular window of `window` characters (one
 * window = the record itself when it is shorter) is Lyndon-factorised; the row of factor LENGTHS is hashed
 * like getHashFingerPrint.  Records are raw bytes back to back (no separators); rec_offsets has n_records+1
 * entries.  out_window_offsets[n_records+1] (required) receives the first window index of each record;
 * out_hashes[n_windows], out_tokens[n_windows][window] (lengths, row-padded) and out_ntokens[n_windows] are
 * optional.  Call once with all three NULL to learn n_windows = out_window_offsets[n_records].             */
int fpm_cfl_fingerprint_batch(fpm_ctx* ctx, const uint8_t* seq, const uint64_t* rec_offsets, uint32_t n_records, uint32_t window,
                              uint32_t seed, int use64, uint64_t* out_hashes, uint16_t* out_tokens, uint16_t* out_ntokens,
                              uint64_t* out_window_offsets);
/* The same with the factorisation chosen (lyn2vec --type_factorization): FPM_FACT_CFL (Duval,
 * factorizations.py:102-126), FPM_FACT_ICFL (inverse Lyndon factorisation, ICFL_recursive, :143-248) or
 * FPM_FACT_CFL_ICFL (CFL_icfl, :265-300: CFL factors longer than sub_len are sub-factorised with ICFL; lyn2vec's
 * CFL_ICFL-10/-20/-30 are sub_len = 10/20/30).  sub_len is ignored by the other two.                           */
#define FPM_FACT_CFL 0
#define FPM_FACT_ICFL 1
#define FPM_FACT_CFL_ICFL 2
/* lyn2vec's *_COMB forms (d_cfl / d_icfl / d_cfl_icfl -> d_duval_, lyn2vec/factorizations_comb.py:178-245): the factor
 * boundaries of the window together with the mirrored boundaries of the factorisation of its reverse complement.        */
#define FPM_FACT_CFL_COMB 3
#define FPM_FACT_ICFL_COMB 4
#define FPM_FACT_CFL_ICFL_COMB 5
int fpm_fingerprint_batch(fpm_ctx* ctx, const uint8_t* seq, const uint64_t* rec_offsets, uint32_t n_records, uint32_t window,
                          int factorization, uint32_t sub_len, uint32_t seed, int use64, uint64_t* out_hashes,
                          uint16_t* out_tokens, uint16_t* out_ntokens, uint64_t* out_window_offsets);

/* ---- FASTA ingestion on the GPU -------------------------------------- */

/* Replaces the byte loop of the reference's reader (kseq.h:170-208, called from sketchFile, Sketch.cpp:1340-1422)
 * for plain FASTA: `raw` holds the bytes of one or more files, EACH FOLLOWED BY ONE 0x00 BYTE (a file that itself
 * contains 0x00 must be read on the host).  On the device the sequence bytes are compacted into the batch layout
 * of fpm_sketch_batch (records back to back, each followed by 0x00) and stay resident in the context; a record
 * table comes back.  Semantics as kseq: a record starts at '>', its header runs to the next '\n', its sequence is
 * every byte 33..126 up to the next '>' (anywhere in a line); bytes before a file's first '>' are skipped.
 * *out_status = FPM_FASTA_NOT_PLAIN when a '+' or '@' ends a record (FASTQ): nothing is produced, use the host
 * reader.  Record i's sequence length = (seq_begin of record i+1, or *out_seq_bytes for the last) - seq_begin - 1. */
typedef struct fpm_fasta_record {
    uint64_t hdr_begin;   /* offset in raw of the '>'                                              */
    uint64_t hdr_end;     /* offset in raw of the '\n' (or the file's 0x00) that ends the header   */
    uint64_t seq_begin;   /* offset of the record's first base in the compacted sequence          */
} fpm_fasta_record;
#define FPM_FASTA_OK 0
#define FPM_FASTA_NOT_PLAIN 1
int fpm_fasta_parse(fpm_ctx* ctx, const uint8_t* raw /* NULL: the batch fpm_gunzip_batch left on the device */, uint64_t n_bytes, uint64_t* out_n_records,
                    uint64_t* out_seq_bytes, int* out_status);
int fpm_fasta_records(fpm_ctx* ctx, fpm_fasta_record* out /* [n_records] */);
int fpm_fasta_sequence(fpm_ctx* ctx, uint8_t* out /* [seq_bytes], tests */);
/* gzip'ed FASTA files inflated on the device (the reference reads every input through zlib's gzopen / gzread,
 * Sketch.cpp:1340-1346, kseq.h:60-75).  `gz` holds the COMPRESSED bytes of n_files files back to back, file i at
 * [gz_offsets[i], gz_offsets[i+1]).  One warp per file inflates it (all members of a multi-member file) into the raw
 * batch of fpm_fasta_parse -- every file followed by one 0x00 -- which stays resident in the context:
 * out_file_end[i] = offset of file i's 0x00, *out_total = bytes of the batch.  Follow with
 * fpm_fasta_parse(ctx, NULL, *out_total, ...) (raw == NULL: parse the resident batch) and fpm_fasta_headers.
 * The trailer of every member is checked like zlib does (CRC-32, ISIZE).  *out_status != FPM_GUNZIP_OK: some file is
 * not a clean gzip stream (or holds a 0x00 byte) and nothing is resident -- read the batch with the host reader.   */
#define FPM_GUNZIP_OK 0
#define FPM_GUNZIP_BAD_STREAM 1
#define FPM_GUNZIP_HAS_NUL 2
int fpm_gunzip_batch(fpm_ctx* ctx, const uint8_t* gz, const uint64_t* gz_offsets /* [n_files+1] */, uint32_t n_files,
                     uint64_t* out_file_end /* [n_files] */, uint64_t* out_total, int* out_status);
int fpm_gunzip_output(fpm_ctx* ctx, uint8_t* out /* [*out_total], tests */);
/* Header text of the records of the last fpm_fasta_parse, from the raw batch on the device (the host has no copy of it
 * after fpm_gunzip_batch): bytes [hdr_begin, hdr_end) of record i go to out[offsets[i] .. offsets[i+1]); the caller
 * computes offsets from the record table (offsets[i+1] - offsets[i] = hdr_end - hdr_begin).                        */
int fpm_fasta_headers(fpm_ctx* ctx, const uint64_t* offsets /* [n_records+1] */, uint8_t* out);
/* Four-line FASTQ (one read set -> one sketch, `mash sketch -r`): a piece of a file -- starting at a record
 * boundary, ending with '\n' -- is parsed on the device and its reads are appended to the HBM-resident stream
 * of fpm_sketch_stream_begin / _end_group / _finish, each read followed by 0x00, exactly as
 * fpm_sketch_stream_append would have received them from the host reader.  Anything the reference's reader
 * (kseq.h:170-208) would treat differently from "header, sequence, '+' line, quality of equal length" is reported
 * as *out_status = FPM_FASTA_NOT_PLAIN and nothing is appended: restart the read set with the host reader.  (CRLF line
 * ends are fine: the '\r' of a sequence or quality line is dropped, as the reader drops it.)
 * out_info: [0] reads in the piece, [1] reads of at least min_len bases (sketchFile skips the others,
 * Sketch.cpp:1374-1378), [2] bytes appended, [3] index of the first read of at least min_len bases (= [0] if none).
 * fpm_fastq_line_ends returns offsets (in the piece) of the '\n' ending lines first_line .. first_line+n_lines-1 of the
 * last piece: read r's header is the text between line 4r-1's end (or the piece start) and line 4r's end.      */
#define FPM_FASTQ_INFO_WORDS 6
int fpm_fastq_stream_append(fpm_ctx* ctx, const uint8_t* raw, uint64_t n_bytes, uint32_t min_len, int* out_status, uint64_t* out_info /* [6] */);
int fpm_fastq_line_ends(fpm_ctx* ctx, uint64_t first_line, uint64_t n_lines, uint64_t* out);
/* FASTA read sets (`mash sketch -r reads.fa`): the compacted sequence of the last fpm_fasta_parse (records back to back,
 * each followed by 0x00 = the stream's layout) is appended to the HBM read stream of fpm_sketch_stream_begin, device to
 * device.  The caller cuts the file into pieces at record boundaries ("\n>") and parses piece after piece.              */
int fpm_sketch_stream_append_parsed(fpm_ctx* ctx);
/* fpm_sketch_batch over the sequence the last fpm_fasta_parse left on the device (group_offsets in compacted
 * coordinates, from 0 to seq_bytes).                                                                        */
int fpm_sketch_parsed(fpm_ctx* ctx, const fpm_sketch_params* p, const uint64_t* group_offsets, uint32_t n_groups, uint64_t* out_hashes,
                      uint32_t* out_counts, uint32_t* out_n);

/* ---- dist ---------------------------------------------------------------------------- */

/* PairOutput (CommandDistance.h:57-64) as a 24-byte POD.                                   */
typedef struct fpm_pair {
    uint32_t numer;           /* shared hashes (common)                                     */
    uint32_t denom;           /* bits 0..30 union size compared; bit 31 = pass              */
    double distance;
    double pvalue;
} fpm_pair;
#define FPM_PAIR_PASS 0x80000000u
#define FPM_PAIR_DENOM(p) ((p).denom & 0x7fffffffu)

typedef struct fpm_dist_params {
    uint32_t sketch_size;     /* min(ref, query minHashesPerWindow), CommandDistance.cpp:342 */
    int32_t kmer_size;
    double kmer_space;        /* alphabetSize^k (Sketch.cpp:661)                             */
    double max_distance;      /* -d                                                          */
    double max_pvalue;        /* -v                                                          */
    uint8_t sorted_unique;    /* 1: lists are ascending and duplicate-free (every nucleotide
                                 sketch); 0: run the reference loop literally (fp-mode lists
                                 are unsorted and may repeat)               */
} fpm_dist_params;

/* A panel is a dense [n][stride] u64 array of hashes plus per-sketch sizes and lengths.   */
typedef struct fpm_panel {
    const uint64_t* hashes;   /* [n][stride]                                                */
    const uint32_t* sizes;    /* [n] hashes used in each row                                */
    const uint64_t* lengths;  /* [n] Reference::length (for the p-value)                    */
    uint64_t n;
    uint64_t stride;
} fpm_panel;

/* Replaces compare() -> compareSketches() -> pValue() (CommandDistance.cpp:335-450) for the
 * whole tile query x ref.  out[q * ref.n + r] (query-major, reference-minor: the order of
 * CommandDistance.cpp:355-359).                                                            */
int fpm_dist_tile(fpm_ctx* ctx, const fpm_dist_params* p, const fpm_panel* ref, const fpm_panel* qry, fpm_pair* out);
int fpm_dist_tile_dev(fpm_ctx* ctx, const fpm_dist_params* p, const fpm_panel* d_ref, const fpm_panel* d_qry,
                      fpm_pair* d_out, uint64_t* d_merge_steps /* nullable: += loop iterations */);

/* A reference panel that stays resident: uploaded, sorted and indexed ONCE, then compared with any number of query panels
 * (chunks of a query set too large for one call, or the rows of a .msh file as they are read).  The reference keeps its
 * whole sketch file loaded while queries stream past, CommandDistance.cpp:163-266; here the index (codes, posting lists)
 * is a property of the reference panel alone, so a query chunk costs only its own lookups.  After this call pass
 * ref = NULL to fpm_dist_tile / fpm_dist_hits; passing a panel again (or NULL here) drops the resident one.  The host
 * arrays may be released when the call returns.                                                                      */
int fpm_dist_set_reference(fpm_ctx* ctx, const fpm_panel* ref /* host; NULL: drop */);

/* The same comparison, but only the pairs that pass the -d / -v filters come back (what `mash dist -d D -v P`
 * prints, CommandDistance.cpp:303-333 skips every other pair): records sorted by (query, ref), i.e. in the
 * reference's output order.  No n x n result matrix exists anywhere: the tile kernel appends passing pairs to
 * `out` directly, and when the filters exclude distance 1 (max_distance < 1 or max_pvalue < 1) pairs that share
 * no hash are not even evaluated.  *n_hits = number of passing pairs; when it exceeds `capacity` the call
 * returns FPM_ERR_CAPACITY, `out` holds nothing useful and the caller retries with room for *n_hits.          */
typedef struct fpm_hit {
    uint32_t query, ref;      /* indices into the two panels                                  */
    uint32_t numer, denom;    /* as fpm_pair (denom carries FPM_PAIR_PASS)                    */
    double distance;
    double pvalue;
} fpm_hit;
#define FPM_ERR_CAPACITY (-6)
int fpm_dist_hits(fpm_ctx* ctx, const fpm_dist_params* p, const fpm_panel* ref, const fpm_panel* qry, fpm_hit* out,
                  uint64_t capacity, uint64_t* n_hits);
int fpm_dist_hits_dev(fpm_ctx* ctx, const fpm_dist_params* p, const fpm_panel* d_ref, const fpm_panel* d_qry, fpm_hit* d_out,
                      uint64_t capacity, uint64_t* n_hits, uint64_t* d_merge_steps /* nullable */);

/* ---- several GPUs -------------------------------------------------------- */

/* The reference spreads `mash dist` over its thread pool by cutting the query x reference pair space into chunks
 * (CommandDistance.cpp:224-261: <= 4096 consecutive pairs each, results consumed in submission order) and `mash sketch`
 * by handing whole files to pool threads (Sketch.cpp:353-355).  Here the pair space is cut into a q_parts x r_parts
 * GRID OF BLOCKS, one per GPU (fpm_dist_grid_shape minimises the rows a GPU has to index: 8 GPUs -> 2 x 4), and whole
 * sketches go to GPUs in contiguous byte-balanced ranges.  Row shard `part` of `parts` of an n-row panel is rows
 * [n*part/parts, n*(part+1)/parts) (fpm_shard_range); rank = q_index * r_parts + r_index; a query block is the union of
 * the row shards of its grid row, a reference block the union of q_parts consecutive row shards.                     */
typedef struct fpm_block {
    uint64_t q_begin, q_end;  /* query rows of the block                                    */
    uint64_t r_begin, r_end;  /* reference rows of the block                                */
} fpm_block;
void fpm_shard_range(uint64_t n, int part, int parts, uint64_t* begin, uint64_t* end);
int fpm_dist_grid_shape(int world, uint64_t n_qry, uint64_t n_ref, int* q_parts, int* r_parts);
int fpm_dist_block(int rank, int world, uint64_t n_qry, uint64_t n_ref, fpm_block* out);

/* One process per GPU: communicator hand-in.  Either adopt an ncclComm_t the host program already has
 * (never destroyed by this library), or let the library create one: rank 0 obtains an id (ncclGetUniqueId), the host
 * distributes its FPM_COMM_ID_BYTES bytes by whatever means it has (MPI, torch.distributed, a file), every rank calls
 * fpm_comm_init_rank -- collectively.  NCCL is loaded with dlopen at that point; FPM_ERR_COMM when it is missing.    */
#define FPM_COMM_ID_BYTES 128
int fpm_comm_get_unique_id(void* out_id /* [FPM_COMM_ID_BYTES] */);
int fpm_comm_init_rank(fpm_ctx* ctx, const void* id, int rank, int world);
int fpm_comm_adopt(fpm_ctx* ctx, void* nccl_comm, int rank, int world);
int fpm_comm_destroy(fpm_ctx* ctx);
int fpm_comm_rank(const fpm_ctx* ctx);               /* -1 without a communicator                */
int fpm_comm_size(const fpm_ctx* ctx);               /* 0 without a communicator                 */

/* compare() over all ranks (collective; replaces the fan-out of CommandDistance.cpp:224-261).  Every rank passes ITS row
 * shard of both panels, resident in HBM (d_*_shard->n must equal its fpm_shard_range of n_*_total; the row stride must be
 * the same on all ranks).  The one exchange step of the path: each shard is sent to exactly the ranks whose block holds
 * it (grouped ncclSend / ncclRecv over NVLink).  The rank then compares its block (fpm_dist_block) and leaves the result
 * in d_out_block as a dense [q_end - q_begin][r_end - r_begin] matrix, query-major like fpm_dist_tile: record (q, r) of
 * the whole comparison is d_out_block[(q - q_begin) * (r_end - r_begin) + (r - r_begin)] on the rank whose block holds
 * it, bit-identical to the single-GPU result.  out_capacity: room in d_out_block, in records.                        */
int fpm_dist_sharded_dev(fpm_ctx* ctx, const fpm_dist_params* p, const fpm_panel* d_ref_shard, uint64_t n_ref_total,
                         const fpm_panel* d_qry_shard, uint64_t n_qry_total, fpm_pair* d_out_block, uint64_t out_capacity,
                         fpm_block* block, uint64_t* d_merge_steps /* nullable */);
/* The same in hits mode (fpm_dist_hits_dev): the block's passing pairs, sorted by (query, ref), indices global.        */
int fpm_dist_hits_sharded_dev(fpm_ctx* ctx, const fpm_dist_params* p, const fpm_panel* d_ref_shard, uint64_t n_ref_total,
                              const fpm_panel* d_qry_shard, uint64_t n_qry_total, fpm_hit* d_out, uint64_t capacity, uint64_t* n_hits,
                              fpm_block* block, uint64_t* d_merge_steps /* nullable */);

/* ONE read set spread over the ranks (`mash sketch -r`: initFromReads + sketchFile turn a whole read set into a single
 * sketch on one thread, Sketch.cpp:203-210).  Collective.  Every rank passes a contiguous part of the read stream, whole
 * records in fpm_sketch_batch's layout, rank order = stream order.  Each rank hashes its part into a counting table of
 * its own; the candidates (hash, count, first position) are all-gathered and merged by key, and the bottom-s with
 * multiplicity >= min_cov is selected from the merged table (MinHashHeap's final set is order independent).
 * The one order-dependent number -- HashSet's count of the largest element of a full sketch -- is settled from the global
 * stream positions of the final hashes (an all-reduce over their position buckets).  Every rank receives the complete
 * result: d_out_hashes [sketch_size], d_out_counts [sketch_size] or NULL, d_out_n [1], d_out_kmers [1] or NULL --
 * bit-identical to fpm_sketch_batch_dev over the concatenated stream.  Nucleotide alphabet only.                      */
int fpm_sketch_reads_sharded_dev(fpm_ctx* ctx, const fpm_sketch_params* p, const uint8_t* d_seq, uint64_t seq_bytes,
                                 uint64_t* d_out_hashes, uint32_t* d_out_counts, uint32_t* d_out_n, uint64_t* d_out_kmers);

/* One process driving all GPUs of the box (what `mash dist` and `mash sketch` do): the panels are in host memory, so each
 * GPU uploads its two blocks and no collective is needed.  One host thread per GPU inside the call; results arrive in
 * `out` exactly as from the single-GPU entry point (full query-major matrix / hits in the reference's output order).  */
typedef struct fpm_multi fpm_multi;
int fpm_multi_create(const int* devices /* nullable: 0..n-1 */, int n_devices /* <= 0: all visible */, fpm_multi** out);
void fpm_multi_destroy(fpm_multi* m);
int fpm_multi_size(const fpm_multi* m);
fpm_ctx* fpm_multi_ctx(fpm_multi* m, int i);
int fpm_dist_tile_multi(fpm_multi* m, const fpm_dist_params* p, const fpm_panel* ref, const fpm_panel* qry, fpm_pair* out);
int fpm_dist_hits_multi(fpm_multi* m, const fpm_dist_params* p, const fpm_panel* ref, const fpm_panel* qry, fpm_hit* out,
                        uint64_t capacity, uint64_t* n_hits);
/* fpm_sketch_batch with the sketches spread over the GPUs (whole sketches, contiguous ranges balanced by bytes).      */
int fpm_sketch_batch_multi(fpm_multi* m, const fpm_sketch_params* p, const uint8_t* seq, uint64_t seq_bytes,
                           const uint64_t* group_offsets, uint32_t n_groups, uint64_t* out_hashes, uint32_t* out_counts,
                           uint32_t* out_n, uint64_t* out_kmers);

/* Replaces the fork's compareFingerprints (CommandTriangle.cpp:265-302, `mash triangle -fp`): positional
 * matches over the first min(|ref|,|qry|) hashes, distance = 1 - matches/min, p = chi-square(1 dof) upper
 * tail at `matches`.  Only max_distance / max_pvalue of *p are used.  Same output layout as fpm_dist_tile. */
int fpm_fp_positional_tile(fpm_ctx* ctx, const fpm_dist_params* p, const fpm_panel* ref, const fpm_panel* qry, fpm_pair* out);

/* Which tile kernel fpm_dist_tile[_dev] uses for ascending duplicate-free panels.
 * FPM_DIST_AUTO (default): 32-bit dense-rank kernel whenever both panels together hold fewer than 2^31
 * hashes, else the 64-bit kernel.  FPM_DIST_FORCE64: always the 64-bit kernel (tests compare the two). */
#define FPM_DIST_AUTO 0
#define FPM_DIST_FORCE64 1
#define FPM_DIST_NO_PRUNE 2   /* 32-bit rank kernel, but merge every pair (no disjoint-pair pruning)           */
#define FPM_DIST_NO_GROUP 3   /* prune, but do not reorder the panels so that related sketches share tiles     */
#define FPM_DIST_SATURATE 4   /* prune, and always bound the marking walks by component sizes (default: only when it pays) */
#define FPM_DIST_GROUP 5      /* prune, and always reorder the panels (default: only when the marked pairs are scattered over the tiles) */
int fpm_ctx_set_dist_mode(fpm_ctx* ctx, int mode);

/* Scalar helpers exported for host code and tests (same code the kernels run).             */
double fpm_pvalue(uint64_t x, uint64_t len_ref, uint64_t len_qry, double kmer_space, uint64_t n);
double fpm_distance(uint64_t common, uint64_t denom, int kmer_size);

/* ---- measurement helpers ------------------------------------------------------------- */

/* Per-kernel device timing (CUDA events recorded on the context's stream around each launch of
 * the named kernel).  Off by default; enabling resets the accumulators.                        */
#define FPM_KERNEL_SKETCH_HASH 0   /* sketch_hash_kernel (pack + canonical roll + Murmur + filter) */
#define FPM_KERNEL_SKETCH_SELECT 1 /* sketch_select_kernel                                         */
#define FPM_KERNEL_DIST_TILE 2     /* dist_tile32_kernel / dist_tile_kernel                        */
#define FPM_KERNEL_DIST_LITERAL 3  /* dist_literal_kernel                                          */
#define FPM_KERNEL_DIST_PACK 4     /* dist_pack_kernel, or the rank pre-pass (keys, sort, scan, scatter) */
#define FPM_KERNEL_DIST_EXCHANGE 5 /* fpm_dist_sharded_dev: the grouped ncclSend / ncclRecv of the row shards      */
#define FPM_KERNEL_GUNZIP 6        /* gunzip_kernel (fpm_gunzip_batch)                             */
int fpm_ctx_set_timing(fpm_ctx* ctx, int enable);
int fpm_ctx_get_timing(fpm_ctx* ctx, int kernel_id, double* out_ms_total, uint64_t* out_launches);

/* Integer-pipe microbenchmark: sustained 32-bit integer ops/s of this GPU (a dependent
 * IMAD+LOP3+SHF mix on every SM), the denominator of the sketch kernel's integer roofline
 * (INT32 peak is not in MEASURED_PEAKS.json and must be measured).           */
int fpm_measure_int32_peak(fpm_ctx* ctx, double* out_ops_per_s);
/* The three figures behind the last fpm_measure_int32_peak: [0] ALU pipe only (LOP3), [1] FMA pipe only
 * (IMAD), [2] alternating; the value returned above is their maximum.                              */
int fpm_get_int32_peaks(fpm_ctx* ctx, double* out3);

#ifdef __cplusplus
}
#endif
#endif /* FPMASH_B200_H */
