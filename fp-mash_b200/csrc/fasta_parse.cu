// fasta_parse.cu -- FASTA ingestion on the GPU (first half): raw file bytes in, the sketch
// kernel's batch layout out (records back to back, each followed by one 0x00), plus a record table.
//
// Replaces, for plain FASTA, the byte loop of the reference's kseq.h reader (2009 version, kseq.h:170-208) as
// mirrored by host/fastx.cpp:
//   * a record starts at the next '>'; its header runs to the next '\n';
//   * the sequence is every isgraph() byte (33..126) up to the next '>' -- ANYWHERE, not only at a line start --
//     everything else ('\n', '\r', blanks) is dropped; bytes before the first '>' of a file are skipped.
// '+' and '@' also end a record there (FASTQ); this parser does not follow them: it reports FPM_FASTA_NOT_PLAIN
// and the caller uses the host reader for that input.  Several files go into one call separated by a 0x00 byte
// (which resets the state machine; a file that itself contains 0x00 must take the host reader).
//
// The reader is a three-state automaton (PRE: before the first header of a file, SEQ, HDR) whose transitions
// are functions on three states, so "state before byte i" is a prefix scan under function composition:
//   fasta_chunk_fn_kernel   one warp per 4 KB chunk: the chunk's transition function from its last reset / '>' / '\n'
//   fasta_scan_kernel       one CTA: exclusive scan of the chunk functions (-> state at every chunk start), and
//                           later of the per-chunk output / record counts
//   fasta_emit_kernel<0|1>  one warp per chunk, 128 bytes per step: a lane composes its 4 bytes, a warp scan gives
//                           every lane its start state; pass 0 counts kept bytes and records, pass 1 writes them
#include "common.h"

namespace fpm {

constexpr int FA_CHUNK = 4096;
constexpr uint32_t ST_PRE = 0, ST_SEQ = 1, ST_HDR = 2;
constexpr uint32_t FN_ID = 0u | (1u << 2) | (2u << 4);         // f(s) in bits [2s+1 : 2s]

__device__ __forceinline__ uint32_t fn_apply(uint32_t f, uint32_t s) { return (f >> (2 * s)) & 3u; }
// first f, then g
__device__ __forceinline__ uint32_t fn_then(uint32_t f, uint32_t g)
{
    return fn_apply(g, f & 3u) | (fn_apply(g, (f >> 2) & 3u) << 2) | (fn_apply(g, (f >> 4) & 3u) << 4);
}
__device__ __forceinline__ uint32_t fn_of_byte(uint32_t b)
{
    if (b == 0) return 0u;                                        // reset: everything -> PRE
    if (b == '>') return ST_HDR | (ST_HDR << 2) | (ST_HDR << 4);
    if (b == '\n') return ST_PRE | (ST_SEQ << 2) | (ST_SEQ << 4);
    return FN_ID;
}
__device__ __forceinline__ uint32_t step_state(uint32_t s, uint32_t b)
{
    if (b == 0) return ST_PRE;
    if (b == '>') return ST_HDR;
    if (b == '\n' && s == ST_HDR) return ST_SEQ;
    return s;
}

// 4 bytes at raw[pos .. pos+4) as a little-endian word; bytes at or beyond n read as ' ' (dropped, no state change)
__device__ __forceinline__ uint32_t load4(const uint8_t* __restrict__ raw, uint64_t pos, uint64_t n)
{
    if (pos + 4 <= n) return *reinterpret_cast<const uint32_t*>(raw + pos);
    uint32_t w = 0x20202020u;
    for (int j = 0; j < 4; j++)
        if (pos + j < n) w = (w & ~(0xffu << (8 * j))) | ((uint32_t)raw[pos + j] << (8 * j));
    return w;
}

__global__ void __launch_bounds__(256) fasta_chunk_fn_kernel(const uint8_t* __restrict__ raw, uint64_t n, uint64_t n_chunks, uint8_t* __restrict__ chunk_fn)
{
    const uint64_t chunk = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (chunk >= n_chunks) return;
    const uint64_t base = chunk * FA_CHUNK;
    int lastR = -1, lastG = -1, lastN = -1;                       // chunk-relative positions of the last reset / '>' / '\n'
    for (int off = 4 * lane; off < FA_CHUNK; off += 128) {
        if (base + off >= n) break;
        const uint32_t w = load4(raw, base + off, n);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t b = (w >> (8 * j)) & 0xffu;
            if (b == 0) lastR = off + j;
            else if (b == '>') lastG = off + j;
            else if (b == '\n') lastN = off + j;
        }
    }
    for (int o = 16; o; o >>= 1) {
        lastR = max(lastR, __shfl_xor_sync(0xffffffffu, lastR, o));
        lastG = max(lastG, __shfl_xor_sync(0xffffffffu, lastG, o));
        lastN = max(lastN, __shfl_xor_sync(0xffffffffu, lastN, o));
    }
    if (lane == 0) {
        uint32_t f = 0;
        for (uint32_t s = 0; s < 3; s++) {
            uint32_t t = lastR >= 0 ? ST_PRE : s;
            if (lastG > lastR) t = lastN > lastG ? ST_SEQ : ST_HDR;
            else if (t == ST_HDR && lastN > lastR) t = ST_SEQ;
            f |= t << (2 * s);
        }
        chunk_fn[chunk] = (uint8_t)f;
    }
}

// One CTA.  mode 0: chunk_fn -> start_state (exclusive scan under composition, applied to PRE).
//           mode 1: counts (bytes | records) -> exclusive prefix sums, totals in totals[0..1].
__global__ void __launch_bounds__(1024) fasta_scan_kernel(int mode, uint64_t n_chunks, const uint8_t* __restrict__ chunk_fn, uint8_t* __restrict__ start_state,
                                                           const uint32_t* __restrict__ cnt_bytes, const uint32_t* __restrict__ cnt_recs,
                                                           uint64_t* __restrict__ off_bytes, uint64_t* __restrict__ off_recs, uint64_t* __restrict__ totals)
{
    __shared__ uint32_t s_fn[32];
    __shared__ uint64_t s_a[32], s_b[32];
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    uint32_t carry_fn = FN_ID;
    uint64_t carry_a = 0, carry_b = 0;
    for (uint64_t tile = 0; tile < n_chunks; tile += 1024) {
        const uint64_t c = tile + t;
        if (mode == 0) {
            uint32_t f = c < n_chunks ? chunk_fn[c] : FN_ID;
            uint32_t inc = f;                                      // inclusive scan within the warp
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc = fn_then(up, inc);
            }
            if (lane == 31) s_fn[wid] = inc;
            __syncthreads();
            if (wid == 0) {
                uint32_t w = s_fn[lane], winc = w;
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t up = __shfl_up_sync(0xffffffffu, winc, o);
                    if (lane >= o) winc = fn_then(up, winc);
                }
                s_fn[lane] = winc;                                  // inclusive over warps
            }
            __syncthreads();
            uint32_t excl = __shfl_up_sync(0xffffffffu, inc, 1);    // exclusive within the warp
            if (lane == 0) excl = FN_ID;
            uint32_t before = wid ? fn_then(s_fn[wid - 1], excl) : excl;
            before = fn_then(carry_fn, before);
            if (c < n_chunks) start_state[c] = (uint8_t)fn_apply(before, ST_PRE);
            carry_fn = fn_then(carry_fn, s_fn[31]);
            __syncthreads();
        } else {
            uint64_t a = c < n_chunks ? cnt_bytes[c] : 0, b = c < n_chunks ? cnt_recs[c] : 0;
            uint64_t ia = a, ib = b;
            for (int o = 1; o < 32; o <<= 1) {
                const uint64_t ua = __shfl_up_sync(0xffffffffu, ia, o), ub = __shfl_up_sync(0xffffffffu, ib, o);
                if (lane >= o) { ia += ua; ib += ub; }
            }
            if (lane == 31) { s_a[wid] = ia; s_b[wid] = ib; }
            __syncthreads();
            if (wid == 0) {
                uint64_t wa = s_a[lane], wb = s_b[lane];
                for (int o = 1; o < 32; o <<= 1) {
                    const uint64_t ua = __shfl_up_sync(0xffffffffu, wa, o), ub = __shfl_up_sync(0xffffffffu, wb, o);
                    if (lane >= o) { wa += ua; wb += ub; }
                }
                s_a[lane] = wa; s_b[lane] = wb;
            }
            __syncthreads();
            const uint64_t pa = carry_a + (wid ? s_a[wid - 1] : 0) + ia - a, pb = carry_b + (wid ? s_b[wid - 1] : 0) + ib - b;
            if (c < n_chunks) { off_bytes[c] = pa; off_recs[c] = pb; }
            carry_a += s_a[31]; carry_b += s_b[31];
            __syncthreads();
        }
    }
    if (mode == 1 && t == 0) { totals[0] = carry_a; totals[1] = carry_b; }
}

// WRITE = false: count kept bytes (sequence + separators) and record starts per chunk.
// WRITE = true:  write them at the scanned offsets and fill the record table.
template <bool WRITE>
__global__ void __launch_bounds__(256) fasta_emit_kernel(const uint8_t* __restrict__ raw, uint64_t n, uint64_t n_chunks, const uint8_t* __restrict__ start_state,
                                                         uint32_t* __restrict__ cnt_bytes, uint32_t* __restrict__ cnt_recs,
                                                         const uint64_t* __restrict__ off_bytes, const uint64_t* __restrict__ off_recs,
                                                         uint8_t* __restrict__ out, fpm_fasta_record* __restrict__ recs, uint32_t* __restrict__ flags)
{
    const uint64_t chunk = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (chunk >= n_chunks) return;
    const uint64_t base = chunk * FA_CHUNK;
    uint32_t state = start_state[chunk];                          // warp-uniform: state before the current 128-byte slice
    uint64_t out_pos = WRITE ? off_bytes[chunk] : 0, rec_idx = WRITE ? off_recs[chunk] : 0;
    uint32_t tot_bytes = 0, tot_recs = 0, bad = 0;
    for (int off = 0; off < FA_CHUNK; off += 128) {
        if (base + off >= n) break;
        const uint64_t pos = base + off + 4 * lane;
        const uint32_t w = pos < n ? load4(raw, pos, n) : 0x20202020u;
        // my four bytes as a transition function, then the state before my first byte
        uint32_t f = FN_ID;
#pragma unroll
        for (int j = 0; j < 4; j++) f = fn_then(f, fn_of_byte((w >> (8 * j)) & 0xffu));
        uint32_t inc = f;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc = fn_then(up, inc);
        }
        uint32_t excl = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) excl = FN_ID;
        uint32_t s = fn_apply(excl, state);
        const uint32_t slice_fn = __shfl_sync(0xffffffffu, inc, 31);
        // walk my four bytes: what do they emit?
        uint32_t nb = 0, nr = 0;
        uint8_t ob[8];                                              // at most a separator + a byte per input byte
        uint32_t rec_at[4], rec_sep[4], hdr_end_at[4];
        uint32_t n_rec_local = 0, n_hend = 0;
        uint32_t hend_before[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t b = (w >> (8 * j)) & 0xffu;
            if (pos + j < n) {
                if (b == '>' && s != ST_HDR) {                      // a record starts; the previous one (if any) gets its separator
                    const uint32_t sep = s == ST_SEQ;
                    if (sep) ob[nb++] = 0;
                    rec_at[n_rec_local] = j; rec_sep[n_rec_local] = nb; n_rec_local++;
                    nr++;
                } else if (b == 0) {                                // end of a file
                    if (s == ST_HDR) { hdr_end_at[n_hend] = j; hend_before[n_hend] = nr; n_hend++; }
                    if (s != ST_PRE) ob[nb++] = 0;
                } else if (s == ST_HDR) {
                    if (b == '\n') { hdr_end_at[n_hend] = j; hend_before[n_hend] = nr; n_hend++; }
                } else if (b == '+' || b == '@') {
                    if (s == ST_SEQ || b == '@') bad = 1;           // FASTQ: not this parser's job
                } else if (s == ST_SEQ && b >= 33 && b <= 126) {
                    ob[nb++] = (uint8_t)b;
                }
            }
            s = step_state(s, b);
        }
        // exclusive prefix of (bytes, records) over the lanes
        uint32_t packed = nb | (nr << 16), pinc = packed;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, pinc, o);
            if (lane >= o) pinc += up;
        }
        const uint32_t pex = pinc - packed, slice_tot = __shfl_sync(0xffffffffu, pinc, 31);
        if (WRITE) {
            const uint64_t my_out = out_pos + (pex & 0xffffu), my_rec = rec_idx + (pex >> 16);
            for (uint32_t i = 0; i < nb; i++) out[my_out + i] = ob[i];
            for (uint32_t r = 0; r < n_rec_local; r++) {
                recs[my_rec + r].hdr_begin = pos + rec_at[r];
                recs[my_rec + r].seq_begin = my_out + rec_sep[r];
            }
            for (uint32_t h = 0; h < n_hend; h++) recs[my_rec + hend_before[h] - 1].hdr_end = pos + hdr_end_at[h];
        }
        out_pos += slice_tot & 0xffffu;
        rec_idx += slice_tot >> 16;
        tot_bytes += slice_tot & 0xffffu;
        tot_recs += slice_tot >> 16;
        state = fn_apply(slice_fn, state);
    }
    if (!WRITE) {
        bad = __any_sync(0xffffffffu, bad);
        if (lane == 0) {
            cnt_bytes[chunk] = tot_bytes;
            cnt_recs[chunk] = tot_recs;
            if (bad) atomicOr(flags, 1u);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// FASTQ, four lines per record ("clean" FASTQ: what sequencers write).  For such input the reference's reader
// (kseq.h:170-208) reduces to: line 4r is the header, line 4r+1 the sequence, line 4r+2 starts with '+', line
// 4r+3 holds as many quality bytes as there are bases.  The role of a byte is then its line number mod 4, and the
// line number is a prefix sum of newlines.  Everything the general reader would treat differently is detected and
// reported as FPM_FASTA_NOT_PLAIN (the caller restarts with the host reader): a header not starting with '@', a
// third line not starting with '+', a sequence byte outside 33..126 or one of '>' '+' '@', a quality byte outside
// 33..127, 0x00 anywhere, a '\r' in a sequence or quality line that is not the last byte before the '\n' (CRLF files: the reader
// drops that '\r' from sequence and quality and keeps it in the header's comment, and so does this parser), sequence and quality
// lines of different lengths, a line count not divisible by four.
// The buffer must start at a record boundary and end with '\n'.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fastq_newline_kernel(const uint8_t* __restrict__ raw, uint64_t n, uint64_t n_chunks, uint32_t* __restrict__ cnt_nl,
                                                            uint32_t* __restrict__ flags)
{
    const uint64_t chunk = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (chunk >= n_chunks) return;
    const uint64_t base = chunk * FA_CHUNK;
    uint32_t c = 0, bad = 0;
    for (int off = 4 * lane; off < FA_CHUNK; off += 128) {
        if (base + off >= n) break;
        const uint32_t w = load4(raw, base + off, n);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t b = (w >> (8 * j)) & 0xffu;
            c += b == '\n';
            bad |= (b == 0);
        }
    }
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        cnt_nl[chunk] = c;
        if (bad) atomicOr(flags, 1u);
    }
}

template <bool WRITE>
__global__ void __launch_bounds__(256) fastq_emit_kernel(const uint8_t* __restrict__ raw, uint64_t n, uint64_t n_chunks, const uint64_t* __restrict__ line_base,
                                                         uint32_t* __restrict__ cnt_bytes, const uint64_t* __restrict__ off_bytes, uint8_t* __restrict__ out,
                                                         uint64_t* __restrict__ nlpos, uint32_t* __restrict__ flags)
{
    const uint64_t chunk = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (chunk >= n_chunks) return;
    const uint64_t base = chunk * FA_CHUNK;
    uint64_t line = line_base[chunk];                              // warp-uniform: lines completed before the current slice
    uint64_t out_pos = WRITE ? off_bytes[chunk] : 0;
    uint32_t prev_last = base ? raw[base - 1] : '\n';             // the byte before the current slice (line-start test)
    uint32_t tot = 0, bad = 0;
    for (int off = 0; off < FA_CHUNK; off += 128) {
        if (base + off >= n) break;
        const uint64_t pos = base + off + 4 * lane;
        const uint32_t w = pos < n ? load4(raw, pos, n) : 0x20202020u;
        uint32_t nl = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) nl += (pos + j < n) && ((w >> (8 * j)) & 0xffu) == '\n';
        uint32_t inc = nl;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += up;
        }
        uint64_t my_line = line + (inc - nl);
        uint32_t before = __shfl_up_sync(0xffffffffu, w >> 24, 1);  // last byte of the previous lane
        if (lane == 0) before = prev_last;
        uint32_t after = __shfl_down_sync(0xffffffffu, w & 0xffu, 1);   // first byte of the next lane
        if (lane == 31) after = pos + 4 < n ? raw[pos + 4] : '\n';
        uint8_t ob[4];
        uint32_t nb = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t b = (w >> (8 * j)) & 0xffu;
            const uint32_t next = j < 3 ? (w >> (8 * (j + 1))) & 0xffu : after;
            if (pos + j < n) {
                const uint32_t role = (uint32_t)my_line & 3u;
                const bool first = before == '\n';
                if (b == '\r' && (role == 1 || role == 3)) {
                    if (!(pos + j + 1 < n && next == '\n')) bad = 1;     // CRLF only: kseq drops the '\r' of a sequence / quality line
                } else if (b == '\n') {
                    if (role == 1) ob[nb++] = 0;                    // end of a read
                    if (first && role != 1 && role != 3) bad = 1;   // empty header or '+' line
                    if (WRITE) nlpos[my_line] = pos + j;
                    my_line++;
                } else if (role == 0) {
                    if (first && b != '@') bad = 1;
                } else if (role == 1) {
                    if (b < 33 || b > 126 || b == '>' || b == '+' || b == '@') bad = 1;
                    ob[nb++] = (uint8_t)b;
                } else if (role == 2) {
                    if (first && b != '+') bad = 1;
                } else {
                    if (b < 33 || b > 127) bad = 1;
                }
            }
            before = b;
        }
        uint32_t pinc = nb;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, pinc, o);
            if (lane >= o) pinc += up;
        }
        const uint32_t slice_tot = __shfl_sync(0xffffffffu, pinc, 31);
        if (WRITE) {
            const uint64_t my_out = out_pos + (pinc - nb);
            for (uint32_t i = 0; i < nb; i++) out[my_out + i] = ob[i];
        }
        out_pos += slice_tot;
        tot += slice_tot;
        line += __shfl_sync(0xffffffffu, inc, 31);
        prev_last = __shfl_sync(0xffffffffu, w >> 24, 31);
    }
    if (!WRITE) {
        bad = __any_sync(0xffffffffu, bad);
        if (lane == 0) {
            cnt_bytes[chunk] = tot;
            if (bad) atomicOr(flags, 1u);
        }
    }
}

// per record: sequence and quality lengths must agree; count the reads of at least min_len bases and find the first one
__global__ void __launch_bounds__(256) fastq_check_kernel(const uint8_t* __restrict__ raw, const uint64_t* __restrict__ nlpos, uint64_t n_reads, uint32_t min_len,
                                                          uint32_t* __restrict__ flags, unsigned long long* __restrict__ n_valid, unsigned long long* __restrict__ first_valid)
{
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = false, bad = false;
    if (r < n_reads) {
        // bases / quality bytes of the record: the line without its '\r' (CRLF files)
        uint64_t len1 = nlpos[4 * r + 1] - nlpos[4 * r] - 1, len3 = nlpos[4 * r + 3] - nlpos[4 * r + 2] - 1;
        if (len1 && raw[nlpos[4 * r + 1] - 1] == '\r') len1--;
        if (len3 && raw[nlpos[4 * r + 3] - 1] == '\r') len3--;
        bad = len1 != len3;
        valid = len1 >= min_len;
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flags, 1u);
    const uint32_t m = __ballot_sync(0xffffffffu, valid);
    if (m && (threadIdx.x & 31) == 0) {
        atomicAdd(n_valid, (unsigned long long)__popc(m));
        atomicMin(first_valid, (unsigned long long)(r + (__ffs(m) - 1)));
    }
}


// header text of every record -> out[off[i] .. off[i+1]) (at most hdr_end - hdr_begin bytes), one warp per record
__global__ void __launch_bounds__(256) fasta_headers_kernel(const uint8_t* __restrict__ raw, uint64_t n_bytes, const fpm_fasta_record* __restrict__ recs, uint64_t n,
                                                            const uint64_t* __restrict__ off, uint8_t* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    for (uint64_t r = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += (uint64_t)gridDim.x * 8) {
        const uint64_t b = recs[r].hdr_begin, e = recs[r].hdr_end < n_bytes ? recs[r].hdr_end : n_bytes;
        const uint64_t len = e > b ? e - b : 0, room = off[r + 1] - off[r];
        for (uint64_t j = lane; j < len && j < room; j += 32) out[off[r] + j] = raw[b + j];
    }
}

}  // namespace fpm

using namespace fpm;

extern "C" {

int fpm_fasta_parse(fpm_ctx* ctx, const uint8_t* raw, uint64_t n_bytes, uint64_t* out_n_records, uint64_t* out_seq_bytes, int* out_status)
{
    if (!ctx) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    if (!out_n_records || !out_seq_bytes || !out_status) { set_error("NULL output"); return FPM_ERR_ARG; }
    // raw == NULL: the batch fpm_gunzip_batch inflated into fa_raw
    if (!raw && n_bytes && ctx->fa_resident != n_bytes) { set_error("raw is NULL and no resident batch of %llu bytes", (unsigned long long)n_bytes); return FPM_ERR_ARG; }
    *out_n_records = 0; *out_seq_bytes = 0; *out_status = FPM_FASTA_OK;
    ctx->fa_records = 0; ctx->fa_seq_bytes = 0;
    if (n_bytes == 0) { ctx->fa_resident = 0; return FPM_OK; }
    FPM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint64_t n_chunks = (n_bytes + FA_CHUNK - 1) / FA_CHUNK;
    if (n_chunks * 32 > 0x7fffffffull * 256) { set_error("FASTA batch too large"); return FPM_ERR_ARG; }
    int rc;
    if ((rc = ctx->fa_raw.ensure(n_bytes + 64))) return rc;
    if ((rc = ctx->fa_seq.ensure(n_bytes + 64))) return rc;          // kept bytes + separators never exceed the input
    if ((rc = ctx->fa_chunk.ensure(n_chunks * (1 + 1 + 4 + 4 + 8 + 8) + 256))) return rc;
    if ((rc = ctx->d_misc.ensure(64))) return rc;
    unsigned char* cb = ctx->fa_chunk.as<unsigned char>();
    uint64_t* off_bytes = (uint64_t*)cb;                             // 8-byte fields first (alignment)
    uint64_t* off_recs = off_bytes + n_chunks;
    uint32_t* cnt_bytes = (uint32_t*)(off_recs + n_chunks);
    uint32_t* cnt_recs = cnt_bytes + n_chunks;
    uint8_t* chunk_fn = (uint8_t*)(cnt_recs + n_chunks);
    uint8_t* start_state = chunk_fn + n_chunks;
    uint64_t* totals = ctx->d_misc.as<uint64_t>() + 2;               // d_misc: [flags | pad | totals[2]]
    FPM_CUDA(cudaMemsetAsync(ctx->d_misc.p, 0, 64, st));
    if (raw) FPM_CUDA(cudaMemcpyAsync(ctx->fa_raw.p, raw, n_bytes, cudaMemcpyHostToDevice, st));
    ctx->fa_resident = n_bytes;
    const uint32_t grid = (uint32_t)((n_chunks * 32 + 255) / 256);
    const uint8_t* d_raw = ctx->fa_raw.as<uint8_t>();
    fasta_chunk_fn_kernel<<<grid, 256, 0, st>>>(d_raw, n_bytes, n_chunks, chunk_fn);
    fasta_scan_kernel<<<1, 1024, 0, st>>>(0, n_chunks, chunk_fn, start_state, nullptr, nullptr, nullptr, nullptr, nullptr);
    fasta_emit_kernel<false><<<grid, 256, 0, st>>>(d_raw, n_bytes, n_chunks, start_state, cnt_bytes, cnt_recs, nullptr, nullptr, nullptr, nullptr,
                                                   ctx->d_misc.as<uint32_t>());
    fasta_scan_kernel<<<1, 1024, 0, st>>>(1, n_chunks, nullptr, nullptr, cnt_bytes, cnt_recs, off_bytes, off_recs, totals);
    ctx->launches += 4;
    FPM_CUDA(cudaGetLastError());
    uint64_t h[4] = {0, 0, 0, 0};
    FPM_CUDA(cudaMemcpyAsync(h, ctx->d_misc.p, 32, cudaMemcpyDeviceToHost, st));
    FPM_CUDA(cudaStreamSynchronize(st));
    if ((uint32_t)h[0] & 1u) { *out_status = FPM_FASTA_NOT_PLAIN; return FPM_OK; }
    const uint64_t seq_bytes = h[2], n_rec = h[3];
    if ((rc = ctx->fa_recs.ensure(sizeof(fpm_fasta_record) * std::max<uint64_t>(n_rec, 1)))) return rc;
    fasta_emit_kernel<true><<<grid, 256, 0, st>>>(d_raw, n_bytes, n_chunks, start_state, nullptr, nullptr, off_bytes, off_recs, ctx->fa_seq.as<uint8_t>(),
                                                  ctx->fa_recs.as<fpm_fasta_record>(), nullptr);
    ctx->launches++;
    FPM_CUDA(cudaGetLastError());
    ctx->fa_records = n_rec;
    ctx->fa_seq_bytes = seq_bytes;
    *out_n_records = n_rec;
    *out_seq_bytes = seq_bytes;
    return FPM_OK;
}

int fpm_fasta_headers(fpm_ctx* ctx, const uint64_t* offsets, uint8_t* out)
{
    if (!ctx || !offsets) { set_error("NULL argument"); return FPM_ERR_ARG; }
    const uint64_t n = ctx->fa_records;
    if (n == 0) return FPM_OK;
    const uint64_t total = offsets[n] - offsets[0];
    if (offsets[0] != 0) { set_error("offsets must start at 0"); return FPM_ERR_ARG; }
    if (total && !out) { set_error("out is NULL"); return FPM_ERR_ARG; }
    FPM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int rc;
    if ((rc = ctx->fa_chunk.ensure((n + 1) * 8 + total + 64))) return rc;
    uint64_t* d_off = ctx->fa_chunk.as<uint64_t>();
    uint8_t* d_out = (uint8_t*)(d_off + n + 1);
    FPM_CUDA(cudaMemcpyAsync(d_off, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    fasta_headers_kernel<<<(uint32_t)std::min<uint64_t>((n + 7) / 8, 65535), 256, 0, st>>>(ctx->fa_raw.as<uint8_t>(), ctx->fa_resident, ctx->fa_recs.as<fpm_fasta_record>(), n, d_off,
                                                                                          d_out);
    ctx->launches++;
    FPM_CUDA(cudaGetLastError());
    if (total) FPM_CUDA(cudaMemcpyAsync(out, d_out, total, cudaMemcpyDeviceToHost, st));
    FPM_CUDA(cudaStreamSynchronize(st));
    return FPM_OK;
}

int fpm_fasta_records(fpm_ctx* ctx, fpm_fasta_record* out)
{
    if (!ctx || !out) { set_error("NULL argument"); return FPM_ERR_ARG; }
    FPM_CUDA(cudaSetDevice(ctx->device));
    if (ctx->fa_records) FPM_CUDA(cudaMemcpyAsync(out, ctx->fa_recs.p, sizeof(fpm_fasta_record) * ctx->fa_records, cudaMemcpyDeviceToHost, ctx->stream));
    FPM_CUDA(cudaStreamSynchronize(ctx->stream));
    return FPM_OK;
}

int fpm_fasta_sequence(fpm_ctx* ctx, uint8_t* out)
{
    if (!ctx || !out) { set_error("NULL argument"); return FPM_ERR_ARG; }
    FPM_CUDA(cudaSetDevice(ctx->device));
    if (ctx->fa_seq_bytes) FPM_CUDA(cudaMemcpyAsync(out, ctx->fa_seq.p, ctx->fa_seq_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    FPM_CUDA(cudaStreamSynchronize(ctx->stream));
    return FPM_OK;
}

int fpm_fastq_stream_append(fpm_ctx* ctx, const uint8_t* raw, uint64_t n_bytes, uint32_t min_len, int* out_status, uint64_t* out_info)
{
    if (!ctx) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    if (!out_status || !out_info) { set_error("NULL output"); return FPM_ERR_ARG; }
    if (ctx->stream_goff.empty()) { set_error("fpm_sketch_stream_begin was not called"); return FPM_ERR_ARG; }
    *out_status = FPM_FASTA_OK;
    for (int i = 0; i < FPM_FASTQ_INFO_WORDS; i++) out_info[i] = 0;
    if (n_bytes == 0) return FPM_OK;
    if (!raw || raw[n_bytes - 1] != '\n') { set_error("a FASTQ piece must end with a newline"); return FPM_ERR_ARG; }
    FPM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint64_t n_chunks = (n_bytes + FA_CHUNK - 1) / FA_CHUNK;
    int rc;
    if ((rc = ctx->fa_raw.ensure(n_bytes + 64))) return rc;
    if ((rc = ctx->fa_chunk.ensure(n_chunks * (8 + 8 + 8 + 4 + 4) + 256))) return rc;
    if ((rc = ctx->d_misc.ensure(64))) return rc;
    if ((rc = stream_reserve(ctx, n_bytes))) return rc;                 // kept bytes + separators never exceed the input
    uint64_t* line_base = ctx->fa_chunk.as<uint64_t>();
    uint64_t* out_off = line_base + n_chunks;
    uint64_t* scratch = out_off + n_chunks;                             // the scan kernel's second output (unused here)
    uint32_t* cnt_nl = (uint32_t*)(scratch + n_chunks);
    uint32_t* cnt_bytes = cnt_nl + n_chunks;
    uint32_t* d_flags = ctx->d_misc.as<uint32_t>();
    uint64_t* totals = ctx->d_misc.as<uint64_t>() + 2;                  // d_misc (u64 words): flags, pad, totals[2], n_valid, first_valid
    unsigned long long* d_valid = (unsigned long long*)(ctx->d_misc.as<uint64_t>() + 4);
    unsigned long long* d_first = d_valid + 1;
    FPM_CUDA(cudaMemsetAsync(ctx->d_misc.p, 0, 40, st));
    FPM_CUDA(cudaMemsetAsync(d_first, 0xff, 8, st));
    FPM_CUDA(cudaMemcpyAsync(ctx->fa_raw.p, raw, n_bytes, cudaMemcpyHostToDevice, st));
    const uint32_t grid = (uint32_t)((n_chunks * 32 + 255) / 256);
    const uint8_t* d_raw = ctx->fa_raw.as<uint8_t>();
    fastq_newline_kernel<<<grid, 256, 0, st>>>(d_raw, n_bytes, n_chunks, cnt_nl, d_flags);
    fasta_scan_kernel<<<1, 1024, 0, st>>>(1, n_chunks, nullptr, nullptr, cnt_nl, cnt_nl, line_base, scratch, totals);
    fastq_emit_kernel<false><<<grid, 256, 0, st>>>(d_raw, n_bytes, n_chunks, line_base, cnt_bytes, nullptr, nullptr, nullptr, d_flags);
    ctx->launches += 3;
    FPM_CUDA(cudaGetLastError());
    uint64_t h[6];
    FPM_CUDA(cudaMemcpyAsync(h, ctx->d_misc.p, 32, cudaMemcpyDeviceToHost, st));
    FPM_CUDA(cudaStreamSynchronize(st));
    const uint64_t n_lines = h[2];
    if (((uint32_t)h[0] & 1u) || (n_lines & 3)) { *out_status = FPM_FASTA_NOT_PLAIN; return FPM_OK; }
    const uint64_t n_reads = n_lines / 4;
    if ((rc = ctx->fa_recs.ensure(sizeof(uint64_t) * std::max<uint64_t>(n_lines, 1)))) return rc;
    uint64_t* nlpos = ctx->fa_recs.as<uint64_t>();
    uint8_t* dst = (uint8_t*)ctx->stream_buf.p + ctx->stream_used;
    fasta_scan_kernel<<<1, 1024, 0, st>>>(1, n_chunks, nullptr, nullptr, cnt_bytes, cnt_bytes, out_off, scratch, totals);
    fastq_emit_kernel<true><<<grid, 256, 0, st>>>(d_raw, n_bytes, n_chunks, line_base, nullptr, out_off, dst, nlpos, nullptr);
    if (n_reads) fastq_check_kernel<<<(uint32_t)((n_reads + 255) / 256), 256, 0, st>>>(d_raw, nlpos, n_reads, min_len, d_flags, d_valid, d_first);
    ctx->launches += 3;
    FPM_CUDA(cudaGetLastError());
    FPM_CUDA(cudaMemcpyAsync(h, ctx->d_misc.p, 48, cudaMemcpyDeviceToHost, st));
    FPM_CUDA(cudaStreamSynchronize(st));
    if ((uint32_t)h[0] & 1u) { *out_status = FPM_FASTA_NOT_PLAIN; return FPM_OK; }      // nothing was committed to the stream
    const uint64_t seq_bytes = h[2], n_valid = h[4], first_valid = h[5];
    out_info[0] = n_reads;
    out_info[1] = n_valid;
    out_info[2] = seq_bytes;
    out_info[3] = n_valid ? first_valid : n_reads;                       // index of the first read of at least min_len bases
    ctx->fa_records = n_lines;                                           // fpm_fastq_line_ends serves the newline table
    ctx->stream_used += seq_bytes;
    return FPM_OK;
}

int fpm_fastq_line_ends(fpm_ctx* ctx, uint64_t first_line, uint64_t n_lines, uint64_t* out)
{
    if (!ctx || !out) { set_error("NULL argument"); return FPM_ERR_ARG; }
    if (first_line + n_lines > ctx->fa_records) { set_error("line range outside the last FASTQ piece"); return FPM_ERR_ARG; }
    FPM_CUDA(cudaSetDevice(ctx->device));
    if (n_lines) FPM_CUDA(cudaMemcpyAsync(out, ctx->fa_recs.as<uint64_t>() + first_line, sizeof(uint64_t) * n_lines, cudaMemcpyDeviceToHost, ctx->stream));
    FPM_CUDA(cudaStreamSynchronize(ctx->stream));
    return FPM_OK;
}

}  // extern "C"
