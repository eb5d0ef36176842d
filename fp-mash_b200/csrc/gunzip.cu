// gunzip.cu -- gzip'ed FASTA files inflated on the GPU ("GPU FASTA/FASTQ(.gz) ingestion").
//
// The reference opens every input through zlib (gzopen / gzread behind kseq, Sketch.cpp:1340-1346, kseq.h:60-75); a
// collection of genomes normally arrives as one .gz per genome, and then the inflate on the host cores -- not the
// sketch kernel, not PCIe -- is what a `mash sketch` run waits for.  Here the COMPRESSED bytes go over PCIe and
// one warp per file inflates them in HBM, straight into the raw-batch layout fpm_fasta_parse takes (every file followed
// by one 0x00), so the FASTA parser and the sketch kernel follow without the decompressed bytes ever visiting the host.
//
// DEFLATE (RFC 1951) is sequential inside one stream: symbol i+1 starts where symbol i ended.  The parallelism is over
// files (a thousand genomes = a thousand warps, seven per SM), and inside a warp over the OUTPUT side:
//   * lane 0 decodes up to 32 tokens (literal, or length/distance) through a 10-bit first-level table per Huffman
//     code in shared memory (longer codes: canonical bit-by-bit decode), from a 2 KB ring of compressed bytes the whole
//     warp keeps filled with 16-byte loads;
//   * all lanes then place the tokens: a warp scan of the lengths gives every token its output position; literals and
//     matches whose source lies entirely before the batch are written by their own lane, all at once; the (rare)
//     matches that read bytes of this very batch follow in token order, copied by the whole warp;
//   * after a member's last block the warp checks the gzip trailer like zlib does: CRC-32 over the member's output
//     (32 lane-local CRCs joined by multiplying with x^(8 len) mod P) and ISIZE.
// Anything zlib would refuse (bad magic, reserved bits, over-subscribed or incomplete code sets, distances beyond the
// output, invalid symbols, length or CRC mismatch, truncation) and a decompressed 0x00 byte (the batch separator) make
// the FILE's status non-zero; the host then reads that batch through its zlib reader, which defines what the reference
// does with a damaged file.  The CRC also means a decoder fault could not pass silently.
//
// Output placement: the uncompressed size is only known at the end of the stream.  Every file first gets a slot of
// the size its trailer claims (ISIZE of the last member -- exact for the usual one-member file below 4 GB); a file
// whose real size differs keeps counting without writing, and the batch is laid out again with the measured sizes
// and inflated once more (bgzip-style multi-member files take that second pass).
#include <algorithm>
#include "common.h"
#include "gunzip_core.cuh"

namespace fpm {

// CRC-32 of out[begin, end) by the whole warp; *has_nul |= a 0x00 byte in there.  Every lane returns the result.
__device__ uint32_t gz_crc_warp(GzShared& sh, const uint8_t* out, uint64_t begin, uint64_t end, int lane, bool* has_nul)
{
    const uint64_t n = end - begin;
    const uint64_t per = ((n + 31) / 32 + 3) & ~3ull;                // equal chunks (the last ones shorter or empty)
    const uint64_t lo = min(begin + per * lane, end), hi = min(lo + per, end);
    uint32_t c = 0xffffffffu;
    bool nul = false;
    uint64_t i = lo;
    for (; i < hi && (i & 3); i++) { const uint32_t v = out[i]; nul |= v == 0; c = gz_crc_byte(sh.crc_tab, c, v); }
    for (; i + 4 <= hi; i += 4) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(out + i);
        nul |= ((w - 0x01010101u) & ~w & 0x80808080u) != 0;
        c = gz_crc_byte(sh.crc_tab, c, w);
        c = gz_crc_byte(sh.crc_tab, c, w >> 8);
        c = gz_crc_byte(sh.crc_tab, c, w >> 16);
        c = gz_crc_byte(sh.crc_tab, c, w >> 24);
    }
    for (; i < hi; i++) { const uint32_t v = out[i]; nul |= v == 0; c = gz_crc_byte(sh.crc_tab, c, v); }
    sh.crc_part[lane] = ~c;
    if (__any_sync(0xffffffffu, nul)) *has_nul = true;
    __syncwarp();
    uint32_t total = 0;
    if (lane == 0) {
        // crc(A || B) = crc(A) x^(8 |B|) + crc(B) (mod P) for the conditioned CRCs
        const uint32_t xfull = gz_x8n_modp(per);
        for (int l = 0; l < 32; l++) {
            const uint64_t llo = min(begin + per * l, end), lhi = min(llo + per, end);
            const uint64_t len = lhi - llo;
            if (len == 0) break;
            const uint32_t x = len == per ? xfull : gz_x8n_modp(len);
            total = l == 0 ? sh.crc_part[0] : gz_multmodp(x, total) ^ sh.crc_part[l];
        }
    }
    total = __shfl_sync(0xffffffffu, total, 0);
    __syncwarp();
    return total;
}

// ---- one warp, one file -----------------------------------------------------------------------------------------
// in_off[f] .. in_off[f+1]: the file's compressed bytes (the buffer continues for >= 8 KB of zeros after the last file);
// out_off[f] .. out_off[f+1] - 1: its slot, one byte longer than the claimed size (the separator).
__global__ void __launch_bounds__(32) gunzip_kernel(const uint8_t* __restrict__ in, const uint64_t* __restrict__ in_off, uint32_t n_files, uint8_t* out,
                                                    const uint64_t* __restrict__ out_off, uint64_t* __restrict__ out_size, uint32_t* __restrict__ out_status)
{
    __shared__ GzShared sh;
    const int lane = threadIdx.x;
    const uint32_t f = blockIdx.x;
    if (f >= n_files) return;
    for (int i = lane; i < 256; i += 32) sh.crc_tab[i] = gz_crc_table_entry((uint32_t)i);
    const uint64_t in_end = in_off[f + 1];
    const uint64_t o_begin = out_off[f], o_limit = out_off[f + 1] - 1;          // writes go to [o_begin, o_limit)
    GzStream s;                            // lane 0's reader; the other lanes only follow wnext / whi for the ring top-ups
    s.init(sh, in, in_off[f], in_end);
    int state = in_end - in_off[f] > GZ_MAX_FILE ? GZ_S_BAD : GZ_S_HEADER;
    uint64_t opos = o_begin;               // next output byte (keeps counting beyond o_limit)
    uint64_t member_begin = o_begin;
    bool has_nul = false;
    __syncwarp();

    while (state != GZ_S_DONE && state != GZ_S_BAD) {
        // ---- keep the ring ahead of the reader: whole-warp 512-byte top-ups ----------------------------------------
        {
            const uint32_t wnext = __shfl_sync(0xffffffffu, s.b.wnext(), 0);
            uint32_t whi = __shfl_sync(0xffffffffu, s.b.whi, 0);
            while ((int32_t)(whi - wnext) <= GZ_RING_WORDS - 128) {          // (negative right after a seek: nothing valid yet)
                const uint4 v = *reinterpret_cast<const uint4*>(s.b.base + 4ull * whi + 16 * lane);
                uint32_t* r = sh.ring + ((whi + 4 * lane) & (GZ_RING_WORDS - 1));
                r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
                whi += 128;
            }
            s.b.whi = whi;
            __syncwarp();
        }
        // ---- lane 0: one step of the stream ---------------------------------------------------------------------------
        uint32_t ntok = 0, ready = 0;
        int next_state = state;
        if (lane == 0) {
            if (state == GZ_S_HEADER) next_state = gz_read_header(s);
            else if (state == GZ_S_BLOCK) next_state = gz_read_block(s, sh);
            else if (state == GZ_S_CODES) next_state = gz_decode_batch(s, sh, &ntok, &ready);
        }
        ntok = __shfl_sync(0xffffffffu, ntok, 0);
        ready = __shfl_sync(0xffffffffu, ready, 0);
        next_state = __shfl_sync(0xffffffffu, next_state, 0);
        __syncwarp();

        if (state == GZ_S_CODES && ntok) {
            // ---- every lane decodes the token whose first bit lane 0 found, then the tokens are placed -------------------
            uint32_t tok = lane < (int)ntok ? sh.tok[lane] : 0;
            if (lane < (int)ntok && !((ready >> lane) & 1u)) tok = gz_token_at(sh, tok);
            const bool is_match = lane < (int)ntok && (tok >> 31);
            const uint32_t len = lane < (int)ntok ? (is_match ? (tok >> 16) & 0x1ffu : 1u) : 0u;
            const uint32_t dst = (tok & 0xffffu) + 1u;
            uint32_t inc = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += up;
            }
            const uint64_t pos = opos + (inc - len);
            const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
            const bool too_far = is_match && (uint64_t)dst > pos - member_begin;
            if (__any_sync(0xffffffffu, too_far)) next_state = GZ_S_BAD;
            else {
                const bool writable = opos + total <= o_limit;              // otherwise this pass only measures the size
                const bool indep = is_match && pos - dst + min(len, dst) <= opos;      // its source was complete before this batch
                if (writable) {
                    if (lane < (int)ntok && !is_match) out[pos] = (uint8_t)tok;
                    if (indep) {
                        const uint8_t* src = out + pos - dst;
                        uint8_t* d8 = out + pos;
                        if (dst >= len) {
                            for (uint32_t j = 0; j < len; j += 16) {          // loads first, then stores: one memory latency per 16 bytes
                                uint8_t v[16];
#pragma unroll
                                for (int u = 0; u < 16; u++) if (j + u < len) v[u] = src[j + u];
#pragma unroll
                                for (int u = 0; u < 16; u++) if (j + u < len) d8[j + u] = v[u];
                            }
                        } else {
                            for (uint32_t j = 0; j < len; j++) d8[j] = src[j % dst];
                        }
                    }
                }
                __syncwarp();
                // matches that read bytes of this batch: in token order, the whole warp on each
                uint32_t dep = __ballot_sync(0xffffffffu, is_match && !indep);
                while (dep) {
                    const int t = __ffs(dep) - 1;
                    dep &= dep - 1;
                    const uint64_t p_t = __shfl_sync(0xffffffffu, pos, t);
                    const uint32_t l_t = __shfl_sync(0xffffffffu, len, t), d_t = __shfl_sync(0xffffffffu, dst, t);
                    if (writable)
                        for (uint32_t j = lane; j < l_t; j += 32) out[p_t + j] = out[p_t - d_t + (j % d_t)];
                    __syncwarp();
                }
            }
            opos += total;
        }
        if (next_state == GZ_S_STORED) {
            // ---- stored block: a plain copy by the whole warp, then the reader restarts behind it ---------------------
            const uint32_t n = __shfl_sync(0xffffffffu, s.stored_len, 0);
            const uint64_t src0 = __shfl_sync(0xffffffffu, (unsigned long long)s.in_pos(), 0);
            const int last = __shfl_sync(0xffffffffu, (int)s.last_block, 0);
            if (src0 + n > in_end) next_state = GZ_S_BAD;
            else {
                if (opos + n <= o_limit)
                    for (uint32_t j = lane; j < n; j += 32) out[opos + j] = in[src0 + j];
                opos += n;
                s.seek(src0 + n);
                next_state = last ? GZ_S_TRAILER : GZ_S_BLOCK;
            }
            __syncwarp();
        }
        if (next_state == GZ_S_TRAILER) {
            // ---- member trailer: CRC-32 and ISIZE, as zlib checks them ------------------------------------------------
            uint32_t crc_stored = 0, isize = 0;
            int ok = 0;
            if (lane == 0) ok = gz_read_trailer(s, &crc_stored, &isize);
            ok = __shfl_sync(0xffffffffu, ok, 0);
            crc_stored = __shfl_sync(0xffffffffu, crc_stored, 0);
            isize = __shfl_sync(0xffffffffu, isize, 0);
            if (!ok || isize != (uint32_t)(opos - member_begin)) next_state = GZ_S_BAD;
            else {
                if (opos <= o_limit && gz_crc_warp(sh, out, member_begin, opos, lane, &has_nul) != crc_stored) next_state = GZ_S_BAD;
                if (next_state != GZ_S_BAD) {
                    int more = 0;
                    if (lane == 0) more = gz_more_members(s);
                    more = __shfl_sync(0xffffffffu, more, 0);
                    member_begin = opos;
                    next_state = more ? GZ_S_HEADER : GZ_S_DONE;
                }
            }
        }
        state = next_state;
    }
    if (lane == 0) {
        const uint64_t size = opos - o_begin;
        out_size[f] = size;
        uint32_t st = GZ_OK;
        if (state == GZ_S_BAD) st = GZ_BAD;
        else if (size != o_limit - o_begin) st = GZ_RESIZE;
        else if (has_nul) st = GZ_HAS_NUL;
        else out[o_limit] = 0;                                   // the batch separator fpm_fasta_parse expects
        out_status[f] = st;
    }
}

}  // namespace fpm

using namespace fpm;

extern "C" {

int fpm_gunzip_batch(fpm_ctx* ctx, const uint8_t* gz, const uint64_t* gz_offsets, uint32_t n_files, uint64_t* out_file_end, uint64_t* out_total, int* out_status)
{
    if (!ctx) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    if (!gz_offsets || !out_file_end || !out_total || !out_status) { set_error("NULL argument"); return FPM_ERR_ARG; }
    *out_total = 0;
    *out_status = FPM_GUNZIP_OK;
    ctx->fa_resident = 0;
    if (n_files == 0) return FPM_OK;
    if (!gz) { set_error("gz is NULL"); return FPM_ERR_ARG; }
    for (uint32_t i = 0; i < n_files; i++)
        if (gz_offsets[i + 1] < gz_offsets[i]) { set_error("gz_offsets must not decrease"); return FPM_ERR_ARG; }
    FPM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint64_t in_bytes = gz_offsets[n_files];
    constexpr uint64_t PAD = 8192;
    int rc;
    if ((rc = ctx->gz_in.ensure(in_bytes + PAD))) return rc;
    if ((rc = ctx->gz_meta.ensure((uint64_t)(n_files + 1) * 8 * 3 + (uint64_t)n_files * 4 + 64))) return rc;
    uint64_t* d_in_off = ctx->gz_meta.as<uint64_t>();
    uint64_t* d_out_off = d_in_off + n_files + 1;
    uint64_t* d_size = d_out_off + n_files + 1;
    uint32_t* d_status = (uint32_t*)(d_size + n_files + 1);
    FPM_CUDA(cudaMemcpyAsync(ctx->gz_in.p, gz, in_bytes, cudaMemcpyHostToDevice, st));
    FPM_CUDA(cudaMemsetAsync(ctx->gz_in.as<uint8_t>() + in_bytes, 0, PAD, st));
    FPM_CUDA(cudaMemcpyAsync(d_in_off, gz_offsets, sizeof(uint64_t) * (n_files + 1), cudaMemcpyHostToDevice, st));
    // first layout: the size each file's trailer claims (RFC 1952 ISIZE)
    std::vector<uint64_t> off(n_files + 1), size(n_files);
    std::vector<uint32_t> status(n_files);
    off[0] = 0;
    for (uint32_t i = 0; i < n_files; i++) {
        const uint64_t b = gz_offsets[i], e = gz_offsets[i + 1];
        uint64_t claim = 0;
        if (e - b >= 18) claim = (uint64_t)gz[e - 4] | ((uint64_t)gz[e - 3] << 8) | ((uint64_t)gz[e - 2] << 16) | ((uint64_t)gz[e - 1] << 24);
        // (a trailer can claim anything: beyond 16 : 1 -- FASTA compresses 3 to 6 : 1 -- the first pass only measures, see below)
        claim = std::min<uint64_t>(claim, (e - b) * 16 + 65536);
        off[i + 1] = off[i] + claim + 1;
    }
    for (int pass = 0; pass < 2; pass++) {
        if ((rc = ctx->fa_raw.ensure(off[n_files] + 64))) return rc;
        FPM_CUDA(cudaMemcpyAsync(d_out_off, off.data(), sizeof(uint64_t) * (n_files + 1), cudaMemcpyHostToDevice, st));
        ctx->time_begin(FPM_KERNEL_GUNZIP);
        gunzip_kernel<<<n_files, 32, 0, st>>>(ctx->gz_in.as<uint8_t>(), d_in_off, n_files, ctx->fa_raw.as<uint8_t>(), d_out_off, d_size, d_status);
        ctx->time_end();
        ctx->launches++;
        FPM_CUDA(cudaGetLastError());
        FPM_CUDA(cudaMemcpyAsync(size.data(), d_size, sizeof(uint64_t) * n_files, cudaMemcpyDeviceToHost, st));
        FPM_CUDA(cudaMemcpyAsync(status.data(), d_status, sizeof(uint32_t) * n_files, cudaMemcpyDeviceToHost, st));
        FPM_CUDA(cudaStreamSynchronize(st));
        bool resize = false;
        for (uint32_t i = 0; i < n_files; i++) {
            if (status[i] == GZ_BAD) { *out_status = FPM_GUNZIP_BAD_STREAM; return FPM_OK; }
            if (status[i] == GZ_HAS_NUL) { *out_status = FPM_GUNZIP_HAS_NUL; return FPM_OK; }
            resize |= status[i] == GZ_RESIZE;
        }
        if (!resize) break;
        if (pass == 1) { set_error("gunzip: sizes changed between passes"); return FPM_ERR_CUDA; }
        for (uint32_t i = 0; i < n_files; i++) off[i + 1] = off[i] + size[i] + 1;
    }
    for (uint32_t i = 0; i < n_files; i++) out_file_end[i] = off[i + 1] - 1;
    *out_total = off[n_files];
    ctx->fa_resident = off[n_files];
    return FPM_OK;
}

int fpm_gunzip_output(fpm_ctx* ctx, uint8_t* out)
{
    if (!ctx || !out) { set_error("NULL argument"); return FPM_ERR_ARG; }
    FPM_CUDA(cudaSetDevice(ctx->device));
    if (ctx->fa_resident) FPM_CUDA(cudaMemcpyAsync(out, ctx->fa_raw.p, ctx->fa_resident, cudaMemcpyDeviceToHost, ctx->stream));
    FPM_CUDA(cudaStreamSynchronize(ctx->stream));
    return FPM_OK;
}

}  // extern "C"
