// cfl.cu -- lyn2vec's "basic" fingerprints on the GPU (the producer of `-fp` inputs):
// --type_factorization CFL, ICFL, CFL_ICFL-<C> and their *_COMB forms (CFL_COMB, ICFL_COMB, CFL_ICFL_COMB-<C>).
//
// Replaces, for --type basic --shift shift (the README recipe, README.md:34-52, uses CFL):
//   shift_string      (lyn2vec/fingerprint_utils.py:95-110)  every circular window of `window` (=100) characters
//   CFL               (lyn2vec/factorizations.py:102-126)    Duval's Lyndon factorisation of each window
//   ICFL_recursive    (factorizations.py:143-248)            the inverse Lyndon factorisation
//   CFL_icfl          (factorizations.py:265-300)            CFL with long factors sub-factorised by ICFL
//   d_duval_          (lyn2vec/factorizations_comb.py:213-245) the *_COMB forms: the factorisation of the window refined by the
//                                                            (mirrored) factorisation of its reverse complement
//   the line writer   (fingerprint_utils.py:443-476)         one row of factor LENGTHS per window
// and fuses getHashFingerPrint (hash.cpp:45-73) over each row, so a FASTA record can go straight to the
// `-fp` sketch (one 32-bit hash per window) without the text file -- which can still be written from
// the token rows, byte-identical with lyn2vec's.
//
// One thread per window: Duval is sequential within a word (O(window) steps) and the windows are
// independent.  Characters compare by byte value, as Python compares single characters.
#include "common.h"
#include "murmur3.cuh"

namespace fpm {

constexpr int CFL_MAX_WINDOW = 256;

// Rows of factor lengths go to the optional token buffer and, two tokens per block, into MurmurHash3
// (getHashFingerPrint, hash.cpp:45-73, hashes the row as an array of uint64).
struct RowSink {
    uint64_t h1, h2, pending;
    uint32_t ntok;
    uint16_t* row;
    __device__ __forceinline__ void emit(uint32_t flen)
    {
        if (row) row[ntok] = (uint16_t)flen;
        if (ntok & 1) mm_block(h1, h2, pending, (uint64_t)flen); else pending = flen;
        ntok++;
    }
    __device__ __forceinline__ uint64_t finish()
    {
        if (ntok & 1) h1 ^= mm_k1(pending);                 // 8-byte tail
        return mm_finish(h1, h2, (uint64_t)ntok * 8);
    }
};

// *_COMB: the factor boundaries of a word as a bit set (bit p = a factor ends after p characters), either as emitted or
// mirrored (the factorisation of the reverse complement, read backwards: a boundary after c characters of the reverse
// complement is a boundary after len - c characters of the word).
struct CutSink {
    uint32_t bits[CFL_MAX_WINDOW / 32 + 1];
    uint32_t pos, len;
    bool mirror;
    __device__ __forceinline__ void mark(uint32_t p) { bits[p >> 5] |= 1u << (p & 31); }
    __device__ __forceinline__ void emit(uint32_t flen)
    {
        if (mirror) mark(len - pos);
        pos += flen;
        if (!mirror) mark(pos);
    }
};

// ICFL, the inverse Lyndon factorisation (lyn2vec/factorizations.py:143-248, ICFL_recursive), without recursion.
// The reference recurses on a suffix: find_pre (:172-189) scans the longest prefix x[0..j) on which
// x[j'] <= x[i'] keeps holding (a non-increasing "anti-Lyndon" run) up to the first character x[j] that breaks
// it; find_bre (:212-232) walks the border chain of x[0..j) (KMP failure function, border(), :235-248) for the
// shortest border b with x[b] < x[j] (`last`), cuts p = x[0 .. j-last) off and recurses on the rest.  On the way
// back (compute_icfl_recursive, :152-169) p becomes a factor of its own iff the first factor found so far is
// longer than `last`, else it is glued to it.  Here: a forward pass records (|p|, last) per level, a backward
// pass decides the cuts, a second forward pass emits the lengths in order.
// scratch: plen/aux/f hold n bytes each (n <= 256, so every value fits a byte: |p| <= j <= 255).
template <typename Sink>
__device__ void icfl_lengths(const uint8_t* w, int n, uint8_t* plen, uint8_t* aux, uint8_t* f, Sink& sink)
{
    int pos = 0, levels = 0, tail_len = 0;
    for (;;) {
        const int m = n - pos;
        const uint8_t* x = w + pos;
        int j = m;
        if (m > 1) {                                         // find_pre
            int i = 0;
            j = 1;
            while (j < m && x[j] <= x[i]) { i = x[j] < x[i] ? 0 : i + 1; j++; }
        }
        if (j >= m) { tail_len = m; break; }                 // the rest is one factor (the "$" case, :158-161)
        f[0] = 0;                                            // border() of x[0..j)
        for (int i = 1, k = 0; i < j; i++) {
            while (k > 0 && x[k] != x[i]) k = f[k - 1];
            if (x[k] == x[i]) k++;
            f[i] = (uint8_t)k;
        }
        int i = j, last = f[j - 1];                          // find_bre
        while (i > 0) {
            if (x[f[i - 1]] < x[j]) last = f[i - 1];
            i = f[i - 1];
        }
        plen[levels] = (uint8_t)(j - last);
        aux[levels] = (uint8_t)last;
        levels++;
        pos += j - last;
    }
    int first = tail_len;                                    // length of the current first factor, unwinding
    for (int t = levels - 1; t >= 0; t--) {
        const bool cut = first > (int)aux[t];                // :165-168
        first = cut ? (int)plen[t] : first + (int)plen[t];
        aux[t] = cut;
    }
    uint32_t acc = 0;
    for (int t = 0; t < levels; t++) {
        acc += plen[t];
        if (aux[t]) { sink.emit(acc); acc = 0; }
    }
    sink.emit(acc + tail_len);
}

// One word through CFL (Duval, factorizations.py:102-126), ICFL or CFL_ICFL; factor lengths go to `sink` in order.
template <typename Sink>
__device__ void factorise(const uint8_t* word, uint32_t len, int mode, uint32_t sub_len, uint8_t* s_plen, uint8_t* s_aux, uint8_t* s_f, Sink& sink)
{
    if (mode == FPM_FACT_ICFL) {
        icfl_lengths(word, (int)len, s_plen, s_aux, s_f, sink);
        return;
    }
    uint32_t i = 0;
    while (i < len) {
        uint32_t j = i + 1, k = i;
        while (j < len && word[k] <= word[j]) {
            k = word[k] < word[j] ? i : k + 1;
            j++;
        }
        const uint32_t flen = j - k;
        while (i <= k) {
            if (mode == FPM_FACT_CFL_ICFL && flen > sub_len) icfl_lengths(word + i, (int)flen, s_plen, s_aux, s_f, sink);
            else sink.emit(flen);
            i += flen;
        }
    }
}

// mode: FPM_FACT_CFL, FPM_FACT_ICFL, FPM_FACT_CFL_ICFL (CFL factors longer than `sub_len` are sub-factorised with
// ICFL, CFL_icfl, factorizations.py:265-300; the "<<" ">>" markers are dropped from the rows, fingerprint_utils.py:459-463)
__global__ void __launch_bounds__(128) cfl_window_kernel(const uint8_t* __restrict__ seq, const uint64_t* __restrict__ rec_off,
                                                         const uint64_t* __restrict__ win_off, uint32_t n_rec, uint32_t window, int mode,
                                                         uint32_t sub_len, uint32_t seed, int use64, uint64_t* __restrict__ out_hash,
                                                         uint16_t* __restrict__ out_tok, uint16_t* __restrict__ out_ntok)
{
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= win_off[n_rec]) return;
    uint32_t lo = 0, hi = n_rec - 1;                       // record holding window w
    while (lo < hi) { uint32_t mid = (lo + hi + 1) >> 1; if (win_off[mid] <= w) lo = mid; else hi = mid - 1; }
    const uint64_t base = rec_off[lo], n = rec_off[lo + 1] - base, shift = w - win_off[lo];
    const uint32_t len = n < window ? (uint32_t)n : window;   // shorter records give one window: the record itself
    uint8_t word[CFL_MAX_WINDOW];
    uint8_t s_plen[CFL_MAX_WINDOW], s_aux[CFL_MAX_WINDOW], s_f[CFL_MAX_WINDOW];   // ICFL scratch (local memory; unused for CFL)
    for (uint32_t j = 0; j < len; j++) {
        uint64_t p = shift + j;
        if (p >= n) p -= n;                                 // circular wrap (fingerprint_utils.py:104-108)
        word[j] = seq[base + p];
    }
    RowSink sink;
    sink.h1 = seed; sink.h2 = seed; sink.pending = 0; sink.ntok = 0;
    sink.row = out_tok ? out_tok + w * window : nullptr;
    if (mode < FPM_FACT_CFL_COMB) {
        factorise(word, len, mode, sub_len, s_plen, s_aux, s_f, sink);
    } else {
        // d_duval_ (factorizations_comb.py:213-245): the boundaries of alg(word) together with the mirrored boundaries of
        // alg(reverse complement).  Its quirk is kept: the reverse complement of CFL_ICFL_COMB-<C> is factorised with the
        // default threshold 30 whatever C is (:221 calls alg(complement) without k).
        const int base = mode - FPM_FACT_CFL_COMB;
        CutSink cuts;
        for (uint32_t i = 0; i < sizeof cuts.bits / 4; i++) cuts.bits[i] = 0;
        cuts.len = len;
        cuts.pos = 0; cuts.mirror = false;
        factorise(word, len, base, sub_len, s_plen, s_aux, s_f, cuts);
        uint8_t rc[CFL_MAX_WINDOW];
        for (uint32_t j = 0; j < len; j++) {                // reverse_complement (:8-10): A<->T, C<->G, N stays (other letters make the reference raise)
            const uint8_t c = word[len - 1 - j];
            rc[j] = c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C' : c;
        }
        cuts.pos = 0; cuts.mirror = true;
        factorise(rc, len, base, 30u, s_plen, s_aux, s_f, cuts);
        uint32_t last = 0;
        for (uint32_t p = 1; p <= len; p++)
            if ((cuts.bits[p >> 5] >> (p & 31)) & 1u) { sink.emit(p - last); last = p; }
    }
    const uint64_t h = sink.finish();
    if (out_hash) out_hash[w] = use64 ? h : (h & 0xffffffffULL);
    if (out_ntok) out_ntok[w] = (uint16_t)sink.ntok;
}

}  // namespace fpm

using namespace fpm;

extern "C" int fpm_fingerprint_batch(fpm_ctx* ctx, const uint8_t* seq, const uint64_t* rec_offsets, uint32_t n_records, uint32_t window,
                                     int factorization, uint32_t sub_len, uint32_t seed, int use64, uint64_t* out_hashes,
                                     uint16_t* out_tokens, uint16_t* out_ntokens, uint64_t* out_window_offsets)
{
    if (!ctx) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    if (factorization < FPM_FACT_CFL || factorization > FPM_FACT_CFL_ICFL_COMB) {
        set_error("unknown factorization %d", factorization);
        return FPM_ERR_ARG;
    }
    if (!rec_offsets || !out_window_offsets) { set_error("NULL buffer"); return FPM_ERR_ARG; }
    if (window < 1 || window > CFL_MAX_WINDOW) { set_error("window %u outside 1..%d", window, CFL_MAX_WINDOW); return FPM_ERR_ARG; }
    FPM_CUDA(cudaSetDevice(ctx->device));
    // windows per record: len(record) circular shifts, or one window when the record is shorter (shift_string)
    std::vector<uint64_t> woff(n_records + 1, 0);
    for (uint32_t r = 0; r < n_records; r++) {
        if (rec_offsets[r + 1] < rec_offsets[r]) { set_error("record offsets not ascending"); return FPM_ERR_ARG; }
        uint64_t n = rec_offsets[r + 1] - rec_offsets[r];
        woff[r + 1] = woff[r] + (n == 0 ? 0 : (n < window ? 1 : n));
    }
    memcpy(out_window_offsets, woff.data(), sizeof(uint64_t) * (n_records + 1));
    const uint64_t n_win = woff[n_records], n_bytes = rec_offsets[n_records];
    if (n_win == 0 || (!out_hashes && !out_tokens && !out_ntokens)) return FPM_OK;
    int rc;
    cudaStream_t st = ctx->stream;
    if ((rc = ctx->seq.ensure(n_bytes + 64))) return rc;
    if ((rc = ctx->goff.ensure(sizeof(uint64_t) * 2 * (n_records + 1)))) return rc;
    if ((rc = ctx->outh.ensure(sizeof(uint64_t) * n_win))) return rc;
    if ((rc = ctx->outc.ensure(out_tokens ? sizeof(uint16_t) * n_win * window : 8))) return rc;
    if ((rc = ctx->outn.ensure(sizeof(uint16_t) * n_win))) return rc;
    uint64_t* d_rec = ctx->goff.as<uint64_t>();
    uint64_t* d_win = d_rec + (n_records + 1);
    FPM_CUDA(cudaMemcpyAsync(ctx->seq.p, seq, n_bytes, cudaMemcpyHostToDevice, st));
    FPM_CUDA(cudaMemcpyAsync(d_rec, rec_offsets, sizeof(uint64_t) * (n_records + 1), cudaMemcpyHostToDevice, st));
    FPM_CUDA(cudaMemcpyAsync(d_win, woff.data(), sizeof(uint64_t) * (n_records + 1), cudaMemcpyHostToDevice, st));
    cfl_window_kernel<<<(uint32_t)((n_win + 127) / 128), 128, 0, st>>>(ctx->seq.as<uint8_t>(), d_rec, d_win, n_records, window, factorization, sub_len, seed, use64,
                                                                       ctx->outh.as<uint64_t>(), out_tokens ? ctx->outc.as<uint16_t>() : nullptr,
                                                                       ctx->outn.as<uint16_t>());
    ctx->launches++;
    FPM_CUDA(cudaGetLastError());
    if (out_hashes) FPM_CUDA(cudaMemcpyAsync(out_hashes, ctx->outh.p, sizeof(uint64_t) * n_win, cudaMemcpyDeviceToHost, st));
    if (out_tokens) FPM_CUDA(cudaMemcpyAsync(out_tokens, ctx->outc.p, sizeof(uint16_t) * n_win * window, cudaMemcpyDeviceToHost, st));
    if (out_ntokens) FPM_CUDA(cudaMemcpyAsync(out_ntokens, ctx->outn.p, sizeof(uint16_t) * n_win, cudaMemcpyDeviceToHost, st));
    FPM_CUDA(cudaStreamSynchronize(st));
    return FPM_OK;
}

extern "C" int fpm_cfl_fingerprint_batch(fpm_ctx* ctx, const uint8_t* seq, const uint64_t* rec_offsets, uint32_t n_records, uint32_t window,
                                         uint32_t seed, int use64, uint64_t* out_hashes, uint16_t* out_tokens, uint16_t* out_ntokens,
                                         uint64_t* out_window_offsets)
{
    return fpm_fingerprint_batch(ctx, seq, rec_offsets, n_records, window, FPM_FACT_CFL, 0, seed, use64, out_hashes, out_tokens, out_ntokens,
                                 out_window_offsets);
}
