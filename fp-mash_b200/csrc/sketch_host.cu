// sketch_host.cu -- C-ABI entry points for sketching: pass orchestration around the kernels.
//
// Exactness does not depend on the threshold guess: a sketch is accepted only when the
// table held at least s qualifying hashes (then the s smallest of them ARE the bottom-s) or
// when the threshold admitted every hash; otherwise that sketch is re-run with a threshold
// scaled from the density just measured.
#include <algorithm>
#include <string.h>
#include "common.h"
#include "sketch_kernels.cuh"
#include "sketch_launch.h"
#include "sketch_select.h"
#include "nccl_dyn.h"
#include <math.h>

namespace fpm {

#define FPM_TAB(K) launch_sketch_hash_k##K,
static const sketch_hash_launcher g_hash_launch[32] = {FPM_FOR_ALL_K(FPM_TAB)};
#undef FPM_TAB
#define FPM_TAB(K) launch_hash_stream_k##K,
static const hash_stream_launcher g_stream_launch[32] = {FPM_FOR_ALL_K(FPM_TAB)};
#undef FPM_TAB
#define FPM_TAB(K) launch_count_windows_k##K,
static const count_windows_launcher g_count_launch[32] = {FPM_FOR_ALL_K(FPM_TAB)};
#undef FPM_TAB

static uint32_t pow2ceil(uint64_t v)
{
    uint64_t p = 64;
    while (p < v) p <<= 1;
    return (uint32_t)std::min<uint64_t>(p, 1ull << 31);
}

static bool is_nucleotide(const fpm_sketch_params* p)
{
    for (int c = 0; c < 256; c++)
        if ((p->alphabet[c] != 0) != (c == 'A' || c == 'C' || c == 'G' || c == 'T')) return false;
    return true;
}

static int check_params(const fpm_sketch_params* p)
{
    if (!p) { set_error("sketch params are NULL"); return FPM_ERR_ARG; }
    if (p->kmer_size < 1 || p->kmer_size > 32) { set_error("k-mer size %d outside 1..32", p->kmer_size); return FPM_ERR_ARG; }
    if (p->sketch_size < 1) { set_error("sketch size must be >= 1"); return FPM_ERR_ARG; }
    if (p->min_cov < 1) { set_error("min_cov must be >= 1"); return FPM_ERR_ARG; }
    // ACGT (alphabetNucleotide, Sketch.h:25) takes the 2-bit kernel; any other alphabet the generic
    // byte kernel, which -- like the reference with -a / -z -- is noncanonical only.
    int asize = 0;
    for (int c = 0; c < 256; c++) asize += p->alphabet[c] != 0;
    if (asize == 0 || p->alphabet[0]) { set_error("alphabet is empty or contains NUL"); return FPM_ERR_ARG; }
    if (!is_nucleotide(p) && !p->noncanonical) {
        set_error("canonical k-mers are only defined for the nucleotide alphabet ACGT; custom alphabets imply -n");
        return FPM_ERR_UNSUPPORTED;
    }
    bool use64 = pow((double)asize, (double)p->kmer_size) > pow(2.0, 32.0);   // Sketch.cpp:1288
    if ((p->use64 != 0) != use64) { set_error("use64=%d inconsistent with k=%d for a %d-letter alphabet", p->use64, p->kmer_size, asize); return FPM_ERR_ARG; }
    if (p->sketch_size > (1u << 24)) { set_error("sketch size %u too large (the reference stores it in a float: exact only up to 2^24)", p->sketch_size); return FPM_ERR_ARG; }
    return FPM_OK;
}

struct GroupPlan {
    uint64_t thresh;     // accept h <= thresh
    bool all;            // thresh admits every hash
    uint32_t cap;        // table capacity (power of two)
};

static const uint64_t kAll64 = ~0ULL;

// [lo, hi) byte ranges covering the active sketches, adjacent active sketches merged into one range
static void active_ranges(const std::vector<uint8_t>& active, const uint64_t* goff, uint32_t n_groups,
                          std::vector<std::pair<uint64_t, uint64_t>>& out)
{
    for (uint32_t g = 0; g < n_groups; g++) {
        if (!active[g] || goff[g + 1] == goff[g]) continue;
        if (!out.empty() && out.back().second == goff[g]) out.back().second = goff[g + 1];
        else out.push_back({goff[g], goff[g + 1]});
    }
}

static uint64_t scale_threshold(uint64_t bits_full /*2^bits - 1*/, double frac)
{
    if (frac >= 1.0) return bits_full;
    long double v = (long double)bits_full * (long double)frac;
    if (v < 1) v = 1;
    return (uint64_t)v;
}

// room for `bytes` more in the HBM-resident read stream: grows by doubling, keeping what is already there
int stream_reserve(fpm_ctx* ctx, uint64_t bytes)
{
    if (ctx->stream_used + bytes + 64 <= ctx->stream_buf.cap) return FPM_OK;
    size_t want = std::max<size_t>(ctx->stream_buf.cap * 2, ctx->stream_used + bytes + 64);
    want = std::max<size_t>(want, (size_t)256 << 20);
    void* np = nullptr;
    cudaError_t e = cudaMalloc(&np, want);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(stream buffer)", __FILE__, __LINE__);
    if (ctx->stream_used) FPM_CUDA(cudaMemcpyAsync(np, ctx->stream_buf.p, ctx->stream_used, cudaMemcpyDeviceToDevice, ctx->stream));
    FPM_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->stream_buf.p) cudaFree(ctx->stream_buf.p);
    ctx->stream_buf.p = np;
    ctx->stream_buf.cap = want;
    return FPM_OK;
}

int sketch_batch_dev_impl(fpm_ctx* ctx, const fpm_sketch_params* p, const uint8_t* d_seq, uint64_t n_bytes,
                          const uint64_t* h_goff, uint32_t n_groups, uint64_t* d_out_hashes, uint32_t* d_out_counts,
                          uint32_t* d_out_n, uint64_t* d_out_kmers)
{
    int rc = check_params(p);
    if (rc) return rc;
    if (n_groups == 0) return FPM_OK;
    if ((!d_seq && n_bytes) || !h_goff || !d_out_hashes || !d_out_n) { set_error("NULL buffer"); return FPM_ERR_ARG; }   // (an empty batch has no buffer: a read set whose reads are all shorter than k)
    if (h_goff[0] != 0 || h_goff[n_groups] != n_bytes) { set_error("group_offsets must start at 0 and end at seq_bytes"); return FPM_ERR_ARG; }
    for (uint32_t g = 0; g < n_groups; g++)
        if (h_goff[g + 1] < h_goff[g]) { set_error("group_offsets not ascending at %u", g); return FPM_ERR_ARG; }
    if (((uintptr_t)d_seq & 15) != 0) { set_error("d_seq must be 16-byte aligned"); return FPM_ERR_ARG; }
    if (n_bytes == 0) {
        FPM_CUDA(cudaMemsetAsync(d_out_n, 0, sizeof(uint32_t) * n_groups, ctx->stream));
        if (d_out_kmers) FPM_CUDA(cudaMemsetAsync(d_out_kmers, 0, sizeof(uint64_t) * n_groups, ctx->stream));
        return FPM_OK;
    }
    const uint64_t n_tiles64 = (n_bytes + SK_TILE_WINDOWS - 1) / SK_TILE_WINDOWS;
    if (n_tiles64 > 0x7fffffffull) { set_error("batch too large (%llu bytes)", (unsigned long long)n_bytes); return FPM_ERR_ARG; }
    const uint32_t n_tiles = (uint32_t)n_tiles64;

    const int K = p->kmer_size;
    const uint32_t s = p->sketch_size;
    const bool canon = !p->noncanonical;
    const bool nucleotide = is_nucleotide(p);
    if (!nucleotide) {
        if ((rc = ctx->alpha.ensure(256))) return rc;
        FPM_CUDA(cudaMemcpyAsync(ctx->alpha.p, p->alphabet, 256, cudaMemcpyHostToDevice, ctx->stream));
    }
    const uint64_t full = p->use64 ? kAll64 : 0xffffffffULL;
    const uint64_t target = s <= 4096 ? 2ull * s + 64 : (uint64_t)s + s / 4 + 256;
    const bool want_counts = p->want_counts && d_out_counts;
    cudaStream_t st = ctx->stream;

    // ---- plan ---------------------------------------------------------------------------
    std::vector<GroupPlan> plan(n_groups);
    for (uint32_t g = 0; g < n_groups; g++) {
        uint64_t n = h_goff[g + 1] - h_goff[g];
        GroupPlan& pl = plan[g];
        // -m > 1 (read sets): only hashes seen min_cov times qualify, so the first bound must admit about
        // coverage x more hash OCCURRENCES than sketch slots.  The coverage is not known yet; a guess of 32x costs a
        // larger table and a few thousand more atomics, and saves the whole second pass whenever the real coverage
        // is below it (bounded so that many-sketch batches do not multiply their table memory by 32).
        const uint64_t cov_guess = p->min_cov > 1 ? std::max<uint64_t>(1, std::min<uint64_t>(32, (64ull << 20) / (4 * target * n_groups))) : 1;
        if (n <= 2 * target * cov_guess) {
            pl.all = true; pl.thresh = full; pl.cap = pow2ceil(2 * n);
        } else {
            pl.all = false;
            pl.thresh = scale_threshold(full, (double)(target * cov_guess) / (double)n);
            pl.cap = pow2ceil(4 * target * cov_guess);
        }
    }

    // persistent per-call device arrays
    if ((rc = ctx->goff.ensure(sizeof(uint64_t) * (n_groups + 1)))) return rc;
    if ((rc = ctx->thresh.ensure(sizeof(uint64_t) * n_groups))) return rc;
    if ((rc = ctx->active.ensure(n_groups))) return rc;
    if ((rc = ctx->toff.ensure(sizeof(uint64_t) * n_groups))) return rc;
    if ((rc = ctx->tmask.ensure(sizeof(uint32_t) * n_groups))) return rc;
    if ((rc = ctx->maxcnt.ensure(sizeof(uint32_t) * n_groups))) return rc;
    if ((rc = ctx->maxpos.ensure(sizeof(uint64_t) * n_groups))) return rc;
    if ((rc = ctx->overflow.ensure(sizeof(uint32_t) * n_groups))) return rc;
    if ((rc = ctx->stat.ensure(sizeof(uint32_t) * 4 * n_groups))) return rc;
    if ((rc = ctx->args.ensure(sizeof(SketchArgs)))) return rc;
    FPM_CUDA(cudaMemcpyAsync(ctx->goff.p, h_goff, sizeof(uint64_t) * (n_groups + 1), cudaMemcpyHostToDevice, st));

    std::vector<uint8_t> active(n_groups, 1);
    std::vector<uint64_t> h_thresh(n_groups), h_toff(n_groups);
    std::vector<uint32_t> h_tmask(n_groups), h_stat(4 * (size_t)n_groups), h_over(n_groups), h_outn(n_groups, 0), h_topcnt(n_groups, 0);
    std::vector<std::pair<uint64_t, uint64_t>> ranges;

    uint32_t* d_stat_nq = ctx->stat.as<uint32_t>();
    uint32_t* d_stat_nd = d_stat_nq + n_groups;
    uint32_t* d_stat_top = d_stat_nd + n_groups;

    int max_smem = 0;
    FPM_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    if ((size_t)max_smem < (size_t)SK_SORT_CAP * 8) { set_error("device offers only %d bytes of opt-in shared memory", max_smem); return FPM_ERR_UNSUPPORTED; }
    if (configure_sketch_select((size_t)SK_SORT_CAP * 8) != 0) { set_error("cudaFuncSetAttribute(select) failed"); return FPM_ERR_CUDA; }

    // survivor log (only when multiplicities are wanted): everything the first pass inserts, so that the order-dependent top
    // count needs no second pass over the input.  Sized for the planned number of insertions; an overflow, or a sketch that
    // needed a re-run, falls back to the trace pass.
    uint64_t log_cap = 0;
    if (want_counts && nucleotide) {
        for (uint32_t g = 0; g < n_groups; g++) {
            const uint64_t n = h_goff[g + 1] - h_goff[g];
            const uint64_t cov_guess = p->min_cov > 1 ? std::max<uint64_t>(1, std::min<uint64_t>(32, (64ull << 20) / (4 * target * n_groups))) : 1;
            log_cap += std::min<uint64_t>(n, 4 * target * cov_guess);
        }
        log_cap = std::min<uint64_t>(log_cap + 1024, (uint64_t)1 << 28);
        if ((rc = ctx->firstpos.ensure(log_cap * 16 + 64))) return rc;
    }
    uint64_t* d_log_h = ctx->firstpos.as<uint64_t>() + 8;
    uint64_t* d_log_p = d_log_h + log_cap;
    unsigned long long* d_log_count = ctx->firstpos.as<unsigned long long>();
    int passes_run = 0;

    for (int pass = 0; pass < 12; pass++) {
        passes_run = pass + 1;
        // tables for the active groups
        uint64_t slots = 0;
        uint32_t max_cap = 64;
        uint32_t n_active = 0;
        for (uint32_t g = 0; g < n_groups; g++) {
            h_thresh[g] = plan[g].thresh;
            if (active[g]) {
                h_toff[g] = slots; h_tmask[g] = plan[g].cap - 1; slots += plan[g].cap;
                max_cap = std::max(max_cap, plan[g].cap); n_active++;
            } else { h_toff[g] = 0; h_tmask[g] = 0; }
        }
        if (n_active == 0) break;
        if ((rc = ctx->tkeys.ensure(slots * 8))) return rc;
        if ((rc = ctx->tcnt.ensure(slots * 4))) return rc;
        if ((rc = ctx->tpos.ensure(slots * 8))) return rc;
        FPM_CUDA(cudaMemsetAsync(ctx->tkeys.p, 0xff, slots * 8, st));
        FPM_CUDA(cudaMemsetAsync(ctx->tcnt.p, 0, slots * 4, st));
        FPM_CUDA(cudaMemsetAsync(ctx->tpos.p, 0xff, slots * 8, st));
        FPM_CUDA(cudaMemsetAsync(ctx->maxcnt.p, 0, sizeof(uint32_t) * n_groups, st));
        FPM_CUDA(cudaMemsetAsync(ctx->maxpos.p, 0xff, sizeof(uint64_t) * n_groups, st));
        FPM_CUDA(cudaMemsetAsync(ctx->overflow.p, 0, sizeof(uint32_t) * n_groups, st));
        FPM_CUDA(cudaMemcpyAsync(ctx->thresh.p, h_thresh.data(), sizeof(uint64_t) * n_groups, cudaMemcpyHostToDevice, st));
        FPM_CUDA(cudaMemcpyAsync(ctx->active.p, active.data(), n_groups, cudaMemcpyHostToDevice, st));
        FPM_CUDA(cudaMemcpyAsync(ctx->toff.p, h_toff.data(), sizeof(uint64_t) * n_groups, cudaMemcpyHostToDevice, st));
        FPM_CUDA(cudaMemcpyAsync(ctx->tmask.p, h_tmask.data(), sizeof(uint32_t) * n_groups, cudaMemcpyHostToDevice, st));

        // byte ranges to hash: everything on the first pass, afterwards the runs of adjacent active sketches
        ranges.clear();
        if (pass == 0) ranges.push_back({0, n_bytes});
        else active_ranges(active, h_goff, n_groups, ranges);

        SketchArgs a;
        memset(&a, 0, sizeof a);
        a.seq = d_seq; a.n_bytes = n_bytes; a.group_off = ctx->goff.as<uint64_t>(); a.n_groups = n_groups;
        a.thresh = ctx->thresh.as<uint64_t>(); a.active = ctx->active.as<uint8_t>();
        a.tkeys = ctx->tkeys.as<uint64_t>(); a.tcnt = ctx->tcnt.as<uint32_t>(); a.tpos = ctx->tpos.as<uint64_t>();
        a.toff = ctx->toff.as<uint64_t>(); a.tmask = ctx->tmask.as<uint32_t>();
        a.maxkey_cnt = ctx->maxcnt.as<uint32_t>(); a.maxkey_pos = ctx->maxpos.as<uint64_t>();
        a.overflow = ctx->overflow.as<uint32_t>();
        a.sketch_size = s; a.seed = p->seed; a.fold_case = !p->preserve_case; a.hash32 = !p->use64;
            a.c_tbl = 0x54474341u; a.c_add1 = 0x52dce729ULL; a.c_add2 = 0x38495ab5ULL; a.c_add1s = a.c_add1 + 5ull * a.seed;
        if (log_cap && pass == 0) {
            a.log_h = d_log_h; a.log_p = d_log_p; a.log_count = d_log_count; a.log_cap = log_cap;
            FPM_CUDA(cudaMemsetAsync(d_log_count, 0, 8, st));
        }
        FPM_CUDA(cudaMemcpyAsync(ctx->args.p, &a, sizeof a, cudaMemcpyHostToDevice, st));
        for (const auto& r : ranges) {
            ctx->time_begin(FPM_KERNEL_SKETCH_HASH);
            if (nucleotide) g_hash_launch[K - 1](canon, st, ctx->args.as<SketchArgs>(), r.first, r.second, 0);
            else launch_sketch_generic(st, ctx->args.as<SketchArgs>(), ctx->alpha.as<uint8_t>(), K, r.first, r.second, 0, nullptr);
            ctx->time_end();
            ctx->launches++;
            FPM_CUDA(cudaGetLastError());
        }

        SelectArgs sa;
        sa.tkeys = a.tkeys; sa.tcnt = a.tcnt; sa.tpos = a.tpos; sa.toff = a.toff; sa.tmask = a.tmask;
        sa.maxkey_cnt = a.maxkey_cnt; sa.maxkey_pos = a.maxkey_pos; sa.active = a.active;
        sa.sketch_size = s; sa.min_cov = p->min_cov;
        sa.sort_cap = std::min<uint32_t>(SK_SORT_CAP, max_cap);
        sa.thresh = a.thresh;
        sa.scratch = nullptr;
        if (max_cap > SK_SORT_CAP) {   // large sketches: sort qualifying keys in global memory
            if ((rc = ctx->scratch.ensure(slots * 8))) return rc;
            sa.scratch = ctx->scratch.as<uint64_t>();
        }
        sa.out_hashes = d_out_hashes; sa.out_counts = want_counts ? d_out_counts : nullptr;
        sa.out_firstpos = nullptr;
        sa.out_n = d_out_n; sa.stat_nq = d_stat_nq; sa.stat_nd = d_stat_nd; sa.stat_topcnt = d_stat_top;
        ctx->time_begin(FPM_KERNEL_SKETCH_SELECT);
        launch_sketch_select(n_groups, (size_t)sa.sort_cap * 8, st, sa);
        ctx->time_end();
        ctx->launches++;
        FPM_CUDA(cudaGetLastError());

        FPM_CUDA(cudaMemcpyAsync(h_stat.data(), ctx->stat.p, sizeof(uint32_t) * 3 * n_groups, cudaMemcpyDeviceToHost, st));
        FPM_CUDA(cudaMemcpyAsync(h_over.data(), ctx->overflow.p, sizeof(uint32_t) * n_groups, cudaMemcpyDeviceToHost, st));
        FPM_CUDA(cudaStreamSynchronize(st));

        // ---- decide ---------------------------------------------------------------------
        bool again = false;
        for (uint32_t g = 0; g < n_groups; g++) {
            if (!active[g]) continue;
            uint32_t nq = h_stat[g], nd = h_stat[n_groups + g];
            GroupPlan& pl = plan[g];
            uint64_t n = h_goff[g + 1] - h_goff[g];
            bool too_many = h_over[g] != 0;
            bool too_few = !pl.all && nq < s;
            if (!too_many && !too_few) {
                active[g] = 0;
                h_outn[g] = std::min<uint32_t>(nq, s);
                h_topcnt[g] = h_stat[2 * (size_t)n_groups + g];
                continue;
            }
            again = true;
            double cur = pl.all ? 1.0 : ((double)pl.thresh + 1.0) / ((double)full + 1.0);
            double next;
            if (too_many) next = cur / 4;
            else next = nq >= 16 ? cur * 1.15 * (double)target / (double)nq : cur * 16;
            if (next >= 1.0) { pl.all = true; pl.thresh = full; next = 1.0; }
            else { pl.all = false; pl.thresh = scale_threshold(full, next); }
            double nd_next = ((double)nd + 16) * (next / cur) * 1.3 + 256;
            if (nd_next > (double)n) nd_next = (double)n;
            pl.cap = std::max(pow2ceil((uint64_t)(2 * nd_next)), pow2ceil(4 * target));
        }
        if (!again) break;
        if (pass == 11) { set_error("bottom-s selection did not converge"); return FPM_ERR_CUDA; }
    }

    // ---- order-dependent multiplicity of the largest element (full sketches only) -------
    if (want_counts) {
        std::vector<uint32_t> tg;
        for (uint32_t g = 0; g < n_groups; g++)
            if (h_outn[g] == s && h_topcnt[g] > p->min_cov) tg.push_back(g);
        if (!tg.empty()) {
            // bucket layout from the counts just written (one copy and one wait, whole rows of all sketches when many need it)
            std::vector<uint64_t> h_troff((size_t)n_groups * s, 0);
            uint64_t total = 0;
            unsigned long long h_logged = ~0ULL;
            {
                const uint32_t g_first = tg.front(), g_last = tg.back();
                std::vector<uint32_t> rows((size_t)(g_last - g_first + 1) * s);
                FPM_CUDA(cudaMemcpyAsync(rows.data(), d_out_counts + (uint64_t)g_first * s, sizeof(uint32_t) * rows.size(), cudaMemcpyDeviceToHost, st));
                if (log_cap) FPM_CUDA(cudaMemcpyAsync(&h_logged, d_log_count, 8, cudaMemcpyDeviceToHost, st));
                FPM_CUDA(cudaStreamSynchronize(st));
                for (uint32_t g : tg)
                    for (uint32_t r = 0; r < s; r++) { h_troff[(uint64_t)g * s + r] = total; total += rows[(size_t)(g - g_first) * s + r]; }
            }
            const bool from_log = log_cap && passes_run == 1 && h_logged <= log_cap;
            if ((rc = ctx->tr_off.ensure(sizeof(uint64_t) * (uint64_t)n_groups * s))) return rc;
            if ((rc = ctx->tr_cursor.ensure(sizeof(uint32_t) * (uint64_t)n_groups * s))) return rc;
            if ((rc = ctx->tr_pos.ensure(sizeof(uint64_t) * std::max<uint64_t>(total, 1)))) return rc;
            if ((rc = ctx->glist.ensure(sizeof(uint32_t) * tg.size()))) return rc;
            FPM_CUDA(cudaMemcpyAsync(ctx->tr_off.p, h_troff.data(), sizeof(uint64_t) * (uint64_t)n_groups * s, cudaMemcpyHostToDevice, st));
            FPM_CUDA(cudaMemsetAsync(ctx->tr_cursor.p, 0, sizeof(uint32_t) * (uint64_t)n_groups * s, st));
            FPM_CUDA(cudaMemcpyAsync(ctx->glist.p, tg.data(), sizeof(uint32_t) * tg.size(), cudaMemcpyHostToDevice, st));
            std::fill(active.begin(), active.end(), 0);
            for (uint32_t g : tg) active[g] = 1;
            FPM_CUDA(cudaMemcpyAsync(ctx->active.p, active.data(), n_groups, cudaMemcpyHostToDevice, st));
            for (uint32_t g = 0; g < n_groups; g++) h_thresh[g] = plan[g].thresh;
            FPM_CUDA(cudaMemcpyAsync(ctx->thresh.p, h_thresh.data(), sizeof(uint64_t) * n_groups, cudaMemcpyHostToDevice, st));
            ranges.clear();
            active_ranges(active, h_goff, n_groups, ranges);
            SketchArgs a;
            memset(&a, 0, sizeof a);
            a.seq = d_seq; a.n_bytes = n_bytes; a.group_off = ctx->goff.as<uint64_t>(); a.n_groups = n_groups;
            a.thresh = ctx->thresh.as<uint64_t>(); a.active = ctx->active.as<uint8_t>();
            a.fin_hashes = d_out_hashes; a.fin_n = d_out_n; a.tr_off = ctx->tr_off.as<uint64_t>(); a.tr_cap = d_out_counts;
            a.tr_cursor = ctx->tr_cursor.as<uint32_t>(); a.tr_pos = ctx->tr_pos.as<uint64_t>();
            a.sketch_size = s; a.seed = p->seed; a.fold_case = !p->preserve_case; a.hash32 = !p->use64;
            a.c_tbl = 0x54474341u; a.c_add1 = 0x52dce729ULL; a.c_add2 = 0x38495ab5ULL; a.c_add1s = a.c_add1 + 5ull * a.seed;
            if (from_log) {
                // positions of the sketches' hashes straight from what the first pass logged
                SketchArgs al = a;
                al.log_h = d_log_h; al.log_p = d_log_p; al.log_count = nullptr; al.log_cap = log_cap;
                FPM_CUDA(cudaMemcpyAsync(ctx->args.p, &al, sizeof al, cudaMemcpyHostToDevice, st));
                launch_sketch_trace_log(h_logged, st, ctx->args.as<SketchArgs>());
                ctx->launches++;
                FPM_CUDA(cudaGetLastError());
            } else {
            FPM_CUDA(cudaMemcpyAsync(ctx->args.p, &a, sizeof a, cudaMemcpyHostToDevice, st));
            for (const auto& r : ranges) {
                ctx->time_begin(FPM_KERNEL_SKETCH_HASH);
                if (nucleotide) g_hash_launch[K - 1](canon, st, ctx->args.as<SketchArgs>(), r.first, r.second, 1);
                else launch_sketch_generic(st, ctx->args.as<SketchArgs>(), ctx->alpha.as<uint8_t>(), K, r.first, r.second, 1, nullptr);
                ctx->time_end();
                ctx->launches++;
                FPM_CUDA(cudaGetLastError());
            }
            }
            launch_sketch_topcount((uint32_t)tg.size(), st, ctx->glist.as<uint32_t>(), s, p->min_cov, a.tr_off, a.tr_cap, a.tr_pos, d_out_counts);
            ctx->launches++;
            FPM_CUDA(cudaGetLastError());
        }
    }

    if (d_out_kmers) {
        FPM_CUDA(cudaMemsetAsync(d_out_kmers, 0, sizeof(uint64_t) * n_groups, st));
        SketchArgs a;
        memset(&a, 0, sizeof a);
        a.seq = d_seq; a.n_bytes = n_bytes; a.group_off = ctx->goff.as<uint64_t>(); a.n_groups = n_groups;
        a.fold_case = !p->preserve_case;
        FPM_CUDA(cudaMemcpyAsync(ctx->args.p, &a, sizeof a, cudaMemcpyHostToDevice, st));
        if (nucleotide) g_count_launch[K - 1](n_tiles, st, ctx->args.as<SketchArgs>(), (unsigned long long*)d_out_kmers);
        else launch_sketch_generic(st, ctx->args.as<SketchArgs>(), ctx->alpha.as<uint8_t>(), K, 0, n_bytes, 2, (unsigned long long*)d_out_kmers);
        ctx->launches++;
        FPM_CUDA(cudaGetLastError());
    }
    FPM_CUDA(cudaStreamSynchronize(st));
    return FPM_OK;
}

// ---------------------------------------------------------------------------------------------------------
// ONE read set spread over the GPUs of a communicator (SURVEY.md 8e, row "read sketch"): `mash sketch -r` turns all
// reads into one sketch on a single thread in the reference (Sketch.cpp:203-210).  Every rank hashes a contiguous part
// of the read stream into a counting table of its own; the candidates -- (hash, count, first position) of every entry
// below the common bound -- are all-gathered and merged by key (counts add up, positions take the minimum), and every
// rank selects the bottom-s of the merged table: the set formulation of MinHashHeap is order
// independent.  The one order-DEPENDENT number, the multiplicity of the largest element of a full sketch, needs the
// stream positions of the final hashes: each rank's trace pass writes its (global) positions into its own slice of
// every hash's bucket, an all-reduce joins the slices, and sketch_topcount_kernel applies the rule as on one GPU.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) table_export_kernel(const uint64_t* __restrict__ tkeys, const uint32_t* __restrict__ tcnt, const uint64_t* __restrict__ tpos,
                                                           uint32_t cap, uint64_t* __restrict__ out_keys, uint32_t* __restrict__ out_cnt, uint64_t* __restrict__ out_pos,
                                                           uint32_t* counter)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool have = i < cap && tkeys[i] != SK_EMPTY;
    const uint32_t m = __ballot_sync(0xffffffffu, have);
    if (!m) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (have) {
        const uint32_t o = base + __popc(m & ((1u << lane) - 1u));
        out_keys[o] = tkeys[i]; out_cnt[o] = tcnt[i]; out_pos[o] = tpos[i];
    }
}

// entries [rank][stride] of all ranks -> one table: counts add up, first positions take the minimum
__global__ void __launch_bounds__(256) table_import_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ cnt, const uint64_t* __restrict__ pos,
                                                           const uint32_t* __restrict__ n_of_rank, uint32_t stride, uint32_t world, uint64_t* tkeys, uint32_t* tcnt,
                                                           uint64_t* tpos, uint32_t mask, uint32_t* overflow)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)world * stride) return;
    const uint32_t r = (uint32_t)(idx / stride), i = (uint32_t)(idx % stride);
    if (i >= n_of_rank[r]) return;
    const uint64_t h = keys[idx];
    uint32_t slot = table_slot(h, mask);
    for (uint32_t probe = 0; probe <= mask; probe++) {
        const unsigned long long prev = atomicCAS((unsigned long long*)&tkeys[slot], (unsigned long long)SK_EMPTY, (unsigned long long)h);
        if (prev == SK_EMPTY || prev == h) {
            atomicAdd(&tcnt[slot], cnt[idx]);
            atomicMin((unsigned long long*)&tpos[slot], (unsigned long long)pos[idx]);
            return;
        }
        slot = (slot + 1) & mask;
    }
    atomicExch(overflow, 1u);
}

// occurrences of each final hash in THIS rank's table
__global__ void __launch_bounds__(256) table_lookup_counts_kernel(const uint64_t* __restrict__ fin, uint32_t n, const uint64_t* __restrict__ tkeys,
                                                                  const uint32_t* __restrict__ tcnt, uint32_t mask, uint32_t maxkey_cnt, uint32_t* __restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t h = fin[i];
    uint32_t c = 0;
    if (h == SK_EMPTY) c = maxkey_cnt;
    else {
        uint32_t slot = table_slot(h, mask);
        for (uint32_t probe = 0; probe <= mask; probe++) {
            const uint64_t k = tkeys[slot];
            if (k == h) { c = tcnt[slot]; break; }
            if (k == SK_EMPTY) break;
            slot = (slot + 1) & mask;
        }
    }
    out[i] = c;
}

int sketch_reads_sharded_impl(fpm_ctx* ctx, const fpm_sketch_params* p, const uint8_t* d_seq, uint64_t n_bytes, uint64_t* d_out_hashes,
                              uint32_t* d_out_counts, uint32_t* d_out_n, uint64_t* d_out_kmers)
{
    int rc = check_params(p);
    if (rc) return rc;
    if (!ctx->comm || ctx->comm_world < 1) { set_error("no communicator: call fpm_comm_init_rank or fpm_comm_adopt first"); return FPM_ERR_ARG; }
    if (!is_nucleotide(p)) { set_error("a read set spread over several GPUs needs the nucleotide alphabet"); return FPM_ERR_UNSUPPORTED; }
    if (!d_out_hashes || !d_out_n || (!d_seq && n_bytes)) { set_error("NULL buffer"); return FPM_ERR_ARG; }
    if (((uintptr_t)d_seq & 15) != 0) { set_error("d_seq must be 16-byte aligned"); return FPM_ERR_ARG; }
    NcclApi* api = nccl_api();
    if (!api) return FPM_ERR_COMM;
    ncclComm_t comm = (ncclComm_t)ctx->comm;
    const int W = ctx->comm_world, me = ctx->comm_rank;
    cudaStream_t st = ctx->stream;
    const int K = p->kmer_size;
    const uint32_t s = p->sketch_size;
    const bool canon = !p->noncanonical;
    const uint64_t full = p->use64 ? kAll64 : 0xffffffffULL;
    const uint64_t target = s <= 4096 ? 2ull * s + 64 : (uint64_t)s + s / 4 + 256;
    const bool want_counts = p->want_counts && d_out_counts;

    // small exchanges go through one scratch block: [W] x 4 u64 of per-rank metadata, sent from slot `me`
    if ((rc = ctx->d_misc.ensure(64 + (size_t)W * 32))) return rc;
    if ((rc = ctx->ensure_pinned((size_t)W * 32 + 64))) return rc;
    uint64_t* d_meta = ctx->d_misc.as<uint64_t>() + 8;
    uint64_t* h_meta = (uint64_t*)ctx->h_pinned;
    auto gather_meta = [&](uint64_t a0, uint64_t a1, uint64_t a2, uint64_t a3) -> int {
        const uint64_t mine[4] = {a0, a1, a2, a3};
        FPM_CUDA(cudaMemcpyAsync(d_meta + 4 * me, mine, 32, cudaMemcpyHostToDevice, st));
        FPM_NCCL(api, api->AllGather(d_meta + 4 * me, d_meta, 32, ncclUint8, comm, st));
        FPM_CUDA(cudaMemcpyAsync(h_meta, d_meta, (size_t)W * 32, cudaMemcpyDeviceToHost, st));
        FPM_CUDA(cudaStreamSynchronize(st));
        return FPM_OK;
    };
    // ---- where this rank's part sits in the stream ---------------------------------------------------------------
    if ((rc = gather_meta(n_bytes, 0, 0, 0))) return rc;
    uint64_t n_total = 0, base = 0;
    for (int r = 0; r < W; r++) { if (r < me) base += h_meta[4 * r]; n_total += h_meta[4 * r]; }
    if (n_total == 0) {
        FPM_CUDA(cudaMemsetAsync(d_out_n, 0, sizeof(uint32_t), st));
        if (d_out_kmers) FPM_CUDA(cudaMemsetAsync(d_out_kmers, 0, sizeof(uint64_t), st));
        return FPM_OK;
    }
    // ---- plan: as fpm_sketch_batch for one group of n_total bytes ------------------------------------------------
    GroupPlan pl;
    const uint64_t cov_guess = p->min_cov > 1 ? std::max<uint64_t>(1, std::min<uint64_t>(32, (64ull << 20) / (4 * target))) : 1;
    if (n_total <= 2 * target * cov_guess) { pl.all = true; pl.thresh = full; pl.cap = pow2ceil(2 * n_total); }
    else { pl.all = false; pl.thresh = scale_threshold(full, (double)(target * cov_guess) / (double)n_total); pl.cap = pow2ceil(4 * target * cov_guess); }

    const uint64_t h_goff[2] = {0, n_bytes};
    if ((rc = ctx->goff.ensure(16))) return rc;
    if ((rc = ctx->thresh.ensure(8))) return rc;
    if ((rc = ctx->active.ensure(1))) return rc;
    if ((rc = ctx->toff.ensure(8))) return rc;
    if ((rc = ctx->tmask.ensure(4))) return rc;
    if ((rc = ctx->maxcnt.ensure(8))) return rc;
    if ((rc = ctx->maxpos.ensure(16))) return rc;
    if ((rc = ctx->overflow.ensure(8))) return rc;
    if ((rc = ctx->stat.ensure(64))) return rc;
    if ((rc = ctx->args.ensure(sizeof(SketchArgs)))) return rc;
    FPM_CUDA(cudaMemcpyAsync(ctx->goff.p, h_goff, 16, cudaMemcpyHostToDevice, st));
    const uint8_t one = 1;
    const uint64_t zero64 = 0;
    FPM_CUDA(cudaMemcpyAsync(ctx->active.p, &one, 1, cudaMemcpyHostToDevice, st));
    FPM_CUDA(cudaMemcpyAsync(ctx->toff.p, &zero64, 8, cudaMemcpyHostToDevice, st));
    int max_smem = 0;
    FPM_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    if (configure_sketch_select((size_t)SK_SORT_CAP * 8) != 0) { set_error("cudaFuncSetAttribute(select) failed"); return FPM_ERR_CUDA; }

    auto fill_args = [&](SketchArgs& a) {
        memset(&a, 0, sizeof a);
        a.seq = d_seq; a.n_bytes = n_bytes; a.group_off = ctx->goff.as<uint64_t>(); a.n_groups = 1;
        a.thresh = ctx->thresh.as<uint64_t>(); a.active = ctx->active.as<uint8_t>();
        a.sketch_size = s; a.seed = p->seed; a.fold_case = !p->preserve_case; a.hash32 = !p->use64;
        a.c_tbl = 0x54474341u; a.c_add1 = 0x52dce729ULL; a.c_add2 = 0x38495ab5ULL; a.c_add1s = a.c_add1 + 5ull * a.seed;
        a.pos_base = base;
    };
    // the merged table lives in ctx->scratch: keys | counts | positions | maxkey count / position | overflow
    uint32_t local_cap = 0, merged_cap = 0, h_outn = 0, h_topcnt = 0;
    uint64_t *m_keys = nullptr, *m_pos = nullptr;
    uint32_t *m_cnt = nullptr;
    uint32_t h_local_maxcnt = 0;
    for (int pass = 0; pass < 12; pass++) {
        // ---- local pass ---------------------------------------------------------------------------------------------
        local_cap = pl.cap;
        if ((rc = ctx->tkeys.ensure((size_t)local_cap * 8))) return rc;
        if ((rc = ctx->tcnt.ensure((size_t)local_cap * 4))) return rc;
        if ((rc = ctx->tpos.ensure((size_t)local_cap * 8))) return rc;
        FPM_CUDA(cudaMemsetAsync(ctx->tkeys.p, 0xff, (size_t)local_cap * 8, st));
        FPM_CUDA(cudaMemsetAsync(ctx->tcnt.p, 0, (size_t)local_cap * 4, st));
        FPM_CUDA(cudaMemsetAsync(ctx->tpos.p, 0xff, (size_t)local_cap * 8, st));
        FPM_CUDA(cudaMemsetAsync(ctx->maxcnt.p, 0, 8, st));
        FPM_CUDA(cudaMemsetAsync(ctx->maxpos.p, 0xff, 16, st));
        FPM_CUDA(cudaMemsetAsync(ctx->overflow.p, 0, 8, st));
        const uint32_t mask = local_cap - 1;
        FPM_CUDA(cudaMemcpyAsync(ctx->thresh.p, &pl.thresh, 8, cudaMemcpyHostToDevice, st));
        FPM_CUDA(cudaMemcpyAsync(ctx->tmask.p, &mask, 4, cudaMemcpyHostToDevice, st));
        SketchArgs a;
        fill_args(a);
        a.tkeys = ctx->tkeys.as<uint64_t>(); a.tcnt = ctx->tcnt.as<uint32_t>(); a.tpos = ctx->tpos.as<uint64_t>();
        a.toff = ctx->toff.as<uint64_t>(); a.tmask = ctx->tmask.as<uint32_t>();
        a.maxkey_cnt = ctx->maxcnt.as<uint32_t>(); a.maxkey_pos = ctx->maxpos.as<uint64_t>(); a.overflow = ctx->overflow.as<uint32_t>();
        FPM_CUDA(cudaMemcpyAsync(ctx->args.p, &a, sizeof a, cudaMemcpyHostToDevice, st));
        if (n_bytes) {
            ctx->time_begin(FPM_KERNEL_SKETCH_HASH);
            g_hash_launch[K - 1](canon, st, ctx->args.as<SketchArgs>(), 0, n_bytes, 0);
            ctx->time_end();
            ctx->launches++;
            FPM_CUDA(cudaGetLastError());
        }
        // ---- export the candidates, exchange, merge ---------------------------------------------------------------------
        // export buffers: room for every rank's entries, this rank's at slot `me` (the in-place form of ncclAllGather)
        const size_t ek = (size_t)local_cap * 8, ec = (size_t)local_cap * 4;
        if ((rc = ctx->tr_pos.ensure((size_t)W * (2 * ek + ec) + 256))) return rc;
        unsigned char* eb = ctx->tr_pos.as<unsigned char>();
        uint64_t* x_keys = (uint64_t*)eb; uint64_t* x_pos = (uint64_t*)(eb + (size_t)W * ek); uint32_t* x_cnt = (uint32_t*)(eb + (size_t)W * 2 * ek);
        uint32_t* d_counter = ctx->overflow.as<uint32_t>() + 1;
        table_export_kernel<<<(local_cap + 255) / 256, 256, 0, st>>>(a.tkeys, a.tcnt, a.tpos, local_cap, x_keys + (size_t)me * local_cap, x_cnt + (size_t)me * local_cap,
                                                                      x_pos + (size_t)me * local_cap, d_counter);
        ctx->launches++;
        FPM_CUDA(cudaGetLastError());
        uint32_t h4[4] = {0, 0, 0, 0};     // overflow, exported entries, maxkey count
        uint64_t h_maxpos = ~0ULL;
        FPM_CUDA(cudaMemcpyAsync(h4, ctx->overflow.p, 8, cudaMemcpyDeviceToHost, st));
        FPM_CUDA(cudaMemcpyAsync(h4 + 2, ctx->maxcnt.p, 4, cudaMemcpyDeviceToHost, st));
        FPM_CUDA(cudaMemcpyAsync(&h_maxpos, ctx->maxpos.p, 8, cudaMemcpyDeviceToHost, st));
        FPM_CUDA(cudaStreamSynchronize(st));
        h_local_maxcnt = h4[2];
        if ((rc = gather_meta(h4[1], h4[0], h4[2], h_maxpos))) return rc;
        bool too_many = false;
        uint64_t total_entries = 0, max_entries = 0, maxkey_cnt = 0, maxkey_pos = ~0ULL;
        std::vector<uint32_t> n_of(W);
        for (int r = 0; r < W; r++) {
            n_of[r] = (uint32_t)h_meta[4 * r]; total_entries += n_of[r]; max_entries = std::max<uint64_t>(max_entries, n_of[r]);
            too_many |= h_meta[4 * r + 1] != 0; maxkey_cnt += h_meta[4 * r + 2]; maxkey_pos = std::min(maxkey_pos, h_meta[4 * r + 3]);
        }
        uint32_t nq = 0, nd = 0;
        if (!too_many) {
            // all ranks send the same number of elements (the largest export); only the first n_of[r] of each slot count
            FPM_NCCL(api, api->AllGather(x_keys + (size_t)me * local_cap, x_keys, ek, ncclUint8, comm, st));
            FPM_NCCL(api, api->AllGather(x_pos + (size_t)me * local_cap, x_pos, ek, ncclUint8, comm, st));
            FPM_NCCL(api, api->AllGather(x_cnt + (size_t)me * local_cap, x_cnt, ec, ncclUint8, comm, st));
            merged_cap = std::max<uint32_t>(pow2ceil(2 * total_entries + 64), 64);
            const size_t mk = (size_t)merged_cap * 8, mc = (size_t)merged_cap * 4;
            if ((rc = ctx->scratch.ensure(2 * mk + mc + 256 + (size_t)merged_cap * 8))) return rc;
            unsigned char* mb = ctx->scratch.as<unsigned char>();
            m_keys = (uint64_t*)mb; m_pos = (uint64_t*)(mb + mk); m_cnt = (uint32_t*)(mb + 2 * mk);
            uint32_t* m_extra = (uint32_t*)(mb + 2 * mk + mc);           // [0] merged mask, [1] overflow, [2] maxkey count, [4..5] maxkey position, [8..] n_of_rank
            uint64_t* m_sort = (uint64_t*)(mb + 2 * mk + mc + 256);      // select's global sort scratch
            FPM_CUDA(cudaMemsetAsync(m_keys, 0xff, mk, st));
            FPM_CUDA(cudaMemsetAsync(m_pos, 0xff, mk, st));
            FPM_CUDA(cudaMemsetAsync(m_cnt, 0, mc, st));
            uint32_t ex[64];
            memset(ex, 0, sizeof ex);
            if (W > 48) { set_error("more than 48 ranks"); return FPM_ERR_UNSUPPORTED; }
            ex[0] = merged_cap - 1; ex[2] = (uint32_t)maxkey_cnt; memcpy(ex + 4, &maxkey_pos, 8);
            for (int r = 0; r < W; r++) ex[8 + r] = n_of[r];
            FPM_CUDA(cudaMemcpyAsync(m_extra, ex, sizeof ex, cudaMemcpyHostToDevice, st));
            table_import_kernel<<<(uint32_t)(((uint64_t)W * local_cap + 255) / 256), 256, 0, st>>>(x_keys, x_cnt, x_pos, m_extra + 8, local_cap, (uint32_t)W, m_keys, m_cnt, m_pos,
                                                                                                 merged_cap - 1, m_extra + 1);
            ctx->launches++;
            FPM_CUDA(cudaGetLastError());
            // ---- bottom-s of the merged table: the same kernel as on one GPU ----------------------------------------------
            SelectArgs sa;
            sa.tkeys = m_keys; sa.tcnt = m_cnt; sa.tpos = m_pos; sa.toff = ctx->toff.as<uint64_t>(); sa.tmask = m_extra;
            sa.maxkey_cnt = m_extra + 2; sa.maxkey_pos = (const uint64_t*)(m_extra + 4); sa.active = ctx->active.as<uint8_t>();
            sa.sketch_size = s; sa.min_cov = p->min_cov;
            sa.thresh = ctx->thresh.as<uint64_t>();
            sa.sort_cap = std::min<uint32_t>(SK_SORT_CAP, merged_cap);
            sa.scratch = merged_cap > SK_SORT_CAP ? m_sort : nullptr;
            sa.out_hashes = d_out_hashes; sa.out_counts = want_counts ? d_out_counts : nullptr; sa.out_firstpos = nullptr;
            sa.out_n = d_out_n;
            uint32_t* d_stat = ctx->stat.as<uint32_t>();
            sa.stat_nq = d_stat; sa.stat_nd = d_stat + 1; sa.stat_topcnt = d_stat + 2;
            ctx->time_begin(FPM_KERNEL_SKETCH_SELECT);
            launch_sketch_select(1, (size_t)sa.sort_cap * 8, st, sa);
            ctx->time_end();
            ctx->launches++;
            FPM_CUDA(cudaGetLastError());
            uint32_t h_stat[3] = {0, 0, 0};
            FPM_CUDA(cudaMemcpyAsync(h_stat, d_stat, 12, cudaMemcpyDeviceToHost, st));
            FPM_CUDA(cudaStreamSynchronize(st));
            nq = h_stat[0]; nd = h_stat[1]; h_topcnt = h_stat[2];
        }
        // ---- decide: the same rule on every rank, from the same numbers -----------------------------------------------
        const bool too_few = !too_many && !pl.all && nq < s;
        if (!too_many && !too_few) { h_outn = std::min<uint32_t>(nq, s); break; }
        if (pass == 11) { set_error("bottom-s selection did not converge"); return FPM_ERR_CUDA; }
        const double cur = pl.all ? 1.0 : ((double)pl.thresh + 1.0) / ((double)full + 1.0);
        double next;
        if (too_many) next = cur / 4;
        else next = nq >= 16 ? cur * 1.15 * (double)target / (double)nq : cur * 16;
        if (next >= 1.0) { pl.all = true; pl.thresh = full; next = 1.0; }
        else { pl.all = false; pl.thresh = scale_threshold(full, next); }
        double nd_next = ((double)nd + 16) * (next / cur) * 1.3 + 256;
        if (too_many) nd_next = (double)local_cap;
        if (nd_next > (double)n_total) nd_next = (double)n_total;
        pl.cap = std::max(pow2ceil((uint64_t)(2 * nd_next)), pow2ceil(4 * target));
    }

    // ---- order-dependent multiplicity of the largest element (full sketch only) ----------------------------------------
    if (want_counts && h_outn == s && h_topcnt > p->min_cov) {
        std::vector<uint32_t> tot(s);
        FPM_CUDA(cudaMemcpyAsync(tot.data(), d_out_counts, sizeof(uint32_t) * s, cudaMemcpyDeviceToHost, st));
        // this rank's occurrences of every final hash, then everybody's
        if ((rc = ctx->tr_cursor.ensure(sizeof(uint32_t) * ((size_t)W + 2) * s))) return rc;
        uint32_t* d_loc_all = ctx->tr_cursor.as<uint32_t>();               // [W][s]; cursors behind it
        table_lookup_counts_kernel<<<(s + 255) / 256, 256, 0, st>>>(d_out_hashes, s, ctx->tkeys.as<uint64_t>(), ctx->tcnt.as<uint32_t>(), local_cap - 1, h_local_maxcnt,
                                                                    d_loc_all + (size_t)me * s);
        ctx->launches++;
        FPM_CUDA(cudaGetLastError());
        FPM_NCCL(api, api->AllGather(d_loc_all + (size_t)me * s, d_loc_all, (size_t)s * 4, ncclUint8, comm, st));
        std::vector<uint32_t> loc((size_t)W * s);
        FPM_CUDA(cudaMemcpyAsync(loc.data(), d_loc_all, sizeof(uint32_t) * (size_t)W * s, cudaMemcpyDeviceToHost, st));
        FPM_CUDA(cudaStreamSynchronize(st));
        std::vector<uint64_t> off_all(s), off_mine(s);
        std::vector<uint32_t> cap_mine(s);
        uint64_t total = 0;
        for (uint32_t b = 0; b < s; b++) {
            off_all[b] = total;
            uint64_t before = 0, sum = 0;
            for (int r = 0; r < W; r++) { if (r < me) before += loc[(size_t)r * s + b]; sum += loc[(size_t)r * s + b]; }
            if (sum != tot[b]) { set_error("internal: per-rank counts of a sketch hash do not add up"); return FPM_ERR_CUDA; }
            off_mine[b] = total + before; cap_mine[b] = loc[(size_t)me * s + b];
            total += tot[b];
        }
        // buckets: [total] positions (zero = not mine), joined by an all-reduce; offsets of both views; capacities of mine
        if ((rc = ctx->tr_off.ensure(sizeof(uint64_t) * 2 * s + sizeof(uint32_t) * s + 64))) return rc;
        uint64_t* d_off_all = ctx->tr_off.as<uint64_t>(); uint64_t* d_off_mine = d_off_all + s; uint32_t* d_cap_mine = (uint32_t*)(d_off_mine + s);
        if ((rc = ctx->glist.ensure(sizeof(uint64_t) * std::max<uint64_t>(total, 1) + 64))) return rc;
        uint64_t* d_pos = ctx->glist.as<uint64_t>();
        uint32_t* d_cursor = d_loc_all + (size_t)W * s;
        FPM_CUDA(cudaMemcpyAsync(d_off_all, off_all.data(), sizeof(uint64_t) * s, cudaMemcpyHostToDevice, st));
        FPM_CUDA(cudaMemcpyAsync(d_off_mine, off_mine.data(), sizeof(uint64_t) * s, cudaMemcpyHostToDevice, st));
        FPM_CUDA(cudaMemcpyAsync(d_cap_mine, cap_mine.data(), sizeof(uint32_t) * s, cudaMemcpyHostToDevice, st));
        FPM_CUDA(cudaMemsetAsync(d_pos, 0, sizeof(uint64_t) * std::max<uint64_t>(total, 1), st));
        FPM_CUDA(cudaMemsetAsync(d_cursor, 0, sizeof(uint32_t) * s, st));
        FPM_CUDA(cudaMemcpyAsync(ctx->thresh.p, &pl.thresh, 8, cudaMemcpyHostToDevice, st));
        SketchArgs a;
        fill_args(a);
        a.fin_hashes = d_out_hashes; a.fin_n = d_out_n; a.tr_off = d_off_mine; a.tr_cap = d_cap_mine; a.tr_cursor = d_cursor; a.tr_pos = d_pos;
        FPM_CUDA(cudaMemcpyAsync(ctx->args.p, &a, sizeof a, cudaMemcpyHostToDevice, st));
        if (n_bytes) {
            ctx->time_begin(FPM_KERNEL_SKETCH_HASH);
            g_hash_launch[K - 1](canon, st, ctx->args.as<SketchArgs>(), 0, n_bytes, 1);
            ctx->time_end();
            ctx->launches++;
            FPM_CUDA(cudaGetLastError());
        }
        FPM_NCCL(api, api->AllReduce(d_pos, d_pos, total, ncclUint64, ncclSum, comm, st));
        const uint32_t zero = 0;
        uint32_t* d_g0 = ctx->overflow.as<uint32_t>();
        FPM_CUDA(cudaMemcpyAsync(d_g0, &zero, 4, cudaMemcpyHostToDevice, st));
        launch_sketch_topcount(1, st, d_g0, s, p->min_cov, d_off_all, d_out_counts, d_pos, d_out_counts);
        ctx->launches++;
        FPM_CUDA(cudaGetLastError());
    }
    if (d_out_kmers) {
        FPM_CUDA(cudaMemsetAsync(d_out_kmers, 0, sizeof(uint64_t), st));
        if (n_bytes) {
            SketchArgs a;
            memset(&a, 0, sizeof a);
            a.seq = d_seq; a.n_bytes = n_bytes; a.group_off = ctx->goff.as<uint64_t>(); a.n_groups = 1;
            a.fold_case = !p->preserve_case;
            FPM_CUDA(cudaMemcpyAsync(ctx->args.p, &a, sizeof a, cudaMemcpyHostToDevice, st));
            g_count_launch[K - 1]((uint32_t)((n_bytes + SK_TILE_WINDOWS - 1) / SK_TILE_WINDOWS), st, ctx->args.as<SketchArgs>(), (unsigned long long*)d_out_kmers);
            ctx->launches++;
            FPM_CUDA(cudaGetLastError());
        }
        FPM_NCCL(api, api->AllReduce(d_out_kmers, d_out_kmers, 1, ncclUint64, ncclSum, comm, st));
    }
    FPM_CUDA(cudaStreamSynchronize(st));
    return FPM_OK;
}

}  // namespace fpm

using namespace fpm;

extern "C" {

int fpm_sketch_batch_dev(fpm_ctx* ctx, const fpm_sketch_params* p, const uint8_t* d_seq, uint64_t seq_bytes,
                         const uint64_t* h_group_offsets, uint32_t n_groups, uint64_t* d_out_hashes, uint32_t* d_out_counts,
                         uint32_t* d_out_n, uint64_t* d_out_kmers)
{
    if (!ctx) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    FPM_CUDA(cudaSetDevice(ctx->device));
    return sketch_batch_dev_impl(ctx, p, d_seq, seq_bytes, h_group_offsets, n_groups, d_out_hashes, d_out_counts, d_out_n, d_out_kmers);
}

int fpm_sketch_reads_sharded_dev(fpm_ctx* ctx, const fpm_sketch_params* p, const uint8_t* d_seq, uint64_t seq_bytes, uint64_t* d_out_hashes,
                                 uint32_t* d_out_counts, uint32_t* d_out_n, uint64_t* d_out_kmers)
{
    if (!ctx) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    FPM_CUDA(cudaSetDevice(ctx->device));
    return sketch_reads_sharded_impl(ctx, p, d_seq, seq_bytes, d_out_hashes, d_out_counts, d_out_n, d_out_kmers);
}

int fpm_sketch_batch(fpm_ctx* ctx, const fpm_sketch_params* p, const uint8_t* seq, uint64_t seq_bytes,
                     const uint64_t* group_offsets, uint32_t n_groups, uint64_t* out_hashes, uint32_t* out_counts,
                     uint32_t* out_n, uint64_t* out_kmers)
{
    if (!ctx) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    int rc = check_params(p);
    if (rc) return rc;
    if (n_groups == 0) return FPM_OK;
    if (!seq && seq_bytes) { set_error("seq is NULL"); return FPM_ERR_ARG; }
    FPM_CUDA(cudaSetDevice(ctx->device));
    const uint64_t s = p->sketch_size;
    if ((rc = ctx->outh.ensure(sizeof(uint64_t) * n_groups * s))) return rc;
    if ((rc = ctx->outc.ensure(sizeof(uint32_t) * n_groups * s))) return rc;
    if ((rc = ctx->outn.ensure(sizeof(uint32_t) * n_groups))) return rc;
    if ((rc = ctx->outk.ensure(sizeof(uint64_t) * n_groups))) return rc;
    if (!group_offsets || group_offsets[0] != 0 || group_offsets[n_groups] != seq_bytes) { set_error("group_offsets must start at 0 and end at seq_bytes"); return FPM_ERR_ARG; }
    bool counts = p->want_counts && out_counts;

    // Chunks of whole sketches, ~kChunkBytes each.  Chunk c+1 is copied on a second stream while
    // chunk c is being hashed, so with pinned host memory the call costs max(H2D, kernels), not
    // their sum.  Each chunk is placed 256-byte aligned on the device and sketched as a batch of
    // its own (positions only need to be ordered within a sketch).
    const uint64_t kChunkBytes = 256ull << 20;
    std::vector<uint32_t> cut{0};
    for (uint32_t g = 0; g < n_groups; g++)
        if (group_offsets[g + 1] - group_offsets[cut.back()] >= kChunkBytes && g + 1 < n_groups) cut.push_back(g + 1);
    cut.push_back(n_groups);
    const size_t n_chunks = cut.size() - 1;
    std::vector<uint64_t> dev_off(n_chunks);
    uint64_t dev_total = 0;
    for (size_t c = 0; c < n_chunks; c++) {
        dev_off[c] = dev_total;
        dev_total += ((group_offsets[cut[c + 1]] - group_offsets[cut[c]]) + 64 + 255) & ~255ull;
    }
    if ((rc = ctx->seq.ensure(dev_total + 64))) return rc;
    if (n_chunks > 1 && !ctx->copy_stream) {
        FPM_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        FPM_CUDA(cudaEventCreateWithFlags(&ctx->copy_done[0], cudaEventDisableTiming));
        FPM_CUDA(cudaEventCreateWithFlags(&ctx->copy_done[1], cudaEventDisableTiming));
    }
    auto enqueue_copy = [&](size_t c) -> int {
        uint64_t h0 = group_offsets[cut[c]], len = group_offsets[cut[c + 1]] - h0;
        cudaStream_t cs = n_chunks > 1 ? ctx->copy_stream : ctx->stream;
        if (len) FPM_CUDA(cudaMemcpyAsync(ctx->seq.as<uint8_t>() + dev_off[c], seq + h0, len, cudaMemcpyHostToDevice, cs));
        if (n_chunks > 1) FPM_CUDA(cudaEventRecord(ctx->copy_done[c & 1], cs));
        return FPM_OK;
    };
    // Any exit below, failure included, first drains the copy stream: a copy of the caller's host buffer may still be in
    // flight there, and the caller is free to release that buffer as soon as this function returns.
    auto drained = [&](int code) -> int {
        if (n_chunks > 1 && ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
        return code;
    };
    if ((rc = enqueue_copy(0))) return drained(rc);
    std::vector<uint64_t> goff;
    rc = FPM_OK;
    for (size_t c = 0; c < n_chunks && rc == FPM_OK; c++) {
        const uint32_t g0 = cut[c], ng = cut[c + 1] - cut[c];
        if (n_chunks > 1) {
            cudaError_t e = cudaStreamWaitEvent(ctx->stream, ctx->copy_done[c & 1], 0);
            if (e != cudaSuccess) return drained(cuda_fail(e, "cudaStreamWaitEvent(copy_done)", __FILE__, __LINE__));
        }
        // the event of chunk c+1 reuses slot (c+1)&1, last used by chunk c-1 whose wait was already enqueued
        if (c + 1 < n_chunks && (rc = enqueue_copy(c + 1))) return drained(rc);
        goff.resize(ng + 1);
        for (uint32_t g = 0; g <= ng; g++) goff[g] = group_offsets[g0 + g] - group_offsets[g0];
        rc = sketch_batch_dev_impl(ctx, p, ctx->seq.as<uint8_t>() + dev_off[c], goff[ng], goff.data(), ng,
                                   ctx->outh.as<uint64_t>() + (uint64_t)g0 * s, counts ? ctx->outc.as<uint32_t>() + (uint64_t)g0 * s : nullptr,
                                   ctx->outn.as<uint32_t>() + g0, out_kmers ? ctx->outk.as<uint64_t>() + g0 : nullptr);
    }
    if (rc) return drained(rc);
    FPM_CUDA(cudaMemcpyAsync(out_hashes, ctx->outh.p, sizeof(uint64_t) * n_groups * s, cudaMemcpyDeviceToHost, ctx->stream));
    if (counts) FPM_CUDA(cudaMemcpyAsync(out_counts, ctx->outc.p, sizeof(uint32_t) * n_groups * s, cudaMemcpyDeviceToHost, ctx->stream));
    FPM_CUDA(cudaMemcpyAsync(out_n, ctx->outn.p, sizeof(uint32_t) * n_groups, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_kmers) FPM_CUDA(cudaMemcpyAsync(out_kmers, ctx->outk.p, sizeof(uint64_t) * n_groups, cudaMemcpyDeviceToHost, ctx->stream));
    FPM_CUDA(cudaStreamSynchronize(ctx->stream));
    return FPM_OK;
}

int fpm_sketch_parsed(fpm_ctx* ctx, const fpm_sketch_params* p, const uint64_t* group_offsets, uint32_t n_groups, uint64_t* out_hashes,
                      uint32_t* out_counts, uint32_t* out_n)
{
    if (!ctx) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    int rc = check_params(p);
    if (rc) return rc;
    if (n_groups == 0) return FPM_OK;
    if (!group_offsets || group_offsets[0] != 0 || group_offsets[n_groups] != ctx->fa_seq_bytes) {
        set_error("group_offsets must start at 0 and end at the parsed sequence's size");
        return FPM_ERR_ARG;
    }
    FPM_CUDA(cudaSetDevice(ctx->device));
    const uint64_t s = p->sketch_size;
    if ((rc = ctx->outh.ensure(sizeof(uint64_t) * n_groups * s))) return rc;
    if ((rc = ctx->outc.ensure(sizeof(uint32_t) * n_groups * s))) return rc;
    if ((rc = ctx->outn.ensure(sizeof(uint32_t) * n_groups))) return rc;
    const bool counts = p->want_counts && out_counts;
    rc = sketch_batch_dev_impl(ctx, p, ctx->fa_seq.as<uint8_t>(), ctx->fa_seq_bytes, group_offsets, n_groups, ctx->outh.as<uint64_t>(),
                               counts ? ctx->outc.as<uint32_t>() : nullptr, ctx->outn.as<uint32_t>(), nullptr);
    if (rc) return rc;
    FPM_CUDA(cudaMemcpyAsync(out_hashes, ctx->outh.p, sizeof(uint64_t) * n_groups * s, cudaMemcpyDeviceToHost, ctx->stream));
    if (counts) FPM_CUDA(cudaMemcpyAsync(out_counts, ctx->outc.p, sizeof(uint32_t) * n_groups * s, cudaMemcpyDeviceToHost, ctx->stream));
    FPM_CUDA(cudaMemcpyAsync(out_n, ctx->outn.p, sizeof(uint32_t) * n_groups, cudaMemcpyDeviceToHost, ctx->stream));
    FPM_CUDA(cudaStreamSynchronize(ctx->stream));
    return FPM_OK;
}

// ---- streaming input: sequence arrives in pieces, is accumulated in HBM, sketched once at the end -------
// (read sets far larger than any pinned staging buffer: `mash sketch -r` reads everything into ONE sketch)

int fpm_sketch_stream_begin(fpm_ctx* ctx)
{
    if (!ctx) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    ctx->stream_used = 0;
    ctx->stream_goff.assign(1, 0);
    return FPM_OK;
}

int fpm_sketch_stream_append(fpm_ctx* ctx, const uint8_t* seq, uint64_t bytes)
{
    if (!ctx) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    if (ctx->stream_goff.empty()) { set_error("fpm_sketch_stream_begin was not called"); return FPM_ERR_ARG; }
    if (bytes == 0) return FPM_OK;
    FPM_CUDA(cudaSetDevice(ctx->device));
    int rc = fpm::stream_reserve(ctx, bytes);
    if (rc) return rc;
    FPM_CUDA(cudaMemcpyAsync((uint8_t*)ctx->stream_buf.p + ctx->stream_used, seq, bytes, cudaMemcpyHostToDevice, ctx->stream));
    FPM_CUDA(cudaStreamSynchronize(ctx->stream));   // the caller may reuse its staging buffer right away
    ctx->stream_used += bytes;
    return FPM_OK;
}

// The compacted sequence the last fpm_fasta_parse left on the device (records back to back, each followed by 0x00 -- the
// stream's own layout) goes to the end of the stream: FASTA read sets never cross PCIe twice.
int fpm_sketch_stream_append_parsed(fpm_ctx* ctx)
{
    if (!ctx) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    if (ctx->stream_goff.empty()) { set_error("fpm_sketch_stream_begin was not called"); return FPM_ERR_ARG; }
    const uint64_t bytes = ctx->fa_seq_bytes;
    if (bytes == 0) return FPM_OK;
    FPM_CUDA(cudaSetDevice(ctx->device));
    int rc = fpm::stream_reserve(ctx, bytes);
    if (rc) return rc;
    FPM_CUDA(cudaMemcpyAsync((uint8_t*)ctx->stream_buf.p + ctx->stream_used, ctx->fa_seq.p, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->stream_used += bytes;
    return FPM_OK;
}

// The same without waiting for the copy: the piece must stay untouched until fpm_sketch_stream_wait(ticket).  With two staging
// buffers the host reads the next piece of the file while this one crosses PCIe.
int fpm_sketch_stream_append_async(fpm_ctx* ctx, const uint8_t* seq, uint64_t bytes, uint64_t* ticket)
{
    if (!ctx || !ticket) { set_error("NULL argument"); return FPM_ERR_ARG; }
    if (ctx->stream_goff.empty()) { set_error("fpm_sketch_stream_begin was not called"); return FPM_ERR_ARG; }
    FPM_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->copy_done[0]) {
        if (!ctx->copy_stream) FPM_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        FPM_CUDA(cudaEventCreateWithFlags(&ctx->copy_done[0], cudaEventDisableTiming));
        FPM_CUDA(cudaEventCreateWithFlags(&ctx->copy_done[1], cudaEventDisableTiming));
    }
    int rc = fpm::stream_reserve(ctx, bytes);      // (growing the HBM buffer waits for the copies in flight: same stream)
    if (rc) return rc;
    if (bytes) FPM_CUDA(cudaMemcpyAsync((uint8_t*)ctx->stream_buf.p + ctx->stream_used, seq, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->stream_used += bytes;
    *ticket = ctx->stream_tickets++;
    FPM_CUDA(cudaEventRecord(ctx->copy_done[*ticket & 1], ctx->stream));
    return FPM_OK;
}

int fpm_sketch_stream_wait(fpm_ctx* ctx, uint64_t ticket)
{
    if (!ctx) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    if (ticket >= ctx->stream_tickets) { set_error("unknown ticket"); return FPM_ERR_ARG; }
    if (ticket + 2 < ctx->stream_tickets) return FPM_OK;                  // two newer copies were queued behind it on the same stream
    FPM_CUDA(cudaEventSynchronize(ctx->copy_done[ticket & 1]));
    return FPM_OK;
}

int fpm_sketch_stream_end_group(fpm_ctx* ctx)
{
    if (!ctx) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    if (ctx->stream_goff.empty()) { set_error("fpm_sketch_stream_begin was not called"); return FPM_ERR_ARG; }
    ctx->stream_goff.push_back(ctx->stream_used);
    return FPM_OK;
}

int fpm_sketch_stream_finish(fpm_ctx* ctx, const fpm_sketch_params* p, uint64_t* out_hashes, uint32_t* out_counts, uint32_t* out_n, uint64_t* out_kmers)
{
    if (!ctx) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    int rc = check_params(p);
    if (rc) return rc;
    if (ctx->stream_goff.empty()) { set_error("fpm_sketch_stream_begin was not called"); return FPM_ERR_ARG; }
    if (ctx->stream_goff.back() != ctx->stream_used) { set_error("the last group was not closed with fpm_sketch_stream_end_group"); return FPM_ERR_ARG; }
    const uint32_t n_groups = (uint32_t)ctx->stream_goff.size() - 1;
    if (n_groups == 0) return FPM_OK;
    FPM_CUDA(cudaSetDevice(ctx->device));
    const uint64_t s = p->sketch_size;
    if ((rc = ctx->outh.ensure(sizeof(uint64_t) * n_groups * s))) return rc;
    if ((rc = ctx->outc.ensure(sizeof(uint32_t) * n_groups * s))) return rc;
    if ((rc = ctx->outn.ensure(sizeof(uint32_t) * n_groups))) return rc;
    if ((rc = ctx->outk.ensure(sizeof(uint64_t) * n_groups))) return rc;
    const bool counts = p->want_counts && out_counts;
    rc = sketch_batch_dev_impl(ctx, p, (const uint8_t*)ctx->stream_buf.p, ctx->stream_used, ctx->stream_goff.data(), n_groups,
                               ctx->outh.as<uint64_t>(), counts ? ctx->outc.as<uint32_t>() : nullptr, ctx->outn.as<uint32_t>(),
                               out_kmers ? ctx->outk.as<uint64_t>() : nullptr);
    if (rc) return rc;
    FPM_CUDA(cudaMemcpyAsync(out_hashes, ctx->outh.p, sizeof(uint64_t) * n_groups * s, cudaMemcpyDeviceToHost, ctx->stream));
    if (counts) FPM_CUDA(cudaMemcpyAsync(out_counts, ctx->outc.p, sizeof(uint32_t) * n_groups * s, cudaMemcpyDeviceToHost, ctx->stream));
    FPM_CUDA(cudaMemcpyAsync(out_n, ctx->outn.p, sizeof(uint32_t) * n_groups, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_kmers) FPM_CUDA(cudaMemcpyAsync(out_kmers, ctx->outk.p, sizeof(uint64_t) * n_groups, cudaMemcpyDeviceToHost, ctx->stream));
    FPM_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stream_goff.clear();
    return FPM_OK;
}

int fpm_kmer_hashes(fpm_ctx* ctx, const fpm_sketch_params* p, const uint8_t* seq, uint64_t seq_bytes, uint64_t* out_hashes, uint64_t* out_count)
{
    if (!ctx) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    int rc = check_params(p);
    if (rc) return rc;
    if (out_count) *out_count = 0;
    if (!is_nucleotide(p)) { set_error("fpm_kmer_hashes covers the nucleotide kernel only"); return FPM_ERR_UNSUPPORTED; }
    if (seq_bytes == 0) return FPM_OK;
    FPM_CUDA(cudaSetDevice(ctx->device));
    if ((rc = ctx->seq.ensure(seq_bytes + 64))) return rc;
    if ((rc = ctx->outh.ensure(sizeof(uint64_t) * seq_bytes))) return rc;
    if ((rc = ctx->outc.ensure(seq_bytes))) return rc;
    FPM_CUDA(cudaMemcpyAsync(ctx->seq.p, seq, seq_bytes, cudaMemcpyHostToDevice, ctx->stream));
    uint32_t grid = (uint32_t)((seq_bytes + SK_TILE_WINDOWS - 1) / SK_TILE_WINDOWS);
    g_stream_launch[p->kmer_size - 1](!p->noncanonical, grid, ctx->stream, ctx->seq.as<uint8_t>(), seq_bytes, p->seed,
                                      !p->preserve_case, !p->use64, ctx->outh.as<uint64_t>(), ctx->outc.as<uint8_t>());
    ctx->launches++;
    FPM_CUDA(cudaGetLastError());
    std::vector<uint64_t> h(seq_bytes);
    std::vector<uint8_t> v(seq_bytes);
    FPM_CUDA(cudaMemcpyAsync(h.data(), ctx->outh.p, sizeof(uint64_t) * seq_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    FPM_CUDA(cudaMemcpyAsync(v.data(), ctx->outc.p, seq_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    FPM_CUDA(cudaStreamSynchronize(ctx->stream));
    uint64_t n = 0;
    for (uint64_t i = 0; i < seq_bytes; i++)
        if (v[i]) out_hashes[n++] = h[i];
    if (out_count) *out_count = n;
    return FPM_OK;
}

int fpm_fp_hash_batch(fpm_ctx* ctx, const uint64_t* tokens, const uint64_t* line_offsets, uint64_t n_lines, uint32_t seed,
                      int use64, uint64_t* out_hashes)
{
    if (!ctx) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    if (n_lines == 0) return FPM_OK;
    if (!line_offsets || !out_hashes) { set_error("NULL buffer"); return FPM_ERR_ARG; }
    FPM_CUDA(cudaSetDevice(ctx->device));
    uint64_t n_tok = line_offsets[n_lines];
    int rc;
    if ((rc = ctx->seq.ensure(sizeof(uint64_t) * std::max<uint64_t>(n_tok, 1)))) return rc;
    if ((rc = ctx->goff.ensure(sizeof(uint64_t) * (n_lines + 1)))) return rc;
    if ((rc = ctx->outh.ensure(sizeof(uint64_t) * n_lines))) return rc;
    if (n_tok) FPM_CUDA(cudaMemcpyAsync(ctx->seq.p, tokens, sizeof(uint64_t) * n_tok, cudaMemcpyHostToDevice, ctx->stream));
    FPM_CUDA(cudaMemcpyAsync(ctx->goff.p, line_offsets, sizeof(uint64_t) * (n_lines + 1), cudaMemcpyHostToDevice, ctx->stream));
    launch_fp_hash(n_lines, ctx->stream, ctx->seq.as<uint64_t>(), ctx->goff.as<uint64_t>(), seed, use64, ctx->outh.as<uint64_t>());
    ctx->launches++;
    FPM_CUDA(cudaGetLastError());
    FPM_CUDA(cudaMemcpyAsync(out_hashes, ctx->outh.p, sizeof(uint64_t) * n_lines, cudaMemcpyDeviceToHost, ctx->stream));
    FPM_CUDA(cudaStreamSynchronize(ctx->stream));
    return FPM_OK;
}

}  // extern "C"
