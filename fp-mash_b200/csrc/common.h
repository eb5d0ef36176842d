// common.h -- context, error handling and scratch memory shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/fpmash_b200.h"

namespace fpm {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define FPM_CUDA(call)                                                          \
    do {                                                                        \
        cudaError_t e__ = (call);                                               \
        if (e__ != cudaSuccess) return fpm::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

// Grow-only device scratch buffer.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);
    void release();
    template <typename T> T* as() const { return (T*)p; }
};

}  // namespace fpm

struct fpm_ctx;
namespace fpm {
int stream_reserve(fpm_ctx* ctx, uint64_t bytes);   // sketch_host.cu
}

struct fpm_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    cudaStream_t copy_stream = nullptr;          // H2D of the next chunk while the current one is hashed
    cudaEvent_t copy_done[2] = {nullptr, nullptr};
    uint64_t launches = 0;
    int sm_count = 0;
    double int_peak[3] = {0, 0, 0};              // last fpm_measure_int32_peak: ALU-only, IMAD-only, alternating
    // sketch scratch
    fpm::DevBuf seq, goff, thresh, active, toff, tmask, tkeys, tcnt, tpos, maxcnt, maxpos, overflow,
        stat, tiles, args, alpha, scratch, outh, outc, outn, outk, firstpos, tr_off, tr_cursor, tr_pos, glist;
    // streaming sketch input (fpm_sketch_stream_*): sequence accumulated in HBM + group ends
    fpm::DevBuf stream_buf;
    uint64_t stream_used = 0;
    uint64_t stream_tickets = 0;                 // fpm_sketch_stream_append_async: copies queued so far (event = ticket & 1)
    std::vector<uint64_t> stream_goff;
    // dist scratch
    fpm::DevBuf d_ref, d_qry, d_rs, d_qs, d_rl, d_ql, d_out, d_misc, d_p32, d_rank;
    // FASTA ingestion (fpm_fasta_parse): raw bytes, compacted sequence, per-chunk scan arrays, record table
    fpm::DevBuf fa_raw, fa_seq, fa_chunk, fa_recs;
    uint64_t fa_records = 0, fa_seq_bytes = 0;
    uint64_t fa_resident = 0;                    // bytes of the raw batch in fa_raw (uploaded by fpm_fasta_parse or inflated by fpm_gunzip_batch)
    fpm::DevBuf gz_in, gz_meta;                  // gunzip.cu: compressed bytes; offsets / sizes / status per file
    fpm::DevBuf d_post, d_marks, d_group;        // dist pruning: posting lists, per-query reference bitmaps, grouped copies
    fpm::DevBuf d_tiles;                         // dist: list of the tiles that hold work
    fpm::DevBuf d_hits, d_hsort;                 // fpm_dist_hits: appended hits, sort keys / sorted copy
    // several GPUs, one process each (dist_multi.cu): NCCL communicator handed in or created from a unique id
    void* comm = nullptr;                        // ncclComm_t
    int comm_rank = 0, comm_world = 0;
    bool comm_owned = false;
    fpm::DevBuf d_xq, d_xr, d_xg, d_xg2;              // this rank's query / reference block after the exchange step; all-gather staging
    // resident reference panel (fpm_dist_set_reference): its device copy and its index survive across query chunks
    bool ref_set = false;
    fpm_panel ref_dev = {nullptr, nullptr, nullptr, 0, 0};
    uint32_t ref_max = 0;
    struct RefIndexState {                       // what dist_rank_panels built for the reference panel (dist_rank.cu)
        bool valid = false, unsorted = false, uf_valid = false;
        const void* hashes = nullptr;            // identity of the panel it was built for
        uint64_t n_r = 0, rows_r = 0, m = 0, pr = 0;
        const uint64_t* ks = nullptr;            // sorted keys
        const uint32_t* rank = nullptr;          // dense rank of every sorted element
        uint32_t n_buckets = 0;
    } rix;
    fpm::DevBuf d_codes;                         // query codes, row-major [n_q][code_stride] (dist_rank.cu)
    uint64_t code_stride = 0;
    fpm::DevBuf d_p32q, d_uf;                    // packed query tiles; union-find parents / component sizes of the references
    bool no_dist_prune = false;                  // tests: merge every pair
    bool no_dist_group = false;                  // tests: prune, but leave the panels in their own order
    bool force_dist_group = false;               // tests: always reorder the panels (default: only when the marked pairs are scattered)
    bool force_dist_saturate = false;            // tests: always bound the marking walks by the components' sizes
    bool force_dist64 = false;                   // tests: run the 64-bit tile kernel although the 32-bit rank path applies
    // optional per-kernel event timing (bench roofline): pairs of events around each launch
    bool timing = false;
    struct Timed { int id; cudaEvent_t e0, e1; };
    std::vector<Timed> pending;
    double kernel_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint64_t kernel_launches[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    void time_begin(int id);
    void time_end();
    void time_resolve();
    // pinned staging for small host<->device exchanges
    void* h_pinned = nullptr;
    size_t h_pinned_cap = 0;
    int ensure_pinned(size_t bytes);
};
