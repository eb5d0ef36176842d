// dist_internal.h -- the pieces of dist.cu that dist_multi.cu (several GPUs) builds on.
#pragma once
#include "common.h"

namespace fpm {

// fpm_dist_hits: passing pairs are appended to a list instead of written to an n x n matrix.
struct HitSink {
    fpm_hit* buf = nullptr;              // nullptr: matrix mode
    unsigned long long* count = nullptr; // all passing pairs, also those beyond cap (the caller retries with room for them)
    unsigned long long cap = 0;
    int skip_unmarked = 0;               // the filters exclude distance 1: a pair without a shared hash cannot pass
    uint32_t q_base = 0, r_base = 0;     // added to the indices written into the records (a rank's block of larger panels)
};

// The comparison proper on device-resident panels (dist.cu).  h_out (nullable): host destination, row pitch h_ld records
// (0 = d_ref->n).  hits (nullable): hits mode.
int run_dist(fpm_ctx* ctx, const fpm_dist_params* p, const fpm_panel* d_ref, const fpm_panel* d_qry, fpm_pair* d_out, uint64_t* d_steps,
             uint32_t max_size_ref, uint32_t max_size_qry, fpm_pair* h_out = nullptr, const HitSink* hits = nullptr, uint64_t h_ld = 0);
int max_size_dev(fpm_ctx* ctx, const fpm_panel* d, uint32_t* out);
int check_dist(const fpm_dist_params* p, const fpm_panel* ref, const fpm_panel* qry);
int dist_hits_run(fpm_ctx* ctx, const fpm_dist_params* p, const fpm_panel* d_ref, const fpm_panel* d_qry, uint32_t mr, uint32_t mq, fpm_hit* d_sorted,
                  uint64_t capacity, uint64_t* n_hits, uint64_t* d_steps, uint32_t q_base = 0, uint32_t r_base = 0);

}  // namespace fpm
