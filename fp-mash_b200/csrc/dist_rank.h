// dist_rank.h -- see dist_rank.cu
#pragma once
#include "common.h"

namespace fpm {

// Fills ctx->d_p32 with both panels as column tiles [ceil(n/16)][rows][16] of 32-bit dense ranks (0xffffffff
// beyond each sketch's size).  *mode says whether the rank path can be used: not with more than 2^31 elements (the
// caller then runs the 64-bit tile kernel) or with input that is not strictly ascending (literal kernel).
enum { DIST_RANK_OK = 0, DIST_RANK_TOO_BIG = 1, DIST_RANK_UNSORTED = 2 };
int dist_rank_panels(fpm_ctx* ctx, const fpm_panel* d_ref, const fpm_panel* d_qry, uint32_t max_size_ref, uint32_t max_size_qry,
                     uint64_t rows_r, uint64_t rows_q, uint32_t sketch_size, uint32_t** packed_ref, uint32_t** packed_qry, int* mode, uint32_t** marks);
// *marks (nullable result): per query one row of ceil(n_ref / 32) words, bit r set iff reference r shares at least one hash
// with the query; nullptr when pruning is off or would not pay.

// Reorders both panels so that related sketches are neighbours (see dist_rank.cu): replaces *p32r / *p32q / *marks by
// their permuted versions and returns the permutations (new index -> original sketch index, device arrays).
int dist_group_panels(fpm_ctx* ctx, uint64_t n_q, uint64_t n_r, uint64_t rows_q, uint64_t rows_r, const uint32_t* d_sizes_q, uint32_t** p32r, uint32_t** p32q,
                      uint32_t** marks, uint32_t** perm_q, uint32_t** perm_r);
// dist_rank_panels leaves the query codes row-major (ctx->d_codes); this writes their column tiles in natural order.  (Grouping
// writes them in grouped order itself.)
int dist_pack_queries(fpm_ctx* ctx, const uint32_t* d_sizes_q, uint64_t n_q, uint64_t rows_q, uint32_t** p32q);

// The 32 x 32 tiles (in the panels' current order: perm_q / perm_r after dist_group_panels, null = natural order) that hold a
// marked pair, or an empty query together with an empty reference (such a pair has distance 0 without sharing a hash), as
// (query tile, reference tile) in device memory; *n_pairs (nullable) = marked pairs in all.  n_listed * 1024 against n_pairs
// says how scattered the marked pairs are: run_dist regroups the panels only when that pays.
int dist_tile_list(fpm_ctx* ctx, const uint32_t* marks, uint64_t n_q, uint64_t n_r, const uint32_t* perm_q, const uint32_t* perm_r,
                   const uint32_t* size_q, const uint32_t* size_r, const uint2** list, uint32_t* n_listed, uint64_t* n_pairs);

// fpm_dist_hits: `n` hits appended in arbitrary order at `in` -> `out` sorted by (query, ref).  in != out.
int dist_sort_hits(fpm_ctx* ctx, const fpm_hit* in, uint64_t n, uint64_t n_qry, uint64_t n_ref, fpm_hit* out, uint32_t q_base = 0, uint32_t r_base = 0);
// (hit records carry indices q_base + q, r_base + r when the call compares one block of larger panels)

}  // namespace fpm
