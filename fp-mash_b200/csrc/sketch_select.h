// sketch_select.h -- argument block + launchers of the selection kernels (sketch_select.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fpm {

struct SelectArgs {
    const uint64_t* tkeys;
    const uint32_t* tcnt;
    const uint64_t* tpos;
    const uint64_t* toff;
    const uint32_t* tmask;
    const uint32_t* maxkey_cnt;
    const uint64_t* maxkey_pos;
    const uint8_t* active;
    uint32_t sketch_size;
    uint32_t min_cov;
    uint32_t sort_cap;       // u64 keys that fit the dynamic shared memory of this launch
    const uint64_t* thresh;  // [n_groups] the bound the pass admitted hashes under (nullable): lets a sketch with many more
                             // qualifying hashes than slots sort only those below a cut instead of all of them
    uint64_t* scratch;       // table-shaped global scratch for sketches with more qualifying keys (nullable)
    uint64_t* out_hashes;    // [n_groups][s]
    uint32_t* out_counts;    // [n_groups][s] (nullable)
    uint64_t* out_firstpos;  // [n_groups][s] (nullable)
    uint32_t* out_n;         // [n_groups]
    uint32_t* stat_nq;       // [n_groups] qualifying distinct (count >= min_cov)
    uint32_t* stat_nd;       // [n_groups] distinct
    uint32_t* stat_topcnt;   // [n_groups] total count of the largest output hash
};

void launch_sketch_select(uint32_t n_groups, size_t smem_bytes, cudaStream_t st, const SelectArgs& a);
int configure_sketch_select(size_t max_smem_bytes);
void launch_sketch_topcount(uint32_t n_list, cudaStream_t st, const uint32_t* d_groups, uint32_t sketch_size, uint32_t min_cov,
                            const uint64_t* tr_off, const uint32_t* tr_cap, const uint64_t* tr_pos, uint32_t* out_counts);
struct SketchArgs;
// the trace pass from the survivor log: every logged (hash, position) of the listed sketches goes through sketch_emit's trace branch
void launch_sketch_trace_log(uint64_t n_log, cudaStream_t st, const SketchArgs* d_args);
void launch_sketch_generic(cudaStream_t st, const SketchArgs* d_args, const uint8_t* d_alphabet, int K, uint64_t range_lo,
                           uint64_t range_hi, int mode, unsigned long long* out_kmers);
void launch_fp_hash(uint64_t n_lines, cudaStream_t st, const uint64_t* tokens, const uint64_t* line_off, uint32_t seed, int use64, uint64_t* out);

}  // namespace fpm
