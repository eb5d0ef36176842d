// sketch_inst.cu -- compiled once per k-mer size (-DFPM_K=1..32) so every k gets fully
// specialised code (constant shifts, static Murmur block/tail structure) and the 32 objects
// build in parallel.
#include "sketch_kernels.cuh"
#include "sketch_launch.h"

#ifndef FPM_K
#error "compile with -DFPM_K=<k>"
#endif

namespace fpm {

#define FPM_CAT2(a, b) a##b
#define FPM_CAT(a, b) FPM_CAT2(a, b)

void FPM_CAT(launch_sketch_hash_k, FPM_K)(bool canon, uint32_t grid, cudaStream_t st, const SketchArgs* d_args, int trace)
{
    if (canon) sketch_hash_kernel<FPM_K, true><<<grid, SK_THREADS, 0, st>>>(d_args, trace);
    else sketch_hash_kernel<FPM_K, false><<<grid, SK_THREADS, 0, st>>>(d_args, trace);
}

void FPM_CAT(launch_hash_stream_k, FPM_K)(bool canon, uint32_t grid, cudaStream_t st, const uint8_t* seq, uint64_t n, uint32_t seed,
                                          int fold, int hash32, uint64_t* out, uint8_t* valid)
{
    if (canon) kmer_hash_stream_kernel<FPM_K, true><<<grid, SK_THREADS, 0, st>>>(seq, n, seed, fold, hash32, out, valid);
    else kmer_hash_stream_kernel<FPM_K, false><<<grid, SK_THREADS, 0, st>>>(seq, n, seed, fold, hash32, out, valid);
}

void FPM_CAT(launch_count_windows_k, FPM_K)(uint32_t grid, cudaStream_t st, const SketchArgs* d_args, unsigned long long* out)
{
    count_windows_kernel<FPM_K><<<grid, SK_THREADS, 0, st>>>(d_args, out);
}

}  // namespace fpm
