// sketch_inst.cu -- compiled once per k-mer size (-DFPM_K=1..32) so every k gets fully
// specialised code (constant shifts, static Murmur block/tail structure) and the 32 objects
// build in parallel.
#include "sketch_hash_v2.cuh"
#include "sketch_launch.h"

#ifndef FPM_K
#error "compile with -DFPM_K=<k>"
#endif

namespace fpm {

#define FPM_CAT2(a, b) a##b
#define FPM_CAT(a, b) FPM_CAT2(a, b)

// hashes every window whose start lies in [range_lo, range_hi); tiles are laid from range_lo rounded
// down to a 32-byte boundary so the 128-bit loads stay aligned
void FPM_CAT(launch_sketch_hash_k, FPM_K)(bool canon, cudaStream_t st, const SketchArgs* d_args, uint64_t range_lo, uint64_t range_hi, int trace)
{
    if (range_hi <= range_lo) return;
    const uint64_t base = range_lo & ~31ull;
    const uint64_t n_tiles = (range_hi - base + WT_WINDOWS - 1) / WT_WINDOWS;
    const uint64_t per_cta = (uint64_t)(SK_THREADS / 32) * WT_TILES_PER_WARP;
    const uint32_t grid = (uint32_t)((n_tiles + per_cta - 1) / per_cta);
    if (canon) sketch_hash_kernel_v2<FPM_K, true><<<grid, SK_THREADS, 0, st>>>(d_args, range_lo, range_hi, base, trace);
    else sketch_hash_kernel_v2<FPM_K, false><<<grid, SK_THREADS, 0, st>>>(d_args, range_lo, range_hi, base, trace);
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
