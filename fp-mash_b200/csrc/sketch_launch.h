// sketch_launch.h -- per-k launcher table (definitions in sketch_inst.cu, one object per k).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fpm {

struct SketchArgs;

#define FPM_FOR_ALL_K(X) \
    X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) \
    X(17) X(18) X(19) X(20) X(21) X(22) X(23) X(24) X(25) X(26) X(27) X(28) X(29) X(30) X(31) X(32)

#define FPM_DECL(K)                                                                                              \
    void launch_sketch_hash_k##K(bool canon, cudaStream_t st, const SketchArgs* d_args, uint64_t range_lo, uint64_t range_hi, int trace); \
    void launch_hash_stream_k##K(bool canon, uint32_t grid, cudaStream_t st, const uint8_t* seq, uint64_t n,      \
                                 uint32_t seed, int fold, int hash32, uint64_t* out, uint8_t* valid);             \
    void launch_count_windows_k##K(uint32_t grid, cudaStream_t st, const SketchArgs* d_args, unsigned long long* out);
FPM_FOR_ALL_K(FPM_DECL)
#undef FPM_DECL

typedef void (*sketch_hash_launcher)(bool, cudaStream_t, const SketchArgs*, uint64_t, uint64_t, int);
typedef void (*hash_stream_launcher)(bool, uint32_t, cudaStream_t, const uint8_t*, uint64_t, uint32_t, int, int, uint64_t*, uint8_t*);
typedef void (*count_windows_launcher)(uint32_t, cudaStream_t, const SketchArgs*, unsigned long long*);

}  // namespace fpm
