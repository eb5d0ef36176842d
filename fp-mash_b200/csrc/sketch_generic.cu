// sketch_generic.cu -- sketching for alphabets other than ACGT (`-a` protein, `-z <alphabet>`).
//
// Same contract as sketch_hash_kernel_v2 (filter by threshold, insert into the per-sketch counting
// table) but the k-mer is hashed as the raw (case-folded) ASCII bytes of the forward strand: with a
// custom alphabet the reference always runs noncanonical (sketchParameterSetup.cpp:85-97), and a
// window is hashed iff all its k bytes are in `alphabet[]` (Sketch.cpp:694-719).  This path is the
// option-surface completion, not the headline: it is a straightforward shared-memory kernel.
#include "sketch_kernels.cuh"
#include "sketch_select.h"

namespace fpm {

constexpr int GEN_THREADS = 256;
constexpr int GEN_WPT = 8;                          // windows per thread
constexpr int GEN_TILE = GEN_THREADS * GEN_WPT;     // 2048 windows per CTA

__device__ __forceinline__ uint64_t murmur3_h1_runtime(const uint64_t (&w)[4], int len, uint32_t seed)
{
    uint64_t h1 = seed, h2 = seed;
    const int nblocks = len >> 4, rem = len & 15;
    if (nblocks >= 1) mm_block(h1, h2, w[0], w[1]);
    if (nblocks >= 2) mm_block(h1, h2, w[2], w[3]);
    const uint64_t t0 = nblocks == 0 ? w[0] : (nblocks == 1 ? w[2] : 0);
    const uint64_t t1 = nblocks == 0 ? w[1] : (nblocks == 1 ? w[3] : 0);
    if (rem > 8) h2 ^= mm_k2(t1);
    if (rem > 0) h1 ^= mm_k1(t0);
    return mm_finish(h1, h2, (uint64_t)len);
}

// mode 0: filter + table insert; mode 1: trace (order-dependent top count); mode 2: count valid windows
__global__ void __launch_bounds__(GEN_THREADS) sketch_generic_kernel(const SketchArgs* __restrict__ ga, const uint8_t* __restrict__ alphabet,
                                                                     int K, uint64_t range_lo, uint64_t range_hi, int mode,
                                                                     unsigned long long* out_kmers)
{
    const SketchArgs& a = *ga;
    __shared__ uint8_t s_byte[GEN_TILE + 32];       // case-folded bytes; 0 where not in the alphabet
    __shared__ uint8_t s_alpha[256];
    __shared__ uint32_t s_g[2];
    __shared__ unsigned long long s_tmax;
    const uint64_t tile_base = range_lo + (uint64_t)blockIdx.x * GEN_TILE;
    for (int i = threadIdx.x; i < 256; i += GEN_THREADS) s_alpha[i] = alphabet[i];
    __syncthreads();
    for (int i = threadIdx.x; i < GEN_TILE + 32; i += GEN_THREADS) {
        uint64_t p = tile_base + i;
        uint8_t c = p < a.n_bytes ? a.seq[p] : 0;
        if (a.fold_case && c > 96 && c < 123) c -= 32;           // Sketch.cpp:676-682
        s_byte[i] = s_alpha[c] ? c : 0;                           // 0x00 is never in an alphabet
    }
    if (threadIdx.x == 0) {
        uint64_t last = tile_base + GEN_TILE - 1;
        if (last >= range_hi) last = range_hi - 1;
        uint32_t g0 = find_group(a.group_off, 0, a.n_groups - 1, tile_base);
        uint32_t g1 = find_group(a.group_off, g0, a.n_groups - 1, last);
        unsigned long long tm = 0;
        if (mode != 2)
            for (uint32_t g = g0; g <= g1; g++)
                if (a.active[g] && a.thresh[g] > tm) tm = a.thresh[g];
        s_g[0] = g0; s_g[1] = g1; s_tmax = tm;
    }
    __syncthreads();
    const uint32_t g_lo = s_g[0], g_hi = s_g[1];
    const uint64_t tmax = s_tmax;
    const int first = threadIdx.x * GEN_WPT;
    // run = number of consecutive in-alphabet bytes ending just before the next byte to add
    int run = 0;
    for (int j = 0; j < K - 1; j++) run = s_byte[first + j] ? run + 1 : 0;
    unsigned long long n_valid = 0;
    for (int i = 0; i < GEN_WPT; i++) {
        const int wdw = first + i;
        run = s_byte[wdw + K - 1] ? run + 1 : 0;
        const uint64_t pos = tile_base + wdw;
        if (run < K || pos >= range_hi) continue;
        if (mode == 2) { n_valid++; if (g_lo != g_hi) atomicAdd(&out_kmers[find_group(a.group_off, g_lo, g_hi, pos)], 1ULL); continue; }
        uint64_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 32; j++)
            if (j < K) w[j >> 3] |= (uint64_t)s_byte[wdw + j] << (8 * (j & 7));
        uint64_t h = murmur3_h1_runtime(w, K, a.seed);
        if (a.hash32) h &= 0xffffffffULL;
        if (h <= tmax) sketch_emit(a, h, pos, g_lo, g_hi, mode == 1);
    }
    if (mode == 2 && g_lo == g_hi) {
        for (int o = 16; o; o >>= 1) n_valid += __shfl_down_sync(0xffffffffu, n_valid, o);
        if ((threadIdx.x & 31) == 0 && n_valid) atomicAdd(&out_kmers[g_lo], n_valid);
    }
}

void launch_sketch_generic(cudaStream_t st, const SketchArgs* d_args, const uint8_t* d_alphabet, int K, uint64_t range_lo,
                           uint64_t range_hi, int mode, unsigned long long* out_kmers)
{
    if (range_hi <= range_lo) return;
    const uint64_t tiles = (range_hi - range_lo + GEN_TILE - 1) / GEN_TILE;
    sketch_generic_kernel<<<(uint32_t)tiles, GEN_THREADS, 0, st>>>(d_args, d_alphabet, K, range_lo, range_hi, mode, out_kmers);
}

}  // namespace fpm
