// sketch_tables.cuh -- MurmurHash3_x64_128 (h1) of a 2-bit packed k-mer through shared-memory tables.
//
// Replaces getHash(kmer, k) -> MurmurHash3_x64_128 (hash.cpp:12-40, MurmurHash3.cpp:255-331) for a key that exists
// only as 2-bit codes: the reference hashes the ASCII letters, so the first thing Murmur does with every 8-byte
// word of the key -- multiply it by c1 (k1 lanes) or c2 (k2 lanes) -- is a function of 8 bases = 16 bits.
// Round 1 expanded the codes back to ASCII (one LDS.32 per four bases) and multiplied (three IMADs per word, one
// of them the half-rate IMAD.WIDE); ncu showed the FMA-heavy pipe as the tightest bound (profiles/r01_sketch_hash_v4.txt,
// DESIGN.md 4.1).  Multiplication mod 2^64 is linear, so the product is assembled from table entries instead:
//   * T0[C][v]  = expand4(v) * C            (64-bit, v = one byte of packed codes = bases 0..3 of a word)
//     word * C  = T0[C][lo group] + ((expand4(hi group) * low32(C)) << 32)        -> one LDS.64, one LDS.32, one IMAD
//   * a tail word of at most five bases (k = 21: bases 16..20) indexes a table of its finished lane value
//     rotl(word * c1, 31) * c2 (or the k2 form): 4^5 entries, no arithmetic at all.
// FPM_LAZYFIN: the last xor-shift of fmix64 only touches the low word, so the filter compares hi(a1) + hi(a2) + 1
// (one IADD3) against the bound and the exact 64-bit hash is finished only for the ~1e-4 of windows that may pass.
#pragma once
#include "sketch_kernels.cuh"

#ifndef FPM_TAILTAB
#define FPM_TAILTAB 1
#endif
#ifndef FPM_T0TAB
#define FPM_T0TAB 1
#endif
#ifndef FPM_T0MASK
#define FPM_T0MASK 0xf   // which key words (bit c = word c) take the T0 table; the others expand + multiply
#endif
#ifndef FPM_LAZYFIN
#define FPM_LAZYFIN 1
#endif
#ifndef FPM_HIONLY
#define FPM_HIONLY 0     // lazy finish: the last fmix multiply produces only its high word (mul.hi instead of mul.wide)
#endif

namespace fpm {

template <int K>
struct KmerPlan {
    static constexpr int nblocks = K / 16, rem = K & 15;
    static constexpr int n_words = (K + 7) / 8;
    static constexpr __host__ __device__ int nbw(int c) { return K - 8 * c >= 8 ? 8 : (K - 8 * c > 0 ? K - 8 * c : 0); }
    static constexpr int tail_word = n_words - 1;                       // only meaningful when rem != 0
    static constexpr int tail_nb = nbw(n_words - 1);
    static constexpr bool tail_tab = FPM_TAILTAB && rem != 0 && tail_nb <= 5;
    static constexpr int tail_entries = tail_tab ? (1 << (2 * tail_nb)) : 1;
    static constexpr bool tail_is_k2 = ((n_words - 1) & 1) != 0;        // odd words feed the k2 lane
};

template <int K>
struct __align__(16) KmerTables {
    uint64_t tail[KmerPlan<K>::tail_entries];
    uint64_t t0c1[FPM_T0TAB ? 256 : 1];
    uint64_t t0c2[FPM_T0TAB ? 256 : 1];
    uint32_t lut[512];                  // expand4: [v] four letters, [256 + v] the first K % 4 letters (zero padded)
};

__device__ __forceinline__ uint64_t ascii_of_codes(uint32_t codes, int nb)   // codes: nb bases, first base most significant
{
    uint64_t w = 0;
    for (int i = 0; i < nb; i++) w |= (uint64_t)(uint8_t)"ACGT"[(codes >> (2 * (nb - 1 - i))) & 3] << (8 * i);
    return w;
}

template <int K>
__device__ __forceinline__ void build_kmer_tables(KmerTables<K>* t, int tid, int nthreads)
{
    using P = KmerPlan<K>;
    build_expand_lut(t->lut, K, tid, nthreads);
    if (FPM_T0TAB)
        for (int v = tid; v < 256; v += nthreads) {
            const uint64_t e = ascii_of_codes((uint32_t)v, 4);
            t->t0c1[v] = e * FPM_MC1;
            t->t0c2[v] = e * FPM_MC2;
        }
    if (P::tail_tab)
        for (int v = tid; v < P::tail_entries; v += nthreads) {
            const uint64_t e = ascii_of_codes((uint32_t)v, P::tail_nb);
            t->tail[v] = P::tail_is_k2 ? mm_k2(e) : mm_k1(e);
        }
}

// word C of the key (bases 8C .. 8C+7; left-aligned codes in chi:clo) times MUL
template <int K, int C, uint64_t MUL>
__device__ __forceinline__ uint64_t word_times(uint32_t chi, uint32_t clo, uint32_t tb)
{
    using P = KmerPlan<K>;
    using T = KmerTables<K>;
    constexpr int nb = P::nbw(C);
    static_assert(nb >= 1, "word holds no base");
    const uint32_t src = C < 2 ? chi : clo;
    constexpr int g0 = 2 * C, g1 = 2 * C + 1;                            // byte 3 of a register holds its first four bases
    constexpr uint32_t off_lut = (uint32_t)offsetof(T, lut);
    const uint32_t i0 = prmt(src, 0u, 0x4440u | (3 - (g0 & 3)));
    uint32_t xh = 0;
    if (nb > 4) {
        const uint32_t i1 = prmt(src, 0u, 0x4440u | (3 - (g1 & 3)));
        uint32_t a1;
        asm("mad.lo.u32 %0, %1, 4, %2;" : "=r"(a1) : "r"(i1), "r"(tb));
        asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(xh) : "r"(a1), "n"(off_lut + (nb < 8 ? 1024u : 0u)));
    }
    if (FPM_T0TAB && ((FPM_T0MASK >> C) & 1) && nb >= 4) {
        uint32_t a0, lo, hi;
        asm("mad.lo.u32 %0, %1, 8, %2;" : "=r"(a0) : "r"(i0), "r"(tb));
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(lo), "=r"(hi) : "r"(a0), "n"((uint32_t)(MUL == FPM_MC1 ? offsetof(T, t0c1) : offsetof(T, t0c2))));
        if (nb > 4) asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(xh), "n"((uint32_t)MUL));
        return ((uint64_t)hi << 32) | lo;
    }
    uint32_t a0, xl;
    asm("mad.lo.u32 %0, %1, 4, %2;" : "=r"(a0) : "r"(i0), "r"(tb));
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(xl) : "r"(a0), "n"(off_lut + (nb < 4 ? 1024u : 0u)));
    return mul64c<MUL>(((uint64_t)xh << 32) | xl);
}

template <int K, int C>
__device__ __forceinline__ uint64_t lane_k1(uint32_t chi, uint32_t clo, uint32_t tb)
{
    return mul64c<FPM_MC2>(rotl64(word_times<K, C, FPM_MC1>(chi, clo, tb), 31));
}

template <int K, int C>
__device__ __forceinline__ uint64_t lane_k2(uint32_t chi, uint32_t clo, uint32_t tb)
{
    return mul64c<FPM_MC1>(rotl64(word_times<K, C, FPM_MC2>(chi, clo, tb), 33));
}

// the tail word's finished lane value (mm_k1 / mm_k2 of its letters)
template <int K>
__device__ __forceinline__ uint64_t lane_tail(uint32_t chi, uint32_t clo, uint32_t tb)
{
    using P = KmerPlan<K>;
    using T = KmerTables<K>;
    constexpr int C = P::tail_word, nb = P::tail_nb;
    if constexpr (!P::tail_tab) {
        if constexpr (P::tail_is_k2) return lane_k2<K, C>(chi, clo, tb);
        else return lane_k1<K, C>(chi, clo, tb);
    }
    const uint32_t src = C < 2 ? chi : clo;
    uint32_t idx;
    if ((C & 1) == 0) idx = src >> (32 - 2 * nb);                        // the word is the register's upper half
    else idx = (src >> (16 - 2 * nb)) & ((1u << (2 * nb)) - 1u);
    uint32_t a, lo, hi;
    asm("mad.lo.u32 %0, %1, 8, %2;" : "=r"(a) : "r"(idx), "r"(tb));
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(lo), "=r"(hi) : "r"(a), "n"((uint32_t)offsetof(T, tail)));
    return ((uint64_t)hi << 32) | lo;
}

// MurmurHash3_x64_128 of the K letters up to the last multiply of both fmix64 calls:
// h1 (the value getHash returns for 64-bit hashes) = (a1 ^ a1 >> 33) + (a2 ^ a2 >> 33)
// add1s = add1 + 5 * seed: in the first block h2 is still the seed, so "h1 += h2; h1 = h1 * 5 + add1" is one multiply-add
template <int K>
__device__ __forceinline__ void kmer_hash_parts(uint32_t chi, uint32_t clo, uint32_t tb, uint32_t seed, uint64_t add1s, uint64_t add1, uint64_t add2,
                                                uint64_t& a1, uint64_t& a2)
{
    using P = KmerPlan<K>;
    uint64_t h1 = seed, h2 = seed;
    if constexpr (P::nblocks >= 1) {
        h1 ^= lane_k1<K, 0>(chi, clo, tb);
        h1 = rotl64(h1, 27); h1 = h1 * 5 + add1s;
        h2 ^= lane_k2<K, 1>(chi, clo, tb);
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + add2;
    }
    if constexpr (P::nblocks >= 2) {
        h1 ^= lane_k1<K, 2>(chi, clo, tb);
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + add1;
        h2 ^= lane_k2<K, 3>(chi, clo, tb);
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + add2;
    }
    if constexpr (P::rem > 8) {                                          // MurmurHash3.cpp:296-314
        h2 ^= lane_tail<K>(chi, clo, tb);
        h1 ^= lane_k1<K, 2 * P::nblocks>(chi, clo, tb);
    } else if constexpr (P::rem > 0) {
        h1 ^= lane_tail<K>(chi, clo, tb);
    }
    h1 ^= (uint64_t)K; h2 ^= (uint64_t)K;
    h1 += h2; h2 += h1;
    h1 ^= h1 >> 33; h1 = mul64c<0xff51afd7ed558ccdULL>(h1); h1 ^= h1 >> 33;
    h2 ^= h2 >> 33; h2 = mul64c<0xff51afd7ed558ccdULL>(h2); h2 ^= h2 >> 33;
#if FPM_HIONLY
    a1 = ((uint64_t)mul64c_hi<0xc4ceb9fe1a85ec53ULL>(h1) << 32) | (uint32_t)h1;     // low word: the multiplicand's, finished on demand
    a2 = ((uint64_t)mul64c_hi<0xc4ceb9fe1a85ec53ULL>(h2) << 32) | (uint32_t)h2;
#else
    a1 = mul64c<0xc4ceb9fe1a85ec53ULL>(h1);
    a2 = mul64c<0xc4ceb9fe1a85ec53ULL>(h2);
#endif
}

__device__ __forceinline__ uint64_t kmer_hash_finish(uint64_t a1, uint64_t a2)
{
#if FPM_HIONLY
    a1 = (a1 & 0xffffffff00000000ULL) | (uint32_t)((uint32_t)a1 * 0x1a85ec53u);
    a2 = (a2 & 0xffffffff00000000ULL) | (uint32_t)((uint32_t)a2 * 0x1a85ec53u);
#endif
    return (a1 ^ (a1 >> 33)) + (a2 ^ (a2 >> 33));
}

template <int K>
__device__ __forceinline__ uint64_t kmer_hash(uint32_t chi, uint32_t clo, uint32_t tb, uint32_t seed, uint64_t add1 = 0x52dce729ULL, uint64_t add2 = 0x38495ab5ULL)
{
    uint64_t a1, a2;
    kmer_hash_parts<K>(chi, clo, tb, seed, add1 + 5ull * seed, add1, add2, a1, a2);
    return kmer_hash_finish(a1, a2);
}

// The 16 windows of one block.  fw0..2: 48 forward bases; window i starts at base i.
// CANON: pick min(forward, reverse complement) (ties are palindromes: identical bytes).
// F is called as F(i, hash) for every window, valid or not (validity is checked only for the
// rare windows that pass the threshold).
template <int K, bool CANON, typename Sink>
__device__ __forceinline__ void hash_block16(uint32_t fw0, uint32_t fw1, uint32_t fw2, uint32_t seed, int hash32, uint32_t tb, Sink&& sink)
{
    uint32_t rc[5];
    if (CANON) {
        rc[0] = revcomp16(fw2); rc[1] = revcomp16(fw1); rc[2] = revcomp16(fw0); rc[3] = 0; rc[4] = 0;
    }
#pragma unroll
    for (int i = 0; i < 16; i++) {
        uint32_t fhi = i ? __funnelshift_l(fw1, fw0, 2 * i) : fw0;
        uint32_t flo = i ? __funnelshift_l(fw2, fw1, 2 * i) : fw1;
        uint32_t chi = fhi, clo = flo;
        if (CANON) {
            const int start = 48 - K - i;            // rc position of the window's first rc base
            const int a = start >> 4, sh = 2 * (start & 15);
            uint32_t rhi = sh ? __funnelshift_l(rc[a + 1], rc[a], sh) : rc[a];
            uint32_t rlo = sh ? __funnelshift_l(rc[a + 2], rc[a + 1], sh) : rc[a + 1];
            uint64_t f64 = ((uint64_t)fhi << 32) | flo, r64 = ((uint64_t)rhi << 32) | rlo;
            // Bits below the k-mer (neighbouring bases) can only decide the comparison when the
            // k-mer equals its own reverse complement, where both choices give the same bytes.
            bool use_r = r64 < f64;
            chi = use_r ? rhi : fhi;
            clo = use_r ? rlo : flo;
        }
        uint64_t h = kmer_hash<K>(chi, clo, tb, seed);      // the production kernel's hash (tables in shared memory)
        if (hash32) h &= 0xffffffffULL;
        sink(i, h);
    }
}

// ---------------------------------------------------------------------------------------
// Every window's hash, in order (tests only; single record, no groups).
// out[pos] = hash or SK_EMPTY-marked invalid via the parallel flag array.
// ---------------------------------------------------------------------------------------
template <int K, bool CANON>
__global__ void __launch_bounds__(SK_THREADS) kmer_hash_stream_kernel(const uint8_t* seq, uint64_t n_bytes, uint32_t seed,
                                                                      int fold_case, int hash32, uint64_t* out, uint8_t* out_valid)
{
    __shared__ uint32_t s_code[2 * SK_TILE_CHUNKS + 2];
    __shared__ uint32_t s_valid[SK_TILE_CHUNKS + 1];
    __shared__ KmerTables<K> s_tab;
    build_kmer_tables<K>(&s_tab, threadIdx.x, SK_THREADS);
    const uint32_t tb = (uint32_t)__cvta_generic_to_shared(&s_tab);
    const uint64_t tile_base = (uint64_t)blockIdx.x * SK_TILE_WINDOWS;
    convert_tile(seq, n_bytes, tile_base, fold_case, s_code, s_valid);
    if (threadIdx.x == 0) { s_code[2 * SK_TILE_CHUNKS] = 0; s_code[2 * SK_TILE_CHUNKS + 1] = 0; s_valid[SK_TILE_CHUNKS] = 0; }
    __syncthreads();
    for (int it = 0; it < SK_BLOCKS_PER_THREAD; it++) {
        const int b = it * SK_THREADS + threadIdx.x;
        const uint64_t block_pos = tile_base + 16ull * b;
        if (block_pos >= n_bytes) break;
        hash_block16<K, CANON>(s_code[b], s_code[b + 1], s_code[b + 2], seed, hash32, tb, [&](int i, uint64_t h) {
            uint64_t pos = block_pos + i;
            if (pos < n_bytes) {
                uint64_t v = ((uint64_t)s_valid[(b >> 1) + 1] << 32) | s_valid[b >> 1];
                v >>= (16 * (b & 1) + i);
                constexpr uint64_t km = (1ULL << K) - 1;
                out[pos] = h;
                out_valid[pos] = ((v & km) == km) ? 1 : 0;
            }
        });
    }
}

}  // namespace fpm
