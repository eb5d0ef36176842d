// murmur3.cuh -- MurmurHash3_x64_128 on device, h1 only.
//
// Replaces MurmurHash3_x64_128 (MurmurHash3.cpp:255-331) + getHash (hash.cpp:12-40).
// Only out[0] (h1) is ever consumed by the reference, so the final "h2 += h1" is dropped.
// All arithmetic is 64-bit integer (IMAD / SHF / LOP3 / IADD3 on sm_100a): the kernel
// built on this is integer-ALU bound, not HBM bound.
#pragma once
#include <stdint.h>

namespace fpm {

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r)
{
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    uint32_t nlo, nhi;
    if (r == 32) { nlo = hi; nhi = lo; }
    else if (r < 32) { nhi = __funnelshift_l(lo, hi, r); nlo = __funnelshift_l(hi, lo, r); }
    else { nhi = __funnelshift_l(hi, lo, r - 32); nlo = __funnelshift_l(lo, hi, r - 32); }
    return ((uint64_t)nhi << 32) | nlo;
}

// x * C mod 2^64 for a compile-time C, pinned as three FMA-pipe instructions (IMAD.WIDE for lo*lo, then two
// dependent IMADs folding hi*lo and lo*hi into the high word); left to itself nvcc emits four (two independent
// IMADs, the IMAD.WIDE and an add).  The longer dependency chain is hidden by the other warps.
template <uint64_t C>
__device__ __forceinline__ uint64_t mul64c(uint64_t x)
{
    const uint32_t xl = (uint32_t)x, xh = (uint32_t)(x >> 32);
    uint32_t lo, hi;
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %4;\n\tmov.b64 {%0, %1}, t;\n\tmad.lo.u32 %1, %3, %4, %1;\n\tmad.lo.u32 %1, %2, %5, %1;\n\t}"
        : "=r"(lo), "=&r"(hi)
        : "r"(xl), "r"(xh), "n"((uint32_t)C), "n"((uint32_t)(C >> 32)));
    return ((uint64_t)hi << 32) | lo;
}

// high word of x * C mod 2^64 only (mul.hi + two IMADs): the low word costs nothing extra with mul64c, but the
// 64-bit result of IMAD.WIDE occupies two destination registers and issues at half rate
template <uint64_t C>
__device__ __forceinline__ uint32_t mul64c_hi(uint64_t x)
{
    const uint32_t xl = (uint32_t)x, xh = (uint32_t)(x >> 32);
    uint32_t hi;
    asm("mul.hi.u32 %0, %1, %3;\n\tmad.lo.u32 %0, %2, %3, %0;\n\tmad.lo.u32 %0, %1, %4, %0;"
        : "=&r"(hi)
        : "r"(xl), "r"(xh), "n"((uint32_t)C), "n"((uint32_t)(C >> 32)));
    return hi;
}

__device__ __forceinline__ uint64_t fmix64(uint64_t k)
{
    k ^= k >> 33;
    k = mul64c<0xff51afd7ed558ccdULL>(k);
    k ^= k >> 33;
    k = mul64c<0xc4ceb9fe1a85ec53ULL>(k);
    k ^= k >> 33;
    return k;
}

#define FPM_MC1 0x87c37b91114253d5ULL
#define FPM_MC2 0x4cf5ad432745937fULL

__device__ __forceinline__ uint64_t mm_k1(uint64_t k) { k = mul64c<FPM_MC1>(k); k = rotl64(k, 31); k = mul64c<FPM_MC2>(k); return k; }
__device__ __forceinline__ uint64_t mm_k2(uint64_t k) { k = mul64c<FPM_MC2>(k); k = rotl64(k, 33); k = mul64c<FPM_MC1>(k); return k; }

// add1/add2 are the two additive constants of the block mix; callers in hot loops pass them in
// registers (an immediate would be re-materialised with two moves per use).
__device__ __forceinline__ void mm_block(uint64_t& h1, uint64_t& h2, uint64_t k1, uint64_t k2,
                                         uint64_t add1 = 0x52dce729ULL, uint64_t add2 = 0x38495ab5ULL)
{
    h1 ^= mm_k1(k1);
    h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + add1;
    h2 ^= mm_k2(k2);
    h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + add2;
}

__device__ __forceinline__ uint64_t mm_finish(uint64_t h1, uint64_t h2, uint64_t len)
{
    h1 ^= len; h2 ^= len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    return h1 + h2;
}

// Hash of a LEN-byte key given as zero-padded little-endian 64-bit words w[0..3]
// (bytes >= LEN are zero).  LEN is a compile-time constant in 1..32, so the block/tail
// structure of MurmurHash3.cpp:270-314 resolves statically.
template <int LEN>
__device__ __forceinline__ uint64_t murmur3_h1_fixed(const uint64_t (&w)[4], uint32_t seed,
                                                     uint64_t add1 = 0x52dce729ULL, uint64_t add2 = 0x38495ab5ULL)
{
    uint64_t h1 = seed, h2 = seed;
    constexpr int nblocks = LEN / 16;
    constexpr int rem = LEN & 15;
    if (nblocks >= 1) mm_block(h1, h2, w[0], w[1], add1, add2);
    if (nblocks >= 2) mm_block(h1, h2, w[2], w[3], add1, add2);
    if (rem > 8) h2 ^= mm_k2(w[2 * nblocks + 1]);
    if (rem > 0) h1 ^= mm_k1(w[2 * nblocks]);
    return mm_finish(h1, h2, (uint64_t)LEN);
}

// Generic length, key read as little-endian u64 words from memory (fingerprint token rows:
// the key IS an array of uint64, so every block is two tokens and the tail is 0 or 8 bytes).
__device__ __forceinline__ uint64_t murmur3_h1_tokens(const uint64_t* tok, uint64_t n_tok, uint32_t seed)
{
    uint64_t h1 = seed, h2 = seed;
    uint64_t i = 0;
    for (; i + 2 <= n_tok; i += 2) mm_block(h1, h2, tok[i], tok[i + 1]);
    if (i < n_tok) h1 ^= mm_k1(tok[i]);          // 8-byte tail -> lane 1 only (case 8)
    // len is an int in the reference; n_tok*8 stays far below 2^31 for any real line
    return mm_finish(h1, h2, (uint64_t)(int64_t)(int32_t)(uint32_t)(n_tok * 8));
}

}  // namespace fpm
