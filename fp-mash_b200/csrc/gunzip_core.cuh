// gunzip_core.cuh -- the sequential half of the GPU inflate (gunzip.cu): bit reader, canonical Huffman codes, the gzip
// member header, DEFLATE block headers (RFC 1951 3.2.3-3.2.7), token decode, CRC-32 arithmetic.  In the kernel one
// lane per warp runs this code; it is plain C++ on purpose (host + device), so that tests/test_host_cpu.py can run the very
// same functions on the CPU against zlib's output (tests/gunzip_sim.cu) before a GPU sees them.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define GZ_HD __host__ __device__
#else
#define GZ_HD
#define __forceinline__ inline
#endif
#ifdef __CUDA_ARCH__
#define GZ_TABLE __constant__
#else
#define GZ_TABLE static const
#endif

namespace fpm {

GZ_HD inline uint32_t gz_brev(uint32_t v)
{
#ifdef __CUDA_ARCH__
    return __brev(v);
#else
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
    v = ((v >> 8) & 0x00ff00ffu) | ((v & 0x00ff00ffu) << 8);
    return (v >> 16) | (v << 16);
#endif
}

constexpr int GZ_RING_WORDS = 512;            // 2 KB of compressed bytes staged per warp
constexpr int GZ_LIT_BITS = 10, GZ_DIST_BITS = 8;
constexpr uint64_t GZ_MAX_FILE = (1ull << 29) - 65536;      // compressed bytes per file the 32-bit bit position covers (larger files: host reader)
constexpr uint32_t GZ_OK = 0, GZ_RESIZE = 1, GZ_BAD = 2, GZ_HAS_NUL = 3;

GZ_TABLE uint16_t c_lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
GZ_TABLE uint8_t c_lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
GZ_TABLE uint16_t c_dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
GZ_TABLE uint8_t c_dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
GZ_TABLE uint8_t c_clorder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// One canonical Huffman code: a first-level table indexed by the next tab_bits bits of the stream, and the (count per
// length, symbols in canonical order) form for the codes that do not fit it.  Table entries carry everything the
// decoder needs for the symbol, so that a token costs two shared-memory lookups and no constant-table reads:
//   bits 0-3 code length (0: not a short code -> canonical decode), 4-7 number of extra bits, 8-9 kind, 10-14 code length +
//   extra bits (what the symbol advances the stream by), 15 "not for the fast loop" (no short code here, end of block,
//   invalid symbol), 16-31 value (the literal byte; the base of a length or distance symbol).
constexpr uint32_t GZ_K_LIT = 0u << 8, GZ_K_LEN = 1u << 8, GZ_K_EOB = 2u << 8, GZ_K_BAD = 3u << 8, GZ_K_MASK = 3u << 8;
constexpr uint32_t GZ_RARE = 1u << 15;
struct GzCode {
    uint16_t* count;           // [16]
    uint16_t* syms;
    uint32_t* tab;
    int tab_bits;
    int kind;                  // 0 literal/length code, 1 distance code, 2 code length code (plain symbols)
};

struct GzShared {
    uint32_t ring[GZ_RING_WORDS];
    uint32_t lit_tab[1 << GZ_LIT_BITS];
    uint32_t dist_tab[1 << GZ_DIST_BITS];
    uint16_t lit_syms[288], dist_syms[32], cl_syms[19];
    uint16_t lit_count[16], dist_count[16], cl_count[16];
    uint8_t lens[320];
    uint32_t tok[32];
    uint32_t crc_part[32];
    uint32_t crc_tab[256];
};

// what the decoder needs to know about symbol `sym` of a code of the given kind, coded with `len` bits
GZ_HD inline uint32_t gz_symbol_entry(int kind, int sym, int len)
{
    uint32_t e, x = 0;
    if (kind == 0) {
        if (sym < 256) e = GZ_K_LIT | ((uint32_t)sym << 16);
        else if (sym == 256) e = GZ_K_EOB | GZ_RARE;
        else if (sym >= 286) e = GZ_K_BAD | GZ_RARE;
        else { x = c_lext[sym - 257]; e = GZ_K_LEN | ((uint32_t)c_lbase[sym - 257] << 16); }
    } else if (kind == 1) {
        if (sym >= 30) e = GZ_K_BAD | GZ_RARE;
        else { x = c_dext[sym]; e = (uint32_t)c_dbase[sym] << 16; }
    } else e = (uint32_t)sym << 16;
    return e | (uint32_t)len | (x << 4) | (((uint32_t)len + x) << 10);
}

// ---- lane 0's bit reader: a bit position over a ring of 32-bit words ---------------------------------------------------
struct GzBits {
    const uint8_t* base;       // 16-byte aligned address at or below the file's first compressed byte
    uint32_t* ring;
    uint32_t bp;               // next bit, counted from base (32 bits: a file of up to GZ_MAX_FILE compressed bytes)
    uint32_t whi;              // words [.., whi) (index from base; a multiple of 4) are in the ring

    GZ_HD __forceinline__ uint32_t wnext() const { return bp >> 5; }
    GZ_HD __forceinline__ void self_fill()                     // the whole-warp top-up fell behind (long header fields): 16 bytes by lane 0 alone
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(base + 4ull * whi);
        uint32_t* r = ring + (whi & (GZ_RING_WORDS - 1));
        r[0] = src[0]; r[1] = src[1]; r[2] = src[2]; r[3] = src[3];
        whi += 4;
    }
    // the next 32 bits of the stream (bit 0 = next bit); the two words must be in the ring
    GZ_HD __forceinline__ uint32_t window() const
    {
        const uint32_t w = bp >> 5;
        const uint32_t w0 = ring[w & (GZ_RING_WORDS - 1)], w1 = ring[(w + 1) & (GZ_RING_WORDS - 1)];
#ifdef __CUDA_ARCH__
        return __funnelshift_r(w0, w1, bp & 31u);
#else
        return (uint32_t)((((uint64_t)w1 << 32) | w0) >> (bp & 31));
#endif
    }
    GZ_HD __forceinline__ uint32_t window_checked() { while (wnext() + 2 > whi) self_fill(); return window(); }
    GZ_HD __forceinline__ uint32_t get(int n) { const uint32_t v = window_checked() & ((1u << n) - 1u); bp += n; return v; }     // n <= 16
    GZ_HD __forceinline__ void align_byte() { bp = (bp + 7) & ~7u; }
};

// canonical decode over the window `cur`, one bit at a time (codes longer than the first-level table, and the 19-symbol
// code length code): the entry of the decoded symbol with its code length, or 0 when no code matches
GZ_HD inline uint32_t gz_decode_slow(uint32_t cur, const GzCode& h)
{
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; len++) {
        code |= (int)((cur >> (len - 1)) & 1u);
        const int c = h.count[len];
        if (code - c < first) return gz_symbol_entry(h.kind, h.syms[index + (code - first)], len);
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return 0;
}

// lengths -> code.  Returns 0 for a complete code, > 0 for an incomplete one, < 0 for an over-subscribed one
GZ_HD inline int gz_build(GzCode& h, const uint8_t* lens, int n)
{
    for (int l = 0; l <= 15; l++) h.count[l] = 0;
    for (int s = 0; s < n; s++) h.count[lens[s]]++;
    const int tab_n = h.tab ? 1 << h.tab_bits : 0;
    for (int i = 0; i < tab_n; i++) h.tab[i] = GZ_RARE;               // no short code ends here
    if (h.count[0] == n) return 0;                         // no codes: complete, every use fails
    int left = 1;
    for (int l = 1; l <= 15; l++) {
        left <<= 1;
        left -= h.count[l];
        if (left < 0) return left;
    }
    uint16_t offs[16], next[16];
    offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = offs[l] + h.count[l];
    int code = 0;
    for (int l = 1; l <= 15; l++) { code = (code + h.count[l - 1] * (l > 1)) << 1; next[l] = (uint16_t)code; }
    for (int s = 0; s < n; s++) {
        const int l = lens[s];
        if (!l) continue;
        h.syms[offs[l]++] = (uint16_t)s;
        const uint32_t c = next[l]++;
        if (h.tab && l <= h.tab_bits) {
            const uint32_t r = gz_brev(c) >> (32 - l);          // codes enter the stream most significant bit first
            const uint32_t e = gz_symbol_entry(h.kind, s, l);
            for (uint32_t i = r; i < (uint32_t)tab_n; i += 1u << l) h.tab[i] = e;
        }
    }
    return left;
}

// ---- CRC-32 (the gzip polynomial, reflected) --------------------------------------------------------------------
constexpr uint32_t GZ_POLY = 0xedb88320u;
GZ_HD inline uint32_t gz_multmodp(uint32_t a, uint32_t b)             // a(x) b(x) mod P, bit 31 = x^0
{
    uint32_t m = 1u << 31, p = 0;
    if (a == 0) return 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ GZ_POLY : b >> 1;
    }
    return p;
}
GZ_HD inline uint32_t gz_x8n_modp(uint64_t n)                          // x^(8 n) mod P
{
    uint32_t sq = 1u << 23, p = 1u << 31;                            // x^8, x^0
    while (n) {
        if (n & 1) p = gz_multmodp(sq, p);
        sq = gz_multmodp(sq, sq);
        n >>= 1;
    }
    return p;
}
GZ_HD inline uint32_t gz_crc_table_entry(uint32_t i)
{
    for (int k = 0; k < 8; k++) i = (i & 1u) ? (i >> 1) ^ GZ_POLY : i >> 1;
    return i;
}
GZ_HD __forceinline__ uint32_t gz_crc_byte(const uint32_t* tab, uint32_t c, uint32_t byte) { return tab[(c ^ byte) & 0xffu] ^ (c >> 8); }

// ---- the states of one file's decode ---------------------------------------------------------------------------
enum GzState { GZ_S_HEADER, GZ_S_BLOCK, GZ_S_CODES, GZ_S_STORED, GZ_S_TRAILER, GZ_S_DONE, GZ_S_BAD };

struct GzStream {
    GzBits b;
    GzCode lit, dist, cl;
    uint64_t base_off;         // offset of b.base in the compressed buffer
    uint64_t in_end;           // offset of the first byte behind the file
    uint32_t stored_len;
    bool last_block;

    GZ_HD void init(GzShared& sh, const uint8_t* in, uint64_t in_begin, uint64_t in_end_)
    {
        base_off = in_begin & ~15ull;
        in_end = in_end_;
        b.base = in + base_off;
        b.ring = sh.ring;
        b.bp = 8 * (uint32_t)(in_begin - base_off);
        b.whi = 0;
        lit.count = sh.lit_count; lit.syms = sh.lit_syms; lit.tab = sh.lit_tab; lit.tab_bits = GZ_LIT_BITS; lit.kind = 0;
        dist.count = sh.dist_count; dist.syms = sh.dist_syms; dist.tab = sh.dist_tab; dist.tab_bits = GZ_DIST_BITS; dist.kind = 1;
        cl.count = sh.cl_count; cl.syms = sh.cl_syms; cl.tab = nullptr; cl.tab_bits = 0; cl.kind = 2;
        stored_len = 0;
        last_block = false;
    }
    GZ_HD uint64_t in_pos() const { return base_off + (uint64_t)(b.bp >> 3); }       // offset of the byte holding the next bit
    // the reader restarts at compressed offset `at` (after a stored block): the caller refills the ring
    GZ_HD void seek(uint64_t at)
    {
        const uint64_t rel = at - base_off;
        b.whi = (uint32_t)((rel & ~15ull) >> 2);
        b.bp = 8 * (uint32_t)rel;
    }
};

// RFC 1952 member header: ID1 ID2 CM FLG MTIME(4) XFL OS [XLEN + extra] [name 0] [comment 0] [CRC16]
GZ_HD inline int gz_read_header(GzStream& s)
{
    GzBits& b = s.b;
    if (s.in_pos() + 18 > s.in_end) return GZ_S_BAD;
    const uint32_t id1 = b.get(8), id2 = b.get(8), cm = b.get(8), flg = b.get(8);
    if (id1 != 0x1f || id2 != 0x8b || cm != 8 || (flg & 0xe0)) return GZ_S_BAD;
    for (int i = 0; i < 6; i++) b.get(8);
    if (flg & 4) {
        uint32_t xlen = b.get(8);
        xlen |= b.get(8) << 8;
        for (uint32_t i = 0; i < xlen && s.in_pos() < s.in_end; i++) b.get(8);
    }
    if (flg & 8) while (s.in_pos() < s.in_end && b.get(8) != 0) {}
    if (flg & 16) while (s.in_pos() < s.in_end && b.get(8) != 0) {}
    if (flg & 2) { b.get(8); b.get(8); }
    return s.in_pos() >= s.in_end ? GZ_S_BAD : GZ_S_BLOCK;
}

// block header: BFINAL, BTYPE; stored -> LEN / NLEN (the raw bytes follow at s.in_pos()); fixed / dynamic -> the two codes
GZ_HD inline int gz_read_block(GzStream& s, GzShared& sh)
{
    GzBits& b = s.b;
    s.last_block = b.get(1) != 0;
    const uint32_t type = b.get(2);
    if (type == 0) {
        b.align_byte();
        const uint32_t len = b.get(16), nlen = b.get(16);
        if ((len ^ 0xffffu) != nlen) return GZ_S_BAD;
        s.stored_len = len;
        return GZ_S_STORED;
    }
    if (type == 1) {
        for (int i = 0; i < 144; i++) sh.lens[i] = 8;
        for (int i = 144; i < 256; i++) sh.lens[i] = 9;
        for (int i = 256; i < 280; i++) sh.lens[i] = 7;
        for (int i = 280; i < 288; i++) sh.lens[i] = 8;
        gz_build(s.lit, sh.lens, 288);
        for (int i = 0; i < 32; i++) sh.lens[i] = 5;
        gz_build(s.dist, sh.lens, 32);
        return GZ_S_CODES;
    }
    if (type != 2) return GZ_S_BAD;
    const int nlen = (int)b.get(5) + 257, ndist = (int)b.get(5) + 1, ncode = (int)b.get(4) + 4;
    if (nlen > 286 || ndist > 30) return GZ_S_BAD;
    uint8_t cll[19];
    for (int i = 0; i < 19; i++) cll[i] = 0;
    for (int i = 0; i < ncode; i++) cll[c_clorder[i]] = (uint8_t)b.get(3);
    if (gz_build(s.cl, cll, 19) != 0) return GZ_S_BAD;                      // the code length code must be complete
    int idx = 0;
    while (idx < nlen + ndist) {
        const uint32_t e = gz_decode_slow(b.window_checked(), s.cl);
        if (!e) return GZ_S_BAD;
        b.bp += e & 15u;
        const int sym = (int)(e >> 16);
        if (sym < 16) { sh.lens[idx++] = (uint8_t)sym; continue; }
        int prev = 0, rep;
        if (sym == 16) {
            if (idx == 0) return GZ_S_BAD;
            prev = sh.lens[idx - 1];
            rep = 3 + (int)b.get(2);
        } else if (sym == 17) rep = 3 + (int)b.get(3);
        else rep = 11 + (int)b.get(7);
        if (idx + rep > nlen + ndist) return GZ_S_BAD;
        while (rep--) sh.lens[idx++] = (uint8_t)prev;
    }
    if (sh.lens[256] == 0) return GZ_S_BAD;                                 // no end-of-block code
    // an incomplete code is accepted only when it is a single code of length one (zlib's inflate_table)
    int err = gz_build(s.lit, sh.lens, nlen);
    if (err < 0 || (err > 0 && !(nlen - s.lit.count[0] == 1 && s.lit.count[1] == 1))) return GZ_S_BAD;
    uint8_t dl[30];
    for (int i = 0; i < ndist; i++) dl[i] = sh.lens[nlen + i];
    err = gz_build(s.dist, dl, ndist);
    if (err < 0 || (err > 0 && !(ndist - s.dist.count[0] == 1 && s.dist.count[1] == 1))) return GZ_S_BAD;
    return GZ_S_CODES;
}

// ---- the token loop ---------------------------------------------------------------------------------------------
// Shared-memory words by byte offset.  On the device these are ld.shared / st.shared on 32-bit shared-space addresses
// computed once per batch: through the generic pointers of GzShared the compiler re-derived the shared window base
// (S2R SR_CgaCtaId ...) at every access of the loop.
#ifdef __CUDA_ARCH__
typedef uint32_t gz_smem_t;
__device__ __forceinline__ gz_smem_t gz_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t gz_ld(gz_smem_t base, uint32_t byte_off)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + byte_off));
    return v;
}
__device__ __forceinline__ void gz_st(gz_smem_t base, uint32_t byte_off, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + byte_off), "r"(v)); }
#else
typedef uint8_t* gz_smem_t;
inline gz_smem_t gz_smem(const void* p) { return (uint8_t*)p; }
inline uint32_t gz_ld(gz_smem_t base, uint32_t byte_off) { return *reinterpret_cast<const uint32_t*>(base + byte_off); }
inline void gz_st(gz_smem_t base, uint32_t byte_off, uint32_t v) { *reinterpret_cast<uint32_t*>(base + byte_off) = v; }
#endif

// One token the careful way (codes longer than the first-level tables, end of block, invalid symbols): the reader is at the
// token's first bit.  Returns the next state; a token was written to sh.tok[*n] (and *n advanced) unless the block ended.
GZ_HD inline int gz_decode_one(GzStream& s, GzShared& sh, uint32_t* n)
{
    GzBits& b = s.b;
    uint32_t cur = b.window();
    uint32_t e = s.lit.tab[cur & ((1u << GZ_LIT_BITS) - 1u)];
    if (!(e & 15u)) e = gz_decode_slow(cur, s.lit);
    const uint32_t kind = e & GZ_K_MASK, cl = e & 15u;
    if (!cl || kind == GZ_K_BAD) return GZ_S_BAD;
    if (kind == GZ_K_LIT) { b.bp += cl; sh.tok[(*n)++] = e >> 16; return GZ_S_CODES; }
    if (kind == GZ_K_EOB) { b.bp += cl; return s.last_block ? GZ_S_TRAILER : GZ_S_BLOCK; }
    const uint32_t xl = (e >> 4) & 15u;
    const uint32_t len = (e >> 16) + ((cur >> cl) & ((1u << xl) - 1u));
    b.bp += cl + xl;
    cur = b.window();
    uint32_t d = s.dist.tab[cur & ((1u << GZ_DIST_BITS) - 1u)];
    if (!(d & 15u)) d = gz_decode_slow(cur, s.dist);
    const uint32_t dl = d & 15u, xd = (d >> 4) & 15u;
    if (!dl || (d & GZ_K_MASK) == GZ_K_BAD) return GZ_S_BAD;
    const uint32_t dd = (d >> 16) + ((cur >> dl) & ((1u << xd) - 1u));
    b.bp += dl + xd;
    sh.tok[(*n)++] = 0x80000000u | (len << 16) | (dd - 1);
    return GZ_S_CODES;
}

// The token that starts at bit `pos` of the stream, decoded through the first-level tables only (the serial loop below has
// checked that they suffice): literal = its byte; match = 1 << 31 | length << 16 | (distance - 1).  Any lane can do this for
// any token of the round once the token's first bit is known -- which is the only thing that is sequential about DEFLATE.
GZ_HD __forceinline__ uint32_t gz_token_at(const GzShared& sh, uint32_t pos)
{
    uint32_t w = pos >> 5;
    uint32_t w0 = sh.ring[w & (GZ_RING_WORDS - 1)], w1 = sh.ring[(w + 1) & (GZ_RING_WORDS - 1)];
#ifdef __CUDA_ARCH__
    uint32_t cur = __funnelshift_r(w0, w1, pos & 31u);
#else
    uint32_t cur = (uint32_t)((((uint64_t)w1 << 32) | w0) >> (pos & 31));
#endif
    const uint32_t e = sh.lit_tab[cur & ((1u << GZ_LIT_BITS) - 1u)];
    const uint32_t cl = e & 15u, xl = (e >> 4) & 15u;
    const uint32_t val = (e >> 16) + ((cur >> cl) & ((1u << xl) - 1u));
    if ((e & GZ_K_MASK) != GZ_K_LEN) return val;
    pos += cl + xl;
    w = pos >> 5;
    w0 = sh.ring[w & (GZ_RING_WORDS - 1)]; w1 = sh.ring[(w + 1) & (GZ_RING_WORDS - 1)];
#ifdef __CUDA_ARCH__
    cur = __funnelshift_r(w0, w1, pos & 31u);
#else
    cur = (uint32_t)((((uint64_t)w1 << 32) | w0) >> (pos & 31));
#endif
    const uint32_t d = sh.dist_tab[cur & ((1u << GZ_DIST_BITS) - 1u)];
    const uint32_t dl = d & 15u, xd = (d >> 4) & 15u;
    const uint32_t dd = (d >> 16) + ((cur >> dl) & ((1u << xd) - 1u));
    return 0x80000000u | (val << 16) | (dd - 1u);
}

// One round of a Huffman block: where up to 32 tokens START.  sh.tok[i] = bit position of token i, for the lanes to decode
// with gz_token_at -- except where bit i of *ready is set: that token needed the careful path and sh.tok[i] is the token.
// Returns the next state (GZ_S_CODES: block goes on); *ntok = entries written.  A token is at most 48 bits long and the
// caller's top-up left more than 1.5 KB in the ring: no ring refill inside.
// This loop bounds a stream's rate -- one thread, one dependent chain (profiles/r02_gunzip_kernel.txt: an instruction every
// 4 to 6 cycles, every branch on a fresh predicate as expensive as a shared-memory load).  So it does nothing but find the
// starts: per token two table lookups whose entries say by how many bits the symbol (code + extra bits) advances the
// stream, on a 64-bit register pair refilled by select from a ring word fetched one refill ahead; a literal runs through
// the distance half with a zero advance instead of branching around it; and everything rare -- a code longer than the
// first-level table, end of block, an invalid symbol: one flag bit in the entries -- leaves through ONE test per token,
// is handled by gz_decode_one from the token's first bit, and the loop is entered again.
GZ_HD inline int gz_decode_batch(GzStream& s, GzShared& sh, uint32_t* ntok, uint32_t* ready)
{
    const gz_smem_t ring = gz_smem(sh.ring), lit_tab = gz_smem(sh.lit_tab), dist_tab = gz_smem(sh.dist_tab), tok = gz_smem(sh.tok);
    constexpr uint32_t RING_MASK = 4u * (GZ_RING_WORDS - 1);
    int next = GZ_S_CODES;
    uint32_t n = 0, rdy = 0;
    while (n < 32 && next == GZ_S_CODES) {
        uint32_t w4 = (s.b.bp >> 5) * 4u;                          // byte offset (unwrapped) of the next ring word to load
        uint64_t bb = (((uint64_t)gz_ld(ring, (w4 + 4u) & RING_MASK) << 32) | gz_ld(ring, w4 & RING_MASK)) >> (s.b.bp & 31u);
        uint32_t nb = 64u - (s.b.bp & 31u);                        // valid bits in bb
        w4 += 8u;
        uint32_t nxt = gz_ld(ring, w4 & RING_MASK);
        // invariant at a token's first bit: nb >= 33.  Both lookups index with bits that are already there (10 of >= 33; 8 of
        // the >= 13 left behind a length symbol), so the two refills wait for nobody: they overlap the lookups' latency, and
        // the chain per token is lookup, shift, lookup, shift.
        {
            const uint32_t fill = nb <= 32u ? 1u : 0u;
            bb |= fill ? (uint64_t)nxt << nb : 0ull;
            nb += fill * 32u;
            w4 += fill * 4u;
            nxt = gz_ld(ring, w4 & RING_MASK);
        }
        for (;;) {
            const uint32_t pos = 8u * w4 - nb;                     // the token's first bit
            const uint32_t e = gz_ld(lit_tab, ((uint32_t)bb & ((1u << GZ_LIT_BITS) - 1u)) * 4u);
            const uint32_t a1 = (e >> 10) & 31u;                   // <= 20
            bb >>= a1;
            nb -= a1;                                              // >= 13
            const uint32_t d = gz_ld(dist_tab, ((uint32_t)bb & ((1u << GZ_DIST_BITS) - 1u)) * 4u);
            uint32_t fill = nb <= 32u ? 1u : 0u;
            bb |= fill ? (uint64_t)nxt << nb : 0ull;
            nb += fill * 32u;                                      // >= 45
            w4 += fill * 4u;
            nxt = gz_ld(ring, w4 & RING_MASK);
            const uint32_t is_len = 0u - ((e >> 8) & 1u);          // all ones for a length symbol (kind bit 0; EOB / BAD carry GZ_RARE)
            const uint32_t dm = d & is_len;
            if ((e | dm) & GZ_RARE) { s.b.bp = pos; break; }       // back to the token's first bit
            const uint32_t a2 = (dm >> 10) & 31u;                  // <= 28
            bb >>= a2;
            nb -= a2;                                              // >= 17
            fill = nb <= 32u ? 1u : 0u;
            bb |= fill ? (uint64_t)nxt << nb : 0ull;
            nb += fill * 32u;                                      // >= 33 again
            w4 += fill * 4u;
            nxt = gz_ld(ring, w4 & RING_MASK);
            gz_st(tok, n * 4u, pos);
            n++;
            if (n == 32) { s.b.bp = 8u * w4 - nb; break; }
        }
        if (n < 32) {
            const uint32_t before = n;
            next = gz_decode_one(s, sh, &n);
            if (n != before) rdy |= 1u << before;
        }
    }
    // reading beyond the file's end means a truncated stream (zeros or the next file follow there)
    if (s.in_pos() > s.in_end + 8) next = GZ_S_BAD;
    *ntok = n;
    *ready = rdy;
    return next;
}

// member trailer: CRC-32 and ISIZE (little-endian) at the next byte boundary
GZ_HD inline bool gz_read_trailer(GzStream& s, uint32_t* crc, uint32_t* isize)
{
    GzBits& b = s.b;
    b.align_byte();
    if (s.in_pos() + 8 > s.in_end) return false;
    *crc = b.get(16); *crc |= b.get(16) << 16;
    *isize = b.get(16); *isize |= b.get(16) << 16;
    return true;
}

// another member?  zlib's gzread goes on when the gzip magic follows and ignores anything else
GZ_HD inline bool gz_more_members(GzStream& s)
{
    if (s.in_pos() + 18 > s.in_end) return false;
    return (s.b.window_checked() & 0xffffu) == 0x8b1fu;
}

}  // namespace fpm
