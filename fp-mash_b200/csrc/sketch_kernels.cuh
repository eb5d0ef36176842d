// sketch_kernels.cuh -- the sm_100a sketching kernels.
//
//   sketch_hash_kernel_v2<K,CANON> (sketch_hash_v2.cuh; this header holds its building blocks)
//                                fused: ASCII -> 2-bit pack (+validity) -> canonical k-mer
//                                (reverse-complement min) -> 2-bit -> ASCII re-expansion
//                                (PRMT) -> MurmurHash3_x64_128 -> threshold filter ->
//                                per-sketch counting hash table (global atomics).
//                                Replaces addMinHashes + getHash + the gate of
//                                MinHashHeap::tryInsert (Sketch.cpp:664-735, hash.cpp:12-40,
//                                MinHashHeap.cpp:70-74).
//   sketch_select_kernel         per sketch: table -> entries with count >= min_cov ->
//                                shared-memory bitonic sort -> bottom-s hashes + counts.
//                                Replaces the rest of tryInsert + HashSet::toHashList
//                                (MinHashHeap.cpp:76-146, HashSet.cpp:78-118) through the
//                                closed form of SURVEY.md 8a/a4 (final set = s smallest
//                                hashes with >= min_cov occurrences, order independent).
//   sketch_topcount_kernel       the one order-DEPENDENT quantity: the multiplicity of the
//                                largest sketch element when the sketch is full (a4).
//   count_windows_kernel         number of valid k-mer windows per sketch (the bench unit).
//   kmer_hash_stream_kernel      every window's hash in order (parity tests vs getHash).
//
// Data layout in HBM
//   seq       : ASCII bytes, records back to back, each followed by 0x00 (not in any
//               alphabet), so a window that crosses a record boundary is invalid by itself
//               and the kernel needs no record table.  1 byte per base is read once
//               (plus a 32-base halo per 8192-window tile, served by L2).
//   tables    : per sketch an open-addressing table {key u64, count u32, firstpos u64},
//               capacity a power of two >= 4x the expected number of distinct survivors.
#pragma once
#include <stdint.h>
#include "murmur3.cuh"

namespace fpm {

constexpr int SK_THREADS = 256;
constexpr int SK_BLOCKS_PER_THREAD = 2;                                   // 16-window blocks per thread
constexpr int SK_TILE_BLOCKS = SK_THREADS * SK_BLOCKS_PER_THREAD;         // 512
constexpr int SK_TILE_WINDOWS = SK_TILE_BLOCKS * 16;                      // 8192
constexpr int SK_TILE_CHUNKS = SK_TILE_WINDOWS / 32 + 1;                  // 257 chunks of 32 bases (1 halo chunk)
constexpr uint64_t SK_EMPTY = ~0ULL;
constexpr uint32_t SK_SORT_CAP = 16384;                                   // u64 keys sortable in shared memory

struct SketchArgs {
    const uint8_t* seq;          // device, 16-byte aligned
    uint64_t n_bytes;
    const uint64_t* group_off;   // [n_groups+1] byte offsets
    uint32_t n_groups;
    // filter + tables
    const uint64_t* thresh;      // [n_groups] accept h <= thresh
    const uint8_t* active;       // [n_groups] group takes part in this pass
    uint64_t* tkeys;
    uint32_t* tcnt;
    uint64_t* tpos;
    const uint64_t* toff;        // [n_groups] first slot of the group's table
    const uint32_t* tmask;       // [n_groups] capacity-1
    uint32_t* maxkey_cnt;        // [n_groups] occurrences of the hash value ~0 (cannot be a table key)
    uint64_t* maxkey_pos;        // [n_groups]
    uint32_t* overflow;          // [n_groups]
    // trace mode (order-dependent top count)
    const uint64_t* fin_hashes;  // [n_groups][s]
    const uint32_t* fin_n;       // [n_groups]
    const uint64_t* tr_off;      // [n_groups*s] bucket start
    const uint32_t* tr_cap;      // [n_groups*s] bucket capacity (= total count)
    uint32_t* tr_cursor;         // [n_groups*s]
    uint64_t* tr_pos;            // positions
    uint32_t sketch_size;
    uint32_t seed;
    int fold_case;               // !preserveCase
    int hash32;                  // !use64
    // hot-loop constants handed over as data ('ACGT' table for PRMT, Murmur block addends)
    uint32_t c_tbl;
    uint64_t c_add1, c_add2, c_add1s;   // c_add1s = c_add1 + 5 * seed (first Murmur block, sketch_tables.cuh)
    uint64_t pos_base;           // added to every stream position stored (a rank's part of one read set spread over GPUs)
    // survivor log (nullable): every (hash, position) that reached a table, so that the order-dependent top count can be
    // settled from the first pass instead of hashing the input a second time
    uint64_t* log_h;
    uint64_t* log_p;
    unsigned long long* log_count;
    uint64_t log_cap;
};

// ---------------------------------------------------------------------------------------
// ASCII -> (2-bit code, valid) for 32 bases.  Codes: A=0 C=1 G=2 T=3 so that numeric order
// of the big-endian packing equals memcmp order of the ASCII k-mer (Sketch.cpp:721).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t nt_codes4(uint32_t x)   // 4 ASCII bytes -> 8 bits, first base most significant
{
    uint32_t c = ((x >> 1) ^ (x >> 2)) & 0x03030303u;
    return (c * 0x40100401u) >> 24;
}

__device__ __forceinline__ uint32_t nt_valid4(uint32_t x, uint32_t fold_mask)
{
    // valid iff (byte & fold_mask) in {'A','C','G','T'}: high 3 bits == 010 and bit (byte&31)
    // of the set {1,3,7,20}.
    const uint32_t SET = (1u << 1) | (1u << 3) | (1u << 7) | (1u << 20);
    uint32_t f = x & fold_mask;
    uint32_t v = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        uint32_t y = (f >> (8 * b)) & 0xffu;
        uint32_t ok = ((y & 0xe0u) == 0x40u) & ((SET >> (y & 31u)) & 1u);
        v |= ok << b;
    }
    return v;
}

static __device__ __noinline__ uint4 load16_tail(const uint8_t* base, uint64_t off, uint64_t n)
{
    uint32_t w[4] = {0, 0, 0, 0};
    for (int i = 0; i < 16; i++)
        if (off + i < n) w[i >> 2] |= (uint32_t)base[off + i] << (8 * (i & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// 16 bytes at `off`; bytes at or beyond n read as 0x00 (never in an alphabet)
__device__ __forceinline__ uint4 load16_guarded(const uint8_t* base, uint64_t off, uint64_t n)
{
    if (off + 16 <= n) return *reinterpret_cast<const uint4*>(base + off);
    return load16_tail(base, off, n);
}

// Convert one tile (SK_TILE_CHUNKS x 32 bases starting at byte tile_base) into shared memory:
// s_code[2*c], s_code[2*c+1] = the 32 bases of chunk c, 16 per word, big-endian 2-bit;
// s_valid[c] bit j = base 32c+j is in the alphabet.
__device__ __forceinline__ void convert_tile(const uint8_t* seq, uint64_t n_bytes, uint64_t tile_base, int fold_case,
                                             uint32_t* s_code, uint32_t* s_valid)
{
    const uint32_t fold_mask = fold_case ? 0xdfdfdfdfu : 0xffffffffu;
    for (int c = threadIdx.x; c < SK_TILE_CHUNKS; c += SK_THREADS) {
        uint64_t off = tile_base + 32ull * c;
        uint4 a = load16_guarded(seq, off, n_bytes);
        uint4 b = load16_guarded(seq, off + 16, n_bytes);
        uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t code0 = 0, code1 = 0, valid = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            code0 |= nt_codes4(w[i]) << (24 - 8 * i);   // bit 5 (case) does not enter the code
            valid |= nt_valid4(w[i], fold_mask) << (4 * i);
        }
#pragma unroll
        for (int i = 4; i < 8; i++) {
            code1 |= nt_codes4(w[i]) << (24 - 8 * (i - 4));
            valid |= nt_valid4(w[i], fold_mask) << (4 * i);
        }
        s_code[2 * c] = code0;
        s_code[2 * c + 1] = code1;
        s_valid[c] = valid;
    }
}

// Reverse-complement of 16 packed bases: reverse the 2-bit fields, complement (c ^ 3).
__device__ __forceinline__ uint32_t revcomp16(uint32_t w)
{
    uint32_t y = __brev(w);
    y = ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
    return ~y;
}

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// 2-bit (left-aligned in hi:lo, first base in bits 63..62) -> zero-padded ASCII, as the four
// little-endian u64 words MurmurHash3 consumes.  PRMT does the 4-entry table lookup
// ("ACGT") four bases at a time; a selector nibble with bit 3 set replicates the sign bit
// of an ASCII byte, i.e. yields 0x00, which zero-pads bytes >= K for free.
template <int K>
__device__ __forceinline__ void expand_ascii(uint32_t hi, uint32_t lo, uint64_t (&w)[4], uint32_t TBL = 0x54474341u /* 'A','C','G','T' */)
{
    uint32_t e[4], o[4];
    uint32_t ehi = hi & 0x33333333u, ohi = (hi >> 2) & 0x33333333u;
    uint32_t elo = lo & 0x33333333u, olo = (lo >> 2) & 0x33333333u;
    e[0] = ehi >> 16; o[0] = ohi >> 16; e[1] = ehi; o[1] = ohi;
    e[2] = elo >> 16; o[2] = olo >> 16; e[3] = elo; o[3] = olo;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const int first = 8 * c;                 // first base of this chunk
        if (first >= K) { w[c] = 0; continue; }
        uint32_t E = prmt(TBL, 0, e[c]);  // bases first+7, +5, +3, +1
        uint32_t O = prmt(TBL, 0, o[c]);  // bases first+6, +4, +2, +0
        // byte j of the low word is base first+j, of the high word base first+4+j
        uint32_t sel_lo = 0, sel_hi = 0;
        const uint32_t lo_idx[4] = {7, 3, 6, 2};
        const uint32_t hi_idx[4] = {5, 1, 4, 0};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            sel_lo |= ((first + j < K) ? lo_idx[j] : 8u) << (4 * j);
            sel_hi |= ((first + 4 + j < K) ? hi_idx[j] : 8u) << (4 * j);
        }
        uint32_t wl = prmt(E, O, sel_lo);
        uint32_t wh = (first + 4 < K) ? prmt(E, O, sel_hi) : 0u;
        w[c] = ((uint64_t)wh << 32) | wl;
    }
}

// The same expansion through a shared-memory table: the canonical k-mer is left-aligned in hi:lo, so every group of
// four bases is one BYTE of hi/lo; lut[v] holds the four ASCII letters of byte value v (first base in the low
// byte), lut[256 + v] only the first K % 4 of them (zero-padded: the Murmur tail).  Per group: one PRMT (byte
// extract, ALU pipe), one IMAD (table address, FMA pipe) and one LDS.32 -- against ~4 ALU-pipe instructions per
// group for expand_ascii.  The sketch kernel is bound by the ALU pipe and leaves the LSU pipe idle.
__device__ __forceinline__ void build_expand_lut(uint32_t* lut, int K, int tid, int nthreads)
{
    for (int v = tid; v < 512; v += nthreads) {
        const int nb = v < 256 ? 4 : (K & 3);
        uint32_t e = 0;
        for (int i = 0; i < nb; i++) e |= (uint32_t)"ACGT"[((v & 255) >> (6 - 2 * i)) & 3] << (8 * i);
        lut[v] = e;
    }
}

template <int K>
__device__ __forceinline__ void expand_lut(uint32_t hi, uint32_t lo, uint64_t (&w)[4], uint32_t lut_addr /* shared-space byte address */)
{
    uint32_t v[8];
#pragma unroll
    for (int g = 0; g < 8; g++) {
        if (4 * g >= K) { v[g] = 0; continue; }
        const uint32_t src = g < 4 ? hi : lo;
        const uint32_t idx = prmt(src, 0u, 0x4440u | (3 - (g & 3)));           // byte 3 holds the first four bases
        const bool partial = 4 * g + 4 > K;
        uint32_t addr;
        asm("mad.lo.u32 %0, %1, 4, %2;" : "=r"(addr) : "r"(idx), "r"(lut_addr + (partial ? 1024u : 0u)));
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[g]) : "r"(addr));
    }
#pragma unroll
    for (int c = 0; c < 4; c++) w[c] = ((uint64_t)v[2 * c + 1] << 32) | v[2 * c];
}

__device__ __forceinline__ uint32_t find_group(const uint64_t* group_off, uint32_t lo, uint32_t hi, uint64_t pos)
{
    // largest g in [lo,hi] with group_off[g] <= pos
    while (lo < hi) {
        uint32_t mid = (lo + hi + 1) >> 1;
        if (group_off[mid] <= pos) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ uint32_t table_slot(uint64_t h, uint32_t mask)
{
    return ((uint32_t)h ^ (uint32_t)(h >> 32) * 0x9e3779b1u) & mask;
}

// Rare path: a window whose hash is below its block's threshold bound.
static __device__ __noinline__ void sketch_emit(const SketchArgs& a, uint64_t h, uint64_t pos, uint32_t g_lo, uint32_t g_hi, int trace)
{
    uint32_t g = find_group(a.group_off, g_lo, g_hi, pos);
    if (!a.active[g] || h > a.thresh[g]) return;
    if (trace) {
        const uint64_t* fin = a.fin_hashes + (uint64_t)g * a.sketch_size;
        uint32_t n = a.fin_n[g], lo = 0, hi = n;
        while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (fin[mid] < h) lo = mid + 1; else hi = mid; }
        if (lo < n && fin[lo] == h) {
            uint64_t b = (uint64_t)g * a.sketch_size + lo;
            uint32_t idx = atomicAdd(&a.tr_cursor[b], 1u);
            if (idx < a.tr_cap[b]) a.tr_pos[a.tr_off[b] + idx] = pos + a.pos_base;
        }
        return;
    }
    if (a.log_h) {
        // warp-aggregated append: one atomic per group of converged lanes
        const unsigned m = __activemask();
        const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
        unsigned long long at = 0;
        if (lane == leader) at = atomicAdd(a.log_count, (unsigned long long)__popc(m));
        at = __shfl_sync(m, at, leader) + __popc(m & ((1u << lane) - 1u));
        if (at < a.log_cap) { a.log_h[at] = h; a.log_p[at] = pos; }
    }
    if (h == SK_EMPTY) {   // the one value that cannot be a table key
        atomicAdd(&a.maxkey_cnt[g], 1u);
        atomicMin((unsigned long long*)&a.maxkey_pos[g], (unsigned long long)(pos + a.pos_base));
        return;
    }
    const uint64_t base = a.toff[g];
    const uint32_t mask = a.tmask[g];
    uint32_t slot = table_slot(h, mask);
    for (uint32_t probe = 0; probe <= mask; probe++) {
        unsigned long long prev = atomicCAS((unsigned long long*)&a.tkeys[base + slot], (unsigned long long)SK_EMPTY, (unsigned long long)h);
        if (prev == SK_EMPTY || prev == h) {
            atomicAdd(&a.tcnt[base + slot], 1u);
            atomicMin((unsigned long long*)&a.tpos[base + slot], (unsigned long long)(pos + a.pos_base));
            return;
        }
        slot = (slot + 1) & mask;
    }
    atomicExch(&a.overflow[g], 1u);
}

// Valid windows per group (the "k-mers sketched" unit).  One thread per 32-base chunk.
template <int K>
__global__ void __launch_bounds__(SK_THREADS) count_windows_kernel(const SketchArgs* __restrict__ ga, unsigned long long* out_kmers)
{
    const SketchArgs& a = *ga;
    __shared__ uint32_t s_code[2 * SK_TILE_CHUNKS + 2];
    __shared__ uint32_t s_valid[SK_TILE_CHUNKS + 1];
    __shared__ uint32_t s_g[2];
    const uint64_t tile_base = (uint64_t)blockIdx.x * SK_TILE_WINDOWS;
    convert_tile(a.seq, a.n_bytes, tile_base, a.fold_case, s_code, s_valid);
    if (threadIdx.x == 0) {
        uint64_t last = tile_base + SK_TILE_WINDOWS - 1;
        if (last >= a.n_bytes) last = a.n_bytes - 1;
        s_g[0] = find_group(a.group_off, 0, a.n_groups - 1, tile_base);
        s_g[1] = find_group(a.group_off, s_g[0], a.n_groups - 1, last);
        s_valid[SK_TILE_CHUNKS] = 0;
    }
    __syncthreads();
    const int c = threadIdx.x;                       // chunk of 32 windows
    const uint64_t chunk_pos = tile_base + 32ull * c;
    // (no early exit for chunks beyond the input: every lane takes part in the warp reduction below)
    const bool in_range = chunk_pos < a.n_bytes;
    // window i valid iff valid bits [i, i+K) all set: AND-shift doubling on 64 bits
    uint64_t v = in_range ? ((uint64_t)s_valid[c + 1] << 32) | s_valid[c] : 0;
    int have = 1;
    // build run-of-K mask by binary decomposition of K
    uint64_t acc = ~0ULL;
    int done = 0;
    uint64_t pw = v;       // pw = run mask of length `have`
#pragma unroll
    for (int bit = 0; bit < 6; bit++) {
        if (K & (1 << bit)) { acc &= (pw >> done); done += have; }
        pw = pw & (pw >> have);
        have <<= 1;
    }
    uint32_t wv = (uint32_t)acc;                     // bit i = window chunk_pos+i valid
    if (s_g[0] == s_g[1]) {
        uint32_t n = __popc(wv);
        // warp-aggregate, one atomic per warp
        for (int o = 16; o; o >>= 1) n += __shfl_down_sync(0xffffffffu, n, o);
        if ((threadIdx.x & 31) == 0 && n) atomicAdd(&out_kmers[s_g[0]], (unsigned long long)n);
    } else {
        while (wv) {
            int i = __ffs(wv) - 1;
            wv &= wv - 1;
            uint32_t g = find_group(a.group_off, s_g[0], s_g[1], chunk_pos + i);
            atomicAdd(&out_kmers[g], 1ULL);
        }
    }
}

}  // namespace fpm
