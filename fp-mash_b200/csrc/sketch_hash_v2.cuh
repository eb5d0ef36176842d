// sketch_hash_v2.cuh -- the production sketching kernel (round-1 second version).
//
// Changes against the first version (kept in sketch_kernels.cuh for the test/count kernels), each
// motivated by the ncu capture in profiles/r01_sketch_hash.txt:
//   * warp-autonomous: a warp loads 64 bases per lane, converts them in registers and gets its
//     right-hand halo from the next lane with three shuffles -- no shared memory, no
//     __syncthreads (the "barrier" stall was 1.6 warps/issue);
//   * 4-window unrolling with byte-wise shifting of the forward / reverse-complement registers
//     between sub-blocks instead of 16 fully unrolled windows: ~10 KB of code instead of 55 KB
//     (the top stall was "no_instruction", i.e. instruction-cache misses);
//   * validity computed four bases at a time (re-expand the 2-bit code with PRMT and compare with
//     the input bytes) instead of a per-byte test;
//   * each warp walks a contiguous run of tiles and caches the sketch it is in, so the per-tile
//     group lookup is a register compare except at sketch boundaries.
// Third pass (profiles/r01_sketch_hash_v3.txt; the kernel is bound by the ALU pipe -- SHF/LOP3/PRMT -- while the
// LSU pipe idles and the FMA pipe has slots left; an experiment with a free expansion ran 18 % faster):
//   * the 2-bit -> ASCII expansion reads a 512-entry shared-memory table, one byte of packed codes = four
//     letters per LDS.32 (expand_lut): 6 ALU + 6 FMA + 6 LSU instructions instead of ~24 ALU per window;
//   * 64-bit constant multiplies are pinned at three IMADs (mul64c);
//   * the per-window threshold reject is ONE compare (the exact 64-bit test sits behind a volatile shared load
//     the compiler cannot hoist), and the survivor-queue check runs only when a vote says a lane queued something.
//   Tried and dropped: right shifts as IMAD.HI (the FMA pipe runs it at a fraction of IMAD's rate: 169 vs 188
//   Gk-mers/s), x*5+c as mad.wide asm (blocks the fold with the preceding add: +3 instructions).
//
// Work decomposition: warp tile = 63 chunks of 32 windows (2016 windows).  Lane l converts
// chunks 2l and 2l+1; lanes 0..30 hash both, lane 31 hashes only its first chunk -- its second
// chunk is the halo of lane 30 and the first chunk of the next warp tile.
#pragma once
#include "sketch_kernels.cuh"
#include "sketch_tables.cuh"

namespace fpm {

constexpr int WT_WINDOWS = 63 * 32;        // windows per warp tile
constexpr int WT_TILES_PER_WARP = 12;      // contiguous warp tiles walked by one warp

// 32 ASCII bases (8 words) -> two big-endian 2-bit code words.  Per word one multiply gathers the four codes into
// the top byte; three PRMTs assemble four such bytes into a register.  (Validity is not tracked here: the hash of a
// window holding a letter outside ACGT is garbage that passes the threshold with probability ~1e-4 like any other,
// and the survivors' letters are re-read from memory -- window_is_valid below -- before anything is inserted.)
__device__ __forceinline__ uint32_t nt_codes4_top(uint32_t x)     // codes of 4 bases in bits 31..24, first base highest
{
    const uint32_t c = ((x >> 1) ^ (x >> 2)) & 0x03030303u;
    return c * 0x40100401u;
}

__device__ __forceinline__ uint32_t gather_top_bytes(uint32_t t0, uint32_t t1, uint32_t t2, uint32_t t3)
{
    const uint32_t u = prmt(t1, t0, 0x0073u);      // byte1 = t0's top byte, byte0 = t1's
    const uint32_t v = prmt(t3, t2, 0x0073u);
    return prmt(v, u, 0x5410u);                    // t0.b3 : t1.b3 : t2.b3 : t3.b3
}

__device__ __forceinline__ void convert32(const uint32_t (&w)[8], uint32_t& c0, uint32_t& c1)
{
    c0 = gather_top_bytes(nt_codes4_top(w[0]), nt_codes4_top(w[1]), nt_codes4_top(w[2]), nt_codes4_top(w[3]));
    c1 = gather_top_bytes(nt_codes4_top(w[4]), nt_codes4_top(w[5]), nt_codes4_top(w[6]), nt_codes4_top(w[7]));
}

// Do the K bytes at `pos` all belong to ACGT (after the optional case fold)?  Rare path: only windows whose hash passed
// the sketch's threshold get here.  Reads the input again (L1/L2 hits: the tile was loaded moments ago).
static __device__ __noinline__ bool window_is_valid(const uint8_t* __restrict__ seq, uint64_t pos, int k, uint64_t n_bytes, uint32_t fold_mask)
{
    if (pos + (uint64_t)k > n_bytes) return false;
    for (int i = 0; i < k; i++) {
        const uint32_t b = seq[pos + i] & (fold_mask & 0xffu);
        if (b != 'A' && b != 'C' && b != 'G' && b != 'T') return false;
    }
    return true;
}

constexpr uint32_t SQ_CAP = 128;     // per-warp survivor queue (hash, position)
constexpr uint32_t SQ_FLUSH = 24;    // flush at a tile boundary once this many are waiting

// All 32 lanes insert queued survivors in parallel (one table round trip per 32 hashes).  Entries may stem
// from earlier tiles, so the sketch is looked up over the whole group table.
// The letters of a queued window are checked here, by all 32 lanes at once, and not where the window was found: there one lane
// would run the byte loop while 31 wait, and with large sketches (s = 10 000: 0.4 % of the windows pass the bound) some lane of
// a warp has a candidate in 40 % of the four-window steps.
__device__ __forceinline__ void flush_survivors(const SketchArgs& a, const uint64_t* qh, const uint64_t* qp, uint32_t* qn, int lane, int trace, int k, uint32_t fold_mask)
{
    uint32_t n = *qn;
    if (n > SQ_CAP) n = SQ_CAP;
    for (uint32_t i = lane; i < n; i += 32)
        if (window_is_valid(a.seq, qp[i], k, a.n_bytes, fold_mask)) sketch_emit(a, qh[i], qp[i], 0, a.n_groups - 1, trace);
    __syncwarp();
    if (lane == 0) *qn = 0;
    __syncwarp();
}

#ifndef FPM_SK_MIN_CTAS
#define FPM_SK_MIN_CTAS 3   // 80 registers, no spills: round 2's table-driven body at k=21 runs 273 Gk-mers/s at 3 CTAs/SM, 263 at 4 (64 registers, spills), 262 at 2
#endif
template <int K, bool CANON>
__global__ void __launch_bounds__(SK_THREADS, FPM_SK_MIN_CTAS) sketch_hash_kernel_v2(const SketchArgs* __restrict__ ga, uint64_t range_lo, uint64_t range_hi,
                                                                   uint64_t range_base, int trace)
{
    const SketchArgs& a = *ga;
    const uint8_t* __restrict__ seq = a.seq;
    const uint64_t n_bytes = a.n_bytes;
    const uint32_t seed = a.seed;
    constexpr bool hash32 = K <= 16;               // 4^K <= 2^32 (Sketch.cpp:1288); the host checks a.hash32 agrees
    const uint32_t fold_mask = a.fold_case ? 0xdfdfdfdfu : 0xffffffffu;
    // constants kept in registers for the whole kernel: as immediates they cost moves per window
    // (read from the argument block so that ptxas cannot fold them back into immediates)
    const uint64_t add1 = a.c_add1, add2 = a.c_add2, add1s = a.c_add1s;
    __shared__ uint64_t s_qh[SK_THREADS / 32][SQ_CAP], s_qp[SK_THREADS / 32][SQ_CAP];
    __shared__ uint32_t s_qn[SK_THREADS / 32];
    __shared__ uint64_t s_tm[SK_THREADS / 32];      // the warp's current threshold bound (see the filter below)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __shared__ KmerTables<K> s_tab;                 // Murmur's first multiply per word / the tail lane as table entries (sketch_tables.cuh)
    build_kmer_tables<K>(&s_tab, threadIdx.x, SK_THREADS);
    uint32_t lut_addr = (uint32_t)__cvta_generic_to_shared(&s_tab);
    asm volatile("" : "+r"(lut_addr));     // keep the table address in a register: re-deriving it costs ~4 instructions per window
    __syncthreads();
    if (lane == 0) s_qn[wid] = 0;
    __syncwarp();
    const uint64_t warp = (uint64_t)blockIdx.x * (SK_THREADS / 32) + (threadIdx.x >> 5);
    const uint64_t n_tiles = (range_hi - range_base + WT_WINDOWS - 1) / WT_WINDOWS;
    uint64_t t0 = warp * WT_TILES_PER_WARP, t1 = t0 + WT_TILES_PER_WARP;
    if (t1 > n_tiles) t1 = n_tiles;

    // the sketch ("group") this warp is currently inside; [g_begin, g_end) in bytes
    uint32_t g_lo = 0, g_hi = 0;
    uint64_t g_begin = 1, g_end = 0, tmax = 0;   // empty interval: forces a lookup on the first tile
    bool uniform_ok = false;

    for (uint64_t t = t0; t < t1; t++) {
        const uint64_t tile_base = range_base + t * WT_WINDOWS;
        uint64_t last = tile_base + WT_WINDOWS - 1;
        if (last >= n_bytes) last = n_bytes - 1;
        if (!(uniform_ok && tile_base >= g_begin && last < g_end)) {
            // leaving the cached sketch: lane 0 looks the tile's sketches up, all lanes take the answer
            uint32_t lo = 0, hi = 0;
            uint64_t tm = 0, gb = 1, ge = 0;
            int uni = 0;
            if (lane == 0) {
                lo = find_group(a.group_off, 0, a.n_groups - 1, tile_base);
                hi = find_group(a.group_off, lo, a.n_groups - 1, last);
                for (uint32_t g = lo; g <= hi; g++)
                    if (a.active[g] && a.thresh[g] > tm) tm = a.thresh[g];
                bool any = false;
                for (uint32_t g = lo; g <= hi; g++) any |= a.active[g] != 0;
                if (!any) tm = 0;
                if (lo == hi) { uni = 1; gb = a.group_off[lo]; ge = a.group_off[lo + 1]; }
                if (!any) uni |= 2;
            }
            g_lo = __shfl_sync(0xffffffffu, lo, 0);
            g_hi = __shfl_sync(0xffffffffu, hi, 0);
            tmax = __shfl_sync(0xffffffffu, tm, 0);
            g_begin = __shfl_sync(0xffffffffu, gb, 0);
            g_end = __shfl_sync(0xffffffffu, ge, 0);
            uni = __shfl_sync(0xffffffffu, uni, 0);
            uniform_ok = (uni & 1) != 0;
            if (uni & 2) { if (!uniform_ok) continue; tmax = 0; }
        }
        const bool idle_tile = uniform_ok && tmax == 0 && !a.active[g_lo];
        if (idle_tile) continue;

        const uint32_t tcut = hash32 ? (uint32_t)tmax : (uint32_t)(tmax >> 32);
        // lazy finish: hi(a1) + hi(a2) + 1 is the hash's high word or one above it (carry from the low words unknown)
        const uint32_t tcut1 = tcut == 0xffffffffu ? tcut : tcut + 1u;
        if (lane == 0) s_tm[wid] = tmax;
        __syncwarp();

        // ---- load + convert 64 bases per lane -------------------------------------------------
        const uint64_t lane_pos = tile_base + 64ull * lane;
        uint32_t q0, q1, q2, q3, q4, q5;
        {
            uint4 l0 = load16_guarded(seq, lane_pos, n_bytes), l1 = load16_guarded(seq, lane_pos + 16, n_bytes);
            uint4 l2 = load16_guarded(seq, lane_pos + 32, n_bytes), l3 = load16_guarded(seq, lane_pos + 48, n_bytes);
            const uint32_t wa[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
            const uint32_t wb[8] = {l2.x, l2.y, l2.z, l2.w, l3.x, l3.y, l3.z, l3.w};
            convert32(wa, q0, q1);
            convert32(wb, q2, q3);
        }
        q4 = __shfl_down_sync(0xffffffffu, q0, 1);
        q5 = __shfl_down_sync(0xffffffffu, q1, 1);
        // (no lane leaves the tile early: the survivor queue below is flushed by the whole warp)
        const int nblk = lane_pos >= n_bytes ? 0 : (lane == 31 ? 2 : 4);

        bool queued = false;
#pragma unroll 1
        for (int blk = 0; blk < 4; blk++) {           // uniform trip count: the warp meets at the flush points
            const bool live = blk < nblk;
            uint32_t fw0 = q0, fw1 = q1, fw2 = q2;
            uint32_t x[5];
            if (CANON) { x[0] = revcomp16(fw2); x[1] = revcomp16(fw1); x[2] = revcomp16(fw0); x[3] = 0; x[4] = 0; }
#ifndef FPM_WIN
#define FPM_WIN 4       // windows in flight per lane (unrolled); 16 / FPM_WIN sub-blocks per 16-window block
#endif
#pragma unroll 1
            for (int sub = 0; sub < 16 / FPM_WIN; sub++) {
                // FPM_WIN windows in flight; ONE filter branch for all of them (a branch per window cost ~16 % of the
                // stall samples: ISETP waiting for the end of the Murmur chain, then branch resolution)
                constexpr bool lazy = FPM_LAZYFIN && !hash32;
                uint64_t hh[FPM_WIN], hb[lazy ? FPM_WIN : 1];
                bool any_below = false;
                if (live) {
#pragma unroll
                    for (int j = 0; j < FPM_WIN; j++) {
                        uint32_t fhi = j ? __funnelshift_l(fw1, fw0, 2 * j) : fw0;
                        uint32_t flo = j ? __funnelshift_l(fw2, fw1, 2 * j) : fw1;
                        uint32_t chi = fhi, clo = flo;
                        if (CANON) {
                            const int start = 48 - K - j;                 // constant: the rc registers move instead
                            const int ia = start >> 4, sh = 2 * (start & 15);
                            uint32_t rhi = sh ? __funnelshift_l(x[ia + 1], x[ia], sh) : x[ia];
                            uint32_t rlo = sh ? __funnelshift_l(x[ia + 2], x[ia + 1], sh) : x[ia + 1];
                            // bits below the k-mer only matter for palindromes, where both strands hash alike
                            bool use_r = (((uint64_t)rhi << 32) | rlo) < (((uint64_t)fhi << 32) | flo);
                            chi = use_r ? rhi : fhi;
                            clo = use_r ? rlo : flo;
                        }
                        if (lazy) {
                            kmer_hash_parts<K>(chi, clo, lut_addr, seed, add1s, add1, add2, hh[j], hb[j]);
                            any_below |= (uint32_t)(hh[j] >> 32) + (uint32_t)(hb[j] >> 32) + 1u <= tcut1;
                        } else {
                            uint64_t a1_, a2_;
                            kmer_hash_parts<K>(chi, clo, lut_addr, seed, add1s, add1, add2, a1_, a2_);
                            uint64_t h = kmer_hash_finish(a1_, a2_);
                            if (hash32) h &= 0xffffffffULL;
                            hh[j] = h;
                            // one-instruction reject on the deciding word; the exact 64-bit test only for the survivors
                            any_below |= (hash32 ? (uint32_t)h : (uint32_t)(h >> 32)) <= tcut;
                        }
                    }
                }
                if (any_below) {
#pragma unroll
                    for (int j = 0; j < FPM_WIN; j++) {
                        const uint64_t h = lazy ? kmer_hash_finish(hh[j], hb[lazy ? j : 0]) : hh[j];
                        // the exact 64-bit test reads its bound back through a volatile shared load, which the compiler
                        // cannot speculate above the branch
                        if (h > *(volatile uint64_t*)&s_tm[wid]) continue;
                        const int b = 16 * blk + FPM_WIN * sub + j;              // window index within the lane's 64
                        const uint64_t pos = lane_pos + b;
                        if (pos >= range_lo && pos < range_hi) {
                            // survivors are queued per warp and inserted 32 at a time: the table atomics cost a
                            // ~1 us round trip that would otherwise stall the whole warp for one lane's hash
                            queued = true;
                            const uint32_t qi = atomicAdd(&s_qn[wid], 1u);
                            if (qi < SQ_CAP) { s_qh[wid][qi] = h; s_qp[wid][qi] = pos; }
                            else if (window_is_valid(seq, pos, K, n_bytes, fold_mask)) sketch_emit(a, h, pos, g_lo, g_hi, trace);      // queue full (accept-all sketches)
                        }
                    }
                }
                // next four windows: forward registers one byte left, reverse-complement one byte right
                fw0 = __funnelshift_l(fw1, fw0, 2 * FPM_WIN); fw1 = __funnelshift_l(fw2, fw1, 2 * FPM_WIN); fw2 <<= 2 * FPM_WIN;
                if (CANON) { x[2] = __funnelshift_r(x[2], x[1], 2 * FPM_WIN); x[1] = __funnelshift_r(x[1], x[0], 2 * FPM_WIN); x[0] >>= 2 * FPM_WIN; }
            }
            // dense survivors (accept-all sketches of short records): insert once a warp-load is waiting.  One vote per
            // 16-window block; the queue can only have grown if some lane queued a survivor (entries beyond the queue's
            // capacity were inserted directly).
            if (__any_sync(0xffffffffu, queued)) {
                queued = false;
                if (s_qn[wid] >= 32) flush_survivors(a, s_qh[wid], s_qp[wid], &s_qn[wid], lane, trace, K, fold_mask);
            }
            q0 = q1; q1 = q2; q2 = q3; q3 = q4; q4 = q5;
        }
        __syncwarp();
        if (s_qn[wid] >= SQ_FLUSH) flush_survivors(a, s_qh[wid], s_qp[wid], &s_qn[wid], lane, trace, K, fold_mask);
    }
    __syncwarp();
    if (s_qn[wid]) flush_survivors(a, s_qh[wid], s_qp[wid], &s_qn[wid], lane, trace, K, fold_mask);
}

}  // namespace fpm
