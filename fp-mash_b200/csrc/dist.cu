// dist.cu -- all-vs-all sketch comparison on sm_100a.
//
// Replaces compare() -> compareSketches() -> pValue() (CommandDistance.cpp:335-450).
//
// Kernels:
//   dist_tile32_kernel    the fast path for ascending duplicate-free lists: 32 queries x 32 references per CTA on the
//                         32-bit dense ranks of dist_rank.cu, two interleaved merges per thread, value-bounded phases,
//                         steps without end tests.  Merges only the pairs whose sketches share a hash (bitmaps from
//                         dist_rank.cu); results as a matrix (dist_fill_unshared_kernel streams the closed-form records
//                         of all other pairs first when the panels were grouped) or as a list of the pairs that pass
//                         the -d / -v filters (HitSink, fpm_dist_hits).  DESIGN.md 4.3.
//   dist_tile_kernel      the first fast path, kept for panels beyond 2^31 hashes and as the implementation the rank
//                         kernel is compared against.  One CTA owns a
//                         16-query x 32-reference tile (512 pairs, one thread each).  The 48
//                         sketches are staged in shared memory as lane-private COLUMNS
//                         (element p of sketch c at [p][c]), so the data-dependent reads of a
//                         sequential merge are bank-conflict free whatever position each lane
//                         has reached.  Sketches are streamed in phases of DT_ROWS rows
//                         bounded by value: V = min over the 48 columns of the first element
//                         that did not fit, so every element < V of every column is resident
//                         and each pair can merge up to V independently; elements >= V are
//                         masked to the +inf sentinel, which makes "phase exhausted" and
//                         "list exhausted" the same cheap test.  Pairs that reach denom == s
//                         stop early (for unrelated same-size genomes one phase suffices).
//   dist_literal_kernel   one thread per pair, the reference loop executed literally from
//                         row-major panels in global memory.  Defines the result for ANY
//                         input (fp-mode lists are unsorted and may repeat) and
//                         is the fallback when the fast path's preconditions do not hold.
//   dist_collect_kernel   hits mode of the two kernels above: scans their matrix for passing pairs.
//   fp_positional_kernel  the fork's positional fingerprint comparison (mash triangle -fp).
//
// This is set intersection, not a contraction: no tensor cores.  The bound is integer
// compare / shared-memory throughput.
#include <algorithm>
#include <string.h>
#include "common.h"
#include "dist_math.h"
#include "dist_rank.h"
#include "dist_internal.h"

namespace fpm {

constexpr uint64_t DT_INF = ~0ULL;
constexpr int DT_Q = 16;            // query columns per CTA
constexpr int DT_R = 32;            // reference columns per CTA (two 16-column planes)
constexpr int DT_COLS = DT_Q + DT_R;
constexpr int DT_ROWS = 288;        // rows resident per phase
constexpr int DT_PAD = 9;           // +inf rows after them: a pointer rests at most on row DT_ROWS and the unchecked blocks look 8 rows ahead
constexpr int DT_COLROWS = DT_ROWS + DT_PAD;   // 297 rows x 48 columns x 8 B = 114.0 KB: two CTAs per SM
constexpr int DT_THREADS = DT_Q * DT_R;   // 512

struct DistArgs {
    uint32_t s;
    int kmer_size;
    double kmer_space, max_distance, max_pvalue;
    uint32_t k128 = 128u, k1 = 1u;   // the row pitch and 1 as run-time values (FPM_D4_FMA_BUMP: a multiply-add ptxas cannot fold into an add)
};

// (the "memory" clobbers matter: these loads read what other threads staged, so they must stay behind the
// __syncthreads() that publishes it -- without the clobber the compiler is free to move them)
__device__ __forceinline__ void lds64(uint32_t addr, uint32_t& lo, uint32_t& hi)
{
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(addr) : "memory");
}

__device__ __forceinline__ void lds32(uint32_t addr, uint32_t& v)
{
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
}

// One step of the reference's merge loop (CommandDistance.cpp:378-386) on 32-bit ranks held in shared-memory
// columns of 128-byte row pitch: the list(s) holding the smaller-or-equal head advance and reload.  Pinned as
// six instructions (two compares, two predicated pointer bumps, two predicated LDS.32).
#ifndef FPM_D4_FMA_BUMP
#define FPM_D4_FMA_BUMP 0     // pointer bumps as IMAD (FMA pipe) instead of IADD (ALU pipe, which also carries the two compares)
#endif
__device__ __forceinline__ void merge_step32(uint32_t& pa, uint32_t& pb, uint32_t& av, uint32_t& bv, uint32_t c128 = 128u, uint32_t c1 = 1u)
{
#if FPM_D4_FMA_BUMP
    asm volatile("{\n\t.reg .pred pa_le, pb_le;\n\t"
                 "setp.le.u32 pa_le, %2, %3;\n\t"
                 "setp.ge.u32 pb_le, %2, %3;\n\t"
                 "@pa_le mad.lo.u32 %0, %4, %5, %0;\n\t"
                 "@pb_le mad.lo.u32 %1, %4, %5, %1;\n\t"
                 "@pa_le ld.shared.u32 %2, [%0];\n\t"
                 "@pb_le ld.shared.u32 %3, [%1];\n\t}"
                 : "+r"(pa), "+r"(pb), "+r"(av), "+r"(bv)
                 : "r"(c128), "r"(c1)
                 : "memory");
    return;
#endif
    asm volatile("{\n\t.reg .pred pa_le, pb_le;\n\t"
                 "setp.le.u32 pa_le, %2, %3;\n\t"
                 "setp.ge.u32 pb_le, %2, %3;\n\t"
                 "@pa_le add.u32 %0, %0, 128;\n\t"
                 "@pb_le add.u32 %1, %1, 128;\n\t"
                 "@pa_le ld.shared.u32 %2, [%0];\n\t"
                 "@pb_le ld.shared.u32 %3, [%1];\n\t}"
                 : "+r"(pa), "+r"(pb), "+r"(av), "+r"(bv)
                 :
                 : "memory");
}

__device__ __forceinline__ void finish_pair(const DistArgs& a, uint64_t common, uint64_t denom, uint64_t len_ref, uint64_t len_qry, fpm_pair* out)
{
    double d = mash_distance(common, denom, a.kmer_size);
    fpm_pair o;
    o.numer = (uint32_t)common;
    o.denom = (uint32_t)denom;
    o.distance = d;
    o.pvalue = 0.;
    bool pass = !(a.max_distance >= 0 && d > a.max_distance);          // CommandDistance.cpp:416-419
    if (pass) {
        o.pvalue = mash_pvalue(common, len_ref, len_qry, a.kmer_space, denom);
        if (a.max_pvalue >= 0 && o.pvalue > a.max_pvalue) pass = false; // :425-428
    }
    if (pass) o.denom |= FPM_PAIR_PASS;
    *out = o;
}

// Called by whole converged warps: one atomicAdd per warp, the lanes with a hit write 32-byte records behind each other.
__device__ __forceinline__ void append_hit(const HitSink& hs, bool pass, uint32_t q, uint32_t r, const fpm_pair& o)
{
    const uint32_t m = __ballot_sync(0xffffffffu, pass);
    if (!m) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(hs.count, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    const unsigned long long at = base + __popc(m & ((1u << lane) - 1u));
    if (pass && at < hs.cap) {
        uint4* dst = reinterpret_cast<uint4*>(hs.buf + at);
        dst[0] = make_uint4(q + hs.q_base, r + hs.r_base, o.numer, o.denom);   // indices in the caller's panels (a rank's block: dist_multi.cu)
        dst[1] = make_uint4((uint32_t)__double_as_longlong(o.distance), (uint32_t)(__double_as_longlong(o.distance) >> 32),
                            (uint32_t)__double_as_longlong(o.pvalue), (uint32_t)(__double_as_longlong(o.pvalue) >> 32));
    }
}

// hits mode of the paths that produce a matrix (64-bit tile kernel, literal kernel): scan it for passing pairs
__global__ void __launch_bounds__(256) dist_collect_kernel(const fpm_pair* __restrict__ mat, uint64_t total, uint64_t n_ref, HitSink hs)
{
    const uint64_t stride = (uint64_t)gridDim.x * 256;
    for (uint64_t base = (uint64_t)blockIdx.x * 256; base < total; base += stride) {   // uniform per CTA: warps stay converged
        const uint64_t p = base + threadIdx.x;
        fpm_pair o = {0, 0, 0., 0.};
        if (p < total && (mat[p].denom & FPM_PAIR_PASS)) o = mat[p];
        append_hit(hs, (o.denom & FPM_PAIR_PASS) != 0, (uint32_t)(p / n_ref), (uint32_t)(p % n_ref), o);   // (bases added inside)
    }
}

// The record of a pair whose sketches share no hash, without going through finish_pair: common = 0, denom = min(s, |A| + |B|)
// (CommandDistance.cpp:376-400 with no equal elements); distance 1, or 0 when both sketches are empty (common == denom);
// pValue(0, ...) = 1 (CommandDistance.cpp:435); the -d / -v tests of :416-428.  Needs no lengths and no logarithm.
__device__ __forceinline__ fpm_pair unshared_record(const DistArgs& a, uint32_t size_qry, uint32_t size_ref)
{
    const uint64_t un = (uint64_t)size_qry + size_ref;
    fpm_pair o;
    o.numer = 0;
    o.denom = (uint32_t)(un < a.s ? un : a.s);
    o.distance = o.denom == 0 ? 0. : 1.;
    bool pass = !(a.max_distance >= 0 && o.distance > a.max_distance);
    o.pvalue = pass ? 1. : 0.;
    if (pass && a.max_pvalue >= 0 && 1. > a.max_pvalue) pass = false;
    if (pass) o.denom |= FPM_PAIR_PASS;
    return o;
}

// Grouped panels scatter the tile kernel's 24-byte records over the matrix (partial sectors, written at different times),
// and 19 of 20 records of a pruned all-vs-all are the closed-form ones of pairs without a shared hash.  So those are
// written first, for ALL pairs, as one coalesced stream in the matrix's own order; the tile kernel then only overwrites
// the pairs it merged.
__global__ void __launch_bounds__(256) dist_fill_unshared_kernel(DistArgs a, uint64_t n_ref, uint64_t total, const uint32_t* __restrict__ size_ref,
                                                                 const uint32_t* __restrict__ size_qry, const uint64_t* __restrict__ len_ref,
                                                                 const uint64_t* __restrict__ len_qry, fpm_pair* __restrict__ out)
{
    // two records = 48 bytes = three 16-byte stores per thread
    // (one 64-bit division per CTA, not two per thread: the divisions were what bounded this kernel)
    __shared__ uint64_t s_q0, s_r0;
    if (threadIdx.x == 0) { const uint64_t p0 = (uint64_t)blockIdx.x * 512; s_q0 = p0 / n_ref; s_r0 = p0 - s_q0 * n_ref; }
    __syncthreads();
    const uint64_t p = 2 * ((uint64_t)blockIdx.x * 256 + threadIdx.x);
    if (p >= total) return;
    fpm_pair rec[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint64_t pp = p + h < total ? p + h : p;
        uint64_t q = s_q0, r = s_r0 + (pp - (uint64_t)blockIdx.x * 512);
        if (r >= n_ref) {
            if (n_ref >= 512) { r -= n_ref; q++; }                        // a CTA's 512 pairs span at most two rows
            else { q = pp / n_ref; r = pp - q * n_ref; }
        }
        rec[h] = unshared_record(a, size_qry[q], size_ref[r]);
    }
    if (p + 1 < total) {
        const uint4* src = reinterpret_cast<const uint4*>(rec);
        uint4* dst = reinterpret_cast<uint4*>(out + p);                  // 48 * (p/2) bytes from a 16-byte aligned base
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
    } else
        out[p] = rec[0];
}

__global__ void __launch_bounds__(256) dist_literal_kernel(fpm_panel ref, fpm_panel qry, DistArgs a, fpm_pair* out, unsigned long long* steps)
{
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t total = ref.n * qry.n;
    unsigned long long my_steps = 0;
    if (p < total) {
        uint64_t q = p / ref.n, r = p % ref.n;
        const uint64_t* A = ref.hashes + r * ref.stride;
        const uint64_t* B = qry.hashes + q * qry.stride;
        uint64_t na = ref.sizes[r], nb = qry.sizes[q];
        uint64_t i = 0, j = 0, common = 0, denom = 0;
        while (denom < a.s && i < na && j < nb) {                       // CommandDistance.cpp:376-387
            uint64_t x = A[i], y = B[j];
            if (x < y) i++;
            else if (y < x) j++;
            else { i++; j++; common++; }
            denom++;
        }
        my_steps = denom;
        if (denom < a.s) {                                              // :389-400
            if (i < na) denom += na - i;
            if (j < nb) denom += nb - j;
            if (denom > a.s) denom = a.s;
        }
        finish_pair(a, common, denom, ref.lengths[r], qry.lengths[q], out + p);
    }
    if (steps) {
        for (int o = 16; o; o >>= 1) my_steps += __shfl_down_sync(0xffffffffu, my_steps, o);
        if ((threadIdx.x & 31) == 0 && my_steps) atomicAdd(steps, my_steps);
    }
}

// The fork's positional fingerprint comparison (CommandTriangle.cpp:265-302, `mash triangle -fp`):
// matches = #{i < min(|A|,|B|) : A[i] == B[i]}, distance = 1 - matches/min, p = chi-square upper tail
// with one degree of freedom at `matches` (gsl_cdf_chisq_Q(matches, 1) = erfc(sqrt(matches/2))).
__global__ void __launch_bounds__(256) fp_positional_kernel(fpm_panel ref, fpm_panel qry, double max_distance, double max_pvalue, fpm_pair* out)
{
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= ref.n * qry.n) return;
    uint64_t q = p / ref.n, r = p % ref.n;
    const uint64_t* A = ref.hashes + r * ref.stride;
    const uint64_t* B = qry.hashes + q * qry.stride;
    uint32_t na = ref.sizes[r], nb = qry.sizes[q], m = na < nb ? na : nb, matches = 0;
    for (uint32_t i = 0; i < m; i++) matches += A[i] == B[i];
    fpm_pair o;
    o.numer = matches;
    o.denom = m;
    o.distance = 1.0 - (double(matches) / double(m));
    o.pvalue = erfc(sqrt(double(matches) * 0.5));
    if (o.distance <= max_distance && o.pvalue <= max_pvalue) o.denom |= FPM_PAIR_PASS;
    out[p] = o;
}

// Row-major [n][stride] -> column tiles [ceil(n/16)][rows][16], +inf beyond each sketch's size.
// flags[0] |= 1 if a real hash collides with the sentinel's high word or a row is not strictly ascending (the
// host then falls back to the literal kernel).
__global__ void __launch_bounds__(256) dist_pack_kernel(fpm_panel pn, uint64_t rows, uint64_t* packed, uint32_t* flags)
{
    uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;     // over n16*16 x rows, column fastest
    uint64_t n16 = (pn.n + 15) / 16;
    if (idx >= n16 * 16 * rows) return;
    uint64_t c = idx & 15, row = (idx >> 4) % rows, tile = (idx >> 4) / rows;
    uint64_t sk = tile * 16 + c;
    uint64_t v = DT_INF;
    if (sk < pn.n && row < pn.sizes[sk]) {
        v = pn.hashes[sk * pn.stride + row];
        bool bad = (v >> 32) == 0xffffffffULL;   // high word all ones is reserved for the +inf sentinel
        if (row > 0 && pn.hashes[sk * pn.stride + row - 1] >= v) bad = true;
        if (bad) atomicOr(flags, 1u);
    }
    packed[idx] = v;
}

__global__ void __launch_bounds__(DT_THREADS, 2)
dist_tile_kernel(const uint64_t* __restrict__ pref, const uint64_t* __restrict__ pqry, uint64_t rows_ref, uint64_t rows_qry,
                 uint64_t n_ref, uint64_t n_qry, const uint64_t* __restrict__ len_ref, const uint64_t* __restrict__ len_qry,
                 DistArgs a, fpm_pair* __restrict__ out, unsigned long long* steps, uint32_t q_tile0)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* colQ = reinterpret_cast<uint64_t*>(smem_raw);                 // [(DT_ROWS+1)][16]
    uint64_t* colR = colQ + DT_COLROWS * 16;                             // 2 planes of [(DT_ROWS+1)][16]
    __shared__ uint32_t s_cursor[DT_COLS];
    __shared__ unsigned long long s_V;
    __shared__ uint32_t s_need;

    const int t = threadIdx.x;
    const int w = t >> 5, l = t & 31, h = l >> 4, i16 = l & 15;
    const int qc = i16;                              // my query column
    const int rc = ((i16 + w) & 15);                 // my reference column within plane h
    const uint64_t q_tile = blockIdx.y + q_tile0, r_tile2 = blockIdx.x;               // r_tile2 indexes pairs of 16-col tiles
    const uint64_t* gQ = pqry + q_tile * rows_qry * 16;
    const uint64_t* gR[2] = {pref + (2 * r_tile2) * rows_ref * 16, pref + (2 * r_tile2 + 1) * rows_ref * 16};
    const uint64_t n_rtiles16 = (n_ref + 15) / 16;
    const bool plane1_exists = 2 * r_tile2 + 1 < n_rtiles16;

    if (t < DT_COLS) s_cursor[t] = 0;
    uint32_t common = 0, denom = 0;
    bool done = false;
    // rows staged per phase: no pair consumes more elements of a list than it has union steps left, so later
    // phases (a few dozen steps for the stragglers) stage only what can still be needed
    int R = (int)(a.s < (uint32_t)DT_ROWS ? (a.s < 32u ? 32u : a.s) : (uint32_t)DT_ROWS);
    __syncthreads();

    for (;;) {
        // ---- V = smallest element that does not fit this phase ------------------------
        __syncthreads();                   // everyone has read s_need / written its cursor for this phase
        if (t == 0) { s_V = DT_INF; s_need = 0; }
        __syncthreads();
        if (t < DT_COLS) {
            uint64_t row = (uint64_t)s_cursor[t] + R;
            uint64_t v = DT_INF;
            if (t < DT_Q) { if (row < rows_qry) v = gQ[row * 16 + t]; }
            else {
                int pl = (t - DT_Q) >> 4, c = (t - DT_Q) & 15;
                if ((pl == 0 || plane1_exists) && row < rows_ref) v = gR[pl][row * 16 + c];
            }
            if (v != DT_INF) atomicMin(&s_V, (unsigned long long)v);
        }
        __syncthreads();
        const uint64_t V = s_V;
        // ---- stage rows [cursor, cursor+R) of every column, masking >= V ----------------
        for (int idx = t; idx < R * 16; idx += DT_THREADS) {
            int c = idx & 15, r = idx >> 4;
            uint64_t row = (uint64_t)s_cursor[c] + r;
            uint64_t v = row < rows_qry ? gQ[row * 16 + c] : DT_INF;
            colQ[r * 16 + c] = v < V ? v : DT_INF;
        }
#pragma unroll
        for (int pl = 0; pl < 2; pl++) {
            uint64_t* dst = colR + pl * DT_COLROWS * 16;
            bool exists = pl == 0 || plane1_exists;
            for (int idx = t; idx < R * 16; idx += DT_THREADS) {
                int c = idx & 15, r = idx >> 4;
                uint64_t row = (uint64_t)s_cursor[DT_Q + 16 * pl + c] + r;
                uint64_t v = (exists && row < rows_ref) ? gR[pl][row * 16 + c] : DT_INF;
                dst[r * 16 + c] = v < V ? v : DT_INF;
            }
        }
        for (int idx = t; idx < DT_PAD * 16; idx += DT_THREADS) {
            colQ[R * 16 + idx] = DT_INF; colR[R * 16 + idx] = DT_INF; colR[DT_COLROWS * 16 + R * 16 + idx] = DT_INF;
        }
        __syncthreads();

        // ---- merge up to V ------------------------------------------------------------
        if (!done) {
            // The reference loop (CommandDistance.cpp:376-387) with two simplifications: an exhausted
            // (or phase-masked) list reads +inf, and the "complete the union" tail (:389-400) is the
            // same loop running on the surviving list one element per step.  Values are handled as
            // 32-bit halves and shared-space byte addresses so one step is ~13 instructions.  The
            // common count is recovered from the pointers: advances(a) + advances(b) = steps + matches.
            const uint32_t pa0 = (uint32_t)__cvta_generic_to_shared(colQ + qc);
            const uint32_t pb0 = (uint32_t)__cvta_generic_to_shared(colR + h * DT_COLROWS * 16 + rc);
            uint32_t pa = pa0, pb = pb0, alo, ahi, blo, bhi;
            lds64(pa, alo, ahi);
            lds64(pb, blo, bhi);
            const uint32_t budget = a.s - denom;
            uint32_t rem = budget;
            // (pack flagged any real hash whose high word is all ones, so hi == ~0 <=> +inf)
            for (;;) {
                // fast block: when at least 8 steps of budget remain and neither list can run out within them
                // (row +8 of both columns is still a real value), run 8 steps without any exit test
                if (rem >= 8) {
                    uint32_t a8, b8;
                    lds32(pa + 8 * 128 + 4, a8);
                    lds32(pb + 8 * 128 + 4, b8);
                    if (a8 != 0xffffffffu && b8 != 0xffffffffu) {
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            const uint64_t x = ((uint64_t)ahi << 32) | alo, y = ((uint64_t)bhi << 32) | blo;
                            const bool lt = x < y, gt = y < x;
                            if (!gt) { pa += 128; lds64(pa, alo, ahi); }
                            if (!lt) { pb += 128; lds64(pb, blo, bhi); }
                        }
                        rem -= 8;
                        continue;
                    }
                }
                if (rem == 0 || (ahi & bhi) == 0xffffffffu) break;
                const uint64_t x = ((uint64_t)ahi << 32) | alo, y = ((uint64_t)bhi << 32) | blo;
                const bool lt = x < y, gt = y < x;
                if (!gt) { pa += 128; lds64(pa, alo, ahi); }
                if (!lt) { pb += 128; lds64(pb, blo, bhi); }
                rem--;
            }
            const uint32_t steps = budget - rem;
            denom += steps;
            common += ((pa - pa0) >> 7) + ((pb - pb0) >> 7) - steps;
            if (denom >= a.s || V == DT_INF) done = true;
            else atomicMax(&s_need, a.s - denom);
        }
        // ---- anyone left?  then advance every column past its elements < V ------------
        __syncthreads();
        const uint32_t need_all = s_need;    // max over the unfinished pairs of the union steps still missing (see dist_tile32_kernel)
        if (need_all == 0) break;
        if (t < DT_COLS) {
            const uint64_t* col = t < DT_Q ? colQ + t : colR + ((t - DT_Q) >> 4) * DT_COLROWS * 16 + ((t - DT_Q) & 15);
            uint32_t lo = 0, hi = R;         // first row holding +inf (masked or exhausted)
            while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (col[mid * 16] != DT_INF) lo = mid + 1; else hi = mid; }
            s_cursor[t] += lo;
        }
        R = (int)min((uint32_t)DT_ROWS, max(32u, need_all));
        // (the barrier at the top of the next phase orders these writes before their readers)
    }

    // ---- results: stage in shared memory, then coalesced row writes ---------------------
    __syncthreads();
    fpm_pair* res = reinterpret_cast<fpm_pair*>(smem_raw);                  // [16][32]
    const uint64_t qg = q_tile * 16 + qc, rg = r_tile2 * 32 + 16 * h + rc;
    unsigned long long my_steps = denom;
    if (qg < n_qry && rg < n_ref) {
        finish_pair(a, common, denom, len_ref[rg], len_qry[qg], &res[qc * 32 + 16 * h + rc]);
    }
    __syncthreads();
    {
        // 16 rows of 32 pairs = 768 bytes each, written as 8-byte words
        const uint64_t* src = reinterpret_cast<const uint64_t*>(res);
        for (int idx = t; idx < 16 * 32 * 3; idx += DT_THREADS) {
            int row = idx / 96, wd = idx % 96, pr = wd / 3;
            uint64_t qg2 = q_tile * 16 + row, rg2 = r_tile2 * 32 + pr;
            if (qg2 < n_qry && rg2 < n_ref)
                reinterpret_cast<uint64_t*>(out + qg2 * n_ref + rg2)[wd % 3] = src[idx];
        }
    }
    if (steps) {
        for (int o = 16; o; o >>= 1) my_steps += __shfl_down_sync(0xffffffffu, my_steps, o);
        if (l == 0 && my_steps) atomicAdd(steps, my_steps);
    }
}

// ---------------------------------------------------------------------------------------------------------
// dist_tile32_kernel: the tile algorithm on 32-bit dense ranks (dist_rank.cu).  What changes against the 64-bit
// kernel above, each from an ncu capture (profiles/r01_dist_tile_v3.txt: shared-memory pipe 80 % busy, every step
// two LDS.64 of two wavefronts each; profiles/r01_dist_tile32_v1.txt: latency bound at 8 warps per scheduler):
//   * elements are 4 bytes: a step's loads are LDS.32, one wavefront each (32 lanes, 32 different columns);
//   * the order test is one ISETP per direction instead of a two-instruction 64-bit compare;
//   * the tile is 32 queries x 32 references and every thread merges TWO pairs (same query, references 16
//     columns apart) in one interleaved instruction stream: twice the loads in flight per warp;
//   * the first +inf row of every column is found once per phase, so a thread knows how many steps it can run
//     without an end test; a list that is through lets the other one jump to its end in O(1);
//   * 432 rows per phase at two CTAs per SM (was 288); all phases are staged with 16-byte vector loads (phase 0
//     row-aligned, later phases as the union of the columns' row ranges per 16-column plane).
// ---------------------------------------------------------------------------------------------------------
constexpr uint32_t D4_INF = 0xffffffffu;
#ifndef FPM_D4_ROWS
#define FPM_D4_ROWS 432
#endif
#ifndef FPM_D4_UNROLL
#define FPM_D4_UNROLL 2      // measured: 1 -> 2.63, 2 -> 2.66, 4 -> 2.62, 8 -> 2.55, 16 -> 2.40 G pairs/s (the compiler unrolls the counted loop itself)
#endif
constexpr int D4_ROWS = FPM_D4_ROWS;      // rows resident per phase
constexpr int D4_UNROLL = FPM_D4_UNROLL;  // unchecked steps per fast block
constexpr int D4_PAD = 1;             // one +inf row after them (a pointer rests at most on row R)
constexpr int D4_COLROWS = D4_ROWS + D4_PAD;   // 433 rows x (32 + 32) columns x 4 B = 108.25 KB: two CTAs per SM
constexpr int D4_COLS = 64;           // 32 query + 32 reference columns

struct Merge32 {
    uint32_t pa, pb, av, bv, ea, eb, rem;
    __device__ __forceinline__ uint32_t safe() const { return min(min((ea - pa) >> 7, (eb - pb) >> 7), rem); }
    __device__ __forceinline__ bool finished() const { return rem == 0 || (av & bv) == D4_INF; }
    // no unchecked step is possible (safe() == 0) although the pair is not finished: one list rests on its first
    // +inf row, so the other one advances alone -- "complete the union" (CommandDistance.cpp:389-400) -- in O(1)
    __device__ __forceinline__ void one_sided()
    {
        if (pa == ea) { const uint32_t k = min((eb - pb) >> 7, rem); pb += k * 128; rem -= k; lds32(pb, bv); }
        else { const uint32_t k = min((ea - pa) >> 7, rem); pa += k * 128; rem -= k; lds32(pa, av); }
    }
};

// One pair to the end of the phase.  Deliberately NOT inlined: with this loop nest inlined twice behind the
// two-pair loop of the kernel, nvcc 12.9 -O3 produced code for sm_100a that ran pointers past their column
// ends (hangs / illegal addresses from n = 288 sketches on; -G and -Xptxas -O0 builds of the same source
// were correct).  tests/test_gpu_dist.py::test_dist_rank32_sizes_sweep pins the shapes that exposed it.
__device__ __noinline__ Merge32 merge32_alone(Merge32 m)
{
    for (;;) {
        uint32_t n = m.safe();
        if (n >= (uint32_t)D4_UNROLL) {
            uint32_t blocks = n / D4_UNROLL;
            m.rem -= blocks * D4_UNROLL;
            do {
#pragma unroll
                for (int u = 0; u < D4_UNROLL; u++) merge_step32(m.pa, m.pb, m.av, m.bv);
            } while (--blocks);
            continue;
        }
        if (n) {                     // fewer than a block: still no end test needed for these n steps
            m.rem -= n;
#pragma unroll 1
            do merge_step32(m.pa, m.pb, m.av, m.bv); while (--n);
            continue;
        }
        if (m.finished()) break;
        m.one_sided();
    }
    return m;
}

__global__ void __launch_bounds__(DT_THREADS, 2)
dist_tile32_kernel(const uint32_t* __restrict__ pref, const uint32_t* __restrict__ pqry, uint64_t rows_ref, uint64_t rows_qry,
                   uint64_t n_ref, uint64_t n_qry, const uint64_t* __restrict__ len_ref, const uint64_t* __restrict__ len_qry,
                   DistArgs a, fpm_pair* __restrict__ out, unsigned long long* steps, uint32_t q_tile0,
                   const uint32_t* __restrict__ marks, uint32_t mark_words, const uint32_t* __restrict__ size_ref, const uint32_t* __restrict__ size_qry,
                   const uint32_t* __restrict__ perm_q, const uint32_t* __restrict__ perm_r, HitSink hs, int prefilled,
                   const uint2* __restrict__ tile_list)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* colQ = reinterpret_cast<uint32_t*>(smem_raw);               // [D4_COLROWS][32] query columns
    uint32_t* colR = colQ + D4_COLROWS * 32;                              // [D4_COLROWS][32] reference columns
    __shared__ uint32_t s_cursor[D4_COLS], s_end[D4_COLS], s_exh[D4_COLS], s_lo[4], s_span[4];
    __shared__ uint32_t s_V;
    __shared__ uint32_t s_need;

    const int t = threadIdx.x;
    const int w = t >> 5, l = t & 31;
    const int rc0 = (l + w) & 31, rc1 = (l + w + 16) & 31;     // my two reference columns; my query column is l
    // both index pairs of 16-column tiles; with a tile list (dist_rank.cu: only the tiles that hold work) the grid is 1-D over it
    const uint64_t q_tile2 = tile_list ? tile_list[blockIdx.x].x : blockIdx.y + q_tile0, r_tile2 = tile_list ? tile_list[blockIdx.x].y : blockIdx.x;
    const uint32_t* gQ0 = pqry + (2 * q_tile2) * rows_qry * 16;           // plane 1 follows at + rows * 16
    const uint32_t* gR0 = pref + (2 * r_tile2) * rows_ref * 16;
    const bool q1_exists = 2 * q_tile2 + 1 < (n_qry + 15) / 16, r1_exists = 2 * r_tile2 + 1 < (n_ref + 15) / 16;

    // column t of the 64 (0..31 query, 32..63 reference): its global base, row count and whether it exists
    const uint32_t* my_col = nullptr;
    uint64_t my_rows = 0;
    if (t < D4_COLS) {
        const int c = t & 31, pl = c >> 4;
        const bool isq = t < 32;
        if (pl == 0 || (isq ? q1_exists : r1_exists)) {
            my_rows = isq ? rows_qry : rows_ref;
            my_col = (isq ? gQ0 : gR0) + (uint64_t)pl * my_rows * 16 + (c & 15);
        }
        s_cursor[t] = 0;
    }
    // ---- which pairs need a merge at all?  (marks: bit r of query q's row = the two lists share a hash; a pair
    // without any shared hash has common = 0 and denom = min(s, |A| + |B|), no merge needed) -----------------------
    __shared__ uint32_t s_mask[32];
    __shared__ uint16_t s_list[1024];
    __shared__ uint32_t s_npairs;
    __shared__ uint32_t s_empty[2];          // bit c: query / reference c of the tile is an empty sketch
    __shared__ uint32_t s_orig[64];          // original sketch index of my 32 queries / 32 references (panels may be grouped, dist_rank.cu)
    if (t >= 64 && t < 128) {
        const int c = t - 64;
        const bool isq = c < 32;
        const uint64_t g = (isq ? q_tile2 : r_tile2) * 32 + (c & 31), n = isq ? n_qry : n_ref;
        const uint32_t* perm = isq ? perm_q : perm_r;
        const uint32_t o = g < n ? (perm ? perm[g] : (uint32_t)g) : 0xffffffffu;
        s_orig[c] = o;
        // (threads 64..95 are one warp = the queries, 96..127 the references) which of them are empty sketches
        const uint32_t em = __ballot_sync(0xffffffffu, o != 0xffffffffu && (isq ? size_qry : size_ref)[o] == 0);
        if ((c & 31) == 0) s_empty[isq ? 0 : 1] = em;
    }
    if (t < 32) {
        const uint64_t qg = q_tile2 * 32 + t;
        uint32_t mk = 0;
        if (qg < n_qry) {
            mk = marks ? marks[qg * mark_words + r_tile2] : 0xffffffffu;
            const uint64_t left = n_ref - r_tile2 * 32;                  // references in this tile
            if (left < 32) mk &= (1u << left) - 1u;
        }
        s_mask[t] = mk;
        const uint32_t c = __popc(mk);
        uint32_t inc = c;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
            if (t >= o) inc += up;
        }
        uint32_t pos = inc - c;
        for (uint32_t m = mk; m; m &= m - 1) s_list[pos++] = (uint16_t)((t << 5) | (__ffs(m) - 1));
        if (t == 31) s_npairs = inc;
    }
    __syncthreads();
    const uint32_t npairs = s_npairs;
    // many pairs: the fixed conflict-free assignment (query l, references (l+w) and (l+w+16) mod 32), unmarked pairs idle;
    // few pairs: the list is dealt out two per thread, so that whole warps have nothing to do
    const bool dense = npairs > 256;
    int qc0 = l, qc1 = l, rcm0 = rc0, rcm1 = rc1;
    bool act0, act1;
    if (dense) {
        act0 = (s_mask[l] >> rc0) & 1u;
        act1 = (s_mask[l] >> rc1) & 1u;
    } else {
        act0 = 2u * t < npairs;
        act1 = 2u * t + 1 < npairs;
        if (act0) { const uint32_t e = s_list[2 * t]; qc0 = e >> 5; rcm0 = e & 31; }
        if (act1) { const uint32_t e = s_list[2 * t + 1]; qc1 = e >> 5; rcm1 = e & 31; }
    }
    uint32_t common0 = 0, common1 = 0, denom0 = 0, denom1 = 0;
    bool done0 = !act0, done1 = !act1;
    int R = (int)(a.s < (uint32_t)D4_ROWS ? (a.s < 32u ? 32u : a.s) : (uint32_t)D4_ROWS);
    __syncthreads();

    for (int phase = 0; npairs != 0; phase++) {
        // ---- V = smallest element that does not fit this phase ------------------------
        __syncthreads();
        if (t == 0) { s_V = D4_INF; s_need = 0; }
        __syncthreads();
        if (t < D4_COLS) {
            const uint64_t row = (uint64_t)s_cursor[t] + R;
            if (row < my_rows) { const uint32_t v = my_col[row * 16]; if (v != D4_INF) atomicMin(&s_V, v); }
        }
        __syncthreads();
        const uint32_t V = s_V;
        // ---- stage rows [cursor, cursor+R) of every column, masking >= V ----------------
        if (phase == 0) {
            // all cursors are 0: rows are contiguous 64-byte lines, moved as 16-byte vectors
            for (int idx = t; idx < R * 16; idx += DT_THREADS) {
                const int r = idx >> 4, c4 = (idx & 7) * 4, pl = c4 >> 4;
                const bool isq = (idx & 8) == 0;
                const uint64_t rows = isq ? rows_qry : rows_ref;
                uint4 v = make_uint4(D4_INF, D4_INF, D4_INF, D4_INF);
                if ((pl == 0 || (isq ? q1_exists : r1_exists)) && (uint64_t)r < rows)
                    v = *reinterpret_cast<const uint4*>((isq ? gQ0 : gR0) + ((uint64_t)pl * rows + r) * 16 + (c4 & 15));
                v.x = v.x < V ? v.x : D4_INF; v.y = v.y < V ? v.y : D4_INF; v.z = v.z < V ? v.z : D4_INF; v.w = v.w < V ? v.w : D4_INF;
                *reinterpret_cast<uint4*>((isq ? colQ : colR) + r * 32 + c4) = v;
            }
        } else {
            // later phases: every column has its own cursor.  Per 16-column plane, load the union of the row ranges
            // [min cursor, max cursor + R) as whole 64-byte lines (16-byte vectors, coalesced) and drop every element
            // into its column's slot; cells fed from beyond the list's end get +inf.  (Per-element loads at the
            // columns' own rows touched a 32-byte sector per 4-byte element.)
            if (t < 4) {
                uint32_t lo = 0xffffffffu, hi = 0;
                for (int c = 0; c < 16; c++) { const uint32_t cu = s_cursor[16 * t + c]; lo = min(lo, cu); hi = max(hi, cu); }
                s_lo[t] = lo;
                s_span[t] = hi - lo + (uint32_t)R;
            }
            __syncthreads();
#pragma unroll 1
            for (int pl4 = 0; pl4 < 4; pl4++) {                       // planes: Q0, Q1, R0, R1
                const bool isq = pl4 < 2;
                const int pl = pl4 & 1;
                const uint64_t rows = isq ? rows_qry : rows_ref;
                const bool exists = pl == 0 || (isq ? q1_exists : r1_exists);
                const uint32_t* src = (isq ? gQ0 : gR0) + (uint64_t)pl * rows * 16;
                uint32_t* dst = (isq ? colQ : colR) + 16 * pl;
                const uint32_t lo = s_lo[pl4], span = s_span[pl4];
                for (uint32_t idx = t; idx < span * 4; idx += DT_THREADS) {
                    const uint32_t row = lo + (idx >> 2), c4 = (idx & 3) * 4;
                    uint4 v = make_uint4(D4_INF, D4_INF, D4_INF, D4_INF);
                    if (exists && (uint64_t)row < rows) v = *reinterpret_cast<const uint4*>(src + (uint64_t)row * 16 + c4);
                    const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const uint32_t r = row - s_cursor[16 * pl4 + c4 + e];       // wraps for rows before the column's cursor
                        if (r < (uint32_t)R) dst[r * 32 + c4 + e] = vv[e] < V ? vv[e] : D4_INF;
                    }
                }
            }
        }
        for (int idx = t; idx < D4_PAD * 64; idx += DT_THREADS) (idx < D4_PAD * 32 ? colQ + R * 32 : colR + R * 32 - D4_PAD * 32)[idx] = D4_INF;
        __syncthreads();
        if (t < D4_COLS) {
            const uint32_t* col = t < 32 ? colQ + t : colR + (t - 32);
            uint32_t lo = 0, hi = R;         // first row holding +inf (masked or exhausted)
            while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (col[mid * 32] != D4_INF) lo = mid + 1; else hi = mid; }
            s_end[t] = lo;
            const uint64_t nxt = (uint64_t)s_cursor[t] + lo;          // nothing at or beyond V either: the list is through for good
            s_exh[t] = !(nxt < my_rows && my_col[nxt * 16] != D4_INF);
        }
        __syncthreads();

        // ---- merge up to V ------------------------------------------------------------
        // the reference loop (CommandDistance.cpp:376-400): an exhausted or phase-masked list reads +inf; matches
        // are recovered from the pointer advances: advances(a) + advances(b) = steps + matches
        {
            const uint32_t pa00 = (uint32_t)__cvta_generic_to_shared(colQ + qc0), pa01 = (uint32_t)__cvta_generic_to_shared(colQ + qc1);
            const uint32_t pb00 = (uint32_t)__cvta_generic_to_shared(colR + rcm0), pb01 = (uint32_t)__cvta_generic_to_shared(colR + rcm1);
            Merge32 m0, m1;
            m0.pa = pa00; m0.pb = pb00; m0.ea = pa00 + s_end[qc0] * 128; m0.eb = pb00 + s_end[32 + rcm0] * 128; m0.rem = done0 ? 0u : a.s - denom0;
            m1.pa = pa01; m1.pb = pb01; m1.ea = pa01 + s_end[qc1] * 128; m1.eb = pb01 + s_end[32 + rcm1] * 128; m1.rem = done1 ? 0u : a.s - denom1;
            const uint32_t budget0 = m0.rem, budget1 = m1.rem;
            lds32(pa00, m0.av);
            lds32(pa01, m1.av);
            lds32(pb00, m0.bv);
            lds32(pb01, m1.bv);
            // both pairs together while both have unchecked steps left; then each one alone
            const uint32_t k128 = a.k128, k1 = a.k1;
            for (;;) {
                uint32_t blocks = min(m0.safe(), m1.safe()) / D4_UNROLL;
                if (!blocks) break;
                m0.rem -= blocks * D4_UNROLL; m1.rem -= blocks * D4_UNROLL;
                do {
#pragma unroll
                    for (int u = 0; u < D4_UNROLL; u++) { merge_step32(m0.pa, m0.pb, m0.av, m0.bv, k128, k1); merge_step32(m1.pa, m1.pb, m1.av, m1.bv, k128, k1); }
                } while (--blocks);
            }
            m0 = merge32_alone(m0);
            m1 = merge32_alone(m1);
            // a pair is finished when its union reached s, or when both lists are through for good (nothing at or
            // beyond V either); everything else goes on in the next phase with the rows it can still need
            const uint32_t exh_a0 = s_exh[qc0], exh_a1 = s_exh[qc1], exh_b0 = s_exh[32 + rcm0], exh_b1 = s_exh[32 + rcm1];
            {
                const uint32_t n = budget0 - m0.rem;
                denom0 += n;
                common0 += ((m0.pa - pa00) >> 7) + ((m0.pb - pb00) >> 7) - n;
                const bool through = (exh_a0 & exh_b0) != 0 & (m0.pa == m0.ea) & (m0.pb == m0.eb);
                done0 = done0 | (denom0 >= a.s) | through;
            }
            {
                const uint32_t n = budget1 - m1.rem;
                denom1 += n;
                common1 += ((m1.pa - pa01) >> 7) + ((m1.pb - pb01) >> 7) - n;
                const bool through = (exh_a1 & exh_b1) != 0 & (m1.pa == m1.ea) & (m1.pb == m1.eb);
                done1 = done1 | (denom1 >= a.s) | through;
            }
            const uint32_t need = max(done0 ? 0u : a.s - denom0, done1 ? 0u : a.s - denom1);
            if (need) atomicMax(&s_need, need);
        }
        // ---- anyone left?  then advance every column past its elements < V ------------
        // (a plain barrier and the shared maximum: the BAR.RED form of this test, __syncthreads_or, faulted with
        // "illegal instruction" on sm_100a in this kernel)
        __syncthreads();
        const uint32_t need_all = s_need;                      // max over the CTA's unfinished pairs of the union steps still missing
        if (need_all == 0) break;
        if (t < D4_COLS) s_cursor[t] += s_end[t];
        R = (int)min((uint32_t)D4_ROWS, max(32u, need_all));
    }

    // ---- results -------------------------------------------------------------------------
    __syncthreads();
    unsigned long long my_steps = 0;       // union steps actually merged
    if (hs.buf) {
        // hits mode: nothing is written for a pair that fails the -d / -v filters
        fpm_pair o = {0, 0, 0., 0.};
        // pairs without a shared hash have distance 1 -- except two EMPTY sketches (denom 0 = common: distance 0, dist_math.h),
        // so even when the filters exclude distance 1 a tile holding empty sketches on both sides looks at those pairs
        const uint32_t eq = s_empty[0], er = s_empty[1];
        if (!hs.skip_unmarked || (eq && er)) {
            const uint32_t qo = s_orig[l], mk = s_mask[l];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int rc = h ? rc1 : rc0;
                const uint32_t ro = s_orig[32 + rc];
                bool pass = false;
                if (qo != 0xffffffffu && ro != 0xffffffffu && !((mk >> rc) & 1u) && (!hs.skip_unmarked || ((eq >> l) & (er >> rc) & 1u))) {
                    const uint64_t un = (uint64_t)size_qry[qo] + size_ref[ro];
                    finish_pair(a, 0, un < a.s ? un : a.s, len_ref[ro], len_qry[qo], &o);
                    pass = (o.denom & FPM_PAIR_PASS) != 0;
                }
                append_hit(hs, pass, qo, ro, o);
            }
        }
        bool pass = false;
        if (act0) { finish_pair(a, common0, denom0, len_ref[s_orig[32 + rcm0]], len_qry[s_orig[qc0]], &o); my_steps += denom0; pass = (o.denom & FPM_PAIR_PASS) != 0; }
        append_hit(hs, pass, s_orig[qc0], s_orig[32 + rcm0], o);
        pass = false;
        if (act1) { finish_pair(a, common1, denom1, len_ref[s_orig[32 + rcm1]], len_qry[s_orig[qc1]], &o); my_steps += denom1; pass = (o.denom & FPM_PAIR_PASS) != 0; }
        append_hit(hs, pass, s_orig[qc1], s_orig[32 + rcm1], o);
    } else if (prefilled) {
        // matrix mode, records of pairs without a shared hash already in place (dist_fill_unshared_kernel): only the merged pairs
        if (act0) { finish_pair(a, common0, denom0, len_ref[s_orig[32 + rcm0]], len_qry[s_orig[qc0]], out + (uint64_t)s_orig[qc0] * n_ref + s_orig[32 + rcm0]); my_steps += denom0; }
        if (act1) { finish_pair(a, common1, denom1, len_ref[s_orig[32 + rcm1]], len_qry[s_orig[qc1]], out + (uint64_t)s_orig[qc1] * n_ref + s_orig[32 + rcm1]); my_steps += denom1; }
    } else {
        // matrix mode: stage in shared memory, then coalesced row writes
        fpm_pair* res = reinterpret_cast<fpm_pair*>(smem_raw);                  // [32][32]
        {
            // pairs that were not merged (by the fixed assignment: query l, references rc0 and rc1): no shared hash
            const uint32_t qo = s_orig[l];
            if (qo != 0xffffffffu) {
                const uint64_t lq = len_qry[qo];
                const uint32_t sq = size_qry[qo], mk = s_mask[l];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int rc = h ? rc1 : rc0;
                    const uint32_t ro = s_orig[32 + rc];
                    if (ro != 0xffffffffu && !((mk >> rc) & 1u)) {
                        const uint64_t un = (uint64_t)sq + size_ref[ro];
                        finish_pair(a, 0, un < a.s ? un : a.s, len_ref[ro], lq, &res[l * 32 + rc]);
                    }
                }
            }
            // the merged ones, by whoever merged them (marked pairs are always real pairs)
            if (act0) { finish_pair(a, common0, denom0, len_ref[s_orig[32 + rcm0]], len_qry[s_orig[qc0]], &res[qc0 * 32 + rcm0]); my_steps += denom0; }
            if (act1) { finish_pair(a, common1, denom1, len_ref[s_orig[32 + rcm1]], len_qry[s_orig[qc1]], &res[qc1 * 32 + rcm1]); my_steps += denom1; }
        }
        __syncthreads();
        {
            // 32 rows of 32 pairs, written as 8-byte words to out[original query][original reference]: 768-byte rows when the
            // panels are in their own order, 24-byte records when they were grouped
            const uint64_t* src = reinterpret_cast<const uint64_t*>(res);
            for (int idx = t; idx < 32 * 32 * 3; idx += DT_THREADS) {
                int row = idx / 96, wd = idx % 96, pr = wd / 3;
                const uint32_t qo = s_orig[row], ro = s_orig[32 + pr];
                if (qo != 0xffffffffu && ro != 0xffffffffu)
                    reinterpret_cast<uint64_t*>(out + (uint64_t)qo * n_ref + ro)[wd % 3] = src[idx];
            }
        }
    }
    if (steps) {
        for (int o = 16; o; o >>= 1) my_steps += __shfl_down_sync(0xffffffffu, my_steps, o);
        if (l == 0 && my_steps) atomicAdd(steps, my_steps);
    }
}

// h_out (nullable): host destination.  When given, the fast path runs in query-row chunks and copies chunk c
// back on a second stream while chunk c+1 is being compared, then waits for all copies.
int run_dist(fpm_ctx* ctx, const fpm_dist_params* p, const fpm_panel* d_ref, const fpm_panel* d_qry, fpm_pair* d_out,
             uint64_t* d_steps, uint32_t max_size_ref, uint32_t max_size_qry, fpm_pair* h_out, const HitSink* hits, uint64_t h_ld)
{
    // hits mode (fpm_dist_hits): d_out and h_out are null.  The 32-bit rank kernel appends passing pairs itself; the other
    // paths need a result matrix, which then lives in ctx->d_out and is scanned by dist_collect_kernel.
    const HitSink no_hits{};
    auto matrix_for_hits = [&]() -> int {
        const uint64_t bytes = d_ref->n * d_qry->n * sizeof(fpm_pair);
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = 0;
        if (bytes > ctx->d_out.cap && bytes > free_b) { set_error("fpm_dist_hits: these panels need a %llu-byte result matrix on this path; pass fewer queries per call", (unsigned long long)bytes); return FPM_ERR_NOMEM; }
        int rc = ctx->d_out.ensure(bytes);
        if (rc) return rc;
        d_out = ctx->d_out.as<fpm_pair>();
        return FPM_OK;
    };
    auto collect = [&]() -> int {
        const uint64_t total = d_ref->n * d_qry->n;
        dist_collect_kernel<<<(uint32_t)std::min<uint64_t>((total + 255) / 256, (uint64_t)ctx->sm_count * 16), 256, 0, ctx->stream>>>(d_out, total, d_ref->n, *hits);
        ctx->launches++;
        FPM_CUDA(cudaGetLastError());
        return FPM_OK;
    };
    DistArgs a;
    a.s = p->sketch_size; a.kmer_size = p->kmer_size; a.kmer_space = p->kmer_space;
    a.max_distance = p->max_distance; a.max_pvalue = p->max_pvalue;
    cudaStream_t st = ctx->stream;
    uint64_t total = d_ref->n * d_qry->n;
    if (total == 0) return FPM_OK;
    bool fast = p->sorted_unique != 0;
    if (fast) {
        int rc;
        const uint64_t rows_r = (uint64_t)max_size_ref + 1, rows_q = (uint64_t)max_size_qry + 1;
        const uint64_t nr16 = (d_ref->n + 15) / 16, nq16 = (d_qry->n + 15) / 16;
        // preferred: 32-bit dense ranks (dist_rank.cu); panels too large for 32-bit indices keep the 64-bit kernel
        uint32_t *p32r = nullptr, *p32q = nullptr;
        int mode = DIST_RANK_TOO_BIG;
        uint32_t* marks = nullptr;
        if (!ctx->force_dist64 && (rc = dist_rank_panels(ctx, d_ref, d_qry, max_size_ref, max_size_qry, rows_r, rows_q, a.s, &p32r, &p32q, &mode, &marks))) return rc;
        if (mode == DIST_RANK_UNSORTED) fast = false;
        uint32_t *perm_q = nullptr, *perm_r = nullptr;
        // When results stay on the device (a grouped query order would break the row-chunked copy-out of the host path) and the
        // panels are large enough to have many tiles, only the tiles that hold a marked pair are launched, over records filled in
        // beforehand (`listed`).  And the panels are regrouped so that related sketches share tiles -- if the marked pairs are
        // scattered: a collection that is already ordered by relatedness (families side by side: 20 465 tiles for 2*10^7 marked
        // pairs at configs[2], 5 % above the minimum) gains nothing from the 1.1 ms that regrouping costs there.
        bool listed = false;
        const uint2* tile_list = nullptr;
        uint32_t n_listed = 0;
        if (mode == DIST_RANK_OK && marks && !h_out && !ctx->no_dist_group && d_ref->n >= 256 && d_qry->n >= 256) {
            uint64_t n_marked = 0;
            listed = true;
            if ((rc = dist_tile_list(ctx, marks, d_qry->n, d_ref->n, nullptr, nullptr, d_qry->sizes, d_ref->sizes, &tile_list, &n_listed, &n_marked))) return rc;
            const bool scattered = (uint64_t)n_listed * 1024 > 2 * n_marked + 64 * 1024;         // tiles less than half full on average
            if (scattered || ctx->force_dist_group) {
                if ((rc = dist_group_panels(ctx, d_qry->n, d_ref->n, rows_q, rows_r, d_qry->sizes, &p32r, &p32q, &marks, &perm_q, &perm_r))) return rc;
                tile_list = nullptr;                                  // listed again in the new order, below
            }
        }
        if (mode == DIST_RANK_OK && !p32q && (rc = dist_pack_queries(ctx, d_qry->sizes, d_qry->n, rows_q, &p32q))) return rc;   // no grouping: natural order
        if (mode == DIST_RANK_TOO_BIG) {
            if ((rc = ctx->d_ref.ensure(nr16 * 16 * rows_r * 8))) return rc;
            if ((rc = ctx->d_qry.ensure(nq16 * 16 * rows_q * 8))) return rc;
            if ((rc = ctx->d_misc.ensure(64))) return rc;
            FPM_CUDA(cudaMemsetAsync(ctx->d_misc.p, 0, 64, st));
            uint64_t n1 = nr16 * 16 * rows_r, n2 = nq16 * 16 * rows_q;
            ctx->time_begin(FPM_KERNEL_DIST_PACK);
            dist_pack_kernel<<<(uint32_t)((n1 + 255) / 256), 256, 0, st>>>(*d_ref, rows_r, ctx->d_ref.as<uint64_t>(), ctx->d_misc.as<uint32_t>());
            dist_pack_kernel<<<(uint32_t)((n2 + 255) / 256), 256, 0, st>>>(*d_qry, rows_q, ctx->d_qry.as<uint64_t>(), ctx->d_misc.as<uint32_t>());
            ctx->time_end();
            ctx->launches += 2;
            FPM_CUDA(cudaGetLastError());
            uint32_t flag = 0;
            FPM_CUDA(cudaMemcpyAsync(&flag, ctx->d_misc.p, 4, cudaMemcpyDeviceToHost, st));
            FPM_CUDA(cudaStreamSynchronize(st));
            if (flag) fast = false;   // sentinel collision or unsorted input: literal loop defines the result
        }
        if (fast) {
            const bool k32 = mode == DIST_RANK_OK;
            if (hits && !k32 && (rc = matrix_for_hits())) return rc;
            const size_t smem = k32 ? (size_t)D4_COLS * D4_COLROWS * 4 : (size_t)DT_COLS * DT_COLROWS * 8;
            if (k32) FPM_CUDA(cudaFuncSetAttribute(dist_tile32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            else FPM_CUDA(cudaFuncSetAttribute(dist_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            // query-row chunks: at most 65535 tiles per launch, and ~16M pairs per chunk when streaming to the host
            const uint64_t qt = k32 ? 32 : 16;                        // query rows per tile
            const uint64_t nqt = (d_qry->n + qt - 1) / qt;
            uint64_t tiles_per_chunk = 65535;
            if (h_out) tiles_per_chunk = std::max<uint64_t>(1, std::min<uint64_t>(65535, (16ull << 20) / (qt * std::max<uint64_t>(d_ref->n, 1))));
            if (h_out && !ctx->copy_stream) {
                FPM_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
                FPM_CUDA(cudaEventCreateWithFlags(&ctx->copy_done[0], cudaEventDisableTiming));
                FPM_CUDA(cudaEventCreateWithFlags(&ctx->copy_done[1], cudaEventDisableTiming));
            }
            const bool prefill = k32 && listed && !hits && ((uintptr_t)d_out & 15) == 0;   // see dist_fill_unshared_kernel
            // nothing has to come out of a tile without a marked pair when its closed-form records are already in place (prefill)
            // or cannot pass the filters (hits): such launches run over the list of tiles that hold work -- 19 of 20 tiles of
            // configs[2] are empty, and an empty 512-thread CTA still costs its launch
            if (k32 && listed && nqt <= tiles_per_chunk && (prefill || (hits && hits->skip_unmarked))) {
                if (!tile_list && (rc = dist_tile_list(ctx, marks, d_qry->n, d_ref->n, perm_q, perm_r, d_qry->sizes, d_ref->sizes, &tile_list, &n_listed, nullptr))) return rc;
            } else tile_list = nullptr;
            uint64_t c = 0;
            for (uint64_t t0 = 0; t0 < nqt; t0 += tiles_per_chunk, c++) {
                const uint64_t nt = std::min(tiles_per_chunk, nqt - t0);
                dim3 grid((uint32_t)((d_ref->n + 31) / 32), (uint32_t)nt);
                if (tile_list) grid = dim3(n_listed, 1);
                ctx->time_begin(FPM_KERNEL_DIST_TILE);
                if (prefill && t0 == 0) {                             // once, all rows: a chunk's grouped tiles write to any row
                    dist_fill_unshared_kernel<<<(uint32_t)((total + 511) / 512), 256, 0, st>>>(a, d_ref->n, total, d_ref->sizes, d_qry->sizes,
                                                                                                d_ref->lengths, d_qry->lengths, d_out);
                    ctx->launches++;
                }
                if (k32 && tile_list && n_listed == 0)
                    ;                                                 // no pair shares a hash
                else if (k32)
                    dist_tile32_kernel<<<grid, DT_THREADS, smem, st>>>(p32r, p32q, rows_r, rows_q, d_ref->n, d_qry->n, d_ref->lengths, d_qry->lengths, a,
                                                                       d_out, (unsigned long long*)d_steps, (uint32_t)t0, marks, (uint32_t)((d_ref->n + 31) / 32), d_ref->sizes, d_qry->sizes, perm_q, perm_r, hits ? *hits : no_hits, prefill ? 1 : 0, tile_list);
                else
                    dist_tile_kernel<<<grid, DT_THREADS, smem, st>>>(ctx->d_ref.as<uint64_t>(), ctx->d_qry.as<uint64_t>(), rows_r, rows_q, d_ref->n,
                                                                     d_qry->n, d_ref->lengths, d_qry->lengths, a, d_out, (unsigned long long*)d_steps, (uint32_t)t0);
                ctx->time_end();
                ctx->launches++;
                FPM_CUDA(cudaGetLastError());
                if (h_out) {
                    const uint64_t q0 = t0 * qt, q1 = std::min<uint64_t>(d_qry->n, (t0 + nt) * qt);
                    FPM_CUDA(cudaEventRecord(ctx->copy_done[c & 1], st));
                    FPM_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_done[c & 1], 0));
                    if (h_ld == 0 || h_ld == d_ref->n)
                        FPM_CUDA(cudaMemcpyAsync(h_out + q0 * d_ref->n, d_out + q0 * d_ref->n, (q1 - q0) * d_ref->n * sizeof(fpm_pair),
                                                 cudaMemcpyDeviceToHost, ctx->copy_stream));
                    else   // this call's matrix is a block of a wider one on the host (h_ld records per row)
                        FPM_CUDA(cudaMemcpy2DAsync(h_out + q0 * h_ld, h_ld * sizeof(fpm_pair), d_out + q0 * d_ref->n, d_ref->n * sizeof(fpm_pair),
                                                   d_ref->n * sizeof(fpm_pair), q1 - q0, cudaMemcpyDeviceToHost, ctx->copy_stream));
                }
            }
            if (h_out) FPM_CUDA(cudaStreamSynchronize(ctx->copy_stream));
            if (hits && !k32 && (rc = collect())) return rc;
        }
    }
    if (!fast) {
        int rc;
        if (hits && (rc = matrix_for_hits())) return rc;
        uint64_t blocks = (total + 255) / 256;
        if (blocks > 0x7fffffffull) { set_error("too many pairs for one launch"); return FPM_ERR_ARG; }
        ctx->time_begin(FPM_KERNEL_DIST_LITERAL);
        dist_literal_kernel<<<(uint32_t)blocks, 256, 0, st>>>(*d_ref, *d_qry, a, d_out, (unsigned long long*)d_steps);
        ctx->time_end();
        ctx->launches++;
        FPM_CUDA(cudaGetLastError());
        if (h_out) {
            if (h_ld == 0 || h_ld == d_ref->n) FPM_CUDA(cudaMemcpyAsync(h_out, d_out, total * sizeof(fpm_pair), cudaMemcpyDeviceToHost, st));
            else FPM_CUDA(cudaMemcpy2DAsync(h_out, h_ld * sizeof(fpm_pair), d_out, d_ref->n * sizeof(fpm_pair), d_ref->n * sizeof(fpm_pair), d_qry->n,
                                            cudaMemcpyDeviceToHost, st));
            FPM_CUDA(cudaStreamSynchronize(st));
        }
        if (hits && (rc = collect())) return rc;
    }
    return FPM_OK;
}

int max_size_dev(fpm_ctx* ctx, const fpm_panel* d, uint32_t* out)
{
    std::vector<uint32_t> h(d->n);
    if (d->n) {
        FPM_CUDA(cudaMemcpyAsync(h.data(), d->sizes, sizeof(uint32_t) * d->n, cudaMemcpyDeviceToHost, ctx->stream));
        FPM_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    uint32_t m = 0;
    for (uint32_t v : h) m = std::max(m, v);
    *out = m;
    return FPM_OK;
}

int check_dist(const fpm_dist_params* p, const fpm_panel* ref, const fpm_panel* qry)
{
    if (!p || !ref || !qry) { set_error("NULL argument"); return FPM_ERR_ARG; }
    if (p->sketch_size >= 0x7fffffffu) { set_error("sketch size too large"); return FPM_ERR_ARG; }
    if (p->kmer_size < 1) { set_error("kmer size must be >= 1"); return FPM_ERR_ARG; }
    return FPM_OK;
}

// one block of a wider host matrix (dist_multi.cu): `out` points at the block's first record, rows are ld records apart
int dist_tile_host_ld(fpm_ctx* ctx, const fpm_dist_params* p, const fpm_panel* ref, const fpm_panel* qry, fpm_pair* out, uint64_t ld);
int dist_hits_host(fpm_ctx* ctx, const fpm_dist_params* p, const fpm_panel* ref, const fpm_panel* qry, fpm_hit* out, uint64_t capacity, uint64_t* n_hits,
                   uint32_t q_base, uint32_t r_base);

}  // namespace fpm

using namespace fpm;

extern "C" {

int fpm_dist_tile_dev(fpm_ctx* ctx, const fpm_dist_params* p, const fpm_panel* d_ref, const fpm_panel* d_qry, fpm_pair* d_out, uint64_t* d_merge_steps)
{
    if (!ctx) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    int rc = check_dist(p, d_ref, d_qry);
    if (rc) return rc;
    FPM_CUDA(cudaSetDevice(ctx->device));
    uint32_t mr = 0, mq = 0;
    if ((rc = max_size_dev(ctx, d_ref, &mr))) return rc;
    if ((rc = max_size_dev(ctx, d_qry, &mq))) return rc;
    if (mr > d_ref->stride || mq > d_qry->stride) { set_error("a sketch size exceeds the panel stride"); return FPM_ERR_ARG; }
    return run_dist(ctx, p, d_ref, d_qry, d_out, d_merge_steps, mr, mq);
}

// host panels -> device copies in ctx->d_rs / d_qs (hashes | lengths | sizes); also the largest sketch of each
// host panel -> device copy in `buf` (hashes | lengths | sizes); also the largest sketch
static int upload_panel(fpm_ctx* ctx, const fpm_panel* h, DevBuf& buf, fpm_panel* d, uint32_t* max_size)
{
    int rc;
    *max_size = 0;
    for (uint64_t i = 0; i < h->n; i++) *max_size = std::max(*max_size, h->sizes[i]);
    if (*max_size > h->stride) { set_error("a sketch size exceeds the panel stride"); return FPM_ERR_ARG; }
    cudaStream_t st = ctx->stream;
    const size_t hb = h->n * h->stride * 8;
    if ((rc = buf.ensure(hb + h->n * 12 + 64))) return rc;
    unsigned char* b = buf.as<unsigned char>();
    *d = *h;
    d->hashes = (const uint64_t*)b; d->lengths = (const uint64_t*)(b + hb); d->sizes = (const uint32_t*)(b + hb + h->n * 8);
    if (hb) FPM_CUDA(cudaMemcpyAsync((void*)d->hashes, h->hashes, hb, cudaMemcpyHostToDevice, st));
    FPM_CUDA(cudaMemcpyAsync((void*)d->lengths, h->lengths, h->n * 8, cudaMemcpyHostToDevice, st));
    FPM_CUDA(cudaMemcpyAsync((void*)d->sizes, h->sizes, h->n * 4, cudaMemcpyHostToDevice, st));
    return FPM_OK;
}

// ref == nullptr: the resident reference panel of fpm_dist_set_reference
static int upload_panels(fpm_ctx* ctx, const fpm_panel* ref, const fpm_panel* qry, fpm_panel* dr, fpm_panel* dq, uint32_t* mr, uint32_t* mq)
{
    int rc;
    if (ref) {
        ctx->ref_set = false;                                            // an explicit reference panel replaces the resident one
        ctx->rix.valid = false;
        if ((rc = upload_panel(ctx, ref, ctx->d_rs, dr, mr))) return rc;
    } else {
        if (!ctx->ref_set) { set_error("no reference panel: pass one, or call fpm_dist_set_reference first"); return FPM_ERR_ARG; }
        *dr = ctx->ref_dev; *mr = ctx->ref_max;
    }
    return upload_panel(ctx, qry, ctx->d_qs, dq, mq);
}

static int dist_tile_host(fpm_ctx* ctx, const fpm_dist_params* p, const fpm_panel* ref, const fpm_panel* qry, fpm_pair* out, bool positional, uint64_t h_ld = 0)
{
    if (!ctx) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    if (!ref && ctx->ref_set) ref = &ctx->ref_dev;                       // resident reference panel (only n / stride are read on the host)
    const bool resident = ref == &ctx->ref_dev;
    int rc = check_dist(p, ref, qry);
    if (rc) return rc;
    if (ref->n == 0 || qry->n == 0) return FPM_OK;
    FPM_CUDA(cudaSetDevice(ctx->device));
    uint32_t mr = 0, mq = 0;
    fpm_panel dr, dq;
    if ((rc = upload_panels(ctx, resident ? nullptr : ref, qry, &dr, &dq, &mr, &mq))) return rc;
    if ((rc = ctx->d_out.ensure(ref->n * qry->n * sizeof(fpm_pair)))) return rc;
    cudaStream_t st = ctx->stream;
    if (positional) {
        uint64_t total = ref->n * qry->n;
        fp_positional_kernel<<<(uint32_t)((total + 255) / 256), 256, 0, st>>>(dr, dq, p->max_distance, p->max_pvalue, ctx->d_out.as<fpm_pair>());
        ctx->launches++;
        FPM_CUDA(cudaGetLastError());
        FPM_CUDA(cudaMemcpyAsync(out, ctx->d_out.p, ref->n * qry->n * sizeof(fpm_pair), cudaMemcpyDeviceToHost, st));
    } else if ((rc = run_dist(ctx, p, &dr, &dq, ctx->d_out.as<fpm_pair>(), nullptr, mr, mq, out, nullptr, h_ld))) return rc;
    FPM_CUDA(cudaStreamSynchronize(st));
    return FPM_OK;
}

}  // extern "C"

// Both fpm_dist_hits entry points: device panels in, sorted hits at d_sorted (room for `capacity`), *n_hits on the host.
int fpm::dist_hits_run(fpm_ctx* ctx, const fpm_dist_params* p, const fpm_panel* d_ref, const fpm_panel* d_qry, uint32_t mr, uint32_t mq,
                       fpm_hit* d_sorted, uint64_t capacity, uint64_t* n_hits, uint64_t* d_steps, uint32_t q_base, uint32_t r_base)
{
    int rc;
    *n_hits = 0;
    if (d_ref->n == 0 || d_qry->n == 0) return FPM_OK;
    if (d_ref->n > 0xfffffffeull || d_qry->n > 0xfffffffeull) { set_error("panel too large for 32-bit hit indices"); return FPM_ERR_ARG; }
    if ((rc = ctx->d_hits.ensure(capacity * sizeof(fpm_hit) + 64))) return rc;
    cudaStream_t st = ctx->stream;
    HitSink hs;
    hs.count = ctx->d_hits.as<unsigned long long>();            // first 32 bytes: the counter; records follow
    hs.buf = reinterpret_cast<fpm_hit*>(ctx->d_hits.as<unsigned char>() + 32);
    hs.cap = capacity;
    hs.q_base = q_base; hs.r_base = r_base;
    // a pair without a shared hash has distance 1 and p-value 1 (dist_math.h): it passes only when neither filter excludes 1
    hs.skip_unmarked = (p->max_distance >= 0 && p->max_distance < 1.) || (p->max_pvalue >= 0 && p->max_pvalue < 1.);
    FPM_CUDA(cudaMemsetAsync(hs.count, 0, 32, st));
    if ((rc = run_dist(ctx, p, d_ref, d_qry, nullptr, d_steps, mr, mq, nullptr, &hs))) return rc;
    unsigned long long n = 0;
    FPM_CUDA(cudaMemcpyAsync(&n, hs.count, 8, cudaMemcpyDeviceToHost, st));
    FPM_CUDA(cudaStreamSynchronize(st));
    *n_hits = n;
    if (n > capacity) { set_error("fpm_dist_hits: %llu pairs pass the filters, room for %llu", n, (unsigned long long)capacity); return FPM_ERR_CAPACITY; }
    return dist_sort_hits(ctx, hs.buf, n, d_qry->n, d_ref->n, d_sorted, q_base, r_base);
}

extern "C" {

int fpm_dist_tile(fpm_ctx* ctx, const fpm_dist_params* p, const fpm_panel* ref, const fpm_panel* qry, fpm_pair* out)
{
    return dist_tile_host(ctx, p, ref, qry, out, false);
}

int fpm_dist_hits_dev(fpm_ctx* ctx, const fpm_dist_params* p, const fpm_panel* d_ref, const fpm_panel* d_qry, fpm_hit* d_out, uint64_t capacity,
                      uint64_t* n_hits, uint64_t* d_merge_steps)
{
    if (!ctx) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    if (!n_hits || (!d_out && capacity)) { set_error("NULL argument"); return FPM_ERR_ARG; }
    int rc = check_dist(p, d_ref, d_qry);
    if (rc) return rc;
    FPM_CUDA(cudaSetDevice(ctx->device));
    uint32_t mr = 0, mq = 0;
    if ((rc = max_size_dev(ctx, d_ref, &mr))) return rc;
    if ((rc = max_size_dev(ctx, d_qry, &mq))) return rc;
    if (mr > d_ref->stride || mq > d_qry->stride) { set_error("a sketch size exceeds the panel stride"); return FPM_ERR_ARG; }
    if ((rc = dist_hits_run(ctx, p, d_ref, d_qry, mr, mq, d_out, capacity, n_hits, d_merge_steps))) return rc;
    FPM_CUDA(cudaStreamSynchronize(ctx->stream));
    return FPM_OK;
}

int fpm_dist_hits(fpm_ctx* ctx, const fpm_dist_params* p, const fpm_panel* ref, const fpm_panel* qry, fpm_hit* out, uint64_t capacity, uint64_t* n_hits)
{
    return dist_hits_host(ctx, p, ref, qry, out, capacity, n_hits, 0, 0);
}

}  // extern "C"

int fpm::dist_tile_host_ld(fpm_ctx* ctx, const fpm_dist_params* p, const fpm_panel* ref, const fpm_panel* qry, fpm_pair* out, uint64_t ld)
{
    return dist_tile_host(ctx, p, ref, qry, out, false, ld);
}

int fpm::dist_hits_host(fpm_ctx* ctx, const fpm_dist_params* p, const fpm_panel* ref, const fpm_panel* qry, fpm_hit* out, uint64_t capacity, uint64_t* n_hits,
                        uint32_t q_base, uint32_t r_base)
{
    if (!ctx) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    if (!n_hits || (!out && capacity)) { set_error("NULL argument"); return FPM_ERR_ARG; }
    if (!ref && ctx->ref_set) ref = &ctx->ref_dev;
    const bool resident = ref == &ctx->ref_dev;
    int rc = check_dist(p, ref, qry);
    if (rc) return rc;
    *n_hits = 0;
    if (ref->n == 0 || qry->n == 0) return FPM_OK;
    FPM_CUDA(cudaSetDevice(ctx->device));
    uint32_t mr = 0, mq = 0;
    fpm_panel dr, dq;
    if ((rc = upload_panels(ctx, resident ? nullptr : ref, qry, &dr, &dq, &mr, &mq))) return rc;
    // d_hits = counter | appended records | sorted records (dist_hits_run's own ensure() is then a no-op: the buffer only grows)
    if ((rc = ctx->d_hits.ensure(2 * capacity * sizeof(fpm_hit) + 64))) return rc;
    fpm_hit* d_sorted = reinterpret_cast<fpm_hit*>(ctx->d_hits.as<unsigned char>() + 32) + capacity;
    if ((rc = dist_hits_run(ctx, p, &dr, &dq, mr, mq, d_sorted, capacity, n_hits, nullptr, q_base, r_base))) return rc;
    if (*n_hits) FPM_CUDA(cudaMemcpyAsync(out, d_sorted, *n_hits * sizeof(fpm_hit), cudaMemcpyDeviceToHost, ctx->stream));
    FPM_CUDA(cudaStreamSynchronize(ctx->stream));
    return FPM_OK;
}

extern "C" {

int fpm_dist_set_reference(fpm_ctx* ctx, const fpm_panel* ref)
{
    if (!ctx) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    ctx->ref_set = false;
    ctx->rix.valid = false;
    if (!ref) return FPM_OK;
    if (!ref->sizes || !ref->lengths || (!ref->hashes && ref->n * ref->stride)) { set_error("NULL panel array"); return FPM_ERR_ARG; }
    FPM_CUDA(cudaSetDevice(ctx->device));
    int rc = upload_panel(ctx, ref, ctx->d_rs, &ctx->ref_dev, &ctx->ref_max);
    if (rc) return rc;
    FPM_CUDA(cudaStreamSynchronize(ctx->stream));                        // the caller may release its host panel now
    ctx->ref_set = true;
    return FPM_OK;
}

int fpm_fp_positional_tile(fpm_ctx* ctx, const fpm_dist_params* p, const fpm_panel* ref, const fpm_panel* qry, fpm_pair* out)
{
    return dist_tile_host(ctx, p, ref, qry, out, true);
}

double fpm_pvalue(uint64_t x, uint64_t len_ref, uint64_t len_qry, double kmer_space, uint64_t n)
{
    return mash_pvalue(x, len_ref, len_qry, kmer_space, n);
}

double fpm_distance(uint64_t common, uint64_t denom, int kmer_size) { return mash_distance(common, denom, kmer_size); }

}  // extern "C"
