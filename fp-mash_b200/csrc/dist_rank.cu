// dist_rank.cu -- order-preserving 32-bit codes for the sketch panels of dist_tile32_kernel, and the inverted index.
//
// compareSketches (CommandDistance.cpp:365-400) only ever compares a hash of the REFERENCE list with a hash of the QUERY
// list, and only for order and equality.  So the 64-bit hashes are replaced by 32-bit codes that keep exactly those
// relations (a merge step then costs one 32-bit shared-memory load and one 32-bit compare per list):
//   * the reference panel alone is sorted (radix sort of (hash, destination) pairs, a library call) and every distinct
//     reference hash gets its dense rank r:      code = 2 r + 1;
//   * a query hash is looked up among the distinct reference hashes (a bucket table over the value range narrows the
//     binary search to a handful of entries): equal to reference hash r -> code 2 r + 1, otherwise 2 * (number of distinct
//     reference hashes below it) -- an even code that sits strictly between its neighbours' odd ones.  Two different
//     query hashes between the same two reference hashes share a code; they are never compared with each other.
//   0xffffffff is the +inf sentinel.
// Round 1 sorted both panels together (dense ranks over the union): twice the sort for an all-vs-all, and a reference
// index that had to be rebuilt for every query chunk.  Now the index is a property of the reference panel alone.
//
// Steps (all on the caller's stream):
//   dist_keys_kernel      every valid reference element -> (hash, destination in the packed tile layout); validates "strictly ascending"
//   cub radix sort        by hash (library call, like the scans below)
//   cub inclusive scan    of "differs from predecessor" = dense rank
//   dist_index_kernel     codes -> packed reference tiles; distinct hashes dk[]; posting lists post[] / run_start[]
//   dist_bucket_* kernels bucket table start[] over dk[] (heads, suffix-min scan)
//   dist_qcode_kernel     every query element: lookup -> code -> packed query tiles; validates; counts the postings a marking pass would walk
// The sorted reference array is the inverted index everything below uses:
//   dist_uf_* kernels    connected components of the "shares a hash" graph (bounds what a query can reach)
//   dist_mark_kernel     per query, one bit per reference that shares a hash with it (pairs without one need no merge)
//   dist_group_panels    both panels re-ordered so that related sketches share tiles; dist_tile_list: tiles with work
//   dist_sort_hits       fpm_dist_hits: appended hits -> the reference's output order
#include <cub/cub.cuh>
#include "common.h"
#include "dist_rank.h"

namespace fpm {

// keys/dst are indexed by out_base + idx, idx over n * max_size (sketch-major)
__global__ void __launch_bounds__(256) dist_keys_kernel(fpm_panel pn, uint32_t max_size, uint64_t rows, uint32_t dst_base, uint64_t out_base,
                                                        uint64_t* __restrict__ keys, uint32_t* __restrict__ dst, uint32_t* flags)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= pn.n * max_size) return;
    const uint64_t sk = idx / max_size, row = idx % max_size;
    uint64_t v = ~0ULL;
    if (row < pn.sizes[sk]) {
        v = pn.hashes[sk * pn.stride + row];
        bool bad = v == ~0ULL;                                            // reserved: sorts with the invalid slots
        if (row > 0 && pn.hashes[sk * pn.stride + row - 1] >= v) bad = true;   // not strictly ascending
        if (bad) atomicOr(flags, 1u);
    }
    keys[out_base + idx] = v;
    dst[out_base + idx] = dst_base + (uint32_t)(((sk >> 4) * rows + row) * 16 + (sk & 15));
}

struct HeadFlag {
    const uint64_t* keys;
    __host__ __device__ uint32_t operator()(uint64_t i) const { return i > 0 && keys[i] != keys[i - 1] ? 1u : 0u; }
};

// ---------------------------------------------------------------------------------------------------------
// Index of the reference panel.  scal[]: [0] D = number of distinct hashes, [1] number of valid elements, [2] shift of the
// bucket table, [3] the largest hash (two words), [5] flags.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dist_index_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ dst, const uint32_t* __restrict__ rank,
                                                         uint64_t m, uint32_t rows_r, uint32_t* __restrict__ packed, uint64_t* __restrict__ dk,
                                                         uint32_t* __restrict__ post, uint32_t* __restrict__ run_start, uint32_t* __restrict__ scal)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint64_t k = keys[i];
    if (k == ~0ULL) return;                                                 // the invalid slots sort behind every hash
    const uint32_t r = rank[i], d = dst[i];
    packed[d] = 2u * r + 1u;
    post[i] = ((d >> 4) / rows_r) * 16 + (d & 15);                          // sketch index from the tile-layout destination
    if (i == 0 || keys[i - 1] != k) { dk[r] = k; run_start[r] = (uint32_t)i; }
    if (i + 1 == m || keys[i + 1] == ~0ULL) {                              // the last valid element
        scal[0] = r + 1; scal[1] = (uint32_t)(i + 1);
        run_start[r + 1] = (uint32_t)(i + 1);
        scal[3] = (uint32_t)k; scal[4] = (uint32_t)(k >> 32);
    }
}

// bucket b of the table holds the distinct hashes h with (h >> shift) == b; shift makes the largest hash land in the last of B
// buckets.  cum[b] = number of distinct hashes in buckets 0..b: written by the last hash of every bucket, completed by a prefix maximum.
__global__ void dist_bucket_setup_kernel(uint32_t* __restrict__ scal, uint32_t log2_buckets)
{
    const uint64_t maxkey = ((uint64_t)scal[4] << 32) | scal[3];
    const int bits = 64 - __clzll((long long)maxkey);
    scal[2] = bits > (int)log2_buckets ? (uint32_t)(bits - (int)log2_buckets) : 0u;
}

__global__ void __launch_bounds__(256) dist_bucket_tails_kernel(const uint64_t* __restrict__ dk, const uint32_t* __restrict__ scal, uint32_t* __restrict__ cum)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x, D = scal[0];
    if (r >= D) return;
    const uint32_t shift = scal[2];
    const uint64_t b = dk[r] >> shift;
    if (r + 1 == D || (dk[r + 1] >> shift) != b) cum[b] = r + 1;
}

struct MaxOp {
    __host__ __device__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

struct RefIndex {
    const uint64_t* dk;          // distinct reference hashes, ascending
    const uint32_t* cum;         // [n_buckets] distinct hashes in buckets 0..b
    const uint32_t* scal;
};

// number of distinct reference hashes below h, and whether h is one of them
__device__ __forceinline__ uint32_t ref_lower_bound(const RefIndex& ix, uint64_t h, uint32_t D, uint32_t shift, uint64_t maxkey, bool* found)
{
    *found = false;
    if (D == 0 || h > maxkey) return D;
    const uint64_t b = h >> shift;
    uint32_t lo = b ? ix.cum[b - 1] : 0u, hi = ix.cum[b];
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (ix.dk[mid] < h) lo = mid + 1; else hi = mid;
    }
    *found = lo < D && ix.dk[lo] == h;
    return lo;
}

// Every query element: its code into the packed query tiles; flags |= 1 for a row that is not strictly ascending or holds
// the reserved value; *total += postings a marking pass would walk for it (references sharing its hash).
// A CTA owns a tile of 32 sketches x 32 rows.  The hashes come in row-major (one sketch's consecutive rows: 256 contiguous
// bytes per warp load) and go through shared memory so that a warp then works on ONE row of 32 sketches:
//   * row r of every sorted sketch is the same quantile of the hash range, so the warp's 32 lookups -- and those of the CTAs
//     running beside it, the grid walks sketch tiles fastest -- fall into one narrow band of the bucket table and of dk[],
//     which the L2 holds; with one sketch per warp the lookups were scattered over the whole index (0.8 GB for configs[4]),
//   * the codes go back through shared memory into a ROW-MAJOR array codes[sketch][row] (128 contiguous bytes per warp store).
//     The column tiles the tile kernel reads are written later by dist_pack_queries_kernel, in natural or in grouped order:
//     packing first and regrouping afterwards gathered single columns out of column tiles, 4 useful bytes per 32-byte sector
//     (10 ms for configs[4]); the marking kernel reads a query's codes contiguously instead of 64 bytes apart.
// configs[4] (10^9 lookups into 10^8 reference hashes): 48 ms before (profiles/r02_c5_launches.csv).
constexpr int QC_TILE = 32;
__global__ void __launch_bounds__(256) dist_qcode_kernel(fpm_panel pn, uint32_t max_size, RefIndex ix, const uint32_t* __restrict__ run_start,
                                                         uint32_t* __restrict__ codes, uint32_t* flags, unsigned long long* total)
{
    __shared__ uint64_t s_h[QC_TILE][QC_TILE + 1];          // [sketch][row], padded: the transposed reads are conflict free
    __shared__ uint32_t s_c[QC_TILE][QC_TILE + 1];          // the codes on their way back
    const uint64_t sk0 = (uint64_t)blockIdx.x * QC_TILE, row0 = (uint64_t)blockIdx.y * QC_TILE;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned long long c = 0;
    // ---- load: warp w takes sketches 4w .. 4w+3, lane = row --------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < QC_TILE / 8; i++) {
        const int s = wid * (QC_TILE / 8) + i;
        const uint64_t sk = sk0 + s, row = row0 + lane;
        const bool in = sk < pn.n && row < pn.sizes[sk];
        const uint64_t v = in ? pn.hashes[sk * pn.stride + row] : ~0ULL;
        uint64_t prev = __shfl_up_sync(0xffffffffu, v, 1);      // (a row inside the sketch has its predecessor inside it too)
        bool bad = false;
        if (in) {
            if (lane == 0 && row > 0) prev = pn.hashes[sk * pn.stride + row - 1];
            bad = v == ~0ULL || (row > 0 && prev >= v);
        }
        s_h[s][lane] = v;
        if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(flags, 1u);
    }
    __syncthreads();
    // ---- look up: warp w takes rows 4w .. 4w+3, lane = sketch ------------------------------------------------------------
    const uint32_t D = ix.scal[0], shift = ix.scal[2];
    const uint64_t maxkey = ((uint64_t)ix.scal[4] << 32) | ix.scal[3];
#pragma unroll
    for (int i = 0; i < QC_TILE / 8; i++) {
        const int r = wid * (QC_TILE / 8) + i;
        const uint64_t sk = sk0 + lane, row = row0 + r;
        const uint64_t v = s_h[lane][r];
        uint32_t code = 0xffffffffu;
        if (sk < pn.n && row < max_size && row < pn.sizes[sk]) {
            bool found;
            const uint32_t lb = ref_lower_bound(ix, v, D, shift, maxkey, &found);
            code = 2u * lb + (found ? 1u : 0u);
            if (found) c += run_start[lb + 1] - run_start[lb];
        }
        s_c[lane][r] = code;
    }
    __syncthreads();
    // ---- store: warp w takes sketches 4w .. 4w+3 again, lane = row: 128 contiguous bytes of the row-major code array ------
#pragma unroll
    for (int i = 0; i < QC_TILE / 8; i++) {
        const int s = wid * (QC_TILE / 8) + i;
        const uint64_t sk = sk0 + s, row = row0 + lane;
        if (sk < pn.n && row < max_size) codes[sk * max_size + row] = s_c[s][lane];
    }
    for (int o = 16; o; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    // one atomic per CTA: a million same-address atomics (one per warp) cost 0.6 ms
    __shared__ unsigned long long s_part[8];
    if (lane == 0) s_part[wid] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; w++) t += s_part[w];
        if (t) atomicAdd(total, t);
    }
}

// Saturation.  Relatives repeat each other's postings: a query from a family of 1000 mutually related genomes walks
// ~600 posting lists of ~600 references each to set the same 1000 bits.  The walk may stop as soon as every reference
// the query can possibly reach is marked, and that set is known cheaply: references that hold a common hash are connected,
// so all references of one posting list lie in one CONNECTED COMPONENT of the "shares a hash" graph over the references,
// and everything a query can reach lies in the components its posting lists start in.  A lock-free union-find over
// neighbours in the sorted reference array gives the components and their sizes in one pass; a query's marking pass first
// adds up the sizes of the (usually one or two) components it touches and stops once it has set that many bits.  Exact
// (a stopped walk could not have set another bit); a query in a huge sparse component simply never stops early.
// Parents only ever decrease (a root is hooked under a smaller node, halving moves to a grandparent), so ANY value a
// parent entry held at some time is an ancestor-or-self: finds may read through L1 (FRESH = false; millions of threads
// read the same few roots, and going to L2 for each of those reads was 2 ms) and only need current data when a hook failed.
template <bool FRESH>
__device__ __forceinline__ uint32_t uf_find(uint32_t* __restrict__ parent, uint32_t x)
{
    for (;;) {
        const uint32_t p = FRESH ? reinterpret_cast<volatile uint32_t*>(parent)[x] : __ldca(parent + x);
        if (p == x) return x;
        const uint32_t g = FRESH ? reinterpret_cast<volatile uint32_t*>(parent)[p] : __ldca(parent + p);
        if (g != p) parent[x] = g;                                        // path halving
        x = p;
    }
}

__global__ void __launch_bounds__(256) dist_uf_init_kernel(uint32_t* __restrict__ parent, uint32_t* __restrict__ ref_count, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { parent[i] = i; ref_count[i] = 0; }
}

__device__ __forceinline__ void uf_union(uint32_t* __restrict__ parent, uint32_t a, uint32_t b)
{
    a = uf_find<false>(parent, a);
    b = uf_find<false>(parent, b);
    for (;;) {                                                             // hook the larger root under the smaller node
        if (a == b) return;
        if (a < b) { const uint32_t t = a; a = b; b = t; }
        if (atomicCAS(&parent[a], a, b) == a) return;
        a = uf_find<true>(parent, a);
        b = uf_find<true>(parent, b);
    }
}

// nodes: the references.  References holding the same hash are neighbours in the sorted array, in ascending sketch order (the
// radix sort is stable and the keys were generated sketch by sketch): every member of a posting list is joined with the
// list's FIRST member, its smallest.  (Joining neighbours instead builds chains that the hooks then have to walk: a few
// hundred dependent L2 round trips, 0.5 ms whatever the panel size.)
__global__ void __launch_bounds__(256) dist_uf_union_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ post, const uint32_t* __restrict__ rank,
                                                            const uint32_t* __restrict__ run_start, uint64_t m, uint32_t* __restrict__ parent)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (i >= m) return;
    const uint64_t k = keys[i];
    if (k == ~0ULL || keys[i - 1] != k) return;
    uf_union(parent, post[i], post[run_start[rank[i]]]);
}

// parent[x] = root for every reference; ref_count[root] = references in the component.  The find here must not compress
// paths: this kernel publishes roots into the array other threads are still walking, and a path-halving store of a
// (non-root) grandparent landing after the final "parent[x] = root" would leave a node pointing at a non-root whose
// ref_count is 0.  With a read-only walk every value ever stored from here on is a root, and a walker that meets one
// early simply arrives sooner.
__device__ __forceinline__ uint32_t uf_root_readonly(const uint32_t* parent, uint32_t x)
{
    for (;;) {
        const uint32_t p = reinterpret_cast<const volatile uint32_t*>(parent)[x];
        if (p == x) return x;
        x = p;
    }
}

__global__ void __launch_bounds__(256) dist_uf_flatten_kernel(uint32_t* __restrict__ parent, uint32_t* __restrict__ ref_count, uint32_t n_r)
{
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_r) return;
    const uint32_t r = uf_root_readonly(parent, x);
    atomicAdd(&ref_count[r], 1u);
    parent[x] = r;
}

// one CTA per query sketch: bit r of its row = reference r shares a hash with it.  (Also tried: one identity for equal posting
// lists, each distinct list walked once per query -- lists of independently mutated relatives are all different; and testing
// the bit with a plain load before the atomic -- 0.7 ms slower.)
constexpr int MARK_ROOTS = 32;       // components a query may touch before the early stop is given up for it
__global__ void __launch_bounds__(256) dist_mark_kernel(const uint32_t* __restrict__ codes, uint64_t code_stride, const uint32_t* __restrict__ sizes_q,
                                                        const uint32_t* __restrict__ run_ref_start, const uint32_t* __restrict__ post, uint32_t words,
                                                        const uint32_t* __restrict__ parent, const uint32_t* __restrict__ ref_count, uint32_t n_r,
                                                        uint32_t* __restrict__ marks)
{
    extern __shared__ uint32_t s_bits[];
    __shared__ uint32_t s_count, s_nroots, s_roots[MARK_ROOTS], s_target;
    const uint32_t q = blockIdx.x;
    for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) s_bits[w] = 0;
    if (threadIdx.x == 0) { s_count = 0; s_nroots = 0; s_target = 0xffffffffu; }
    __syncthreads();
    const uint32_t n = sizes_q[q];
    const uint32_t* col = codes + (uint64_t)q * code_stride;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (parent) {
        // references this query can reach at all: the sizes of the components its posting lists start in.  A component listed
        // twice (two warps racing) only makes the bound larger, which keeps it a bound.
        for (uint32_t e0 = 0; e0 < n; e0 += blockDim.x) {                   // (uniform trip count: the warp votes below)
            const uint32_t e = e0 + threadIdx.x;
            uint32_t root = 0xffffffffu;
            if (e < n) {
                const uint32_t code = col[e];
                if (code & 1u) root = parent[post[run_ref_start[code >> 1]]];
            }
            // one lane per distinct component of the warp looks it up in the list and appends it if it is new
            const uint32_t peers = __match_any_sync(0xffffffffu, root);
            if (root != 0xffffffffu && __ffs(peers) - 1 == lane) {
                bool seen = false;
                const uint32_t have = min(*reinterpret_cast<volatile uint32_t*>(&s_nroots), (uint32_t)MARK_ROOTS);
                for (uint32_t i = 0; i < have; i++) seen |= reinterpret_cast<volatile uint32_t*>(s_roots)[i] == root;
                if (!seen) {
                    const uint32_t slot = atomicAdd(&s_nroots, 1u);
                    if (slot < MARK_ROOTS) s_roots[slot] = root;
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 && s_nroots <= MARK_ROOTS) {
            // (duplicates from races are dropped here: the list is tiny)
            uint32_t t = 0;
            for (uint32_t i = 0; i < s_nroots; i++) {
                bool dup = false;
                for (uint32_t j = 0; j < i; j++) dup |= s_roots[j] == s_roots[i];
                if (!dup) t += ref_count[s_roots[i]];
            }
            s_target = t;
        }
        __syncthreads();
    }
    const uint32_t target = s_target;
    for (uint32_t e = wid; e < n; e += nw) {                              // a warp per element, lanes over its postings
        if (*reinterpret_cast<volatile uint32_t*>(&s_count) >= target) break;
        const uint32_t code = col[e];
        if (!(code & 1u)) continue;                                       // even code: no reference holds this hash
        const uint32_t r = code >> 1;
        const uint32_t lo = run_ref_start[r], hi = run_ref_start[r + 1];
        uint32_t fresh = 0;
        for (uint32_t j0 = lo; j0 < hi; j0 += 32) {
            const uint32_t j = j0 + lane;
            bool isnew = false;
            if (j < hi) {
                const uint32_t sk = post[j], bit = 1u << (sk & 31);
                isnew = !(atomicOr(&s_bits[sk >> 5], bit) & bit);
            }
            fresh += __popc(__ballot_sync(0xffffffffu, isnew));
        }
        if (lane == 0 && fresh) atomicAdd(&s_count, fresh);
    }
    __syncthreads();
    for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) marks[(uint64_t)q * words + w] = s_bits[w];
}

// ---------------------------------------------------------------------------------------------------------
// Grouping.  A 32 x 32 tile is either skipped (no marked pair), merged at full efficiency (many marked pairs) or
// merged by a nearly idle CTA (a few).  Related sketches that sit next to each other make tiles of the first two
// kinds, so both panels are put in an order in which they do: queries by their first marked reference, references
// by the first query that marks them (stable sorts).  A heuristic -- results never depend on it.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dist_group_key_kernel(const uint32_t* __restrict__ marks, uint32_t n_q, uint32_t n_r, uint32_t words,
                                                             uint32_t* __restrict__ key_q, uint32_t* __restrict__ key_r)
{
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= n_q) return;
    uint32_t first = n_r;
    for (uint32_t w = lane; w < words; w += 32) {
        uint32_t m = marks[(uint64_t)q * words + w];
        if (m && first == n_r) first = w * 32 + (__ffs(m) - 1);
        for (; m; m &= m - 1) atomicMin(&key_r[w * 32 + (__ffs(m) - 1)], q);
    }
    for (int o = 16; o; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    if (lane == 0) key_q[q] = first;
}

__global__ void __launch_bounds__(256) dist_iota_kernel(uint32_t* v, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}

// column tiles in the new order: sketch sk' of the output is sketch perm[sk'] of the input
__global__ void __launch_bounds__(256) dist_repack_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, const uint32_t* __restrict__ perm,
                                                          uint64_t n, uint64_t rows)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;     // over n16*16 x rows, column fastest
    const uint64_t n16 = (n + 15) / 16;
    if (idx >= n16 * 16 * rows) return;
    const uint64_t c = idx & 15, row = (idx >> 4) % rows, tile = (idx >> 4) / rows, sk = tile * 16 + c;
    uint32_t v = 0xffffffffu;
    if (sk < n) { const uint64_t o = perm[sk]; v = src[(((o >> 4) * rows + row) << 4) + (o & 15)]; }
    dst[idx] = v;
}

// one CTA per output row: the source row sits in shared memory, a warp builds one output word per ballot
__global__ void __launch_bounds__(256) dist_marks_permute_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, const uint32_t* __restrict__ perm_q,
                                                                 const uint32_t* __restrict__ perm_r, uint32_t n_q, uint32_t n_r, uint32_t words)
{
    extern __shared__ uint32_t s_row[];
    const uint32_t q2 = blockIdx.x;
    const uint32_t* row = src + (uint64_t)perm_q[q2] * words;
    for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) s_row[w] = row[w];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (uint32_t w2 = threadIdx.x >> 5; w2 < words; w2 += blockDim.x >> 5) {
        const uint32_t i = w2 * 32 + lane;
        bool bit = false;
        if (i < n_r) { const uint32_t r = perm_r[i]; bit = (s_row[r >> 5] >> (r & 31)) & 1u; }
        const uint32_t out = __ballot_sync(0xffffffffu, bit);
        if (lane == 0) dst[(uint64_t)q2 * words + w2] = out;
    }
}

// Row-major query codes -> column tiles [n/16][rows][16], sketch sk' of the output = sketch perm[sk'] of the input (perm null:
// natural order); +inf beyond each sketch's size.  A CTA builds 16 columns x 32 rows: coalesced row-major reads (a warp reads
// 32 consecutive codes of one sketch), transposed in shared memory, written as 64-byte row segments.
__global__ void __launch_bounds__(256) dist_pack_queries_kernel(const uint32_t* __restrict__ codes, uint64_t code_stride, const uint32_t* __restrict__ sizes,
                                                                const uint32_t* __restrict__ perm, uint64_t n, uint64_t rows, uint32_t* __restrict__ dst)
{
    __shared__ uint32_t s_t[16][33];
    const uint64_t tile = blockIdx.x, row0 = (uint64_t)blockIdx.y * 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int c = wid * 2 + i;
        const uint64_t sk = tile * 16 + c, row = row0 + lane;
        uint32_t v = 0xffffffffu;
        if (sk < n) {
            const uint64_t o = perm ? perm[sk] : sk;
            if (row < sizes[o]) v = codes[o * code_stride + row];
        }
        s_t[c][lane] = v;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int r = (threadIdx.x >> 4) + 16 * i, c = threadIdx.x & 15;
        const uint64_t row = row0 + r;
        if (row < rows) dst[(tile * rows + row) * 16 + c] = s_t[c][r];
    }
}

static void launch_pack_queries(cudaStream_t st, const uint32_t* codes, uint64_t code_stride, const uint32_t* sizes, const uint32_t* perm, uint64_t n, uint64_t rows,
                                uint32_t* dst)
{
    const dim3 grid((uint32_t)((n + 15) / 16), (uint32_t)((rows + 31) / 32));
    if (n && rows) dist_pack_queries_kernel<<<grid, 256, 0, st>>>(codes, code_stride, sizes, perm, n, rows, dst);
}

// natural order (no grouping): into ctx->d_p32q
int dist_pack_queries(fpm_ctx* ctx, const uint32_t* d_sizes_q, uint64_t n_q, uint64_t rows_q, uint32_t** p32q)
{
    const uint64_t pq = ((n_q + 15) / 16) * 16 * rows_q;
    int rc;
    if ((rc = ctx->d_p32q.ensure(pq * 4 + 64))) return rc;
    *p32q = ctx->d_p32q.as<uint32_t>();
    launch_pack_queries(ctx->stream, ctx->d_codes.as<uint32_t>(), ctx->code_stride, d_sizes_q, nullptr, n_q, rows_q, *p32q);
    ctx->launches++;
    FPM_CUDA(cudaGetLastError());
    return FPM_OK;
}

int dist_group_panels(fpm_ctx* ctx, uint64_t n_q, uint64_t n_r, uint64_t rows_q, uint64_t rows_r, const uint32_t* d_sizes_q, uint32_t** p32r, uint32_t** p32q,
                      uint32_t** marks, uint32_t** perm_q, uint32_t** perm_r)
{
    cudaStream_t st = ctx->stream;
    const uint32_t words = (uint32_t)((n_r + 31) / 32);
    const uint64_t nr16 = (n_r + 15) / 16, nq16 = (n_q + 15) / 16;
    const uint64_t pr = nr16 * 16 * rows_r, pq = nq16 * 16 * rows_q;
    size_t tmp_q = 0, tmp_r = 0;
    FPM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_q, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)n_q, 0, 32, st));
    FPM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_r, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)n_r, 0, 32, st));
    const size_t tmp = std::max(tmp_q, tmp_r);
    const size_t aq = (n_q * 4 + 255) & ~(size_t)255, ar = (n_r * 4 + 255) & ~(size_t)255, am = ((size_t)n_q * words * 4 + 255) & ~(size_t)255;
    const size_t ap = ((pr + pq) * 4 + 255) & ~(size_t)255;
    int rc;
    if ((rc = ctx->d_group.ensure(4 * aq + 4 * ar + am + ap + tmp + 256))) return rc;
    unsigned char* b = ctx->d_group.as<unsigned char>();
    uint32_t* key_q = (uint32_t*)b; uint32_t* key_q2 = (uint32_t*)(b + aq); uint32_t* idx_q = (uint32_t*)(b + 2 * aq); uint32_t* pq_out = (uint32_t*)(b + 3 * aq);
    b += 4 * aq;
    uint32_t* key_r = (uint32_t*)b; uint32_t* key_r2 = (uint32_t*)(b + ar); uint32_t* idx_r = (uint32_t*)(b + 2 * ar); uint32_t* pr_out = (uint32_t*)(b + 3 * ar);
    b += 4 * ar;
    uint32_t* marks2 = (uint32_t*)b; b += am;
    uint32_t* packed2 = (uint32_t*)b; b += ap;
    void* d_tmp = b;
    FPM_CUDA(cudaMemsetAsync(key_r, 0xff, n_r * 4, st));
    dist_group_key_kernel<<<(uint32_t)((n_q * 32 + 255) / 256), 256, 0, st>>>(*marks, (uint32_t)n_q, (uint32_t)n_r, words, key_q, key_r);
    dist_iota_kernel<<<(uint32_t)((n_q + 255) / 256), 256, 0, st>>>(idx_q, (uint32_t)n_q);
    dist_iota_kernel<<<(uint32_t)((n_r + 255) / 256), 256, 0, st>>>(idx_r, (uint32_t)n_r);
    FPM_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_q, key_q, key_q2, idx_q, pq_out, (int64_t)n_q, 0, 32, st));
    FPM_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_r, key_r, key_r2, idx_r, pr_out, (int64_t)n_r, 0, 32, st));
    dist_repack_kernel<<<(uint32_t)((pr + 255) / 256), 256, 0, st>>>(*p32r, packed2, pr_out, n_r, rows_r);
    launch_pack_queries(st, ctx->d_codes.as<uint32_t>(), ctx->code_stride, d_sizes_q, pq_out, n_q, rows_q, packed2 + pr);   // the queries go straight from their row-major codes into grouped tiles
    FPM_CUDA(cudaFuncSetAttribute(dist_marks_permute_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(words * 4)));
    dist_marks_permute_kernel<<<(uint32_t)n_q, 256, words * 4, st>>>(*marks, marks2, pq_out, pr_out, (uint32_t)n_q, (uint32_t)n_r, words);
    ctx->launches += 8;
    FPM_CUDA(cudaGetLastError());
    *p32r = packed2; *p32q = packed2 + pr; *marks = marks2; *perm_q = pq_out; *perm_r = pr_out;
    return FPM_OK;
}

// ---- tiles that hold work ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dist_tile_empty_kernel(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ sizes, uint32_t n, uint32_t* __restrict__ has_empty)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;            // grouped position
    if (g < n && sizes[perm ? perm[g] : g] == 0) has_empty[g >> 5] = 1;
}

// a thread per (query tile, reference tile = one word of the permuted bitmaps)
__global__ void __launch_bounds__(256) dist_tile_list_kernel(const uint32_t* __restrict__ marks, uint32_t n_q, uint32_t words, const uint32_t* __restrict__ empty_q,
                                                             const uint32_t* __restrict__ empty_r, uint2* __restrict__ list, uint32_t* __restrict__ count, unsigned long long* __restrict__ n_pairs)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t q_tiles = (n_q + 31) / 32;
    bool work = false;
    uint32_t qt = 0, rt = 0, my_pairs = 0;
    if (idx < (uint64_t)q_tiles * words) {
        qt = (uint32_t)(idx / words); rt = (uint32_t)(idx % words);
        uint32_t any = 0, pairs = 0;
        for (uint32_t q = qt * 32; q < min(n_q, qt * 32 + 32); q++) { const uint32_t mk = marks[(uint64_t)q * words + rt]; any |= mk; pairs += __popc(mk); }
        work = any != 0 || (empty_q[qt] && empty_r[rt]);
        my_pairs = pairs;
    }
    // marked pairs in all: one atomic per warp (in natural order every tile of an unordered collection holds some)
    for (int o = 16; o; o >>= 1) my_pairs += __shfl_down_sync(0xffffffffu, my_pairs, o);
    if ((threadIdx.x & 31) == 0 && my_pairs) atomicAdd(n_pairs, (unsigned long long)my_pairs);
    // warp-aggregated append: the list stays roughly in row-major tile order
    const uint32_t m = __ballot_sync(0xffffffffu, work);
    if (!m) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(count, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (work) list[base + __popc(m & ((1u << lane) - 1u))] = make_uint2(qt, rt);
}

int dist_tile_list(fpm_ctx* ctx, const uint32_t* marks, uint64_t n_q, uint64_t n_r, const uint32_t* perm_q, const uint32_t* perm_r,
                   const uint32_t* size_q, const uint32_t* size_r, const uint2** list, uint32_t* n_listed, uint64_t* n_pairs)
{
    cudaStream_t st = ctx->stream;
    const uint32_t words = (uint32_t)((n_r + 31) / 32), q_tiles = (uint32_t)((n_q + 31) / 32);
    const uint64_t tiles = (uint64_t)q_tiles * words;
    const size_t a_e = (((size_t)q_tiles + words) * 4 + 255) & ~(size_t)255;
    int rc;
    if ((rc = ctx->d_tiles.ensure(256 + a_e + tiles * sizeof(uint2)))) return rc;
    unsigned char* b = ctx->d_tiles.as<unsigned char>();
    uint32_t* count = (uint32_t*)b;
    uint32_t* empty_q = (uint32_t*)(b + 256); uint32_t* empty_r = empty_q + q_tiles;
    uint2* out = (uint2*)(b + 256 + a_e);
    FPM_CUDA(cudaMemsetAsync(b, 0, 256 + a_e, st));
    dist_tile_empty_kernel<<<(uint32_t)((n_q + 255) / 256), 256, 0, st>>>(perm_q, size_q, (uint32_t)n_q, empty_q);
    dist_tile_empty_kernel<<<(uint32_t)((n_r + 255) / 256), 256, 0, st>>>(perm_r, size_r, (uint32_t)n_r, empty_r);
    dist_tile_list_kernel<<<(uint32_t)((tiles + 255) / 256), 256, 0, st>>>(marks, (uint32_t)n_q, words, empty_q, empty_r, out, count, (unsigned long long*)(b + 8));
    ctx->launches += 3;
    FPM_CUDA(cudaGetLastError());
    uint64_t h[2] = {0, 0};
    FPM_CUDA(cudaMemcpyAsync(h, b, 16, cudaMemcpyDeviceToHost, st));
    FPM_CUDA(cudaStreamSynchronize(st));
    *n_listed = (uint32_t)h[0];
    if (n_pairs) *n_pairs = h[1];
    *list = out;
    return FPM_OK;
}

// ---- fpm_dist_hits: order the appended hits as the reference prints them (query-major, CommandDistance.cpp:303-333) ----
__global__ void __launch_bounds__(256) hit_keys_kernel(const fpm_hit* __restrict__ hits, uint64_t n, uint64_t n_ref, uint32_t q_base, uint32_t r_base,
                                                       uint64_t* __restrict__ keys, uint32_t* __restrict__ idx)
{
    const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const uint2 qr = *reinterpret_cast<const uint2*>(hits + i);
    keys[i] = (uint64_t)(qr.x - q_base) * n_ref + (qr.y - r_base);          // the pair's position in the reference's output order
    idx[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) hit_gather_kernel(const fpm_hit* __restrict__ in, const uint32_t* __restrict__ idx, uint64_t n, fpm_hit* __restrict__ out)
{
    // two threads per 32-byte record, 16 bytes each
    const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= 2 * n) return;
    reinterpret_cast<uint4*>(out)[i] = reinterpret_cast<const uint4*>(in)[2ull * idx[i >> 1] + (i & 1)];
}

int dist_sort_hits(fpm_ctx* ctx, const fpm_hit* in, uint64_t n, uint64_t n_qry, uint64_t n_ref, fpm_hit* out, uint32_t q_base, uint32_t r_base)
{
    int bits = 1;
    while (bits < 64 && ((n_qry * n_ref - 1) >> bits)) bits++;   // radix passes only over the bits a pair index can have

    if (n == 0) return FPM_OK;
    if (n >= 0xffffffffull) { set_error("more than 2^32 hits in one call"); return FPM_ERR_ARG; }
    cudaStream_t st = ctx->stream;
    size_t tmp = 0;
    cub::DoubleBuffer<uint64_t> kb(nullptr, nullptr);
    cub::DoubleBuffer<uint32_t> vb(nullptr, nullptr);
    FPM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, kb, vb, (int64_t)n, 0, bits, st));
    const size_t ak = (n * 8 + 255) & ~(size_t)255, av = (n * 4 + 255) & ~(size_t)255;
    int rc;
    if ((rc = ctx->d_hsort.ensure(2 * ak + 2 * av + tmp + 256))) return rc;
    unsigned char* b = ctx->d_hsort.as<unsigned char>();
    kb = cub::DoubleBuffer<uint64_t>((uint64_t*)b, (uint64_t*)(b + ak));
    vb = cub::DoubleBuffer<uint32_t>((uint32_t*)(b + 2 * ak), (uint32_t*)(b + 2 * ak + av));
    void* d_tmp = b + 2 * ak + 2 * av;
    hit_keys_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, st>>>(in, n, n_ref, q_base, r_base, kb.Current(), vb.Current());
    FPM_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmp, kb, vb, (int64_t)n, 0, bits, st));
    hit_gather_kernel<<<(uint32_t)((2 * n + 255) / 256), 256, 0, st>>>(in, vb.Current(), n, out);
    ctx->launches += 3;
    FPM_CUDA(cudaGetLastError());
    return FPM_OK;
}

int dist_rank_panels(fpm_ctx* ctx, const fpm_panel* d_ref, const fpm_panel* d_qry, uint32_t max_size_ref, uint32_t max_size_qry,
                     uint64_t rows_r, uint64_t rows_q, uint32_t sketch_size, uint32_t** packed_ref, uint32_t** packed_qry, int* mode, uint32_t** marks)
{
    *marks = nullptr;
    cudaStream_t st = ctx->stream;
    *mode = DIST_RANK_TOO_BIG;
    const uint64_t nr16 = (d_ref->n + 15) / 16, nq16 = (d_qry->n + 15) / 16;
    const uint64_t pr = nr16 * 16 * rows_r, pq = nq16 * 16 * rows_q;      // packed elements
    const uint64_t mr = d_ref->n * (uint64_t)max_size_ref, mq = d_qry->n * (uint64_t)max_size_qry, m = mr;
    if (pr >= 0xffffffffull || pq >= 0xffffffffull || mr >= 0x7fffffffull) { ctx->rix.valid = false; return FPM_OK; }   // codes / indices would not fit: 64-bit path
    int rc;
    if ((rc = ctx->d_misc.ensure(64))) return rc;
    FPM_CUDA(cudaMemsetAsync(ctx->d_misc.p, 0, 64, st));
    uint32_t* d_flag_r = ctx->d_misc.as<uint32_t>();
    uint32_t* d_flag_q = d_flag_r + 1;
    unsigned long long* d_total = (unsigned long long*)(ctx->d_misc.as<uint64_t>() + 1);

    // bucket table over the distinct reference hashes: about four elements per bucket
    uint32_t log2b = 10;
    while ((1ull << log2b) < mr / 4 && log2b < 27) log2b++;
    const uint32_t n_buckets = 1u << log2b;
    // the index: post [m] | run_start [m + 1] | dk [m] | cum [n_buckets] | scalars
    const size_t a4 = (m * 4 + 4 + 255) & ~(size_t)255, a8 = (m * 8 + 255) & ~(size_t)255, ab = (((size_t)n_buckets + 1) * 4 + 255) & ~(size_t)255;

    // A resident reference panel (fpm_dist_set_reference) keeps its packed tiles and its index from one query chunk to the next
    fpm_ctx::RefIndexState& rix = ctx->rix;
    const bool reuse = ctx->ref_set && rix.valid && rix.hashes == (const void*)d_ref->hashes && rix.n_r == d_ref->n && rix.rows_r == rows_r && rix.m == m;
    ctx->time_begin(FPM_KERNEL_DIST_PACK);
    if (!reuse) {
        rix = fpm_ctx::RefIndexState();
        if ((rc = ctx->d_p32.ensure(pr * 4 + 64))) return rc;
        FPM_CUDA(cudaMemsetAsync(ctx->d_p32.p, 0xff, pr * 4, st));
        size_t sort_tmp = 0, scan_tmp = 0, scan2_tmp = 0;
        cub::DoubleBuffer<uint64_t> kb(nullptr, nullptr);
        cub::DoubleBuffer<uint32_t> vb(nullptr, nullptr);
        HeadFlag hf{nullptr};
        cub::CountingInputIterator<uint64_t> cnt(0);
        cub::TransformInputIterator<uint32_t, HeadFlag, cub::CountingInputIterator<uint64_t>> it0(cnt, hf);
        if (m) {
            FPM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, kb, vb, (int64_t)m, 0, 64, st));
            FPM_CUDA(cub::DeviceScan::InclusiveSum(nullptr, scan_tmp, it0, (uint32_t*)nullptr, (int64_t)m, st));
        }
        FPM_CUDA(cub::DeviceScan::InclusiveScan(nullptr, scan2_tmp, (uint32_t*)nullptr, (uint32_t*)nullptr, MaxOp(), (int64_t)n_buckets, st));
        const size_t tmp = std::max(std::max(sort_tmp, scan_tmp), scan2_tmp);
        // one scratch block: keys x2 | dst x2 | cub temp (the rank array reuses the spare key buffer)
        const size_t ka = (m * 8 + 255) & ~(size_t)255, va = (m * 4 + 255) & ~(size_t)255;
        if ((rc = ctx->d_rank.ensure(2 * ka + 2 * va + tmp + 256))) return rc;
        unsigned char* base = ctx->d_rank.as<unsigned char>();
        uint64_t* k0 = (uint64_t*)base; uint64_t* k1 = (uint64_t*)(base + ka);
        uint32_t* v0 = (uint32_t*)(base + 2 * ka); uint32_t* v1 = (uint32_t*)(base + 2 * ka + va);
        void* d_tmp = base + 2 * ka + 2 * va;
        if ((rc = ctx->d_post.ensure(2 * a4 + a8 + ab + 256))) return rc;
        unsigned char* pb = ctx->d_post.as<unsigned char>();
        uint32_t* post = (uint32_t*)pb; uint32_t* run_start = (uint32_t*)(pb + a4); uint64_t* dk = (uint64_t*)(pb + 2 * a4);
        uint32_t* cum = (uint32_t*)(pb + 2 * a4 + a8); uint32_t* scal = (uint32_t*)(pb + 2 * a4 + a8 + ab);
        FPM_CUDA(cudaMemsetAsync(scal, 0, 64, st));
        FPM_CUDA(cudaMemsetAsync(cum, 0, ((size_t)n_buckets + 1) * 4, st));
        FPM_CUDA(cudaMemsetAsync(run_start, 0, 4, st));
        const uint64_t* ks = k0;
        if (m) {
            dist_keys_kernel<<<(uint32_t)((mr + 255) / 256), 256, 0, st>>>(*d_ref, max_size_ref, rows_r, 0u, 0, k0, v0, d_flag_r);
            FPM_CUDA(cudaGetLastError());
            kb = cub::DoubleBuffer<uint64_t>(k0, k1);
            vb = cub::DoubleBuffer<uint32_t>(v0, v1);
            FPM_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, sort_tmp, kb, vb, (int64_t)m, 0, 64, st));
            ks = kb.Current();
            uint32_t* rank = (uint32_t*)kb.Alternate();
            hf.keys = ks;
            cub::TransformInputIterator<uint32_t, HeadFlag, cub::CountingInputIterator<uint64_t>> it(cnt, hf);
            FPM_CUDA(cub::DeviceScan::InclusiveSum(d_tmp, scan_tmp, it, rank, (int64_t)m, st));
            dist_index_kernel<<<(uint32_t)((m + 255) / 256), 256, 0, st>>>(ks, vb.Current(), rank, m, (uint32_t)rows_r, ctx->d_p32.as<uint32_t>(), dk, post, run_start, scal);
            ctx->launches += 2;
        }
        dist_bucket_setup_kernel<<<1, 1, 0, st>>>(scal, log2b);
        if (m) dist_bucket_tails_kernel<<<(uint32_t)((m + 255) / 256), 256, 0, st>>>(dk, scal, cum);
        FPM_CUDA(cub::DeviceScan::InclusiveScan(d_tmp, scan2_tmp, cum, cum, MaxOp(), (int64_t)n_buckets, st));
        ctx->launches += 1 + (m ? 1 : 0);
        FPM_CUDA(cudaGetLastError());
        rix.valid = true; rix.hashes = d_ref->hashes; rix.n_r = d_ref->n; rix.rows_r = rows_r; rix.m = m; rix.pr = pr; rix.ks = ks; rix.n_buckets = n_buckets;
        rix.rank = m ? (const uint32_t*)kb.Alternate() : nullptr;
    }
    unsigned char* pb = ctx->d_post.as<unsigned char>();
    uint32_t* post = (uint32_t*)pb; uint32_t* run_start = (uint32_t*)(pb + a4); uint64_t* dk = (uint64_t*)(pb + 2 * a4);
    uint32_t* cum = (uint32_t*)(pb + 2 * a4 + a8); uint32_t* scal = (uint32_t*)(pb + 2 * a4 + a8 + ab);
    const uint64_t* ks = rix.ks;
    *packed_ref = ctx->d_p32.as<uint32_t>();

    // ---- the query panel: codes by lookup, row-major; the caller packs them (dist_pack_queries / dist_group_panels) ---------
    (void)pq;
    *packed_qry = nullptr;
    if ((rc = ctx->d_codes.ensure(std::max<uint64_t>(mq, 1) * 4 + 64))) return rc;
    ctx->code_stride = max_size_qry;
    RefIndex ix{dk, cum, scal};
    if (mq) {
        // sketch tiles fastest, row tiles slowest: the CTAs in flight at any time work on the same rows (= the same band of the index)
        const dim3 qgrid((uint32_t)((d_qry->n + QC_TILE - 1) / QC_TILE), (uint32_t)((max_size_qry + QC_TILE - 1) / QC_TILE));
        if (qgrid.y > 65535) { set_error("sketches of more than %u hashes are not supported", 65535u * QC_TILE); return FPM_ERR_UNSUPPORTED; }
        dist_qcode_kernel<<<qgrid, 256, 0, st>>>(*d_qry, max_size_qry, ix, run_start, ctx->d_codes.as<uint32_t>(), d_flag_q, d_total);
        ctx->launches++;
        FPM_CUDA(cudaGetLastError());
    }

    // ---- pruning (optional: skipped for huge reference panels or when switched off) ------------------------------
    const uint32_t words = (uint32_t)((d_ref->n + 31) / 32);
    const bool want_prune = !ctx->no_dist_prune && d_ref->n <= 1000000 && mr > 0 && mq > 0;
    uint64_t h2[2] = {0, 0};
    FPM_CUDA(cudaMemcpyAsync(h2, ctx->d_misc.p, 16, cudaMemcpyDeviceToHost, st));
    FPM_CUDA(cudaStreamSynchronize(st));
    if (!reuse) rix.unsorted = (uint32_t)h2[0] != 0;
    const bool bad = rix.unsorted || (uint32_t)(h2[0] >> 32) != 0;
    *mode = bad ? DIST_RANK_UNSORTED : DIST_RANK_OK;           // unsorted input or a hash equal to 2^64-1: the literal kernel defines the result
    if (*mode == DIST_RANK_OK && want_prune) {
        // walk the postings only if that costs well below the merges it can save: a posting is ~4 bytes of L2 traffic and
        // one shared-memory atomic, a pair's merge up to sketch_size steps
        const double postings = (double)h2[1], full = (double)d_ref->n * (double)d_qry->n * (double)std::max<uint32_t>(sketch_size, 1);
        // (and only while the bitmaps stay small next to the panels: n_q x n_r bits, twice when grouped)
        if (postings <= 0.05 * full && (size_t)words * 4 <= 200 * 1024 && (double)d_qry->n * words * 4 <= 8e9) {
            if ((rc = ctx->d_marks.ensure((size_t)d_qry->n * words * 4 + 64))) return rc;
            const uint32_t nodes = (uint32_t)d_ref->n;
            const size_t a_nodes = ((size_t)nodes * 4 + 255) & ~(size_t)255;
            if ((rc = ctx->d_uf.ensure(2 * a_nodes + 64))) return rc;
            uint32_t* parent = ctx->d_uf.as<uint32_t>();
            uint32_t* ref_count = (uint32_t*)(ctx->d_uf.as<unsigned char>() + a_nodes);
            // The components cost a fixed ~0.5 ms (a few long chains of dependent L2 round trips while the trees form, whatever
            // the panel size), a posting costs ~1 ps of marking (measured: 9e8 postings of a 10000 x 5000 block, 1.0 ms without
            // the early stop against 0.11 ms + 0.54 ms with it): the early stop pays from about 5e8 postings on.
            const bool saturate = postings > 4e8 || ctx->force_dist_saturate;
            if (saturate && !rix.uf_valid) {
                dist_uf_init_kernel<<<(nodes + 255) / 256, 256, 0, st>>>(parent, ref_count, nodes);
                dist_uf_union_kernel<<<(uint32_t)((m + 255) / 256), 256, 0, st>>>(ks, post, rix.rank, run_start, m, parent);
                dist_uf_flatten_kernel<<<(nodes + 255) / 256, 256, 0, st>>>(parent, ref_count, nodes);
                ctx->launches += 3;
                rix.uf_valid = true;
            }
            FPM_CUDA(cudaFuncSetAttribute(dist_mark_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(words * 4)));
            dist_mark_kernel<<<(uint32_t)d_qry->n, 256, words * 4, st>>>(ctx->d_codes.as<uint32_t>(), ctx->code_stride, d_qry->sizes, run_start, post, words, saturate ? parent : nullptr,
                                                                         ref_count, (uint32_t)d_ref->n, ctx->d_marks.as<uint32_t>());
            ctx->launches += 1;
            FPM_CUDA(cudaGetLastError());
            *marks = ctx->d_marks.as<uint32_t>();
        }
    }
    ctx->time_end();
    return FPM_OK;
}

}  // namespace fpm
