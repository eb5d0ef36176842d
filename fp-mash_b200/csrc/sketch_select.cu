// sketch_select.cu -- bottom-s selection, the order-dependent top count, and the -fp line hash.
#include "sketch_kernels.cuh"
#include "sketch_select.h"

namespace fpm {

// ---------------------------------------------------------------------------------------
// Selection: one CTA per sketch.
// ---------------------------------------------------------------------------------------

// One CTA walks a whole table: with one dependent load per iteration that is latency bound (256 iterations x ~0.6 us for a
// 262144-slot table: 0.15 ms per walk, three walks per sketch).  So eight slots are fetched before any is looked at.
template <typename F>
__device__ __forceinline__ void for_each_slot(const uint64_t* __restrict__ tkeys, const uint32_t* __restrict__ tcnt, uint64_t base, uint32_t mask, F&& f)
{
    for (uint32_t i0 = threadIdx.x; i0 <= mask; i0 += blockDim.x * 8) {
        uint64_t k[8];
        uint32_t c[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const uint32_t i = i0 + u * blockDim.x;
            const bool in = i <= mask && i >= i0;                      // (i >= i0: no wrap-around at 2^32)
            k[u] = in ? tkeys[base + i] : SK_EMPTY;
            c[u] = in ? tcnt[base + i] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) f(k[u], c[u]);
    }
}

__global__ void __launch_bounds__(1024) sketch_select_kernel(const SelectArgs a)
{
    extern __shared__ uint64_t s_keys[];
    __shared__ uint32_t s_nq, s_nd;
    const uint32_t g = blockIdx.x;
    if (!a.active[g]) return;
    const uint64_t base = a.toff[g];
    const uint32_t mask = a.tmask[g];
    // pass 1: how many distinct hashes, how many qualify (count >= min_cov)
    if (threadIdx.x == 0) { s_nq = 0; s_nd = 0; }
    __syncthreads();
    {
        uint32_t nd = 0, nq1 = 0;
        for_each_slot(a.tkeys, a.tcnt, base, mask, [&](uint64_t key, uint32_t cnt) {
            if (key != SK_EMPTY) { nd++; nq1 += cnt >= a.min_cov; }
        });
        if (threadIdx.x == 0 && a.maxkey_cnt[g]) { nd++; nq1 += a.maxkey_cnt[g] >= a.min_cov; }
        if (nd) atomicAdd(&s_nd, nd);
        if (nq1) atomicAdd(&s_nq, nq1);
    }
    __syncthreads();
    uint32_t nq = s_nq;
    if (threadIdx.x == 0) { a.stat_nq[g] = nq; a.stat_nd[g] = s_nd; a.stat_topcnt[g] = 0; }
    __syncthreads();
    // Many more qualifying hashes than sketch slots (a read set at 30x under a bound guessed for 32x): the hashes are uniform
    // below the bound, so about 2 s of them lie below cut = bound * 2 s / nq.  If at least s do, the bottom-s is among them and
    // only they are sorted -- in shared memory instead of a one-CTA bitonic sort of all nq keys in global memory.
    uint64_t cut = ~0ULL;
    if (a.thresh && nq > 4 * a.sketch_size && a.thresh[g] != ~0ULL) {
        const uint64_t c = (uint64_t)((double)a.thresh[g] * (2.0 * (double)a.sketch_size / (double)nq));
        if (threadIdx.x == 0) s_nq = 0;
        __syncthreads();
        uint32_t below = 0;
        for_each_slot(a.tkeys, a.tcnt, base, mask, [&](uint64_t key, uint32_t cnt) {
            below += key <= c && cnt >= a.min_cov;                       // (SK_EMPTY is above every cut)
        });
        if (below) atomicAdd(&s_nq, below);
        __syncthreads();
        const uint32_t nb = s_nq;
        __syncthreads();
        if (nb >= a.sketch_size) { cut = c; nq = nb; }
    }
    // keys are sorted in shared memory when they fit, else in this sketch's slice of the global scratch
    // (same capacity as its table, so P = pow2ceil(nq) always fits)
    uint64_t* keys = nq <= a.sort_cap ? s_keys : (a.scratch ? a.scratch + base : nullptr);
    if (!keys) { if (threadIdx.x == 0) a.out_n[g] = 0; return; }
    if (threadIdx.x == 0) s_nq = 0;
    __syncthreads();
    for_each_slot(a.tkeys, a.tcnt, base, mask, [&](uint64_t key, uint32_t cnt) {
        if (key != SK_EMPTY && key <= cut && cnt >= a.min_cov) keys[atomicAdd(&s_nq, 1u)] = key;
    });
    if (threadIdx.x == 0 && cut == ~0ULL && a.maxkey_cnt[g] >= a.min_cov && a.maxkey_cnt[g]) keys[atomicAdd(&s_nq, 1u)] = SK_EMPTY;
    uint32_t P = 1;
    while (P < nq) P <<= 1;
    for (uint32_t i = nq + threadIdx.x; i < P; i += blockDim.x) keys[i] = SK_EMPTY;
    __syncthreads();
    // bitonic sort, ascending
    for (uint32_t k = 2; k <= P; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
                uint32_t ixj = i ^ j;
                if (ixj > i) {
                    uint64_t x = keys[i], y = keys[ixj];
                    bool up = (i & k) == 0;
                    if ((x > y) == up) { keys[i] = y; keys[ixj] = x; }
                }
            }
            __syncthreads();
        }
    }
    const uint32_t n_out = nq < a.sketch_size ? nq : a.sketch_size;
    if (threadIdx.x == 0) a.out_n[g] = n_out;
    for (uint32_t i = threadIdx.x; i < n_out; i += blockDim.x) {
        uint64_t key = keys[i];
        uint64_t o = (uint64_t)g * a.sketch_size + i;
        a.out_hashes[o] = key;
        {
            uint32_t cnt; uint64_t fp;
            if (key == SK_EMPTY) { cnt = a.maxkey_cnt[g]; fp = a.maxkey_pos[g]; }
            else {
                uint32_t slot = table_slot(key, mask);
                while (a.tkeys[base + slot] != key) slot = (slot + 1) & mask;
                cnt = a.tcnt[base + slot]; fp = a.tpos[base + slot];
            }
            if (a.out_counts) a.out_counts[o] = cnt;
            if (i == n_out - 1) a.stat_topcnt[g] = cnt;
            if (a.out_firstpos) a.out_firstpos[o] = fp;
        }
    }
}

// Order-dependent multiplicity of the largest element of a FULL sketch:
// counts(T) = #occurrences of T at stream position <= t*, t* = max over the sketch of the
// position of each hash's min_cov-th occurrence.  One CTA per listed group; positions of
// every occurrence of the sketch's hashes were bucketed by the trace pass.
__global__ void __launch_bounds__(256) sketch_topcount_kernel(const uint32_t* groups, uint32_t sketch_size, uint32_t min_cov,
                                                              const uint64_t* tr_off, const uint32_t* tr_cap,
                                                              const uint64_t* tr_pos, uint32_t* out_counts)
{
    __shared__ unsigned long long s_tstar;
    __shared__ uint32_t s_cnt;
    const uint32_t g = groups[blockIdx.x];
    if (threadIdx.x == 0) { s_tstar = 0; s_cnt = 0; }
    __syncthreads();
    for (uint32_t r = threadIdx.x; r < sketch_size; r += blockDim.x) {
        uint64_t b = (uint64_t)g * sketch_size + r;
        const uint64_t* p = tr_pos + tr_off[b];
        uint32_t n = tr_cap[b];
        // min_cov-th smallest position (min_cov is tiny: repeated minimum extraction)
        uint64_t prev = 0; bool first = true; uint64_t kth = 0;
        uint32_t need = min_cov;
        while (need) {
            uint64_t best = ~0ULL; uint32_t mult = 0;
            for (uint32_t i = 0; i < n; i++) {
                uint64_t v = p[i];
                if (!first && v <= prev) continue;
                if (v < best) { best = v; mult = 1; } else if (v == best) mult++;
            }
            kth = best; prev = best; first = false;
            need = mult >= need ? 0 : need - mult;
        }
        atomicMax(&s_tstar, (unsigned long long)kth);
    }
    __syncthreads();
    const uint64_t tstar = s_tstar;
    uint64_t b = (uint64_t)g * sketch_size + (sketch_size - 1);
    const uint64_t* p = tr_pos + tr_off[b];
    uint32_t n = tr_cap[b], c = 0;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) c += p[i] <= tstar;
    atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) out_counts[b] = s_cnt;
}

// The trace pass without a second pass over the input: the first pass logged every (hash, position) it inserted.
__global__ void __launch_bounds__(256) sketch_trace_log_kernel(const SketchArgs* __restrict__ ga, uint64_t n_log)
{
    const SketchArgs& a = *ga;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_log) return;
    sketch_emit(a, a.log_h[i], a.log_p[i], 0, a.n_groups - 1, 1);
}

void launch_sketch_trace_log(uint64_t n_log, cudaStream_t st, const SketchArgs* d_args)
{
    if (n_log) sketch_trace_log_kernel<<<(uint32_t)((n_log + 255) / 256), 256, 0, st>>>(d_args, n_log);
}

// -fp mode: one thread per fingerprint line.
__global__ void fp_hash_kernel(const uint64_t* tokens, const uint64_t* line_off, uint64_t n_lines, uint32_t seed, int use64, uint64_t* out)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lines) return;
    uint64_t h = murmur3_h1_tokens(tokens + line_off[i], line_off[i + 1] - line_off[i], seed);
    out[i] = use64 ? h : (h & 0xffffffffULL);
}


void launch_sketch_select(uint32_t n_groups, size_t smem_bytes, cudaStream_t st, const SelectArgs& a)
{
    // large tables (big sketch sizes) get a full 1024-thread CTA for the scan and the bitonic passes
    sketch_select_kernel<<<n_groups, smem_bytes > 32768 ? 1024 : 256, smem_bytes, st>>>(a);
}

int configure_sketch_select(size_t max_smem_bytes)
{
    return (int)cudaFuncSetAttribute(sketch_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem_bytes);
}

void launch_sketch_topcount(uint32_t n_list, cudaStream_t st, const uint32_t* d_groups, uint32_t sketch_size, uint32_t min_cov,
                            const uint64_t* tr_off, const uint32_t* tr_cap, const uint64_t* tr_pos, uint32_t* out_counts)
{
    sketch_topcount_kernel<<<n_list, 256, 0, st>>>(d_groups, sketch_size, min_cov, tr_off, tr_cap, tr_pos, out_counts);
}

void launch_fp_hash(uint64_t n_lines, cudaStream_t st, const uint64_t* tokens, const uint64_t* line_off, uint32_t seed, int use64, uint64_t* out)
{
    uint32_t grid = (uint32_t)((n_lines + 255) / 256);
    fp_hash_kernel<<<grid, 256, 0, st>>>(tokens, line_off, n_lines, seed, use64, out);
}

}  // namespace fpm
