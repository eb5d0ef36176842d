// capi.cu -- context management, error reporting, pinned memory, INT32 peak microbenchmark.
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <ctype.h>
#include <sched.h>
#include <stdarg.h>
#include <string.h>
#include "common.h"

namespace fpm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line)
{
    const char* base = strrchr(file, '/');
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), base ? base + 1 : file, line, what);
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return FPM_ERR_NO_DEVICE;
    if (e == cudaErrorMemoryAllocation) return FPM_ERR_NOMEM;
    return FPM_ERR_CUDA;
}

int DevBuf::ensure(size_t bytes)
{
    if (bytes <= cap) return FPM_OK;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { p = nullptr; return cuda_fail(e, "cudaMalloc(scratch)", __FILE__, __LINE__); }
    cap = want;
    return FPM_OK;
}

void DevBuf::release()
{
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

// Integer-pipe microbenchmark.  Sixteen independent register chains per thread, instructions pinned with inline
// PTX so ptxas keeps them 1:1.  MODE 0: ALU pipe only (LOP3 / SHF / IADD3-class), MODE 1: FMA pipe only (IMAD),
// MODE 2: strictly alternating.  Every asm statement is one counted op.  (profiles/ubench/int_mix.cu explores the
// patterns: each pipe alone issues 0.48-0.50 warp instructions per clock and SM sub-partition; alternating reaches
// 0.94 with sixteen chains at 32 warps per SM but only 0.78-0.80 with eight -- the first version of this kernel
// used eight and under-reported the dual-pipe peak as 27 T/s instead of 35 T/s.)
constexpr int PEAK_CHAINS = 16;
template <int MODE>
__global__ void __launch_bounds__(256) int32_peak_kernel(uint32_t iters, uint32_t seed, uint32_t* sink)
{
    uint32_t r[PEAK_CHAINS];
#pragma unroll
    for (int c = 0; c < PEAK_CHAINS; c++) r[c] = seed * (2 * c + 1) + threadIdx.x + blockIdx.x;
    const uint32_t k1 = seed | 0x9e3779b1u, k2 = seed ^ 0x7f4a7c15u;
#pragma unroll 1
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int c = 0; c < PEAK_CHAINS; c++) {
                if (MODE == 0 || (MODE == 2 && ((c + u) & 1)))
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[c]) : "r"(k1), "r"(k2));
                else
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[c]) : "r"(k1), "r"(k2));
            }
        }
    }
    uint32_t x = 0;
#pragma unroll
    for (int c = 0; c < PEAK_CHAINS; c++) x ^= r[c];
    if (x == 0x12345u) sink[0] = x;   // practically never; keeps the chains alive
}

}  // namespace fpm

using namespace fpm;

int fpm_ctx::ensure_pinned(size_t bytes)
{
    if (bytes <= h_pinned_cap) return FPM_OK;
    if (h_pinned) cudaFreeHost(h_pinned);
    h_pinned = nullptr; h_pinned_cap = 0;
    FPM_CUDA(cudaHostAlloc(&h_pinned, bytes, cudaHostAllocDefault));
    h_pinned_cap = bytes;
    return FPM_OK;
}

void fpm_ctx::time_begin(int id)
{
    if (!timing) return;
    Timed t;
    t.id = id;
    cudaEventCreate(&t.e0);
    cudaEventCreate(&t.e1);
    cudaEventRecord(t.e0, stream);
    pending.push_back(t);
}

void fpm_ctx::time_end()
{
    if (!timing || pending.empty()) return;
    cudaEventRecord(pending.back().e1, stream);
}

void fpm_ctx::time_resolve()
{
    for (auto& t : pending) {
        float ms = 0;
        if (cudaEventSynchronize(t.e1) == cudaSuccess && cudaEventElapsedTime(&ms, t.e0, t.e1) == cudaSuccess) {
            kernel_ms[t.id] += ms;
            kernel_launches[t.id]++;
        }
        cudaEventDestroy(t.e0);
        cudaEventDestroy(t.e1);
    }
    pending.clear();
}

extern "C" {

int fpm_ctx_set_timing(fpm_ctx* c, int enable)
{
    if (!c) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    c->time_resolve();
    c->timing = enable != 0;
    for (int i = 0; i < 8; i++) { c->kernel_ms[i] = 0; c->kernel_launches[i] = 0; }
    return FPM_OK;
}

int fpm_ctx_set_dist_mode(fpm_ctx* c, int mode)
{
    if (!c) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    if (mode != FPM_DIST_AUTO && mode != FPM_DIST_FORCE64 && mode != FPM_DIST_NO_PRUNE && mode != FPM_DIST_NO_GROUP && mode != FPM_DIST_SATURATE && mode != FPM_DIST_GROUP) { set_error("unknown dist mode %d", mode); return FPM_ERR_ARG; }
    c->force_dist64 = mode == FPM_DIST_FORCE64;
    c->no_dist_prune = mode == FPM_DIST_NO_PRUNE;
    c->no_dist_group = mode == FPM_DIST_NO_GROUP;
    c->force_dist_saturate = mode == FPM_DIST_SATURATE;
    c->force_dist_group = mode == FPM_DIST_GROUP;
    return FPM_OK;
}

int fpm_ctx_get_timing(fpm_ctx* c, int kernel_id, double* out_ms_total, uint64_t* out_launches)
{
    if (!c || kernel_id < 0 || kernel_id >= 8) { set_error("bad argument"); return FPM_ERR_ARG; }
    c->time_resolve();
    if (out_ms_total) *out_ms_total = c->kernel_ms[kernel_id];
    if (out_launches) *out_launches = c->kernel_launches[kernel_id];
    return FPM_OK;
}

int fpm_abi_version(void) { return FPM_ABI_VERSION; }

const char* fpm_last_error(void) { return g_err; }

int fpm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int fpm_ctx_create(int device, fpm_ctx** out)
{
    if (!out) { set_error("out is NULL"); return FPM_ERR_ARG; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s): the fp-mash B200 path has no CPU fallback", e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
        return FPM_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) { set_error("device %d out of range (0..%d)", device, n - 1); return FPM_ERR_ARG; }
    FPM_CUDA(cudaSetDevice(device));
    fpm_ctx* c = new fpm_ctx();
    c->device = device;
    cudaError_t se = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (se != cudaSuccess) { delete c; return cuda_fail(se, "cudaStreamCreate", __FILE__, __LINE__); }
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = c;
    return FPM_OK;
}

void fpm_ctx_destroy(fpm_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    fpm::DevBuf* bufs[] = {&c->seq, &c->goff, &c->thresh, &c->active, &c->toff, &c->tmask, &c->tkeys, &c->tcnt, &c->tpos, &c->maxcnt,
                           &c->maxpos, &c->overflow, &c->stat, &c->tiles, &c->args, &c->alpha, &c->scratch, &c->stream_buf, &c->outh, &c->outc, &c->outn, &c->outk, &c->firstpos,
                           &c->tr_off, &c->tr_cursor, &c->tr_pos, &c->glist, &c->d_ref, &c->d_qry, &c->d_rs, &c->d_qs, &c->d_rl, &c->d_ql,
                           &c->d_out, &c->d_misc, &c->d_p32, &c->d_rank, &c->fa_raw, &c->fa_seq, &c->fa_chunk, &c->fa_recs, &c->d_post, &c->d_marks, &c->d_group, &c->d_hits, &c->d_hsort, &c->d_tiles, &c->d_xq, &c->d_xr, &c->d_xg, &c->d_xg2, &c->d_p32q, &c->d_uf, &c->d_codes, &c->gz_in, &c->gz_meta};
    fpm_comm_destroy(c);
    for (auto* b : bufs) b->release();
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    if (c->copy_stream) { cudaStreamDestroy(c->copy_stream); cudaEventDestroy(c->copy_done[0]); cudaEventDestroy(c->copy_done[1]); }
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int fpm_ctx_sync(fpm_ctx* c)
{
    if (!c) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    FPM_CUDA(cudaStreamSynchronize(c->stream));
    return FPM_OK;
}

void* fpm_ctx_stream(fpm_ctx* c) { return c ? (void*)c->stream : nullptr; }

int fpm_ctx_set_stream(fpm_ctx* c, void* s)
{
    if (!c) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    c->stream = (cudaStream_t)s;
    c->own_stream = false;
    return FPM_OK;
}

uint64_t fpm_ctx_launch_count(const fpm_ctx* c) { return c ? c->launches : 0; }

// CPUs of the NUMA node the current CUDA device hangs off (sysfs); false when that cannot be told.
static bool device_local_cpus(cpu_set_t* set)
{
    int dev = 0;
    char bus[64] = "";
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetPCIBusId(bus, sizeof bus, dev) != cudaSuccess) { cudaGetLastError(); return false; }
    for (char* c = bus; *c; c++) *c = (char)tolower((unsigned char)*c);
    char path[160];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE* f = fopen(path, "r");
    if (!f) return false;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    if (node < 0) return false;
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    f = fopen(path, "r");
    if (!f) return false;
    CPU_ZERO(set);
    int a, b, n = 0;
    while (fscanf(f, "%d", &a) == 1) {                      // "0-31,64-95"
        b = a;
        int ch = fgetc(f);
        if (ch == '-') { if (fscanf(f, "%d", &b) != 1) break; ch = fgetc(f); }
        for (int c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET(c, set); n++; }
        if (ch != ',') break;
    }
    fclose(f);
    return n > 0;
}

// Pinned host memory, placed on the NUMA node of the current device where the platform tells it: the allocation runs
// with the calling thread confined to that node's CPUs (first-touch placement), so that the H2D streams of several
// GPUs of one box do not all cross the socket interconnect.  (This pool's boxes are single-node VMs whose PCI
// devices report numa_node = -1, so it is a no-op there: 8 ranks x 5 GB reach 181 Gk-mers/s end to end.)
int fpm_host_alloc(size_t bytes, void** out)
{
    if (!out) { set_error("out is NULL"); return FPM_ERR_ARG; }
    *out = nullptr;
    cpu_set_t old_set, node_set, use;
    bool bound = false;
    if (sched_getaffinity(0, sizeof old_set, &old_set) == 0 && device_local_cpus(&node_set)) {
        CPU_AND(&use, &old_set, &node_set);
        if (CPU_COUNT(&use) > 0 && !CPU_EQUAL(&use, &old_set)) bound = sched_setaffinity(0, sizeof use, &use) == 0;
    }
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
    if (bound) sched_setaffinity(0, sizeof old_set, &old_set);
    if (e != cudaSuccess) return cuda_fail(e, "cudaHostAlloc", __FILE__, __LINE__);
    return FPM_OK;
}

void fpm_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

int fpm_get_int32_peaks(fpm_ctx* c, double* out3)
{
    if (!c || !out3) { set_error("NULL argument"); return FPM_ERR_ARG; }
    for (int i = 0; i < 3; i++) out3[i] = c->int_peak[i];
    return FPM_OK;
}

int fpm_measure_int32_peak(fpm_ctx* c, double* out_ops_per_s)
{
    if (!c || !out_ops_per_s) { set_error("NULL argument"); return FPM_ERR_ARG; }
    FPM_CUDA(cudaSetDevice(c->device));
    int rc = c->d_misc.ensure(64);
    if (rc) return rc;
    const uint32_t iters = 1 << 11;
    const uint32_t grid = (uint32_t)c->sm_count * 4;          // 32 warps per SM
    cudaEvent_t e0, e1;
    FPM_CUDA(cudaEventCreate(&e0));
    FPM_CUDA(cudaEventCreate(&e1));
    double best = 0;
    for (int mode = 0; mode < 3; mode++) {
        double best_mode = 0;
        for (int rep = 0; rep < 4; rep++) {
            FPM_CUDA(cudaEventRecord(e0, c->stream));
            if (mode == 0) int32_peak_kernel<0><<<grid, 256, 0, c->stream>>>(iters, (uint32_t)rep + 3, c->d_misc.as<uint32_t>());
            else if (mode == 1) int32_peak_kernel<1><<<grid, 256, 0, c->stream>>>(iters, (uint32_t)rep + 3, c->d_misc.as<uint32_t>());
            else int32_peak_kernel<2><<<grid, 256, 0, c->stream>>>(iters, (uint32_t)rep + 3, c->d_misc.as<uint32_t>());
            c->launches++;
            FPM_CUDA(cudaEventRecord(e1, c->stream));
            FPM_CUDA(cudaEventSynchronize(e1));
            float ms = 0;
            FPM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            double ops = (double)grid * 256.0 * iters * 4.0 * PEAK_CHAINS;   // counted instructions per iteration
            if (rep > 0 && ms > 0) best_mode = std::max(best_mode, ops / (ms * 1e-3));
        }
        c->int_peak[mode] = best_mode;
        best = std::max(best, best_mode);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *out_ops_per_s = best;
    return FPM_OK;
}

}  // extern "C"
