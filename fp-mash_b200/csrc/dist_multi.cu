// dist_multi.cu -- the hot path on several GPUs of one box.
//
// The reference fans its work out itself: `mash dist` cuts the query x reference pair space into chunks for its thread
// pool (CommandDistance.cpp:224-261), `mash sketch` hands whole files to pool threads (Sketch.cpp:353-355).  Here the
// same two cuts are made across GPUs, inside the library:
//
//   * dist: the pair space is cut into a q_parts x r_parts GRID OF BLOCKS, one block per GPU (8 GPUs: 2 x 4).  A GPU
//     then rank-compresses, indexes and prunes only its query block and its reference block -- (1/q_parts + 1/r_parts)
//     of the two panels instead of all references plus its query rows -- so the pre-pass shrinks with the GPU count
//     like the merges do.  (Round 1 sharded query rows only, in bench.py: every rank sorted the whole reference panel,
//     and 8 GPUs gave 3.1x.)
//   * sketch: whole sketches (groups) go to GPUs in contiguous, byte-balanced ranges; no exchange at all.
//
// Two deployment shapes, same grid:
//   - ONE PROCESS PER GPU (fpm_comm_*, fpm_dist_sharded_dev): every rank holds 1/world of the rows of both panels in
//     HBM; the one exchange step of the path sends every row shard to exactly the ranks whose block contains it
//     (grouped ncclSend/ncclRecv over NVLink: a rank receives n_q/q_parts + n_r/r_parts rows, not the whole panels);
//     results stay on the rank as a dense block.  The communicator is handed in (fpm_comm_adopt) or created from a
//     unique id the host distributes (fpm_comm_init_rank).  NCCL is loaded with dlopen at that moment: the library
//     has no link-time dependency on it, and inside a process that already loaded a libnccl (torch) that one is used.
//   - ONE PROCESS, ALL GPUS (fpm_multi_*): what `mash dist` / `mash sketch` use.  Panels are in host memory, so every
//     GPU simply uploads its two blocks: no collective.  One host thread per GPU.
#include <dlfcn.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>
#include "common.h"
#include "dist_internal.h"
#include "nccl_dyn.h"

namespace fpm {

// ---- NCCL, bound at run time (nccl_dyn.h) ------------------------------------------------------------------------
NcclApi* nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.handle ? &api : nullptr;
    tried = true;
    const char* names[] = {getenv("FPMASH_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n || !*n) continue;
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) { set_error("NCCL is not available: %s", dlerror()); return nullptr; }
    bool ok = true;
#define FPM_NCCL_SYM(field, sym) \
    do { *(void**)(&api.field) = dlsym(api.handle, sym); if (!api.field) ok = false; } while (0)
    FPM_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    FPM_NCCL_SYM(CommInitRank, "ncclCommInitRank");
    FPM_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    FPM_NCCL_SYM(GroupStart, "ncclGroupStart");
    FPM_NCCL_SYM(GroupEnd, "ncclGroupEnd");
    FPM_NCCL_SYM(Send, "ncclSend");
    FPM_NCCL_SYM(Recv, "ncclRecv");
    FPM_NCCL_SYM(AllGather, "ncclAllGather");
    FPM_NCCL_SYM(AllReduce, "ncclAllReduce");
    FPM_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef FPM_NCCL_SYM
    if (!ok) { set_error("the NCCL library lacks a required symbol"); dlclose(api.handle); api.handle = nullptr; return nullptr; }
    return &api;
}

// ---- the grid ---------------------------------------------------------------------------------------------------
static uint64_t shard_begin(uint64_t n, uint64_t part, uint64_t parts) { return (uint64_t)((unsigned __int128)n * part / parts); }

// q_parts x r_parts = world minimising the rows a GPU has to rank-compress and index, n_q / q_parts + n_r / r_parts
static void grid_shape(int world, uint64_t n_q, uint64_t n_r, int* qp, int* rp)
{
    double best = -1;
    *qp = 1; *rp = world;
    for (int q = 1; q <= world; q++) {
        if (world % q) continue;
        const int r = world / q;
        const double cost = (double)n_q / q + (double)n_r / r;
        if (best < 0 || cost < best * (1 - 1e-12)) { best = cost; *qp = q; *rp = r; }
    }
}

struct Block { uint64_t q0, q1, r0, r1; int qi, rj, qp, rp; };

static Block block_of(int rank, int world, uint64_t n_q, uint64_t n_r)
{
    Block b;
    grid_shape(world, n_q, n_r, &b.qp, &b.rp);
    b.qi = rank / b.rp; b.rj = rank % b.rp;
    // a query block is the union of the rp row shards of its grid row, a reference block of qp consecutive row shards
    b.q0 = shard_begin(n_q, (uint64_t)b.qi * b.rp, world); b.q1 = shard_begin(n_q, (uint64_t)(b.qi + 1) * b.rp, world);
    b.r0 = shard_begin(n_r, (uint64_t)b.rj * b.qp, world); b.r1 = shard_begin(n_r, (uint64_t)(b.rj + 1) * b.qp, world);
    return b;
}

// block panel in one scratch buffer: hashes [rows][stride] | lengths [rows] | sizes [rows]
static int block_panel(DevBuf& buf, uint64_t rows, uint64_t stride, fpm_panel* out)
{
    const size_t hb = rows * stride * 8;
    int rc = buf.ensure(hb + rows * 12 + 64);
    if (rc) return rc;
    unsigned char* b = buf.as<unsigned char>();
    out->hashes = (const uint64_t*)b;
    out->lengths = (const uint64_t*)(b + hb);
    out->sizes = (const uint32_t*)(b + hb + rows * 8);
    out->n = rows;
    out->stride = stride;
    return FPM_OK;
}

// The exchange step for panels of moderate size: ONE all-gather per panel (NCCL's best-tuned collective: rings over all
// NVLink channels / NVLS through the switch), then every rank keeps the shards of its block.  Measured at 8 GPUs on
// configs[2] (2 x 20 000 x 8 KB rows): 0.57 ms for the grouped send/recv below against the all-gather's time in
// profiles/r02_bench_8gpu_*.json.  A panel used in both roles (all-vs-all) is gathered once.
static int exchange_blocks_allgather(fpm_ctx* ctx, NcclApi* api, const fpm_panel* d_ref_shard, uint64_t n_ref, const fpm_panel* d_qry_shard, uint64_t n_qry,
                                     const Block& b, fpm_panel* blk_ref, fpm_panel* blk_qry)
{
    const int W = ctx->comm_world, me = ctx->comm_rank;
    ncclComm_t comm = (ncclComm_t)ctx->comm;
    cudaStream_t st = ctx->stream;
    const bool same = d_qry_shard->hashes == d_ref_shard->hashes && d_qry_shard->sizes == d_ref_shard->sizes && d_qry_shard->lengths == d_ref_shard->lengths &&
                      n_qry == n_ref && d_qry_shard->stride == d_ref_shard->stride;
    int rc;
    ctx->time_begin(FPM_KERNEL_DIST_EXCHANGE);
    for (int role = 0; role < 2; role++) {                                // 0: query panel, 1: reference panel
        const fpm_panel* mine = role == 0 ? d_qry_shard : d_ref_shard;
        const uint64_t n = role == 0 ? n_qry : n_ref;
        const uint64_t my0 = shard_begin(n, me, W), my1 = shard_begin(n, me + 1, W);
        if (mine->n != my1 - my0) { set_error("rank %d holds %llu %s rows, its shard of %llu rows over %d ranks has %llu", me, (unsigned long long)mine->n,
                                              role ? "reference" : "query", (unsigned long long)n, W, (unsigned long long)(my1 - my0)); return FPM_ERR_ARG; }
        const uint64_t max_rows = (n + W - 1) / W, stride = mine->stride;
        const size_t hb = max_rows * stride * 8, lb = max_rows * 8, sb = max_rows * 4;
        DevBuf& gbuf = (role == 1 && !same) ? ctx->d_xg2 : ctx->d_xg;      // (a block may keep pointing into its panel's gather buffer)
        if (!(role == 1 && same)) {
            if ((rc = gbuf.ensure((size_t)W * (hb + lb + sb) + 256))) return rc;
            unsigned char* g = gbuf.as<unsigned char>();
            unsigned char* gh = g; unsigned char* gl = g + (size_t)W * hb; unsigned char* gs = gl + (size_t)W * lb;
            if (my1 > my0) {
                FPM_CUDA(cudaMemcpyAsync(gh + (size_t)me * hb, mine->hashes, (my1 - my0) * stride * 8, cudaMemcpyDeviceToDevice, st));
                FPM_CUDA(cudaMemcpyAsync(gl + (size_t)me * lb, mine->lengths, (my1 - my0) * 8, cudaMemcpyDeviceToDevice, st));
                FPM_CUDA(cudaMemcpyAsync(gs + (size_t)me * sb, mine->sizes, (my1 - my0) * 4, cudaMemcpyDeviceToDevice, st));
            }
            if (hb) FPM_NCCL(api, api->AllGather(gh + (size_t)me * hb, gh, hb, ncclUint8, comm, st));
            if (lb) FPM_NCCL(api, api->AllGather(gl + (size_t)me * lb, gl, lb, ncclUint8, comm, st));
            if (sb) FPM_NCCL(api, api->AllGather(gs + (size_t)me * sb, gs, sb, ncclUint8, comm, st));
        }
        unsigned char* g = gbuf.as<unsigned char>();
        unsigned char* gh = g; unsigned char* gl = g + (size_t)W * hb; unsigned char* gs = gl + (size_t)W * lb;
        // keep the shards of this rank's block
        for (int keep = (role == 1 && same) ? 1 : role; keep <= ((role == 0 && same) ? 0 : role); keep++) {
            fpm_panel* blk = keep == 0 ? blk_qry : blk_ref;
            const int first = keep == 0 ? b.qi * b.rp : b.rj * b.qp, cnt = keep == 0 ? b.rp : b.qp;
            const uint64_t blk0 = keep == 0 ? b.q0 : b.r0;
            if (n % W == 0) {
                // equal shards: consecutive slots of the gather buffer ARE the block, nothing to copy
                blk->hashes = (const uint64_t*)(gh + (size_t)first * hb);
                blk->lengths = (const uint64_t*)(gl + (size_t)first * lb);
                blk->sizes = (const uint32_t*)(gs + (size_t)first * sb);
                continue;
            }
            for (int sh = first; sh < first + cnt; sh++) {
                const uint64_t s0 = shard_begin(n, sh, W), s1 = shard_begin(n, sh + 1, W);
                if (s1 == s0) continue;
                FPM_CUDA(cudaMemcpyAsync((void*)(blk->hashes + (s0 - blk0) * stride), gh + (size_t)sh * hb, (s1 - s0) * stride * 8, cudaMemcpyDeviceToDevice, st));
                FPM_CUDA(cudaMemcpyAsync((void*)(blk->lengths + (s0 - blk0)), gl + (size_t)sh * lb, (s1 - s0) * 8, cudaMemcpyDeviceToDevice, st));
                FPM_CUDA(cudaMemcpyAsync((void*)(blk->sizes + (s0 - blk0)), gs + (size_t)sh * sb, (s1 - s0) * 4, cudaMemcpyDeviceToDevice, st));
            }
        }
    }
    ctx->time_end();
    return FPM_OK;
}

// The exchange step: every rank's row shard of a panel goes to the ranks whose block holds it.
static int exchange_blocks(fpm_ctx* ctx, const fpm_panel* d_ref_shard, uint64_t n_ref, const fpm_panel* d_qry_shard, uint64_t n_qry, const Block& b,
                           fpm_panel* blk_ref, fpm_panel* blk_qry)
{
    NcclApi* api = nccl_api();
    if (!api) return FPM_ERR_COMM;
    const int world = ctx->comm_world, me = ctx->comm_rank;
    ncclComm_t comm = (ncclComm_t)ctx->comm;
    cudaStream_t st = ctx->stream;
    int rc;
    if ((rc = block_panel(ctx->d_xq, b.q1 - b.q0, d_qry_shard->stride, blk_qry))) return rc;
    if ((rc = block_panel(ctx->d_xr, b.r1 - b.r0, d_ref_shard->stride, blk_ref))) return rc;
    // panels that fit a few GB whole: all-gather, keep the block; larger ones (the 8 GB query panel of configs[4]): only the
    // shards of the block travel, by grouped send / receive
    const uint64_t kAllGatherBytes = 4ull << 30;
    const char* xenv = getenv("FPMASH_EXCHANGE");                            // tests: "p2p" / "allgather" pin the method
    const bool small = n_qry * d_qry_shard->stride * 8 <= kAllGatherBytes && n_ref * d_ref_shard->stride * 8 <= kAllGatherBytes;
    if ((small && !(xenv && xenv[0] == 'p')) || (xenv && xenv[0] == 'a'))
        return exchange_blocks_allgather(ctx, api, d_ref_shard, n_ref, d_qry_shard, n_qry, b, blk_ref, blk_qry);

    struct Msg { const void* src; void* dst; size_t bytes; };
    auto three = [](const fpm_panel* from, uint64_t from_row, const fpm_panel* to, uint64_t to_row, uint64_t rows, Msg* m) {
        m[0] = {from ? from->hashes + from_row * from->stride : nullptr, to ? (void*)(to->hashes + to_row * to->stride) : nullptr, rows * (from ? from->stride : to->stride) * 8};
        m[1] = {from ? from->lengths + from_row : nullptr, to ? (void*)(to->lengths + to_row) : nullptr, rows * 8};
        m[2] = {from ? from->sizes + from_row : nullptr, to ? (void*)(to->sizes + to_row) : nullptr, rows * 4};
    };
    ctx->time_begin(FPM_KERNEL_DIST_EXCHANGE);
    FPM_NCCL(api, api->GroupStart());
    int fail = FPM_OK;
    for (int role = 0; role < 2 && !fail; role++) {                       // 0: query panel, 1: reference panel
        const fpm_panel* mine = role == 0 ? d_qry_shard : d_ref_shard;
        const fpm_panel* blk = role == 0 ? blk_qry : blk_ref;
        const uint64_t n = role == 0 ? n_qry : n_ref;
        const uint64_t my0 = shard_begin(n, me, world), my1 = shard_begin(n, me + 1, world);
        if (mine->n != my1 - my0) { set_error("rank %d holds %llu %s rows, its shard of %llu rows over %d ranks has %llu", me, (unsigned long long)mine->n,
                                              role ? "reference" : "query", (unsigned long long)n, world, (unsigned long long)(my1 - my0)); fail = FPM_ERR_ARG; break; }
        // sends: my shard belongs to query block me / rp (wanted by that grid row) or reference block me / qp (wanted by that grid column)
        const int n_dst = role == 0 ? b.rp : b.qp;
        for (int d = 0; d < n_dst && !fail; d++) {
            const int dst = role == 0 ? (me / b.rp) * b.rp + d : d * b.rp + me / b.qp;
            if (dst == me || my1 == my0) continue;
            Msg m[3];
            three(mine, 0, nullptr, 0, my1 - my0, m);
            for (int i = 0; i < 3; i++)
                if (api->Send(m[i].src, m[i].bytes, ncclUint8, dst, comm, st) != ncclSuccess) { set_error("ncclSend failed"); fail = FPM_ERR_COMM; }
        }
        // receives: the shards that make up my block
        const int first = role == 0 ? b.qi * b.rp : b.rj * b.qp, cnt = role == 0 ? b.rp : b.qp;
        const uint64_t blk0 = role == 0 ? b.q0 : b.r0;
        for (int s = first; s < first + cnt && !fail; s++) {
            const uint64_t s0 = shard_begin(n, s, world), s1 = shard_begin(n, s + 1, world);
            if (s1 == s0) continue;
            Msg m[3];
            if (s == me) {
                three(mine, 0, blk, s0 - blk0, s1 - s0, m);
                for (int i = 0; i < 3; i++)
                    if (cudaMemcpyAsync(m[i].dst, m[i].src, m[i].bytes, cudaMemcpyDeviceToDevice, st) != cudaSuccess) { set_error("device copy of the own shard failed"); fail = FPM_ERR_CUDA; }
                continue;
            }
            three(nullptr, 0, blk, s0 - blk0, s1 - s0, m);
            for (int i = 0; i < 3; i++)
                if (api->Recv(m[i].dst, m[i].bytes, ncclUint8, s, comm, st) != ncclSuccess) { set_error("ncclRecv failed"); fail = FPM_ERR_COMM; }
        }
    }
    ncclResult_t ge = api->GroupEnd();
    ctx->time_end();
    if (fail) return fail;
    if (ge != ncclSuccess) { set_error("NCCL error %d (%s): ncclGroupEnd", (int)ge, api->GetErrorString(ge)); return FPM_ERR_COMM; }
    ctx->launches += 0;   // (NCCL's kernels are not ours)
    return FPM_OK;
}

static int check_sharded(fpm_ctx* ctx, const fpm_dist_params* p, const fpm_panel* d_ref_shard, const fpm_panel* d_qry_shard, fpm_block* blk)
{
    if (!ctx) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    if (!blk) { set_error("NULL argument"); return FPM_ERR_ARG; }
    int rc = check_dist(p, d_ref_shard, d_qry_shard);
    if (rc) return rc;
    if (!ctx->comm || ctx->comm_world < 1) { set_error("no communicator: call fpm_comm_init_rank or fpm_comm_adopt first"); return FPM_ERR_ARG; }
    return FPM_OK;
}

}  // namespace fpm

using namespace fpm;

// ---- single process, several GPUs -------------------------------------------------------------------------------
struct fpm_multi {
    std::vector<fpm_ctx*> ctx;
};

namespace fpm {

// run fn(i) on one host thread per GPU; the first failure's status and message are handed to the calling thread
template <typename F>
static int on_all(fpm_multi* m, int n_used, F fn)
{
    std::vector<int> rc(n_used, FPM_OK);
    std::vector<std::string> msg(n_used);
    std::vector<std::thread> th;
    for (int i = 1; i < n_used; i++)
        th.emplace_back([&, i]() { rc[i] = fn(i); if (rc[i]) msg[i] = fpm_last_error(); });
    rc[0] = fn(0);
    if (rc[0]) msg[0] = fpm_last_error();
    for (auto& t : th) t.join();
    for (int i = 0; i < n_used; i++)
        if (rc[i] && rc[i] != FPM_ERR_CAPACITY) { set_error("GPU %d: %s", m->ctx[i]->device, msg[i].c_str()); return rc[i]; }
    for (int i = 0; i < n_used; i++)
        if (rc[i]) { set_error("GPU %d: %s", m->ctx[i]->device, msg[i].c_str()); return rc[i]; }
    return FPM_OK;
}

static fpm_panel sub_panel(const fpm_panel* p, uint64_t r0, uint64_t r1)
{
    fpm_panel s = *p;
    s.hashes = p->hashes + r0 * p->stride;
    s.sizes = p->sizes + r0;
    s.lengths = p->lengths + r0;
    s.n = r1 - r0;
    return s;
}

// GPUs that get a block: no more than there are rows to cut
static int gpus_for(const fpm_multi* m, uint64_t n_q, uint64_t n_r)
{
    int n = (int)m->ctx.size();
    while (n > 1) {
        int qp, rp;
        grid_shape(n, n_q, n_r, &qp, &rp);
        if ((uint64_t)qp <= n_q && (uint64_t)rp <= n_r && n_q * n_r >= (uint64_t)n * 4096) break;   // tiny jobs: one GPU (start-up dominates)
        n--;
    }
    return n;
}

int dist_tile_host_ld(fpm_ctx* ctx, const fpm_dist_params* p, const fpm_panel* ref, const fpm_panel* qry, fpm_pair* out, uint64_t ld);   // dist.cu
int dist_hits_host(fpm_ctx* ctx, const fpm_dist_params* p, const fpm_panel* ref, const fpm_panel* qry, fpm_hit* out, uint64_t capacity, uint64_t* n_hits,
                   uint32_t q_base, uint32_t r_base);   // dist.cu

}  // namespace fpm

extern "C" {

// ---- communicator -----------------------------------------------------------------------------------------------
int fpm_comm_get_unique_id(void* out_id)
{
    if (!out_id) { set_error("NULL argument"); return FPM_ERR_ARG; }
    NcclApi* api = nccl_api();
    if (!api) return FPM_ERR_COMM;
    static_assert(sizeof(ncclUniqueId) == FPM_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    FPM_NCCL(api, api->GetUniqueId(&id));
    memcpy(out_id, &id, sizeof id);
    return FPM_OK;
}

int fpm_comm_init_rank(fpm_ctx* ctx, const void* id_bytes, int rank, int world)
{
    if (!ctx || !id_bytes) { set_error("NULL argument"); return FPM_ERR_ARG; }
    if (world < 1 || rank < 0 || rank >= world) { set_error("rank %d of %d", rank, world); return FPM_ERR_ARG; }
    NcclApi* api = nccl_api();
    if (!api) return FPM_ERR_COMM;
    fpm_comm_destroy(ctx);
    FPM_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof id);
    ncclComm_t comm = nullptr;
    FPM_NCCL(api, api->CommInitRank(&comm, world, id, rank));
    ctx->comm = comm; ctx->comm_rank = rank; ctx->comm_world = world; ctx->comm_owned = true;
    return FPM_OK;
}

int fpm_comm_adopt(fpm_ctx* ctx, void* nccl_comm, int rank, int world)
{
    if (!ctx || !nccl_comm) { set_error("NULL argument"); return FPM_ERR_ARG; }
    if (world < 1 || rank < 0 || rank >= world) { set_error("rank %d of %d", rank, world); return FPM_ERR_ARG; }
    if (!nccl_api()) return FPM_ERR_COMM;
    fpm_comm_destroy(ctx);
    ctx->comm = nccl_comm; ctx->comm_rank = rank; ctx->comm_world = world; ctx->comm_owned = false;
    return FPM_OK;
}

int fpm_comm_destroy(fpm_ctx* ctx)
{
    if (!ctx) { set_error("ctx is NULL"); return FPM_ERR_ARG; }
    if (ctx->comm && ctx->comm_owned) {
        NcclApi* api = nccl_api();
        if (api) { cudaSetDevice(ctx->device); api->CommDestroy((ncclComm_t)ctx->comm); }
    }
    ctx->comm = nullptr; ctx->comm_rank = 0; ctx->comm_world = 0; ctx->comm_owned = false;
    return FPM_OK;
}

int fpm_comm_rank(const fpm_ctx* ctx) { return ctx && ctx->comm ? ctx->comm_rank : -1; }
int fpm_comm_size(const fpm_ctx* ctx) { return ctx && ctx->comm ? ctx->comm_world : 0; }

// ---- grid -------------------------------------------------------------------------------------------------------
void fpm_shard_range(uint64_t n, int part, int parts, uint64_t* begin, uint64_t* end)
{
    if (parts < 1) parts = 1;
    if (begin) *begin = shard_begin(n, (uint64_t)part, (uint64_t)parts);
    if (end) *end = shard_begin(n, (uint64_t)part + 1, (uint64_t)parts);
}

int fpm_dist_grid_shape(int world, uint64_t n_qry, uint64_t n_ref, int* q_parts, int* r_parts)
{
    if (world < 1 || !q_parts || !r_parts) { set_error("bad argument"); return FPM_ERR_ARG; }
    grid_shape(world, n_qry, n_ref, q_parts, r_parts);
    return FPM_OK;
}

int fpm_dist_block(int rank, int world, uint64_t n_qry, uint64_t n_ref, fpm_block* out)
{
    if (world < 1 || rank < 0 || rank >= world || !out) { set_error("bad argument"); return FPM_ERR_ARG; }
    const Block b = block_of(rank, world, n_qry, n_ref);
    out->q_begin = b.q0; out->q_end = b.q1; out->r_begin = b.r0; out->r_end = b.r1;
    return FPM_OK;
}

// ---- one process per GPU ----------------------------------------------------------------------------------------
int fpm_dist_sharded_dev(fpm_ctx* ctx, const fpm_dist_params* p, const fpm_panel* d_ref_shard, uint64_t n_ref_total, const fpm_panel* d_qry_shard,
                         uint64_t n_qry_total, fpm_pair* d_out_block, uint64_t out_capacity, fpm_block* blk, uint64_t* d_merge_steps)
{
    int rc = check_sharded(ctx, p, d_ref_shard, d_qry_shard, blk);
    if (rc) return rc;
    FPM_CUDA(cudaSetDevice(ctx->device));
    const Block b = block_of(ctx->comm_rank, ctx->comm_world, n_qry_total, n_ref_total);
    blk->q_begin = b.q0; blk->q_end = b.q1; blk->r_begin = b.r0; blk->r_end = b.r1;
    if ((b.q1 - b.q0) * (b.r1 - b.r0) > out_capacity) { set_error("the block has %llu pairs, room for %llu", (unsigned long long)((b.q1 - b.q0) * (b.r1 - b.r0)), (unsigned long long)out_capacity); return FPM_ERR_CAPACITY; }
    fpm_panel br, bq;
    if ((rc = exchange_blocks(ctx, d_ref_shard, n_ref_total, d_qry_shard, n_qry_total, b, &br, &bq))) return rc;
    if (br.n == 0 || bq.n == 0) return FPM_OK;
    uint32_t mr = 0, mq = 0;
    if ((rc = max_size_dev(ctx, &br, &mr))) return rc;
    if ((rc = max_size_dev(ctx, &bq, &mq))) return rc;
    if (mr > br.stride || mq > bq.stride) { set_error("a sketch size exceeds the panel stride"); return FPM_ERR_ARG; }
    return run_dist(ctx, p, &br, &bq, d_out_block, d_merge_steps, mr, mq);
}

int fpm_dist_hits_sharded_dev(fpm_ctx* ctx, const fpm_dist_params* p, const fpm_panel* d_ref_shard, uint64_t n_ref_total, const fpm_panel* d_qry_shard,
                              uint64_t n_qry_total, fpm_hit* d_out, uint64_t capacity, uint64_t* n_hits, fpm_block* blk, uint64_t* d_merge_steps)
{
    int rc = check_sharded(ctx, p, d_ref_shard, d_qry_shard, blk);
    if (rc) return rc;
    if (!n_hits || (!d_out && capacity)) { set_error("NULL argument"); return FPM_ERR_ARG; }
    if (n_ref_total > 0xfffffffeull || n_qry_total > 0xfffffffeull) { set_error("panel too large for 32-bit hit indices"); return FPM_ERR_ARG; }
    FPM_CUDA(cudaSetDevice(ctx->device));
    const Block b = block_of(ctx->comm_rank, ctx->comm_world, n_qry_total, n_ref_total);
    blk->q_begin = b.q0; blk->q_end = b.q1; blk->r_begin = b.r0; blk->r_end = b.r1;
    *n_hits = 0;
    fpm_panel br, bq;
    if ((rc = exchange_blocks(ctx, d_ref_shard, n_ref_total, d_qry_shard, n_qry_total, b, &br, &bq))) return rc;
    if (br.n == 0 || bq.n == 0) return FPM_OK;
    uint32_t mr = 0, mq = 0;
    if ((rc = max_size_dev(ctx, &br, &mr))) return rc;
    if ((rc = max_size_dev(ctx, &bq, &mq))) return rc;
    if (mr > br.stride || mq > bq.stride) { set_error("a sketch size exceeds the panel stride"); return FPM_ERR_ARG; }
    if ((rc = dist_hits_run(ctx, p, &br, &bq, mr, mq, d_out, capacity, n_hits, d_merge_steps, (uint32_t)b.q0, (uint32_t)b.r0))) return rc;
    FPM_CUDA(cudaStreamSynchronize(ctx->stream));
    return FPM_OK;
}

// ---- one process, all GPUs --------------------------------------------------------------------------------------
int fpm_multi_create(const int* devices, int n_devices, fpm_multi** out)
{
    if (!out) { set_error("out is NULL"); return FPM_ERR_ARG; }
    *out = nullptr;
    const int have = fpm_device_count();
    if (have == 0) { set_error("no CUDA device available: the fp-mash B200 path has no CPU fallback"); return FPM_ERR_NO_DEVICE; }
    if (n_devices <= 0) n_devices = have;
    fpm_multi* m = new fpm_multi();
    for (int i = 0; i < n_devices; i++) {
        fpm_ctx* c = nullptr;
        const int rc = fpm_ctx_create(devices ? devices[i] : i, &c);
        if (rc) { fpm_multi_destroy(m); return rc; }
        m->ctx.push_back(c);
    }
    *out = m;
    return FPM_OK;
}

void fpm_multi_destroy(fpm_multi* m)
{
    if (!m) return;
    for (fpm_ctx* c : m->ctx) fpm_ctx_destroy(c);
    delete m;
}

int fpm_multi_size(const fpm_multi* m) { return m ? (int)m->ctx.size() : 0; }
fpm_ctx* fpm_multi_ctx(fpm_multi* m, int i) { return m && i >= 0 && i < (int)m->ctx.size() ? m->ctx[i] : nullptr; }

int fpm_dist_tile_multi(fpm_multi* m, const fpm_dist_params* p, const fpm_panel* ref, const fpm_panel* qry, fpm_pair* out)
{
    if (!m || m->ctx.empty()) { set_error("no GPUs"); return FPM_ERR_ARG; }
    int rc = check_dist(p, ref, qry);
    if (rc) return rc;
    if (ref->n == 0 || qry->n == 0) return FPM_OK;
    const int n = gpus_for(m, qry->n, ref->n);
    return on_all(m, n, [&](int i) -> int {
        const Block b = block_of(i, n, qry->n, ref->n);
        if (b.q1 == b.q0 || b.r1 == b.r0) return FPM_OK;
        const fpm_panel sr = sub_panel(ref, b.r0, b.r1), sq = sub_panel(qry, b.q0, b.q1);
        return dist_tile_host_ld(m->ctx[i], p, &sr, &sq, out + b.q0 * ref->n + b.r0, ref->n);
    });
}

int fpm_dist_hits_multi(fpm_multi* m, const fpm_dist_params* p, const fpm_panel* ref, const fpm_panel* qry, fpm_hit* out, uint64_t capacity, uint64_t* n_hits)
{
    if (!m || m->ctx.empty()) { set_error("no GPUs"); return FPM_ERR_ARG; }
    if (!n_hits || (!out && capacity)) { set_error("NULL argument"); return FPM_ERR_ARG; }
    int rc = check_dist(p, ref, qry);
    if (rc) return rc;
    *n_hits = 0;
    if (ref->n == 0 || qry->n == 0) return FPM_OK;
    if (ref->n > 0xfffffffeull || qry->n > 0xfffffffeull) { set_error("panel too large for 32-bit hit indices"); return FPM_ERR_ARG; }
    const int n = gpus_for(m, qry->n, ref->n);
    std::vector<std::vector<fpm_hit>> part(n);
    std::vector<uint64_t> got(n, 0);
    std::vector<Block> blocks(n);
    rc = on_all(m, n, [&](int i) -> int {
        const Block b = blocks[i] = block_of(i, n, qry->n, ref->n);
        if (b.q1 == b.q0 || b.r1 == b.r0) return FPM_OK;
        const fpm_panel sr = sub_panel(ref, b.r0, b.r1), sq = sub_panel(qry, b.q0, b.q1);
        uint64_t cap = std::max<uint64_t>(4096, std::min<uint64_t>(2 * capacity / n + 4096, (b.q1 - b.q0) * (b.r1 - b.r0)));
        for (int attempt = 0; attempt < 3; attempt++) {
            part[i].resize(cap);
            uint64_t k = 0;
            const int r = dist_hits_host(m->ctx[i], p, &sr, &sq, part[i].data(), cap, &k, (uint32_t)b.q0, (uint32_t)b.r0);
            got[i] = k;
            if (r != FPM_ERR_CAPACITY) return r;
            cap = k;                                                  // the call reported how many pairs pass
        }
        return FPM_ERR_CAPACITY;
    });
    if (rc) return rc;
    uint64_t total = 0;
    for (int i = 0; i < n; i++) total += got[i];
    *n_hits = total;
    if (total > capacity) { set_error("fpm_dist_hits_multi: %llu pairs pass the filters, room for %llu", (unsigned long long)total, (unsigned long long)capacity); return FPM_ERR_CAPACITY; }
    // the reference's output order is query-major (CommandDistance.cpp:355-359): per grid row, the blocks' lists -- each
    // sorted by (query, ref), reference ranges ascending with the block column -- are interleaved query by query
    const int rp = blocks[0].rp, qp = blocks[0].qp;
    uint64_t w = 0;
    for (int qi = 0; qi < qp; qi++) {
        std::vector<uint64_t> at(rp, 0);
        const Block& row = blocks[qi * rp];
        for (uint64_t q = row.q0; q < row.q1; q++)
            for (int rj = 0; rj < rp; rj++) {
                const int i = qi * rp + rj;
                const fpm_hit* h = part[i].data();
                uint64_t a = at[rj];
                while (a < got[i] && h[a].query == q) out[w++] = h[a++];
                at[rj] = a;
            }
    }
    return FPM_OK;
}

int fpm_sketch_batch_multi(fpm_multi* m, const fpm_sketch_params* p, const uint8_t* seq, uint64_t seq_bytes, const uint64_t* group_offsets, uint32_t n_groups,
                           uint64_t* out_hashes, uint32_t* out_counts, uint32_t* out_n, uint64_t* out_kmers)
{
    if (!m || m->ctx.empty()) { set_error("no GPUs"); return FPM_ERR_ARG; }
    if (!p || !group_offsets || (!seq && seq_bytes)) { set_error("NULL argument"); return FPM_ERR_ARG; }
    if (n_groups == 0) return FPM_OK;
    // contiguous ranges of whole sketches, balanced by bytes (Sketch.cpp:353-355 hands whole files to its pool threads)
    const int n = (int)std::min<uint64_t>(m->ctx.size(), n_groups);
    std::vector<uint32_t> cut(n + 1, n_groups);
    cut[0] = 0;
    const uint64_t base = group_offsets[0], total = group_offsets[n_groups] - base;
    uint32_t g = 0;
    for (int i = 1; i < n; i++) {
        const uint64_t want = base + (uint64_t)((unsigned __int128)total * i / n);
        while (g < n_groups && group_offsets[g] < want) g++;
        cut[i] = std::max(g, cut[i - 1]);
    }
    const uint32_t s = p->sketch_size;
    return on_all(m, n, [&](int i) -> int {
        const uint32_t g0 = cut[i], g1 = cut[i + 1];
        if (g1 <= g0) return FPM_OK;
        std::vector<uint64_t> off(g1 - g0 + 1);
        for (uint32_t k = g0; k <= g1; k++) off[k - g0] = group_offsets[k] - group_offsets[g0];
        return fpm_sketch_batch(m->ctx[i], p, seq + group_offsets[g0], off.back(), off.data(), g1 - g0, out_hashes + (uint64_t)g0 * s,
                                out_counts ? out_counts + (uint64_t)g0 * s : nullptr, out_n + g0, out_kmers ? out_kmers + g0 : nullptr);
    });
}

}  // extern "C"
