// dist.cu -- K10: row-partitioned mode for one graph spread over several GPUs (no reference analogue): the communicator,
// the peer mapping, the collectives of the partitioned build.
//
// Every rank owns a contiguous block of rows of W^T (the destination nodes) with its own edge stream (stream.cu).  The
// internal labels are dealt over the slices (k_deal_labels, graph.cu): hot nodes one by one, clustered cold nodes in
// blocks, so the slices hold near-equal row counts and near-equal link counts and both the SpMV and the exchange are
// balanced.  An iteration on a rank needs the whole gather vector x and produces the slice of the next one for its rows.
// The two gather vectors of every rank (and a page of arrival tags) are mapped into every other rank through CUDA IPC:
//   * overlapped exchange (default from 5 ranks on): k_finish_ws writes the slice locally, the NEXT k_spmv_ws pushes it to
//     the peers with a TMA-driving warp per CTA while the other warps gather block by block as the slices arrive (stream.cu);
//   * peer stores (up to 4 ranks, or RWR_DIST_LEGACY=1): k_finish_ws stores each new entry into all copies over NVLink;
//   * NCCL path (fallback: RWR_DIST_NO_P2P=1, more than 8 ranks, or a mapping that fails on any rank): one grouped set
//     of in-place ncclBroadcast calls per iteration (the slices have unequal lengths).
// In every form the two scalars every rank needs -- the restart mass S and the L1 residual -- are summed with one 16-byte
// ncclAllReduce, which is also the barrier of the iteration: a rank can only be one allReduce ahead of its peers, so the
// vector it writes remotely is never the one a peer still gathers from.
// The partitioned build (graph.cu) uses dist_allreduce_sum / dist_alltoallv below.
// NCCL is bound at run time (dlopen) so that a single-GPU host needs no NCCL at all.
#include <dlfcn.h>

#include <algorithm>
#include <mutex>

#include "dist.h"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclChar = 0, ncclFloat64 = 8 };      // nccl.h (2.x): ncclDataType_t
enum { ncclSum = 0, ncclMax = 2 };

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

NcclApi g_nccl;
std::once_flag g_nccl_once;
std::string g_nccl_err;

void load_nccl() {
    const char* names[] = {getenv("RWR_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        if (!nm || !*nm) continue;
        g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) { g_nccl_err = "libnccl.so.2 not found (set RWR_NCCL_LIB)"; return; }
#define BIND(field, sym)                                                        \
    *(void**)(&g_nccl.field) = dlsym(g_nccl.lib, sym);                          \
    if (!g_nccl.field) { g_nccl_err = std::string("NCCL symbol missing: ") + sym; g_nccl.lib = nullptr; return; }
    BIND(GetUniqueId, "ncclGetUniqueId")
    BIND(CommInitRank, "ncclCommInitRank")
    BIND(CommDestroy, "ncclCommDestroy")
    BIND(Broadcast, "ncclBroadcast")
    BIND(AllReduce, "ncclAllReduce")
    BIND(Send, "ncclSend")
    BIND(Recv, "ncclRecv")
    BIND(GroupStart, "ncclGroupStart")
    BIND(GroupEnd, "ncclGroupEnd")
    BIND(GetErrorString, "ncclGetErrorString")
#undef BIND
}

NcclApi& nccl() {
    std::call_once(g_nccl_once, load_nccl);
    if (!g_nccl.lib) RWR_FAIL(RWR_E_NCCL, "%s", g_nccl_err.c_str());
    return g_nccl;
}

#define NCCL_CHECK(expr)                                                                              \
    do {                                                                                              \
        int rc__ = (expr);                                                                            \
        if (rc__ != ncclSuccess) RWR_FAIL(RWR_E_NCCL, "%s failed: %s", #expr, nccl().GetErrorString(rc__)); \
    } while (0)

}  // namespace

struct rwr_comm {
    int rank = 0, n_ranks = 1, device = 0;
    ncclComm_t comm = nullptr;
    bool fake = false;      // RWR_FAKE_COMM probe: a slice of a partitioned graph on one GPU, no exchange (timing only)
};

bool dist_is_fake(const rwr_comm* c) { return c && c->fake; }
int dist_rank(const rwr_comm* c) { return c ? c->rank : 0; }
int dist_n_ranks(const rwr_comm* c) { return c ? c->n_ranks : 1; }

void dist_allgather_rows(rwr_graph* g, void* vec, size_t elt) {
    rwr_comm* c = g->comm;
    if (!c || c->n_ranks < 2 || c->fake) return;
    NcclApi& api = nccl();
    NCCL_CHECK(api.GroupStart());
    for (int r = 0; r < c->n_ranks; r++) {
        const size_t b = (size_t)g->part_rows[r], e = (size_t)g->part_rows[r + 1];
        if (e == b) continue;
        unsigned char* ptr = (unsigned char*)vec + b * elt;
        NCCL_CHECK(api.Broadcast(ptr, ptr, (e - b) * elt, ncclInt8, r, c->comm, g->stream));
    }
    NCCL_CHECK(api.GroupEnd());
}

void dist_exchange(rwr_graph* g, void* x_next, size_t elt, double* two_doubles) {
    rwr_comm* c = g->comm;
    if (!c || c->n_ranks < 2 || c->fake) return;
    if (x_next) dist_allgather_rows(g, x_next, elt);
    NCCL_CHECK(nccl().AllReduce(two_doubles, two_doubles, 2, ncclFloat64, ncclSum, c->comm, g->stream));
}

void dist_allreduce_sum(rwr_graph* g, void* buf, size_t count, int dtype) {
    rwr_comm* c = g->comm;
    if (!c || c->n_ranks < 2 || c->fake || count == 0) return;
    // NCCL counts are size_t, but keep single calls below 2^31 elements
    const size_t elt = dtype == DIST_U32 ? 4 : 8, step = (size_t)1 << 30;
    for (size_t off = 0; off < count; off += step)
        NCCL_CHECK(nccl().AllReduce((char*)buf + off * elt, (char*)buf + off * elt, std::min(step, count - off), dtype, ncclSum, c->comm, g->stream));
}

void dist_allreduce_max_u32(rwr_graph* g, u32* buf, size_t count) {
    rwr_comm* c = g->comm;
    if (!c || c->n_ranks < 2 || c->fake || count == 0) return;
    NCCL_CHECK(nccl().AllReduce(buf, buf, count, DIST_U32, ncclMax, c->comm, g->stream));
}

void dist_alltoallv(rwr_graph* g, const void* send, const size_t* send_off, const size_t* send_cnt, void* recv,
                    const size_t* recv_off, const size_t* recv_cnt, size_t elt) {
    rwr_comm* c = g->comm;
    if (!c || c->fake) return;
    NcclApi& api = nccl();
    NCCL_CHECK(api.GroupStart());
    for (int r = 0; r < c->n_ranks; r++) {
        if (send_cnt[r]) NCCL_CHECK(api.Send((const char*)send + send_off[r] * elt, send_cnt[r] * elt, ncclInt8, r, c->comm, g->stream));
        if (recv_cnt[r]) NCCL_CHECK(api.Recv((char*)recv + recv_off[r] * elt, recv_cnt[r] * elt, ncclInt8, r, c->comm, g->stream));
    }
    NCCL_CHECK(api.GroupEnd());
}

void synth_generate_device(rwr_graph* g, const rwr_synth_spec* spec);     // synth.cu

// Two persistent gather vectors per rank, mapped by every peer (one process per GPU -> CUDA IPC handles, exchanged
// through the communicator itself).  RWR_DIST_NO_P2P=1, more than 8 ranks or a failing mapping fall back to NCCL.
bool dist_overlap_wanted(const rwr_graph* g) {
    const rwr_comm* c = g->comm;
    if (!c || c->n_ranks < 2 || c->n_ranks > 8 || getenv("RWR_DIST_LEGACY") || getenv("RWR_DIST_NO_P2P")) return false;
    // Few ranks: the (P - 1) slices a rank sends ride inside the epilogue kernel at NVLink rate, cheaper than the block
    // bookkeeping and the gathering warp the overlapped form gives up (measured, profiles/r02_part2.txt, r02_part8.txt:
    // P = 2 0.414 ms against 0.45 ms overlapped, P = 4 0.585 against 0.599, P = 8 1.41 against 1.09).  RWR_DIST_OVERLAP=1 forces it.
    return c->n_ranks >= 5 || getenv("RWR_DIST_OVERLAP") != nullptr;
}

void dist_setup_p2p(rwr_graph* g) {
    rwr_comm* c = g->comm;
    g->p2p = false;
    g->overlap = false;
    if (!c || c->n_ranks < 2 || c->n_ranks > 8 || c->fake || getenv("RWR_DIST_NO_P2P")) return;
    cudaStream_t st = g->stream;
    const int P = c->n_ranks;
    const size_t bytes = ((size_t)g->n + 8) * 8;
    for (int b = 0; b < 2; b++) {
        CUDA_CHECK(cudaMalloc(&g->px[b], bytes));
        g->pool.bytes += (int64_t)bytes;
        CUDA_CHECK(cudaMemsetAsync(g->px[b], 0, bytes, st));
    }
    CUDA_CHECK(cudaMalloc(&g->psync, sizeof(DistSync)));
    CUDA_CHECK(cudaMemsetAsync(g->psync, 0, sizeof(DistSync), st));
    // handles: [P][3] x 64 bytes (the two gather vectors, the tag page), every rank broadcasts its own triple
    constexpr int HB = 3 * 64;
    DevBuf<unsigned char> hb;
    hb.alloc((size_t)P * HB);
    std::vector<unsigned char> host((size_t)P * HB, 0);
    void* mine[3] = {g->px[0], g->px[1], g->psync};
    for (int b = 0; b < 3; b++) {
        cudaIpcMemHandle_t h;
        CUDA_CHECK(cudaIpcGetMemHandle(&h, mine[b]));
        static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t size");
        memcpy(host.data() + (size_t)c->rank * HB + b * 64, &h, 64);
    }
    CUDA_CHECK(cudaMemcpyAsync(hb.p, host.data(), host.size(), cudaMemcpyHostToDevice, st));
    NcclApi& api = nccl();
    NCCL_CHECK(api.GroupStart());
    for (int r = 0; r < P; r++) NCCL_CHECK(api.Broadcast(hb.p + (size_t)r * HB, hb.p + (size_t)r * HB, HB, ncclInt8, r, c->comm, st));
    NCCL_CHECK(api.GroupEnd());
    CUDA_CHECK(cudaMemcpyAsync(host.data(), hb.p, host.size(), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    int ok = 1;
    for (int b = 0; b < 2; b++) g->peer_px[b].assign(P, nullptr);
    g->peer_psync.assign(P, nullptr);
    for (int r = 0; r < P && ok; r++) {
        for (int b = 0; b < 3; b++) {
            void** slot = b < 2 ? &g->peer_px[b][r] : &g->peer_psync[r];
            if (r == c->rank) { *slot = mine[b]; continue; }
            cudaIpcMemHandle_t h;
            memcpy(&h, host.data() + (size_t)r * HB + b * 64, 64);
            if (cudaIpcOpenMemHandle(slot, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                *slot = nullptr;
                ok = 0;
                break;
            }
        }
    }
    // all ranks must agree: one failing mapping anywhere sends everybody to the NCCL path, and the overlapped exchange needs
    // the slice-aligned stream on EVERY rank (a rank without a single link in its rows builds a plain, empty stream)
    DevBuf<double> flag;
    flag.alloc(2);
    const double mine_bad[2] = {ok ? 0.0 : 1.0, (g->ws_compact && dist_overlap_wanted(g)) ? 0.0 : 1.0};
    CUDA_CHECK(cudaMemcpyAsync(flag.p, mine_bad, sizeof(mine_bad), cudaMemcpyHostToDevice, st));
    NCCL_CHECK(api.AllReduce(flag.p, flag.p, 2, ncclFloat64, ncclSum, c->comm, st));
    double bad[2] = {0.0, 0.0};
    CUDA_CHECK(cudaMemcpyAsync(bad, flag.p, sizeof(bad), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    if (bad[0] != 0.0) { dist_release_p2p(g); return; }
    g->p2p = true;
    if (bad[1] == 0.0) {
        CUDA_CHECK(cudaMalloc((void**)&g->push_done, 8 * sizeof(unsigned)));
        CUDA_CHECK(cudaMemsetAsync(g->push_done, 0, 8 * sizeof(unsigned), st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        g->overlap = true;
    }
}

// Teardown in three steps (collective on a partitioned handle): every rank closes its mappings of the peers' buffers,
// all ranks meet on the communicator, and only then does each rank free the buffers it exported -- cudaFree of memory a
// peer still has open through cudaIpcOpenMemHandle is undefined behaviour.
void dist_release_p2p(rwr_graph* g) {
    if (g->push_done) { cudaFree(g->push_done); g->push_done = nullptr; }
    g->overlap = false;
    if (!g->px[0] && !g->px[1] && !g->psync && g->peer_px[0].empty() && g->peer_px[1].empty()) { g->p2p = false; return; }
    const int me = dist_rank(g->comm);
    for (int b = 0; b < 2; b++) {
        for (int r = 0; r < (int)g->peer_px[b].size(); r++)
            if (r != me && g->peer_px[b][r]) cudaIpcCloseMemHandle(g->peer_px[b][r]);
        g->peer_px[b].clear();
    }
    for (int r = 0; r < (int)g->peer_psync.size(); r++)
        if (r != me && g->peer_psync[r]) cudaIpcCloseMemHandle(g->peer_psync[r]);
    g->peer_psync.clear();
    rwr_comm* c = g->comm;
    if (c && c->n_ranks > 1 && !c->fake && c->comm && g_nccl.lib && g->px[0]) {
        // px[0] doubles as the 8-byte payload of the barrier: nobody reads it any more
        if (g_nccl.AllReduce(g->px[0], g->px[0], 1, ncclFloat64, ncclSum, c->comm, g->stream) == ncclSuccess)
            cudaStreamSynchronize(g->stream);
        else
            cudaGetLastError();
    }
    for (int b = 0; b < 2; b++)
        if (g->px[b]) { cudaFree(g->px[b]); g->pool.bytes -= (int64_t)(((size_t)g->n + 8) * 8); g->px[b] = nullptr; }
    if (g->psync) { cudaFree(g->psync); g->psync = nullptr; }
    g->p2p = false;
}

extern "C" {

int rwr_comm_unique_id(void* id128) {
    RWR_API_BEGIN
    if (!id128) RWR_FAIL(RWR_E_INVALID, "NULL argument");
    ncclUniqueId id;
    NCCL_CHECK(nccl().GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return RWR_OK;
    RWR_API_END
}

int rwr_comm_create(int32_t rank, int32_t n_ranks, const void* id128, const rwr_opts* opts, rwr_comm** out) {
    rwr_comm* c = nullptr;
    try {
        if (!out || !id128) RWR_FAIL(RWR_E_INVALID, "NULL argument");
        *out = nullptr;
        if (n_ranks < 1 || rank < 0 || rank >= n_ranks) RWR_FAIL(RWR_E_INVALID, "rank %d of %d", rank, n_ranks);
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) RWR_FAIL(RWR_E_CUDA, "no CUDA device");
        c = new rwr_comm();
        c->rank = rank;
        c->n_ranks = n_ranks;
        if (opts && opts->device >= 0) CUDA_CHECK(cudaSetDevice(opts->device));
        CUDA_CHECK(cudaGetDevice(&c->device));
        // probe knob: one slice of a partitioned graph on a single GPU, no NCCL and no exchange --
        // kernel timing and ncu captures of a slice only, the results of a run are wrong
        c->fake = getenv("RWR_FAKE_COMM") != nullptr;
        if (!c->fake) {
            ncclUniqueId id;
            memcpy(&id, id128, sizeof(id));
            NCCL_CHECK(nccl().CommInitRank(&c->comm, n_ranks, id, rank));
        }
        *out = c;
        return RWR_OK;
    } catch (const RwrError& e) {
        delete c;
        return e.code;
    } catch (...) {
        delete c;
        rwr_set_error("unexpected exception");
        return RWR_E_INVALID;
    }
}

void rwr_comm_destroy(rwr_comm* c) {
    if (!c) return;
    if (c->comm && g_nccl.lib) {
        cudaSetDevice(c->device);
        g_nccl.CommDestroy(c->comm);
    }
    delete c;
}

int rwr_graph_create_flat(int32_t n_nodes, const int64_t* node_id, const int32_t* node_type, int64_t n_links, const int32_t* src,
                          const int32_t* dst, const int32_t* etype, const double* w, const rwr_opts* opts, rwr_comm* comm,
                          rwr_graph** out);                               // graph.cu

// Every rank runs the same deterministic generator and keeps the links of the sources it owns; rwr_graph_build exchanges
// them so that every rank ends up with the rows of W^T it owns.
int rwr_synth_create_partitioned(const rwr_synth_spec* spec, const rwr_opts* opts, rwr_comm* comm, rwr_graph** out) {
    rwr_graph* g = nullptr;
    try {
        if (!out || !spec || !comm) RWR_FAIL(RWR_E_INVALID, "NULL argument");
        *out = nullptr;
        g = new rwr_graph();
        rwr_opts o;
        if (opts) o = *opts; else { memset(&o, 0, sizeof(o)); o.hub_entries = -1; }
        o.device = comm->device;
        graph_init_device(g, &o);
        g->comm = comm;
        // every rank generates, sorts and keeps only the links of the sources it owns (the RWR_FAKE_COMM probe and
        // RWR_PART_REPLICATED=1 keep the whole graph on every rank, as round 1 did)
        g->part_build = comm->n_ranks > 1 && !comm->fake && !getenv("RWR_PART_REPLICATED");
        AllocStream alloc_on(g->stream);
        synth_generate_device(g, spec);
        *out = g;
        return RWR_OK;
    } catch (const RwrError& e) {
        rwr_graph_destroy(g);
        return e.code;
    } catch (...) {
        rwr_graph_destroy(g);
        rwr_set_error("unexpected exception");
        return RWR_E_INVALID;
    }
}

// Every rank passes the same link list.
int rwr_graph_create_partitioned(int32_t n_nodes, const int64_t* node_id, const int32_t* node_type, int64_t n_links,
                                 const int32_t* src, const int32_t* dst, const int32_t* etype, const double* w,
                                 const rwr_opts* opts, rwr_comm* comm, rwr_graph** out) {
    if (!comm) { rwr_set_error("NULL communicator"); return RWR_E_INVALID; }
    rwr_opts o;
    if (opts) o = *opts; else { memset(&o, 0, sizeof(o)); o.hub_entries = -1; }
    o.device = comm->device;
    return rwr_graph_create_flat(n_nodes, node_id, node_type, n_links, src, dst, etype, w, &o, comm, out);
}

}  // extern "C"
