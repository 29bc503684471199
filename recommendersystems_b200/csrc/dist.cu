// dist.cu -- K10: row-partitioned mode for one graph spread over several GPUs (no reference analogue).
//
// Every rank owns a contiguous block of rows of W^T (the destination nodes) with its own edge stream (stream.cu).  The
// internal labels are dealt over the slices (k_deal_labels, graph.cu): hot nodes one by one, clustered cold nodes in
// blocks, so the slices hold near-equal row counts and near-equal link counts and both the SpMV and the exchange are
// balanced.  An iteration on a rank needs the whole gather vector x and produces the slice of the next one for its rows:
//   * peer path (default, up to 8 ranks): the two gather vectors of every rank are mapped into every other rank through
//     CUDA IPC, and k_finish_ws stores each new entry into all copies over NVLink / NVSwitch while it computes it;
//   * NCCL path (fallback: RWR_DIST_NO_P2P=1, more than 8 ranks, or a mapping that fails on any rank): one grouped set
//     of in-place ncclBroadcast calls per iteration (the slices have unequal lengths).
// Either way the two scalars every rank needs -- the restart mass S and the L1 residual -- are summed with one 16-byte
// ncclAllReduce, which is also the barrier that orders the peer stores of one iteration before the gathers of the next.
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

__global__ void k_dist_noop() {}
// The push of the overlapped exchange: this rank's slice of the next x into every peer's gather vector, peer me+1 first,
// each followed by the arrival tag.  It has to run BESIDE the next k_spmv_ws, whose CTA leaves an SM 4096 registers and a
// few KB of shared memory, and it has to drive NVLink at full rate: the copy engines reached ~350 GB/s for these
// peer-mapped buffers (profiles/r02_part8.txt), 128 plain load/store threads per SM are latency-bound.  So one warp per
// SM lets the TMA do the work: bulk copies global -> shared (mbarrier) and shared -> peer memory (bulk groups), three 4-KB
// buffers in flight, no registers to speak of.
constexpr int PUSH_CHUNK = 4096;
constexpr int PUSH_BUFS = 3;
constexpr int PUSH_SMEM = PUSH_CHUNK * PUSH_BUFS + 64;
struct PushArgs {
    const unsigned char* src;     // this rank's slice in its own vector
    unsigned char* dst[7];        // the same rows in the peers' vectors, in push order
    unsigned long long* flag[7];  // arrive[me] on those peers
    int n_peers;
    size_t bytes16;               // multiple of 16
    int n_tail;                   // 4-byte words after that
    unsigned long long tag;
    unsigned* done;               // [7] CTAs that have finished a peer
};
__device__ __forceinline__ u32 push_smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(32, 1) k_push_slices(const PushArgs a) {
    __shared__ __align__(128) unsigned char buf[PUSH_BUFS][PUSH_CHUNK];
    __shared__ __align__(8) u64 bar[PUSH_BUFS];
    const int lane = threadIdx.x;
    if (lane == 0) {
        for (int s = 0; s < PUSH_BUFS; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(push_smem_addr(&bar[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const size_t n_chunks = (a.bytes16 + PUSH_CHUNK - 1) / PUSH_CHUNK;
    u32 phase[PUSH_BUFS] = {0, 0, 0};
    for (int j = 0; j < a.n_peers; j++) {
        if (lane == 0) {
            unsigned char* d = a.dst[j];
            // chunks c = blockIdx.x, + gridDim.x, ...; the load of chunk k+1 is issued before the store of chunk k
            auto chunk_len = [&](size_t c) { return (u32)((c + 1) * PUSH_CHUNK <= a.bytes16 ? PUSH_CHUNK : a.bytes16 - c * PUSH_CHUNK); };
            auto issue_load = [&](size_t c, int s) {
                // the bulk store that read buffer s (PUSH_BUFS chunks ago) must be done reading it
                asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PUSH_BUFS - 2) : "memory");
                const u32 len = chunk_len(c);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(push_smem_addr(&bar[s])), "r"(len) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(push_smem_addr(buf[s])),
                             "l"(a.src + c * PUSH_CHUNK), "r"(len), "r"(push_smem_addr(&bar[s]))
                             : "memory");
            };
            size_t c = blockIdx.x;
            int s = 0;
            if (c < n_chunks) issue_load(c, s);
            while (c < n_chunks) {
                const size_t cn = c + gridDim.x;
                const int sn = (s + 1) % PUSH_BUFS;
                if (cn < n_chunks) issue_load(cn, sn);
                // wait for chunk c in buffer s
                asm volatile(
                    "{\n.reg .pred p;\nPW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra PD;\nbra PW;\nPD:\n}\n" ::"r"(push_smem_addr(&bar[s])),
                    "r"(phase[s])
                    : "memory");
                phase[s] ^= 1u;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d + c * PUSH_CHUNK), "r"(push_smem_addr(buf[s])),
                             "r"(chunk_len(c))
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                c = cn;
                s = sn;
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // every store to this peer has completed
            asm volatile("fence.proxy.async;" ::: "memory");
        }
        if (blockIdx.x == 0 && lane < a.n_tail)
            reinterpret_cast<unsigned*>(a.dst[j] + a.bytes16)[lane] = reinterpret_cast<const unsigned*>(a.src + a.bytes16)[lane];
        __threadfence_system();
        __syncwarp();
        if (lane == 0) {
            const unsigned cdone = atomicAdd(&a.done[j], 1u);
            if (cdone == gridDim.x - 1) {                  // every CTA's stores to this peer are out: hand over the tag
                a.done[j] = 0;
                __threadfence_system();
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.flag[j]), "l"(a.tag) : "memory");
            }
        }
        __syncwarp();
    }
}

// test knob RWR_DIST_PUSH_DELAY=<microseconds>: holds the copy engines back so that a gather that does not wait for its
// slice reads stale data for sure (tests/test_gpu_partitioned.py, mode "overlapped_slow")
__global__ void k_dist_delay(long long cycles) {
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) __nanosleep(1000);
}

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
    // two ranks: the one 44-MB-class slice a rank sends rides inside the epilogue kernel almost for free (peer stores at
    // NVLink rate while the kernel streams its rows), cheaper than the block bookkeeping of the overlapped form
    // (profiles/r02_slice_probe.txt); from three ranks on the (P - 1)-fold egress would be exposed.  RWR_DIST_OVERLAP=1 forces it.
    return c->n_ranks >= 3 || getenv("RWR_DIST_OVERLAP") != nullptr;
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
    // all ranks must agree: one failing mapping anywhere sends everybody to the NCCL path
    DevBuf<double> flag;
    flag.alloc(1);
    const double mine_bad = ok ? 0.0 : 1.0;
    CUDA_CHECK(cudaMemcpyAsync(flag.p, &mine_bad, sizeof(double), cudaMemcpyHostToDevice, st));
    NCCL_CHECK(api.AllReduce(flag.p, flag.p, 1, ncclFloat64, ncclSum, c->comm, st));
    double bad = 0.0;
    CUDA_CHECK(cudaMemcpyAsync(&bad, flag.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    if (bad != 0.0) { dist_release_p2p(g); return; }
    g->p2p = true;
    if (g->ws_compact && dist_overlap_wanted(g)) {
        int prio_lo = 0, prio_hi = 0;
        CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        for (int k = 0; k < rwr_graph::XSTREAMS; k++)
            CUDA_CHECK(cudaStreamCreateWithPriority(&g->xstream[k], cudaStreamNonBlocking, prio_hi));
        CUDA_CHECK(cudaMalloc((void**)&g->push_done, 8 * sizeof(unsigned)));
        CUDA_CHECK(cudaMemsetAsync(g->push_done, 0, 8 * sizeof(unsigned), st));
        CUDA_CHECK(cudaEventCreateWithFlags(&g->ev_fin, cudaEventDisableTiming));
        for (int b = 0; b < 2; b++)
            for (int k = 0; k < rwr_graph::XSTREAMS; k++) CUDA_CHECK(cudaEventCreateWithFlags(&g->ev_push[b][k], cudaEventDisableTiming));
        g->overlap = true;
    }
}

void dist_before_iteration(rwr_graph* g, int b) {
    if (!g->overlap || !g->push_pending[b]) return;
    for (int k = 0; k < g->push_streams[b]; k++) CUDA_CHECK(cudaStreamWaitEvent(g->stream, g->ev_push[b][k], 0));
    g->push_pending[b] = false;
}

void dist_push_slice(rwr_graph* g, int b, size_t elt, unsigned long long tag) {
    if (!g->overlap) return;
    rwr_comm* c = g->comm;
    const int P = c->n_ranks, me = c->rank;
    const size_t off = (size_t)g->row_begin * elt, len = (size_t)(g->row_end - g->row_begin) * elt;
    DistSync* mine = (DistSync*)g->psync;
    CUDA_CHECK(cudaEventRecord(g->ev_fin, g->stream));
    static const bool use_ce = getenv("RWR_DIST_PUSH_CE") != nullptr;      // probe: copy engines instead of the push kernel
    const int n_streams = use_ce ? rwr_graph::XSTREAMS : 1;
    for (int k = 0; k < n_streams; k++) CUDA_CHECK(cudaStreamWaitEvent(g->xstream[k], g->ev_fin, 0));
    static const long long delay_us = getenv("RWR_DIST_PUSH_DELAY") ? atoll(getenv("RWR_DIST_PUSH_DELAY")) : 0;
    if (delay_us > 0)
        for (int k = 0; k < n_streams; k++) k_dist_delay<<<1, 1, 0, g->xstream[k]>>>(delay_us * 1900);
    if (use_ce) {
        // peer me+1 first (it gathers from this slice first), the streams take the peers in turn
        for (int j = 1; j < P; j++) {
            const int peer = (me + j) % P;
            cudaStream_t xs = g->xstream[(j - 1) % rwr_graph::XSTREAMS];
            if (len)
                CUDA_CHECK(cudaMemcpyAsync((unsigned char*)g->peer_px[b][peer] + off, (unsigned char*)g->px[b] + off, len, cudaMemcpyDefault, xs));
            DistSync* theirs = (DistSync*)g->peer_psync[peer];
            CUDA_CHECK(cudaMemcpyAsync(&theirs->arrive[me], &mine->tag_out[b], sizeof(unsigned long long), cudaMemcpyDefault, xs));
        }
    } else {
        PushArgs a{};
        a.src = (const unsigned char*)g->px[b] + off;
        a.bytes16 = len & ~(size_t)15;
        a.n_tail = (int)((len % 16) / 4);
        a.n_peers = P - 1;
        for (int j = 1; j < P; j++) {
            const int peer = (me + j) % P;
            a.dst[j - 1] = (unsigned char*)g->peer_px[b][peer] + off;
            a.flag[j - 1] = &((DistSync*)g->peer_psync[peer])->arrive[me];
        }
        a.tag = tag;
        a.done = g->push_done;
        // the SM keeps ONE shared-memory / L1 split while CTAs of both kernels live on it: ask for the split k_spmv_ws runs with
        CUDA_CHECK(cudaFuncSetAttribute(k_push_slices, cudaFuncAttributePreferredSharedMemoryCarveout, g->xchg_carveout_pct));
        k_push_slices<<<g->sm_count, 32, 0, g->xstream[0]>>>(a);
        KERNEL_CHECK();
    }
    for (int k = 0; k < n_streams; k++) CUDA_CHECK(cudaEventRecord(g->ev_push[b][k], g->xstream[k]));
    g->push_streams[b] = n_streams;
    g->push_pending[b] = true;
}

void dist_drain_pushes(rwr_graph* g) {
    if (!g->overlap) return;
    for (int b = 0; b < 2; b++) dist_before_iteration(g, b);
}

void dist_barrier(rwr_graph* g) {
    rwr_comm* c = g->comm;
    if (!c || c->n_ranks < 2 || c->fake) return;
    DevBuf<double> d;
    d.alloc(1);
    CUDA_CHECK(cudaMemsetAsync(d.p, 0, sizeof(double), g->stream));
    NCCL_CHECK(nccl().AllReduce(d.p, d.p, 1, ncclFloat64, ncclSum, c->comm, g->stream));
    CUDA_CHECK(cudaStreamSynchronize(g->stream));
}

// Teardown in three steps (collective on a partitioned handle): every rank closes its mappings of the peers' buffers,
// all ranks meet on the communicator, and only then does each rank free the buffers it exported -- cudaFree of memory a
// peer still has open through cudaIpcOpenMemHandle is undefined behaviour.
void dist_release_p2p(rwr_graph* g) {
    for (int k = 0; k < rwr_graph::XSTREAMS; k++)
        if (g->xstream[k]) { cudaStreamSynchronize(g->xstream[k]); cudaStreamDestroy(g->xstream[k]); g->xstream[k] = nullptr; }
    if (g->push_done) { cudaFree(g->push_done); g->push_done = nullptr; }
    if (g->ev_fin) { cudaEventDestroy(g->ev_fin); g->ev_fin = nullptr; }
    for (int b = 0; b < 2; b++)
        for (int k = 0; k < rwr_graph::XSTREAMS; k++)
            if (g->ev_push[b][k]) { cudaEventDestroy(g->ev_push[b][k]); g->ev_push[b][k] = nullptr; }
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
