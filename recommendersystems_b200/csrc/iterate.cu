// iterate.cu -- the power iteration r <- (1-c) W^T r + S q  (K6, K7) with its convergence test.
//
// Restates Recommenders/RWRBased/Model.cs:
//   :33-50   seeded constructor  (rank[seed] = N, restart = e_seed)        -> k_init
//   :14-31   uniform constructor (rank = 1, restart = 1/N)                  -> k_init (seed == -1)
//   :76-100  deliverRanks  push loop, as a pull over CSR(W^T)               -> stream.cu (k_spmv_ws, k_cutrows_ws, k_finish_ws)
//   :103-108 updateRanks   rank <- next, next <- 0                          -> buffer swap (nothing to zero)
//   :110-115 checkConvergence  sum |rank - next|                            -> fused into the epilogue
//   :52-73   run() / run(double) / run(int)                                 -> rwr_run_threshold / rwr_run_fixed
//
// The iteration kernels live in stream.cu (K7: k_spmv_ws, k_cutrows_ws, k_finish_ws); this file holds the run modes,
// K6 (k_init), the workspaces and the result handle.
#include <cmath>

#include "iterate.h"

#include "iterate_dev.cuh"
#include "stream.h"
#include "dist.h"


// ------------------------------------------------------------------------------------------------ K6: init
template <typename T>
__global__ void k_init(int n, int seed, T omc, const T* __restrict__ inv, T* __restrict__ r0, T* __restrict__ x0,
                       T* __restrict__ x1, IterCtl* ctl, double S_uniform) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < 8) { x0[n + j] = (T)0; x1[n + j] = (T)0; }     // x[n] is the always-zero entry the edge stream pads with
    if (j == 0) {
        ctl->resid = 0.0; ctl->done = 0; ctl->iters = 0; ctl->ticket = 0; ctl->fault = 0; ctl->wait_clk = 0;
        ctl->tile_ctr = 0;
        ctl->seed = seed;
        if (seed < 0) ctl->S = S_uniform;
    }
    if (j >= n) return;
    const T r = (seed < 0) ? (T)1 : ((j == seed) ? (T)n : (T)0);           // Model.cs:24 / :44
    const T invj = inv[j];
    const T rw = mul_rn(omc, r);
    r0[j] = r;
    x0[j] = mul_rn(rw, invj);
    if (j == seed) ctl->S = (invj == (T)0) ? (double)r : (double)sub_rn(r, rw);
}

__global__ void k_partition(const u32* __restrict__ in_ptr, int n, u32 nnz, int n_chunks, int2* __restrict__ part) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > n_chunks) return;
    u64 d = (u64)k * CHUNK_ITEMS;
    const u64 total = (u64)n + nnz;
    if (d > total) d = total;
    u64 lo = d > nnz ? d - nnz : 0, hi = d < (u64)n ? d : (u64)n;
    while (lo < hi) {
        const u64 mid = (lo + hi) >> 1;
        if ((u64)in_ptr[mid + 1] <= d - mid - 1) lo = mid + 1; else hi = mid;
    }
    part[k] = make_int2((int)lo, (int)(d - lo));
}

__global__ void k_f64_to_f32(const double* __restrict__ in, float* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}

// row slice: after the allReduce of ctl->red, the sums over all ranks become S and the residual; threshold mode then
// applies the test of Model.cs:114.  A converged run ignores the (stale) sums of its no-op launches.
__global__ void k_after_reduce(IterCtl* ctl, double thr, int use_thr) {
    if (ctl->done) return;
    ctl->S = ctl->red[0];
    ctl->resid = ctl->red[1];
    if (use_thr && ctl->red[1] < thr) ctl->done = 1;
}

template <typename T>
__global__ void k_unpermute(const T* __restrict__ y_int, const int32_t* __restrict__ new_of_old, int n, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)y_int[new_of_old[i]];
}

// ------------------------------------------------------------------------------------------------ host side
int hub_entries_for(const rwr_graph* g, int precision) { return ws_hub_entries(g, precision); }

void iterate_prepare(rwr_graph* g) {
    cudaStream_t st = g->stream;
    const u64 total = (u64)g->n + (u64)g->nnz_in;
    g->n_chunks = (int)std::max<u64>(1, (total + CHUNK_ITEMS - 1) / CHUNK_ITEMS);
    g->part.alloc((size_t)g->n_chunks + 1, &g->pool);
    k_partition<<<div_up((size_t)g->n_chunks + 1, 256), 256, 0, st>>>(g->in_ptr.p, g->n, (u32)g->nnz_in, g->n_chunks, g->part.p);
    KERNEL_CHECK();
    CUDA_CHECK(cudaStreamSynchronize(st));
    stream_prepare(g);
}

void ensure_fp32_arrays(rwr_graph* g) {
    cudaStream_t st = g->stream;
    if (!g->inv32.p) {
        g->inv32.alloc((size_t)g->n + 4, &g->pool);
        if (g->n) k_f64_to_f32<<<div_up(g->n, 256), 256, 0, st>>>(g->inv64.p, g->inv32.p, g->n);
        KERNEL_CHECK();
    }
    if (g->layout == RWR_LAYOUT_VALUED && !g->in_val32.p && g->in_val64.p) {     // (a row slice keeps no whole-graph pull arrays)
        g->in_val32.alloc((size_t)g->nnz_in + IDX_PAD, &g->pool);
        if (g->nnz_in) k_f64_to_f32<<<div_up((size_t)g->nnz_in, 256), 256, 0, st>>>(g->in_val64.p, g->in_val32.p, (size_t)g->nnz_in);
        KERNEL_CHECK();
    }
    if (g->layout == RWR_LAYOUT_VALUED && !g->ws_val32.p && g->ws_val64.n) {
        g->ws_val32.alloc(g->ws_val64.n, &g->pool);
        k_f64_to_f32<<<div_up(g->ws_val64.n, 256), 256, 0, st>>>(g->ws_val64.p, g->ws_val32.p, g->ws_val64.n);
        KERNEL_CHECK();
    }
}

template <typename T> struct Prec;
template <> struct Prec<double> {
    static const double* inv(rwr_graph* g) { return g->inv64.p; }
    static const double* val(rwr_graph* g) { return g->in_val64.p; }
    static const double* wsval(rwr_graph* g) { return g->ws_val64.p; }
    static DevBuf<double>& ybuf(rwr_result* r) { return r->y64; }
    static constexpr int id = RWR_FP64;
};
template <> struct Prec<float> {
    static const float* inv(rwr_graph* g) { return g->inv32.p; }
    static const float* val(rwr_graph* g) { return g->in_val32.p; }
    static const float* wsval(rwr_graph* g) { return g->ws_val32.p; }
    static DevBuf<float>& ybuf(rwr_result* r) { return r->y32; }
    static constexpr int id = RWR_FP32;
};

// One iteration: K7 on this handle's stream, plus the exchange of a row-partitioned graph.
template <typename T>
static void launch_iteration(rwr_graph* g, const IterParams<T>& p, bool resid, double thr, int use_thr) {
    cudaStream_t st = g->stream;
    const bool parted = dist_n_ranks(g->comm) > 1;
    // on a row slice the convergence test needs the residual of all slices: it runs after the exchange
    ws_launch_iteration<T>(g, p, resid, thr, parted ? 0 : use_thr);
    static const bool skip_exchange = getenv("RWR_DIST_SKIP") != nullptr;      // timing probe only: wrong results
    if (parted && !skip_exchange) {
        // x_next: pushed by the next k_spmv_ws while it gathers (overlapped exchange), or already in every peer's copy when
        // the epilogue stored it there (p.n_peers > 0), else NCCL
        dist_exchange(g, (p.n_peers || g->overlap) ? nullptr : p.x_next, sizeof(T), p.ctl->red);
        k_after_reduce<<<1, 1, 0, st>>>(p.ctl, thr, use_thr);
        KERNEL_CHECK();
    }
}

struct RunWorkspace {
    Scratch<unsigned char> xa, xb, ya, yv;
    Scratch<double> carry, head, slot_S, slot_R;
    Scratch<IterCtl> ctl;
    void* x[2] = {nullptr, nullptr};      // the two gather vectors: scratch, or the peer-mapped buffers of a partitioned graph
    void alloc_x(rwr_graph* g, size_t vec_bytes) {
        if (g->p2p) { x[0] = g->px[0]; x[1] = g->px[1]; return; }
        xa.alloc(&g->scratch, vec_bytes); xb.alloc(&g->scratch, vec_bytes);
        x[0] = xa.p; x[1] = xb.p;
    }
    // column blocking of x (experimental): partial row sums of the virtual rows
    void alloc_yv(rwr_graph* g, size_t elt) {
        if (g->ws_compact) yv.alloc(&g->scratch, (size_t)g->v_compact * elt + 16);
        else if (g->x_blocks > 1) yv.alloc(&g->scratch, (size_t)g->x_blocks * (size_t)g->v_rows * elt + 16);
    }
};

// rows below this label are hot (their x entries are stored evict-last by the epilogue): the degree-sorted label prefix,
// or -- on a slice of a partitioned graph, whose labels are dealt over the slices -- the hot head of this rank's own rows
static int hot_limit(const rwr_graph* g) {
    const int parts = dist_n_ranks(g->comm);
    if (parts > 1 && (int)g->part_hot.size() == parts) return g->deal_rows[dist_rank(g->comm)] + g->part_hot[dist_rank(g->comm)];
    return g->n_hot;
}

// peers' copies of the buffer `x_next` is (row-partitioned graphs with peer-mapped gather vectors)
template <typename T>
static void set_peers(rwr_graph* g, IterParams<T>& p, const void* x_next) {
    p.parted = dist_n_ranks(g->comm) > 1;
    p.n_peers = 0;
    if (!g->p2p || g->overlap) return;          // overlapped exchange: the copy engines carry the slice, not the epilogue
    const int b = (x_next == g->px[0]) ? 0 : 1, me = dist_rank(g->comm);
    for (int r = 0; r < (int)g->peer_px[b].size(); r++)
        if (r != me) p.peer_next[p.n_peers++] = g->peer_px[b][r];
}

// Block layout of the stream and the tags of the overlapped exchange for the iteration that writes `x_next`.  `first`: the
// gather vector of this iteration comes from k_init (every rank computes all of it), nothing is in flight.
template <typename T>
static void set_exchange(rwr_graph* g, IterParams<T>& p, const void* x_next, bool first, bool live) {
    const int P = dist_n_ranks(g->comm), me = dist_rank(g->comm);
    p.compact = g->ws_compact ? 1 : 0;
    p.vrow_ptr = g->vrow_ptr.p; p.vpair = g->vpair.p;
    for (int k = 0; k < 8; k++) { p.blk_first_tile[k] = g->blk_first_tile[k]; p.blk_src[k] = ((me - k) % P + P) % P; }
    p.arrive = nullptr; p.wait_tag = 0;
    p.push_src = nullptr; p.push_peers = 0; p.push_tail = 0; p.push_bytes16 = 0; p.push_done = g->push_done; p.push_delay = 0;
    if (g->overlap && live) {
        // tag t marks the slices produced by the t-th iteration of this handle's life; all ranks count alike (collective calls)
        const uint64_t tag = ++g->xtag;
        if (!first) {
            // this launch gathers from the vector the previous iteration produced (tag - 1): it pushes this rank's slice of it
            // to the peers and waits for theirs
            DistSync* ds = (DistSync*)g->psync;
            p.arrive = ds->arrive;
            p.wait_tag = tag - 1;
            const int b = ((const void*)p.x == g->px[0]) ? 0 : 1;
            const size_t off = (size_t)g->row_begin * sizeof(T), len = (size_t)(g->row_end - g->row_begin) * sizeof(T);
            p.push_src = (const unsigned char*)g->px[b] + off;
            p.push_bytes16 = len & ~(size_t)15;
            p.push_tail = (int)((len % 16) / 4);
            p.push_peers = P - 1;
            for (int j = 1; j < P; j++) {
                const int peer = (me + j) % P;
                p.push_dst[j - 1] = (unsigned char*)g->peer_px[b][peer] + off;
                p.push_flag[j - 1] = &((DistSync*)g->peer_psync[peer])->arrive[me];
            }
            static const long long delay_us = getenv("RWR_DIST_PUSH_DELAY") ? atoll(getenv("RWR_DIST_PUSH_DELAY")) : 0;
            p.push_delay = delay_us * 1900;
        }
    }
}

// Runs one seed.  mode 0: fixed n_iter; mode 1: threshold.  Final rank lands in y_out (internal labels).
template <typename T>
static void run_one(rwr_graph* g, RunWorkspace& ws, int seed_orig, double c, int mode, int n_iter, double thr, int max_iter,
                    T* y_out, int* iters_out, double* resid_out, float* iter_ms, cudaEvent_t ev0, cudaEvent_t ev1) {
    cudaStream_t st = g->stream;
    const int n = g->n;
    int seed_int = -1;
    if (seed_orig >= 0) {
        CUDA_CHECK(cudaMemcpyAsync(&seed_int, g->new_of_old.p + seed_orig, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
    }
    const int hub = hub_entries_for(g, Prec<T>::id);

    T* xa = reinterpret_cast<T*>(ws.x[0]);
    T* xb = reinterpret_cast<T*>(ws.x[1]);
    T* ya = reinterpret_cast<T*>(ws.ya.p);
    IterParams<T> p;
    p.n = n; p.inv = Prec<T>::inv(g);
    p.ws_src = g->ws_src.p; p.ws_val = Prec<T>::wsval(g); p.ws_tile = g->ws_tile.p; p.ws_tiles = g->ws_tiles; p.tile_links = g->ws_tile_links;
    p.row_begin = g->row_begin; p.row_end = g->row_end;
    p.omc = (T)(1.0 - c);                                      // Model.cs:84 `(1 - dampingFactor)`
    p.seed = seed_int; p.inv_n = n ? 1.0 / (double)n : 0.0; p.hub = hub; p.n_hot = hot_limit(g); p.debug = 0;
    p.head_partial = ws.head.p; p.carry = ws.carry.p;
    p.slot_S = ws.slot_S.p; p.slot_R = ws.slot_R.p; p.ctl = ws.ctl.p;
    p.yv = reinterpret_cast<T*>(ws.yv.p); p.x_blocks = g->x_blocks; p.v_rows = g->v_rows;

    // uniform constructor: S0 = sum over nodes of (dangling ? 1 : 1 - fl((1-c)*1))
    const double omc_d = (double)p.omc;
    const double S_uniform = (double)g->n_dangling + (double)(n - g->n_dangling) * (1.0 - omc_d);
    // r0 goes to y_out in fixed mode (also the answer for n_iter == 0); threshold mode ping-pongs ya <-> y_out
    T* r_cur = (mode == 0) ? y_out : ya;
    k_init<T><<<div_up(std::max(n, 1), 256), 256, 0, st>>>(n, seed_int, p.omc, p.inv, r_cur, xa, xb, ws.ctl.p, S_uniform);
    KERNEL_CHECK();
    g->pool.launches += 1;

    CUDA_CHECK(cudaEventRecord(ev0, st));
    T* x_cur = xa;
    T* x_nxt = xb;
    int launched = 0;
    if (mode == 0) {
        auto launch_all = [&]() {
            for (int it = 0; it < n_iter; it++) {
                p.x = x_cur; p.x_next = x_nxt; p.r_prev = nullptr; p.y = y_out;
                set_peers<T>(g, p, x_nxt);
                set_exchange<T>(g, p, x_nxt, it == 0, true);
                launch_iteration<T>(g, p, /*resid=*/false, 0.0, 0);
                std::swap(x_cur, x_nxt);
            }
        };
        static const bool no_graph = getenv("RWR_NO_GRAPH") != nullptr;
        rwr_graph::IterGraph& ig = g->iter_graph[Prec<T>::id];
        const void* key[8] = {xa, xb, y_out, ws.ctl.p, ws.head.p, ws.carry.p, ws.slot_S.p, ws.slot_R.p};
        const bool graphable = n_iter >= 2 && dist_n_ranks(g->comm) == 1 && !no_graph && g->x_blocks == 1;
        const bool same = ig.n_iter == n_iter && ig.hub == hub && ig.c == c && memcmp(ig.ptr, key, sizeof(key)) == 0;
        if (graphable && same && (ig.exec || ig.seen)) {
            // second and later runs with this key: replay the graph of the whole loop (captured now if this is the second
            // run); nothing in it depends on the seed
            if (!ig.exec) {
                const int64_t counted = g->pool.launches;
                cudaGraph_t graph = nullptr;
                CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                try {
                    launch_all();
                } catch (...) {
                    cudaStreamEndCapture(st, &graph);
                    if (graph) cudaGraphDestroy(graph);
                    throw;
                }
                CUDA_CHECK(cudaStreamEndCapture(st, &graph));
                const cudaError_t err = cudaGraphInstantiate(&ig.exec, graph, 0);
                cudaGraphDestroy(graph);
                if (err != cudaSuccess) { ig.exec = nullptr; CUDA_CHECK(err); }
                g->pool.launches = counted;
            }
            CUDA_CHECK(cudaGraphLaunch(ig.exec, st));
            g->pool.launches += 3 * (int64_t)n_iter;
        } else {
            // first run with this key (or no graph at all): launch kernel by kernel.  Capturing and instantiating costs more
            // than one run saves, and the reference asks each graph for a single recommendation (Experiment.cs:109).
            if (graphable) {
                if (ig.exec) { cudaGraphExecDestroy(ig.exec); ig.exec = nullptr; }
                ig.seen = true; ig.n_iter = n_iter; ig.hub = hub; ig.c = c;
                memcpy(ig.ptr, key, sizeof(key));
            }
            launch_all();
        }
        launched = n_iter;
        *iters_out = n_iter;
        *resid_out = NAN;
    } else {
        // Model.cs:57-66.  Launch in batches; converged launches are no-ops, the flag is read between batches.
        IterCtl h{};
        const int BATCH = 8;
        bool done = false;
        while (!done) {
            for (int b = 0; b < BATCH; b++) {
                if (max_iter > 0 && launched >= max_iter) break;
                T* target = (r_cur == ya) ? y_out : ya;
                p.x = x_cur; p.x_next = x_nxt; p.r_prev = r_cur; p.y = target;
                set_peers<T>(g, p, x_nxt);
                set_exchange<T>(g, p, x_nxt, launched == 0, true);
                launch_iteration<T>(g, p, /*resid=*/true, thr, 1);
                std::swap(x_cur, x_nxt);
                r_cur = target;
                launched++;
            }
            CUDA_CHECK(cudaMemcpyAsync(&h, ws.ctl.p, sizeof(h), cudaMemcpyDeviceToHost, st));
            CUDA_CHECK(cudaStreamSynchronize(st));
            done = h.done || (max_iter > 0 && launched >= max_iter);
        }
        *iters_out = h.iters;
        *resid_out = h.resid;
        // the rank of iteration h.iters sits in ya when h.iters is even (r0 was in ya), else in y_out
        if ((h.iters & 1) == 0 && n) CUDA_CHECK(cudaMemcpyAsync(y_out, ya, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, st));
    }
    if (launched > 0) {
        // (overlapped exchange: the pushes are part of the k_spmv_ws launches, nothing is in flight once they have run)
        dist_allgather_rows(g, y_out, sizeof(T));                     // row-partitioned: every rank gets the whole rank vector
    }
    CUDA_CHECK(cudaEventRecord(ev1, st));
    if (!iter_ms) return;                          // the caller keeps enqueueing and reads the events after its own sync
    CUDA_CHECK(cudaEventSynchronize(ev1));
    float ms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, ev0, ev1));
    *iter_ms += ms;
}

// a k_spmv_ws that gave up waiting for a peer's slice leaves a mark: the results of that run are void
static void check_exchange_fault(rwr_graph* g, const IterCtl* ctl) {
#ifndef RWR_CHECKED
    if (!g->overlap) return;
#endif
    IterCtl h{};
    CUDA_CHECK(cudaMemcpyAsync(&h, ctl, sizeof(h), cudaMemcpyDeviceToHost, g->stream));
    CUDA_CHECK(cudaStreamSynchronize(g->stream));
    if (getenv("RWR_XCHG_TRACE"))
        fprintf(stderr, "[rwr xchg r%d] %d iterations, warps waited %.3f ms in total for slices (%.4f ms per warp and iteration)\n",
                dist_rank(g->comm), h.iters, (double)h.wait_clk / 1.9e6,
                h.iters ? (double)h.wait_clk / 1.9e6 / (148.0 * 16.0) / h.iters : 0.0);
    if (h.fault > 1) RWR_FAIL(RWR_E_INVALID, "checked build: index out of bounds in the iteration kernels (code %d)", h.fault);
    if (h.fault) RWR_FAIL(RWR_E_NCCL, "row-partitioned exchange: a peer's slice of x did not arrive in time (rank out of step or down)");
}

template <typename T>
static void run_all(rwr_graph* g, rwr_result* res, const int32_t* seeds, int n_seeds, double c, int mode, int n_iter,
                    double thr, int max_iter, int32_t* iters_out) {
    cudaStream_t st = g->stream;
    const size_t n = (size_t)g->n;
    if (Prec<T>::id == RWR_FP32) ensure_fp32_arrays(g);
    const size_t ld = (n + 3) & ~(size_t)3;
    res->ld = ld;
    if (Prec<T>::ybuf(res).n != std::max<size_t>(1, ld * (size_t)n_seeds))      // a re-run keeps its rank buffers
        Prec<T>::ybuf(res).alloc(std::max<size_t>(1, ld * (size_t)n_seeds), nullptr);
    RunWorkspace ws;
    const size_t vec_bytes = (n + 8) * sizeof(T);
    ws.alloc_x(g, vec_bytes); ws.alloc_yv(g, sizeof(T)); ws.ya.alloc(&g->scratch, vec_bytes);
    ws.carry.alloc(&g->scratch, (size_t)g->ws_tiles + 1); ws.head.alloc(&g->scratch, (size_t)g->ws_tiles + 1);
    CUDA_CHECK(cudaMemsetAsync(ws.carry.p, 0, ((size_t)g->ws_tiles + 1) * sizeof(double), st));
    CUDA_CHECK(cudaMemsetAsync(ws.head.p, 0, ((size_t)g->ws_tiles + 1) * sizeof(double), st));
    const size_t slots = (size_t)g->sm_count * 8 + 8;
    ws.slot_S.alloc(&g->scratch, slots); ws.slot_R.alloc(&g->scratch, slots);
    CUDA_CHECK(cudaMemsetAsync(ws.slot_S.p, 0, slots * sizeof(double), st));
    CUDA_CHECK(cudaMemsetAsync(ws.slot_R.p, 0, slots * sizeof(double), st));
    ws.ctl.alloc(&g->scratch, 1);
    CUDA_CHECK(cudaMemsetAsync(ws.ctl.p, 0, sizeof(IterCtl), st));
    DevEvent ev0, ev1, evA, evB;
    const int64_t launches0 = g->pool.launches;
    CUDA_CHECK(cudaEventRecord(evA, st));
    res->iterate_ms = 0.f;
    for (int s = 0; s < n_seeds; s++) {
        int it = 0;
        double rs = NAN;
        run_one<T>(g, ws, seeds[s], c, mode, n_iter, thr, max_iter, Prec<T>::ybuf(res).p + (size_t)s * ld, &it, &rs,
                   &res->iterate_ms, ev0, ev1);
        res->iters[s] = it;
        res->residual = rs;
        if (iters_out) iters_out[s] = it;
    }
    CUDA_CHECK(cudaEventRecord(evB, st));
    CUDA_CHECK(cudaEventSynchronize(evB));
    CUDA_CHECK(cudaEventElapsedTime(&res->total_ms, evA, evB));
    res->launches = g->pool.launches - launches0;
    check_exchange_fault(g, ws.ctl.p);
}

template <typename T>
static void profile_impl(rwr_graph* g, int seed_orig, double c, int reps, float* spmv_ms, float* fixup_ms) {
    cudaStream_t st = g->stream;
    const size_t n = (size_t)g->n;
    if (Prec<T>::id == RWR_FP32) ensure_fp32_arrays(g);
    RunWorkspace ws;
    const size_t vec_bytes = (n + 8) * sizeof(T);
    ws.alloc_x(g, vec_bytes); ws.alloc_yv(g, sizeof(T)); ws.ya.alloc(&g->scratch, vec_bytes);
    ws.carry.alloc(&g->scratch, (size_t)g->ws_tiles + 1); ws.head.alloc(&g->scratch, (size_t)g->ws_tiles + 1);
    CUDA_CHECK(cudaMemsetAsync(ws.carry.p, 0, ((size_t)g->ws_tiles + 1) * sizeof(double), st));
    CUDA_CHECK(cudaMemsetAsync(ws.head.p, 0, ((size_t)g->ws_tiles + 1) * sizeof(double), st));
    const size_t slots = (size_t)g->sm_count * 8 + 8;
    ws.slot_S.alloc(&g->scratch, slots); ws.slot_R.alloc(&g->scratch, slots);
    CUDA_CHECK(cudaMemsetAsync(ws.slot_S.p, 0, slots * sizeof(double), st));
    CUDA_CHECK(cudaMemsetAsync(ws.slot_R.p, 0, slots * sizeof(double), st));
    ws.ctl.alloc(&g->scratch, 1);
    int seed_int = 0;
    CUDA_CHECK(cudaMemcpyAsync(&seed_int, g->new_of_old.p + seed_orig, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    const int hub = hub_entries_for(g, Prec<T>::id);
    T* xa = reinterpret_cast<T*>(ws.x[0]);
    T* xb = reinterpret_cast<T*>(ws.x[1]);
    T* ya = reinterpret_cast<T*>(ws.ya.p);
    IterParams<T> p;
    p.n = g->n; p.inv = Prec<T>::inv(g);
    p.ws_src = g->ws_src.p; p.ws_val = Prec<T>::wsval(g); p.ws_tile = g->ws_tile.p; p.ws_tiles = g->ws_tiles; p.tile_links = g->ws_tile_links;
    p.row_begin = g->row_begin; p.row_end = g->row_end;
    p.omc = (T)(1.0 - c); p.seed = seed_int; p.inv_n = n ? 1.0 / (double)n : 0.0; p.hub = hub; p.n_hot = hot_limit(g);
    p.head_partial = ws.head.p; p.carry = ws.carry.p; p.slot_S = ws.slot_S.p; p.slot_R = ws.slot_R.p; p.ctl = ws.ctl.p;
    p.r_prev = nullptr; p.y = ya;
    p.yv = reinterpret_cast<T*>(ws.yv.p); p.x_blocks = g->x_blocks; p.v_rows = g->v_rows;
    { const char* dm = getenv("RWR_DEBUG_MODE"); p.debug = dm ? atoi(dm) : 0; }
    k_init<T><<<div_up(std::max((int)n, 1), 256), 256, 0, st>>>((int)n, seed_int, p.omc, p.inv, ya, xa, xb, ws.ctl.p, 0.0);
    KERNEL_CHECK();
    std::vector<DevEvent> ev(3 * (size_t)reps);
    T* x_cur = xa;
    T* x_nxt = xb;
    for (int it = 0; it < 3 + reps; it++) {
        p.x = x_cur; p.x_next = x_nxt;
        set_peers<T>(g, p, x_nxt);
        set_exchange<T>(g, p, x_nxt, true, false);       // kernel timing only: no pushes, nothing to wait for
        const int r = it - 3;
        if (r >= 0) CUDA_CHECK(cudaEventRecord(ev[3 * r], st));
        ws_launch_spmv_only<T>(g, p);
        if (r >= 0) CUDA_CHECK(cudaEventRecord(ev[3 * r + 1], st));
        ws_launch_finish_only<T>(g, p, false, 0.0, 0);
        KERNEL_CHECK();
        if (r >= 0) CUDA_CHECK(cudaEventRecord(ev[3 * r + 2], st));
        std::swap(x_cur, x_nxt);
    }
    CUDA_CHECK(cudaStreamSynchronize(st));
    double a = 0, b = 0;
    for (int r = 0; r < reps; r++) {
        float m1 = 0, m2 = 0;
        CUDA_CHECK(cudaEventElapsedTime(&m1, ev[3 * r], ev[3 * r + 1]));
        CUDA_CHECK(cudaEventElapsedTime(&m2, ev[3 * r + 1], ev[3 * r + 2]));
        a += m1; b += m2;
    }
    *spmv_ms = (float)(a / reps);
    *fixup_ms = (float)(b / reps);
    g->pool.launches += 1 + 2 * (3 + reps);
}

// One fixed-iteration run of one seed into a caller-provided rank vector (internal labels); used by the fused
// request path (rwr_recommend) so that no result object and no cudaMalloc / cudaFree sit on that path.
template <typename T>
void iterate_single_into(rwr_graph* g, int seed_orig, double c, int n_iter, T* y_out, float* iter_ms, int64_t* launches,
                         cudaEvent_t ext0, cudaEvent_t ext1) {
    const size_t n = (size_t)g->n;
    if (Prec<T>::id == RWR_FP32) ensure_fp32_arrays(g);
    RunWorkspace ws;
    const size_t vec_bytes = (n + 8) * sizeof(T);
    ws.alloc_x(g, vec_bytes); ws.alloc_yv(g, sizeof(T)); ws.ya.alloc(&g->scratch, 16);
    ws.carry.alloc(&g->scratch, (size_t)g->ws_tiles + 1); ws.head.alloc(&g->scratch, (size_t)g->ws_tiles + 1);
    const size_t slots = (size_t)g->sm_count * 8 + 8;
    ws.slot_S.alloc(&g->scratch, slots); ws.slot_R.alloc(&g->scratch, slots);
    ws.ctl.alloc(&g->scratch, 1);
    const int64_t l0 = g->pool.launches;
    int it = 0;
    double rs = 0;
#ifdef RWR_CHECKED
    const bool must_check = true;
#else
    const bool must_check = g->overlap;
#endif
    if (ext0 && ext1 && !must_check) {
        // no host synchronisation: the workspace goes back to the scratch pool while the kernels are still queued, which is
        // safe because every later user of those blocks enqueues on the same stream
        run_one<T>(g, ws, seed_orig, c, 0, n_iter, 0.0, 0, y_out, &it, &rs, nullptr, ext0, ext1);
    } else {
        DevEvent ev0, ev1;
        run_one<T>(g, ws, seed_orig, c, 0, n_iter, 0.0, 0, y_out, &it, &rs, iter_ms, ev0, ev1);
        check_exchange_fault(g, ws.ctl.p);
        if (ext0 && ext1) {                       // the caller reads its own events: bracket the (finished) run for it
            CUDA_CHECK(cudaEventRecord(ext0, g->stream));
            CUDA_CHECK(cudaEventRecord(ext1, g->stream));
        }
    }
    *launches += g->pool.launches - l0;
}
template void iterate_single_into<double>(rwr_graph*, int, double, int, double*, float*, int64_t*, cudaEvent_t, cudaEvent_t);
template void iterate_single_into<float>(rwr_graph*, int, double, int, float*, float*, int64_t*, cudaEvent_t, cudaEvent_t);

static int run_entry(rwr_graph* g, const int32_t* seeds, int32_t n_seeds, double c, int mode, int32_t n_iter, double thr,
                     int32_t max_iter, int32_t precision, int32_t* iters_out, rwr_result** out, rwr_result* reuse = nullptr) {
    rwr_result* res = nullptr;
    try {
        if (!out && !reuse) RWR_FAIL(RWR_E_INVALID, "out is NULL");
        if (out) *out = nullptr;
        if (!g) RWR_FAIL(RWR_E_INVALID, "graph is NULL");
        if (!g->built) RWR_FAIL(RWR_E_NOT_BUILT, "buildGraph() has not run (KeyNotFoundException at Model.cs:79)");
        if (n_seeds < 0 || (n_seeds && !seeds)) RWR_FAIL(RWR_E_INVALID, "bad seed list");
        if (mode == 0 && n_iter < 0) n_iter = 0;                     // `for (n = 0; n < nIterations; ..)` runs zero times
        if (precision != RWR_FP64 && precision != RWR_FP32) RWR_FAIL(RWR_E_INVALID, "unknown precision %d", precision);
        if (!(c == c)) RWR_FAIL(RWR_E_INVALID, "c is NaN");
        for (int s = 0; s < n_seeds; s++)
            if (seeds[s] < -1 || seeds[s] >= g->n) RWR_FAIL(RWR_E_BADSEED, "seed %d outside [0, %d)", seeds[s], g->n);
        CUDA_CHECK(cudaSetDevice(g->device));
        if (mode == 1 && !(thr > 0.0)) thr = (1.0 / 1.7976931348623157e308) * (double)g->n;   // Model.cs:53
        AllocStream alloc_on(g->stream);
        res = reuse ? reuse : new rwr_result();
        res->y64.plain = true;                    // rank buffers may be freed after the graph (and its stream) are gone
        res->y32.plain = true;
        res->g = g;
        res->device = g->device;
        res->n_seeds = n_seeds;
        res->precision = precision;
        res->seeds.assign(seeds, seeds + n_seeds);
        res->iters.assign(n_seeds, 0);
        if (precision == RWR_FP64) run_all<double>(g, res, seeds, n_seeds, c, mode, n_iter, thr, max_iter, iters_out);
        else run_all<float>(g, res, seeds, n_seeds, c, mode, n_iter, thr, max_iter, iters_out);
        if (out) *out = res;
        return RWR_OK;
    } catch (const RwrError& e) {
        if (!reuse) delete res;
        return e.code;
    } catch (...) {
        if (!reuse) delete res;
        rwr_set_error("unexpected exception");
        return RWR_E_INVALID;
    }
}

extern "C" {

int rwr_run_fixed(rwr_graph* g, const int32_t* seeds, int32_t n_seeds, double c, int32_t n_iter, int32_t precision,
                  rwr_result** out) {
    return run_entry(g, seeds, n_seeds, c, 0, n_iter, 0.0, 0, precision, nullptr, out);
}

int rwr_run_threshold(rwr_graph* g, const int32_t* seeds, int32_t n_seeds, double c, double thr, int32_t max_iter,
                      int32_t precision, int32_t* iters_out, rwr_result** out) {
    return run_entry(g, seeds, n_seeds, c, 1, 0, thr, max_iter, precision, iters_out, out);
}

// `new Model(graph, c, seed).run(n)` again on a live Model object: same graph, seed count and precision, the rank
// buffers are reused (no device allocation on the call).
int rwr_rerun_fixed(rwr_result* r, const int32_t* seeds, double c, int32_t n_iter) {
    if (!r || !r->g) { rwr_set_error("NULL result"); return RWR_E_INVALID; }
    return run_entry(r->g, seeds, r->n_seeds, c, 0, n_iter, 0.0, 0, r->precision, nullptr, nullptr, r);
}

int rwr_result_get_info(rwr_result* r, rwr_run_info* info) {
    RWR_API_BEGIN
    if (!r || !info) RWR_FAIL(RWR_E_INVALID, "NULL argument");
    memset(info, 0, sizeof(*info));
    info->n_seeds = r->n_seeds;
    info->n_nodes = r->g->n;
    info->precision = r->precision;
    info->iterations = r->n_seeds ? r->iters[r->n_seeds - 1] : 0;
    info->residual = r->residual;
    info->iterate_ms = r->iterate_ms;
    info->total_ms = r->total_ms;
    info->kernel_launches = r->launches;
    return RWR_OK;
    RWR_API_END
}

int rwr_scores(rwr_result* r, int32_t seed_slot, double* out_n) {
    RWR_API_BEGIN
    if (!r || !out_n) RWR_FAIL(RWR_E_INVALID, "NULL argument");
    if (seed_slot < 0 || seed_slot >= r->n_seeds) RWR_FAIL(RWR_E_INVALID, "seed slot %d outside [0, %d)", seed_slot, r->n_seeds);
    rwr_graph* g = r->g;
    CUDA_CHECK(cudaSetDevice(g->device));
    const int n = g->n;
    if (n == 0) return RWR_OK;
    Scratch<double> tmp;
    tmp.alloc(&g->scratch, n);
    if (r->precision == RWR_FP64)
        k_unpermute<double><<<div_up(n, 256), 256, 0, g->stream>>>(r->y64.p + (size_t)seed_slot * r->ld, g->new_of_old.p, n, tmp.p);
    else
        k_unpermute<float><<<div_up(n, 256), 256, 0, g->stream>>>(r->y32.p + (size_t)seed_slot * r->ld, g->new_of_old.p, n, tmp.p);
    KERNEL_CHECK();
    CUDA_CHECK(cudaMemcpyAsync(out_n, tmp.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
    CUDA_CHECK(cudaStreamSynchronize(g->stream));
    return RWR_OK;
    RWR_API_END
}

// Times the two kernels of one iteration separately (CUDA events on the handle's stream), `reps` iterations after
// 3 warm-up iterations.  bench.py uses it for the roofline of the dominant kernel.
int rwr_profile_iteration(rwr_graph* g, int32_t seed, double c, int32_t precision, int32_t reps, float* spmv_ms,
                          float* fixup_ms) {
    RWR_API_BEGIN
    if (!g || !spmv_ms || !fixup_ms) RWR_FAIL(RWR_E_INVALID, "NULL argument");
    if (!g->built) RWR_FAIL(RWR_E_NOT_BUILT, "buildGraph() has not run");
    if (seed < 0 || seed >= g->n) RWR_FAIL(RWR_E_BADSEED, "seed %d outside [0, %d)", seed, g->n);
    if (reps < 1) reps = 1;
    CUDA_CHECK(cudaSetDevice(g->device));
    if (precision == RWR_FP32) profile_impl<float>(g, seed, c, reps, spmv_ms, fixup_ms);
    else profile_impl<double>(g, seed, c, reps, spmv_ms, fixup_ms);
    return RWR_OK;
    RWR_API_END
}

#ifdef RWR_PROFILE_CLOCKS
int rwr_debug_clocks(unsigned long long* out16, int reset) {
    if (out16) cudaMemcpyFromSymbol(out16, g_clk, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_clk, z, sizeof(z)); }
    return 0;
}
#endif

void rwr_result_destroy(rwr_result* r) {
    if (!r) return;
    cudaSetDevice(r->device);       // never dereferences r->g: the graph may already be gone
    delete r;
}

}  // extern "C"
