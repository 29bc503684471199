// iterate.cu -- the power iteration r <- (1-c) W^T r + S q  (K6, K7) with its convergence test.
//
// Restates Recommenders/RWRBased/Model.cs:
//   :33-50   seeded constructor  (rank[seed] = N, restart = e_seed)        -> k_init
//   :14-31   uniform constructor (rank = 1, restart = 1/N)                  -> k_init (seed == -1)
//   :76-100  deliverRanks  push loop, as a pull over CSR(W^T)               -> k_spmv + k_fixup
//   :103-108 updateRanks   rank <- next, next <- 0                          -> buffer swap (nothing to zero)
//   :110-115 checkConvergence  sum |rank - next|                            -> fused into the epilogue
//   :52-73   run() / run(double) / run(int)                                 -> rwr_run_threshold / rwr_run_fixed
//
// One iteration = 2 launches:
//   k_spmv   persistent, one 1024-thread CTA per SM = 4 groups of 256 threads.  Each group walks merge-path chunks
//            (<= 2044 rows+nnz): coalesced int4 index loads (evict-first), gathers of the pre-scaled vector
//            x_i = fl(fl((1-c) r_i) * w_i) from a TMA-staged shared-memory table (the hottest sources after the
//            degree relabel) or from L2 (evict-last), products parked in shared memory, then one thread per row sums
//            its products in storage order (== the reference's accumulation order), rows >= 64 nnz by a warp.
//            Epilogue per finished row: y_t, next x_t, restart-mass and L1-residual partials (warp shuffle -> block).
//   k_fixup  rows cut by a chunk boundary (<= 1 per chunk) and the seed row (+S); last block reduces the partials
//            in a fixed order -> next S, residual, convergence flag.  Deterministic: no floating-point atomics.
#include <cmath>

#include "iterate.h"

constexpr int GROUPS = 4;
constexpr int CTA_THREADS = GROUPS * GROUP_THREADS;       // 1024
constexpr int LONG_ROW = 64;                              // rows with >= LONG_ROW nnz inside a chunk: one warp
constexpr int LONG_CAP = CHUNK_ITEMS / LONG_ROW + 1;      // 32
constexpr int FIX_THREADS = 256;

template <typename T>
struct IterParams {
    const u32* in_ptr;
    const int32_t* in_src;
    const T* in_val;        // valued layout only
    const int2* part;
    int n_chunks;
    int n;
    const T* x;             // gather source, internal labels
    const T* inv;
    const T* r_prev;        // previous rank (residual)
    T* y;
    T* x_next;
    T omc;                  // (1 - c)
    int seed;               // internal label, -1: uniform restart
    double inv_n;           // 1/N (uniform restart)
    int hub;                // x entries staged in shared memory
    double* head_partial;   // [n_chunks] row sums are accumulated in double in both precisions
    double* carry;          // [n_chunks]
    double* slot_S;         // [main_grid + fix_grid]
    double* slot_R;
    IterCtl* ctl;
};

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void group_sync(int group) {
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(GROUP_THREADS) : "memory");
}
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src, u32 bytes, u64* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void load4_stream(const double* p, u64 pol, double out[4]) {
    out[0] = ld_stream(p, pol); out[1] = ld_stream(p + 1, pol); out[2] = ld_stream(p + 2, pol); out[3] = ld_stream(p + 3, pol);
}
__device__ __forceinline__ void load4_stream(const float* p, u64 pol, float out[4]) {
    out[0] = ld_stream(p, pol); out[1] = ld_stream(p + 1, pol); out[2] = ld_stream(p + 2, pol); out[3] = ld_stream(p + 3, pol);
}

// ------------------------------------------------------------------------------------------------ epilogue
// One finished row t with pull sum y (Model.cs:84, :91, :96-97 folded into per-row form):
//   next x_t = fl(fl((1-c) y) * inv_t);  restart mass += inv_t == 0 ? y : y - fl((1-c) y);  residual += |r_t - y|
template <typename T, bool WRITE_Y, bool RESID>
__device__ __forceinline__ void finalize_row(const IterParams<T>& p, int row, T y, double uni_add, double& accS, double& accR) {
    if (p.seed < 0) y = add_rn(y, (T)uni_add);
    const T invr = p.inv[row];
    if (WRITE_Y) p.y[row] = y;
    const T rw = mul_rn(p.omc, y);
    p.x_next[row] = mul_rn(rw, invr);
    accS += (invr == (T)0) ? (double)y : (double)sub_rn(y, rw);
    if (RESID) {
        const T rp = p.r_prev[row];
        accR += (double)((rp > y) ? sub_rn(rp, y) : sub_rn(y, rp));
    }
}

// fixed-order block reduction of two doubles; result valid in thread 0
template <int THREADS>
__device__ __forceinline__ void block_sum2(double& a, double& b, double* scratch /* 2 * THREADS/32 */) {
    a = warp_sum(a);
    b = warp_sum(b);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { scratch[warp] = a; scratch[THREADS / 32 + warp] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sa = 0, sb = 0;
        for (int w = 0; w < THREADS / 32; w++) { sa += scratch[w]; sb += scratch[THREADS / 32 + w]; }
        a = sa; b = sb;
    }
}

// ------------------------------------------------------------------------------------------------ K7: SpMV
template <typename T, bool VALUED, bool WRITE_Y, bool RESID>
__global__ void __launch_bounds__(CTA_THREADS, 1) k_spmv(const IterParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (p.ctl->done) return;
    u64* bar = reinterpret_cast<u64*>(smem_raw);
    double* scratch = reinterpret_cast<double*>(smem_raw + 16);                   // 64 doubles
    int* long_list = reinterpret_cast<int*>(smem_raw + 16 + 512);                  // GROUPS * (LONG_CAP + 1)
    T* prod_base = reinterpret_cast<T*>(smem_raw + 16 + 512 + 1024);
    T* hub = prod_base + GROUPS * CHUNK_SPAN;

    const int group = threadIdx.x / GROUP_THREADS, gtid = threadIdx.x % GROUP_THREADS;
    const int lane = threadIdx.x & 31, gwarp = gtid >> 5;
    T* prod = prod_base + group * CHUNK_SPAN;
    int* llist = long_list + group * (LONG_CAP + 1);
    int* lcount = llist + LONG_CAP;

    // ---- stage the hot prefix of x in shared memory with TMA bulk copies
    if (p.hub > 0) {
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            const u32 bytes = (u32)p.hub * (u32)sizeof(T);
            mbar_expect_tx(bar, bytes);
            for (u32 off = 0; off < bytes; off += 32768) {
                u32 len = bytes - off < 32768 ? bytes - off : 32768;
                tma_load_1d(reinterpret_cast<unsigned char*>(hub) + off, reinterpret_cast<const unsigned char*>(p.x) + off, len, bar);
            }
        }
    }
    if (gtid == 0) *lcount = 0;
    const double uni_add = (p.seed < 0) ? p.ctl->S * p.inv_n : 0.0;
    const u64 pol_stream = policy_evict_first(), pol_keep = policy_evict_last();
    double accS = 0.0, accR = 0.0;
    if (p.hub > 0) mbar_wait(bar, 0);
    __syncthreads();

    const int n_groups = gridDim.x * GROUPS;
    for (int chunk = blockIdx.x * GROUPS + group; chunk < p.n_chunks; chunk += n_groups) {
        const int2 c0 = p.part[chunk], c1 = p.part[chunk + 1];
        const u32 nnz0 = (u32)c0.y, nnz1 = (u32)c1.y;
        const int row0 = c0.x, row1 = c1.x;
        const u32 base = nnz0 & ~3u;

        // ---- phase 1: coalesced index loads, gathers, products -> shared memory
        int4 iv[CHUNK_ROUNDS];
#pragma unroll
        for (int j = 0; j < CHUNK_ROUNDS; j++) {
            const u32 pos = base + (u32)(j * GROUP_THREADS + gtid) * 4u;
            iv[j] = make_int4(0, 0, 0, 0);
            if (pos < nnz1) iv[j] = ld_stream_int4(reinterpret_cast<const int4*>(p.in_src + pos), pol_stream);
        }
#pragma unroll
        for (int j = 0; j < CHUNK_ROUNDS; j++) {
            const u32 pos = base + (u32)(j * GROUP_THREADS + gtid) * 4u;
            if (pos < nnz1) {
                const int s4[4] = {iv[j].x, iv[j].y, iv[j].z, iv[j].w};
                T v[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const u32 e = pos + k;
                    v[k] = (T)0;
                    if (e >= nnz0 && e < nnz1) v[k] = (s4[k] < p.hub) ? hub[s4[k]] : ld_keep(p.x + s4[k], pol_keep);
                }
                if (VALUED) {
                    T w[4];
                    load4_stream(p.in_val + pos, pol_stream, w);
#pragma unroll
                    for (int k = 0; k < 4; k++) v[k] = mul_rn(v[k], w[k]);
                }
                T* dst = prod + (pos - base);
#pragma unroll
                for (int k = 0; k < 4; k++) dst[k] = v[k];
            }
        }
        group_sync(group);

        // ---- phase 2: one thread per row, products summed in storage order
        for (int r = row0 + gtid; r <= row1 && r < p.n; r += GROUP_THREADS) {
            const u32 rs = p.in_ptr[r];
            const bool complete = r < row1;
            const u32 s = rs > nnz0 ? rs : nnz0;
            u32 e = complete ? p.in_ptr[r + 1] : nnz1;
            if (e < s) e = s;
            if (e - s >= (u32)LONG_ROW) {
                llist[atomicAdd(lcount, 1)] = r;
                continue;
            }
            double sum = 0.0;
            for (u32 q = s - base; q < e - base; q++) sum = __dadd_rn(sum, (double)prod[q]);
            if (!complete) p.carry[chunk] = sum;
            else if (rs < nnz0) p.head_partial[chunk] = sum;
            else if (r == p.seed) { p.ctl->seed_sum = sum; p.ctl->seed_flag = 1; }
            else finalize_row<T, WRITE_Y, RESID>(p, r, (T)sum, uni_add, accS, accR);
        }
        group_sync(group);

        // ---- phase 2b: long rows, one warp each (lane-strided partials, fixed shuffle tree)
        const int n_long = *lcount;
        for (int li = gwarp; li < n_long; li += GROUP_THREADS / 32) {
            const int r = llist[li];
            const u32 rs = p.in_ptr[r];
            const bool complete = r < row1;
            const u32 s = rs > nnz0 ? rs : nnz0;
            const u32 e = complete ? p.in_ptr[r + 1] : nnz1;
            double part = 0.0;
            for (u32 q = s - base + lane; q < e - base; q += 32) part = __dadd_rn(part, (double)prod[q]);
            const double sum = warp_sum_down<double>(part);
            if (lane == 0) {
                if (!complete) p.carry[chunk] = sum;
                else if (rs < nnz0) p.head_partial[chunk] = sum;
                else if (r == p.seed) { p.ctl->seed_sum = sum; p.ctl->seed_flag = 1; }
                else finalize_row<T, WRITE_Y, RESID>(p, r, (T)sum, uni_add, accS, accR);
            }
        }
        group_sync(group);
        if (gtid == 0) *lcount = 0;
    }

    __syncthreads();
    block_sum2<CTA_THREADS>(accS, accR, scratch);
    if (threadIdx.x == 0) {
        p.slot_S[blockIdx.x] = accS;
        p.slot_R[blockIdx.x] = accR;
    }
}

// ------------------------------------------------------------------------------------------------ fix-up
template <typename T, bool WRITE_Y, bool RESID>
__global__ void __launch_bounds__(FIX_THREADS) k_fixup(const IterParams<T> p, int main_grid, double thr, int use_thr) {
    __shared__ double scratch[2 * FIX_THREADS / 32];
    __shared__ int is_last;
    IterCtl* ctl = p.ctl;
    if (ctl->done) return;
    const double S = ctl->S;
    const double uni_add = (p.seed < 0) ? S * p.inv_n : 0.0;
    double accS = 0.0, accR = 0.0;
    const int k = blockIdx.x * FIX_THREADS + threadIdx.x;
    if (k < p.n_chunks) {
        const int2 c0 = p.part[k], c1 = p.part[k + 1];
        if (c0.x < c1.x && p.in_ptr[c0.x] < (u32)c0.y) {       // first row of the chunk started in an earlier chunk
            const int row = c0.x;
            int m0 = k - 1;
            while (m0 > 0 && p.part[m0].x == row) m0--;
            double total = 0.0;
            for (int m = m0; m < k; m++) total = __dadd_rn(total, p.carry[m]);
            total = __dadd_rn(total, p.head_partial[k]);
            if (row == p.seed) total = __dadd_rn(total, S);
            finalize_row<T, WRITE_Y, RESID>(p, row, (T)total, uni_add, accS, accR);
        }
    }
    if (k == 0 && ctl->seed_flag) {                            // seed row finished inside one chunk
        const T y = (T)__dadd_rn(ctl->seed_sum, S);
        finalize_row<T, WRITE_Y, RESID>(p, p.seed, y, uni_add, accS, accR);
        ctl->seed_flag = 0;
    }
    block_sum2<FIX_THREADS>(accS, accR, scratch);
    if (threadIdx.x == 0) {
        p.slot_S[main_grid + blockIdx.x] = accS;
        p.slot_R[main_grid + blockIdx.x] = accR;
        __threadfence();
        const unsigned t = atomicAdd(&ctl->ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        const int total = main_grid + (int)gridDim.x;
        double a = 0.0, b = 0.0;
        for (int i = threadIdx.x; i < total; i += FIX_THREADS) {
            a += __ldcg(p.slot_S + i);
            b += __ldcg(p.slot_R + i);
        }
        __syncthreads();
        block_sum2<FIX_THREADS>(a, b, scratch);
        if (threadIdx.x == 0) {
            ctl->S = a;
            ctl->resid = b;
            ctl->iters += 1;
            ctl->ticket = 0;
            if (use_thr && b < thr) ctl->done = 1;            // strict `<` (Model.cs:114)
        }
    }
}

// ------------------------------------------------------------------------------------------------ K6: init
template <typename T>
__global__ void k_init(int n, int seed, T omc, const T* __restrict__ inv, T* __restrict__ r0, T* __restrict__ x0,
                       IterCtl* ctl, double S_uniform) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) {
        ctl->resid = 0.0; ctl->seed_sum = 0.0; ctl->seed_flag = 0; ctl->done = 0; ctl->iters = 0; ctl->ticket = 0;
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

template <typename T>
__global__ void k_unpermute(const T* __restrict__ y_int, const int32_t* __restrict__ new_of_old, int n, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)y_int[new_of_old[i]];
}

// ------------------------------------------------------------------------------------------------ host side
static size_t smem_fixed_bytes(size_t elt) { return 16 + 512 + 1024 + (size_t)GROUPS * CHUNK_SPAN * elt; }

int hub_entries_for(const rwr_graph* g, int precision) {
    const size_t elt = precision == RWR_FP32 ? 4 : 8;
    const size_t fixed = smem_fixed_bytes(elt);
    if ((size_t)g->max_smem_optin <= fixed) return 0;
    long cap = (long)(((size_t)g->max_smem_optin - fixed) / elt) & ~3L;
    long want = g->opts.hub_entries < 0 ? cap : std::min<long>(cap, (long)g->opts.hub_entries & ~3L);
    long n4 = ((long)g->n + 3) & ~3L;
    return (int)std::max<long>(0, std::min(want, n4));
}

void iterate_prepare(rwr_graph* g) {
    cudaStream_t st = g->stream;
    const u64 total = (u64)g->n + (u64)g->nnz;
    g->n_chunks = (int)std::max<u64>(1, (total + CHUNK_ITEMS - 1) / CHUNK_ITEMS);
    g->part.alloc((size_t)g->n_chunks + 1, &g->pool);
    k_partition<<<div_up((size_t)g->n_chunks + 1, 256), 256, 0, st>>>(g->in_ptr.p, g->n, (u32)g->nnz, g->n_chunks, g->part.p);
    KERNEL_CHECK();
    CUDA_CHECK(cudaStreamSynchronize(st));
}

static void ensure_fp32(rwr_graph* g) {
    cudaStream_t st = g->stream;
    if (!g->inv32.p) {
        g->inv32.alloc((size_t)g->n + 4, &g->pool);
        if (g->n) k_f64_to_f32<<<div_up(g->n, 256), 256, 0, st>>>(g->inv64.p, g->inv32.p, g->n);
        KERNEL_CHECK();
    }
    if (g->layout == RWR_LAYOUT_VALUED && !g->in_val32.p) {
        g->in_val32.alloc((size_t)g->nnz + IDX_PAD, &g->pool);
        if (g->nnz) k_f64_to_f32<<<div_up((size_t)g->nnz, 256), 256, 0, st>>>(g->in_val64.p, g->in_val32.p, (size_t)g->nnz);
        KERNEL_CHECK();
    }
}

template <typename T> struct Prec;
template <> struct Prec<double> {
    static const double* inv(rwr_graph* g) { return g->inv64.p; }
    static const double* val(rwr_graph* g) { return g->in_val64.p; }
    static DevBuf<double>& ybuf(rwr_result* r) { return r->y64; }
    static constexpr int id = RWR_FP64;
};
template <> struct Prec<float> {
    static const float* inv(rwr_graph* g) { return g->inv32.p; }
    static const float* val(rwr_graph* g) { return g->in_val32.p; }
    static DevBuf<float>& ybuf(rwr_result* r) { return r->y32; }
    static constexpr int id = RWR_FP32;
};

template <typename T, bool VALUED, bool WRITE_Y, bool RESID>
static void launch_spmv(const IterParams<T>& p, int grid, size_t smem, cudaStream_t st) {
    auto kern = k_spmv<T, VALUED, WRITE_Y, RESID>;
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, CTA_THREADS, smem, st>>>(p);
    KERNEL_CHECK();
}

template <typename T>
static void launch_iteration(rwr_graph* g, const IterParams<T>& p, bool write_y, bool resid, int main_grid, int fix_grid,
                             size_t smem, double thr, int use_thr) {
    cudaStream_t st = g->stream;
    const bool valued = g->layout == RWR_LAYOUT_VALUED;
#define LAUNCH(V, W, R)                                                                  \
    do {                                                                                 \
        launch_spmv<T, V, W, R>(p, main_grid, smem, st);                                 \
        k_fixup<T, W, R><<<fix_grid, FIX_THREADS, 0, st>>>(p, main_grid, thr, use_thr);  \
    } while (0)
    if (resid) {
        if (valued) LAUNCH(true, true, true); else LAUNCH(false, true, true);
    } else if (write_y) {
        if (valued) LAUNCH(true, true, false); else LAUNCH(false, true, false);
    } else {
        if (valued) LAUNCH(true, false, false); else LAUNCH(false, false, false);
    }
#undef LAUNCH
    KERNEL_CHECK();
    g->pool.launches += 2;
}

struct RunWorkspace {
    DevBuf<unsigned char> xa, xb, ya;
    DevBuf<double> carry, head, slot_S, slot_R;
    DevBuf<IterCtl> ctl;
};

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
    const int main_grid = std::max(1, std::min(g->sm_count, (g->n_chunks + GROUPS - 1) / GROUPS));
    const int fix_grid = div_up((size_t)g->n_chunks, FIX_THREADS);
    const int hub = hub_entries_for(g, Prec<T>::id);
    const size_t smem = smem_fixed_bytes(sizeof(T)) + (size_t)hub * sizeof(T);

    T* xa = reinterpret_cast<T*>(ws.xa.p);
    T* xb = reinterpret_cast<T*>(ws.xb.p);
    T* ya = reinterpret_cast<T*>(ws.ya.p);
    IterParams<T> p;
    p.in_ptr = g->in_ptr.p; p.in_src = g->in_src.p; p.in_val = Prec<T>::val(g); p.part = g->part.p;
    p.n_chunks = g->n_chunks; p.n = n; p.inv = Prec<T>::inv(g);
    p.omc = (T)(1.0 - c);                                      // Model.cs:84 `(1 - dampingFactor)`
    p.seed = seed_int; p.inv_n = n ? 1.0 / (double)n : 0.0; p.hub = hub;
    p.head_partial = ws.head.p; p.carry = ws.carry.p;
    p.slot_S = ws.slot_S.p; p.slot_R = ws.slot_R.p; p.ctl = ws.ctl.p;

    // uniform constructor: S0 = sum over nodes of (dangling ? 1 : 1 - fl((1-c)*1))
    const double omc_d = (double)p.omc;
    const double S_uniform = (double)g->n_dangling + (double)(n - g->n_dangling) * (1.0 - omc_d);
    // r0 goes to y_out in fixed mode (also the answer for n_iter == 0); threshold mode ping-pongs ya <-> y_out
    T* r_cur = (mode == 0) ? y_out : ya;
    k_init<T><<<div_up(std::max(n, 1), 256), 256, 0, st>>>(n, seed_int, p.omc, p.inv, r_cur, xa, ws.ctl.p, S_uniform);
    KERNEL_CHECK();
    g->pool.launches += 1;

    CUDA_CHECK(cudaEventRecord(ev0, st));
    T* x_cur = xa;
    T* x_nxt = xb;
    int launched = 0;
    if (mode == 0) {
        for (int it = 0; it < n_iter; it++) {
            p.x = x_cur; p.x_next = x_nxt; p.r_prev = nullptr; p.y = y_out;
            launch_iteration<T>(g, p, /*write_y=*/it == n_iter - 1, /*resid=*/false, main_grid, fix_grid, smem, 0.0, 0);
            std::swap(x_cur, x_nxt);
            launched++;
        }
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
                launch_iteration<T>(g, p, true, true, main_grid, fix_grid, smem, thr, 1);
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
    CUDA_CHECK(cudaEventRecord(ev1, st));
    CUDA_CHECK(cudaEventSynchronize(ev1));
    float ms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, ev0, ev1));
    *iter_ms += ms;
}

template <typename T>
static void run_all(rwr_graph* g, rwr_result* res, const int32_t* seeds, int n_seeds, double c, int mode, int n_iter,
                    double thr, int max_iter, int32_t* iters_out) {
    cudaStream_t st = g->stream;
    const size_t n = (size_t)g->n;
    if (Prec<T>::id == RWR_FP32) ensure_fp32(g);
    const size_t ld = (n + 3) & ~(size_t)3;
    res->ld = ld;
    Prec<T>::ybuf(res).alloc(std::max<size_t>(1, ld * (size_t)n_seeds), nullptr);
    RunWorkspace ws;
    const size_t vec_bytes = (n + 8) * sizeof(T);
    ws.xa.alloc(vec_bytes); ws.xb.alloc(vec_bytes); ws.ya.alloc(vec_bytes);
    CUDA_CHECK(cudaMemsetAsync(ws.xa.p, 0, vec_bytes, st));
    CUDA_CHECK(cudaMemsetAsync(ws.xb.p, 0, vec_bytes, st));
    ws.carry.alloc((size_t)g->n_chunks); ws.head.alloc((size_t)g->n_chunks);
    CUDA_CHECK(cudaMemsetAsync(ws.carry.p, 0, (size_t)g->n_chunks * sizeof(double), st));
    CUDA_CHECK(cudaMemsetAsync(ws.head.p, 0, (size_t)g->n_chunks * sizeof(double), st));
    const size_t slots = (size_t)g->sm_count + div_up((size_t)g->n_chunks, FIX_THREADS) + 8;
    ws.slot_S.alloc(slots); ws.slot_R.alloc(slots);
    CUDA_CHECK(cudaMemsetAsync(ws.slot_S.p, 0, slots * sizeof(double), st));
    CUDA_CHECK(cudaMemsetAsync(ws.slot_R.p, 0, slots * sizeof(double), st));
    ws.ctl.alloc(1);
    CUDA_CHECK(cudaMemsetAsync(ws.ctl.p, 0, sizeof(IterCtl), st));
    cudaEvent_t ev0, ev1, evA, evB;
    CUDA_CHECK(cudaEventCreate(&ev0)); CUDA_CHECK(cudaEventCreate(&ev1));
    CUDA_CHECK(cudaEventCreate(&evA)); CUDA_CHECK(cudaEventCreate(&evB));
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
    cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(evA); cudaEventDestroy(evB);
}

template <typename T>
static void profile_impl(rwr_graph* g, int seed_orig, double c, int reps, float* spmv_ms, float* fixup_ms) {
    cudaStream_t st = g->stream;
    const size_t n = (size_t)g->n;
    if (Prec<T>::id == RWR_FP32) ensure_fp32(g);
    RunWorkspace ws;
    const size_t vec_bytes = (n + 8) * sizeof(T);
    ws.xa.alloc(vec_bytes); ws.xb.alloc(vec_bytes); ws.ya.alloc(vec_bytes);
    CUDA_CHECK(cudaMemsetAsync(ws.xa.p, 0, vec_bytes, st));
    CUDA_CHECK(cudaMemsetAsync(ws.xb.p, 0, vec_bytes, st));
    ws.carry.alloc((size_t)g->n_chunks); ws.head.alloc((size_t)g->n_chunks);
    CUDA_CHECK(cudaMemsetAsync(ws.carry.p, 0, (size_t)g->n_chunks * sizeof(double), st));
    CUDA_CHECK(cudaMemsetAsync(ws.head.p, 0, (size_t)g->n_chunks * sizeof(double), st));
    const size_t slots = (size_t)g->sm_count + div_up((size_t)g->n_chunks, FIX_THREADS) + 8;
    ws.slot_S.alloc(slots); ws.slot_R.alloc(slots);
    CUDA_CHECK(cudaMemsetAsync(ws.slot_S.p, 0, slots * sizeof(double), st));
    CUDA_CHECK(cudaMemsetAsync(ws.slot_R.p, 0, slots * sizeof(double), st));
    ws.ctl.alloc(1);
    int seed_int = 0;
    CUDA_CHECK(cudaMemcpyAsync(&seed_int, g->new_of_old.p + seed_orig, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    const int main_grid = std::max(1, std::min(g->sm_count, (g->n_chunks + GROUPS - 1) / GROUPS));
    const int fix_grid = div_up((size_t)g->n_chunks, FIX_THREADS);
    const int hub = hub_entries_for(g, Prec<T>::id);
    const size_t smem = smem_fixed_bytes(sizeof(T)) + (size_t)hub * sizeof(T);
    T* xa = reinterpret_cast<T*>(ws.xa.p);
    T* xb = reinterpret_cast<T*>(ws.xb.p);
    T* ya = reinterpret_cast<T*>(ws.ya.p);
    IterParams<T> p;
    p.in_ptr = g->in_ptr.p; p.in_src = g->in_src.p; p.in_val = Prec<T>::val(g); p.part = g->part.p;
    p.n_chunks = g->n_chunks; p.n = g->n; p.inv = Prec<T>::inv(g);
    p.omc = (T)(1.0 - c); p.seed = seed_int; p.inv_n = n ? 1.0 / (double)n : 0.0; p.hub = hub;
    p.head_partial = ws.head.p; p.carry = ws.carry.p; p.slot_S = ws.slot_S.p; p.slot_R = ws.slot_R.p; p.ctl = ws.ctl.p;
    p.r_prev = nullptr; p.y = ya;
    k_init<T><<<div_up(std::max((int)n, 1), 256), 256, 0, st>>>((int)n, seed_int, p.omc, p.inv, ya, xa, ws.ctl.p, 0.0);
    KERNEL_CHECK();
    const bool valued = g->layout == RWR_LAYOUT_VALUED;
    std::vector<cudaEvent_t> ev(3 * (size_t)reps);
    for (auto& e : ev) CUDA_CHECK(cudaEventCreate(&e));
    T* x_cur = xa;
    T* x_nxt = xb;
    for (int it = 0; it < 3 + reps; it++) {
        p.x = x_cur; p.x_next = x_nxt;
        const int r = it - 3;
        if (r >= 0) CUDA_CHECK(cudaEventRecord(ev[3 * r], st));
        if (valued) launch_spmv<T, true, false, false>(p, main_grid, smem, st);
        else launch_spmv<T, false, false, false>(p, main_grid, smem, st);
        if (r >= 0) CUDA_CHECK(cudaEventRecord(ev[3 * r + 1], st));
        k_fixup<T, false, false><<<fix_grid, FIX_THREADS, 0, st>>>(p, main_grid, 0.0, 0);
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
    for (auto& e : ev) cudaEventDestroy(e);
    *spmv_ms = (float)(a / reps);
    *fixup_ms = (float)(b / reps);
    g->pool.launches += 1 + 2 * (3 + reps);
}

static int run_entry(rwr_graph* g, const int32_t* seeds, int32_t n_seeds, double c, int mode, int32_t n_iter, double thr,
                     int32_t max_iter, int32_t precision, int32_t* iters_out, rwr_result** out) {
    rwr_result* res = nullptr;
    try {
        if (!out) RWR_FAIL(RWR_E_INVALID, "out is NULL");
        *out = nullptr;
        if (!g) RWR_FAIL(RWR_E_INVALID, "graph is NULL");
        if (!g->built) RWR_FAIL(RWR_E_NOT_BUILT, "buildGraph() has not run (KeyNotFoundException at Model.cs:79)");
        if (g->comm) RWR_FAIL(RWR_E_UNSUPPORTED, "row-partitioned graphs run through rwr_run_fixed_partitioned");
        if (n_seeds < 0 || (n_seeds && !seeds)) RWR_FAIL(RWR_E_INVALID, "bad seed list");
        if (mode == 0 && n_iter < 0) n_iter = 0;                     // `for (n = 0; n < nIterations; ..)` runs zero times
        if (precision != RWR_FP64 && precision != RWR_FP32) RWR_FAIL(RWR_E_INVALID, "unknown precision %d", precision);
        if (!(c == c)) RWR_FAIL(RWR_E_INVALID, "c is NaN");
        for (int s = 0; s < n_seeds; s++)
            if (seeds[s] < -1 || seeds[s] >= g->n) RWR_FAIL(RWR_E_BADSEED, "seed %d outside [0, %d)", seeds[s], g->n);
        CUDA_CHECK(cudaSetDevice(g->device));
        if (mode == 1 && !(thr > 0.0)) thr = (1.0 / 1.7976931348623157e308) * (double)g->n;   // Model.cs:53
        res = new rwr_result();
        res->g = g;
        res->device = g->device;
        res->n_seeds = n_seeds;
        res->precision = precision;
        res->seeds.assign(seeds, seeds + n_seeds);
        res->iters.assign(n_seeds, 0);
        if (precision == RWR_FP64) run_all<double>(g, res, seeds, n_seeds, c, mode, n_iter, thr, max_iter, iters_out);
        else run_all<float>(g, res, seeds, n_seeds, c, mode, n_iter, thr, max_iter, iters_out);
        *out = res;
        return RWR_OK;
    } catch (const RwrError& e) {
        delete res;
        return e.code;
    } catch (...) {
        delete res;
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
    DevBuf<double> tmp;
    tmp.alloc(n);
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

void rwr_result_destroy(rwr_result* r) {
    if (!r) return;
    cudaSetDevice(r->device);       // never dereferences r->g: the graph may already be gone
    delete r;
}

}  // extern "C"
