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

#include "iterate_dev.cuh"
#include "stream.h"
#include "dist.h"


// ------------------------------------------------------------------------------------------------ K7 (variant): pipelined SpMV, rwr_opts.kernel = 1
// Warp-specialised: per CTA 2 producer groups and 2 consumer groups of 256 threads, paired through a ring of
// stage buffers guarded by mbarriers (full / empty).  Producers do every global load of a tile -- index int4s and row
// metadata one tile ahead, gathers of x (shared-memory hub table or L2) -- and park products, row pointers and the
// rows' inv values in the stage.  Consumers read shared memory only: one thread per row sums its products in storage
// order, rows of 64..1023 nnz by a warp, longer ones by the whole group, then the fused epilogue.
// All hot shared-memory traffic uses explicit ld/st.shared with 32-bit addresses computed once: through generic
// pointers ptxas re-derived the shared window (S2UR SR_CgaCtaId) for every access, which serialised the gathers.
constexpr int PAIRS = 4;
constexpr int STAGES = 2;
constexpr int HUGE_ROW = 512;                             // rows with >= HUGE_ROW nnz inside a tile: the whole group
constexpr int STAGE_ROWS = GROUP_THREADS;                 // rows whose metadata travels through the stage
constexpr int HDR_BARS = 256, HDR_SCRATCH = 512, HDR_LISTS = 3072;
constexpr int PIPE_HDR = HDR_BARS + HDR_SCRATCH + HDR_LISTS;   // barriers | block-reduce scratch | per-group lists

template <typename T>
struct Stage {
    T prod[CHUNK_SPAN];
    T inv[STAGE_ROWS];
    u32 rp[STAGE_ROWS + 4];                               // in_ptr[row0 .. row0 + STAGE_ROWS]
    int4 coord;                                           // row0, nnz0, row1, nnz1
};
template <typename T> struct StageOff {
    static constexpr u32 inv = CHUNK_SPAN * sizeof(T);
    static constexpr u32 rp = inv + STAGE_ROWS * sizeof(T);
    static constexpr u32 coord = rp + (STAGE_ROWS + 4) * sizeof(u32);
    static constexpr u32 size = coord + 16;
};
static_assert(StageOff<double>::size == sizeof(Stage<double>) && StageOff<float>::size == sizeof(Stage<float>), "stage layout");

// sum of prod[q0 .. q1) in storage order; the shared loads are issued 8 at a time (adding +0.0 is exact)
template <typename T>
__device__ __forceinline__ double sum_run(u32 prod_addr, u32 q0, u32 q1) {
    double sum = 0.0;
    for (u32 q = q0; q < q1; q += 8) {
        T v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {                 // unconditional loads from a clamped slot, masked afterwards
            const u32 qq = (q + k < q1) ? q + k : q1 - 1;
            v[k] = lds_t(prod_addr + qq * (u32)sizeof(T), T());
        }
#pragma unroll
        for (int k = 0; k < 8; k++) sum = __dadd_rn(sum, (q + k < q1) ? (double)v[k] : 0.0);
    }
    return sum;
}
// lane/thread-strided partial: elements q0 + lane, + stride, ...; loads issued 4 at a time
template <typename T>
__device__ __forceinline__ double sum_strided(u32 prod_addr, u32 q0, u32 q1, u32 stride) {
    double sum = 0.0;
    for (u32 q = q0; q < q1; q += 4 * stride) {
        T v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const u32 qq = (q + k * stride < q1) ? q + k * stride : q;
            v[k] = lds_t(prod_addr + qq * (u32)sizeof(T), T());
        }
#pragma unroll
        for (int k = 0; k < 4; k++) sum = __dadd_rn(sum, (q + k * stride < q1) ? (double)v[k] : 0.0);
    }
    return sum;
}

// everything a producer thread prefetches for one tile
template <typename T>
struct TileLoad {
    int2 c0, c1;
    int4 iv[CHUNK_ROUNDS];
    u32 rp, rp_last;
    T inv;
};

template <typename T>
__device__ __forceinline__ void load_tile(const IterParams<T>& p, int chunk, int gtid, u64 pol, TileLoad<T>& t) {
    t.c0 = p.part[chunk];
    t.c1 = p.part[chunk + 1];
    const u32 nnz1 = (u32)t.c1.y, base = (u32)t.c0.y & ~3u;
#pragma unroll
    for (int j = 0; j < CHUNK_ROUNDS; j++) {
        const u32 pos = base + (u32)(j * GROUP_THREADS + gtid) * 4u;
        t.iv[j] = make_int4(0, 0, 0, 0);
        if (pos < nnz1) t.iv[j] = ld_stream_int4(reinterpret_cast<const int4*>(p.in_src + pos), pol);
    }
    const int r = t.c0.x + gtid;
    t.rp = 0; t.rp_last = 0; t.inv = (T)0;
    if (r <= p.n) t.rp = p.in_ptr[r];
    if (r < p.n) t.inv = p.inv[r];
    if (gtid == 0 && t.c0.x + STAGE_ROWS <= p.n) t.rp_last = p.in_ptr[t.c0.x + STAGE_ROWS];
}

#ifdef RWR_PROFILE_CLOCKS
__device__ unsigned long long g_clk[16];
#define CLK_DECL long long clk_t0 = clock64(), clk_t1
#define CLK_ADD(slot) do { clk_t1 = clock64(); if (gtid == 0) atomicAdd(&g_clk[slot], (unsigned long long)(clk_t1 - clk_t0)); clk_t0 = clk_t1; } while (0)
#else
#define CLK_DECL
#define CLK_ADD(slot)
#endif

template <typename T, bool VALUED, bool WRITE_Y, bool RESID>
__global__ void __launch_bounds__(CTA_THREADS, 1) k_spmv(const IterParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (p.ctl->done) return;
    double* scratch = reinterpret_cast<double*>(smem_raw + HDR_BARS);
    // opaque to ptxas on purpose: a plain cvta result is rematerialised (S2UR SR_CgaCtaId + ULEA) at every use
    u32 smem0;
    asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(smem0) : "l"(smem_raw));
    const u32 bars = smem0;                                                    // [0] hub, then full/empty pairs
    const u32 stages = smem0 + PIPE_HDR;
    const u32 hub_addr = stages + PAIRS * STAGES * StageOff<T>::size;

    const int group = threadIdx.x / GROUP_THREADS, gtid = threadIdx.x % GROUP_THREADS;
    const int lane = threadIdx.x & 31, gwarp = gtid >> 5;

    if (threadIdx.x == 0) {
        mbar_init(reinterpret_cast<u64*>(smem_raw), 1);
        for (int i = 0; i < PAIRS * STAGES; i++) {
            mbar_init(reinterpret_cast<u64*>(smem_raw) + 1 + 2 * i, GROUP_THREADS);
            mbar_init(reinterpret_cast<u64*>(smem_raw) + 2 + 2 * i, GROUP_THREADS);
        }
    }
    __syncthreads();
    if (p.hub > 0 && threadIdx.x == 0) {
        const u32 bytes = (u32)p.hub * (u32)sizeof(T);
        mbar_expect_tx(reinterpret_cast<u64*>(smem_raw), bytes);
        for (u32 off = 0; off < bytes; off += 32768) {
            u32 len = bytes - off < 32768 ? bytes - off : 32768;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(hub_addr + off),
                         "l"(reinterpret_cast<const unsigned char*>(p.x) + off), "r"(len), "r"(bars)
                         : "memory");
        }
    }
    const int stride = gridDim.x * PAIRS;
    double accS = 0.0, accR = 0.0;

    if (group < PAIRS) {
        // ================================================================== producers
        const int pair = group;
        const u64 pol_stream = policy_evict_first(), pol_keep = policy_evict_last();
        int chunk = blockIdx.x * PAIRS + pair;
        TileLoad<T> cur;
        if (chunk < p.n_chunks) load_tile<T>(p, chunk, gtid, pol_stream, cur);
        if (p.hub > 0) mbar_wait_a(bars, 0);
        CLK_DECL;
        for (int t = 0; chunk < p.n_chunks; t++, chunk += stride) {
            const int stage = t % STAGES;
            const u32 par = (u32)(t / STAGES) & 1u;
            const u32 full_b = bars + 8u * (1 + (pair * STAGES + stage) * 2), empty_b = full_b + 8;
            const u32 st = stages + (u32)(pair * STAGES + stage) * StageOff<T>::size;
            const u32 nnz0 = (u32)cur.c0.y, nnz1 = (u32)cur.c1.y, base = nnz0 & ~3u;
            // ---- gathers of this tile (its indices arrived during the previous tile)
            T v[CHUNK_ROUNDS][4];
            int take_hub[CHUNK_ROUNDS][4], take_glob[CHUNK_ROUNDS][4];
#pragma unroll
            for (int j = 0; j < CHUNK_ROUNDS; j++) {          // every predicate first: they only need the index registers
                const u32 pos = base + (u32)(j * GROUP_THREADS + gtid) * 4u;
                const int s4[4] = {cur.iv[j].x, cur.iv[j].y, cur.iv[j].z, cur.iv[j].w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const u32 e = pos + k;
                    const bool valid = e >= nnz0 && e < nnz1;
                    take_hub[j][k] = valid && s4[k] < p.hub;
                    take_glob[j][k] = valid && s4[k] >= p.hub;
                }
            }
#pragma unroll
            for (int j = 0; j < CHUNK_ROUNDS; j++) {
                const int s4[4] = {cur.iv[j].x, cur.iv[j].y, cur.iv[j].z, cur.iv[j].w};
#pragma unroll
                for (int k = 0; k < 4; k++)
                    v[j][k] = gather_sel(take_hub[j][k], take_glob[j][k], hub_addr + (u32)s4[k] * (u32)sizeof(T), p.x + s4[k],
                                         s4[k] < p.n_hot ? pol_keep : pol_stream);
            }
            if (VALUED) {
#pragma unroll
                for (int j = 0; j < CHUNK_ROUNDS; j++) {
                    const u32 pos = base + (u32)(j * GROUP_THREADS + gtid) * 4u;
                    if (pos < nnz1) {
                        T w[4];
                        load4_stream(p.in_val + pos, pol_stream, w);
#pragma unroll
                        for (int k = 0; k < 4; k++) v[j][k] = mul_rn(v[j][k], w[k]);
                    }
                }
            }
            CLK_ADD(1);                                       // producer: issuing gathers
            const int4 coord = make_int4(cur.c0.x, cur.c0.y, cur.c1.x, cur.c1.y);
            const u32 my_rp = cur.rp, my_rp_last = cur.rp_last;
            const T my_inv = cur.inv;
            // ---- prefetch everything of the next tile
            const int next = chunk + stride;
            if (next < p.n_chunks) load_tile<T>(p, next, gtid, pol_stream, cur);
            CLK_ADD(2);                                       // producer: issuing the next tile's loads (part latency)
            // ---- fill the stage once the consumer has released it
            mbar_wait_a(empty_b, par ^ 1u);
            CLK_ADD(3);                                       // producer: waiting for the consumer
#pragma unroll
            for (int j = 0; j < CHUNK_ROUNDS; j++) {
                const u32 pos = base + (u32)(j * GROUP_THREADS + gtid) * 4u;
                if (pos < nnz1) sts4(st + (pos - base) * (u32)sizeof(T), v[j]);
            }
            sts_u32(st + StageOff<T>::rp + 4u * gtid, my_rp);
            sts_t(st + StageOff<T>::inv + (u32)sizeof(T) * gtid, my_inv);
            if (gtid == 0) {
                sts_u32(st + StageOff<T>::rp + 4u * STAGE_ROWS, my_rp_last);
                sts_int4(st + StageOff<T>::coord, coord);
            }
            mbar_arrive_a(full_b);
            CLK_ADD(4);                                       // producer: waiting for gathers + stores
        }
    } else {
        // ================================================================== consumers
        const int pair = group - PAIRS;
        // two copies of the long-row list, used by alternate tiles: the copy of tile t+1 is reset during tile t,
        // between two group barriers, so no reset can race with a reader or an appender
        int* lists2 = reinterpret_cast<int*>(smem_raw + HDR_BARS + HDR_SCRATCH + pair * 512);   // 2 x (LONG_CAP ids, count, huge)
        double* hpart = reinterpret_cast<double*>(smem_raw + HDR_BARS + HDR_SCRATCH + pair * 512 + 320);   // warp partials
        const double uni_add = (p.seed < 0) ? p.ctl->S * p.inv_n : 0.0;
        if (gtid < 2) { lists2[gtid * (LONG_CAP + 2) + LONG_CAP] = 0; lists2[gtid * (LONG_CAP + 2) + LONG_CAP + 1] = -1; }
        group_sync(group);
        int chunk = blockIdx.x * PAIRS + pair;
        CLK_DECL;
        for (int t = 0; chunk < p.n_chunks; t++, chunk += stride) {
            const int stage = t % STAGES;
            const u32 par = (u32)(t / STAGES) & 1u;
            const u32 full_b = bars + 8u * (1 + (pair * STAGES + stage) * 2), empty_b = full_b + 8;
            const u32 st = stages + (u32)(pair * STAGES + stage) * StageOff<T>::size;
            int* llist = lists2 + (t & 1) * (LONG_CAP + 2);
            int* lcount = llist + LONG_CAP;
            int* huge = lcount + 1;
            mbar_wait_a(full_b, par);
            CLK_ADD(8);                                       // consumer: waiting for the producer
            const int4 coord = lds_int4(st + StageOff<T>::coord);
            const int row0 = coord.x, row1 = coord.z;
            const u32 nnz0 = (u32)coord.y, nnz1 = (u32)coord.w, base = nnz0 & ~3u;
            // row r of the tile: pointers and inv come from the stage for the first STAGE_ROWS rows
            auto row_begin = [&](int r) -> u32 {
                return (r - row0 <= STAGE_ROWS) ? lds_u32(st + StageOff<T>::rp + 4u * (u32)(r - row0)) : p.in_ptr[r];
            };
            auto row_inv = [&](int r) -> T {
                return (r - row0 < STAGE_ROWS) ? lds_t(st + StageOff<T>::inv + (u32)sizeof(T) * (u32)(r - row0), T()) : p.inv[r];
            };
            for (int r = row0 + gtid; r <= row1 && r < p.n; r += GROUP_THREADS) {
                const bool complete = r < row1;
                const u32 rs = row_begin(r);
                const u32 s = rs > nnz0 ? rs : nnz0;
                u32 e = complete ? row_begin(r + 1) : nnz1;
                if (e < s) e = s;
                const u32 len = e - s;
                if (len >= (u32)HUGE_ROW) {
                    *huge = r;
                } else if (len >= (u32)LONG_ROW) {
                    llist[atomicAdd(lcount, 1)] = r;
                } else {
                    const double sum = sum_run<T>(st, s - base, e - base);
                    if (!complete) p.carry[chunk] = sum;
                    else if (rs < nnz0) p.head_partial[chunk] = sum;
                    else if (r == p.seed) { p.ctl->seed_sum = sum; p.ctl->seed_flag = 1; }
                    else finalize_row<T, WRITE_Y, RESID>(p, r, (T)sum, row_inv(r), uni_add, accS, accR);
                }
            }
            CLK_ADD(9);                                       // consumer: own short rows
            group_sync(group);
            CLK_ADD(10);                                      // consumer: barrier after the row loop
            if (gtid == 0) {                               // reset the other copy for the next tile
                int* other = lists2 + ((t + 1) & 1) * (LONG_CAP + 2);
                other[LONG_CAP] = 0;
                other[LONG_CAP + 1] = -1;
            }
            // ---- rows of 64..1023 nnz: one warp each (lane-strided partials, fixed shuffle tree)
            const int n_long = *lcount;
            for (int li = gwarp; li < n_long; li += GROUP_THREADS / 32) {
                const int lr = llist[li];
                const u32 lrs = row_begin(lr);
                const bool complete = lr < row1;
                const u32 s = lrs > nnz0 ? lrs : nnz0;
                const u32 e = complete ? row_begin(lr + 1) : nnz1;
                const double part = sum_strided<T>(st, s - base + lane, e - base, 32);
                const double sum = warp_sum_down<double>(part);
                if (lane == 0) {
                    if (!complete) p.carry[chunk] = sum;
                    else if (lrs < nnz0) p.head_partial[chunk] = sum;
                    else if (lr == p.seed) { p.ctl->seed_sum = sum; p.ctl->seed_flag = 1; }
                    else finalize_row<T, WRITE_Y, RESID>(p, lr, (T)sum, row_inv(lr), uni_add, accS, accR);
                }
            }
            CLK_ADD(11);                                      // consumer: long rows
            // ---- a row of >= 1024 nnz (at most one per tile): the whole group
            const int hr = *huge;
            if (hr >= 0) {
                const u32 hrs = row_begin(hr);
                const bool complete = hr < row1;
                const u32 s = hrs > nnz0 ? hrs : nnz0;
                const u32 e = complete ? row_begin(hr + 1) : nnz1;
                double part = sum_strided<T>(st, s - base + gtid, e - base, GROUP_THREADS);
                part = warp_sum_down<double>(part);
                if (lane == 0) hpart[gwarp] = part;
                group_sync(group);
                if (gtid == 0) {
                    double sum = 0.0;
                    for (int w = 0; w < GROUP_THREADS / 32; w++) sum = __dadd_rn(sum, hpart[w]);
                    if (!complete) p.carry[chunk] = sum;
                    else if (hrs < nnz0) p.head_partial[chunk] = sum;
                    else if (hr == p.seed) { p.ctl->seed_sum = sum; p.ctl->seed_flag = 1; }
                    else finalize_row<T, WRITE_Y, RESID>(p, hr, (T)sum, row_inv(hr), uni_add, accS, accR);
                }
            }
            CLK_ADD(12);                                      // consumer: huge row
            group_sync(group);                         // every read of the stage and of the lists is done
            mbar_arrive_a(empty_b);
            CLK_ADD(13);                                      // consumer: final barrier
        }
    }

    __syncthreads();
    block_sum2<CTA_THREADS>(accS, accR, scratch);
    if (threadIdx.x == 0) {
        p.slot_S[blockIdx.x] = accS;
        p.slot_R[blockIdx.x] = accR;
    }
}

// ------------------------------------------------------------------------------------------------ K7: phased SpMV (default)
// 8 symmetric groups of 128 threads per CTA, each walking its own tiles: prefetch (next tile's coordinates, indices,
// row metadata) -> gathers -> products to shared memory -> group barrier -> row sums + epilogue.  Eight independent
// tiles per SM keep the LSU queue fed while any one group sits in a dependent step.  (rwr_opts.kernel = 0)
template <typename T>
struct PhasedLoad {
    int2 c0, c1;
    int4 iv[CHUNK_ROUNDS];
    u32 rs, re;          // in_ptr[row0 + gtid], in_ptr[row0 + gtid + 1]
    T inv;
};
template <typename T>
__device__ __forceinline__ void phased_load(const IterParams<T>& p, int2 c0, int2 c1, int gtid, u64 pol, PhasedLoad<T>& t) {
    t.c0 = c0; t.c1 = c1;
    const u32 nnz1 = (u32)c1.y, base = (u32)c0.y & ~3u;
#pragma unroll
    for (int j = 0; j < CHUNK_ROUNDS; j++) {
        const u32 pos = base + (u32)(j * GROUP_THREADS + gtid) * 4u;
        t.iv[j] = make_int4(0, 0, 0, 0);
        if (pos < nnz1) t.iv[j] = ld_stream_int4(reinterpret_cast<const int4*>(p.in_src + pos), pol);
    }
    const int r = c0.x + gtid;
    t.rs = 0; t.re = 0; t.inv = (T)0;
    if (r < p.n) { t.rs = ld_stream_u32(p.in_ptr + r, pol); t.re = ld_stream_u32(p.in_ptr + r + 1, pol); t.inv = ld_stream(p.inv + r, pol); }
}

template <typename T, bool VALUED, bool WRITE_Y, bool RESID, bool DEBUG>
__global__ void __launch_bounds__(CTA_THREADS, 1) k_spmv_phased(const IterParams<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (p.ctl->done) return;
    double* scratch = reinterpret_cast<double*>(smem_raw + HDR_BARS);
    u32 smem0;
    asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(smem0) : "l"(smem_raw));
    const u32 hub_addr = smem0 + PIPE_HDR + PAIRS * STAGES * StageOff<T>::size;   // same place as in the pipelined kernel

    const int group = threadIdx.x / GROUP_THREADS, gtid = threadIdx.x % GROUP_THREADS;
    const int lane = threadIdx.x & 31, gwarp = gtid >> 5;
    const u32 prod = smem0 + PIPE_HDR + (u32)group * CHUNK_SPAN * (u32)sizeof(T);
    // per-group lists: r[16] s[16] e[16] rs[16] | count, huge r, s, e, rs | warp partials
    int* gl = reinterpret_cast<int*>(smem_raw + HDR_BARS + HDR_SCRATCH + group * 384);
    int* l_r = gl; int* l_s = gl + 16; int* l_e = gl + 32; int* l_rs = gl + 48;
    int* l_cnt = gl + 64; int* h_r = gl + 65; int* h_s = gl + 66; int* h_e = gl + 67; int* h_rs = gl + 68;
    double* hpart = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(gl) + 320);   // 4 warp partials

    if (threadIdx.x == 0) mbar_init(reinterpret_cast<u64*>(smem_raw), 1);
    if (gtid == 0) { *l_cnt = 0; *h_r = -1; }
    __syncthreads();
    if (p.hub > 0 && threadIdx.x == 0) {
        const u32 bytes = (u32)p.hub * (u32)sizeof(T);
        mbar_expect_tx(reinterpret_cast<u64*>(smem_raw), bytes);
        for (u32 off = 0; off < bytes; off += 32768) {
            u32 len = bytes - off < 32768 ? bytes - off : 32768;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(hub_addr + off),
                         "l"(reinterpret_cast<const unsigned char*>(p.x) + off), "r"(len), "r"(smem0)
                         : "memory");
        }
    }
    const double uni_add = (p.seed < 0) ? p.ctl->S * p.inv_n : 0.0;
    const u64 pol_stream = policy_evict_first(), pol_keep = policy_evict_last();
    double accS = 0.0, accR = 0.0;
    // Tiles are walked in runs of RUN consecutive chunks: inside a run the coordinates of the next two tiles are
    // already in registers (c1, c2), so the index and row-metadata loads of tile t+1 never wait on another load.
    constexpr int RUN = 8;
    const int n_groups = gridDim.x * GROUPS;
    PhasedLoad<T> cur;
    if (p.hub > 0) mbar_wait_a(smem0, 0);

    for (int run0 = (blockIdx.x * GROUPS + group) * RUN; run0 < p.n_chunks; run0 += n_groups * RUN) {
    const int run1 = run0 + RUN < p.n_chunks ? run0 + RUN : p.n_chunks;
    int2 c2 = p.part[run0 + 2 <= p.n_chunks ? run0 + 2 : p.n_chunks];
    phased_load<T>(p, p.part[run0], p.part[run0 + 1], gtid, pol_stream, cur);
    CLK_DECL;
    for (int chunk = run0; chunk < run1; chunk++) {
        const u32 nnz0 = (u32)cur.c0.y, nnz1 = (u32)cur.c1.y, base = nnz0 & ~3u;
        const int row0 = cur.c0.x, row1 = cur.c1.x;
        // ---- coordinates two tiles ahead: one load, first needed a whole tile from now
        const int2 c1 = cur.c1;
        const int2 c3 = p.part[chunk + 3 <= p.n_chunks ? chunk + 3 : p.n_chunks];
        if (cur.iv[0].x == -7 && cur.iv[CHUNK_ROUNDS - 1].w == -7 && cur.rs == 77u) accS += 1.0;   // (profiling) lands the prefetch
        CLK_ADD(0);                                           // waiting for the prefetched tile
        // ---- gathers (branch-free, all 8 in flight)
        T v[CHUNK_ROUNDS][4];
        int take_hub[CHUNK_ROUNDS][4], take_glob[CHUNK_ROUNDS][4];
#pragma unroll
        for (int j = 0; j < CHUNK_ROUNDS; j++) {
            const u32 pos = base + (u32)(j * GROUP_THREADS + gtid) * 4u;
            const int s4[4] = {cur.iv[j].x, cur.iv[j].y, cur.iv[j].z, cur.iv[j].w};
            const bool inner = pos >= nnz0 && pos + 3 < nnz1;      // the whole int4 lies inside the tile (the usual case)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const u32 e = pos + k;
                const bool valid = inner || (e >= nnz0 && e < nnz1);
                take_hub[j][k] = valid && s4[k] < p.hub;
                take_glob[j][k] = valid && s4[k] >= p.hub;
            }
        }
#pragma unroll
        for (int j = 0; j < CHUNK_ROUNDS; j++) {
            int s4[4] = {cur.iv[j].x, cur.iv[j].y, cur.iv[j].z, cur.iv[j].w};
            if (DEBUG && p.debug) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (p.debug == 1) { take_hub[j][k] = 0; take_glob[j][k] = 0; }                       // no gathers
                    else if (p.debug == 2) { s4[k] = p.hub ? s4[k] % p.hub : 0; take_hub[j][k] |= take_glob[j][k]; take_glob[j][k] = 0; }   // all shared
                    else if (p.debug == 3) { s4[k] = p.hub + (s4[k] & 4095); take_glob[j][k] |= take_hub[j][k]; take_hub[j][k] = 0; }      // all global, 32 KB
                    else if (p.debug == 5) { s4[k] = p.hub + (s4[k] & 0x3fffff); take_glob[j][k] |= take_hub[j][k]; take_hub[j][k] = 0; }  // all global, 32 MB
                }
            }
#pragma unroll
            for (int k = 0; k < 4; k++)
                v[j][k] = gather_sel(take_hub[j][k], take_glob[j][k], hub_addr + (u32)s4[k] * (u32)sizeof(T), p.x + s4[k],
                                         s4[k] < p.n_hot ? pol_keep : pol_stream);
        }
        if (VALUED) {
#pragma unroll
            for (int j = 0; j < CHUNK_ROUNDS; j++) {
                const u32 pos = base + (u32)(j * GROUP_THREADS + gtid) * 4u;
                if (pos < nnz1) {
                    T w[4];
                    load4_stream(p.in_val + pos, pol_stream, w);
#pragma unroll
                    for (int k = 0; k < 4; k++) v[j][k] = mul_rn(v[j][k], w[k]);
                }
            }
        }
        CLK_ADD(1);                                           // issuing gathers
        const u32 my_rs = cur.rs, my_re = cur.re;
        const T my_inv = cur.inv;
        // ---- prefetch the next tile's indices and row metadata (arrive while this tile is being summed)
        if (chunk + 1 < run1) phased_load<T>(p, c1, c2, gtid, pol_stream, cur);
        c2 = c3;
        CLK_ADD(2);                                           // issuing the prefetch
        if (row0 == row1) {
            // ---- the tile lies inside one long row: nothing completes here, its sum is this chunk's carry.
            //      Reduced in registers (fixed order), no trip through the product buffer.
            double part = 0.0;
#pragma unroll
            for (int j = 0; j < CHUNK_ROUNDS; j++)
#pragma unroll
                for (int k = 0; k < 4; k++) part = __dadd_rn(part, (double)v[j][k]);     // out-of-range slots hold +0
            part = warp_sum_down<double>(part);
            if (lane == 0) hpart[gwarp] = part;
            group_sync(group);
            if (gtid == 0) {
                double sum = 0.0;
                for (int w = 0; w < GROUP_THREADS / 32; w++) sum = __dadd_rn(sum, hpart[w]);
                p.carry[chunk] = sum;
            }
            group_sync(group);
            CLK_ADD(3);                                       // single-row tile: gather wait + reduce
            continue;
        }
        // ---- products to shared memory
#pragma unroll
        for (int j = 0; j < CHUNK_ROUNDS; j++) {
            const u32 pos = base + (u32)(j * GROUP_THREADS + gtid) * 4u;
            if (pos < nnz1) sts4(prod + (pos - base) * (u32)sizeof(T), v[j]);
        }
        CLK_ADD(4);                                           // gather wait + STS
        group_sync(group);
        CLK_ADD(5);                                           // barrier 1

        // ---- rows: one thread per row, products summed in storage order
        bool first = true;
        for (int r = row0 + gtid; r <= row1 && r < p.n && !(DEBUG && p.debug == 4); r += GROUP_THREADS) {
            const bool complete = r < row1;
            u32 rs, re;
            T invr;
            if (first) { rs = my_rs; re = my_re; invr = my_inv; first = false; }
            else { rs = ld_stream_u32(p.in_ptr + r, pol_stream); re = ld_stream_u32(p.in_ptr + r + 1, pol_stream); invr = ld_stream(p.inv + r, pol_stream); }
            const u32 s = rs > nnz0 ? rs : nnz0;
            u32 e = complete ? re : nnz1;
            if (e < s) e = s;
            const u32 len = e - s;
            if (len >= (u32)HUGE_ROW) {
                *h_r = r; *h_s = (int)(s - base); *h_e = (int)(e - base); *h_rs = (int)rs;
            } else if (len >= (u32)LONG_ROW) {
                const int slot = atomicAdd(l_cnt, 1);
                l_r[slot] = r; l_s[slot] = (int)(s - base); l_e[slot] = (int)(e - base); l_rs[slot] = (int)rs;
            } else {
                const double sum = sum_run<T>(prod, s - base, e - base);
                if (!complete) p.carry[chunk] = sum;
                else if (rs < nnz0) p.head_partial[chunk] = sum;
                else if (r == p.seed) { p.ctl->seed_sum = sum; p.ctl->seed_flag = 1; }
                else finalize_row<T, WRITE_Y, RESID>(p, r, (T)sum, invr, uni_add, accS, accR);
            }
        }
        CLK_ADD(6);                                           // short rows
        group_sync(group);
        CLK_ADD(7);                                           // barrier 2
        // ---- rows of 64..511 nnz: one warp each
        const int n_long = *l_cnt;
        for (int li = gwarp; li < n_long; li += GROUP_THREADS / 32) {
            const int lr = l_r[li];
            const u32 s = (u32)l_s[li], e = (u32)l_e[li], lrs = (u32)l_rs[li];
            const bool complete = lr < row1;
            const double part = sum_strided<T>(prod, s + lane, e, 32);
            const double sum = warp_sum_down<double>(part);
            if (lane == 0) {
                if (!complete) p.carry[chunk] = sum;
                else if (lrs < nnz0) p.head_partial[chunk] = sum;
                else if (lr == p.seed) { p.ctl->seed_sum = sum; p.ctl->seed_flag = 1; }
                else finalize_row<T, WRITE_Y, RESID>(p, lr, (T)sum, p.inv[lr], uni_add, accS, accR);
            }
        }
        // ---- a row of >= 512 nnz (at most one per tile): the whole group
        const int hr = *h_r;
        if (hr >= 0) {
            const u32 s = (u32)*h_s, e = (u32)*h_e, hrs = (u32)*h_rs;
            const bool complete = hr < row1;
            double part = sum_strided<T>(prod, s + gtid, e, GROUP_THREADS);
            part = warp_sum_down<double>(part);
            if (lane == 0) hpart[gwarp] = part;
            group_sync(group);
            if (gtid == 0) {
                double sum = 0.0;
                for (int w = 0; w < GROUP_THREADS / 32; w++) sum = __dadd_rn(sum, hpart[w]);
                if (!complete) p.carry[chunk] = sum;
                else if (hrs < nnz0) p.head_partial[chunk] = sum;
                else if (hr == p.seed) { p.ctl->seed_sum = sum; p.ctl->seed_flag = 1; }
                else finalize_row<T, WRITE_Y, RESID>(p, hr, (T)sum, p.inv[hr], uni_add, accS, accR);
            }
        }
        CLK_ADD(8);                                           // long + huge rows
        group_sync(group);                              // products and lists are free again
        CLK_ADD(9);                                           // barrier 3
        if (gtid == 0) { *l_cnt = 0; *h_r = -1; }
    }
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
            finalize_row<T, WRITE_Y, RESID>(p, row, (T)total, p.inv[row], uni_add, accS, accR);
        }
    }
    if (k == 0 && ctl->seed_flag) {                            // seed row finished inside one chunk
        const T y = (T)__dadd_rn(ctl->seed_sum, S);
        finalize_row<T, WRITE_Y, RESID>(p, p.seed, y, p.inv[p.seed], uni_add, accS, accR);
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
                       T* __restrict__ x1, IterCtl* ctl, double S_uniform) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < 8) { x0[n + j] = (T)0; x1[n + j] = (T)0; }     // x[n] is the always-zero entry the edge stream pads with
    if (j == 0) {
        ctl->resid = 0.0; ctl->seed_sum = 0.0; ctl->seed_flag = 0; ctl->done = 0; ctl->iters = 0; ctl->ticket = 0;
        ctl->tile_ctr = 0;
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
static size_t smem_fixed_bytes(size_t elt) {     // the v1 kernel needs a little less; one size keeps the hub identical
    return PIPE_HDR + (size_t)PAIRS * STAGES * (elt == 4 ? sizeof(Stage<float>) : sizeof(Stage<double>));
}

int hub_entries_for(const rwr_graph* g, int precision) {
    if (g->opts.kernel == 0) return ws_hub_entries(g, precision);
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
        g->in_val32.alloc((size_t)g->nnz + IDX_PAD, &g->pool);
        if (g->nnz) k_f64_to_f32<<<div_up((size_t)g->nnz, 256), 256, 0, st>>>(g->in_val64.p, g->in_val32.p, (size_t)g->nnz);
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

// rwr_opts.kernel: 0 = the warp-streamed kernel of stream.cu (default); 1 = pipelined producer/consumer; 2 = phased
template <typename T, bool VALUED, bool WRITE_Y, bool RESID>
static void launch_spmv(const IterParams<T>& p, int variant, int grid, size_t smem, cudaStream_t st) {
    auto kern = variant != 1 ? (p.debug ? k_spmv_phased<T, VALUED, WRITE_Y, RESID, true> : k_spmv_phased<T, VALUED, WRITE_Y, RESID, false>)
                             : k_spmv<T, VALUED, WRITE_Y, RESID>;
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, CTA_THREADS, smem, st>>>(p);
    KERNEL_CHECK();
}

template <typename T>
static void launch_iteration(rwr_graph* g, const IterParams<T>& p, bool write_y, bool resid, int main_grid, int fix_grid,
                             size_t smem, double thr, int use_thr) {
    cudaStream_t st = g->stream;
    if (g->opts.kernel == 0) {
        const bool parted = dist_n_ranks(g->comm) > 1;
        // on a row slice the convergence test needs the residual of all slices: it runs after the exchange
        ws_launch_iteration<T>(g, p, resid, thr, parted ? 0 : use_thr);
        static const bool skip_exchange = getenv("RWR_DIST_SKIP") != nullptr;      // timing probe only: wrong results
        if (parted && !skip_exchange) {
            // x_next: already in every peer's copy when the epilogue stored it there (p.n_peers > 0), else NCCL
            dist_exchange(g, p.n_peers ? nullptr : p.x_next, sizeof(T), p.ctl->red);
            k_after_reduce<<<1, 1, 0, st>>>(p.ctl, thr, use_thr);
            KERNEL_CHECK();
        }
        return;
    }
    const bool valued = g->layout == RWR_LAYOUT_VALUED;
#define LAUNCH(V, W, R)                                                                  \
    do {                                                                                 \
        launch_spmv<T, V, W, R>(p, g->opts.kernel, main_grid, smem, st);                                 \
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
    Scratch<unsigned char> xa, xb, ya;
    Scratch<double> carry, head, slot_S, slot_R;
    Scratch<IterCtl> ctl;
    void* x[2] = {nullptr, nullptr};      // the two gather vectors: scratch, or the peer-mapped buffers of a partitioned graph
    void alloc_x(rwr_graph* g, size_t vec_bytes) {
        if (g->p2p) { x[0] = g->px[0]; x[1] = g->px[1]; return; }
        xa.alloc(&g->scratch, vec_bytes); xb.alloc(&g->scratch, vec_bytes);
        x[0] = xa.p; x[1] = xb.p;
    }
};

// peers' copies of the buffer `x_next` is (row-partitioned graphs with peer-mapped gather vectors)
template <typename T>
static void set_peers(rwr_graph* g, IterParams<T>& p, const void* x_next) {
    p.parted = dist_n_ranks(g->comm) > 1;
    p.n_peers = 0;
    if (!g->p2p) return;
    const int b = (x_next == g->px[0]) ? 0 : 1, me = dist_rank(g->comm);
    for (int r = 0; r < (int)g->peer_px[b].size(); r++)
        if (r != me) p.peer_next[p.n_peers++] = g->peer_px[b][r];
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
    const int per_cta = g->opts.kernel == 1 ? PAIRS : GROUPS;
    const int main_grid = std::max(1, std::min(g->sm_count, (g->n_chunks + per_cta - 1) / per_cta));
    const int fix_grid = div_up((size_t)g->n_chunks, FIX_THREADS);
    const int hub = hub_entries_for(g, Prec<T>::id);
    const size_t smem = smem_fixed_bytes(sizeof(T)) + (size_t)hub * sizeof(T);

    T* xa = reinterpret_cast<T*>(ws.x[0]);
    T* xb = reinterpret_cast<T*>(ws.x[1]);
    T* ya = reinterpret_cast<T*>(ws.ya.p);
    IterParams<T> p;
    p.in_ptr = g->in_ptr.p; p.in_src = g->in_src.p; p.in_val = Prec<T>::val(g); p.part = g->part.p;
    p.n_chunks = g->n_chunks; p.n = n; p.inv = Prec<T>::inv(g);
    p.ws_src = g->ws_src.p; p.ws_val = Prec<T>::wsval(g); p.ws_tile = g->ws_tile.p; p.ws_tiles = g->ws_tiles;
    p.row_begin = g->row_begin; p.row_end = g->row_end;
    p.omc = (T)(1.0 - c);                                      // Model.cs:84 `(1 - dampingFactor)`
    p.seed = seed_int; p.inv_n = n ? 1.0 / (double)n : 0.0; p.hub = hub; p.n_hot = g->n_hot; p.debug = 0;
    p.head_partial = ws.head.p; p.carry = ws.carry.p;
    p.slot_S = ws.slot_S.p; p.slot_R = ws.slot_R.p; p.ctl = ws.ctl.p;

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
        for (int it = 0; it < n_iter; it++) {
            p.x = x_cur; p.x_next = x_nxt; p.r_prev = nullptr; p.y = y_out;
            set_peers<T>(g, p, x_nxt);
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
                set_peers<T>(g, p, x_nxt);
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
    if (launched > 0) dist_allgather_rows(g, y_out, sizeof(T));       // row-partitioned: every rank gets the whole rank vector
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
    if (Prec<T>::id == RWR_FP32) ensure_fp32_arrays(g);
    const size_t ld = (n + 3) & ~(size_t)3;
    res->ld = ld;
    if (Prec<T>::ybuf(res).n != std::max<size_t>(1, ld * (size_t)n_seeds))      // a re-run keeps its rank buffers
        Prec<T>::ybuf(res).alloc(std::max<size_t>(1, ld * (size_t)n_seeds), nullptr);
    RunWorkspace ws;
    const size_t vec_bytes = (n + 8) * sizeof(T);
    ws.alloc_x(g, vec_bytes); ws.ya.alloc(&g->scratch, vec_bytes);
    ws.carry.alloc(&g->scratch, (size_t)g->n_chunks); ws.head.alloc(&g->scratch, (size_t)g->n_chunks);
    CUDA_CHECK(cudaMemsetAsync(ws.carry.p, 0, (size_t)g->n_chunks * sizeof(double), st));
    CUDA_CHECK(cudaMemsetAsync(ws.head.p, 0, (size_t)g->n_chunks * sizeof(double), st));
    const size_t slots = (size_t)g->sm_count * 8 + div_up((size_t)g->n_chunks, FIX_THREADS) + 8;
    ws.slot_S.alloc(&g->scratch, slots); ws.slot_R.alloc(&g->scratch, slots);
    CUDA_CHECK(cudaMemsetAsync(ws.slot_S.p, 0, slots * sizeof(double), st));
    CUDA_CHECK(cudaMemsetAsync(ws.slot_R.p, 0, slots * sizeof(double), st));
    ws.ctl.alloc(&g->scratch, 1);
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
    if (Prec<T>::id == RWR_FP32) ensure_fp32_arrays(g);
    RunWorkspace ws;
    const size_t vec_bytes = (n + 8) * sizeof(T);
    ws.alloc_x(g, vec_bytes); ws.ya.alloc(&g->scratch, vec_bytes);
    ws.carry.alloc(&g->scratch, (size_t)g->n_chunks); ws.head.alloc(&g->scratch, (size_t)g->n_chunks);
    CUDA_CHECK(cudaMemsetAsync(ws.carry.p, 0, (size_t)g->n_chunks * sizeof(double), st));
    CUDA_CHECK(cudaMemsetAsync(ws.head.p, 0, (size_t)g->n_chunks * sizeof(double), st));
    const size_t slots = (size_t)g->sm_count * 8 + div_up((size_t)g->n_chunks, FIX_THREADS) + 8;
    ws.slot_S.alloc(&g->scratch, slots); ws.slot_R.alloc(&g->scratch, slots);
    CUDA_CHECK(cudaMemsetAsync(ws.slot_S.p, 0, slots * sizeof(double), st));
    CUDA_CHECK(cudaMemsetAsync(ws.slot_R.p, 0, slots * sizeof(double), st));
    ws.ctl.alloc(&g->scratch, 1);
    int seed_int = 0;
    CUDA_CHECK(cudaMemcpyAsync(&seed_int, g->new_of_old.p + seed_orig, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    const int per_cta = g->opts.kernel == 1 ? PAIRS : GROUPS;
    const int main_grid = std::max(1, std::min(g->sm_count, (g->n_chunks + per_cta - 1) / per_cta));
    const int fix_grid = div_up((size_t)g->n_chunks, FIX_THREADS);
    const int hub = hub_entries_for(g, Prec<T>::id);
    const size_t smem = smem_fixed_bytes(sizeof(T)) + (size_t)hub * sizeof(T);
    T* xa = reinterpret_cast<T*>(ws.x[0]);
    T* xb = reinterpret_cast<T*>(ws.x[1]);
    T* ya = reinterpret_cast<T*>(ws.ya.p);
    IterParams<T> p;
    p.in_ptr = g->in_ptr.p; p.in_src = g->in_src.p; p.in_val = Prec<T>::val(g); p.part = g->part.p;
    p.n_chunks = g->n_chunks; p.n = g->n; p.inv = Prec<T>::inv(g);
    p.ws_src = g->ws_src.p; p.ws_val = Prec<T>::wsval(g); p.ws_tile = g->ws_tile.p; p.ws_tiles = g->ws_tiles;
    p.row_begin = g->row_begin; p.row_end = g->row_end;
    p.omc = (T)(1.0 - c); p.seed = seed_int; p.inv_n = n ? 1.0 / (double)n : 0.0; p.hub = hub; p.n_hot = g->n_hot;
    p.head_partial = ws.head.p; p.carry = ws.carry.p; p.slot_S = ws.slot_S.p; p.slot_R = ws.slot_R.p; p.ctl = ws.ctl.p;
    p.r_prev = nullptr; p.y = ya;
    { const char* dm = getenv("RWR_DEBUG_MODE"); p.debug = dm ? atoi(dm) : 0; }
    k_init<T><<<div_up(std::max((int)n, 1), 256), 256, 0, st>>>((int)n, seed_int, p.omc, p.inv, ya, xa, xb, ws.ctl.p, 0.0);
    KERNEL_CHECK();
    const bool valued = g->layout == RWR_LAYOUT_VALUED;
    std::vector<cudaEvent_t> ev(3 * (size_t)reps);
    for (auto& e : ev) CUDA_CHECK(cudaEventCreate(&e));
    T* x_cur = xa;
    T* x_nxt = xb;
    for (int it = 0; it < 3 + reps; it++) {
        p.x = x_cur; p.x_next = x_nxt;
        set_peers<T>(g, p, x_nxt);
        const int r = it - 3;
        if (r >= 0) CUDA_CHECK(cudaEventRecord(ev[3 * r], st));
        if (g->opts.kernel == 0) ws_launch_spmv_only<T>(g, p);
        else if (valued) launch_spmv<T, true, false, false>(p, g->opts.kernel, main_grid, smem, st);
        else launch_spmv<T, false, false, false>(p, g->opts.kernel, main_grid, smem, st);
        if (r >= 0) CUDA_CHECK(cudaEventRecord(ev[3 * r + 1], st));
        if (g->opts.kernel == 0) ws_launch_finish_only<T>(g, p, false, 0.0, 0);
        else k_fixup<T, false, false><<<fix_grid, FIX_THREADS, 0, st>>>(p, main_grid, 0.0, 0);
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

// One fixed-iteration run of one seed into a caller-provided rank vector (internal labels); used by the fused
// request path (rwr_recommend) so that no result object and no cudaMalloc / cudaFree sit on that path.
template <typename T>
void iterate_single_into(rwr_graph* g, int seed_orig, double c, int n_iter, T* y_out, float* iter_ms, int64_t* launches) {
    cudaStream_t st = g->stream;
    const size_t n = (size_t)g->n;
    if (Prec<T>::id == RWR_FP32) ensure_fp32_arrays(g);
    RunWorkspace ws;
    const size_t vec_bytes = (n + 8) * sizeof(T);
    ws.alloc_x(g, vec_bytes); ws.ya.alloc(&g->scratch, 16);
    ws.carry.alloc(&g->scratch, (size_t)g->n_chunks); ws.head.alloc(&g->scratch, (size_t)g->n_chunks);
    const size_t slots = (size_t)g->sm_count * 8 + div_up((size_t)g->n_chunks, FIX_THREADS) + 8;
    ws.slot_S.alloc(&g->scratch, slots); ws.slot_R.alloc(&g->scratch, slots);
    ws.ctl.alloc(&g->scratch, 1);
    cudaEvent_t ev0, ev1;
    CUDA_CHECK(cudaEventCreateWithFlags(&ev0, cudaEventDefault));
    CUDA_CHECK(cudaEventCreateWithFlags(&ev1, cudaEventDefault));
    const int64_t l0 = g->pool.launches;
    int it = 0;
    double rs = 0;
    run_one<T>(g, ws, seed_orig, c, 0, n_iter, 0.0, 0, y_out, &it, &rs, iter_ms, ev0, ev1);
    *launches += g->pool.launches - l0;
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
}
template void iterate_single_into<double>(rwr_graph*, int, double, int, double*, float*, int64_t*);
template void iterate_single_into<float>(rwr_graph*, int, double, int, float*, float*, int64_t*);

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
        res = reuse ? reuse : new rwr_result();
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
