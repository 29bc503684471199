// iterate_dev.cuh -- device-side pieces shared by the power-iteration kernels (iterate.cu, stream.cu):
// launch parameters, PTX helpers (mbarrier / TMA bulk copy / cache-hinted loads), the fused per-row epilogue and
// the fixed-order block reduction.
#pragma once

#include "iterate.h"

constexpr int GROUPS = 8;
constexpr int CTA_THREADS = GROUPS * GROUP_THREADS;       // 1024
constexpr int LONG_ROW = 64;                              // rows with >= LONG_ROW nnz inside a chunk: one warp
constexpr int LONG_CAP = CHUNK_ITEMS / LONG_ROW + 1;      // 16
constexpr int FIX_THREADS = 256;

template <typename T>
struct IterParams {
    const u32* in_ptr;
    const int32_t* in_src;
    const T* in_val;        // valued layout only
    const int2* part;
    int n_chunks;
    int n;
    const int32_t* ws_src;  // edge stream (stream.cu): source | end-of-row flag in bit 31
    const T* ws_val;        // valued layout only
    const u32* ws_tile;     // [ws_tiles + 1] first row of a tile | continued-row flag in bit 31
    int ws_tiles;
    int row_begin, row_end; // rows of W^T this rank owns ([0, n) unless the graph is row-partitioned)
    int parted;             // row-partitioned: the epilogue leaves its partial sums in ctl->red for the allReduce
    int n_peers;            // peers whose copy of x_next the epilogue writes directly (NVLink peer stores)
    void* peer_next[7];
    const T* x;             // gather source, internal labels
    const T* inv;
    const T* r_prev;        // previous rank (residual)
    T* y;
    T* x_next;
    T omc;                  // (1 - c)
    int seed;               // internal label, -1: uniform restart
    double inv_n;           // 1/N (uniform restart)
    int hub;                // x entries staged in shared memory
    int n_hot;              // labels below: hot, L2-resident (evict-last); above: clustered cold nodes (streamed)
    int debug;              // measurement-only ablations of the phased kernel (RWR_DEBUG_MODE, profile hook only)
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
__device__ __forceinline__ void finalize_row(const IterParams<T>& p, int row, T y, T invr, double uni_add, double& accS,
                                             double& accR) {
    if (p.seed < 0) y = add_rn(y, (T)uni_add);
    const u64 pol_first = policy_evict_first();
    if (WRITE_Y) st_policy(p.y + row, y, pol_first);
    const T rw = mul_rn(p.omc, y);
    // the next iteration gathers x_next: hot rows should still be in L2 then, cold rows are streamed
    st_policy(p.x_next + row, mul_rn(rw, invr), row < p.n_hot ? policy_evict_last() : pol_first);
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

// ------------------------------------------------------------------------------------------------ shared-space PTX
__device__ __forceinline__ void mbar_arrive_a(u32 bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(u32 bar, u32 parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ double lds_t(u32 a, double) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ float lds_t(u32 a, float) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ u32 lds_u32(u32 a) { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ int4 lds_int4(u32 a) {
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u32(u32 a, u32 v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_int4(u32 a, int4 v) {
    asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts_t(u32 a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_t(u32 a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts4(u32 a, const double (&v)[4]) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v[0]), "d"(v[1]) : "memory");
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a + 16), "d"(v[2]), "d"(v[3]) : "memory");
}
__device__ __forceinline__ void sts4(u32 a, const float (&v)[4]) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}

// One gather, branch-free: hub hit -> ld.shared, miss -> ld.global (evict-last), out of range -> 0.  Divergent
// branches here made ptxas re-wait on the gathers' scoreboard slot before every index compare, which serialised the
// 8 gathers of a thread; predicated straight-line code lets all of them issue back to back.
__device__ __forceinline__ double gather_sel(int take_hub, int take_glob, u32 saddr, const double* gaddr, u64 pol) {
    double v;
    asm volatile(
        "{\n\t.reg .pred ph, pg;\n\t"
        "setp.ne.b32 ph, %1, 0;\n\t"
        "setp.ne.b32 pg, %2, 0;\n\t"
        "mov.f64 %0, 0d0000000000000000;\n\t"
        "@ph ld.shared.f64 %0, [%3];\n\t"
        "@pg ld.global.nc.L2::cache_hint.f64 %0, [%4], %5;\n\t}"
        : "=d"(v)
        : "r"(take_hub), "r"(take_glob), "r"(saddr), "l"(gaddr), "l"(pol));
    return v;
}
__device__ __forceinline__ float gather_sel(int take_hub, int take_glob, u32 saddr, const float* gaddr, u64 pol) {
    float v;
    asm volatile(
        "{\n\t.reg .pred ph, pg;\n\t"
        "setp.ne.b32 ph, %1, 0;\n\t"
        "setp.ne.b32 pg, %2, 0;\n\t"
        "mov.f32 %0, 0f00000000;\n\t"
        "@ph ld.shared.f32 %0, [%3];\n\t"
        "@pg ld.global.nc.L2::cache_hint.f32 %0, [%4], %5;\n\t}"
        : "=f"(v)
        : "r"(take_hub), "r"(take_glob), "r"(saddr), "l"(gaddr), "l"(pol));
    return v;
}

