// common.cuh -- shared host/device helpers of librwr_b200 (sm_100a only, no fallback paths).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rwr_b200.h"

typedef uint32_t u32;
typedef uint64_t u64;
typedef uint8_t u8;

// ------------------------------------------------------------------ error plumbing
void rwr_set_error(const char* fmt, ...);

struct RwrError {
    int code;
};

#define RWR_FAIL(code_, ...)          \
    do {                              \
        rwr_set_error(__VA_ARGS__);   \
        throw RwrError{(code_)};      \
    } while (0)

#define CUDA_CHECK(expr)                                                                              \
    do {                                                                                              \
        cudaError_t err__ = (expr);                                                                   \
        if (err__ != cudaSuccess) {                                                                   \
            int code__ = (err__ == cudaErrorMemoryAllocation) ? RWR_E_OOM : RWR_E_CUDA;               \
            rwr_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
            throw RwrError{code__};                                                                   \
        }                                                                                             \
    } while (0)

#define KERNEL_CHECK() CUDA_CHECK(cudaGetLastError())

// Every exported function body runs inside this guard: exceptions never cross the C ABI.
#define RWR_API_BEGIN try {
#define RWR_API_END                                     \
    }                                                   \
    catch (const RwrError& e__) { return e__.code; }    \
    catch (const std::bad_alloc&) {                     \
        rwr_set_error("host allocation failed");        \
        return RWR_E_OOM;                               \
    }                                                   \
    catch (...) {                                       \
        rwr_set_error("unexpected exception");          \
        return RWR_E_INVALID;                           \
    }

// ------------------------------------------------------------------ device memory
// Tracks bytes so rwr_graph_info.device_bytes is exact.
struct DevPool {
    int64_t bytes = 0;
    int64_t launches = 0;   // kernels launched through this handle (bench "gpu_launches")
};

// Device allocations made while an AllocStream guard is alive on the calling thread come from the device's memory pool,
// ordered on that stream (cudaMallocAsync / cudaFreeAsync): no device-wide synchronisation on free, and microseconds
// instead of ~0.1 ms per call -- the reference builds one graph per ego network, from up to ten threads at once.
// Without a guard (or for buffers marked `plain`, which may outlive the stream) plain cudaMalloc / cudaFree are used.
extern thread_local cudaStream_t tls_alloc_stream;
struct AllocStream {
    cudaStream_t prev;
    explicit AllocStream(cudaStream_t s) : prev(tls_alloc_stream) { tls_alloc_stream = s; }
    ~AllocStream() { tls_alloc_stream = prev; }
    AllocStream(const AllocStream&) = delete;
    AllocStream& operator=(const AllocStream&) = delete;
};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevPool* pool = nullptr;
    cudaStream_t astream = nullptr;   // stream the block was allocated on (pool allocation), null: cudaMalloc
    bool plain = false;               // always cudaMalloc / cudaFree (the buffer may be freed after its stream is gone)
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept { *this = std::move(o); }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            release();
            p = o.p; n = o.n; pool = o.pool; astream = o.astream; plain = o.plain;
            o.p = nullptr; o.n = 0; o.astream = nullptr;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count, DevPool* pl = nullptr) {
        release();
        pool = pl;
        n = count;
        size_t bytes = (count ? count : 1) * sizeof(T);
        if (!plain && tls_alloc_stream) {
            CUDA_CHECK(cudaMallocAsync((void**)&p, bytes, tls_alloc_stream));
            astream = tls_alloc_stream;
        } else {
            CUDA_CHECK(cudaMalloc((void**)&p, bytes));
            astream = nullptr;
        }
        if (pool) pool->bytes += (int64_t)bytes;
    }
    void release() {
        if (p) {
            if (astream) cudaFreeAsync(p, astream); else cudaFree(p);
            if (pool) pool->bytes -= (int64_t)((n ? n : 1) * sizeof(T));
        }
        p = nullptr;
        n = 0;
        astream = nullptr;
    }
    operator T*() const { return p; }
};

// Per-handle cache of device scratch blocks: request-path workspaces are reused instead of going through
// cudaMalloc / cudaFree (which synchronise the device) on every call.
struct ScratchPool {
    struct Block { void* p; size_t bytes; bool used; cudaStream_t astream; };
    static cudaError_t raw_alloc(void** p, size_t bytes, cudaStream_t* as) {
        *as = tls_alloc_stream;
        return tls_alloc_stream ? cudaMallocAsync(p, bytes, tls_alloc_stream) : cudaMalloc(p, bytes);
    }
    static void raw_free(const Block& b) { if (b.astream) cudaFreeAsync(b.p, b.astream); else cudaFree(b.p); }
    std::vector<Block> blocks;
    DevPool* pool = nullptr;
    void* get(size_t bytes) {
        if (bytes == 0) bytes = 16;
        int best = -1;
        for (int i = 0; i < (int)blocks.size(); i++)
            if (!blocks[i].used && blocks[i].bytes >= bytes && blocks[i].bytes <= 2 * bytes + 4096 &&
                (best < 0 || blocks[i].bytes < blocks[best].bytes)) best = i;
        if (best >= 0) { blocks[best].used = true; return blocks[best].p; }
        void* p = nullptr;
        cudaStream_t as = nullptr;
        cudaError_t err = raw_alloc(&p, bytes, &as);
        if (err != cudaSuccess) {            // drop every cached free block and retry once
            cudaGetLastError();
            trim();
            err = raw_alloc(&p, bytes, &as);
        }
        if (err != cudaSuccess) {
            rwr_set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(err));
            throw RwrError{err == cudaErrorMemoryAllocation ? RWR_E_OOM : RWR_E_CUDA};
        }
        if (pool) pool->bytes += (int64_t)bytes;
        blocks.push_back({p, bytes, true, as});
        return p;
    }
    void put(void* p) {
        for (auto& b : blocks) if (b.p == p) { b.used = false; return; }
    }
    void trim() {
        std::vector<Block> keep;
        for (auto& b : blocks) {
            if (b.used) keep.push_back(b);
            else { raw_free(b); if (pool) pool->bytes -= (int64_t)b.bytes; }
        }
        blocks.swap(keep);
    }
    ~ScratchPool() { for (auto& b : blocks) raw_free(b); }
};

template <typename T>
struct Scratch {
    T* p = nullptr;
    ScratchPool* sp = nullptr;
    Scratch() {}
    Scratch(const Scratch&) = delete;
    Scratch& operator=(const Scratch&) = delete;
    ~Scratch() { release(); }
    void alloc(ScratchPool* pool_, size_t count) { release(); sp = pool_; p = (T*)sp->get((count ? count : 1) * sizeof(T)); }
    void release() { if (p && sp) sp->put(p); p = nullptr; }
    operator T*() const { return p; }
};

// CUDA event destroyed on every exit of its scope (error paths throw through the C-ABI guard)
struct DevEvent {
    cudaEvent_t e = nullptr;
    DevEvent() { CUDA_CHECK(cudaEventCreate(&e)); }
    ~DevEvent() { if (e) cudaEventDestroy(e); }
    DevEvent(const DevEvent&) = delete;
    DevEvent& operator=(const DevEvent&) = delete;
    operator cudaEvent_t() const { return e; }
};

static inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }
static inline int ceil_log2_u64(u64 n) {
    int L = 0;
    while ((1ULL << L) < n) L++;
    return L;
}

// ------------------------------------------------------------------ device helpers
#ifdef __CUDACC__

__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ u64 policy_evict_first() {
    u64 pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ u64 policy_evict_last() {
    u64 pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// streamed, read-once data: bypass L1, first to leave L2
__device__ __forceinline__ int4 ld_stream_int4(const int4* ptr, u64 pol) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(ptr), "l"(pol));
    return r;
}
__device__ __forceinline__ double ld_stream(const double* ptr, u64 pol) {
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(ptr), "l"(pol));
    return r;
}
__device__ __forceinline__ float ld_stream(const float* ptr, u64 pol) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(ptr), "l"(pol));
    return r;
}
// stores with an L2 eviction policy (streamed results must not push the gathered vector out of L2)
__device__ __forceinline__ void st_policy(double* ptr, double v, u64 pol) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(ptr), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_policy(float* ptr, float v, u64 pol) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(ptr), "f"(v), "l"(pol) : "memory");
}

#endif  // __CUDACC__
