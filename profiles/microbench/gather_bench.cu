// gather_bench.cu -- microbenchmark behind the SpMV design: how many random 8-byte (and 4-byte) reads per cycle
// per SM does a B200 sustain from (a) local shared memory, (b) distributed shared memory of a cluster, (c) L2?
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_bench gather_bench.cu ; run: ./gather_bench
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned hash32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

__device__ __forceinline__ double ld_hint(const double* ptr, unsigned long long pol) {
    double r;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(ptr), "l"(pol));
    return r;
}
__device__ __forceinline__ float ld_hint(const float* ptr, unsigned long long pol) {
    float r;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(ptr), "l"(pol));
    return r;
}
// mode 0: local smem table; mode 1: DSMEM (table striped over the cluster); mode 2: global (L2-resident) table
// mode 3: global with the evict_last cache-hint asm; mode 4: 40 % smem / 60 % global, chosen per lane (the SpMV mix)
// mode 5: like 4 but only threads < active gather (the rest idle)
template <typename T, int MODE, int UNROLL>
__global__ void k_gather(const T* __restrict__ gtable, unsigned gmask, int table_entries, int iters, T* out) {
    extern __shared__ __align__(16) unsigned char smem[];
    T* table = reinterpret_cast<T*>(smem);
    for (int i = threadIdx.x; i < table_entries; i += blockDim.x) table[i] = (T)(i & 1023);
    unsigned csize = 1, crank = 0;
    cg::cluster_group cluster = cg::this_cluster();
    if (MODE == 1) { csize = cluster.num_blocks(); crank = cluster.block_rank(); cluster.sync(); } else __syncthreads();
    const T* peer[16];
    if (MODE == 1) for (unsigned r = 0; r < 16; r++) peer[r] = cluster.map_shared_rank(table, r % csize);
    unsigned h = hash32(blockIdx.x * 1315423911u + threadIdx.x * 2654435761u + 17u);
    T acc = 0;
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    if (MODE == 5 && threadIdx.x >= 512) return;
    const unsigned tmask = (unsigned)table_entries - 1;      // table_entries is a power of two
    for (int it = 0; it < iters; it++) {
        T v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            h = hash32(h + u);
            if (MODE == 0) v[u] = table[h & tmask];
            else if (MODE == 1) v[u] = peer[(h >> 24) % csize][h & tmask];
            else if (MODE == 2) v[u] = gtable[h & gmask];
            else if (MODE == 3) v[u] = ld_hint(gtable + (h & gmask), pol);
            else v[u] = ((h >> 27) < 13) ? table[h & tmask] : ld_hint(gtable + ((h >> 3) & gmask), pol);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) acc += v[u];
    }
    if (MODE == 1) cluster.sync();
    if (acc == (T)-12345) out[0] = acc;
    (void)crank;
}

template <typename T, int MODE>
static void run(const char* name, int cluster, int table_bytes, const T* gtable, unsigned gmask, T* out) {
    int entries = table_bytes / sizeof(T);
    auto kern = k_gather<T, MODE, 8>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, table_bytes));
    if (cluster > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    int sms = 148, threads = 1024, iters = 2000;
    int grid = (sms / cluster) * cluster;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = table_bytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; rep++) {
        CK(cudaEventRecord(e0));
        CK(cudaLaunchKernelEx(&cfg, kern, gtable, gmask, entries, iters, out));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
    }
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double reads = (double)grid * threads * iters * 8.0;
    printf("%-28s cluster %2d  elt %zu B  table %3d KB/CTA  %7.3f ms  %7.1f Greads/s  %5.2f reads/ns/SM\n", name, cluster,
           sizeof(T), table_bytes / 1024, ms, reads / ms / 1e6, reads / ms / 1e6 / grid);
}

int main() {
    size_t gbytes = 64u << 20;        // 64 MB global table: L2-resident on B200
    {   // tables larger than L2: random sector reads from HBM
        size_t big = (size_t)2 << 30;
        double* gb; CK(cudaMalloc(&gb, big)); CK(cudaMemset(gb, 0, big));
        double* o2; CK(cudaMalloc(&o2, 64));
        run<double, 2>("global f64, 2 GB table (HBM)", 1, 1024, gb, (unsigned)(big / 8 - 1), o2);
        run<double, 2>("global f64, 256 MB table", 1, 1024, gb, (unsigned)((256u << 20) / 8 - 1), o2);
        run<double, 2>("global f64, 128 MB table", 1, 1024, gb, (unsigned)((128u << 20) / 8 - 1), o2);
        run<double, 2>("global f64, 96 MB table", 1, 1024, gb, (unsigned)((96u << 20) / 8 - 1) , o2);
        CK(cudaFree(gb));
    }
    double* g64; CK(cudaMalloc(&g64, gbytes)); CK(cudaMemset(g64, 0, gbytes));
    double* out; CK(cudaMalloc(&out, 64));
    const int TB = 128 * 1024;
    run<double, 0>("local smem f64", 1, TB, g64, 0, out);
    run<float, 0>("local smem f32", 1, TB, (float*)g64, 0, (float*)out);
    for (int c : {2, 4, 8, 16}) run<double, 1>("dsmem f64", c, TB, g64, 0, out);
    for (int c : {8, 16}) run<float, 1>("dsmem f32", c, TB, (float*)g64, 0, (float*)out);
    run<double, 2>("global f64, 64 MB table", 1, 1024, g64, (unsigned)(gbytes / 8 - 1), out);
    run<double, 2>("global f64, 8 MB table", 1, 1024, g64, (unsigned)((8u << 20) / 8 - 1), out);
    run<double, 2>("global f64, 1 MB table", 1, 1024, g64, (unsigned)((1u << 20) / 8 - 1), out);
    run<float, 2>("global f32, 64 MB table", 1, 1024, (float*)g64, (unsigned)(gbytes / 4 - 1), (float*)out);
    run<double, 3>("global f64 hint, 64 MB", 1, 1024, g64, (unsigned)(gbytes / 8 - 1), out);
    run<double, 4>("mix 40% smem/60% glob f64", 1, TB, g64, (unsigned)(gbytes / 8 - 1), out);
    run<double, 5>("mix, 512 of 1024 threads", 1, TB, g64, (unsigned)(gbytes / 8 - 1), out);
    run<double, 2>("global f64, 88 MB table", 1, 1024, g64, (unsigned)(gbytes / 8 - 1), out);
    return 0;
}
