// Random-sector gather microbenchmark: the measured denominator of the "random-sector roofline".
// Uniformly random, 32-byte-aligned, read-only gathers over a table of T MiB; optionally the
// random addresses are confined to a window of W MiB that slides over the table (what a
// region-partitioned probe order would produce).
//   ./randsector <table_MiB> <gathers_M> [variant|all] [window_MiB]
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t sm64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

enum { V_NC_NA_256 = 0, V_NC_256, V_PLAIN_256, V_NC_2x128, V_NC_NA_64, V_NC_4x64, V_CG_2x128, V_LU_256, V_NC_NA_EF_256, V_NC_NA_L2_64B, V_NC_NA_L2_128B, V_NC_NA_L2_256B, V_COUNT };
static const char* kNames[] = {"nc.L1no_alloc.v4u64", "nc.v4u64", "plain.v4u64", "nc.2x v2u64", "nc.L1no_alloc.u64 (8B)",
                               "nc.4x u64", "cg.2x v2u64", "lu.v4u64", "nc.na.L2evict_first.v4u64", "nc.na.L2::64B.v4u64",
                               "nc.na.L2::128B.v4u64", "nc.na.L2::256B.v4u64"};

template <int V>
__device__ __forceinline__ uint64_t ld(const char* p, uint64_t pol) {
    uint64_t a = 0, b = 0, c = 0, d = 0;
    if (V == V_NC_NA_256) asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (V == V_NC_256) asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (V == V_PLAIN_256) asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (V == V_NC_2x128) {
        asm volatile("ld.global.nc.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
        asm volatile("ld.global.nc.v2.u64 {%0,%1}, [%2];" : "=l"(c), "=l"(d) : "l"(p + 16));
    }
    if (V == V_NC_NA_64) asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(a) : "l"(p));
    if (V == V_NC_4x64) {
        asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(a) : "l"(p));
        asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(b) : "l"(p + 8));
        asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(c) : "l"(p + 16));
        asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(d) : "l"(p + 24));
    }
    if (V == V_CG_2x128) {
        asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
        asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(c), "=l"(d) : "l"(p + 16));
    }
    if (V == V_LU_256) asm volatile("ld.global.lu.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (V == V_NC_NA_EF_256) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p), "l"(pol));
    // prefetch-size qualifiers (SASS: LDG.E.NA.ENL2.LTC64B/LTC128B/LTC256B.256.CONSTANT): do they set the fill size of an L2 miss?
    if (V == V_NC_NA_L2_64B) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (V == V_NC_NA_L2_128B) asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    if (V == V_NC_NA_L2_256B) asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    return a ^ b ^ c ^ d;
}

template <int V, int U>
__global__ void gather(const char* __restrict__ t, uint64_t nslots, uint64_t wslots, uint64_t n, uint64_t seed,
                       unsigned long long* sink) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t pol = 0;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    uint64_t acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += U * stride) {
        uint64_t v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t g = i + u * stride;
            const uint64_t r = sm64(seed + g);
            // window start slides linearly with the gather index; the address is random inside it
            const uint64_t wstart = wslots >= nslots ? 0 : (uint64_t)((double)g / (double)n * (double)(nslots - wslots));
            const uint64_t slot = wstart + __umul64hi(r, wslots >= nslots ? nslots : wslots);
            v[u] = g < n ? ld<V>(t + slot * 32, pol) : 0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u];
    }
    if (acc == 0x1234567887654321ull) atomicAdd(sink, 1ull);
}

template <int V>
void run(const char* t, uint64_t bytes, uint64_t wbytes, uint64_t n, unsigned long long* sink) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int blocks = 148 * 8, threads = 256;
    gather<V, 8><<<blocks, threads>>>(t, bytes / 32, wbytes / 32, n, 1, sink);
    float best = 1e30f;
    for (int it = 0; it < 3; ++it) {
        CK(cudaEventRecord(e0));
        gather<V, 8><<<blocks, threads>>>(t, bytes / 32, wbytes / 32, n, 1000 + it * 7919, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    printf("table %7llu MiB window %7llu MiB  %-28s : %8.2f G sectors/s  %8.1f GB/s\n", (unsigned long long)(bytes >> 20),
           (unsigned long long)(wbytes >> 20), kNames[V], n / (best * 1e6), n * 32.0 / (best * 1e6));
    fflush(stdout);
}

int main(int argc, char** argv) {
    const uint64_t mib = argc > 1 ? strtoull(argv[1], 0, 10) : 1024;
    const uint64_t n = (argc > 2 ? strtoull(argv[2], 0, 10) : 256) * 1000000ull;
    const char* var = argc > 3 ? argv[3] : "0";
    const uint64_t wmib = argc > 4 ? strtoull(argv[4], 0, 10) : mib;
    const uint64_t bytes = mib << 20, wbytes = (wmib > mib ? mib : wmib) << 20;
    char* t;
    if (getenv("RS_VMM")) {
        // virtual-memory-management allocation: one physical handle, address range aligned to RS_VMM MiB —
        // does the driver map it with pages larger than 2 MiB (address-translation reach of a >64 GiB table)?
        CK(cudaFree(0));
        CUmemAllocationProp prop = {};
        prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        prop.location.id = 0;
        size_t gmin = 0, grec = 0;
        cuMemGetAllocationGranularity(&gmin, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM);
        cuMemGetAllocationGranularity(&grec, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
        const size_t align = strtoull(getenv("RS_VMM"), 0, 10) << 20;
        printf("vmm: granularity min %zu recommended %zu, alignment %zu\n", gmin, grec, align);
        CUmemGenericAllocationHandle h;
        CUresult r = cuMemCreate(&h, bytes, &prop, 0);
        if (r != CUDA_SUCCESS) { printf("cuMemCreate %d\n", (int)r); return 1; }
        CUdeviceptr va = 0;
        r = cuMemAddressReserve(&va, bytes, align, 0, 0);
        if (r != CUDA_SUCCESS) { printf("cuMemAddressReserve %d\n", (int)r); return 1; }
        r = cuMemMap(va, bytes, 0, h, 0);
        if (r != CUDA_SUCCESS) { printf("cuMemMap %d\n", (int)r); return 1; }
        CUmemAccessDesc acc = {};
        acc.location = prop.location;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        r = cuMemSetAccess(va, bytes, &acc, 1);
        if (r != CUDA_SUCCESS) { printf("cuMemSetAccess %d\n", (int)r); return 1; }
        t = (char*)va;
    } else {
        CK(cudaMalloc(&t, bytes));
    }
    CK(cudaMemset(t, 1, bytes));
    unsigned long long* sink; CK(cudaMalloc(&sink, 8)); CK(cudaMemset(sink, 0, 8));
    const bool all = !strcmp(var, "all");
    const int v = atoi(var);
#define RUN(V) if (all || v == V) run<V>(t, bytes, wbytes, n, sink);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11)
    return 0;
}
