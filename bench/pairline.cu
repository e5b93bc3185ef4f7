// Two small loads from one random 128-byte line: does the second one cost a DRAM access of its own?
// (the question behind the pair filter of table.cuh: an even and the following odd sampled position ask one line)
//   ./pairline <table_MiB> <lines_M>
// variant 0: one u32 per random line; 1: two u32, byte offsets 0 and 64 (different halves); 2: 0 and 32 (same half,
// different sectors); 3: 0 and 16 (same sector); 4: two u32 from two different random lines.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t sm64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ uint32_t ld32(const char* p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

template <int V, int U>
__global__ void gather(const char* __restrict__ t, uint64_t nlines, uint64_t n, uint64_t seed, unsigned long long* sink) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += U * stride) {
        uint32_t v[U], w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t g = i + u * stride;
            const uint64_t r = sm64(seed + g);
            const char* p = t + __umul64hi(r, nlines) * 128;
            const uint32_t o = (uint32_t)(r & 3u) * 4;  // some word of the chosen part
            v[u] = w[u] = 0;
            if (g < n) {
                v[u] = ld32(p + o);
                if (V == 1) w[u] = ld32(p + 64 + o);
                if (V == 2) w[u] = ld32(p + 32 + o);
                if (V == 3) w[u] = ld32(p + 16 + o);
                if (V == 4) w[u] = ld32(t + __umul64hi(sm64(r), nlines) * 128 + o);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u] ^ w[u];
    }
    if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

template <int V>
void run(const char* t, uint64_t bytes, uint64_t n, unsigned long long* sink, const char* name) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    gather<V, 8><<<148 * 8, 256>>>(t, bytes / 128, n, 1, sink);
    float best = 1e30f;
    for (int it = 0; it < 3; ++it) {
        CK(cudaEventRecord(e0));
        gather<V, 8><<<148 * 8, 256>>>(t, bytes / 128, n, 1000 + it * 7919, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    printf("table %6llu MiB  %-44s : %7.2f G lines/s (%.3f ms)\n", (unsigned long long)(bytes >> 20), name, n / (best * 1e6), best);
    fflush(stdout);
}

int main(int argc, char** argv) {
    const uint64_t mib = argc > 1 ? strtoull(argv[1], 0, 10) : 4096;
    const uint64_t n = (argc > 2 ? strtoull(argv[2], 0, 10) : 128) * 1000000ull;
    char* t;
    CK(cudaMalloc(&t, mib << 20));
    CK(cudaMemset(t, 1, mib << 20));
    unsigned long long* sink; CK(cudaMalloc(&sink, 8)); CK(cudaMemset(sink, 0, 8));
    run<0>(t, mib << 20, n, sink, "one u32 per random line");
    run<1>(t, mib << 20, n, sink, "two u32, offsets 0 and 64 of one line");
    run<2>(t, mib << 20, n, sink, "two u32, offsets 0 and 32 of one line");
    run<3>(t, mib << 20, n, sink, "two u32, offsets 0 and 16 (one sector)");
    run<4>(t, mib << 20, n, sink, "two u32 from two random lines");
    return 0;
}
