// Random 32-byte gathers from a PEER GPU's memory over NVLink (single process, 2 GPUs).
//   ./peer_gather <table_MiB> <gathers_M>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint64_t sm64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}
template <int V>
__global__ void gather(const char* __restrict__ t, uint64_t nslots, uint64_t n, uint64_t seed, unsigned long long* sink) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 4 * stride) {
        uint64_t v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t g = i + u * stride;
            const char* p = t + __umul64hi(sm64(seed + g), nslots) * 32;
            uint64_t a = 0, b = 0, c = 0, d = 0;
            if (g < n) {
                if (V == 0) asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
                if (V == 1) asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
                if (V == 2) { asm volatile("ld.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p)); asm volatile("ld.global.v2.u64 {%0,%1}, [%2];" : "=l"(c), "=l"(d) : "l"(p + 16)); }
                if (V == 3) asm volatile("ld.global.relaxed.sys.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
                if (V == 4) asm volatile("ld.global.u64 %0, [%1];" : "=l"(a) : "l"(p));
            }
            v[u] = a ^ b ^ c ^ d;
        }
        acc += v[0] + v[1] + v[2] + v[3];
    }
    if (acc == 0x1234567887654321ull) atomicAdd(sink, 1ull);
}
template <int V>
void run(const char* name, const char* t, uint64_t bytes, uint64_t n, unsigned long long* sink) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    gather<V><<<148 * 8, 256>>>(t, bytes / 32, n / 8, 1, sink);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    gather<V><<<148 * 8, 256>>>(t, bytes / 32, n, 7, sink);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-34s : %8.3f G sectors/s  %8.1f GB/s\n", name, n / (ms * 1e6), n * 32.0 / (ms * 1e6)); fflush(stdout);
}
int main(int argc, char** argv) {
    const uint64_t mib = argc > 1 ? strtoull(argv[1], 0, 10) : 4096;
    const uint64_t n = (argc > 2 ? strtoull(argv[2], 0, 10) : 50) * 1000000ull;
    int ndev = 0; CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) { printf("needs 2 GPUs\n"); return 0; }
    int can = 0; CK(cudaDeviceCanAccessPeer(&can, 0, 1)); printf("canAccessPeer(0,1) = %d\n", can);
    int attr = 0; CK(cudaDeviceGetP2PAttribute(&attr, cudaDevP2PAttrPerformanceRank, 0, 1)); printf("P2P performance rank %d\n", attr);
    CK(cudaDeviceGetP2PAttribute(&attr, cudaDevP2PAttrNativeAtomicSupported, 0, 1)); printf("native atomics %d\n", attr);
    const uint64_t bytes = mib << 20;
    char *local, *remote;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&remote, bytes)); CK(cudaMemset(remote, 1, bytes)); CK(cudaDeviceSynchronize());
    CK(cudaSetDevice(0)); CK(cudaMalloc(&local, bytes)); CK(cudaMemset(local, 1, bytes));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    unsigned long long* sink; CK(cudaMalloc(&sink, 8)); CK(cudaMemset(sink, 0, 8));
    // streaming copy for reference
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaMemcpy(local, remote, bytes, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e0)); CK(cudaMemcpy(local, remote, bytes, cudaMemcpyDeviceToDevice)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); printf("peer memcpy %llu MiB: %.1f GB/s\n", (unsigned long long)mib, bytes / (ms * 1e6));
    run<0>("local  nc.L1no_alloc.v4u64", local, bytes, n, sink);
    run<0>("remote nc.L1no_alloc.v4u64", remote, bytes, n, sink);
    run<1>("remote plain v4u64", remote, bytes, n, sink);
    run<2>("remote plain 2x v2u64", remote, bytes, n, sink);
    run<3>("remote relaxed.sys v2u64 (16 B)", remote, bytes, n, sink);
    run<4>("remote plain u64 (8 B)", remote, bytes, n, sink);
    return 0;
}
