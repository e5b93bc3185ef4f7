// Does a bulk-async / TMA gather of 32-byte rows avoid the 128-byte L2 line fill that plain LDG
// gathers pay on B200?  Variants: 0 = LDG.256 (reference), 1 = cp.async.bulk 32 B per lane,
// 2 = cp.async.bulk.tensor.2d box {32 B, 1 row} with L2 promotion NONE, 3 = same with 128 B
// promotion, 4 = tile::gather4 (4 rows per instruction, promotion NONE).
//   ./tma_gather <table_MiB> <gathers_M> <variant>
#include <cuda.h>
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
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t phase) {
    asm volatile("{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}" ::"r"(smem_u32(b)), "r"(phase) : "memory");
}

constexpr int WARPS = 4, U = 2, RS = 128;  // row slot stride in shared memory (tensor copies need 128 B alignment)

template <int V>
__global__ void __launch_bounds__(WARPS * 32) gather(const char* __restrict__ t, const __grid_constant__ CUtensorMap tm,
                                                     uint64_t nrows, uint64_t n, uint64_t seed, unsigned long long* sink) {
    __shared__ __align__(128) unsigned char buf[WARPS][U][32][RS];
    __shared__ __align__(8) uint64_t bars[WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) mbar_init(&bars[warp], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t acc = 0;
    uint32_t phase = 0;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + warp * 32; i0 < n; i0 += U * stride) {
        if (V == 0) {
            uint64_t v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint64_t row = __umul64hi(sm64(seed + i0 + lane + u * stride), nrows);
                uint64_t a, b, c, d;
                asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(t + row * 32));
                v[u] = a ^ b ^ c ^ d;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) acc += v[u];
        } else {
            if (lane == 0) mbar_expect(&bars[warp], U * 32 * 32);
            __syncwarp();
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint64_t row = __umul64hi(sm64(seed + i0 + lane + u * stride), nrows);
                const uint32_t dst = smem_u32(&buf[warp][u][lane][0]);
                const uint32_t mb = smem_u32(&bars[warp]);
                if (V == 1) {
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 32, [%2];" ::"r"(dst), "l"(t + row * 32), "r"(mb) : "memory");
                } else if (V == 2 || V == 3) {
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(&tm), "r"(0), "r"((int)row), "r"(mb) : "memory");
                } else if (V == 4) {
                    // lanes 0..7 gather 4 rows each: 8 x 128 B = the same 32 rows per u
                    const uint64_t r1 = __shfl_sync(0xffffffffu, row, (lane * 4 + 1) & 31);
                    const uint64_t r2 = __shfl_sync(0xffffffffu, row, (lane * 4 + 2) & 31);
                    const uint64_t r3 = __shfl_sync(0xffffffffu, row, (lane * 4 + 3) & 31);
                    const uint64_t r0 = __shfl_sync(0xffffffffu, row, (lane * 4) & 31);
                    if (lane < 8) {
                        const uint32_t d4 = smem_u32(&buf[warp][u][lane][0]);
                        asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(d4), "l"(&tm), "r"(0), "r"((int)r0), "r"((int)r1), "r"((int)r2), "r"((int)r3), "r"(mb) : "memory");
                    }
                }
            }
            mbar_wait(&bars[warp], phase);
            phase ^= 1;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint64_t* p = V == 4 ? reinterpret_cast<const uint64_t*>(&buf[warp][u][lane >> 2][(lane & 3) * 32])
                                           : reinterpret_cast<const uint64_t*>(&buf[warp][u][lane][0]);
                acc += p[0] ^ p[1] ^ p[2] ^ p[3];
            }
            __syncwarp();
        }
    }
    if (acc == 0x1234567887654321ull) atomicAdd(sink, 1ull);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int V>
void run(const char* t, const CUtensorMap& tm, uint64_t nrows, uint64_t n, unsigned long long* sink, const char* name) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int blocks = 148 * 6;  // 24 warps / SM
    gather<V><<<blocks, WARPS * 32>>>(t, tm, nrows, n, 1, sink);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int it = 0; it < 3; ++it) {
        CK(cudaEventRecord(e0));
        gather<V><<<blocks, WARPS * 32>>>(t, tm, nrows, n, 1000 + it * 7919, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    printf("%-44s : %8.2f G rows/s  %8.1f GB/s\n", name, n / (best * 1e6), n * 32.0 / (best * 1e6));
    fflush(stdout);
}

int main(int argc, char** argv) {
    const uint64_t mib = argc > 1 ? strtoull(argv[1], 0, 10) : 1024;
    const uint64_t n = (argc > 2 ? strtoull(argv[2], 0, 10) : 100) * 1000000ull;
    const int v = argc > 3 ? atoi(argv[3]) : -1;
    const uint64_t bytes = mib << 20, nrows = bytes / 32;
    char* t; CK(cudaMalloc(&t, bytes)); CK(cudaMemset(t, 1, bytes));
    unsigned long long* sink; CK(cudaMalloc(&sink, 8)); CK(cudaMemset(sink, 0, 8));
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qr));
    auto make = [&](CUtensorMapL2promotion promo) {
        CUtensorMap tm;
        cuuint64_t dims[2] = {32, nrows};
        cuuint64_t strides[1] = {32};
        cuuint32_t box[2] = {32, 1};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, t, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
        return tm;
    };
    const CUtensorMap tm_none = make(CU_TENSOR_MAP_L2_PROMOTION_NONE);
    const CUtensorMap tm_128 = make(CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    printf("table %llu MiB, %llu M gathers\n", (unsigned long long)mib, (unsigned long long)(n / 1000000));
    if (v < 0 || v == 0) run<0>(t, tm_none, nrows, n, sink, "LDG.256");
    if (v < 0 || v == 1) run<1>(t, tm_none, nrows, n, sink, "cp.async.bulk 32 B");
    if (v < 0 || v == 2) run<2>(t, tm_none, nrows, n, sink, "cp.async.bulk.tensor.2d promo NONE");
    if (v < 0 || v == 3) run<3>(t, tm_128, nrows, n, sink, "cp.async.bulk.tensor.2d promo 128B");
    if (v < 0 || v == 4) run<4>(t, tm_none, nrows, n, sink, "cp.async.bulk.tensor.2d tile::gather4 NONE");
    return 0;
}
