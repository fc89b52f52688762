// Micro-benchmark (measurement aid, not product code): how fast can a B200 gather RANDOM rows of a given width from HBM?
// The IBF count kernels are random row gathers -- 512-byte rows per warp in the flat K3, 128-byte rows (top level of the
// c4 HIBF) and 8-byte rows (its 64-bin children) in K3h.  This gives the roofline that applies to each width: rows are
// drawn with splitmix64 from a 32 GiB buffer, a row of R bytes is loaded by R/16 lanes with one 128-bit
// ld.global.nc.L1::no_allocate each (8-byte rows: one lane, 64-bit load), 16 independent loads in flight per lane.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/rowgather_bench tools/rowgather_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ uint4 ld16(const void *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint2 ld8(const void *p)
{
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

template <int ROW_BYTES>
__global__ void __launch_bounds__(256) k_gather(const uint8_t *data, uint64_t n_rows, uint32_t iters, uint32_t *sink)
{
    constexpr int LPR = ROW_BYTES >= 16 ? ROW_BYTES / 16 : 1; // lanes per row
    const uint32_t lane = threadIdx.x & 31, sub = lane % LPR, grp = lane / LPR;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    uint32_t acc = 0;
    // one splitmix per (warp, group), then a 64-bit LCG per load and a fastrange multiply: the address arithmetic must stay
    // far below the issue budget, or the benchmark measures the ALUs instead of the DRAM
    uint64_t state = mix((warp << 8) ^ grp);
    for (uint32_t it = 0; it < iters; ++it)
    {
        uint4 v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
        {
            state              = state * 6364136223846793005ull + 1442695040888963407ull;
            const uint64_t row = __umul64hi(state, n_rows);
            const uint8_t *p   = data + row * ROW_BYTES + sub * 16;
            if (ROW_BYTES >= 16)
                v[j] = ld16(p);
            else
            {
                const uint2 a = ld8(data + row * ROW_BYTES);
                v[j]          = make_uint4(a.x, a.y, 0, 0);
            }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j)
            acc ^= v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
    }
    if (acc == 0x12345678u)
        sink[0] = acc;
}

template <int ROW_BYTES>
void run(const uint8_t *d, uint64_t bytes, uint32_t *sink)
{
    constexpr int LPR = ROW_BYTES >= 16 ? ROW_BYTES / 16 : 1;
    const uint64_t n_rows = bytes / ROW_BYTES;
    const int      grid = 148 * 8, block = 256;
    const uint64_t warps = (uint64_t)grid * block / 32;
    // about 24 GB of useful bytes per launch
    const uint32_t iters = (uint32_t)((24ull << 30) / (warps * 16 * (32 / LPR) * ROW_BYTES));
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k_gather<ROW_BYTES><<<grid, block>>>(d, n_rows, 4, sink);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep)
    {
        cudaEventRecord(a);
        k_gather<ROW_BYTES><<<grid, block>>>(d, n_rows, iters, sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        best = ms < best ? ms : best;
    }
    const double useful = (double)warps * iters * 16 * (32 / LPR) * ROW_BYTES;
    printf("{\"row_bytes\": %d, \"useful_GBps\": %.1f, \"ms\": %.3f, \"useful_bytes\": %.0f}\n", ROW_BYTES, useful / 1e9 / (best / 1e3), best, useful);
}

int main()
{
    const uint64_t bytes = 32ull << 30;
    uint8_t       *d = nullptr;
    uint32_t      *sink = nullptr;
    if (cudaMalloc(&d, bytes) != cudaSuccess || cudaMalloc(&sink, 64) != cudaSuccess)
    {
        printf("{\"error\": \"cudaMalloc\"}\n");
        return 1;
    }
    cudaMemset(d, 0x5a, bytes);
    run<8>(d, bytes, sink);
    run<32>(d, bytes, sink);
    run<64>(d, bytes, sink);
    run<128>(d, bytes, sink);
    run<256>(d, bytes, sink);
    run<512>(d, bytes, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess)
        printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e));
    return 0;
}
