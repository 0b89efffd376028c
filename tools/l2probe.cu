// L2 -> SM read-bandwidth probe (development tool, not part of the product path).
// Every warp reads 512-byte rows (one float4 per lane) at pseudo-random row positions of a buffer of
// `mb` megabytes; for mb << 126 the rows are L2 hits, for mb >> 126 they come from HBM.  This is the
// access pattern of the rspmm gather, without index loads or arithmetic: it bounds what any gather
// kernel can reach (DESIGN.md "Rooflines").
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int UNROLL, bool NOALLOC>
__global__ void __launch_bounds__(256) probe(const float4 *__restrict__ buf, long long rows, int iters, float *sink, long long row_stride = 32) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned long long state = warp * 0x9e3779b97f4a7c15ULL + 12345;
    float acc = 0.f;
    for (int it = 0; it < iters; it += UNROLL) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            state = state * 6364136223846793005ULL + 1442695040888963407ULL;
            const long long row = (long long)__umulhi((unsigned)(state >> 32), (unsigned)rows);  // cheap range reduction
            const float4 *p = buf + row * row_stride + lane;
            if (NOALLOC)
                asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                             : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(p));
            else
                v[u] = __ldg(p);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (acc == 123.456f) sink[0] = acc;
}

int main(int argc, char **argv) {
    const int sizes[] = {4, 8, 32, 64, 96, 240, 2048};
    float *sink;
    cudaMalloc(&sink, 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    printf("mb,blocks_per_sm,unroll,no_allocate,GBps\n");
    for (int mb : sizes) {
        float4 *buf;
        const size_t bytes = (size_t)mb << 20;
        if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed %d\n", mb); return 1; }
        cudaMemset(buf, 0, bytes);
        const long long rows = bytes / 512;
        for (int bps : {4, 8}) {
            const int blocks = 148 * bps, iters = 4096;
            for (int variant = 0; variant < 3; ++variant) {
                auto run = [&]() {
                    if (variant == 0) probe<4, true><<<blocks, 256>>>(buf, rows, iters, sink);
                    else if (variant == 1) probe<8, true><<<blocks, 256>>>(buf, rows, iters, sink);
                    else probe<4, false><<<blocks, 256>>>(buf, rows, iters, sink);
                };
                run(); run();
                cudaEventRecord(a);
                for (int r = 0; r < 5; ++r) run();
                cudaEventRecord(b);
                cudaEventSynchronize(b);
                float ms;
                cudaEventElapsedTime(&ms, a, b);
                const double gb = 5.0 * blocks * 8 * (double)iters * 512 / 1e9;
                printf("%d,%d,%d,%d,%.1f\n", mb, bps, variant == 1 ? 8 : 4, variant != 2, gb / (ms / 1e3));
            }
        }
        cudaFree(buf);
    }
    // strided slabs: 512 B of every 16 KB row (the column block of a (rows, 4096) fp32 matrix), footprint = rows * 512 B
    printf("strided: footprint_mb,rows,GBps\n");
    for (int mb : {8, 16, 32, 48, 64, 80, 96}) {
        const long long rows = ((long long)mb << 20) / 512;
        float4 *buf;
        if (cudaMalloc(&buf, rows * 16384) != cudaSuccess) { printf("alloc failed %d\n", mb); return 1; }
        cudaMemset(buf, 0, rows * 16384);
        const int blocks = 148 * 8, iters = 4096;
        for (int r = 0; r < 2; ++r) probe<4, true><<<blocks, 256>>>(buf, rows, iters, sink, 1024);
        cudaEventRecord(a);
        for (int r = 0; r < 5; ++r) probe<4, true><<<blocks, 256>>>(buf, rows, iters, sink, 1024);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        printf("%d,%lld,%.1f\n", mb, rows, 5.0 * blocks * 8 * (double)iters * 512 / 1e9 / (ms / 1e3));
        cudaFree(buf);
    }
    return cudaGetLastError() != cudaSuccess;
}
