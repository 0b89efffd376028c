// HBM bandwidth of random row PIECES (development tool, not part of the product path).
// A (rows, 4096) fp32 matrix is gathered in column slabs: each access reads `piece` bytes (128 ... 1024) at a fixed column
// offset of a pseudo-random row (row stride 16 KB), 32 * 16 / piece rows per warp instruction.  With rows * piece >> 126 MB
// the pieces come from DRAM: this measures what slab width a gather kernel needs before DRAM stops being activate-bound.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int LANES, int UNROLL>   // LANES lanes (16 bytes each) share one row piece
__global__ void __launch_bounds__(256) probe(const char *__restrict__ buf, unsigned rows, int iters, float *sink) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int group = lane / LANES, l = lane % LANES;
    unsigned long long state = (warp * (32 / LANES) + group) * 0x9e3779b97f4a7c15ULL + 12345;
    float acc = 0.f;
    for (int it = 0; it < iters; it += UNROLL) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            state = state * 6364136223846793005ULL + 1442695040888963407ULL;
            const long long row = (long long)__umulhi((unsigned)(state >> 32), rows);
            const char *p = buf + row * 16384 + l * 16;
            asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(p));
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (acc == 123.456f) sink[0] = acc;
}

template <int LANES, int UNROLL> double run(const char *buf, unsigned rows, float *sink, int blocks) {
    const int iters = 1024;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    probe<LANES, UNROLL><<<blocks, 256>>>(buf, rows, iters, sink);
    cudaEventRecord(a);
    for (int r = 0; r < 3; ++r) probe<LANES, UNROLL><<<blocks, 256>>>(buf, rows, iters, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return 3.0 * blocks * 8 * (double)iters * 512 / 1e9 / (ms / 1e3);
}

int main() {
    const unsigned rows = 2u << 20;                       // 2 M rows x 16 KB = 32 GB; a 256-byte slab of it is 512 MB
    char *buf;
    float *sink;
    if (cudaMalloc(&buf, (size_t)rows * 16384) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(buf, 0, (size_t)rows * 16384);
    cudaMalloc(&sink, 4);
    printf("piece_bytes,blocks_per_sm,loads_in_flight_per_lane,GBps\n");
    for (int bps : {4, 8}) {
        const int blocks = 148 * bps;
        printf("128,%d,4,%.1f\n", bps, run<8, 4>(buf, rows, sink, blocks));
        printf("128,%d,8,%.1f\n", bps, run<8, 8>(buf, rows, sink, blocks));
        printf("256,%d,4,%.1f\n", bps, run<16, 4>(buf, rows, sink, blocks));
        printf("256,%d,8,%.1f\n", bps, run<16, 8>(buf, rows, sink, blocks));
        printf("512,%d,4,%.1f\n", bps, run<32, 4>(buf, rows, sink, blocks));
        printf("512,%d,8,%.1f\n", bps, run<32, 8>(buf, rows, sink, blocks));
    }
    return cudaGetLastError() != cudaSuccess;
}
