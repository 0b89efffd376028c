// Gather-bandwidth probe (diagnostics; bench.py measures its rooflines with it on the box it runs on).
// Every warp reads 512-byte row pieces (one LDG.128 per lane, the access of the rspmm gather) at pseudo-random rows of
// a caller-provided buffer, without index loads or arithmetic.  With a footprint (rows x 512 B) well inside the 126 MB
// L2 the result is the L2 -> SM ceiling of any gather kernel; with a footprint far beyond it, the HBM ceiling for
// random 512-byte rows.  row_stride_bytes = 512 packs the rows; 4 * D reproduces the column slab of a (rows, D) matrix.
#include "rspmm_common.cuh"

namespace ultra {
namespace {

__global__ void __launch_bounds__(256) gather_probe_kernel(const char *__restrict__ buffer, unsigned rows, long long row_stride,
                                                           int iters, float *sink) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned long long state = (unsigned long long)warp * 0x9e3779b97f4a7c15ULL + 12345;
    float acc = 0.f;
    for (int it = 0; it < iters; it += 4) {
        Vec<float, 4> v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            state = state * 6364136223846793005ULL + 1442695040888963407ULL;
            const long long row = (long long)__umulhi((unsigned)(state >> 32), rows);
            const float *p = reinterpret_cast<const float *>(buffer + row * row_stride) + lane * 4;
            asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(v[u].v[0]), "=f"(v[u].v[1]), "=f"(v[u].v[2]), "=f"(v[u].v[3]) : "l"(p));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc += v[u].v[0] + v[u].v[1] + v[u].v[2] + v[u].v[3];
    }
    if (acc == 123.456f) sink[0] = acc;   // never true for the zero / random buffers used; keeps the loads alive
}

}  // namespace
}  // namespace ultra

using namespace ultra;

extern "C" int ultra_probe_gather(const void *dev_buffer, int64_t rows, int64_t row_stride_bytes, int32_t iters,
                                  int32_t blocks, float *dev_sink, int64_t *bytes_read, void *stream) {
    if (!dev_buffer || !dev_sink || rows <= 0 || rows > 0xffffffffLL || row_stride_bytes < 512 || row_stride_bytes % 16 ||
        iters <= 0 || iters % 4 || blocks <= 0)
        return ULTRA_RSPMM_ERR_ARG;
    gather_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const char *)dev_buffer, (unsigned)rows, row_stride_bytes,
                                                                  iters, dev_sink);
    note_launch();
    if (bytes_read) *bytes_read = (int64_t)blocks * 8 * iters * 512;
    ULTRA_CUDA_OK(cudaGetLastError());
    return ULTRA_RSPMM_OK;
}
