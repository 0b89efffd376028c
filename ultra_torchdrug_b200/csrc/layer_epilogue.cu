// Layer epilogue of GeneralizedRelationalConvNBF{,Mod}.combine (reference ultra/layer.py:184-190, 386-392) and
// the short-cut of the bellmanford loops (reference ultra/model.py:126-127, rel_model.py:371-372):
//     out = relu(layer_norm(x + linear_bias) * gamma + beta) + residual
// where x = cat[input, update] @ W^T stays a cuBLAS fp32 GEMM in PyTorch (north_star); the Linear's bias is added
// here because cuBLASLt runs it as a separate pass for this SIMT GEMM (profiles/: `cublasLt::epilogue::globalKernel`).  SURVEY.md section 8
// row f1: with 64 features per row PyTorch's LayerNorm kernel runs at ~7 % of HBM bandwidth and, together with
// the separate ReLU and add passes, costs more than the rspmm kernels of a layer; fused here it is one pass
// (read x, read residual, write out) at HBM speed.  Inference path (no autograd).
#include "rspmm_common.cuh"

namespace ultra {

namespace {

// LANES lanes of a warp own one row of `dim = 4 * LANES` floats (float4 per lane); 32 / LANES rows per warp.
template <int LANES>
__global__ void __launch_bounds__(256) norm_relu_residual_kernel(const float4 *__restrict__ x,
                                                                  const float4 *__restrict__ linear_bias,
                                                                  const float4 *__restrict__ gamma,
                                                                  const float4 *__restrict__ beta,
                                                                  const float4 *__restrict__ residual,
                                                                  float4 *__restrict__ out, long long rows, float eps,
                                                                  int relu) {
    const int lane = threadIdx.x & 31;
    const int sub = lane % LANES;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const bool live = row < rows;
    const long long at = (live ? row : 0) * LANES + sub;
    float4 v = live ? __ldcs(x + at) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (linear_bias) {
        const float4 b = __ldg(linear_bias + sub);
        v = make_float4(v.x + b.x, v.y + b.y, v.z + b.z, v.w + b.w);
    }
    constexpr float inv = 1.0f / (4 * LANES);
    float sum = (v.x + v.y) + (v.z + v.w);
#pragma unroll
    for (int off = LANES / 2; off; off >>= 1) sum += __shfl_xor_sync(kFullMask, sum, off);
    const float mean = sum * inv;
    const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
    float sq = (dx * dx + dy * dy) + (dz * dz + dw * dw);
#pragma unroll
    for (int off = LANES / 2; off; off >>= 1) sq += __shfl_xor_sync(kFullMask, sq, off);
    const float rstd = rsqrtf(sq * inv + eps);
    if (!live) return;
    float4 r = make_float4(dx * rstd, dy * rstd, dz * rstd, dw * rstd);
    if (gamma) {
        const float4 g = __ldg(gamma + sub), b = __ldg(beta + sub);
        r = make_float4(fmaf(r.x, g.x, b.x), fmaf(r.y, g.y, b.y), fmaf(r.z, g.z, b.z), fmaf(r.w, g.w, b.w));
    }
    if (relu) r = make_float4(fmaxf(r.x, 0.f), fmaxf(r.y, 0.f), fmaxf(r.z, 0.f), fmaxf(r.w, 0.f));
    if (residual) {
        const float4 s = __ldcs(residual + at);
        r = make_float4(r.x + s.x, r.y + s.y, r.z + s.z, r.w + s.w);
    }
    out[at] = r;
}

}  // namespace

}  // namespace ultra

using namespace ultra;

extern "C" int ultra_layer_norm_relu_residual(const float *dev_x, const float *dev_linear_bias, const float *dev_gamma, const float *dev_beta,
                                              const float *dev_residual, float *dev_out, int64_t rows, int32_t dim,
                                              float eps, int32_t relu, void *stream) {
    if (rows < 0 || dim <= 0 || (rows > 0 && (!dev_x || !dev_out))) return ULTRA_RSPMM_ERR_ARG;
    if ((dev_gamma == nullptr) != (dev_beta == nullptr)) return ULTRA_RSPMM_ERR_ARG;
    if (dim % 4 || dim > 128 || (dim & (dim - 1))) return ULTRA_RSPMM_ERR_RANGE;   // 4, 8, ..., 128 features per row
    if (((uintptr_t)dev_x | (uintptr_t)dev_out | (uintptr_t)dev_gamma | (uintptr_t)dev_beta | (uintptr_t)dev_residual |
         (uintptr_t)dev_linear_bias) & 15)
        return ULTRA_RSPMM_ERR_ARG;
    if (rows == 0) return ULTRA_RSPMM_OK;
    const int lanes = dim / 4;
    const long long threads = rows * lanes;
    const long long blocks = (threads + 255) / 256;
    if (blocks > 0x7fffffffLL) return ULTRA_RSPMM_ERR_RANGE;
    cudaStream_t s = (cudaStream_t)stream;
#define ULTRA_LAUNCH(L)                                                                                              \
    norm_relu_residual_kernel<L><<<(unsigned)blocks, 256, 0, s>>>((const float4 *)dev_x, (const float4 *)dev_linear_bias, (const float4 *)dev_gamma, \
                                                                    (const float4 *)dev_beta, (const float4 *)dev_residual, \
                                                                    (float4 *)dev_out, rows, eps, relu)
    switch (lanes) {
        case 1: ULTRA_LAUNCH(1); break;
        case 2: ULTRA_LAUNCH(2); break;
        case 4: ULTRA_LAUNCH(4); break;
        case 8: ULTRA_LAUNCH(8); break;
        case 16: ULTRA_LAUNCH(16); break;
        default: ULTRA_LAUNCH(32); break;
    }
#undef ULTRA_LAUNCH
    note_launch();
    ULTRA_CUDA_OK(cudaGetLastError());
    return ULTRA_RSPMM_OK;
}
