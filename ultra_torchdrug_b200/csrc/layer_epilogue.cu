// Layer epilogue of GeneralizedRelationalConvNBF{,Mod}.combine (reference ultra/layer.py:184-190, 386-392) and
// the short-cut of the bellmanford loops (reference ultra/model.py:126-127, rel_model.py:371-372):
//     out = relu(layer_norm(x + linear_bias) * gamma + beta) + residual
// where x = cat[input, update] @ W^T stays a cuBLAS fp32 GEMM in PyTorch (north_star); the Linear's bias is added
// here because cuBLASLt runs it as a separate pass for this SIMT GEMM (profiles/: `cublasLt::epilogue::globalKernel`).  SURVEY.md section 8
// row f1: with 64 features per row PyTorch's LayerNorm kernel runs at ~7 % of HBM bandwidth and, together with
// the separate ReLU and add passes, costs more than the rspmm kernels of a layer; fused here it is one pass
// (read x, read residual, write out) at HBM speed.  The backward kernel below makes it usable under autograd
// (fine-tuning): it recomputes the row statistics from x instead of saving them.
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
                                                                  int relu, long long residual_stride4,
                                                                  long long out_stride4) {
    const int lane = threadIdx.x & 31;
    const int sub = lane % LANES;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const bool live = row < rows;
    const long long at = (live ? row : 0) * LANES + sub;
    const long long residual_at = (live ? row : 0) * residual_stride4 + sub;   // strides in float4 units
    const long long out_at = (live ? row : 0) * out_stride4 + sub;
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
        const float4 s = __ldcs(residual + residual_at);
        r = make_float4(r.x + s.x, r.y + s.y, r.z + s.z, r.w + s.w);
    }
    out[out_at] = r;
}

// Backward of the fused epilogue.  Per row (recomputing mean / rstd / xhat from x):
//   dy = relu ? (xhat * gamma + beta > 0 ? dout : 0) : dout          d_residual = dout (returned by the caller as is)
//   g  = dy * gamma;   dx = rstd * (g - mean(g) - xhat * mean(g * xhat))
//   d_gamma += dy * xhat;  d_beta += dy;  d_linear_bias += dx        (column sums over all rows)
// Rows are walked grid-stride with a fixed grid, column sums are accumulated per thread, folded per block in shared
// memory in a fixed order and written as one partial row per block; `column_sums_kernel` folds the partials in
// block order.  No atomics: gradients are bit-reproducible.
constexpr int kBackwardBlocks = 148 * 4;

template <int LANES>
__global__ void __launch_bounds__(256) norm_relu_residual_backward_kernel(
    const float4 *__restrict__ x, const float4 *__restrict__ linear_bias, const float4 *__restrict__ gamma,
    const float4 *__restrict__ beta, const float4 *__restrict__ dout, float4 *__restrict__ dx_out,
    float *__restrict__ partial, long long rows, float eps, int relu) {
    __shared__ float4 s_sum[3][256];
    const int lane = threadIdx.x & 31;
    const int sub = lane % LANES;
    constexpr int kGroups = 256 / LANES;
    const long long group = (long long)blockIdx.x * kGroups + threadIdx.x / LANES;
    const long long stride = (long long)gridDim.x * kGroups;
    constexpr float inv = 1.0f / (4 * LANES);
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 lb = linear_bias ? __ldg(linear_bias + sub) : zero;
    const float4 gm = gamma ? __ldg(gamma + sub) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 bt = beta ? __ldg(beta + sub) : zero;
    float4 acc_gamma = zero, acc_beta = zero, acc_bias = zero;
    // every group runs the same number of iterations so that the shuffles stay warp-convergent
    const long long iterations = (rows + stride - 1) / stride;
    for (long long it = 0; it < iterations; ++it) {
        const long long row = group + it * stride;
        const bool live = row < rows;
        const long long at = (live ? row : 0) * LANES + sub;
        float4 v = live ? __ldcs(x + at) : zero;
        v = make_float4(v.x + lb.x, v.y + lb.y, v.z + lb.z, v.w + lb.w);
        float sum = (v.x + v.y) + (v.z + v.w);
#pragma unroll
        for (int off = LANES / 2; off; off >>= 1) sum += __shfl_xor_sync(kFullMask, sum, off);
        const float mean = sum * inv;
        const float cx = v.x - mean, cy = v.y - mean, cz = v.z - mean, cw = v.w - mean;
        float sq = (cx * cx + cy * cy) + (cz * cz + cw * cw);
#pragma unroll
        for (int off = LANES / 2; off; off >>= 1) sq += __shfl_xor_sync(kFullMask, sq, off);
        const float rstd = rsqrtf(sq * inv + eps);
        const float4 xh = make_float4(cx * rstd, cy * rstd, cz * rstd, cw * rstd);
        float4 dy = live ? __ldcs(dout + at) : zero;
        if (relu) {
            dy.x = fmaf(xh.x, gm.x, bt.x) > 0.f ? dy.x : 0.f;
            dy.y = fmaf(xh.y, gm.y, bt.y) > 0.f ? dy.y : 0.f;
            dy.z = fmaf(xh.z, gm.z, bt.z) > 0.f ? dy.z : 0.f;
            dy.w = fmaf(xh.w, gm.w, bt.w) > 0.f ? dy.w : 0.f;
        }
        const float4 g = make_float4(dy.x * gm.x, dy.y * gm.y, dy.z * gm.z, dy.w * gm.w);
        float m1 = (g.x + g.y) + (g.z + g.w);
        float m2 = (g.x * xh.x + g.y * xh.y) + (g.z * xh.z + g.w * xh.w);
#pragma unroll
        for (int off = LANES / 2; off; off >>= 1) {
            m1 += __shfl_xor_sync(kFullMask, m1, off);
            m2 += __shfl_xor_sync(kFullMask, m2, off);
        }
        m1 *= inv;
        m2 *= inv;
        const float4 dx = make_float4(rstd * (g.x - m1 - xh.x * m2), rstd * (g.y - m1 - xh.y * m2),
                                      rstd * (g.z - m1 - xh.z * m2), rstd * (g.w - m1 - xh.w * m2));
        if (live) {
            dx_out[at] = dx;
            acc_gamma = make_float4(acc_gamma.x + dy.x * xh.x, acc_gamma.y + dy.y * xh.y, acc_gamma.z + dy.z * xh.z, acc_gamma.w + dy.w * xh.w);
            acc_beta = make_float4(acc_beta.x + dy.x, acc_beta.y + dy.y, acc_beta.z + dy.z, acc_beta.w + dy.w);
            acc_bias = make_float4(acc_bias.x + dx.x, acc_bias.y + dx.y, acc_bias.z + dx.z, acc_bias.w + dx.w);
        }
    }
    s_sum[0][threadIdx.x] = acc_gamma;
    s_sum[1][threadIdx.x] = acc_beta;
    s_sum[2][threadIdx.x] = acc_bias;
    __syncthreads();
    // thread t < 3 * LANES folds column block (t % LANES) of quantity (t / LANES) over the block's groups, in order
    if (threadIdx.x < 3 * LANES) {
        const int which = threadIdx.x / LANES, column = threadIdx.x % LANES;
        float4 total = zero;
        for (int g = 0; g < kGroups; ++g) {
            const float4 t = s_sum[which][g * LANES + column];
            total = make_float4(total.x + t.x, total.y + t.y, total.z + t.z, total.w + t.w);
        }
        reinterpret_cast<float4 *>(partial)[((long long)blockIdx.x * 3 + which) * LANES + column] = total;
    }
}

// out[c] = sum over blocks of partial[block][c]  (c < 3 * dim), fixed order
__global__ void column_sums_kernel(const float *__restrict__ partial, int blocks, int width, float *__restrict__ d_gamma,
                                   float *__restrict__ d_beta, float *__restrict__ d_bias, int dim) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= width) return;
    float total = 0.f;
    for (int b = 0; b < blocks; ++b) total += partial[(long long)b * width + c];
    float *target = c < dim ? d_gamma : (c < 2 * dim ? d_beta : d_bias);
    if (target) target[c % dim] = total;
}

// Scoring head of TransferNBFNet.forward (reference ultra/model.py:177-193): score = w2 . relu(W1 [hidden | query] + b1) + b2.
// W1 [hidden | query] = W1h hidden + W1q query, and the query part is one row per QUERY, not per (node, query): the caller
// runs z = hidden @ W1h^T as a K = d GEMM (half the flops of the K = 2d one) and passes query_bias = query @ W1q^T + b1
// (batch rows).  This kernel is the rest of the head in one pass over z: add the row's query bias, ReLU, dot with w2.
// Row r of z belongs to query r % batch (rows are (node, query) pairs, query fastest).
template <int LANES>
__global__ void __launch_bounds__(256) score_head_kernel(const float4 *__restrict__ z, const float4 *__restrict__ query_bias,
                                                          const float4 *__restrict__ weight, const float *__restrict__ bias,
                                                          float *__restrict__ score, long long rows, int batch) {
    const int lane = threadIdx.x & 31;
    const int sub = lane % LANES;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const bool live = row < rows;
    const long long safe = live ? row : 0;
    const float4 v = live ? __ldcs(z + safe * LANES + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 q = __ldg(query_bias + (safe % batch) * LANES + sub);
    const float4 w = __ldg(weight + sub);
    float dot = fmaxf(v.x + q.x, 0.f) * w.x;
    dot = fmaf(fmaxf(v.y + q.y, 0.f), w.y, dot);
    dot = fmaf(fmaxf(v.z + q.z, 0.f), w.z, dot);
    dot = fmaf(fmaxf(v.w + q.w, 0.f), w.w, dot);
#pragma unroll
    for (int off = LANES / 2; off; off >>= 1) dot += __shfl_xor_sync(kFullMask, dot, off);
    if (live && sub == 0) score[row] = dot + (bias ? __ldg(bias) : 0.f);
}

}  // namespace

}  // namespace ultra

using namespace ultra;

static int epilogue_forward(const float *dev_x, const float *dev_linear_bias, const float *dev_gamma, const float *dev_beta,
                            const float *dev_residual, float *dev_out, int64_t rows, int32_t dim, int64_t residual_stride,
                            int64_t out_stride, float eps, int32_t relu, void *stream) {
    if (rows < 0 || dim <= 0 || (rows > 0 && (!dev_x || !dev_out))) return ULTRA_RSPMM_ERR_ARG;
    if ((dev_gamma == nullptr) != (dev_beta == nullptr)) return ULTRA_RSPMM_ERR_ARG;
    if (dim % 4 || dim > 128 || (dim & (dim - 1))) return ULTRA_RSPMM_ERR_RANGE;   // 4, 8, ..., 128 features per row
    if (residual_stride % 4 || out_stride % 4 || out_stride < dim || (dev_residual && residual_stride < dim))
        return ULTRA_RSPMM_ERR_ARG;
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
    norm_relu_residual_kernel<L><<<(unsigned)blocks, 256, 0, s>>>((const float4 *)dev_x, (const float4 *)dev_linear_bias, \
                                                                    (const float4 *)dev_gamma, (const float4 *)dev_beta, \
                                                                    (const float4 *)dev_residual, (float4 *)dev_out, rows, \
                                                                    eps, relu, residual_stride / 4, out_stride / 4)
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

extern "C" int ultra_layer_norm_relu_residual(const float *dev_x, const float *dev_linear_bias, const float *dev_gamma, const float *dev_beta,
                                              const float *dev_residual, float *dev_out, int64_t rows, int32_t dim,
                                              float eps, int32_t relu, void *stream) {
    return epilogue_forward(dev_x, dev_linear_bias, dev_gamma, dev_beta, dev_residual, dev_out, rows, dim, dim, dim, eps, relu,
                            stream);
}

extern "C" int ultra_layer_norm_relu_residual_strided(const float *dev_x, const float *dev_linear_bias,
                                                      const float *dev_gamma, const float *dev_beta,
                                                      const float *dev_residual, float *dev_out, int64_t rows,
                                                      int32_t dim, int64_t residual_stride, int64_t out_stride, float eps,
                                                      int32_t relu, void *stream) {
    return epilogue_forward(dev_x, dev_linear_bias, dev_gamma, dev_beta, dev_residual, dev_out, rows, dim, residual_stride,
                            out_stride, eps, relu, stream);
}

extern "C" int ultra_layer_norm_relu_residual_backward_bytes(int32_t dim, size_t *workspace_bytes) {
    if (!workspace_bytes || dim <= 0) return ULTRA_RSPMM_ERR_ARG;
    *workspace_bytes = (size_t)kBackwardBlocks * 3 * dim * sizeof(float);
    return ULTRA_RSPMM_OK;
}

extern "C" int ultra_layer_norm_relu_residual_backward(const float *dev_x, const float *dev_linear_bias,
                                                       const float *dev_gamma, const float *dev_beta,
                                                       const float *dev_grad_out, float *dev_grad_x,
                                                       float *dev_grad_linear_bias, float *dev_grad_gamma,
                                                       float *dev_grad_beta, int64_t rows, int32_t dim, float eps,
                                                       int32_t relu, void *workspace, size_t workspace_bytes,
                                                       void *stream) {
    if (rows < 0 || dim <= 0 || (rows > 0 && (!dev_x || !dev_grad_out || !dev_grad_x))) return ULTRA_RSPMM_ERR_ARG;
    if ((dev_gamma == nullptr) != (dev_beta == nullptr)) return ULTRA_RSPMM_ERR_ARG;
    if (dim % 4 || dim > 128 || (dim & (dim - 1))) return ULTRA_RSPMM_ERR_RANGE;
    if (((uintptr_t)dev_x | (uintptr_t)dev_grad_out | (uintptr_t)dev_grad_x | (uintptr_t)dev_gamma | (uintptr_t)dev_beta |
         (uintptr_t)dev_linear_bias | (uintptr_t)workspace) & 15)
        return ULTRA_RSPMM_ERR_ARG;
    const size_t need = (size_t)kBackwardBlocks * 3 * dim * sizeof(float);
    if (!workspace || workspace_bytes < need) return ULTRA_RSPMM_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    const int lanes = dim / 4;
#define ULTRA_LAUNCH(L)                                                                                        \
    norm_relu_residual_backward_kernel<L><<<kBackwardBlocks, 256, 0, s>>>(                                     \
        (const float4 *)dev_x, (const float4 *)dev_linear_bias, (const float4 *)dev_gamma, (const float4 *)dev_beta, \
        (const float4 *)dev_grad_out, (float4 *)dev_grad_x, (float *)workspace, rows, eps, relu)
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
    column_sums_kernel<<<(3 * dim + 127) / 128, 128, 0, s>>>((const float *)workspace, kBackwardBlocks, 3 * dim,
                                                             dev_grad_gamma, dev_grad_beta, dev_grad_linear_bias, dim);
    note_launch();
    ULTRA_CUDA_OK(cudaGetLastError());
    return ULTRA_RSPMM_OK;
}

extern "C" int ultra_score_head(const float *dev_z, const float *dev_query_bias, const float *dev_weight, const float *dev_bias,
                                float *dev_score, int64_t rows, int32_t batch, int32_t dim, void *stream) {
    if (rows < 0 || batch <= 0 || dim <= 0 || (rows > 0 && (!dev_z || !dev_query_bias || !dev_weight || !dev_score)))
        return ULTRA_RSPMM_ERR_ARG;
    if (dim % 4 || dim > 128 || (dim & (dim - 1))) return ULTRA_RSPMM_ERR_RANGE;   // 4, 8, ..., 128 features per row
    if (((uintptr_t)dev_z | (uintptr_t)dev_query_bias | (uintptr_t)dev_weight) & 15) return ULTRA_RSPMM_ERR_ARG;
    if (rows == 0) return ULTRA_RSPMM_OK;
    const int lanes = dim / 4;
    const long long blocks = (rows * lanes + 255) / 256;
    if (blocks > 0x7fffffffLL) return ULTRA_RSPMM_ERR_RANGE;
    cudaStream_t s = (cudaStream_t)stream;
#define ULTRA_LAUNCH(L)                                                                                                    \
    score_head_kernel<L><<<(unsigned)blocks, 256, 0, s>>>((const float4 *)dev_z, (const float4 *)dev_query_bias,          \
                                                           (const float4 *)dev_weight, dev_bias, dev_score, rows, batch)
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
