// Fused `combine` of GeneralizedRelationalConvNBF{,Mod} for inference (reference ultra/layer.py:184-190, 386-392 and
// the short-cut of ultra/model.py:126-127):
//     out[r, 0:N] = relu(layer_norm(A[r, 0:2N] @ W^T + b) * gamma + beta) + A[r, 0:N]
// where a row of A is [layer input | update + boundary] - the (N_nodes * B, 2N) layer buffer the blocked operator writes
// (ultra_rspmm_forward_blocked) - and W is the layer's Linear weight (N, 2N).  SURVEY.md section 8 row f1.
//
// Why: as separate passes the Linear (cuBLAS SIMT SGEMM, 0.36 ms at the C2 shape) and the LayerNorm epilogue (0.12 ms)
// move 1.43 GB per layer; fused, a row is read once and written once (0.71 GB, 0.11 ms at HBM speed), which needs the
// 15 GFLOP of the Linear in well under that time - out of reach of the fp32 FMA pipe (0.2 ms at its peak).
//
// Arithmetic: fp32 accuracy on the tensor cores by the 3xTF32 split.  Each fp32 operand x is split into
// hi = tf32(x) and lo = tf32(x - hi) (together 21+ mantissa bits), and a . b is accumulated in fp32 as
// lo_a hi_b + hi_a lo_b + hi_a hi_b (the dropped lo_a lo_b term is below 2^-22 relative).  This is NOT the TF32 mode the
// reference switches off (script/run_full.py:19-20: a single 10-bit-mantissa product); tests/test_rspmm_gpu.py bounds the
// error against a float64 Linear at the level of cuBLAS's fp32 SGEMM.  ULTRA_FUSED_LINEAR=0 keeps the cuBLAS path.
//
// Shape: persistent CTAs (one per SM, 16 warps).  W is split once per CTA into shared memory in mma-fragment order.  Each
// warp then runs its own pipeline over 16-row tiles - cp.async its rows into its own 8 KB of shared memory (zero-filled
// past the last row), 3 x N/8 x K/8 MMAs (m16n8k8, all N columns), LayerNorm epilogue in registers (a row's statistics
// live in the 4 lanes that share it) - with no block-wide barrier, so the loads and epilogues of some warps overlap the
// MMAs of the others (the first version, 128-row tiles behind __syncthreads, left the tensor pipe 53 % idle).
#include "rspmm_common.cuh"

namespace ultra {

namespace {

constexpr int kDefaultLinearKernel = 2;                // 1 = mma.sync (this file), 2 = tcgen05 + TMA (layer_linear_tc.cu)
constexpr int kWarpRows = 16;                          // rows per warp tile (one m16 MMA row block)
constexpr int kLinearWarps = 16;
constexpr int kLinearThreads = 32 * kLinearWarps;

template <int N> struct LinearShape {
    static constexpr int K = 2 * N;
    static constexpr int kPad = K + 4;                 // staged row stride (floats): fragment reads hit 32 distinct banks
    static constexpr int kSteps = K / 8;
    static constexpr int kNTiles = N / 8;
    static constexpr size_t kWeightBytes = (size_t)kSteps * kNTiles * 32 * sizeof(float4);
    static constexpr size_t kTileBytes = (size_t)kWarpRows * kPad * sizeof(float);
    static constexpr size_t kSmemBytes = kWeightBytes + kLinearWarps * kTileBytes;
};

__device__ __forceinline__ float to_tf32(float x) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__device__ __forceinline__ void cp_async_16(void *smem, const void *global, unsigned bytes) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(global), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// D(16x8, fp32) += A(16x8, tf32, row) * B(8x8, tf32, col).  Lane l: g = l / 4, t = l % 4.
//   a0 = A[g][t], a1 = A[g+8][t], a2 = A[g][t+4], a3 = A[g+8][t+4];  b0 = B[t][g], b1 = B[t+4][g];
//   d0 = D[g][2t], d1 = D[g][2t+1], d2 = D[g+8][2t], d3 = D[g+8][2t+1].
__device__ __forceinline__ void mma_tf32(float (&d)[4], const float (&a)[4], float b0, float b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])),
                   "r"(__float_as_uint(a[3])), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

template <int N>
__global__ void __launch_bounds__(kLinearThreads, 1)
linear_norm_relu_residual_kernel(const float *__restrict__ A, long long lda, const float *__restrict__ W,
                                 const float *__restrict__ linear_bias, const float *__restrict__ gamma,
                                 const float *__restrict__ beta, float *__restrict__ out, long long ldo, long long rows,
                                 float eps, int relu, int shortcut) {
    using Shape = LinearShape<N>;
    constexpr int K = Shape::K, kPad = Shape::kPad, kSteps = Shape::kSteps, kNTiles = Shape::kNTiles;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *w_frag = reinterpret_cast<float4 *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    float *mine = reinterpret_cast<float *>(smem_raw + Shape::kWeightBytes) + warp * (kWarpRows * kPad);   // this warp's rows

    // W (N, K) row-major -> fragment order [k-step][n-tile][lane] = (b0 hi, b1 hi, b0 lo, b1 lo), once per CTA
    for (int idx = tid; idx < kSteps * kNTiles * 32; idx += kLinearThreads) {
        const int l = idx & 31, j = (idx >> 5) % kNTiles, s = idx / (32 * kNTiles);
        const int n = 8 * j + (l >> 2), k = 8 * s + (l & 3);
        const float b0 = __ldg(W + n * K + k), b1 = __ldg(W + n * K + k + 4);
        const float b0_hi = to_tf32(b0), b1_hi = to_tf32(b1);
        w_frag[idx] = make_float4(b0_hi, b1_hi, to_tf32(b0 - b0_hi), to_tf32(b1 - b1_hi));
    }
    __syncthreads();                                   // the only block-wide barrier: from here on warps run on their own

    // Every warp is its own pipeline over 16-row tiles (load -> 3 x N/8 x K/8 MMAs -> epilogue): the warps of an SM drift
    // out of phase, so loads and epilogues of some overlap the MMAs of others without any block-wide synchronisation.
    constexpr float inv = 1.0f / N;
    constexpr int kChunks = K / 4;                     // 16-byte chunks per row
    const long long n_tiles = (rows + kWarpRows - 1) / kWarpRows;
    const long long stride = (long long)gridDim.x * kLinearWarps;
    for (long long tile = (long long)blockIdx.x * kLinearWarps + warp; tile < n_tiles; tile += stride) {
        const long long row0 = tile * kWarpRows;
#pragma unroll 4
        for (int c = lane; c < kWarpRows * kChunks; c += 32) {
            const int r = c / kChunks, q = c % kChunks;
            const bool live = row0 + r < rows;
            cp_async_16(mine + r * kPad + 4 * q, A + (live ? row0 + r : 0) * lda + 4 * q, live ? 16u : 0u);
        }
        cp_async_wait_all();
        __syncwarp();

        float acc[kNTiles][4];
#pragma unroll
        for (int j = 0; j < kNTiles; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
        const float *row_upper = mine + g * kPad + t;
        const float *row_lower = row_upper + 8 * kPad;
#pragma unroll 2
        for (int s = 0; s < kSteps; ++s) {
            const float a[4] = {row_upper[8 * s], row_lower[8 * s], row_upper[8 * s + 4], row_lower[8 * s + 4]};
            float big[4], small[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                big[i] = to_tf32(a[i]);
                small[i] = to_tf32(a[i] - big[i]);
            }
            const float4 *w = w_frag + s * kNTiles * 32 + lane;
            float4 b[kNTiles];
#pragma unroll
            for (int j = 0; j < kNTiles; ++j) b[j] = w[j * 32];
#pragma unroll
            for (int j = 0; j < kNTiles; ++j) mma_tf32(acc[j], small, b[j].x, b[j].y);   // lo_a hi_b
#pragma unroll
            for (int j = 0; j < kNTiles; ++j) mma_tf32(acc[j], big, b[j].z, b[j].w);     // hi_a lo_b
#pragma unroll
            for (int j = 0; j < kNTiles; ++j) mma_tf32(acc[j], big, b[j].x, b[j].y);     // hi_a hi_b
        }

        // epilogue: this thread holds columns {8j + 2t, 8j + 2t + 1} of tile rows g (acc[j][0..1]) and g + 8
        // (acc[j][2..3]); the 4 lanes with the same g hold a full row between them
        float sum[2] = {0.f, 0.f};
#pragma unroll
        for (int j = 0; j < kNTiles; ++j) {
            if (linear_bias) {
                const float2 lb = __ldg(reinterpret_cast<const float2 *>(linear_bias + 8 * j + 2 * t));
                acc[j][0] += lb.x; acc[j][1] += lb.y; acc[j][2] += lb.x; acc[j][3] += lb.y;
            }
            sum[0] += acc[j][0] + acc[j][1];
            sum[1] += acc[j][2] + acc[j][3];
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            sum[h] += __shfl_xor_sync(kFullMask, sum[h], 1);
            sum[h] += __shfl_xor_sync(kFullMask, sum[h], 2);
        }
        const float mean[2] = {sum[0] * inv, sum[1] * inv};
        float sq[2] = {0.f, 0.f};
#pragma unroll
        for (int j = 0; j < kNTiles; ++j) {
            acc[j][0] -= mean[0]; acc[j][1] -= mean[0]; acc[j][2] -= mean[1]; acc[j][3] -= mean[1];
            sq[0] += acc[j][0] * acc[j][0] + acc[j][1] * acc[j][1];
            sq[1] += acc[j][2] * acc[j][2] + acc[j][3] * acc[j][3];
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            sq[h] += __shfl_xor_sync(kFullMask, sq[h], 1);
            sq[h] += __shfl_xor_sync(kFullMask, sq[h], 2);
        }
        const float rstd[2] = {rsqrtf(sq[0] * inv + eps), rsqrtf(sq[1] * inv + eps)};
#pragma unroll
        for (int j = 0; j < kNTiles; ++j) {
            const int col = 8 * j + 2 * t;
            float2 scale = make_float2(1.f, 1.f), shift = make_float2(0.f, 0.f);
            if (gamma) {
                scale = __ldg(reinterpret_cast<const float2 *>(gamma + col));
                shift = __ldg(reinterpret_cast<const float2 *>(beta + col));
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float2 y = make_float2(fmaf(acc[j][2 * h] * rstd[h], scale.x, shift.x),
                                       fmaf(acc[j][2 * h + 1] * rstd[h], scale.y, shift.y));
                if (relu) y = make_float2(fmaxf(y.x, 0.f), fmaxf(y.y, 0.f));
                if (shortcut) {
                    const float2 skip = *reinterpret_cast<const float2 *>(mine + (g + 8 * h) * kPad + col);
                    y.x += skip.x; y.y += skip.y;
                }
                const long long row = row0 + g + 8 * h;
                if (row < rows) *reinterpret_cast<float2 *>(out + row * ldo + col) = y;
            }
        }
        __syncwarp();                                  // all lanes are done with the staged rows before the next load
    }
}

// ---- scoring head (reference ultra/model.py:177-193): score[r] = b2 + w2 . relu(W1h hidden[r] + query_bias[r % batch]) -------
// The K = d GEMM of the split head (nbf._split_head) and ultra_score_head's pass fused: the (rows, H) activations of the
// MLP's hidden layer never reach memory.  Same machinery as the layer kernel above (3xTF32 mma.sync, one pipeline per
// warp over 16-row tiles); H = 2K hidden units are accumulated as H / 8 tiles in two halves of 8 to bound the registers.
template <int K> struct HeadShape {
    static constexpr int H = 2 * K;                    // hidden units of the MLP = width of [hidden | query]
    static constexpr int kPad = K + 4;
    static constexpr int kSteps = K / 8;
    static constexpr int kNTiles = H / 8;
    static constexpr size_t kWeightBytes = (size_t)kSteps * kNTiles * 32 * sizeof(float4);
    static constexpr size_t kTileBytes = (size_t)kWarpRows * kPad * sizeof(float);
    static constexpr size_t kSmemBytes = kWeightBytes + kLinearWarps * kTileBytes;
};

template <int K>
__global__ void __launch_bounds__(kLinearThreads, 1)
score_head_linear_kernel(const float *__restrict__ A, long long lda, const float *__restrict__ W, long long ldw,
                         const float *__restrict__ query_bias, const float *__restrict__ w2, const float *__restrict__ b2,
                         float *__restrict__ score, long long rows, int batch) {
    using Shape = HeadShape<K>;
    constexpr int H = Shape::H, kPad = Shape::kPad, kSteps = Shape::kSteps, kNTiles = Shape::kNTiles;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *w_frag = reinterpret_cast<float4 *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    float *mine = reinterpret_cast<float *>(smem_raw + Shape::kWeightBytes) + warp * (kWarpRows * kPad);

    // W1[:, :K] (H rows of ldw floats) -> fragment order [k-step][n-tile][lane] = (b0 hi, b1 hi, b0 lo, b1 lo)
    for (int idx = tid; idx < kSteps * kNTiles * 32; idx += kLinearThreads) {
        const int l = idx & 31, j = (idx >> 5) % kNTiles, s = idx / (32 * kNTiles);
        const int n = 8 * j + (l >> 2), k = 8 * s + (l & 3);
        const float b0 = __ldg(W + n * ldw + k), b1 = __ldg(W + n * ldw + k + 4);
        const float b0_hi = to_tf32(b0), b1_hi = to_tf32(b1);
        w_frag[idx] = make_float4(b0_hi, b1_hi, to_tf32(b0 - b0_hi), to_tf32(b1 - b1_hi));
    }
    __syncthreads();

    constexpr int kChunks = K / 4;
    const float bias = b2 ? __ldg(b2) : 0.f;
    const long long n_tiles = (rows + kWarpRows - 1) / kWarpRows;
    const long long stride = (long long)gridDim.x * kLinearWarps;
    for (long long tile = (long long)blockIdx.x * kLinearWarps + warp; tile < n_tiles; tile += stride) {
        const long long row0 = tile * kWarpRows;
#pragma unroll 4
        for (int c = lane; c < kWarpRows * kChunks; c += 32) {
            const int r = c / kChunks, q = c % kChunks;
            const bool live = row0 + r < rows;
            cp_async_16(mine + r * kPad + 4 * q, A + (live ? row0 + r : 0) * lda + 4 * q, live ? 16u : 0u);
        }
        cp_async_wait_all();
        __syncwarp();

        const float *row_upper = mine + g * kPad + t;
        const float *row_lower = row_upper + 8 * kPad;
        // the two rows of this thread and their queries (rows are (node, query) pairs, query fastest)
        const long long row_a = row0 + g, row_b = row0 + g + 8;
        const float *qb_a = query_bias + (row_a % batch) * H + 2 * t, *qb_b = query_bias + (row_b % batch) * H + 2 * t;
        float dot[2] = {0.f, 0.f};
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            constexpr int kHalf = kNTiles / 2;
            float acc[kHalf][4];
#pragma unroll
            for (int j = 0; j < kHalf; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll 2
            for (int s = 0; s < kSteps; ++s) {
                const float a[4] = {row_upper[8 * s], row_lower[8 * s], row_upper[8 * s + 4], row_lower[8 * s + 4]};
                float big[4], small[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    big[i] = to_tf32(a[i]);
                    small[i] = to_tf32(a[i] - big[i]);
                }
                const float4 *w = w_frag + (s * kNTiles + half * kHalf) * 32 + lane;
                float4 b[kHalf];
#pragma unroll
                for (int j = 0; j < kHalf; ++j) b[j] = w[j * 32];
#pragma unroll
                for (int j = 0; j < kHalf; ++j) mma_tf32(acc[j], small, b[j].x, b[j].y);
#pragma unroll
                for (int j = 0; j < kHalf; ++j) mma_tf32(acc[j], big, b[j].z, b[j].w);
#pragma unroll
                for (int j = 0; j < kHalf; ++j) mma_tf32(acc[j], big, b[j].x, b[j].y);
            }
            // hidden units {8j + 2t, 8j + 2t + 1} of rows g (acc[j][0..1]) and g + 8 (acc[j][2..3])
#pragma unroll
            for (int j = 0; j < kHalf; ++j) {
                const int col = 8 * (half * kHalf + j);
                const float2 w = __ldg(reinterpret_cast<const float2 *>(w2 + col + 2 * t));
                const float2 qa = __ldg(reinterpret_cast<const float2 *>(qb_a + col));
                const float2 qb = __ldg(reinterpret_cast<const float2 *>(qb_b + col));
                dot[0] = fmaf(fmaxf(acc[j][0] + qa.x, 0.f), w.x, dot[0]);
                dot[0] = fmaf(fmaxf(acc[j][1] + qa.y, 0.f), w.y, dot[0]);
                dot[1] = fmaf(fmaxf(acc[j][2] + qb.x, 0.f), w.x, dot[1]);
                dot[1] = fmaf(fmaxf(acc[j][3] + qb.y, 0.f), w.y, dot[1]);
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            dot[h] += __shfl_xor_sync(kFullMask, dot[h], 1);
            dot[h] += __shfl_xor_sync(kFullMask, dot[h], 2);
        }
        if (t == 0) {
            if (row_a < rows) score[row_a] = dot[0] + bias;
            if (row_b < rows) score[row_b] = dot[1] + bias;
        }
        __syncwarp();
    }
}

template <int K>
int launch_score_head(const float *A, long long lda, const float *W, long long ldw, const float *query_bias, const float *w2,
                      const float *b2, float *score, long long rows, int batch, cudaStream_t stream) {
    using Shape = HeadShape<K>;
    static int sm_count = 0;
    auto kernel = score_head_linear_kernel<K>;
    if (sm_count == 0) {                               // once per process (one process per GPU), outside any graph capture
        int device = 0, count = 0;
        ULTRA_CUDA_OK(cudaGetDevice(&device));
        ULTRA_CUDA_OK(cudaDeviceGetAttribute(&count, cudaDevAttrMultiProcessorCount, device));
        ULTRA_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Shape::kSmemBytes));
        sm_count = count;
    }
    const long long n_blocks = (rows + kWarpRows * kLinearWarps - 1) / (kWarpRows * kLinearWarps);
    const unsigned grid = (unsigned)(n_blocks < sm_count ? n_blocks : sm_count);
    kernel<<<grid, kLinearThreads, Shape::kSmemBytes, stream>>>(A, lda, W, ldw, query_bias, w2, b2, score, rows, batch);
    note_launch();
    ULTRA_CUDA_OK(cudaGetLastError());
    return ULTRA_RSPMM_OK;
}

template <int N>
int launch_linear(const float *A, long long lda, const float *W, const float *linear_bias, const float *gamma,
                  const float *beta, float *out, long long ldo, long long rows, float eps, int relu, int shortcut,
                  cudaStream_t stream) {
    using Shape = LinearShape<N>;
    static int sm_count = 0;
    if (sm_count == 0) {
        int device = 0;
        ULTRA_CUDA_OK(cudaGetDevice(&device));
        ULTRA_CUDA_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
    }
    auto kernel = linear_norm_relu_residual_kernel<N>;
    static bool configured = false;                    // once per process (one process per GPU), outside any graph capture
    if (!configured) {
        ULTRA_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Shape::kSmemBytes));
        configured = true;
    }
    const long long n_blocks = (rows + kWarpRows * kLinearWarps - 1) / (kWarpRows * kLinearWarps);
    const unsigned grid = (unsigned)(n_blocks < sm_count ? n_blocks : sm_count);
    kernel<<<grid, kLinearThreads, Shape::kSmemBytes, stream>>>(A, lda, W, linear_bias, gamma, beta, out, ldo, rows, eps,
                                                                 relu, shortcut);
    note_launch();
    ULTRA_CUDA_OK(cudaGetLastError());
    return ULTRA_RSPMM_OK;
}

}  // namespace

}  // namespace ultra

namespace ultra {
// layer_linear_tc.cu
int layer_linear_tc(const float *A, long long lda, const float *A1, long long lda1, const float *W, const float *linear_bias, const float *gamma,
                    const float *beta, float *out, long long ldo, long long rows, int out_dim, float eps, int relu,
                    int shortcut, cudaStream_t stream, float *pre_out = nullptr, long long ld_pre = 0);
int g_linear_kernel = 0;   // 0 = default, 1 = mma.sync, 2 = tcgen05
}  // namespace ultra

using namespace ultra;

extern "C" int ultra_layer_linear_set_kernel(int32_t kind) {
    if (kind < 0 || kind > 2) return ULTRA_RSPMM_ERR_ARG;
    g_linear_kernel = kind;
    return ULTRA_RSPMM_OK;
}

extern "C" int ultra_layer_linear_get_kernel(void) { return g_linear_kernel ? g_linear_kernel : kDefaultLinearKernel; }

extern "C" int ultra_layer_linear_norm_relu_residual(const float *dev_input, int64_t input_stride, const float *dev_weight,
                                                     const float *dev_linear_bias, const float *dev_gamma,
                                                     const float *dev_beta, float *dev_out, int64_t out_stride,
                                                     int64_t rows, int32_t out_dim, float eps, int32_t relu,
                                                     int32_t shortcut, void *stream) {
    if (rows < 0 || (rows > 0 && (!dev_input || !dev_weight || !dev_out))) return ULTRA_RSPMM_ERR_ARG;
    if ((dev_gamma == nullptr) != (dev_beta == nullptr)) return ULTRA_RSPMM_ERR_ARG;
    if (out_dim != 32 && out_dim != 64) return ULTRA_RSPMM_ERR_RANGE;
    if (input_stride < 2 * out_dim || input_stride % 4 || out_stride < out_dim || out_stride % 2) return ULTRA_RSPMM_ERR_ARG;
    if (((uintptr_t)dev_input & 15) || (((uintptr_t)dev_out | (uintptr_t)dev_linear_bias | (uintptr_t)dev_gamma |
                                         (uintptr_t)dev_beta) & 7))
        return ULTRA_RSPMM_ERR_ARG;
    if (rows == 0) return ULTRA_RSPMM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const bool wide_aligned = !(((uintptr_t)dev_out | (uintptr_t)dev_linear_bias | (uintptr_t)dev_gamma | (uintptr_t)dev_beta) & 15) &&
                              out_stride % 4 == 0;
    const int kind = g_linear_kernel ? g_linear_kernel : kDefaultLinearKernel;
    if (kind == 2 && wide_aligned)
        return layer_linear_tc(dev_input, input_stride, nullptr, 0, dev_weight, dev_linear_bias, dev_gamma, dev_beta, dev_out, out_stride,
                               rows, out_dim, eps, relu, shortcut, s);
    if (out_dim == 64)
        return launch_linear<64>(dev_input, input_stride, dev_weight, dev_linear_bias, dev_gamma, dev_beta, dev_out,
                                 out_stride, rows, eps, relu, shortcut, s);
    return launch_linear<32>(dev_input, input_stride, dev_weight, dev_linear_bias, dev_gamma, dev_beta, dev_out, out_stride,
                             rows, eps, relu, shortcut, s);
}

extern "C" int ultra_layer_linear_norm_relu_residual_two(const float *dev_input, int64_t input_stride, const float *dev_update,
                                                         int64_t update_stride, const float *dev_weight,
                                                         const float *dev_linear_bias, const float *dev_gamma,
                                                         const float *dev_beta, float *dev_out, int64_t out_stride, int64_t rows,
                                                         int32_t out_dim, float eps, int32_t relu, int32_t shortcut, void *stream) {
    if (rows < 0 || (rows > 0 && (!dev_input || !dev_update || !dev_weight || !dev_out))) return ULTRA_RSPMM_ERR_ARG;
    if ((dev_gamma == nullptr) != (dev_beta == nullptr)) return ULTRA_RSPMM_ERR_ARG;
    if (out_dim != 32 && out_dim != 64) return ULTRA_RSPMM_ERR_RANGE;
    if (input_stride < out_dim || input_stride % 4 || update_stride < out_dim || update_stride % 4 || out_stride < out_dim ||
        out_stride % 4)
        return ULTRA_RSPMM_ERR_ARG;
    if (((uintptr_t)dev_input | (uintptr_t)dev_update | (uintptr_t)dev_out | (uintptr_t)dev_linear_bias | (uintptr_t)dev_gamma |
         (uintptr_t)dev_beta | (uintptr_t)dev_weight) & 15)
        return ULTRA_RSPMM_ERR_ARG;
    if (rows == 0) return ULTRA_RSPMM_OK;
    return layer_linear_tc(dev_input, input_stride, dev_update, update_stride, dev_weight, dev_linear_bias, dev_gamma, dev_beta,
                           dev_out, out_stride, rows, out_dim, eps, relu, shortcut, (cudaStream_t)stream);
}

extern "C" int ultra_layer_linear_norm_relu_residual_two_pre(const float *dev_input, int64_t input_stride, const float *dev_update,
                                                             int64_t update_stride, const float *dev_weight,
                                                             const float *dev_linear_bias, const float *dev_gamma,
                                                             const float *dev_beta, float *dev_out, int64_t out_stride,
                                                             float *dev_pre_out, int64_t pre_stride, int64_t rows, int32_t out_dim,
                                                             float eps, int32_t relu, int32_t shortcut, void *stream) {
    if (!dev_pre_out || pre_stride < out_dim || pre_stride % 4 || ((uintptr_t)dev_pre_out & 15)) return ULTRA_RSPMM_ERR_ARG;
    if (rows < 0 || (rows > 0 && (!dev_input || !dev_update || !dev_weight || !dev_out))) return ULTRA_RSPMM_ERR_ARG;
    if ((dev_gamma == nullptr) != (dev_beta == nullptr)) return ULTRA_RSPMM_ERR_ARG;
    if (out_dim != 32 && out_dim != 64) return ULTRA_RSPMM_ERR_RANGE;
    if (input_stride < out_dim || input_stride % 4 || update_stride < out_dim || update_stride % 4 || out_stride < out_dim ||
        out_stride % 4)
        return ULTRA_RSPMM_ERR_ARG;
    if (((uintptr_t)dev_input | (uintptr_t)dev_update | (uintptr_t)dev_out | (uintptr_t)dev_linear_bias | (uintptr_t)dev_gamma |
         (uintptr_t)dev_beta | (uintptr_t)dev_weight) & 15)
        return ULTRA_RSPMM_ERR_ARG;
    if (rows == 0) return ULTRA_RSPMM_OK;
    return layer_linear_tc(dev_input, input_stride, dev_update, update_stride, dev_weight, dev_linear_bias, dev_gamma, dev_beta,
                           dev_out, out_stride, rows, out_dim, eps, relu, shortcut, (cudaStream_t)stream, dev_pre_out, pre_stride);
}

extern "C" int ultra_score_head_linear(const float *dev_input, int64_t input_stride, const float *dev_weight,
                                       int64_t weight_stride, const float *dev_query_bias, const float *dev_out_weight,
                                       const float *dev_out_bias, float *dev_score, int64_t rows, int32_t batch,
                                       int32_t in_dim, void *stream) {
    if (rows < 0 || batch <= 0 || (rows > 0 && (!dev_input || !dev_weight || !dev_query_bias || !dev_out_weight || !dev_score)))
        return ULTRA_RSPMM_ERR_ARG;
    if (in_dim != 32 && in_dim != 64) return ULTRA_RSPMM_ERR_RANGE;
    if (input_stride < in_dim || input_stride % 4 || weight_stride < in_dim) return ULTRA_RSPMM_ERR_ARG;
    if (((uintptr_t)dev_input & 15) || (((uintptr_t)dev_query_bias | (uintptr_t)dev_out_weight) & 7)) return ULTRA_RSPMM_ERR_ARG;
    if (rows == 0) return ULTRA_RSPMM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (in_dim == 64)
        return launch_score_head<64>(dev_input, input_stride, dev_weight, weight_stride, dev_query_bias, dev_out_weight,
                                     dev_out_bias, dev_score, rows, batch, s);
    return launch_score_head<32>(dev_input, input_stride, dev_weight, weight_stride, dev_query_bias, dev_out_weight,
                                 dev_out_bias, dev_score, rows, batch, s);
}
