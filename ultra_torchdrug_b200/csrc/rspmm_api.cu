// Diagnostics, tuning knobs and the host-buffer convenience layer of the C ABI (include/ultra_rspmm.h).
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <new>

#include "rspmm_common.cuh"

namespace ultra {

static std::atomic<long long> g_launches{0};
static thread_local int t_last_cuda_error = 0;
int g_chunk = 256;
int g_variant = 0;

void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int fail_cuda(cudaError_t error) {
    t_last_cuda_error = (int)error;
    return ULTRA_RSPMM_ERR_CUDA;
}

}  // namespace ultra

using namespace ultra;

extern "C" int ultra_rspmm_abi_version(void) { return ULTRA_RSPMM_ABI_VERSION; }
extern "C" int ultra_rspmm_last_cuda_error(void) { return t_last_cuda_error; }
extern "C" int64_t ultra_rspmm_launch_count(void) { return g_launches.load(); }
extern "C" void ultra_rspmm_launch_count_reset(void) { g_launches.store(0); }

extern "C" const char *ultra_rspmm_status_string(int status) {
    switch (status) {
        case ULTRA_RSPMM_OK: return "ok";
        case ULTRA_RSPMM_ERR_ARG: return "invalid argument (null pointer, negative size, unknown op or dtype code)";
        case ULTRA_RSPMM_ERR_WORKSPACE: return "caller-provided buffer is smaller than the size query answered";
        case ULTRA_RSPMM_ERR_CUDA: return "a CUDA runtime call failed (see ultra_rspmm_last_cuda_error)";
        case ULTRA_RSPMM_ERR_INDEX: return "a sparse index is out of range for the operand shape";
        case ULTRA_RSPMM_ERR_DTYPE: return "operand dtype differs from the dtype the graph index was built for";
        case ULTRA_RSPMM_ERR_RANGE: return "operand shape exceeds the 63-bit sort key / int32 edge-id range";
        default: return "unknown status";
    }
}

extern "C" int ultra_rspmm_set_tuning(int32_t chunk, int32_t variant) {
    if (chunk < 0 || chunk > (1 << 20) || variant < 0 || variant > 2) return ULTRA_RSPMM_ERR_ARG;
    if (chunk > 0) g_chunk = chunk;
    g_variant = variant;
    return ULTRA_RSPMM_OK;
}

extern "C" int ultra_rspmm_host_alloc(void **ptr, size_t bytes) {
    if (!ptr) return ULTRA_RSPMM_ERR_ARG;
    ULTRA_CUDA_OK(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return ULTRA_RSPMM_OK;
}

extern "C" int ultra_rspmm_host_free(void *ptr) {
    if (ptr) ULTRA_CUDA_OK(cudaFreeHost(ptr));
    return ULTRA_RSPMM_OK;
}

// ---- host-buffer context ----------------------------------------------------------------------
struct ultra_rspmm_ctx {
    int device;
    cudaStream_t stream;
    cudaEvent_t start, stop;
    float last_ms;
    bool has_graph;
    ultra_rspmm_index_t index;
    void *index_buffer;
    // grow-only device buffers
    void *buf[8];
    size_t cap[8];
};

enum { BUF_REL = 0, BUF_IN, BUF_OUT, BUF_GOUT, BUF_GREL, BUF_GIN, BUF_WS, BUF_TMP };

static int ctx_reserve(ultra_rspmm_ctx *ctx, int which, size_t bytes) {
    if (bytes <= ctx->cap[which]) return ULTRA_RSPMM_OK;
    if (ctx->buf[which]) ULTRA_CUDA_OK(cudaFree(ctx->buf[which]));
    ctx->buf[which] = nullptr;
    ctx->cap[which] = 0;
    ULTRA_CUDA_OK(cudaMalloc(&ctx->buf[which], bytes));
    ctx->cap[which] = bytes;
    return ULTRA_RSPMM_OK;
}

extern "C" int ultra_rspmm_ctx_create(ultra_rspmm_ctx_t **out, int32_t device) {
    if (!out || device < 0) return ULTRA_RSPMM_ERR_ARG;
    ULTRA_CUDA_OK(cudaSetDevice(device));
    ultra_rspmm_ctx *ctx = new (std::nothrow) ultra_rspmm_ctx();
    if (!ctx) return ULTRA_RSPMM_ERR_ARG;
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device;
    ULTRA_CUDA_OK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ULTRA_CUDA_OK(cudaEventCreate(&ctx->start));
    ULTRA_CUDA_OK(cudaEventCreate(&ctx->stop));
    *out = ctx;
    return ULTRA_RSPMM_OK;
}

extern "C" int ultra_rspmm_ctx_destroy(ultra_rspmm_ctx_t *ctx) {
    if (!ctx) return ULTRA_RSPMM_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 8; ++i)
        if (ctx->buf[i]) cudaFree(ctx->buf[i]);
    if (ctx->index_buffer) cudaFree(ctx->index_buffer);
    cudaEventDestroy(ctx->start);
    cudaEventDestroy(ctx->stop);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return ULTRA_RSPMM_OK;
}

extern "C" int ultra_rspmm_ctx_set_graph(ultra_rspmm_ctx_t *ctx, const int64_t *host_indices, const void *host_values,
                                         int64_t nnz_raw, int32_t n_out, int32_t n_in, int32_t n_rel, int32_t dtype) {
    if (!ctx || nnz_raw < 0 || (nnz_raw > 0 && (!host_indices || !host_values))) return ULTRA_RSPMM_ERR_ARG;
    ULTRA_CUDA_OK(cudaSetDevice(ctx->device));
    size_t index_bytes = 0, scratch_bytes = 0;
    int status = ultra_rspmm_index_bytes(nnz_raw, n_out, n_in, n_rel, dtype, &index_bytes, &scratch_bytes);
    if (status) return status;
    const size_t elem = dtype == ULTRA_RSPMM_F32 ? 4 : 8;
    const size_t idx_bytes = align_up((size_t)nnz_raw * 3 * sizeof(int64_t));
    const size_t val_bytes = align_up((size_t)nnz_raw * elem);
    status = ctx_reserve(ctx, BUF_TMP, idx_bytes + val_bytes + scratch_bytes + 256);
    if (status) return status;
    ctx->has_graph = false;
    if (ctx->index_buffer) ULTRA_CUDA_OK(cudaFree(ctx->index_buffer));
    ctx->index_buffer = nullptr;
    ULTRA_CUDA_OK(cudaMalloc(&ctx->index_buffer, index_bytes ? index_bytes : 256));
    char *tmp = (char *)ctx->buf[BUF_TMP];
    if (nnz_raw > 0) {
        ULTRA_CUDA_OK(cudaMemcpyAsync(tmp, host_indices, (size_t)nnz_raw * 3 * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
        ULTRA_CUDA_OK(cudaMemcpyAsync(tmp + idx_bytes, host_values, (size_t)nnz_raw * elem, cudaMemcpyHostToDevice, ctx->stream));
    }
    status = ultra_rspmm_index_build((const int64_t *)tmp, nnz_raw, tmp + idx_bytes, nnz_raw, n_out, n_in, n_rel, dtype,
                                     ctx->index_buffer, index_bytes, tmp + idx_bytes + val_bytes, scratch_bytes,
                                     &ctx->index, ctx->stream);
    if (status) return status;
    ctx->has_graph = true;
    return ULTRA_RSPMM_OK;
}

static int ctx_run(ultra_rspmm_ctx_t *ctx, const void *host_relation, const void *host_input,
                   const void *host_grad_output, void *host_output, void *host_grad_relation, void *host_grad_input,
                   int64_t dim, int32_t sum_op, int32_t mul_op, bool with_backward) {
    if (!ctx || !ctx->has_graph || dim < 0) return ULTRA_RSPMM_ERR_ARG;
    if (!host_relation || !host_input || !host_output) return ULTRA_RSPMM_ERR_ARG;
    if (with_backward && (!host_grad_output || !host_grad_relation || !host_grad_input)) return ULTRA_RSPMM_ERR_ARG;
    ULTRA_CUDA_OK(cudaSetDevice(ctx->device));
    const ultra_rspmm_index_t &ix = ctx->index;
    const size_t elem = ix.dtype == ULTRA_RSPMM_F32 ? 4 : 8;
    const size_t rel_bytes = (size_t)ix.n_rel * dim * elem, in_bytes = (size_t)ix.n_in * dim * elem,
                 out_bytes = (size_t)ix.n_out * dim * elem;
    size_t fwd_ws = 0, bwd_ws = 0;
    int status = ultra_rspmm_workspace_bytes(&ix, dim, ix.dtype, &fwd_ws, &bwd_ws);
    if (status) return status;
    const size_t ws = with_backward && bwd_ws > fwd_ws ? bwd_ws : fwd_ws;
    if ((status = ctx_reserve(ctx, BUF_REL, rel_bytes + 256))) return status;
    if ((status = ctx_reserve(ctx, BUF_IN, in_bytes + 256))) return status;
    if ((status = ctx_reserve(ctx, BUF_OUT, out_bytes + 256))) return status;
    if ((status = ctx_reserve(ctx, BUF_WS, ws + 256))) return status;
    if (with_backward) {
        if ((status = ctx_reserve(ctx, BUF_GOUT, out_bytes + 256))) return status;
        if ((status = ctx_reserve(ctx, BUF_GREL, rel_bytes + 256))) return status;
        if ((status = ctx_reserve(ctx, BUF_GIN, in_bytes + 256))) return status;
    }
    cudaStream_t s = ctx->stream;
    ULTRA_CUDA_OK(cudaMemcpyAsync(ctx->buf[BUF_REL], host_relation, rel_bytes, cudaMemcpyHostToDevice, s));
    ULTRA_CUDA_OK(cudaMemcpyAsync(ctx->buf[BUF_IN], host_input, in_bytes, cudaMemcpyHostToDevice, s));
    if (with_backward)
        ULTRA_CUDA_OK(cudaMemcpyAsync(ctx->buf[BUF_GOUT], host_grad_output, out_bytes, cudaMemcpyHostToDevice, s));
    ULTRA_CUDA_OK(cudaEventRecord(ctx->start, s));
    status = ultra_rspmm_forward(&ix, ctx->buf[BUF_REL], ctx->buf[BUF_IN], ctx->buf[BUF_OUT], nullptr, dim, ix.dtype,
                                 sum_op, mul_op, ctx->buf[BUF_WS], ctx->cap[BUF_WS], s);
    if (status) return status;
    if (with_backward) {
        status = ultra_rspmm_backward(&ix, ctx->buf[BUF_REL], ctx->buf[BUF_IN], ctx->buf[BUF_OUT], ctx->buf[BUF_GOUT],
                                      ctx->buf[BUF_GREL], ctx->buf[BUF_GIN], dim, ix.dtype, sum_op, mul_op,
                                      ctx->buf[BUF_WS], ctx->cap[BUF_WS], s);
        if (status) return status;
    }
    ULTRA_CUDA_OK(cudaEventRecord(ctx->stop, s));
    ULTRA_CUDA_OK(cudaMemcpyAsync(host_output, ctx->buf[BUF_OUT], out_bytes, cudaMemcpyDeviceToHost, s));
    if (with_backward) {
        ULTRA_CUDA_OK(cudaMemcpyAsync(host_grad_relation, ctx->buf[BUF_GREL], rel_bytes, cudaMemcpyDeviceToHost, s));
        ULTRA_CUDA_OK(cudaMemcpyAsync(host_grad_input, ctx->buf[BUF_GIN], in_bytes, cudaMemcpyDeviceToHost, s));
    }
    ULTRA_CUDA_OK(cudaStreamSynchronize(s));
    ULTRA_CUDA_OK(cudaEventElapsedTime(&ctx->last_ms, ctx->start, ctx->stop));
    return ULTRA_RSPMM_OK;
}

extern "C" int ultra_rspmm_ctx_forward(ultra_rspmm_ctx_t *ctx, const void *host_relation, const void *host_input,
                                       void *host_output, int64_t dim, int32_t sum_op, int32_t mul_op) {
    return ctx_run(ctx, host_relation, host_input, nullptr, host_output, nullptr, nullptr, dim, sum_op, mul_op, false);
}

extern "C" int ultra_rspmm_ctx_forward_backward(ultra_rspmm_ctx_t *ctx, const void *host_relation,
                                                const void *host_input, const void *host_grad_output,
                                                void *host_output, void *host_grad_relation, void *host_grad_input,
                                                int64_t dim, int32_t sum_op, int32_t mul_op) {
    return ctx_run(ctx, host_relation, host_input, host_grad_output, host_output, host_grad_relation, host_grad_input,
                   dim, sum_op, mul_op, true);
}

extern "C" float ultra_rspmm_ctx_last_kernel_ms(const ultra_rspmm_ctx_t *ctx) { return ctx ? ctx->last_ms : -1.0f; }
extern "C" int64_t ultra_rspmm_ctx_nnz(const ultra_rspmm_ctx_t *ctx) {
    return ctx && ctx->has_graph ? ctx->index.nnz : -1;
}
