// Diagnostics, tuning knobs and the host-buffer convenience layer of the C ABI (include/ultra_rspmm.h).
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "rspmm_common.cuh"

namespace ultra {

static std::atomic<long long> g_launches{0};
static thread_local int t_last_cuda_error = 0;
int g_chunk = 256;
int g_variant = 0;
// off by default: measured slower than the 512-byte generic kernel on every shape tried (DESIGN.md "Tried and rejected")
long long g_narrow_bytes = (getenv("ULTRA_RSPMM_NARROW_MB") ? atoll(getenv("ULTRA_RSPMM_NARROW_MB")) : 0) << 20;
int g_narrow_sub = getenv("ULTRA_RSPMM_NARROW_SUB") ? atoi(getenv("ULTRA_RSPMM_NARROW_SUB")) : 0;
int g_staged = getenv("ULTRA_RSPMM_STAGED") ? atoi(getenv("ULTRA_RSPMM_STAGED")) : 1;
int g_group_edges = getenv("ULTRA_RSPMM_GROUP") ? atoi(getenv("ULTRA_RSPMM_GROUP")) : -1;
long long g_l2_budget = 1ll << 40;   // slab narrowing off by default: it lost on every measured shape (profiles/)

static ultra_rspmm_pass_info_t g_pass_info[3] = {};

void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void note_pass(int pass, const ultra_rspmm_pass_info_t &info) {
    if (pass >= 0 && pass < 3) g_pass_info[pass] = info;
}

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

extern "C" int ultra_rspmm_last_pass_info(int32_t pass, ultra_rspmm_pass_info_t *info) {
    if (pass < 0 || pass > 2 || !info) return ULTRA_RSPMM_ERR_ARG;
    *info = g_pass_info[pass];
    return ULTRA_RSPMM_OK;
}

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

extern "C" int ultra_rspmm_set_tuning(int32_t chunk, int32_t variant, int64_t l2_budget_bytes) {
    if (chunk < 0 || chunk > (1 << 20) || variant < 0 || variant > 2 || l2_budget_bytes < 0) return ULTRA_RSPMM_ERR_ARG;
    if (chunk > 0) g_chunk = chunk;
    g_variant = variant;
    if (l2_budget_bytes > 0) g_l2_budget = l2_budget_bytes;
    return ULTRA_RSPMM_OK;
}

extern "C" int ultra_rspmm_set_staged(int32_t mode) {
    if (mode < 0 || mode > 2) return ULTRA_RSPMM_ERR_ARG;
    g_staged = mode;
    return ULTRA_RSPMM_OK;
}

extern "C" int ultra_rspmm_set_narrow(int64_t slab_bytes, int32_t sub) {
    if (slab_bytes < 0 || (sub != 0 && sub != 2 && sub != 4)) return ULTRA_RSPMM_ERR_ARG;
    g_narrow_bytes = slab_bytes;
    g_narrow_sub = sub;
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
// The feature axis is the query batch (feature = b * 64 + c), so a call on host operands is pipelined
// over column chunks (query groups): chunk c+1 is uploaded (2-D copies, pinned host memory -> compact
// device buffers) while chunk c is reduced and chunk c-1 is downloaded.  PCIe is full duplex, so the
// call costs about max(H2D, D2H) instead of H2D + kernels + D2H.
namespace {

constexpr int kSets = 3;            // buffer sets in flight: upload / compute / download
constexpr int64_t kChunkCols = 512; // 8 queries of 64 features; a multiple of the 128-feature slab

enum { BUF_REL = 0, BUF_IN, BUF_OUT, BUF_GOUT, BUF_GREL, BUF_GIN, BUF_WS, BUF_KINDS };

}  // namespace

struct ultra_rspmm_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, stream_in = nullptr, stream_out = nullptr;
    float last_ms = 0.f;
    bool has_graph = false;
    ultra_rspmm_index_t index;
    void *index_buffer = nullptr;
    void *extension = nullptr;
    void *tmp = nullptr;
    size_t tmp_cap = 0;
    void *buf[kSets][BUF_KINDS] = {};
    size_t cap[kSets][BUF_KINDS] = {};
    cudaEvent_t uploaded[kSets] = {}, computed[kSets] = {}, downloaded[kSets] = {};
    std::vector<cudaEvent_t> tick;   // per-chunk kernel timing: 2 events per chunk
};

static int reserve(void **ptr, size_t *cap, size_t bytes) {
    if (bytes <= *cap) return ULTRA_RSPMM_OK;
    if (*ptr) ULTRA_CUDA_OK(cudaFree(*ptr));
    *ptr = nullptr;
    *cap = 0;
    ULTRA_CUDA_OK(cudaMalloc(ptr, bytes));
    *cap = bytes;
    return ULTRA_RSPMM_OK;
}

extern "C" int ultra_rspmm_ctx_create(ultra_rspmm_ctx_t **out, int32_t device) {
    if (!out || device < 0) return ULTRA_RSPMM_ERR_ARG;
    ULTRA_CUDA_OK(cudaSetDevice(device));
    ultra_rspmm_ctx *ctx = new (std::nothrow) ultra_rspmm_ctx();
    if (!ctx) return ULTRA_RSPMM_ERR_ARG;
    memset(&ctx->index, 0, sizeof(ctx->index));
    ctx->device = device;
    ULTRA_CUDA_OK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ULTRA_CUDA_OK(cudaStreamCreateWithFlags(&ctx->stream_in, cudaStreamNonBlocking));
    ULTRA_CUDA_OK(cudaStreamCreateWithFlags(&ctx->stream_out, cudaStreamNonBlocking));
    for (int s = 0; s < kSets; ++s) {
        ULTRA_CUDA_OK(cudaEventCreateWithFlags(&ctx->uploaded[s], cudaEventDisableTiming));
        ULTRA_CUDA_OK(cudaEventCreateWithFlags(&ctx->computed[s], cudaEventDisableTiming));
        ULTRA_CUDA_OK(cudaEventCreateWithFlags(&ctx->downloaded[s], cudaEventDisableTiming));
    }
    *out = ctx;
    return ULTRA_RSPMM_OK;
}

extern "C" int ultra_rspmm_ctx_destroy(ultra_rspmm_ctx_t *ctx) {
    if (!ctx) return ULTRA_RSPMM_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int s = 0; s < kSets; ++s) {
        for (int k = 0; k < BUF_KINDS; ++k)
            if (ctx->buf[s][k]) cudaFree(ctx->buf[s][k]);
        cudaEventDestroy(ctx->uploaded[s]);
        cudaEventDestroy(ctx->computed[s]);
        cudaEventDestroy(ctx->downloaded[s]);
    }
    for (cudaEvent_t e : ctx->tick) cudaEventDestroy(e);
    if (ctx->index_buffer) cudaFree(ctx->index_buffer);
    if (ctx->extension) cudaFree(ctx->extension);
    if (ctx->tmp) cudaFree(ctx->tmp);
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->stream_in);
    cudaStreamDestroy(ctx->stream_out);
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
    status = reserve(&ctx->tmp, &ctx->tmp_cap, idx_bytes + val_bytes + scratch_bytes + 256);
    if (status) return status;
    ctx->has_graph = false;
    if (ctx->index_buffer) ULTRA_CUDA_OK(cudaFree(ctx->index_buffer));
    ctx->index_buffer = nullptr;
    ULTRA_CUDA_OK(cudaMalloc(&ctx->index_buffer, index_bytes ? index_bytes : 256));
    char *tmp = (char *)ctx->tmp;
    if (nnz_raw > 0) {
        ULTRA_CUDA_OK(cudaMemcpyAsync(tmp, host_indices, (size_t)nnz_raw * 3 * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
        ULTRA_CUDA_OK(cudaMemcpyAsync(tmp + idx_bytes, host_values, (size_t)nnz_raw * elem, cudaMemcpyHostToDevice, ctx->stream));
    }
    status = ultra_rspmm_index_build((const int64_t *)tmp, nnz_raw, tmp + idx_bytes, nnz_raw, n_out, n_in, n_rel, dtype,
                                     ctx->index_buffer, index_bytes, tmp + idx_bytes + val_bytes, scratch_bytes,
                                     &ctx->index, ctx->stream);
    if (status) return status;
    size_t extend_bytes = 0;
    if ((status = ultra_rspmm_index_extend_bytes(&ctx->index, &extend_bytes))) return status;
    if (ctx->extension) ULTRA_CUDA_OK(cudaFree(ctx->extension));
    ctx->extension = nullptr;
    if (extend_bytes) {
        ULTRA_CUDA_OK(cudaMalloc(&ctx->extension, extend_bytes));
        if ((status = ultra_rspmm_index_extend(&ctx->index, ctx->extension, extend_bytes, ctx->stream))) return status;
    }
    ctx->has_graph = true;
    return ULTRA_RSPMM_OK;
}

static int ctx_run_pipeline(ultra_rspmm_ctx_t *ctx, const void *host_relation, const void *host_input,
                            const void *host_grad_output, void *host_output, void *host_grad_relation, void *host_grad_input,
                            int64_t dim, int32_t sum_op, int32_t mul_op, bool with_backward);

// Any failure in the middle of the pipeline leaves asynchronous copies in flight on the caller's host buffers: drain
// the three streams (best effort) before the status is returned, so the caller may free or reuse them right away.
static int ctx_run(ultra_rspmm_ctx_t *ctx, const void *host_relation, const void *host_input,
                   const void *host_grad_output, void *host_output, void *host_grad_relation, void *host_grad_input,
                   int64_t dim, int32_t sum_op, int32_t mul_op, bool with_backward) {
    const int status = ctx_run_pipeline(ctx, host_relation, host_input, host_grad_output, host_output, host_grad_relation,
                                        host_grad_input, dim, sum_op, mul_op, with_backward);
    if (status != ULTRA_RSPMM_OK && ctx) {
        const int kept = t_last_cuda_error;
        cudaStreamSynchronize(ctx->stream_in);
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->stream_out);
        cudaGetLastError();
        t_last_cuda_error = kept;
    }
    return status;
}

static int ctx_run_pipeline(ultra_rspmm_ctx_t *ctx, const void *host_relation, const void *host_input,
                   const void *host_grad_output, void *host_output, void *host_grad_relation, void *host_grad_input,
                   int64_t dim, int32_t sum_op, int32_t mul_op, bool with_backward) {
    if (!ctx || !ctx->has_graph || dim < 0) return ULTRA_RSPMM_ERR_ARG;
    if (!host_relation || !host_input || !host_output) return ULTRA_RSPMM_ERR_ARG;
    if (with_backward && (!host_grad_output || !host_grad_relation || !host_grad_input)) return ULTRA_RSPMM_ERR_ARG;
    ULTRA_CUDA_OK(cudaSetDevice(ctx->device));
    const ultra_rspmm_index_t &ix = ctx->index;
    const size_t elem = ix.dtype == ULTRA_RSPMM_F32 ? 4 : 8;
    int64_t chunk_target = kChunkCols;
    if (const char *env = getenv("ULTRA_RSPMM_CHUNK_COLS")) chunk_target = atoll(env) > 0 ? atoll(env) : kChunkCols;
    const int64_t chunk_cols = dim < chunk_target ? dim : chunk_target;
    ctx->last_ms = 0.f;
    if (dim == 0) return ULTRA_RSPMM_OK;
    // chunk boundaries.  The upload of the first chunk and the download of the last one overlap with nothing, so with
    // four or more chunks those two are half as wide (a multiple of the 128-feature slab)
    std::vector<int64_t> bounds;
    bounds.push_back(0);
    {
        const int64_t full = (dim + chunk_cols - 1) / chunk_cols;
        const int64_t edge = full >= 4 && chunk_cols % 256 == 0 ? chunk_cols / 2 : chunk_cols;
        int64_t at = edge < dim ? edge : dim;
        bounds.push_back(at);
        while (at < dim) {
            int64_t next = at + chunk_cols;
            if (dim - at <= chunk_cols + edge && dim - at > edge && edge != chunk_cols) next = dim - edge;   // leave a half chunk
            if (next > dim) next = dim;
            bounds.push_back(next);
            at = next;
        }
    }
    const int n_chunk = (int)bounds.size() - 1;

    size_t fwd_ws = 0, bwd_ws = 0;
    int status = ultra_rspmm_workspace_bytes(&ix, chunk_cols, ix.dtype, &fwd_ws, &bwd_ws);
    if (status) return status;
    const size_t ws = with_backward && bwd_ws > fwd_ws ? bwd_ws : fwd_ws;
    const size_t rel_bytes = (size_t)ix.n_rel * chunk_cols * elem, in_bytes = (size_t)ix.n_in * chunk_cols * elem,
                 out_bytes = (size_t)ix.n_out * chunk_cols * elem;
    const int sets = n_chunk < kSets ? n_chunk : kSets;
    for (int s = 0; s < sets; ++s) {
        if ((status = reserve(&ctx->buf[s][BUF_REL], &ctx->cap[s][BUF_REL], rel_bytes + 256))) return status;
        if ((status = reserve(&ctx->buf[s][BUF_IN], &ctx->cap[s][BUF_IN], in_bytes + 256))) return status;
        if ((status = reserve(&ctx->buf[s][BUF_OUT], &ctx->cap[s][BUF_OUT], out_bytes + 256))) return status;
        if ((status = reserve(&ctx->buf[s][BUF_WS], &ctx->cap[s][BUF_WS], ws + 256))) return status;
        if (with_backward) {
            if ((status = reserve(&ctx->buf[s][BUF_GOUT], &ctx->cap[s][BUF_GOUT], out_bytes + 256))) return status;
            if ((status = reserve(&ctx->buf[s][BUF_GREL], &ctx->cap[s][BUF_GREL], rel_bytes + 256))) return status;
            if ((status = reserve(&ctx->buf[s][BUF_GIN], &ctx->cap[s][BUF_GIN], in_bytes + 256))) return status;
        }
    }
    while ((int)ctx->tick.size() < 2 * n_chunk) {
        cudaEvent_t e;
        ULTRA_CUDA_OK(cudaEventCreate(&e));
        ctx->tick.push_back(e);
    }
    const size_t host_pitch = (size_t)dim * elem;
    for (int c = 0; c < n_chunk; ++c) {
        const int s = c % kSets;
        const int64_t col0 = bounds[c];
        const int64_t cols = bounds[c + 1] - col0;
        const size_t width = (size_t)cols * elem, offset = (size_t)col0 * elem;
        void **b = ctx->buf[s];
        // upload (the set's previous chunk must have been downloaded)
        if (c >= kSets) ULTRA_CUDA_OK(cudaStreamWaitEvent(ctx->stream_in, ctx->downloaded[s], 0));
        ULTRA_CUDA_OK(cudaMemcpy2DAsync(b[BUF_REL], width, (const char *)host_relation + offset, host_pitch, width,
                                        ix.n_rel, cudaMemcpyHostToDevice, ctx->stream_in));
        ULTRA_CUDA_OK(cudaMemcpy2DAsync(b[BUF_IN], width, (const char *)host_input + offset, host_pitch, width, ix.n_in,
                                        cudaMemcpyHostToDevice, ctx->stream_in));
        if (with_backward)
            ULTRA_CUDA_OK(cudaMemcpy2DAsync(b[BUF_GOUT], width, (const char *)host_grad_output + offset, host_pitch, width,
                                            ix.n_out, cudaMemcpyHostToDevice, ctx->stream_in));
        ULTRA_CUDA_OK(cudaEventRecord(ctx->uploaded[s], ctx->stream_in));
        // compute
        ULTRA_CUDA_OK(cudaStreamWaitEvent(ctx->stream, ctx->uploaded[s], 0));
        ULTRA_CUDA_OK(cudaEventRecord(ctx->tick[2 * c], ctx->stream));
        status = ultra_rspmm_forward(&ix, b[BUF_REL], b[BUF_IN], nullptr, b[BUF_OUT], nullptr, cols, ix.dtype, sum_op, mul_op,
                                     b[BUF_WS], ctx->cap[s][BUF_WS], ctx->stream);
        if (status) return status;
        if (with_backward) {
            status = ultra_rspmm_backward(&ix, b[BUF_REL], b[BUF_IN], b[BUF_OUT], b[BUF_GOUT], b[BUF_GREL], b[BUF_GIN], cols,
                                          ix.dtype, sum_op, mul_op, b[BUF_WS], ctx->cap[s][BUF_WS], ctx->stream);
            if (status) return status;
        }
        ULTRA_CUDA_OK(cudaEventRecord(ctx->tick[2 * c + 1], ctx->stream));
        ULTRA_CUDA_OK(cudaEventRecord(ctx->computed[s], ctx->stream));
        // download
        ULTRA_CUDA_OK(cudaStreamWaitEvent(ctx->stream_out, ctx->computed[s], 0));
        ULTRA_CUDA_OK(cudaMemcpy2DAsync((char *)host_output + offset, host_pitch, b[BUF_OUT], width, width, ix.n_out,
                                        cudaMemcpyDeviceToHost, ctx->stream_out));
        if (with_backward) {
            ULTRA_CUDA_OK(cudaMemcpy2DAsync((char *)host_grad_relation + offset, host_pitch, b[BUF_GREL], width, width,
                                            ix.n_rel, cudaMemcpyDeviceToHost, ctx->stream_out));
            ULTRA_CUDA_OK(cudaMemcpy2DAsync((char *)host_grad_input + offset, host_pitch, b[BUF_GIN], width, width, ix.n_in,
                                            cudaMemcpyDeviceToHost, ctx->stream_out));
        }
        ULTRA_CUDA_OK(cudaEventRecord(ctx->downloaded[s], ctx->stream_out));
    }
    ULTRA_CUDA_OK(cudaStreamSynchronize(ctx->stream_out));
    ULTRA_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    ULTRA_CUDA_OK(cudaStreamSynchronize(ctx->stream_in));
    for (int c = 0; c < n_chunk; ++c) {
        float ms = 0.f;
        ULTRA_CUDA_OK(cudaEventElapsedTime(&ms, ctx->tick[2 * c], ctx->tick[2 * c + 1]));
        ctx->last_ms += ms;
    }
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
