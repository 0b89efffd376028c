/* ultra_rspmm.h - C ABI of the B200-native relational SpMM ("rspmm") hot path.
 *
 * This is the drop-in boundary underneath
 *     torchdrug.layers.functional.generalized_rspmm(sparse, relation, input, sum=, mul=)
 * as called by the reference at ultra/layer.py:134-167 (GeneralizedRelationalConvNBF) and
 * ultra/layer.py:336-369 (GeneralizedRelationalConvNBFMod).  In the reference that Python entry
 * dispatches to torchdrug's pybind functions rspmm_{add,min,max}_{mul,add}_{forward,backward}_{cpu,cuda}
 * (torchdrug/layers/functional/extension/rspmm.{h,cpp,cu}; un-vendored).
 * The entry points below are what a binding for this path binds instead: plain pointers and sizes,
 * no torch types, one CUDA stream argument, integer status codes.
 *
 * Conventions
 *   - all `dev_*` pointers are CUDA device memory on the current device; nothing here allocates or
 *     frees device memory except the ultra_rspmm_ctx_* convenience layer (host-buffer API);
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*) and return without
 *     synchronising, except where stated;
 *   - dense operands are row-major, contiguous: relation (n_rel, dim), input (n_in, dim),
 *     output / grad_output (n_out, dim); feature index = batch * d + channel (reference layer.py:118,306);
 *   - the sparse operand is COO (n_out, n_in, n_rel): row = node_out (destination), col = node_in (source),
 *     layer = relation (reference layer.py:127,328 `graph.adjacency.transpose(0, 1)`); it need not be
 *     coalesced - ultra_rspmm_index_build sorts by (row, col, layer) and merges duplicates by summing
 *     their values, exactly what `sparse.coalesce()` does before torchdrug's coo2csr3d;
 *   - empty rows produce the reduction identity: 0 (add), -FLT_MAX / -DBL_MAX (max), +FLT_MAX / +DBL_MAX (min);
 *   - arg-index = position in coalesce() order (dst, src, rel) of the first edge attaining the extremum,
 *     -1 for empty rows;
 *   - max/min backward follows the reference's all-ties rule (gate `output == message`).
 */
#ifndef ULTRA_RSPMM_H_
#define ULTRA_RSPMM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ULTRA_RSPMM_ABI_VERSION 7

/* status codes (0 = ok).  For ULTRA_RSPMM_ERR_CUDA the cudaError_t is kept per thread, see
 * ultra_rspmm_last_cuda_error(). */
enum {
    ULTRA_RSPMM_OK = 0,
    ULTRA_RSPMM_ERR_ARG = 1,        /* null pointer, negative size, unknown op / dtype code            */
    ULTRA_RSPMM_ERR_WORKSPACE = 2,  /* caller-provided buffer smaller than the *_bytes query answered   */
    ULTRA_RSPMM_ERR_CUDA = 3,       /* a CUDA runtime call failed                                       */
    ULTRA_RSPMM_ERR_INDEX = 4,      /* a COO index is outside (n_out, n_in, n_rel)                      */
    ULTRA_RSPMM_ERR_DTYPE = 5,      /* operand dtype does not match the dtype the index was built for   */
    ULTRA_RSPMM_ERR_RANGE = 6       /* n_out * n_in * n_rel does not fit the 63-bit sort key / int32 ids */
};

/* op codes: torchdrug's NaryOp {add,min,max} x BinaryOp {mul,add} (reference layer.py:18-21 message2mul) */
enum { ULTRA_RSPMM_SUM_ADD = 0, ULTRA_RSPMM_SUM_MIN = 1, ULTRA_RSPMM_SUM_MAX = 2 };
enum { ULTRA_RSPMM_MUL_MUL = 0, ULTRA_RSPMM_MUL_ADD = 1 };
enum { ULTRA_RSPMM_F32 = 0, ULTRA_RSPMM_F64 = 1 };

/* One ordering of the coalesced edges, cut into segments (the rows a pass reduces into) and into
 * tasks (at most `chunk` consecutive edges of one segment; one warp per task and feature slab).
 * A segment with more than one task is "split": its tasks write partial rows (slots) that a
 * fixed-order combine pass folds into the result row, so results do not depend on scheduling. */
typedef struct ultra_rspmm_order {
    int32_t n_seg;        /* rows of the result this order reduces into                                 */
    int32_t n_task;       /* >= n_seg (an empty segment still has one task: it writes the identity)     */
    int32_t n_slot;       /* partial rows needed by the split segments                                  */
    int32_t n_split;      /* number of split segments                                                   */
    int32_t max_seg_nnz;  /* longest segment                                                            */
    int32_t pack_shift;   /* > 0: `packed` holds edge.x | edge.y << pack_shift; 0: ids do not fit 32 bits      */
    int32_t n_gtask;      /* tasks of the grouped list (0: no grouped list was built for this order)            */
    int32_t group_edges;  /* rows of at most this many edges are grouped (0: grouping off)                      */
    const int32_t *ptr;   /* n_seg + 1: edges of segment s are [ptr[s], ptr[s+1])                        */
    const int32_t *edge;  /* M x int2: the two row ids each edge gathers from (see ultra_rspmm_index_t)  */
    const void *w;        /* M merged values, element type = index dtype                                */
    const int32_t *eid;   /* M: position of the edge in canonical coalesce() order (dst, src, rel)        */
    const uint32_t *packed; /* M: both ids of an edge in one word (see pack_shift)                        */
    const int32_t *task;  /* n_task x int4 {seg, begin, end, (slot + 1) | 0x40000000 if any edge weight != 1};
                             slot = -1: the task writes the result row itself.  Sorted by descending edge
                             count (longest first).                                                     */
    const int32_t *split; /* n_split x int4 {seg, first_slot, n_slots, 0}                                */
    const int32_t *gtask; /* n_gtask x int4: the same work with short rows grouped - a group task
                             {first seg, begin, end, 0x20000000 | (rows - 1) << 24 | non-unit flag} covers up to 16
                             consecutive segments of <= group_edges edges each (one warp walks them back to back,
                             so a short row does not pay its own task start-up); longer segments appear as in `task` */
} ultra_rspmm_order_t;

/* Optional extension of the csr / csc order for graphs with at most 4 relation types and at most 864 nodes (the graph
 * of relations of reference rel_model.py:91-147, which has exactly 4): the edges of a segment that share their other
 * endpoint are merged into one *pair* word  other-node id | relation mask << id_bits.  A dense graph of relations has 4
 * edges per pair; the pair kernel reads the other node's row once for all of them. */
typedef struct ultra_rspmm_pairs {
    int32_t n_pair;       /* 0: not built                                                     */
    int32_t id_bits;      /* pair word = other-node id | (bit k set: an edge of relation k) << id_bits */
    const int32_t *ptr;   /* n_seg + 1: pairs of segment s are [ptr[s], ptr[s + 1]), ascending other-node id */
    const uint32_t *pair; /* n_pair words                                                     */
    const int32_t *rows;  /* n_seg: the segments by descending pair count (work order)        */
} ultra_rspmm_pairs_t;

/* Graph index: int32 device arrays describing the coalesced operand in the three edge orders the
 * kernels walk.  A POD the caller keeps on the host; every pointer points into the caller's
 * `index_buffer` (see ultra_rspmm_index_build).  Replaces coo2csr3d + the per-call
 * `sparse.transpose(0,1).coalesce()` of the reference. */
typedef struct ultra_rspmm_index {
    int64_t nnz;          /* coalesced edge count M                                                     */
    int64_t nnz_raw;      /* edge count before merging duplicates                                       */
    int32_t n_out, n_in, n_rel;
    int32_t dtype;        /* ULTRA_RSPMM_F32 / F64: element type of the w arrays                          */
    int32_t unit_weight;  /* 1 if every merged value == 1 (the w stream is not read)                    */
    int32_t chunk;        /* maximum edges per task                                                     */
    ultra_rspmm_order_t csr;  /* forward: segments = destination rows, sorted (dst, rel, src); edge = {src, rel} */
    ultra_rspmm_order_t csc;  /* backward w.r.t. input: segments = source rows, sorted (src, rel, dst); edge = {dst, rel} */
    ultra_rspmm_order_t rel;  /* backward w.r.t. relation: segments = relations, sorted (rel, dst, src); edge = {dst, src} */
    const int32_t *merge_perm;  /* nnz_raw: the caller's edge positions in the order the build merged them            */
    const int32_t *merge_start; /* nnz + 1: csr edge m is the sum of merge_perm[merge_start[m] .. merge_start[m + 1])  */
    /* ---- optional extensions, filled by ultra_rspmm_index_extend (all zero / null when absent) ---- */
    ultra_rspmm_pairs_t pairs[2];   /* [0] csr order (forward), [1] csc order (gradient w.r.t. input)                 */
    const int32_t *block_ptr;   /* n_rel x (2 n_block + 1): position in the rel order where the edges of relation k with
                                   destination >= h * block_rows / 2 start - a (relation, destination block) run is the
                                   contiguous range between entries 2 b and 2 b + 2, its two half blocks split at 2 b + 1 */
    const int32_t *block_split; /* n_rel x int4 {rel, rel * n_block, n_block, 0}: combine list of the blocked pass     */
    int32_t block_rows;         /* destination rows per block (their grad_output slab is staged in shared memory)      */
    int32_t n_block;
} ultra_rspmm_index_t;

/* ---- version / diagnostics ------------------------------------------------------------------- */
int ultra_rspmm_abi_version(void);
int ultra_rspmm_last_cuda_error(void);          /* cudaError_t of the last failure on this thread */
const char *ultra_rspmm_status_string(int status);
/* number of kernels this library has enqueued since load / last reset (process-wide counter) */
int64_t ultra_rspmm_launch_count(void);
void ultra_rspmm_launch_count_reset(void);
/* How the most recent pass of each kind was launched (process-wide, for tests and profiles): which kernel served it
 * and with which template switches.  pass: 0 = forward (csr order), 1 = gradient w.r.t. input (csc order),
 * 2 = gradient w.r.t. relation. */
enum { ULTRA_RSPMM_PASS_FORWARD = 0, ULTRA_RSPMM_PASS_GRAD_INPUT = 1, ULTRA_RSPMM_PASS_GRAD_RELATION = 2 };
enum {
    ULTRA_RSPMM_KERNEL_NONE = 0,
    ULTRA_RSPMM_KERNEL_SEG_REDUCE = 1,    /* generic gather-combine-reduce (sum / min / max, all three orders)        */
    ULTRA_RSPMM_KERNEL_SEG_GATED = 2,     /* min / max backward (all-ties gate)                                       */
    ULTRA_RSPMM_KERNEL_SEG_PNA = 3,       /* four PNA aggregates in one pass                                          */
    ULTRA_RSPMM_KERNEL_ROWS_IN_SMEM = 4,  /* few-row operands: the gathered slab is staged in shared memory           */
    ULTRA_RSPMM_KERNEL_DST_BLOCKED = 5,   /* grad_relation with grad_output rows of a destination block in shared memory */
    ULTRA_RSPMM_KERNEL_PAIRS_IN_SMEM = 6, /* few-row operands with <= 4 relation types: one row read per (node, node) pair */
    ULTRA_RSPMM_KERNEL_SUBWARP_ROWS = 7,  /* slabs beyond L2: 2 or 4 tasks per warp, 256- / 128-byte slabs (n_slab tells which) */
    ULTRA_RSPMM_KERNEL_DST_BLOCKED_GATED = 8 /* min / max grad_relation with grad_output and output rows of half a block staged */
};
typedef struct ultra_rspmm_pass_info {
    int32_t kernel;    /* ULTRA_RSPMM_KERNEL_*                                        */
    int32_t vec;       /* features per lane (slab = 32 * vec features; 64 for ROWS_IN_SMEM) */
    int32_t keep;      /* L2 eviction-priority hints on (slab > 24 MiB)               */
    int32_t grouped;   /* grouped task list (short rows share a warp task)            */
    int32_t packed;    /* edge ids packed into one 32-bit word                        */
    int32_t n_task;    /* tasks per slab                                              */
    int32_t n_slab;    /* feature slabs                                               */
    int32_t n_split;   /* segments folded from partial rows by the combine pass       */
} ultra_rspmm_pass_info_t;
int ultra_rspmm_last_pass_info(int32_t pass, ultra_rspmm_pass_info_t *info);
/* tuning knobs (process-wide; 0 keeps the current value).  chunk: edges per task for indexes built
 * afterwards (default 256; rows of up to chunk / 4 edges are grouped, see `gtask`).  variant: L2 eviction-priority hints of the gather kernels - 0 = automatic
 * (on when the gathered slab exceeds 24 MiB), 1 = never, 2 = always.
 * l2_budget_bytes: L2 bytes the gathered operand's slab (rows x slab width) may occupy; the slab width
 * (512 / 256 / 128 bytes per row) is the widest that fits (default: unlimited, i.e. always 512). */
int ultra_rspmm_set_tuning(int32_t chunk, int32_t variant, int64_t l2_budget_bytes);
/* Rows-in-shared-memory kernel for operands with at most 864 rows (the graph of relations, reference
 * rel_model.py:253-257): 0 = never, 1 = automatic (default: when the slab refills are small against the edge work),
 * 2 = whenever the operand fits (tests). */
int ultra_rspmm_set_staged(int32_t mode);
/* Sub-warp rows kernel (experimental, off by default: slab_bytes = 0) for graphs whose 512-byte column slab of the gathered
 * operand exceeds `slab_bytes`: 2 or 4 tasks per warp over 256- or 128-byte slabs, so that the slab is L2-resident again.
 * Measured slower than the generic kernel at every shape of the configs[4] sweep.  sub: 0 = chosen by the slab
 * size, 2 / 4 = forced (tests). */
int ultra_rspmm_set_narrow(int64_t slab_bytes, int32_t sub);
/* Index extensions built by ultra_rspmm_index_extend afterwards (0 = never, 1 = automatic, 2 = whenever the graph
 * qualifies structurally): pair lists; destination-block table. */
int ultra_rspmm_set_extensions(int32_t pairs, int32_t blocked);

/* Gather-bandwidth probe (diagnostics, used by bench.py for the roofline denominators it reports): every warp of
 * `blocks` x 8 reads `iters` (multiple of 4) 512-byte row pieces at pseudo-random rows of dev_buffer (rows x
 * row_stride_bytes, stride >= 512, % 16 == 0) - the access pattern of the rspmm gather without ids or arithmetic.
 * A footprint of rows x 512 B well inside L2 measures the L2 -> SM ceiling, one far beyond L2 the HBM ceiling for
 * random rows.  *bytes_read = blocks * 8 * iters * 512.  Asynchronous; the caller times the stream. */
int ultra_probe_gather(const void *dev_buffer, int64_t rows, int64_t row_stride_bytes, int32_t iters, int32_t blocks,
                       float *dev_sink, int64_t *bytes_read, void *stream);

/* ---- index build (replaces sparse.coalesce() + coo2csr3d) ---------- */
/* Bytes needed for the index arrays (upper bound, from the raw edge count) and for scratch. */
int ultra_rspmm_index_bytes(int64_t nnz_raw, int32_t n_out, int32_t n_in, int32_t n_rel, int32_t dtype,
                            size_t *index_bytes, size_t *scratch_bytes);
/* dev_indices: int64 (3, nnz_raw) = rows [node_out; node_in; relation], row r starting at
 * dev_indices + r * index_stride (torch `sparse._indices()` of the transposed adjacency; stride =
 * nnz_raw when contiguous).  dev_values: nnz_raw values of `dtype`.  Both buffers must be 256-byte
 * aligned (cudaMalloc / torch allocations are).  Fills *index (host POD).  Synchronises `stream`
 * (the merged edge and task counts must reach the host). */
int ultra_rspmm_index_build(const int64_t *dev_indices, int64_t index_stride, const void *dev_values,
                            int64_t nnz_raw, int32_t n_out, int32_t n_in, int32_t n_rel, int32_t dtype,
                            void *index_buffer, size_t index_bytes, void *scratch, size_t scratch_bytes,
                            ultra_rspmm_index_t *index, void *stream);
/* The same edge structure with other values: `dev_values` holds nnz_raw new values in the
 * caller's original edge order (the order of dev_indices at build time).  Fills *derived, a copy of *base whose w arrays
 * and task lists (per-task "all weights are 1" flags) live in `buffer` (ultra_rspmm_index_derive_bytes; 256-byte aligned)
 * and whose structure arrays still point into base's buffer - keep both alive.  Duplicates are summed in the build's
 * order.  Asynchronous, no host synchronisation: a weight-0 mask over a fixed graph (what `remove_easy_edges`,
 * reference model.py:57-74, amounts to under sum aggregation) costs three small kernels instead of an index rebuild. */
int ultra_rspmm_index_derive_bytes(const ultra_rspmm_index_t *base, size_t *bytes);
int ultra_rspmm_index_derive(const ultra_rspmm_index_t *base, const void *dev_values, void *buffer, size_t bytes,
                             ultra_rspmm_index_t *derived, void *stream);
/* Optional extensions of a built index (call once after ultra_rspmm_index_build, before the first forward):
 *   - pair lists (see ultra_rspmm_pairs_t) when n_rel <= 4, n_out and n_in <= 864, all weights 1 and the edges average
 *     at least 1.5 per pair: the forward and grad_input passes then run the pair kernel;
 *   - the destination-block table of the rel order when the two gathered slabs of the grad_relation pass exceed L2
 *     ((n_out + n_in) x 512 B > 96 MB): grad_relation then stages grad_output rows block by block in shared memory
 *     and gathers only input rows (3 row gathers per edge and fwd+bwd step instead of 4).
 * Fills the extension fields of *index; `buffer` (ultra_rspmm_index_extend_bytes, 256-byte aligned) must outlive the
 * index.  Synchronises `stream`.  Indexes made by ultra_rspmm_index_derive inherit the block table, not the pairs. */
int ultra_rspmm_index_extend_bytes(const ultra_rspmm_index_t *index, size_t *buffer_bytes);
int ultra_rspmm_index_extend(ultra_rspmm_index_t *index, void *buffer, size_t buffer_bytes, void *stream);

/* 128-bit content fingerprint of (indices, values) written to dev_out[2] (uint64); lets a caller
 * recognise an edge set it already indexed without a sort.  dev_out must be zero on entry is NOT
 * required (the call clears it).  Asynchronous. */
int ultra_rspmm_fingerprint(const int64_t *dev_indices, int64_t index_stride, const void *dev_values,
                            int64_t nnz_raw, int32_t dtype, uint64_t *dev_out, void *stream);

/* ---- forward: rspmm_{sum}_{mul}_forward_cuda -------------------------------------------------- */
/* workspace bytes for forward / backward at feature width `dim` (may be 0) */
int ultra_rspmm_workspace_bytes(const ultra_rspmm_index_t *index, int64_t dim, int32_t dtype,
                                size_t *forward_bytes, size_t *backward_bytes);
/* output[i,:] = (sum)_{(i,j,k)} w * (relation[k,:] (mul) input[j,:]);  dev_argidx (n_out, dim) int32 may be
 * NULL (only meaningful for min/max).  dev_addend (n_out, dim), sum_op = add only, may be NULL: it is added
 * to the reduced rows in the kernel epilogue - the `update + boundary` of reference layer.py:156,358. */
int ultra_rspmm_forward(const ultra_rspmm_index_t *index, const void *dev_relation, const void *dev_input,
                        const void *dev_addend, void *dev_output, int32_t *dev_argidx, int64_t dim, int32_t dtype,
                        int32_t sum_op, int32_t mul_op, void *workspace, size_t workspace_bytes, void *stream);

/* Forward (sum = add, fp32) on operands that live inside a wider buffer: a row of `input` / `output` is dim / block
 * blocks of `block` features, consecutive blocks `*_block_stride` elements apart, the output shifted by
 * `output_block_offset` inside its stride.  With block = 64, strides = 128, offset = 64 the operator reads the layer
 * input from the left halves and writes `update + boundary` into the right halves of the (N, B, 128) buffer that the
 * layer's Linear consumes - the `torch.cat([input, update], dim=-1)` of reference layer.py:387 disappears.
 * relation (n_rel, dim) and addend (n_out, dim, may be NULL) are plain matrices.  All strides / offsets % 4 == 0. */
int ultra_rspmm_forward_blocked(const ultra_rspmm_index_t *index, const void *dev_relation, const void *dev_input,
                                const void *dev_addend, void *dev_output, int64_t dim, int32_t dtype, int32_t mul_op,
                                int64_t block, int64_t input_block_stride, int64_t output_block_stride,
                                int64_t output_block_offset, void *workspace, size_t workspace_bytes, void *stream);

/* PNA aggregation in one pass: the four operator calls of reference layer.py:141-144 / 164-167 / 343-346 / 366-369
 * (add, add of the squared operands, max, min) over one gather per edge.  Outputs (n_out, dim) each.
 * workspace: 4 x the forward partial-row bytes (4 * csr.n_slot * dim * sizeof(element)); forward only. */
int ultra_rspmm_forward_pna(const ultra_rspmm_index_t *index, const void *dev_relation, const void *dev_input,
                            void *dev_sum, void *dev_square_sum, void *dev_max, void *dev_min, int64_t dim,
                            int32_t dtype, int32_t mul_op, void *workspace, size_t workspace_bytes, void *stream);

/* ---- backward: rspmm_{sum}_{mul}_backward_cuda (overload without value_grad) ------------------- */
/* dev_output is read only for min/max (may be NULL for add).  Either gradient pointer may be NULL to
 * skip that pass.  Results are written (not accumulated). */
int ultra_rspmm_backward(const ultra_rspmm_index_t *index, const void *dev_relation, const void *dev_input,
                         const void *dev_output, const void *dev_grad_output, void *dev_grad_relation,
                         void *dev_grad_input, int64_t dim, int32_t dtype, int32_t sum_op, int32_t mul_op,
                         void *workspace, size_t workspace_bytes, void *stream);
/* The same for sum_op = add with  grad_input = (operator's gradient) + dev_grad_input_addend  ((n_in, dim), not aliasing
 * dev_grad_input; NULL = plain backward): the layer input of an NBFNet layer also feeds the Linear of `combine` and the
 * short-cut (reference layer.py:386-392, model.py:126-127), and their gradients arrive here instead of in two more passes
 * over (N, D) tensors (the framework's gradient accumulation). */
int ultra_rspmm_backward_addend(const ultra_rspmm_index_t *index, const void *dev_relation, const void *dev_input,
                                const void *dev_grad_output, void *dev_grad_relation, void *dev_grad_input,
                                const void *dev_grad_input_addend, int64_t dim, int32_t dtype, int32_t mul_op,
                                void *workspace, size_t workspace_bytes, void *stream);

/* ---- host-buffer convenience layer (what a non-torch host binds; used for end-to-end timing) --- */
/* Owns device copies of one graph's index and of the dense operands; every call below takes HOST
 * pointers, copies host->device, runs the kernels and copies the results back, synchronising before
 * it returns.  Pinned host memory makes the copies asynchronous DMA; pageable memory also works. */
typedef struct ultra_rspmm_ctx ultra_rspmm_ctx_t;
int ultra_rspmm_ctx_create(ultra_rspmm_ctx_t **ctx, int32_t device);
int ultra_rspmm_ctx_destroy(ultra_rspmm_ctx_t *ctx);
/* host_indices: int64 (3, nnz_raw) contiguous; host_values: nnz_raw of dtype */
int ultra_rspmm_ctx_set_graph(ultra_rspmm_ctx_t *ctx, const int64_t *host_indices, const void *host_values,
                              int64_t nnz_raw, int32_t n_out, int32_t n_in, int32_t n_rel, int32_t dtype);
int ultra_rspmm_ctx_forward(ultra_rspmm_ctx_t *ctx, const void *host_relation, const void *host_input,
                            void *host_output, int64_t dim, int32_t sum_op, int32_t mul_op);
/* forward + backward in one call: uploads relation, input, grad_output; downloads output,
 * grad_relation, grad_input */
int ultra_rspmm_ctx_forward_backward(ultra_rspmm_ctx_t *ctx, const void *host_relation, const void *host_input,
                                     const void *host_grad_output, void *host_output, void *host_grad_relation,
                                     void *host_grad_input, int64_t dim, int32_t sum_op, int32_t mul_op);
/* device time (ms, CUDA events on the context's stream) of the kernels of the last ctx_* call */
float ultra_rspmm_ctx_last_kernel_ms(const ultra_rspmm_ctx_t *ctx);
/* coalesced edge count of the context's graph (-1 before set_graph) */
int64_t ultra_rspmm_ctx_nnz(const ultra_rspmm_ctx_t *ctx);
/* pinned host memory helpers for callers without a CUDA runtime binding of their own */
int ultra_rspmm_host_alloc(void **ptr, size_t bytes);
int ultra_rspmm_host_free(void *ptr);

/* ---- layer epilogue (reference layer.py:184-190, 386-392 + model.py:126-127) ---- */
/* out = relu(layer_norm(x + linear_bias; eps) * gamma + beta) + residual over `rows` rows of `dim` fp32
 * features (dim in {4, 8, ..., 128}); linear_bias, gamma/beta (both or neither) and residual may be NULL;
 * relu = 0 skips the activation.  x is the output of the layer's Linear without its bias (a cuBLAS GEMM
 * that stays in PyTorch). */
int ultra_layer_norm_relu_residual(const float *dev_x, const float *dev_linear_bias, const float *dev_gamma, const float *dev_beta,
                                   const float *dev_residual, float *dev_out, int64_t rows, int32_t dim, float eps,
                                   int32_t relu, void *stream);

/* Backward of the fused epilogue (fine-tuning): given d(out) it writes d(x) and the column sums d(linear_bias),
 * d(gamma), d(beta) (any of the three may be NULL); d(residual) = d(out).  Row statistics are recomputed from x.
 * workspace: ultra_layer_norm_relu_residual_backward_bytes(dim).  Deterministic (no atomics). */
/* Strided form of the epilogue: row r of out / residual starts at r * out_stride / r * residual_stride elements
 * (strides % 4 == 0, >= dim), x stays a plain (rows, dim) matrix - lets the epilogue read its short-cut from, and write
 * its result into, the left halves of the (N, B, 128) layer buffers (see ultra_rspmm_forward_blocked). */
int ultra_layer_norm_relu_residual_strided(const float *dev_x, const float *dev_linear_bias, const float *dev_gamma,
                                           const float *dev_beta, const float *dev_residual, float *dev_out,
                                           int64_t rows, int32_t dim, int64_t residual_stride, int64_t out_stride,
                                           float eps, int32_t relu, void *stream);
int ultra_layer_norm_relu_residual_backward_bytes(int32_t dim, size_t *workspace_bytes);
int ultra_layer_norm_relu_residual_backward(const float *dev_x, const float *dev_linear_bias, const float *dev_gamma,
                                            const float *dev_beta, const float *dev_grad_out, float *dev_grad_x,
                                            float *dev_grad_linear_bias, float *dev_grad_gamma, float *dev_grad_beta,
                                            int64_t rows, int32_t dim, float eps, int32_t relu, void *workspace,
                                            size_t workspace_bytes, void *stream);

/* Fused Linear + epilogue of the layer (inference): out[r, 0:out_dim] = relu(layer_norm(input[r, 0:2*out_dim] @ W^T +
 * linear_bias) * gamma + beta) + (shortcut ? input[r, 0:out_dim] : 0) for `rows` rows; a row of input is
 * [layer input | update + boundary] (row stride input_stride >= 2 * out_dim, % 4 == 0), W is (out_dim, 2 * out_dim)
 * row-major, out rows are out_stride apart.  out_dim in {32, 64}.  fp32 accuracy on the tensor cores (3xTF32 split with
 * fp32 accumulation, see csrc/layer_linear.cu); replaces F.linear + ultra_layer_norm_relu_residual_strided. */
int ultra_layer_linear_norm_relu_residual(const float *dev_input, int64_t input_stride, const float *dev_weight,
                                          const float *dev_linear_bias, const float *dev_gamma, const float *dev_beta,
                                          float *dev_out, int64_t out_stride, int64_t rows, int32_t out_dim, float eps,
                                          int32_t relu, int32_t shortcut, void *stream);

/* The same with the two halves of the Linear's input in two tensors: input[r, 0:out_dim] = layer input (also the short-cut),
 * update[r, 0:out_dim] = update + boundary - the operator then reads and writes plain (N, B * d) matrices and neither a
 * `cat` nor an interleaved (N, B, 2d) buffer exists.  tcgen05 + TMA only (two tensor maps); 16-byte aligned pointers. */
int ultra_layer_linear_norm_relu_residual_two(const float *dev_input, int64_t input_stride, const float *dev_update,
                                              int64_t update_stride, const float *dev_weight, const float *dev_linear_bias,
                                              const float *dev_gamma, const float *dev_beta, float *dev_out,
                                              int64_t out_stride, int64_t rows, int32_t out_dim, float eps, int32_t relu,
                                              int32_t shortcut, void *stream);
/* The same with a second output: dev_pre_out[r, 0:out_dim] = the Linear's output WITHOUT its bias (the accumulators), which
 * the backward of the layer epilogue (ultra_layer_norm_relu_residual_backward) needs - the training forward of a layer is
 * then one kernel instead of a GEMM and an epilogue pass. */
int ultra_layer_linear_norm_relu_residual_two_pre(const float *dev_input, int64_t input_stride, const float *dev_update,
                                                  int64_t update_stride, const float *dev_weight, const float *dev_linear_bias,
                                                  const float *dev_gamma, const float *dev_beta, float *dev_out,
                                                  int64_t out_stride, float *dev_pre_out, int64_t pre_stride, int64_t rows,
                                                  int32_t out_dim, float eps, int32_t relu, int32_t shortcut, void *stream);

/* ---- the Linear of `combine` under autograd (fine-tuning; reference layer.py:386-392) on the tensor cores -------------------
 * fp32 accuracy through the 3xTF32 split, no cuBLAS SIMT SGEMM, no `cat([input, update])`.
 * ultra_layer_rows_gemm:  out[r, 0:n_out] = [a0[r, :] | a1[r, :]] @ weight^T  for `rows` rows (tcgen05 + TMA).
 *     (n_out, k_in) = (64, 128): the forward Linear; a0 = layer input, a1 = update (64 columns each, a1 may be NULL when a0
 *                     already holds all 128 columns), weight (64, 128) row-major, out0 (rows, 64).
 *     (n_out, k_in) = (128, 64): gradient w.r.t. [input | update] = dx @ W; a0 = dx (64 columns), a1 = NULL, weight = W^T
 *                     (128, 64) row-major; columns [0, 64) go to out0 (+ addend0 when given: the short-cut's gradient),
 *                     columns [64, 128) to out1 (or to out0 + 64 when out1 is NULL).
 *     Row strides (ld*) in floats, multiples of 4; all pointers 16-byte aligned.
 * ultra_layer_rows_gemm_weight:  weight_grad[n, k] = sum_r dx[r, n] * [a0[r, :] | a1[r, :]][k]  -> (64, 128); mma.sync over
 *     row tiles with per-CTA partial results folded in fixed order (deterministic).  workspace: *_weight_bytes. */
int ultra_layer_rows_gemm(const float *dev_a0, int64_t lda0, const float *dev_a1, int64_t lda1, const float *dev_weight,
                          float *dev_out0, int64_t ld_out0, float *dev_out1, int64_t ld_out1, const float *dev_addend0,
                          int64_t ld_addend0, int64_t rows, int32_t n_out, int32_t k_in, void *stream);
int ultra_layer_rows_gemm_weight_bytes(size_t *workspace_bytes);
int ultra_layer_rows_gemm_weight(const float *dev_dx, int64_t ld_dx, const float *dev_a0, int64_t lda0, const float *dev_a1,
                                 int64_t lda1, int64_t rows, float *dev_weight_grad, void *workspace, size_t workspace_bytes,
                                 void *stream);

/* which implementation serves ultra_layer_linear_norm_relu_residual: 0 = library default, 1 = mma.sync, 2 = tcgen05 */
int ultra_layer_linear_set_kernel(int32_t kind);
int ultra_layer_linear_get_kernel(void);   /* the implementation in effect: 1 or 2 */
/* development: when set (8 int64 per CTA, >= 8 * SM count), the tcgen05 kernel records the cycles each warp role spent
 * waiting on its barriers: [tma/empty, split/landed, split/lo_empty, mma/tmem_empty, mma/full, epilogue/tmem_full, total
 * cycles, tiles].  NULL switches it off (default). */
int ultra_layer_linear_set_debug(long long *dev_buffer);

/* ---- scoring head (reference model.py:177-193, the 2-layer MLP over [hidden | query]) ---- */
/* score[r] = bias[0] + sum_c weight[c] * relu(z[r, c] + query_bias[r % batch, c]) over `rows` rows of `dim` fp32 features
 * (dim in {4, 8, ..., 128}).  z = hidden @ W1[:, :d]^T (a cuBLAS GEMM that stays in PyTorch), query_bias (batch, dim) =
 * query @ W1[:, d:]^T + b1, weight (dim) = W2[0], bias (1 element, may be NULL) = b2.  Rows are (node, query) pairs with
 * the query index fastest. */
int ultra_score_head(const float *dev_z, const float *dev_query_bias, const float *dev_weight, const float *dev_bias,
                     float *dev_score, int64_t rows, int32_t batch, int32_t dim, void *stream);

/* The split scoring head in one kernel (inference): score[r] = out_bias[0] + out_weight . relu(W1[:, 0:in_dim] input[r, 0:in_dim]
 * + query_bias[r % batch]) for `rows` rows; W1 rows are weight_stride floats apart (the first in_dim columns of the MLP's
 * (2 in_dim, 2 in_dim) first Linear), query_bias is (batch, 2 in_dim) = query @ W1[:, in_dim:]^T + b1, out_weight (2 in_dim).
 * in_dim in {32, 64}.  fp32 accuracy on the tensor cores (3xTF32, as ultra_layer_linear_norm_relu_residual); replaces the
 * K = in_dim cuBLAS GEMM + ultra_score_head. */
int ultra_score_head_linear(const float *dev_input, int64_t input_stride, const float *dev_weight, int64_t weight_stride,
                            const float *dev_query_bias, const float *dev_out_weight, const float *dev_out_bias,
                            float *dev_score, int64_t rows, int32_t batch, int32_t in_dim, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ULTRA_RSPMM_H_ */
