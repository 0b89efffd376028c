// Gather-combine-reduce kernels of the rspmm hot path (sm_100a) and their C-ABI launchers.
//
// One kernel shape serves the forward pass and both atomic-free backward passes
// (DESIGN.md "Kernels"): a *task* is at most `chunk` consecutive edges of one segment of an edge
// order; a warp owns (task, feature slab of 32*VEC features) and reduces
//     acc[:] (+)= w_e * f(A[edge.x, slab], B[edge.y, slab])
// in registers, then writes the result row (or a partial row for split segments).
//
//   pass                      order  segment   A (gathered)     B                     f
//   forward                   csr    dst i     input[src j]     relation[k] (table)   mul | add
//   backward wrt input  (add) csc    src j     grad_out[dst i]  relation[k] (table)   mul | copy A
//   backward wrt relation(add) rel   rel k     grad_out[dst i]  input[src j] (gather) mul | copy A
//   backward (min/max)        csc / rel: gated variant, see seg_gated_kernel
//
// Warps are numbered slab-major (all tasks of slab 0, then slab 1, ...), so the CTAs resident at
// any moment read one (rows x SLAB) column block of the gathered operand - it stays in the 126 MB
// L2 - and one (n_rel x SLAB) block of the relation table, which stays in L1.  HBM traffic is then
// the compulsory bytes; the gathers are L2 hits.  No atomics anywhere: results are bit-reproducible.
#include <cstdlib>
#include <initializer_list>
#include <type_traits>

#include "rspmm_common.cuh"

namespace ultra {

namespace {

constexpr int kUnroll = 4;
static_assert(kUnroll == 4, "the tail switch in seg_reduce_kernel enumerates 1..3 leftover edges");   // 4 edges in flight per warp at 4 CTAs/SM beat 2x5, 6x3, 8x3, 8x2 (profiles/README.md)

template <typename T> struct SegArgs {
    const int32_t *ptr;       // segment pointers of the order (group tasks read their rows' boundaries here)
    const int4 *task;
    const int2 *edge;
    const unsigned *packed;   // edge ids packed as x | y << pack_shift (null when they do not fit 32 bits)
    int pack_shift;
    const int32_t *eid;       // ARG only: canonical coalesce() position of each edge
    const T *w;        // null when all weights are 1
    const T *A;
    const T *B;
    T *out;
    const T *addend;        // optional (n_seg, dim): added to the reduced row (the layer's `update + boundary`)
    T *partial;
    int32_t *arg_out;       // ARG only
    int32_t *partial_arg;   // ARG only
    long long dim;
    int n_task;
    int n_slab;
    int keep;   // 1: mark the gathered slab evict_last in L2 and the edge-id stream evict_first
    int grouped;   // 1: `task` is the grouped list (short rows share a task)
    // blocked layout (block == 0: A and out are plain (rows, dim) matrices).  Otherwise a row of A / out is dim / block
    // blocks of `block` features, consecutive blocks a_stride / o_stride elements apart, out shifted by o_offset:
    // the (N, B, 2d) buffer whose halves are the layer input and the layer update (reference layer.py:387 `cat`)
    // block_shift = log2(block) when block is a power of two, else -1; a_row / o_row = row strides in elements (= dim for
    // plain matrices), precomputed on the host so that the kernel does no 64-bit division.
    long long a_stride, o_stride, o_offset, a_row, o_row;
    int block, block_shift;
};

template <typename T, int SUM> __device__ __forceinline__ void reduce_into(T &acc, T m) {
    if (SUM == ULTRA_RSPMM_SUM_ADD) acc += m;
    else if (SUM == ULTRA_RSPMM_SUM_MAX) acc = acc > m ? acc : m;   // torchdrug NaryMax::forward
    else acc = acc < m ? acc : m;                                  // torchdrug NaryMin::forward
}

__device__ __forceinline__ int id_x(const int2 &e) { return e.x; }
__device__ __forceinline__ int id_y(const int2 &e) { return e.y; }
__device__ __forceinline__ int id_x(const unsigned &) { return 0; }
__device__ __forceinline__ int id_y(const unsigned &) { return 0; }
__device__ __forceinline__ unsigned id_bits(const unsigned &e) { return e; }
__device__ __forceinline__ unsigned id_bits(const int2 &) { return 0; }

// row address = base + row * row_bytes as one IMAD.WIDE.U32 (row ids are int32 >= 0, row_bytes < 2^32)
template <typename T> __device__ __forceinline__ const T *row_ptr(const char *base, int row, unsigned row_bytes) {
    return reinterpret_cast<const T *>(base + (unsigned long long)(unsigned)row * row_bytes);
}

// Column c of a logical (rows, dim) matrix stored in blocks: (c / block) * stride + c % block, in 32-bit arithmetic
// (c < 2^30: launch_seg checks dim * sizeof(T) < 2^32) and with a shift when the block is a power of two.
__device__ __forceinline__ long long blocked_col(long long col, int block, int shift, long long stride) {
    const unsigned c = (unsigned)col;
    const unsigned q = shift >= 0 ? c >> shift : c / (unsigned)block;
    return (long long)q * stride + (c - q * (unsigned)block);
}

// The issue-slot and L1-data-pipe budgets of the inner loop are what bound these kernels once the gathers are L2
// hits (profiles/README.md).  Per edge and warp the loop below is: a broadcast shared-memory read of the edge
// ids (staged by the whole warp; one LDS.128 serves 4 edges when the two ids pack into 32 bits), 2 IMAD.WIDE
// (row addresses), 2 LDG.128, VEC FFMA - no shuffles, no predicates; the ragged tail of a task runs in a separate
// single-edge loop.  Tasks whose edges all have weight 1 (flag in task.w) skip the weight stream and multiply.
template <typename T, int VEC, int SUM, int MSG, bool B_TABLE, bool ARG, bool PACKED, bool KEEP, bool GROUPED>
__global__ void __launch_bounds__(kThreadsPerBlock, 4) seg_reduce_kernel(const SegArgs<T> a) {
    using Ids = typename std::conditional<PACKED, unsigned, int2>::type;
    __shared__ __align__(16) Ids s_edge[kWarpsPerBlock][32];
    __shared__ __align__(16) T s_w[kWarpsPerBlock][32];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long gw = (long long)blockIdx.x * kWarpsPerBlock + warp;
    if (gw >= (long long)a.n_task * a.n_slab) return;   // warps never meet at a block barrier
    const int slab = (int)(gw / a.n_task);
    const int4 task = __ldg(a.task + (gw - (long long)slab * a.n_task));
    const int slot = task_slot(task.w);
    // a group task walks `rows` consecutive short segments back to back: lane l keeps the start of row l (lane `rows`
    // the end of the last one); a plain task is the case rows == 1
    const int rows = GROUPED ? task_rows(task.w) : 1;
    int boundary_of_lane = task.z;
    if (GROUPED && rows > 1 && lane <= rows) boundary_of_lane = __ldg(a.ptr + task.x + lane);
    int row_local = 0;
    int next_boundary = GROUPED && rows > 1 ? __shfl_sync(kFullMask, boundary_of_lane, 1) : task.z;
    const long long col = (long long)slab * (32 * VEC) + lane * VEC;
    const bool active = col < a.dim;
    const long long safe_col = active ? col : 0;   // idle lanes (dim % (32 * VEC) != 0) read column 0, store nothing
    const unsigned row_bytes = (unsigned)(a.dim * sizeof(T));
    // blocked operands: column c of the logical (rows, dim) matrix lives at (c / block) * stride + c % block
    const long long a_col = a.block ? blocked_col(safe_col, a.block, a.block_shift, a.a_stride) : safe_col;
    const unsigned a_row_bytes = (unsigned)(a.a_row * sizeof(T));
    const char *A = reinterpret_cast<const char *>(a.A + a_col);
    const char *B = reinterpret_cast<const char *>(a.B + safe_col);
    const Ids *ids = reinterpret_cast<const Ids *>(PACKED ? (const void *)a.packed : (const void *)a.edge);
    const int shift = a.pack_shift;
    const unsigned low = PACKED ? (shift >= 32 ? 0xffffffffu : ((1u << shift) - 1u)) : 0u;
    // KEEP (large slabs): gathered rows are marked evict_last in L2, the edge-id stream evict_first
    const unsigned long long keep_policy = KEEP ? policy_evict_last() : 0, once_policy = KEEP ? policy_evict_first() : 0;
    auto gather = [&](const T *p, Vec<T, VEC> &v) {
        if (KEEP) gather_load_keep(p, v, keep_policy);
        else gather_load(p, v);
    };
    auto load_ids = [&](const Ids *p) { return KEEP ? edge_load_once(p, once_policy) : __ldg(p); };
    auto first_id = [&](const Ids &e) { return PACKED ? (int)(id_bits(e) & low) : id_x(e); };
    auto second_id = [&](const Ids &e) { return PACKED ? (int)(id_bits(e) >> shift) : id_y(e); };

    T acc[VEC];
    int32_t arg[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        acc[v] = reduce_identity<T, SUM>();
        arg[v] = -1;
    }
    // write the finished row (or partial row) and start the next one; warp-uniform
    auto flush = [&]() {
        if (active) {
            Vec<T, VEC> r;
#pragma unroll
            for (int v = 0; v < VEC; ++v) r.v[v] = acc[v];
            const long long row = task.x + row_local;
            if (slot < 0) {
                if (a.addend) {
                    Vec<T, VEC> b;
                    gather_load(a.addend + row * a.dim + col, b);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) r.v[v] += b.v[v];
                }
                // (the blocked output position is recomputed here rather than kept live across the edge loop)
                const long long o_col = a.block ? blocked_col(col, a.block, a.block_shift, a.o_stride) + a.o_offset : col;
                stream_store(a.out + row * a.o_row + o_col, r);
            } else {
                T *p = a.partial + (long long)slot * a.dim + col;   // re-read soon by the combine pass: default policy
#pragma unroll
                for (int v = 0; v < VEC; ++v) p[v] = r.v[v];
            }
            if (ARG) {
                int32_t *p = slot < 0 ? a.arg_out + row * a.dim + col : a.partial_arg + (long long)slot * a.dim + col;
#pragma unroll
                for (int v = 0; v < VEC; ++v) p[v] = arg[v];
            }
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            acc[v] = reduce_identity<T, SUM>();
            arg[v] = -1;
        }
        ++row_local;
        if (GROUPED) {
            const int upcoming = __shfl_sync(kFullMask, boundary_of_lane, (row_local + 1) & 31);
            next_boundary = row_local < rows ? upcoming : 0x7fffffff;
        }
    };

    auto reduce_task = [&](auto unit_tag) {
        constexpr bool UNIT = decltype(unit_tag)::value;
        // the edges of a segment are grouped by relation (index build), so the relation row of the previous edge
        // is kept in registers: a batch whose edges all use it issues no table loads (half the L1 wavefronts)
        constexpr bool CACHE = B_TABLE && MSG != MSG_COPY;
        Vec<T, VEC> cached;
        int cached_row = -1;
        auto accumulate = [&](const Vec<T, VEC> &va, const Vec<T, VEC> &vb, T w, int position) {
            const int canonical = ARG ? __ldg(a.eid + position) : 0;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const T m = UNIT ? message<T, MSG>(vb.v[v], va.v[v]) : message<T, MSG>(w, vb.v[v], va.v[v]);
                if (ARG) {   // first edge in canonical order attaining the extremum
                    const bool better = SUM == ULTRA_RSPMM_SUM_MAX ? (m > acc[v]) : (m < acc[v]);
                    if (better || (m == acc[v] && canonical < arg[v])) arg[v] = canonical;
                }
                reduce_into<T, SUM>(acc[v], m);
            }
        };
        auto load_table = [&](int row, Vec<T, VEC> &vb) {
            if (B_TABLE) table_load(row_ptr<T>(B, row, row_bytes), vb);
            else gather(row_ptr<T>(B, row, row_bytes), vb);
        };
        Ids ahead = Ids();
        T ahead_w = T(1);
        if (task.y + lane < task.z) {
            ahead = load_ids(ids + task.y + lane);
            if (!UNIT) ahead_w = __ldg(a.w + task.y + lane);
        }
        for (int base = task.y; base < task.z; base += 32) {
            const int n = min(32, task.z - base);
            __syncwarp();
            s_edge[warp][lane] = ahead;
            if (!UNIT) s_w[warp][lane] = ahead_w;
            __syncwarp();
            if (base + 32 + lane < task.z) {   // next batch's edge ids travel while this batch is reduced
                ahead = load_ids(ids + base + 32 + lane);
                if (!UNIT) ahead_w = __ldg(a.w + base + 32 + lane);
            }
            // COUNT edges at once: all loads are issued before the first use; the ragged tail of a task (1..3 edges)
            // runs as one specialised batch instead of one memory round trip per edge
            auto run_batch = [&](auto count_tag, int u) {
                constexpr int COUNT = decltype(count_tag)::value;
                Vec<T, VEC> va[COUNT], vb[COUNT];
                T w[COUNT];
                Ids e[COUNT];
                bool same = CACHE;
#pragma unroll
                for (int q = 0; q < COUNT; ++q) {
                    e[q] = s_edge[warp][u + q];
                    w[q] = UNIT ? T(1) : s_w[warp][u + q];
                    same = same && second_id(e[q]) == cached_row;
                    gather(row_ptr<T>(A, first_id(e[q]), a_row_bytes), va[q]);
                }
                if (same) {   // warp-uniform
#pragma unroll
                    for (int q = 0; q < COUNT; ++q) accumulate(va[q], cached, w[q], base + u + q);
                } else {
#pragma unroll
                    for (int q = 0; q < COUNT; ++q) {
                        if (MSG != MSG_COPY) load_table(second_id(e[q]), vb[q]);
                        else vb[q] = va[q];
                    }
#pragma unroll
                    for (int q = 0; q < COUNT; ++q) accumulate(va[q], vb[q], w[q], base + u + q);
                    if (CACHE) {
                        cached = vb[COUNT - 1];
                        cached_row = second_id(e[COUNT - 1]);
                    }
                }
            };
            int u = 0;
            while (u < n) {
                if (GROUPED) {
                    while (base + u == next_boundary) flush();   // a group task: the next row(s) start here
                }
                const int stop = GROUPED ? min(n, next_boundary - base) : n;   // the current row's edges inside this batch
                for (; u + kUnroll <= stop; u += kUnroll) run_batch(std::integral_constant<int, kUnroll>(), u);
                switch (stop - u) {
                    case 1: run_batch(std::integral_constant<int, 1>(), u); break;
                    case 2: run_batch(std::integral_constant<int, 2>(), u); break;
                    case 3: run_batch(std::integral_constant<int, 3>(), u); break;
                    default: break;
                }
                u = stop;
            }
        }
    };
    if (a.w == nullptr || !(task.w & kNonUnitTask)) reduce_task(std::true_type());
    else reduce_task(std::false_type());
    if (GROUPED) {
        while (row_local < rows) flush();   // the last row, and empty rows that close a group
    } else {
        flush();
    }
}

// Gated variant for the min/max backward (reference all-ties rule: NaryMin/NaryMax::backward = (out == y)).
//   csc order: segment = src j, own row S = input[j], P = relation[edge.y] (table)  -> grad_input[j]
//   rel order: segment = rel k, own row S = relation[k], P = input[edge.y] (gather)  -> grad_relation[k]
// per edge (dst i = edge.x):  y = w * (P (x) S);  if (output[i] == y)  acc += grad_out[i] * w * (mul ? P : 1)
template <typename T> struct GatedArgs {
    const int4 *task;
    const int2 *edge;
    const unsigned *packed;
    int pack_shift;
    int keep;
    const T *w;
    const T *G;   // grad_output, gathered by edge.x
    const T *O;   // output, gathered by edge.x
    const T *P;   // gathered / table by edge.y
    const T *S;   // own row, by segment
    T *out;
    T *partial;
    long long dim;
    int n_task;
    int n_slab;
};

template <typename T, int VEC, int MSG, bool P_TABLE, bool PACKED, bool KEEP>
__global__ void __launch_bounds__(kThreadsPerBlock, 3) seg_gated_kernel(const GatedArgs<T> a) {
    using Ids = typename std::conditional<PACKED, unsigned, int2>::type;
    __shared__ __align__(16) Ids s_edge[kWarpsPerBlock][32];
    __shared__ __align__(16) T s_w[kWarpsPerBlock][32];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long gw = (long long)blockIdx.x * kWarpsPerBlock + warp;
    if (gw >= (long long)a.n_task * a.n_slab) return;
    const int slab = (int)(gw / a.n_task);
    const int4 task = __ldg(a.task + (gw - (long long)slab * a.n_task));
    const int slot = task_slot(task.w);
    const long long col = (long long)slab * (32 * VEC) + lane * VEC;
    const bool active = col < a.dim;
    const long long safe_col = active ? col : 0;
    const unsigned row_bytes = (unsigned)(a.dim * sizeof(T));
    const char *G = reinterpret_cast<const char *>(a.G + safe_col);
    const char *O = reinterpret_cast<const char *>(a.O + safe_col);
    const char *P = reinterpret_cast<const char *>(a.P + safe_col);
    const Ids *ids = reinterpret_cast<const Ids *>(PACKED ? (const void *)a.packed : (const void *)a.edge);
    const int shift = a.pack_shift;
    const unsigned low = PACKED ? (shift >= 32 ? 0xffffffffu : ((1u << shift) - 1u)) : 0u;
    const unsigned long long keep_policy = KEEP ? policy_evict_last() : 0, once_policy = KEEP ? policy_evict_first() : 0;
    auto gather = [&](const T *p, Vec<T, VEC> &v) {
        if (KEEP) gather_load_keep(p, v, keep_policy);
        else gather_load(p, v);
    };
    auto load_ids = [&](const Ids *p) { return KEEP ? edge_load_once(p, once_policy) : __ldg(p); };
    auto first_id = [&](const Ids &e) { return PACKED ? (int)(id_bits(e) & low) : id_x(e); };
    auto second_id = [&](const Ids &e) { return PACKED ? (int)(id_bits(e) >> shift) : id_y(e); };

    Vec<T, VEC> own;
    T acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { acc[v] = T(0); own.v[v] = T(0); }
    if (task.z > task.y) gather_load(a.S + (long long)task.x * a.dim + safe_col, own);

    auto reduce_task = [&](auto unit_tag) {
        constexpr bool UNIT = decltype(unit_tag)::value;
        Ids ahead = Ids();
        T ahead_w = T(1);
        if (task.y + lane < task.z) {
            ahead = load_ids(ids + task.y + lane);
            if (!UNIT) ahead_w = __ldg(a.w + task.y + lane);
        }
        for (int base = task.y; base < task.z; base += 32) {
            const int n = min(32, task.z - base);
            __syncwarp();
            s_edge[warp][lane] = ahead;
            if (!UNIT) s_w[warp][lane] = ahead_w;
            __syncwarp();
            if (base + 32 + lane < task.z) {
                ahead = load_ids(ids + base + 32 + lane);
                if (!UNIT) ahead_w = __ldg(a.w + base + 32 + lane);
            }
            auto run_batch = [&](auto count_tag, int u) {
                constexpr int COUNT = decltype(count_tag)::value;
                Vec<T, VEC> vg[COUNT], vo[COUNT], vp[COUNT];
                T w[COUNT];
#pragma unroll
                for (int q = 0; q < COUNT; ++q) {
                    const Ids e = s_edge[warp][u + q];
                    w[q] = UNIT ? T(1) : s_w[warp][u + q];
                    gather(row_ptr<T>(G, first_id(e), row_bytes), vg[q]);
                    gather(row_ptr<T>(O, first_id(e), row_bytes), vo[q]);
                    if (P_TABLE) table_load(row_ptr<T>(P, second_id(e), row_bytes), vp[q]);
                    else gather(row_ptr<T>(P, second_id(e), row_bytes), vp[q]);
                }
#pragma unroll
                for (int q = 0; q < COUNT; ++q) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        const T y = UNIT ? message<T, MSG>(vp[q].v[v], own.v[v]) : message<T, MSG>(w[q], vp[q].v[v], own.v[v]);
                        const T up = UNIT ? vg[q].v[v] : vg[q].v[v] * w[q];
                        const T term = MSG == MSG_MUL ? up * vp[q].v[v] : up;
                        if (vo[q].v[v] == y) acc[v] += term;
                    }
                }
            };
            int u = 0;
            for (; u + kUnroll <= n; u += kUnroll) run_batch(std::integral_constant<int, kUnroll>(), u);
            switch (n - u) {
                case 1: run_batch(std::integral_constant<int, 1>(), u); break;
                case 2: run_batch(std::integral_constant<int, 2>(), u); break;
                case 3: run_batch(std::integral_constant<int, 3>(), u); break;
                default: break;
            }
        }
    };
    if (a.w == nullptr || !(task.w & kNonUnitTask)) reduce_task(std::true_type());
    else reduce_task(std::false_type());

    if (!active) return;
    T *p = slot < 0 ? a.out + (long long)task.x * a.dim + col : a.partial + (long long)slot * a.dim + col;
    Vec<T, VEC> r;
#pragma unroll
    for (int v = 0; v < VEC; ++v) r.v[v] = acc[v];
    if (slot < 0) stream_store(p, r);
    else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) p[v] = r.v[v];
    }
}

// PNA aggregation in one pass (reference layer.py:141-144, 164-167, 343-346, 366-369 issue four operator calls over the
// same edges: add, add of squared operands, max, min).  One gather per edge feeds four accumulators:
//     sum += w (r (x) x)    sq += w (r^2 (x) x^2)    max / min of w (r (x) x)
// with the reference's rounding (the squares are formed first, as `relation ** 2` / `input ** 2` are in the reference).
// Outputs and partial rows are stacked [4][rows][dim] in the order sum, squares, max, min.
template <typename T> struct PnaArgs {
    const int4 *task;
    const int2 *edge;
    const unsigned *packed;
    int pack_shift;
    int keep;
    const T *w;
    const T *A;
    const T *B;
    T *out[4];
    T *partial;          // [4][n_slot][dim]
    long long dim;
    long long slot_stride;   // n_slot * dim
    int n_task;
    int n_slab;
};

template <typename T, int VEC, int MSG, bool PACKED, bool KEEP>
__global__ void __launch_bounds__(kThreadsPerBlock, 2) seg_pna_kernel(const PnaArgs<T> a) {
    using Ids = typename std::conditional<PACKED, unsigned, int2>::type;
    __shared__ __align__(16) Ids s_edge[kWarpsPerBlock][32];
    __shared__ __align__(16) T s_w[kWarpsPerBlock][32];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long gw = (long long)blockIdx.x * kWarpsPerBlock + warp;
    if (gw >= (long long)a.n_task * a.n_slab) return;
    const int slab = (int)(gw / a.n_task);
    const int4 task = __ldg(a.task + (gw - (long long)slab * a.n_task));
    const int slot = task_slot(task.w);
    const long long col = (long long)slab * (32 * VEC) + lane * VEC;
    const bool active = col < a.dim;
    const long long safe_col = active ? col : 0;
    const unsigned row_bytes = (unsigned)(a.dim * sizeof(T));
    const char *A = reinterpret_cast<const char *>(a.A + safe_col);
    const char *B = reinterpret_cast<const char *>(a.B + safe_col);
    const Ids *ids = reinterpret_cast<const Ids *>(PACKED ? (const void *)a.packed : (const void *)a.edge);
    const int shift = a.pack_shift;
    const unsigned low = PACKED ? (shift >= 32 ? 0xffffffffu : ((1u << shift) - 1u)) : 0u;
    const unsigned long long keep_policy = KEEP ? policy_evict_last() : 0, once_policy = KEEP ? policy_evict_first() : 0;
    auto gather = [&](const T *p, Vec<T, VEC> &v) {
        if (KEEP) gather_load_keep(p, v, keep_policy);
        else gather_load(p, v);
    };
    auto load_ids = [&](const Ids *p) { return KEEP ? edge_load_once(p, once_policy) : __ldg(p); };
    auto first_id = [&](const Ids &e) { return PACKED ? (int)(id_bits(e) & low) : id_x(e); };
    auto second_id = [&](const Ids &e) { return PACKED ? (int)(id_bits(e) >> shift) : id_y(e); };

    T acc_sum[VEC], acc_sq[VEC], acc_max[VEC], acc_min[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        acc_sum[v] = acc_sq[v] = T(0);
        acc_max[v] = Limits<T>::lowest();
        acc_min[v] = Limits<T>::highest();
    }
    const bool weighted = a.w != nullptr && (task.w & kNonUnitTask);
    Vec<T, VEC> cached;
    int cached_row = -1;
    Ids ahead = Ids();
    T ahead_w = T(1);
    if (task.y + lane < task.z) {
        ahead = load_ids(ids + task.y + lane);
        if (weighted) ahead_w = __ldg(a.w + task.y + lane);
    }
    for (int base = task.y; base < task.z; base += 32) {
        const int n = min(32, task.z - base);
        __syncwarp();
        s_edge[warp][lane] = ahead;
        s_w[warp][lane] = ahead_w;
        __syncwarp();
        if (base + 32 + lane < task.z) {
            ahead = load_ids(ids + base + 32 + lane);
            if (weighted) ahead_w = __ldg(a.w + base + 32 + lane);
        }
        auto run_batch = [&](auto count_tag, int u) {
            constexpr int COUNT = decltype(count_tag)::value;
            Vec<T, VEC> va[COUNT], vb[COUNT];
            T w[COUNT];
            Ids e[COUNT];
            bool same = true;
#pragma unroll
            for (int q = 0; q < COUNT; ++q) {
                e[q] = s_edge[warp][u + q];
                w[q] = s_w[warp][u + q];
                same = same && second_id(e[q]) == cached_row;
                gather(row_ptr<T>(A, first_id(e[q]), row_bytes), va[q]);
            }
            if (!same) {
#pragma unroll
                for (int q = 0; q < COUNT; ++q) table_load(row_ptr<T>(B, second_id(e[q]), row_bytes), vb[q]);
            }
#pragma unroll
            for (int q = 0; q < COUNT; ++q) {
                const Vec<T, VEC> &rel = same ? cached : vb[q];
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const T r = rel.v[v], x = va[q].v[v];
                    const T m = message<T, MSG>(w[q], r, x);                 // weight 1 multiplies exactly
                    const T msq = message<T, MSG>(w[q], r * r, x * x);
                    acc_sum[v] += m;
                    acc_sq[v] += msq;
                    acc_max[v] = acc_max[v] > m ? acc_max[v] : m;
                    acc_min[v] = acc_min[v] < m ? acc_min[v] : m;
                }
            }
            if (!same) {
                cached = vb[COUNT - 1];
                cached_row = second_id(e[COUNT - 1]);
            }
        };
        int u = 0;
        for (; u + kUnroll <= n; u += kUnroll) run_batch(std::integral_constant<int, kUnroll>(), u);
        switch (n - u) {
            case 1: run_batch(std::integral_constant<int, 1>(), u); break;
            case 2: run_batch(std::integral_constant<int, 2>(), u); break;
            case 3: run_batch(std::integral_constant<int, 3>(), u); break;
            default: break;
        }
    }
    if (!active) return;
    const T *results[4] = {acc_sum, acc_sq, acc_max, acc_min};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        Vec<T, VEC> r;
#pragma unroll
        for (int v = 0; v < VEC; ++v) r.v[v] = results[k][v];
        if (slot < 0) {
            stream_store(a.out[k] + (long long)task.x * a.dim + col, r);
        } else {
            T *p = a.partial + k * a.slot_stride + (long long)slot * a.dim + col;
#pragma unroll
            for (int v = 0; v < VEC; ++v) p[v] = r.v[v];
        }
    }
}

// Fixed-order fold of the partial rows of split segments (deterministic second stage).
template <typename T, int SUM, bool ARG>
__global__ void combine_kernel(const int4 *__restrict__ split, int n_split, const T *__restrict__ partial,
                               const int32_t *__restrict__ partial_arg, const T *__restrict__ addend,
                               T *__restrict__ out, int32_t *__restrict__ arg_out, long long dim, long long block,
                               int block_shift, long long o_stride, long long o_offset, long long o_row) {
    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= dim) return;
    const int4 s = __ldg(split + blockIdx.y);
    T acc = reduce_identity<T, SUM>();
    int32_t arg = -1;
    for (int q = 0; q < s.z; ++q) {
        const long long at = (long long)(s.y + q) * dim + col;
        const T m = partial[at];
        if (ARG) {
            const int32_t candidate = partial_arg[at];
            const bool better = SUM == ULTRA_RSPMM_SUM_MAX ? (m > acc) : (m < acc);
            if (better || (m == acc && candidate >= 0 && (arg < 0 || candidate < arg))) arg = candidate;
        }
        reduce_into<T, SUM>(acc, m);
    }
    if (addend) acc += addend[(long long)s.x * dim + col];
    if (block) out[(long long)s.x * o_row + blocked_col(col, (int)block, block_shift, o_stride) + o_offset] = acc;
    else out[(long long)s.x * dim + col] = acc;
    if (ARG) arg_out[(long long)s.x * dim + col] = arg;
}

// ------------------------------------------------------------------------------------------------
// host-side dispatch
// ------------------------------------------------------------------------------------------------
template <typename T> constexpr int wide_vec() { return 16 / sizeof(T); }
constexpr int FWD = ULTRA_RSPMM_PASS_FORWARD, GIN = ULTRA_RSPMM_PASS_GRAD_INPUT, GREL = ULTRA_RSPMM_PASS_GRAD_RELATION;

// Instantiation budget: the L2-hint (KEEP) and grouped-task variants exist only where they matter - float operands at the
// full slab width; group tasks never carry an arg-index.  args.keep / args.grouped are set accordingly by run_pass.
template <typename T, int VEC, int SUM, int MSG, bool B_TABLE, bool ARG>
int launch_seg(const SegArgs<T> &args, cudaStream_t stream) {
    const long long warps = (long long)args.n_task * args.n_slab;
    if (warps == 0) return ULTRA_RSPMM_OK;
    const long long blocks = (warps + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (blocks > 0x7fffffffLL || args.dim * (long long)sizeof(T) > 0xffffffffLL) return ULTRA_RSPMM_ERR_RANGE;
    constexpr bool TUNED = sizeof(T) == 4 && VEC == 4;
    const unsigned grid = (unsigned)blocks;
#define ULTRA_SEG(PACKED, KEEP, GROUPED) \
    seg_reduce_kernel<T, VEC, SUM, MSG, B_TABLE, ARG, PACKED, KEEP, GROUPED><<<grid, kThreadsPerBlock, 0, stream>>>(args)
    if (TUNED && args.grouped && !ARG) {
        if (args.keep) { if (args.packed) ULTRA_SEG(true, TUNED, (TUNED && !ARG)); else ULTRA_SEG(false, TUNED, (TUNED && !ARG)); }
        else { if (args.packed) ULTRA_SEG(true, false, (TUNED && !ARG)); else ULTRA_SEG(false, false, (TUNED && !ARG)); }
    } else if (TUNED && args.keep) {
        if (args.packed) ULTRA_SEG(true, TUNED, false); else ULTRA_SEG(false, TUNED, false);
    } else {
        if (args.packed) ULTRA_SEG(true, false, false); else ULTRA_SEG(false, false, false);
    }
#undef ULTRA_SEG
    note_launch();
    return ULTRA_RSPMM_OK;
}

struct BlockedLayout {
    long long block = 0, a_stride = 0, o_stride = 0, o_offset = 0;
    int shift() const { return block > 0 && (block & (block - 1)) == 0 ? __builtin_ctzll((unsigned long long)block) : -1; }
};

template <typename T, int SUM, bool ARG>
int launch_combine(const ultra_rspmm_order_t &order, const T *partial, const int32_t *partial_arg, T *out,
                   int32_t *arg_out, long long dim, cudaStream_t stream, const T *addend = nullptr,
                   const BlockedLayout &layout = BlockedLayout()) {
    const long long block = layout.block, o_stride = layout.o_stride, o_offset = layout.o_offset;
    const long long o_row = layout.block ? (dim / layout.block) * layout.o_stride : dim;
    const int block_shift = layout.shift();
    if (order.n_split == 0 || dim == 0) return ULTRA_RSPMM_OK;
    const dim3 grid((unsigned)((dim + 255) / 256), (unsigned)order.n_split);
    if (order.n_split > 65535) {
        // y-dimension limit: fold in slices
        for (int at = 0; at < order.n_split; at += 65535) {
            const int n = order.n_split - at < 65535 ? order.n_split - at : 65535;
            combine_kernel<T, SUM, ARG><<<dim3(grid.x, n), 256, 0, stream>>>((const int4 *)order.split + at, n, partial,
                                                                              partial_arg, addend, out, arg_out, dim, block,
                                                                              block_shift, o_stride, o_offset, o_row);
            note_launch();
        }
        return ULTRA_RSPMM_OK;
    }
    combine_kernel<T, SUM, ARG><<<grid, 256, 0, stream>>>((const int4 *)order.split, order.n_split, partial, partial_arg,
                                                          addend, out, arg_out, dim, block, block_shift, o_stride,
                                                          o_offset, o_row);
    note_launch();
    return ULTRA_RSPMM_OK;
}

// Slab width: the widest vector (16 / 8 / 4 bytes per lane) that the operands' shape and alignment allow and whose
// column block of the gathered operand (rows x 32 lanes x VEC) fits the L2 budget - a slab that does not stay in L2
// turns every gather into an HBM access (profiles/r01: C4 forward moved 19 GB of DRAM at 512 B slabs).
template <typename T>
int pick_vec(long long dim, long long rows, std::initializer_list<const void *> pointers) {
    int vec = wide_vec<T>();
    while (vec > 1) {
        bool ok = dim % vec == 0;
        for (const void *p : pointers) ok = ok && ((uintptr_t)p % (vec * sizeof(T)) == 0);
        if (ok) break;
        vec /= 2;
    }
    const int widest = vec;
    while (vec > 1 && rows * 32 * vec * (long long)sizeof(T) > g_l2_budget) vec /= 2;
    if (rows * 32 * vec * (long long)sizeof(T) > g_l2_budget) vec = widest;   // nothing fits: HBM-bound, widest rows are best
    return vec;
}

// one reduction pass + its combine
// bytes of the partial rows (+ partial arg-indices) of one pass; the work counter of the staged kernel sits right after
static size_t slot_bytes(const ultra_rspmm_order_t &order, long long dim, size_t elem, bool with_arg) {
    size_t bytes = align_up((size_t)order.n_slot * dim * elem);
    if (with_arg) bytes += align_up((size_t)order.n_slot * dim * sizeof(int32_t));
    return bytes;
}

// largest workspace the destination-blocked pass may ask for (partial rows: n_rel x n_block x dim floats)
static size_t blocked_cap() {
    static const size_t cap = (size_t)(getenv("ULTRA_RSPMM_BLOCKED_MAX_GB") ? atoll(getenv("ULTRA_RSPMM_BLOCKED_MAX_GB")) : 24) << 30;
    return cap;
}

// workspace of the destination-blocked grad_relation pass: n_rel x n_block partial rows + its work counter
static size_t blocked_bytes(const ultra_rspmm_index_t &ix, long long dim) {
    if (!ix.block_ptr || ix.dtype != ULTRA_RSPMM_F32) return 0;
    return align_up((size_t)ix.n_rel * ix.n_block * dim * sizeof(float)) + 256;
}

// When the blocked pass beats the generic one (configs[4] sweep, profiles/r02_sweep_c5.csv): it trades one row gather per
// edge for a partial row per (relation, block) run, so the runs must be long enough; DistMult saves the grad_output gather
// once the two slabs exceed L2 (slightly: runs >= 24 edges; far: runs >= 8); TransE gathers grad_output only, which the
// staged blocks replace by streaming loads - that pays once that one slab is far beyond L2.
static bool blocked_pays(const ultra_rspmm_index_t &ix, long long dim, int msg) {
    if (g_blocked == 2) return true;
    const double run = (double)ix.nnz / ((double)ix.n_rel * (ix.n_block > 0 ? ix.n_block : 1));
    if (msg == MSG_COPY) return (long long)ix.n_out * 512 > (150ll << 20) && run >= 8;
    const long long both = ((long long)ix.n_out + ix.n_in) * 512;
    return (both > (200ll << 20) && run >= 8) || (dim >= 512 && run >= 24);
}

// Few-row operands (the graph of relations): the gathered operand's 64-feature slab fits shared memory and every edge
// reads it from there (rspmm_staged.cu).  Worth it when the refills (one per CTA and slab change) are small against the
// edge work: at least ~16 edges per staged row and SM.
template <typename T, int SUM, int MSG, bool B_TABLE, bool ARG>
bool staged_applies(const ultra_rspmm_order_t &order, long long rows_gathered, long long dim, int vec, long long nnz,
                    const void *workspace, size_t workspace_bytes, size_t counter_at) {
    if (!std::is_same<T, float>::value || SUM != ULTRA_RSPMM_SUM_ADD || !B_TABLE || ARG) return false;
    if (g_staged == 0 || vec != 4 || order.pack_shift <= 0 || rows_gathered <= 0 || rows_gathered > kStagedMaxRows) return false;
    if (!workspace || workspace_bytes < counter_at + sizeof(unsigned)) return false;
    if (g_staged == 2) return true;
    const long long slabs = (dim + kStagedSlab - 1) / kStagedSlab;
    return nnz * slabs >= 16ll * 148 * rows_gathered;
}

template <typename T, int SUM, int MSG, bool B_TABLE, bool ARG>
int run_pass(int pass, const ultra_rspmm_index_t &ix, const ultra_rspmm_order_t &order, bool unit_weight, const T *A, const T *B,
             long long rows_gathered, T *out, int32_t *arg_out, long long dim, void *workspace, size_t workspace_bytes,
             cudaStream_t stream,
             const T *addend = nullptr, const BlockedLayout layout = BlockedLayout()) {
    SegArgs<T> args;
    args.ptr = order.ptr;
    args.task = (const int4 *)order.task;
    args.edge = (const int2 *)order.edge;
    args.packed = order.pack_shift > 0 ? (const unsigned *)order.packed : nullptr;
    args.pack_shift = order.pack_shift;
    args.eid = order.eid;
    args.w = unit_weight ? nullptr : (const T *)order.w;
    args.A = A;
    args.B = B;
    args.out = out;
    args.addend = addend;
    args.partial = (T *)workspace;
    args.arg_out = arg_out;
    args.partial_arg = ARG ? (int32_t *)((char *)workspace + align_up((size_t)order.n_slot * dim * sizeof(T))) : nullptr;
    args.dim = dim;
    args.n_task = order.n_task;
    args.block = (int)layout.block;
    args.block_shift = layout.shift();
    args.a_row = layout.block ? (dim / layout.block) * layout.a_stride : dim;
    args.o_row = layout.block ? (dim / layout.block) * layout.o_stride : dim;
    args.a_stride = layout.a_stride;
    args.o_stride = layout.o_stride;
    args.o_offset = layout.o_offset;
    const int vec = pick_vec<T>(dim, rows_gathered, {A, B, out, workspace, addend});
    args.n_slab = (int)((dim + 32 * vec - 1) / (32 * vec));
    ultra_rspmm_pass_info_t info = {};
    info.packed = order.pack_shift > 0;
    info.n_split = order.n_split;
    const size_t counter_at = slot_bytes(order, dim, sizeof(T), ARG);
    const bool counter_ok = workspace && workspace_bytes >= counter_at + sizeof(unsigned);
    // <= 4 relation types, few rows, unit weights, pair lists built (ultra_rspmm_index_extend): one row read per pair
    if (std::is_same<T, float>::value && SUM == ULTRA_RSPMM_SUM_ADD && B_TABLE && !ARG && pass != GREL && vec == 4 && g_pairs != 0 &&
        unit_weight && ix.unit_weight && ix.pairs[pass == GIN ? 1 : 0].n_pair > 0 && rows_gathered <= kStagedMaxRows && counter_ok) {
        const ultra_rspmm_pairs_t &pairs = ix.pairs[pass == GIN ? 1 : 0];
        PairArgs pa = {};
        pa.ptr = pairs.ptr; pa.pair = pairs.pair; pa.rows = pairs.rows; pa.id_bits = pairs.id_bits;
        pa.A = (const float *)A; pa.B = (const float *)B; pa.out = (float *)out; pa.addend = (const float *)addend;
        pa.counter = (unsigned *)((char *)workspace + counter_at);
        pa.dim = dim; pa.n_seg = order.n_seg; pa.n_rows = (int)rows_gathered; pa.n_rel = ix.n_rel;
        pa.a_stride = args.a_stride; pa.o_stride = args.o_stride; pa.o_offset = args.o_offset;
        pa.a_row = args.a_row; pa.o_row = args.o_row; pa.block = args.block; pa.block_shift = args.block_shift;
        if (int status = launch_pairs_in_smem(pa, MSG, stream)) return status;
        info.kernel = ULTRA_RSPMM_KERNEL_PAIRS_IN_SMEM;
        info.vec = 4;
        info.n_task = order.n_seg;
        info.n_slab = (int)((dim + kStagedSlab - 1) / kStagedSlab);
        info.n_split = 0;
        note_pass(pass, info);
        return ULTRA_RSPMM_OK;     // whole rows per warp: no partial rows, no combine
    }
    // grad_relation of graphs whose gathered slabs exceed L2: destination-blocked pass (3 row gathers per edge and step)
    if (std::is_same<T, float>::value && SUM == ULTRA_RSPMM_SUM_ADD && !ARG && pass == GREL && vec == 4 && g_blocked != 0 &&
        ix.block_ptr && !layout.block && !addend && workspace && blocked_pays(ix, dim, MSG) &&
        workspace_bytes >= blocked_bytes(ix, dim) && blocked_bytes(ix, dim) <= blocked_cap()) {
        BlockedRelArgs ba = {};
        ba.block_ptr = ix.block_ptr;
        ba.edge = (const int2 *)order.edge;
        ba.packed = order.pack_shift > 0 ? (const unsigned *)order.packed : nullptr;
        ba.pack_shift = order.pack_shift;
        ba.w = unit_weight ? nullptr : (const float *)order.w;
        ba.G = (const float *)A;
        ba.X = (const float *)B;
        ba.partial = (float *)workspace;
        ba.counter = (unsigned *)((char *)workspace + blocked_bytes(ix, dim) - 256);
        ba.dim = dim; ba.n_rel = ix.n_rel; ba.n_block = ix.n_block; ba.block_rows = ix.block_rows; ba.n_out = ix.n_out;
        if (int status = launch_dst_blocked(ba, MSG, stream)) return status;
        info.kernel = ULTRA_RSPMM_KERNEL_DST_BLOCKED;
        info.vec = 4;
        info.n_task = ix.n_rel * ix.n_block;
        info.n_slab = (int)((dim + kStagedSlab - 1) / kStagedSlab);
        info.n_split = ix.n_rel;
        note_pass(pass, info);
        ultra_rspmm_order_t folded = order;      // combine list: relation k <- its n_block partial rows, in block order
        folded.split = ix.block_split;
        folded.n_split = ix.n_rel;
        return launch_combine<T, SUM, ARG>(folded, (const T *)workspace, nullptr, out, nullptr, dim, stream);
    }
    // slab of the gathered operand far beyond L2: sub-warp rows kernel with a 256- or 128-byte slab that is L2-resident
    // again (rspmm_narrow.cu).  grad_relation with two gathered operands only when the narrow slabs of both fit.
    if (std::is_same<T, float>::value && SUM == ULTRA_RSPMM_SUM_ADD && !ARG && vec == 4 && !layout.block && g_narrow_bytes > 0 &&
        rows_gathered * 512 > g_narrow_bytes && (B_TABLE || rows_gathered * 128 <= (96ll << 20)) &&
        (order.n_slot == 0 || workspace)) {
        const int sub = g_narrow_sub ? g_narrow_sub : (rows_gathered * 256 <= (72ll << 20) ? 2 : 4);
        NarrowArgs na = {};
        na.task = (const int4 *)order.task;
        na.edge = (const int2 *)order.edge;
        na.packed = order.pack_shift > 0 ? (const unsigned *)order.packed : nullptr;
        na.pack_shift = order.pack_shift;
        na.w = unit_weight ? nullptr : (const float *)order.w;
        na.A = (const float *)A; na.B = (const float *)B; na.out = (float *)out; na.addend = (const float *)addend;
        na.partial = (float *)workspace;
        na.dim = dim;
        na.n_task = order.n_task;
        if (int status = launch_narrow(na, MSG, B_TABLE, sub, stream)) return status;
        info.kernel = ULTRA_RSPMM_KERNEL_SUBWARP_ROWS;
        info.vec = 4;
        info.keep = 1;
        info.n_task = order.n_task;
        info.n_slab = (int)((dim + 128 / sub - 1) / (128 / sub));
        note_pass(pass, info);
        return launch_combine<T, SUM, ARG>(order, args.partial, args.partial_arg, out, arg_out, dim, stream, addend, layout);
    }
    if (staged_applies<T, SUM, MSG, B_TABLE, ARG>(order, rows_gathered, dim, vec, ix.nnz, workspace, workspace_bytes, counter_at)) {
        StagedArgs staged = {};
        staged.task = (const int4 *)order.task;
        staged.packed = (const unsigned *)order.packed;
        staged.pack_shift = order.pack_shift;
        staged.w = unit_weight ? nullptr : (const float *)order.w;
        staged.A = (const float *)A;
        staged.B = (const float *)B;
        staged.out = (float *)out;
        staged.addend = (const float *)addend;
        staged.partial = (float *)workspace;
        staged.counter = (unsigned *)((char *)workspace + counter_at);
        staged.dim = dim;
        staged.n_task = order.n_task;
        staged.n_rows = (int)rows_gathered;
        staged.a_stride = args.a_stride; staged.o_stride = args.o_stride; staged.o_offset = args.o_offset;
        staged.a_row = args.a_row; staged.o_row = args.o_row;
        staged.block = args.block; staged.block_shift = args.block_shift;
        int status = launch_rows_in_smem(staged, MSG, stream);
        if (status) return status;
        info.kernel = ULTRA_RSPMM_KERNEL_ROWS_IN_SMEM;
        info.vec = 4;
        info.n_task = order.n_task;
        info.n_slab = (int)((dim + kStagedSlab - 1) / kStagedSlab);
        note_pass(pass, info);
        return launch_combine<T, SUM, ARG>(order, args.partial, args.partial_arg, out, arg_out, dim, stream, addend, layout);
    }
    const bool tuned = sizeof(T) == 4 && vec == 4;
    args.keep = tuned && (g_variant == 2 || (g_variant == 0 && rows_gathered * 32 * vec * (long long)sizeof(T) > (24ll << 20)));
    // short rows are walked several per warp: the grouped task list (built only for orders whose segments average
    // fewer than ~30 edges, see rspmm_index.cu)
    args.grouped = tuned && !ARG && order.n_gtask > 0;
    if (args.grouped) {
        args.task = (const int4 *)order.gtask;
        args.n_task = order.n_gtask;
    }
    int status;
    if (vec == 4) status = launch_seg<T, sizeof(T) == 4 ? 4 : 2, SUM, MSG, B_TABLE, ARG>(args, stream);
    else if (vec == 2) status = launch_seg<T, 2, SUM, MSG, B_TABLE, ARG>(args, stream);
    else status = launch_seg<T, 1, SUM, MSG, B_TABLE, ARG>(args, stream);
    if (status) return status;
    info.kernel = ULTRA_RSPMM_KERNEL_SEG_REDUCE;
    info.vec = vec;
    info.keep = args.keep;
    info.grouped = args.grouped;
    info.n_task = args.n_task;
    info.n_slab = args.n_slab;
    note_pass(pass, info);
    return launch_combine<T, SUM, ARG>(order, args.partial, args.partial_arg, out, arg_out, dim, stream, addend,
                                       layout);
}

template <typename T, int VEC, int MSG, bool P_TABLE>
int launch_gated(const GatedArgs<T> &args, cudaStream_t stream) {
    const long long warps = (long long)args.n_task * args.n_slab;
    if (warps == 0) return ULTRA_RSPMM_OK;
    const long long blocks = (warps + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (blocks > 0x7fffffffLL || args.dim * (long long)sizeof(T) > 0xffffffffLL) return ULTRA_RSPMM_ERR_RANGE;
    if (args.keep) {
        if (args.packed) seg_gated_kernel<T, VEC, MSG, P_TABLE, true, true><<<(unsigned)blocks, kThreadsPerBlock, 0, stream>>>(args);
        else seg_gated_kernel<T, VEC, MSG, P_TABLE, false, true><<<(unsigned)blocks, kThreadsPerBlock, 0, stream>>>(args);
    } else {
        if (args.packed) seg_gated_kernel<T, VEC, MSG, P_TABLE, true, false><<<(unsigned)blocks, kThreadsPerBlock, 0, stream>>>(args);
        else seg_gated_kernel<T, VEC, MSG, P_TABLE, false, false><<<(unsigned)blocks, kThreadsPerBlock, 0, stream>>>(args);
    }
    note_launch();
    return ULTRA_RSPMM_OK;
}

template <typename T, int MSG, bool P_TABLE>
int run_gated(int pass, const ultra_rspmm_order_t &order, bool unit_weight, const T *G, const T *O, const T *P, const T *S,
              long long rows_gathered, T *out, long long dim, void *workspace, size_t, cudaStream_t stream) {
    GatedArgs<T> args;
    args.task = (const int4 *)order.task;
    args.edge = (const int2 *)order.edge;
    args.packed = order.pack_shift > 0 ? (const unsigned *)order.packed : nullptr;
    args.pack_shift = order.pack_shift;
    args.w = unit_weight ? nullptr : (const T *)order.w;
    args.G = G; args.O = O; args.P = P; args.S = S;
    args.out = out;
    args.partial = (T *)workspace;
    args.dim = dim;
    args.n_task = order.n_task;
    const int vec = pick_vec<T>(dim, 2 * rows_gathered, {G, O, P, S, out, workspace});   // two gathered operands share L2
    args.n_slab = (int)((dim + 32 * vec - 1) / (32 * vec));
    args.keep = g_variant == 2 || (g_variant == 0 && 2 * rows_gathered * 32 * vec * (long long)sizeof(T) > (24ll << 20));
    int status;
    if (vec == 4) status = launch_gated<T, sizeof(T) == 4 ? 4 : 2, MSG, P_TABLE>(args, stream);
    else if (vec == 2) status = launch_gated<T, 2, MSG, P_TABLE>(args, stream);
    else status = launch_gated<T, 1, MSG, P_TABLE>(args, stream);
    if (status) return status;
    ultra_rspmm_pass_info_t info = {};
    info.kernel = ULTRA_RSPMM_KERNEL_SEG_GATED;
    info.vec = vec;
    info.keep = args.keep;
    info.packed = order.pack_shift > 0;
    info.n_task = args.n_task;
    info.n_slab = args.n_slab;
    info.n_split = order.n_split;
    note_pass(pass, info);
    return launch_combine<T, ULTRA_RSPMM_SUM_ADD, false>(order, args.partial, nullptr, out, nullptr, dim, stream);
}

// grad_relation of min / max on graphs whose slabs exceed L2: destination-blocked gated pass (one gathered row per edge
// instead of three).  Returns true when it ran (status in *status).
template <typename T, int MSG>
bool run_gated_blocked(const ultra_rspmm_index_t &ix, bool unit_weight, const T *G, const T *O, const T *X, const T *R, T *out,
                       long long dim, void *workspace, size_t workspace_bytes, cudaStream_t stream, int *status) {
    if (!std::is_same<T, float>::value || g_blocked == 0 || !ix.block_ptr || !workspace || ix.block_rows % 2) return false;
    if (workspace_bytes < blocked_bytes(ix, dim) || blocked_bytes(ix, dim) > blocked_cap()) return false;
    if (pick_vec<T>(dim, (long long)ix.n_out + ix.n_in, {G, O, X, R, out, workspace}) != 4) return false;
    const double run = (double)ix.nnz / ((double)ix.n_rel * (ix.n_block > 0 ? ix.n_block : 1));
    const long long both = ((long long)ix.n_out + ix.n_in) * 512;
    // a run is walked in two halves: below ~24 edges per run the per-run overhead eats the saved gathers (R' = 2,000 at
    // 8.4 M edges, 12 edges per run: 62-68 % of the HBM peak fwd+bwd generic, 58-60 % blocked)
    if (g_blocked != 2 && !(both > (96ll << 20) && run >= 24)) return false;
    BlockedRelArgs ba = {};
    ba.block_ptr = ix.block_ptr;
    ba.edge = (const int2 *)ix.rel.edge;
    ba.packed = ix.rel.pack_shift > 0 ? (const unsigned *)ix.rel.packed : nullptr;
    ba.pack_shift = ix.rel.pack_shift;
    ba.w = unit_weight ? nullptr : (const float *)ix.rel.w;
    ba.G = (const float *)G; ba.O = (const float *)O; ba.X = (const float *)X; ba.R = (const float *)R;
    ba.partial = (float *)workspace;
    ba.counter = (unsigned *)((char *)workspace + blocked_bytes(ix, dim) - 256);
    ba.dim = dim; ba.n_rel = ix.n_rel; ba.n_block = ix.n_block; ba.block_rows = ix.block_rows; ba.n_out = ix.n_out;
    *status = launch_dst_blocked_gated(ba, MSG, stream);
    if (*status) return true;
    ultra_rspmm_pass_info_t info = {};
    info.kernel = ULTRA_RSPMM_KERNEL_DST_BLOCKED_GATED;
    info.vec = 4;
    info.packed = ix.rel.pack_shift > 0;
    info.n_task = ix.n_rel * ix.n_block;
    info.n_slab = (int)((dim + kStagedSlab - 1) / kStagedSlab);
    info.n_split = ix.n_rel;
    note_pass(GREL, info);
    ultra_rspmm_order_t folded = ix.rel;         // combine list: relation k <- its n_block partial rows, in block order
    folded.split = ix.block_split;
    folded.n_split = ix.n_rel;
    *status = launch_combine<T, ULTRA_RSPMM_SUM_ADD, false>(folded, (const T *)workspace, nullptr, out, nullptr, dim, stream);
    return true;
}

template <typename T, int SUM>
int forward_sum(const ultra_rspmm_index_t &ix, const T *relation, const T *input, T *output, int32_t *argidx,
                long long dim, int mul_op, void *ws, size_t ws_bytes, cudaStream_t stream, const T *addend = nullptr) {
    const bool unit = ix.unit_weight != 0;
    if (SUM != ULTRA_RSPMM_SUM_ADD && argidx) {
        if (mul_op == ULTRA_RSPMM_MUL_MUL)
            return run_pass<T, SUM, MSG_MUL, true, true>(FWD, ix, ix.csr, unit, input, relation, ix.n_in, output, argidx, dim, ws, ws_bytes, stream);
        return run_pass<T, SUM, MSG_ADD, true, true>(FWD, ix, ix.csr, unit, input, relation, ix.n_in, output, argidx, dim, ws, ws_bytes, stream);
    }
    if (mul_op == ULTRA_RSPMM_MUL_MUL)
        return run_pass<T, SUM, MSG_MUL, true, false>(FWD, ix, ix.csr, unit, input, relation, ix.n_in, output, nullptr, dim, ws, ws_bytes, stream, addend);
    return run_pass<T, SUM, MSG_ADD, true, false>(FWD, ix, ix.csr, unit, input, relation, ix.n_in, output, nullptr, dim, ws, ws_bytes, stream, addend);
}

template <typename T>
int forward_typed(const ultra_rspmm_index_t &ix, const void *relation, const void *input, const void *addend, void *output,
                  int32_t *argidx, long long dim, int sum_op, int mul_op, void *ws, size_t ws_bytes, cudaStream_t stream) {
    const T *r = (const T *)relation, *x = (const T *)input;
    T *o = (T *)output;
    switch (sum_op) {
        case ULTRA_RSPMM_SUM_ADD:
            return forward_sum<T, ULTRA_RSPMM_SUM_ADD>(ix, r, x, o, nullptr, dim, mul_op, ws, ws_bytes, stream, (const T *)addend);
        case ULTRA_RSPMM_SUM_MIN: return forward_sum<T, ULTRA_RSPMM_SUM_MIN>(ix, r, x, o, argidx, dim, mul_op, ws, ws_bytes, stream);
        default: return forward_sum<T, ULTRA_RSPMM_SUM_MAX>(ix, r, x, o, argidx, dim, mul_op, ws, ws_bytes, stream);
    }
}

template <typename T>
int backward_typed(const ultra_rspmm_index_t &ix, const void *relation, const void *input, const void *output,
                   const void *grad_output, void *grad_relation, void *grad_input, long long dim, int sum_op,
                   int mul_op, void *ws, size_t ws_bytes, cudaStream_t stream, const void *grad_input_addend = nullptr) {
    const T *r = (const T *)relation, *x = (const T *)input, *o = (const T *)output, *g = (const T *)grad_output;
    T *gr = (T *)grad_relation, *gx = (T *)grad_input;
    const bool unit = ix.unit_weight != 0;
    constexpr int ADD = ULTRA_RSPMM_SUM_ADD;
    int status = ULTRA_RSPMM_OK;
    if (sum_op == ULTRA_RSPMM_SUM_ADD) {
        if (gx) {
            status = mul_op == ULTRA_RSPMM_MUL_MUL
                         ? run_pass<T, ADD, MSG_MUL, true, false>(GIN, ix, ix.csc, unit, g, r, ix.n_out, gx, nullptr, dim, ws, ws_bytes, stream, (const T *)grad_input_addend)
                         : run_pass<T, ADD, MSG_COPY, true, false>(GIN, ix, ix.csc, unit, g, r, ix.n_out, gx, nullptr, dim, ws, ws_bytes, stream, (const T *)grad_input_addend);
            if (status) return status;
        }
        if (gr) {
            status = mul_op == ULTRA_RSPMM_MUL_MUL
                         ? run_pass<T, ADD, MSG_MUL, false, false>(GREL, ix, ix.rel, unit, g, x, (long long)ix.n_out + ix.n_in, gr, nullptr, dim, ws, ws_bytes, stream)
                         : run_pass<T, ADD, MSG_COPY, true, false>(GREL, ix, ix.rel, unit, g, x, ix.n_out, gr, nullptr, dim, ws, ws_bytes, stream);
        }
        return status;
    }
    if (gx) {
        status = mul_op == ULTRA_RSPMM_MUL_MUL ? run_gated<T, MSG_MUL, true>(GIN, ix.csc, unit, g, o, r, x, ix.n_out, gx, dim, ws, ws_bytes, stream)
                                               : run_gated<T, MSG_ADD, true>(GIN, ix.csc, unit, g, o, r, x, ix.n_out, gx, dim, ws, ws_bytes, stream);
        if (status) return status;
    }
    if (gr) {
        const bool blocked = mul_op == ULTRA_RSPMM_MUL_MUL
                                 ? run_gated_blocked<T, MSG_MUL>(ix, unit, g, o, x, r, gr, dim, ws, ws_bytes, stream, &status)
                                 : run_gated_blocked<T, MSG_ADD>(ix, unit, g, o, x, r, gr, dim, ws, ws_bytes, stream, &status);
        if (blocked) return status;
        status = mul_op == ULTRA_RSPMM_MUL_MUL ? run_gated<T, MSG_MUL, false>(GREL, ix.rel, unit, g, o, x, r, (long long)ix.n_out + ix.n_in, gr, dim, ws, ws_bytes, stream)
                                               : run_gated<T, MSG_ADD, false>(GREL, ix.rel, unit, g, o, x, r, (long long)ix.n_out + ix.n_in, gr, dim, ws, ws_bytes, stream);
    }
    return status;
}


template <typename T, int VEC, int MSG>
int launch_pna(const PnaArgs<T> &args, cudaStream_t stream) {
    const long long warps = (long long)args.n_task * args.n_slab;
    if (warps == 0) return ULTRA_RSPMM_OK;
    const long long blocks = (warps + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (blocks > 0x7fffffffLL || args.dim * (long long)sizeof(T) > 0xffffffffLL) return ULTRA_RSPMM_ERR_RANGE;
    if (args.keep) {
        if (args.packed) seg_pna_kernel<T, VEC, MSG, true, true><<<(unsigned)blocks, kThreadsPerBlock, 0, stream>>>(args);
        else seg_pna_kernel<T, VEC, MSG, false, true><<<(unsigned)blocks, kThreadsPerBlock, 0, stream>>>(args);
    } else {
        if (args.packed) seg_pna_kernel<T, VEC, MSG, true, false><<<(unsigned)blocks, kThreadsPerBlock, 0, stream>>>(args);
        else seg_pna_kernel<T, VEC, MSG, false, false><<<(unsigned)blocks, kThreadsPerBlock, 0, stream>>>(args);
    }
    note_launch();
    return ULTRA_RSPMM_OK;
}

template <typename T, int MSG>
int forward_pna_typed(const ultra_rspmm_index_t &ix, const T *relation, const T *input, T *const out[4], long long dim,
                      void *workspace, cudaStream_t stream) {
    const ultra_rspmm_order_t &order = ix.csr;
    PnaArgs<T> args;
    args.task = (const int4 *)order.task;
    args.edge = (const int2 *)order.edge;
    args.packed = order.pack_shift > 0 ? (const unsigned *)order.packed : nullptr;
    args.pack_shift = order.pack_shift;
    args.w = ix.unit_weight ? nullptr : (const T *)order.w;
    args.A = input;
    args.B = relation;
    for (int k = 0; k < 4; ++k) args.out[k] = out[k];
    args.partial = (T *)workspace;
    args.dim = dim;
    args.slot_stride = (long long)order.n_slot * dim;
    args.n_task = order.n_task;
    const int vec = pick_vec<T>(dim, ix.n_in, {input, relation, out[0], out[1], out[2], out[3], workspace});
    args.n_slab = (int)((dim + 32 * vec - 1) / (32 * vec));
    args.keep = g_variant == 2 || (g_variant == 0 && (long long)ix.n_in * 32 * vec * (long long)sizeof(T) > (24ll << 20));
    int status;
    if (vec == 4) status = launch_pna<T, sizeof(T) == 4 ? 4 : 2, MSG>(args, stream);
    else if (vec == 2) status = launch_pna<T, 2, MSG>(args, stream);
    else status = launch_pna<T, 1, MSG>(args, stream);
    if (status) return status;
    ultra_rspmm_pass_info_t info = {};
    info.kernel = ULTRA_RSPMM_KERNEL_SEG_PNA;
    info.vec = vec;
    info.keep = args.keep;
    info.packed = order.pack_shift > 0;
    info.n_task = args.n_task;
    info.n_slab = args.n_slab;
    info.n_split = order.n_split;
    note_pass(FWD, info);
    const T *partial = (const T *)workspace;
    const long long stride = args.slot_stride;
    if ((status = launch_combine<T, ULTRA_RSPMM_SUM_ADD, false>(order, partial, nullptr, out[0], nullptr, dim, stream))) return status;
    if ((status = launch_combine<T, ULTRA_RSPMM_SUM_ADD, false>(order, partial + stride, nullptr, out[1], nullptr, dim, stream))) return status;
    if ((status = launch_combine<T, ULTRA_RSPMM_SUM_MAX, false>(order, partial + 2 * stride, nullptr, out[2], nullptr, dim, stream))) return status;
    return launch_combine<T, ULTRA_RSPMM_SUM_MIN, false>(order, partial + 3 * stride, nullptr, out[3], nullptr, dim, stream);
}

int check_call(const ultra_rspmm_index_t *index, int64_t dim, int32_t dtype, int32_t sum_op, int32_t mul_op) {
    if (!index || dim < 0) return ULTRA_RSPMM_ERR_ARG;
    if (dtype != ULTRA_RSPMM_F32 && dtype != ULTRA_RSPMM_F64) return ULTRA_RSPMM_ERR_ARG;
    if (dtype != index->dtype) return ULTRA_RSPMM_ERR_DTYPE;
    if (sum_op < 0 || sum_op > 2 || mul_op < 0 || mul_op > 1) return ULTRA_RSPMM_ERR_ARG;
    return ULTRA_RSPMM_OK;
}

size_t pass_bytes(const ultra_rspmm_order_t &order, int64_t dim, size_t elem, bool with_arg) {
    return slot_bytes(order, dim, elem, with_arg) + 256;   // + the work counter of the staged kernel
}

}  // namespace

}  // namespace ultra

using namespace ultra;

extern "C" int ultra_rspmm_workspace_bytes(const ultra_rspmm_index_t *index, int64_t dim, int32_t dtype,
                                           size_t *forward_bytes, size_t *backward_bytes) {
    if (!index || dim < 0 || (dtype != ULTRA_RSPMM_F32 && dtype != ULTRA_RSPMM_F64)) return ULTRA_RSPMM_ERR_ARG;
    const size_t elem = dtype == ULTRA_RSPMM_F32 ? 4 : 8;
    if (forward_bytes) *forward_bytes = pass_bytes(index->csr, dim, elem, true);
    if (backward_bytes) {
        const size_t a = pass_bytes(index->csc, dim, elem, false), b = pass_bytes(index->rel, dim, elem, false);
        const size_t c = dtype == ULTRA_RSPMM_F32 && blocked_bytes(*index, dim) <= blocked_cap() ? blocked_bytes(*index, dim) : 0;
        *backward_bytes = a > b ? (a > c ? a : c) : (b > c ? b : c);
    }
    return ULTRA_RSPMM_OK;
}

extern "C" int ultra_rspmm_forward(const ultra_rspmm_index_t *index, const void *dev_relation, const void *dev_input,
                                   const void *dev_addend, void *dev_output, int32_t *dev_argidx, int64_t dim,
                                   int32_t dtype, int32_t sum_op, int32_t mul_op, void *workspace, size_t workspace_bytes,
                                   void *stream) {
    int status = check_call(index, dim, dtype, sum_op, mul_op);
    if (status) return status;
    if (dev_addend && sum_op != ULTRA_RSPMM_SUM_ADD) return ULTRA_RSPMM_ERR_ARG;
    if (index->n_out == 0 || dim == 0) return ULTRA_RSPMM_OK;
    if (!dev_output || ((index->n_rel > 0 && !dev_relation) || (index->n_in > 0 && !dev_input)))
        return ULTRA_RSPMM_ERR_ARG;
    const bool with_arg = dev_argidx && sum_op != ULTRA_RSPMM_SUM_ADD;
    const size_t need = pass_bytes(index->csr, dim, dtype == ULTRA_RSPMM_F32 ? 4 : 8, with_arg);
    if (index->csr.n_slot > 0 && (!workspace || workspace_bytes < need)) return ULTRA_RSPMM_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    status = dtype == ULTRA_RSPMM_F32
                 ? forward_typed<float>(*index, dev_relation, dev_input, dev_addend, dev_output, dev_argidx, dim, sum_op, mul_op, workspace, workspace_bytes, s)
                 : forward_typed<double>(*index, dev_relation, dev_input, dev_addend, dev_output, dev_argidx, dim, sum_op, mul_op, workspace, workspace_bytes, s);
    if (status) return status;
    ULTRA_CUDA_OK(cudaGetLastError());
    return ULTRA_RSPMM_OK;
}

static int backward_call(const ultra_rspmm_index_t *index, const void *dev_relation, const void *dev_input,
                         const void *dev_output, const void *dev_grad_output, void *dev_grad_relation, void *dev_grad_input,
                         const void *dev_grad_input_addend, int64_t dim, int32_t dtype, int32_t sum_op, int32_t mul_op,
                         void *workspace, size_t workspace_bytes, void *stream) {
    int status = check_call(index, dim, dtype, sum_op, mul_op);
    if (status) return status;
    if (dim == 0 || (!dev_grad_relation && !dev_grad_input)) return ULTRA_RSPMM_OK;
    if (index->n_out > 0 && !dev_grad_output) return ULTRA_RSPMM_ERR_ARG;
    if (sum_op != ULTRA_RSPMM_SUM_ADD && index->n_out > 0 && !dev_output) return ULTRA_RSPMM_ERR_ARG;
    if ((index->n_rel > 0 && !dev_relation) || (index->n_in > 0 && !dev_input)) return ULTRA_RSPMM_ERR_ARG;
    if (dev_grad_input_addend && (sum_op != ULTRA_RSPMM_SUM_ADD || !dev_grad_input || dev_grad_input_addend == dev_grad_input))
        return ULTRA_RSPMM_ERR_ARG;
    const size_t elem = dtype == ULTRA_RSPMM_F32 ? 4 : 8;
    size_t need = 0;
    if (dev_grad_input) need = pass_bytes(index->csc, dim, elem, false);
    if (dev_grad_relation) {
        const size_t b = pass_bytes(index->rel, dim, elem, false);
        need = b > need ? b : need;
    }
    const bool uses_slots = (dev_grad_input && index->csc.n_slot > 0) || (dev_grad_relation && index->rel.n_slot > 0);
    if (uses_slots && (!workspace || workspace_bytes < need)) return ULTRA_RSPMM_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    status = dtype == ULTRA_RSPMM_F32
                 ? backward_typed<float>(*index, dev_relation, dev_input, dev_output, dev_grad_output, dev_grad_relation,
                                         dev_grad_input, dim, sum_op, mul_op, workspace, workspace_bytes, s, dev_grad_input_addend)
                 : backward_typed<double>(*index, dev_relation, dev_input, dev_output, dev_grad_output, dev_grad_relation,
                                          dev_grad_input, dim, sum_op, mul_op, workspace, workspace_bytes, s, dev_grad_input_addend);
    if (status) return status;
    ULTRA_CUDA_OK(cudaGetLastError());
    return ULTRA_RSPMM_OK;
}

extern "C" int ultra_rspmm_backward(const ultra_rspmm_index_t *index, const void *dev_relation, const void *dev_input,
                                    const void *dev_output, const void *dev_grad_output, void *dev_grad_relation,
                                    void *dev_grad_input, int64_t dim, int32_t dtype, int32_t sum_op, int32_t mul_op,
                                    void *workspace, size_t workspace_bytes, void *stream) {
    return backward_call(index, dev_relation, dev_input, dev_output, dev_grad_output, dev_grad_relation, dev_grad_input, nullptr,
                         dim, dtype, sum_op, mul_op, workspace, workspace_bytes, stream);
}

extern "C" int ultra_rspmm_backward_addend(const ultra_rspmm_index_t *index, const void *dev_relation, const void *dev_input,
                                           const void *dev_grad_output, void *dev_grad_relation, void *dev_grad_input,
                                           const void *dev_grad_input_addend, int64_t dim, int32_t dtype, int32_t mul_op,
                                           void *workspace, size_t workspace_bytes, void *stream) {
    return backward_call(index, dev_relation, dev_input, nullptr, dev_grad_output, dev_grad_relation, dev_grad_input,
                         dev_grad_input_addend, dim, dtype, ULTRA_RSPMM_SUM_ADD, mul_op, workspace, workspace_bytes, stream);
}

extern "C" int ultra_rspmm_forward_pna(const ultra_rspmm_index_t *index, const void *dev_relation, const void *dev_input,
                                       void *dev_sum, void *dev_square_sum, void *dev_max, void *dev_min, int64_t dim,
                                       int32_t dtype, int32_t mul_op, void *workspace, size_t workspace_bytes, void *stream) {
    int status = check_call(index, dim, dtype, ULTRA_RSPMM_SUM_ADD, mul_op);
    if (status) return status;
    if (index->n_out == 0 || dim == 0) return ULTRA_RSPMM_OK;
    if (!dev_sum || !dev_square_sum || !dev_max || !dev_min) return ULTRA_RSPMM_ERR_ARG;
    if ((index->n_rel > 0 && !dev_relation) || (index->n_in > 0 && !dev_input)) return ULTRA_RSPMM_ERR_ARG;
    const size_t elem = dtype == ULTRA_RSPMM_F32 ? 4 : 8;
    const size_t need = 4 * (size_t)index->csr.n_slot * dim * elem;
    if (index->csr.n_slot > 0 && (!workspace || workspace_bytes < need)) return ULTRA_RSPMM_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == ULTRA_RSPMM_F32) {
        float *const out[4] = {(float *)dev_sum, (float *)dev_square_sum, (float *)dev_max, (float *)dev_min};
        status = mul_op == ULTRA_RSPMM_MUL_MUL
                     ? forward_pna_typed<float, MSG_MUL>(*index, (const float *)dev_relation, (const float *)dev_input, out, dim, workspace, s)
                     : forward_pna_typed<float, MSG_ADD>(*index, (const float *)dev_relation, (const float *)dev_input, out, dim, workspace, s);
    } else {
        double *const out[4] = {(double *)dev_sum, (double *)dev_square_sum, (double *)dev_max, (double *)dev_min};
        status = mul_op == ULTRA_RSPMM_MUL_MUL
                     ? forward_pna_typed<double, MSG_MUL>(*index, (const double *)dev_relation, (const double *)dev_input, out, dim, workspace, s)
                     : forward_pna_typed<double, MSG_ADD>(*index, (const double *)dev_relation, (const double *)dev_input, out, dim, workspace, s);
    }
    if (status) return status;
    ULTRA_CUDA_OK(cudaGetLastError());
    return ULTRA_RSPMM_OK;
}

extern "C" int ultra_rspmm_forward_blocked(const ultra_rspmm_index_t *index, const void *dev_relation, const void *dev_input,
                                           const void *dev_addend, void *dev_output, int64_t dim, int32_t dtype,
                                           int32_t mul_op, int64_t block, int64_t input_block_stride,
                                           int64_t output_block_stride, int64_t output_block_offset, void *workspace,
                                           size_t workspace_bytes, void *stream) {
    int status = check_call(index, dim, dtype, ULTRA_RSPMM_SUM_ADD, mul_op);
    if (status) return status;
    if (dtype != ULTRA_RSPMM_F32) return ULTRA_RSPMM_ERR_DTYPE;
    if (block <= 0 || dim % block || block % 4 || input_block_stride < block || output_block_stride < block ||
        input_block_stride % 4 || output_block_stride % 4 || output_block_offset % 4 || output_block_offset < 0 ||
        output_block_offset + block > output_block_stride)
        return ULTRA_RSPMM_ERR_ARG;
    if (index->n_out == 0 || dim == 0) return ULTRA_RSPMM_OK;
    if (!dev_output || !dev_relation || !dev_input) return ULTRA_RSPMM_ERR_ARG;
    if ((dim / block) * input_block_stride * 4 > 0xffffffffLL) return ULTRA_RSPMM_ERR_RANGE;
    const size_t need = pass_bytes(index->csr, dim, 4, false);
    if (index->csr.n_slot > 0 && (!workspace || workspace_bytes < need)) return ULTRA_RSPMM_ERR_WORKSPACE;
    BlockedLayout layout;
    layout.block = block;
    layout.a_stride = input_block_stride;
    layout.o_stride = output_block_stride;
    layout.o_offset = output_block_offset;
    const bool unit = index->unit_weight != 0;
    cudaStream_t s = (cudaStream_t)stream;
    const float *r = (const float *)dev_relation, *x = (const float *)dev_input, *b = (const float *)dev_addend;
    float *o = (float *)dev_output;
    constexpr int ADD = ULTRA_RSPMM_SUM_ADD;
    status = mul_op == ULTRA_RSPMM_MUL_MUL
                 ? run_pass<float, ADD, MSG_MUL, true, false>(FWD, *index, index->csr, unit, x, r, index->n_in, o, nullptr, dim, workspace, workspace_bytes, s, b, layout)
                 : run_pass<float, ADD, MSG_ADD, true, false>(FWD, *index, index->csr, unit, x, r, index->n_in, o, nullptr, dim, workspace, workspace_bytes, s, b, layout);
    if (status) return status;
    ULTRA_CUDA_OK(cudaGetLastError());
    return ULTRA_RSPMM_OK;
}
