// Shared device helpers for the rspmm kernels (sm_100a).
//
// Data layout in HBM (DESIGN.md "Data layout"): dense operands are row-major (rows, dim) with the
// query batch folded into the feature axis (feature = batch * 64 + channel, reference layer.py:118,306);
// a "slab" is a contiguous run of SLAB = LPE * VEC features of every row.  Kernels walk slabs in
// slab-major block order so that the co-resident CTAs share one N x SLAB column block of the gathered
// operand (L2-resident) and one R' x SLAB column block of the relation table (L1/L2-resident).
#pragma once

#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "ultra_rspmm.h"

namespace ultra {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreadsPerBlock = kWarpsPerBlock * 32;
constexpr unsigned kFullMask = 0xffffffffu;

enum MsgKind { MSG_MUL = 0, MSG_ADD = 1, MSG_COPY = 2 };  // rel*x, rel+x, x (relation not read)

template <typename T> struct Limits;
template <> struct Limits<float> {
    static __host__ __device__ constexpr float lowest() { return -FLT_MAX; }
    static __host__ __device__ constexpr float highest() { return FLT_MAX; }
};
template <> struct Limits<double> {
    static __host__ __device__ constexpr double lowest() { return -DBL_MAX; }
    static __host__ __device__ constexpr double highest() { return DBL_MAX; }
};

template <typename T, int SUM> __device__ __forceinline__ T reduce_identity() {
    return SUM == ULTRA_RSPMM_SUM_ADD ? T(0)
                                      : (SUM == ULTRA_RSPMM_SUM_MAX ? Limits<T>::lowest() : Limits<T>::highest());
}

// ---------------------------------------------------------------------------------------------
// 128-bit (or scalar) global loads with explicit cache policy.
//   gather_*     : rows of the gathered operand (input / grad_output / output) when its column slab is small or
//                  mid-sized: read-only path with L1 allocation.  Reuse is mainly through L2 (slab-major order), but on
//                  graphs with few nodes (the relation graph: 474 rows x 512 B = 243 KB per slab) a large share of the
//                  gathers hits L1 and L2->SM bandwidth stops being the bound (profiles/README.md: -10..16 %);
//   gather_*_keep: the same rows when the slab approaches the L2 capacity: no L1 allocation, evict_last in L2;
//   table_*      : rows of the small relation table: read-only path with L1 allocation, so the R' x SLAB block that
//                  the CTAs of one SM share stays on chip.
// ---------------------------------------------------------------------------------------------
template <typename T, int VEC> struct Vec { T v[VEC]; };

__device__ __forceinline__ void gather_load(const float *p, Vec<float, 4> &out) {
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(out.v[0]), "=f"(out.v[1]), "=f"(out.v[2]), "=f"(out.v[3]) : "l"(p));
}
__device__ __forceinline__ void gather_load(const float *p, Vec<float, 2> &out) {
    asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(out.v[0]), "=f"(out.v[1]) : "l"(p));
}
__device__ __forceinline__ void gather_load(const float *p, Vec<float, 1> &out) {
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(out.v[0]) : "l"(p));
}
__device__ __forceinline__ void gather_load(const double *p, Vec<double, 2> &out) {
    asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(out.v[0]), "=d"(out.v[1]) : "l"(p));
}
__device__ __forceinline__ void gather_load(const double *p, Vec<double, 1> &out) {
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(out.v[0]) : "l"(p));
}

// L2 eviction-priority policies (createpolicy): the gathered slab is marked evict_last, the streams that pass
// through once (edge ids, result rows) evict_first, so that a slab close to the L2 capacity is not pushed out
// by them (profiles/r01: C4's 63 MB slab + 63 MB of result rows + 17 MB of edge ids per slab pass).
__device__ __forceinline__ unsigned long long policy_evict_last() {
    unsigned long long policy;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    return policy;
}
__device__ __forceinline__ unsigned long long policy_evict_first() {
    unsigned long long policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    return policy;
}
__device__ __forceinline__ void gather_load_keep(const float *p, Vec<float, 4> &out, unsigned long long policy) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(out.v[0]), "=f"(out.v[1]), "=f"(out.v[2]), "=f"(out.v[3]) : "l"(p), "l"(policy));
}
__device__ __forceinline__ void gather_load_keep(const float *p, Vec<float, 2> &out, unsigned long long policy) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;"
                 : "=f"(out.v[0]), "=f"(out.v[1]) : "l"(p), "l"(policy));
}
__device__ __forceinline__ void gather_load_keep(const float *p, Vec<float, 1> &out, unsigned long long policy) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(out.v[0]) : "l"(p), "l"(policy));
}
__device__ __forceinline__ void gather_load_keep(const double *p, Vec<double, 2> &out, unsigned long long policy) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;"
                 : "=d"(out.v[0]), "=d"(out.v[1]) : "l"(p), "l"(policy));
}
__device__ __forceinline__ void gather_load_keep(const double *p, Vec<double, 1> &out, unsigned long long policy) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(out.v[0]) : "l"(p), "l"(policy));
}
__device__ __forceinline__ unsigned edge_load_once(const unsigned *p, unsigned long long policy) {
    unsigned e;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(e) : "l"(p), "l"(policy));
    return e;
}
__device__ __forceinline__ int2 edge_load_once(const int2 *p, unsigned long long policy) {
    int2 e;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.s32 {%0,%1}, [%2], %3;" : "=r"(e.x), "=r"(e.y) : "l"(p), "l"(policy));
    return e;
}

__device__ __forceinline__ void table_load(const float *p, Vec<float, 4> &out) {
    const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
    out.v[0] = t.x; out.v[1] = t.y; out.v[2] = t.z; out.v[3] = t.w;
}
__device__ __forceinline__ void table_load(const float *p, Vec<float, 2> &out) {
    const float2 t = __ldg(reinterpret_cast<const float2 *>(p));
    out.v[0] = t.x; out.v[1] = t.y;
}
__device__ __forceinline__ void table_load(const float *p, Vec<float, 1> &out) { out.v[0] = __ldg(p); }
__device__ __forceinline__ void table_load(const double *p, Vec<double, 2> &out) {
    const double2 t = __ldg(reinterpret_cast<const double2 *>(p));
    out.v[0] = t.x; out.v[1] = t.y;
}
__device__ __forceinline__ void table_load(const double *p, Vec<double, 1> &out) { out.v[0] = __ldg(p); }

// streaming (evict-first) stores for results that are written once and not re-read by this kernel
__device__ __forceinline__ void stream_store(float *p, const Vec<float, 4> &v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.v[0]), "f"(v.v[1]), "f"(v.v[2]), "f"(v.v[3]) : "memory");
}
__device__ __forceinline__ void stream_store(float *p, const Vec<float, 2> &v) {
    asm volatile("st.global.cs.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.v[0]), "f"(v.v[1]) : "memory");
}
__device__ __forceinline__ void stream_store(float *p, const Vec<float, 1> &v) {
    asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v.v[0]) : "memory");
}
__device__ __forceinline__ void stream_store(double *p, const Vec<double, 2> &v) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.v[0]), "d"(v.v[1]) : "memory");
}
__device__ __forceinline__ void stream_store(double *p, const Vec<double, 1> &v) {
    asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p), "d"(v.v[0]) : "memory");
}
__device__ __forceinline__ void stream_store(int32_t *p, const Vec<int32_t, 4> &v) {
    asm volatile("st.global.cs.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.v[0]), "r"(v.v[1]), "r"(v.v[2]), "r"(v.v[3]) : "memory");
}
__device__ __forceinline__ void stream_store(int32_t *p, const Vec<int32_t, 2> &v) {
    asm volatile("st.global.cs.v2.s32 [%0], {%1,%2};" ::"l"(p), "r"(v.v[0]), "r"(v.v[1]) : "memory");
}
__device__ __forceinline__ void stream_store(int32_t *p, const Vec<int32_t, 1> &v) {
    asm volatile("st.global.cs.s32 [%0], %1;" ::"l"(p), "r"(v.v[0]) : "memory");
}

__device__ __forceinline__ double shfl_value(double v, int src) { return __shfl_sync(kFullMask, v, src); }
__device__ __forceinline__ float shfl_value(float v, int src) { return __shfl_sync(kFullMask, v, src); }
__device__ __forceinline__ double shfl_xor_value(double v, int m) { return __shfl_xor_sync(kFullMask, v, m); }
__device__ __forceinline__ float shfl_xor_value(float v, int m) { return __shfl_xor_sync(kFullMask, v, m); }

// message = w * (rel (x) x) with the rounding sequence of the reference (two roundings, no contraction
// into the following add for min/max, where the value itself is the result).
template <typename T, int MSG> __device__ __forceinline__ T message(T w, T r, T x) {
    if (MSG == MSG_MUL) return w * (r * x);
    if (MSG == MSG_ADD) return w * (r + x);
    return w * x;
}

// unit-weight form: 1 * v == v exactly, so the multiply is dropped
template <typename T, int MSG> __device__ __forceinline__ T message(T r, T x) {
    if (MSG == MSG_MUL) return r * x;
    if (MSG == MSG_ADD) return r + x;
    return x;
}

// task.w = (slot + 1) | kNonUnitTask: slot of the partial row the task writes (-1 = the result row itself) and
// whether any of its edges has a merged weight different from 1
constexpr int kNonUnitTask = 0x40000000;
// group tasks (grouped task list): task.w = kGroupTask | (rows - 1) << 24 | kNonUnitTask?; they never write partial rows
constexpr int kGroupTask = 0x20000000;
constexpr int kGroupRows = 16;   // most segments per group task (4 bits at bit 24)
__host__ __device__ __forceinline__ int task_slot(int encoded) {
    return (encoded & kGroupTask) ? -1 : (encoded & (kGroupTask - 1)) - 1;
}
__host__ __device__ __forceinline__ int task_rows(int encoded) {
    return (encoded & kGroupTask) ? ((encoded >> 24) & (kGroupRows - 1)) + 1 : 1;
}

// ---- rows-in-shared-memory kernel (rspmm_staged.cu): operands with few rows, e.g. the graph of relations ------------
constexpr int kStagedSlab = 64;         // features per slab: a half-warp reads one 256-byte row slab per LDS.128
constexpr int kStagedMaxRows = 864;     // 864 x 256 B = 216 KB of dynamic shared memory (+ 8 KB static) <= 227 KB per CTA
constexpr size_t kStagedMaxSmem = (size_t)kStagedMaxRows * kStagedSlab * sizeof(float);
struct StagedArgs {
    const int4 *task;         // plain task list of the order
    const unsigned *packed;   // edge ids, x | y << pack_shift
    int pack_shift;
    const float *w;           // null when all weights are 1
    const float *A;           // gathered operand: n_rows rows, staged slab by slab
    const float *B;           // relation table (unused for MSG_COPY)
    float *out;
    const float *addend;
    float *partial;
    unsigned *counter;        // work counter in the workspace (zeroed by the launcher)
    long long dim;
    int n_task, n_slab, n_rows, tasks_per_item;
    long long a_stride, o_stride, o_offset, a_row, o_row;   // blocked layout, as SegArgs (rspmm_kernels.cu)
    int block, block_shift;
};
int launch_rows_in_smem(StagedArgs args, int msg, cudaStream_t stream);

// pair kernel (<= 4 relation types, unit weights): one shared-memory row read per (segment, other node) pair
struct PairArgs {
    const int32_t *ptr;       // n_seg + 1
    const uint32_t *pair;     // other | mask << id_bits
    const int32_t *rows;      // segments by descending pair count
    int id_bits;
    const float *A;           // gathered operand (n_rows rows, staged slab by slab)
    const float *B;           // relation table, n_rel <= 4 rows (unused for MSG_COPY)
    float *out;
    const float *addend;
    unsigned *counter;
    long long dim;
    int n_seg, n_rows, n_rel, n_slab;
    long long a_stride, o_stride, o_offset, a_row, o_row;
    int block, block_shift;
};
int launch_pairs_in_smem(PairArgs args, int msg, cudaStream_t stream);

// grad_relation with the grad_output rows of a destination block staged in shared memory (rel order + block table)
struct BlockedRelArgs {
    const int32_t *block_ptr; // n_rel x (n_block + 1)
    const int2 *edge;         // rel order: {dst, src}
    const unsigned *packed;   // or null
    int pack_shift;
    const float *w;           // rel-order weights, null when all 1
    const float *G;           // grad_output (n_out, dim): staged
    const float *X;           // input (n_in, dim): gathered (unused for MSG_COPY)
    const float *O;           // gated pass (min / max): output (n_out, dim), staged beside G
    const float *R;           // gated pass: relation (n_rel, dim), the run's own row
    float *partial;           // (n_rel * n_block, dim)
    unsigned *counter;
    long long dim;
    int n_rel, n_block, block_rows, n_out, n_slab;
};
int launch_dst_blocked(BlockedRelArgs args, int msg, cudaStream_t stream);
int launch_dst_blocked_gated(BlockedRelArgs args, int msg, cudaStream_t stream);   // msg: MSG_MUL or MSG_ADD
extern int g_pairs, g_blocked;

// sub-warp rows kernel (rspmm_narrow.cu): graphs whose 512-byte slab exceeds L2 - SUB tasks per warp, 256 / 128-byte slabs
struct NarrowArgs {
    const int4 *task;         // plain task list of the order
    const int2 *edge;
    const unsigned *packed;   // or null
    int pack_shift;
    const float *w;           // null when all weights are 1
    const float *A;           // gathered by edge.x
    const float *B;           // table / gathered by edge.y (unused for MSG_COPY)
    float *out;
    const float *addend;
    float *partial;
    long long dim;
    int n_task, n_slab;
};
int launch_narrow(const NarrowArgs &args, int msg, bool b_table, int sub, cudaStream_t stream);
extern int g_narrow_sub;           // 0: by slab size, 2 / 4: forced
extern long long g_narrow_bytes;   // a 512-byte slab of the gathered operand above this size takes the sub-warp kernel (0: never)

// launch bookkeeping (claimed in bench.py as `gpu_launches`)
void note_launch();
// how the last pass of each kind was launched (ultra_rspmm_last_pass_info; read by the parity tests)
void note_pass(int pass, const ultra_rspmm_pass_info_t &info);
// per-thread status plumbing shared by the translation units
int fail_cuda(cudaError_t error);
extern int g_chunk;    // edges per task for new indexes
extern int g_variant;  // 0 auto, 1 generic, 2 staged
extern int g_group_edges;  // rows of up to this many edges are grouped in the grouped task list (-1: chunk / 4, 0: off)
extern int g_staged;   // rows-in-shared-memory kernel: 0 off, 1 automatic, 2 whenever the operand fits
extern long long g_l2_budget;  // bytes of L2 the gathered operand's slab may occupy (slab width is chosen to fit)

#define ULTRA_CUDA_OK(expr)                                      \
    do {                                                         \
        cudaError_t ultra_err_ = (expr);                         \
        if (ultra_err_ != cudaSuccess) return ::ultra::fail_cuda(ultra_err_); \
    } while (0)

static inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

}  // namespace ultra
