// Gather-combine-reduce for graphs with so many rows that a 512-byte column slab of the gathered operand exceeds L2.
//
// The generic kernel (rspmm_kernels.cu) gives a warp one (task, 128-feature slab): with N = 524,288 rows the slab of the
// gathered operand is 268 MB, twice the 126 MB L2, and the pass becomes HBM-bound (ncu, 16.8 M edges: 194 GB of DRAM reads
// for 275 GB of gathers, L2 hit rate 32 %, DRAM 72 % busy, 34 ms).  Narrowing the slab makes it L2-resident again - but
// the library's narrow variants (VEC = 2 / 1: 8 or 4 bytes per lane) double or quadruple the instructions per byte and lost
// on every shape.  Here the slab narrows WITHOUT narrowing the per-lane access: a warp is split into SUB = 2 or 4
// sub-warps of 16 or 8 lanes, each sub-warp owns its own task (row) and a 256- or 128-byte slab of it, every lane still
// moves 16 bytes per load.  A warp instruction then gathers SUB different rows; no cross-lane reduction is needed at all
// (a sub-warp writes its own result row piece), and the slab of the gathered operand is N x 256 B or N x 128 B (67 MB at
// N = 524,288: L2-resident with the evict_last hint), so DRAM sees about the compulsory bytes again.
//
// Serves sum aggregation (forward on the csr order, grad_input on the csc order, grad_relation on the rel order when the
// destination-blocked kernel does not apply) of fp32 operands over the plain task list; tasks are sorted by length, so the
// SUB tasks of a warp have (almost) the same number of edges.  Same arithmetic and the same fixed summation order per
// task as the generic kernel's (4 edges per step, in edge order): deterministic, no atomics.
#include <type_traits>

#include "rspmm_common.cuh"

namespace ultra {

namespace {

__device__ __forceinline__ int narrow_x(const int2 &e) { return e.x; }
__device__ __forceinline__ int narrow_y(const int2 &e) { return e.y; }
__device__ __forceinline__ int narrow_x(const unsigned &) { return 0; }
__device__ __forceinline__ int narrow_y(const unsigned &) { return 0; }
__device__ __forceinline__ unsigned narrow_bits(const unsigned &e) { return e; }
__device__ __forceinline__ unsigned narrow_bits(const int2 &) { return 0; }

template <int MSG, bool B_TABLE, bool PACKED, int SUB>
__global__ void __launch_bounds__(kThreadsPerBlock, 4) seg_rows_kernel(const NarrowArgs a) {
    using Ids = typename std::conditional<PACKED, unsigned, int2>::type;
    constexpr int LPS = 32 / SUB;                 // lanes per sub-warp; sub-slab = LPS * 4 features
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPS, l = lane % LPS;
    const long long gw = (long long)blockIdx.x * kWarpsPerBlock + warp;
    const int n_group = (a.n_task + SUB - 1) / SUB;
    if (gw >= (long long)n_group * a.n_slab) return;
    const int slab = (int)(gw / n_group);
    const int t = (int)(gw - (long long)slab * n_group) * SUB + sub;
    const bool has = t < a.n_task;
    const int4 task = has ? __ldg(a.task + t) : make_int4(0, 0, 0, 0);
    const int slot = task_slot(task.w);
    const long long col = (long long)slab * (LPS * 4) + l * 4;
    const bool active = has && col < a.dim;
    const unsigned row_bytes = (unsigned)(a.dim * sizeof(float));
    const char *A = reinterpret_cast<const char *>(a.A + (col < a.dim ? col : 0));
    const char *B = reinterpret_cast<const char *>(a.B + (col < a.dim ? col : 0));
    const Ids *ids = reinterpret_cast<const Ids *>(PACKED ? (const void *)a.packed : (const void *)a.edge);
    const int shift = a.pack_shift;
    const unsigned low = PACKED ? (shift >= 32 ? 0xffffffffu : ((1u << shift) - 1u)) : 0u;
    const unsigned long long keep_policy = policy_evict_last(), once_policy = policy_evict_first();
    const bool weighted = a.w != nullptr && (task.w & kNonUnitTask);
    auto first_id = [&](const Ids &e) { return PACKED ? (int)(narrow_bits(e) & low) : narrow_x(e); };
    auto second_id = [&](const Ids &e) { return PACKED ? (int)(narrow_bits(e) >> shift) : narrow_y(e); };

    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    constexpr int kStep = 4;                      // edges in flight per sub-warp
    Ids ahead[kStep];
#pragma unroll
    for (int k = 0; k < kStep; ++k) ahead[k] = task.y + k < task.z ? edge_load_once(ids + task.y + k, once_policy) : Ids();
    for (int pos = task.y; __any_sync(kFullMask, pos < task.z); pos += kStep) {
        Ids e[kStep];
        Vec<float, 4> va[kStep], vb[kStep];
        float w[kStep];
#pragma unroll
        for (int k = 0; k < kStep; ++k) {
            e[k] = ahead[k];
            const bool live = pos + k < task.z;
            w[k] = 1.f;
#pragma unroll
            for (int v = 0; v < 4; ++v) va[k].v[v] = vb[k].v[v] = 0.f;
            if (live) {
                gather_load_keep(reinterpret_cast<const float *>(A + (unsigned long long)(unsigned)first_id(e[k]) * row_bytes), va[k], keep_policy);
                if (weighted) w[k] = __ldg(a.w + pos + k);
            }
        }
#pragma unroll
        for (int k = 0; k < kStep; ++k)      // the next step's ids travel while this step's rows do
            ahead[k] = pos + kStep + k < task.z ? edge_load_once(ids + pos + kStep + k, once_policy) : Ids();
        if (MSG != MSG_COPY) {
#pragma unroll
            for (int k = 0; k < kStep; ++k) {
                if (pos + k < task.z) {
                    const float *p = reinterpret_cast<const float *>(B + (unsigned long long)(unsigned)second_id(e[k]) * row_bytes);
                    if (B_TABLE) table_load(p, vb[k]);
                    else gather_load_keep(p, vb[k], keep_policy);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kStep; ++k) {
            if (pos + k < task.z) {
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const float m = weighted ? message<float, MSG>(w[k], vb[k].v[v], va[k].v[v]) : message<float, MSG>(vb[k].v[v], va[k].v[v]);
                    acc[v] += m;
                }
            }
        }
    }
    if (!active) return;
    Vec<float, 4> r;
#pragma unroll
    for (int v = 0; v < 4; ++v) r.v[v] = acc[v];
    if (slot < 0) {
        const long long row = task.x;
        if (a.addend) {
            Vec<float, 4> b;
            gather_load(a.addend + row * a.dim + col, b);
#pragma unroll
            for (int v = 0; v < 4; ++v) r.v[v] += b.v[v];
        }
        stream_store(a.out + row * a.dim + col, r);
    } else {
        float *p = a.partial + (long long)slot * a.dim + col;
        *reinterpret_cast<float4 *>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    }
}

template <int MSG, bool B_TABLE, int SUB> int launch_rows(const NarrowArgs &args, cudaStream_t stream) {
    const int n_group = (args.n_task + SUB - 1) / SUB;
    const long long warps = (long long)n_group * args.n_slab;
    const long long blocks = (warps + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (blocks > 0x7fffffffLL) return ULTRA_RSPMM_ERR_RANGE;
    if (args.packed) seg_rows_kernel<MSG, B_TABLE, true, SUB><<<(unsigned)blocks, kThreadsPerBlock, 0, stream>>>(args);
    else seg_rows_kernel<MSG, B_TABLE, false, SUB><<<(unsigned)blocks, kThreadsPerBlock, 0, stream>>>(args);
    note_launch();
    return ULTRA_RSPMM_OK;
}

template <int MSG, bool B_TABLE> int launch_sub(NarrowArgs args, int sub, cudaStream_t stream) {
    const int features = 32 / sub * 4;
    args.n_slab = (int)((args.dim + features - 1) / features);
    return sub == 2 ? launch_rows<MSG, B_TABLE, 2>(args, stream) : launch_rows<MSG, B_TABLE, 4>(args, stream);
}

}  // namespace

int launch_narrow(const NarrowArgs &args, int msg, bool b_table, int sub, cudaStream_t stream) {
    if (args.n_task == 0 || args.dim == 0) return ULTRA_RSPMM_OK;
    if ((sub != 2 && sub != 4) || args.dim * (long long)sizeof(float) > 0xffffffffLL) return ULTRA_RSPMM_ERR_ARG;
    if (msg == MSG_MUL) return b_table ? launch_sub<MSG_MUL, true>(args, sub, stream) : launch_sub<MSG_MUL, false>(args, sub, stream);
    if (msg == MSG_ADD) return launch_sub<MSG_ADD, true>(args, sub, stream);
    return launch_sub<MSG_COPY, true>(args, sub, stream);
}

}  // namespace ultra
