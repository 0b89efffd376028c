// Gather-combine-reduce for operands with FEW rows (sm_100a): the gathered operand's column slab lives in shared memory.
//
// The graph of relations that ULTRA's relation model walks (reference ultra/rel_model.py:253-257, 369-374, built by
// :91-147) has N_r = 2R nodes (474 at the FB15k-237 shape), 4 edge types and up to 4 N_r^2 edges: every destination row
// gathers ~1,900 source rows out of the same 474.  The generic kernel (rspmm_kernels.cu) serves those gathers from
// L1/L2 and is bound by L2->SM bandwidth there.  Here a persistent CTA copies the whole (rows x 64 features) slab of the
// gathered operand into shared memory once (121 KB at 474 rows) and its 32 warps then walk tasks of that slab:
//
//   * per edge one LDS.128 per lane - a half-warp reads the 256-byte slab of one source row, the two halves of a warp
//     work on different edges, so one instruction moves 512 B with the minimum of 4 shared-memory wavefronts;
//   * the relation row of the current run stays in registers (edges of a segment are grouped by relation);
//   * edge ids are staged per warp exactly as in the generic kernel (coalesced load, LDS.128 of 4 packed ids);
//   * work is handed out dynamically: an item = (slab, 32..128 consecutive tasks of the task list), items are numbered
//     slab-major and fetched with one atomicAdd per item and CTA, so the CTAs drain the list evenly and refill their
//     slab only when the slab changes.  The counter only schedules: which CTA runs a task never changes its result, and a
//     task's edges are always summed in the same order (deterministic, no atomics on data).
//
// Serves sum aggregation with mul / add / copy messages (forward on the csr order, grad_input on the csc order) of fp32
// operands whose ids pack into 32 bits; run_pass (rspmm_kernels.cu) selects it when the slab fits shared memory.
#include <cstdlib>
#include <type_traits>

#include "rspmm_common.cuh"

namespace ultra {

namespace {

constexpr int kStagedWarps = 32;
constexpr int kStagedThreads = kStagedWarps * 32;
constexpr int kEdgesPerHalf = 4;   // edges in flight per half-warp

__device__ __forceinline__ int id_first(const int2 &e) { return e.x; }
__device__ __forceinline__ int id_second(const int2 &e) { return e.y; }
__device__ __forceinline__ int id_first(const unsigned &) { return 0; }
__device__ __forceinline__ int id_second(const unsigned &) { return 0; }
__device__ __forceinline__ unsigned id_bits_of(const unsigned &e) { return e; }
__device__ __forceinline__ unsigned id_bits_of(const int2 &) { return 0; }

__device__ __forceinline__ long long staged_blocked_col(long long col, int block, int shift, long long stride) {
    const unsigned c = (unsigned)col;
    const unsigned q = shift >= 0 ? c >> shift : c / (unsigned)block;
    return (long long)q * stride + (c - q * (unsigned)block);
}

template <int MSG, bool UNIT>
__device__ __forceinline__ void staged_accumulate(const float4 &x, const float4 &r, float w, float (&acc)[4]) {
    const float xv[4] = {x.x, x.y, x.z, x.w};
    const float rv[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const float m = UNIT ? message<float, MSG>(rv[v], xv[v]) : message<float, MSG>(w, rv[v], xv[v]);
        acc[v] += m;
    }
}

// one edge, relation row looked up if it differs from the cached one (run boundaries and the ragged tail of a task)
template <int MSG, bool UNIT>
__device__ __forceinline__ void staged_edge(unsigned id, float w, unsigned low, int shift, const float4 *s_rows, int l16,
                                            const char *B, unsigned row_bytes, float4 &cached, int &cached_rel, float (&acc)[4]) {
    const float4 x = s_rows[(id & low) * (kStagedSlab / 4) + l16];
    if (MSG != MSG_COPY) {
        const int rel = (int)(id >> shift);
        if (rel != cached_rel) {   // at most n_rel times per segment: the edges of a segment are grouped by relation
            cached = __ldg(reinterpret_cast<const float4 *>(B + (unsigned long long)(unsigned)rel * row_bytes));
            cached_rel = rel;
        }
    }
    staged_accumulate<MSG, UNIT>(x, cached, w, acc);
}

// four consecutive edges of one half-warp.  Common case: all of them use the cached relation row - four LDS.128 in
// flight, then 16 FFMA, no table access
template <int MSG, bool UNIT>
__device__ __forceinline__ void staged_edges4(const unsigned (&ids)[4], const float (&ws)[4], unsigned low, int shift,
                                              const float4 *s_rows, int l16, const char *B, unsigned row_bytes,
                                              float4 &cached, int &cached_rel, float (&acc)[4]) {
    bool same = true;
    if (MSG != MSG_COPY) {
        const unsigned c = (unsigned)cached_rel << shift;   // cached_rel = -1 never matches: ids have no bits above the relation
        same = cached_rel >= 0 && ((((ids[0] ^ c) | (ids[1] ^ c)) | ((ids[2] ^ c) | (ids[3] ^ c))) >> shift) == 0;
    }
    if (same) {
        float4 x[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) x[k] = s_rows[(ids[k] & low) * (kStagedSlab / 4) + l16];
#pragma unroll
        for (int k = 0; k < 4; ++k) staged_accumulate<MSG, UNIT>(x[k], cached, ws[k], acc);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) staged_edge<MSG, UNIT>(ids[k], ws[k], low, shift, s_rows, l16, B, row_bytes, cached, cached_rel, acc);
    }
}

template <int MSG>
__global__ void __launch_bounds__(kStagedThreads, 1) rows_in_smem_kernel(const StagedArgs a) {
    extern __shared__ __align__(16) float4 s_rows[];   // n_rows x 16 float4: the slab of the gathered operand
    __shared__ __align__(16) unsigned s_edge[kStagedWarps][32];
    __shared__ __align__(16) float s_w[kStagedWarps][32];
    __shared__ int s_item;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = lane >> 4, l16 = lane & 15;
    const int shift = a.pack_shift;
    const unsigned low = shift >= 32 ? 0xffffffffu : ((1u << shift) - 1u);
    const unsigned row_bytes = (unsigned)(a.dim * sizeof(float));
    const int items_per_slab = (a.n_task + a.tasks_per_item - 1) / a.tasks_per_item;
    const int n_item = items_per_slab * a.n_slab;
    int resident = -1;
    for (;;) {
        if (threadIdx.x == 0) s_item = (int)atomicAdd(a.counter, 1u);
        __syncthreads();
        const int item = s_item;
        if (item >= n_item) break;
        const int slab = item / items_per_slab;
        const int first = (item - slab * items_per_slab) * a.tasks_per_item;
        const int last = min(a.n_task, first + a.tasks_per_item);
        if (slab != resident) {   // (re)fill: 256 contiguous bytes per row, 16 lanes each
            for (int i = threadIdx.x; i < a.n_rows * (kStagedSlab / 4); i += kStagedThreads) {
                const int row = i >> 4;
                const long long col = (long long)slab * kStagedSlab + (i & 15) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (col < a.dim) {
                    const long long at = a.block ? staged_blocked_col(col, a.block, a.block_shift, a.a_stride) : col;
                    v = __ldg(reinterpret_cast<const float4 *>(a.A + (long long)row * a.a_row + at));
                }
                s_rows[i] = v;
            }
            resident = slab;
            __syncthreads();
        }
        const long long col = (long long)slab * kStagedSlab + l16 * 4;
        const bool active = col < a.dim;
        const char *B = reinterpret_cast<const char *>(a.B + (active ? col : 0));
        for (int t = first + warp; t < last; t += kStagedWarps) {
            const int4 task = __ldg(a.task + t);
            const int slot = task_slot(task.w);
            const bool unit = a.w == nullptr || !(task.w & kNonUnitTask);
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            float4 cached = make_float4(0.f, 0.f, 0.f, 0.f);
            int cached_rel = -1;
            unsigned ahead = 0;
            float ahead_w = 1.f;
            if (task.y + lane < task.z) {
                ahead = __ldg(a.packed + task.y + lane);
                if (!unit) ahead_w = __ldg(a.w + task.y + lane);
            }
            for (int base = task.y; base < task.z; base += 32) {
                const int n = min(32, task.z - base);
                __syncwarp();
                s_edge[warp][lane] = ahead;
                if (!unit) s_w[warp][lane] = ahead_w;
                __syncwarp();
                if (base + 32 + lane < task.z) {
                    ahead = __ldg(a.packed + base + 32 + lane);
                    if (!unit) ahead_w = __ldg(a.w + base + 32 + lane);
                }
                int u = 0;
                // full groups: each half-warp takes 4 consecutive edges (one LDS.128 of ids), no predicates
                for (; u + 2 * kEdgesPerHalf <= n; u += 2 * kEdgesPerHalf) {
                    const int mine = u + half * kEdgesPerHalf;
                    const uint4 id = *reinterpret_cast<const uint4 *>(&s_edge[warp][mine]);
                    const unsigned ids[4] = {id.x, id.y, id.z, id.w};
                    if (unit) {
                        const float ws[4] = {1.f, 1.f, 1.f, 1.f};
                        staged_edges4<MSG, true>(ids, ws, low, shift, s_rows, l16, B, row_bytes, cached, cached_rel, acc);
                    } else {
                        const float4 w4 = *reinterpret_cast<const float4 *>(&s_w[warp][mine]);
                        const float ws[4] = {w4.x, w4.y, w4.z, w4.w};
                        staged_edges4<MSG, false>(ids, ws, low, shift, s_rows, l16, B, row_bytes, cached, cached_rel, acc);
                    }
                }
                // ragged tail (< 8 edges): the halves alternate, idle half-warps skip
                for (; u < n; u += 2) {
                    const int mine = u + half;
                    if (mine < n) {
                        if (unit) staged_edge<MSG, true>(s_edge[warp][mine], 1.f, low, shift, s_rows, l16, B, row_bytes, cached, cached_rel, acc);
                        else staged_edge<MSG, false>(s_edge[warp][mine], s_w[warp][mine], low, shift, s_rows, l16, B, row_bytes, cached, cached_rel, acc);
                    }
                }
            }
            // fold the two half-warps (fixed order: lower half + upper half), lanes 0..15 write 256 bytes
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[v] += __shfl_xor_sync(kFullMask, acc[v], 16);
            if (half == 0 && active) {
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
                    const long long o_col = a.block ? staged_blocked_col(col, a.block, a.block_shift, a.o_stride) + a.o_offset : col;
                    stream_store(a.out + row * a.o_row + o_col, r);
                } else {
                    float *p = a.partial + (long long)slot * a.dim + col;
                    *reinterpret_cast<float4 *>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
                }
            }
        }
        __syncthreads();   // every warp is done with s_item and with the resident slab before the next fetch / refill
    }
}

// ---- pair kernel -----------------------------------------------------------------------------------------------------
// One warp per segment (rows are handed out longest first); per pair one LDS.128 of the other node's row and, for each
// relation bit of the mask, 4 predicated FMAs with that relation's row, which stays in registers for the whole slab.
template <int MSG>
__device__ __forceinline__ void pair_accumulate(unsigned word, unsigned low, int id_bits, const float4 *s_rows, int l16,
                                                const float4 (&rel)[4], float (&acc)[4]) {
    const float4 x = s_rows[(word & low) * (kStagedSlab / 4) + l16];
    const unsigned mask = word >> id_bits;
    const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (mask & (1u << k)) {
            const float rv[4] = {rel[k].x, rel[k].y, rel[k].z, rel[k].w};
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[v] += message<float, MSG>(rv[v], xv[v]);
        }
    }
}

template <int MSG>
__global__ void __launch_bounds__(kStagedThreads, 1) pairs_in_smem_kernel(const PairArgs a) {
    extern __shared__ __align__(16) float4 s_rows[];
    __shared__ __align__(16) unsigned s_pair[kStagedWarps][32];
    __shared__ int s_item;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = lane >> 4, l16 = lane & 15;
    const unsigned low = (1u << a.id_bits) - 1u;
    const int items_per_slab = (a.n_seg + kStagedWarps - 1) / kStagedWarps;
    const int n_item = items_per_slab * a.n_slab;
    int resident = -1;
    float4 rel[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) rel[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (;;) {
        if (threadIdx.x == 0) s_item = (int)atomicAdd(a.counter, 1u);
        __syncthreads();
        const int item = s_item;
        if (item >= n_item) break;
        const int slab = item / items_per_slab;
        const long long col = (long long)slab * kStagedSlab + l16 * 4;
        const bool active = col < a.dim;
        if (slab != resident) {
            for (int i = threadIdx.x; i < a.n_rows * (kStagedSlab / 4); i += kStagedThreads) {
                const int row = i >> 4;
                const long long c = (long long)slab * kStagedSlab + (i & 15) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c < a.dim) {
                    const long long at = a.block ? staged_blocked_col(c, a.block, a.block_shift, a.a_stride) : c;
                    v = __ldg(reinterpret_cast<const float4 *>(a.A + (long long)row * a.a_row + at));
                }
                s_rows[i] = v;
            }
            if (MSG != MSG_COPY) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    rel[k] = (k < a.n_rel && active) ? __ldg(reinterpret_cast<const float4 *>(a.B + (long long)k * a.dim + col))
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            resident = slab;
            __syncthreads();
        }
        const int position = (item - slab * items_per_slab) * kStagedWarps + warp;
        if (position < a.n_seg) {
            const int seg = __ldg(a.rows + position);
            const int begin = __ldg(a.ptr + seg), end = __ldg(a.ptr + seg + 1);
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            unsigned ahead = begin + lane < end ? __ldg(a.pair + begin + lane) : 0u;
            for (int base = begin; base < end; base += 32) {
                const int n = min(32, end - base);
                __syncwarp();
                s_pair[warp][lane] = ahead;
                __syncwarp();
                if (base + 32 + lane < end) ahead = __ldg(a.pair + base + 32 + lane);
                int u = 0;
                for (; u + 2 * kEdgesPerHalf <= n; u += 2 * kEdgesPerHalf) {
                    const uint4 w = *reinterpret_cast<const uint4 *>(&s_pair[warp][u + half * kEdgesPerHalf]);
                    const unsigned words[4] = {w.x, w.y, w.z, w.w};
                    float4 x[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) x[k] = s_rows[(words[k] & low) * (kStagedSlab / 4) + l16];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const unsigned mask = words[k] >> a.id_bits;
                        const float xv[4] = {x[k].x, x[k].y, x[k].z, x[k].w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (mask & (1u << q)) {
                                const float rv[4] = {rel[q].x, rel[q].y, rel[q].z, rel[q].w};
#pragma unroll
                                for (int v = 0; v < 4; ++v) acc[v] += message<float, MSG>(rv[v], xv[v]);
                            }
                        }
                    }
                }
                for (; u < n; u += 2) {
                    const int mine = u + half;
                    if (mine < n) pair_accumulate<MSG>(s_pair[warp][mine], low, a.id_bits, s_rows, l16, rel, acc);
                }
            }
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[v] += __shfl_xor_sync(kFullMask, acc[v], 16);
            if (half == 0 && active) {
                Vec<float, 4> r;
#pragma unroll
                for (int v = 0; v < 4; ++v) r.v[v] = acc[v];
                if (a.addend) {
                    Vec<float, 4> b;
                    gather_load(a.addend + (long long)seg * a.dim + col, b);
#pragma unroll
                    for (int v = 0; v < 4; ++v) r.v[v] += b.v[v];
                }
                const long long o_col = a.block ? staged_blocked_col(col, a.block, a.block_shift, a.o_stride) + a.o_offset : col;
                stream_store(a.out + (long long)seg * a.o_row + o_col, r);
            }
        }
        __syncthreads();
    }
}

// ---- destination-blocked grad_relation ------------------------------------------------------------------------------------
// item = (slab, destination block b): the CTA stages grad_output rows [b * block_rows, (b + 1) * block_rows) of the slab,
// then warp w reduces the (relation k, block b) runs for k = w, w + 32, ...:  acc += g[dst] (x) x[src]  with g from shared
// memory and x gathered (half-warp per edge, LDG.128).  Every (k, b) run writes one partial row (zeros when empty); the
// combine pass folds the n_block partial rows of a relation in block order - deterministic, no atomics on data.
template <int MSG, bool PACKED, int WARPS, int CTAS>   // WARPS x 32 threads per CTA, CTAS CTAs per SM (WARPS * CTAS = 32)
__global__ void __launch_bounds__(WARPS * 32, CTAS) dst_blocked_kernel(const BlockedRelArgs a) {
    using Ids = typename std::conditional<PACKED, unsigned, int2>::type;
    constexpr int kThreads = WARPS * 32;
    extern __shared__ __align__(16) float4 s_rows[];
    __shared__ __align__(16) Ids s_edge[WARPS][32];
    __shared__ __align__(16) float s_w[WARPS][32];
    __shared__ int s_item;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = lane >> 4, l16 = lane & 15;
    const int shift = a.pack_shift;
    const unsigned low = PACKED ? ((1u << shift) - 1u) : 0u;
    const unsigned row_bytes = (unsigned)(a.dim * sizeof(float));
    const int n_item = a.n_block * a.n_slab;
    const Ids *ids = reinterpret_cast<const Ids *>(PACKED ? (const void *)a.packed : (const void *)a.edge);
    // the gathered slab of X (n_in x 256 B, re-read once per edge) should stay in L2; grad_output rows, edge ids and partial
    // rows pass through once (16.8 M edges: 47.6 -> 45.0 ms; C4: 4.36 -> 4.02 ms)
    const unsigned long long keep_policy = policy_evict_last(), once_policy = policy_evict_first();
    for (;;) {
        if (threadIdx.x == 0) s_item = (int)atomicAdd(a.counter, 1u);
        __syncthreads();
        const int item = s_item;
        if (item >= n_item) break;
        const int slab = item / a.n_block, b = item - slab * a.n_block;
        const int first_row = b * a.block_rows;
        const int rows = min(a.block_rows, a.n_out - first_row);
        for (int i = threadIdx.x; i < rows * (kStagedSlab / 4); i += kThreads) {
            const long long c = (long long)slab * kStagedSlab + (i & 15) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < a.dim) {
                Vec<float, 4> g;
                gather_load_keep(a.G + (long long)(first_row + (i >> 4)) * a.dim + c, g, once_policy);
                v = make_float4(g.v[0], g.v[1], g.v[2], g.v[3]);
            }
            s_rows[i] = v;
        }
        __syncthreads();
        const long long col = (long long)slab * kStagedSlab + l16 * 4;
        const bool active = col < a.dim;
        const char *X = reinterpret_cast<const char *>(a.X + (active ? col : 0));
        for (int k = warp; k < a.n_rel; k += WARPS) {
            const int begin = __ldg(a.block_ptr + (long long)k * (2 * a.n_block + 1) + 2 * b);   // table in half blocks
            const int end = __ldg(a.block_ptr + (long long)k * (2 * a.n_block + 1) + 2 * b + 2);
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            auto one = [&](const Ids &e, float w) {
                const int dst = PACKED ? (int)(id_bits_of(e) & low) : id_first(e);
                const int src = PACKED ? (int)(id_bits_of(e) >> shift) : id_second(e);
                const float4 g = s_rows[(dst - first_row) * (kStagedSlab / 4) + l16];
                float4 x = make_float4(1.f, 1.f, 1.f, 1.f);
                if (MSG == MSG_MUL) {
                    Vec<float, 4> v;
                    gather_load_keep(reinterpret_cast<const float *>(X + (unsigned long long)(unsigned)src * row_bytes), v, keep_policy);
                    x = make_float4(v.v[0], v.v[1], v.v[2], v.v[3]);
                }
                const float gv[4] = {g.x, g.y, g.z, g.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[v] += MSG == MSG_MUL ? w * (gv[v] * xv[v]) : w * gv[v];
            };
            Ids ahead = Ids();
            float ahead_w = 1.f;
            if (begin + lane < end) {
                ahead = edge_load_once(ids + begin + lane, once_policy);
                if (a.w) ahead_w = __ldg(a.w + begin + lane);
            }
            for (int base = begin; base < end; base += 32) {
                const int n = min(32, end - base);
                __syncwarp();
                s_edge[warp][lane] = ahead;
                s_w[warp][lane] = ahead_w;
                __syncwarp();
                if (base + 32 + lane < end) {
                    ahead = edge_load_once(ids + base + 32 + lane, once_policy);
                    if (a.w) ahead_w = __ldg(a.w + base + 32 + lane);
                }
                int u = 0;
                for (; u + 2 * kEdgesPerHalf <= n; u += 2 * kEdgesPerHalf) {
                    const int mine = u + half * kEdgesPerHalf;
                    Ids e[4];
                    float w[4];
                    float4 g[4];
                    Vec<float, 4> x[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        e[q] = s_edge[warp][mine + q];
                        w[q] = s_w[warp][mine + q];
                        const int dst = PACKED ? (int)(id_bits_of(e[q]) & low) : id_first(e[q]);
                        const int src = PACKED ? (int)(id_bits_of(e[q]) >> shift) : id_second(e[q]);
                        if (MSG == MSG_MUL) gather_load_keep(reinterpret_cast<const float *>(X + (unsigned long long)(unsigned)src * row_bytes), x[q], keep_policy);
                        g[q] = s_rows[(dst - first_row) * (kStagedSlab / 4) + l16];
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float gv[4] = {g[q].x, g[q].y, g[q].z, g[q].w};
#pragma unroll
                        for (int v = 0; v < 4; ++v) acc[v] += MSG == MSG_MUL ? w[q] * (gv[v] * x[q].v[v]) : w[q] * gv[v];
                    }
                }
                for (; u < n; u += 2) {
                    const int mine = u + half;
                    if (mine < n) one(s_edge[warp][mine], s_w[warp][mine]);
                }
            }
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[v] += __shfl_xor_sync(kFullMask, acc[v], 16);
            if (half == 0 && active) {
                float *p = a.partial + ((long long)k * a.n_block + b) * a.dim + col;
                __stcs(reinterpret_cast<float4 *>(p), make_float4(acc[0], acc[1], acc[2], acc[3]));
            }
        }
        __syncthreads();
    }
}

// ---- destination-blocked grad_relation of the min / max aggregation (all-ties rule) ----------------------------------------
// The gated pass needs grad_output[dst] AND output[dst] per edge (`if (output[dst] == w (x[src] (x) relation[k])) acc +=
// grad_output[dst] w (x[src] | 1)`): the generic kernel gathers three rows per edge.  Here both are staged - 2 x 256 B per
// destination row, so a CTA works on HALF a block at a time (the table is kept at that granularity) - and only x[src] is
// gathered.  A (k, b) run's partial row is written after the first half and completed after the second by the same lanes.
template <int MSG, bool PACKED, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) dst_blocked_gated_kernel(const BlockedRelArgs a) {
    using Ids = typename std::conditional<PACKED, unsigned, int2>::type;
    constexpr int kThreads = WARPS * 32;
    extern __shared__ __align__(16) float4 s_rows[];   // [half_rows][16] grad_output, then [half_rows][16] output
    __shared__ __align__(16) Ids s_edge[WARPS][32];
    __shared__ __align__(16) float s_w[WARPS][32];
    __shared__ int s_item;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = lane >> 4, l16 = lane & 15;
    const int shift = a.pack_shift;
    const unsigned low = PACKED ? ((1u << shift) - 1u) : 0u;
    const unsigned row_bytes = (unsigned)(a.dim * sizeof(float));
    const int half_rows = a.block_rows / 2;
    float4 *s_out = s_rows + half_rows * (kStagedSlab / 4);
    const int n_item = a.n_block * a.n_slab;
    const Ids *ids = reinterpret_cast<const Ids *>(PACKED ? (const void *)a.packed : (const void *)a.edge);
    const unsigned long long keep_policy = policy_evict_last(), once_policy = policy_evict_first();
    for (;;) {
        if (threadIdx.x == 0) s_item = (int)atomicAdd(a.counter, 1u);
        __syncthreads();
        const int item = s_item;
        if (item >= n_item) break;
        const int slab = item / a.n_block, b = item - slab * a.n_block;
        const long long col = (long long)slab * kStagedSlab + l16 * 4;
        const bool active = col < a.dim;
        const char *X = reinterpret_cast<const char *>(a.X + (active ? col : 0));
        for (int part = 0; part < 2; ++part) {
            const int first_row = b * a.block_rows + part * half_rows;
            const int rows = max(0, min(half_rows, a.n_out - first_row));
            for (int i = threadIdx.x; i < rows * (kStagedSlab / 4); i += kThreads) {
                const long long c = (long long)slab * kStagedSlab + (i & 15) * 4;
                float4 g = make_float4(0.f, 0.f, 0.f, 0.f), o = g;
                if (c < a.dim) {
                    Vec<float, 4> vg, vo;
                    gather_load_keep(a.G + (long long)(first_row + (i >> 4)) * a.dim + c, vg, once_policy);
                    gather_load_keep(a.O + (long long)(first_row + (i >> 4)) * a.dim + c, vo, once_policy);
                    g = make_float4(vg.v[0], vg.v[1], vg.v[2], vg.v[3]);
                    o = make_float4(vo.v[0], vo.v[1], vo.v[2], vo.v[3]);
                }
                s_rows[i] = g;
                s_out[i] = o;
            }
            __syncthreads();
            for (int k = warp; k < a.n_rel; k += WARPS) {
                const int begin = __ldg(a.block_ptr + (long long)k * (2 * a.n_block + 1) + 2 * b + part);
                const int end = __ldg(a.block_ptr + (long long)k * (2 * a.n_block + 1) + 2 * b + part + 1);
                float *p = a.partial + ((long long)k * a.n_block + b) * a.dim + col;
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                float4 own = make_float4(0.f, 0.f, 0.f, 0.f);
                if (begin < end && active) own = __ldg(reinterpret_cast<const float4 *>(a.R + (long long)k * a.dim + col));
                const float ov[4] = {own.x, own.y, own.z, own.w};
                auto accumulate = [&](const Ids &e, float w, const Vec<float, 4> &x) {
                    const int dst = PACKED ? (int)(id_bits_of(e) & low) : id_first(e);
                    const float4 g = s_rows[(dst - first_row) * (kStagedSlab / 4) + l16];
                    const float4 o = s_out[(dst - first_row) * (kStagedSlab / 4) + l16];
                    const float gv[4] = {g.x, g.y, g.z, g.w}, out_v[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        // the same expressions as seg_gated_kernel: y = w * (x (x) own) (unit weight: no multiply)
                        const float y = a.w ? message<float, MSG>(w, x.v[v], ov[v]) : message<float, MSG>(x.v[v], ov[v]);
                        const float up = a.w ? gv[v] * w : gv[v];
                        const float term = MSG == MSG_MUL ? up * x.v[v] : up;
                        if (out_v[v] == y) acc[v] += term;
                    }
                };
                Ids ahead = Ids();
                float ahead_w = 1.f;
                if (begin + lane < end) {
                    ahead = edge_load_once(ids + begin + lane, once_policy);
                    if (a.w) ahead_w = __ldg(a.w + begin + lane);
                }
                for (int base = begin; base < end; base += 32) {
                    const int n = min(32, end - base);
                    __syncwarp();
                    s_edge[warp][lane] = ahead;
                    s_w[warp][lane] = ahead_w;
                    __syncwarp();
                    if (base + 32 + lane < end) {
                        ahead = edge_load_once(ids + base + 32 + lane, once_policy);
                        if (a.w) ahead_w = __ldg(a.w + base + 32 + lane);
                    }
                    int u = 0;
                    for (; u + 2 * kEdgesPerHalf <= n; u += 2 * kEdgesPerHalf) {
                        const int mine = u + half * kEdgesPerHalf;
                        Vec<float, 4> x[kEdgesPerHalf];
#pragma unroll
                        for (int q = 0; q < kEdgesPerHalf; ++q) {
                            const Ids e = s_edge[warp][mine + q];
                            const int src = PACKED ? (int)(id_bits_of(e) >> shift) : id_second(e);
                            gather_load_keep(reinterpret_cast<const float *>(X + (unsigned long long)(unsigned)src * row_bytes), x[q], keep_policy);
                        }
#pragma unroll
                        for (int q = 0; q < kEdgesPerHalf; ++q) accumulate(s_edge[warp][mine + q], s_w[warp][mine + q], x[q]);
                    }
                    for (; u < n; u += 2) {
                        const int mine = u + half;
                        if (mine < n) {
                            const Ids e = s_edge[warp][mine];
                            const int src = PACKED ? (int)(id_bits_of(e) >> shift) : id_second(e);
                            Vec<float, 4> x;
                            gather_load_keep(reinterpret_cast<const float *>(X + (unsigned long long)(unsigned)src * row_bytes), x, keep_policy);
                            accumulate(e, s_w[warp][mine], x);
                        }
                    }
                }
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[v] += __shfl_xor_sync(kFullMask, acc[v], 16);
                if (half == 0 && active) {
                    float4 r = make_float4(acc[0], acc[1], acc[2], acc[3]);
                    if (part == 1) {                     // this lane wrote the first half's sum: complete the run's partial row
                        const float4 before = *reinterpret_cast<const float4 *>(p);
                        r = make_float4(before.x + r.x, before.y + r.y, before.z + r.z, before.w + r.w);
                    }
                    *reinterpret_cast<float4 *>(p) = r;
                }
            }
            __syncthreads();
        }
    }
}

template <int MSG> int launch_typed(const StagedArgs &args, size_t smem, int blocks, cudaStream_t stream) {
    // every launch: the attribute belongs to the current device's copy of the function (one process may drive several)
    ULTRA_CUDA_OK(cudaFuncSetAttribute(rows_in_smem_kernel<MSG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rows_in_smem_kernel<MSG><<<blocks, kStagedThreads, smem, stream>>>(args);
    note_launch();
    return ULTRA_RSPMM_OK;
}

}  // namespace

int launch_rows_in_smem(StagedArgs args, int msg, cudaStream_t stream) {
    if (args.n_task == 0 || args.dim == 0) return ULTRA_RSPMM_OK;
    const size_t smem = (size_t)args.n_rows * kStagedSlab * sizeof(float);
    if (smem > kStagedMaxSmem || !args.packed || !args.counter) return ULTRA_RSPMM_ERR_ARG;
    args.n_slab = (int)((args.dim + kStagedSlab - 1) / kStagedSlab);
    int device = 0, sms = 0;
    ULTRA_CUDA_OK(cudaGetDevice(&device));
    ULTRA_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    // items: about 8 per CTA so that the tail is short, 32..128 tasks (1..4 per warp) each
    const long long work = (long long)args.n_task * args.n_slab;
    long long per_item = work / ((long long)sms * 8);
    per_item = per_item < kStagedWarps ? kStagedWarps : (per_item > 4 * kStagedWarps ? 4 * kStagedWarps : per_item);
    args.tasks_per_item = (int)((per_item + kStagedWarps - 1) / kStagedWarps * kStagedWarps);
    const long long items = (long long)((args.n_task + args.tasks_per_item - 1) / args.tasks_per_item) * args.n_slab;
    const int blocks = (int)(items < sms ? items : sms);
    ULTRA_CUDA_OK(cudaMemsetAsync(args.counter, 0, sizeof(unsigned), stream));
    switch (msg) {
        case MSG_MUL: return launch_typed<MSG_MUL>(args, smem, blocks, stream);
        case MSG_ADD: return launch_typed<MSG_ADD>(args, smem, blocks, stream);
        default: return launch_typed<MSG_COPY>(args, smem, blocks, stream);
    }
}

static int sm_count(int *sms) {
    int device = 0;
    ULTRA_CUDA_OK(cudaGetDevice(&device));
    ULTRA_CUDA_OK(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, device));
    return ULTRA_RSPMM_OK;
}

int launch_pairs_in_smem(PairArgs args, int msg, cudaStream_t stream) {
    if (args.n_seg == 0 || args.dim == 0) return ULTRA_RSPMM_OK;
    const size_t smem = (size_t)args.n_rows * kStagedSlab * sizeof(float);
    if (smem > kStagedMaxSmem || !args.pair || !args.counter || args.n_rel > 4) return ULTRA_RSPMM_ERR_ARG;
    args.n_slab = (int)((args.dim + kStagedSlab - 1) / kStagedSlab);
    int sms = 0;
    if (int status = sm_count(&sms)) return status;
    const long long items = (long long)((args.n_seg + kStagedWarps - 1) / kStagedWarps) * args.n_slab;
    const int blocks = (int)(items < sms ? items : sms);
    ULTRA_CUDA_OK(cudaMemsetAsync(args.counter, 0, sizeof(unsigned), stream));
#define ULTRA_PAIRS(M)                                                                                                        \
    do {                                                                                                                       \
        ULTRA_CUDA_OK(cudaFuncSetAttribute(pairs_in_smem_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        pairs_in_smem_kernel<M><<<blocks, kStagedThreads, smem, stream>>>(args);                                              \
    } while (0)
    if (msg == MSG_MUL) ULTRA_PAIRS(MSG_MUL);
    else if (msg == MSG_ADD) ULTRA_PAIRS(MSG_ADD);
    else ULTRA_PAIRS(MSG_COPY);
#undef ULTRA_PAIRS
    note_launch();
    return ULTRA_RSPMM_OK;
}

int launch_dst_blocked(BlockedRelArgs args, int msg, cudaStream_t stream) {
    if (args.n_rel == 0 || args.n_block == 0 || args.dim == 0) return ULTRA_RSPMM_OK;
    const size_t smem = (size_t)args.block_rows * kStagedSlab * sizeof(float);   // + 12.3 KB static (unpacked ids): <= 227 KB
    if (smem > kStagedMaxSmem - 8192 || !args.block_ptr || !args.counter || (msg != MSG_MUL && msg != MSG_COPY)) return ULTRA_RSPMM_ERR_ARG;
    args.n_slab = (int)((args.dim + kStagedSlab - 1) / kStagedSlab);
    int sms = 0;
    if (int status = sm_count(&sms)) return status;
    const long long items = (long long)args.n_block * args.n_slab;
    ULTRA_CUDA_OK(cudaMemsetAsync(args.counter, 0, sizeof(unsigned), stream));
    // blocks of <= 400 rows (<= 100 KB of shared memory): two CTAs of 16 warps per SM, so that one CTA's block refill and
    // end-of-item barrier overlap with the other's gathers; larger blocks: one CTA of 32 warps
    const bool two = smem <= 100 * 1024;
#define ULTRA_BLOCKED_LAUNCH(M, P, W, C)                                                                                       \
    do {                                                                                                                       \
        ULTRA_CUDA_OK(cudaFuncSetAttribute(dst_blocked_kernel<M, P, W, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        dst_blocked_kernel<M, P, W, C><<<(items < C * sms ? (int)items : C * sms), W * 32, smem, stream>>>(args);              \
    } while (0)
#define ULTRA_BLOCKED(M, P)                                                                                                   \
    do {                                                                                                                       \
        if (two) ULTRA_BLOCKED_LAUNCH(M, P, 16, 2);                                                                            \
        else ULTRA_BLOCKED_LAUNCH(M, P, 32, 1);                                                                                \
    } while (0)
    const bool packed = args.packed != nullptr && args.pack_shift > 0;
    if (msg == MSG_MUL) { if (packed) ULTRA_BLOCKED(MSG_MUL, true); else ULTRA_BLOCKED(MSG_MUL, false); }
    else { if (packed) ULTRA_BLOCKED(MSG_COPY, true); else ULTRA_BLOCKED(MSG_COPY, false); }
#undef ULTRA_BLOCKED
#undef ULTRA_BLOCKED_LAUNCH
    note_launch();
    return ULTRA_RSPMM_OK;
}

int launch_dst_blocked_gated(BlockedRelArgs args, int msg, cudaStream_t stream) {
    if (args.n_rel == 0 || args.n_block == 0 || args.dim == 0) return ULTRA_RSPMM_OK;
    const size_t smem = (size_t)(args.block_rows / 2) * 2 * kStagedSlab * sizeof(float);   // grad_output and output of half a block
    if (smem > kStagedMaxSmem - 8192 || args.block_rows % 2 || !args.block_ptr || !args.counter || !args.O || !args.R ||
        (msg != MSG_MUL && msg != MSG_ADD))
        return ULTRA_RSPMM_ERR_ARG;
    args.n_slab = (int)((args.dim + kStagedSlab - 1) / kStagedSlab);
    int sms = 0;
    if (int status = sm_count(&sms)) return status;
    const long long items = (long long)args.n_block * args.n_slab;
    ULTRA_CUDA_OK(cudaMemsetAsync(args.counter, 0, sizeof(unsigned), stream));
#define ULTRA_GATED_LAUNCH(M, P)                                                                                               \
    do {                                                                                                                       \
        ULTRA_CUDA_OK(cudaFuncSetAttribute(dst_blocked_gated_kernel<M, P, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        dst_blocked_gated_kernel<M, P, 32><<<(items < sms ? (int)items : sms), 32 * 32, smem, stream>>>(args);                 \
    } while (0)
    const bool packed = args.packed != nullptr && args.pack_shift > 0;
    if (msg == MSG_MUL) { if (packed) ULTRA_GATED_LAUNCH(MSG_MUL, true); else ULTRA_GATED_LAUNCH(MSG_MUL, false); }
    else { if (packed) ULTRA_GATED_LAUNCH(MSG_ADD, true); else ULTRA_GATED_LAUNCH(MSG_ADD, false); }
#undef ULTRA_GATED_LAUNCH
    note_launch();
    return ULTRA_RSPMM_OK;
}

}  // namespace ultra
