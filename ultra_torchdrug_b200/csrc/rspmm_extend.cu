// Optional extensions of a built graph index (ultra_rspmm_index_extend, include/ultra_rspmm.h):
//
//  * pair lists of the csr / csc order for graphs with <= 4 relation types and <= 864 nodes - the graph of relations of
//    reference ultra/rel_model.py:91-147 (h2h, t2t, h2t, t2h).  The edges of a segment that share their other endpoint
//    become one word  other | mask << id_bits;  the pair kernel (rspmm_staged.cu) then reads the other node's row once
//    for up to 4 edges.  Built with one CTA per segment (a byte table of relation masks in shared memory), a single-CTA
//    scan + rank sort of the segments by pair count (there are at most 864), and an emit pass;
//  * the destination-block table of the rel order for graphs whose gathered slabs exceed L2: the rel order is sorted by
//    (rel, dst, src), so the edges of relation k whose destination lies in block b are one contiguous range, found by a
//    binary search per (k, b).  The blocked grad_relation kernel (rspmm_staged.cu) stages the grad_output rows of a block
//    in shared memory and gathers only the input rows.
#include <cstdlib>
#include <cstring>

#include "rspmm_common.cuh"

namespace ultra {

int g_pairs = getenv("ULTRA_RSPMM_PAIRS") ? atoi(getenv("ULTRA_RSPMM_PAIRS")) : 1;        // 0 off, 1 automatic, 2 always
int g_blocked = getenv("ULTRA_RSPMM_BLOCKED") ? atoi(getenv("ULTRA_RSPMM_BLOCKED")) : 1;  // 0 off, 1 automatic, 2 always

namespace {

constexpr int kPairThreads = 128;
constexpr int kPairTable = 1024;          // >= kStagedMaxRows, one byte of relation mask per node
// destination rows per block: 768 x 256 B = 192 KB of shared memory (one CTA per SM); ULTRA_RSPMM_BLOCK_ROWS=384: two CTAs
const int kBlockRows = getenv("ULTRA_RSPMM_BLOCK_ROWS") ? (atoi(getenv("ULTRA_RSPMM_BLOCK_ROWS")) > 1 && atoi(getenv("ULTRA_RSPMM_BLOCK_ROWS")) <= 768 ? atoi(getenv("ULTRA_RSPMM_BLOCK_ROWS")) & ~1 : 768) : 768;   // even: half blocks

// relation masks of one segment: table[other] |= 1 << rel (bytes packed four to a word, set with shared-memory atomics -
// integer OR is order-independent)
__device__ void fill_mask_table(const int32_t *ptr, const int2 *edge, int seg, unsigned *table) {
    for (int i = threadIdx.x; i < kPairTable / 4; i += blockDim.x) table[i] = 0u;
    __syncthreads();
    for (int e = ptr[seg] + threadIdx.x; e < ptr[seg + 1]; e += blockDim.x) {
        const int2 id = edge[e];
        atomicOr(&table[id.x >> 2], (1u << id.y) << (8 * (id.x & 3)));
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kPairThreads) pair_count_kernel(const int32_t *__restrict__ ptr, const int2 *__restrict__ edge,
                                                                 int32_t *__restrict__ count) {
    __shared__ unsigned table[kPairTable / 4];
    __shared__ int total;
    if (threadIdx.x == 0) total = 0;
    fill_mask_table(ptr, edge, blockIdx.x, table);
    int mine = 0;
    for (int i = threadIdx.x; i < kPairTable / 4; i += blockDim.x) {
        const unsigned w = table[i];
        mine += ((w & 0xffu) != 0) + ((w & 0xff00u) != 0) + ((w & 0xff0000u) != 0) + ((w & 0xff000000u) != 0);
    }
    atomicAdd(&total, mine);
    __syncthreads();
    if (threadIdx.x == 0) count[blockIdx.x] = total;
}

// single CTA: exclusive scan of the pair counts and the segments ranked by descending count (ties by segment id)
__global__ void __launch_bounds__(1024) pair_layout_kernel(const int32_t *__restrict__ count, int n_seg, int32_t *__restrict__ ptr,
                                                           int32_t *__restrict__ rows, int32_t *__restrict__ total) {
    __shared__ int32_t c[kPairTable];
    for (int i = threadIdx.x; i < n_seg; i += blockDim.x) c[i] = count[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        int32_t run = 0;
        for (int i = 0; i < n_seg; ++i) { ptr[i] = run; run += c[i]; }
        ptr[n_seg] = run;
        *total = run;
    }
    for (int i = threadIdx.x; i < n_seg; i += blockDim.x) {
        int rank = 0;
        for (int j = 0; j < n_seg; ++j) rank += c[j] > c[i] || (c[j] == c[i] && j < i);
        rows[rank] = i;
    }
}

__global__ void __launch_bounds__(kPairThreads) pair_emit_kernel(const int32_t *__restrict__ ptr, const int2 *__restrict__ edge,
                                                                const int32_t *__restrict__ pair_ptr, int id_bits,
                                                                uint32_t *__restrict__ pair) {
    __shared__ unsigned table[kPairTable / 4];
    __shared__ int offset[kPairThreads + 1];
    fill_mask_table(ptr, edge, blockIdx.x, table);
    // thread t owns nodes [8 t, 8 t + 8): count, exclusive scan over the CTA, write in ascending node order
    constexpr int kPerThread = kPairTable / kPairThreads;
    const unsigned char *bytes = reinterpret_cast<const unsigned char *>(table);
    int mine = 0;
    for (int q = 0; q < kPerThread; ++q) mine += bytes[threadIdx.x * kPerThread + q] != 0;
    offset[threadIdx.x + 1] = mine;
    if (threadIdx.x == 0) offset[0] = 0;
    __syncthreads();
    if (threadIdx.x == 0)
        for (int i = 1; i <= kPairThreads; ++i) offset[i] += offset[i - 1];
    __syncthreads();
    int at = pair_ptr[blockIdx.x] + offset[threadIdx.x];
    for (int q = 0; q < kPerThread; ++q) {
        const int node = threadIdx.x * kPerThread + q;
        const unsigned mask = bytes[node];
        if (mask) pair[at++] = (unsigned)node | (mask << id_bits);
    }
}

// block_ptr[k * (2 n_block + 1) + h] = first position in [ptr[k], ptr[k + 1]) of the rel order whose destination >= h * rows / 2:
// half-block granularity - the sum pass takes blocks (entries 2 b and 2 b + 2), the gated pass, which stages two rows per
// destination, half blocks
__global__ void block_ptr_kernel(const int32_t *__restrict__ ptr, const int2 *__restrict__ edge, int n_rel, int n_block,
                                 int block_rows, int32_t *__restrict__ block_ptr, int4 *__restrict__ split) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int per_rel = 2 * n_block + 1;
    if (i >= (long long)n_rel * per_rel) return;
    const int k = (int)(i / per_rel), h = (int)(i - (long long)k * per_rel);
    int lo = ptr[k], hi = ptr[k + 1];
    const long long bound = (long long)h * (block_rows / 2);
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (edge[mid].x < bound) lo = mid + 1;
        else hi = mid;
    }
    block_ptr[i] = lo;
    if (h == 0) split[k] = make_int4(k, k * n_block, n_block, 0);
}

struct ExtendLayout {
    bool pairs, blocks;
    int n_block;
    size_t count, total_slot, pair_ptr[2], pair_rows[2], pair[2], block_ptr, block_split, total;
};

ExtendLayout extend_layout(const ultra_rspmm_index_t &ix) {
    ExtendLayout L = {};
    L.pairs = g_pairs != 0 && ix.dtype == ULTRA_RSPMM_F32 && ix.unit_weight && ix.nnz > 0 && ix.n_rel <= 4 &&
              ix.n_out <= kStagedMaxRows && ix.n_in <= kStagedMaxRows;
    const long long slab_bytes = ((long long)ix.n_out + ix.n_in) * 512;
    L.blocks = g_blocked != 0 && ix.dtype == ULTRA_RSPMM_F32 && ix.nnz > 0 && ix.n_rel > 0 &&
               (g_blocked == 2 || slab_bytes > (96ll << 20));
    L.n_block = L.blocks ? (ix.n_out + kBlockRows - 1) / kBlockRows : 0;
    size_t at = 0;
    if (L.pairs) {
        L.count = at; at = align_up(at + 4 * (size_t)kPairTable);
        L.total_slot = at; at = align_up(at + 8);
        const int32_t n_seg[2] = {ix.n_out, ix.n_in};
        for (int o = 0; o < 2; ++o) {
            L.pair_ptr[o] = at; at = align_up(at + 4 * ((size_t)n_seg[o] + 1));
            L.pair_rows[o] = at; at = align_up(at + 4 * ((size_t)n_seg[o] + 1));
            L.pair[o] = at; at = align_up(at + 4 * (size_t)ix.nnz);
        }
    }
    if (L.blocks) {
        L.block_ptr = at; at = align_up(at + 4 * (size_t)ix.n_rel * (2 * (size_t)L.n_block + 1));
        L.block_split = at; at = align_up(at + sizeof(int4) * (size_t)ix.n_rel);
    }
    L.total = at;
    return L;
}

}  // namespace
}  // namespace ultra

using namespace ultra;

extern "C" int ultra_rspmm_index_extend_bytes(const ultra_rspmm_index_t *index, size_t *buffer_bytes) {
    if (!index || !buffer_bytes) return ULTRA_RSPMM_ERR_ARG;
    *buffer_bytes = extend_layout(*index).total;
    return ULTRA_RSPMM_OK;
}

extern "C" int ultra_rspmm_index_extend(ultra_rspmm_index_t *index, void *buffer, size_t buffer_bytes, void *stream_) {
    if (!index) return ULTRA_RSPMM_ERR_ARG;
    const ExtendLayout L = extend_layout(*index);
    memset(index->pairs, 0, sizeof(index->pairs));
    index->block_ptr = nullptr;
    index->block_split = nullptr;
    index->block_rows = index->n_block = 0;
    if (L.total == 0) return ULTRA_RSPMM_OK;
    if (!buffer || buffer_bytes < L.total) return ULTRA_RSPMM_ERR_WORKSPACE;
    if ((uintptr_t)buffer & 255) return ULTRA_RSPMM_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    char *base = (char *)buffer;
    if (L.pairs) {
        const ultra_rspmm_order_t *orders[2] = {&index->csr, &index->csc};
        int id_bits = 1;
        while ((1 << id_bits) < kPairTable) ++id_bits;
        for (int o = 0; o < 2; ++o) {
            const ultra_rspmm_order_t &order = *orders[o];
            if (order.n_seg <= 0) continue;
            int32_t *count = (int32_t *)(base + L.count), *total = (int32_t *)(base + L.total_slot);
            int32_t *pair_ptr = (int32_t *)(base + L.pair_ptr[o]), *rows = (int32_t *)(base + L.pair_rows[o]);
            uint32_t *pair = (uint32_t *)(base + L.pair[o]);
            pair_count_kernel<<<order.n_seg, kPairThreads, 0, stream>>>(order.ptr, (const int2 *)order.edge, count);
            note_launch();
            pair_layout_kernel<<<1, 1024, 0, stream>>>(count, order.n_seg, pair_ptr, rows, total);
            note_launch();
            pair_emit_kernel<<<order.n_seg, kPairThreads, 0, stream>>>(order.ptr, (const int2 *)order.edge, pair_ptr, id_bits, pair);
            note_launch();
            int32_t n_pair = 0;
            ULTRA_CUDA_OK(cudaMemcpyAsync(&n_pair, total, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
            ULTRA_CUDA_OK(cudaStreamSynchronize(stream));
            // worth it when the pairs merge enough edges: one row read + 16 predicated FMAs per pair against one row read
            // + 4 FMAs per edge
            if (n_pair > 0 && (g_pairs == 2 || 2ll * index->nnz >= 3ll * n_pair)) {
                index->pairs[o].n_pair = n_pair;
                index->pairs[o].id_bits = id_bits;
                index->pairs[o].ptr = pair_ptr;
                index->pairs[o].pair = pair;
                index->pairs[o].rows = rows;
            }
        }
    }
    if (L.blocks) {
        int32_t *block_ptr = (int32_t *)(base + L.block_ptr);
        int4 *split = (int4 *)(base + L.block_split);
        const long long n = (long long)index->n_rel * (2 * L.n_block + 1);
        block_ptr_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(index->rel.ptr, (const int2 *)index->rel.edge, index->n_rel,
                                                                        L.n_block, kBlockRows, block_ptr, split);
        note_launch();
        index->block_ptr = block_ptr;
        index->block_split = (const int32_t *)split;
        index->block_rows = kBlockRows;
        index->n_block = L.n_block;
    }
    ULTRA_CUDA_OK(cudaGetLastError());
    ULTRA_CUDA_OK(cudaStreamSynchronize(stream));
    return ULTRA_RSPMM_OK;
}

extern "C" int ultra_rspmm_set_extensions(int32_t pairs, int32_t blocked) {
    if (pairs < 0 || pairs > 2 || blocked < 0 || blocked > 2) return ULTRA_RSPMM_ERR_ARG;
    g_pairs = pairs;
    g_blocked = blocked;
    return ULTRA_RSPMM_OK;
}
