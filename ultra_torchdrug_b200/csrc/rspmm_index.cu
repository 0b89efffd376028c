// Graph index build: un-coalesced int64 COO  ->  three int32 edge orders + task lists.
//
// Replaces what the reference redoes on every operator call: `graph.adjacency.transpose(0,1)`
// (reference ultra/layer.py:127,328), `sparse.coalesce()` and torchdrug's `coo2csr3d`
// (SURVEY.md section 8 row a5).  Done once per distinct edge set and cached by the caller.
//
//   1. key = (row * n_rel + rel) * n_in + col  (row = node_out, col = node_in), stable radix sort
//   2. runs of equal keys are merged, values summed in fp64   ->  CSR order (dst, rel, src): the edges of a
//      destination are grouped by relation so that the kernel can keep a relation row in registers across
//      consecutive edges; eid = the edge's position in canonical coalesce() order (dst, src, rel)
//   3. second sort by (col, rel, row)  -> CSC order (backward w.r.t. input), same grouping
//   4. third, stable sort by rel        -> relation order (backward w.r.t. relation)
//   5. per order: segment pointers by binary search, tasks of <= chunk edges, split-segment slots,
//      tasks sorted longest first
// Sorting and scans use CUB (library code, like calling cuBLAS for a plain GEMM); this is not the
// measured hot path.
#include <cub/cub.cuh>

#include "rspmm_common.cuh"

namespace ultra {

namespace {

constexpr int kBuildThreads = 256;
inline int blocks_for(int64_t n) { return (int)((n + kBuildThreads - 1) / kBuildThreads); }

enum Counter {
    CNT_ERROR = 0,    // an index was out of range
    CNT_NONUNIT = 1,  // a merged value differs from 1
    CNT_NNZ = 2,      // merged edge count
    CNT_MAXSEG = 3,   // +order: longest segment
    CNT_TOTALS = 8,   // +3*order: n_task, n_slot, n_split
    CNT_GTASKS = 20,  // +order: n_gtask
    CNT_SIZE = 32
};

__global__ void make_keys_kernel(const int64_t *__restrict__ indices, int64_t stride, int64_t nnz, int64_t n_out,
                                 int64_t n_in, int64_t n_rel, unsigned long long *__restrict__ keys,
                                 int32_t *__restrict__ vals, int32_t *__restrict__ counters) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    const int64_t r = indices[e], c = indices[stride + e], k = indices[2 * stride + e];
    if (r < 0 || r >= n_out || c < 0 || c >= n_in || k < 0 || k >= n_rel) {
        atomicOr(&counters[CNT_ERROR], 1);
        keys[e] = 0;
    } else {
        keys[e] = ((unsigned long long)r * n_rel + k) * n_in + c;
    }
    vals[e] = (int32_t)e;
}

__global__ void head_flags_kernel(const unsigned long long *__restrict__ keys, int64_t nnz, int32_t *__restrict__ head) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e > nnz) return;
    head[e] = (e < nnz && (e == 0 || keys[e] != keys[e - 1])) ? 1 : 0;
}

// one thread per run head: sums the run's values (fp64 accumulate), decodes the key, emits the CSR edge
template <typename T>
__global__ void merge_kernel(const unsigned long long *__restrict__ keys, const int32_t *__restrict__ vals,
                             const int32_t *__restrict__ pos, const T *__restrict__ values, int64_t nnz,
                             int64_t n_in, int64_t n_rel, int2 *__restrict__ csr_edge, T *__restrict__ csr_w,
                             int32_t *__restrict__ row_of, int32_t *__restrict__ merge_start,
                             int32_t *__restrict__ counters) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    const unsigned long long key = keys[e];
    if (e > 0 && keys[e - 1] == key) return;
    double total = 0.0;
    for (int64_t q = e; q < nnz && keys[q] == key; ++q) total += (double)values[vals[q]];
    const T w = (T)total;
    const int32_t m = pos[e];
    const unsigned long long rk = key / (unsigned long long)n_in;
    csr_edge[m] = make_int2((int32_t)(key % (unsigned long long)n_in), (int32_t)(rk % (unsigned long long)n_rel));
    csr_w[m] = w;
    row_of[m] = (int32_t)(rk / (unsigned long long)n_rel);
    merge_start[m] = (int32_t)e;          // csr edge m = sum of the sorted raw edges [merge_start[m], merge_start[m + 1])
    if (w != T(1)) atomicOr(&counters[CNT_NONUNIT], 1);
    if (e == 0) {
        counters[CNT_NNZ] = pos[nnz];
        merge_start[pos[nnz]] = (int32_t)nnz;
    }
}

__global__ void csc_keys_kernel(const int2 *__restrict__ csr_edge, const int32_t *__restrict__ row_of, int32_t nnz,
                                int64_t n_out, int64_t n_rel, unsigned long long *__restrict__ keys,
                                int32_t *__restrict__ vals) {
    const int32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nnz) return;
    const int2 e = csr_edge[m];
    keys[m] = ((unsigned long long)e.x * n_rel + e.y) * n_out + row_of[m];
    vals[m] = m;
}

// canonical coalesce() order (row, col, rel): used only to number the edges (eid)
__global__ void canonical_keys_kernel(const int2 *__restrict__ csr_edge, const int32_t *__restrict__ row_of, int32_t nnz,
                                      int64_t n_in, int64_t n_rel, unsigned long long *__restrict__ keys,
                                      int32_t *__restrict__ vals) {
    const int32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nnz) return;
    const int2 e = csr_edge[m];
    keys[m] = ((unsigned long long)row_of[m] * n_in + e.x) * n_rel + e.y;
    vals[m] = m;
}

__global__ void rank_scatter_kernel(const int32_t *__restrict__ sorted_positions, int32_t nnz, int32_t *__restrict__ eid) {
    const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nnz) eid[sorted_positions[r]] = r;
}

__global__ void rel_keys_kernel(const int2 *__restrict__ csr_edge, int32_t nnz, unsigned long long *__restrict__ keys,
                                int32_t *__restrict__ vals) {
    const int32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nnz) return;
    keys[m] = (unsigned long long)csr_edge[m].y;
    vals[m] = m;
}

// permute the coalesced edges into another order; MODE 0: CSC {dst, rel} / segment = src,
// MODE 1: relation order {dst, src} / segment = rel
template <typename T, int MODE>
__global__ void permute_kernel(const int32_t *__restrict__ perm, int32_t nnz, const int2 *__restrict__ csr_edge,
                               const T *__restrict__ csr_w, const int32_t *__restrict__ row_of,
                               const int32_t *__restrict__ csr_eid, int2 *__restrict__ edge, T *__restrict__ w,
                               int32_t *__restrict__ eid, int32_t *__restrict__ seg_of) {
    const int32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    const int32_t m = perm[p];
    const int2 e = csr_edge[m];
    const int32_t r = row_of[m];
    edge[p] = MODE == 0 ? make_int2(r, e.y) : make_int2(r, e.x);
    seg_of[p] = MODE == 0 ? e.x : e.y;
    w[p] = csr_w[m];
    eid[p] = csr_eid[m];
}

// both ids of an edge in one word, and a 0/1 flag per edge for "weight differs from 1" (n + 1 entries, last = 0)
template <typename T>
__global__ void pack_kernel(const int2 *__restrict__ edge, const T *__restrict__ w, int32_t nnz, int shift,
                            uint32_t *__restrict__ packed, int32_t *__restrict__ nonunit) {
    const int32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > nnz) return;
    if (p == nnz) {
        nonunit[p] = 0;
        return;
    }
    const int2 e = edge[p];
    if (shift > 0) packed[p] = (uint32_t)e.x | ((uint32_t)e.y << shift);
    nonunit[p] = w[p] != T(1) ? 1 : 0;
}

// ptr[s] = first position whose segment id is >= s   (s in [0, n_seg])
__global__ void segment_ptr_kernel(const int32_t *__restrict__ seg_of, int32_t nnz, int32_t n_seg,
                                   int32_t *__restrict__ ptr) {
    const int32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > n_seg) return;
    int32_t lo = 0, hi = nnz;
    while (lo < hi) {
        const int32_t mid = (lo + hi) >> 1;
        if (seg_of[mid] < s) lo = mid + 1; else hi = mid;
    }
    ptr[s] = lo;
}

__global__ void task_count_kernel(const int32_t *__restrict__ ptr, int32_t n_seg, int32_t chunk,
                                  int32_t *__restrict__ cnt_task, int32_t *__restrict__ cnt_slot,
                                  int32_t *__restrict__ cnt_split, int32_t *__restrict__ max_seg) {
    const int32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > n_seg) return;
    if (s == n_seg) {
        cnt_task[s] = cnt_slot[s] = cnt_split[s] = 0;
        return;
    }
    const int32_t deg = ptr[s + 1] - ptr[s];
    const int32_t c = deg <= chunk ? 1 : (deg + chunk - 1) / chunk;
    cnt_task[s] = c;
    cnt_slot[s] = c > 1 ? c : 0;
    cnt_split[s] = c > 1 ? 1 : 0;
    atomicMax(max_seg, deg);
}

__global__ void gather_totals_kernel(const int32_t *__restrict__ off_task, const int32_t *__restrict__ off_slot,
                                     const int32_t *__restrict__ off_split, int32_t n_seg,
                                     int32_t *__restrict__ totals) {
    totals[0] = off_task[n_seg];
    totals[1] = off_slot[n_seg];
    totals[2] = off_split[n_seg];
}

__global__ void task_emit_kernel(const int32_t *__restrict__ ptr, int32_t n_seg, int32_t chunk,
                                 const int32_t *__restrict__ off_task, const int32_t *__restrict__ off_slot,
                                 const int32_t *__restrict__ off_split, const int32_t *__restrict__ nonunit_before,
                                 int4 *__restrict__ task,
                                 uint32_t *__restrict__ task_key, int32_t *__restrict__ task_val,
                                 int4 *__restrict__ split) {
    const int32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    const int32_t begin = ptr[s], end = ptr[s + 1];
    const int32_t deg = end - begin;
    const int32_t t0 = off_task[s];
    auto encode = [&](int32_t slot, int32_t from, int32_t to) {
        const bool nonunit = nonunit_before && nonunit_before[to] != nonunit_before[from];
        return (slot + 1) | (nonunit ? kNonUnitTask : 0);
    };
    if (deg <= chunk) {
        task[t0] = make_int4(s, begin, end, encode(-1, begin, end));
        task_key[t0] = (uint32_t)(chunk - deg);
        task_val[t0] = t0;
        return;
    }
    const int32_t c = (deg + chunk - 1) / chunk;
    const int32_t slot0 = off_slot[s];
    // equal-sized chunks (sizes differ by at most one) so that no task of a split row is tiny
    const int32_t base = deg / c, extra = deg % c;
    int32_t at = begin;
    for (int32_t q = 0; q < c; ++q) {
        const int32_t len = base + (q < extra ? 1 : 0);
        task[t0 + q] = make_int4(s, at, at + len, encode(slot0 + q, at, at + len));
        task_key[t0 + q] = (uint32_t)(chunk - len);
        task_val[t0 + q] = t0 + q;
        at += len;
    }
    split[off_split[s]] = make_int4(s, slot0, c, 0);
}

// ---- grouped task list: consecutive short segments share one task -------------------------------------------------
// A segment of at most `group` edges is "short".  Consecutive short segments whose first edges fall into the same
// window of `group` edges (and the same block of kGroupRows segments) form one group task; every other segment
// gets the tasks it has in the plain list.
__device__ __forceinline__ bool group_head(const int32_t *__restrict__ ptr, int32_t s, int32_t group) {
    if (s == 0) return true;
    if (ptr[s] - ptr[s - 1] > group) return true;                 // the previous segment is not short
    return ptr[s] / group != ptr[s - 1] / group || s / kGroupRows != (s - 1) / kGroupRows;
}

__global__ void group_count_kernel(const int32_t *__restrict__ ptr, int32_t n_seg, int32_t chunk, int32_t group,
                                   int32_t *__restrict__ cnt_gtask) {
    const int32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > n_seg) return;
    if (s == n_seg) {
        cnt_gtask[s] = 0;
        return;
    }
    const int32_t deg = ptr[s + 1] - ptr[s];
    if (deg > group) cnt_gtask[s] = deg <= chunk ? 1 : (deg + chunk - 1) / chunk;
    else cnt_gtask[s] = group_head(ptr, s, group) ? 1 : 0;
}

__global__ void group_emit_kernel(const int32_t *__restrict__ ptr, int32_t n_seg, int32_t chunk, int32_t group,
                                  const int32_t *__restrict__ off_gtask, const int32_t *__restrict__ off_slot,
                                  const int32_t *__restrict__ nonunit_before, int4 *__restrict__ task,
                                  uint32_t *__restrict__ task_key, int32_t *__restrict__ task_val) {
    const int32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    const int32_t begin = ptr[s], deg = ptr[s + 1] - begin;
    const int32_t t0 = off_gtask[s];
    auto nonunit = [&](int32_t from, int32_t to) {
        return nonunit_before && nonunit_before[to] != nonunit_before[from] ? kNonUnitTask : 0;
    };
    if (deg <= group) {
        if (!group_head(ptr, s, group)) return;
        int32_t rows = 1;
        while (rows < kGroupRows && s + rows < n_seg && ptr[s + rows + 1] - ptr[s + rows] <= group &&
               !group_head(ptr, s + rows, group))
            ++rows;
        const int32_t end = ptr[s + rows];
        task[t0] = make_int4(s, begin, end, kGroupTask | ((rows - 1) << 24) | nonunit(begin, end));
        task_key[t0] = (uint32_t)(chunk - (end - begin));
        task_val[t0] = t0;
        return;
    }
    const int32_t end = begin + deg;
    if (deg <= chunk) {
        task[t0] = make_int4(s, begin, end, 0 | nonunit(begin, end));
        task_key[t0] = (uint32_t)(chunk - deg);
        task_val[t0] = t0;
        return;
    }
    const int32_t c = (deg + chunk - 1) / chunk, slot0 = off_slot[s];
    const int32_t base = deg / c, extra = deg % c;
    int32_t at = begin;
    for (int32_t q = 0; q < c; ++q) {
        const int32_t len = base + (q < extra ? 1 : 0);
        task[t0 + q] = make_int4(s, at, at + len, (slot0 + q + 1) | nonunit(at, at + len));
        task_key[t0 + q] = (uint32_t)(chunk - len);
        task_val[t0 + q] = t0 + q;
        at += len;
    }
}

__global__ void task_gather_kernel(const int4 *__restrict__ task_in, const int32_t *__restrict__ order, int32_t n_task,
                                   int4 *__restrict__ task_out) {
    const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_task) task_out[t] = task_in[order[t]];
}

// FNV-style 2 x 64-bit fingerprint; order-sensitive (position mixed into each term), combined with
// integer atomics so that the result does not depend on scheduling.
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

template <typename T>
__global__ void fingerprint_kernel(const int64_t *__restrict__ indices, int64_t stride, const T *__restrict__ values,
                                   int64_t nnz, unsigned long long *__restrict__ out) {
    unsigned long long a = 0, b = 0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long h = mix64((unsigned long long)e + 0x9e3779b97f4a7c15ULL);
        h = mix64(h ^ (unsigned long long)indices[e]);
        h = mix64(h ^ ((unsigned long long)indices[stride + e] << 1));
        h = mix64(h ^ ((unsigned long long)indices[2 * stride + e] << 2));
        unsigned long long bits = 0;
        const T v = values[e];
        memcpy(&bits, &v, sizeof(T));
        h = mix64(h ^ bits);
        a += h;
        b ^= mix64(h + 0x632be59bd9b4e019ULL);
    }
    for (int off = 16; off; off >>= 1) {
        a += __shfl_xor_sync(kFullMask, a, off);
        b ^= __shfl_xor_sync(kFullMask, b, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&out[0], a);
        atomicXor(&out[1], b);
    }
}

// Stable radix sort of (key, value) pairs from the `in` buffers into the `out` buffers.  Uses CUB's DoubleBuffer
// form (the plain form allocates a second copy of both arrays inside its temporary storage: 12 bytes per edge);
// the result is copied over when the last pass happened to land in the `in` buffers.
template <typename K, typename V>
int sort_pairs(void *tmp, size_t tmp_bytes, K *keys_in, K *keys_out, V *vals_in, V *vals_out, int n, int bits,
               cudaStream_t stream) {
    cub::DoubleBuffer<K> keys(keys_in, keys_out);
    cub::DoubleBuffer<V> vals(vals_in, vals_out);
    size_t need = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, need, keys, vals, n, 0, bits > 0 ? bits : 1, stream);
    if (need > tmp_bytes) return ULTRA_RSPMM_ERR_WORKSPACE;
    ULTRA_CUDA_OK(cub::DeviceRadixSort::SortPairs(tmp, need, keys, vals, n, 0, bits > 0 ? bits : 1, stream));
    note_launch();
    if (keys.Current() != keys_out)
        ULTRA_CUDA_OK(cudaMemcpyAsync(keys_out, keys.Current(), sizeof(K) * (size_t)n, cudaMemcpyDeviceToDevice, stream));
    if (vals.Current() != vals_out)
        ULTRA_CUDA_OK(cudaMemcpyAsync(vals_out, vals.Current(), sizeof(V) * (size_t)n, cudaMemcpyDeviceToDevice, stream));
    return ULTRA_RSPMM_OK;
}

int bit_length(unsigned long long v) {
    int bits = 0;
    while (v) { ++bits; v >>= 1; }
    return bits;
}

struct OrderLayout {
    size_t ptr, edge, w, eid, packed, task, split, gtask;
};

struct IndexLayout {
    OrderLayout order[3];
    size_t merge_perm, merge_start;   // which raw edges were summed into which csr edge (ultra_rspmm_index_derive)
    size_t total;
};

struct ScratchLayout {
    size_t keys_a, keys_b, vals_a, vals_b, pos, row_of, seg_of, cnt, off, task_tmp, tkey_a, tkey_b, tval_a, tval_b,
        nonunit, counters, cub, cub_bytes, total;
};

int64_t task_upper(int64_t nnz_raw, int64_t n_seg, int chunk) { return n_seg + nnz_raw / chunk + 2; }
int64_t split_upper(int64_t nnz_raw, int chunk) { return nnz_raw / chunk + 2; }

IndexLayout index_layout(int64_t nnz_raw, const int32_t n_seg[3], size_t elem, int chunk) {
    IndexLayout L;
    size_t at = 0;
    const int64_t e = nnz_raw > 0 ? nnz_raw : 1;
    for (int o = 0; o < 3; ++o) {
        OrderLayout &q = L.order[o];
        q.ptr = at; at = align_up(at + sizeof(int32_t) * ((size_t)n_seg[o] + 1));
        q.edge = at; at = align_up(at + sizeof(int2) * e);
        q.w = at; at = align_up(at + elem * e);
        q.eid = at; at = align_up(at + sizeof(int32_t) * e);
        q.packed = at; at = align_up(at + sizeof(uint32_t) * e);
        q.task = at; at = align_up(at + sizeof(int4) * task_upper(nnz_raw, n_seg[o], chunk));
        q.split = at; at = align_up(at + sizeof(int4) * split_upper(nnz_raw, chunk));
        q.gtask = at; at = align_up(at + (o == 2 ? 0 : sizeof(int4) * task_upper(nnz_raw, n_seg[o], chunk)));
    }
    L.merge_perm = at; at = align_up(at + sizeof(int32_t) * e);
    L.merge_start = at; at = align_up(at + sizeof(int32_t) * (e + 1));
    L.total = at;
    return L;
}

ScratchLayout scratch_layout(int64_t nnz_raw, int32_t n_seg_max, int chunk) {
    ScratchLayout S;
    size_t at = 0;
    const size_t e = nnz_raw > 0 ? nnz_raw : 1;
    const size_t nt = task_upper(nnz_raw, n_seg_max, chunk);
    S.keys_a = at; at = align_up(at + 8 * e);
    S.keys_b = at; at = align_up(at + 8 * e);
    S.vals_a = at; at = align_up(at + 4 * (e + 1));  // doubles as the head-flag array (nnz + 1)
    S.vals_b = at; at = align_up(at + 4 * e);
    S.pos = at; at = align_up(at + 4 * (e + 1));
    S.row_of = at; at = align_up(at + 4 * e);
    S.seg_of = at; at = align_up(at + 4 * e);
    S.cnt = at; at = align_up(at + 4 * 4 * ((size_t)n_seg_max + 1));
    S.off = at; at = align_up(at + 12 * 4 * ((size_t)n_seg_max + 1));
    S.task_tmp = at; at = align_up(at + 16 * nt);
    S.tkey_a = at; at = align_up(at + 4 * nt);
    S.tkey_b = at; at = align_up(at + 4 * nt);
    S.tval_a = at; at = align_up(at + 4 * nt);
    S.tval_b = at; at = align_up(at + 4 * nt);
    S.nonunit = at; at = align_up(at + 3 * 4 * (e + 1));
    S.counters = at; at = align_up(at + 4 * CNT_SIZE);
    S.cub = at;
    // CUB temporary storage: radix sort / scan need O(tiles) words; provision generously and verify at build time
    S.cub_bytes = align_up((size_t)(16u << 20) + e + nt);   // radix-sort histograms / scan tile states: O(items / tile)
    at += S.cub_bytes;
    S.total = at;
    return S;
}

template <typename T>
int build_typed(const int64_t *dev_indices, int64_t stride, const T *dev_values, int64_t nnz_raw, int32_t n_out,
                int32_t n_in, int32_t n_rel, char *ibuf, char *sbuf, ultra_rspmm_index_t *index, cudaStream_t stream) {
    const int chunk = g_chunk;
    const int group = g_group_edges < 0 ? chunk / 4 : (g_group_edges > chunk / 2 ? chunk / 2 : g_group_edges);
    const int32_t n_seg[3] = {n_out, n_in, n_rel};
    const int32_t n_seg_max = n_out > n_in ? (n_out > n_rel ? n_out : n_rel) : (n_in > n_rel ? n_in : n_rel);
    const IndexLayout L = index_layout(nnz_raw, n_seg, sizeof(T), chunk);
    const ScratchLayout S = scratch_layout(nnz_raw, n_seg_max, chunk);

    unsigned long long *keys_a = (unsigned long long *)(sbuf + S.keys_a), *keys_b = (unsigned long long *)(sbuf + S.keys_b);
    int32_t *vals_a = (int32_t *)(sbuf + S.vals_a), *vals_b = (int32_t *)(sbuf + S.vals_b);
    int32_t *pos = (int32_t *)(sbuf + S.pos), *row_of = (int32_t *)(sbuf + S.row_of), *seg_of = (int32_t *)(sbuf + S.seg_of);
    int32_t *counters = (int32_t *)(sbuf + S.counters);
    void *cub_tmp = sbuf + S.cub;
    size_t need = 0;

    int2 *csr_edge = (int2 *)(ibuf + L.order[0].edge);
    T *csr_w = (T *)(ibuf + L.order[0].w);

    ULTRA_CUDA_OK(cudaMemsetAsync(counters, 0, 4 * CNT_SIZE, stream));
    int32_t host_counters[CNT_SIZE] = {0};
    int32_t nnz = 0;

    if (nnz_raw > 0) {
        make_keys_kernel<<<blocks_for(nnz_raw), kBuildThreads, 0, stream>>>(dev_indices, stride, nnz_raw, n_out, n_in,
                                                                           n_rel, keys_a, vals_a, counters);
        note_launch();
        const unsigned long long span = (unsigned long long)n_out * (unsigned long long)n_in * (unsigned long long)n_rel;
        const int bits = bit_length(span > 0 ? span - 1 : 0);
        if (int status = sort_pairs(cub_tmp, S.cub_bytes, keys_a, keys_b, vals_a, vals_b, (int)nnz_raw, bits, stream))
            return status;
        head_flags_kernel<<<blocks_for(nnz_raw + 1), kBuildThreads, 0, stream>>>(keys_b, nnz_raw, vals_a);
        note_launch();
        need = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, need, vals_a, pos, (int)nnz_raw + 1, stream);
        if (need > S.cub_bytes) return ULTRA_RSPMM_ERR_WORKSPACE;
        need = S.cub_bytes;
        ULTRA_CUDA_OK(cub::DeviceScan::ExclusiveSum(cub_tmp, need, vals_a, pos, (int)nnz_raw + 1, stream));
        note_launch();
        merge_kernel<T><<<blocks_for(nnz_raw), kBuildThreads, 0, stream>>>(keys_b, vals_b, pos, dev_values, nnz_raw, n_in,
                                                                          n_rel, csr_edge, csr_w, row_of,
                                                                          (int32_t *)(ibuf + L.merge_start), counters);
        note_launch();
        ULTRA_CUDA_OK(cudaMemcpyAsync(ibuf + L.merge_perm, vals_b, sizeof(int32_t) * (size_t)nnz_raw, cudaMemcpyDeviceToDevice, stream));
        ULTRA_CUDA_OK(cudaMemcpyAsync(host_counters, counters, 4 * CNT_SIZE, cudaMemcpyDeviceToHost, stream));
        ULTRA_CUDA_OK(cudaStreamSynchronize(stream));
        if (host_counters[CNT_ERROR]) return ULTRA_RSPMM_ERR_INDEX;
        nnz = host_counters[CNT_NNZ];
    }

    // ---- canonical numbering of the coalesced edges (arg-index contract: position in coalesce() order) ---------
    int32_t *csr_eid = (int32_t *)(ibuf + L.order[0].eid);
    if (nnz > 0) {
        canonical_keys_kernel<<<blocks_for(nnz), kBuildThreads, 0, stream>>>(csr_edge, row_of, nnz, n_in, n_rel, keys_a, vals_a);
        note_launch();
        const unsigned long long span = (unsigned long long)n_out * (unsigned long long)n_in * (unsigned long long)n_rel;
        const int bits = bit_length(span > 0 ? span - 1 : 0);
        if (int status = sort_pairs(cub_tmp, S.cub_bytes, keys_a, keys_b, vals_a, vals_b, nnz, bits, stream)) return status;
        rank_scatter_kernel<<<blocks_for(nnz), kBuildThreads, 0, stream>>>(vals_b, nnz, csr_eid);
        note_launch();
    }

    // ---- the two other orders -------------------------------------------------------------------
    int pack_shift[3] = {0, 0, 0};
    const bool any_nonunit = host_counters[CNT_NONUNIT] != 0;
    for (int o = 0; o < 3; ++o) {
        int32_t *ptr = (int32_t *)(ibuf + L.order[o].ptr);
        const int32_t *segments = row_of;
        if (o > 0 && nnz > 0) {
            int2 *edge = (int2 *)(ibuf + L.order[o].edge);
            T *w = (T *)(ibuf + L.order[o].w);
            int32_t *eid = (int32_t *)(ibuf + L.order[o].eid);
            int bits;
            if (o == 1) {
                csc_keys_kernel<<<blocks_for(nnz), kBuildThreads, 0, stream>>>(csr_edge, row_of, nnz, n_out, n_rel, keys_a, vals_a);
                const unsigned long long span = (unsigned long long)n_out * (unsigned long long)n_in * (unsigned long long)n_rel;
                bits = bit_length(span > 0 ? span - 1 : 0);
            } else {
                rel_keys_kernel<<<blocks_for(nnz), kBuildThreads, 0, stream>>>(csr_edge, nnz, keys_a, vals_a);
                bits = bit_length(n_rel > 0 ? (unsigned long long)n_rel - 1 : 0);
            }
            note_launch();
            if (int status = sort_pairs(cub_tmp, S.cub_bytes, keys_a, keys_b, vals_a, vals_b, nnz, bits, stream)) return status;
            if (o == 1)
                permute_kernel<T, 0><<<blocks_for(nnz), kBuildThreads, 0, stream>>>(vals_b, nnz, csr_edge, csr_w, row_of, csr_eid, edge, w, eid, seg_of);
            else
                permute_kernel<T, 1><<<blocks_for(nnz), kBuildThreads, 0, stream>>>(vals_b, nnz, csr_edge, csr_w, row_of, csr_eid, edge, w, eid, seg_of);
            note_launch();
            segments = seg_of;
        }
        segment_ptr_kernel<<<blocks_for((int64_t)n_seg[o] + 1), kBuildThreads, 0, stream>>>(segments, nnz, n_seg[o], ptr);
        note_launch();
        {   // packed edge ids and the running count of non-unit weights (for the per-task flag)
            const int32_t x_range = o == 0 ? n_in : n_out, y_range = o == 2 ? n_in : n_rel;
            const int x_bits = bit_length(x_range > 1 ? (unsigned long long)x_range - 1 : 1);
            const int y_bits = bit_length(y_range > 1 ? (unsigned long long)y_range - 1 : 1);
            pack_shift[o] = x_bits + y_bits <= 32 ? x_bits : 0;
            int32_t *flags = vals_a;
            int32_t *before = (int32_t *)(sbuf + S.nonunit) + (size_t)o * ((size_t)(nnz_raw > 0 ? nnz_raw : 1) + 1);
            pack_kernel<T><<<blocks_for((int64_t)nnz + 1), kBuildThreads, 0, stream>>>(
                (const int2 *)(ibuf + L.order[o].edge), (const T *)(ibuf + L.order[o].w), nnz, pack_shift[o],
                (uint32_t *)(ibuf + L.order[o].packed), flags);
            note_launch();
            if (host_counters[CNT_NONUNIT]) {
                need = S.cub_bytes;
                ULTRA_CUDA_OK(cub::DeviceScan::ExclusiveSum(cub_tmp, need, flags, before, nnz + 1, stream));
                note_launch();
            }
        }
        int32_t *cnt = (int32_t *)(sbuf + S.cnt);
        int32_t *off = (int32_t *)(sbuf + S.off) + (size_t)o * 4 * ((size_t)n_seg_max + 1);
        const size_t span = (size_t)n_seg_max + 1;
        task_count_kernel<<<blocks_for((int64_t)n_seg[o] + 1), kBuildThreads, 0, stream>>>(
            ptr, n_seg[o], chunk, cnt, cnt + span, cnt + 2 * span, counters + CNT_MAXSEG + o);
        note_launch();
        for (int q = 0; q < 3; ++q) {
            need = S.cub_bytes;
            ULTRA_CUDA_OK(cub::DeviceScan::ExclusiveSum(cub_tmp, need, cnt + q * span, off + q * span, n_seg[o] + 1, stream));
            note_launch();
        }
        gather_totals_kernel<<<1, 1, 0, stream>>>(off, off + span, off + 2 * span, n_seg[o], counters + CNT_TOTALS + 3 * o);
        note_launch();
        if (o < 2 && group > 0 && (long long)nnz < 30ll * n_seg[o]) {   // grouped task list: worth it for short segments only
            group_count_kernel<<<blocks_for((int64_t)n_seg[o] + 1), kBuildThreads, 0, stream>>>(ptr, n_seg[o], chunk, group,
                                                                                              cnt + 3 * span);
            note_launch();
            need = S.cub_bytes;
            ULTRA_CUDA_OK(cub::DeviceScan::ExclusiveSum(cub_tmp, need, cnt + 3 * span, off + 3 * span, n_seg[o] + 1, stream));
            note_launch();
            ULTRA_CUDA_OK(cudaMemcpyAsync(counters + CNT_GTASKS + o, off + 3 * span + n_seg[o], sizeof(int32_t),
                                          cudaMemcpyDeviceToDevice, stream));
        }
    }
    ULTRA_CUDA_OK(cudaMemcpyAsync(host_counters, counters, 4 * CNT_SIZE, cudaMemcpyDeviceToHost, stream));
    ULTRA_CUDA_OK(cudaStreamSynchronize(stream));

    // ---- tasks ------------------------------------------------------------------------------------
    ultra_rspmm_order_t *orders[3] = {&index->csr, &index->csc, &index->rel};
    for (int o = 0; o < 3; ++o) {
        ultra_rspmm_order_t &out = *orders[o];
        const size_t span = (size_t)n_seg_max + 1;
        const int32_t *off = (int32_t *)(sbuf + S.off) + (size_t)o * 4 * span;
        out.n_seg = n_seg[o];
        out.n_task = host_counters[CNT_TOTALS + 3 * o];
        out.n_slot = host_counters[CNT_TOTALS + 3 * o + 1];
        out.n_split = host_counters[CNT_TOTALS + 3 * o + 2];
        out.max_seg_nnz = host_counters[CNT_MAXSEG + o];
        out.pack_shift = pack_shift[o];
        out.packed = (const uint32_t *)(ibuf + L.order[o].packed);
        out.ptr = (const int32_t *)(ibuf + L.order[o].ptr);
        out.edge = (const int32_t *)(ibuf + L.order[o].edge);
        out.w = ibuf + L.order[o].w;
        out.eid = (const int32_t *)(ibuf + L.order[o].eid);
        out.task = (const int32_t *)(ibuf + L.order[o].task);
        out.split = (const int32_t *)(ibuf + L.order[o].split);
        if (out.n_task > task_upper(nnz_raw, n_seg[o], chunk) || out.n_split > split_upper(nnz_raw, chunk))
            return ULTRA_RSPMM_ERR_WORKSPACE;
        if (out.n_task == 0) continue;
        int4 *task_tmp = (int4 *)(sbuf + S.task_tmp);
        uint32_t *tkey_a = (uint32_t *)(sbuf + S.tkey_a), *tkey_b = (uint32_t *)(sbuf + S.tkey_b);
        int32_t *tval_a = (int32_t *)(sbuf + S.tval_a), *tval_b = (int32_t *)(sbuf + S.tval_b);
        task_emit_kernel<<<blocks_for(n_seg[o]), kBuildThreads, 0, stream>>>(
            out.ptr, n_seg[o], chunk, off, off + span, off + 2 * span,
            any_nonunit ? (const int32_t *)(sbuf + S.nonunit) + (size_t)o * ((size_t)(nnz_raw > 0 ? nnz_raw : 1) + 1) : nullptr,
            task_tmp, tkey_a, tval_a, (int4 *)(ibuf + L.order[o].split));
        note_launch();
        if (int status = sort_pairs(cub_tmp, S.cub_bytes, tkey_a, tkey_b, tval_a, tval_b, out.n_task,
                                    bit_length((unsigned long long)chunk), stream))
            return status;
        task_gather_kernel<<<blocks_for(out.n_task), kBuildThreads, 0, stream>>>(task_tmp, tval_b, out.n_task,
                                                                               (int4 *)(ibuf + L.order[o].task));
        note_launch();
        out.n_gtask = o < 2 && group > 0 && (long long)nnz < 30ll * n_seg[o] ? host_counters[CNT_GTASKS + o] : 0;
        out.group_edges = out.n_gtask > 0 ? group : 0;
        out.gtask = out.n_gtask > 0 ? (const int32_t *)(ibuf + L.order[o].gtask) : nullptr;
        if (out.n_gtask > 0) {
            if (out.n_gtask > task_upper(nnz_raw, n_seg[o], chunk)) return ULTRA_RSPMM_ERR_WORKSPACE;
            group_emit_kernel<<<blocks_for(n_seg[o]), kBuildThreads, 0, stream>>>(
                out.ptr, n_seg[o], chunk, group, off + 3 * span, off + span,
                any_nonunit ? (const int32_t *)(sbuf + S.nonunit) + (size_t)o * ((size_t)(nnz_raw > 0 ? nnz_raw : 1) + 1) : nullptr,
                task_tmp, tkey_a, tval_a);
            note_launch();
            if (int status = sort_pairs(cub_tmp, S.cub_bytes, tkey_a, tkey_b, tval_a, tval_b, out.n_gtask,
                                        bit_length((unsigned long long)chunk), stream))
                return status;
            task_gather_kernel<<<blocks_for(out.n_gtask), kBuildThreads, 0, stream>>>(task_tmp, tval_b, out.n_gtask,
                                                                                    (int4 *)(ibuf + L.order[o].gtask));
            note_launch();
        }
    }
    ULTRA_CUDA_OK(cudaGetLastError());
    ULTRA_CUDA_OK(cudaStreamSynchronize(stream));  // scratch may be released by the caller on return

    index->nnz = nnz;
    index->nnz_raw = nnz_raw;
    index->n_out = n_out;
    index->n_in = n_in;
    index->n_rel = n_rel;
    index->dtype = sizeof(T) == 4 ? ULTRA_RSPMM_F32 : ULTRA_RSPMM_F64;
    index->unit_weight = host_counters[CNT_NONUNIT] ? 0 : 1;
    index->chunk = chunk;
    index->merge_perm = (const int32_t *)(ibuf + L.merge_perm);
    index->merge_start = (const int32_t *)(ibuf + L.merge_start);
    return ULTRA_RSPMM_OK;
}

// ---- derived index: the same edge structure with other values (ultra_rspmm_index_derive) -----------------------------
// csr edge m <- sum of its raw edges in the order the build summed them (sequential, fp64, rounded once); the value is
// also scattered to the edge's canonical rank, from where the two other orders pick it up through their eid arrays.
template <typename T>
__global__ void derive_merge_kernel(const int32_t *__restrict__ merge_perm, const int32_t *__restrict__ merge_start,
                                    const T *__restrict__ values, int32_t nnz, const int32_t *__restrict__ csr_eid,
                                    T *__restrict__ csr_w, T *__restrict__ by_rank) {
    const int32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nnz) return;
    double total = 0.0;
    for (int32_t q = merge_start[m]; q < merge_start[m + 1]; ++q) total += (double)values[merge_perm[q]];
    const T w = (T)total;
    csr_w[m] = w;
    by_rank[csr_eid[m]] = w;
}

template <typename T>
__global__ void derive_gather_kernel(const int32_t *__restrict__ eid, const T *__restrict__ by_rank, int32_t nnz,
                                     T *__restrict__ w) {
    const int32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < nnz) w[m] = by_rank[eid[m]];
}

// one warp per task: copy the task with its non-unit flag recomputed from the new values
template <typename T>
__global__ void derive_task_flags_kernel(const int4 *__restrict__ task_in, int32_t n_task, const T *__restrict__ w,
                                         int4 *__restrict__ task_out) {
    const int32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= n_task) return;
    int4 task = task_in[t];
    bool nonunit = false;
    for (int32_t e = task.y + lane; e < task.z; e += 32) nonunit = nonunit || w[e] != T(1);
    nonunit = __any_sync(0xffffffffu, nonunit);
    task.w = nonunit ? (task.w | kNonUnitTask) : (task.w & ~kNonUnitTask);
    if (lane == 0) task_out[t] = task;
}

struct DerivedLayout {
    size_t by_rank, w[3], task[3], gtask[3], total;
};

DerivedLayout derived_layout(const ultra_rspmm_index_t &base) {
    DerivedLayout D;
    const size_t elem = base.dtype == ULTRA_RSPMM_F32 ? 4 : 8;
    const size_t e = base.nnz > 0 ? (size_t)base.nnz : 1;
    const ultra_rspmm_order_t *orders[3] = {&base.csr, &base.csc, &base.rel};
    size_t at = 0;
    D.by_rank = at; at = align_up(at + elem * e);
    for (int o = 0; o < 3; ++o) {
        D.w[o] = at; at = align_up(at + elem * e);
        D.task[o] = at; at = align_up(at + sizeof(int4) * (size_t)(orders[o]->n_task > 0 ? orders[o]->n_task : 1));
        D.gtask[o] = at; at = align_up(at + sizeof(int4) * (size_t)(orders[o]->n_gtask > 0 ? orders[o]->n_gtask : 1));
    }
    D.total = at;
    return D;
}

template <typename T>
int derive_typed(const ultra_rspmm_index_t &base, const T *values, char *buffer, ultra_rspmm_index_t *derived,
                 cudaStream_t stream) {
    const DerivedLayout D = derived_layout(base);
    *derived = base;
    derived->unit_weight = 0;               // decided per task (flag in task.w), without a host round trip
    const int32_t nnz = (int32_t)base.nnz;
    ultra_rspmm_order_t *orders[3] = {&derived->csr, &derived->csc, &derived->rel};
    T *by_rank = (T *)(buffer + D.by_rank);
    if (nnz > 0) {
        derive_merge_kernel<T><<<blocks_for(nnz), kBuildThreads, 0, stream>>>(base.merge_perm, base.merge_start, values, nnz,
                                                                              base.csr.eid, (T *)(buffer + D.w[0]), by_rank);
        note_launch();
    }
    for (int o = 0; o < 3; ++o) {
        ultra_rspmm_order_t &order = *orders[o];
        T *w = (T *)(buffer + D.w[o]);
        if (o > 0 && nnz > 0) {
            derive_gather_kernel<T><<<blocks_for(nnz), kBuildThreads, 0, stream>>>(order.eid, by_rank, nnz, w);
            note_launch();
        }
        order.w = w;
        if (order.n_task > 0) {
            int4 *task = (int4 *)(buffer + D.task[o]);
            derive_task_flags_kernel<T><<<blocks_for((int64_t)order.n_task * 32), kBuildThreads, 0, stream>>>(
                (const int4 *)order.task, order.n_task, w, task);
            note_launch();
            order.task = (const int32_t *)task;
        }
        if (order.n_gtask > 0) {
            int4 *gtask = (int4 *)(buffer + D.gtask[o]);
            derive_task_flags_kernel<T><<<blocks_for((int64_t)order.n_gtask * 32), kBuildThreads, 0, stream>>>(
                (const int4 *)order.gtask, order.n_gtask, w, gtask);
            note_launch();
            order.gtask = (const int32_t *)gtask;
        }
    }
    ULTRA_CUDA_OK(cudaGetLastError());
    return ULTRA_RSPMM_OK;
}

int check_shape(int64_t nnz_raw, int32_t n_out, int32_t n_in, int32_t n_rel, int32_t dtype) {
    if (nnz_raw < 0 || n_out < 0 || n_in < 0 || n_rel < 0) return ULTRA_RSPMM_ERR_ARG;
    if (dtype != ULTRA_RSPMM_F32 && dtype != ULTRA_RSPMM_F64) return ULTRA_RSPMM_ERR_ARG;
    if (nnz_raw >= (int64_t)1 << 31) return ULTRA_RSPMM_ERR_RANGE;
    const unsigned __int128 span = (unsigned __int128)n_out * (unsigned __int128)n_in * (unsigned __int128)n_rel;
    if (span >> 63) return ULTRA_RSPMM_ERR_RANGE;
    return ULTRA_RSPMM_OK;
}

}  // namespace

}  // namespace ultra

using namespace ultra;

extern "C" int ultra_rspmm_index_bytes(int64_t nnz_raw, int32_t n_out, int32_t n_in, int32_t n_rel, int32_t dtype,
                                       size_t *index_bytes, size_t *scratch_bytes) {
    const int status = check_shape(nnz_raw, n_out, n_in, n_rel, dtype);
    if (status) return status;
    if (!index_bytes || !scratch_bytes) return ULTRA_RSPMM_ERR_ARG;
    const int32_t n_seg[3] = {n_out, n_in, n_rel};
    const int32_t n_seg_max = n_out > n_in ? (n_out > n_rel ? n_out : n_rel) : (n_in > n_rel ? n_in : n_rel);
    *index_bytes = index_layout(nnz_raw, n_seg, dtype == ULTRA_RSPMM_F32 ? 4 : 8, g_chunk).total;
    *scratch_bytes = scratch_layout(nnz_raw, n_seg_max, g_chunk).total;
    return ULTRA_RSPMM_OK;
}

extern "C" int ultra_rspmm_index_build(const int64_t *dev_indices, int64_t index_stride, const void *dev_values,
                                       int64_t nnz_raw, int32_t n_out, int32_t n_in, int32_t n_rel, int32_t dtype,
                                       void *index_buffer, size_t index_bytes, void *scratch, size_t scratch_bytes,
                                       ultra_rspmm_index_t *index, void *stream) {
    int status = check_shape(nnz_raw, n_out, n_in, n_rel, dtype);
    if (status) return status;
    if (!index || !index_buffer || !scratch) return ULTRA_RSPMM_ERR_ARG;
    if (nnz_raw > 0 && (!dev_indices || !dev_values || index_stride < nnz_raw)) return ULTRA_RSPMM_ERR_ARG;
    size_t need_index = 0, need_scratch = 0;
    status = ultra_rspmm_index_bytes(nnz_raw, n_out, n_in, n_rel, dtype, &need_index, &need_scratch);
    if (status) return status;
    if (index_bytes < need_index || scratch_bytes < need_scratch) return ULTRA_RSPMM_ERR_WORKSPACE;
    if (((uintptr_t)index_buffer | (uintptr_t)scratch) & 255) return ULTRA_RSPMM_ERR_ARG;
    memset(index, 0, sizeof(*index));
    if (dtype == ULTRA_RSPMM_F32)
        return build_typed<float>(dev_indices, index_stride, (const float *)dev_values, nnz_raw, n_out, n_in, n_rel,
                                  (char *)index_buffer, (char *)scratch, index, (cudaStream_t)stream);
    return build_typed<double>(dev_indices, index_stride, (const double *)dev_values, nnz_raw, n_out, n_in, n_rel,
                               (char *)index_buffer, (char *)scratch, index, (cudaStream_t)stream);
}

extern "C" int ultra_rspmm_fingerprint(const int64_t *dev_indices, int64_t index_stride, const void *dev_values,
                                       int64_t nnz_raw, int32_t dtype, uint64_t *dev_out, void *stream) {
    if (nnz_raw < 0 || !dev_out || (nnz_raw > 0 && (!dev_indices || !dev_values))) return ULTRA_RSPMM_ERR_ARG;
    if (dtype != ULTRA_RSPMM_F32 && dtype != ULTRA_RSPMM_F64) return ULTRA_RSPMM_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    ULTRA_CUDA_OK(cudaMemsetAsync(dev_out, 0, 16, s));
    if (nnz_raw == 0) return ULTRA_RSPMM_OK;
    int blocks = blocks_for(nnz_raw);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (dtype == ULTRA_RSPMM_F32)
        fingerprint_kernel<float><<<blocks, kBuildThreads, 0, s>>>(dev_indices, index_stride, (const float *)dev_values,
                                                                  nnz_raw, (unsigned long long *)dev_out);
    else
        fingerprint_kernel<double><<<blocks, kBuildThreads, 0, s>>>(dev_indices, index_stride, (const double *)dev_values,
                                                                   nnz_raw, (unsigned long long *)dev_out);
    note_launch();
    ULTRA_CUDA_OK(cudaGetLastError());
    return ULTRA_RSPMM_OK;
}

extern "C" int ultra_rspmm_index_derive_bytes(const ultra_rspmm_index_t *base, size_t *bytes) {
    if (!base || !bytes) return ULTRA_RSPMM_ERR_ARG;
    *bytes = derived_layout(*base).total;
    return ULTRA_RSPMM_OK;
}

extern "C" int ultra_rspmm_index_derive(const ultra_rspmm_index_t *base, const void *dev_values, void *buffer, size_t bytes,
                                        ultra_rspmm_index_t *derived, void *stream) {
    if (!base || !derived || !buffer || (base->nnz_raw > 0 && !dev_values)) return ULTRA_RSPMM_ERR_ARG;
    if (base->nnz > 0 && (!base->merge_perm || !base->merge_start)) return ULTRA_RSPMM_ERR_ARG;
    if (bytes < derived_layout(*base).total) return ULTRA_RSPMM_ERR_WORKSPACE;
    if ((uintptr_t)buffer & 255) return ULTRA_RSPMM_ERR_ARG;
    if (base->dtype == ULTRA_RSPMM_F32)
        return derive_typed<float>(*base, (const float *)dev_values, (char *)buffer, derived, (cudaStream_t)stream);
    return derive_typed<double>(*base, (const double *)dev_values, (char *)buffer, derived, (cudaStream_t)stream);
}
