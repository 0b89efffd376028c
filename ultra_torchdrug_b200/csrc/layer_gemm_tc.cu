// The Linear of `combine` under autograd (reference ultra/layer.py:386-392 in the fine-tuning step), fp32 accuracy on the
// tensor cores (3xTF32 split, see layer_linear.cu) instead of cuBLAS SIMT SGEMMs and without `cat([input, update])`:
//
//   ultra_layer_rows_gemm          out[r, 0:N] = [A0[r, :] | A1[r, :]] @ W^T        rows x K  ->  rows x N      (tcgen05 + TMA)
//        forward   N = 64,  K = 128:  x = [input | update] @ W^T   (A0 = input, A1 = update: the cat is never materialised)
//        grad rows N = 128, K = 64:   [d input | d update] = dx @ W  (A0 = dx, W passed transposed; the two 64-column
//                                     halves go to two tensors, the first optionally + addend: the short-cut's gradient)
//   ultra_layer_rows_gemm_weight   dW[n, k] = sum_r dx[r, n] * [A0[r, :] | A1[r, :]][k]   (tcgen05 with MN-major operands, fixed-order fold)
//
// rows_gemm_tc_kernel is the fused-Linear kernel of layer_linear_tc.cu with a plain epilogue: persistent CTAs, one per SM;
// warp 0 = TMA producer (16 KB SWIZZLE_128B boxes of 32 columns x 128 rows into a ring of 3 slots, from one or two tensor
// maps), warps 2-5 = hi / lo split into A slots in tensor memory (tcgen05.st), warp 1 = MMA issuer (tcgen05.mma.kind::tf32,
// A from tensor memory, M 128, N 64 or 128, two accumulator stages in TMEM), warps 6-9 = epilogue (tcgen05.ld 64 columns at a time, staged through warp-private shared
// memory, coalesced 16-byte stores).
//
// weight_grad_tc_kernel: see its own header below.  weight_grad_kernel (the first version, ULTRA_WEIGHT_GRAD=mma) reduces over
// millions of rows into a 64 x 128 result: persistent CTAs walk 64-row tiles (cp.async
// double buffer), every warp owns a 16 x 64 block of the result in registers (mma.sync.m16n8k8 tf32, fragments gathered
// from shared memory - the transposition dx^T costs nothing there), 3 MMAs per product; each CTA writes its partial
// result, a second kernel folds the partials in CTA order: deterministic, no atomics.
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace ultra {

namespace {

using namespace tcx;

namespace gm {
constexpr int kRows = 128;
constexpr int kSlotK = 32;
constexpr int kSlots = 3;                               // landing slots of the TMA ring (16 KB each)
constexpr int kASlots = 2;                              // A slots in tensor memory (64 columns each: hi | lo), see layer_linear_tc.cu
constexpr int kTmaWarp = 0, kMmaWarp = 1, kSplitWarp0 = 2, kSplitWarps = 4, kEpilogueWarp0 = 6, kEpilogueWarps = 4;
constexpr int kThreads = 32 * (kEpilogueWarp0 + kEpilogueWarps);
constexpr int kSlotHalfBytes = kRows * kSlotK * 4;
constexpr int kBarriers = 2 * kSlots + 2 * kASlots + 4;
constexpr int kHalf = 64;                               // epilogue works on 64 output columns at a time
constexpr int kStageStride = kHalf + 4;

template <int N, int K> struct Shape {
    static constexpr int kSlotsPerTile = K / kSlotK;
    static constexpr int kWeightHalfBytes = N * K * 4;
    static constexpr int kCoreBytesW = N * 16;
    static constexpr int kRingOffset = (2 * kWeightHalfBytes + 1023) / 1024 * 1024;
    static constexpr int kStagingOffset = kRingOffset + kSlots * kSlotHalfBytes;
    static constexpr int kBarrierOffset = kStagingOffset + kEpilogueWarps * 32 * kStageStride * 4;
    static constexpr int kSmemBytes = kBarrierOffset + kBarriers * 8 + 16;
    static constexpr int kOperandColumn0 = 2 * N;       // two accumulator stages (128 or 256 columns), then the A slots
    static constexpr int kTmemColumns = 512;
    static_assert(2 * N + 64 * kASlots <= 512, "accumulators and A slots exceed tensor memory");
    static constexpr unsigned kInstr = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(kRows >> 4) << 24);
    static_assert(kSmemBytes <= 227 * 1024, "does not fit shared memory");
};
}  // namespace gm

struct RowsGemmArgs {
    const float *W;          // (N, K) row-major
    float *out0, *out1;      // columns [0, 64) -> out0; [64, 128) -> out1 (or out0 + 64 when out1 is null)
    const float *addend0;    // optional, added to the first 64 columns (rows ld_addend apart)
    long long ld0, ld1, ld_addend, rows;
    int two_sources;         // K slots [0, S/2) from map 0, [S/2, S) from map 1
};

template <int N, int K>
__global__ void __launch_bounds__(gm::kThreads, 1)
rows_gemm_tc_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1, const RowsGemmArgs a) {
    using S = gm::Shape<N, K>;
    constexpr int kSlotsPerTile = S::kSlotsPerTile;
    extern __shared__ __align__(1024) unsigned char smem[];
    const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem);
    const unsigned bar_base = smem_base + S::kBarrierOffset;
    // landed / empty per landing slot (empty = the split warps have read it); full / a_free per tensor-memory A slot
    auto landed_bar = [&](int slot) { return bar_base + 8u * slot; };
    auto empty_bar = [&](int slot) { return bar_base + 8u * (gm::kSlots + slot); };
    auto full_bar = [&](int slot) { return bar_base + 8u * (2 * gm::kSlots + slot); };
    auto a_free_bar = [&](int slot) { return bar_base + 8u * (2 * gm::kSlots + gm::kASlots + slot); };
    auto tmem_full_bar = [&](int stage) { return bar_base + 8u * (2 * gm::kSlots + 2 * gm::kASlots + stage); };
    auto tmem_empty_bar = [&](int stage) { return bar_base + 8u * (2 * gm::kSlots + 2 * gm::kASlots + 2 + stage); };
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(smem + S::kBarrierOffset + gm::kBarriers * 8);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < gm::kSlots; ++s) {
            mbar_init(empty_bar(s), gm::kSplitWarps);
            mbar_init(landed_bar(s), 1);
        }
        for (int s = 0; s < gm::kASlots; ++s) {
            mbar_init(full_bar(s), gm::kSplitWarps);
            mbar_init(a_free_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tmem_full_bar(s), 1);
            mbar_init(tmem_empty_bar(s), gm::kEpilogueWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == gm::kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (unsigned)__cvta_generic_to_shared(tmem_slot)), "r"((unsigned)S::kTmemColumns) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // W (N, K) split into hi / lo, canonical no-swizzle K-major UMMA layout (8-row x 16-byte core matrices)
    for (int idx = tid; idx < N * (K / 4); idx += gm::kThreads) {
        const int n = idx / (K / 4), kc = idx % (K / 4);
        const float4 w = __ldg(reinterpret_cast<const float4 *>(a.W + n * K + 4 * kc));
        const float4 hi = make_float4(tc_tf32(w.x), tc_tf32(w.y), tc_tf32(w.z), tc_tf32(w.w));
        const float4 lo = make_float4(tc_tf32(w.x - hi.x), tc_tf32(w.y - hi.y), tc_tf32(w.z - hi.z), tc_tf32(w.w - hi.w));
        unsigned char *at = smem + kc * S::kCoreBytesW + n * 16;
        *reinterpret_cast<float4 *>(at) = hi;
        *reinterpret_cast<float4 *>(at + S::kWeightHalfBytes) = lo;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = *tmem_slot;

    const long long n_tiles = (a.rows + gm::kRows - 1) / gm::kRows;
    const long long first = blockIdx.x;
    const long long my_tiles = first < n_tiles ? (n_tiles - first + gridDim.x - 1) / gridDim.x : 0;
    const long long total = my_tiles * kSlotsPerTile;

    if (warp == gm::kTmaWarp) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map0) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map1) : "memory");
            for (long long it = 0; it < total; ++it) {
                const int slot = (int)(it % gm::kSlots);
                const unsigned phase = (unsigned)((it / gm::kSlots) & 1);
                const long long tile = first + (it / kSlotsPerTile) * gridDim.x;
                const int q = (int)(it % kSlotsPerTile);
                const bool second = a.two_sources && q >= kSlotsPerTile / 2;
                const int k0 = (second ? q - kSlotsPerTile / 2 : q) * gm::kSlotK;
                mbar_wait(empty_bar(slot), phase ^ 1u);
                mbar_expect_tx(landed_bar(slot), gm::kSlotHalfBytes);
                tma_load_2d(smem_base + S::kRingOffset + slot * gm::kSlotHalfBytes, second ? &map1 : &map0, k0,
                            (int)(tile * gm::kRows), landed_bar(slot));
            }
        }
    } else if (warp >= gm::kSplitWarp0 && warp < gm::kEpilogueWarp0) {
        // one thread = one row of a landed slot (the tensor-memory lane its warp may write): 8 conflict-free 16-byte reads
        // of the swizzled row, hi = tf32(x) and lo = tf32(x - hi) to the A slot's columns with tcgen05.st
        const int quadrant = warp & 3, r = 32 * quadrant + lane;
        for (long long it = 0; it < total; ++it) {
            const int slot = (int)(it % gm::kSlots), a_slot = (int)(it % gm::kASlots);
            mbar_wait(landed_bar(slot), (unsigned)((it / gm::kSlots) & 1));
            mbar_wait(a_free_bar(a_slot), (unsigned)((it / gm::kASlots) & 1) ^ 1u);   // the MMAs that read this A slot have completed
            const unsigned char *at = smem + S::kRingOffset + slot * gm::kSlotHalfBytes + r * 128;
            float4 x[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) x[c] = *reinterpret_cast<const float4 *>(at + ((c ^ (r & 7)) << 4));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned taddr = tmem_base + ((unsigned)(32 * quadrant) << 16) + (unsigned)(S::kOperandColumn0 + 64 * a_slot);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float hi[16], lo[16];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float4 v = x[4 * h + c];
                    hi[4 * c] = tc_tf32(v.x); hi[4 * c + 1] = tc_tf32(v.y); hi[4 * c + 2] = tc_tf32(v.z); hi[4 * c + 3] = tc_tf32(v.w);
                    lo[4 * c] = tc_tf32(v.x - hi[4 * c]); lo[4 * c + 1] = tc_tf32(v.y - hi[4 * c + 1]);
                    lo[4 * c + 2] = tc_tf32(v.z - hi[4 * c + 2]); lo[4 * c + 3] = tc_tf32(v.w - hi[4 * c + 3]);
                }
                tmem_store16(taddr + 16 * h, hi);
                tmem_store16(taddr + 32 + 16 * h, lo);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_bar(slot));                 // the slot's bytes are in registers: TMA may refill it
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar(a_slot));
        }
    } else if (warp == gm::kMmaWarp) {
        if (lane == 0) {
            const unsigned w_hi = smem_base, w_lo = smem_base + S::kWeightHalfBytes;
            long long it = 0;
            for (long long t = 0; t < my_tiles; ++t) {
                const int stage = (int)(t & 1);
                const unsigned accum_phase = (unsigned)((t >> 1) & 1);
                mbar_wait(tmem_empty_bar(stage), accum_phase ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned tmem_d = tmem_base + (unsigned)(stage * N);
                for (int q = 0; q < kSlotsPerTile; ++q, ++it) {
                    const int a_slot = (int)(it % gm::kASlots);
                    mbar_wait(full_bar(a_slot), (unsigned)((it / gm::kASlots) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const unsigned a_hi = tmem_base + (unsigned)(S::kOperandColumn0 + 64 * a_slot), a_lo = a_hi + 32u;
#pragma unroll
                    for (int ks = 0; ks < gm::kSlotK / 8; ++ks) {
                        const int kg = q * (gm::kSlotK / 8) + ks;
                        const unsigned long long db_hi = umma_desc(w_hi + kg * 2 * S::kCoreBytesW, S::kCoreBytesW, 128);
                        const unsigned long long db_lo = umma_desc(w_lo + kg * 2 * S::kCoreBytesW, S::kCoreBytesW, 128);
                        umma_tf32_ts(tmem_d, a_lo + 8u * ks, db_hi, S::kInstr, kg > 0 ? 1u : 0u);
                        umma_tf32_ts(tmem_d, a_hi + 8u * ks, db_lo, S::kInstr, 1u);
                        umma_tf32_ts(tmem_d, a_hi + 8u * ks, db_hi, S::kInstr, 1u);
                    }
                    umma_commit(a_free_bar(a_slot));                     // the A slot may be overwritten once these MMAs have read it
                }
                umma_commit(tmem_full_bar(stage));
            }
        }
    } else if (warp >= gm::kEpilogueWarp0) {
        // one thread = one row (tcgen05.ld.32x32b); 64 columns at a time through warp-private staging rows, then the warp
        // writes its 32 rows two at a time with lane = (row, 16-byte chunk): coalesced stores (and addend loads)
        constexpr int kChunks = gm::kHalf / 4, kRowsPerPass = 32 / kChunks, kStride = gm::kStageStride, kPasses = 32 / kRowsPerPass;
        const int quadrant = warp & 3;
        float *staged = reinterpret_cast<float *>(smem + S::kStagingOffset) + quadrant * 32 * kStride;
        const int my_chunk = lane % kChunks, sub_row = lane / kChunks;
        for (long long t = 0; t < my_tiles; ++t) {
            const int stage = (int)(t & 1);
            const unsigned accum_phase = (unsigned)((t >> 1) & 1);
            const long long row0 = (first + t * gridDim.x) * gm::kRows + 32 * quadrant;
            mbar_wait(tmem_full_bar(stage), accum_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int h = 0; h < N / gm::kHalf; ++h) {
                float v[gm::kHalf];
#pragma unroll
                for (int c = 0; c < gm::kHalf / 16; ++c) {
                    float part[16];
                    tmem_load16(tmem_base + ((unsigned)(32 * quadrant) << 16) + (unsigned)(stage * N + h * gm::kHalf + 16 * c), part);
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[16 * c + i] = part[i];
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (h == N / gm::kHalf - 1) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tmem_empty_bar(stage));
                }
#pragma unroll
                for (int c = 0; c < gm::kHalf; c += 4)
                    *reinterpret_cast<float4 *>(staged + lane * kStride + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
                __syncwarp();
                float *out = h == 0 ? a.out0 : (a.out1 ? a.out1 : a.out0 + gm::kHalf);
                const long long ld = h == 0 || !a.out1 ? a.ld0 : a.ld1;
                const float *addend = h == 0 ? a.addend0 : nullptr;
                constexpr int kBatch = 8;                                // addend loads in flight before their first use
#pragma unroll
                for (int first_pass = 0; first_pass < kPasses; first_pass += kBatch) {
                    float4 extra[kBatch];
#pragma unroll
                    for (int j = 0; j < kBatch; ++j) {
                        const long long row = row0 + (first_pass + j) * kRowsPerPass + sub_row;
                        extra[j] = addend && row < a.rows ? __ldg(reinterpret_cast<const float4 *>(addend + row * a.ld_addend + 4 * my_chunk))
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int j = 0; j < kBatch; ++j) {
                        const int r = (first_pass + j) * kRowsPerPass + sub_row;
                        const long long row = row0 + r;
                        float4 y = *reinterpret_cast<const float4 *>(staged + r * kStride + 4 * my_chunk);
                        y = make_float4(y.x + extra[j].x, y.y + extra[j].y, y.z + extra[j].z, y.w + extra[j].w);
                        if (row < a.rows) *reinterpret_cast<float4 *>(out + row * ld + 4 * my_chunk) = y;
                    }
                }
                __syncwarp();
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == gm::kMmaWarp) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((unsigned)S::kTmemColumns) : "memory");
    }
}

template <int N, int K>
int launch_rows_gemm(const float *A0, long long lda0, const float *A1, long long lda1, const RowsGemmArgs &args, cudaStream_t stream) {
    using S = gm::Shape<N, K>;
    auto kernel = rows_gemm_tc_kernel<N, K>;
    int device = 0, sm_count = 0;
    ULTRA_CUDA_OK(cudaGetDevice(&device));
    ULTRA_CUDA_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
    ULTRA_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kSmemBytes));
    CUtensorMap map0, map1;
    const long long cols = args.two_sources ? K / 2 : K;
    if (int status = encode_rows_map(&map0, A0, args.rows, cols, lda0)) return status;
    if (int status = encode_rows_map(&map1, args.two_sources ? A1 : A0, args.rows, cols, args.two_sources ? lda1 : lda0)) return status;
    const long long n_tiles = (args.rows + gm::kRows - 1) / gm::kRows;
    const unsigned grid = (unsigned)(n_tiles < sm_count ? n_tiles : sm_count);
    kernel<<<grid, gm::kThreads, S::kSmemBytes, stream>>>(map0, map1, args);
    note_launch();
    ULTRA_CUDA_OK(cudaGetLastError());
    return ULTRA_RSPMM_OK;
}

// ---- weight gradient --------------------------------------------------------------------------------------------------
namespace wg {
constexpr int kTileRows = 64;
constexpr int kWarps = 8;                  // warp w: result rows [16 (w % 4), +16) x columns [64 (w / 4), +64)
constexpr int kThreads = kWarps * 32;
constexpr int kPitchD = 64 + 8;            // dx tile row pitch (floats): conflict-free transposed fragment reads
constexpr int kPitchJ = 128 + 8;
constexpr int kStageFloats = kTileRows * (kPitchD + kPitchJ);
constexpr int kSmemBytes = 2 * kStageFloats * 4;
}  // namespace wg

__device__ __forceinline__ void cp_async16(void *dst, const void *src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int bytes = valid ? 16 : 0;      // zero-fill past the last row
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}

__device__ __forceinline__ void split_tf32(float x, unsigned &hi, unsigned &lo) {
    const float h = tc_tf32(x);
    hi = __float_as_uint(h);
    lo = __float_as_uint(tc_tf32(x - h));
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// partial[cta][n][k] = sum over the CTA's rows of dx[r][n] * J[r][k],  J = [A0 | A1] (64 columns each)
__global__ void __launch_bounds__(wg::kThreads, 1)
weight_grad_kernel(const float *__restrict__ dx, long long ld_dx, const float *__restrict__ A0, long long lda0,
                   const float *__restrict__ A1, long long lda1, long long rows, float *__restrict__ partial) {
    extern __shared__ __align__(16) float wsm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int m0 = 16 * (warp & 3), n0 = 64 * (warp >> 2);
    const long long n_tiles = (rows + wg::kTileRows - 1) / wg::kTileRows;
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[j][q] = 0.f;

    auto issue = [&](long long tile, int stage) {
        float *sd = wsm + stage * wg::kStageFloats, *sj = sd + wg::kTileRows * wg::kPitchD;
        const long long row0 = tile * wg::kTileRows;
        // dx: 64 rows x 16 chunks; J: 64 rows x 32 chunks (16 from A0, 16 from A1)
        for (int i = tid; i < wg::kTileRows * 16; i += wg::kThreads) {
            const int r = i >> 4, c = i & 15;
            const bool valid = row0 + r < rows;
            cp_async16(sd + r * wg::kPitchD + 4 * c, dx + (valid ? row0 + r : 0) * ld_dx + 4 * c, valid);
        }
        for (int i = tid; i < wg::kTileRows * 32; i += wg::kThreads) {
            const int r = i >> 5, c = i & 31;
            const bool valid = row0 + r < rows;
            const float *src = c < 16 ? A0 + (valid ? row0 + r : 0) * lda0 + 4 * c : A1 + (valid ? row0 + r : 0) * lda1 + 4 * (c - 16);
            cp_async16(sj + r * wg::kPitchJ + 4 * c, src, valid);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    long long tile = blockIdx.x;
    int stage = 0;
    if (tile < n_tiles) issue(tile, 0);
    for (; tile < n_tiles; tile += gridDim.x, stage ^= 1) {
        const long long next = tile + gridDim.x;
        if (next < n_tiles) {
            issue(next, stage ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const float *sd = wsm + stage * wg::kStageFloats, *sj = sd + wg::kTileRows * wg::kPitchD;
#pragma unroll 2
        for (int ks = 0; ks < wg::kTileRows / 8; ++ks) {
            // A fragment (16 result rows x 8 tile rows): A[m][k] = dx[k][m] - read transposed from the row-major tile
            unsigned a_hi[4], a_lo[4];
            split_tf32(sd[(8 * ks + t) * wg::kPitchD + m0 + g], a_hi[0], a_lo[0]);
            split_tf32(sd[(8 * ks + t) * wg::kPitchD + m0 + g + 8], a_hi[1], a_lo[1]);
            split_tf32(sd[(8 * ks + t + 4) * wg::kPitchD + m0 + g], a_hi[2], a_lo[2]);
            split_tf32(sd[(8 * ks + t + 4) * wg::kPitchD + m0 + g + 8], a_hi[3], a_lo[3]);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                unsigned b_hi[2], b_lo[2];
                split_tf32(sj[(8 * ks + t) * wg::kPitchJ + n0 + 8 * j + g], b_hi[0], b_lo[0]);
                split_tf32(sj[(8 * ks + t + 4) * wg::kPitchJ + n0 + 8 * j + g], b_hi[1], b_lo[1]);
                mma_tf32(acc[j], a_lo, b_hi[0], b_hi[1]);
                mma_tf32(acc[j], a_hi, b_lo[0], b_lo[1]);
                mma_tf32(acc[j], a_hi, b_hi[0], b_hi[1]);
            }
        }
        __syncthreads();          // the stage is refilled two iterations from now
    }
    // accumulator layout of m16n8: d0:(g, 2t) d1:(g, 2t+1) d2:(g+8, 2t) d3:(g+8, 2t+1)
    float *mine = partial + (long long)blockIdx.x * 64 * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int col = n0 + 8 * j + 2 * t;
        *reinterpret_cast<float2 *>(mine + (m0 + g) * 128 + col) = make_float2(acc[j][0], acc[j][1]);
        *reinterpret_cast<float2 *>(mine + (m0 + g + 8) * 128 + col) = make_float2(acc[j][2], acc[j][3]);
    }
}

// ---- the weight gradient on tcgen05 ----------------------------------------------------------------------------------------
// dW[n, k] = sum_r dx[r, n] J[r, k] is a GEMM whose reduction runs over the ROWS: both operands are MN-major (the row index is
// the K index of the MMA), which tcgen05 takes directly from what TMA writes (32-column boxes in the 128B_ATOM_32B swizzle
// mode = UMMA layout SWIZZLE_128B_BASE32B, the only one for MN-major tf32; a_major = b_major = 1).  Per 32-row tile a persistent CTA (two per SM) loads 6 boxes (dx: 2, J = [A0 | A1]: 4), the split warps rewrite them in
// place as hi = tf32(x) and put lo = tf32(x - hi) behind them, and ONE thread issues per 8 rows two MMAs of M 128, N 128:
//     [dx_hi ; dx_lo]^T (stacked along M: 4 atoms)  x  J_hi     and     [dx_hi ; dx_lo]^T  x  J_lo
// - all four terms of (hi + lo)(hi + lo) with two full-size instructions.  The accumulator (128 lanes x 128 columns of
// tensor memory) lives for the CTA's whole row range; at the end rows n and n + 64 (the hi and lo halves) are added and the
// CTA writes its partial result; weight_grad_fold_kernel adds the partials in CTA order (deterministic, no atomics).
// tile rows x stages x CTAs per SM (measured per call in the C3 kernel table: 64 x 2 x 1 0.509 ms, 48 x 3 x 1 0.503, 32 x 4 x 1
// 0.544, 32 x 2 x 2 0.381: two small CTAs per SM overlap each other's TMA wait / split / MMA phases, a deeper ring does not)
#ifndef WT_ROWS
#define WT_ROWS 32
#define WT_STAGES 2
#define WT_CTAS 2
#endif
namespace wt {
constexpr int kTileRows = WT_ROWS;
constexpr int kStages = WT_STAGES;
constexpr int kAtomBytes = kTileRows * 128;                       // 32 columns x 64 rows
constexpr int kAHi = 0, kALo = 2 * kAtomBytes, kBHi = 4 * kAtomBytes, kBLo = 8 * kAtomBytes;
constexpr int kStageBytes = 12 * kAtomBytes;                      // 96 KB
constexpr int kRawBytes = 6 * kAtomBytes;                         // what TMA delivers per tile
constexpr int kBarrierOffset = kStages * kStageBytes;
constexpr int kSmemBytes = kBarrierOffset + (3 * kStages + 1) * 8 + 16;
constexpr int kThreads = 6 * 32;                                  // warp 0 TMA, warp 1 MMA, warps 2-5 split + epilogue
constexpr int kScratchPitch = 132;
// D = F32, A = B = TF32, both MN-major (bits 15, 16), N = 128 (>> 3 at bit 17), M = 128 (>> 4 at bit 24)
constexpr unsigned kInstr = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
static_assert(kSmemBytes <= 227 * 1024, "does not fit shared memory");
static_assert(64 * kScratchPitch * 4 <= kStageBytes, "epilogue scratch lives in stage 0");
}  // namespace wt

__global__ void __launch_bounds__(wt::kThreads, WT_CTAS)
weight_grad_tc_kernel(const __grid_constant__ CUtensorMap map_dx, const __grid_constant__ CUtensorMap map_a0,
                      const __grid_constant__ CUtensorMap map_a1, long long rows, float *__restrict__ partial) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem);
    const unsigned bar_base = smem_base + wt::kBarrierOffset;
    auto landed_bar = [&](int s) { return bar_base + 8u * s; };
    auto full_bar = [&](int s) { return bar_base + 8u * (wt::kStages + s); };
    auto empty_bar = [&](int s) { return bar_base + 8u * (2 * wt::kStages + s); };
    const unsigned done_bar = bar_base + 8u * (3 * wt::kStages);
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(smem + wt::kBarrierOffset + (3 * wt::kStages + 1) * 8);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < wt::kStages; ++s) {
            mbar_init(landed_bar(s), 1);
            mbar_init(full_bar(s), 4);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (unsigned)__cvta_generic_to_shared(tmem_slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = *tmem_slot;
    const long long n_tiles = (rows + wt::kTileRows - 1) / wt::kTileRows;
    const long long first = blockIdx.x;
    const long long my_tiles = first < n_tiles ? (n_tiles - first + gridDim.x - 1) / gridDim.x : 0;

    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dx) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a0) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a1) : "memory");
            for (long long t = 0; t < my_tiles; ++t) {
                const int s = (int)(t % wt::kStages);
                const int row0 = (int)((first + t * gridDim.x) * wt::kTileRows);
                mbar_wait(empty_bar(s), (unsigned)((t / wt::kStages) & 1) ^ 1u);
                mbar_expect_tx(landed_bar(s), wt::kRawBytes);
                const unsigned at = smem_base + s * wt::kStageBytes;
                tma_load_2d(at + wt::kAHi, &map_dx, 0, row0, landed_bar(s));          // rows past the end: zeros
                tma_load_2d(at + wt::kAHi + wt::kAtomBytes, &map_dx, 32, row0, landed_bar(s));
                tma_load_2d(at + wt::kBHi, &map_a0, 0, row0, landed_bar(s));
                tma_load_2d(at + wt::kBHi + wt::kAtomBytes, &map_a0, 32, row0, landed_bar(s));
                tma_load_2d(at + wt::kBHi + 2 * wt::kAtomBytes, &map_a1, 0, row0, landed_bar(s));
                tma_load_2d(at + wt::kBHi + 3 * wt::kAtomBytes, &map_a1, 32, row0, landed_bar(s));
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (long long t = 0; t < my_tiles; ++t) {
                const int s = (int)(t % wt::kStages);
                mbar_wait(full_bar(s), (unsigned)((t / wt::kStages) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned at = smem_base + s * wt::kStageBytes;
#pragma unroll
                for (int ks = 0; ks < wt::kTileRows / 8; ++ks) {
                    const unsigned long long a = umma_desc_mn_sw128_32b(at + wt::kAHi + ks * 1024, wt::kAtomBytes);   // hi, hi, lo, lo atoms
                    const unsigned long long b_hi = umma_desc_mn_sw128_32b(at + wt::kBHi + ks * 1024, wt::kAtomBytes);
                    const unsigned long long b_lo = umma_desc_mn_sw128_32b(at + wt::kBLo + ks * 1024, wt::kAtomBytes);
                    umma_tf32(tmem_base, a, b_lo, wt::kInstr, (t > 0 || ks > 0) ? 1u : 0u);
                    umma_tf32(tmem_base, a, b_hi, wt::kInstr, 1u);
                }
                umma_commit(empty_bar(s));
            }
            umma_commit(done_bar);
        }
    } else {
        // ---- split: hi in place, lo behind it (elementwise: the swizzled positions carry over) -------------------------
        const int t128 = tid - 64;
        for (long long t = 0; t < my_tiles; ++t) {
            const int s = (int)(t % wt::kStages);
            mbar_wait(landed_bar(s), (unsigned)((t / wt::kStages) & 1));
            unsigned char *at = smem + s * wt::kStageBytes;
            constexpr int kChunks = wt::kRawBytes / 16 / 128;              // 24 chunks of 16 bytes per thread
#pragma unroll 4
            for (int q = 0; q < kChunks; ++q) {
                const int chunk = t128 + q * 128;                         // [0, 1024): dx, [1024, 3072): J
                const bool is_a = chunk < 2 * wt::kAtomBytes / 16;
                unsigned char *hi_at = at + (is_a ? wt::kAHi + chunk * 16 : wt::kBHi + (chunk - 2 * wt::kAtomBytes / 16) * 16);
                const int lo_offset = is_a ? wt::kALo - wt::kAHi : wt::kBLo - wt::kBHi;
                const float4 x = *reinterpret_cast<const float4 *>(hi_at);
                const float4 hi = make_float4(tc_tf32(x.x), tc_tf32(x.y), tc_tf32(x.z), tc_tf32(x.w));
                const float4 lo = make_float4(tc_tf32(x.x - hi.x), tc_tf32(x.y - hi.y), tc_tf32(x.z - hi.z), tc_tf32(x.w - hi.w));
                *reinterpret_cast<float4 *>(hi_at) = hi;
                *reinterpret_cast<float4 *>(hi_at + lo_offset) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar(s));
        }
        // ---- epilogue: rows n (hi half) + n + 64 (lo half) of the accumulator -> this CTA's partial result --------------
        const int quadrant = warp & 3, row = 32 * quadrant + lane;        // TMEM lane = row of the stacked operand
        float *scratch = reinterpret_cast<float *>(smem);
        float *mine = partial + (long long)blockIdx.x * 64 * 128;
        if (my_tiles > 0) {
            mbar_wait(done_bar, 0u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float v[64];
            if (my_tiles > 0) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float part[16];
                    tmem_load16(tmem_base + ((unsigned)(32 * quadrant) << 16) + (unsigned)(64 * h + 16 * c), part);
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[16 * c + i] = part[i];
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int i = 0; i < 64; ++i) v[i] = 0.f;
            }
            if (row >= 64) {
#pragma unroll
                for (int c = 0; c < 64; c += 4)
                    *reinterpret_cast<float4 *>(scratch + (row - 64) * wt::kScratchPitch + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (row < 64) {
#pragma unroll
                for (int c = 0; c < 64; c += 4) {
                    const float4 lo = *reinterpret_cast<const float4 *>(scratch + row * wt::kScratchPitch + c);
                    *reinterpret_cast<float4 *>(mine + row * 128 + 64 * h + c) = make_float4(v[c] + lo.x, v[c + 1] + lo.y, v[c + 2] + lo.z, v[c + 3] + lo.w);
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
    }
}

__global__ void weight_grad_fold_kernel(const float *__restrict__ partial, int n_partial, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 64 * 128) return;
    float total = 0.f;
    for (int p = 0; p < n_partial; ++p) total += partial[(long long)p * 64 * 128 + i];   // fixed order
    out[i] = total;
}

}  // namespace

}  // namespace ultra

using namespace ultra;

extern "C" int ultra_layer_rows_gemm(const float *dev_a0, int64_t lda0, const float *dev_a1, int64_t lda1, const float *dev_weight,
                               float *dev_out0, int64_t ld_out0, float *dev_out1, int64_t ld_out1, const float *dev_addend0,
                               int64_t ld_addend0, int64_t rows, int32_t n_out, int32_t k_in, void *stream) {
    if (rows < 0 || (rows > 0 && (!dev_a0 || !dev_weight || !dev_out0))) return ULTRA_RSPMM_ERR_ARG;
    if (!((n_out == 64 && k_in == 128) || (n_out == 128 && k_in == 64))) return ULTRA_RSPMM_ERR_RANGE;
    const bool two = dev_a1 != nullptr;
    const int64_t cols = two ? k_in / 2 : k_in;
    if (lda0 < cols || lda0 % 4 || (two && (lda1 < cols || lda1 % 4)) || ld_out0 % 4 || (dev_out1 && ld_out1 % 4) ||
        (dev_addend0 && ld_addend0 % 4))
        return ULTRA_RSPMM_ERR_ARG;
    if (((uintptr_t)dev_a0 | (uintptr_t)dev_a1 | (uintptr_t)dev_weight | (uintptr_t)dev_out0 | (uintptr_t)dev_out1 |
         (uintptr_t)dev_addend0) & 15)
        return ULTRA_RSPMM_ERR_ARG;
    if (rows == 0) return ULTRA_RSPMM_OK;
    RowsGemmArgs args = {};
    args.W = dev_weight;
    args.out0 = dev_out0; args.out1 = dev_out1; args.addend0 = dev_addend0;
    args.ld0 = ld_out0; args.ld1 = ld_out1; args.ld_addend = ld_addend0;
    args.rows = rows;
    args.two_sources = two ? 1 : 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (n_out == 64) return launch_rows_gemm<64, 128>(dev_a0, lda0, dev_a1, lda1, args, s);
    return launch_rows_gemm<128, 64>(dev_a0, lda0, dev_a1, lda1, args, s);
}

extern "C" int ultra_layer_rows_gemm_weight_bytes(size_t *workspace_bytes) {
    if (!workspace_bytes) return ULTRA_RSPMM_ERR_ARG;
    *workspace_bytes = (size_t)512 * 64 * 128 * sizeof(float);      // one partial result per CTA (at most 512)
    return ULTRA_RSPMM_OK;
}

extern "C" int ultra_layer_rows_gemm_weight(const float *dev_dx, int64_t ld_dx, const float *dev_a0, int64_t lda0, const float *dev_a1,
                                      int64_t lda1, int64_t rows, float *dev_weight_grad, void *workspace, size_t workspace_bytes,
                                      void *stream) {
    if (rows < 0 || !dev_weight_grad || (rows > 0 && (!dev_dx || !dev_a0 || !dev_a1))) return ULTRA_RSPMM_ERR_ARG;
    if (ld_dx < 64 || lda0 < 64 || lda1 < 64 || ld_dx % 4 || lda0 % 4 || lda1 % 4) return ULTRA_RSPMM_ERR_ARG;
    if (((uintptr_t)dev_dx | (uintptr_t)dev_a0 | (uintptr_t)dev_a1 | (uintptr_t)dev_weight_grad | (uintptr_t)workspace) & 15)
        return ULTRA_RSPMM_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    int device = 0, sm_count = 0;
    ULTRA_CUDA_OK(cudaGetDevice(&device));
    ULTRA_CUDA_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
    const long long n_tiles = (rows + wg::kTileRows - 1) / wg::kTileRows;
    static const bool legacy = getenv("ULTRA_WEIGHT_GRAD") && !strcmp(getenv("ULTRA_WEIGHT_GRAD"), "mma");
    const long long n_tiles_tc = (rows + wt::kTileRows - 1) / wt::kTileRows;
    int grid = legacy || rows == 0 ? (int)(n_tiles < sm_count ? n_tiles : sm_count)
                                   : (int)(n_tiles_tc < (long long)WT_CTAS * sm_count ? n_tiles_tc : (long long)WT_CTAS * sm_count);
    if (grid > 512) grid = 512;
    if (grid < 1) grid = 1;
    if (!workspace || workspace_bytes < (size_t)grid * 64 * 128 * sizeof(float)) return ULTRA_RSPMM_ERR_WORKSPACE;
    // ULTRA_WEIGHT_GRAD=mma keeps the mma.sync kernel (the first version: bound by the legacy tensor pipe at 3.5 TB/s)
    if (legacy || rows == 0) {
        ULTRA_CUDA_OK(cudaFuncSetAttribute(weight_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wg::kSmemBytes));
        weight_grad_kernel<<<grid, wg::kThreads, wg::kSmemBytes, s>>>(dev_dx, ld_dx, dev_a0, lda0, dev_a1, lda1, rows, (float *)workspace);
    } else {
        CUtensorMap map_dx, map_a0, map_a1;
        if (int status = encode_rows_map(&map_dx, dev_dx, rows, 64, ld_dx, wt::kTileRows, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return status;
        if (int status = encode_rows_map(&map_a0, dev_a0, rows, 64, lda0, wt::kTileRows, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return status;
        if (int status = encode_rows_map(&map_a1, dev_a1, rows, 64, lda1, wt::kTileRows, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return status;
        ULTRA_CUDA_OK(cudaFuncSetAttribute(weight_grad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wt::kSmemBytes));
        weight_grad_tc_kernel<<<grid, wt::kThreads, wt::kSmemBytes, s>>>(map_dx, map_a0, map_a1, rows, (float *)workspace);
    }
    note_launch();
    weight_grad_fold_kernel<<<(64 * 128 + 255) / 256, 256, 0, s>>>((const float *)workspace, grid, dev_weight_grad);
    note_launch();
    ULTRA_CUDA_OK(cudaGetLastError());
    return ULTRA_RSPMM_OK;
}
