// tcgen05 / TMEM / TMA version of the fused `combine` (see layer_linear.cu for what is computed and for the 3xTF32 arithmetic):
//     out[r, 0:N] = relu(layer_norm(A[r, 0:2N] @ W^T + b) * gamma + beta) + A[r, 0:N]
// The mma.sync version is bound by the legacy HMMA pipe (0.18 ms per C2 layer at best, 0.27 ms measured).  Here the MMAs
// run on the 5th-generation tensor cores (tcgen05.mma.kind::tf32, one issuing thread, accumulators in tensor memory:
// ~0.04 ms for the same work) and the 64 KB row tiles reach shared memory by TMA, so the load/store unit only sees the
// hi / lo split and the epilogue.
//
// Persistent CTAs, one per SM, warp-specialised (10 warps):
//   warp  0     TMA producer (one thread): per 128-row tile and 32-column K slot one cp.async.bulk.tensor.2d (box
//               32 floats x 128 rows = 16 KB, SWIZZLE_128B) into a ring of landing slots;
//   warps 2-5   split: one thread = one row of a landed slot (8 conflict-free 16-byte reads), hi = tf32(x) and the
//               remainder lo = tf32(x - hi) are written with tcgen05.st into a ring of A slots in TENSOR memory (64 columns
//               each: hi | lo); the landing slot is released as soon as its bytes are in registers;
//   warp  1     one elected thread issues, per slot, 4 k-steps x 3 MMAs (lo_a hi_b, hi_a lo_b, hi_a hi_b; M = 128, N, K = 8;
//               A from tensor memory, B = W by shared-memory descriptor) into one of two accumulator stages in TMEM and
//               commits them to the A slot's `lo_empty` barrier;
//   warps 6-9   epilogue: tcgen05.ld of their 32 TMEM lanes (one thread = one row, all N columns), LayerNorm statistics in
//               the thread, centred rows through warp-private shared memory, then coalesced: affine, ReLU, short-cut
//               (fp32 row re-read from global memory: an L2 hit), 16-byte streaming stores.
// Hand-offs are mbarriers: landed[slot] (TMA complete_tx), empty[slot] (one arrival per split warp), full[a_slot] (one
// arrival per split warp after tcgen05.wait::st), lo_empty[a_slot] (tcgen05.commit), tmem_full[stage] (tcgen05.commit),
// tmem_empty[stage] (one arrival per epilogue warp).  W is split into hi / lo once per CTA (no-swizzle K-major layout).
//
// Why A lives in tensor memory: the first version (kept as ring code 133) rewrote each landed slot in place as hi, put lo
// beside it in shared memory and issued the MMAs with both operands from shared memory.  Barrier-wait counters
// (ultra_layer_linear_set_debug, tools/linear_stalls.py) showed the TMA ring always full (the producer waited 79 % of the
// kernel) while the split warps and the MMA issuer were busy: a K = 8 tf32 MMA re-reads 4 KB of A and 2 KB of B from shared
// memory, 48 MMAs per tile = 288 KB, plus the split's 192 KB - shared-memory bandwidth, not HBM, was the bound (0.19 ms at
// C2).  With A in tensor memory the shared-memory traffic per tile drops to the landing write, one read by the split, W and
// the epilogue's staging (about 290 KB -> 2,300 cycles against the 4,270 the tile's HBM bytes take): 0.142 ms, 5.0 TB/s of
// compulsory bytes.  Ring shape 3 landing slots + 2 A slots measured best (sweep in profiles/r02_linear_ring_sweep.txt).
#include <cstdlib>

#include "tc_common.cuh"

namespace ultra {

long long *g_linear_debug = nullptr;   // development: per-CTA barrier wait cycles (ultra_layer_linear_set_debug)

namespace {

using namespace tcx;

// Development (compile with -DULTRA_LINEAR_KNOCKOUT): ULTRA_LINEAR_KNOCK = bit set of the work to leave out, to find which
// role bounds the kernel - 1 MMAs, 2 the split's arithmetic and stores, 4 the epilogue's global loads / stores, 8 the
// epilogue's arithmetic and staging too, 16 the TMA loads.  Results are garbage with any bit set; never in the shipped build.
#ifdef ULTRA_LINEAR_KNOCKOUT
__device__ int g_knock;
#define ULTRA_KNOCK(bit) (knock & (bit))
#else
#define ULTRA_KNOCK(bit) false
#endif

namespace tc {
constexpr int kRows = 128;                             // UMMA M
constexpr int kSlotK = 32;                             // K columns per ring slot = 4 k-steps of 8
constexpr int kTmaWarp = 0;
constexpr int kMmaWarp = 1;
constexpr int kSplitWarp0 = 2;
constexpr int kSplitWarps = 4;
constexpr int kEpilogueWarp0 = kSplitWarp0 + kSplitWarps;
constexpr int kEpilogueWarps = 4;
constexpr int kThreads = 32 * (kEpilogueWarp0 + kEpilogueWarps);
constexpr int kSlotHalfBytes = kRows * kSlotK * 4;     // hi (or lo) part of a slot: 16 KB = one TMA box

// HI landing / hi slots (16 KB each, the TMA loads in flight), LO lo tiles (live from the split to the end of their MMAs)
// TS: the split writes hi / lo into TENSOR memory (LO slots of 64 columns) and the MMAs take A from there (see the header)
template <int N, int HI, int LO, bool TS> struct Shape {
    static constexpr int kSlots = HI, kLoSlots = LO;
    static constexpr int kBarriers = TS ? 2 * HI + 2 * LO + 4 : 3 * HI + LO + 4;
    static constexpr int K = 2 * N;
    static constexpr int kSlotsPerTile = K / kSlotK;
    static constexpr int kWeightHalfBytes = N * K * 4;
    static constexpr int kCoreBytesW = N * 16;          // LBO of the W operand
    static constexpr int kRingOffset = (2 * kWeightHalfBytes + 1023) / 1024 * 1024;   // SWIZZLE_128B slots: 1024-byte aligned
    static constexpr int kStageStride = N + 4;          // floats per staged output row (bank-conflict-free both ways)
    static constexpr int kLoOffset = kRingOffset + kSlots * kSlotHalfBytes;
    static constexpr int kStagingOffset = kLoOffset + (TS ? 0 : kLoSlots) * kSlotHalfBytes;
    static constexpr int kBarrierOffset = kStagingOffset + kEpilogueWarps * 32 * kStageStride * 4;
    static constexpr int kSmemBytes = kBarrierOffset + kBarriers * 8 + 16;
    static_assert(kSmemBytes <= 227 * 1024, "ring does not fit the 227 KB of shared memory a CTA may use");
    static constexpr int kOperandColumn0 = 2 * N;      // TS: A slot s = columns [2N + 64 s, +32) hi, [.. + 32, +32) lo
    static constexpr int kTmemColumns = TS ? 512 : (2 * N < 32 ? 32 : 2 * N);   // two accumulator stages (+ A slots); power of two >= 32
    static_assert(!TS || 2 * N + 64 * LO <= 512, "accumulators and A slots exceed the 512 columns of tensor memory");
    // instruction descriptor: D = F32 (bits 4-5), A = B = TF32 (bits 7-9, 10-12), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    static constexpr unsigned kInstr = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(kRows >> 4) << 24);
};
}  // namespace tc

template <int N, int HI, int LO, bool TS, bool PRE>   // PRE: also write the Linear's output (training forward)
// 10 warps = 3 on two of the four SM sub-partitions: 16,384 / 3 / 32 = 170 registers per thread is the most a launch can
// get (ptxas stops at 168; __maxnreg__(192) compiles but fails to launch)
__global__ void __launch_bounds__(tc::kThreads, 1)
linear_norm_relu_residual_tc_kernel(const __grid_constant__ CUtensorMap a_map, const __grid_constant__ CUtensorMap a_map2,
                                    int two_sources, const float *__restrict__ A, long long lda,
                                    const float *__restrict__ W,
                                    const float *__restrict__ linear_bias, const float *__restrict__ gamma,
                                    const float *__restrict__ beta, float *__restrict__ out, long long ldo, long long rows,
                                    float eps, int relu, int shortcut, long long *debug, float *__restrict__ pre_out,
                                    long long ld_pre) {
    using S = tc::Shape<N, HI, LO, TS>;
#ifdef ULTRA_LINEAR_KNOCKOUT
    const int knock = g_knock;
#endif
    // debug (development): cycles each role spent waiting on its barriers, per CTA: [tma/empty, split/landed, split/lo_empty,
    // mma/tmem_empty, mma/full, epilogue/tmem_full, total]
    long long waited = 0, waited2 = 0;
    const long long kernel_start = debug ? clock64() : 0;
#define ULTRA_TIMED_WAIT(counter, ...)                 \
    do {                                               \
        if (debug) {                                   \
            const long long t0_ = clock64();           \
            __VA_ARGS__;                               \
            counter += clock64() - t0_;                \
        } else {                                       \
            __VA_ARGS__;                               \
        }                                              \
    } while (0)
    constexpr int K = S::K, kSlotsPerTile = S::kSlotsPerTile;
    extern __shared__ __align__(1024) unsigned char smem[];
    const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem);
    const unsigned bar_base = smem_base + S::kBarrierOffset;
    // SS: full / empty / landed per ring slot, lo_empty per lo tile.  TS: landed / empty per ring slot (empty = the split
    // warps have read it), full / lo_empty per tensor-memory A slot (full = hi and lo stored, lo_empty = its MMAs completed).
    auto landed_bar = [&](int slot) { return bar_base + 8u * slot; };
    auto empty_bar = [&](int slot) { return bar_base + 8u * (S::kSlots + slot); };
    auto full_bar = [&](int slot) { return bar_base + 8u * (2 * S::kSlots + slot); };
    constexpr int kFullBars = TS ? S::kLoSlots : S::kSlots;
    auto lo_empty_bar = [&](int slot) { return bar_base + 8u * (2 * S::kSlots + kFullBars + slot); };
    auto tmem_full_bar = [&](int stage) { return bar_base + 8u * (2 * S::kSlots + kFullBars + S::kLoSlots + stage); };
    auto tmem_empty_bar = [&](int stage) { return bar_base + 8u * (2 * S::kSlots + kFullBars + S::kLoSlots + 2 + stage); };
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(smem + S::kBarrierOffset + S::kBarriers * 8);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- one-time setup: barriers, tensor memory, W split into hi / lo in UMMA layout ----------------------------------
    if (tid == 0) {
        for (int s = 0; s < S::kSlots; ++s) {
            mbar_init(empty_bar(s), TS ? tc::kSplitWarps : 1);
            mbar_init(landed_bar(s), 1);
        }
        for (int s = 0; s < kFullBars; ++s) mbar_init(full_bar(s), tc::kSplitWarps);
        for (int s = 0; s < S::kLoSlots; ++s) mbar_init(lo_empty_bar(s), 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(tmem_full_bar(s), 1);
            mbar_init(tmem_empty_bar(s), tc::kEpilogueWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == tc::kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (unsigned)__cvta_generic_to_shared(tmem_slot)), "r"((unsigned)S::kTmemColumns) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int idx = tid; idx < N * (K / 4); idx += tc::kThreads) {
        const int n = idx / (K / 4), kc = idx % (K / 4);
        const float4 w = __ldg(reinterpret_cast<const float4 *>(W + n * K + 4 * kc));
        const float4 hi = make_float4(tc_tf32(w.x), tc_tf32(w.y), tc_tf32(w.z), tc_tf32(w.w));
        const float4 lo = make_float4(tc_tf32(w.x - hi.x), tc_tf32(w.y - hi.y), tc_tf32(w.z - hi.z), tc_tf32(w.w - hi.w));
        unsigned char *at = smem + kc * S::kCoreBytesW + n * 16;
        *reinterpret_cast<float4 *>(at) = hi;
        *reinterpret_cast<float4 *>(at + S::kWeightHalfBytes) = lo;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> visible to the tensor cores
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = *tmem_slot;

    const long long n_tiles = (rows + tc::kRows - 1) / tc::kRows;
    const long long first = blockIdx.x;
    const long long my_tiles = first < n_tiles ? (n_tiles - first + gridDim.x - 1) / gridDim.x : 0;

    if (warp == tc::kTmaWarp) {
        // ===== TMA producer (one thread) ================================================================================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&a_map) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&a_map2) : "memory");
            const long long total = my_tiles * kSlotsPerTile;
            for (long long it = 0; it < total; ++it) {
                const int slot = (int)(it % S::kSlots);
                const unsigned phase = (unsigned)((it / S::kSlots) & 1);
                const long long tile = first + (it / kSlotsPerTile) * gridDim.x;
                // two sources: the K slots of the first half come from the layer input, those of the second from the update
                const int q = (int)(it % kSlotsPerTile);
                const bool second = two_sources && q >= kSlotsPerTile / 2;
                const int k0 = (second ? q - kSlotsPerTile / 2 : q) * tc::kSlotK;
                ULTRA_TIMED_WAIT(waited, mbar_wait(empty_bar(slot), phase ^ 1u));   // the MMAs that read this slot have completed
                if (ULTRA_KNOCK(16)) { mbar_arrive(landed_bar(slot)); continue; }
                mbar_expect_tx(landed_bar(slot), tc::kSlotHalfBytes);
                tma_load_2d(smem_base + S::kRingOffset + slot * tc::kSlotHalfBytes, second ? &a_map2 : &a_map, k0, (int)(tile * tc::kRows),
                            landed_bar(slot));                           // rows past the end are filled with zeros
            }
        }
    } else if (warp >= tc::kSplitWarp0 && warp < tc::kEpilogueWarp0) {
        // ===== split: hi in place, lo beside it (elementwise, so the swizzled positions carry over) ========================
        const int t = tid - 32 * tc::kSplitWarp0;                        // 0 .. 127
        const long long total = my_tiles * kSlotsPerTile;
        for (long long it = 0; it < total; ++it) {
            const int slot = (int)(it % S::kSlots);
            const unsigned phase = (unsigned)((it / S::kSlots) & 1);
            const int lo_slot = (int)(it % S::kLoSlots);
            const unsigned lo_phase = (unsigned)((it / S::kLoSlots) & 1);
            ULTRA_TIMED_WAIT(waited, mbar_wait(landed_bar(slot), phase));
            ULTRA_TIMED_WAIT(waited2, mbar_wait(lo_empty_bar(lo_slot), lo_phase ^ 1u));   // the MMAs that read this lo tile have completed
            unsigned char *hi_at = smem + S::kRingOffset + slot * tc::kSlotHalfBytes;
            if constexpr (TS) {
                // one thread = one row of the slot (the tensor-memory lane this warp may write): its 128 bytes are 8 chunks
                // at (chunk ^ row % 8) - a quarter-warp reads 8 different bank groups - then hi / lo go to TMEM columns
                const int quadrant = warp & 3, r = 32 * quadrant + lane;
                if (ULTRA_KNOCK(2)) {
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(empty_bar(slot)); mbar_arrive(full_bar(lo_slot)); }
                    continue;
                }
                float4 x[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) x[c] = *reinterpret_cast<const float4 *>(hi_at + r * 128 + ((c ^ (r & 7)) << 4));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned taddr = tmem_base + ((unsigned)(32 * quadrant) << 16) + (unsigned)(S::kOperandColumn0 + 64 * lo_slot);
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
                if (lane == 0) mbar_arrive(empty_bar(slot));             // the slot's bytes are in registers: TMA may refill it
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(full_bar(lo_slot));
                continue;
            }
            unsigned char *lo_at = smem + S::kLoOffset + lo_slot * tc::kSlotHalfBytes;
            constexpr int kChunksPerThread = tc::kSlotHalfBytes / 16 / (32 * tc::kSplitWarps);
            float4 x[kChunksPerThread];
#pragma unroll
            for (int q = 0; q < kChunksPerThread; ++q)
                x[q] = *reinterpret_cast<const float4 *>(hi_at + 16 * (t + q * 32 * tc::kSplitWarps));
#pragma unroll
            for (int q = 0; q < kChunksPerThread; ++q) {
                const float4 hi = make_float4(tc_tf32(x[q].x), tc_tf32(x[q].y), tc_tf32(x[q].z), tc_tf32(x[q].w));
                const float4 lo = make_float4(tc_tf32(x[q].x - hi.x), tc_tf32(x[q].y - hi.y), tc_tf32(x[q].z - hi.z),
                                              tc_tf32(x[q].w - hi.w));
                const int offset = 16 * (t + q * 32 * tc::kSplitWarps);
                *reinterpret_cast<float4 *>(hi_at + offset) = hi;
                *reinterpret_cast<float4 *>(lo_at + offset) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor cores
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar(slot));
        }
    } else if (warp == tc::kMmaWarp) {
        // ===== MMA issuer (one thread) ==================================================================================
        if (lane == 0) {
            const unsigned w_hi = smem_base, w_lo = smem_base + S::kWeightHalfBytes;
            long long it = 0;
            for (long long t = 0; t < my_tiles; ++t) {
                const int stage = (int)(t & 1);
                const unsigned accum_phase = (unsigned)((t >> 1) & 1);
                ULTRA_TIMED_WAIT(waited, mbar_wait(tmem_empty_bar(stage), accum_phase ^ 1u));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned tmem_d = tmem_base + (unsigned)(stage * N);
                for (int q = 0; q < kSlotsPerTile; ++q, ++it) {
                    if constexpr (TS) {
                        const int a_slot = (int)(it % S::kLoSlots);
                        ULTRA_TIMED_WAIT(waited2, mbar_wait(full_bar(a_slot), (unsigned)((it / S::kLoSlots) & 1)));
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const unsigned a_hi = tmem_base + (unsigned)(S::kOperandColumn0 + 64 * a_slot), a_lo = a_hi + 32u;
#pragma unroll
                        for (int ks = 0; ks < tc::kSlotK / 8; ++ks) {
                            const int kg = q * (tc::kSlotK / 8) + ks;
                            const unsigned long long db_hi = umma_desc(w_hi + kg * 2 * S::kCoreBytesW, S::kCoreBytesW, 128);
                            const unsigned long long db_lo = umma_desc(w_lo + kg * 2 * S::kCoreBytesW, S::kCoreBytesW, 128);
                            if (ULTRA_KNOCK(1)) continue;
                            umma_tf32_ts(tmem_d, a_lo + 8u * ks, db_hi, S::kInstr, kg > 0 ? 1u : 0u);
                            umma_tf32_ts(tmem_d, a_hi + 8u * ks, db_lo, S::kInstr, 1u);
                            umma_tf32_ts(tmem_d, a_hi + 8u * ks, db_hi, S::kInstr, 1u);
                        }
                        umma_commit(lo_empty_bar(a_slot));               // the A slot may be overwritten once these MMAs have read it
                        continue;
                    }
                    const int slot = (int)(it % S::kSlots);
                    const unsigned phase = (unsigned)((it / S::kSlots) & 1);
                    ULTRA_TIMED_WAIT(waited2, mbar_wait(full_bar(slot), phase));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const int lo_slot = (int)(it % S::kLoSlots);
                    const unsigned a_hi = smem_base + S::kRingOffset + slot * tc::kSlotHalfBytes;
                    const unsigned a_lo = smem_base + S::kLoOffset + lo_slot * tc::kSlotHalfBytes;
#pragma unroll
                    for (int ks = 0; ks < tc::kSlotK / 8; ++ks) {
                        const int kg = q * (tc::kSlotK / 8) + ks;        // k-step within the tile: two core matrices each
                        const unsigned long long da_hi = umma_desc_sw128(a_hi + ks * 32);
                        const unsigned long long da_lo = umma_desc_sw128(a_lo + ks * 32);
                        const unsigned long long db_hi = umma_desc(w_hi + kg * 2 * S::kCoreBytesW, S::kCoreBytesW, 128);
                        const unsigned long long db_lo = umma_desc(w_lo + kg * 2 * S::kCoreBytesW, S::kCoreBytesW, 128);
                        umma_tf32(tmem_d, da_lo, db_hi, S::kInstr, kg > 0 ? 1u : 0u);
                        umma_tf32(tmem_d, da_hi, db_lo, S::kInstr, 1u);
                        umma_tf32(tmem_d, da_hi, db_hi, S::kInstr, 1u);
                    }
                    umma_commit(empty_bar(slot));                        // slot free once these MMAs have read it
                    umma_commit(lo_empty_bar(lo_slot));
                }
                umma_commit(tmem_full_bar(stage));                       // accumulator complete
            }
        }
    } else if (warp >= tc::kEpilogueWarp0) {
        // ===== epilogue =================================================================================================
        // Row phase: one thread = one row (what tcgen05.ld.32x32b hands out): bias, mean, rstd in the thread, centred row
        // into this warp's private staging rows.  Write-back phase: the warp walks its 32 rows kRowsPerPass at a time with
        // lane = (row, 16-byte chunk), so the short-cut loads and the result stores are coalesced (the first version stored
        // straight from the row phase: 32 cache lines per instruction, L1 data pipe 87 % busy).
        constexpr int kChunks = N / 4, kRowsPerPass = 32 / kChunks, kStride = S::kStageStride;
        const int quadrant = warp & 3;                                   // the TMEM lanes this warp may read
        float *staged = reinterpret_cast<float *>(smem + S::kStagingOffset) + quadrant * 32 * kStride;
        const int my_chunk = lane % kChunks, sub_row = lane / kChunks;
        float4 scale = make_float4(1.f, 1.f, 1.f, 1.f), shift = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gamma) {
            scale = __ldg(reinterpret_cast<const float4 *>(gamma + 4 * my_chunk));
            shift = __ldg(reinterpret_cast<const float4 *>(beta + 4 * my_chunk));
        }
        if (linear_bias) bias4 = __ldg(reinterpret_cast<const float4 *>(linear_bias + 4 * my_chunk));
        constexpr float inv = 1.0f / N;
        for (long long t = 0; t < my_tiles; ++t) {
            const int stage = (int)(t & 1);
            const unsigned accum_phase = (unsigned)((t >> 1) & 1);
            const long long row0 = (first + t * gridDim.x) * tc::kRows + 32 * quadrant;
            ULTRA_TIMED_WAIT(waited, mbar_wait(tmem_full_bar(stage), accum_phase));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float v[N];
#pragma unroll
            for (int c = 0; c < N / 16; ++c) {
                float part[16];
                tmem_load16(tmem_base + ((unsigned)(32 * quadrant) << 16) + (unsigned)(stage * N + 16 * c), part);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[16 * c + i] = part[i];
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar(stage));           // the MMAs of tile t + 2 may overwrite this stage
            if (ULTRA_KNOCK(8)) {
                if (v[0] + v[N - 1] == 123.456f) out[0] = v[0];
                continue;
            }
            float mean = 0.f, rstd = 0.f;
            if constexpr (PRE) {
                // the accumulators themselves (the Linear's output without its bias) are what is staged - a training caller keeps
                // them for the backward (pre_out) - the bias and the mean are applied again in the write-back phase, with the
                // same operations in the same order
                float sum = 0.f;
#pragma unroll
                for (int c = 0; c < N; c += 4) {
                    float4 lb = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (linear_bias) lb = __ldg(reinterpret_cast<const float4 *>(linear_bias + c));
                    sum += ((v[c] + lb.x) + (v[c + 1] + lb.y)) + ((v[c + 2] + lb.z) + (v[c + 3] + lb.w));
                }
                mean = sum * inv;
                float sq = 0.f;
#pragma unroll
                for (int c = 0; c < N; c += 4) {
                    float4 lb = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (linear_bias) lb = __ldg(reinterpret_cast<const float4 *>(linear_bias + c));
                    const float c0 = (v[c] + lb.x) - mean, c1 = (v[c + 1] + lb.y) - mean, c2 = (v[c + 2] + lb.z) - mean,
                                c3 = (v[c + 3] + lb.w) - mean;
                    sq = fmaf(c0, c0, sq); sq = fmaf(c1, c1, sq); sq = fmaf(c2, c2, sq); sq = fmaf(c3, c3, sq);
                }
                rstd = rsqrtf(sq * inv + eps);
#pragma unroll
                for (int c = 0; c < N; c += 4)
                    *reinterpret_cast<float4 *>(staged + lane * kStride + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
            } else {
                float sum = 0.f;
#pragma unroll
                for (int c = 0; c < N; c += 4) {
                    if (linear_bias) {
                        const float4 lb = __ldg(reinterpret_cast<const float4 *>(linear_bias + c));
                        v[c] += lb.x; v[c + 1] += lb.y; v[c + 2] += lb.z; v[c + 3] += lb.w;
                    }
                    sum += (v[c] + v[c + 1]) + (v[c + 2] + v[c + 3]);
                }
                mean = sum * inv;
                float sq = 0.f;
#pragma unroll
                for (int c = 0; c < N; ++c) {
                    v[c] -= mean;
                    sq = fmaf(v[c], v[c], sq);
                }
                rstd = rsqrtf(sq * inv + eps);
#pragma unroll
                for (int c = 0; c < N; c += 4)
                    *reinterpret_cast<float4 *>(staged + lane * kStride + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
            }
            __syncwarp();
            constexpr int kPasses = 32 / kRowsPerPass;
            float4 skip[kPasses];                                        // all short-cut loads in flight before the first use
#pragma unroll
            for (int pass = 0; pass < kPasses; ++pass) {
                const long long row = row0 + pass * kRowsPerPass + sub_row;
                skip[pass] = shortcut && row < rows && !ULTRA_KNOCK(4) ? __ldg(reinterpret_cast<const float4 *>(A + row * lda + 4 * my_chunk))
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int pass = 0; pass < kPasses; ++pass) {
                const int r = pass * kRowsPerPass + sub_row;
                float4 x;
                const float rs = __shfl_sync(kFullMask, rstd, r);
                if constexpr (PRE) {
                    const float mu = __shfl_sync(kFullMask, mean, r);
                    const float4 raw = *reinterpret_cast<const float4 *>(staged + r * kStride + 4 * my_chunk);
                    if (pre_out && row0 + r < rows) __stcs(reinterpret_cast<float4 *>(pre_out + (row0 + r) * ld_pre + 4 * my_chunk), raw);
                    x = make_float4((raw.x + bias4.x) - mu, (raw.y + bias4.y) - mu, (raw.z + bias4.z) - mu, (raw.w + bias4.w) - mu);
                } else {
                    x = *reinterpret_cast<const float4 *>(staged + r * kStride + 4 * my_chunk);
                }
                float4 y = make_float4(fmaf(x.x * rs, scale.x, shift.x), fmaf(x.y * rs, scale.y, shift.y),
                                       fmaf(x.z * rs, scale.z, shift.z), fmaf(x.w * rs, scale.w, shift.w));
                if (relu) y = make_float4(fmaxf(y.x, 0.f), fmaxf(y.y, 0.f), fmaxf(y.z, 0.f), fmaxf(y.w, 0.f));
                y = make_float4(y.x + skip[pass].x, y.y + skip[pass].y, y.z + skip[pass].z, y.w + skip[pass].w);
                if (row0 + r < rows && (!ULTRA_KNOCK(4) || y.x == 123.456f))
                    __stcs(reinterpret_cast<float4 *>(out + (row0 + r) * ldo + 4 * my_chunk), y);
            }
            __syncwarp();                                                // staging rows are rewritten by the next tile
        }
    }

    if (debug) {
        long long *mine = debug + 8 * (long long)blockIdx.x;
        if (warp == tc::kTmaWarp && lane == 0) mine[0] = waited;
        if (warp == tc::kSplitWarp0 && lane == 0) { mine[1] = waited; mine[2] = waited2; }
        if (warp == tc::kMmaWarp && lane == 0) { mine[3] = waited; mine[4] = waited2; }
        if (warp == tc::kEpilogueWarp0 && lane == 0) { mine[5] = waited; mine[6] = clock64() - kernel_start; mine[7] = my_tiles; }
    }
#undef ULTRA_TIMED_WAIT
    // ---- teardown -------------------------------------------------------------------------------------------------------
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == tc::kMmaWarp) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((unsigned)S::kTmemColumns) : "memory");
    }
}

template <int N, int HI, int LO, bool TS, bool PRE>
int launch_linear_tc(const float *A, long long lda, const float *A1, long long lda1, const float *W, const float *linear_bias, const float *gamma,
                     const float *beta, float *out, long long ldo, long long rows, float eps, int relu, int shortcut,
                     cudaStream_t stream, float *pre_out, long long ld_pre) {
    using S = tc::Shape<N, HI, LO, TS>;
    auto kernel = linear_norm_relu_residual_tc_kernel<N, HI, LO, TS, PRE>;
    int device = 0, sm_count = 0;
    ULTRA_CUDA_OK(cudaGetDevice(&device));
    ULTRA_CUDA_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
    ULTRA_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kSmemBytes));
    // the (rows, 2N) operand as a 2-D tensor: inner dimension = the 2N columns the Linear reads, rows lda floats apart
    CUtensorMap map, map2;
    const bool two = A1 != nullptr;
    if (int status = encode_rows_map(&map, A, rows, two ? S::K / 2 : S::K, lda)) return status;
    if (int status = encode_rows_map(&map2, two ? A1 : A, rows, two ? S::K / 2 : S::K, two ? lda1 : lda)) return status;
    const long long n_tiles = (rows + tc::kRows - 1) / tc::kRows;
    const unsigned grid = (unsigned)(n_tiles < sm_count ? n_tiles : sm_count);
#ifdef ULTRA_LINEAR_KNOCKOUT
    const int knock = getenv("ULTRA_LINEAR_KNOCK") ? atoi(getenv("ULTRA_LINEAR_KNOCK")) : 0;
    ULTRA_CUDA_OK(cudaMemcpyToSymbolAsync(g_knock, &knock, sizeof(int), 0, cudaMemcpyHostToDevice, stream));
#endif
    kernel<<<grid, tc::kThreads, S::kSmemBytes, stream>>>(map, map2, two ? 1 : 0, A, lda, W, linear_bias, gamma, beta, out, ldo, rows, eps, relu,
                                                         shortcut, g_linear_debug, pre_out, ld_pre);
    note_launch();
    ULTRA_CUDA_OK(cudaGetLastError());
    return ULTRA_RSPMM_OK;
}

}  // namespace

// called by ultra_layer_linear_norm_relu_residual (layer_linear.cu) after its argument checks
int layer_linear_tc(const float *A, long long lda, const float *A1, long long lda1, const float *W, const float *linear_bias, const float *gamma,
                    const float *beta, float *out, long long ldo, long long rows, int out_dim, float eps, int relu,
                    int shortcut, cudaStream_t stream, float *pre_out, long long ld_pre) {
    // ULTRA_LINEAR_RING (development knob): "<ring slots><A slots>" with A in tensor memory, or 133 = the first version
    // (hi / lo tiles in shared memory, both operands read from there)
    static const int ring = getenv("ULTRA_LINEAR_RING") ? atoi(getenv("ULTRA_LINEAR_RING")) : 32;
#define ULTRA_TC(N, HI, LO, TS) launch_linear_tc<N, HI, LO, TS, false>(A, lda, A1, lda1, W, linear_bias, gamma, beta, out, ldo, rows, eps, relu, shortcut, stream, pre_out, ld_pre)
    if (pre_out)    // training forward: default ring only
        return out_dim == 64 ? launch_linear_tc<64, 3, 2, true, true>(A, lda, A1, lda1, W, linear_bias, gamma, beta, out, ldo, rows, eps, relu, shortcut, stream, pre_out, ld_pre)
                             : launch_linear_tc<32, 3, 2, true, true>(A, lda, A1, lda1, W, linear_bias, gamma, beta, out, ldo, rows, eps, relu, shortcut, stream, pre_out, ld_pre);
    if (out_dim == 64) {
        switch (ring) {
            case 133: return ULTRA_TC(64, 3, 3, false);
            case 22: return ULTRA_TC(64, 2, 2, true);
            case 33: return ULTRA_TC(64, 3, 3, true);
            case 42: return ULTRA_TC(64, 4, 2, true);
            case 44: return ULTRA_TC(64, 4, 4, true);
            case 53: return ULTRA_TC(64, 5, 3, true);
            case 64: return ULTRA_TC(64, 6, 4, true);
            case 43: return ULTRA_TC(64, 4, 3, true);
            default: return ULTRA_TC(64, 3, 2, true);
        }
    }
    return ring == 133 ? ULTRA_TC(32, 3, 3, false) : ULTRA_TC(32, 3, 2, true);
#undef ULTRA_TC
}

}  // namespace ultra

extern "C" int ultra_layer_linear_set_debug(long long *dev_buffer) {
    ultra::g_linear_debug = dev_buffer;    // 8 int64 per CTA (at least 8 * SM count), or NULL to switch the instrumentation off
    return ULTRA_RSPMM_OK;
}
