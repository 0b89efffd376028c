// tcgen05 / TMEM version of the fused `combine` (see layer_linear.cu for what is computed and for the 3xTF32 arithmetic):
//     out[r, 0:N] = relu(layer_norm(A[r, 0:2N] @ W^T + b) * gamma + beta) + A[r, 0:N]
// The mma.sync version is bound by the legacy HMMA pipe (0.18 ms per C2 layer at best, 0.27 ms measured); here the MMAs run
// on the 5th-generation tensor cores (tcgen05.mma.kind::tf32, one issuing thread, accumulators in tensor memory), which
// need ~0.04 ms for the same work, so the kernel is bound by what it must move: 0.71 GB per C2 layer.
//
// Persistent CTAs, one per SM, warp-specialised (13 warps):
//   warps 0-7   producers: read 128-row tiles of A from global memory (each warp instruction = 8 rows x 64 B), split every
//               value into hi = tf32(x) and lo = tf32(x - hi), and store both in shared memory in the canonical K-major
//               no-swizzle UMMA layout (8-row x 16-byte core matrices), a ring of 3 slots of 32 K-columns each;
//   warp  8     one elected thread issues, per slot, 4 k-steps x 3 MMAs (lo_a hi_b, hi_a lo_b, hi_a hi_b; M = 128, N, K = 8)
//               into one of two accumulator stages in TMEM and commits them to the slot's `empty` barrier;
//   warps 9-12  epilogue: tcgen05.ld of their 32 TMEM lanes (one thread = one row, all N columns), LayerNorm statistics in
//               the thread, centred rows through warp-private shared memory, then coalesced: affine, ReLU, short-cut
//               (fp32 row re-read from global memory: an L2 hit), 16-byte streaming stores.
// Hand-offs are mbarriers: full[slot] (one arrival per producer warp), empty[slot] (tcgen05.commit), tmem_full[stage]
// (tcgen05.commit), tmem_empty[stage] (one arrival per epilogue warp).  W is split into hi / lo once per CTA.
#include "rspmm_common.cuh"

namespace ultra {

namespace {

namespace tc {
constexpr int kRows = 128;                             // UMMA M
constexpr int kSlotK = 32;                             // K columns per ring slot = 4 k-steps of 8
constexpr int kSlots = 3;
constexpr int kProducerWarps = 8;
constexpr int kMmaWarp = kProducerWarps;
constexpr int kEpilogueWarps = 4;
constexpr int kThreads = 32 * (kProducerWarps + 1 + kEpilogueWarps);
constexpr int kSlotHalfBytes = kRows * kSlotK * 4;     // hi (or lo) part of a slot: 16 KB
constexpr int kCoreBytesA = kRows * 16;                // K-direction core-matrix stride of an A slot (LBO)
constexpr int kBarriers = 2 * kSlots + 4;

template <int N> struct Shape {
    static constexpr int K = 2 * N;
    static constexpr int kSlotsPerTile = K / kSlotK;
    static constexpr int kWeightHalfBytes = N * K * 4;
    static constexpr int kCoreBytesW = N * 16;          // LBO of the W operand
    static constexpr int kRingOffset = 2 * kWeightHalfBytes;
    static constexpr int kStageStride = N + 4;          // floats per staged output row (bank-conflict-free both ways)
    static constexpr int kStagingOffset = kRingOffset + kSlots * 2 * kSlotHalfBytes;
    static constexpr int kBarrierOffset = kStagingOffset + kEpilogueWarps * 32 * kStageStride * 4;
    static constexpr int kSmemBytes = kBarrierOffset + kBarriers * 8 + 16;
    static constexpr int kTmemColumns = 2 * N < 32 ? 32 : 2 * N;   // two accumulator stages; power of two >= 32
    // instruction descriptor: D = F32 (bits 4-5), A = B = TF32 (bits 7-9, 10-12), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    static constexpr unsigned kInstr = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(kRows >> 4) << 24);
};
}  // namespace tc

__device__ __forceinline__ float tc_tf32(float x) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle: 8-row x 16-byte core matrices; LBO = byte distance between core
// matrices adjacent in K, SBO = between 8-row groups (both / 16); bits 46-47 = descriptor version 1 (Blackwell)
__device__ __forceinline__ unsigned long long umma_desc(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes) {
    return (unsigned long long)((smem_addr >> 4) & 0x3fff) | ((unsigned long long)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((unsigned long long)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long a, unsigned long long b, unsigned instr,
                                          unsigned accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}" ::"r"(tmem_d), "l"(a), "l"(b), "r"(instr), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_load16(unsigned taddr, float (&v)[16]) {
    unsigned r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

template <int N>
__global__ void __launch_bounds__(tc::kThreads, 1)
linear_norm_relu_residual_tc_kernel(const float *__restrict__ A, long long lda, const float *__restrict__ W,
                                    const float *__restrict__ linear_bias, const float *__restrict__ gamma,
                                    const float *__restrict__ beta, float *__restrict__ out, long long ldo, long long rows,
                                    float eps, int relu, int shortcut) {
    using S = tc::Shape<N>;
    constexpr int K = S::K, kSlotsPerTile = S::kSlotsPerTile;
    extern __shared__ __align__(128) unsigned char smem[];
    const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem);
    const unsigned bar_base = smem_base + S::kBarrierOffset;
    auto full_bar = [&](int slot) { return bar_base + 8u * slot; };
    auto empty_bar = [&](int slot) { return bar_base + 8u * (tc::kSlots + slot); };
    auto tmem_full_bar = [&](int stage) { return bar_base + 8u * (2 * tc::kSlots + stage); };
    auto tmem_empty_bar = [&](int stage) { return bar_base + 8u * (2 * tc::kSlots + 2 + stage); };
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(smem + S::kBarrierOffset + tc::kBarriers * 8);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- one-time setup: barriers, tensor memory, W split into hi / lo in UMMA layout ----------------------------------
    if (tid == 0) {
        for (int s = 0; s < tc::kSlots; ++s) {
            mbar_init(full_bar(s), tc::kProducerWarps);
            mbar_init(empty_bar(s), 1);
        }
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

    if (warp < tc::kProducerWarps) {
        // ===== producers ================================================================================================
        // A warp instruction covers 8 rows x 4 chunks of 16 B; thread (warp, i) handles row group (4 warp + i) / 2 and the
        // K-half (4 warp + i) % 2 of the slot: global reads of 64 B per row, shared-memory stores of 128 contiguous bytes
        // per quarter warp (conflict-free).
        const long long total = my_tiles * kSlotsPerTile;
        const int r_in_group = lane & 7, chunk_in_half = lane >> 3;
        auto issue = [&](long long it, float4 (&regs)[4]) {
            const long long tile = first + (it / kSlotsPerTile) * gridDim.x;
            const int k0 = (int)(it % kSlotsPerTile) * tc::kSlotK;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int idx = 4 * warp + i;
                const long long row = tile * tc::kRows + (idx >> 1) * 8 + r_in_group;
                const int kc = (idx & 1) * 4 + chunk_in_half;
                regs[i] = row < rows ? __ldcs(reinterpret_cast<const float4 *>(A + row * lda + k0 + 4 * kc))
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto consume = [&](long long it, const float4 (&regs)[4]) {
            const int slot = (int)(it % tc::kSlots);
            const unsigned phase = (unsigned)((it / tc::kSlots) & 1);
            mbar_wait(empty_bar(slot), phase ^ 1u);
            unsigned char *slot_hi = smem + S::kRingOffset + slot * 2 * tc::kSlotHalfBytes;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int idx = 4 * warp + i;
                const int row = (idx >> 1) * 8 + r_in_group, kc = (idx & 1) * 4 + chunk_in_half;
                const float4 x = regs[i];
                const float4 hi = make_float4(tc_tf32(x.x), tc_tf32(x.y), tc_tf32(x.z), tc_tf32(x.w));
                const float4 lo = make_float4(tc_tf32(x.x - hi.x), tc_tf32(x.y - hi.y), tc_tf32(x.z - hi.z), tc_tf32(x.w - hi.w));
                unsigned char *at = slot_hi + kc * tc::kCoreBytesA + row * 16;
                *reinterpret_cast<float4 *>(at) = hi;
                *reinterpret_cast<float4 *>(at + tc::kSlotHalfBytes) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar(slot));                  // one arrival per warp: 8 per slot, not 256
        };
        float4 r0[4], r1[4], r2[4];                     // three slots of loads in flight per thread
        if (total > 0) issue(0, r0);
        if (total > 1) issue(1, r1);
        for (long long it = 0; it < total; it += 3) {
            if (it + 2 < total) issue(it + 2, r2);
            consume(it, r0);
            if (it + 1 < total) {
                if (it + 3 < total) issue(it + 3, r0);
                consume(it + 1, r1);
            }
            if (it + 2 < total) {
                if (it + 4 < total) issue(it + 4, r1);
                consume(it + 2, r2);
            }
        }
    } else if (warp == tc::kMmaWarp) {
        // ===== MMA issuer (one thread) ==================================================================================
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
                    const int slot = (int)(it % tc::kSlots);
                    const unsigned phase = (unsigned)((it / tc::kSlots) & 1);
                    mbar_wait(full_bar(slot), phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const unsigned a_hi = smem_base + S::kRingOffset + slot * 2 * tc::kSlotHalfBytes;
                    const unsigned a_lo = a_hi + tc::kSlotHalfBytes;
#pragma unroll
                    for (int ks = 0; ks < tc::kSlotK / 8; ++ks) {
                        const int kg = q * (tc::kSlotK / 8) + ks;        // k-step within the tile: two core matrices each
                        const unsigned long long da_hi = umma_desc(a_hi + ks * 2 * tc::kCoreBytesA, tc::kCoreBytesA, 128);
                        const unsigned long long da_lo = umma_desc(a_lo + ks * 2 * tc::kCoreBytesA, tc::kCoreBytesA, 128);
                        const unsigned long long db_hi = umma_desc(w_hi + kg * 2 * S::kCoreBytesW, S::kCoreBytesW, 128);
                        const unsigned long long db_lo = umma_desc(w_lo + kg * 2 * S::kCoreBytesW, S::kCoreBytesW, 128);
                        umma_tf32(tmem_d, da_lo, db_hi, S::kInstr, kg > 0 ? 1u : 0u);
                        umma_tf32(tmem_d, da_hi, db_lo, S::kInstr, 1u);
                        umma_tf32(tmem_d, da_hi, db_hi, S::kInstr, 1u);
                    }
                    umma_commit(empty_bar(slot));                        // slot free once these MMAs have read it
                }
                umma_commit(tmem_full_bar(stage));                       // accumulator complete
            }
        }
    } else {
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
        if (gamma) {
            scale = __ldg(reinterpret_cast<const float4 *>(gamma + 4 * my_chunk));
            shift = __ldg(reinterpret_cast<const float4 *>(beta + 4 * my_chunk));
        }
        constexpr float inv = 1.0f / N;
        for (long long t = 0; t < my_tiles; ++t) {
            const int stage = (int)(t & 1);
            const unsigned accum_phase = (unsigned)((t >> 1) & 1);
            const long long row0 = (first + t * gridDim.x) * tc::kRows + 32 * quadrant;
            mbar_wait(tmem_full_bar(stage), accum_phase);
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
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < N; c += 4) {
                if (linear_bias) {
                    const float4 lb = __ldg(reinterpret_cast<const float4 *>(linear_bias + c));
                    v[c] += lb.x; v[c + 1] += lb.y; v[c + 2] += lb.z; v[c + 3] += lb.w;
                }
                sum += (v[c] + v[c + 1]) + (v[c + 2] + v[c + 3]);
            }
            const float mean = sum * inv;
            float sq = 0.f;
#pragma unroll
            for (int c = 0; c < N; ++c) {
                v[c] -= mean;
                sq = fmaf(v[c], v[c], sq);
            }
            const float rstd = rsqrtf(sq * inv + eps);
#pragma unroll
            for (int c = 0; c < N; c += 4)
                *reinterpret_cast<float4 *>(staged + lane * kStride + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
            __syncwarp();
            constexpr int kPasses = 32 / kRowsPerPass;
            float4 skip[kPasses];                                        // all short-cut loads in flight before the first use
#pragma unroll
            for (int pass = 0; pass < kPasses; ++pass) {
                const long long row = row0 + pass * kRowsPerPass + sub_row;
                skip[pass] = shortcut && row < rows ? __ldg(reinterpret_cast<const float4 *>(A + row * lda + 4 * my_chunk))
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int pass = 0; pass < kPasses; ++pass) {
                const int r = pass * kRowsPerPass + sub_row;
                const float rs = __shfl_sync(kFullMask, rstd, r);
                const float4 x = *reinterpret_cast<const float4 *>(staged + r * kStride + 4 * my_chunk);
                float4 y = make_float4(fmaf(x.x * rs, scale.x, shift.x), fmaf(x.y * rs, scale.y, shift.y),
                                       fmaf(x.z * rs, scale.z, shift.z), fmaf(x.w * rs, scale.w, shift.w));
                if (relu) y = make_float4(fmaxf(y.x, 0.f), fmaxf(y.y, 0.f), fmaxf(y.z, 0.f), fmaxf(y.w, 0.f));
                y = make_float4(y.x + skip[pass].x, y.y + skip[pass].y, y.z + skip[pass].z, y.w + skip[pass].w);
                if (row0 + r < rows) __stcs(reinterpret_cast<float4 *>(out + (row0 + r) * ldo + 4 * my_chunk), y);
            }
            __syncwarp();                                                // staging rows are rewritten by the next tile
        }
    }

    // ---- teardown -------------------------------------------------------------------------------------------------------
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == tc::kMmaWarp) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((unsigned)S::kTmemColumns) : "memory");
    }
}

template <int N>
int launch_linear_tc(const float *A, long long lda, const float *W, const float *linear_bias, const float *gamma,
                     const float *beta, float *out, long long ldo, long long rows, float eps, int relu, int shortcut,
                     cudaStream_t stream) {
    using S = tc::Shape<N>;
    static int sm_count = 0;
    auto kernel = linear_norm_relu_residual_tc_kernel<N>;
    if (sm_count == 0) {                               // once per process (one process per GPU), outside any graph capture
        int device = 0, count = 0;
        ULTRA_CUDA_OK(cudaGetDevice(&device));
        ULTRA_CUDA_OK(cudaDeviceGetAttribute(&count, cudaDevAttrMultiProcessorCount, device));
        ULTRA_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kSmemBytes));
        sm_count = count;
    }
    const long long n_tiles = (rows + tc::kRows - 1) / tc::kRows;
    const unsigned grid = (unsigned)(n_tiles < sm_count ? n_tiles : sm_count);
    kernel<<<grid, tc::kThreads, S::kSmemBytes, stream>>>(A, lda, W, linear_bias, gamma, beta, out, ldo, rows, eps, relu,
                                                         shortcut);
    note_launch();
    ULTRA_CUDA_OK(cudaGetLastError());
    return ULTRA_RSPMM_OK;
}

}  // namespace

// called by ultra_layer_linear_norm_relu_residual (layer_linear.cu) after its argument checks
int layer_linear_tc(const float *A, long long lda, const float *W, const float *linear_bias, const float *gamma,
                    const float *beta, float *out, long long ldo, long long rows, int out_dim, float eps, int relu,
                    int shortcut, cudaStream_t stream) {
    if (out_dim == 64) return launch_linear_tc<64>(A, lda, W, linear_bias, gamma, beta, out, ldo, rows, eps, relu, shortcut, stream);
    return launch_linear_tc<32>(A, lda, W, linear_bias, gamma, beta, out, ldo, rows, eps, relu, shortcut, stream);
}

}  // namespace ultra
