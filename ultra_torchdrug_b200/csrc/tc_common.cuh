// tcgen05 / TMEM / TMA building blocks shared by the fused Linear (layer_linear_tc.cu) and the training GEMMs
// (layer_gemm_tc.cu): mbarrier and UMMA wrappers (inline PTX for sm_100a), shared-memory matrix descriptors, and the host
// side of the tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point - the library does not link
// libcuda).
#pragma once

#include <cuda.h>

#include <mutex>

#include "rspmm_common.cuh"

namespace ultra {
namespace tcx {

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
// K-major SWIZZLE_128B operand (rows of 128 bytes, 16-byte chunks XOR-ed with row % 8 - what TMA writes): SBO = 1024 bytes
// between 8-row groups, LBO unused, layout type 2 in bits 61-63; a k-step of 8 tf32 advances the start address by 32 bytes
__device__ __forceinline__ unsigned long long umma_desc_sw128(unsigned smem_addr) {
    return (unsigned long long)((smem_addr >> 4) & 0x3fff) | ((unsigned long long)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major tf32 operand.  The only shared-memory layout tcgen05 takes for it is SWIZZLE_128B_BASE32B (layout type 1; with
// plain SWIZZLE_128B the MMAs return zeros): atoms of 32 tf32 along M / N (128 contiguous bytes) x 4 K rows 128 bytes
// apart, the four 32-byte chunks of a row XOR-ed with row % 4 - what TMA writes for a 32-column box in the
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B mode (the K index is the box row).  LBO = byte distance between atoms along M / N,
// SBO = between 4-row K groups (512 bytes when the rows are contiguous).  Needs a_major / b_major = 1 in the instruction
// descriptor.
__device__ __forceinline__ unsigned long long umma_desc_mn_sw128_32b(unsigned smem_addr, unsigned lbo_bytes) {
    return (unsigned long long)((smem_addr >> 4) & 0x3fff) | ((unsigned long long)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((unsigned long long)(512 >> 4) << 32) | (1ull << 46) | (1ull << 61);
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned smem_dst, const CUtensorMap *map, int c0, int c1, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
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
// A operand in tensor memory (lane = row, one 32-bit column per K element), B by shared-memory descriptor
__device__ __forceinline__ void umma_tf32_ts(unsigned tmem_d, unsigned tmem_a, unsigned long long b, unsigned instr,
                                             unsigned accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(b), "r"(instr), "r"(accumulate) : "memory");
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

__device__ __forceinline__ void tmem_store16(unsigned taddr, const float (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                 "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                 "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
                 : "memory");
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline int encode_tiled(EncodeTiled *out) {
    static EncodeTiled cached = nullptr;
    static std::mutex guard;
    std::lock_guard<std::mutex> lock(guard);
    if (!cached) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult found;
        ULTRA_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &found));
        if (found != cudaDriverEntryPointSuccess || !fn) return fail_cuda(cudaErrorNotSupported);
        cached = (EncodeTiled)fn;
    }
    *out = cached;
    return ULTRA_RSPMM_OK;
}


// (rows, cols) fp32 matrix with rows `ld` floats apart as a 2-D tensor, box = 32 columns x box_rows rows, SWIZZLE_128B
inline int encode_rows_map(CUtensorMap *map, const float *base, long long rows, long long cols, long long ld, unsigned box_rows = 128,
                           CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiled encode = nullptr;
    if (int status = encode_tiled(&encode)) return status;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {32u, box_rows};
    const cuuint32_t element_strides[2] = {1, 1};
    if (encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, element_strides, CU_TENSOR_MAP_INTERLEAVE_NONE,
               swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return fail_cuda(cudaErrorInvalidValue);
    return ULTRA_RSPMM_OK;
}

}  // namespace tcx
}  // namespace ultra
