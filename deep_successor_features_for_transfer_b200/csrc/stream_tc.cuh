// tcgen05 kind::tf32 building blocks of the fp32-precision tensor-core path (mlp_stream_tc.cu, mlp_wgrad_tf32.cu).
//
// Precision modes on 32-bit operands (measured on B200, scripts/tf32_probe.cu):
//   TF32   one pass.  The tensor core TRUNCATES the low 13 mantissa bits of an fp32 operand (error vs truncated inputs 1.6e-7,
//          vs round-to-nearest inputs 8e-4), so operands are stored already rounded to nearest (cvt.rna): ~2e-4 per product.
//   TF32X3 three passes  a_hi.b_hi + a_hi.b_lo + a_lo.b_hi  with hi = rna_tf32(x), lo = rna_tf32(x - hi) (so that the core's truncation
//          of lo is a no-op and the split error is an unbiased 2^-22 instead of a one-sided 2^-21), fp32 accumulation
//          in TMEM: 4e-7 relative error of a K = 32 product against fp64 -- the reference's fp32 arithmetic on the tensor cores.
//   One M128 N256 K8 instruction executes in 128 cycles (half the bf16 rate): a 128 x 256 x 256 tile-layer is 4096 cycles per pass.
// Operand layouts: K-major SWIZZLE_128B exactly as for bf16 (a 128-byte span holds 32 fp32 = four K = 8 steps, 32 bytes apart);
// MN-major fp32 has ONE legal swizzled layout, SWIZZLE_128B_BASE32B (descriptor layout type 1): 32-byte chunks of a 128-byte row
// XORed with (row mod 4), k atoms of 4 rows 512 B apart -- what TMA produces with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
#pragma once
#include "tc_common.cuh"

namespace sfgpi {
namespace tc {

constexpr int kPrecTf32 = 2, kPrecTf32x3 = 3;           // SFGPI_PREC_* of include/sfgpi.h
constexpr int kK32 = 32;                                 // fp32 elements per 128-byte swizzle span = k extent of one ring stage

__device__ __forceinline__ uint32_t idesc_tf32(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
    // c_format F32 (1) @4, a/b format TF32 (2) @7/@10, A / B MN-major @15/@16, N>>3 @17, M>>4 @24
    return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32_e(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                            uint32_t leader) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
        : "memory");
}
// MN-major fp32 operand tile, SWIZZLE_128B_BASE32B: 32-wide M/N blocks LBO bytes apart, 4-row k atoms 512 B apart; a K = 8 step
// (two atoms) advances the start address by 1024 B.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128_32b(uint32_t smem_addr, uint32_t lbo_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)1 << 61);
}
// hi = x rounded to nearest tf32 (ties away), as a float with the low 13 mantissa bits clear
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// byte offset of the 16-byte chunk holding fp32 columns [c, c+4) (c < 32, multiple of 4) of row r inside a [rows][32] K-major SW128 tile
__device__ __forceinline__ uint32_t a32_chunk_off(int r, int c) { return (uint32_t)(r * 128 + (((c >> 2) ^ (r & 7)) << 4)); }

// generic fp32 tensor map (rank 2 or 3), box given, swizzle SWIZZLE_128B (K-major use) or SWIZZLE_128B_ATOM_32B (MN-major use)
static inline int make_tmap_f32(CUtensorMap *tm, const void *base, int rank, const uint64_t *dims, const uint32_t *box, bool atom32) {
    EncodeTiledFn encode = get_encode_fn();
    if (!encode) { set_error("cuTensorMapEncodeTiled entry point not found"); return SFGPI_E_CUDA; }
    cuuint64_t gdim[3], gstride[2];
    cuuint32_t bx[3], estride[3] = {1, 1, 1};
    uint64_t pitch = 4;
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        pitch *= dims[i];
        if (i + 1 < rank) gstride[i] = pitch;
    }
    CUresult cr = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstride, bx, estride,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(f32) failed (%d)", (int)cr); return SFGPI_E_CUDA; }
    return SFGPI_OK;
}

}  // namespace tc
}  // namespace sfgpi
