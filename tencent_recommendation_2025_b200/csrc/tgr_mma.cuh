// Error-compensated TF32 tensor-core helpers (3xTF32) for the small fp32 contractions of the factored path.
//
// The parity bar is 1e-5 of tensor scale against the reference's fp32 Linear (model/BaseLine/model.py:297,303,306);
// single-pass TF32 (10 explicit mantissa bits) gives ~5e-4. Every fp32 operand x is therefore split as
//     x = hi + lo,   hi = tf32(x),   lo = tf32(x - hi)          (hi + lo carries ~21 mantissa bits)
// and a product a.b is accumulated as  a_lo.b_hi + a_hi.b_lo + a_hi.b_hi  in the tensor core's fp32 accumulator (the
// dropped a_lo.b_lo term is 2^-22 relative). Three MMAs per tile instead of one: still far above the FFMA rate.
//
// mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 fragment layout (g = lane >> 2, t = lane & 3):
//   A (16 x 8, row):  a0 = A[g][t]      a1 = A[g+8][t]    a2 = A[g][t+4]    a3 = A[g+8][t+4]
//   B ( 8 x 8, col):  b0 = B[t][g]      b1 = B[t+4][g]
//   C (16 x 8)     :  c0 = C[g][2t]     c1 = C[g][2t+1]   c2 = C[g+8][2t]   c3 = C[g+8][2t+1]
#pragma once
#include "tgr_common.cuh"

namespace tgr {

// Round-to-nearest (ties away from zero) to 10 explicit mantissa bits: what cvt.rna.tf32.f32 returns for every finite
// value, as two integer instructions. On sm_100a ptxas expands the cvt into an ~8-instruction FSETP/IMAD/LOP3/SEL
// sequence; with two conversions per split and a split per fragment element that was 75 % of fact_dz_kernel's issue
// slots (profiles/README.md r2, ncu source page). A mantissa carry runs into the exponent exactly as rounding up to the
// next binade must; inf stays inf (the mask clears the added bits), NaN stays NaN.
__device__ __forceinline__ uint32_t to_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = to_tf32(x);
  lo = to_tf32(__fsub_rn(x, __uint_as_float(hi)));
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// d += a . b with a = a_hi + a_lo, b = b_hi + b_lo (small terms first)
__device__ __forceinline__ void mma_3xtf32(float (&d)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4],
                                           const uint32_t (&bhi)[2], const uint32_t (&blo)[2]) {
  mma_tf32(d, alo, bhi);
  mma_tf32(d, ahi, blo);
  mma_tf32(d, ahi, bhi);
}

}  // namespace tgr
