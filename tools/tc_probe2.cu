// tc_probe2 — standalone check of TMA (cp.async.bulk.tensor, 128-byte swizzle) feeding tcgen05.mma kind::f16 (bf16):
//   nvcc -gencode arch=compute_100a,code=sm_100a -cudart shared -o tools/tc_probe2 tools/tc_probe2.cu
//   tools/tc_probe2 <variant>      variant 0: LBO field 1 (CUTLASS), 1: LBO field 0
// D[128 x 64] = X[128 x K] . W[64 x K]^T, K = 256, bf16 inputs that are multiples of 1/8 (exact), fp32 accumulate.
// One CTA: thread 0 drives a 2-slot TMA ring (X box 64 x 128, W box 64 x 64) and issues 4 MMAs (K = 16 each, +32 B on
// the descriptor start address inside the 128-byte swizzle atom) per 64-wide K block; 4 warps read TMEM back.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_field, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)(lbo_field & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ bool wait_bounded(uint64_t* bar, uint32_t parity) {
  for (int it = 0; it < (1 << 22); ++it)
    if (try_wait(bar, parity)) return true;
  return false;
}

constexpr int kK = 256, kBK = 64, kStages = 2;
constexpr int kABytes = 128 * kBK * 2, kBBytes = 64 * kBK * 2;

__global__ void __launch_bounds__(128) probe2_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                                                     float* __restrict__ D, uint32_t lbo_field, int* __restrict__ status) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* As = base;                          // [stages][16 KB]
  uint8_t* Bs = base + kStages * kABytes;      // [stages][8 KB]
  __shared__ __align__(8) uint64_t full[kStages], empty[kStages], accbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[s])));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&accbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tacc = tmem_base;
  bool ok = true;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);   // bf16 x bf16 -> f32
    auto load = [&](int kb) {
      const int s = kb % kStages;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(kABytes + kBBytes) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                       smem_u32(As + s * kABytes)),
                   "l"(&tm_x), "r"(smem_u32(&full[s])), "r"(kb * kBK), "r"(0)
                   : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                       smem_u32(Bs + s * kBBytes)),
                   "l"(&tm_w), "r"(smem_u32(&full[s])), "r"(kb * kBK), "r"(0)
                   : "memory");
    };
    const int nkb = kK / kBK;
    for (int kb = 0; kb < kStages && kb < nkb; ++kb) load(kb);
    for (int kb = 0; kb < nkb && ok; ++kb) {
      const int s = kb % kStages;
      const uint32_t par = (kb / kStages) & 1;
      ok = wait_bounded(&full[s], par);
      if (!ok) break;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int k = 0; k < kBK / 16; ++k) {
        const uint64_t da = make_desc(smem_u32(As + s * kABytes) + k * 32, lbo_field, 1024);
        const uint64_t db = make_desc(smem_u32(Bs + s * kBBytes) + k * 32, lbo_field, 1024);
        const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tacc),
            "l"(da), "l"(db), "r"(idesc), "r"(acc)
            : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
      if (kb + kStages < nkb) {
        ok = wait_bounded(&empty[s], par);
        if (!ok) break;
        load(kb + kStages);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&accbar)) : "memory");
    if (!ok) *status = 3;
  }
  const bool done = wait_bounded(&accbar, 0);
  if (!done) {
    if (tid == 0 && *status == 0) *status = 2;
  } else {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tacc + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < 64; c0 += 8) {
      uint32_t r[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(taddr + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) D[tid * 64 + c0 + j] = __uint_as_float(r[j]);
    }
    if (tid == 0 && *status == 0) *status = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tacc) : "memory");
}

// ---- variants 2 / 3: both operands MN-major (the split-K dW = dZ^T x GEMM of the mm backward) -------------------------
// D[h][j] = sum_t Z[t][h] X[t][j], Z bf16 [T x 64] and X bf16 [T x 128] row-major, T = 256. TMA boxes {64 inner, 64 t}:
// a box is 64 k-rows of 128 bytes = 8 MN-major SWIZZLE_128B atoms (8 k-rows x 64 MN elements). A: one box (M = 128 reads a
// second MN group at +LBO_A; LBO_A = 0 repeats the first, rows 64..127 of D are don't-care); B: two boxes, MN groups 8 KB apart.
// One K = 16 MMA spans two 8-row k-groups: +2048 bytes per MMA. variant 2: LBO = MN-group stride, SBO = k-group stride
// (1024); variant 3: swapped.
constexpr int kT2 = 256, kStageA2 = 64 * 64 * 2, kStageB2 = 2 * 64 * 64 * 2;
__global__ void __launch_bounds__(128) probe_mn_kernel(const __grid_constant__ CUtensorMap tm_z, const __grid_constant__ CUtensorMap tm_x,
                                                       float* __restrict__ D, uint32_t swap, int* __restrict__ status) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* As = base;                              // [stages][8 KB]
  uint8_t* Bs = base + kStages * kStageA2;         // [stages][16 KB]
  __shared__ __align__(8) uint64_t full[kStages], empty[kStages], accbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[s])));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&accbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tacc = tmem_base;
  bool ok = true;
  if (tid == 0) {
    // bf16 x bf16 -> f32, A and B MN-major (bits 15, 16), N = 128, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    auto mk = [&](uint32_t addr, uint32_t mn_stride, uint32_t k_stride) {
      const uint32_t lbo = swap ? k_stride : mn_stride, sbo = swap ? mn_stride : k_stride;
      uint64_t d = 0;
      d |= (uint64_t)((addr >> 4) & 0x3FFF);
      d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
      d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
      d |= (uint64_t)1 << 46;
      d |= (uint64_t)2 << 61;
      return d;
    };
    auto load = [&](int kb) {
      const int s = kb % kStages;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(kStageA2 + kStageB2) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                       smem_u32(As + s * kStageA2)), "l"(&tm_z), "r"(smem_u32(&full[s])), "r"(0), "r"(kb * 64) : "memory");
      for (int g = 0; g < 2; ++g)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                         smem_u32(Bs + s * kStageB2 + g * 8192)), "l"(&tm_x), "r"(smem_u32(&full[s])), "r"(g * 64), "r"(kb * 64) : "memory");
    };
    const int nkb = kT2 / 64;
    for (int kb = 0; kb < kStages && kb < nkb; ++kb) load(kb);
    for (int kb = 0; kb < nkb && ok; ++kb) {
      const int s = kb % kStages;
      const uint32_t par = (kb / kStages) & 1;
      ok = wait_bounded(&full[s], par);
      if (!ok) break;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int k = 0; k < 4; ++k) {
        const uint64_t da = mk(smem_u32(As + s * kStageA2) + k * 2048, 0, 1024);
        const uint64_t db = mk(smem_u32(Bs + s * kStageB2) + k * 2048, 8192, 1024);
        const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tacc), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
      if (kb + kStages < nkb) {
        ok = wait_bounded(&empty[s], par);
        if (!ok) break;
        load(kb + kStages);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&accbar)) : "memory");
    if (!ok) *status = 3;
  }
  const bool done = wait_bounded(&accbar, 0);
  if (!done) {
    if (tid == 0 && *status == 0) *status = 2;
  } else {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tacc + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < 128; c0 += 8) {
      uint32_t r[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(taddr + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) D[tid * 128 + c0 + j] = __uint_as_float(r[j]);
    }
    if (tid == 0 && *status == 0) *status = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tacc) : "memory");
}

static float val(int i, int j, int salt) {
  uint32_t h = (uint32_t)(i * 1315423911u) ^ (uint32_t)(j * 2654435761u) ^ (uint32_t)(salt * 97531u);
  h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
  return (float)((int)(h % 33) - 16) / 8.0f;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);



static int run_mn(int variant) {
  const int T = kT2;
  std::vector<__nv_bfloat16> hz((size_t)T * 64), hx((size_t)T * 128);
  for (int t = 0; t < T; ++t) {
    for (int h = 0; h < 64; ++h) hz[(size_t)t * 64 + h] = __float2bfloat16(val(t, h, 7));
    for (int j = 0; j < 128; ++j) hx[(size_t)t * 128 + j] = __float2bfloat16(val(t, j, 8));
  }
  std::vector<float> ref((size_t)64 * 128);
  for (int h = 0; h < 64; ++h)
    for (int j = 0; j < 128; ++j) {
      float s = 0.f;
      for (int t = 0; t < T; ++t) s += val(t, h, 7) * val(t, j, 8);
      ref[(size_t)h * 128 + j] = s;
    }
  __nv_bfloat16 *dz, *dx;
  float* dD;
  int* dst;
  cudaMalloc(&dz, hz.size() * 2);
  cudaMalloc(&dx, hx.size() * 2);
  cudaMalloc(&dD, 128 * 128 * 4);
  cudaMalloc(&dst, 4);
  cudaMemcpy(dz, hz.data(), hz.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, 128 * 128 * 4);
  cudaMemset(dst, 0, 4);
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode) return 1;
  CUtensorMap tmz, tmx;
  cuuint32_t box[2] = {64, 64}, es[2] = {1, 1};
  {
    cuuint64_t dims[2] = {64, (cuuint64_t)T}, strides[1] = {64 * 2};
    if (encode(&tmz, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dz, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 1;
  }
  {
    cuuint64_t dims[2] = {128, (cuuint64_t)T}, strides[1] = {128 * 2};
    if (encode(&tmx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dx, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 1;
  }
  const size_t smem = kStages * (kStageA2 + kStageB2) + 1024;
  cudaFuncSetAttribute(probe_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_mn_kernel<<<1, 128, smem>>>(tmz, tmx, dD, variant == 3 ? 1u : 0u, dst);
  cudaError_t e = cudaDeviceSynchronize();
  int status = 0;
  std::vector<float> out(128 * 128);
  if (e == cudaSuccess) {
    cudaMemcpy(&status, dst, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
  }
  double maxerr = 0;
  int bad = 0;
  for (size_t i = 0; i < ref.size(); ++i) {
    const double d = fabs((double)out[i] - (double)ref[i]);
    if (d > maxerr) maxerr = d;
    if (d > 1e-3) ++bad;
  }
  printf("probe2 variant %d (MN-major): cuda=%s status=%d max_err=%g mismatches=%d / %zu  D[0][0..3]=%g %g %g %g ref=%g %g %g %g  D[64][0]=%g\n",
         variant, cudaGetErrorString(e), status, maxerr, bad, ref.size(), out[0], out[1], out[2], out[3], ref[0], ref[1], ref[2], ref[3], out[64 * 128]);
  return 0;
}

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  if (variant >= 2) return run_mn(variant);
  const int T = 128, N = 64, K = kK;
  std::vector<__nv_bfloat16> hx((size_t)T * K), hw((size_t)N * K);
  for (int t = 0; t < T; ++t)
    for (int k = 0; k < K; ++k) hx[(size_t)t * K + k] = __float2bfloat16(val(t, k, 5));
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) hw[(size_t)n * K + k] = __float2bfloat16(val(n, k, 6));
  std::vector<float> ref((size_t)T * N);
  for (int t = 0; t < T; ++t)
    for (int n = 0; n < N; ++n) {
      float s = 0.f;
      for (int k = 0; k < K; ++k) s += val(t, k, 5) * val(n, k, 6);
      ref[(size_t)t * N + n] = s;
    }
  __nv_bfloat16 *dx, *dw;
  float* dD;
  int* dst;
  cudaMalloc(&dx, hx.size() * 2);
  cudaMalloc(&dw, hw.size() * 2);
  cudaMalloc(&dD, ref.size() * 4);
  cudaMalloc(&dst, 4);
  cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, ref.size() * 4);
  cudaMemset(dst, 0, 4);
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres);
  if (ge != cudaSuccess || encode == nullptr) { printf("no cuTensorMapEncodeTiled: %s\n", cudaGetErrorString(ge)); return 1; }
  CUtensorMap tmx, tmw;
  {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)T};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&tmx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dx, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode x failed %d\n", (int)r); return 1; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&tmw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dw, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode w failed %d\n", (int)r); return 1; }
  }
  const size_t smem = kStages * (kABytes + kBBytes) + 1024;
  cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe2_kernel<<<1, 128, smem>>>(tmx, tmw, dD, variant == 0 ? 1u : 0u, dst);
  cudaError_t e = cudaDeviceSynchronize();
  int status = 0;
  std::vector<float> out(ref.size());
  if (e == cudaSuccess) {
    cudaMemcpy(&status, dst, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
  }
  double maxerr = 0;
  int bad = 0;
  for (size_t i = 0; i < ref.size(); ++i) {
    const double d = fabs((double)out[i] - (double)ref[i]);
    if (d > maxerr) maxerr = d;
    if (d > 1e-3) ++bad;
  }
  printf("probe2 variant %d: cuda=%s status=%d (1 ok, 2 acc timeout, 3 ring timeout) max_err=%g mismatches=%d / %zu  D[0][0..3]=%g %g %g %g ref=%g %g %g %g\n",
         variant, cudaGetErrorString(e), status, maxerr, bad, ref.size(), out[0], out[1], out[2], out[3], ref[0], ref[1], ref[2], ref[3]);
  return 0;
}
