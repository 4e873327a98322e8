// tc_probe — standalone check of the tcgen05 building blocks the row kernels use (dev tool, runs on the GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -cudart shared -o tools/tc_probe tools/tc_probe.cu
//   tools/tc_probe <variant>
// One CTA, 128 threads: copy host-built shared-memory images of A and B (no-swizzle canonical core-matrix layouts),
// issue K/8 tcgen05.mma kind::tf32 (M = 128, N = 64) from one thread, commit to an mbarrier, read the accumulator back
// with tcgen05.ld.32x32b and compare with the host result. Values are multiples of 1/8 in [-2, 2]: exact in tf32, so
// any mismatch is a layout / descriptor error, not rounding. Variants sweep the LBO / SBO reading of the descriptor.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

struct Params {
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;   // bytes
  uint32_t a_step, b_step;               // descriptor start-address advance per MMA (bytes)
  uint32_t a_major, b_major;             // 0 = K-major, 1 = MN-major
  uint32_t n_mma;                        // K / 8
  uint32_t a_bytes, b_bytes;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  return d;                 // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

__global__ void __launch_bounds__(128) probe_kernel(const uint8_t* __restrict__ a_img, const uint8_t* __restrict__ b_img,
                                                    float* __restrict__ D, Params p, int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base;
  uint8_t* As = smem;
  uint8_t* Bs = smem + ((p.a_bytes + 1023) / 1024) * 1024;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (uint32_t i = tid * 16; i < p.a_bytes; i += 128 * 16) *(uint4*)(As + i) = *(const uint4*)(a_img + i);
  for (uint32_t i = tid * 16; i < p.b_bytes; i += 128 * 16) *(uint4*)(Bs + i) = *(const uint4*)(b_img + i);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tacc = tmem_base;
  if (tid == 0) {
    // instruction descriptor: c = f32 (1 << 4), a = b = tf32 (2 << 7, 2 << 10), majors, N >> 3 at 17, M >> 4 at 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (p.a_major << 15) | (p.b_major << 16) | ((64u >> 3) << 17) |
                           ((128u >> 4) << 24);
    for (uint32_t k = 0; k < p.n_mma; ++k) {
      const uint64_t da = make_desc(smem_u32(As) + k * p.a_step, p.a_lbo, p.a_sbo);
      const uint64_t db = make_desc(smem_u32(Bs) + k * p.b_step, p.b_lbo, p.b_sbo);
      const uint32_t acc = k > 0 ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tacc),
          "l"(da), "l"(db), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  // bounded wait on phase 0
  uint32_t done = 0;
  for (int it = 0; it < (1 << 22) && !done; ++it) {
    asm volatile(
        "{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
        : "=r"(done)
        : "r"(smem_u32(&mbar)), "r"(0u)
        : "memory");
  }
  if (!done) {
    if (tid == 0) *status = 2;
  } else {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[64];
    const uint32_t taddr = tacc + ((uint32_t)(warp * 32) << 16);
#define L8(o)                                                                                                     \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                           \
               : "=r"(r[o]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]),      \
                 "=r"(r[o + 6]), "=r"(r[o + 7])                                                                   \
               : "r"(taddr + o))
    L8(0); L8(8); L8(16); L8(24); L8(32); L8(40); L8(48); L8(56);
#undef L8
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 64; ++j) D[tid * 64 + j] = __uint_as_float(r[j]);
    if (tid == 0) *status = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tacc) : "memory");
}

static float val(int i, int j, int salt) {
  uint32_t h = (uint32_t)(i * 1315423911u) ^ (uint32_t)(j * 2654435761u) ^ (uint32_t)(salt * 97531u);
  h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
  return (float)((int)(h % 33) - 16) / 8.0f;
}

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  // variant 0/1: K-major A [128 x 64] and B [64 x 64]; core (r/8, k/4) at (r/8)*2048 + (k/4)*128 + (r%8)*16
  //              0: LBO = 128 (K dir), SBO = 2048 (MN dir), +256 B per MMA;   1: LBO / SBO swapped
  // variant 2/3: MN-major operands from row-major G [u=128][h=64], R [u=128][k=64]: chunk (u, h/4) at
  //              (u/8)*2048 + (h/4)*128 + (u%8)*16; D[h][k] = sum_u G[u][h] R[u][k], rows h >= 64 are don't-care
  //              2: LBO = 2048 (K dir), SBO = 128 (MN dir), +2048 B per MMA;  3: swapped
  Params p{};
  std::vector<uint8_t> a_img, b_img;
  std::vector<float> ref(128 * 64, 0.f);
  int rows_checked = 128;
  if (variant < 2) {
    const int K = 64;
    a_img.assign(128 * K * 4, 0);
    b_img.assign(64 * K * 4, 0);
    for (int r = 0; r < 128; ++r)
      for (int k = 0; k < K; ++k) {
        const float v = val(r, k, 1);
        memcpy(&a_img[(r / 8) * 2048 + (k / 4) * 128 + (r % 8) * 16 + (k % 4) * 4], &v, 4);
      }
    for (int n = 0; n < 64; ++n)
      for (int k = 0; k < K; ++k) {
        const float v = val(n, k, 2);
        memcpy(&b_img[(n / 8) * 2048 + (k / 4) * 128 + (n % 8) * 16 + (k % 4) * 4], &v, 4);
      }
    for (int r = 0; r < 128; ++r)
      for (int n = 0; n < 64; ++n) {
        float s = 0.f;
        for (int k = 0; k < K; ++k) s += val(r, k, 1) * val(n, k, 2);
        ref[r * 64 + n] = s;
      }
    p.a_lbo = p.b_lbo = variant == 0 ? 128 : 2048;
    p.a_sbo = p.b_sbo = variant == 0 ? 2048 : 128;
    p.a_step = p.b_step = 256;
    p.a_major = p.b_major = 0;
    p.n_mma = K / 8;
  } else {
    const int Ku = 128;
    a_img.assign(Ku * 64 * 4 + 4096, 0);   // + slack: the M = 128 read runs 2 KB past the 64 real MN rows
    b_img.assign(Ku * 64 * 4, 0);
    for (int u = 0; u < Ku; ++u)
      for (int h = 0; h < 64; ++h) {
        const float g = val(u, h, 3), r = val(u, h, 4);
        memcpy(&a_img[(u / 8) * 2048 + (h / 4) * 128 + (u % 8) * 16 + (h % 4) * 4], &g, 4);
        memcpy(&b_img[(u / 8) * 2048 + (h / 4) * 128 + (u % 8) * 16 + (h % 4) * 4], &r, 4);
      }
    for (int h = 0; h < 64; ++h)
      for (int k = 0; k < 64; ++k) {
        float s = 0.f;
        for (int u = 0; u < Ku; ++u) s += val(u, h, 3) * val(u, k, 4);
        ref[h * 64 + k] = s;
      }
    rows_checked = 64;
    p.a_lbo = p.b_lbo = variant == 2 ? 2048 : 128;
    p.a_sbo = p.b_sbo = variant == 2 ? 128 : 2048;
    p.a_step = p.b_step = 2048;
    p.a_major = p.b_major = 1;
    p.n_mma = Ku / 8;
  }
  p.a_bytes = (uint32_t)a_img.size();
  p.b_bytes = (uint32_t)b_img.size();
  uint8_t *da, *db;
  float* dD;
  int* dst;
  cudaMalloc(&da, a_img.size());
  cudaMalloc(&db, b_img.size());
  cudaMalloc(&dD, 128 * 64 * 4);
  cudaMalloc(&dst, 4);
  cudaMemcpy(da, a_img.data(), a_img.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(db, b_img.data(), b_img.size(), cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, 128 * 64 * 4);
  cudaMemset(dst, 0, 4);
  const size_t smem = ((a_img.size() + 1023) / 1024) * 1024 + b_img.size() + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<<<1, 128, smem>>>(da, db, dD, p, dst);
  cudaError_t e = cudaDeviceSynchronize();
  int status = 0;
  std::vector<float> out(128 * 64);
  if (e == cudaSuccess) {
    cudaMemcpy(&status, dst, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
  }
  double maxerr = 0;
  int bad = 0;
  for (int i = 0; i < rows_checked * 64; ++i) {
    const double d = fabs((double)out[i] - (double)ref[i]);
    if (d > maxerr) maxerr = d;
    if (d > 1e-3) ++bad;
  }
  printf("variant %d: cuda=%s status=%d (1 = ok, 2 = mbarrier timeout) max_err=%g mismatches=%d / %d  D[0][0..3]=%g %g %g %g ref=%g %g %g %g\n",
         variant, cudaGetErrorString(e), status, maxerr, bad, rows_checked * 64, out[0], out[1], out[2], out[3], ref[0],
         ref[1], ref[2], ref[3]);
  return 0;
}
