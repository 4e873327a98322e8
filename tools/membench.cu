// Micro-benchmarks that bound the embedding kernels on B200: streaming write / read / copy and random 256 B row
// gather with 128-bit vs 256-bit accesses at several load depths.
//   nvcc -arch=sm_100a -O3 -std=c++17 -o tools/membench tools/membench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

struct f8 { float v[8]; };
__device__ __forceinline__ f8 ld256(const void* p) {
  f8 r;
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
  return r;
}
__device__ __forceinline__ void st256(void* p, const f8& r) {
  asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]), "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7]) : "memory");
}
__device__ __forceinline__ float4 ld128(const void* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st128(void* p, const float4& v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void write128(float4* out, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
    st128(out + i, make_float4(1.f, 2.f, 3.f, 4.f));
}
__global__ void write256(f8* out, size_t n8) {
  f8 v; for (int k = 0; k < 8; ++k) v.v[k] = (float)k;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) st256(out + i, v);
}
__global__ void read128(const float4* in, size_t n4, float* sink) {
  float a = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) { float4 v = ld128(in + i); a += v.x + v.y + v.z + v.w; }
  if (a == 123.456f) *sink = a;
}
__global__ void copy128(const float4* in, float4* out, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) st128(out + i, ld128(in + i));
}

// gather rows of 256 B: LANES lanes per row (16 -> 128-bit, 8 -> 256-bit), U rows in flight per group, optional copy-out
template <int LANES, int U, bool STORE>
__global__ void gather_rows(const char* table, const uint32_t* idx, size_t n_rows, char* out, float* sink) {
  const int lane = threadIdx.x % LANES;
  const size_t grp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  const size_t ngrp = ((size_t)gridDim.x * blockDim.x) / LANES;
  float acc = 0.f;
  for (size_t r0 = grp * U; r0 < n_rows; r0 += ngrp * U) {
    if constexpr (LANES == 16) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) if (r0 + u < n_rows) v[u] = ld128(table + (size_t)idx[r0 + u] * 256 + lane * 16);
#pragma unroll
      for (int u = 0; u < U; ++u) if (r0 + u < n_rows) { if (STORE) st128(out + (r0 + u) * 256 + lane * 16, v[u]); else acc += v[u].x + v[u].w; }
    } else {
      f8 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) if (r0 + u < n_rows) v[u] = ld256(table + (size_t)idx[r0 + u] * 256 + lane * 32);
#pragma unroll
      for (int u = 0; u < U; ++u) if (r0 + u < n_rows) { if (STORE) st256(out + (r0 + u) * 256 + lane * 32, v[u]); else acc += v[u].v[0] + v[u].v[7]; }
    }
  }
  if (acc == 123.456f) *sink = acc;
}

template <class F> float timeit(F f, int iters = 10) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int i = 0; i < iters; ++i) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  return best;
}

int main() {
  const size_t bytes = (size_t)2 << 30;  // 2 GiB buffers
  char *A, *B; float* sink; CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(A, 1, bytes)); CK(cudaMemset(B, 0, bytes));
  const int grid = 148 * 8, blk = 256;
  float ms;
  ms = timeit([&] { write128<<<grid, blk>>>((float4*)B, bytes / 16); }); printf("write128  %7.1f GB/s\n", bytes / ms / 1e6);
  ms = timeit([&] { write256<<<grid, blk>>>((f8*)B, bytes / 32); });     printf("write256  %7.1f GB/s\n", bytes / ms / 1e6);
  ms = timeit([&] { cudaMemsetAsync(B, 0, bytes); });                    printf("memset    %7.1f GB/s\n", bytes / ms / 1e6);
  ms = timeit([&] { read128<<<grid, blk>>>((float4*)A, bytes / 16, sink); }); printf("read128   %7.1f GB/s\n", bytes / ms / 1e6);
  ms = timeit([&] { copy128<<<grid, blk>>>((float4*)A, (float4*)B, bytes / 16); }); printf("copy128   %7.1f GB/s (r+w)\n", 2.0 * bytes / ms / 1e6);
  // random rows from a 1.3 GB region (like the concat-gradient buffers), 2.8M rows (one backward)
  const size_t region_rows = (size_t)1300 * 1000 * 1000 / 256, n_rows = 2800000;
  std::vector<uint32_t> h(n_rows); uint64_t s = 88172645463325252ull;
  for (auto& x : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x = (uint32_t)(s % region_rows); }
  uint32_t* idx; CK(cudaMalloc(&idx, n_rows * 4)); CK(cudaMemcpy(idx, h.data(), n_rows * 4, cudaMemcpyHostToDevice));
  const double gb = n_rows * 256.0;
#define G(L, U, S, g) ms = timeit([&] { gather_rows<L, U, S><<<g, blk>>>(A, idx, n_rows, B, sink); }); \
  printf("gather lanes=%2d U=%d store=%d grid=%5d  %7.1f GB/s (%s)\n", L, U, S, g, (S ? 2 : 1) * gb / ms / 1e6, S ? "r+w" : "read");
  for (int g : {148 * 4, 148 * 8, 148 * 16}) {
    G(16, 2, false, g) G(16, 4, false, g) G(16, 8, false, g) G(8, 2, false, g) G(8, 4, false, g) G(8, 8, false, g)
  }
  G(16, 4, true, 148 * 8) G(8, 4, true, 148 * 8) G(16, 8, true, 148 * 8) G(8, 8, true, 148 * 8)
  CK(cudaDeviceSynchronize());
  return 0;
}
