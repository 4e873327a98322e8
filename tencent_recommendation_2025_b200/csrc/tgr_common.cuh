// Shared device/host helpers for the sm_100a sparse-feature embedding kernels.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "tgr_embed.h"

namespace tgr {

// ---- error plumbing (thread-local, no global mutable state shared across threads) ----------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define TGR_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      ::tgr::set_error(__VA_ARGS__);  \
      return -1;                      \
    }                                 \
  } while (0)

// RAII CUDA-event pair around one C-ABI entry's launches; a no-op unless tgr_timing_enable(1) was called.
class TimedScope {
 public:
  TimedScope(const char* name, void* stream);
  ~TimedScope();
 private:
  const char* name_;
  void* stream_;
  void* a_;
};

// Every kernel launch of the library goes through TGR_K(kernel)<<<...>>>(...): a comma expression bumps a process-wide
// counter ahead of the launch, so tgr_launch_count() is a COUNT of launches, not an estimate (bench.py gpu_launches).
void count_launch();
// variants whose entry count lives in device memory (n = capacity, *n_dev <= n the count when the kernels run): the
// launch sequence is then a function of the call SHAPES only and can be captured in a CUDA graph (tgr_fact_group_t.n_is_capacity)
int sort_pairs_dn(const uint32_t* keys_in, const uint32_t* srcs_in, uint32_t* keys_out, uint32_t* srcs_out, int64_t n,
                  int key_bits, void* workspace, size_t workspace_bytes, const int32_t* n_dev, void* stream);
int dedup_dn(const uint32_t* keys_sorted, int64_t n, uint32_t* uniq, int32_t* seg_off, int32_t* seg_of_entry,
             int32_t* n_unique_dev, void* workspace, size_t workspace_bytes, const int32_t* n_dev, void* stream);
int dedup_remap_dn(const uint32_t* keys_sorted, const uint32_t* srcs_sorted, int64_t n, uint32_t* uniq, int32_t* seg_off,
                   int32_t* seg_of_entry, int32_t* n_unique_dev, void* workspace, size_t workspace_bytes, const tgr_call_t* calls,
                   int n_calls, int32_t* const* ids_out, const int32_t* n_dev, void* stream);
int remap_scatter_dn(const uint32_t* srcs_sorted, const int32_t* seg_of_entry, int64_t n, const int32_t* perm,
                     const tgr_call_t* calls, int n_calls, int32_t* const* ids_out, const int32_t* n_dev, void* stream);
int bwd_reduce_dn(const tgr_table_t* tables, int n_tables, int H, const tgr_call_t* calls, int n_calls,
                  const uint32_t* keys_sorted, const uint32_t* srcs_sorted, int64_t n, int mode, const int32_t* seg_of_entry,
                  float* grads_out, const tgr_adam_t* adam, void* workspace, size_t workspace_bytes, const int32_t* n_dev,
                  void* stream);
#define TGR_K(...) ::tgr::count_launch(), __VA_ARGS__

constexpr int kNumSMs = 148;  // B200

// ---- cache-hinted 128-bit accesses --------------------------------------------------------
// Table rows: default caching (hot Zipf rows and small tables live in the 126 MB L2).
__device__ __forceinline__ float4 ld_row(const float4* p) { return __ldg(p); }

// Streaming data touched exactly once (concat output, concat gradients): keep it out of L1 and
// (st.global.cs marks the stores evict-first) so it does not displace table rows.
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ld_stream_u2(const uint2* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream_u2(uint2* p, const uint2& v) {
  asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

__device__ __forceinline__ uint2 pack_bf16x4(const float4& v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&lo);
  r.y = *reinterpret_cast<uint32_t*>(&hi);
  return r;
}
__device__ __forceinline__ float4 unpack_bf16x4(const uint2& u) {
  float4 r;
  r.x = __uint_as_float(u.x << 16);
  r.y = __uint_as_float(u.x & 0xFFFF0000u);
  r.z = __uint_as_float(u.y << 16);
  r.w = __uint_as_float(u.y & 0xFFFF0000u);
  return r;
}

__device__ __forceinline__ float4 f4_add(const float4& a, const float4& b) {
  // plain IEEE adds, no contraction possible
  return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}

}  // namespace tgr
