// Row-sharded multi-GPU routing (no reference counterpart — the reference is single-GPU).
//
// Tables are sharded by global key: owner = key mod W, local_row = key div W (round-robin spreads the
// Zipf-hot rows; NVSwitch is uniform so there is no topology term). Given this rank's sorted unique keys,
// produce the all-to-all send layout: local rows grouped by owner (ascending key inside each bucket),
// the per-owner counts, and each key's position in that layout (to expand the returned rows).
// Stable counting partition: count -> scan -> emit, no atomics, bit-reproducible.
#include "tgr_common.cuh"

namespace tgr {

constexpr int kRouteBlock = 1024;
constexpr int kMaxW = 64;

__global__ void __launch_bounds__(kRouteBlock) route_count_kernel(const uint32_t* __restrict__ uniq,
                                                                  const int32_t* __restrict__ n_dev, int W, int nb,
                                                                  int32_t* __restrict__ cnt /*[W][nb]*/) {
  const int n = *n_dev;
  const int i = blockIdx.x * kRouteBlock + threadIdx.x;
  const int owner = i < n ? (int)(__ldg(uniq + i) % (uint32_t)W) : -1;
  for (int w = 0; w < W; ++w) {
    const int c = __syncthreads_count(owner == w);
    if (threadIdx.x == 0) cnt[w * nb + blockIdx.x] = c;
  }
}

// exclusive scan over the w-major [W*nb] array (single CTA), then per-owner totals
__global__ void __launch_bounds__(kRouteBlock) route_scan_kernel(int32_t* __restrict__ cnt, int W, int nb,
                                                                 int32_t* __restrict__ counts_out) {
  __shared__ int32_t warp_sum[32];
  __shared__ int32_t carry_s;
  __shared__ int32_t base_s[kMaxW + 1];
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int total = W * nb;
  for (int base = 0; base < total; base += kRouteBlock) {
    const int i = base + threadIdx.x;
    const int v = i < total ? cnt[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sum[wid] = x;
    __syncthreads();
    if (wid == 0) {
      int s = warp_sum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += y;
      }
      warp_sum[lane] = s;
    }
    __syncthreads();
    const int incl = x + (wid ? warp_sum[wid - 1] : 0) + carry_s;
    if (i < total) {
      cnt[i] = incl - v;
      if (i % nb == 0) base_s[i / nb] = incl - v;
    }
    __syncthreads();
    if (threadIdx.x == kRouteBlock - 1) carry_s = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) base_s[W] = carry_s;
  __syncthreads();
  if (threadIdx.x < W) counts_out[threadIdx.x] = base_s[threadIdx.x + 1] - base_s[threadIdx.x];
}

__global__ void __launch_bounds__(kRouteBlock) route_emit_kernel(const uint32_t* __restrict__ uniq,
                                                                 const int32_t* __restrict__ n_dev, int W, int nb,
                                                                 const int32_t* __restrict__ off /*[W][nb]*/,
                                                                 uint32_t* __restrict__ local_rows,
                                                                 int32_t* __restrict__ perm) {
  __shared__ int32_t warp_cnt[32];
  const int n = *n_dev;
  const int i = blockIdx.x * kRouteBlock + threadIdx.x;
  const uint32_t key = i < n ? __ldg(uniq + i) : 0u;
  const int owner = i < n ? (int)(key % (uint32_t)W) : -1;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int w = 0; w < W; ++w) {
    const int v = owner == w;
    const unsigned b = __ballot_sync(0xffffffffu, v);
    const int r = __popc(b & ((1u << lane) - 1u));
    if (lane == 0) warp_cnt[wid] = __popc(b);
    __syncthreads();
    if (wid == 0) {
      int s = warp_cnt[lane];
      const int orig = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += y;
      }
      warp_cnt[lane] = s - orig;
    }
    __syncthreads();
    if (v) {
      const int pos = off[w * nb + blockIdx.x] + warp_cnt[wid] + r;
      local_rows[pos] = key / (uint32_t)W;
      perm[i] = pos;
    }
    __syncthreads();
  }
}

}  // namespace tgr

using namespace tgr;

extern "C" size_t tgr_route_workspace_bytes(int64_t max_unique, int W) {
  const size_t nb = (size_t)((max_unique + kRouteBlock - 1) / kRouteBlock) + 1;
  return (nb * (size_t)W * sizeof(int32_t) + 255) / 256 * 256;
}

extern "C" int tgr_route_bucket(const uint32_t* uniq, const int32_t* n_unique_dev, int64_t max_unique, int W,
                                uint32_t* bucketed_local_rows, int32_t* perm, int32_t* counts_dev, void* workspace,
                                size_t workspace_bytes, void* stream) {
  TGR_REQUIRE(uniq && n_unique_dev && bucketed_local_rows && perm && counts_dev && workspace, "null argument");
  TGR_REQUIRE(W >= 1 && W <= kMaxW, "W=%d out of range", W);
  TGR_REQUIRE(max_unique >= 0 && max_unique < (1ll << 31), "max_unique out of range");
  TGR_REQUIRE(workspace_bytes >= tgr_route_workspace_bytes(max_unique, W), "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (max_unique == 0) {
    cudaMemsetAsync(counts_dev, 0, sizeof(int32_t) * W, st);
    return check_launch("route(empty)");
  }
  const int nb = (int)((max_unique + kRouteBlock - 1) / kRouteBlock);
  int32_t* cnt = (int32_t*)workspace;
  route_count_kernel<<<nb, kRouteBlock, 0, st>>>(uniq, n_unique_dev, W, nb, cnt);
  route_scan_kernel<<<1, kRouteBlock, 0, st>>>(cnt, W, nb, counts_dev);
  route_emit_kernel<<<nb, kRouteBlock, 0, st>>>(uniq, n_unique_dev, W, nb, cnt, bucketed_local_rows, perm);
  return check_launch("route_bucket");
}
