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
  __shared__ int32_t s_cnt[kMaxW];
  const int n = *n_dev;
  if ((int)blockIdx.x * kRouteBlock >= n) {   // the grid covers the capacity bound; blocks past *n_dev just report zeros
    if (threadIdx.x < W) cnt[threadIdx.x * nb + blockIdx.x] = 0;
    return;
  }
  if (threadIdx.x < W) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int i = blockIdx.x * kRouteBlock + threadIdx.x;
  const int owner = i < n ? (int)(__ldg(uniq + i) % (uint32_t)W) : -1;
  const int lane = threadIdx.x & 31;
  for (int w = 0; w < W; ++w) {               // warp ballots; integer shared-memory atomics (totals are order independent)
    const unsigned b = __ballot_sync(0xffffffffu, owner == w);
    if (lane == 0 && b) atomicAdd(&s_cnt[w], __popc(b));
  }
  __syncthreads();
  if (threadIdx.x < W) cnt[threadIdx.x * nb + blockIdx.x] = s_cnt[threadIdx.x];
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
  __shared__ uint16_t warp_cnt[kRouteBlock / 32][kMaxW];   // keys of owner w in warp i
  const int n = *n_dev;
  if ((int)blockIdx.x * kRouteBlock >= n) return;
  const int i = blockIdx.x * kRouteBlock + threadIdx.x;
  const uint32_t key = i < n ? __ldg(uniq + i) : 0u;
  const int owner = i < n ? (int)(key % (uint32_t)W) : -1;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int r = 0;                                  // rank among the warp's earlier keys of the same owner
  for (int w = 0; w < W; ++w) {
    const unsigned b = __ballot_sync(0xffffffffu, owner == w);
    if (owner == w) r = __popc(b & ((1u << lane) - 1u));
    if (lane == 0) warp_cnt[wid][w] = (uint16_t)__popc(b);
  }
  __syncthreads();
  if (owner >= 0) {
    int before = 0;                           // stable: earlier warps first
    for (int q = 0; q < wid; ++q) before += warp_cnt[q][owner];
    const int pos = off[owner * nb + blockIdx.x] + before + r;
    local_rows[pos] = key / (uint32_t)W;
    perm[i] = pos;
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
  tgr::TimedScope tgr_timed_("route_bucket", stream);
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
  TGR_K(route_count_kernel)<<<nb, kRouteBlock, 0, st>>>(uniq, n_unique_dev, W, nb, cnt);
  TGR_K(route_scan_kernel)<<<1, kRouteBlock, 0, st>>>(cnt, W, nb, counts_dev);
  TGR_K(route_emit_kernel)<<<nb, kRouteBlock, 0, st>>>(uniq, n_unique_dev, W, nb, cnt, bucketed_local_rows, perm);
  return check_launch("route_bucket");
}

// ---- id remap / row permutation for the sharded exchange ------------------------------------------------
namespace tgr {

struct RemapCols {
  uint32_t key_base[TGR_MAX_SLOTS];
  int32_t rows[TGR_MAX_SLOTS];
};

// out[i] = 1 + perm[index of (key_base[col] + ids[i]) in uniq]   (0 for padding / out-of-range ids)
__global__ void __launch_bounds__(256) remap_ids_kernel(const int32_t* __restrict__ ids, int64_t n, int n_cols,
                                                        const __grid_constant__ RemapCols cols,
                                                        const uint32_t* __restrict__ uniq,
                                                        const int32_t* __restrict__ n_unique_dev,
                                                        const int32_t* __restrict__ perm, int32_t* __restrict__ out) {
  const int U = *n_unique_dev;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % n_cols);
    const int id = __ldg(ids + i);
    int r = 0;
    if (id > 0 && id < cols.rows[c]) {
      const uint32_t key = cols.key_base[c] + (uint32_t)id;
      int lo = 0, hi = U;  // first index with uniq[idx] >= key
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(uniq + mid) < key) lo = mid + 1; else hi = mid;
      }
      if (lo < U && __ldg(uniq + lo) == key) r = 1 + (perm ? __ldg(perm + lo) : lo);
    }
    out[i] = r;
  }
}

// out[perm[u], :] = in[u, :]  (u < *n_dev)   or, inverse = 1:  out[u, :] = in[perm[u], :]
__global__ void __launch_bounds__(256) permute_rows_kernel(const float* __restrict__ in, int H4,
                                                           const int32_t* __restrict__ perm,
                                                           const int32_t* __restrict__ n_dev, int inverse,
                                                           float* __restrict__ out) {
  const int n = *n_dev;
  const int64_t total = (int64_t)n * H4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int u = (int)(i / H4), c = (int)(i - (int64_t)u * H4);
    const int64_t p = __ldg(perm + u);
    if (inverse) reinterpret_cast<float4*>(out)[i] = __ldg(reinterpret_cast<const float4*>(in) + p * H4 + c);
    else reinterpret_cast<float4*>(out)[p * H4 + c] = __ldg(reinterpret_cast<const float4*>(in) + i);
  }
}

}  // namespace tgr

extern "C" int tgr_remap_ids(const int32_t* ids, int64_t n, int n_cols, const uint32_t* col_key_base,
                             const int32_t* col_rows, const uint32_t* uniq, const int32_t* n_unique_dev,
                             const int32_t* perm, int32_t* out, void* stream) {
  tgr::TimedScope tgr_timed_("remap_ids", stream);
  TGR_REQUIRE(n >= 0 && n_cols > 0 && n_cols <= TGR_MAX_SLOTS, "bad n / n_cols");
  if (n == 0) return 0;
  TGR_REQUIRE(ids && col_key_base && col_rows && uniq && n_unique_dev && out, "null argument");
  RemapCols cols{};
  for (int c = 0; c < n_cols; ++c) { cols.key_base[c] = col_key_base[c]; cols.rows[c] = col_rows[c]; }
  int64_t blocks = (n + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  TGR_K(remap_ids_kernel)<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(ids, n, n_cols, cols, uniq, n_unique_dev, perm, out);
  return check_launch("remap_ids");
}

extern "C" int tgr_permute_rows(const float* in, int H, const int32_t* perm, const int32_t* n_dev, int64_t max_n,
                                int inverse, float* out, void* stream) {
  tgr::TimedScope tgr_timed_("permute_rows", stream);
  TGR_REQUIRE(in && perm && n_dev && out, "null argument");
  TGR_REQUIRE(H > 0 && H % 4 == 0, "bad H");
  if (max_n <= 0) return 0;
  int64_t blocks = (max_n * (H / 4) + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  TGR_K(permute_rows_kernel)<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, H / 4, perm, n_dev, inverse, out);
  return check_launch("permute_rows");
}

// ---- remap by scatter: the sorted (key, src) pairs already know where every id lives -------------------------
namespace tgr {
struct ScatterParams {
  int32_t* out[TGR_MAX_CALLS];
  int32_t n_cols[TGR_MAX_CALLS];
  int8_t col_of_slot[TGR_MAX_CALLS][TGR_MAX_SLOTS];  // ids column of a SINGLE slot, -1 otherwise
};
__global__ void __launch_bounds__(256) remap_scatter_kernel(const uint32_t* __restrict__ srcs,
                                                            const int32_t* __restrict__ seg_of_entry, int64_t n,
                                                            const int32_t* __restrict__ perm,
                                                            const __grid_constant__ ScatterParams p,
                                                            const int32_t* __restrict__ n_dev) {
  if (n_dev) n = min(n, (int64_t)__ldg(n_dev));
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t src = __ldg(srcs + e);
    const int call = src >> TGR_SRC_CALL_SHIFT;
    const int slot = (src >> TGR_SRC_SLOT_SHIFT) & 31;
    const int col = p.col_of_slot[call][slot];
    if (col < 0) continue;  // array values are remapped by tgr_remap_ids (a token may hold several)
    const uint32_t tok = src & TGR_SRC_TOKEN_MASK;
    const int u = __ldg(seg_of_entry + e);
    p.out[call][(size_t)tok * p.n_cols[call] + col] = 1 + (perm ? __ldg(perm + u) : u);
  }
}
}  // namespace tgr

extern "C" int tgr_remap_scatter(const uint32_t* srcs_sorted, const int32_t* seg_of_entry, int64_t n, const int32_t* perm,
                                 const tgr_call_t* calls, int n_calls, int32_t* const* ids_out, void* stream) {
  return tgr::remap_scatter_dn(srcs_sorted, seg_of_entry, n, perm, calls, n_calls, ids_out, nullptr, stream);
}

int tgr::remap_scatter_dn(const uint32_t* srcs_sorted, const int32_t* seg_of_entry, int64_t n, const int32_t* perm,
                          const tgr_call_t* calls, int n_calls, int32_t* const* ids_out, const int32_t* n_dev, void* stream) {
  tgr::TimedScope tgr_timed_("remap_scatter", stream);
  TGR_REQUIRE(n >= 0 && n_calls > 0 && n_calls <= TGR_MAX_CALLS, "bad n / n_calls");
  if (n == 0) return 0;
  TGR_REQUIRE(srcs_sorted && seg_of_entry && calls && ids_out, "null argument");
  ScatterParams p{};
  for (int c = 0; c < n_calls; ++c) {
    TGR_REQUIRE(ids_out[c] != nullptr, "ids_out[%d] is NULL", c);
    p.out[c] = ids_out[c];
    p.n_cols[c] = calls[c].n_single;
    for (int i = 0; i < TGR_MAX_SLOTS; ++i) p.col_of_slot[c][i] = -1;
    for (int i = 0; i < calls[c].n_slots; ++i)
      if (calls[c].slots[i].kind == TGR_KIND_SINGLE) p.col_of_slot[c][i] = (int8_t)calls[c].slots[i].src;
  }
  int64_t blocks = (n + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  TGR_K(remap_scatter_kernel)<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(srcs_sorted, seg_of_entry, n, perm, p, n_dev);
  return check_launch("remap_scatter");
}

// ---- all ARRAY slots of up to 4 calls in ONE launch (they are few values each: one launch instead of 4..16) ------
namespace tgr {
constexpr int kMaxArrSegs = TGR_MAX_CALLS * TGR_MAX_ARRAYS;
struct ArrSegs {
  const int32_t* vals[kMaxArrSegs];
  int32_t* out[kMaxArrSegs];
  int32_t first[kMaxArrSegs + 1];   // prefix of the value counts
  uint32_t key_base[kMaxArrSegs];
  int32_t rows[kMaxArrSegs];
  int32_t n_seg;
};
__global__ void __launch_bounds__(256) remap_arrays_kernel(const __grid_constant__ ArrSegs a, const uint32_t* __restrict__ uniq,
                                                           const int32_t* __restrict__ n_unique_dev,
                                                           const int32_t* __restrict__ perm) {
  const int U = *n_unique_dev;
  const int total = a.first[a.n_seg];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int s = 0;
    while (i >= a.first[s + 1]) ++s;
    const int j = i - a.first[s];
    const int id = __ldg(a.vals[s] + j);
    int r = 0;
    if (id > 0 && id < a.rows[s]) {
      const uint32_t key = a.key_base[s] + (uint32_t)id;
      int lo = 0, hi = U;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(uniq + mid) < key) lo = mid + 1; else hi = mid;
      }
      if (lo < U && __ldg(uniq + lo) == key) r = 1 + (perm ? __ldg(perm + lo) : lo);
    }
    a.out[s][j] = r;
  }
}
}  // namespace tgr

extern "C" int tgr_remap_arrays(const tgr_table_t* tables, int n_tables, const tgr_call_t* calls, int n_calls,
                                const uint32_t* uniq, const int32_t* n_unique_dev, const int32_t* perm,
                                int32_t* const* arr_out, void* stream) {
  tgr::TimedScope tgr_timed_("remap_arrays", stream);
  TGR_REQUIRE(tables && calls && uniq && n_unique_dev && arr_out, "null argument");
  TGR_REQUIRE(n_calls > 0 && n_calls <= TGR_MAX_CALLS, "bad n_calls");
  ArrSegs a{};
  int ns = 0, total = 0;
  for (int c = 0; c < n_calls; ++c) {
    const tgr_call_t& cl = calls[c];
    for (int i = 0; i < cl.n_slots; ++i) {
      const tgr_slot_t& s = cl.slots[i];
      if (s.kind != TGR_KIND_ARRAY) continue;
      TGR_REQUIRE(s.src >= 0 && s.src < cl.n_arrays && s.src < TGR_MAX_ARRAYS, "call %d slot %d: bad array index", c, i);
      if (cl.arr_nnz[s.src] <= 0) continue;
      TGR_REQUIRE(s.table >= 0 && s.table < n_tables, "call %d slot %d: bad table", c, i);
      TGR_REQUIRE(cl.arr_val && arr_out[c], "call %d: array pointers NULL", c);
      a.vals[ns] = cl.arr_val + cl.arr_begin[s.src];
      a.out[ns] = arr_out[c] + cl.arr_begin[s.src];
      a.first[ns] = total;
      a.key_base[ns] = (uint32_t)tables[s.table].key_base;
      a.rows[ns] = (int32_t)tables[s.table].rows;
      total += cl.arr_nnz[s.src];
      ++ns;
    }
  }
  a.first[ns] = total;
  a.n_seg = ns;
  if (total == 0) return 0;
  int blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  TGR_K(remap_arrays_kernel)<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, uniq, n_unique_dev, perm);
  return check_launch("remap_arrays");
}

// ---- rows of this step's unique keys fetched in place from their owners' shards (NVLink peer memory) -----------------
// out[u, :] = peer[key % W][key / W, :]. A pure latency-hiding kernel (remote reads take ~2 us and bypass the local L2):
// one H/4-lane group per row, 4 rows in flight per lane, so it is launched on a side stream next to value-independent
// work instead of stalling the projection GEMM behind the fabric.
namespace tgr {
struct PeerPtrs { const float* p[TGR_MAX_PEERS]; };
template <int LANES>
__global__ void __launch_bounds__(256) fetch_peer_rows_kernel(const __grid_constant__ PeerPtrs peers, int W, int H4,
                                                              const uint32_t* __restrict__ uniq,
                                                              const int32_t* __restrict__ n_dev, float* __restrict__ out) {
  constexpr int G = 256 / LANES, UN = 4;
  const int n = *n_dev;
  const int lane = threadIdx.x % LANES, grp = threadIdx.x / LANES;
  for (int u0 = (blockIdx.x * G + grp) * UN; u0 < n; u0 += gridDim.x * G * UN) {
    for (int c = lane; c < H4; c += LANES) {
      float4 v[UN];
#pragma unroll
      for (int j = 0; j < UN; ++j) {
        if (u0 + j < n) {
          const uint32_t key = __ldg(uniq + u0 + j);
          const float4* src = reinterpret_cast<const float4*>(peers.p[key % (uint32_t)W] + (size_t)(key / (uint32_t)W) * (H4 * 4));
          v[j] = ld_stream(src + c);
        }
      }
#pragma unroll
      for (int j = 0; j < UN; ++j)
        if (u0 + j < n) reinterpret_cast<float4*>(out)[(size_t)(u0 + j) * H4 + c] = v[j];
    }
  }
}
}  // namespace tgr

extern "C" int tgr_fetch_peer_rows(const float* const* peer_rows, int n_peers, int H, const uint32_t* uniq,
                                   const int32_t* n_unique_dev, int64_t max_unique, float* out, void* stream) {
  tgr::TimedScope tgr_timed_("fetch_peer_rows", stream);
  TGR_REQUIRE(peer_rows && uniq && n_unique_dev && out, "null argument");
  TGR_REQUIRE(n_peers > 0 && n_peers <= TGR_MAX_PEERS, "n_peers out of range");
  TGR_REQUIRE(H > 0 && H % 4 == 0, "bad H");
  if (max_unique <= 0) return 0;
  PeerPtrs pp{};
  for (int r = 0; r < n_peers; ++r) {
    TGR_REQUIRE(peer_rows[r] != nullptr, "peer %d: shard pointer is NULL", r);
    pp.p[r] = peer_rows[r];
  }
  const int H4 = H / 4;
  const int lanes = H4 <= 8 ? 8 : (H4 <= 16 ? 16 : 32);
  int64_t blocks = (max_unique + (256 / lanes) * 4 - 1) / ((256 / lanes) * 4);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (lanes == 8) TGR_K(fetch_peer_rows_kernel<8>)<<<(unsigned)blocks, 256, 0, st>>>(pp, n_peers, H4, uniq, n_unique_dev, out);
  else if (lanes == 16) TGR_K(fetch_peer_rows_kernel<16>)<<<(unsigned)blocks, 256, 0, st>>>(pp, n_peers, H4, uniq, n_unique_dev, out);
  else TGR_K(fetch_peer_rows_kernel<32>)<<<(unsigned)blocks, 256, 0, st>>>(pp, n_peers, H4, uniq, n_unique_dev, out);
  return check_launch("fetch_peer_rows");
}
