// Deterministic segmented reduction of concat-gradient rows by sorted key, and the AdamW row update.
//
// Replaces embedding_dense_backward's per-table sort + segmented sum into DENSE grads and the dense AdamW
// pass over every row with work proportional to the rows a step touched.
//
// Fixed tiling of the SORTED (key, src) list — the tiling depends on positions only, never on the data, so
// the summation order is a pure function of the sorted list (bitwise reproducible, no float atomics) and the
// load is perfectly balanced under Zipf skew (a 60k-entry run is simply 60 consecutive group tiles):
//
//   group tile = kC entries, one group of LANES = H/4 threads; runs inside the tile are summed sequentially
//   in sorted (= ascending position) order with kU gradient rows in flight; runs that cross group-tile borders
//   are stitched from head/tail partials in shared memory, runs that cross CTA tiles by a tiny fix-up kernel.
//
//   mode 0  a finished row goes to grads_out[seg_of_entry[...]]           (compact [U, H]; parity / sharded)
//   mode 1  a finished row goes to a per-CTA region (row_buf/row_keys/row_cnt) and a second, embarrassingly
//           parallel kernel applies AdamW to all finished rows. ncu on the first version (AdamW inline in the
//           streaming loop) showed 34 % active warps and long-scoreboard stalls on the dependent w/m/v loads at
//           3.0 TB/s, while random 256 B gathers reach 6.6+ TB/s on B200 (tools/membench): the update was
//           split out so the streaming loop only ever issues independent loads and fire-and-forget stores.
#include "tgr_common.cuh"
#include "tgr_rows.cuh"

namespace tgr {

constexpr int kRedThreads = 256;
constexpr int kC = 64;  // sorted entries per group tile
constexpr int kU = 8;   // gradient rows in flight per thread

struct RedParams {
  const char* chunk_base[TGR_MAX_CALLS][TGR_MAX_SLOTS];  // d(concat) base of (call, slot), offset to the slot's column
  int64_t ld_bytes[TGR_MAX_CALLS][TGR_MAX_SLOTS];
  const uint32_t* keys;
  const uint32_t* srcs;
  const int32_t* seg_of_entry;  // mode 0
  float* grads_out;             // mode 0: [U, H]
  uint32_t* row_keys;           // mode 1: [n_cta * TILE]
  float* row_buf;               // mode 1: [n_cta * TILE, H]
  int32_t* row_cnt;             // mode 1: [n_cta]
  float* cta_head;              // [n_cta, H]
  float* cta_tail;              // [n_cta, H]
  int64_t n;
  const int32_t* n_dev;         // != NULL: n is a capacity, the entry count is *n_dev when the kernels run
  int32_t H4;
  int32_t mode;
  __device__ __forceinline__ int64_t count() const { return n_dev ? min(n, (int64_t)__ldg(n_dev)) : n; }
};

template <int LANES>
__device__ __forceinline__ unsigned group_mask() {
  if constexpr (LANES == 32) return 0xffffffffu;
  const int g = (threadIdx.x & 31) / LANES;
  return ((1u << LANES) - 1u) << (g * LANES);
}

// destination of a finished row
template <int LANES, int TILE>
__device__ __forceinline__ float4* finish_dst(const RedParams& p, uint32_t key, int64_t entry, int lane, int* s_cnt) {
  const int H4 = p.H4;
  if (p.mode == 0) {
    const int u = __ldg(p.seg_of_entry + entry);
    return reinterpret_cast<float4*>(p.grads_out + (size_t)u * (size_t)(H4 * 4));
  }
  int slot = 0;
  if (lane == 0) slot = atomicAdd(s_cnt, 1);  // shared-memory counter: order of slots is irrelevant to the result
  slot = __shfl_sync(group_mask<LANES>(), slot, (threadIdx.x & 31) / LANES * LANES);
  const size_t r = (size_t)blockIdx.x * TILE + slot;
  if (lane == 0) p.row_keys[r] = key;
  return reinterpret_cast<float4*>(p.row_buf + r * (size_t)(H4 * 4));
}

template <bool BF16>
__device__ __forceinline__ float4 load_grad4(const RedParams& p, uint32_t src, int c) {
  const int call = src >> TGR_SRC_CALL_SHIFT;
  const int slot = (src >> TGR_SRC_SLOT_SHIFT) & 31;
  const uint32_t tok = src & TGR_SRC_TOKEN_MASK;
  const char* row = p.chunk_base[call][slot] + (size_t)tok * p.ld_bytes[call][slot];
  if constexpr (BF16) return unpack_bf16x4(ld_stream_u2(reinterpret_cast<const uint2*>(row) + c));
  else return ld_stream(reinterpret_cast<const float4*>(row) + c);
}

#define TGR_FOR_COLS(j, c) _Pragma("unroll") for (int j = 0, c = lane; j < NJ; ++j, c += LANES) if (c < H4)

// One CTA = G groups x kC sorted entries. NJ = float4 columns per lane (1 for H <= 128).
template <int LANES, int NJ, bool BF16>
__global__ void __launch_bounds__(kRedThreads, 4) reduce_tiles_kernel(const __grid_constant__ RedParams p) {
  constexpr int G = kRedThreads / LANES;
  constexpr int TILE = G * kC;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* s_keys = reinterpret_cast<uint32_t*>(smem_raw);        // [TILE + 2]  (index 0 = entry before the tile)
  uint32_t* s_srcs = s_keys + TILE + 2;                            // [TILE] (+2 pad keeps 16 B alignment)
  float4* s_head = reinterpret_cast<float4*>(s_srcs + TILE + 2);   // [G][H4]
  float4* s_tail = s_head + G * p.H4;                              // [G][H4]
  int32_t* s_flag = reinterpret_cast<int32_t*>(s_tail + G * p.H4); // [G] bit0 has_head, bit1 head_through, bit2 has_tail
  uint32_t* s_tkey = reinterpret_cast<uint32_t*>(s_flag + G);      // [G] key of the tail run
  int* s_cnt = reinterpret_cast<int*>(s_tkey + G);                 // finished rows of this CTA (mode 1)

  const int tid = threadIdx.x, lane = tid % LANES, grp = tid / LANES;
  const int H4 = p.H4;
  const int64_t n = p.count();
  const int64_t cta_a = (int64_t)blockIdx.x * TILE;
  if (cta_a >= n) {            // tile past the data (capacity-sized grid)
    if (p.mode == 1 && tid == 0) p.row_cnt[blockIdx.x] = 0;
    return;
  }
  const int cnt_cta = (int)(min(n, cta_a + TILE) - cta_a);

  for (int i = tid; i < cnt_cta + 2; i += kRedThreads) {
    const int64_t e = cta_a - 1 + i;
    s_keys[i] = (e >= 0 && e < n) ? __ldg(p.keys + e) : 0xFFFFFFFFu;  // 0xFFFFFFFF never equals a real key
  }
  for (int i = tid; i < cnt_cta; i += kRedThreads) s_srcs[i] = __ldg(p.srcs + cta_a + i);
  if (tid < G) s_flag[tid] = 0;
  if (tid == 0) *s_cnt = 0;
  __syncthreads();

  const int ga = grp * kC;
  const int gb = min(cnt_cta, ga + kC);
  if (ga < gb) {
    float4 acc[NJ];
    TGR_FOR_COLS(j, c) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t cur = s_keys[ga + 1];
    const bool from_prev = (s_keys[ga] == cur);
    int run_start = ga;
    for (int e0 = ga; e0 < gb; e0 += kU) {
      float4 g[kU][NJ];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        if (e0 + u < gb) {
          const uint32_t src = s_srcs[e0 + u];
          TGR_FOR_COLS(j, c) g[u][j] = load_grad4<BF16>(p, src, c);
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int e = e0 + u;
        if (e < gb) {
          const uint32_t k = s_keys[e + 1];
          if (k != cur) {  // the run [run_start, e) ended inside this tile
            if (run_start == ga && from_prev) {
              TGR_FOR_COLS(j, c) s_head[grp * H4 + c] = acc[j];
              if (lane == 0) s_flag[grp] |= 1;
            } else {
              float4* dst = finish_dst<LANES, TILE>(p, cur, cta_a + run_start, lane, s_cnt);
              TGR_FOR_COLS(j, c) dst[c] = acc[j];
            }
            cur = k;
            run_start = e;
            TGR_FOR_COLS(j, c) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          TGR_FOR_COLS(j, c) acc[j] = f4_add(acc[j], g[u][j]);
        }
      }
    }
    // last run reaches the end of the group tile
    const bool to_next = (s_keys[gb + 1] == cur);  // s_keys[cnt_cta + 1] is the entry after the CTA tile (or sentinel)
    if (run_start == ga && from_prev) {
      TGR_FOR_COLS(j, c) s_head[grp * H4 + c] = acc[j];
      if (lane == 0) s_flag[grp] |= to_next ? 3 : 1;
    } else if (to_next) {
      TGR_FOR_COLS(j, c) s_tail[grp * H4 + c] = acc[j];
      if (lane == 0) { s_flag[grp] |= 4; s_tkey[grp] = cur; }
    } else {
      float4* dst = finish_dst<LANES, TILE>(p, cur, cta_a + run_start, lane, s_cnt);
      TGR_FOR_COLS(j, c) dst[c] = acc[j];
    }
  }
  __syncthreads();

  // ---- stitch runs that cross group tiles, in fixed order ----
  const int g_active = (cnt_cta + kC - 1) / kC;
  if (grp < g_active) {
    const int fl = s_flag[grp];
    if (grp == 0 && (fl & 1)) {  // run entering the CTA from the previous one: partial for the cross-CTA fix-up
      float4 acc[NJ];
      TGR_FOR_COLS(j, c) acc[j] = s_head[c];
      if (fl & 2) {
        for (int q = 1; q < g_active; ++q) {
          TGR_FOR_COLS(j, c) acc[j] = f4_add(acc[j], s_head[q * H4 + c]);
          if (!(s_flag[q] & 2)) break;
        }
      }
      float4* dst = reinterpret_cast<float4*>(p.cta_head + (size_t)blockIdx.x * (size_t)(H4 * 4));
      TGR_FOR_COLS(j, c) dst[c] = acc[j];
    }
    if (fl & 4) {
      float4 acc[NJ];
      TGR_FOR_COLS(j, c) acc[j] = s_tail[grp * H4 + c];
      bool open = true;  // the run still continues past what has been summed
      for (int q = grp + 1; q < g_active; ++q) {
        TGR_FOR_COLS(j, c) acc[j] = f4_add(acc[j], s_head[q * H4 + c]);
        if (!(s_flag[q] & 2)) { open = false; break; }
      }
      float4* dst;
      if (open) dst = reinterpret_cast<float4*>(p.cta_tail + (size_t)blockIdx.x * (size_t)(H4 * 4));  // continues into the next CTA tile
      else dst = finish_dst<LANES, TILE>(p, s_tkey[grp], cta_a + min(cnt_cta, (grp + 1) * kC) - 1, lane, s_cnt);
      TGR_FOR_COLS(j, c) dst[c] = acc[j];
    }
  }
  if (p.mode == 1) {
    __syncthreads();
    if (tid == 0) p.row_cnt[blockIdx.x] = *s_cnt;
  }
}

// cross-CTA fix-up: one CTA per tile; the tile that holds the START of a run leaving it owns the run. The run's extent
// is found by galloping + bisection over the tile borders (the keys are sorted, so "this tile is entirely key K" is
// monotone), and the head partials of the tiles it covers are summed by the CTA's G groups in a fixed strided order
// (group q takes tiles first + q, first + q + G, ... ascending; the group sums are then added in group order). The
// first version walked the tiles one by one from a single group: a 56 k-entry run of a 100-row table cost 55 dependent
// round trips (38 us, 6 % active warps — profiles/README.md r1e).
template <int LANES, int NJ>
__global__ void __launch_bounds__(kRedThreads) reduce_fixup_kernel(const __grid_constant__ RedParams p, int n_cta) {
  constexpr int G = kRedThreads / LANES;
  constexpr int64_t TILE = (int64_t)G * kC;
  __shared__ float4 s_part[G][NJ * LANES];
  const int lane = threadIdx.x % LANES, grp = threadIdx.x / LANES;
  const int c0 = blockIdx.x;
  const int64_t n = p.count();
  const int H4 = p.H4;
  const int64_t a = (int64_t)c0 * TILE, b = min(n, a + TILE);
  if (b >= n) return;
  const uint32_t K = __ldg(p.keys + b - 1);
  if (__ldg(p.keys + b) != K) return;                                        // nothing leaves this tile
  if (a > 0 && __ldg(p.keys + a) == K && __ldg(p.keys + a - 1) == K) return;  // "through" tile: an earlier tile owns it
  // through(t): tile t lies inside the run and the run continues past it. First tile that is not: the run's last one.
  auto through = [&](int t) {
    const int64_t te = min(n, (int64_t)(t + 1) * TILE);
    return te < n && __ldg(p.keys + te - 1) == K && __ldg(p.keys + te) == K;
  };
  int lo = c0 + 1, hi = c0 + 1, step = 1;
  while (through(hi)) {              // through(n_cta - 1) is false: terminates
    lo = hi + 1;
    hi = min(hi + step, n_cta - 1);
    step <<= 1;
  }
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (through(mid)) lo = mid + 1; else hi = mid;
  }
  const int c1 = lo;                 // heads of tiles c0 + 1 .. c1 belong to the run
  float4 acc[NJ];
  TGR_FOR_COLS(j, c) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = c0 + 1 + grp; t <= c1; t += G) {
    const float4* hs = reinterpret_cast<const float4*>(p.cta_head + (size_t)t * (size_t)(H4 * 4));
    TGR_FOR_COLS(j, c) acc[j] = f4_add(acc[j], hs[c]);
  }
  TGR_FOR_COLS(j, c) s_part[grp][j * LANES + lane] = acc[j];
  __syncthreads();
  if (grp != 0) return;
  const float4* src = reinterpret_cast<const float4*>(p.cta_tail + (size_t)c0 * (size_t)(H4 * 4));
  TGR_FOR_COLS(j, c) acc[j] = src[c];
  const int used = min(G, c1 - c0);
  for (int q = 0; q < used; ++q) {
    TGR_FOR_COLS(j, c) acc[j] = f4_add(acc[j], s_part[q][j * LANES + lane]);
  }
  float4* dst;
  if (p.mode == 0) {
    dst = reinterpret_cast<float4*>(p.grads_out + (size_t)__ldg(p.seg_of_entry + b - 1) * (size_t)(H4 * 4));
  } else {
    // this tile finished at most TILE-1 rows itself (its tail run was left open), so one slot is free
    const int slot = p.row_cnt[c0];
    const size_t r = (size_t)c0 * TILE + slot;
    if (lane == 0) { p.row_keys[r] = K; }
    dst = reinterpret_cast<float4*>(p.row_buf + r * (size_t)(H4 * 4));
    __syncwarp(group_mask<LANES>());
    if (lane == 0) p.row_cnt[c0] = slot + 1;
  }
  TGR_FOR_COLS(j, c) dst[c] = acc[j];
}

// mode 1, second phase: AdamW on every finished row. grid = (regions, kAdamSplit): a region holds up to TILE finished rows
// (all of them when every key occurs once, as on the owner side of the sharded exchange), so its rows are split over
// kAdamSplit CTAs; two rows in flight per group.
constexpr int kAdamSplit = 4;
template <int LANES, int NJ>
__global__ void __launch_bounds__(kRedThreads) adam_regions_kernel(const __grid_constant__ RowParams rp,
                                                                   const uint32_t* __restrict__ row_keys,
                                                                   const float* __restrict__ row_buf,
                                                                   const int32_t* __restrict__ row_cnt) {
  constexpr int G = kRedThreads / LANES;
  constexpr int TILE = G * kC;
  const int lane = threadIdx.x % LANES, grp = threadIdx.x / LANES;
  const int H4 = rp.H4;
  const int cnt = row_cnt[blockIdx.x];
  const size_t base = (size_t)blockIdx.x * TILE;
  for (int j0 = grp + 2 * G * blockIdx.y; j0 < cnt; j0 += 2 * G * kAdamSplit) {
    float4 g[2][NJ], w[2][NJ], m[2][NJ], v[2][NJ];
    float4 *wp[2], *mp[2], *vp[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int jj = j0 + u * G;
      wp[u] = nullptr;
      if (jj < cnt) {
        const uint32_t key = __ldg(row_keys + base + jj);
        const int t = find_table(rp.key_base, rp.n_tables, key);
        const size_t row = (size_t)(key - rp.key_base[t]) * (size_t)(H4 * 4);
        wp[u] = reinterpret_cast<float4*>(rp.w[t] + row);
        mp[u] = reinterpret_cast<float4*>(rp.m[t] + row);
        vp[u] = reinterpret_cast<float4*>(rp.v[t] + row);
        const float4* gp = reinterpret_cast<const float4*>(row_buf + (base + jj) * (size_t)(H4 * 4));
        TGR_FOR_COLS(q, c) { g[u][q] = gp[c]; w[u][q] = wp[u][c]; m[u][q] = mp[u][c]; v[u][q] = vp[u][c]; }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (wp[u] != nullptr) {
        TGR_FOR_COLS(q, c) {
          float4 gg = g[u][q];
          gg.x *= rp.adam.grad_scale; gg.y *= rp.adam.grad_scale; gg.z *= rp.adam.grad_scale; gg.w *= rp.adam.grad_scale;
          adam_elem(w[u][q].x, m[u][q].x, v[u][q].x, gg.x, rp.adam);
          adam_elem(w[u][q].y, m[u][q].y, v[u][q].y, gg.y, rp.adam);
          adam_elem(w[u][q].z, m[u][q].z, v[u][q].z, gg.z, rp.adam);
          adam_elem(w[u][q].w, m[u][q].w, v[u][q].w, gg.w, rp.adam);
          wp[u][c] = w[u][q]; mp[u][c] = m[u][q]; vp[u][c] = v[u][q];
        }
      }
    }
  }
}

static int red_lanes(int H4) { return H4 <= 8 ? 8 : (H4 <= 16 ? 16 : 32); }
static int red_tile(int H4) { return (kRedThreads / red_lanes(H4)) * kC; }

template <int LANES, int NJ>
static int launch_reduce(RedParams& p, const RowParams* rp, bool bf16, int n_cta, cudaStream_t st) {
  constexpr int G = kRedThreads / LANES;
  constexpr int TILE = G * kC;
  const size_t smem = (size_t)(2 * TILE + 4) * 4 + (size_t)2 * G * p.H4 * 16 + (size_t)2 * G * 4 + 16;
  if (bf16) {
    { static bool tgr_attr_once_ = false; if (!tgr_attr_once_) { cudaFuncSetAttribute(reduce_tiles_kernel<LANES, NJ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); tgr_attr_once_ = true; } }
    TGR_K(reduce_tiles_kernel<LANES, NJ, true>)<<<n_cta, kRedThreads, smem, st>>>(p);
  } else {
    { static bool tgr_attr_once_ = false; if (!tgr_attr_once_) { cudaFuncSetAttribute(reduce_tiles_kernel<LANES, NJ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); tgr_attr_once_ = true; } }
    TGR_K(reduce_tiles_kernel<LANES, NJ, false>)<<<n_cta, kRedThreads, smem, st>>>(p);
  }
  if (int rc = check_launch("reduce_tiles")) return rc;
  if (n_cta > 1) {
    TGR_K(reduce_fixup_kernel<LANES, NJ>)<<<n_cta - 1, kRedThreads, 0, st>>>(p, n_cta);
    if (int rc = check_launch("reduce_fixup")) return rc;
  }
  if (p.mode == 1) {
    TGR_K(adam_regions_kernel<LANES, NJ>)<<<dim3(n_cta, kAdamSplit), kRedThreads, 0, st>>>(*rp, p.row_keys, p.row_buf, p.row_cnt);
    return check_launch("adam_regions");
  }
  return 0;
}

}  // namespace tgr

using namespace tgr;

// workspace: per-CTA head/tail partials; mode 1 additionally needs one row slot per sorted entry (sparsely
// touched: only the U finished rows are ever written) + its key + a counter per CTA tile
extern "C" size_t tgr_reduce_workspace_bytes(int64_t n, int H) {
  const int H4 = H / 4;
  const int tile = red_tile(H4);
  const int64_t n_cta = (n + tile - 1) / tile;
  const size_t part = align_up((size_t)(n_cta + 1) * H * sizeof(float));
  const size_t rows = align_up((size_t)n_cta * tile * H * sizeof(float)) + align_up((size_t)n_cta * tile * 4) +
                      align_up((size_t)(n_cta + 1) * 4);
  return 2 * part + rows;
}

extern "C" int tgr_bwd_reduce(const tgr_table_t* tables, int n_tables, int H, const tgr_call_t* calls, int n_calls,
                              const uint32_t* keys_sorted, const uint32_t* srcs_sorted, int64_t n, int mode,
                              const int32_t* seg_of_entry, float* grads_out, const tgr_adam_t* adam, void* workspace,
                              size_t workspace_bytes, void* stream) {
  return tgr::bwd_reduce_dn(tables, n_tables, H, calls, n_calls, keys_sorted, srcs_sorted, n, mode, seg_of_entry, grads_out, adam,
                            workspace, workspace_bytes, nullptr, stream);
}

int tgr::bwd_reduce_dn(const tgr_table_t* tables, int n_tables, int H, const tgr_call_t* calls, int n_calls,
                       const uint32_t* keys_sorted, const uint32_t* srcs_sorted, int64_t n, int mode,
                       const int32_t* seg_of_entry, float* grads_out, const tgr_adam_t* adam, void* workspace,
                       size_t workspace_bytes, const int32_t* n_dev, void* stream) {
  tgr::TimedScope tgr_timed_("bwd_reduce", stream);
  TGR_REQUIRE(calls && n_calls > 0 && n_calls <= TGR_MAX_CALLS, "bad calls");
  TGR_REQUIRE(H > 0 && H % 4 == 0 && H <= 512, "H=%d unsupported (multiple of 4, <= 512)", H);
  TGR_REQUIRE(mode == 0 || mode == 1, "bad mode");
  TGR_REQUIRE(n >= 0 && n < (1ll << 31), "n out of range");
  if (n == 0) return 0;
  TGR_REQUIRE(keys_sorted && srcs_sorted && workspace, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  RedParams p{};
  RowParams rp{};
  const int dtype = calls[0].cat_dtype;
  const size_t esz = dtype == TGR_DTYPE_BF16 ? 2 : 4;
  for (int c = 0; c < n_calls; ++c) {
    const tgr_call_t& cl = calls[c];
    TGR_REQUIRE(cl.cat_dtype == dtype, "all calls must share the concat-gradient dtype");
    TGR_REQUIRE(cl.n_slots >= 0 && cl.n_slots <= TGR_MAX_SLOTS, "call %d: bad n_slots", c);
    for (int i = 0; i < cl.n_slots; ++i) {
      const tgr_slot_t& s = cl.slots[i];
      if (s.kind == TGR_KIND_MM) continue;
      const char* base = (const char*)(s.side == TGR_SIDE_ITEM ? cl.item_cat : cl.user_cat);
      const int64_t ld = s.side == TGR_SIDE_ITEM ? cl.item_ld : cl.user_ld;
      TGR_REQUIRE(base != nullptr, "call %d slot %d: concat gradient is NULL", c, i);
      TGR_REQUIRE(s.col % 4 == 0 && ld % 4 == 0, "call %d slot %d: col/ld not 128-bit tileable", c, i);
      p.chunk_base[c][i] = base + (size_t)s.col * esz;
      p.ld_bytes[c][i] = ld * (int64_t)esz;
    }
  }
  p.keys = keys_sorted;
  p.srcs = srcs_sorted;
  p.n = n;
  p.n_dev = n_dev;
  p.H4 = H / 4;
  p.mode = mode;
  const int tile = red_tile(p.H4);
  const int n_cta = (int)((n + tile - 1) / tile);
  TGR_REQUIRE(workspace_bytes >= (mode == 1 ? tgr_reduce_workspace_bytes(n, H)
                                            : 2 * align_up((size_t)(n_cta + 1) * H * sizeof(float))),
              "workspace too small");
  const size_t part = align_up((size_t)(n_cta + 1) * H * sizeof(float));
  char* ws = (char*)workspace;
  p.cta_head = (float*)ws;
  p.cta_tail = (float*)(ws + part);
  if (mode == 0) {
    TGR_REQUIRE(seg_of_entry && grads_out, "mode 0 needs seg_of_entry / grads_out");
    p.seg_of_entry = seg_of_entry;
    p.grads_out = grads_out;
  } else {
    TGR_REQUIRE(adam != nullptr, "adam is NULL");
    if (int rc = fill_row_params(rp, tables, n_tables, H)) return rc;
    for (int t = 0; t < n_tables; ++t) TGR_REQUIRE(rp.w[t] && rp.m[t] && rp.v[t], "table %d: weight/exp_avg/exp_avg_sq NULL", t);
    rp.adam = *adam;
    p.row_buf = (float*)(ws + 2 * part);
    p.row_keys = (uint32_t*)(ws + 2 * part + align_up((size_t)n_cta * tile * H * sizeof(float)));
    p.row_cnt = (int32_t*)((char*)p.row_keys + align_up((size_t)n_cta * tile * 4));
  }
  const bool bf16 = dtype == TGR_DTYPE_BF16;
  const int H4 = p.H4;
  if (H4 <= 8) return launch_reduce<8, 1>(p, &rp, bf16, n_cta, st);
  if (H4 <= 16) return launch_reduce<16, 1>(p, &rp, bf16, n_cta, st);
  if (H4 <= 32) return launch_reduce<32, 1>(p, &rp, bf16, n_cta, st);
  if (H4 <= 64) return launch_reduce<32, 2>(p, &rp, bf16, n_cta, st);
  return launch_reduce<32, 4>(p, &rp, bf16, n_cta, st);
}

// Owner side of the row-sharded exchange without the gradient all-to-all: the contributions of source rank b are
// rows of that rank's bucketed gradient buffer, read IN PLACE over NVLink peer memory by the same fixed-tile
// segmented reduction (+ AdamW) — src = b << 24 | row inside row_bases[b].
extern "C" int tgr_bwd_reduce_rows(const tgr_table_t* tables, int n_tables, int H, const float* const* row_bases,
                                   int n_bases, const uint32_t* keys_sorted, const uint32_t* srcs_sorted, int64_t n,
                                   const tgr_adam_t* adam, void* workspace, size_t workspace_bytes, void* stream) {
  tgr::TimedScope tgr_timed_("bwd_reduce_rows", stream);
  TGR_REQUIRE(H > 0 && H % 4 == 0 && H <= 512, "H=%d unsupported (multiple of 4, <= 512)", H);
  TGR_REQUIRE(n >= 0 && n < (1ll << 31), "n out of range");
  TGR_REQUIRE(n_bases > 0 && n_bases <= TGR_MAX_SLOTS && row_bases, "n_bases out of range");
  if (n == 0) return 0;
  TGR_REQUIRE(keys_sorted && srcs_sorted && workspace && adam, "null argument");
  TGR_REQUIRE(workspace_bytes >= tgr_reduce_workspace_bytes(n, H), "workspace too small");
  RedParams p{};
  RowParams rp{};
  for (int b = 0; b < n_bases; ++b) {
    TGR_REQUIRE(row_bases[b] != nullptr, "row base %d is NULL", b);
    p.chunk_base[0][b] = (const char*)row_bases[b];     // call field 0, slot field = source rank
    p.ld_bytes[0][b] = (int64_t)H * 4;
  }
  p.keys = keys_sorted;
  p.srcs = srcs_sorted;
  p.n = n;
  p.H4 = H / 4;
  p.mode = 1;
  const int tile = red_tile(p.H4);
  const int n_cta = (int)((n + tile - 1) / tile);
  const size_t part = align_up((size_t)(n_cta + 1) * H * sizeof(float));
  char* ws = (char*)workspace;
  p.cta_head = (float*)ws;
  p.cta_tail = (float*)(ws + part);
  if (int rc = fill_row_params(rp, tables, n_tables, H)) return rc;
  for (int t = 0; t < n_tables; ++t) TGR_REQUIRE(rp.w[t] && rp.m[t] && rp.v[t], "table %d: weight/exp_avg/exp_avg_sq NULL", t);
  rp.adam = *adam;
  p.row_buf = (float*)(ws + 2 * part);
  p.row_keys = (uint32_t*)(ws + 2 * part + align_up((size_t)n_cta * tile * H * sizeof(float)));
  p.row_cnt = (int32_t*)((char*)p.row_keys + align_up((size_t)n_cta * tile * 4));
  cudaStream_t st = (cudaStream_t)stream;
  const int H4 = p.H4;
  if (H4 <= 8) return launch_reduce<8, 1>(p, &rp, false, n_cta, st);
  if (H4 <= 16) return launch_reduce<16, 1>(p, &rp, false, n_cta, st);
  if (H4 <= 32) return launch_reduce<32, 1>(p, &rp, false, n_cta, st);
  if (H4 <= 64) return launch_reduce<32, 2>(p, &rp, false, n_cta, st);
  return launch_reduce<32, 4>(p, &rp, false, n_cta, st);
}
