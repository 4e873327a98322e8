// Backward of the sparse-feature embedding path: deterministic sparse gradient + fused AdamW row update.
//
// Replaces what autograd + the optimizer do under the reference's feat2emb:
//   embedding_dense_backward x 54 per step (sort + segmented sum into DENSE zero-filled [rows, H] grads)
//   + dense AdamW over EVERY row of every table (model/BaseLine/main.py:131,188-190)
// with a pipeline that touches only the rows a step actually used:
//
//   build_keys   (key = table.key_base + id, src = call|slot|token) for every non-padding id of up to 4
//                calls, compacted in (call, token, slot) order (count -> scan -> emit, no atomics)
//   sort_pairs   (tgr_sort.cu) stable LSD radix sort on the key bits only (stability keeps each row's contributions
//                in ascending (call, token) order = embedding_dense_backward's per-row order, F16)
//   dedup        run-length encode -> unique keys / segment offsets / per-entry segment index (8 keys per thread; the
//                emit pass can also write the remapped ids of the SINGLE slots)
//   reduce       fixed-tile segmented sum of the concat-gradient rows the sources point at; short runs
//                are summed sequentially in order, runs crossing tile borders are stitched from
//                per-tile partials in a fixed order (bitwise reproducible, no float atomics);
//                mode 1 applies the AdamW row update in the same pass (w, m, v read+written once)
//
// All kernels are HBM/L2-bound integer/byte movers: 128-bit row accesses, one LANES=H/4 thread
// group per gradient row, grids sized by the entry count — host-known, or an upper bound with the count read from device
// memory (the *_dn variants: what makes the step capturable in a CUDA graph).
#include "tgr_common.cuh"
#include "tgr_rows.cuh"

namespace tgr {

// =================================================================================================
// single-CTA exclusive scan of per-block counts (shared by build_keys and dedup)
// =================================================================================================
constexpr int kScanBlock = 1024;

// exclusive scan of block counts in place; total -> *total_out (and optional second copy)
__global__ void __launch_bounds__(kScanBlock) block_scan_kernel(int32_t* __restrict__ block_count, int n_blocks,
                                                                int32_t* __restrict__ total_out) {
  __shared__ int32_t warp_sum[32];
  __shared__ int32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < n_blocks; base += kScanBlock) {
    const int i = base + threadIdx.x;
    const int v = i < n_blocks ? block_count[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sum[wid] = x;
    __syncthreads();
    if (wid == 0) {
      int w = warp_sum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      warp_sum[lane] = w;  // inclusive
    }
    __syncthreads();
    const int carry = carry_s;
    const int incl = x + (wid ? warp_sum[wid - 1] : 0) + carry;
    if (i < n_blocks) block_count[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == kScanBlock - 1) carry_s = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry_s;
}

// =================================================================================================
// build_keys
// =================================================================================================
constexpr int kMaxKeySegs = TGR_MAX_CALLS * (1 + TGR_MAX_ARRAYS);

struct KeySeg {
  const int32_t* vals;  // SINGLE: ids matrix; ARRAY: arr_val + arr_begin
  const int32_t* toks;  // ARRAY: token of each value
  int64_t start;        // first global entry
  int64_t count;
  int32_t call;
  int32_t n_cols;       // SINGLE: n_single (>0); ARRAY: 0
  int32_t slot;         // ARRAY: slot index in the call
  uint32_t key_base;    // ARRAY
  int32_t rows;         // ARRAY
  int32_t pad;
};

struct KeyParams {
  KeySeg seg[kMaxKeySegs];
  uint32_t col_key_base[TGR_MAX_CALLS][TGR_MAX_SLOTS];  // SINGLE: by ids column
  int32_t col_rows[TGR_MAX_CALLS][TGR_MAX_SLOTS];
  uint8_t col_slot[TGR_MAX_CALLS][TGR_MAX_SLOTS];
  int32_t n_seg;
};

// ---- block-wise key builder: 2048 consecutive entries of ONE segment per CTA, 8 per thread ------------------------
// (the first version decoded one entry per thread through the generic compaction functor: 126 us per step, bound by
//  per-entry divisions and divergent constant-bank lookups; here the per-column tables sit in shared memory, the
//  (token, column) pair is advanced incrementally and the compacted pairs leave through a coalesced copy-out)
constexpr int kKT = 256;             // threads
constexpr int kKE = 8;               // entries per thread
constexpr int kKB = kKT * kKE;       // entries per CTA

struct KeyBlockParams {
  KeyParams kp;
  int32_t seg_first_block[kMaxKeySegs + 1];
};

template <bool EMIT>
__global__ void __launch_bounds__(kKT) keys_block_kernel(const __grid_constant__ KeyBlockParams P,
                                                         int32_t* __restrict__ block_cnt,   // count: out; emit: exclusive offsets
                                                         uint32_t* __restrict__ keys, uint32_t* __restrict__ srcs) {
  __shared__ uint32_t s_base[TGR_MAX_SLOTS];
  __shared__ int32_t s_rows[TGR_MAX_SLOTS];
  __shared__ uint32_t s_slot[TGR_MAX_SLOTS];
  __shared__ int32_t s_warp[kKT / 32];
  __shared__ uint32_t s_k[EMIT ? kKB : 1], s_s[EMIT ? kKB : 1];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int s = 0;
  while (s + 1 < P.kp.n_seg && (int)blockIdx.x >= P.seg_first_block[s + 1]) ++s;
  const KeySeg& g = P.kp.seg[s];
  if (g.n_cols > 0 && tid < g.n_cols) {
    s_base[tid] = P.kp.col_key_base[g.call][tid];
    s_rows[tid] = P.kp.col_rows[g.call][tid];
    s_slot[tid] = ((uint32_t)g.call << TGR_SRC_CALL_SHIFT) | ((uint32_t)P.kp.col_slot[g.call][tid] << TGR_SRC_SLOT_SHIFT);
  }
  __syncthreads();
  const uint32_t cnt = (uint32_t)g.count;
  const uint32_t i0 = (uint32_t)(blockIdx.x - P.seg_first_block[s]) * kKB + tid * kKE;
  int id[kKE];
  const int32_t* src = g.vals + i0;
  if (i0 + kKE <= cnt && (((uintptr_t)src) & 15) == 0) {
    const int4 a = __ldg(reinterpret_cast<const int4*>(src)), b = __ldg(reinterpret_cast<const int4*>(src) + 1);
    id[0] = a.x; id[1] = a.y; id[2] = a.z; id[3] = a.w; id[4] = b.x; id[5] = b.y; id[6] = b.z; id[7] = b.w;
  } else {
#pragma unroll
    for (int k = 0; k < kKE; ++k) id[k] = i0 + k < cnt ? __ldg(src + k) : 0;
  }
  uint32_t key[kKE], sc[kKE];
  unsigned vm = 0;
  if (g.n_cols > 0) {
    uint32_t t = i0 / (uint32_t)g.n_cols;
    int c = (int)(i0 - t * (uint32_t)g.n_cols);
#pragma unroll
    for (int k = 0; k < kKE; ++k) {
      if (id[k] > 0 && id[k] < s_rows[c]) {
        vm |= 1u << k;
        key[k] = s_base[c] + (uint32_t)id[k];
        sc[k] = s_slot[c] | t;
      }
      if (++c == g.n_cols) { c = 0; ++t; }
    }
  } else {
    const uint32_t hi = ((uint32_t)g.call << TGR_SRC_CALL_SHIFT) | ((uint32_t)g.slot << TGR_SRC_SLOT_SHIFT);
#pragma unroll
    for (int k = 0; k < kKE; ++k) {
      if (id[k] > 0 && id[k] < g.rows) {
        vm |= 1u << k;
        key[k] = g.key_base + (uint32_t)id[k];
        sc[k] = EMIT ? (hi | (uint32_t)__ldg(g.toks + i0 + k)) : 0u;
      }
    }
  }
  const int mine = __popc(vm);
  int x = mine;   // inclusive warp scan
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_warp[wid] = x;
  __syncthreads();
  int before = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kKT / 32; ++w) {
    const int v = s_warp[w];
    if (w < wid) before += v;
    total += v;
  }
  if (!EMIT) {
    if (tid == 0) block_cnt[blockIdx.x] = total;
    return;
  }
  int pos = before + x - mine;
#pragma unroll
  for (int k = 0; k < kKE; ++k)
    if (vm & (1u << k)) { s_k[pos] = key[k]; s_s[pos] = sc[k]; ++pos; }
  __syncthreads();
  const size_t off = (size_t)block_cnt[blockIdx.x];
  for (int j = tid; j < total; j += kKT) { keys[off + j] = s_k[j]; srcs[off + j] = s_s[j]; }
}

// =================================================================================================
// dedup (run-length encode of sorted keys)
// =================================================================================================
// ---- dedup of the sorted keys: 8 keys per thread, 2048 per CTA (the generic one-entry-per-thread compaction above spent
//      17 + 26 us on 2.85 M keys; profiles/README.md r2 graph timeline). The emit pass optionally does the id remap of the
//      SINGLE slots as well (what tgr_remap_scatter does from seg_of_entry): one read of the sorted payloads instead of two
//      passes over the entry list. ----
constexpr int kDdThreads = 256, kDdItems = 8, kDdTile = kDdThreads * kDdItems;
struct DedupScatter {
  int32_t* out[TGR_MAX_CALLS];                          // NULL table => no remap
  int32_t n_cols[TGR_MAX_CALLS];
  int8_t col_of_slot[TGR_MAX_CALLS][TGR_MAX_SLOTS];
  const uint32_t* srcs;
};

__device__ __forceinline__ void dd_load(const uint32_t* __restrict__ k, int64_t base, int64_t n, bool vec, uint32_t (&key)[kDdItems],
                                        uint32_t& prev) {
  if (vec && base + kDdItems <= n) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(k + base)), b = __ldg(reinterpret_cast<const uint4*>(k + base) + 1);
    key[0] = a.x; key[1] = a.y; key[2] = a.z; key[3] = a.w; key[4] = b.x; key[5] = b.y; key[6] = b.z; key[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < kDdItems; ++j) key[j] = base + j < n ? __ldg(k + base + j) : 0u;
  }
  prev = (base > 0 && base < n) ? __ldg(k + base - 1) : ~key[0];   // entry 0 always opens a run
}

__global__ void __launch_bounds__(kDdThreads) dedup_count_kernel(const uint32_t* __restrict__ k, int64_t n, int32_t* __restrict__ block_count,
                                                                 const int32_t* __restrict__ n_dev, int vec) {
  __shared__ int32_t ws[kDdThreads / 32];
  if (n_dev) n = min(n, (int64_t)__ldg(n_dev));
  const int64_t base = (int64_t)blockIdx.x * kDdTile + threadIdx.x * kDdItems;
  int c = 0;
  if (base < n) {
    uint32_t key[kDdItems], prev;
    dd_load(k, base, n, vec != 0, key, prev);
#pragma unroll
    for (int j = 0; j < kDdItems; ++j) {
      if (base + j < n) c += key[j] != prev;
      prev = key[j];
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < kDdThreads / 32; ++w) t += ws[w];
    block_count[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(kDdThreads) dedup_emit_kernel(const uint32_t* __restrict__ k, int64_t n, const int32_t* __restrict__ block_off,
                                                                uint32_t* __restrict__ uniq, int32_t* __restrict__ seg_off,
                                                                int32_t* __restrict__ seg_of_entry, const __grid_constant__ DedupScatter sc,
                                                                const int32_t* __restrict__ n_dev, int vec) {
  __shared__ int32_t ws[kDdThreads / 32];
  if (n_dev) n = min(n, (int64_t)__ldg(n_dev));
  const int64_t base = (int64_t)blockIdx.x * kDdTile + threadIdx.x * kDdItems;
  uint32_t key[kDdItems], prev = 0;
  int c = 0;
  unsigned heads = 0;
  if (base < n) {
    dd_load(k, base, n, vec != 0, key, prev);
#pragma unroll
    for (int j = 0; j < kDdItems; ++j) {
      if (base + j < n && key[j] != prev) { heads |= 1u << j; ++c; }
      prev = key[j];
    }
  }
  // exclusive scan of the per-thread head counts over the CTA, in thread order
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int x = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) ws[wid] = x;
  __syncthreads();
  int before = 0;
#pragma unroll
  for (int w = 0; w < kDdThreads / 32; ++w) before += w < wid ? ws[w] : 0;
  if (base >= n) return;
  int pos = __ldg(block_off + blockIdx.x) + before + x - c;   // heads before this thread's first entry
  int32_t seg[kDdItems];
#pragma unroll
  for (int j = 0; j < kDdItems; ++j) {
    if (heads & (1u << j)) {
      uniq[pos] = key[j];
      seg_off[pos] = (int32_t)(base + j);
      ++pos;
    }
    seg[j] = pos - 1;                                          // index of the last head at or before the entry
  }
  const bool full = vec && base + kDdItems <= n;
  if (seg_of_entry) {
    if (full) {
      reinterpret_cast<int4*>(seg_of_entry + base)[0] = make_int4(seg[0], seg[1], seg[2], seg[3]);
      reinterpret_cast<int4*>(seg_of_entry + base)[1] = make_int4(seg[4], seg[5], seg[6], seg[7]);
    } else {
#pragma unroll
      for (int j = 0; j < kDdItems; ++j) if (base + j < n) seg_of_entry[base + j] = seg[j];
    }
  }
  if (sc.srcs) {   // ids of the SINGLE slots -> 1 + unique index, scattered to (call, token, column)
    uint32_t src[kDdItems];
    if (full) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(sc.srcs + base)), b = __ldg(reinterpret_cast<const uint4*>(sc.srcs + base) + 1);
      src[0] = a.x; src[1] = a.y; src[2] = a.z; src[3] = a.w; src[4] = b.x; src[5] = b.y; src[6] = b.z; src[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < kDdItems; ++j) src[j] = base + j < n ? __ldg(sc.srcs + base + j) : 0u;
    }
#pragma unroll
    for (int j = 0; j < kDdItems; ++j) {
      if (base + j >= n) continue;
      const int call = src[j] >> TGR_SRC_CALL_SHIFT;
      const int col = sc.col_of_slot[call][(src[j] >> TGR_SRC_SLOT_SHIFT) & 31];
      if (col < 0) continue;    // array values are remapped by tgr_remap_arrays (a token may hold several)
      sc.out[call][(size_t)(src[j] & TGR_SRC_TOKEN_MASK) * sc.n_cols[call] + col] = 1 + seg[j];
    }
  }
}

__global__ void dedup_finish_kernel(int32_t* seg_off, int32_t* n_unique_dev, int64_t n, const int32_t* n_dev) {
  if (n_dev) n = min(n, (int64_t)*n_dev);
  if (n == 0) *n_unique_dev = 0;   // (entry 0 always opens a run: an empty list must not report one)
  seg_off[*n_unique_dev] = (int32_t)n;
}

// =================================================================================================
// row-wise kernels over the compact unique list
// =================================================================================================
template <int MODE>  // 0: adam, 1: scatter-add into dense grads
__global__ void __launch_bounds__(256) rows_kernel(const __grid_constant__ RowParams p, const uint32_t* __restrict__ uniq,
                                                   const float* __restrict__ grads, const int32_t* __restrict__ n_dev) {
  const int n = *n_dev;
  const int H4 = p.H4;
  const tgr_adam_t ad = (MODE == 0 && p.adam_dev) ? *p.adam_dev : p.adam;
  const int64_t total = (int64_t)n * H4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int u = (int)(i / H4), c = (int)(i - (int64_t)u * H4);
    const uint32_t key = __ldg(uniq + u);
    const int t = find_table(p.key_base, p.n_tables, key);
    const size_t row = (size_t)(key - p.key_base[t]) * (size_t)(H4 * 4);
    const float4 g = __ldg(reinterpret_cast<const float4*>(grads) + i);
    if (MODE == 0) {
      adam_row4(reinterpret_cast<float4*>(p.w[t] + row) + c, reinterpret_cast<float4*>(p.m[t] + row) + c,
                reinterpret_cast<float4*>(p.v[t] + row) + c, g, ad);
    } else if (p.grad[t] != nullptr) {  // tables without a dense gradient target are skipped
      float4* d = reinterpret_cast<float4*>(p.grad[t] + row) + c;
      float4 o = *d;
      *d = f4_add(o, g);
    }
  }
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ table, int H4,
                                                          const uint32_t* __restrict__ rows,
                                                          const int32_t* __restrict__ n_dev, float* __restrict__ out) {
  const int n = *n_dev;
  const int64_t total = (int64_t)n * H4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int u = (int)(i / H4), c = (int)(i - (int64_t)u * H4);
    const size_t r = __ldg(rows + u);
    reinterpret_cast<float4*>(out)[i] = __ldg(reinterpret_cast<const float4*>(table + r * (size_t)(H4 * 4)) + c);
  }
}

// ---- helpers -------------------------------------------------------------------------------------
}  // namespace tgr

// =================================================================================================
// C ABI
// =================================================================================================
using namespace tgr;

extern "C" int64_t tgr_bwd_max_entries(const tgr_call_t* calls, int n_calls) {
  int64_t n = 0;
  for (int c = 0; c < n_calls; ++c) {
    n += (int64_t)calls[c].T * calls[c].n_single;
    for (int a = 0; a < calls[c].n_arrays; ++a) n += calls[c].arr_nnz[a];
  }
  return n;
}

extern "C" size_t tgr_build_keys_workspace_bytes(int64_t max_entries) {
  const size_t nb = (size_t)((max_entries + kScanBlock - 1) / kScanBlock) + kMaxKeySegs + 1;
  return align_up(nb * sizeof(int32_t));
}

extern "C" int tgr_bwd_build_keys(const tgr_table_t* tables, int n_tables, const tgr_call_t* calls, int n_calls,
                                  uint32_t* keys, uint32_t* srcs, int32_t* n_valid_dev, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  tgr::TimedScope tgr_timed_("build_keys", stream);
  TGR_REQUIRE(tables && calls && keys && srcs && n_valid_dev && workspace, "null argument");
  TGR_REQUIRE(n_calls > 0 && n_calls <= TGR_MAX_CALLS, "n_calls=%d out of range", n_calls);
  TGR_REQUIRE(n_tables > 0 && n_tables <= TGR_MAX_TABLES, "n_tables=%d out of range", n_tables);
  cudaStream_t st = (cudaStream_t)stream;
  KeyBlockParams P{};
  KeyParams& kp = P.kp;
  int64_t total = 0;
  int ns = 0;
  for (int c = 0; c < n_calls; ++c) {
    const tgr_call_t& cl = calls[c];
    TGR_REQUIRE(cl.T >= 0 && (uint32_t)cl.T <= TGR_SRC_TOKEN_MASK, "call %d: T=%d does not fit the 24-bit token field", c, cl.T);
    TGR_REQUIRE(cl.n_slots <= TGR_MAX_SLOTS && cl.n_single <= TGR_MAX_SLOTS && cl.n_arrays <= TGR_MAX_ARRAYS, "call %d: too many slots", c);
    for (int i = 0; i < cl.n_slots; ++i) {
      const tgr_slot_t& s = cl.slots[i];
      if (s.kind == TGR_KIND_MM) continue;
      TGR_REQUIRE(s.table >= 0 && s.table < n_tables, "call %d slot %d: bad table", c, i);
      const tgr_table_t& tb = tables[s.table];
      TGR_REQUIRE(tb.key_base + tb.rows <= 0xFFFFFFFFll, "key space exceeds 32 bits");
      if (s.kind == TGR_KIND_SINGLE) {
        TGR_REQUIRE(s.src >= 0 && s.src < cl.n_single, "call %d slot %d: bad ids column", c, i);
        kp.col_key_base[c][s.src] = (uint32_t)tb.key_base;
        kp.col_rows[c][s.src] = (int32_t)tb.rows;
        kp.col_slot[c][s.src] = (uint8_t)i;
      }
    }
    if (cl.T > 0 && cl.n_single > 0) {
      TGR_REQUIRE(cl.ids != nullptr, "call %d: ids is NULL", c);
      KeySeg& g = kp.seg[ns++];
      g.vals = cl.ids; g.toks = nullptr; g.start = total; g.count = (int64_t)cl.T * cl.n_single;
      g.call = c; g.n_cols = cl.n_single; g.slot = 0; g.key_base = 0; g.rows = 0;
      total += g.count;
    }
    for (int i = 0; i < cl.n_slots; ++i) {
      const tgr_slot_t& s = cl.slots[i];
      if (s.kind != TGR_KIND_ARRAY) continue;
      TGR_REQUIRE(s.src >= 0 && s.src < cl.n_arrays, "call %d slot %d: bad array index", c, i);
      const int a = s.src;
      if (cl.arr_nnz[a] <= 0) continue;
      TGR_REQUIRE(cl.arr_val && cl.arr_tok[a], "call %d array %d: NULL pointers", c, a);
      KeySeg& g = kp.seg[ns++];
      g.vals = cl.arr_val + cl.arr_begin[a]; g.toks = cl.arr_tok[a]; g.start = total; g.count = cl.arr_nnz[a];
      g.call = c; g.n_cols = 0; g.slot = i; g.key_base = (uint32_t)tables[s.table].key_base; g.rows = (int32_t)tables[s.table].rows;
      total += g.count;
    }
  }
  kp.n_seg = ns;
  TGR_REQUIRE(workspace_bytes >= tgr_build_keys_workspace_bytes(total), "workspace too small");
  if (total == 0) {
    cudaMemsetAsync(n_valid_dev, 0, sizeof(int32_t), st);
    return check_launch("build_keys(empty)");
  }
  TGR_REQUIRE(total < (1ll << 31), "too many entries");
  int32_t* block_cnt = (int32_t*)workspace;
  int nb = 0;
  for (int i = 0; i < ns; ++i) {
    P.seg_first_block[i] = nb;
    nb += (int)((kp.seg[i].count + kKB - 1) / kKB);
  }
  P.seg_first_block[ns] = nb;
  TGR_K(keys_block_kernel<false>)<<<nb, kKT, 0, st>>>(P, block_cnt, nullptr, nullptr);
  TGR_K(block_scan_kernel)<<<1, kScanBlock, 0, st>>>(block_cnt, nb, n_valid_dev);
  TGR_K(keys_block_kernel<true>)<<<nb, kKT, 0, st>>>(P, block_cnt, keys, srcs);
  return check_launch("build_keys");
}

extern "C" size_t tgr_dedup_workspace_bytes(int64_t n) {
  return align_up((size_t)((n + kDdTile - 1) / kDdTile + 1) * sizeof(int32_t));
}

extern "C" int tgr_dedup(const uint32_t* keys_sorted, int64_t n, uint32_t* uniq, int32_t* seg_off, int32_t* seg_of_entry,
                         int32_t* n_unique_dev, void* workspace, size_t workspace_bytes, void* stream) {
  return tgr::dedup_dn(keys_sorted, n, uniq, seg_off, seg_of_entry, n_unique_dev, workspace, workspace_bytes, nullptr, stream);
}

int tgr::dedup_dn(const uint32_t* keys_sorted, int64_t n, uint32_t* uniq, int32_t* seg_off, int32_t* seg_of_entry,
                  int32_t* n_unique_dev, void* workspace, size_t workspace_bytes, const int32_t* n_dev, void* stream) {
  return dedup_remap_dn(keys_sorted, nullptr, n, uniq, seg_off, seg_of_entry, n_unique_dev, workspace, workspace_bytes, nullptr, 0,
                        nullptr, n_dev, stream);
}

// srcs_sorted / calls / ids_out != NULL: the emit pass also writes ids_out[call][token, column] = 1 + unique index for every
// SINGLE-slot entry (the caller has zeroed ids_out: padding ids stay 0)
int tgr::dedup_remap_dn(const uint32_t* keys_sorted, const uint32_t* srcs_sorted, int64_t n, uint32_t* uniq, int32_t* seg_off,
                        int32_t* seg_of_entry, int32_t* n_unique_dev, void* workspace, size_t workspace_bytes,
                        const tgr_call_t* calls, int n_calls, int32_t* const* ids_out, const int32_t* n_dev, void* stream) {
  tgr::TimedScope tgr_timed_("dedup", stream);
  TGR_REQUIRE(uniq && seg_off && n_unique_dev && workspace, "null argument");
  TGR_REQUIRE(n >= 0 && n < (1ll << 31), "n out of range");
  TGR_REQUIRE(workspace_bytes >= tgr_dedup_workspace_bytes(n), "workspace too small");
  const int vec = (((uintptr_t)keys_sorted | (uintptr_t)seg_of_entry | (uintptr_t)srcs_sorted) & 15) == 0;   // 128-bit accesses
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    cudaMemsetAsync(n_unique_dev, 0, sizeof(int32_t), st);
    cudaMemsetAsync(seg_off, 0, sizeof(int32_t), st);
    return check_launch("dedup(empty)");
  }
  TGR_REQUIRE(keys_sorted, "null keys");
  DedupScatter sc{};
  if (srcs_sorted != nullptr) {
    TGR_REQUIRE(calls && ids_out && n_calls > 0 && n_calls <= TGR_MAX_CALLS, "dedup+remap: bad calls");
    sc.srcs = srcs_sorted;
    for (int c = 0; c < n_calls; ++c) {
      TGR_REQUIRE(ids_out[c] != nullptr, "ids_out[%d] is NULL", c);
      sc.out[c] = ids_out[c];
      sc.n_cols[c] = calls[c].n_single;
      for (int i = 0; i < TGR_MAX_SLOTS; ++i) sc.col_of_slot[c][i] = -1;
      for (int i = 0; i < calls[c].n_slots; ++i)
        if (calls[c].slots[i].kind == TGR_KIND_SINGLE) sc.col_of_slot[c][i] = (int8_t)calls[c].slots[i].src;
    }
  }
  int32_t* block_cnt = (int32_t*)workspace;
  const int nb = (int)((n + kDdTile - 1) / kDdTile);
  TGR_K(dedup_count_kernel)<<<nb, kDdThreads, 0, st>>>(keys_sorted, n, block_cnt, n_dev, vec);
  TGR_K(block_scan_kernel)<<<1, kScanBlock, 0, st>>>(block_cnt, nb, n_unique_dev);
  TGR_K(dedup_emit_kernel)<<<nb, kDdThreads, 0, st>>>(keys_sorted, n, block_cnt, uniq, seg_off, seg_of_entry, sc, n_dev, vec);
  TGR_K(dedup_finish_kernel)<<<1, 1, 0, st>>>(seg_off, n_unique_dev, n, n_dev);
  return check_launch("dedup");
}

static int adam_rows_impl(const tgr_table_t* tables, int n_tables, int H, const uint32_t* uniq, const float* grads,
                          const int32_t* n_unique_dev, int64_t max_unique, const tgr_adam_t* adam, const tgr_adam_t* adam_dev,
                          void* stream);
// resident CTAs of the row update per SM: HBM-bound at four already (32 warps x 4 independent 16-byte loads per thread);
// the other half of the SM's thread slots stays free for the next step's key processing running beside it
static const int kAdamRowsBlocksPerSM = [] { const char* e = getenv("TGR_ADAM_BPS"); return e ? atoi(e) : 4; }();

extern "C" int tgr_adam_rows(const tgr_table_t* tables, int n_tables, int H, const uint32_t* uniq, const float* grads,
                             const int32_t* n_unique_dev, int64_t max_unique, const tgr_adam_t* adam, void* stream) {
  TGR_REQUIRE(adam, "null argument");
  return adam_rows_impl(tables, n_tables, H, uniq, grads, n_unique_dev, max_unique, adam, nullptr, stream);
}

extern "C" int tgr_adam_rows_dev(const tgr_table_t* tables, int n_tables, int H, const uint32_t* uniq, const float* grads,
                                 const int32_t* n_unique_dev, int64_t max_unique, const tgr_adam_t* adam_dev, void* stream) {
  TGR_REQUIRE(adam_dev, "null argument");
  return adam_rows_impl(tables, n_tables, H, uniq, grads, n_unique_dev, max_unique, nullptr, adam_dev, stream);
}

static int adam_rows_impl(const tgr_table_t* tables, int n_tables, int H, const uint32_t* uniq, const float* grads,
                          const int32_t* n_unique_dev, int64_t max_unique, const tgr_adam_t* adam, const tgr_adam_t* adam_dev,
                          void* stream) {
  tgr::TimedScope tgr_timed_("adam_rows", stream);
  TGR_REQUIRE(uniq && grads && n_unique_dev, "null argument");
  RowParams rp{};
  if (int rc = fill_row_params(rp, tables, n_tables, H)) return rc;
  for (int t = 0; t < n_tables; ++t) TGR_REQUIRE(rp.w[t] && rp.m[t] && rp.v[t], "table %d: weight/exp_avg/exp_avg_sq NULL", t);
  if (adam) rp.adam = *adam;
  rp.adam_dev = adam_dev;
  if (max_unique <= 0) return 0;
  int64_t blocks = (max_unique * rp.H4 + 255) / 256;
  if (blocks > kNumSMs * kAdamRowsBlocksPerSM) blocks = kNumSMs * kAdamRowsBlocksPerSM;
  TGR_K(rows_kernel<0>)<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rp, uniq, grads, n_unique_dev);
  return check_launch("adam_rows");
}

namespace tgr {
struct DenseParams {
  tgr_dense_list_t l;
  int32_t first_block[TGR_MAX_DENSE + 1];
  tgr_adam_t adam;
  const tgr_adam_t* adam_dev;
};
constexpr int kDenseChunk = 1024;   // elements per CTA (256 threads x 4)
__global__ void __launch_bounds__(256) adam_dense_kernel(const __grid_constant__ DenseParams p) {
  int t = 0;
  while (t + 1 < p.l.n && (int)blockIdx.x >= p.first_block[t + 1]) ++t;
  const tgr_adam_t ad = p.adam_dev ? *p.adam_dev : p.adam;
  const int64_t i0 = (int64_t)(blockIdx.x - p.first_block[t]) * kDenseChunk;
  float* w = p.l.w[t];
  float* m = p.l.m[t];
  float* v = p.l.v[t];
  const float* g = p.l.g[t];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t i = i0 + k * 256 + threadIdx.x;
    if (i < p.l.numel[t]) {
      float ww = w[i], mm = m[i], vv = v[i];
      adam_elem(ww, mm, vv, __ldg(g + i) * ad.grad_scale, ad);
      w[i] = ww; m[i] = mm; v[i] = vv;
    }
  }
}
}  // namespace tgr

extern "C" int tgr_adam_dense(const tgr_dense_list_t* list, const tgr_adam_t* adam, const tgr_adam_t* adam_dev, void* stream) {
  tgr::TimedScope tgr_timed_("adam_dense", stream);
  TGR_REQUIRE(list && list->n > 0 && list->n <= TGR_MAX_DENSE, "bad dense list");
  TGR_REQUIRE((adam != nullptr) != (adam_dev != nullptr), "exactly one of adam / adam_dev");
  DenseParams p{};
  p.l = *list;
  int blocks = 0;
  for (int t = 0; t < list->n; ++t) {
    TGR_REQUIRE(list->w[t] && list->g[t] && list->m[t] && list->v[t] && list->numel[t] > 0, "dense tensor %d: null / empty", t);
    p.first_block[t] = blocks;
    blocks += (int)((list->numel[t] + kDenseChunk - 1) / kDenseChunk);
  }
  p.first_block[list->n] = blocks;
  if (adam) p.adam = *adam;
  p.adam_dev = adam_dev;
  TGR_K(adam_dense_kernel)<<<blocks, 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("adam_dense");
}

extern "C" int tgr_scatter_rows(const tgr_table_t* tables, int n_tables, int H, const uint32_t* uniq, const float* grads,
                                const int32_t* n_unique_dev, int64_t max_unique, void* stream) {
  tgr::TimedScope tgr_timed_("scatter_rows", stream);
  TGR_REQUIRE(uniq && grads && n_unique_dev, "null argument");
  RowParams rp{};
  if (int rc = fill_row_params(rp, tables, n_tables, H)) return rc;
  if (max_unique <= 0) return 0;
  int64_t blocks = (max_unique * rp.H4 + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  TGR_K(rows_kernel<1>)<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rp, uniq, grads, n_unique_dev);
  return check_launch("scatter_rows");
}

extern "C" int tgr_gather_rows(const float* table, int H, const uint32_t* rows, const int32_t* n_dev, int64_t max_n,
                               float* out, void* stream) {
  tgr::TimedScope tgr_timed_("gather_rows", stream);
  TGR_REQUIRE(table && rows && n_dev && out, "null argument");
  TGR_REQUIRE(H > 0 && H % 4 == 0, "bad H");
  if (max_n <= 0) return 0;
  int64_t blocks = (max_n * (H / 4) + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  TGR_K(gather_rows_kernel)<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(table, H / 4, rows, n_dev, out);
  return check_launch("gather_rows");
}
