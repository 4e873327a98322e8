// Device-resident item features (second half): in the reference every item-side sparse feature and every
// frozen mm vector is a FUNCTION OF THE ITEM ID — the dataset reads them from item_feat_dict[str(id)] and the mm store by
// creative id (model/BaseLine/dataset.py:159,260-263) and re-materialises them per token on the host. Keeping the
// [items + 1, n_feat] int32 feature table and the [items + 1, mm_dim] mm tables in HBM turns a step's host feed into the id
// tensors plus the few user tokens (~3 MB instead of 62 MB at B = 1024) and removes the item side of the dict walk.
// These kernels expand the ids on the device into exactly the packed call the host tensorizer would have produced.
#include "tgr_common.cuh"

namespace tgr {

// ids[t, :] = 0 except ids[t, id_col] = item_ids[t] and ids[t, col0 + j] = feat[item_ids[t], j].
// CTA = kExpTok tokens: thread t reads its id and the item's feature row (all loads of a thread independent, 8-byte pieces when
// the row allows), stages the packed row in shared memory, and the CTA copies the [256, n_single] block out linearly (it is
// contiguous in `ids`), 16 bytes per lane. The first version (one thread per element: id load -> dependent feature load ->
// store) took 10 us per call at T = 103 k.
constexpr int kExpTok = 64;    // 5 KB of staging per CTA: small enough to run beside the 218 KB row-gradient kernel (PipelinedStep branch)
__global__ void __launch_bounds__(kExpTok) expand_items_kernel(const int32_t* __restrict__ item_ids, int64_t T, int n_single, int id_col,
                                                               int col0, const int32_t* __restrict__ feat, int n_feat, int64_t n_items,
                                                               int32_t* __restrict__ ids) {
  extern __shared__ __align__(16) int32_t s_rows[];   // [kExpTok][n_single]
  const int64_t t0 = (int64_t)blockIdx.x * kExpTok;
  const int nt = (int)min((int64_t)kExpTok, T - t0);
  const int tid = threadIdx.x;
  if (tid < nt) {
    const int raw = __ldg(item_ids + t0 + tid);
    const int id = (raw < 0 || raw >= n_items) ? 0 : raw;   // out-of-range ids read the padding row here; the range check proper is the kernels'
    int32_t* row = s_rows + tid * n_single;
    for (int c = 0; c < col0; ++c) row[c] = 0;
    for (int c = col0 + n_feat; c < n_single; ++c) row[c] = 0;
    const int32_t* src = feat + (size_t)id * n_feat;
    if ((n_feat & 1) == 0) {                                  // rows are 8-byte aligned
      for (int j = 0; j < n_feat; j += 2) {
        const int2 v = __ldg(reinterpret_cast<const int2*>(src + j));
        row[col0 + j] = v.x;
        row[col0 + j + 1] = v.y;
      }
    } else {
      for (int j = 0; j < n_feat; ++j) row[col0 + j] = __ldg(src + j);
    }
    row[id_col] = raw;
  }
  __syncthreads();
  const int total = nt * n_single;
  int32_t* dst = ids + t0 * n_single;                         // 256 * n_single * 4 bytes per CTA: 16-byte aligned
  const int total4 = total >> 2;
  for (int i = tid; i < total4; i += kExpTok) reinterpret_cast<int4*>(dst)[i] = reinterpret_cast<const int4*>(s_rows)[i];
  for (int i = (total4 << 2) + tid; i < total; i += kExpTok) dst[i] = s_rows[i];
}

// ids[tok[u], col0 + j] = vals[u, j] for the few user tokens of a call (one per sequence, dataset.py:119)
__global__ void __launch_bounds__(256) scatter_user_kernel(const int32_t* __restrict__ tok, const int32_t* __restrict__ vals, int n_tok,
                                                           int n_cols, int n_single, int col0, int32_t* __restrict__ ids) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n_tok * n_cols) return;
  const int u = i / n_cols, j = i - u * n_cols;
  const int t = __ldg(tok + u);
  if (t < 0) return;                       // padding entry of a fixed-shape call
  ids[(size_t)t * n_single + col0 + j] = __ldg(vals + i);
}

// out[t, :] = table[item_ids[t], :] (16-byte pieces; rows are mm_dim * esz bytes, a multiple of 16); four pieces in flight
// per thread
__global__ void __launch_bounds__(256) gather_mm_kernel(const int32_t* __restrict__ item_ids, int64_t T, const uint4* __restrict__ table,
                                                        int row16, int64_t n_items, uint4* __restrict__ out) {
  const int64_t total = T * row16;
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t i0 = (int64_t)blockIdx.x * 256 + threadIdx.x; i0 < total; i0 += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < total) {
        const int64_t t = i / row16;
        const int c = (int)(i - t * row16);
        int id = __ldg(item_ids + t);
        if (id < 0 || id >= n_items) id = 0;
        v[u] = __ldg(table + (size_t)id * row16 + c);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < total) out[i] = v[u];
    }
  }
}

}  // namespace tgr

using namespace tgr;

extern "C" int tgr_expand_item_features(const int32_t* item_ids, int64_t T, int n_single, int id_col, int col0, const int32_t* feat_table,
                                        int n_feat, int64_t n_items, int32_t* ids_out, void* stream) {
  tgr::TimedScope tgr_timed_("expand_item_features", stream);
  TGR_REQUIRE(T >= 0 && n_single > 0 && id_col >= 0 && id_col < n_single && n_feat >= 0 && col0 >= 0 && col0 + n_feat <= n_single,
              "bad column layout");
  if (T == 0) return 0;
  TGR_REQUIRE(item_ids && ids_out && (n_feat == 0 || feat_table) && n_items > 0, "null argument");
  TGR_REQUIRE(((uintptr_t)ids_out & 15) == 0 && ((uintptr_t)feat_table & 7) == 0, "misaligned buffers");
  TGR_REQUIRE(n_single <= 48, "n_single too large for the staging tile");
  const int64_t blocks = (T + kExpTok - 1) / kExpTok;
  TGR_K(expand_items_kernel)<<<(unsigned)blocks, kExpTok, (size_t)kExpTok * n_single * sizeof(int32_t), (cudaStream_t)stream>>>(
      item_ids, T, n_single, id_col, col0, feat_table, n_feat, n_items, ids_out);
  return check_launch("expand_item_features");
}

extern "C" int tgr_scatter_user_tokens(const int32_t* tok, const int32_t* vals, int n_tok, int n_cols, int n_single, int col0,
                                       int32_t* ids, void* stream) {
  tgr::TimedScope tgr_timed_("scatter_user_tokens", stream);
  TGR_REQUIRE(n_tok >= 0 && n_cols > 0 && col0 >= 0 && col0 + n_cols <= n_single, "bad column layout");
  if (n_tok == 0) return 0;
  TGR_REQUIRE(tok && vals && ids, "null argument");
  TGR_K(scatter_user_kernel)<<<(n_tok * n_cols + 255) / 256, 256, 0, (cudaStream_t)stream>>>(tok, vals, n_tok, n_cols, n_single, col0, ids);
  return check_launch("scatter_user_tokens");
}

extern "C" int tgr_gather_mm_rows(const int32_t* item_ids, int64_t T, const void* table, int dtype, int mm_dim, int64_t n_items, void* out,
                                  void* stream) {
  tgr::TimedScope tgr_timed_("gather_mm_rows", stream);
  const int esz = dtype == TGR_DTYPE_BF16 ? 2 : 4;
  TGR_REQUIRE(T >= 0 && mm_dim > 0 && (mm_dim * esz) % 16 == 0, "mm rows must be a multiple of 16 bytes");
  if (T == 0) return 0;
  TGR_REQUIRE(item_ids && table && out && n_items > 0, "null argument");
  TGR_REQUIRE(((uintptr_t)table & 15) == 0 && ((uintptr_t)out & 15) == 0, "misaligned buffers");
  const int row16 = mm_dim * esz / 16;
  int64_t blocks = (T * row16 + 1023) / 1024;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  TGR_K(gather_mm_kernel)<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(item_ids, T, (const uint4*)table, row16, n_items, (uint4*)out);
  return check_launch("gather_mm_rows");
}
