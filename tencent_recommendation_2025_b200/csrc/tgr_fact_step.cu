// Step-level driver of the factored path: one C-ABI call per phase instead of one per kernel.
//
// A training step of the reference issues ~70 feat2emb-side torch ops from Python (model/BaseLine/model.py:226-310 x3,
// autograd, main.py:188-190); the factored pipeline is ~45 small kernels whose launch sequence is fixed once the
// packed calls are known. Driving them one ctypes call at a time left the GPU idle behind the Python interpreter
// (1.85 ms of host time per 1.4 ms of kernels, profiles/README.md), so the sequencing lives here:
//
//   tgr_fact_prepare        carve the group's arena, keys -> sort -> dedup + id remap                 (value independent)
//   tgr_fact_mm_branch      (optional) mm fold + projection of every call on an internal side stream
//   tgr_fact_call_forward   [first call: project unique rows, fold mm weights] mm projection, gather-sum forward
//   tgr_fact_call_backward  relu mask + bias grads; with `finish`: mm chain rule (side stream), segmented reduce, row grads, dW
//
// With tgr_fact_group_t.n_is_capacity the sequence depends on the calls' SHAPES only (the lookup count is read from device
// memory), so a whole step can be captured in a CUDA graph (graphed.py).
//
// Every buffer is carved from ONE caller-provided arena (torch-allocated); nothing is allocated here and all work is
// enqueued on the caller's stream in a fixed order, so results are identical to the per-kernel entry points.
#include "tgr_common.cuh"
#include "tgr_rows.cuh"

namespace tgr {

struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* b) : base((char*)b) {}
  template <class T>
  T* take(size_t count) {
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += align_up(count * sizeof(T));
    return p;
  }
};

static bool call_has_user(const tgr_call_t& c) {
  for (int i = 0; i < c.n_slots; ++i)
    if (c.slots[i].side == TGR_SIDE_USER) return true;
  return false;
}

static int64_t call_nnz(const tgr_call_t& c) {
  int64_t z = 0;
  for (int a = 0; a < c.n_arrays; ++a) z += c.arr_nnz[a];
  return z;
}

// arr_u is indexed exactly like arr_val: extent = end of the last array
static int64_t call_arr_extent(const tgr_call_t& c) {
  int64_t z = 0;
  for (int a = 0; a < c.n_arrays; ++a) z = max(z, (int64_t)c.arr_begin[a] + c.arr_nnz[a]);
  return z;
}

// lays the group's buffers out over `arena` (or just measures when arena == nullptr)
static size_t carve(tgr_fact_group_t* g, void* arena, int n_tables) {
  Carver cv(arena);
  const int H = g->H;
  const size_t cap = (size_t)(g->n > 0 ? g->n : 1);
  g->cap = (int64_t)cap;
  g->keys_in = cv.take<uint32_t>(cap);
  g->srcs_in = cv.take<uint32_t>(cap);
  g->keys = cv.take<uint32_t>(cap);
  g->srcs = cv.take<uint32_t>(cap);
  g->uniq = cv.take<uint32_t>(cap);
  g->seg_off = cv.take<int32_t>(cap + 1);
  g->seg_of = cv.take<int32_t>(cap);
  g->n_unique = cv.take<int32_t>(4);
  g->n_valid = cv.take<int32_t>(4);
  g->P = cv.take<float>(cap * H);
  g->rows_local = cv.take<float>(cap * H);
  g->G = cv.take<float>(cap * H);
  size_t ws = 0;
  int64_t max_entries = 0;
  size_t max_T = 0;
  bool any_tc_bwd = false;
  for (int f = 0; f < g->n_mm; ++f) any_tc_bwd = any_tc_bwd || tgr_mm_proj_bwd_tc_supported(g->mm_x_dtype, g->mm_dim[f], H);
  for (int c = 0; c < g->n_calls; ++c) {
    const tgr_call_t& cl = g->calls[c];
    const size_t T = (size_t)cl.T;
    max_entries += (int64_t)T * cl.n_single + call_nnz(cl);
    g->ids_u[c] = cv.take<int32_t>(T * (size_t)cl.n_single);
    g->arr_u[c] = cv.take<int32_t>((size_t)call_arr_extent(cl));
    g->mask[c] = cv.take<uint8_t>(T * (H / 4));
    g->dz_item[c] = cv.take<float>(T * H);
    g->dz_user[c] = call_has_user(cl) ? cv.take<float>(T * H) : nullptr;
    for (int f = 0; f < g->n_mm; ++f) g->mmz[c][f] = cv.take<float>(T * H);
    ws = max(ws, tgr_fact_relu_mask_workspace_bytes((int64_t)T, H));
    for (int f = 0; f < g->n_mm; ++f) {
      if (tgr_mm_proj_bwd_tc_supported(g->mm_x_dtype, g->mm_dim[f], H))
        ws = max(ws, tgr_mm_proj_bwd_tc_workspace_bytes((int64_t)T, g->mm_dim[f], H));
      else
        ws = max(ws, tgr_mm_proj_bwd_workspace_bytes((int64_t)T, g->mm_dim[f], H));
    }
    max_T = max(max_T, T);
  }
  g->dzb = any_tc_bwd ? (void*)cv.take<uint16_t>(2 * ((max_T + 63) / 64 * 64) * H) : nullptr;   // two bf16 planes
  for (int f = 0; f < g->n_mm; ++f) {
    g->fold_M[f] = cv.take<float>((size_t)H * g->mm_dim[f]);
    g->fold_c[f] = cv.take<float>(H);
    g->mm_A[f] = cv.take<float>((size_t)H * g->mm_dim[f]);
    g->mm_s[f] = cv.take<float>(H);
    g->fold_Mb[f] = cv.take<uint16_t>((size_t)2 * H * g->mm_dim[f]);   // two bf16 planes (hi, lo)
  }
  ws = max(ws, tgr_build_keys_workspace_bytes(max_entries));
  ws = max(ws, tgr_sort_workspace_bytes(g->n));
  ws = max(ws, tgr_dedup_workspace_bytes(g->n));
  ws = max(ws, 2 * align_up((size_t)(g->n / 512 + 2) * H * sizeof(float)));   // reduce mode 0: per-CTA head/tail partials
  ws = max(ws, tgr_fact_backward_workspace_bytes(n_tables, H));
  g->ws = cv.take<char>(ws);
  g->ws_bytes = ws;
  return cv.off;
}

// side stream + fork / join events of the mm branch, one set per device (created on first use, i.e. in a warm-up step and
// never during a graph capture)
struct Branch {
  cudaStream_t side = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
static Branch* branch_of_device() {
  static Branch b[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (b[dev].side == nullptr) {
    if (cudaStreamCreateWithFlags(&b[dev].side, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    cudaEventCreateWithFlags(&b[dev].fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&b[dev].join, cudaEventDisableTiming);
  }
  return &b[dev];
}

}  // namespace tgr

using namespace tgr;

static int check_group(const tgr_fact_group_t* g) {
  TGR_REQUIRE(g != nullptr, "group is NULL");
  TGR_REQUIRE(g->n_calls > 0 && g->n_calls <= TGR_MAX_CALLS, "n_calls=%d out of range", g->n_calls);
  TGR_REQUIRE(g->H == 32 || g->H == 64 || g->H == 128, "the factored path supports H in {32, 64, 128} (H=%d)", g->H);
  TGR_REQUIRE(g->n >= 0 && g->n < (1ll << 31) && g->n * (g->H / 4) < (1ll << 31), "n out of range");
  TGR_REQUIRE(g->n_mm >= 0 && g->n_mm <= TGR_MAX_MM, "n_mm out of range");
  return 0;
}

extern "C" size_t tgr_fact_group_bytes(const tgr_fact_group_t* g, int n_tables) {
  if (check_group(g)) return 0;
  tgr_fact_group_t tmp = *g;
  return carve(&tmp, nullptr, n_tables) + 256;
}

extern "C" int tgr_fact_prepare(const tgr_table_t* tables, int n_tables, tgr_fact_group_t* g, void* arena,
                                size_t arena_bytes, void* stream) {
  if (int rc = check_group(g)) return rc;
  TGR_REQUIRE(tables && arena, "null argument");
  TGR_REQUIRE(((uintptr_t)arena & 255) == 0, "arena must be 256-byte aligned");
  const size_t need = carve(g, arena, n_tables);
  TGR_REQUIRE(arena_bytes >= need, "arena too small (%zu < %zu)", arena_bytes, need);
  g->projected = 0;
  g->n_backward = 0;
  g->mm_done = g->mm_joined = 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (g->n_mm > 0)
    if (Branch* br = branch_of_device()) cudaEventRecord(br->fork, st);   // where tgr_fact_mm_branch (if called) forks from
  for (int f = 0; f < g->n_mm; ++f) {   // A = dz^T x and s = colsum(dz) accumulate over the group's calls
    cudaMemsetAsync(g->mm_A[f], 0, (size_t)g->H * g->mm_dim[f] * sizeof(float), st);
    cudaMemsetAsync(g->mm_s[f], 0, (size_t)g->H * sizeof(float), st);
  }
  // remapped ids start from zero (padding): cleared FIRST — when the group is prepared on a branch next to the previous
  // step's row-gradient kernel these copy-engine nodes run at once instead of sitting on the tail of the key chain
  for (int c = 0; c < g->n_calls; ++c)
    cudaMemsetAsync(g->ids_u[c], 0, (size_t)g->calls[c].T * g->calls[c].n_single * sizeof(int32_t), st);
  if (int rc = tgr_bwd_build_keys(tables, n_tables, g->calls, g->n_calls, g->keys_in, g->srcs_in, g->n_valid, g->ws,
                                  g->ws_bytes, stream)) return rc;
  // n_is_capacity: g->n only bounds the count; the kernels read the count build_keys left in g->n_valid, so the launch
  // sequence depends on the calls' shapes alone (CUDA-graph capture, graphed.py)
  const int32_t* nd = g->n_is_capacity ? g->n_valid : nullptr;
  if (int rc = sort_pairs_dn(g->keys_in, g->srcs_in, g->keys, g->srcs, g->n, g->key_bits, g->ws, g->ws_bytes, nd, stream))
    return rc;
  // dedup + id remap of the SINGLE slots in one pass over the sorted pairs (ids_u were zeroed above: padding ids stay 0)
  int32_t* outs[TGR_MAX_CALLS];
  for (int c = 0; c < g->n_calls; ++c) outs[c] = g->ids_u[c];
  if (int rc = dedup_remap_dn(g->keys, g->srcs, g->n, g->uniq, g->seg_off, g->seg_of, g->n_unique, g->ws, g->ws_bytes, g->calls,
                              g->n_calls, outs, nd, stream)) return rc;
  // array values (a token may hold several): searching remap, they are few — one launch for all of them
  if (int rc = tgr_remap_arrays(tables, n_tables, g->calls, g->n_calls, g->uniq, g->n_unique, nullptr, g->arr_u, stream))
    return rc;
  return check_launch("fact_prepare");
}

static int mm_fold_all(const tgr_fact_params_t* prm, tgr_fact_group_t* g, void* stream) {
  const int H = g->H;
  for (int f = 0; f < g->n_mm; ++f) {
    const tgr_mm_feat_t& m = prm->mm[f];
    TGR_REQUIRE(m.mm_dim == g->mm_dim[f], "mm_dim mismatch");
    if (int rc = tgr_fact_mm_fold(prm->dnn.w_item + m.col, prm->dnn.item_ld, m.w, m.b, H, m.mm_dim, g->fold_M[f],
                                  g->fold_c[f], stream)) return rc;
    if (tgr_mm_proj_fwd_tc_supported(g->mm_x_dtype, m.mm_dim, H))   // wide bf16 feature: tcgen05 path wants a bf16 M
      if (int rc = tgr_split_bf16(g->fold_M[f], (int64_t)H * m.mm_dim, g->fold_Mb[f],
                                  (uint16_t*)g->fold_Mb[f] + (size_t)H * m.mm_dim, stream)) return rc;
  }
  return 0;
}

static int mm_project_call(tgr_fact_group_t* g, int c, void* stream) {
  const int H = g->H;
  const tgr_call_t& cl = g->calls[c];
  for (int f = 0; f < g->n_mm; ++f) {
    TGR_REQUIRE(g->mm_x[c][f] != nullptr, "mm input %d of call %d is NULL", f, c);
    if (tgr_mm_proj_fwd_tc_supported(g->mm_x_dtype, g->mm_dim[f], H)) {
      if (int rc = tgr_mm_proj_fwd_tc(g->mm_x[c][f], cl.T, g->mm_dim[f], g->fold_Mb[f], 2, g->fold_c[f], H, g->mmz[c][f], H,
                                      TGR_DTYPE_F32, stream)) return rc;
      continue;
    }
    if (int rc = tgr_mm_proj_fwd(g->mm_x[c][f], g->mm_x_dtype, cl.T, g->mm_dim[f], g->fold_M[f], g->fold_c[f], H,
                                 g->mmz[c][f], H, TGR_DTYPE_F32, stream)) return rc;
  }
  return 0;
}

extern "C" int tgr_fact_mm_branch(const tgr_fact_params_t* prm, tgr_fact_group_t* g, int fork_now, void* stream) {
  if (int rc = check_group(g)) return rc;
  TGR_REQUIRE(prm != nullptr && prm->n_mm == g->n_mm, "bad params");
  if (g->n_mm == 0 || g->mm_done) return 0;
  TGR_REQUIRE(!g->projected, "tgr_fact_mm_branch must precede the group's first forward");
  Branch* br = branch_of_device();
  TGR_REQUIRE(br != nullptr, "could not create the side stream");
  if (fork_now) cudaEventRecord(br->fork, (cudaStream_t)stream);
  cudaStreamWaitEvent(br->side, br->fork, 0);          // (else) recorded by tgr_fact_prepare on the caller's stream
  if (int rc = mm_fold_all(prm, g, br->side)) return rc;
  for (int c = 0; c < g->n_calls; ++c)
    if (int rc = mm_project_call(g, c, br->side)) return rc;
  cudaEventRecord(br->join, br->side);
  g->mm_done = 1;
  return check_launch("fact_mm_branch");
}

extern "C" int tgr_fact_call_forward(const tgr_table_t* tables, int n_tables, const tgr_fact_params_t* prm,
                                     tgr_fact_group_t* g, int c, float* out, void* stream) {
  if (int rc = check_group(g)) return rc;
  TGR_REQUIRE(tables && prm && out, "null argument");
  TGR_REQUIRE(c >= 0 && c < g->n_calls, "call index out of range");
  TGR_REQUIRE(prm->n_mm == g->n_mm, "n_mm mismatch");
  const int H = g->H;
  if (!g->projected) {
    tgr_row_source_t src = g->src;
    if (src.n_peers > 0) src.save_rows = g->rows_local;   // read the owners' shards once; the backward uses the copy
    if (int rc = tgr_fact_project_rows(tables, n_tables, H, &prm->dnn, g->uniq, g->n_unique, g->n, &src, g->P, stream)) return rc;
    if (!g->mm_done)
      if (int rc = mm_fold_all(prm, g, stream)) return rc;
    g->projected = 1;
  }
  tgr_call_t& cl = g->calls[c];
  if (g->mm_done) {
    if (!g->mm_joined) {
      Branch* br = branch_of_device();
      TGR_REQUIRE(br != nullptr, "side stream missing");
      cudaStreamWaitEvent((cudaStream_t)stream, br->join, 0);
      g->mm_joined = 1;
    }
  } else if (int rc = mm_project_call(g, c, stream)) {
    return rc;
  }
  return tgr_fact_forward(&cl, H, g->ids_u[c], call_arr_extent(cl) ? g->arr_u[c] : nullptr, g->P, (const float* const*)g->mmz[c], g->n_mm, prm->b_item,
                          call_has_user(cl) ? prm->b_user : nullptr, out, g->mask[c], stream);
}

extern "C" int tgr_fact_call_backward(const tgr_table_t* tables, int n_tables, const tgr_fact_params_t* prm,
                                      tgr_fact_group_t* g, int c, const float* d_out, const tgr_fact_grads_t* gr,
                                      int finish, void* stream) {
  if (int rc = check_group(g)) return rc;
  TGR_REQUIRE(tables && prm && gr, "null argument");
  const int H = g->H;
  if (c >= 0) {
    TGR_REQUIRE(c < g->n_calls && d_out, "bad call index / d_out");
    const tgr_call_t& cl = g->calls[c];
    const bool user = call_has_user(cl);
    // one 32-wide mm feature rides along in the dZ pass (A = dz^T x, s = colsum); the others use the generic kernels
    int fused = -1;
    for (int f = 0; f < g->n_mm; ++f)
      if (prm->mm[f].mm_dim == 32 && gr->dW_mm[f] != nullptr) { fused = f; break; }
    if (int rc = tgr_fact_relu_mask(d_out, g->mask[c], cl.T, H, g->dz_item[c], user ? g->dz_user[c] : nullptr, gr->db_item,
                                    user ? gr->db_user : nullptr, fused >= 0 ? g->mm_x[c][fused] : nullptr, g->mm_x_dtype,
                                    fused >= 0 ? 32 : 0, fused >= 0 ? g->mm_A[fused] : nullptr,
                                    fused >= 0 ? g->mm_s[fused] : nullptr, g->ws, g->ws_bytes, stream)) return rc;
    bool dz_cast = false;
    for (int f = 0; f < g->n_mm; ++f) {
      if (gr->dW_mm[f] == nullptr || f == fused) continue;
      if (tgr_mm_proj_bwd_tc_supported(g->mm_x_dtype, prm->mm[f].mm_dim, H)) {
        // wide bf16 feature: A += dz^T x on the tensor cores (s = colsum(dz) is read off db_item at the finish)
        const int64_t plane_rows = ((int64_t)cl.T + 63) / 64 * 64;
        if (!dz_cast) {   // dz as two bf16 planes (16 mantissa bits); the pad rows of plane 0 must be finite
          uint16_t* hi = (uint16_t*)g->dzb;
          if (plane_rows > cl.T) cudaMemsetAsync(hi + (size_t)cl.T * H, 0, (size_t)(plane_rows - cl.T) * H * 2, (cudaStream_t)stream);
          if (int rc = tgr_split_bf16(g->dz_item[c], (int64_t)cl.T * H, hi, hi + (size_t)plane_rows * H, stream)) return rc;
          dz_cast = true;
        }
        if (int rc = tgr_mm_proj_bwd_tc(g->mm_x[c][f], cl.T, prm->mm[f].mm_dim, g->dzb, 2, plane_rows, H, g->mm_A[f], 1, g->ws,
                                        g->ws_bytes, stream)) return rc;
        continue;
      }
      if (int rc = tgr_mm_proj_bwd(g->mm_x[c][f], g->mm_x_dtype, cl.T, prm->mm[f].mm_dim, g->dz_item[c], H, TGR_DTYPE_F32,
                                   H, g->mm_A[f], g->mm_s[f], 1, g->ws, g->ws_bytes, stream)) return rc;
    }
    g->n_backward++;
  }
  if (!finish) return 0;
  // the chain rule through emb_transform / the item-DNN block is linear in (A, s): ONCE per group on the sums. Nothing in
  // the reduce / row-gradient kernels below depends on it (it adds into the mm features' own columns of dW_item), so it
  // runs on the side stream next to them and is joined at the end of this call.
  Branch* br = (g->n_mm > 0 && g->n > 0) ? branch_of_device() : nullptr;
  void* chain_stream = stream;
  if (br != nullptr) {
    cudaEventRecord(br->fork, (cudaStream_t)stream);
    cudaStreamWaitEvent(br->side, br->fork, 0);
    chain_stream = br->side;
  }
  for (int f = 0; f < g->n_mm; ++f) {
    const tgr_mm_feat_t& m = prm->mm[f];
    if (gr->dW_mm[f] == nullptr) continue;
    TGR_REQUIRE(gr->dW_item != nullptr, "dW_item is NULL");
    // tensor-core features: s = sum over the group's calls of colsum(dz_item) is exactly what the (zero-initialised)
    // db_item accumulator holds by now
    const bool tcb = tgr_mm_proj_bwd_tc_supported(g->mm_x_dtype, m.mm_dim, H) &&
                     !(m.mm_dim == 32);
    TGR_REQUIRE(!tcb || gr->db_item != nullptr, "the tensor-core mm backward reads colsum(dz) from db_item: it must not be NULL");
    if (int rc = tgr_fact_mm_chain_bwd(prm->dnn.w_item + m.col, prm->dnn.item_ld, m.w, m.b, g->mm_A[f], tcb ? gr->db_item : g->mm_s[f], H,
                                       m.mm_dim, gr->dW_mm[f], gr->db_mm[f], gr->dW_item + m.col, prm->dnn.item_ld,
                                       chain_stream)) return rc;
  }
  if (br != nullptr) cudaEventRecord(br->join, br->side);
  if (g->n == 0) return 0;
  // "concat gradient" of the reduction = dZ [T, H]: every slot at column 0 of its side, row pitch H
  tgr_call_t calls[TGR_MAX_CALLS];
  for (int i = 0; i < g->n_calls; ++i) {
    calls[i] = g->calls[i];
    for (int s = 0; s < calls[i].n_slots; ++s) calls[i].slots[s].col = 0;
    calls[i].item_cat = g->dz_item[i];
    calls[i].item_ld = H;
    calls[i].user_cat = g->dz_user[i];
    calls[i].user_ld = H;
    calls[i].cat_dtype = TGR_DTYPE_F32;
  }
  if (int rc = bwd_reduce_dn(tables, n_tables, H, calls, g->n_calls, g->keys, g->srcs, g->n, 0, g->seg_of, g->G, nullptr,
                             g->ws, g->ws_bytes, g->n_is_capacity ? g->n_valid : nullptr, stream)) return rc;
  if (g->reduce_done_event != nullptr) cudaEventRecord((cudaEvent_t)g->reduce_done_event, (cudaStream_t)stream);
  tgr_row_source_t src = g->src;
  if (src.n_peers > 0) {   // the rows were copied out of the peers' shards by the forward projection
    src = tgr_row_source_t{};
    src.fetched_rows = g->rows_local;
  }
  const int rc = tgr_fact_unique_backward(tables, n_tables, H, &prm->dnn, g->uniq, g->n_unique, g->n, &src, g->G, gr->dW_item,
                                          gr->dW_user, g->ws, g->ws_bytes, stream);
  if (br != nullptr) cudaStreamWaitEvent((cudaStream_t)stream, br->join, 0);
  return rc;
}
