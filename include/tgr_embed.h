/* tgr_embed.h — C ABI of the B200 (sm_100a) sparse-feature embedding path for the TencentGR
 * baseline models (Puiching-Memory/Tencent_Recommendation_2025).
 *
 * The reference has NO FFI / plugin interface: its boundary is the Python method
 *     BaselineModel.feat2emb(seq, feature_array, mask=None, include_user=False)
 * (model/BaseLine/model.py:226-310, model/BaseLineO1/model.py:327-416) plus what autograd and
 * torch.optim.AdamW do underneath it (model/BaseLine/main.py:131,188-190). This header is the
 * boundary the build introduces UNDER that method; each entry cites the reference lines whose
 * work it replaces. Host code (tencent_recommendation_2025_b200/engine.py) binds it with ctypes;
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain C types only; every pointer named *_dev / inside the structs is a DEVICE pointer
 *     borrowed for the duration of the call (torch owns all memory; the library allocates nothing);
 *   - every entry takes the CUDA stream as `void* stream` (cudaStream_t), is asynchronous, does no
 *     host synchronisation and keeps no global mutable state (re-entrant across streams);
 *   - return 0 on success, negative on error; tgr_last_error() gives the thread's last message;
 *   - ids are int32, id 0 is the padding row of every table (nn.Embedding(padding_idx=0),
 *     model.py:115-116,158-165); tables are row-major fp32 [rows, H], H % 4 == 0.
 */
#ifndef TGR_EMBED_H_
#define TGR_EMBED_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TGR_ABI_VERSION 1

#define TGR_MAX_TABLES 64
#define TGR_MAX_SLOTS 32
#define TGR_MAX_ARRAYS 8
#define TGR_MAX_CALLS 4
#define TGR_MAX_PEERS 16

/* slot kinds */
#define TGR_KIND_SINGLE 0 /* one id per token: copy one table row            (model.py:242-247,275) */
#define TGR_KIND_ARRAY 1  /* ragged id list: rows summed left to right       (model.py:277)         */
#define TGR_KIND_MM 2     /* dense mm vector: projected by tgr_mm_proj_*     (model.py:281-299)     */

#define TGR_SIDE_ITEM 0
#define TGR_SIDE_USER 1

#define TGR_DTYPE_F32 0
#define TGR_DTYPE_BF16 1

/* source code of one gradient contribution: call << 29 | slot << 24 | token */
#define TGR_SRC_CALL_SHIFT 29
#define TGR_SRC_SLOT_SHIFT 24
#define TGR_SRC_TOKEN_MASK 0x00FFFFFFu

/* One embedding table = one nn.Embedding(rows, H, padding_idx=0) (model.py:115-116,158-165).
 * Global key of a row = key_base + id; key bases are cumulative row counts in table order. */
typedef struct tgr_table {
  float* weight;     /* [rows, H] */
  float* exp_avg;    /* Adam m, [rows, H]; NULL unless a row update is requested */
  float* exp_avg_sq; /* Adam v, [rows, H] */
  float* grad;       /* dense [rows, H] gradient target for tgr_scatter_rows (parity mode), else NULL */
  int64_t rows;
  int64_t key_base;
} tgr_table_t;

/* One H-wide column block of a concat buffer (concat order: model.py:244-245,252-263,281-299). */
typedef struct tgr_slot {
  int32_t kind;  /* TGR_KIND_* */
  int32_t side;  /* TGR_SIDE_* */
  int32_t col;   /* first column (elements) in that side's concat buffer */
  int32_t table; /* index into the table array (SINGLE/ARRAY) */
  int32_t src;   /* SINGLE: column of ids; ARRAY: index into arr_*; MM: unused */
} tgr_slot_t;

/* One feat2emb call in packed form (replaces the 22/14 feat2tensor tensors of model.py:186-224,272). */
typedef struct tgr_call {
  int32_t T;       /* tokens = B*L */
  int32_t n_slots; /* entries of slots[] */
  int32_t n_single; /* columns of ids */
  int32_t n_arrays;
  tgr_slot_t slots[TGR_MAX_SLOTS];
  const int32_t* ids;                     /* [T, n_single] token-major; already type-masked (model.py:240-243) */
  const int32_t* arr_off[TGR_MAX_ARRAYS]; /* [T+1] CSR offsets into arr_val (absolute) */
  const int32_t* arr_tok[TGR_MAX_ARRAYS]; /* [nnz_a] token of each value (COO row), used by the backward */
  int32_t arr_begin[TGR_MAX_ARRAYS];      /* first value of array a inside arr_val */
  int32_t arr_nnz[TGR_MAX_ARRAYS];
  const int32_t* arr_val;                 /* concatenated values of all arrays, padding id 0 dropped */
  void* item_cat;                         /* fwd: OUT concat [T, item_ld]; bwd: IN d(concat) */
  void* user_cat;                         /* NULL when the call has no user side */
  int64_t item_ld;                        /* row pitch in elements */
  int64_t user_ld;
  int32_t cat_dtype;                      /* TGR_DTYPE_* of item_cat / user_cat */
  int32_t reserved;
  int32_t* err_flag;                      /* optional device int32: set to 1 + slot index on an out-of-range id
                                             (the reference raises IndexError / device-asserts); the offending
                                             lookup reads row 0 instead of faulting. NULL = no report. */
} tgr_call_t;

/* AdamW hyper-parameters of the row update (model/BaseLine/main.py:131; torch/optim/adam.py:416-419,457,476,531-547).
 * step counts from 1; bias corrections are formed on the host in double exactly as torch does. */
typedef struct tgr_adam {
  float lr, beta1, beta2, eps, weight_decay;
  float step_size;       /* lr / (1 - beta1^step) */
  float bc2_sqrt;        /* sqrt(1 - beta2^step) */
  float grad_scale;      /* multiplied into the reduced gradient before the update (1/loss_scale; 1 otherwise) */
  float decay;           /* 1 - lr*weight_decay, formed in double then rounded (param.mul_) */
  float one_minus_beta1; /* float(1 - beta1) in double (lerp_ weight) */
  float one_minus_beta2; /* float(1 - beta2) in double (addcmul_ value) */
  float reserved;
} tgr_adam_t;

int tgr_abi_version(void);
const char* tgr_last_error(void);

/* Instrumentation for bench.py / profiling (off by default; the one piece of process-wide state in the library):
 * after tgr_timing_enable(1) every kernel-launching entry brackets its launches with a CUDA-event pair on the
 * caller's stream; tgr_timing_collect synchronises those events and returns, per entry name ('\n'-joined in
 * `names`), the summed milliseconds and the number of calls. Returns the number of distinct names. */
int64_t tgr_launch_count(void); /* kernels launched by this library in this process so far (a count, monotone) */
int tgr_timing_enable(int on);
int tgr_timing_collect(char* names, size_t names_bytes, float* ms, int32_t* counts, int max_entries);

/* ---- forward ---------------------------------------------------------------------------------
 * Fused multi-table gather + array sum-pool + concat write: every SINGLE/ARRAY slot of the call in
 * one launch, straight into item_cat/user_cat (fp32 or bf16 RNE). Replaces aten::embedding x(15|24),
 * sum(2) x4 and cat x(1|2) of model.py:240-247,267-279,302,305 (SURVEY.md §2.2 K1,K3,K4,K6). */
int tgr_fwd_gather_pool_concat(const tgr_table_t* tables, int n_tables, int H, const tgr_call_t* call, void* stream);

/* mm projection forward: out[t, 0:H] = x[t, :] . W^T + b, written at `out` (already offset to the slot's
 * column) with row pitch out_ld. Replaces emb_transform[k](x) + its cat copy (model.py:297,299,302). */
int tgr_mm_proj_fwd(const void* x, int x_dtype, int64_t T, int mm_dim, const float* W, const float* bias, int H,
                    void* out, int64_t out_ld, int out_dtype, void* stream);

/* The same projection on the 5th-generation tensor cores (tcgen05.mma kind::f16 + TMEM accumulator, x and W tiles
 * streamed by TMA with 128-byte swizzle; csrc/tgr_mm_tc.cu) for the wide frozen mm features kept in bf16 (BASELINE.json
 * config 3: '82' = 1024-d ... '84' = 4096-d, model.py:183): x bf16 [T, mm_dim], w_bf16 = bf16 copy of W [H, mm_dim]
 * (tgr_cast_bf16; w_planes = 1) or TWO bf16 planes stacked along the rows, [2 H, mm_dim] = hi then bf16(W - hi)
 * (tgr_split_bf16; w_planes = 2: 16 mantissa bits, the result matches fp32 math on the bf16-stored features), fp32
 * accumulate, bias fp32 or NULL. Needs mm_dim % 64 == 0, mm_dim >= 128, H in {32, 64, 128} (tgr_mm_proj_fwd_tc_supported). */
int tgr_mm_proj_fwd_tc_supported(int x_dtype, int mm_dim, int H);
int tgr_mm_proj_fwd_tc(const void* x_bf16, int64_t T, int mm_dim, const void* w_bf16, int w_planes, const float* bias, int H,
                       void* out, int64_t out_ld, int out_dtype, void* stream);
/* Backward of the wide bf16 projection on the tensor cores: dW[H, mm_dim] (+)= dz^T . x with dz_bf16 a bf16 copy of the
 * fp32 dz [T, H] (tgr_cast_bf16). Split-K over tokens, both operands MN-major straight from TMA (csrc/tgr_mm_tc.cu), chunk
 * partials reduced in fixed order (bitwise reproducible). Needs mm_dim % 64 == 0, mm_dim >= 128, H in {64, 128}. */
int tgr_mm_proj_bwd_tc_supported(int x_dtype, int mm_dim, int H);
size_t tgr_mm_proj_bwd_tc_workspace_bytes(int64_t T, int mm_dim, int H);
int tgr_mm_proj_bwd_tc(const void* x_bf16, int64_t T, int mm_dim, const void* dz_bf16, int dz_planes, int64_t plane_rows, int H,
                       float* dW, int accumulate, void* workspace, size_t workspace_bytes, void* stream);
/* hi[i] = bf16(src[i]), lo[i] = bf16(src[i] - hi[i]): the two-plane operand form of the tensor-core projection. With
 * dz_planes = 2 the planes are [plane_rows, H] blocks of one buffer (plane_rows >= T, a multiple of 64; rows [T, plane_rows)
 * of plane 0 must be finite, e.g. zero). */
int tgr_split_bf16(const float* src, int64_t n, void* hi_bf16, void* lo_bf16, void* stream);
/* dst_bf16[i] = bf16(src[i]) (round to nearest even), n elements; src 16-byte, dst 8-byte aligned. */
int tgr_cast_bf16(const float* src, int64_t n, void* dst_bf16, void* stream);

/* mm projection backward: dW[H, mm_dim] (+)= dY^T . x ; db[H] (+)= sum_t dY. dY is read in place from the
 * concat gradient (pointer already offset to the slot's column). Deterministic split-T reduction.
 * workspace: tgr_mm_proj_bwd_workspace_bytes(). (autograd of model.py:297) */
size_t tgr_mm_proj_bwd_workspace_bytes(int64_t T, int mm_dim, int H);
int tgr_mm_proj_bwd(const void* x, int x_dtype, int64_t T, int mm_dim, const void* dy, int64_t dy_ld, int dy_dtype,
                    int H, float* dW, float* db, int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* ---- backward --------------------------------------------------------------------------------
 * Replaces embedding_dense_backward x54 + dense AdamW over every table row (SURVEY.md §2.2 K8,K10).
 * Pipeline: build_keys -> sort_pairs -> [dedup] -> segment_reduce(+adam) . All stages deterministic. */

/* Upper bound of gradient contributions of `n_calls` calls (all id slots incl. padding + all array values). */
int64_t tgr_bwd_max_entries(const tgr_call_t* calls, int n_calls);

/* Emit (global key, source code) for every non-padding id, compacted, in (call, token, slot) order.
 * n_valid_dev (int32, device) receives the count. keys/srcs must hold tgr_bwd_max_entries() items.
 * workspace >= tgr_build_keys_workspace_bytes(max_entries). */
size_t tgr_build_keys_workspace_bytes(int64_t max_entries);
int tgr_bwd_build_keys(const tgr_table_t* tables, int n_tables, const tgr_call_t* calls, int n_calls,
                       uint32_t* keys, uint32_t* srcs, int32_t* n_valid_dev, void* workspace, size_t workspace_bytes,
                       void* stream);

/* Stable LSD radix sort of (key, src) pairs on key bits [0, key_bits). n is the host-known entry count. */
size_t tgr_sort_workspace_bytes(int64_t n);
int tgr_sort_pairs(const uint32_t* keys_in, const uint32_t* srcs_in, uint32_t* keys_out, uint32_t* srcs_out, int64_t n,
                   int key_bits, void* workspace, size_t workspace_bytes, void* stream);

/* Run-length encode sorted keys: unique keys, segment offsets [U+1] (counts = diff), per-entry segment index
 * (optional, [n]) and U to n_unique_dev. Bit-exact with torch.unique(sorted=True, return_inverse/return_counts). */
size_t tgr_dedup_workspace_bytes(int64_t n);
int tgr_dedup(const uint32_t* keys_sorted, int64_t n, uint32_t* uniq, int32_t* seg_off, int32_t* seg_of_entry,
              int32_t* n_unique_dev, void* workspace, size_t workspace_bytes, void* stream);

/* Segmented reduction of the concat-gradient rows over the sorted pairs; fixed tiling => bitwise reproducible.
 *   mode 0: write the reduced row of the u-th unique key to grads_out[u, 0:H] (u from seg_of_entry, tgr_dedup)
 *   mode 1: fused AdamW row update in place on tables[].weight/exp_avg/exp_avg_sq (no grads_out, no dedup)
 * calls[].item_cat/user_cat are the concat GRADIENTS here. */
size_t tgr_reduce_workspace_bytes(int64_t n, int H);
int tgr_bwd_reduce(const tgr_table_t* tables, int n_tables, int H, const tgr_call_t* calls, int n_calls,
                   const uint32_t* keys_sorted, const uint32_t* srcs_sorted, int64_t n, int mode,
                   const int32_t* seg_of_entry, float* grads_out, const tgr_adam_t* adam, void* workspace,
                   size_t workspace_bytes, void* stream);

/* The same fixed-tile reduction + AdamW (mode 1) over rows of up to 32 flat buffers instead of concat gradients:
 * src = b << 24 | row, contribution = row_bases[b][row, 0:H]. Used by the owner side of the row-sharded exchange with
 * row_bases[b] pointing INTO rank b's bucketed gradient buffer (NVLink peer memory): the reduction pulls the rows in
 * place, there is no gradient all-to-all. row_bases is a HOST array of device pointers. */
int tgr_bwd_reduce_rows(const tgr_table_t* tables, int n_tables, int H, const float* const* row_bases, int n_bases,
                        const uint32_t* keys_sorted, const uint32_t* srcs_sorted, int64_t n, const tgr_adam_t* adam,
                        void* workspace, size_t workspace_bytes, void* stream);

/* AdamW row update from already-reduced rows: for u < *n_unique_dev: row(uniq[u]) <- adamw(row, grads[u]). */
int tgr_adam_rows(const tgr_table_t* tables, int n_tables, int H, const uint32_t* uniq, const float* grads,
                  const int32_t* n_unique_dev, int64_t max_unique, const tgr_adam_t* adam, void* stream);

/* The same update with the hyper-parameter block in DEVICE memory (read when the kernel runs): a captured CUDA graph
 * replays it while the host refreshes the bias corrections of the step with a 48-byte copy. */
int tgr_adam_rows_dev(const tgr_table_t* tables, int n_tables, int H, const uint32_t* uniq, const float* grads,
                      const int32_t* n_unique_dev, int64_t max_unique, const tgr_adam_t* adam_dev, void* stream);

/* AdamW on up to TGR_MAX_DENSE small dense tensors in ONE launch (the path's own Linear layers — itemdnn, userdnn,
 * emb_transform, model/BaseLine/model.py:150-151,166-167 — ~0.1 M parameters; torch's multi-tensor AdamW spends 37 us
 * of device time on them). Same arithmetic as the row update. Exactly one of adam / adam_dev is non-NULL. */
#define TGR_MAX_DENSE 16
typedef struct tgr_dense_list {
  float* w[TGR_MAX_DENSE];
  const float* g[TGR_MAX_DENSE];
  float* m[TGR_MAX_DENSE];
  float* v[TGR_MAX_DENSE];
  int64_t numel[TGR_MAX_DENSE];
  int32_t n;
  int32_t reserved;
} tgr_dense_list_t;
int tgr_adam_dense(const tgr_dense_list_t* list, const tgr_adam_t* adam, const tgr_adam_t* adam_dev, void* stream);

/* Parity mode: add reduced rows into dense per-table gradients tables[].grad (each key once => plain store-add). */
int tgr_scatter_rows(const tgr_table_t* tables, int n_tables, int H, const uint32_t* uniq, const float* grads,
                     const int32_t* n_unique_dev, int64_t max_unique, void* stream);

/* ---- row-sharded multi-GPU helpers (no reference counterpart) ------------------
 * owner = key mod W, local_row = key div W. Input keys sorted ascending & unique. Outputs: keys grouped by
 * owner (ascending inside each bucket), bucket counts [W], and for each input position its slot in the
 * bucketed order (perm). */
size_t tgr_route_workspace_bytes(int64_t max_unique, int W);
int tgr_route_bucket(const uint32_t* uniq, const int32_t* n_unique_dev, int64_t max_unique, int W,
                     uint32_t* bucketed_local_rows, int32_t* perm, int32_t* counts_dev, void* workspace,
                     size_t workspace_bytes, void* stream);

/* Replace every id of an [n / n_cols, n_cols] id matrix by 1 + perm[index of its global key in uniq] (0 for
 * padding): turns the packed ids into row numbers of the buffer the all-to-all returned, so the same fused
 * gather kernel runs on it. col_key_base / col_rows are HOST arrays [n_cols]. */
int tgr_remap_ids(const int32_t* ids, int64_t n, int n_cols, const uint32_t* col_key_base, const int32_t* col_rows,
                  const uint32_t* uniq, const int32_t* n_unique_dev, const int32_t* perm, int32_t* out, void* stream);
/* (perm == NULL means the identity in tgr_remap_ids / tgr_remap_scatter.) */

/* Same remap for all SINGLE slots of up to 4 calls at once, without searching: every sorted (key, src) entry
 * knows its (call, slot, token), so ids_out[call][token, col(slot)] = 1 + perm[seg_of_entry[e]]. ids_out[] must be
 * zero-filled by the caller (padding ids stay 0). ids_out is a HOST array of device pointers. */
int tgr_remap_scatter(const uint32_t* srcs_sorted, const int32_t* seg_of_entry, int64_t n, const int32_t* perm,
                      const tgr_call_t* calls, int n_calls, int32_t* const* ids_out, void* stream);

/* out[u, :] = peer_rows[uniq[u] % n_peers][uniq[u] / n_peers, :] for u < *n_unique_dev: this step's rows copied in
 * place out of their owners' shards over NVLink peer memory (replaces owner-side gather + row all-to-all). A latency-
 * bound copy: meant for a side stream next to value-independent work. peer_rows is a HOST array of device pointers. */
int tgr_fetch_peer_rows(const float* const* peer_rows, int n_peers, int H, const uint32_t* uniq, const int32_t* n_unique_dev,
                        int64_t max_unique, float* out, void* stream);

/* The same remap for the values of every ARRAY slot of up to 4 calls in one launch (searching: a token may hold
 * several values, so the pairs' (call, slot, token) code does not address them). arr_out[c] is indexed like
 * calls[c].arr_val; arr_out is a HOST array of device pointers. */
int tgr_remap_arrays(const tgr_table_t* tables, int n_tables, const tgr_call_t* calls, int n_calls, const uint32_t* uniq,
                     const int32_t* n_unique_dev, const int32_t* perm, int32_t* const* arr_out, void* stream);

/* out[perm[u], :] = in[u, :] for u < *n_dev (inverse = 0), or out[u, :] = in[perm[u], :] (inverse = 1). */
int tgr_permute_rows(const float* in, int H, const int32_t* perm, const int32_t* n_dev, int64_t max_n, int inverse,
                     float* out, void* stream);

/* Gather rows of a flat table by local row index: out[i, :] = table[rows[i], :] for i < *n_dev. */
int tgr_gather_rows(const float* table, int H, const uint32_t* rows, const int32_t* n_dev, int64_t max_n, float* out,
                    void* stream);

/* ---- device-resident item features (csrc/tgr_resident.cu) ---------------------------------------------
 * Item-side sparse features and frozen mm vectors are functions of the item id (model/BaseLine/dataset.py:159,260-263): with
 * the [items + 1, n_feat] int32 feature table and the [items + 1, mm_dim] mm tables resident in HBM, a call is expanded on the
 * device from its item-id column into exactly the packed ids / mm inputs the host tensorizer produces (model.py:186-224,283-296). */
/* ids_out [T, n_single]: column id_col = item_ids[t], columns [col0, col0 + n_feat) = feat_table[item_ids[t], :], the rest 0. */
int tgr_expand_item_features(const int32_t* item_ids, int64_t T, int n_single, int id_col, int col0, const int32_t* feat_table,
                             int n_feat, int64_t n_items, int32_t* ids_out, void* stream);
/* ids[tok[u], col0 + j] = vals[u, j]: the user-side columns of the few user tokens (one per sequence, dataset.py:119). */
int tgr_scatter_user_tokens(const int32_t* tok, const int32_t* vals, int n_tok, int n_cols, int n_single, int col0, int32_t* ids,
                            void* stream);
/* out[t, :] = table[item_ids[t], :], rows of mm_dim elements (fp32 or bf16, a multiple of 16 bytes). */
int tgr_gather_mm_rows(const int32_t* item_ids, int64_t T, const void* table, int dtype, int mm_dim, int64_t n_items, void* out,
                       void* stream);

/* ---- peer-memory transport of the sharded exchange (csrc/tgr_symm.cu; no reference counterpart) -----------------------
 * Small messages move by kernels that store to / load from the other ranks' symmetric-memory buffers over NVLink, ordered by
 * device-side barriers: no NCCL collective and no host round trip on the step's critical path. All pointer arrays are HOST
 * arrays of device pointers mapped into this process (own buffer included). */
/* dst_bases[p][rank * n + i] = src[i] for every peer p: all-gather of one small int32 vector (per-owner counts) by stores. */
int tgr_peer_put(void* const* dst_bases, int n_peers, int rank, const int32_t* src, int n, void* stream);
/* dst = concat over s of src_ptrs[s][0 : counts[s]]: this owner's bucket out of every source's bucketed local-row list. */
int tgr_peer_pull(const uint32_t* const* src_ptrs, const int64_t* counts, int n_peers, uint32_t* dst, void* stream);
/* Stable n_buckets-way merge of sorted buckets stored back to back in `rows` (counts on the host): keys_out = merged keys,
 * code_out = bucket << 24 | index inside the bucket (with_code) or the position in `rows`. Bit-identical to a stable sort by
 * key of the concatenation (ties keep bucket order) — replaces the owner-side second radix sort. */
int tgr_merge_buckets(const uint32_t* rows, const int64_t* counts, int n_buckets, int with_code, uint32_t* keys_out,
                      uint32_t* code_out, void* stream);
/* Device-side barrier between the ranks of one node over symmetric memory: flags[r] = rank r's uint32 flag array
 * (>= n_peers entries, zero-initialised, mapped into this process), epoch = 1, 2, 3, ... identical on every rank. Orders
 * everything enqueued before it on `stream` (on every rank) before everything enqueued after it. */
int tgr_peer_barrier(uint32_t* const* flags, int rank, int n_peers, uint32_t epoch, void* stream);
/* out[i] = scale * sum over r (ascending) of peers[r][i]: one-shot pull all-reduce of a small replicated buffer. */
int tgr_allreduce_peers(const float* const* peers, int n_peers, int64_t n, float scale, float* out, void* stream);

/* ---- factored path: the item/user DNN applied to DEDUPLICATED rows -----------------
 * model.py:302-307 computes out = relu(itemdnn(cat(item slots))) + relu(userdnn(cat(user slots))); a Linear over
 * a concat is a sum of per-slot H x H blocks applied to the slot's row, and every table feeds exactly one slot
 * (model.py:244-245,252-263), so the block product is formed ONCE per unique row of the step and the concat
 * buffers, their gradients and the [T, 1024] GEMMs never exist. Pipeline (keys/sort/dedup shared with the backward):
 *   build_keys -> sort_pairs -> dedup -> remap_scatter/remap_ids (ids -> 1 + unique index)
 *   fact_project_rows -> [fact_mm_fold, mm_proj_fwd] -> fact_forward                      (forward)
 *   fact_relu_mask -> bwd_reduce(mode 0 on dZ) -> fact_unique_backward [-> mm_proj_bwd, fact_mm_chain_bwd]
 *   -> adam_rows | scatter_rows                                                            (backward + update)
 * Supported H: 32, 64, 128. All fp32; every reduction order is fixed by the sorted unique list. */
typedef struct tgr_dnn {
  const float* w_item; /* itemdnn.weight [H, item_ld] (model.py:151) */
  const float* w_user; /* userdnn.weight [H, user_ld] (model.py:150) or NULL */
  int64_t item_ld, user_ld;
  int32_t table_side[TGR_MAX_TABLES]; /* TGR_SIDE_* of the one slot table t feeds */
  int32_t table_col[TGR_MAX_TABLES];  /* first DNN-input column of that slot */
} tgr_dnn_t;

/* Where the factored kernels read table rows from. All NULL / 0 (or src == NULL): tables[].weight.
 *   fetched_rows [+ fetched_perm]: row-sharded tables, the step's rows already fetched from their owners (all-to-all):
 *       row(uniq[u]) = fetched_rows[fetched_perm[u]] (fetched_perm NULL: fetched_rows[u]);
 *   peer_rows[0..n_peers): row-sharded tables read IN PLACE over NVLink peer memory (one kernel does the fetch and the
 *       projection: no gather kernel, no row all-to-all): row(key) = peer_rows[key % n_peers][key / n_peers];
 *   save_rows: tgr_fact_project_rows also stores the raw rows there, [n_unique, H] (the backward needs them again).
 * With a non-table source the table weights may be NULL (key bases / rows are still read). */
typedef struct tgr_row_source {
  const float* fetched_rows;
  const int32_t* fetched_perm;
  const float* peer_rows[TGR_MAX_PEERS];
  int32_t n_peers;
  int32_t reserved;
  float* save_rows;
} tgr_row_source_t;

/* P[u, :] = W_side[:, col : col+H] . row(uniq[u]) for u < *n_unique_dev (uniq sorted ascending). */
int tgr_fact_project_rows(const tgr_table_t* tables, int n_tables, int H, const tgr_dnn_t* dnn, const uint32_t* uniq,
                          const int32_t* n_unique_dev, int64_t max_unique, const tgr_row_source_t* src, float* P,
                          void* stream);

/* One feat2emb call from the projected rows: out[t] = relu(b_item + sum P[ids_u[t, item slots]-1] + sum_f mmz[f][t])
 * + relu(b_user + sum P[ids_u[t, user slots]-1]); arrays via call->arr_off and arr_u (remapped arr_val). ids 0 add
 * nothing. mask[t, H/4]: bit j / 4+j = z_item / z_user element 4c+j > 0 (the ReLU masks of model.py:303,306). */
int tgr_fact_forward(const tgr_call_t* call, int H, const int32_t* ids_u, const int32_t* arr_u, const float* P,
                     const float* const* mmz, int n_mm, const float* bias_item, const float* bias_user, float* out,
                     uint8_t* mask, void* stream);

/* dz_item = d_out * mask_item, dz_user = d_out * mask_user (NULL without users); db_* += column sums
 * (autograd of relu + the Linear bias, model.py:303-307). Optional fused mm statistics for ONE 32-wide mm feature
 * ('81'): mm_A [H, 32] += dz_item^T . mm_x and mm_s [H] += colsum(dz_item), from the same pass (mm_x NULL = off). */
size_t tgr_fact_relu_mask_workspace_bytes(int64_t T, int H);
int tgr_fact_relu_mask(const float* d_out, const uint8_t* mask, int64_t T, int H, float* dz_item, float* dz_user,
                       float* db_item, float* db_user, const void* mm_x, int mm_x_dtype, int mm_dim, float* mm_A,
                       float* mm_s, void* workspace, size_t workspace_bytes, void* stream);

/* G[u, :] (sum of dz over the row's lookups, from tgr_bwd_reduce mode 0) -> row gradient G[u] . W[:, cols] in place;
 * dW_item / dW_user [H, ld] += sum_u G[u]^T (x) row[u] in the slot's columns (autograd of the Linear weight). */
size_t tgr_fact_backward_workspace_bytes(int n_tables, int H);
int tgr_fact_unique_backward(const tgr_table_t* tables, int n_tables, int H, const tgr_dnn_t* dnn, const uint32_t* uniq,
                             const int32_t* n_unique_dev, int64_t max_unique, const tgr_row_source_t* src, float* G,
                             float* dW_item, float* dW_user, void* workspace, size_t workspace_bytes, void* stream);

/* mm feature folded through its item-DNN block: M = W_slot . W_mm [H, mm_dim], c = W_slot . b_mm [H]
 * (emb_transform then itemdnn, model.py:297-303); x . M^T + c is then produced by tgr_mm_proj_fwd. */
int tgr_fact_mm_fold(const float* w_slot, int64_t ld, const float* w_mm, const float* b_mm, int H, int mm_dim, float* M,
                     float* c, void* stream);
/* Chain rule back from A = dz^T x [H, mm_dim], s = colsum(dz) [H] (tgr_mm_proj_bwd):
 * dW_mm += W_slot^T A ; db_mm += W_slot^T s ; dW_slot += A W_mm^T + s b_mm^T. */
int tgr_fact_mm_chain_bwd(const float* w_slot, int64_t ld, const float* w_mm, const float* b_mm, const float* A,
                          const float* s, int H, int mm_dim, float* dW_mm, float* db_mm, float* dW_slot, int64_t dld,
                          void* stream);

/* ---- factored path, step-level driver: one call per phase instead of one per kernel ------------------------
 * The launch sequence of the factored pipeline is fixed once the packed calls are known, so it is sequenced in
 * the library (tgr_fact_step.cu) and the host makes ~8 calls per training step instead of ~45. All buffers are
 * carved from ONE caller-provided arena; results are identical to the per-kernel entries above. */
#define TGR_MAX_MM 6

typedef struct tgr_mm_feat { /* one emb_transform[k] (model.py:167) and its slot's first item-DNN input column */
  const float* w;            /* [H, mm_dim] */
  const float* b;            /* [H] or NULL */
  int32_t mm_dim;
  int32_t col;
} tgr_mm_feat_t;

typedef struct tgr_fact_params { /* parameters a group's forward / backward reads (borrowed per call) */
  tgr_dnn_t dnn;
  const float* b_item; /* itemdnn.bias */
  const float* b_user; /* userdnn.bias */
  tgr_mm_feat_t mm[TGR_MAX_MM];
  int32_t n_mm;
  int32_t reserved;
} tgr_fact_params_t;

typedef struct tgr_fact_grads { /* zero-initialised accumulators (+=), NULL = not wanted */
  float* dW_item;
  float* db_item;
  float* dW_user;
  float* db_user;
  float* dW_mm[TGR_MAX_MM];
  float* db_mm[TGR_MAX_MM];
} tgr_fact_grads_t;

typedef struct tgr_fact_group {
  /* ---- filled by the caller before tgr_fact_group_bytes / tgr_fact_prepare ---- */
  int32_t n_calls, H, key_bits, n_mm;
  int64_t n;                           /* count of non-padding in-range ids over the calls: exact (host-known), or an
                                          upper bound when n_is_capacity != 0 */
  int32_t mm_dim[TGR_MAX_MM];
  int32_t mm_x_dtype;                  /* TGR_DTYPE_* of the mm inputs */
  int32_t n_is_capacity;               /* != 0: the kernels take the count from device memory (what build_keys emitted),
                                          n only sizes buffers and grids — the launch sequence is then a function of
                                          the calls' shapes alone and can be captured in a CUDA graph */
  tgr_call_t calls[TGR_MAX_CALLS];     /* ids / arrays of every call; item_cat / user_cat unused */
  const void* mm_x[TGR_MAX_CALLS][TGR_MAX_MM]; /* [T, mm_dim] inputs of every call */
  tgr_row_source_t src;                /* row-sharded tables: set BEFORE the first forward (fetched rows or peer
                                          shards; src.save_rows is carved by the driver); all zero = own tables */
  /* ---- carved from the arena by tgr_fact_prepare (device pointers; read-only for the caller) ---- */
  int64_t cap;                         /* rows of uniq / P / G (= max(n, 1)) */
  uint32_t *keys_in, *srcs_in, *keys, *srcs; /* unsorted / stably sorted (key, src) pairs [n] */
  uint32_t* uniq;                      /* sorted unique keys [*n_unique] */
  int32_t *seg_off, *seg_of, *n_unique, *n_valid;
  float* P;                            /* projected rows [*n_unique, H] */
  float* rows_local;                   /* peer source only: raw rows kept for the backward [*n_unique, H] */
  float* G;                            /* after the finishing backward: row gradients [*n_unique, H] */
  int32_t* ids_u[TGR_MAX_CALLS];
  int32_t* arr_u[TGR_MAX_CALLS];
  uint8_t* mask[TGR_MAX_CALLS];
  float* dz_item[TGR_MAX_CALLS];
  float* dz_user[TGR_MAX_CALLS];
  float* mmz[TGR_MAX_CALLS][TGR_MAX_MM];
  float *fold_M[TGR_MAX_MM], *fold_c[TGR_MAX_MM], *mm_A[TGR_MAX_MM], *mm_s[TGR_MAX_MM];
  void* fold_Mb[TGR_MAX_MM];           /* bf16 copy of fold_M for the tcgen05 projection (wide bf16 mm features) */
  void* dzb;                           /* bf16 copy of one call's dz_item for the tcgen05 mm backward, [max T, H] */
  void* ws;
  size_t ws_bytes;
  int32_t projected, n_backward;       /* progress */
  int32_t mm_done, mm_joined;          /* tgr_fact_mm_branch ran / the caller's stream has waited for it */
  void* reduce_done_event;             /* optional cudaEvent_t (caller-owned): recorded on the stream right after the
                                          segmented reduce of the finishing backward — the L2-sensitive gathers of the step
                                          are over, what follows (row gradients, row update) is latency- / HBM-bound and
                                          a good neighbour for the NEXT step's key processing (graphed.PipelinedStep) */
} tgr_fact_group_t;

/* Arena bytes tgr_fact_prepare needs for this group (depends on n, the calls' T / n_single / arr_nnz, H, mm dims). */
size_t tgr_fact_group_bytes(const tgr_fact_group_t* g, int n_tables);
/* keys -> sort -> dedup -> ids remapped to 1 + unique index, for all calls of the group. Value independent. */
int tgr_fact_prepare(const tgr_table_t* tables, int n_tables, tgr_fact_group_t* g, void* arena, size_t arena_bytes,
                     void* stream);
/* Optional, right after tgr_fact_prepare on the same stream, when the forwards follow immediately (the weights the mm
 * projection reads are already final): fold + mm projection of EVERY call of the group on an internal side stream that
 * forks from the point where tgr_fact_prepare started, so the small value-dependent mm kernels run next to the issue-bound
 * key processing (sort / dedup) instead of after it. The first tgr_fact_call_forward joins. Capturable in a CUDA graph.
 * fork_now != 0: the group was prepared EARLIER (one step ahead, possibly on another stream): the side stream forks from
 * `stream` at this call instead. */
int tgr_fact_mm_branch(const tgr_fact_params_t* prm, tgr_fact_group_t* g, int fork_now, void* stream);
/* feat2emb forward of call c into out [T, H]; the first forward of a group projects the unique rows. */
int tgr_fact_call_forward(const tgr_table_t* tables, int n_tables, const tgr_fact_params_t* prm, tgr_fact_group_t* g,
                          int c, float* out, void* stream);
/* Backward of call c from d_out [T, H] (c < 0: none). finish != 0 (after the group's last call): segmented reduce of
 * dZ by key, row gradients into g->G, DNN weight gradients into gr. */
int tgr_fact_call_backward(const tgr_table_t* tables, int n_tables, const tgr_fact_params_t* prm, tgr_fact_group_t* g,
                           int c, const float* d_out, const tgr_fact_grads_t* gr, int finish, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TGR_EMBED_H_ */
