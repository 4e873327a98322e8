// Parameter block of the factored row kernels (shared by tgr_factored.cu and tgr_rows_ws.cu).
#pragma once
#include "tgr_common.cuh"

namespace tgr {

struct FactParams {
  const float* w[TGR_MAX_TABLES];       // table rows
  uint32_t key_base[TGR_MAX_TABLES + 1];
  int32_t col[TGR_MAX_TABLES];          // first DNN-input column of the table's slot
  int8_t side[TGR_MAX_TABLES];          // which DNN the table's slot feeds
  const float* dnn_w[2];                // itemdnn.weight [H, item_dim], userdnn.weight [H, user_dim]
  int64_t dnn_ld[2];
  const float* fetched;                 // row-sharded tables: rows of this step fetched from their owners, or NULL
  const int32_t* fetched_perm;          // row of unique key u = fetched[fetched_perm[u]] (NULL: fetched[u])
  const float* peer[TGR_MAX_PEERS];     // row-sharded tables read in place over NVLink: shard of owner r (peer memory)
  int32_t n_peers;                      // > 0: row(key) = peer[key % n_peers][key / n_peers]
  float* save_rows;                     // MODE 0: also keep the raw rows, [U, H] (the backward's dW needs them again)
  int32_t n_tables;
};

// Warp-specialised tcgen05 row kernels (tgr_rows_ws.cu), H = 64. Return 0 / negative like the C ABI entries.
bool rows_ws_supported(int H);
int launch_rows_ws_fwd(const FactParams& p, const uint32_t* uniq, const int32_t* n_unique_dev, float* P, cudaStream_t st);
int launch_rows_ws_bwd(const FactParams& p, const uint32_t* uniq, const int32_t* n_unique_dev, float* G, float* dw_part,
                       cudaStream_t st);
int rows_ws_bwd_grid();
int rows_ws_bwd_rt();

}  // namespace tgr
