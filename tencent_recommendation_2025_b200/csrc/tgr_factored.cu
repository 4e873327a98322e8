// Factored feat2emb: the item/user DNN is applied to DEDUPLICATED table rows, so the [T,1024]/[T,576] concat
// buffers, their gradients and the [T,1024]x[1024,64] GEMMs (forward, dX, dW) never exist.
//
// Reference (model/BaseLine/model.py:302-307):  out = relu(itemdnn(cat(item slots))) + relu(userdnn(cat(user slots)))
// and  itemdnn(cat(...)) = b + sum_slots W[:, cols(slot)] . row_slot(id)  — linear in every gathered row. A training
// step looks up ~2.8 M rows but only ~0.23 M DISTINCT ones (Zipf ids, small-vocabulary features; each table feeds
// exactly one slot, model.py:244-245,252-263), hence:
//
//   project_rows     P[u]      = W[:, cols(table(u))] . row[u]             once per unique row (fp32 FFMA tile GEMM)
//   forward          z_side[t] = b_side + sum_slots P[idx(t, slot)] (+ folded mm projections);
//                    out       = relu(z_item) + relu(z_user);  mask = sign bits for the backward
//                                                                           gather-sum of 256 B rows that live in L2
//   relu_mask        dZ_side   = dOut * mask_side ; db_side = column sums   (fixed-order partials)
//   (tgr_bwd_reduce) G[u]      = sum over the row's lookups of dZ_side[token]  (mode 0 on the dZ buffers, ld = H)
//   unique_backward  g_row[u]  = G[u] . W[:, cols]      -> AdamW row update / dense scatter
//                    dW[:, cols] += sum_u G[u]^T (x) row[u]                  split-K partials, reduced in fixed order
//   mm features      z_item   += x . (W_s Wmm)^T + W_s bmm   with the folded [H, mm_dim] matrix (mm_fold);
//                    dWmm, dbmm, dW_s from A = dZ^T x, s = colsum(dZ)        (mm_chain_bwd)
//
// Sums are re-associated (per-slot H-term dot products, then a <= 25-term sum) — fp32 throughout, inside the 1e-5
// parity bar. Every reduction order is a function of the sorted unique-key list and the launch geometry only.
// This is SURVEY.md §8(f) N4 obtained through deduplication instead of a gather-prologue GEMM.
#include "tgr_common.cuh"
#include "tgr_mma.cuh"
#include "tgr_rows.cuh"
#include "tgr_tc.cuh"
#include "tgr_fact_params.cuh"

#include <stdlib.h>

namespace tgr {

constexpr int kFT = 256;          // threads per CTA
constexpr int kRowsGridFwd = 6 * kNumSMs;   // H = 64: 35 KB smem, 64 threads x 163 regs -> 6 CTAs / SM
constexpr int kRowsGridBwd = 4 * kNumSMs;   // H = 64: 52 KB smem, 64 threads x 225 regs -> 4 CTAs / SM


// 16-byte global -> shared copy that bypasses the register file (LDGSTS); src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ void fma4(float4& a, float x, const float4& w) {
  a.x = fmaf(x, w.x, a.x); a.y = fmaf(x, w.y, a.y); a.z = fmaf(x, w.z, a.z); a.w = fmaf(x, w.w, a.w);
}

// MODE 0: PG[u] = W_s . row[u]                       (forward projection; PG is write-only)
// MODE 1: PG[u] (holding G[u]) <- G[u] . W_s  in place, and the CTA's split-K partial of dW_s = sum_u G[u]^T row[u]
// Persistent: CTA b owns the contiguous tile range [b*tpc, (b+1)*tpc) of the sorted unique list, so the rows of one
// table are consecutive and its H x H weight block / dW accumulator stay on chip across tiles.
// fp32 FFMA with 8 x 8 register tiles: 4 LDS.128 feed 64 FMAs, which balances the shared-memory pipe against the
// FMA pipe (the first version used 4 x 4 tiles and was LDS-bound at ~30 % of the FFMA rate, profiles/README.md).
// threads = (H/8) column groups x (RT/8) row groups over an RT-row tile; the dW GEMM uses (H/8)^2 8x8 tiles and
// RG interleaved row groups whose partials are combined in fixed order by fact_dw_reduce_kernel. Measured on
// B200 (tools/rows_bench.py, 227 k rows): 67 us forward / 130 us backward; with the loads or the GEMMs disabled the
// phases take 25 / 43 us — the GEMM loop runs at ~55 % of the FFMA rate with the LDS and FMA pipes both saturated.
template <int H>
struct RowsCfg {
  static constexpr int RT = H == 64 ? 64 : 128;   // unique rows per tile (64-row tiles: more, smaller CTAs per SM in
                                                  // different load / compute phases)
  static constexpr int TXN = H / 8;               // column groups
  static constexpr int TYN = RT / 8;              // row groups
  static constexpr int NT = TXN * TYN;            // threads
  static constexpr int DT = (H / 8) * (H / 8);
  static constexpr int RG = NT / DT;
  static constexpr int LD = H + 4;
  static_assert(NT % DT == 0 && RG >= 1, "tile shape");
};

template <int H, int MODE>
__global__ void __launch_bounds__(RowsCfg<H>::NT) fact_rows_kernel(const __grid_constant__ FactParams p,
                                                          const uint32_t* __restrict__ uniq,
                                                          const int32_t* __restrict__ n_unique_dev,
                                                          float* __restrict__ PG, float* __restrict__ dw_part) {
  using Cfg = RowsCfg<H>;
  constexpr int NT = Cfg::NT, TXN = Cfg::TXN, TYN = Cfg::TYN, DT = Cfg::DT, RG = Cfg::RG, LD = Cfg::LD, HH = H / 2;
  constexpr int kRT = Cfg::RT;
  extern __shared__ __align__(16) float sm[];
  float* Ws = sm;              // [H][LD]   MODE 0: Ws[k][h] = W[h][col+k]   MODE 1: Ws[h][k] = W[h][col+k]
  float* Xs = Ws + H * LD;     // [kRT][LD] MODE 0: table rows               MODE 1: G rows
  float* Rs = Xs + kRT * LD;   // [kRT][LD] MODE 1: table rows
  __shared__ uint32_t s_key[kRT];
  __shared__ int32_t s_perm[kRT];
  const int tid = threadIdx.x, tx = tid % TXN, ty = tid / TXN;          // row GEMM: rows ty + TYN i, cols tx*4 (+ H/2)
  const int hx = tid % TXN, hy = (tid % DT) / TXN, rg = tid / DT;        // dW GEMM: h = hy*4 (+H/2), k = hx*4 (+H/2)
  const int U = *n_unique_dev;
  const int n_tiles = (U + kRT - 1) / kRT;
  const int tpc = (n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int tile_a = blockIdx.x * tpc, tile_b = min(n_tiles, tile_a + tpc);
  int cur_t = -1;
  float4 dw[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) dw[i][0] = dw[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);

  auto flush_dw = [&](int t) {
    float* dst = dw_part + ((size_t)(blockIdx.x + t) * RG + rg) * H * H;   // (cta, table) pairs are monotone => unique slots
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int h = (i < 4 ? 0 : HH) + hy * 4 + (i & 3);
      *reinterpret_cast<float4*>(dst + h * H + hx * 4) = dw[i][0];
      *reinterpret_cast<float4*>(dst + h * H + HH + hx * 4) = dw[i][1];
      dw[i][0] = dw[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };

  for (int tile = tile_a; tile < tile_b; ++tile) {
    const int r0 = tile * kRT;
    const int nr = min(kRT, U - r0);
    __syncthreads();
    for (int i = tid; i < kRT; i += NT) {
      s_key[i] = i < nr ? __ldg(uniq + r0 + i) : 0xFFFFFFFFu;
      if (p.fetched != nullptr) s_perm[i] = i >= nr ? 0 : (p.fetched_perm ? __ldg(p.fetched_perm + r0 + i) : r0 + i);
    }
    __syncthreads();
    int seg_a = 0;
    while (seg_a < nr) {
      const int t = find_table(p.key_base, p.n_tables, s_key[seg_a]);
      const uint32_t kend = p.key_base[t + 1];
      int seg_b;
      if (s_key[nr - 1] < kend) {
        seg_b = nr;
      } else {   // first row of the next table, by bisection (uniform across the CTA: shared-memory broadcast reads)
        int lo = seg_a + 1, hi = nr - 1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_key[mid] < kend) lo = mid + 1; else hi = mid; }
        seg_b = lo;
      }
      const int ns = seg_b - seg_a;
      __syncthreads();   // previous segment's readers of Xs / Rs / Ws are done
      if (t != cur_t) {
        if (MODE == 1 && cur_t >= 0) flush_dw(cur_t);
        const float* W = p.dnn_w[p.side[t]];
        const int64_t ld = p.dnn_ld[p.side[t]];
        const int col = p.col[t];
        for (int i = tid; i < H * H; i += NT) {
          const int h = i / H, k = i - h * H;
          const float wv = __ldg(W + (size_t)h * ld + col + k);
          if (MODE == 1) Ws[h * LD + k] = wv; else Ws[k * LD + h] = wv;
        }
        cur_t = t;
      }
      const float* tab = p.w[t];
      const uint32_t kb = p.key_base[t];
      // all of the tile's 16-byte pieces go global -> shared asynchronously (every copy in flight at once: the rows
      // are scattered 256 B reads out of HBM, so one latency round instead of one per register-staged batch)
      for (int i = tid; i < kRT * (H / 4); i += NT) {
        const int r = i / (H / 4), c = i - r * (H / 4);
        const bool ok = r < ns;
        const float* rsrc;
        if (p.n_peers > 0) {          // the owner's shard, read in place through NVLink peer memory
          const uint32_t key = ok ? s_key[seg_a + r] : 0u;
          rsrc = p.peer[key % (uint32_t)p.n_peers] + (size_t)(key / (uint32_t)p.n_peers) * H + c * 4;
        } else if (p.fetched != nullptr) {
          rsrc = ok ? p.fetched + (size_t)s_perm[seg_a + r] * H + c * 4 : p.fetched;
        } else {
          rsrc = ok ? tab + (size_t)(s_key[seg_a + r] - kb) * H + c * 4 : tab;
        }
        if (MODE == 1) {
          cp_async16(Rs + r * LD + c * 4, rsrc, ok ? 16 : 0);
          cp_async16(Xs + r * LD + c * 4, ok ? PG + (size_t)(r0 + seg_a + r) * H + c * 4 : PG, ok ? 16 : 0);
        } else {
          cp_async16(Xs + r * LD + c * 4, rsrc, ok ? 16 : 0);
        }
      }
      cp_async_wait_all();
      __syncthreads();
      if (MODE == 0 && p.save_rows != nullptr) {
        for (int i = tid; i < ns * (H / 4); i += NT) {
          const int r = i / (H / 4), c = i - r * (H / 4);
          st_stream(reinterpret_cast<float4*>(p.save_rows + (size_t)(r0 + seg_a + r) * H) + c,
                    *reinterpret_cast<const float4*>(Xs + r * LD + c * 4));
        }
      }
      // ---- row GEMM: out[r][j] = sum_q Xs[r][q] * Ws[q][j];  r = ty + TYN i, j in {tx*4.., H/2 + tx*4..}
      if (ty < ns) {   // (rows are interleaved by TYN: a row group with no valid row at all only exists in short tails)
        float4 acc[8][2];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int q = 0; q < H; q += 4) {
          float4 xv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) xv[i] = *reinterpret_cast<const float4*>(Xs + (ty + TYN * i) * LD + q);
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            const float4 w0 = *reinterpret_cast<const float4*>(Ws + (q + qq) * LD + tx * 4);
            const float4 w1 = *reinterpret_cast<const float4*>(Ws + (q + qq) * LD + HH + tx * 4);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float x = qq == 0 ? xv[i].x : (qq == 1 ? xv[i].y : (qq == 2 ? xv[i].z : xv[i].w));
              fma4(acc[i][0], x, w0);
              fma4(acc[i][1], x, w1);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = ty + TYN * i;
          if (r < ns) {
            float* dst = PG + (size_t)(r0 + seg_a + r) * H;
            *reinterpret_cast<float4*>(dst + tx * 4) = acc[i][0];
            *reinterpret_cast<float4*>(dst + HH + tx * 4) = acc[i][1];
          }
        }
      }
      // ---- dW GEMM: dw[h][k] += sum_r G[r][h] * R[r][k]; row group rg takes rows rg, rg + RG, ... in order
      if (MODE == 1) {
#pragma unroll 2
        for (int r = rg; r < ns; r += RG) {
          const float4 k0 = *reinterpret_cast<const float4*>(Rs + r * LD + hx * 4);
          const float4 k1 = *reinterpret_cast<const float4*>(Rs + r * LD + HH + hx * 4);
          const float4 g0 = *reinterpret_cast<const float4*>(Xs + r * LD + hy * 4);
          const float4 g1 = *reinterpret_cast<const float4*>(Xs + r * LD + HH + hy * 4);
          fma4(dw[0][0], g0.x, k0); fma4(dw[0][1], g0.x, k1);
          fma4(dw[1][0], g0.y, k0); fma4(dw[1][1], g0.y, k1);
          fma4(dw[2][0], g0.z, k0); fma4(dw[2][1], g0.z, k1);
          fma4(dw[3][0], g0.w, k0); fma4(dw[3][1], g0.w, k1);
          fma4(dw[4][0], g1.x, k0); fma4(dw[4][1], g1.x, k1);
          fma4(dw[5][0], g1.y, k0); fma4(dw[5][1], g1.y, k1);
          fma4(dw[6][0], g1.z, k0); fma4(dw[6][1], g1.z, k1);
          fma4(dw[7][0], g1.w, k0); fma4(dw[7][1], g1.w, k1);
        }
      }
      seg_a = seg_b;
    }
  }
  if (MODE == 1 && cur_t >= 0) flush_dw(cur_t);
}

// ---- tensor-core variant of the two row GEMMs (H = 32, 64): 3xTF32 mma.sync, fp32 accumulate (tgr_mma.cuh) ---------
// Same persistent tiling and the same (CTA, table) split-K partial scheme as fact_rows_kernel; what changes is the
// contraction: a CTA is 4 warps over a 128-row tile, warp w owns rows [32w, 32w+32) x all H columns of the row GEMM
// (2 x H/8 m16n8 accumulator tiles) and, in MODE 1, one 16-row slab (x H/8 / n_groups column tiles) of the H x H dW
// block, accumulated over the tile's rows (K = 128 per tile) and kept in registers across the CTA's tiles.
// Shared memory: the table's weight block already split into tf32 hi / lo parts, laid out [n][k] with pitch H + 4 so
// that every fragment load of the row GEMM is bank-conflict free (rows 4 banks apart); the row tiles stay fp32 and are
// split after the load. 70 KB (MODE 0, 3 CTAs / SM) / 104 KB (MODE 1, 2 CTAs / SM) at H = 64.
template <int H>
struct MmaRowsCfg {
  static constexpr int RT = 128, NT = 128, LD = H + 4, NTILES = H / 8, KSTEPS = H / 8;
  static constexpr int MT_DW = H / 16;            // 16-row slabs of the dW block
  static constexpr int NG_DW = 4 / MT_DW;         // warps sharing one slab split its column tiles
  static constexpr int NPW = NTILES / NG_DW;      // column tiles of the dW block per warp
  static_assert(H == 32 || H == 64, "tile shape");
};
constexpr int kMmaGridFwd = 3 * kNumSMs;
constexpr int kMmaGridBwd = 2 * kNumSMs;

template <int H, int MODE>
__global__ void __launch_bounds__(MmaRowsCfg<H>::NT) fact_rows_mma_kernel(const __grid_constant__ FactParams p,
                                                                          const uint32_t* __restrict__ uniq,
                                                                          const int32_t* __restrict__ n_unique_dev,
                                                                          float* __restrict__ PG,
                                                                          float* __restrict__ dw_part) {
  using Cfg = MmaRowsCfg<H>;
  constexpr int kRT = Cfg::RT, NT = Cfg::NT, LD = Cfg::LD, NTILES = Cfg::NTILES, KSTEPS = Cfg::KSTEPS;
  constexpr int MT_DW = Cfg::MT_DW, NPW = Cfg::NPW;
  extern __shared__ __align__(16) float sm[];
  uint32_t* Whi = reinterpret_cast<uint32_t*>(sm);   // [H][LD] tf32 hi part; MODE 0: [n = h][k]  MODE 1: [n = k][h]
  uint32_t* Wlo = Whi + H * LD;                      // [H][LD] tf32 lo part
  float* Xs = reinterpret_cast<float*>(Wlo + H * LD);   // [kRT][LD] MODE 0: table rows   MODE 1: G rows
  float* Rs = Xs + kRT * LD;                            // [kRT][LD] MODE 1: table rows
  __shared__ uint32_t s_key[kRT];
  __shared__ int32_t s_perm[kRT];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const int U = *n_unique_dev;
  const int n_tiles = (U + kRT - 1) / kRT;
  const int tpc = (n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int tile_a = blockIdx.x * tpc, tile_b = min(n_tiles, tile_a + tpc);
  const int mt_dw = warp % MT_DW, ng_dw = warp / MT_DW;
  int cur_t = -1;
  float dw[NPW][4];
#pragma unroll
  for (int j = 0; j < NPW; ++j) dw[j][0] = dw[j][1] = dw[j][2] = dw[j][3] = 0.f;

  auto flush_dw = [&](int t) {
    float* dst = dw_part + (size_t)(blockIdx.x + t) * H * H;   // (cta, table) pairs are monotone => unique slots
#pragma unroll
    for (int j = 0; j < NPW; ++j) {
      const int h = 16 * mt_dw + g, k = 8 * (ng_dw * NPW + j) + 2 * t4;
      *reinterpret_cast<float2*>(dst + h * H + k) = make_float2(dw[j][0], dw[j][1]);
      *reinterpret_cast<float2*>(dst + (h + 8) * H + k) = make_float2(dw[j][2], dw[j][3]);
      dw[j][0] = dw[j][1] = dw[j][2] = dw[j][3] = 0.f;
    }
  };

  for (int tile = tile_a; tile < tile_b; ++tile) {
    const int r0 = tile * kRT;
    const int nr = min(kRT, U - r0);
    __syncthreads();
    for (int i = tid; i < kRT; i += NT) {
      s_key[i] = i < nr ? __ldg(uniq + r0 + i) : 0xFFFFFFFFu;
      if (p.fetched != nullptr) s_perm[i] = i >= nr ? 0 : (p.fetched_perm ? __ldg(p.fetched_perm + r0 + i) : r0 + i);
    }
    __syncthreads();
    int seg_a = 0;
    while (seg_a < nr) {
      const int t = find_table(p.key_base, p.n_tables, s_key[seg_a]);
      const uint32_t kend = p.key_base[t + 1];
      int seg_b;
      if (s_key[nr - 1] < kend) {
        seg_b = nr;
      } else {   // first row of the next table, by bisection (uniform across the CTA: shared-memory broadcast reads)
        int lo = seg_a + 1, hi = nr - 1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_key[mid] < kend) lo = mid + 1; else hi = mid; }
        seg_b = lo;
      }
      const int ns = seg_b - seg_a;
      __syncthreads();   // previous segment's readers of Xs / Rs / W are done
      if (t != cur_t) {
        if (MODE == 1 && cur_t >= 0) flush_dw(cur_t);
        const float* W = p.dnn_w[p.side[t]];
        const int64_t ld = p.dnn_ld[p.side[t]];
        const int col = p.col[t];
        constexpr int WPT = H * (H / 4) / NT;   // all 128-bit loads in flight before the first use
        float4 wv[WPT];
#pragma unroll
        for (int q = 0; q < WPT; ++q) {
          const int i = tid + q * NT, h = i / (H / 4), c = i - h * (H / 4);
          wv[q] = __ldg(reinterpret_cast<const float4*>(W + (size_t)h * ld + col) + c);   // col % H == 0, ld % 4 == 0
        }
#pragma unroll
        for (int q = 0; q < WPT; ++q) {
          const int i = tid + q * NT, h = i / (H / 4), k = (i - h * (H / 4)) * 4;
          const float w4[4] = {wv[q].x, wv[q].y, wv[q].z, wv[q].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t hi, lo;
            split_tf32(w4[j], hi, lo);
            const int idx = MODE == 1 ? (k + j) * LD + h : h * LD + k + j;
            Whi[idx] = hi;
            Wlo[idx] = lo;
          }
        }
        cur_t = t;
      }
      const float* tab = p.w[t];
      const uint32_t kb = p.key_base[t];
      for (int i = tid; i < kRT * (H / 4); i += NT) {   // every 16-byte piece of the tile in flight at once (zero-filled past ns)
        const int r = i / (H / 4), c = i - r * (H / 4);
        const bool ok = r < ns;
        const float* rsrc;
        if (p.n_peers > 0) {          // the owner's shard, read in place through NVLink peer memory
          const uint32_t key = ok ? s_key[seg_a + r] : 0u;
          rsrc = p.peer[key % (uint32_t)p.n_peers] + (size_t)(key / (uint32_t)p.n_peers) * H + c * 4;
        } else if (p.fetched != nullptr) {
          rsrc = ok ? p.fetched + (size_t)s_perm[seg_a + r] * H + c * 4 : p.fetched;
        } else {
          rsrc = ok ? tab + (size_t)(s_key[seg_a + r] - kb) * H + c * 4 : tab;
        }
        if (MODE == 1) {
          cp_async16(Rs + r * LD + c * 4, rsrc, ok ? 16 : 0);
          cp_async16(Xs + r * LD + c * 4, ok ? PG + (size_t)(r0 + seg_a + r) * H + c * 4 : PG, ok ? 16 : 0);
        } else {
          cp_async16(Xs + r * LD + c * 4, rsrc, ok ? 16 : 0);
        }
      }
      cp_async_wait_all();
      __syncthreads();
      if (MODE == 0 && p.save_rows != nullptr) {
        for (int i = tid; i < ns * (H / 4); i += NT) {
          const int r = i / (H / 4), c = i - r * (H / 4);
          st_stream(reinterpret_cast<float4*>(p.save_rows + (size_t)(r0 + seg_a + r) * H) + c,
                    *reinterpret_cast<const float4*>(Xs + r * LD + c * 4));
        }
      }
      // ---- row GEMM: out[r][n] = sum_k Xs[r][k] * Wsm[n][k]; warp rows [32 warp, 32 warp + 32)
      const int m0 = warp * 32;
      if (m0 < ns) {
        float acc[2][NTILES][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < NTILES; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
#pragma unroll 2
        for (int ks = 0; ks < KSTEPS; ++ks) {
          uint32_t ahi[2][4], alo[2][4];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const float* xr = Xs + (m0 + 16 * mt + g) * LD + 8 * ks + t4;
            split_tf32(xr[0], ahi[mt][0], alo[mt][0]);
            split_tf32(xr[8 * LD], ahi[mt][1], alo[mt][1]);
            split_tf32(xr[4], ahi[mt][2], alo[mt][2]);
            split_tf32(xr[8 * LD + 4], ahi[mt][3], alo[mt][3]);
          }
#pragma unroll
          for (int nt = 0; nt < NTILES; ++nt) {
            const int wi = (8 * nt + g) * LD + 8 * ks + t4;
            const uint32_t bhi[2] = {Whi[wi], Whi[wi + 4]};
            const uint32_t blo[2] = {Wlo[wi], Wlo[wi + 4]};
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) mma_3xtf32(acc[mt][nt], ahi[mt], alo[mt], bhi, blo);
          }
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int ra = m0 + 16 * mt + g, rb = ra + 8;
          float* da = PG + (size_t)(r0 + seg_a + ra) * H + 2 * t4;
          float* db = PG + (size_t)(r0 + seg_a + rb) * H + 2 * t4;
#pragma unroll
          for (int nt = 0; nt < NTILES; ++nt) {
            if (ra < ns) *reinterpret_cast<float2*>(da + 8 * nt) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
            if (rb < ns) *reinterpret_cast<float2*>(db + 8 * nt) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
          }
        }
      }
      // ---- dW GEMM: dw[h][k] += sum_r G[r][h] * R[r][k] over the tile's rows (zero-filled past ns), 8 rows per step
      if (MODE == 1) {
        const int ksn = (ns + 7) >> 3;
#pragma unroll 2
        for (int ks = 0; ks < ksn; ++ks) {
          const float* ga = Xs + (8 * ks + t4) * LD + 16 * mt_dw + g;
          uint32_t ahi[4], alo[4];
          split_tf32(ga[0], ahi[0], alo[0]);
          split_tf32(ga[8], ahi[1], alo[1]);
          split_tf32(ga[4 * LD], ahi[2], alo[2]);
          split_tf32(ga[4 * LD + 8], ahi[3], alo[3]);
          const float* rb = Rs + (8 * ks + t4) * LD + 8 * (ng_dw * NPW) + g;
#pragma unroll
          for (int j = 0; j < NPW; ++j) {
            uint32_t bhi[2], blo[2];
            split_tf32(rb[8 * j], bhi[0], blo[0]);
            split_tf32(rb[8 * j + 4 * LD], bhi[1], blo[1]);
            mma_3xtf32(dw[j], ahi, alo, bhi, blo);
          }
        }
      }
      seg_a = seg_b;
    }
  }
  if (MODE == 1 && cur_t >= 0) flush_dw(cur_t);
}

// ---- tcgen05 variant of the forward projection (H = 32, 64): P[u] = W_s . row[u] on the 5th-generation tensor cores ----
// M = 128 unique rows x N = H x K = H per tile, kind::tf32 with the 3xTF32 split (tgr_mma.cuh): 3 * H/8 MMAs issued by one
// thread into ONE TMEM accumulator [128 lanes x H fp32 columns], committed to an mbarrier; the four warps then read
// their 32 lanes back with tcgen05.ld and store the rows. Operands are staged in shared memory in the K-major no-swizzle
// core-matrix layout (tgr_tc.cuh): rows are loaded with 128-bit LDGs (lanes 0-7 = 8 consecutive rows of one 64-byte
// column block, so each row segment is two full sectors and each STS.128 phase writes one 128-byte core matrix),
// split into tf32 hi / lo in registers and stored to the A_hi / A_lo tiles. mma.sync (fact_rows_mma_kernel) tops out at
// the legacy tensor path's rate (measured 54 us for 227 k rows vs 70 us FFMA); the tcgen05 pipe runs kind::tf32 ~8x faster,
// which leaves the kernel bound by the scattered row reads.
template <int H>
struct TcRowsCfg {
  static constexpr int RT = 128, NT = 128;
  static constexpr int LBO = 128, SBO = (H / 4) * 128;          // bytes
  static constexpr int A_BYTES = RT * H * 4, B_BYTES = H * H * 4;
  static constexpr int TMEM_COLS = H < 32 ? 32 : H;
  static constexpr size_t SMEM = 2 * (size_t)A_BYTES + 2 * (size_t)B_BYTES + 1024;   // + alignment slack
  static_assert(H == 32 || H == 64, "tile shape");
};
constexpr int kTcGridFwd = 2 * kNumSMs;

// Work item of the tc kernel: rows [seg_a, seg_b) of one tile, all of one table.
struct TcItem { int tile, seg_a, seg_b, t; };

template <int H>
__global__ void __launch_bounds__(TcRowsCfg<H>::NT) fact_rows_tc_kernel(const __grid_constant__ FactParams p,
                                                                        const uint32_t* __restrict__ uniq,
                                                                        const int32_t* __restrict__ n_unique_dev,
                                                                        float* __restrict__ P, int dbg) {
  using Cfg = TcRowsCfg<H>;
  constexpr int kRT = Cfg::RT, NT = Cfg::NT, H4 = H / 4;
  constexpr int KT = 8;                                  // tiles whose keys are staged at once
  constexpr int UNITS = (kRT / 8) * (H4 / 4), UPW = UNITS / 4;
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* Ahi = base;
  uint8_t* Alo = Ahi + Cfg::A_BYTES;
  uint8_t* Bhi = Alo + Cfg::A_BYTES;
  uint8_t* Blo = Bhi + Cfg::B_BYTES;
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t s_tmem;
  __shared__ uint32_t s_key[KT * kRT];
  __shared__ int32_t s_perm[KT * kRT];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) tc::mbar_init(&mbar, 1);
  if (warp == 0) tc::tmem_alloc<Cfg::TMEM_COLS>(&s_tmem);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tacc = s_tmem;
  const uint32_t idesc = tc::make_idesc(2u, 128, H);
  uint32_t phase = 0;
  const int U = *n_unique_dev;
  const int n_tiles = (U + kRT - 1) / kRT;
  const int tpc = (n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int tile_a = blockIdx.x * tpc, tile_b = min(n_tiles, tile_a + tpc);
  int cur_t = -1;

  for (int win_a = tile_a; win_a < tile_b; win_a += KT) {
    const int win_b = min(tile_b, win_a + KT);
    __syncthreads();                                     // previous window's readers of s_key are done
    for (int i = tid; i < (win_b - win_a) * kRT; i += NT) {
      const int u = win_a * kRT + i;
      s_key[i] = u < U ? __ldg(uniq + u) : 0xFFFFFFFFu;
      if (p.fetched != nullptr) s_perm[i] = u >= U ? 0 : (p.fetched_perm ? __ldg(p.fetched_perm + u) : u);
    }
    __syncthreads();

    // item after (tile, seg_b): the next run of same-table rows, or tile == win_b when the window is exhausted
    auto item_at = [&](int tile, int seg_a) {
      TcItem it;
      it.tile = tile; it.seg_a = seg_a; it.seg_b = 0; it.t = 0;
      if (tile >= win_b) return it;
      const uint32_t* k = s_key + (tile - win_a) * kRT;
      const int nr = min(kRT, U - tile * kRT);
      it.t = find_table(p.key_base, p.n_tables, k[seg_a]);
      const uint32_t kend = p.key_base[it.t + 1];
      if (k[nr - 1] < kend) {
        it.seg_b = nr;
      } else {   // first row of the next table, by bisection (uniform across the CTA: shared-memory broadcast reads)
        int lo = seg_a + 1, hi = nr - 1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (k[mid] < kend) lo = mid + 1; else hi = mid; }
        it.seg_b = lo;
      }
      return it;
    };
    auto next_of = [&](const TcItem& it) {
      const int nr = min(kRT, U - it.tile * kRT);
      return it.seg_b < nr ? item_at(it.tile, it.seg_b) : item_at(it.tile + 1, 0);
    };
    // unit = (8-row group rg, block of 4 chunks): lane -> (row rg * 8 + lane % 8, chunk 4 * cq + lane / 8); every 128-bit
    // load of the item is issued here, the data is consumed one pipeline stage later
    float4 v[UPW];
    auto issue_loads = [&](const TcItem& it) {
      const int ns = it.seg_b - it.seg_a;
      const uint32_t* k = s_key + (it.tile - win_a) * kRT + it.seg_a;
      const int32_t* pm = s_perm + (it.tile - win_a) * kRT + it.seg_a;
      const float* tab = p.w[it.t];
      const uint32_t kb = p.key_base[it.t];
#pragma unroll
      for (int q = 0; q < UPW; ++q) {
        const int unit = warp * UPW + q;
        const int rg = unit / (H4 / 4), cq = unit % (H4 / 4);
        const int r = rg * 8 + (lane & 7), c = cq * 4 + (lane >> 3);
        v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < ns && !(dbg & 1)) {
          const float* rsrc;
          if (p.n_peers > 0) {
            const uint32_t key = k[r];
            rsrc = p.peer[key % (uint32_t)p.n_peers] + (size_t)(key / (uint32_t)p.n_peers) * H;
          } else if (p.fetched != nullptr) {
            rsrc = p.fetched + (size_t)pm[r] * H;
          } else {
            rsrc = tab + (size_t)(k[r] - kb) * H;
          }
          v[q] = __ldg(reinterpret_cast<const float4*>(rsrc) + c);
        }
      }
    };

    TcItem cur = item_at(win_a, 0);
    issue_loads(cur);
    while (cur.tile < win_b) {
      const int ns = cur.seg_b - cur.seg_a;
      const int row0 = cur.tile * kRT + cur.seg_a;       // first unique row of the item
      // (the previous item's MMAs have completed — every thread waited on the mbarrier — so the tiles are free)
      if (cur.t != cur_t) {
        const float* W = p.dnn_w[p.side[cur.t]];
        const int64_t ld = p.dnn_ld[p.side[cur.t]];
        const int col = p.col[cur.t];
        // B[n = h][k]: core matrix (h / 8, k / 4); all 128-bit loads in flight before the first use
        constexpr int WPT = H * H4 / NT;
        float4 wv[WPT];
#pragma unroll
        for (int q = 0; q < WPT; ++q) {
          const int i = tid + q * NT, h = i / H4, c = i - h * H4;
          wv[q] = __ldg(reinterpret_cast<const float4*>(W + (size_t)h * ld + col) + c);   // col % 4 == 0, ld % 4 == 0
        }
#pragma unroll
        for (int q = 0; q < WPT; ++q) {
          const int i = tid + q * NT, h = i / H4, c = i - h * H4;
          uint4 hi, lo;
          split_tf32(wv[q].x, hi.x, lo.x);
          split_tf32(wv[q].y, hi.y, lo.y);
          split_tf32(wv[q].z, hi.z, lo.z);
          split_tf32(wv[q].w, hi.w, lo.w);
          const int off = (h >> 3) * Cfg::SBO + c * Cfg::LBO + (h & 7) * 16;
          *reinterpret_cast<uint4*>(Bhi + off) = hi;
          *reinterpret_cast<uint4*>(Blo + off) = lo;
        }
        cur_t = cur.t;
      }
#pragma unroll
      for (int q = 0; q < UPW; ++q) {
        const int unit = warp * UPW + q;
        const int rg = unit / (H4 / 4), cq = unit % (H4 / 4);
        const int r = rg * 8 + (lane & 7), c = cq * 4 + (lane >> 3);
        if (p.save_rows != nullptr && r < ns)
          st_stream(reinterpret_cast<float4*>(p.save_rows + (size_t)(row0 + r) * H) + c, v[q]);
        uint4 hi, lo;
        split_tf32(v[q].x, hi.x, lo.x);
        split_tf32(v[q].y, hi.y, lo.y);
        split_tf32(v[q].z, hi.z, lo.z);
        split_tf32(v[q].w, hi.w, lo.w);
        const int off = rg * Cfg::SBO + c * Cfg::LBO + (lane & 7) * 16;
        *reinterpret_cast<uint4*>(Ahi + off) = hi;
        *reinterpret_cast<uint4*>(Alo + off) = lo;
      }
      tc::fence_smem_to_async();
      tc::fence_before_sync();
      __syncthreads();
      if (tid == 0) {
        tc::fence_after_sync();
        const uint32_t a_hi = tc::smem_u32(Ahi), a_lo = tc::smem_u32(Alo), b_hi = tc::smem_u32(Bhi), b_lo = tc::smem_u32(Blo);
#pragma unroll
        for (int ks = 0; ks < H / 8; ++ks) {
          if (dbg & 4) break;
          const uint32_t o = ks * 2 * Cfg::LBO;
          const uint64_t dah = tc::make_desc(a_hi + o, Cfg::LBO, Cfg::SBO, 0), dal = tc::make_desc(a_lo + o, Cfg::LBO, Cfg::SBO, 0);
          const uint64_t dbh = tc::make_desc(b_hi + o, Cfg::LBO, Cfg::SBO, 0), dbl = tc::make_desc(b_lo + o, Cfg::LBO, Cfg::SBO, 0);
          tc::mma_tf32(tacc, dal, dbh, idesc, ks > 0);
          tc::mma_tf32(tacc, dah, dbl, idesc, true);
          tc::mma_tf32(tacc, dah, dbh, idesc, true);
        }
        tc::commit(&mbar);
      }
      // software pipeline: the next item's rows are requested now and travel while the MMAs and the epilogue run
      const TcItem nxt = next_of(cur);
      if (nxt.tile < win_b) issue_loads(nxt);
      tc::mbar_wait(&mbar, phase);
      phase ^= 1u;
      tc::fence_after_sync();
      {
        const int r = warp * 32 + lane;
        float* dst = P + (size_t)(row0 + r) * H;
        const uint32_t taddr = tacc + ((uint32_t)(warp * 32) << 16);
#pragma unroll
        for (int c0 = 0; c0 < H; c0 += 16) {
          uint32_t rr[16];
          tc::ld16(taddr + c0, rr);
          tc::ld_wait();
          if (r < ns && !(dbg & 2)) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(__uint_as_float(rr[j]), __uint_as_float(rr[j + 1]),
                                                                     __uint_as_float(rr[j + 2]), __uint_as_float(rr[j + 3]));
          }
        }
      }
      tc::fence_before_sync();
      __syncthreads();   // every warp has drained its TMEM lanes before the next item's first MMA overwrites them
      cur = nxt;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<Cfg::TMEM_COLS>(tacc);
}

// dW[:, col(t) : col(t)+H] (+)= sum of table t's split-K partials in CTA order. grid = (n_tables, H*H/64);
// 4 strided lanes per output element (ascending inside a lane, unrolled loads), combined in fixed order.
template <int H, int kRT, int RG>
__global__ void __launch_bounds__(kFT) fact_dw_reduce_kernel(const __grid_constant__ FactParams p,
                                                             const uint32_t* __restrict__ uniq,
                                                             const int32_t* __restrict__ n_unique_dev,
                                                             const float* __restrict__ dw_part, int rows_grid,
                                                             float* dW_item, float* dW_user) {
  __shared__ float s_p[4][64];
  __shared__ int s_ab[2];
  const int t = blockIdx.x;
  const int U = *n_unique_dev;
  // unique-row range [a, b) of table t: two 32-ary searches on the sorted keys, one warp each (a thread-serial binary
  // search was 2 x 18 dependent L2 round trips = most of this kernel's 21 us, profiles/README.md r2 graph timeline)
  if (threadIdx.x < 64) {
    const int which = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t key = p.key_base[t + which];
    int lo = 0, hi = U;
    while (lo < hi) {
      const int step = (hi - lo + 31) / 32;
      const int i = lo + lane * step;
      const bool lt = i < hi && __ldg(uniq + i) < key;
      const int c = __popc(__ballot_sync(0xffffffffu, lt));
      if (c == 0) break;
      hi = min(hi, lo + c * step);
      lo = lo + (c - 1) * step + 1;
    }
    if (lane == 0) s_ab[which] = lo;
  }
  __syncthreads();
  const int a = s_ab[0], b = s_ab[1];
  if (b <= a) return;
  float* dW = p.side[t] == 0 ? dW_item : dW_user;
  if (dW == nullptr) return;
  const int64_t ld = p.dnn_ld[p.side[t]];
  const int n_tiles = (U + kRT - 1) / kRT;
  const int tpc = (n_tiles + rows_grid - 1) / rows_grid;
  const int cta_a = (a / kRT) / tpc, cta_b = ((b - 1) / kRT) / tpc;
  const int ol = threadIdx.x % 64, pl = threadIdx.x / 64;
  const int i = blockIdx.y * 64 + ol;
  // partial slots of table t: ((cta + t) * RG + rg), cta in [cta_a, cta_b], rg in [0, RG)  => one contiguous run
  const float* src = dw_part + (size_t)t * RG * H * H + i;
  const int ca = cta_a * RG, cb = cta_b * RG + RG - 1;
  float s = 0.f;
  int c = ca + pl;
  for (; c + 12 <= cb; c += 16) {
    const float v0 = src[(size_t)c * H * H], v1 = src[(size_t)(c + 4) * H * H];
    const float v2 = src[(size_t)(c + 8) * H * H], v3 = src[(size_t)(c + 12) * H * H];
    s = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(s, v0), v1), v2), v3);
  }
  for (; c <= cb; c += 4) s = __fadd_rn(s, src[(size_t)c * H * H]);
  s_p[pl][ol] = s;
  __syncthreads();
  if (pl != 0) return;
  s = __fadd_rn(__fadd_rn(s_p[0][ol], s_p[1][ol]), __fadd_rn(s_p[2][ol], s_p[3][ol]));
  const int h = i / H, k = i - h * H;
  float* d = dW + (size_t)h * ld + p.col[t] + k;
  *d = __fadd_rn(*d, s);
}

// ---- per-token fused forward ----------------------------------------------------------------------------------
constexpr int kFTok = 32;   // tokens per CTA tile
constexpr int kFU = 8;      // P rows in flight per lane

struct FwdFactParams {
  const int32_t* ids_u;     // [T, n_single]  1 + unique index, 0 = padding
  const int32_t* arr_off[TGR_MAX_ARRAYS];
  const int32_t* arr_u;     // remapped array values (same indexing as arr_val)
  const float* mmz[6];      // per mm feature: x . Mfold^T + cfold, [T, H]
  const float* bias[2];     // itemdnn.bias, userdnn.bias
  const float* P;           // [U, H]
  float* out;               // [T, H]
  uint8_t* mask;            // [T, H/4]: bit j = z_item[4c+j] > 0, bit 4+j = z_user[4c+j] > 0
  int8_t s_col[TGR_MAX_SLOTS];   // ids column of the SINGLE slots, item side first then user side
  int8_t a_idx[TGR_MAX_ARRAYS];  // array index of the ARRAY slots, item side first then user side
  int32_t n_item_single, n_user_single, n_item_array, n_user_array, n_mm, n_single, T, H4, include_user;
};

__device__ __forceinline__ void gather_sum(float4& z, const int32_t* idr, const int8_t* cols, int s_a, int s_b,
                                           const float4* __restrict__ P4, int H4, int c) {
  for (int s0 = s_a; s0 < s_b; s0 += kFU) {
    float4 v[kFU];
#pragma unroll
    for (int u = 0; u < kFU; ++u) {
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (s0 + u < s_b) {
        const int r = idr[cols[s0 + u]];
        if (r) v[u] = __ldg(P4 + (size_t)(r - 1) * H4 + c);
      }
    }
#pragma unroll
    for (int u = 0; u < kFU; ++u)
      if (s0 + u < s_b) z = f4_add(z, v[u]);   // slot order
  }
}

__device__ __forceinline__ void array_sum(float4& z, const FwdFactParams& p, int a_a, int a_b, int t,
                                          const float4* __restrict__ P4, int H4, int c) {
  for (int a = a_a; a < a_b; ++a) {
    const int ai = p.a_idx[a];
    const int lo = __ldg(p.arr_off[ai] + t), hi = __ldg(p.arr_off[ai] + t + 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = lo; e < hi; ++e) {
      const int r = __ldg(p.arr_u + e);
      if (r) acc = f4_add(acc, __ldg(P4 + (size_t)(r - 1) * H4 + c));
    }
    z = f4_add(z, acc);
  }
}

__device__ __forceinline__ unsigned pos_bits(const float4& z) {
  return (z.x > 0.f ? 1u : 0u) | (z.y > 0.f ? 2u : 0u) | (z.z > 0.f ? 4u : 0u) | (z.w > 0.f ? 8u : 0u);
}
__device__ __forceinline__ float4 relu4(const float4& z) {
  return make_float4(fmaxf(z.x, 0.f), fmaxf(z.y, 0.f), fmaxf(z.z, 0.f), fmaxf(z.w, 0.f));
}

// N compile-time consecutive id columns starting at idr[0]: no bounds checks, no column indirection
template <int N>
__device__ __forceinline__ void gather_sum_fixed(float4& z, const int32_t* idr, const float4* __restrict__ P4, int H4, int c) {
#pragma unroll
  for (int s0 = 0; s0 < N; s0 += kFU) {
    float4 v[kFU];
#pragma unroll
    for (int u = 0; u < kFU; ++u) {
      if (s0 + u < N) {
        const int r = idr[s0 + u];
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r) v[u] = __ldg(P4 + (uint32_t)((r - 1) * H4 + c));
      }
    }
#pragma unroll
    for (int u = 0; u < kFU; ++u)
      if (s0 + u < N) z = f4_add(z, v[u]);   // slot order
  }
}

// NSI / NSU > 0: the call's SINGLE slots are ids columns [0, NSI) (item side) and [NSI, NSI + NSU) (user side) — the
// reference's default feature lists (dataset.py:191-212) give (15, 0) and (15, 5); NSI < 0: counts / columns at run time.
template <int LANES, int NSI, int NSU>
__global__ void __launch_bounds__(kFT) fact_forward_kernel(const __grid_constant__ FwdFactParams p) {
  extern __shared__ int32_t s_ids[];   // [kFTok * n_single]
  constexpr int G = kFT / LANES;
  const int tid = threadIdx.x, lane = tid % LANES, grp = tid / LANES;
  const int H4 = p.H4;
  const int t0 = blockIdx.x * kFTok;
  const int nt = min(kFTok, p.T - t0);
  {
    const int n = nt * p.n_single;
    const int32_t* src = p.ids_u + (size_t)t0 * p.n_single;   // 16 B aligned: kFTok * n_single * 4 % 16 == 0
    const int n4 = n >> 2;
    const int4* src4 = reinterpret_cast<const int4*>(src);
    int4* dst4 = reinterpret_cast<int4*>(s_ids);
    for (int i = tid; i < n4; i += kFT) dst4[i] = __ldg(src4 + i);
    for (int i = (n4 << 2) + tid; i < n; i += kFT) s_ids[i] = __ldg(src + i);
  }
  __syncthreads();
  const float4* P4 = reinterpret_cast<const float4*>(p.P);
  const int ns_i = p.n_item_single, ns = ns_i + p.n_user_single;
  const int na_i = p.n_item_array, na = na_i + p.n_user_array;
  const bool user = NSI >= 0 ? NSU > 0 : p.include_user != 0;
  for (int tl = grp; tl < nt; tl += G) {
    const int t = t0 + tl;
    const int32_t* idr = s_ids + tl * p.n_single;
    for (int c = lane; c < H4; c += LANES) {
      float4 zi = __ldg(reinterpret_cast<const float4*>(p.bias[0]) + c);
      if constexpr (NSI >= 0) gather_sum_fixed<NSI>(zi, idr, P4, H4, c);
      else gather_sum(zi, idr, p.s_col, 0, ns_i, P4, H4, c);
      array_sum(zi, p, 0, na_i, t, P4, H4, c);
      for (int f = 0; f < p.n_mm; ++f)
        zi = f4_add(zi, ld_stream(reinterpret_cast<const float4*>(p.mmz[f]) + (size_t)t * H4 + c));
      unsigned bits = pos_bits(zi);
      float4 o = relu4(zi);
      if (user) {
        float4 zu = __ldg(reinterpret_cast<const float4*>(p.bias[1]) + c);
        if constexpr (NSI >= 0) gather_sum_fixed<(NSU > 0 ? NSU : 1)>(zu, idr + NSI, P4, H4, c);
        else gather_sum(zu, idr, p.s_col, ns_i, ns, P4, H4, c);
        array_sum(zu, p, na_i, na, t, P4, H4, c);
        bits |= pos_bits(zu) << 4;
        o = f4_add(o, relu4(zu));
      }
      st_stream(reinterpret_cast<float4*>(p.out) + (size_t)t * H4 + c, o);
      p.mask[(size_t)t * H4 + c] = (uint8_t)bits;
    }
  }
}

// dZ_side = dOut * mask_side, per-CTA column partial sums, and (MM) the CTA's partial of A = dZ_item^T x for ONE
// 32-wide mm feature — all from one pass over dOut (fixed token chunks, fixed order => bitwise reproducible).
// thread = (column c = tid % H4, row lane rl = tid / H4); CTA b owns tokens [b*chunk, (b+1)*chunk), chunk % kDzTok == 0
constexpr int kDzTok = 64;   // tokens per smem tile
constexpr int kDzMM = 32;    // mm width the fused A accumulation supports ('81', model.py:183)

template <int H, bool MM, bool XBF16>
__global__ void __launch_bounds__(kFT) fact_dz_kernel(const float4* __restrict__ d_out, const uint8_t* __restrict__ mask,
                                                      float4* __restrict__ dzi, float4* __restrict__ dzu,
                                                      const void* __restrict__ x, int T, int chunk,
                                                      float* __restrict__ part /* [grid][2H (+ H*32)] */) {
  constexpr int H4 = H / 4, RL = kFT / H4, NJ = kDzTok / RL, LD = H + 4, XLD = kDzMM + 4, TPW = H / 32;
  constexpr int PART = 2 * H + (MM ? H * kDzMM : 0);
  extern __shared__ __align__(16) float dsm[];
  float* Ds = dsm;                      // [kDzTok][LD]   (MM)
  float* Xs = Ds + kDzTok * LD;         // [kDzTok][XLD]  (MM)
  __shared__ float4 s_acc[2][kFT];
  const int tid = threadIdx.x, c = tid % H4, rl = tid / H4;
  // A = dZ_item^T x on the tensor cores (3xTF32, tgr_mma.cuh): the [H x 32] block is (H/16) x 4 m16n8 tiles, TPW per warp
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const int ta = blockIdx.x * chunk, tb = min(T, ta + chunk);
  float4 si = make_float4(0.f, 0.f, 0.f, 0.f), su = si;
  float acc[TPW][4];
#pragma unroll
  for (int i = 0; i < TPW; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  // software pipeline: the loads of tile i + 1 are issued right after tile i has been staged, so they travel during the
  // barrier + MMA phase instead of in front of it
  constexpr int XPT = kDzTok * (kDzMM / 4) / kFT;
  float4 d[NJ];
  unsigned m[NJ];
  float4 xv[MM ? XPT : 1];
  auto load_tile = [&](int t0) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int t = t0 + rl + RL * j;
      d[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      m[j] = 0;
      if (t < tb) {
        const size_t i = (size_t)t * H4 + c;
        d[j] = ld_stream(d_out + i);
        m[j] = mask[i];
      }
    }
    if (MM) {
#pragma unroll
      for (int q = 0; q < XPT; ++q) {
        const int i = tid + q * kFT, r = i / (kDzMM / 4), c4 = i % (kDzMM / 4);
        xv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t0 + r < tb) {
          const size_t e = (size_t)(t0 + r) * kDzMM + c4 * 4;
          if constexpr (XBF16) xv[q] = unpack_bf16x4(__ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(x) + e)));
          else xv[q] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + e));
        }
      }
    }
  };
  if (ta < tb) load_tile(ta);
  for (int t0 = ta; t0 < tb; t0 += kDzTok) {
    if (MM) {
#pragma unroll
      for (int q = 0; q < XPT; ++q) {
        const int i = tid + q * kFT, r = i / (kDzMM / 4), c4 = i % (kDzMM / 4);
        *reinterpret_cast<float4*>(Xs + r * XLD + c4 * 4) = xv[q];
      }
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int t = t0 + rl + RL * j;
      const float4 dd = d[j];
      const unsigned mm = m[j];
      const float4 a = make_float4(mm & 1u ? dd.x : 0.f, mm & 2u ? dd.y : 0.f, mm & 4u ? dd.z : 0.f, mm & 8u ? dd.w : 0.f);
      if (MM) *reinterpret_cast<float4*>(Ds + (rl + RL * j) * LD + c * 4) = a;   // zero beyond tb
      if (t < tb) {
        const size_t i = (size_t)t * H4 + c;
        dzi[i] = a;
        si = f4_add(si, a);
        if (dzu != nullptr) {
          const float4 b = make_float4(mm & 16u ? dd.x : 0.f, mm & 32u ? dd.y : 0.f, mm & 64u ? dd.z : 0.f, mm & 128u ? dd.w : 0.f);
          dzu[i] = b;
          su = f4_add(su, b);
        }
      }
    }
    if (t0 + kDzTok < tb) load_tile(t0 + kDzTok);
    if (MM) {
      __syncthreads();
#pragma unroll 2
      for (int ks = 0; ks < kDzTok / 8; ++ks) {
#pragma unroll
        for (int i = 0; i < TPW; ++i) {
          const int q = warp * TPW + i, mt = q >> 2, nt = q & 3;
          const float* da = Ds + (8 * ks + t4) * LD + 16 * mt + g;     // A[m = h][k = token] = Ds[token][h]
          const float* xb = Xs + (8 * ks + t4) * XLD + 8 * nt + g;     // B[k = token][n = j] = Xs[token][j]
          uint32_t ahi[4], alo[4], bhi[2], blo[2];
          split_tf32(da[0], ahi[0], alo[0]);
          split_tf32(da[8], ahi[1], alo[1]);
          split_tf32(da[4 * LD], ahi[2], alo[2]);
          split_tf32(da[4 * LD + 8], ahi[3], alo[3]);
          split_tf32(xb[0], bhi[0], blo[0]);
          split_tf32(xb[4 * XLD], bhi[1], blo[1]);
          mma_3xtf32(acc[i], ahi, alo, bhi, blo);
        }
      }
      __syncthreads();
    }
  }
  s_acc[0][tid] = si;
  s_acc[1][tid] = su;
  __syncthreads();
  float* dst = part + (size_t)blockIdx.x * PART;
  if (tid < 2 * H4) {
    const int side = tid / H4, cc = tid % H4;
    float4 s = s_acc[side][cc];
    for (int r = 1; r < RL; ++r) s = f4_add(s, s_acc[side][r * H4 + cc]);
    *reinterpret_cast<float4*>(dst + side * H + cc * 4) = s;
  }
  if (MM) {
#pragma unroll
    for (int i = 0; i < TPW; ++i) {
      const int q = warp * TPW + i, mt = q >> 2, nt = q & 3;
      float* d0 = dst + 2 * H + (16 * mt + g) * kDzMM + 8 * nt + 2 * t4;
      *reinterpret_cast<float2*>(d0) = make_float2(acc[i][0], acc[i][1]);
      *reinterpret_cast<float2*>(d0 + 8 * kDzMM) = make_float2(acc[i][2], acc[i][3]);
    }
  }
}

// Ordered sum of the per-CTA partial vectors: output o = sum_b part[b][o]; 16 strided lanes per output (b ascending
// inside a lane, 4 independent loads in flight), lanes combined in fixed order.
// o < H: db_item += , mm_s += ; o < 2H: db_user += ; else mm_A += .
__global__ void __launch_bounds__(256) fact_dz_finish_kernel(const float* __restrict__ part, int n_part, int part_ld, int H,
                                                             float* db_item, float* db_user, float* mm_A, float* mm_s) {
  __shared__ float s_p[16][17];
  const int ol = threadIdx.x % 16, pl = threadIdx.x / 16;
  const int o = blockIdx.x * 16 + ol;
  float s = 0.f;
  if (o < part_ld) {
    int b = pl;
    for (; b + 48 < n_part; b += 64) {
      const float v0 = part[(size_t)b * part_ld + o], v1 = part[(size_t)(b + 16) * part_ld + o];
      const float v2 = part[(size_t)(b + 32) * part_ld + o], v3 = part[(size_t)(b + 48) * part_ld + o];
      s = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(s, v0), v1), v2), v3);
    }
    for (; b < n_part; b += 16) s = __fadd_rn(s, part[(size_t)b * part_ld + o]);
  }
  s_p[pl][ol] = s;
  __syncthreads();
  if (pl != 0 || o >= part_ld) return;
  s = s_p[0][ol];
#pragma unroll
  for (int l = 1; l < 16; ++l) s = __fadd_rn(s, s_p[l][ol]);
  if (o < H) {
    if (db_item) db_item[o] = __fadd_rn(db_item[o], s);
    if (mm_s) mm_s[o] = __fadd_rn(mm_s[o], s);
  } else if (o < 2 * H) {
    if (db_user) db_user[o - H] = __fadd_rn(db_user[o - H], s);
  } else if (mm_A) {
    mm_A[o - 2 * H] = __fadd_rn(mm_A[o - 2 * H], s);
  }
}

// ---- mm features folded through the item DNN -------------------------------------------------------------------
// M[h][j] = sum_q Ws[h][q] Wmm[q][j] ; c[h] = sum_q Ws[h][q] bmm[q]         (Ws = itemdnn.weight[:, col:col+H])
__global__ void __launch_bounds__(256) fact_mm_fold_kernel(const float* __restrict__ Ws, int64_t ld, const float* __restrict__ Wmm,
                                                           const float* __restrict__ bmm, int H, int mm_dim,
                                                           float* __restrict__ M, float* __restrict__ cvec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < H * mm_dim) {
    const int h = i / mm_dim, j = i - h * mm_dim;
    float s = 0.f;
#pragma unroll 8
    for (int q = 0; q < H; ++q) s = fmaf(__ldg(Ws + (size_t)h * ld + q), __ldg(Wmm + (size_t)q * mm_dim + j), s);
    M[i] = s;
  } else if (i < H * mm_dim + H) {
    const int h = i - H * mm_dim;
    float s = 0.f;
    if (bmm != nullptr)
      for (int q = 0; q < H; ++q) s = fmaf(__ldg(Ws + (size_t)h * ld + q), __ldg(bmm + q), s);
    cvec[h] = s;
  }
}

// From A = dZ^T x [H, mm_dim] and s = colsum(dZ) [H]:
//   dWmm[q][j] += sum_h Ws[h][q] A[h][j] ; dbmm[q] += sum_h Ws[h][q] s[h] ; dWs[h][q] += sum_j A[h][j] Wmm[q][j] + s[h] bmm[q]
__global__ void __launch_bounds__(256) fact_mm_chain_kernel(const float* __restrict__ Ws, int64_t ld, const float* __restrict__ Wmm,
                                                            const float* __restrict__ bmm, const float* __restrict__ A,
                                                            const float* __restrict__ s, int H, int mm_dim, float* dWmm,
                                                            float* dbmm, float* dWs, int64_t dld) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n1 = H * mm_dim, n2 = n1 + H, n3 = n2 + H * H;
  if (i < n1) {
    const int q = i / mm_dim, j = i - q * mm_dim;
    float a = 0.f;
#pragma unroll 8
    for (int h = 0; h < H; ++h) a = fmaf(__ldg(Ws + (size_t)h * ld + q), __ldg(A + (size_t)h * mm_dim + j), a);
    dWmm[i] = __fadd_rn(dWmm[i], a);
  } else if (i < n2) {
    const int q = i - n1;
    if (dbmm != nullptr) {
      float a = 0.f;
      for (int h = 0; h < H; ++h) a = fmaf(__ldg(Ws + (size_t)h * ld + q), __ldg(s + h), a);
      dbmm[q] = __fadd_rn(dbmm[q], a);
    }
  }
}

// dWs[h][q] += sum_j A[h][j] Wmm[q][j] + s[h] bmm[q]: one warp per output, lanes stride j (both rows coalesced), fixed
// shuffle tree. (One THREAD per output walked Wmm with a stride of mm_dim floats: 0.15 ms at mm_dim = 1024.)
__global__ void __launch_bounds__(256) fact_mm_chain_ws_kernel(const float* __restrict__ Wmm, const float* __restrict__ bmm,
                                                               const float* __restrict__ A, const float* __restrict__ s, int H,
                                                               int mm_dim, float* dWs, int64_t dld) {
  const int lane = threadIdx.x & 31;
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (e >= H * H) return;
  const int h = e / H, q = e - h * H;
  float a = 0.f;
  for (int j = lane; j < mm_dim; j += 32) a = fmaf(__ldg(A + (size_t)h * mm_dim + j), __ldg(Wmm + (size_t)q * mm_dim + j), a);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) {
    if (bmm != nullptr) a = fmaf(__ldg(s + h), __ldg(bmm + q), a);
    float* d = dWs + (size_t)h * dld + q;
    *d = __fadd_rn(*d, a);
  }
}

static int fill_fact(FactParams& p, const tgr_table_t* tables, int n_tables, int H, const tgr_dnn_t* dnn,
                     const tgr_row_source_t* src) {
  const float* fetched = src ? src->fetched_rows : nullptr;
  const int32_t* fetched_perm = src ? src->fetched_perm : nullptr;
  const int n_peers = src ? src->n_peers : 0;
  TGR_REQUIRE(n_peers >= 0 && n_peers <= TGR_MAX_PEERS, "n_peers out of range");
  TGR_REQUIRE(!(n_peers > 0 && fetched), "row source: peers and fetched rows are exclusive");
  p.n_peers = n_peers;
  for (int r = 0; r < n_peers; ++r) {
    TGR_REQUIRE(src->peer_rows[r] != nullptr, "peer %d: shard pointer is NULL", r);
    p.peer[r] = src->peer_rows[r];
  }
  p.save_rows = src ? src->save_rows : nullptr;
  TGR_REQUIRE(tables && n_tables > 0 && n_tables <= TGR_MAX_TABLES, "bad table array");
  TGR_REQUIRE(H == 32 || H == 64 || H == 128, "the factored path supports H in {32, 64, 128} (H=%d)", H);
  TGR_REQUIRE(dnn && dnn->w_item, "dnn / itemdnn weight is NULL");
  p.n_tables = n_tables;
  for (int t = 0; t < n_tables; ++t) {
    p.w[t] = tables[t].weight;
    TGR_REQUIRE(p.w[t] != nullptr || fetched != nullptr || n_peers > 0, "table %d: weight NULL", t);
    p.key_base[t] = (uint32_t)tables[t].key_base;
    p.side[t] = (int8_t)dnn->table_side[t];
    p.col[t] = dnn->table_col[t];
    TGR_REQUIRE(p.side[t] == 0 || (p.side[t] == 1 && dnn->w_user), "table %d: bad side %d", t, (int)p.side[t]);
    const int64_t ld = p.side[t] == 0 ? dnn->item_ld : dnn->user_ld;
    TGR_REQUIRE(p.col[t] >= 0 && p.col[t] + H <= ld, "table %d: DNN columns out of range", t);
    TGR_REQUIRE(p.col[t] % 4 == 0 && ld % 4 == 0, "table %d: DNN column / pitch not 128-bit tileable", t);
    if (t) TGR_REQUIRE(tables[t].key_base == tables[t - 1].key_base + tables[t - 1].rows, "key bases must be cumulative");
  }
  p.key_base[n_tables] = (uint32_t)(tables[n_tables - 1].key_base + tables[n_tables - 1].rows);
  TGR_REQUIRE(fetched != nullptr || fetched_perm == nullptr, "a permutation without fetched rows");
  p.fetched = fetched;
  p.fetched_perm = fetched_perm;
  TGR_REQUIRE(((uintptr_t)dnn->w_item & 15) == 0 && ((uintptr_t)dnn->w_user & 15) == 0, "DNN weights must be 16-byte aligned");
  p.dnn_w[0] = dnn->w_item; p.dnn_ld[0] = dnn->item_ld;
  p.dnn_w[1] = dnn->w_user; p.dnn_ld[1] = dnn->user_ld;
  return 0;
}

static int rows_rt(int H) { return H == 64 ? 64 : 128; }
static size_t fact_smem(int H, bool bwd) { return (size_t)(H * (H + 4) + (bwd ? 2 : 1) * rows_rt(H) * (H + 4)) * sizeof(float); }
static int rows_rg(int H) { return (H / 8) * (rows_rt(H) / 8) / ((H / 8) * (H / 8)); }
static size_t mma_smem(int H, bool bwd) { return (size_t)(2 * H * (H + 4) + (bwd ? 2 : 1) * 128 * (H + 4)) * sizeof(float); }
// TGR_ROWS_FFMA=1 selects the fp32 CUDA-core kernels for every H (A/B timing, tools/rows_bench.py)
static bool rows_use_mma(int H) {
  static const bool ffma = [] { const char* e = getenv("TGR_ROWS_FFMA"); return e && e[0] == '1'; }();
  return (H == 32 || H == 64) && !ffma;
}

template <int H, int MODE>
static int launch_rows(const FactParams& p, const uint32_t* uniq, const int32_t* n_unique_dev, float* PG, float* dw_part,
                       cudaStream_t st) {
  const size_t smem = fact_smem(H, MODE == 1);
  { static bool tgr_attr_once_ = false; if (!tgr_attr_once_) { cudaFuncSetAttribute(fact_rows_kernel<H, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); tgr_attr_once_ = true; } }
  TGR_K(fact_rows_kernel<H, MODE>)<<<MODE ? kRowsGridBwd : kRowsGridFwd, RowsCfg<H>::NT, smem, st>>>(p, uniq, n_unique_dev, PG, dw_part);
  return check_launch(MODE ? "fact_unique_backward" : "fact_project_rows");
}

// TGR_ROWS_TC=0 keeps the forward projection on mma.sync (A/B timing)
static bool rows_use_tc(int H) {
  static const bool off = [] { const char* e = getenv("TGR_ROWS_TC"); return e && e[0] == '0'; }();
  return rows_use_mma(H) && !off;
}

template <int H>
static int launch_rows_tc(const FactParams& p, const uint32_t* uniq, const int32_t* n_unique_dev, float* P, cudaStream_t st) {
  const size_t smem = TcRowsCfg<H>::SMEM;
  { static bool tgr_attr_once_ = false; if (!tgr_attr_once_) { cudaFuncSetAttribute(fact_rows_tc_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); tgr_attr_once_ = true; } }
  static const int dbg = [] { const char* e = getenv("TGR_TC_DBG"); return e ? atoi(e) : 0; }();   // dev: 1 no loads, 2 no stores, 4 no MMA
  TGR_K(fact_rows_tc_kernel<H>)<<<kTcGridFwd, TcRowsCfg<H>::NT, smem, st>>>(p, uniq, n_unique_dev, P, dbg);
  return check_launch("fact_project_rows");
}

template <int H, int MODE>
static int launch_rows_mma(const FactParams& p, const uint32_t* uniq, const int32_t* n_unique_dev, float* PG, float* dw_part,
                           cudaStream_t st) {
  const size_t smem = mma_smem(H, MODE == 1);
  { static bool tgr_attr_once_ = false; if (!tgr_attr_once_) { cudaFuncSetAttribute(fact_rows_mma_kernel<H, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); tgr_attr_once_ = true; } }
  TGR_K(fact_rows_mma_kernel<H, MODE>)<<<MODE ? kMmaGridBwd : kMmaGridFwd, MmaRowsCfg<H>::NT, smem, st>>>(p, uniq, n_unique_dev, PG, dw_part);
  return check_launch(MODE ? "fact_unique_backward" : "fact_project_rows");
}

}  // namespace tgr

using namespace tgr;

extern "C" int tgr_fact_project_rows(const tgr_table_t* tables, int n_tables, int H, const tgr_dnn_t* dnn,
                                     const uint32_t* uniq, const int32_t* n_unique_dev, int64_t max_unique,
                                     const tgr_row_source_t* src, float* P, void* stream) {
  tgr::TimedScope tgr_timed_("fact_project_rows", stream);
  FactParams p{};
  if (int rc = fill_fact(p, tables, n_tables, H, dnn, src)) return rc;
  TGR_REQUIRE(uniq && n_unique_dev && P, "null argument");
  if (max_unique <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (rows_use_mma(H) && rows_ws_supported(H)) return launch_rows_ws_fwd(p, uniq, n_unique_dev, P, st);
  if (rows_use_tc(H)) {
    if (H == 32) return launch_rows_tc<32>(p, uniq, n_unique_dev, P, st);
    return launch_rows_tc<64>(p, uniq, n_unique_dev, P, st);
  }
  if (rows_use_mma(H)) {
    if (H == 32) return launch_rows_mma<32, 0>(p, uniq, n_unique_dev, P, nullptr, st);
    return launch_rows_mma<64, 0>(p, uniq, n_unique_dev, P, nullptr, st);
  }
  if (H == 32) return launch_rows<32, 0>(p, uniq, n_unique_dev, P, nullptr, st);
  if (H == 64) return launch_rows<64, 0>(p, uniq, n_unique_dev, P, nullptr, st);
  return launch_rows<128, 0>(p, uniq, n_unique_dev, P, nullptr, st);
}

extern "C" size_t tgr_fact_backward_workspace_bytes(int n_tables, int H) {
  return (size_t)(kRowsGridBwd + n_tables + 1) * rows_rg(H) * H * H * sizeof(float);
}

extern "C" int tgr_fact_unique_backward(const tgr_table_t* tables, int n_tables, int H, const tgr_dnn_t* dnn,
                                        const uint32_t* uniq, const int32_t* n_unique_dev, int64_t max_unique,
                                        const tgr_row_source_t* src, float* G, float* dW_item, float* dW_user,
                                        void* workspace, size_t workspace_bytes, void* stream) {
  tgr::TimedScope tgr_timed_("fact_unique_backward", stream);
  FactParams p{};
  if (int rc = fill_fact(p, tables, n_tables, H, dnn, src)) return rc;
  TGR_REQUIRE(uniq && n_unique_dev && G && workspace, "null argument");
  TGR_REQUIRE(workspace_bytes >= tgr_fact_backward_workspace_bytes(n_tables, H), "workspace too small");
  if (max_unique <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  float* part = (float*)workspace;
  int rc;
  const bool mma = rows_use_mma(H);
  if (mma && rows_ws_supported(H)) {     // warp-specialised tcgen05 kernel (tgr_rows_ws.cu), 96-row tiles, one CTA per SM
    if (int rc2 = launch_rows_ws_bwd(p, uniq, n_unique_dev, G, part, st)) return rc2;
    if (dW_item == nullptr && dW_user == nullptr) return 0;
    const dim3 grid_ws(n_tables, H * H / 64);
    TGR_K(fact_dw_reduce_kernel<64, 96, 1>)<<<grid_ws, kFT, 0, st>>>(p, uniq, n_unique_dev, part, rows_ws_bwd_grid(), dW_item, dW_user);
    return check_launch("fact_dw_reduce");
  }
  if (mma) rc = H == 32 ? launch_rows_mma<32, 1>(p, uniq, n_unique_dev, G, part, st) : launch_rows_mma<64, 1>(p, uniq, n_unique_dev, G, part, st);
  else if (H == 32) rc = launch_rows<32, 1>(p, uniq, n_unique_dev, G, part, st);
  else if (H == 64) rc = launch_rows<64, 1>(p, uniq, n_unique_dev, G, part, st);
  else rc = launch_rows<128, 1>(p, uniq, n_unique_dev, G, part, st);
  if (rc) return rc;
  if (dW_item == nullptr && dW_user == nullptr) return 0;
  const dim3 grid(n_tables, H * H / 64);
  if (mma && H == 32) TGR_K(fact_dw_reduce_kernel<32, 128, 1>)<<<grid, kFT, 0, st>>>(p, uniq, n_unique_dev, part, kMmaGridBwd, dW_item, dW_user);
  else if (mma) TGR_K(fact_dw_reduce_kernel<64, 128, 1>)<<<grid, kFT, 0, st>>>(p, uniq, n_unique_dev, part, kMmaGridBwd, dW_item, dW_user);
  else if (H == 32) TGR_K(fact_dw_reduce_kernel<32, RowsCfg<32>::RT, RowsCfg<32>::RG>)<<<grid, kFT, 0, st>>>(p, uniq, n_unique_dev, part, kRowsGridBwd, dW_item, dW_user);
  else if (H == 64) TGR_K(fact_dw_reduce_kernel<64, RowsCfg<64>::RT, RowsCfg<64>::RG>)<<<grid, kFT, 0, st>>>(p, uniq, n_unique_dev, part, kRowsGridBwd, dW_item, dW_user);
  else TGR_K(fact_dw_reduce_kernel<128, RowsCfg<128>::RT, RowsCfg<128>::RG>)<<<grid, kFT, 0, st>>>(p, uniq, n_unique_dev, part, kRowsGridBwd, dW_item, dW_user);
  return check_launch("fact_dw_reduce");
}

extern "C" int tgr_fact_forward(const tgr_call_t* call, int H, const int32_t* ids_u, const int32_t* arr_u, const float* P,
                                const float* const* mmz, int n_mm, const float* bias_item, const float* bias_user,
                                float* out, uint8_t* mask, void* stream) {
  tgr::TimedScope tgr_timed_("fact_forward", stream);
  TGR_REQUIRE(call && P && bias_item && out && mask, "null argument");
  TGR_REQUIRE(H == 32 || H == 64 || H == 128, "the factored path supports H in {32, 64, 128}");
  TGR_REQUIRE(n_mm >= 0 && n_mm <= 6, "n_mm out of range");
  TGR_REQUIRE(call->T >= 0 && call->n_slots >= 0 && call->n_slots <= TGR_MAX_SLOTS, "bad call");
  if (call->T == 0) return 0;
  FwdFactParams p{};
  p.ids_u = ids_u; p.arr_u = arr_u; p.P = P; p.out = out; p.mask = mask;
  p.bias[0] = bias_item; p.bias[1] = bias_user;
  p.T = call->T; p.H4 = H / 4; p.n_single = call->n_single; p.n_mm = n_mm;
  for (int f = 0; f < n_mm; ++f) { TGR_REQUIRE(mmz && mmz[f], "mmz[%d] NULL", f); p.mmz[f] = mmz[f]; }
  int ns = 0, na = 0;
  bool user = false;
  for (int side = 0; side < 2; ++side) {
    for (int i = 0; i < call->n_slots; ++i) {
      const tgr_slot_t& s = call->slots[i];
      if (s.side != side) continue;
      if (s.kind == TGR_KIND_SINGLE) {
        TGR_REQUIRE(s.src >= 0 && s.src < call->n_single, "slot %d: ids column out of range", i);
        p.s_col[ns++] = (int8_t)s.src;
        (side == 0 ? p.n_item_single : p.n_user_single)++;
      } else if (s.kind == TGR_KIND_ARRAY) {
        TGR_REQUIRE(s.src >= 0 && s.src < call->n_arrays && s.src < TGR_MAX_ARRAYS, "slot %d: array index out of range", i);
        TGR_REQUIRE(call->arr_off[s.src] != nullptr && (arr_u != nullptr || call->arr_nnz[s.src] == 0), "array pointers NULL");
        p.a_idx[na++] = (int8_t)s.src;
        p.arr_off[s.src] = call->arr_off[s.src];
        (side == 0 ? p.n_item_array : p.n_user_array)++;
      } else {
        TGR_REQUIRE(s.kind == TGR_KIND_MM && side == 0, "slot %d: bad kind/side", i);
      }
      if (side == 1) user = true;
    }
  }
  TGR_REQUIRE(ns == 0 || ids_u != nullptr, "ids_u is NULL");
  TGR_REQUIRE(((uintptr_t)ids_u & 15) == 0, "ids_u must be 16-byte aligned");
  p.include_user = user ? 1 : 0;
  TGR_REQUIRE(!user || bias_user, "user side needs the userdnn bias");
  const int grid = (call->T + kFTok - 1) / kFTok;
  const size_t smem = (size_t)kFTok * (p.n_single > 0 ? p.n_single : 1) * sizeof(int32_t);
  cudaStream_t st = (cudaStream_t)stream;
  // identity column layout with the reference's default slot counts -> fully unrolled gather loops
  bool ident = true;   // (the fixed variant indexes P with 32-bit float4 offsets: unique rows * H/4 < 2^31)
  for (int i = 0; i < ns; ++i) ident = ident && p.s_col[i] == i;
  const int nsi = p.n_item_single, nsu = p.n_user_single;
#define TGR_FWD(L)                                                                                         \
  do {                                                                                                     \
    if (ident && nsi == 15 && nsu == 0) TGR_K(fact_forward_kernel<L, 15, 0>)<<<grid, kFT, smem, st>>>(p);        \
    else if (ident && nsi == 15 && nsu == 5) TGR_K(fact_forward_kernel<L, 15, 5>)<<<grid, kFT, smem, st>>>(p);   \
    else TGR_K(fact_forward_kernel<L, -1, 0>)<<<grid, kFT, smem, st>>>(p);                                       \
  } while (0)
  if (H == 32) TGR_FWD(8);
  else if (H == 64) TGR_FWD(16);
  else TGR_FWD(32);
#undef TGR_FWD
  return check_launch("fact_forward");
}

static int relu_grid(int64_t T, int* chunk) {
  int64_t g = (T + 4 * kDzTok - 1) / (4 * kDzTok);   // >= 256 tokens per CTA
  if (g > 4 * kNumSMs) g = 4 * kNumSMs;
  if (g < 1) g = 1;
  int64_t ch = (T + g - 1) / g;
  ch = (ch + kDzTok - 1) / kDzTok * kDzTok;
  if (ch < kDzTok) ch = kDzTok;
  *chunk = (int)ch;
  return (int)((T + ch - 1) / ch);
}

extern "C" size_t tgr_fact_relu_mask_workspace_bytes(int64_t T, int H) {
  int chunk;
  return (size_t)relu_grid(T, &chunk) * (2 * H + H * kDzMM) * sizeof(float) + 256;
}

template <int H>
static int launch_dz(const float* d_out, const uint8_t* mask, int T, float* dz_item, float* dz_user, const void* mm_x,
                     int mm_x_dtype, float* part, int grid, int chunk, cudaStream_t st) {
  const size_t smem = (size_t)(kDzTok * (H + 4) + kDzTok * (kDzMM + 4)) * sizeof(float);
  if (mm_x == nullptr) {
    TGR_K(fact_dz_kernel<H, false, false>)<<<grid, kFT, 0, st>>>((const float4*)d_out, mask, (float4*)dz_item, (float4*)dz_user,
                                                         nullptr, T, chunk, part);
  } else if (mm_x_dtype == TGR_DTYPE_BF16) {
    { static bool tgr_attr_once_ = false; if (!tgr_attr_once_) { cudaFuncSetAttribute(fact_dz_kernel<H, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); tgr_attr_once_ = true; } }
    TGR_K(fact_dz_kernel<H, true, true>)<<<grid, kFT, smem, st>>>((const float4*)d_out, mask, (float4*)dz_item, (float4*)dz_user,
                                                          mm_x, T, chunk, part);
  } else {
    { static bool tgr_attr_once_ = false; if (!tgr_attr_once_) { cudaFuncSetAttribute(fact_dz_kernel<H, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); tgr_attr_once_ = true; } }
    TGR_K(fact_dz_kernel<H, true, false>)<<<grid, kFT, smem, st>>>((const float4*)d_out, mask, (float4*)dz_item, (float4*)dz_user,
                                                           mm_x, T, chunk, part);
  }
  return check_launch("fact_dz");
}

extern "C" int tgr_fact_relu_mask(const float* d_out, const uint8_t* mask, int64_t T, int H, float* dz_item, float* dz_user,
                                  float* db_item, float* db_user, const void* mm_x, int mm_x_dtype, int mm_dim, float* mm_A,
                                  float* mm_s, void* workspace, size_t workspace_bytes, void* stream) {
  tgr::TimedScope tgr_timed_("fact_relu_mask", stream);
  TGR_REQUIRE(H == 32 || H == 64 || H == 128, "the factored path supports H in {32, 64, 128}");
  TGR_REQUIRE(T >= 0 && T < (1ll << 31), "T out of range");
  if (T == 0) return 0;
  TGR_REQUIRE(d_out && mask && dz_item && workspace, "null argument");
  TGR_REQUIRE(workspace_bytes >= tgr_fact_relu_mask_workspace_bytes(T, H), "workspace too small");
  TGR_REQUIRE(mm_x == nullptr || (mm_dim == kDzMM && mm_A && mm_s), "the fused A = dz^T x supports mm_dim == 32 only");
  int chunk;
  const int grid = relu_grid(T, &chunk);
  cudaStream_t st = (cudaStream_t)stream;
  float* part = (float*)workspace;
  int rc;
  if (H == 32) rc = launch_dz<32>(d_out, mask, (int)T, dz_item, dz_user, mm_x, mm_x_dtype, part, grid, chunk, st);
  else if (H == 64) rc = launch_dz<64>(d_out, mask, (int)T, dz_item, dz_user, mm_x, mm_x_dtype, part, grid, chunk, st);
  else rc = launch_dz<128>(d_out, mask, (int)T, dz_item, dz_user, mm_x, mm_x_dtype, part, grid, chunk, st);
  if (rc) return rc;
  if (db_item == nullptr && db_user == nullptr && mm_x == nullptr) return 0;
  const int part_ld = 2 * H + (mm_x ? H * kDzMM : 0);
  TGR_K(fact_dz_finish_kernel)<<<(part_ld + 15) / 16, 256, 0, st>>>(part, grid, part_ld, H, db_item, dz_user ? db_user : nullptr,
                                                           mm_x ? mm_A : nullptr, mm_x ? mm_s : nullptr);
  return check_launch("fact_dz_finish");
}

extern "C" int tgr_fact_mm_fold(const float* w_slot, int64_t ld, const float* w_mm, const float* b_mm, int H, int mm_dim,
                                float* M, float* c, void* stream) {
  tgr::TimedScope tgr_timed_("fact_mm_fold", stream);
  TGR_REQUIRE(w_slot && w_mm && M && c && H > 0 && mm_dim > 0, "bad argument");
  const int n = H * mm_dim + H;
  TGR_K(fact_mm_fold_kernel)<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w_slot, ld, w_mm, b_mm, H, mm_dim, M, c);
  return check_launch("fact_mm_fold");
}

extern "C" int tgr_fact_mm_chain_bwd(const float* w_slot, int64_t ld, const float* w_mm, const float* b_mm, const float* A,
                                     const float* s, int H, int mm_dim, float* dW_mm, float* db_mm, float* dW_slot,
                                     int64_t dld, void* stream) {
  tgr::TimedScope tgr_timed_("fact_mm_chain_bwd", stream);
  TGR_REQUIRE(w_slot && w_mm && A && s && dW_mm && dW_slot && H > 0 && mm_dim > 0, "bad argument");
  const int n = H * mm_dim + H;
  TGR_K(fact_mm_chain_kernel)<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w_slot, ld, w_mm, b_mm, A, s, H, mm_dim, dW_mm,
                                                                          db_mm, dW_slot, dld);
  TGR_K(fact_mm_chain_ws_kernel)<<<(H * H * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w_mm, b_mm, A, s, H, mm_dim, dW_slot, dld);
  return check_launch("fact_mm_chain_bwd");
}
