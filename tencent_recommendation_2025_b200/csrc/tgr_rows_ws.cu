// Warp-specialised tcgen05 kernels for the two row GEMMs of the factored path (H = 64):
//
//   forward  (MODE 0)  P[u]     = W_s . row[u]                         once per unique row of the step
//   backward (MODE 1)  g_row[u] = G[u] . W_s   (in place over G)       and   dW_s += sum_u G[u]^T (x) row[u]
//
// Why this shape. The first tcgen05 version (fact_rows_tc_kernel, tgr_factored.cu) ran load -> split -> STS -> MMA ->
// TMEM read -> store strictly one after the other inside a CTA and measured 64 us for 227 k rows with the tensor pipe 9 %
// busy: with loads, stores AND MMAs switched off the skeleton alone still took 25 us (profiles/README.md, r2 ablation) —
// the phases have to overlap. Here one persistent CTA per SM runs three roles on mbarrier rings:
//
//   loader warps    gather the tile's rows with 128-bit loads (the NEXT tile's loads are in flight while the current
//                   one is converted), split every fp32 value into three bf16 pieces (x = p1 + p2 + p3 carries 24 mantissa
//                   bits) and store the pieces as 128-byte-swizzled [rows x 64] bf16 tiles — the very layout TMA writes,
//                   so the same tile serves as a K-major operand (contraction over its 64 columns) and as an MN-major
//                   operand (contraction over its rows; tools/tc_probe2.cu variants 0 and 2 pin both descriptors);
//   MMA thread      one lane issues tcgen05.mma kind::f16 for the six piece products that matter
//                   (11, 12, 21, 22, 13, 31 — the dropped ones are below 2^-24), fp32 accumulation in tensor memory:
//                   row GEMM  D[128 x 64] (+)= A_piece[rows x h] . W_piece[n x h]^T            (both K-major)
//                   dW GEMM   D[ 64 x 64]  += G_piece[rows x h]^T . R_piece[rows x k]          (both MN-major, K = rows)
//                   and commits to the barriers that free the operand slot / publish the accumulator;
//   epilogue warps  read their 32 TMEM lanes (tcgen05.ld) and store the rows; per table they also drain the dW accumulator
//                   into the CTA's split-K partial (reduced in CTA order by fact_dw_reduce_kernel => reproducible).
//
// Two operand slots and two accumulator slots: the loader works on tile i+1 while the tensor core runs tile i and the
// epilogue drains tile i-1. Accuracy is fp32-level (1e-5 of tensor scale asserted against fp64 in tests/).
#include <stdlib.h>

#include "tgr_common.cuh"
#include "tgr_fact_params.cuh"
#include "tgr_rows.cuh"
#include "tgr_tc.cuh"

namespace tgr {
namespace ws {

constexpr int H = 64, H4 = 16;
constexpr int kStagePitch = H * 4 + 16;   // bytes per row of the output staging tile (272: conflict-free 128-bit accesses)

template <int MODE>
struct Cfg {
  static constexpr int RT = MODE ? 96 : 128;                 // rows per tile (the MMA is M = 128 either way)
  static constexpr int NLOAD = 8;                            // loader warps: two groups of four, alternating tiles
  static constexpr int NT = (NLOAD + 1 + 4) * 32;            // + MMA warp + 4 epilogue warps
  static constexpr int TILE = RT * 128;                      // bytes of one bf16 piece tile
  static constexpr int WTILE = H * 128;
  static constexpr int OPS = MODE ? 2 : 1;                   // operand matrices per slot (G and R | rows)
  static constexpr int WIN = MODE ? 960 : 2048;              // unique keys staged in shared memory at once
  static constexpr int STAGE = RT * kStagePitch;             // output staging tile (coalesced copy-out)
  static constexpr size_t SMEM = (size_t)2 * OPS * 3 * TILE + (size_t)2 * 3 * WTILE + STAGE + 1024;
  static constexpr int TMEM_COLS = MODE ? 256 : 128;
  static constexpr int TPW = WIN / RT;                       // tiles per key window
  static constexpr int LPT = RT * H4 / 128;                  // 128-bit loads per loader thread and operand (a group = 128 threads)
};

struct Item { int tile, seg_a, seg_b, t; };

__device__ __forceinline__ uint32_t sw128(int row, int chunk16) { return (uint32_t)(row * 128 + ((chunk16 ^ (row & 7)) << 4)); }

__device__ __forceinline__ void split3(const float4& v, uint2& p1, uint2& p2, uint2& p3) {
  p1 = pack_bf16x4(v);
  const float4 f1 = unpack_bf16x4(p1);
  const float4 r1 = make_float4(v.x - f1.x, v.y - f1.y, v.z - f1.z, v.w - f1.w);
  p2 = pack_bf16x4(r1);
  const float4 f2 = unpack_bf16x4(p2);
  p3 = pack_bf16x4(make_float4(r1.x - f2.x, r1.y - f2.y, r1.z - f2.z, r1.w - f2.w));
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(Cfg<MODE>::NT, 1) fact_rows_ws_kernel(const __grid_constant__ FactParams p,
                                                                       const uint32_t* __restrict__ uniq,
                                                                       const int32_t* __restrict__ n_unique_dev,
                                                                       float* __restrict__ PG, float* __restrict__ dw_part) {
  using C = Cfg<MODE>;
  constexpr int RT = C::RT, NLOAD = C::NLOAD, TILE = C::TILE, WTILE = C::WTILE, LPT = C::LPT;
  extern __shared__ uint8_t ws_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)ws_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* Abuf = base;                                          // [2 slots][3 pieces][TILE]   rows (MODE 0) / G (MODE 1)
  uint8_t* Rbuf = Abuf + 2 * 3 * TILE;                           // [2][3][TILE]                table rows (MODE 1)
  uint8_t* Wbuf = base + (size_t)2 * C::OPS * 3 * TILE;          // [2 weight slots][3][WTILE]
  uint8_t* Stage = Wbuf + (size_t)2 * 3 * WTILE;                 // [RT][kStagePitch] fp32 rows on their way out
  __shared__ __align__(8) uint64_t a_full[2], a_empty[2], t_full[2], t_empty[2], dw_full, dw_empty;
  __shared__ uint32_t s_tmem;
  __shared__ uint32_t s_key[C::WIN];
  __shared__ int32_t s_perm[C::WIN];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&a_full[b], 128);                  // one loader group fills a slot
      tc::mbar_init(&a_empty[b], 1);
      tc::mbar_init(&t_full[b], 1);
      tc::mbar_init(&t_empty[b], 4);
    }
    tc::mbar_init(&dw_full, 1);
    tc::mbar_init(&dw_empty, 4);
  }
  if (warp == 0) tc::tmem_alloc<C::TMEM_COLS>(&s_tmem);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem;
  const int U = *n_unique_dev;
  const int n_tiles = (U + RT - 1) / RT;
  const int tpc = (n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int tile_a = blockIdx.x * tpc, tile_b = min(n_tiles, tile_a + tpc);
  const bool is_loader = warp < NLOAD, is_mma = warp == NLOAD, is_epi = warp > NLOAD;

  uint32_t it = 0;          // items processed so far by this role (slot = it & 1, use = it >> 1)
  int prev_t = -1;          // table of the previous item (per role)
  int wsel = 1;             // weight slot of the current table (toggles at every table change; first table -> 0)
  uint32_t flushes = 0;     // dW drains so far (MMA / epilogue roles)

  for (int win_a = tile_a; win_a < tile_b; win_a += C::TPW) {
    const int win_b = min(tile_b, win_a + C::TPW);
    __syncthreads();                                     // every role is done reading the previous window's keys
    for (int i = tid; i < (win_b - win_a) * RT; i += C::NT) {
      const int u = win_a * RT + i;
      s_key[i] = u < U ? __ldg(uniq + u) : 0xFFFFFFFFu;
      if (p.fetched != nullptr) s_perm[i] = u >= U ? 0 : (p.fetched_perm ? __ldg(p.fetched_perm + u) : u);
    }
    __syncthreads();

    auto item_at = [&](int tile, int seg_a) {
      Item x;
      x.tile = tile; x.seg_a = seg_a; x.seg_b = 0; x.t = 0;
      if (tile >= win_b) return x;
      const uint32_t* k = s_key + (tile - win_a) * RT;
      const int nr = min(RT, U - tile * RT);
      x.t = find_table(p.key_base, p.n_tables, k[seg_a]);
      const uint32_t kend = p.key_base[x.t + 1];
      if (k[nr - 1] < kend) {
        x.seg_b = nr;
      } else {
        int lo = seg_a + 1, hi = nr - 1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (k[mid] < kend) lo = mid + 1; else hi = mid; }
        x.seg_b = lo;
      }
      return x;
    };
    auto next_of = [&](const Item& x) {
      const int nr = min(RT, U - x.tile * RT);
      return x.seg_b < nr ? item_at(x.tile, x.seg_b) : item_at(x.tile + 1, 0);
    };

    if (is_loader) {
      // =================================================== LOADER ===================================================
      // Two groups of four warps take alternate tiles (group g fills operand slot g). A group runs its tile start to end:
      // loads -> split -> shared-memory stores -> proxy fence -> arrive. The fence orders the generic-proxy stores before
      // the tensor core's reads and compiles to a CTA-wide memory barrier that also waits for the thread's outstanding
      // global loads — prefetching the next tile inside the same thread would therefore be serialised again (measured:
      // 87 us); the overlap comes from the OTHER group being in its load phase meanwhile.
      const int grp = warp >> 2, gw = warp & 3, gtid = tid & 127;
      for (Item x = item_at(win_a, 0); x.tile < win_b; x = next_of(x), ++it) {
        const bool new_table = x.t != prev_t;
        if (new_table) { wsel ^= 1; prev_t = x.t; }
        if ((int)(it & 1u) != grp) continue;
        const int b = grp;
        const int ns = x.seg_b - x.seg_a;
        const int row0 = x.tile * RT + x.seg_a;
        const uint32_t* k = s_key + (x.tile - win_a) * RT + x.seg_a;
        const int32_t* pm = s_perm + (x.tile - win_a) * RT + x.seg_a;
        const float* tab = p.w[x.t];
        const uint32_t kb = p.key_base[x.t];
        // lane -> (row 8 * rg + lane % 8, float4 column 4 * cq + lane / 8) of unit (rg, cq); LPT units per thread
        float4 va[LPT];
        float4 vr[MODE ? LPT : 1];
#pragma unroll
        for (int q = 0; q < LPT; ++q) {
          const int unit = gw * LPT + q;
          const int rg = unit >> 2, cq = unit & 3;
          const int r = rg * 8 + (lane & 7), c = cq * 4 + (lane >> 3);
          va[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (MODE) vr[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r < ns) {
            const float* rsrc;
            if (p.n_peers > 0) {
              const uint32_t key = k[r];
              rsrc = p.peer[key % (uint32_t)p.n_peers] + (size_t)(key / (uint32_t)p.n_peers) * H;
            } else if (p.fetched != nullptr) {
              rsrc = p.fetched + (size_t)pm[r] * H;
            } else {
              rsrc = tab + (size_t)(k[r] - kb) * H;
            }
            if (MODE) {
              vr[q] = __ldg(reinterpret_cast<const float4*>(rsrc) + c);
              va[q] = *(reinterpret_cast<const float4*>(PG + (size_t)(row0 + r) * H) + c);   // G row (written by this step)
            } else {
              va[q] = __ldg(reinterpret_cast<const float4*>(rsrc) + c);
            }
          }
        }
        if (it >= 2) tc::mbar_wait(&a_empty[b], ((it >> 1) - 1u) & 1u);   // the MMAs that read this slot have completed
        if (new_table) {
          // weight block of the new table -> three bf16 pieces in the other weight slot. MODE 0: B[n = h][k] = W[h][col + k];
          // MODE 1: B[n = k][h] (transposed: the row GEMM contracts over h)
          const float* W = p.dnn_w[p.side[x.t]];
          const int64_t ld = p.dnn_ld[p.side[x.t]];
          const int col = p.col[x.t];
          uint8_t* wb = Wbuf + (size_t)wsel * 3 * WTILE;
          constexpr int WPT = H * H4 / 128;
          float4 wv[WPT];
#pragma unroll
          for (int q = 0; q < WPT; ++q) {
            const int i = gtid + q * 128, h = i >> 4, c = i & 15;
            wv[q] = __ldg(reinterpret_cast<const float4*>(W + (size_t)h * ld + col) + c);
          }
#pragma unroll
          for (int q = 0; q < WPT; ++q) {
            const int i = gtid + q * 128, h = i >> 4, c = i & 15;
            uint2 p1, p2, p3;
            split3(wv[q], p1, p2, p3);
            if (MODE == 0) {
              const uint32_t off = sw128(h, c >> 1) + (c & 1) * 8;
              *reinterpret_cast<uint2*>(wb + off) = p1;
              *reinterpret_cast<uint2*>(wb + WTILE + off) = p2;
              *reinterpret_cast<uint2*>(wb + 2 * WTILE + off) = p3;
            } else {
              const uint16_t e1[4] = {(uint16_t)p1.x, (uint16_t)(p1.x >> 16), (uint16_t)p1.y, (uint16_t)(p1.y >> 16)};
              const uint16_t e2[4] = {(uint16_t)p2.x, (uint16_t)(p2.x >> 16), (uint16_t)p2.y, (uint16_t)(p2.y >> 16)};
              const uint16_t e3[4] = {(uint16_t)p3.x, (uint16_t)(p3.x >> 16), (uint16_t)p3.y, (uint16_t)(p3.y >> 16)};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t off = sw128(4 * c + j, h >> 3) + (h & 7) * 2;
                *reinterpret_cast<uint16_t*>(wb + off) = e1[j];
                *reinterpret_cast<uint16_t*>(wb + WTILE + off) = e2[j];
                *reinterpret_cast<uint16_t*>(wb + 2 * WTILE + off) = e3[j];
              }
            }
          }
        }
        uint8_t* ab = Abuf + (size_t)b * 3 * TILE;
        uint8_t* rb = Rbuf + (size_t)b * 3 * TILE;
#pragma unroll
        for (int q = 0; q < LPT; ++q) {
          const int unit = gw * LPT + q;
          const int rg = unit >> 2, cq = unit & 3;
          const int r = rg * 8 + (lane & 7), c = cq * 4 + (lane >> 3);
          const uint32_t off = sw128(r, c >> 1) + (c & 1) * 8;
          uint2 p1, p2, p3;
          split3(va[q], p1, p2, p3);
          *reinterpret_cast<uint2*>(ab + off) = p1;
          *reinterpret_cast<uint2*>(ab + TILE + off) = p2;
          *reinterpret_cast<uint2*>(ab + 2 * TILE + off) = p3;
          if (MODE) {
            split3(vr[q], p1, p2, p3);
            *reinterpret_cast<uint2*>(rb + off) = p1;
            *reinterpret_cast<uint2*>(rb + TILE + off) = p2;
            *reinterpret_cast<uint2*>(rb + 2 * TILE + off) = p3;
          } else if (p.save_rows != nullptr && r < ns) {
            st_stream(reinterpret_cast<float4*>(p.save_rows + (size_t)(row0 + r) * H) + c, va[q]);
          }
        }
        tc::fence_smem_to_async();
        mbar_arrive(&a_full[b]);
      }
    } else if (is_mma) {
      // ================================================= MMA ISSUER =================================================
      if (lane == 0) {
        const uint32_t idesc_k = tc::make_idesc(1u, 128, H);                          // bf16, A and B K-major
        const uint32_t idesc_mn = tc::make_idesc(1u, 128, H) | (1u << 15) | (1u << 16);   // A and B MN-major
        // piece products, smallest first: (a3 w1) (a1 w3) (a2 w2) (a2 w1) (a1 w2) (a1 w1)
        constexpr int PA[6] = {2, 0, 1, 1, 0, 0}, PB[6] = {0, 2, 1, 0, 1, 0};
        for (Item x = item_at(win_a, 0); x.tile < win_b; x = next_of(x)) {
          const int b = it & 1;
          const bool new_table = x.t != prev_t;
          if (new_table) {
            wsel ^= 1;
            if (MODE && prev_t >= 0) {                      // the previous table's dW block is complete: publish it,
              tc::commit(&dw_full);                         // and wait until the epilogue has drained it
              tc::mbar_wait(&dw_empty, flushes & 1u);
              ++flushes;
              tc::fence_after_sync();
            }
            prev_t = x.t;
          }
          tc::mbar_wait(&a_full[b], (it >> 1) & 1u);
          if (it >= 2) tc::mbar_wait(&t_empty[b], ((it >> 1) - 1u) & 1u);
          tc::fence_after_sync();
          const uint32_t a0 = tc::smem_u32(Abuf + (size_t)b * 3 * TILE), r0 = tc::smem_u32(Rbuf + (size_t)b * 3 * TILE);
          const uint32_t w0 = tc::smem_u32(Wbuf + (size_t)wsel * 3 * WTILE);
          const uint32_t tacc = tmem + b * H;
          bool first = true;
#pragma unroll
          for (int pr = 0; pr < 6; ++pr) {
#pragma unroll
            for (int ks = 0; ks < H / 16; ++ks) {
              const uint64_t da = tc::make_desc(a0 + PA[pr] * TILE + ks * 32, 16, 1024, 2);
              const uint64_t db = tc::make_desc(w0 + PB[pr] * WTILE + ks * 32, 16, 1024, 2);
              tc::mma_f16(tacc, da, db, idesc_k, !first);
              first = false;
            }
          }
          tc::commit(&t_full[b]);
          if (MODE) {
            const uint32_t tdw = tmem + 2 * H;
            bool facc = !new_table;                          // a new table starts a fresh dW accumulator
#pragma unroll
            for (int pr = 0; pr < 6; ++pr) {
#pragma unroll
              for (int ks = 0; ks < RT / 16; ++ks) {
                // MN-major: 16 contraction rows = two 8-row groups of 1024 bytes; M = 128 reads a second 64-wide MN group
                // at +LBO: LBO = 0 repeats the first (accumulator rows 64..127 are never read back)
                const uint64_t da = tc::make_desc(a0 + PA[pr] * TILE + ks * 2048, 0, 1024, 2);
                const uint64_t db = tc::make_desc(r0 + PB[pr] * TILE + ks * 2048, 0, 1024, 2);
                tc::mma_f16(tdw, da, db, idesc_mn, facc);
                facc = true;
              }
            }
          }
          tc::commit(&a_empty[b]);
          ++it;
        }
      }
    } else if (is_epi) {
      // ================================================== EPILOGUE ==================================================
      const int quarter = warp & 3;                          // TMEM lanes [32 quarter, 32 quarter + 32)
      const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
      auto drain_dw = [&](int t) {
        tc::mbar_wait(&dw_full, flushes & 1u);
        ++flushes;
        tc::fence_after_sync();
        const int h = quarter * 32 + lane;
        float* dst = dw_part + (size_t)(blockIdx.x + t) * H * H + (size_t)h * H;   // (cta, table) pairs are monotone => unique slots
#pragma unroll
        for (int c0 = 0; c0 < H; c0 += 16) {
          uint32_t rr[16];
          tc::ld16(tmem + lane_base + 2 * H + c0, rr);
          tc::ld_wait();
          if (quarter < 2) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(__uint_as_float(rr[j]), __uint_as_float(rr[j + 1]),
                                                                     __uint_as_float(rr[j + 2]), __uint_as_float(rr[j + 3]));
          }
        }
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&dw_empty);
      };
      for (Item x = item_at(win_a, 0); x.tile < win_b; x = next_of(x)) {
        const int b = it & 1;
        if (x.t != prev_t) {
          if (MODE && prev_t >= 0) drain_dw(prev_t);
          prev_t = x.t;
        }
        tc::mbar_wait(&t_full[b], (it >> 1) & 1u);
        tc::fence_after_sync();
        const int ns = x.seg_b - x.seg_a;
        const int r = quarter * 32 + lane;
        // TMEM -> registers -> padded staging tile (every lane owns one row: written straight to global memory that is one
        // 16-byte piece of 32 different lines per instruction, which kept L1TEX 66 % busy and the loaders' gathers queued
        // behind it — profiles/README.md r2 ws capture); the accumulator slot is released as soon as it is in registers
#pragma unroll
        for (int c0 = 0; c0 < H; c0 += 16) {
          uint32_t rr[16];
          tc::ld16(tmem + lane_base + b * H + c0, rr);
          tc::ld_wait();
          if (r < RT) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<uint4*>(Stage + r * kStagePitch + (c0 + j) * 4) = make_uint4(rr[j], rr[j + 1], rr[j + 2], rr[j + 3]);
          }
        }
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[b]);
        asm volatile("bar.sync 1, 128;" ::: "memory");          // the four epilogue warps: staging tile complete
        // the item's rows are one contiguous block of PG: linear, fully coalesced copy-out (16 lanes per 256-byte row)
        {
          float4* dst = reinterpret_cast<float4*>(PG + (size_t)(x.tile * RT + x.seg_a) * H);
          const int et = (warp - NLOAD - 1) * 32 + lane;       // 0 .. 127
          for (int i = et; i < ns * H4; i += 128) {
            const int rr_ = i >> 4, cc = i & 15;
            dst[i] = *reinterpret_cast<const float4*>(Stage + rr_ * kStagePitch + cc * 16);
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");          // staging tile free for the next item
        ++it;
      }
    }
  }
  // ---- the last table's dW block ----
  if (MODE) {
    if (is_mma && lane == 0 && prev_t >= 0) tc::commit(&dw_full);
    if (is_epi && prev_t >= 0) {
      const int quarter = warp & 3;
      tc::mbar_wait(&dw_full, flushes & 1u);
      tc::fence_after_sync();
      const int h = quarter * 32 + lane;
      float* dst = dw_part + (size_t)(blockIdx.x + prev_t) * H * H + (size_t)h * H;
#pragma unroll
      for (int c0 = 0; c0 < H; c0 += 16) {
        uint32_t rr[16];
        tc::ld16(tmem + ((uint32_t)(quarter * 32) << 16) + 2 * H + c0, rr);
        tc::ld_wait();
        if (quarter < 2) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(__uint_as_float(rr[j]), __uint_as_float(rr[j + 1]),
                                                                   __uint_as_float(rr[j + 2]), __uint_as_float(rr[j + 3]));
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<C::TMEM_COLS>(tmem);
}

template <int MODE>
static int launch(const FactParams& p, const uint32_t* uniq, const int32_t* n_unique_dev, float* PG, float* dw_part, cudaStream_t st) {
  using C = Cfg<MODE>;
  { static bool once = false; if (!once) { cudaFuncSetAttribute(fact_rows_ws_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM); once = true; } }
  TGR_K(fact_rows_ws_kernel<MODE>)<<<kNumSMs, C::NT, C::SMEM, st>>>(p, uniq, n_unique_dev, PG, dw_part);
  return check_launch(MODE ? "fact_unique_backward" : "fact_project_rows");
}

}  // namespace ws

// TGR_ROWS_WS=0 falls back to the earlier kernels (A/B timing); H = 64 only
bool rows_ws_supported(int H) {
  static const bool off = [] { const char* e = getenv("TGR_ROWS_WS"); return e && e[0] == '0'; }();
  return H == 64 && !off;
}
int launch_rows_ws_fwd(const FactParams& p, const uint32_t* uniq, const int32_t* n_unique_dev, float* P, cudaStream_t st) {
  return ws::launch<0>(p, uniq, n_unique_dev, P, nullptr, st);
}
int launch_rows_ws_bwd(const FactParams& p, const uint32_t* uniq, const int32_t* n_unique_dev, float* G, float* dw_part,
                       cudaStream_t st) {
  return ws::launch<1>(p, uniq, n_unique_dev, G, dw_part, st);
}
int rows_ws_bwd_grid() { return kNumSMs; }
int rows_ws_bwd_rt() { return ws::Cfg<1>::RT; }

}  // namespace tgr
