// Forward of the sparse-feature embedding path: fused multi-table gather + array sum-pool + concat write.
//
// Replaces, for one feat2emb call (model/BaseLine/model.py:240-247,267-279,302,305):
//   aten::embedding x (15 | 24)  +  sum(dim=2) x 4  +  cat(dim=2) x (1 | 2)
// with ONE launch that writes every SINGLE/ARRAY slot straight into the item/user concat buffers.
//
// Mapping (HBM-bound byte mover, no tensor cores):
//   * one CTA = TT consecutive tokens; its [TT, n_single] block of the token-major id matrix is one
//     contiguous span, staged into shared memory with 128-bit loads;
//   * one "group" of LANES = H/4 threads moves one H-float row as one 128-bit access per thread
//     (H=64: 16 lanes x 16 B = two full 128 B lines per row, read and written fully coalesced);
//   * each group keeps UNROLL independent row loads in flight before the first store;
//   * table rows use the default cache policy (Zipf-hot rows / small tables stay in L2), the concat
//     output is written with st.global.cs (streamed, evict-first);
//   * ARRAY slots: CSR list summed left to right in fp32 — the order of torch's sum(2) on the
//     reference's zero-padded [B,L,A,H] gather, padding ids contribute row 0 == 0.
#include "tgr_common.cuh"

namespace tgr {

struct FwdSlot {
  const float* w;        // table base
  char* out;             // concat base of this slot's side, already offset to the slot's first column
  const int32_t* arr_off;  // ARRAY: CSR offsets [T+1]
  int64_t ld_bytes;      // row pitch of the concat buffer in bytes
  int32_t rows;          // table rows (range check)
  int32_t src;           // SINGLE: ids column
};

struct FwdParams {
  FwdSlot slot[TGR_MAX_SLOTS];  // singles first [0, n_s), arrays after [n_s, n_s + n_a)
  const int32_t* ids;
  const int32_t* arr_val;
  int32_t* err;  // optional: receives 1 + offending slot on an out-of-range id
  int32_t T, n_single, n_s, n_a, H4;
};

constexpr int kThreads = 256;
constexpr int kTT = 16;      // tokens per CTA
constexpr int kUnroll = 4;   // independent row loads in flight per thread

template <bool BF16>
__device__ __forceinline__ void store_chunk(char* out_row, int c, const float4& v) {
  if constexpr (BF16) {
    st_stream_u2(reinterpret_cast<uint2*>(out_row) + c, pack_bf16x4(v));
  } else {
    st_stream(reinterpret_cast<float4*>(out_row) + c, v);
  }
}

template <int LANES, bool BF16>
__global__ void __launch_bounds__(kThreads) fwd_gather_pool_concat_kernel(const __grid_constant__ FwdParams p) {
  extern __shared__ int32_t sm_ids[];  // [kTT * n_single]
  constexpr int G = kThreads / LANES;  // groups per CTA
  const int tid = threadIdx.x;
  const int lane = tid % LANES;
  const int grp = tid / LANES;
  const int t0 = blockIdx.x * kTT;
  const int nt = min(kTT, p.T - t0);
  const int H4 = p.H4;

  // ---- stage the id block (contiguous in global memory) ----
  {
    const int n = nt * p.n_single;
    const int32_t* src = p.ids + (size_t)t0 * p.n_single;
    // the block start is 16 B aligned when kTT * n_single * 4 is (kTT = 16 => always)
    const int n4 = n >> 2;
    const int4* src4 = reinterpret_cast<const int4*>(src);
    int4* dst4 = reinterpret_cast<int4*>(sm_ids);
    for (int i = tid; i < n4; i += kThreads) dst4[i] = __ldg(src4 + i);
    for (int i = (n4 << 2) + tid; i < n; i += kThreads) sm_ids[i] = __ldg(src + i);
  }
  __syncthreads();

  // ---- pass 1: SINGLE slots, kUnroll rows in flight per group ----
  const int n_s = p.n_s;
  const int items = nt * n_s;
  for (int i0 = grp; i0 < items; i0 += G * kUnroll) {
    float4 v[kUnroll];
    char* dst[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int i = i0 + u * G;
      dst[u] = nullptr;
      if (i < items) {
        const int tl = i / n_s;
        const int s = i - tl * n_s;
        const FwdSlot& sl = p.slot[s];
        int id = sm_ids[tl * p.n_single + sl.src];
        if ((unsigned)id >= (unsigned)sl.rows) {
          if (p.err && lane == 0) atomicMax(p.err, s + 1);
          id = 0;
        }
        dst[u] = sl.out + (size_t)(t0 + tl) * sl.ld_bytes;
        if (lane < H4) v[u] = ld_row(reinterpret_cast<const float4*>(sl.w + (size_t)id * (H4 * 4)) + lane);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (dst[u] != nullptr && lane < H4) store_chunk<BF16>(dst[u], lane, v[u]);
    }
    if (H4 > LANES) {  // H > 128: remaining 128-bit columns, same rows
      for (int c = lane + LANES; c < H4; c += LANES) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int i = i0 + u * G;
          if (i < items) {
            const int tl = i / n_s;
            const int s = i - tl * n_s;
            const FwdSlot& sl = p.slot[s];
            int id = sm_ids[tl * p.n_single + sl.src];
            if ((unsigned)id >= (unsigned)sl.rows) id = 0;
            float4 x = ld_row(reinterpret_cast<const float4*>(sl.w + (size_t)id * (H4 * 4)) + c);
            store_chunk<BF16>(dst[u], c, x);
          }
        }
      }
    }
  }

  // ---- pass 2: ARRAY slots (ragged, mostly empty: one user token per sequence) ----
  const int n_a = p.n_a;
  const int aitems = nt * n_a;
  for (int i = grp; i < aitems; i += G) {
    const int tl = i / n_a;
    const int a = i - tl * n_a;
    const FwdSlot& sl = p.slot[n_s + a];
    const int t = t0 + tl;
    const int lo = __ldg(sl.arr_off + t), hi = __ldg(sl.arr_off + t + 1);
    char* dst = sl.out + (size_t)t * sl.ld_bytes;
    for (int c = lane; c < H4; c += LANES) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int e0 = lo; e0 < hi; e0 += kUnroll) {
        float4 x[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (e0 + u < hi) {
            int id = __ldg(p.arr_val + e0 + u);
            if ((unsigned)id >= (unsigned)sl.rows) {
              if (p.err && lane == 0) atomicMax(p.err, n_s + a + 1);
              id = 0;
            }
            x[u] = ld_row(reinterpret_cast<const float4*>(sl.w + (size_t)id * (H4 * 4)) + c);
          }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
          if (e0 + u < hi) acc = f4_add(acc, x[u]);  // left to right
      }
      store_chunk<BF16>(dst, c, acc);
    }
  }
}

template <int LANES>
static int launch_fwd(const FwdParams& p, bool bf16, cudaStream_t st) {
  const int grid = (p.T + kTT - 1) / kTT;
  const size_t smem = (size_t)kTT * p.n_single * sizeof(int32_t);
  if (bf16)
    TGR_K(fwd_gather_pool_concat_kernel<LANES, true>)<<<grid, kThreads, smem, st>>>(p);
  else
    TGR_K(fwd_gather_pool_concat_kernel<LANES, false>)<<<grid, kThreads, smem, st>>>(p);
  return check_launch("fwd_gather_pool_concat");
}

}  // namespace tgr

extern "C" int tgr_fwd_gather_pool_concat(const tgr_table_t* tables, int n_tables, int H, const tgr_call_t* call,
                                          void* stream) {
  tgr::TimedScope tgr_timed_("fwd_gather_pool_concat", stream);
  using namespace tgr;
  TGR_REQUIRE(tables && call, "null argument");
  TGR_REQUIRE(H > 0 && H % 4 == 0, "H=%d must be a positive multiple of 4", H);
  TGR_REQUIRE(n_tables > 0 && n_tables <= TGR_MAX_TABLES, "n_tables=%d out of range", n_tables);
  TGR_REQUIRE(call->n_slots > 0 && call->n_slots <= TGR_MAX_SLOTS, "n_slots=%d out of range", call->n_slots);
  TGR_REQUIRE(call->T >= 0, "T=%d negative", call->T);
  TGR_REQUIRE(call->cat_dtype == TGR_DTYPE_F32 || call->cat_dtype == TGR_DTYPE_BF16, "bad cat_dtype");
  if (call->T == 0) return 0;
  const size_t esz = call->cat_dtype == TGR_DTYPE_BF16 ? 2 : 4;
  FwdParams p{};
  p.ids = call->ids;
  p.arr_val = call->arr_val;
  p.err = call->err_flag;
  p.T = call->T;
  p.n_single = call->n_single;
  p.H4 = H / 4;
  int n_s = 0, n_a = 0;
  for (int pass = 0; pass < 2; ++pass) {
    for (int i = 0; i < call->n_slots; ++i) {
      const tgr_slot_t& s = call->slots[i];
      if (s.kind == TGR_KIND_MM) continue;
      TGR_REQUIRE(s.kind == TGR_KIND_SINGLE || s.kind == TGR_KIND_ARRAY, "slot %d: bad kind %d", i, s.kind);
      if ((s.kind == TGR_KIND_SINGLE) != (pass == 0)) continue;
      TGR_REQUIRE(s.table >= 0 && s.table < n_tables, "slot %d: table %d out of range", i, s.table);
      char* base = (char*)(s.side == TGR_SIDE_ITEM ? call->item_cat : call->user_cat);
      const int64_t ld = s.side == TGR_SIDE_ITEM ? call->item_ld : call->user_ld;
      TGR_REQUIRE(base != nullptr, "slot %d: concat buffer of side %d is NULL", i, s.side);
      TGR_REQUIRE(s.col % 4 == 0 && ld % 4 == 0 && s.col + H <= ld, "slot %d: col/ld not 128-bit tileable", i);
      FwdSlot& d = p.slot[n_s + n_a];
      d.w = tables[s.table].weight;
      d.rows = (int32_t)tables[s.table].rows;
      d.out = base + (size_t)s.col * esz;
      d.ld_bytes = ld * (int64_t)esz;
      d.src = s.src;
      d.arr_off = nullptr;
      TGR_REQUIRE(d.w != nullptr, "slot %d: table weight is NULL", i);
      if (s.kind == TGR_KIND_SINGLE) {
        TGR_REQUIRE(s.src >= 0 && s.src < call->n_single, "slot %d: ids column %d out of range", i, s.src);
        ++n_s;
      } else {
        TGR_REQUIRE(s.src >= 0 && s.src < call->n_arrays && s.src < TGR_MAX_ARRAYS, "slot %d: array %d out of range", i, s.src);
        d.arr_off = call->arr_off[s.src];
        TGR_REQUIRE(d.arr_off != nullptr && (call->arr_val != nullptr || call->arr_nnz[s.src] == 0), "slot %d: array pointers NULL", i);
        ++n_a;
      }
    }
  }
  p.n_s = n_s;
  p.n_a = n_a;
  TGR_REQUIRE(n_s == 0 || p.ids != nullptr, "ids is NULL");
  TGR_REQUIRE(((uintptr_t)p.ids & 15) == 0, "ids must be 16-byte aligned");
  const bool bf16 = call->cat_dtype == TGR_DTYPE_BF16;
  cudaStream_t st = (cudaStream_t)stream;
  const int H4 = p.H4;
  if (H4 <= 8) return launch_fwd<8>(p, bf16, st);
  if (H4 <= 16) return launch_fwd<16>(p, bf16, st);
  return launch_fwd<32>(p, bf16, st);
}
