// bf16 tensor-core multimodal projection for the wide mm features ('82' 1024-d ... '84' 4096-d; model/BaseLine/model.py:183,
// 281-299; BASELINE.json config 3): out[t, 0:H] = x[t, :] . Wb^T + bias with x bf16 [T, mm_dim] (frozen features kept in
// bf16), Wb a bf16 copy of the fp32 weight — emb_transform[k].weight, or on the factored path the folded
// W_slot . W_mm (tgr_fact_mm_fold) — and fp32 accumulation in tensor memory.
//
// Blackwell-native pipeline (tcgen05 + TMEM + TMA; sm_100a only):
//   * TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) streams [128 tokens x 64 k] tiles of x and [H x 64 k] tiles of Wb into
//     a 4-slot shared-memory ring; one thread arms the slot's mbarrier with the byte count and issues both copies;
//   * the same thread issues 4 tcgen05.mma kind::f16 (M = 128, N = H, K = 16) per slot straight from the swizzled tiles
//     (K-major SWIZZLE_128B descriptors, +32 bytes per K step inside the swizzle atom), accumulating into one TMEM tile
//     [128 lanes x H fp32 columns]; tcgen05.commit releases the slot / signals the accumulator;
//   * all four warps read their 32 TMEM lanes back (tcgen05.ld), add the bias and write the output rows;
//   * persistent CTAs, 2 per SM: the next tile's first four k-blocks are already in flight during the epilogue.
// The kernel is HBM-bound on x (T * mm_dim * 2 bytes; intensity 2H / 2 = 64 flop/B at H = 64): the tensor pipe is busy a
// few percent of the time by construction — see profiles/ for the sm__pipe_tensor counters.
#include <cuda.h>
#include <stdlib.h>

#include "tgr_common.cuh"
#include "tgr_tc.cuh"

namespace tgr {

constexpr int kTcBM = 128, kTcBK = 64, kTcStages = 4;
constexpr int kTcFwdStages = 3;   // forward ring: 3 x (16 KB x + 2 planes x H x 128 B) = 96 KB at H = 64 -> 2 CTAs / SM

typedef CUresult (*TmEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static TmEncodeFn tm_encode() {
  static TmEncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
    return (TmEncodeFn)p;
  }();
  return fn;
}

// row-major bf16 [rows, cols] -> 2-D tensor map with a [box_rows x 64] box, 128-byte swizzle, OOB rows zero-filled
static int make_map(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int box_rows) {
  TmEncodeFn enc = tm_encode();
  TGR_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)kTcBK, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TGR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c_inner, int c_outer) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   tc::smem_u32(smem_dst)),
               "l"(tm), "r"(tc::smem_u32(bar)), "r"(c_inner), "r"(c_outer)
               : "memory");
}

template <int H, bool OBF16>
__global__ void __launch_bounds__(128) mm_proj_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_x,
                                                             const __grid_constant__ CUtensorMap tm_w,
                                                             const float* __restrict__ bias, char* __restrict__ out,
                                                             int64_t out_ld_bytes, int64_t T, int K, int n_tiles, int planes) {
  // W arrives as `planes` bf16 planes stacked along the rows ([planes * H, K]: hi, then the bf16 of the remainder): with
  // two planes the weights carry 16 mantissa bits and the product with the bf16-STORED features matches fp32 math.
  constexpr int A_BYTES = kTcBM * kTcBK * 2, P_BYTES = H * kTcBK * 2, B_BYTES = 2 * P_BYTES, S = kTcFwdStages;
  constexpr int TMEM_COLS = H < 32 ? 32 : H;
  extern __shared__ uint8_t mmtc_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)mmtc_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* As = base;                 // [S][128 x 64 bf16], rows 128 bytes, 128-byte swizzle (as TMA wrote it)
  uint8_t* Bs = base + S * A_BYTES;   // [S][H x 64 bf16]
  __shared__ __align__(8) uint64_t full[S], empty[S], accbar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    tc::mbar_init(&accbar, 1);
  }
  if (warp == 0) tc::tmem_alloc<TMEM_COLS>(&s_tmem);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tacc = s_tmem;
  const uint32_t idesc = tc::make_idesc(1u, 128, H);   // bf16 x bf16 -> f32
  const int nkb = K / kTcBK;
  const int my_tiles = (int)blockIdx.x < n_tiles ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const uint32_t total = (uint32_t)my_tiles * (uint32_t)nkb;
  uint32_t consumed = 0;   // k-blocks consumed so far (thread 0)

  auto issue_load = [&](uint32_t item) {   // thread 0 only
    const uint32_t s = item % S;
    if (item >= (uint32_t)S) tc::mbar_wait(&empty[s], ((item / S) - 1u) & 1u);   // the MMAs that read this slot's previous tile
    const int tile = (int)blockIdx.x + (int)(item / (uint32_t)nkb) * (int)gridDim.x;
    const int kb = (int)(item % (uint32_t)nkb);
    tc::mbar_expect_tx(&full[s], A_BYTES + planes * P_BYTES);
    tma_load_2d(As + s * A_BYTES, &tm_x, &full[s], kb * kTcBK, tile * kTcBM);
    for (int pl = 0; pl < planes; ++pl) tma_load_2d(Bs + s * B_BYTES + pl * P_BYTES, &tm_w, &full[s], kb * kTcBK, pl * H);
  };
  if (tid == 0)
    for (uint32_t it = 0; it < (uint32_t)S && it < total; ++it) issue_load(it);

  for (int ti = 0; ti < my_tiles; ++ti) {
    const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
    if (tid == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const uint32_t s = consumed % S;
        tc::mbar_wait(&full[s], (consumed / S) & 1u);
        tc::fence_after_sync();
        const uint32_t a0 = tc::smem_u32(As + s * A_BYTES), b0 = tc::smem_u32(Bs + s * B_BYTES);
#pragma unroll
        for (int k = 0; k < kTcBK / 16; ++k) {
          const uint64_t da = tc::make_desc(a0 + k * 32, 16, 1024, 2);   // SWIZZLE_128B, 8-row groups 1024 B apart
          for (int pl = planes - 1; pl >= 0; --pl) {                      // small terms first
            const uint64_t db = tc::make_desc(b0 + pl * P_BYTES + k * 32, 16, 1024, 2);
            tc::mma_f16(tacc, da, db, idesc, kb > 0 || k > 0 || pl < planes - 1);
          }
        }
        tc::commit(&empty[s]);
        if (kb == nkb - 1) tc::commit(&accbar);
        if (consumed + S < total) issue_load(consumed + S);
        ++consumed;
      }
    }
    tc::mbar_wait(&accbar, (uint32_t)ti & 1u);
    tc::fence_after_sync();
    {
      const int64_t t = (int64_t)tile * kTcBM + warp * 32 + lane;
      char* row = out + (size_t)t * out_ld_bytes;
      const uint32_t taddr = tacc + ((uint32_t)(warp * 32) << 16);
#pragma unroll
      for (int c0 = 0; c0 < H; c0 += 16) {
        uint32_t rr[16];
        tc::ld16(taddr + c0, rr);
        tc::ld_wait();
        if (t < T) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 v = make_float4(__uint_as_float(rr[j]), __uint_as_float(rr[j + 1]), __uint_as_float(rr[j + 2]),
                                   __uint_as_float(rr[j + 3]));
            if (bias != nullptr) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c0 + j));
              v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            }
            if constexpr (OBF16) *(reinterpret_cast<uint2*>(row) + ((c0 + j) >> 2)) = pack_bf16x4(v);
            else *(reinterpret_cast<float4*>(row) + ((c0 + j) >> 2)) = v;
          }
        }
      }
    }
    tc::fence_before_sync();
    __syncthreads();   // the accumulator is drained before the next tile's first MMA (accumulate = 0) overwrites it
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<TMEM_COLS>(tacc);
}

// ---- backward: A[h][j] (+)= sum_t dZ[t][h] x[t][j]  (dW of the projection: autograd of model.py:297 through the fold) --------
// Split-K (over tokens) tcgen05 GEMM with BOTH operands MN-major: dZ (bf16 copy, [T, H]) and x (bf16, [T, mm_dim]) are
// row-major with the contraction index t as the ROW, so a TMA box {64 columns, 64 tokens} lands in shared memory as eight
// 128-byte-swizzled atoms of 8 token rows x 64 MN elements — the canonical MN-major SWIZZLE_128B operand (verified on B200,
// tools/tc_probe2.cu variant 2: LBO = stride between 64-element MN groups, SBO = 1024 between 8-row K groups, +2048 bytes
// per K = 16 MMA). M = 128 accumulator rows of which H are real: for H = 64 the second MN group of A repeats the first
// (LBO = 0), rows 64..127 are never read back. CTA = (token chunk, BN-column tile): 4-slot TMA ring, one thread issues
// the copies and the MMAs, D [128 x BN] fp32 in tensor memory, the H real lanes write the chunk's partial; the existing
// fixed-order reduction (mm_proj_bwd_reduce_kernel) sums the chunks => bitwise reproducible.
template <int H, int BN>
__global__ void __launch_bounds__(128) mm_proj_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_z,
                                                             const __grid_constant__ CUtensorMap tm_x,
                                                             float* __restrict__ ws_dw, int64_t T, int K, int n_chunks,
                                                             int planes, int64_t plane_rows) {
  // dz arrives as `planes` bf16 planes stacked along the rows ([planes * plane_rows, H]: hi, then bf16(dz - hi))
  constexpr int S = kTcStages, BT = 64;
  constexpr int PA_BYTES = (H / 64) * BT * 128, A_BYTES = 2 * PA_BYTES, B_BYTES = (BN / 64) * BT * 128;   // boxes of 64 tokens x 128 bytes
  constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t mmtc_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)mmtc_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* As = base;
  uint8_t* Bs = base + S * A_BYTES;
  __shared__ __align__(8) uint64_t full[S], empty[S], accbar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int chunk = blockIdx.x, j0 = blockIdx.y * BN;
  const int64_t per = ((T + n_chunks - 1) / n_chunks + BT - 1) / BT * BT;   // tokens per chunk, multiple of 64
  const int64_t tb = (int64_t)chunk * per, te = min(T, tb + per);
  const int nkb = te > tb ? (int)((te - tb + BT - 1) / BT) : 0;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    tc::mbar_init(&accbar, 1);
  }
  if (warp == 0) tc::tmem_alloc<TMEM_COLS>(&s_tmem);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tacc = s_tmem;
  if (tid == 0 && nkb > 0) {
    // bf16 x bf16 -> f32, A and B MN-major (bits 15 / 16)
    const uint32_t idesc = tc::make_idesc(1u, 128, BN) | (1u << 15) | (1u << 16);
    auto issue_load = [&](int kb) {
      const int s = kb % S;
      if (kb >= S) tc::mbar_wait(&empty[s], (uint32_t)((kb / S) - 1) & 1u);
      tc::mbar_expect_tx(&full[s], planes * PA_BYTES + B_BYTES);
      const int t0 = (int)(tb + (int64_t)kb * BT);
      for (int pl = 0; pl < planes; ++pl)
#pragma unroll
        for (int g = 0; g < H / 64; ++g)
          tma_load_2d(As + s * A_BYTES + pl * PA_BYTES + g * (BT * 128), &tm_z, &full[s], g * 64, (int)(pl * plane_rows) + t0);
#pragma unroll
      for (int g = 0; g < BN / 64; ++g) tma_load_2d(Bs + s * B_BYTES + g * (BT * 128), &tm_x, &full[s], j0 + g * 64, t0);
    };
    for (int kb = 0; kb < S && kb < nkb; ++kb) issue_load(kb);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % S;
      tc::mbar_wait(&full[s], (uint32_t)(kb / S) & 1u);
      tc::fence_after_sync();
      const uint32_t a0 = tc::smem_u32(As + s * A_BYTES), b0 = tc::smem_u32(Bs + s * B_BYTES);
#pragma unroll
      for (int k = 0; k < BT / 16; ++k) {
        const uint64_t db = tc::make_desc(b0 + k * 2048, BT * 128, 1024, 2);
        for (int pl = planes - 1; pl >= 0; --pl) {   // small terms first
          const uint64_t da = tc::make_desc(a0 + pl * PA_BYTES + k * 2048, H > 64 ? BT * 128 : 0, 1024, 2);
          tc::mma_f16(tacc, da, db, idesc, kb > 0 || k > 0 || pl < planes - 1);
        }
      }
      tc::commit(&empty[s]);
      if (kb + S < nkb) issue_load(kb + S);
    }
    tc::commit(&accbar);
  }
  const int h = warp * 32 + lane;          // accumulator lane = output row h
  float* dst = ws_dw + ((size_t)chunk * H + h) * (size_t)K + j0;
  if (nkb > 0) {
    tc::mbar_wait(&accbar, 0);
    tc::fence_after_sync();
    const uint32_t taddr = tacc + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t rr[16];
      tc::ld16(taddr + c0, rr);
      tc::ld_wait();
      if (h < H) {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(__uint_as_float(rr[j]), __uint_as_float(rr[j + 1]),
                                                                 __uint_as_float(rr[j + 2]), __uint_as_float(rr[j + 3]));
      }
    }
  } else if (h < H) {
    for (int c0 = 0; c0 < BN; c0 += 4) *reinterpret_cast<float4*>(dst + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<TMEM_COLS>(tacc);
}

// out[o] (+)= sum over chunks of ws[c][o], chunks ascending per lane, fixed shuffle tree (deterministic)
__global__ void __launch_bounds__(256) chunk_reduce_kernel(const float* __restrict__ ws, int n_chunks, int64_t n_out,
                                                           float* __restrict__ out, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n_out) return;
  float s = 0.f;
  for (int c = lane; c < n_chunks; c += 32) s += ws[(size_t)c * n_out + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[i] = accumulate ? out[i] + s : s;
}

// column tile: 256 where the ring fits (H = 64: 4 x 48 KB), 128 for H = 128 (4 x 48 KB)
static int bwd_tc_bn(int mm_dim, int H) { return mm_dim % 256 == 0 && H <= 64 ? 256 : (mm_dim % 128 == 0 ? 128 : 64); }
static int bwd_tc_chunks(int64_t T, int mm_dim, int H) {
  const int nt = mm_dim / bwd_tc_bn(mm_dim, H);
  int n = (kNumSMs + nt - 1) / nt;
  const int64_t by_T = (T + 63) / 64;
  if (n > by_T) n = (int)by_T;
  return n < 1 ? 1 : n;
}

template <int H, int BN>
static int launch_bwd_tc(const CUtensorMap& tz, const CUtensorMap& tx, float* ws, int64_t T, int K, int n_chunks, int planes,
                         int64_t plane_rows, cudaStream_t st) {
  const size_t smem = (size_t)kTcStages * (2 * (H / 64) * 64 * 128 + (BN / 64) * 64 * 128) + 1024;
  { static bool tgr_attr_once_ = false; if (!tgr_attr_once_) { cudaFuncSetAttribute(mm_proj_bwd_tc_kernel<H, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); tgr_attr_once_ = true; } }
  TGR_K(mm_proj_bwd_tc_kernel<H, BN>)<<<dim3(n_chunks, K / BN), 128, smem, st>>>(tz, tx, ws, T, K, n_chunks, planes, plane_rows);
  return check_launch("mm_proj_bwd_tc");
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ src, int64_t n, __nv_bfloat16* __restrict__ dst) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + i));
    *reinterpret_cast<uint2*>(dst + i) = pack_bf16x4(v);
  } else {
    for (int64_t j = i; j < n; ++j) dst[j] = __float2bfloat16(src[j]);
  }
}

// hi[i] = bf16(src[i]); lo[i] = bf16(src[i] - hi[i])  (two bf16 planes = 16 mantissa bits)
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ src, int64_t n, __nv_bfloat16* __restrict__ hi,
                                                         __nv_bfloat16* __restrict__ lo) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + i));
    const uint2 h = pack_bf16x4(v);
    const float4 hf = unpack_bf16x4(h);
    *reinterpret_cast<uint2*>(hi + i) = h;
    *reinterpret_cast<uint2*>(lo + i) = pack_bf16x4(make_float4(v.x - hf.x, v.y - hf.y, v.z - hf.z, v.w - hf.w));
  } else {
    for (int64_t j = i; j < n; ++j) {
      const __nv_bfloat16 h = __float2bfloat16(src[j]);
      hi[j] = h;
      lo[j] = __float2bfloat16(src[j] - __bfloat162float(h));
    }
  }
}

template <int H, bool OBF16>
static int launch_tc(const CUtensorMap& tx, const CUtensorMap& tw, const float* bias, char* out, int64_t ldb, int64_t T, int K,
                     int planes, cudaStream_t st) {
  const size_t smem = (size_t)kTcFwdStages * (kTcBM * kTcBK * 2 + 2 * H * kTcBK * 2) + 1024;
  { static bool tgr_attr_once_ = false; if (!tgr_attr_once_) { cudaFuncSetAttribute(mm_proj_fwd_tc_kernel<H, OBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); tgr_attr_once_ = true; } }
  const int n_tiles = (int)((T + kTcBM - 1) / kTcBM);
  const int grid = n_tiles < 2 * kNumSMs ? n_tiles : 2 * kNumSMs;
  TGR_K(mm_proj_fwd_tc_kernel<H, OBF16>)<<<grid, 128, smem, st>>>(tx, tw, bias, out, ldb, T, K, n_tiles, planes);
  return check_launch("mm_proj_fwd_tc");
}

}  // namespace tgr

using namespace tgr;

extern "C" int tgr_cast_bf16(const float* src, int64_t n, void* dst_bf16, void* stream) {
  tgr::TimedScope tgr_timed_("cast_bf16", stream);
  TGR_REQUIRE(n >= 0, "n out of range");
  if (n == 0) return 0;
  TGR_REQUIRE(src && dst_bf16, "null argument");
  TGR_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst_bf16 & 7) == 0, "cast_bf16: misaligned buffers");
  const int64_t groups = (n + 3) / 4;
  TGR_K(cast_bf16_kernel)<<<(unsigned)((groups + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, n, (__nv_bfloat16*)dst_bf16);
  return check_launch("cast_bf16");
}

extern "C" int tgr_split_bf16(const float* src, int64_t n, void* hi_bf16, void* lo_bf16, void* stream) {
  tgr::TimedScope tgr_timed_("cast_bf16", stream);
  TGR_REQUIRE(n >= 0, "n out of range");
  if (n == 0) return 0;
  TGR_REQUIRE(src && hi_bf16 && lo_bf16, "null argument");
  TGR_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)hi_bf16 & 7) == 0 && ((uintptr_t)lo_bf16 & 7) == 0, "split_bf16: misaligned buffers");
  const int64_t groups = (n + 3) / 4;
  TGR_K(split_bf16_kernel)<<<(unsigned)((groups + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, n, (__nv_bfloat16*)hi_bf16, (__nv_bfloat16*)lo_bf16);
  return check_launch("split_bf16");
}

extern "C" int tgr_mm_proj_fwd_tc_supported(int x_dtype, int mm_dim, int H) {
  return x_dtype == TGR_DTYPE_BF16 && mm_dim >= 128 && mm_dim % kTcBK == 0 && (H == 32 || H == 64 || H == 128) ? 1 : 0;
}

extern "C" int tgr_mm_proj_fwd_tc(const void* x_bf16, int64_t T, int mm_dim, const void* w_bf16, int w_planes, const float* bias,
                                  int H, void* out, int64_t out_ld, int out_dtype, void* stream) {
  tgr::TimedScope tgr_timed_("mm_proj_fwd_tc", stream);
  TGR_REQUIRE(x_bf16 && w_bf16 && out, "null argument");
  TGR_REQUIRE(tgr_mm_proj_fwd_tc_supported(TGR_DTYPE_BF16, mm_dim, H), "mm_proj_fwd_tc: mm_dim %% 64 == 0, >= 128 and H in {32, 64, 128} (mm_dim=%d, H=%d)", mm_dim, H);
  TGR_REQUIRE(out_ld % 4 == 0, "out_ld must be a multiple of 4 elements");
  TGR_REQUIRE(((uintptr_t)x_bf16 & 15) == 0 && ((uintptr_t)w_bf16 & 15) == 0 && ((uintptr_t)out & 15) == 0, "mm_proj_fwd_tc: misaligned buffers");
  TGR_REQUIRE(T >= 0 && T < (1ll << 31), "T out of range");
  TGR_REQUIRE(w_planes == 1 || w_planes == 2, "w_planes must be 1 or 2");
  if (T == 0) return 0;
  CUtensorMap tx, tw;
  if (int rc = make_map(&tx, x_bf16, T, mm_dim, kTcBM)) return rc;
  if (int rc = make_map(&tw, w_bf16, (int64_t)w_planes * H, mm_dim, H)) return rc;
  const bool ob = out_dtype == TGR_DTYPE_BF16;
  const int64_t ldb = out_ld * (ob ? 2 : 4);
  cudaStream_t st = (cudaStream_t)stream;
  char* o = (char*)out;
  if (H == 32) return ob ? launch_tc<32, true>(tx, tw, bias, o, ldb, T, mm_dim, w_planes, st) : launch_tc<32, false>(tx, tw, bias, o, ldb, T, mm_dim, w_planes, st);
  if (H == 64) return ob ? launch_tc<64, true>(tx, tw, bias, o, ldb, T, mm_dim, w_planes, st) : launch_tc<64, false>(tx, tw, bias, o, ldb, T, mm_dim, w_planes, st);
  return ob ? launch_tc<128, true>(tx, tw, bias, o, ldb, T, mm_dim, w_planes, st) : launch_tc<128, false>(tx, tw, bias, o, ldb, T, mm_dim, w_planes, st);
}

/* ---- backward on the tensor cores (see mm_proj_bwd_tc_kernel) ---- */
extern "C" int tgr_mm_proj_bwd_tc_supported(int x_dtype, int mm_dim, int H) {
  return x_dtype == TGR_DTYPE_BF16 && mm_dim >= 128 && mm_dim % kTcBK == 0 && (H == 64 || H == 128) ? 1 : 0;
}

extern "C" size_t tgr_mm_proj_bwd_tc_workspace_bytes(int64_t T, int mm_dim, int H) {
  return (size_t)bwd_tc_chunks(T, mm_dim, H) * H * mm_dim * sizeof(float) + 256;
}

extern "C" int tgr_mm_proj_bwd_tc(const void* x_bf16, int64_t T, int mm_dim, const void* dz_bf16, int dz_planes, int64_t plane_rows,
                                  int H, float* dW, int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  tgr::TimedScope tgr_timed_("mm_proj_bwd_tc", stream);
  TGR_REQUIRE(x_bf16 && dz_bf16 && dW && workspace, "null argument");
  TGR_REQUIRE(tgr_mm_proj_bwd_tc_supported(TGR_DTYPE_BF16, mm_dim, H), "mm_proj_bwd_tc: mm_dim %% 64 == 0, >= 128 and H in {64, 128} (mm_dim=%d, H=%d)", mm_dim, H);
  TGR_REQUIRE(T >= 0 && T < (1ll << 31), "T out of range");
  TGR_REQUIRE(workspace_bytes >= tgr_mm_proj_bwd_tc_workspace_bytes(T, mm_dim, H), "workspace too small");
  TGR_REQUIRE(((uintptr_t)x_bf16 & 15) == 0 && ((uintptr_t)dz_bf16 & 15) == 0 && ((uintptr_t)workspace & 15) == 0, "mm_proj_bwd_tc: misaligned buffers");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_out = (int64_t)H * mm_dim;
  if (T == 0) {
    if (!accumulate) cudaMemsetAsync(dW, 0, (size_t)n_out * sizeof(float), st);
    return 0;
  }
  const int bn = bwd_tc_bn(mm_dim, H), n_chunks = bwd_tc_chunks(T, mm_dim, H);
  TGR_REQUIRE(dz_planes == 1 || (dz_planes == 2 && plane_rows >= T), "dz_planes must be 1, or 2 with plane_rows >= T");
  CUtensorMap tz, tx;
  // every plane is its own [T, H] view inside the stacked buffer: rows past T of plane 0 must read as zero, so the map covers
  // plane_rows * planes rows and the kernel only asks for token rows < T of each plane (chunks end at T)
  if (int rc = make_map(&tz, dz_bf16, dz_planes == 2 ? plane_rows + T : T, H, 64)) return rc;
  if (int rc = make_map(&tx, x_bf16, T, mm_dim, 64)) return rc;
  float* ws = (float*)workspace;
  int rc;
  if (H == 64) rc = bn == 256 ? launch_bwd_tc<64, 256>(tz, tx, ws, T, mm_dim, n_chunks, dz_planes, plane_rows, st) : (bn == 128 ? launch_bwd_tc<64, 128>(tz, tx, ws, T, mm_dim, n_chunks, dz_planes, plane_rows, st) : launch_bwd_tc<64, 64>(tz, tx, ws, T, mm_dim, n_chunks, dz_planes, plane_rows, st));
  else rc = bn == 128 ? launch_bwd_tc<128, 128>(tz, tx, ws, T, mm_dim, n_chunks, dz_planes, plane_rows, st) : launch_bwd_tc<128, 64>(tz, tx, ws, T, mm_dim, n_chunks, dz_planes, plane_rows, st);
  if (rc) return rc;
  TGR_K(chunk_reduce_kernel)<<<(unsigned)((n_out * 32 + 255) / 256), 256, 0, st>>>(ws, n_chunks, n_out, dW, accumulate);
  return check_launch("mm_proj_bwd_tc_reduce");
}
