// Multimodal-embedding projection (emb_transform[k], model/BaseLine/model.py:167,281-299) written straight
// into / read straight from the item concat buffer.
//
//   forward : out[t, 0:H] = x[t, :] . W^T + b            W is nn.Linear.weight [H, mm_dim]
//   backward: dW[h, k] (+)= sum_t dY[t, h] x[t, k] ;  db[h] (+)= sum_t dY[t, h]      (x is frozen data: no dX)
//
// fp32 path: the forward runs on the tensor cores with error-compensated TF32 (3xTF32 mma.sync, tgr_mma.cuh — single-pass
// TF32 would miss the 1e-5 fp32 bar); the first version's fp32 FFMA kernel is kept for other H and as
// the A/B reference (TGR_MM_FFMA=1). Arithmetic intensity is 2H/4 = 32 flop/B on x, so for mm_dim = 32 ('81', the
// benchmark config) the kernel is HBM/L2-bound on the x read + output write. The bf16 tcgen05 + TMA kernel for the
// 1024..4096-wide features lives in tgr_mm_tc.cu.
#include <stdlib.h>

#include "tgr_common.cuh"
#include "tgr_mma.cuh"

namespace tgr {

constexpr int kMmThreads = 256;
constexpr int kBK = 32;

template <bool BF16>
__device__ __forceinline__ float4 load_x4(const void* x, size_t elem_index) {
  if constexpr (BF16) {
    uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(x) + elem_index));
    return unpack_bf16x4(u);
  } else {
    return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + elem_index));
  }
}

// ------------------------------------------------------------------------------------------------
// forward: BM tokens x BN outputs per CTA, K swept in chunks of kBK through shared memory
// ------------------------------------------------------------------------------------------------
template <int BM, int BN, bool XBF16, bool OBF16>
__global__ void __launch_bounds__(kMmThreads) mm_proj_fwd_kernel(const void* __restrict__ x, int64_t T, int K,
                                                                 const float* __restrict__ W,
                                                                 const float* __restrict__ bias, int H,
                                                                 char* __restrict__ out, int64_t out_ld_bytes) {
  constexpr int CG = BN / 4;               // column groups (float4 each)
  constexpr int RG = kMmThreads / CG;      // row groups
  constexpr int TM = BM / RG;              // tokens per thread
  __shared__ __align__(16) float Xs[BM][kBK + 4];
  __shared__ __align__(16) float Ws[kBK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % CG, ty = tid / CG;
  const int64_t t0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  float acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

  for (int k0 = 0; k0 < K; k0 += kBK) {
    // X tile: BM rows x 32 k (8 float4 per row)
    for (int i = tid; i < BM * (kBK / 4); i += kMmThreads) {
      const int r = i / (kBK / 4), c4 = i % (kBK / 4);
      const int64_t t = t0 + r;
      const int k = k0 + c4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < T) {
        if (k + 3 < K) {
          v = load_x4<XBF16>(x, (size_t)t * K + k);
        } else {
          float tmp[4] = {0.f, 0.f, 0.f, 0.f};
          for (int j = 0; j < 4 && k + j < K; ++j)
            tmp[j] = XBF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[(size_t)t * K + k + j])
                           : reinterpret_cast<const float*>(x)[(size_t)t * K + k + j];
          v = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]);
        }
      }
      *reinterpret_cast<float4*>(&Xs[r][c4 * 4]) = v;
    }
    // W tile, transposed to [k][n]
    for (int i = tid; i < BN * kBK; i += kMmThreads) {
      const int n = i / kBK, k = i % kBK;
      float w = 0.f;
      if (n0 + n < H && k0 + k < K) w = __ldg(W + (size_t)(n0 + n) * K + k0 + k);
      Ws[k][n] = w;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < kBK; ++k) {
      const float4 w = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const float xv = Xs[ty * TM + i][k];
        acc[i][0] = fmaf(xv, w.x, acc[i][0]);
        acc[i][1] = fmaf(xv, w.y, acc[i][1]);
        acc[i][2] = fmaf(xv, w.z, acc[i][2]);
        acc[i][3] = fmaf(xv, w.w, acc[i][3]);
      }
    }
    __syncthreads();
  }
  const int n = n0 + tx * 4;
  if (n < H) {  // H % 4 == 0
    const float4 b = bias ? __ldg(reinterpret_cast<const float4*>(bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int64_t t = t0 + ty * TM + i;
      if (t < T) {
        float4 v = make_float4(acc[i][0] + b.x, acc[i][1] + b.y, acc[i][2] + b.z, acc[i][3] + b.w);
        char* row = out + (size_t)t * out_ld_bytes;
        if constexpr (OBF16)
          st_stream_u2(reinterpret_cast<uint2*>(row) + (n >> 2), pack_bf16x4(v));
        else
          st_stream(reinterpret_cast<float4*>(row) + (n >> 2), v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// forward on the tensor cores: 128 tokens x H outputs per CTA (4 warps x 32 tokens), K swept in chunks of 32;
// x fp32 (split hi/lo after the shared-memory load) or bf16 (exact in tf32: lo = 0), W split once per chunk
// ------------------------------------------------------------------------------------------------
template <int H, bool XBF16, bool OBF16>
__global__ void __launch_bounds__(128) mm_proj_fwd_mma_kernel(const void* __restrict__ x, int64_t T, int K,
                                                              const float* __restrict__ W,
                                                              const float* __restrict__ bias,
                                                              char* __restrict__ out, int64_t out_ld_bytes) {
  constexpr int BM = 128, BK = 32, LD = BK + 4, NTILES = H / 8, NT = 128;
  __shared__ __align__(16) float Xs[BM * LD];
  __shared__ __align__(16) uint32_t Whi[H * LD];
  __shared__ __align__(16) uint32_t Wlo[H * LD];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const int64_t t0 = (int64_t)blockIdx.x * BM;
  float acc[2][NTILES][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NTILES; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
  for (int k0 = 0; k0 < K; k0 += BK) {
    // every 128-bit load of the x tile and of the W tile is issued before the first store (a load -> convert -> store loop
    // exposes one L2 round trip per iteration: 16 of them per CTA in the first version)
    constexpr int XPT = BM * (BK / 4) / NT, WPT = H * (BK / 4) / NT;
    float4 xv[XPT], wv[WPT];
#pragma unroll
    for (int q = 0; q < XPT; ++q) {
      const int i = tid + q * NT, r = i / (BK / 4), c4 = i % (BK / 4);
      const int64_t t = t0 + r;
      const int k = k0 + c4 * 4;
      xv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < T && k < K) xv[q] = load_x4<XBF16>(x, (size_t)t * K + k);   // K % 4 == 0
    }
#pragma unroll
    for (int q = 0; q < WPT; ++q) {
      const int i = tid + q * NT, n = i / (BK / 4), c4 = i % (BK / 4);
      const int k = k0 + c4 * 4;
      wv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < K) wv[q] = __ldg(reinterpret_cast<const float4*>(W + (size_t)n * K + k));
    }
#pragma unroll
    for (int q = 0; q < XPT; ++q) {
      const int i = tid + q * NT, r = i / (BK / 4), c4 = i % (BK / 4);
      *reinterpret_cast<float4*>(Xs + r * LD + c4 * 4) = xv[q];
    }
#pragma unroll
    for (int q = 0; q < WPT; ++q) {
      const int i = tid + q * NT, n = i / (BK / 4), c4 = i % (BK / 4);
      uint4 hi, lo;
      split_tf32(wv[q].x, hi.x, lo.x);
      split_tf32(wv[q].y, hi.y, lo.y);
      split_tf32(wv[q].z, hi.z, lo.z);
      split_tf32(wv[q].w, hi.w, lo.w);
      *reinterpret_cast<uint4*>(Whi + n * LD + c4 * 4) = hi;
      *reinterpret_cast<uint4*>(Wlo + n * LD + c4 * 4) = lo;
    }
    __syncthreads();
    const int m0 = warp * 32;
#pragma unroll
    for (int ks = 0; ks < BK / 8; ++ks) {
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
        for (int mt = 0; mt < 2; ++mt) {
          if constexpr (XBF16) { mma_tf32(acc[mt][nt], ahi[mt], blo); mma_tf32(acc[mt][nt], ahi[mt], bhi); }
          else mma_3xtf32(acc[mt][nt], ahi[mt], alo[mt], bhi, blo);
        }
      }
    }
    __syncthreads();
  }
  const int m0 = warp * 32;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const int64_t ra = t0 + m0 + 16 * mt + g, rb = ra + 8;
#pragma unroll
    for (int nt = 0; nt < NTILES; ++nt) {
      const int n = 8 * nt + 2 * t4;
      const float b0 = bias ? __ldg(bias + n) : 0.f, b1 = bias ? __ldg(bias + n + 1) : 0.f;
      if (ra < T) {
        char* row = out + (size_t)ra * out_ld_bytes;
        if constexpr (OBF16) reinterpret_cast<__nv_bfloat162*>(row)[n >> 1] = __floats2bfloat162_rn(acc[mt][nt][0] + b0, acc[mt][nt][1] + b1);
        else reinterpret_cast<float2*>(row)[n >> 1] = make_float2(acc[mt][nt][0] + b0, acc[mt][nt][1] + b1);
      }
      if (rb < T) {
        char* row = out + (size_t)rb * out_ld_bytes;
        if constexpr (OBF16) reinterpret_cast<__nv_bfloat162*>(row)[n >> 1] = __floats2bfloat162_rn(acc[mt][nt][2] + b0, acc[mt][nt][3] + b1);
        else reinterpret_cast<float2*>(row)[n >> 1] = make_float2(acc[mt][nt][2] + b0, acc[mt][nt][3] + b1);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward: split-T partial products (fixed chunking => deterministic), then an ordered reduction
// ------------------------------------------------------------------------------------------------
constexpr int kBT = 32;   // tokens per smem stage
constexpr int kBKb = 32;  // k columns per CTA
constexpr int kMaxRH = 4; // H <= 128

template <bool XBF16, bool DBF16>
__global__ void __launch_bounds__(kMmThreads) mm_proj_bwd_partial_kernel(const void* __restrict__ x, int64_t T, int K,
                                                                         const char* __restrict__ dy,
                                                                         int64_t dy_ld_bytes, int H, int n_chunks,
                                                                         float* __restrict__ ws_dw,
                                                                         float* __restrict__ ws_db) {
  __shared__ __align__(16) float Ds[kBT][128 + 4];
  __shared__ __align__(16) float Xs[kBT][kBKb + 4];
  const int tid = threadIdx.x;
  const int chunk = blockIdx.x;
  const int k0 = blockIdx.y * kBKb;
  const int64_t per = (T + n_chunks - 1) / n_chunks;
  const int64_t tb = (int64_t)chunk * per;
  const int64_t te = min(T, tb + per);
  const int tk = tid % 8;       // k cols tk*4 .. +3
  const int th = tid / 8;       // 0..31 ; rows th, th+32, th+64, th+96
  float acc[kMaxRH][4];
#pragma unroll
  for (int i = 0; i < kMaxRH; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  float dbacc = 0.f;  // thread tid < H accumulates db[tid] (only blockIdx.y == 0)
  const int H4 = H >> 2;

  for (int64_t t0 = tb; t0 < te; t0 += kBT) {
    for (int i = tid; i < kBT * H4; i += kMmThreads) {
      const int r = i / H4, c4 = i % H4;
      const int64_t t = t0 + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < te) {
        const char* row = dy + (size_t)t * dy_ld_bytes;
        if constexpr (DBF16)
          v = unpack_bf16x4(ld_stream_u2(reinterpret_cast<const uint2*>(row) + c4));
        else
          v = ld_stream(reinterpret_cast<const float4*>(row) + c4);
      }
      *reinterpret_cast<float4*>(&Ds[r][c4 * 4]) = v;
    }
    for (int i = tid; i < kBT * (kBKb / 4); i += kMmThreads) {
      const int r = i / (kBKb / 4), c4 = i % (kBKb / 4);
      const int64_t t = t0 + r;
      const int k = k0 + c4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < te && k + 3 < K) v = load_x4<XBF16>(x, (size_t)t * K + k);
      else if (t < te) {
        float tmp[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < 4 && k + j < K; ++j)
          tmp[j] = XBF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[(size_t)t * K + k + j])
                         : reinterpret_cast<const float*>(x)[(size_t)t * K + k + j];
        v = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]);
      }
      *reinterpret_cast<float4*>(&Xs[r][c4 * 4]) = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < kBT; ++r) {
      const float4 xv = *reinterpret_cast<const float4*>(&Xs[r][tk * 4]);
#pragma unroll
      for (int i = 0; i < kMaxRH; ++i) {
        const int h = th + 32 * i;
        if (h < H) {
          const float d = Ds[r][h];
          acc[i][0] = fmaf(d, xv.x, acc[i][0]);
          acc[i][1] = fmaf(d, xv.y, acc[i][1]);
          acc[i][2] = fmaf(d, xv.z, acc[i][2]);
          acc[i][3] = fmaf(d, xv.w, acc[i][3]);
        }
      }
    }
    if (blockIdx.y == 0 && tid < H) {
#pragma unroll 4
      for (int r = 0; r < kBT; ++r) dbacc += Ds[r][tid];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < kMaxRH; ++i) {
    const int h = th + 32 * i;
    const int k = k0 + tk * 4;
    if (h < H) {
      float* dst = ws_dw + ((size_t)chunk * H + h) * K + k;
      for (int j = 0; j < 4; ++j)
        if (k + j < K) dst[j] = acc[i][j];
    }
  }
  if (blockIdx.y == 0 && tid < H) ws_db[(size_t)chunk * H + tid] = dbacc;
}

// one warp per output element: lanes stride the chunk partials, then a fixed shuffle tree (deterministic)
__global__ void __launch_bounds__(256) mm_proj_bwd_reduce_kernel(const float* __restrict__ ws_dw,
                                                                 const float* __restrict__ ws_db, int n_chunks, int H,
                                                                 int K, float* __restrict__ dW, float* __restrict__ db,
                                                                 int accumulate) {
  const int lane = threadIdx.x & 31;
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nW = (int64_t)H * K;
  if (i >= nW + H) return;
  const float* src = i < nW ? ws_dw + i : ws_db + (i - nW);
  const int64_t stride = i < nW ? nW : H;
  float s = 0.f;
  for (int c = lane; c < n_chunks; c += 32) s += src[(size_t)c * stride];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    if (i < nW) dW[i] = accumulate ? dW[i] + s : s;
    else if (db) db[i - nW] = accumulate ? db[i - nW] + s : s;
  }
}

static int bwd_chunks(int64_t T, int K) {
  const int ky = (K + kBKb - 1) / kBKb;
  int n = (kNumSMs * 4 + ky - 1) / ky;
  const int64_t max_by_T = (T + kBT - 1) / kBT;
  if (n > max_by_T) n = (int)max_by_T;
  if (n < 1) n = 1;
  return n;
}

}  // namespace tgr

extern "C" int tgr_mm_proj_fwd(const void* x, int x_dtype, int64_t T, int mm_dim, const float* W, const float* bias,
                               int H, void* out, int64_t out_ld, int out_dtype, void* stream) {
  tgr::TimedScope tgr_timed_("mm_proj_fwd", stream);
  using namespace tgr;
  TGR_REQUIRE(x && W && out, "null argument");
  TGR_REQUIRE(H > 0 && H % 4 == 0 && mm_dim > 0, "bad H=%d / mm_dim=%d", H, mm_dim);
  TGR_REQUIRE(out_ld % 4 == 0, "out_ld must be a multiple of 4 elements");
  TGR_REQUIRE((x_dtype == TGR_DTYPE_F32 && mm_dim % 4 == 0) || (x_dtype == TGR_DTYPE_BF16 && mm_dim % 4 == 0),
              "mm_dim must be a multiple of 4");
  TGR_REQUIRE(((uintptr_t)W & 15) == 0, "W must be 16-byte aligned");
  if (T == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t ldb = out_ld * (out_dtype == TGR_DTYPE_BF16 ? 2 : 4);
  const bool xb = x_dtype == TGR_DTYPE_BF16, ob = out_dtype == TGR_DTYPE_BF16;
  static const bool ffma = [] { const char* e = getenv("TGR_MM_FFMA"); return e && e[0] == '1'; }();
  if (!ffma && (H == 32 || H == 64)) {
    const unsigned grid = (unsigned)((T + 127) / 128);
#define TGR_LAUNCH_MMA(HH)                                                                                                     \
    do {                                                                                                                       \
      if (xb && ob) TGR_K(mm_proj_fwd_mma_kernel<HH, true, true>)<<<grid, 128, 0, st>>>(x, T, mm_dim, W, bias, (char*)out, ldb);   \
      else if (xb) TGR_K(mm_proj_fwd_mma_kernel<HH, true, false>)<<<grid, 128, 0, st>>>(x, T, mm_dim, W, bias, (char*)out, ldb);   \
      else if (ob) TGR_K(mm_proj_fwd_mma_kernel<HH, false, true>)<<<grid, 128, 0, st>>>(x, T, mm_dim, W, bias, (char*)out, ldb);   \
      else TGR_K(mm_proj_fwd_mma_kernel<HH, false, false>)<<<grid, 128, 0, st>>>(x, T, mm_dim, W, bias, (char*)out, ldb);          \
    } while (0)
    if (H == 32) TGR_LAUNCH_MMA(32);
    else TGR_LAUNCH_MMA(64);
#undef TGR_LAUNCH_MMA
    return check_launch("mm_proj_fwd");
  }
#define TGR_LAUNCH_FWD(BM, BN, XB, OB)                                                                  \
  TGR_K(mm_proj_fwd_kernel<BM, BN, XB, OB>)<<<dim3((unsigned)((T + BM - 1) / BM), (H + BN - 1) / BN), kMmThreads, 0, st>>>( \
      x, T, mm_dim, W, bias, H, (char*)out, ldb)
  if (H >= 64) {
    if (xb && ob) TGR_LAUNCH_FWD(128, 64, true, true);
    else if (xb) TGR_LAUNCH_FWD(128, 64, true, false);
    else if (ob) TGR_LAUNCH_FWD(128, 64, false, true);
    else TGR_LAUNCH_FWD(128, 64, false, false);
  } else {
    if (xb && ob) TGR_LAUNCH_FWD(128, 32, true, true);
    else if (xb) TGR_LAUNCH_FWD(128, 32, true, false);
    else if (ob) TGR_LAUNCH_FWD(128, 32, false, true);
    else TGR_LAUNCH_FWD(128, 32, false, false);
  }
#undef TGR_LAUNCH_FWD
  return check_launch("mm_proj_fwd");
}

extern "C" size_t tgr_mm_proj_bwd_workspace_bytes(int64_t T, int mm_dim, int H) {
  const int n = tgr::bwd_chunks(T, mm_dim);
  return ((size_t)n * H * mm_dim + (size_t)n * H) * sizeof(float) + 256;
}

extern "C" int tgr_mm_proj_bwd(const void* x, int x_dtype, int64_t T, int mm_dim, const void* dy, int64_t dy_ld,
                               int dy_dtype, int H, float* dW, float* db, int accumulate, void* workspace,
                               size_t workspace_bytes, void* stream) {
  tgr::TimedScope tgr_timed_("mm_proj_bwd", stream);
  using namespace tgr;
  TGR_REQUIRE(x && dy && dW && workspace, "null argument");
  TGR_REQUIRE(H > 0 && H % 4 == 0 && H <= 128, "mm backward supports H <= 128 (H=%d)", H);
  TGR_REQUIRE(mm_dim % 4 == 0 && dy_ld % 4 == 0, "mm_dim / dy_ld must be multiples of 4");
  TGR_REQUIRE(workspace_bytes >= tgr_mm_proj_bwd_workspace_bytes(T, mm_dim, H), "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int n = bwd_chunks(T, mm_dim);
  float* ws_dw = (float*)workspace;
  float* ws_db = ws_dw + (size_t)n * H * mm_dim;
  const int64_t ldb = dy_ld * (dy_dtype == TGR_DTYPE_BF16 ? 2 : 4);
  const bool xb = x_dtype == TGR_DTYPE_BF16, dbf = dy_dtype == TGR_DTYPE_BF16;
  dim3 grid(n, (mm_dim + kBKb - 1) / kBKb);
  if (T > 0) {
    if (xb && dbf) TGR_K(mm_proj_bwd_partial_kernel<true, true>)<<<grid, kMmThreads, 0, st>>>(x, T, mm_dim, (const char*)dy, ldb, H, n, ws_dw, ws_db);
    else if (xb) TGR_K(mm_proj_bwd_partial_kernel<true, false>)<<<grid, kMmThreads, 0, st>>>(x, T, mm_dim, (const char*)dy, ldb, H, n, ws_dw, ws_db);
    else if (dbf) TGR_K(mm_proj_bwd_partial_kernel<false, true>)<<<grid, kMmThreads, 0, st>>>(x, T, mm_dim, (const char*)dy, ldb, H, n, ws_dw, ws_db);
    else TGR_K(mm_proj_bwd_partial_kernel<false, false>)<<<grid, kMmThreads, 0, st>>>(x, T, mm_dim, (const char*)dy, ldb, H, n, ws_dw, ws_db);
    if (int rc = check_launch("mm_proj_bwd_partial")) return rc;
  }
  const int64_t total = (int64_t)H * mm_dim + H;
  TGR_K(mm_proj_bwd_reduce_kernel)<<<(unsigned)((total * 32 + 255) / 256), 256, 0, st>>>(ws_dw, ws_db, T > 0 ? n : 0, H, mm_dim, dW, db, accumulate);
  return check_launch("mm_proj_bwd_reduce");
}
