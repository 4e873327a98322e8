// Stable LSD radix sort of (key, payload) uint32 pairs on the low `key_bits` bits — hand-written for this path.
//
// The pairs are (global row key, source code) of every non-padding lookup of a step (2.85 M at the C2 benchmark,
// keys < 2^24) or (local row, source) on the owner side of the sharded exchange (0.2 M). A generic 8-bit onesweep sort
// needs 3 passes + a histogram pass and is launch/latency-bound at these sizes (128 us per step with the library sort
// the first version used; this one: 122 us at 2.85 M pairs / 24 bits, 53 us at 0.23 M / 21 bits on B200); here:
//   * 12-bit digits: 24-bit keys sort in TWO passes (key_bits <= 12: one), digits sized ceil(key_bits / passes);
//   * one tile per SM-sized block (n / 148 rounded to 256), so a pass is one wave;
//   * pass = block histogram (shared-memory integer atomics) -> per-digit scan over blocks (one warp per digit) ->
//     scan of the 4096 digit totals -> scatter;
//   * scatter keeps the sort STABLE without a local sort: every warp owns a contiguous chunk and walks it in order,
//     32 entries per round; `match.any` groups the lanes with equal digits (one ballot per digit bit measured 8 %
//     slower, tools/sort_bench.py), a lane's position is
//     base[digit] + running[warp][digit] + rank inside its group, and the group's leader advances
//     running[warp][digit]. A block's running offsets are 16-bit (relative to the block's digit base) and live in
//     shared memory: 16 warps x 4096 digits x 2 B = 128 KB + 16 KB of bases out of the SM's 227 KB; keys are
//     register-prefetched four rounds ahead so the in-order walk is not a chain of dependent global loads.
// Stability is what keeps each row's gradient contributions in ascending (call, token) order — the summation order
// of embedding_dense_backward — and makes the backward bitwise reproducible.
#include "tgr_common.cuh"
#include "tgr_rows.cuh"

namespace tgr {

constexpr int kSortThreads = 512;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortMaxBits = 12;
constexpr int kSortMaxTile = 32768;   // 16-bit in-block offsets
constexpr int kSortAhead = 4;         // rounds of 32 entries prefetched into registers

struct SortGeom {
  int tile;      // entries per block, multiple of 512 (=> every warp chunk is a multiple of 32)
  int n_blocks;
};

__host__ __device__ inline int sort_tile_of(int64_t n) {
  int64_t tile = (n + kNumSMs - 1) / kNumSMs;
  tile = (tile + kSortThreads - 1) / kSortThreads * kSortThreads;
  if (tile < kSortThreads) tile = kSortThreads;
  if (tile > kSortMaxTile) tile = kSortMaxTile;
  return (int)tile;
}

// n_blocks is the LAUNCHED grid and the row pitch of the block histograms. It covers every n' <= n: when the entry count
// is only known on the device (n = capacity, the kernels read n' and derive the tile from it) a smaller n' can need MORE
// blocks than n does (tile(n') shrinks in steps of 512), but never more than min(SMs, ceil(n / 512)) while one wave
// suffices. Blocks past the data see an empty range and write zero histograms.
static SortGeom sort_geom(int64_t n) {
  SortGeom g;
  g.tile = sort_tile_of(n);
  int64_t nb = (n + g.tile - 1) / g.tile;
  const int64_t wave = (n + kSortThreads - 1) / kSortThreads < kNumSMs ? (n + kSortThreads - 1) / kSortThreads : kNumSMs;
  if (nb < wave) nb = wave;
  if (nb < 1) nb = 1;
  g.n_blocks = (int)nb;
  return g;
}

template <bool IN_PAIRS>   // keys[i] or the .x of interleaved (key, payload) pairs
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const uint32_t* __restrict__ keys, int n, int tile,
                                                                  int shift, int bits, int nb,
                                                                  int32_t* __restrict__ block_hist /*[R][nb]*/,
                                                                  const int32_t* __restrict__ n_dev) {
  extern __shared__ int32_t sh[];   // [R]
  if (n_dev) { n = min(n, __ldg(n_dev)); tile = sort_tile_of(n); }   // count known on the device only (n = capacity)
  const int R = 1 << bits;
  const uint32_t mask = (uint32_t)R - 1u;
  for (int i = threadIdx.x; i < R; i += kSortThreads) sh[i] = 0;
  __syncthreads();
  const int a = blockIdx.x * tile, b = min(n, a + tile);
  for (int i0 = a + threadIdx.x; i0 < b; i0 += kSortAhead * kSortThreads) {
    uint32_t k[kSortAhead];
#pragma unroll
    for (int j = 0; j < kSortAhead; ++j)
      k[j] = i0 + j * kSortThreads < b ? __ldg(keys + (size_t)(i0 + j * kSortThreads) * (IN_PAIRS ? 2 : 1)) : 0u;
#pragma unroll
    for (int j = 0; j < kSortAhead; ++j)
      if (i0 + j * kSortThreads < b) atomicAdd(&sh[(k[j] >> shift) & mask], 1);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < R; d += kSortThreads) block_hist[(size_t)d * nb + blockIdx.x] = sh[d];
}

// one warp per digit: exclusive prefix of the digit's counts over the blocks (in place) + the digit's total
__global__ void __launch_bounds__(kSortThreads) radix_binscan_kernel(int32_t* __restrict__ block_hist, int nb, int R,
                                                                     int32_t* __restrict__ bin_total) {
  const int bin = (blockIdx.x * kSortThreads + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (bin >= R) return;
  int32_t* row = block_hist + (size_t)bin * nb;
  int carry = 0;
  for (int j0 = 0; j0 < nb; j0 += 32) {
    const int j = j0 + lane;
    const int v = j < nb ? row[j] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (j < nb) row[j] = carry + x - v;
    carry += __shfl_sync(0xffffffffu, x, 31);
  }
  if (lane == 0) bin_total[bin] = carry;
}

// exclusive scan of the R <= 4096 digit totals (single CTA of 1024 threads, 4 per thread)
__global__ void __launch_bounds__(1024) radix_totals_kernel(const int32_t* __restrict__ bin_total, int R,
                                                            int32_t* __restrict__ bin_base) {
  __shared__ int32_t warp_sum[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int v[4], s = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = threadIdx.x * 4 + k;
    v[k] = i < R ? bin_total[i] : 0;
    s += v[k];
  }
  int x = s;
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
    warp_sum[lane] = w;   // inclusive over warps
  }
  __syncthreads();
  int run = x - s + (wid ? warp_sum[wid - 1] : 0);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = threadIdx.x * 4 + k;
    if (i < R) bin_base[i] = run;
    run += v[k];
  }
}

// IN_PAIRS / OUT_PAIRS: the pass reads / writes interleaved (key, payload) pairs (keys_in / keys_out point at uint2[n],
// vals_* unused). The intermediate buffer of a two-pass sort is interleaved: the first pass scatters its entries to 4096 bins
// in effectively random order, so every entry is its own partial-sector write — one 8-byte write per entry instead of two
// 4-byte ones halves the L2 write transactions of that pass.
template <bool IN_PAIRS, bool OUT_PAIRS>
__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const uint32_t* __restrict__ keys_in,
                                                                     const uint32_t* __restrict__ vals_in,
                                                                     uint32_t* __restrict__ keys_out,
                                                                     uint32_t* __restrict__ vals_out, int n, int tile,
                                                                     int shift, int bits, int nb,
                                                                     const int32_t* __restrict__ block_off /*[R][nb]*/,
                                                                     const int32_t* __restrict__ bin_base /*[R]*/,
                                                                     const int32_t* __restrict__ n_dev) {
  extern __shared__ __align__(16) unsigned char sort_smem[];
  if (n_dev) { n = min(n, __ldg(n_dev)); tile = sort_tile_of(n); }
  const int R = 1 << bits;
  int32_t* base = reinterpret_cast<int32_t*>(sort_smem);                  // [R] first output position of (block, digit)
  uint16_t* wh = reinterpret_cast<uint16_t*>(sort_smem + (size_t)R * 4);  // [kSortWarps][R] counts, then running offsets
  const uint32_t mask = (uint32_t)R - 1u;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kSortWarps * R / 2; i += kSortThreads) reinterpret_cast<uint32_t*>(wh)[i] = 0u;
  __syncthreads();
  const int chunk = tile / kSortWarps;
  const int a = min(n, blockIdx.x * tile + w * chunk), b = min(n, a + chunk);
  uint16_t* mine = wh + w * R;
  // phase 1: this warp's digit histogram — integer shared-memory atomics on the 32-bit word holding two 16-bit
  // counters (counts stay below 2^15, so the halves never carry into each other); order does not matter for counts
  {
    unsigned* mine32 = reinterpret_cast<unsigned*>(mine);
    for (int i0 = a; i0 < b; i0 += 32 * kSortAhead) {
      uint32_t k[kSortAhead];
#pragma unroll
      for (int j = 0; j < kSortAhead; ++j)
        k[j] = i0 + 32 * j + lane < b ? __ldg(keys_in + (size_t)(i0 + 32 * j + lane) * (IN_PAIRS ? 2 : 1)) : 0u;
#pragma unroll
      for (int j = 0; j < kSortAhead; ++j) {
        if (i0 + 32 * j + lane < b) {
          const uint32_t d = (k[j] >> shift) & mask;
          atomicAdd(mine32 + (d >> 1), (d & 1u) ? 65536u : 1u);
        }
      }
    }
  }
  __syncthreads();
  // phase 2: counts -> offset of (warp, digit) inside the block's digit run; base = digit base + earlier blocks
  for (int d = threadIdx.x; d < R; d += kSortThreads) {
    base[d] = __ldg(bin_base + d) + __ldg(block_off + (size_t)d * nb + blockIdx.x);
    unsigned run = 0;
#pragma unroll
    for (int q = 0; q < kSortWarps; ++q) {
      const unsigned t = wh[q * R + d];
      wh[q * R + d] = (uint16_t)run;
      run += t;
    }
  }
  __syncthreads();
  // phase 3: scatter, same walk
  for (int i0 = a; i0 < b; i0 += 32 * kSortAhead) {
    uint32_t k[kSortAhead], v[kSortAhead];
#pragma unroll
    for (int j = 0; j < kSortAhead; ++j) {
      const int i = i0 + 32 * j + lane;
      if constexpr (IN_PAIRS) {
        const uint2 kv = i < b ? __ldg(reinterpret_cast<const uint2*>(keys_in) + i) : make_uint2(0u, 0u);
        k[j] = kv.x;
        v[j] = kv.y;
      } else {
        k[j] = i < b ? __ldg(keys_in + i) : 0u;
        v[j] = i < b ? __ldg(vals_in + i) : 0u;
      }
    }
#pragma unroll
    for (int j = 0; j < kSortAhead; ++j) {
      const bool ok = i0 + 32 * j + lane < b;
      const uint32_t d = ok ? ((k[j] >> shift) & mask) : (uint32_t)R;
      const unsigned m = __match_any_sync(0xffffffffu, d);
      int pos = 0;
      if (ok) pos = base[d] + (int)mine[d] + __popc(m & ((1u << lane) - 1u));
      __syncwarp();
      if (ok && lane == __ffs(m) - 1) mine[d] = (uint16_t)(mine[d] + __popc(m));
      __syncwarp();
      if (ok) {
        if constexpr (OUT_PAIRS) reinterpret_cast<uint2*>(keys_out)[pos] = make_uint2(k[j], v[j]);
        else { keys_out[pos] = k[j]; vals_out[pos] = v[j]; }
      }
    }
  }
}

static int sort_passes(int key_bits) { return (key_bits + kSortMaxBits - 1) / kSortMaxBits; }

}  // namespace tgr

using namespace tgr;

// temp (key, payload) pair buffer for multi-pass sorts + block histograms [4096][n_blocks] + digit totals / bases
extern "C" size_t tgr_sort_workspace_bytes(int64_t n) {
  if (n < 1) n = 1;
  const SortGeom g = sort_geom(n);
  return 2 * align_up((size_t)n * 4) + align_up((size_t)(1 << kSortMaxBits) * g.n_blocks * 4) +
         2 * align_up((size_t)(1 << kSortMaxBits) * 4);
}

extern "C" int tgr_sort_pairs(const uint32_t* keys_in, const uint32_t* srcs_in, uint32_t* keys_out, uint32_t* srcs_out,
                              int64_t n, int key_bits, void* workspace, size_t workspace_bytes, void* stream) {
  return tgr::sort_pairs_dn(keys_in, srcs_in, keys_out, srcs_out, n, key_bits, workspace, workspace_bytes, nullptr, stream);
}

// n_dev != NULL: n is a capacity, the entry count is *n_dev (<= n) when the kernels run (CUDA-graph replay)
int tgr::sort_pairs_dn(const uint32_t* keys_in, const uint32_t* srcs_in, uint32_t* keys_out, uint32_t* srcs_out,
                       int64_t n, int key_bits, void* workspace, size_t workspace_bytes, const int32_t* n_dev, void* stream) {
  tgr::TimedScope tgr_timed_("sort_pairs", stream);
  TGR_REQUIRE(n >= 0 && n < (1ll << 31), "n out of range");
  TGR_REQUIRE(key_bits > 0 && key_bits <= 32, "key_bits=%d out of range", key_bits);
  if (n == 0) return 0;
  TGR_REQUIRE(keys_in && srcs_in && keys_out && srcs_out && workspace, "null argument");
  TGR_REQUIRE(keys_in != keys_out && srcs_in != srcs_out, "sort_pairs is out of place");
  TGR_REQUIRE(workspace_bytes >= tgr_sort_workspace_bytes(n), "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const SortGeom g = sort_geom(n);
  char* ws = (char*)workspace;
  uint32_t* keys_tmp = (uint32_t*)ws;
  uint32_t* srcs_tmp = (uint32_t*)(ws + align_up((size_t)n * 4));
  int32_t* block_hist = (int32_t*)(ws + 2 * align_up((size_t)n * 4));
  int32_t* bin_total = (int32_t*)((char*)block_hist + align_up((size_t)(1 << kSortMaxBits) * g.n_blocks * 4));
  int32_t* bin_base = (int32_t*)((char*)bin_total + align_up((size_t)(1 << kSortMaxBits) * 4));
  const int P = sort_passes(key_bits);
  const int bits0 = (key_bits + P - 1) / P;
  const int kScatterSmem = (1 << kSortMaxBits) * 4 + kSortWarps * (1 << kSortMaxBits) * 2;
  { static bool once = false;
    if (!once) {
      cudaFuncSetAttribute(radix_scatter_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kScatterSmem);
      cudaFuncSetAttribute(radix_scatter_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kScatterSmem);
      cudaFuncSetAttribute(radix_scatter_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kScatterSmem);
      once = true; } }
  const uint32_t* ksrc = keys_in;
  const uint32_t* vsrc = srcs_in;
  int shift = 0;
  const bool pairs = P == 2;   // the intermediate buffer holds interleaved pairs (keys_tmp | srcs_tmp are contiguous: uint2[n])
  for (int p = 0; p < P; ++p) {
    const int bits = min(bits0, key_bits - shift);
    const int R = 1 << bits;
    const bool to_out = ((P - 1 - p) % 2) == 0;   // the last pass lands in the caller's output
    uint32_t* kdst = to_out ? keys_out : keys_tmp;
    uint32_t* vdst = to_out ? srcs_out : srcs_tmp;
    const bool in_pairs = pairs && p == 1, out_pairs = pairs && p == 0;
    if (in_pairs) TGR_K(radix_hist_kernel<true>)<<<g.n_blocks, kSortThreads, (size_t)R * 4, st>>>(ksrc, (int)n, g.tile, shift, bits, g.n_blocks, block_hist, n_dev);
    else TGR_K(radix_hist_kernel<false>)<<<g.n_blocks, kSortThreads, (size_t)R * 4, st>>>(ksrc, (int)n, g.tile, shift, bits, g.n_blocks, block_hist, n_dev);
    TGR_K(radix_binscan_kernel)<<<(R * 32 + kSortThreads - 1) / kSortThreads, kSortThreads, 0, st>>>(block_hist, g.n_blocks, R, bin_total);
    TGR_K(radix_totals_kernel)<<<1, 1024, 0, st>>>(bin_total, R, bin_base);
    const size_t smem = (size_t)R * 4 + (size_t)kSortWarps * R * 2;
    if (in_pairs)
      TGR_K(radix_scatter_kernel<true, false>)<<<g.n_blocks, kSortThreads, smem, st>>>(ksrc, vsrc, kdst, vdst, (int)n, g.tile, shift, bits, g.n_blocks, block_hist, bin_base, n_dev);
    else if (out_pairs)
      TGR_K(radix_scatter_kernel<false, true>)<<<g.n_blocks, kSortThreads, smem, st>>>(ksrc, vsrc, kdst, vdst, (int)n, g.tile, shift, bits, g.n_blocks, block_hist, bin_base, n_dev);
    else
      TGR_K(radix_scatter_kernel<false, false>)<<<g.n_blocks, kSortThreads, smem, st>>>(ksrc, vsrc, kdst, vdst, (int)n, g.tile, shift, bits, g.n_blocks, block_hist, bin_base, n_dev);
    if (int rc = check_launch("sort_pairs")) return rc;
    ksrc = kdst;
    vsrc = vdst;
    shift += bits;
  }
  return 0;
}
