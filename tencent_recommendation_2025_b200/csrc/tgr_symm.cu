// Peer-memory transport of the row-sharded exchange: the small messages of a sharded step (per-owner counts, bucketed
// local-row ids, replicated dense gradients) move by kernels that store to / load from the other ranks' symmetric-memory
// buffers over NVLink, ordered by device-side barriers (signal pads) — no NCCL collective, no host round trip, no
// data-dependent split sizes on the host's critical path. Round 1 used NCCL for these (all-gather of counts, all-to-all of
// ids, a 4-byte all-reduce as barrier, ring all-reduce of 420 KB of dense gradients): ~0.25 ms of exposed latency per step
// at 2-8 GPUs for < 2 MB of payload ("scaling limiter is latency and the host, not NVLink").
// The owner-side order of the received ids is produced by a W-way MERGE of the per-source buckets (each arrives sorted),
// not by a second radix sort.
#include "tgr_common.cuh"

namespace tgr {

struct PeerBases { void* p[TGR_MAX_PEERS]; };
struct PullSegs {
  const uint32_t* src[TGR_MAX_PEERS];   // peer s' bucketed-row buffer, already offset to this rank's bucket
  int64_t first[TGR_MAX_PEERS + 1];     // prefix of the counts (positions in dst)
};

// dst_p[rank * n + i] = src[i] for every peer p (n small: W counts)
__global__ void peer_put_kernel(const __grid_constant__ PeerBases dst, int W, int rank, const int32_t* __restrict__ src, int n) {
  const int p = blockIdx.x;
  int32_t* d = reinterpret_cast<int32_t*>(dst.p[p]) + (size_t)rank * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = src[i];
  __threadfence_system();
}

// dst = concat_s src[s][0 : cnt_s]   (remote loads, 4 in flight per thread)
__global__ void __launch_bounds__(256) peer_pull_kernel(const __grid_constant__ PullSegs seg, int W, uint32_t* __restrict__ dst) {
  const int s = blockIdx.y;
  const int64_t n = seg.first[s + 1] - seg.first[s];
  const uint32_t* src = seg.src[s];
  uint32_t* d = dst + seg.first[s];
  for (int64_t i0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4; i0 < n; i0 += (int64_t)gridDim.x * 256 * 4) {
    uint32_t v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = i0 + j < n ? src[i0 + j] : 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (i0 + j < n) d[i0 + j] = v[j];
  }
}

struct MergeSegs { int64_t first[TGR_MAX_PEERS + 1]; };

// Stable W-way merge of sorted buckets by rank counting: element i of bucket s with key k lands at
//   i + sum_{s' < s} |{x in bucket s' : x <= k}| + sum_{s' > s} |{x in bucket s' : x < k}|
// — exactly the position a stable sort by key of the concatenation gives it (ties keep source-rank order).
__global__ void __launch_bounds__(256) merge_buckets_kernel(const uint32_t* __restrict__ rows, const __grid_constant__ MergeSegs seg,
                                                            int W, int with_code, uint32_t* __restrict__ keys_out,
                                                            uint32_t* __restrict__ code_out) {
  const int64_t R = seg.first[W];
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < R; e += (int64_t)gridDim.x * 256) {
    int s = 0;
    while (e >= seg.first[s + 1]) ++s;
    const int64_t i = e - seg.first[s];
    const uint32_t k = __ldg(rows + e);
    int64_t pos = i;
    for (int q = 0; q < W; ++q) {
      if (q == s) continue;
      int64_t lo = seg.first[q], hi = seg.first[q + 1];
      if (q < s) { while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (__ldg(rows + mid) <= k) lo = mid + 1; else hi = mid; } }
      else       { while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (__ldg(rows + mid) < k) lo = mid + 1; else hi = mid; } }
      pos += lo - seg.first[q];
    }
    keys_out[pos] = k;
    code_out[pos] = with_code ? ((uint32_t)s << 24) | (uint32_t)i : (uint32_t)e;
  }
}

struct PeerF32 { const float* p[TGR_MAX_PEERS]; };
// out[i] = scale * (p_0[i] + p_1[i] + ... + p_{W-1}[i]): every rank adds in rank order => identical results everywhere
__global__ void __launch_bounds__(256) allreduce_peers_kernel(const __grid_constant__ PeerF32 peers, int W, int64_t n4, float scale,
                                                              float4* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    float4 v[TGR_MAX_PEERS];
#pragma unroll
    for (int r = 0; r < TGR_MAX_PEERS; ++r)
      if (r < W) v[r] = ld_stream(reinterpret_cast<const float4*>(peers.p[r]) + i);
    float4 s = v[0];
#pragma unroll
    for (int r = 1; r < TGR_MAX_PEERS; ++r)
      if (r < W) s = f4_add(s, v[r]);
    out[i] = make_float4(s.x * scale, s.y * scale, s.z * scale, s.w * scale);
  }
}

}  // namespace tgr

using namespace tgr;

extern "C" int tgr_peer_put(void* const* dst_bases, int n_peers, int rank, const int32_t* src, int n, void* stream) {
  tgr::TimedScope tgr_timed_("peer_put", stream);
  TGR_REQUIRE(dst_bases && src && n_peers > 0 && n_peers <= TGR_MAX_PEERS && rank >= 0 && rank < n_peers && n > 0, "bad argument");
  PeerBases b{};
  for (int p = 0; p < n_peers; ++p) { TGR_REQUIRE(dst_bases[p] != nullptr, "peer %d: NULL", p); b.p[p] = dst_bases[p]; }
  TGR_K(peer_put_kernel)<<<n_peers, 64, 0, (cudaStream_t)stream>>>(b, n_peers, rank, src, n);
  return check_launch("peer_put");
}

extern "C" int tgr_peer_pull(const uint32_t* const* src_ptrs, const int64_t* counts, int n_peers, uint32_t* dst, void* stream) {
  tgr::TimedScope tgr_timed_("peer_pull", stream);
  TGR_REQUIRE(src_ptrs && counts && dst && n_peers > 0 && n_peers <= TGR_MAX_PEERS, "bad argument");
  PullSegs seg{};
  int64_t tot = 0, mx = 0;
  for (int s = 0; s < n_peers; ++s) {
    TGR_REQUIRE(counts[s] >= 0 && (counts[s] == 0 || src_ptrs[s] != nullptr), "peer %d: bad segment", s);
    seg.src[s] = src_ptrs[s];
    seg.first[s] = tot;
    tot += counts[s];
    if (counts[s] > mx) mx = counts[s];
  }
  seg.first[n_peers] = tot;
  if (tot == 0) return 0;
  int64_t bx = (mx + 1023) / 1024;
  if (bx > 64) bx = 64;
  TGR_K(peer_pull_kernel)<<<dim3((unsigned)bx, n_peers), 256, 0, (cudaStream_t)stream>>>(seg, n_peers, dst);
  return check_launch("peer_pull");
}

extern "C" int tgr_merge_buckets(const uint32_t* rows, const int64_t* counts, int n_buckets, int with_code, uint32_t* keys_out,
                                 uint32_t* code_out, void* stream) {
  tgr::TimedScope tgr_timed_("merge_buckets", stream);
  TGR_REQUIRE(rows && counts && keys_out && code_out && n_buckets > 0 && n_buckets <= TGR_MAX_PEERS, "bad argument");
  MergeSegs seg{};
  int64_t tot = 0;
  for (int s = 0; s < n_buckets; ++s) {
    TGR_REQUIRE(counts[s] >= 0 && counts[s] < (1ll << 24), "bucket %d: count out of range", s);
    seg.first[s] = tot;
    tot += counts[s];
  }
  seg.first[n_buckets] = tot;
  if (tot == 0) return 0;
  int64_t blocks = (tot + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  TGR_K(merge_buckets_kernel)<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rows, seg, n_buckets, with_code, keys_out, code_out);
  return check_launch("merge_buckets");
}

// Device-side barrier between the W ranks over a symmetric flag array (uint32 [>= W] per rank, zero-initialised; every rank
// passes the same monotonically increasing epoch): thread r publishes everything this rank's stream has written so far
// (system-scope fence + release store of the epoch into slot [rank] of rank r's array) and waits until rank r's epoch has
// arrived in its own array. torch's symmetric-memory barrier does the same on its signal pads but costs ~0.12 ms of HOST time
// per call (three per sharded step, tools/profile_sharded.py CPROFILE=1). One CTA; the wait is bounded (~20 s) and traps.
namespace tgr {
struct PeerU32 { uint32_t* p[TGR_MAX_PEERS]; };
__global__ void __launch_bounds__(32) peer_barrier_kernel(const __grid_constant__ PeerU32 flags, int rank, int W, uint32_t epoch) {
  const int r = threadIdx.x;
  if (r < W) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags.p[r] + rank), "r"(epoch) : "memory");
    const uint32_t* mine = flags.p[rank] + r;
    const long long t0 = clock64();
    uint32_t v;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((int32_t)(v - epoch) >= 0) break;
      if (clock64() - t0 > 40000000000ll) __trap();   // a peer never arrived
      __nanosleep(64);
    }
  }
  __syncwarp();
  __threadfence_system();
}
}  // namespace tgr

extern "C" int tgr_peer_barrier(uint32_t* const* flags, int rank, int n_peers, uint32_t epoch, void* stream) {
  tgr::TimedScope tgr_timed_("peer_barrier", stream);
  TGR_REQUIRE(flags && n_peers > 0 && n_peers <= TGR_MAX_PEERS && n_peers <= 32 && rank >= 0 && rank < n_peers, "bad argument");
  PeerU32 f{};
  for (int r = 0; r < n_peers; ++r) {
    TGR_REQUIRE(flags[r] != nullptr, "peer %d: NULL flag array", r);
    f.p[r] = flags[r];
  }
  TGR_K(peer_barrier_kernel)<<<1, 32, 0, (cudaStream_t)stream>>>(f, rank, n_peers, epoch);
  return check_launch("peer_barrier");
}

extern "C" int tgr_allreduce_peers(const float* const* peers, int n_peers, int64_t n, float scale, float* out, void* stream) {
  tgr::TimedScope tgr_timed_("allreduce_peers", stream);
  TGR_REQUIRE(peers && out && n_peers > 0 && n_peers <= TGR_MAX_PEERS && n >= 0 && n % 4 == 0, "bad argument (n must be a multiple of 4)");
  if (n == 0) return 0;
  PeerF32 pp{};
  for (int r = 0; r < n_peers; ++r) {
    TGR_REQUIRE(peers[r] != nullptr && ((uintptr_t)peers[r] & 15) == 0, "peer %d: NULL / misaligned", r);
    pp.p[r] = peers[r];
  }
  TGR_REQUIRE(((uintptr_t)out & 15) == 0, "out misaligned");
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
  TGR_K(allreduce_peers_kernel)<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pp, n_peers, n / 4, scale, (float4*)out);
  return check_launch("allreduce_peers");
}
