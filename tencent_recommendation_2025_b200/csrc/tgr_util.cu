// Error plumbing and version of the C ABI.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "tgr_common.cuh"

namespace tgr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

// ---- optional per-entry CUDA-event timing (bench / profiling only; off by default) ---------------------
struct TimingRec { const char* name; cudaEvent_t a, b; };
static std::mutex g_tm_mu;
static bool g_tm_on = false;
static std::vector<TimingRec> g_tm_recs;

TimedScope::TimedScope(const char* name, void* stream) : name_(name), stream_(stream), a_(nullptr) {
  if (!g_tm_on) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, (cudaStream_t)stream);
  a_ = e;
}
TimedScope::~TimedScope() {
  if (a_ == nullptr) return;
  cudaEvent_t b;
  if (cudaEventCreate(&b) != cudaSuccess) return;
  cudaEventRecord(b, (cudaStream_t)stream_);
  std::lock_guard<std::mutex> lk(g_tm_mu);
  g_tm_recs.push_back({name_, (cudaEvent_t)a_, b});
}

}  // namespace tgr

extern "C" int tgr_timing_enable(int on) {
  std::lock_guard<std::mutex> lk(tgr::g_tm_mu);
  for (auto& r : tgr::g_tm_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  tgr::g_tm_recs.clear();
  tgr::g_tm_on = on != 0;
  return 0;
}

// Aggregates the records since tgr_timing_enable(1) by entry name: names = '\n'-joined, ms / counts per name.
extern "C" int tgr_timing_collect(char* names, size_t names_bytes, float* ms, int32_t* counts, int max_entries) {
  std::lock_guard<std::mutex> lk(tgr::g_tm_mu);
  std::map<std::string, std::pair<double, int>> agg;
  std::vector<std::string> order;
  for (auto& r : tgr::g_tm_recs) {
    cudaEventSynchronize(r.b);
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    auto it = agg.find(r.name);
    if (it == agg.end()) { agg[r.name] = {t, 1}; order.push_back(r.name); }
    else { it->second.first += t; it->second.second += 1; }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  tgr::g_tm_recs.clear();
  size_t off = 0;
  int n = 0;
  for (auto& k : order) {
    if (n >= max_entries || off + k.size() + 2 > names_bytes) break;
    memcpy(names + off, k.data(), k.size());
    off += k.size();
    names[off++] = '\n';
    ms[n] = (float)agg[k].first;
    counts[n] = agg[k].second;
    ++n;
  }
  if (names_bytes) names[off < names_bytes ? off : names_bytes - 1] = 0;
  return n;
}

extern "C" int64_t tgr_launch_count(void) { return (int64_t)tgr::g_launches.load(std::memory_order_relaxed); }
extern "C" int tgr_abi_version(void) { return TGR_ABI_VERSION; }
extern "C" const char* tgr_last_error(void) { return tgr::g_err; }
