// Error plumbing and version of the C ABI.
#include <stdarg.h>
#include <string.h>

#include "tgr_common.cuh"

namespace tgr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

}  // namespace tgr

extern "C" int tgr_abi_version(void) { return TGR_ABI_VERSION; }
extern "C" const char* tgr_last_error(void) { return tgr::g_err; }
