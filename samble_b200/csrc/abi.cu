// ABI plumbing: version, thread-local error text, launch counter.
#include <stdarg.h>

#include "common.cuh"

namespace samble {

static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch() { ++g_launches; }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return SAMBLE_OK;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return SAMBLE_E_CUDA;
}

}  // namespace samble

extern "C" int samble_abi_version(void) { return 1; }
extern "C" const char* samble_last_error(void) { return samble::g_err; }
extern "C" long long samble_launch_count(void) { return samble::g_launches; }
extern "C" void samble_reset_launch_count(void) { samble::g_launches = 0; }
