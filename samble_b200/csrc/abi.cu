// ABI plumbing: version, thread-local error text, launch counter, optional per-kernel event timing.
#include <stdarg.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace samble {

static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;

// ---- per-kernel timing (bench.py / DESIGN.md evidence; off by default) ----
struct ProfRec {
  const char* name;
  cudaEvent_t e0, e1;
};
static thread_local bool g_prof = false;
static thread_local cudaStream_t g_prof_stream = nullptr;
static thread_local cudaEvent_t g_pending = nullptr;
static thread_local std::vector<ProfRec>* g_recs = nullptr;

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

void prof_pre(cudaStream_t st) {
  if (!g_prof) return;
  g_prof_stream = st;
  cudaEventCreate(&g_pending);
  cudaEventRecord(g_pending, st);
}

void prof_post(const char* what) {
  if (!g_prof || !g_pending) return;
  ProfRec r{what, g_pending, nullptr};
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e1, g_prof_stream);
  if (!g_recs) g_recs = new std::vector<ProfRec>();
  g_recs->push_back(r);
  g_pending = nullptr;
}

}  // namespace samble

using namespace samble;

extern "C" int samble_abi_version(void) { return 1; }
extern "C" const char* samble_last_error(void) { return g_err; }
extern "C" long long samble_launch_count(void) { return g_launches; }
extern "C" void samble_reset_launch_count(void) { g_launches = 0; }

extern "C" void samble_profile_enable(int on) { g_prof = on != 0; }

extern "C" int samble_profile_report(char* buf, size_t cap) {
  // "name,launches,total_ms\n" per kernel; synchronises the recorded events; clears the records.
  if (!buf || cap == 0) return SAMBLE_E_INVALID;
  buf[0] = 0;
  if (!g_recs) return SAMBLE_OK;
  std::map<std::string, std::pair<long long, double>> agg;
  for (auto& r : *g_recs) {
    float ms = 0.f;
    cudaEventSynchronize(r.e1);
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    auto& a = agg[r.name];
    a.first += 1;
    a.second += ms;
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_recs->clear();
  size_t off = 0;
  for (auto& kv : agg) {
    int n = snprintf(buf + off, cap - off, "%s,%lld,%.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    if (n < 0 || (size_t)n >= cap - off) break;
    off += n;
  }
  return SAMBLE_OK;
}
