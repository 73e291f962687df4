// Shared device/host helpers for the sm_100a kernels behind include/samble_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/samble_b200.h"

namespace samble {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// ---- error plumbing (host) ----
void set_error(const char* fmt, ...);
void count_launch();
int check_launch(const char* what);   // cudaGetLastError -> SAMBLE_E_CUDA + message
void prof_pre(cudaStream_t st);       // no-ops unless samble_profile_enable(1)
void prof_post(const char* what);

#define SAMBLE_REQUIRE(cond, ...)                    \
  do {                                               \
    if (!(cond)) {                                   \
      ::samble::set_error(__VA_ARGS__);              \
      return SAMBLE_E_INVALID;                       \
    }                                                \
  } while (0)

#define SAMBLE_PRE(st) ::samble::prof_pre(st)

#define SAMBLE_LAUNCHED(what)                        \
  do {                                               \
    ::samble::prof_post(what);                       \
    ::samble::count_launch();                        \
    int _e = ::samble::check_launch(what);           \
    if (_e) return _e;                               \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// bump allocator over the caller's workspace
struct Workspace {
  char* base;
  size_t size, used;
  Workspace(void* p, size_t n) : base((char*)p), size(n), used(0) {}
  template <class T>
  T* take(size_t n) {
    size_t off = align_up(used, 256);
    used = off + n * sizeof(T);
    return (T*)(base + off);
  }
  bool ok() const { return used <= size; }
};

// ---- index access, int32 or int64 ----
template <class I>
__device__ __forceinline__ int ld_idx(const I* p, long long i) { return (int)p[i]; }

// ---- warp primitives ----
__device__ __forceinline__ float warp_sum(float v) {   // fixed butterfly order => deterministic
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

// ---- cp.async (LDGSTS) with zero-fill predicate ----
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int n = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(n));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---- running "k smallest" set held one entry per lane (k <= 32) ----
// Entry order is the total order (distance bits, index): distances are clamped to >= +0 so their IEEE
// bit patterns order like unsigned ints; ties go to the lower index.
// CONTRACT: candidates are offered in ASCENDING index order (lane order inside a call, calls in index
// order).  A candidate that ties the current worst distance therefore has the higher index and loses,
// so the hot path is a single 32-bit compare; only eviction among equal worst distances looks at indices.
struct LaneTopK {
  unsigned d;   // distance bits of this lane's entry (0xffffffff = empty, 0 on inactive lanes)
  int i;        // candidate index of this lane's entry
  unsigned td;  // warp-uniform: distance bits of the current worst (largest) entry
  int tl;       // warp-uniform: lane holding the entry to evict next
  int lane;
  bool active;  // lane < k

  __device__ __forceinline__ void init(int lane_, int k) {
    lane = lane_;
    active = lane < k;
    d = active ? 0xffffffffu : 0u;
    i = 0x7ffffff0 - lane;   // distinct per lane
    td = 0xffffffffu;
    tl = 0;                  // lane 0 holds the largest placeholder index
  }
  __device__ __forceinline__ void load(int lane_, int k, unsigned d_, int i_) {
    lane = lane_;
    active = lane < k;
    d = d_;
    i = i_;
    refresh();
  }
  // seed the set directly (no serial inserts): lanes with `take` adopt their own candidate.
  // Only valid while every taken slot is still empty; candidates must be in ascending index order by lane.
  __device__ __forceinline__ void fill(unsigned cd, int ci, bool take) {
    if (take && active) { d = cd; i = ci; }
    refresh();
  }
  __device__ __forceinline__ void refresh() {
    td = __reduce_max_sync(kFull, d);
    const unsigned m = __ballot_sync(kFull, active && d == td);
    tl = __ffs(m) - 1;
    if (m & (m - 1)) {       // several entries share the worst distance: evict the highest index
      const int mi = __reduce_max_sync(kFull, (active && d == td) ? i : (int)0x80000000);
      tl = __ffs(__ballot_sync(kFull, active && d == td && i == mi)) - 1;
    }
  }
  // All 32 lanes call this with their own candidate (cd, ci); `want` says whether it is real.
  __device__ __forceinline__ void offer(unsigned cd, int ci, bool want) {
    unsigned pass = __ballot_sync(kFull, want && cd < td);
    while (pass) {
      const int src = __ffs(pass) - 1;
      pass &= pass - 1;
      const unsigned sd = __shfl_sync(kFull, cd, src);
      if (sd < td) {                        // warp-uniform: the worst may have improved meanwhile
        const int si = __shfl_sync(kFull, ci, src);
        const bool hit = lane == tl;
        d = hit ? sd : d;
        i = hit ? si : i;
        refresh();
      }
    }
  }
  // rank of this lane's entry among the k entries (0 = nearest); entries are distinct.
  __device__ __forceinline__ int rank() const {
    int r = 0;
#pragma unroll
    for (int l = 0; l < kWarp; ++l) {
      const unsigned od = __shfl_sync(kFull, active ? d : 0xffffffffu, l);
      const int oi = __shfl_sync(kFull, active ? i : 0x7fffffff, l);
      r += (od < d || (od == d && oi < i)) ? 1 : 0;
    }
    return r;
  }
};

__device__ __forceinline__ unsigned dist_bits(float d2) {
  // torch.cdist clamps the GEMM-form squared distance at 0 before sqrt
  // (reference utils/ops.py:35 -> at::_euclidean_dist clamp_min_(0)); also maps -0/NaN to +0.
  return __float_as_uint(d2 > 0.f ? d2 : 0.f);
}

}  // namespace samble
