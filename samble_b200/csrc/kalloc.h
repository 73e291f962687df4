// Sequential fp32 logic of the DownSample bin stage, shared verbatim between the CUDA kernels
// (sampler.cu) and a host build (hostcheck.cpp) so it can be tested on a box without a GPU.
//
// Mirrors reference utils/ops.py:385-432 (calculate_num_points_to_choose) op for op: the same
// fp32 operations in the same order, including ATen's summation order for a short contiguous
// row (probed on torch 2.11 CPU: n<=4 and n==8 left-to-right; 5<=n<=7: x0, then x4..x[n-1],
// then x1, x2, x3), the .int() truncation and the first-maximum remainder rule.
#pragma once
#if defined(__CUDACC__)
#define SAMBLE_HD __host__ __device__ __forceinline__
#else
#define SAMBLE_HD inline
#endif

namespace samble {

constexpr int kMaxBins = 8;

// no FMA contraction may sneak in: every product and sum below is a separately rounded fp32 op.
#if defined(__CUDA_ARCH__)
#define SAMBLE_MUL(a, b) __fmul_rn((a), (b))
#define SAMBLE_ADD(a, b) __fadd_rn((a), (b))
#define SAMBLE_SUB(a, b) __fsub_rn((a), (b))
#define SAMBLE_DIV(a, b) __fdiv_rn((a), (b))
#else
static inline float samble_mul(volatile float a, volatile float b) { volatile float r = a * b; return r; }
static inline float samble_add(volatile float a, volatile float b) { volatile float r = a + b; return r; }
static inline float samble_sub(volatile float a, volatile float b) { volatile float r = a - b; return r; }
static inline float samble_div(volatile float a, volatile float b) { volatile float r = a / b; return r; }
#define SAMBLE_MUL(a, b) samble_mul((a), (b))
#define SAMBLE_ADD(a, b) samble_add((a), (b))
#define SAMBLE_SUB(a, b) samble_sub((a), (b))
#define SAMBLE_DIV(a, b) samble_div((a), (b))
#endif

SAMBLE_HD float aten_row_sum(const float* x, int n) {
  float s = x[0];
  if (n <= 4 || n == 8) {
    for (int j = 1; j < n; ++j) s = SAMBLE_ADD(s, x[j]);
    return s;
  }
  for (int j = 4; j < n; ++j) s = SAMBLE_ADD(s, x[j]);
  for (int j = 1; j < 4; ++j) s = SAMBLE_ADD(s, x[j]);
  return s;
}

// w: relu'd bin weights, cnt: points per bin, total: M.  k_out: points to take per bin.
SAMBLE_HD void num_points_to_choose(const float* w, const long long* cnt, int nb, int total, int* k_out) {
  float p[kMaxBins], chosen[kMaxBins], cap[kMaxBins];
  for (int j = 0; j < nb; ++j) {
    cap[j] = (float)cnt[j];
    p[j] = SAMBLE_ADD(SAMBLE_MUL(w[j], cap[j]), 1e-10f);
    chosen[j] = 0.f;
  }
  for (int it = 0; it < nb; ++it) {
    const float ps = aten_row_sum(p, nb);
    for (int j = 0; j < nb; ++j) p[j] = SAMBLE_DIV(p[j], ps);
    const float left = SAMBLE_SUB((float)total, aten_row_sum(chosen, nb));
    // the reference leaves the loop only when EVERY cloud of the batch has nothing left
    // (:409); for one cloud with left == 0 a further round adds p*0 and changes nothing,
    // so stopping per cloud gives the same integers.
    if (left == 0.f) break;
    for (int j = 0; j < nb; ++j) {
      chosen[j] = SAMBLE_ADD(chosen[j], SAMBLE_MUL(p[j], left));
      const bool full = chosen[j] >= cap[j];
      if (full) chosen[j] = cap[j];
      p[j] = SAMBLE_MUL(p[j], full ? 0.f : 1.f);
    }
  }
  long long sum = 0, best = 0;
  int arg = 0;
  for (int j = 0; j < nb; ++j) {
    k_out[j] = (int)chosen[j];
    sum += k_out[j];
  }
  for (int j = 0; j < nb; ++j) {
    long long room = cnt[j] - k_out[j];
    if (j == 0 || room > best) best = room, arg = j;
  }
  k_out[arg] += (int)(total - sum);
}

}  // namespace samble
