// Shared pieces of the tcgen05 point-wise linear kernels (linear_tc.cu: cp.async loaders, any X layout;
// linear_tma.cu: TMA loaders, row-major X).
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace samble {

constexpr int kLinThreads = 288;
constexpr int kLinStages = 3;     // smem ring depth; loads run kLinAhead stages ahead of the split/arrive step
constexpr int kLinAhead = 2;
// The tensor core adds into its fp32 accumulator with truncation, a bias that grows with the length of the
// accumulation chain (measured: 9e-5 abs at K=1024 vs 2e-5 at K=128 on O(10) outputs).  Chains are therefore cut
// every kLinChain K-blocks: each chunk gets its own TMEM accumulator and the epilogue adds the chunks in fp32.
constexpr int kLinChain = 8;

struct LinArgs {
  const float* X; long long ldx;      // row-major: X[m*ldx + k];  channel-major (x_cm): X[(b*K + k)*npc + n], m = b*npc + n
  const float* W; long long ldw;      // Nout x K
  const float* Wlo;                   // W - tf32_trunc(W), same shape/stride (weights are constants: split once on the host side)
  const float* scale;                 // [Nout] or null (=1)
  const float* shift;                 // [Nout] (+ b*shift_ldb) or null (=0)
  const float* residual; long long ldr;   // same indexing as out, or null
  float* out; long long ldo;          // row-major: out[m*ldo + c];  channel-major (out_cm): out[(b*Nout + c)*npc + n]
  int M, K, Nout, npc;                // npc = points per cloud (needed by either channel-major side and by shift_ldb)
  int lrelu, x_cm, out_cm, res_first; // res_first: y = (acc + res)*scale + shift  (else residual is added last)
  int res_cm;                         // residual layout (row-major with ldr, or channel-major), independent of out's
  long long shift_ldb;                // per-cloud shift stride (0 = shared)
};


// Epilogue of one 128 x NT output tile, thread = output row (TMEM lane):
//   y = acc*scale[c] + shift[c] -> LeakyReLU -> + residual      (or (acc + residual)*scale + shift when res_first)
// summed over the tile's accumulation chunks, stored row-major or channel-major.
template <int NT>
__device__ __forceinline__ void linear_epilogue_tile(const LinArgs& a, uint32_t tmem, int set, int nacc, int m0, int n0, int warp,
                                                     int lane) {
  const int m = m0 + warp * 32 + lane;
  const uint32_t lane_base = ((uint32_t)(warp * 32) << 16) + set * nacc * NT;
  const bool live = m < a.M;
  long long ob = 0, on = 0;
  if (a.npc > 0) {
    ob = (live ? m : 0) / a.npc;
    on = (live ? m : 0) % a.npc;
  }
  const float* shift = a.shift ? a.shift + ob * a.shift_ldb : nullptr;
#pragma unroll 1
  for (int c0 = 0; c0 < NT; c0 += 32) {
    float v[32];
    tc::tmem_ld32(tmem + lane_base + c0, v);      // warp-collective: every lane takes part, stores are predicated
    for (int ac = 1; ac < nacc; ++ac) {           // add the accumulation chunks in fp32
      float w[32];
      tc::tmem_ld32(tmem + lane_base + ac * NT + c0, w);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += w[i];
    }
    if (!live || n0 + c0 >= a.Nout) continue;
    const bool full32 = n0 + c0 + 32 <= a.Nout;
    float r[32];
    if (a.residual) {
      if (a.res_cm) {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = (n0 + c0 + i < a.Nout) ? __ldg(a.residual + (ob * a.Nout + n0 + c0 + i) * a.npc + on) : 0.f;
      } else {
        const float* rrow = a.residual + (long long)m * a.ldr + n0 + c0;
        if (full32 && a.ldr % 4 == 0 && reinterpret_cast<uintptr_t>(a.residual) % 16 == 0) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(rrow + i));
            r[i] = t.x, r[i + 1] = t.y, r[i + 2] = t.z, r[i + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = (n0 + c0 + i < a.Nout) ? __ldg(rrow + i) : 0.f;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int c = min(n0 + c0 + i, a.Nout - 1);
      float y = v[i];
      if (a.residual && a.res_first) y += r[i];
      if (a.scale) y *= __ldg(a.scale + c);
      if (shift) y += __ldg(shift + c);
      if (a.lrelu) y = y > 0.f ? y : 0.2f * y;
      if (a.residual && !a.res_first) y += r[i];
      v[i] = y;
    }
    if (a.out_cm) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (n0 + c0 + i < a.Nout) a.out[(ob * a.Nout + n0 + c0 + i) * a.npc + on] = v[i];
    } else {
      float* orow = a.out + (long long)m * a.ldo + n0 + c0;
      if (full32 && a.ldo % 4 == 0 && reinterpret_cast<uintptr_t>(a.out) % 16 == 0) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(orow + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (n0 + c0 + i < a.Nout) orow[i] = v[i];
      }
    }
  }
}

}  // namespace samble
