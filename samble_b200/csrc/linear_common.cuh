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
  // pooling mode (linear_tma.cu only): instead of storing y, write per-32-row-group column maxima / sums
  // [M/32][Nout] (a group never straddles two clouds: npc % 32 == 0); samble_linear_pool reduces them per cloud
  float* pool_max;
  float* pool_sum;
  // per-cloud weights (linear_tma.cu only): cloud b = m / npc multiplies by W[b] (Nout x K slices, pitch Nout*ldw)
  int w_batched;
  // row-softmax epilogue: y = exp(acc / logit_div - row_max[m]) / row_sum[m]   (row statistics known beforehand)
  const float* row_max;
  const float* row_sum;
  float logit_div;
  int chain;                          // K-blocks per accumulation chain (linear_tma.cu; 0 = kLinChain)
  // row-statistics mode (linear_tma.cu): nothing is stored but, per row and 128-column tile, the maximum of
  // acc / logit_div and the sum of exp(. - max):  stat_out[m * ntiles + n_tile] = (max, sum)
  float2* stat_out;
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
    // per-column constants: warp-uniform addresses, fetched 4 at a time when the block is whole and aligned
    float sc[32], sh[32];
    const bool cvec = full32 && (n0 + c0) % 4 == 0 && (!a.scale || reinterpret_cast<uintptr_t>(a.scale) % 16 == 0) &&
                      (!shift || (reinterpret_cast<uintptr_t>(shift) % 16 == 0));
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f), h4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cvec) {
        if (a.scale) s4 = __ldg(reinterpret_cast<const float4*>(a.scale + n0 + c0 + i));
        if (shift) h4 = __ldg(reinterpret_cast<const float4*>(shift + n0 + c0 + i));
      } else {
        const int c = n0 + c0 + i, last = a.Nout - 1;
        if (a.scale) s4 = make_float4(__ldg(a.scale + min(c, last)), __ldg(a.scale + min(c + 1, last)), __ldg(a.scale + min(c + 2, last)), __ldg(a.scale + min(c + 3, last)));
        if (shift) h4 = make_float4(__ldg(shift + min(c, last)), __ldg(shift + min(c + 1, last)), __ldg(shift + min(c + 2, last)), __ldg(shift + min(c + 3, last)));
      }
      sc[i] = s4.x, sc[i + 1] = s4.y, sc[i + 2] = s4.z, sc[i + 3] = s4.w;
      sh[i] = h4.x, sh[i + 1] = h4.y, sh[i + 2] = h4.z, sh[i + 3] = h4.w;
    }
    if (a.row_max) {
      // 4 instructions per element (FFMA, FMUL, MUFU.EX2, FMUL): the row statistics themselves were accumulated with
      // the same ex2-based exponential (ds_rowstats_tc.cu), relative error ~2^-21
      const float mu = __ldg(a.row_max + (live ? m : 0)), inv_s = 1.f / __ldg(a.row_sum + (live ? m : 0)), inv_div = 1.f / a.logit_div;
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __expf(fmaf(v[i], inv_div, -mu)) * inv_s;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float y = v[i];
      if (a.residual && a.res_first) y += r[i];
      if (a.scale) y *= sc[i];
      if (shift) y += sh[i];
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

// Row-major outputs: the same epilogue, staged through a per-warp shared-memory slab so that global traffic is
// coalesced and the per-column constants sit in registers.  Phase 1 (thread = row) dumps a 32x32 block of raw
// accumulators into the slab; phase 2 re-reads it with 8 lanes per row (one float4 = 4 fixed columns per lane), so a
// warp instruction moves four full 128-byte row segments, scale/shift are loaded once per block, and a row-major
// residual is read coalesced.  The slab is private to the warp (warp w owns rows 32w..32w+31): __syncwarp only.
constexpr int kLinSlabLd = 36;                                   // floats per slab row (pad 4: conflict-free 128-bit access)
constexpr int kLinSlabBytes = 4 * 32 * kLinSlabLd * 4;           // four epilogue warps

template <int NT>
__device__ __forceinline__ void linear_epilogue_tile_staged(const LinArgs& a, uint32_t tmem, int set, int nacc, int m0, int n0,
                                                            int warp, int lane, float* slab_all) {
  float* slab = slab_all + warp * 32 * kLinSlabLd;
  const uint32_t lane_base = ((uint32_t)(warp * 32) << 16) + set * nacc * NT;
  const int rq = lane >> 3, cq = (lane & 7) * 4;                 // phase 2: row within a group of 4, first of 4 columns
  const bool vec_ok = a.ldo % 4 == 0 && reinterpret_cast<uintptr_t>(a.out) % 16 == 0 &&
                      (!a.residual || (a.ldr % 4 == 0 && reinterpret_cast<uintptr_t>(a.residual) % 16 == 0));
#pragma unroll 1
  for (int c0 = 0; c0 < NT; c0 += 32) {
    float v[32];
    tc::tmem_ld32(tmem + lane_base + c0, v);
    for (int ac = 1; ac < nacc; ++ac) {
      float w[32];
      tc::tmem_ld32(tmem + lane_base + ac * NT + c0, w);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += w[i];
    }
    if (n0 + c0 >= a.Nout) continue;                             // warp-uniform
#pragma unroll
    for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(slab + lane * kLinSlabLd + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    __syncwarp();
    const int c = n0 + c0 + cq;                                  // this lane's 4 columns: c .. c+3
    float sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cc = min(c + j, a.Nout - 1);
      if (a.scale) sc[j] = __ldg(a.scale + cc);
      if (a.shift && a.shift_ldb == 0) sh[j] = __ldg(a.shift + cc);
    }
    const bool full4 = c + 4 <= a.Nout;
    // residual rows of the 8 row groups: all eight loads are issued before any is used (one latency, not eight)
    float4 res4[8];
    if (a.residual) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int m = m0 + warp * 32 + it * 4 + rq;
        res4[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < a.M && c < a.Nout) {
          const float* rp = a.residual + (long long)m * a.ldr + c;
          if (full4 && vec_ok) {
            res4[it] = __ldg(reinterpret_cast<const float4*>(rp));
          } else {
            res4[it].x = __ldg(rp);
            if (c + 1 < a.Nout) res4[it].y = __ldg(rp + 1);
            if (c + 2 < a.Nout) res4[it].z = __ldg(rp + 2);
            if (c + 3 < a.Nout) res4[it].w = __ldg(rp + 3);
          }
        }
      }
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int r = it * 4 + rq;
      const int m = m0 + warp * 32 + r;
      if (m >= a.M || c >= a.Nout) continue;
      const float4 t = *reinterpret_cast<const float4*>(slab + r * kLinSlabLd + cq);
      float y[4] = {t.x, t.y, t.z, t.w};
      if (a.shift && a.shift_ldb != 0) {
        const float* shp = a.shift + (long long)(m / a.npc) * a.shift_ldb;
#pragma unroll
        for (int j = 0; j < 4; ++j) sh[j] = __ldg(shp + min(c + j, a.Nout - 1));
      }
      const float res[4] = {res4[it].x, res4[it].y, res4[it].z, res4[it].w};
      if (a.row_max) {                                           // softmax row with known max / sum (downsample.py:242-250)
        const float mu = __ldg(a.row_max + m), inv_s = 1.f / __ldg(a.row_sum + m), inv_div = 1.f / a.logit_div;
#pragma unroll
        for (int j = 0; j < 4; ++j) y[j] = __expf(fmaf(y[j], inv_div, -mu)) * inv_s;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float z = y[j];
        if (a.residual && a.res_first) z += res[j];
        if (a.scale) z *= sc[j];
        if (a.shift) z += sh[j];
        if (a.lrelu) z = z > 0.f ? z : 0.2f * z;
        if (a.residual && !a.res_first) z += res[j];
        y[j] = z;
      }
      float* op = a.out + (long long)m * a.ldo + c;
      if (full4 && vec_ok) {
        *reinterpret_cast<float4*>(op) = make_float4(y[0], y[1], y[2], y[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c + j < a.Nout) op[j] = y[j];
      }
    }
    __syncwarp();                                                // slab is rewritten by the next block
  }
}

// Pooling epilogue: y as above (no residual), reduced over the 32 rows of this warp per column with a halving
// butterfly (31 shuffles per reduction instead of 160: at each step a lane keeps the half of its columns selected by
// one of its lane bits and hands the other half to its partner), after which lane l holds column l of the block.
template <int NT>
__device__ __forceinline__ void linear_epilogue_tile_pool(const LinArgs& a, uint32_t tmem, int set, int nacc, int m0, int n0,
                                                          int warp, int lane) {
  const int m = m0 + warp * 32 + lane;
  const uint32_t lane_base = ((uint32_t)(warp * 32) << 16) + set * nacc * NT;
  const bool live = m < a.M;
  const int grp = (m0 >> 5) + warp;                                    // 32-row group of this warp
  if (grp * 32 >= a.M) {                                               // whole warp past the end: only take part in the loads
    float v[32];
    for (int c0 = 0; c0 < NT; c0 += 32)
      for (int ac = 0; ac < nacc; ++ac) tc::tmem_ld32(tmem + lane_base + ac * NT + c0, v);
    return;
  }
  const float* shift = a.shift ? a.shift + (long long)((grp * 32) / a.npc) * a.shift_ldb : nullptr;
#pragma unroll 1
  for (int c0 = 0; c0 < NT; c0 += 32) {
    float v[32];
    tc::tmem_ld32(tmem + lane_base + c0, v);
    for (int ac = 1; ac < nacc; ++ac) {
      float w[32];
      tc::tmem_ld32(tmem + lane_base + ac * NT + c0, w);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += w[i];
    }
    if (n0 + c0 >= a.Nout) continue;
    float mx[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int c = min(n0 + c0 + i, a.Nout - 1);
      float y = v[i];
      if (a.scale) y *= __ldg(a.scale + c);
      if (shift) y += __ldg(shift + c);
      if (a.lrelu) y = y > 0.f ? y : 0.2f * y;
      v[i] = live ? y : 0.f;
      mx[i] = live ? y : -INFINITY;
    }
#pragma unroll
    for (int w = 16; w >= 1; w >>= 1) {
      const bool upper = (lane & w) != 0;
#pragma unroll
      for (int i = 0; i < w; ++i) {
        const float keep_s = upper ? v[i + w] : v[i], send_s = upper ? v[i] : v[i + w];
        const float keep_m = upper ? mx[i + w] : mx[i], send_m = upper ? mx[i] : mx[i + w];
        v[i] = keep_s + __shfl_xor_sync(kFull, send_s, w);
        mx[i] = fmaxf(keep_m, __shfl_xor_sync(kFull, send_m, w));
      }
    }
    const int c = n0 + c0 + lane;
    if (c < a.Nout) {
      a.pool_max[(long long)grp * a.Nout + c] = mx[0];
      if (a.pool_sum) a.pool_sum[(long long)grp * a.Nout + c] = v[0];
    }
  }
}

// Row-statistics epilogue (DownSampleToken pass 1, models/downsample.py:139-153): the N x N logits are never stored;
// each row keeps an online (max, sum of exp) over the tile's columns.
template <int NT>
__device__ __forceinline__ void linear_epilogue_tile_rowstat(const LinArgs& a, uint32_t tmem, int set, int nacc, int m0, int n0,
                                                             int warp, int lane, int ntiles) {
  const int m = m0 + warp * 32 + lane;
  const uint32_t lane_base = ((uint32_t)(warp * 32) << 16) + set * nacc * NT;
  const float inv_div = 1.f / a.logit_div;
  float mx = -INFINITY, sum = 0.f;
#pragma unroll 1
  for (int c0 = 0; c0 < NT; c0 += 32) {
    float v[32];
    tc::tmem_ld32(tmem + lane_base + c0, v);
    for (int ac = 1; ac < nacc; ++ac) {
      float w[32];
      tc::tmem_ld32(tmem + lane_base + ac * NT + c0, w);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += w[i];
    }
    if (n0 + c0 >= a.Nout) continue;
    float cm = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      v[i] = (n0 + c0 + i < a.Nout) ? v[i] * inv_div : -INFINITY;
      cm = fmaxf(cm, v[i]);
    }
    const float mn = fmaxf(mx, cm);
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) part += __expf(v[i] - mn);
    sum = sum * __expf(mx - mn) + part;
    mx = mn;
  }
  if (m < a.M) a.stat_out[(long long)m * ntiles + n0 / NT] = make_float2(mx, sum);
}

}  // namespace samble
