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
  int dbg;                            // measurement switches (samble_set_linear_debug): 8 = no residual loads, 16 = no stores
  int pool_rows;                      // rows per pooling group: 32 (row-per-thread epilogue) or 128 (swapped orientation); 0 = 32
};


// Epilogue of one 128 x NT output tile, thread = output row (TMEM lane):
//   y = acc*scale[c] + shift[c] -> LeakyReLU -> + residual      (or (acc + residual)*scale + shift when res_first)
// summed over the tile's accumulation chunks, stored row-major or channel-major.
template <int NT>
__device__ __forceinline__ void linear_epilogue_tile(const LinArgs& a, uint32_t tmem, int set, int nacc, int m0, int n0, int warp,
                                                     int lane, int c_begin = 0, int c_end = NT) {
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
  for (int c0 = c_begin; c0 < c_end; c0 += 32) {   // (a column range: csrc/mlp2.cu shares a tile between two warps per lane quadrant)
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
    if (a.dbg & 8) {
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = 0.f;
    } else if (a.residual) {
      if (a.res_cm) {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = (n0 + c0 + i < a.Nout) ? __ldg(a.residual + (ob * a.Nout + n0 + c0 + i) * a.npc + on) : 0.f;
      } else {
        const float* rrow = a.residual + (long long)m * a.ldr + n0 + c0;
        if (full32 && a.ldr % 8 == 0 && (n0 + c0) % 8 == 0 && reinterpret_cast<uintptr_t>(a.residual) % 32 == 0) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) tc::ldg256(rrow + i, r + i);          // one whole sector per lane and instruction
        } else if (full32 && a.ldr % 4 == 0 && reinterpret_cast<uintptr_t>(a.residual) % 16 == 0) {
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
    if (a.dbg & 16) {
      if (v[0] == 1234.5f) a.out[0] = v[1];                    // (keeps the math alive)
    } else if (a.out_cm) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (n0 + c0 + i < a.Nout) a.out[(ob * a.Nout + n0 + c0 + i) * a.npc + on] = v[i];
    } else {
      float* orow = a.out + (long long)m * a.ldo + n0 + c0;
      if (full32 && a.ldo % 8 == 0 && (n0 + c0) % 8 == 0 && reinterpret_cast<uintptr_t>(a.out) % 32 == 0) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) tc::stg256(orow + i, v + i);            // 32 whole sectors per warp instruction
      } else if (full32 && a.ldo % 4 == 0 && reinterpret_cast<uintptr_t>(a.out) % 16 == 0) {
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

// Row-major outputs through TMA stores (round 2).  Phase ablation of linear_tma_kernel (tools/probe_linear.py,
// profiles/r2_linear_tma.md) showed the thread-per-row global stores ADDING their whole duration to the kernel instead
// of overlapping with the main loop: 32 scattered sectors per instruction queue up in the LSU, which the operand
// splitters share.  Here a warp parks its 32 x 32 block in shared memory (128-byte-swizzled rows: conflict-free for
// thread = row) and one lane hands it to the TMA engine: whole 128-byte lines, no LSU involvement, rows / columns past
// the end clipped by the tensor map.  Two 4 KB buffers per warp: the store of block c drains while block c+1 is computed.
constexpr int kLinStoreBytes = 4 * 2 * 4096;                     // four epilogue warps, double buffered

template <int NT>
__device__ __forceinline__ void linear_epilogue_tile_tma(const LinArgs& a, const CUtensorMap* map_out, uint32_t tmem, int set,
                                                         int nacc, int m0, int n0, int warp, int lane, uint8_t* stage_all,
                                                         int& parity) {
  const int m = m0 + warp * 32 + lane;
  const uint32_t lane_base = ((uint32_t)(warp * 32) << 16) + set * nacc * NT;
  const bool live = m < a.M;
  const int mm = live ? m : 0;
  const float* shift = a.shift ? a.shift + (a.shift_ldb ? (long long)(mm / a.npc) * a.shift_ldb : 0) : nullptr;
  const bool res_vec = a.residual && a.ldr % 8 == 0 && reinterpret_cast<uintptr_t>(a.residual) % 32 == 0;
#pragma unroll 1
  for (int c0 = 0; c0 < NT; c0 += 32) {
    if (n0 + c0 >= a.Nout) break;                                // warp-uniform (tail tile of the n range)
    const bool full32 = n0 + c0 + 32 <= a.Nout;
    float r[32];
    if (a.residual) {                                            // issued first: in flight during the TMEM load
      const float* rrow = a.residual + (long long)mm * a.ldr + n0 + c0;
      if (full32 && res_vec && (n0 + c0) % 8 == 0) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) tc::ldg256(rrow + i, r + i);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = (n0 + c0 + i < a.Nout) ? __ldg(rrow + i) : 0.f;
      }
    }
    float v[32];
    tc::tmem_ld32(tmem + lane_base + c0, v);
    for (int ac = 1; ac < nacc; ++ac) {
      float w[32];
      tc::tmem_ld32(tmem + lane_base + ac * NT + c0, w);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += w[i];
    }
    if (a.row_max) {                                             // softmax row with known max / sum (downsample.py:242-250)
      const float mu = __ldg(a.row_max + mm), inv_s = 1.f / __ldg(a.row_sum + mm), inv_div = 1.f / a.logit_div;
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __expf(fmaf(v[i], inv_div, -mu)) * inv_s;
    }
    const bool cvec = full32 && (n0 + c0) % 4 == 0 && (!a.scale || reinterpret_cast<uintptr_t>(a.scale) % 16 == 0) &&
                      (!shift || (reinterpret_cast<uintptr_t>(shift) % 16 == 0));
    uint8_t* buf = stage_all + warp * 8192 + (parity & 1) * 4096;
    if (lane == 0) tc::bulk_wait_read<1>();                      // the store issued from this buffer two blocks ago has read it
    __syncwarp();
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
      const float sc[4] = {s4.x, s4.y, s4.z, s4.w}, sh[4] = {h4.x, h4.y, h4.z, h4.w};
      float y[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float z = v[i + j];
        if (a.residual && a.res_first) z += r[i + j];
        if (a.scale) z *= sc[j];
        if (shift) z += sh[j];
        if (a.lrelu) z = z > 0.f ? z : 0.2f * z;
        if (a.residual && !a.res_first) z += r[i + j];
        y[j] = z;
      }
      *reinterpret_cast<float4*>(buf + tc::sw128_offset(lane, i >> 2)) = make_float4(y[0], y[1], y[2], y[3]);
    }
    tc::fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      tc::tma_store_3d(map_out, buf, n0 + c0, m0 + warp * 32, 0);
      tc::bulk_commit();
    }
    ++parity;
  }
}

// ---- swapped orientation (linear_tma.cu, EPI 5): the MMAs are issued with the WEIGHT tile as the A operand and the X tile as B,
// so the accumulator holds the transposed tile: TMEM lane = output channel n0 + lane, column = point m0 + col; the products and
// their accumulation order per output element are unchanged (bit-identical pre-activations).
// (A direct-store epilogue in this orientation -- one whole 128-byte line per store instruction instead of 32 sectors of 32
// different lines -- was built and measured SLOWER than the row-per-thread stores: 54.8 vs 43.6 us at 32768 x 128 -> 384,
// 114 vs 88 us at -> 1024.  Without any store the kernel takes 27.9 / 49.6 us, so the 50 / 134 MB of output drain at
// 3.1 - 3.4 TB/s either way: the epilogue of these layers is bound by the HBM write stream, not by the LSU.  Removed.)
// swapped pooling: a thread owns one channel and 128 points of ONE cloud (npc % 128 == 0): the reduction over the points runs
// down the thread's own columns -- no shuffles at all (the row-per-thread version needs a 31-shuffle butterfly per 32 columns
// for each of max and sum).  One partial per (128-row tile, channel): pool_max / pool_sum [M/128][Nout].
template <int NT>
__device__ __forceinline__ void linear_epilogue_tile_pool_swapped(const LinArgs& a, uint32_t tmem, int set, int nacc, int m0, int n0,
                                                                  int warp, int lane) {
  const int c = n0 + warp * 32 + lane;
  const bool live = c < a.Nout;
  const int cc = live ? c : a.Nout - 1;
  const uint32_t lane_base = ((uint32_t)(warp * 32) << 16) + set * nacc * NT;
  const float sc = a.scale ? __ldg(a.scale + cc) : 1.f;
  const float sh = a.shift ? __ldg(a.shift + (long long)(m0 / a.npc) * a.shift_ldb + cc) : 0.f;
  const int cols = min(128, a.M - m0);
  float mx = -INFINITY, sum = 0.f;
#pragma unroll 1
  for (int c0 = 0; c0 < 128; c0 += 32) {
    float v[32];
    tc::tmem_ld32(tmem + lane_base + c0, v);
    for (int ac = 1; ac < nacc; ++ac) {
      float w[32];
      tc::tmem_ld32(tmem + lane_base + ac * NT + c0, w);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += w[i];
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float y = v[i];
      if (a.scale) y *= sc;
      if (a.shift) y += sh;
      if (a.lrelu) y = y > 0.f ? y : 0.2f * y;
      if (c0 + i < cols) {
        mx = fmaxf(mx, y);
        sum += y;
      }
    }
  }
  if (live) {
    const long long o = (long long)(m0 >> 7) * a.Nout + c;
    a.pool_max[o] = mx;
    if (a.pool_sum) a.pool_sum[o] = sum;
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
