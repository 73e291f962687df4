// DownSampleToken, the attention rows of the M selected points times V (reference models/downsample.py:242-252),
// flash-style: the (B, M, N) probability slab is never written.
//
// The reference gathers M rows of the N x (N + nb) softmax map and multiplies them by V^T.  The row statistics
// (max, sum of exp) of EVERY row are already known from pass 1 (samble_ds_row_stats_exact), so no online rescaling is
// needed: one CTA owns 128 selected rows of one cloud and walks the keys in tiles of 64,
//     S = Q_sel K_tile^T        exact-product digit GEMM (xgemm.cu; the first three digits: six kind::f16 MMAs per K step)
//     P = exp(S / sqrt(D) - max_i) / sum_i        in the epilogue warps, straight from TMEM, split into bf16 hi + lo
//     O += P_hi V_hi + P_lo V_hi + P_hi V_lo      second UMMA, P as the A operand from shared memory, V^T tiles by TMA
// and finally adds the nb token columns' share sum_t softmax(row)[N + t] * v_tok[t] and stores the (B, M, C) rows.
// TMEM: two sets of three 64-column S accumulators (384) + the 128-column O accumulator = 512 columns.
// Warps 0-3: gather Q rows, S -> P, final epilogue (thread = selected row = TMEM lane); warp 4: MMA issuer; warp 5: TMA.
// Small layers under-fill the chip (M/128 * B CTAs), so the key range can be split in two: both halves atomicAdd their
// partial O onto a zeroed output -- two addends commute, so the result stays bit-reproducible.
#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"

namespace samble {

constexpr int kDaThreads = 192;
constexpr int kDaDigits = 3;          // digits of q and k used for S (of the 4 stored)
constexpr int kDaTileN = 64;          // keys per tile
constexpr int kDaKSlots = 6;          // K ring: one tile of boxes [64 keys x 64 bf16] (8 KB each)

struct DaMaps {
  CUtensorMap k[kDaDigits];           // digit planes of k, boxes of 64 keys x 64 channels
  CUtensorMap vh, vl;                 // V^T bf16 hi / lo planes (B, C, N), boxes of 128 channels x 64 keys
};

struct DaArgs {
  const __nv_bfloat16* q_planes;      // [digit][B][N][Cp]
  long long q_plane_stride;
  const float* q_scale;               // [B]
  const float* k_scale;               // [B]
  const long long* idx;               // (B, M) selected points
  const float* rowmax;                // (B, N)
  const float* rowsum;                // (B, N)
  const float* tok_logits;            // (B, N, nb) pre-softmax token columns
  const float* v_tok;                 // (nb, C)
  float* out;                         // (B, M, C) rows
  int B, N, M, Cp, C, nb, ksplit;
  float inv_sqrt_d;
};

// v (B, N, C) fp32 rows -> V^T as two bf16 planes (B, C, Npad): hi = bf16(v), lo = bf16(v - hi).  32 x 32 tiles via smem.
__global__ void __launch_bounds__(256) ds_vt_split_kernel(const float* __restrict__ v, long long ldv, int N, int C, int Npad,
                                                          __nv_bfloat16* __restrict__ vt_hi, __nv_bfloat16* __restrict__ vt_lo) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + r, c = c0 + tx;
    tile[r][tx] = (n < N && c < C) ? v[((long long)b * N + n) * ldv + c] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, n = n0 + tx;
    if (c < C && n < Npad) {
      const float x = tile[tx][r];
      const __nv_bfloat16 hi = __float2bfloat16_rn(x);
      const long long o = ((long long)b * C + c) * Npad + n;
      vt_hi[o] = hi;
      vt_lo[o] = __float2bfloat16_rn(x - __bfloat162float(hi));
    }
  }
}

__global__ void __launch_bounds__(kDaThreads, 1) ds_attend_rows_kernel(const __grid_constant__ DaMaps maps, DaArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = tc::smem_align1024(smem_raw);
  const int nkt = a.Cp >> 6;
  uint8_t* sQ = base;                                           // [digit][kt] tiles of 128 rows x 128 B (gathered)
  uint8_t* sK = sQ + (size_t)kDaDigits * nkt * 16384;           // ring of 8 KB boxes
  uint8_t* sV = sK + (size_t)kDaKSlots * 8192;                  // V^T hi | lo: 128 channels x 128 B each
  uint8_t* sP = sV + 32768;                                     // P hi | lo: 128 rows x 128 B each
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 32768);
  uint64_t* kfull = bars;                      // [kDaKSlots]
  uint64_t* kempty = bars + kDaKSlots;         // [kDaKSlots]
  uint64_t* sfull = bars + 2 * kDaKSlots;      // [2] S accumulator set complete
  uint64_t* sempty = sfull + 2;                // [2] ... drained by the 4 epilogue warps
  uint64_t* vfull = sempty + 2;
  uint64_t* vempty = vfull + 1;
  uint64_t* pfull = vempty + 1;                // P tile written (4 warp arrivals)
  uint64_t* pempty = pfull + 1;                // ... consumed by the P V MMAs
  uint64_t* qfull = pempty + 1;                // Q rows gathered (4 warp arrivals)
  uint64_t* ofull = qfull + 1;                 // every P V MMA retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ofull + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, m0 = blockIdx.x * 128;
  const int ntiles = (a.N + kDaTileN - 1) / kDaTileN;
  const int t_begin = (int)((long long)blockIdx.z * ntiles / a.ksplit), t_end = (int)((long long)(blockIdx.z + 1) * ntiles / a.ksplit);

  if (tid == 0) {
    for (int s = 0; s < kDaKSlots; ++s) {
      tc::mbar_init(&kfull[s], 1);
      tc::mbar_init(&kempty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&sfull[i], 1);
      tc::mbar_init(&sempty[i], 4);
    }
    tc::mbar_init(vfull, 1);
    tc::mbar_init(vempty, 1);
    tc::mbar_init(pfull, 4);
    tc::mbar_init(pempty, 1);
    tc::mbar_init(qfull, 4);
    tc::mbar_init(ofull, 1);
    tc::mbar_init_fence();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_o = tmem + 2 * kDaDigits * kDaTileN;       // columns 384..511

  if (warp == 5) {
    // ================= TMA producer: K(t0); then K(t+1), V(t) =================
    if (tc::elect_one()) {
      tc::tma_prefetch_desc(&maps.k[0]);
      tc::tma_prefetch_desc(&maps.vh);
      int s = 0, ph = 0;
      auto load_k = [&](int t) {
        for (int kt = 0; kt < nkt; ++kt) {
#pragma unroll
          for (int p = 0; p < kDaDigits; ++p) {
            tc::mbar_wait(&kempty[s], ph ^ 1);
            tc::mbar_arrive_expect_tx(&kfull[s], 8192u);
            tc::tma_load_3d(sK + (size_t)s * 8192, &maps.k[p], &kfull[s], kt * 64, t * kDaTileN, b);
            if (++s == kDaKSlots) { s = 0; ph ^= 1; }
          }
        }
      };
      if (t_begin < t_end) load_k(t_begin);
      for (int t = t_begin; t < t_end; ++t) {
        if (t + 1 < t_end) load_k(t + 1);
        const int u = t - t_begin;
        tc::mbar_wait(vempty, (u & 1) ^ 1);                    // P V MMAs of the previous tile retired
        tc::mbar_arrive_expect_tx(vfull, 32768u);
        tc::tma_load_3d(sV, &maps.vh, vfull, t * kDaTileN, 0, b);
        tc::tma_load_3d(sV + 16384, &maps.vl, vfull, t * kDaTileN, 0, b);
      }
    }
    __syncwarp();
  } else if (warp == 4) {
    // ================= MMA issuer: S(t0); then S(t+1), P V(t) =================
    // whole warp walks the loops (uniform control flow keeps descriptors in uniform registers), one elected lane issues
    {
      const uint32_t idesc_s = tc::instr_desc(1, 128, kDaTileN);
      const uint32_t idesc_o = tc::instr_desc(1, 128, 128);
      const uint32_t q_base = tc::smem_u32(sQ), k_base = tc::smem_u32(sK);
      int s = 0, ph = 0;
      auto issue_s = [&](int u) {                               // u = tile ordinal within this CTA
        const int set = u & 1, use = u >> 1;
        tc::mbar_wait(&sempty[set], (use & 1) ^ 1);
        tc::tc_fence_after();
        const uint32_t g0 = tmem + set * kDaDigits * kDaTileN;
        for (int kt = 0; kt < nkt; ++kt) {
#pragma unroll
          for (int p = 0; p < kDaDigits; ++p) {
            tc::mbar_wait(&kfull[s], ph);
            tc::tc_fence_after();
            if (tc::elect_one()) {
              const uint64_t kd = tc::smem_desc_sw128(k_base + s * 8192);
#pragma unroll
              for (int k16 = 0; k16 < 4; ++k16) {
#pragma unroll
                for (int d = 0; d + p < kDaDigits; ++d)
                  tc::mma_bf16(g0 + (d + p) * kDaTileN, tc::smem_desc_sw128(q_base + (d * nkt + kt) * 16384) + 2 * k16, kd + 2 * k16, idesc_s,
                               p == 0 ? (uint32_t)((kt | k16) != 0) : 1u);
              }
              tc::mma_commit(&kempty[s]);
              if (p == kDaDigits - 1 && kt == nkt - 1) tc::mma_commit(&sfull[set]);
            }
            __syncwarp();
            if (++s == kDaKSlots) { s = 0; ph ^= 1; }
          }
        }
      };
      tc::mbar_wait(qfull, 0);
      tc::tc_fence_after();
      const int nt = t_end - t_begin;
      if (nt > 0) issue_s(0);
      const uint64_t ph_d = tc::smem_desc_sw128(tc::smem_u32(sP)), pl_d = tc::smem_desc_sw128(tc::smem_u32(sP + 16384));
      const uint64_t vh_d = tc::smem_desc_sw128(tc::smem_u32(sV)), vl_d = tc::smem_desc_sw128(tc::smem_u32(sV + 16384));
      for (int u = 0; u < nt; ++u) {
        if (u + 1 < nt) issue_s(u + 1);
        tc::mbar_wait(pfull, u & 1);
        tc::mbar_wait(vfull, u & 1);
        tc::tc_fence_after();
        if (tc::elect_one()) {
#pragma unroll
          for (int k16 = 0; k16 < 4; ++k16) {
            tc::mma_bf16(tmem_o, ph_d + 2 * k16, vh_d + 2 * k16, idesc_o, (uint32_t)((u | k16) != 0));
            tc::mma_bf16(tmem_o, pl_d + 2 * k16, vh_d + 2 * k16, idesc_o, 1u);
            tc::mma_bf16(tmem_o, ph_d + 2 * k16, vl_d + 2 * k16, idesc_o, 1u);
          }
          tc::mma_commit(pempty);
          tc::mma_commit(vempty);
          if (u == nt - 1) tc::mma_commit(ofull);
        }
        __syncwarp();
      }
    }
  } else {
    // ================= warps 0-3: thread = selected row =================
    const int row = warp * 32 + lane;
    const int m = m0 + row;
    const bool live = m < a.M;
    const long long src = live ? a.idx[(long long)b * a.M + m] : 0;
    // ---- gather the row's digits into the swizzled A tiles
    {
      const __nv_bfloat16* qrow = a.q_planes + ((long long)b * a.N + src) * a.Cp;
#pragma unroll
      for (int d = 0; d < kDaDigits; ++d) {
        for (int kt = 0; kt < nkt; ++kt) {
          uint4 c[8];
#pragma unroll
          for (int ch = 0; ch < 8; ++ch)
            c[ch] = live ? __ldg(reinterpret_cast<const uint4*>(qrow + d * a.q_plane_stride + kt * 64) + ch) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int ch = 0; ch < 8; ++ch)
            *reinterpret_cast<uint4*>(sQ + (size_t)(d * nkt + kt) * 16384 + tc::sw128_offset(row, ch)) = c[ch];
        }
      }
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(qfull);
    }
    const float mu = live ? a.rowmax[(long long)b * a.N + src] : 0.f;
    const float inv_s = live ? 1.f / a.rowsum[(long long)b * a.N + src] : 0.f;
    const float sc = __ldg(a.q_scale + b) * __ldg(a.k_scale + b) * a.inv_sqrt_d;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const int nt = t_end - t_begin;
    for (int u = 0; u < nt; ++u) {
      const int set = u & 1, use = u >> 1;
      tc::mbar_wait(&sfull[set], use & 1);
      tc::tc_fence_after();
      const uint32_t tb = tmem + lane_base + set * kDaDigits * kDaTileN;
      uint32_t hi[32], lo[32];                                  // 64 probabilities as packed bf16 pairs
#pragma unroll
      for (int c0 = 0; c0 < kDaTileN; c0 += 32) {
        float v[32], w[32];
        tc::tmem_ld32(tb + (kDaDigits - 1) * kDaTileN + c0, v);
#pragma unroll
        for (int g = kDaDigits - 2; g >= 0; --g) {
          tc::tmem_ld32(tb + g * kDaTileN + c0, w);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaf(v[i], 0.00390625f, w[i]);
        }
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = __expf(fmaf(v[i], sc, -mu)) * inv_s, p1 = __expf(fmaf(v[i + 1], sc, -mu)) * inv_s;
          const __nv_bfloat16 h0 = __float2bfloat16_rn(p0), h1 = __float2bfloat16_rn(p1);
          const __nv_bfloat16 l0 = __float2bfloat16_rn(p0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(p1 - __bfloat162float(h1));
          hi[(c0 + i) >> 1] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          lo[(c0 + i) >> 1] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&sempty[set]);             // the S set may be overwritten
      tc::mbar_wait(pempty, (u & 1) ^ 1);                       // P V MMAs of the previous tile retired
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const uint32_t off = tc::sw128_offset(row, ch);
        *reinterpret_cast<uint4*>(sP + off) = make_uint4(hi[4 * ch], hi[4 * ch + 1], hi[4 * ch + 2], hi[4 * ch + 3]);
        *reinterpret_cast<uint4*>(sP + 16384 + off) = make_uint4(lo[4 * ch], lo[4 * ch + 1], lo[4 * ch + 2], lo[4 * ch + 3]);
      }
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(pfull);
    }
    // ---- final epilogue: O (+ the token columns' share) -> out rows
    if (nt > 0) {
      tc::mbar_wait(ofull, 0);
      tc::tc_fence_after();
    }
    float ptok[8];
    const bool add_tok = blockIdx.z == 0 && a.nb > 0;
#pragma unroll
    for (int t = 0; t < 8; ++t)
      ptok[t] = (add_tok && live && t < a.nb) ? __fdiv_rn(expf(a.tok_logits[((long long)b * a.N + src) * a.nb + t] - mu), a.rowsum[(long long)b * a.N + src]) : 0.f;
    float* orow = a.out + ((long long)b * a.M + (live ? m : 0)) * a.C;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      float v[32];
      if (nt > 0) {
        tc::tmem_ld32(tmem_o + lane_base + c0, v);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
      }
      if (c0 >= a.C || !live) continue;
      if (add_tok) {
        for (int t = 0; t < a.nb && t < 8; ++t) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < a.C) v[i] = fmaf(ptok[t], __ldg(a.v_tok + (long long)t * a.C + c0 + i), v[i]);
        }
      }
      if (a.ksplit > 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c0 + i < a.C) atomicAdd(orow + c0 + i, v[i]);
      } else if (c0 + 32 <= a.C && a.C % 4 == 0) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(orow + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c0 + i < a.C) orow[c0 + i] = v[i];
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

static size_t da_smem_bytes(int nkt) { return (size_t)kDaDigits * nkt * 16384 + (size_t)kDaKSlots * 8192 + 32768 + 32768 + 512 + 1024; }
static int da_npad(int N) { return (int)align_up(N, 8); }

}  // namespace samble

using namespace samble;

extern "C" size_t samble_ds_attend_rows_workspace_bytes(int B, int N, int C) {
  if (B <= 0 || N <= 0 || C <= 0) return 0;
  return 2 * align_up((size_t)B * C * da_npad(N) * sizeof(__nv_bfloat16), 256) + 256;
}

extern "C" int samble_ds_attend_rows(const void* q_planes, const float* q_scale, const void* k_planes, const float* k_scale,
                                     const float* v, long long ldv, const long long* idx, const float* rowmax, const float* rowsum,
                                     const float* token_logits, const float* v_tok, int B, int N, int M, int D, int C, int nb,
                                     float* out, void* ws, size_t ws_bytes, samble_stream_t stream) {
  SAMBLE_REQUIRE(q_planes && q_scale && k_planes && k_scale && v && idx && rowmax && rowsum && out && ws,
                 "samble_ds_attend_rows: null pointer");
  SAMBLE_REQUIRE(nb == 0 || (token_logits && v_tok), "samble_ds_attend_rows: token pointers required when nb > 0");
  SAMBLE_REQUIRE(B > 0 && N > 0 && M > 0 && B <= 65535, "samble_ds_attend_rows: bad shape");
  SAMBLE_REQUIRE(D > 0 && D <= 128 && C > 0 && C <= 128 && nb >= 0 && nb <= 8, "samble_ds_attend_rows: D, C <= 128 and nb <= 8");
  SAMBLE_REQUIRE(((uintptr_t)q_planes | (uintptr_t)k_planes | (uintptr_t)out) % 16 == 0, "samble_ds_attend_rows: 16-byte alignment");
  SAMBLE_REQUIRE(ws_bytes >= samble_ds_attend_rows_workspace_bytes(B, N, C), "samble_ds_attend_rows: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int Cp = (int)align_up(D, 64), Npad = da_npad(N);
  Workspace w(ws, ws_bytes);
  __nv_bfloat16* vt_hi = w.take<__nv_bfloat16>((size_t)B * C * Npad);
  __nv_bfloat16* vt_lo = w.take<__nv_bfloat16>((size_t)B * C * Npad);
  SAMBLE_PRE(st);
  ds_vt_split_kernel<<<dim3(ceil_div(Npad, 32), ceil_div(C, 32), B), 256, 0, st>>>(v, ldv, N, C, Npad, vt_hi, vt_lo);
  SAMBLE_LAUNCHED("ds_vt_split_kernel");
  alignas(64) DaMaps maps;
  const long long plane = (long long)B * N * Cp;
  for (int p = 0; p < kDaDigits; ++p)
    if (int e = make_tile_map(&maps.k[p], (const __nv_bfloat16*)k_planes + p * plane, Cp, Cp, N, B, kDaTileN, 2)) return e;
  if (int e = make_tile_map(&maps.vh, vt_hi, Npad, Npad, C, B, 128, 2)) return e;
  if (int e = make_tile_map(&maps.vl, vt_lo, Npad, Npad, C, B, 128, 2)) return e;
  const int rt = ceil_div(M, 128);
  const int ksplit = (rt * B * 2 <= 148 && ceil_div(N, kDaTileN) >= 4) ? 2 : 1;
  if (ksplit > 1) {
    if (cudaMemsetAsync(out, 0, (size_t)B * M * C * sizeof(float), st) != cudaSuccess) return check_launch("samble_ds_attend_rows memset");
    count_launch();
  }
  DaArgs a{(const __nv_bfloat16*)q_planes, plane, q_scale, k_scale, idx, rowmax, rowsum, token_logits, v_tok, out,
           B, N, M, Cp, C, nb, ksplit, 1.f / sqrtf((float)D)};
  const size_t smem = da_smem_bytes(Cp >> 6);
  if (cudaFuncSetAttribute(ds_attend_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return check_launch("ds_attend_rows smem attribute");
  SAMBLE_PRE(st);
  ds_attend_rows_kernel<<<dim3(rt, B, ksplit), kDaThreads, smem, st>>>(maps, a);
  SAMBLE_LAUNCHED("ds_attend_rows_kernel");
  return SAMBLE_OK;
}
