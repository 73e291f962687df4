// Exact-product GEMM on the bf16 tensor cores: C[b] = A[b] B[b]^T with fp32-class (<= 1 ulp-of-operand) accuracy and an
// accumulation that is EXACT, hence independent of tile order, chain length and tensor-core rounding mode.
//
// Why.  DownSampleToken's point score exponentiates q.k logits that reach |l| ~ 100 (models/downsample.py:139-147):
// a relative dot-product error eps becomes a relative score error of eps * sum|q_c k_c| / sqrt(D).  The 3xTF32 kernels
// (linear_tma.cu) lose ~2^-21 per product and their tensor-core accumulator truncates on every update (measured
// round 2: score error 1e-4 median / 7e-4 max vs 1e-5 for the reference's fp32 GEMM), which moved sampled indices and
// scaled whole attention rows.  Sampled indices must be bit-exact wherever fp32 can decide, so the two contractions
// that feed the score -- the q/k/v projection and q k^T -- run here instead.
//
// How (an Ozaki-style splitting).  Per cloud, x is scaled by a power of two so that |t| < 128 and cut into FOUR signed
// 8-bit digits  t = d0 + d1/2^8 + d2/2^16 + d3/2^24 (+ residual < 2^-25),  each stored as a bf16 INTEGER in [-128, 128]
// (exact).  32 bits below the cloud's ceiling: every element within a factor 2^8 of the largest is represented EXACTLY,
// the rest to 2^-32 of the ceiling.  A digit product is an integer < 2^14, a K=128 contraction of them < 2^21, and the
// partial sums grouped by weight,
//     G_g = sum over (i + j = g) of sum_c dA_i[c] dB_j[c],      g = 0..3       (at most 4 products: < 2^23)
// are integers the fp32 accumulator of tcgen05.mma holds EXACTLY, whatever its internal rounding.  Ten kind::f16 MMAs
// per K step (the tensor-pipe time of five kind::tf32 MMAs; 3xTF32 spends three) then
//     C = sA sB (G0 + G1/2^8 + G2/2^16 + G3/2^24)            (three fused roundings in the epilogue)
// with the dropped digit pairs (i + j >= 4) below 2^-30 K max|A| max|B|: C is the fp64 dot product rounded to fp32 up
// to ~1 ulp.  (Three digits -- six MMAs -- were measured first: 24 bits below the CLOUD's ceiling leave ~20 bits on a
// typical element, LSE error 2e-5 against 2e-6 for a true fp32 GEMM; the fourth digit removes that.)
// xgemm_ref_kernel restates the same integer arithmetic with int32 sums: the tests require BIT-IDENTICAL results,
// which is the proof that the tensor-core accumulation is exact.
//
// Kernel shape: persistent CTAs over (128-row tile of A) x (64-row tile of B) items, A's four digit planes resident in
// shared memory per row tile, B's planes streamed by TMA through an 8-stage ring of 8 KB boxes (128-byte swizzle), two
// TMEM sets of four 64-column accumulators (all 512 columns) so that the epilogue of an item overlaps the MMAs of the next.
// Warps 0-3 epilogue (thread = row = TMEM lane), warp 4 MMA issuer (elect.sync), warp 5 TMA producer.
// Epilogues: 0 = store fp32 rows (smem-staged, coalesced) + optional per-cloud |max| per column group (feeds the next
// slicing without another pass); 1 = online row statistics (max, sum exp) per item for DownSampleToken pass 1.
#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"

namespace samble {

constexpr int kXgThreads = 192;
constexpr int kXgStageBytes = 4 * 8192;   // one B stage = the four digit boxes [64 rows x 64 bf16] of one K tile
// ring depth: every stage costs the issuing threads a ~360-cycle mbarrier round trip (measured, tools/probe_xgemm.py:
// "handshakes only" = 93 us for 65k stages), so a stage must carry far more MMA time than that: 40 MMAs = 1920 cycles.
// (The first version used one 8 KB box per stage -- 4 to 16 MMAs -- and ran at 47 % tensor-pipe activity.)
template <int EPI> constexpr int xg_stages() { return EPI == 1 ? 3 : 2; }
constexpr int kXgTileN = 64;
constexpr int kXgDigits = 4;
constexpr int kXgSlabLd = 36;         // floats per staged row (pad 4: conflict-free 128-bit access)
constexpr int kXgSlabBytes = 4 * 32 * kXgSlabLd * 4;

// ------------------------------------------------------------------ digit planes
// exponent e with 2^e > amax (so |x * 2^(7-e)| < 128), clamped to a range whose powers of two are normal floats
__device__ __forceinline__ int xg_exponent(unsigned amax_bits) {
  const float amax = __uint_as_float(amax_bits);
  if (!(amax > 0.f)) return 0;
  int e = (int)((amax_bits >> 23) & 0xff) - 127 + 1;       // floor(log2(amax)) + 1   (denormals: e = -126)
  return max(-100, min(100, e));
}
__device__ __forceinline__ float xg_pow2(int e) { return __int_as_float((127 + e) << 23); }

__global__ void __launch_bounds__(256) xg_absmax_kernel(const float* __restrict__ x, long long ld, long long bs, int R, int C,
                                                        unsigned* __restrict__ amax) {
  const int b = blockIdx.y;
  const float* xb = x + (long long)b * bs;
  const int c4n = C >> 2;
  float mx = 0.f;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < (long long)R * c4n; v += (long long)gridDim.x * blockDim.x) {
    const long long r = v / c4n;
    const int c = (int)(v % c4n) * 4;
    const float4 t = __ldg(reinterpret_cast<const float4*>(xb + r * ld + c));
    mx = fmaxf(fmaxf(mx, fmaxf(fabsf(t.x), fabsf(t.y))), fmaxf(fabsf(t.z), fabsf(t.w)));
  }
  const unsigned m = __reduce_max_sync(kFull, __float_as_uint(mx));        // non-negative floats order like their bits
  if ((threadIdx.x & 31) == 0 && m) atomicMax(amax + b, m);
}

// x (B, R, C) rows -> kXgDigits bf16 digit planes (B, R, Cp), Cp = C rounded up to 64 (zero padded), and scale[b] = 2^(e-7).
// amax holds, per cloud and per `groups` column groups of C columns each (groups > 1: x is a (B, R, groups*C) buffer
// and this call slices group `g`), the bits of max|x|.
__global__ void __launch_bounds__(256) xg_slice_kernel(const float* __restrict__ x, long long ld, long long bs, int R, int C, int Cp,
                                                       const unsigned* __restrict__ amax, int amax_stride,
                                                       __nv_bfloat16* __restrict__ planes, long long plane_stride,
                                                       float* __restrict__ scale) {
  const int b = blockIdx.y;
  const int e = xg_exponent(amax[(long long)b * amax_stride]);
  const float up = xg_pow2(7 - e);
  if (blockIdx.x == 0 && threadIdx.x == 0) scale[b] = xg_pow2(e - 7);
  const float* xb = x + (long long)b * bs;
  const int c8n = Cp >> 3;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < (long long)R * c8n; v += (long long)gridDim.x * blockDim.x) {
    const long long r = v / c8n;
    const int c = (int)(v % c8n) * 8;
    float in[8];
#pragma unroll
    for (int h4 = 0; h4 < 2; ++h4) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c + 4 * h4 < C) t = __ldg(reinterpret_cast<const float4*>(xb + r * ld + c + 4 * h4));
      in[4 * h4] = t.x, in[4 * h4 + 1] = t.y, in[4 * h4 + 2] = t.z, in[4 * h4 + 3] = t.w;
    }
    uint32_t pk[kXgDigits][4];
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      float dg[2][kXgDigits];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float r = in[i + u] * up;                       // exact (power of two)
#pragma unroll
        for (int d = 0; d < kXgDigits; ++d) {
          dg[u][d] = rintf(r);
          r = (r - dg[u][d]) * 256.f;                   // exact: |r - digit| <= 1/2
        }
      }
      // integers of magnitude <= 128 are exact in bf16; low half-word = the lower channel
#pragma unroll
      for (int d = 0; d < kXgDigits; ++d) pk[d][i >> 1] = (__float_as_uint(dg[0][d]) >> 16) | (__float_as_uint(dg[1][d]) & 0xffff0000u);
    }
    const long long o = ((long long)b * R + r) * Cp + c;
#pragma unroll
    for (int d = 0; d < kXgDigits; ++d) *reinterpret_cast<uint4*>(planes + d * plane_stride + o) = make_uint4(pk[d][0], pk[d][1], pk[d][2], pk[d][3]);
  }
}

// ------------------------------------------------------------------ the GEMM
struct XgArgs {
  const float* a_scale;     // [Ba]
  const float* b_scale;     // [Bb]
  int Ba, Ra;               // clouds and rows per cloud of A
  int Bb, Rb;               // Bb == Ba (per-cloud B) or 1 (shared B); rows of B = output columns
  int Cp;                   // padded contraction length (64 or 128)
  // EPI 0
  float* out;               // [(b*Ra + i)*ldo + j]
  long long ldo;
  unsigned* amax_out;       // [Ba][ceil(Rb / amax_group)] bits of max|out| per cloud and column group, or null
  int amax_group;
  // EPI 1
  float2* stat_out;         // [(b*Ra + i)*nct + ct] = (max, sum exp) of out * logit_mul over the item's columns
  float logit_mul;
};

struct XgMaps {
  CUtensorMap a[kXgDigits], b[kXgDigits];       // digit planes of A (boxes of 128 rows) and of B (boxes of 64 rows)
};

template <int EPI, int NKT>
__global__ void __launch_bounds__(kXgThreads, 1)
    xgemm_kernel(const __grid_constant__ XgMaps maps, XgArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = tc::smem_align1024(smem_raw);
  constexpr int nkt = NKT;                                    // K tiles of 64 bf16 (128-byte rows): Cp = 64 * NKT
  uint8_t* sA = base;                                         // [plane][kt] tiles of 128 rows x 128 B
  uint8_t* sB = sA + (size_t)kXgDigits * nkt * 16384;         // ring
  constexpr int kStages = xg_stages<EPI>();
  float* slab = reinterpret_cast<float*>(sB + (size_t)kStages * kXgStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(slab) + (EPI == 0 ? kXgSlabBytes : 0));
  uint64_t* full = bars;                        // [kStages]
  uint64_t* empty = bars + kStages;             // [kStages]
  uint64_t* tfull = bars + 2 * kStages;         // [2]
  uint64_t* tempty = tfull + 2;                 // [2]
  uint64_t* afull = tempty + 2;                 // A row tile landed
  uint64_t* aempty = afull + 1;                 // ... no longer read by any MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tpc = (a.Ra + 127) >> 7;                          // row tiles per cloud
  const int nct = (a.Rb + kXgTileN - 1) / kXgTileN;
  const long long total = (long long)a.Ba * tpc * nct;
  const long long t0 = (long long)blockIdx.x * total / gridDim.x, t1 = (long long)(blockIdx.x + 1) * total / gridDim.x;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&tfull[i], 1);
      tc::mbar_init(&tempty[i], 4);
    }
    tc::mbar_init(afull, 1);
    tc::mbar_init(aempty, 1);
    tc::mbar_init_fence();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 5) {
    // ================= TMA producer (whole warp loops, one elected lane issues) =================
    {
      if (tc::elect_one()) {
        tc::tma_prefetch_desc(&maps.a[0]);
        tc::tma_prefetch_desc(&maps.b[0]);
      }
      int s = 0, ph = 0, aruns = 0;
      long long cur_rt = -1;
      for (long long item = t0; item < t1; ++item) {
        const long long rt = item / nct;
        const int ct = (int)(item % nct);
        const int b = (int)(rt / tpc), r0 = (int)(rt % tpc) * 128;
        if (rt != cur_rt) {
          tc::mbar_wait(aempty, (aruns & 1) ^ 1);             // MMAs of the previous row tile retired
          if (tc::elect_one()) {
            tc::mbar_arrive_expect_tx(afull, (uint32_t)(kXgDigits * nkt) * 16384u);
            for (int kt = 0; kt < nkt; ++kt) {
#pragma unroll
              for (int d = 0; d < kXgDigits; ++d) tc::tma_load_3d(sA + (size_t)(d * nkt + kt) * 16384, &maps.a[d], afull, kt * 64, r0, b);
            }
          }
          __syncwarp();
          cur_rt = rt;
          ++aruns;
        }
        const int bb = a.Bb > 1 ? b : 0;
        for (int kt = 0; kt < nkt; ++kt) {
          tc::mbar_wait(&empty[s], ph ^ 1);
          if (tc::elect_one()) {
            tc::mbar_arrive_expect_tx(&full[s], (uint32_t)kXgStageBytes);
#pragma unroll
            for (int p = 0; p < kXgDigits; ++p)
              tc::tma_load_3d(sB + (size_t)s * kXgStageBytes + p * 8192, &maps.b[p], &full[s], kt * 64, ct * kXgTileN, bb);
          }
          __syncwarp();
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 4) {
    // ================= MMA issuer =================
    // The WHOLE warp walks the loop (uniform control flow: stage indices, phases and descriptors stay in uniform
    // registers) and one elected lane issues each batch of MMAs + its commit.  With the loop inside `if (elect_one())`
    // every descriptor went through R2UR moves in front of each UTCHMMA and the single issuing thread paced the
    // kernel (ncu round 2: tensor pipe 47 % active, the epilogue warps idle 39 % of their samples on tfull).
    {
      const uint32_t idesc = tc::instr_desc(1, 128, kXgTileN);              // bf16 x bf16 -> fp32
      int s = 0, ph = 0, aruns = 0, it = 0;
      long long cur_rt = -1;
      const uint32_t a_lo0 = tc::smem_desc_sw128_lo(tc::smem_u32(sA)), b_lo0 = tc::smem_desc_sw128_lo(tc::smem_u32(sB));
      for (long long item = t0; item < t1; ++item, ++it) {
        const long long rt = item / nct;
        const int set = it & 1, use = it >> 1;
        if (rt != cur_rt) {
          tc::mbar_wait(afull, aruns & 1);
          cur_rt = rt;
          ++aruns;
        }
        tc::mbar_wait(&tempty[set], (use & 1) ^ 1);
        tc::tc_fence_after();
        const uint32_t g0 = tmem + set * kXgDigits * kXgTileN;            // accumulator G_g at g0 + g*64
#pragma unroll
        for (int kt = 0; kt < nkt; ++kt) {
          tc::mbar_wait(&full[s], ph);
          tc::tc_fence_after();
          if (tc::elect_one()) {
            const uint32_t b_lo = b_lo0 + (uint32_t)s * (kXgStageBytes >> 4);
#pragma unroll
            for (int p = 0; p < kXgDigits; ++p) {                         // B digit p pairs with A digits 0 .. D-1-p
#pragma unroll
              for (int k16 = 0; k16 < 4; ++k16) {
#pragma unroll
                for (int d = 0; d + p < kXgDigits; ++d)
                  // (the first write of every G_g happens with B digit 0 at the first K step; every operand tile is a
                  // compile-time offset from its base descriptor word: one 32-bit add per operand)
                  tc::mma_bf16_lo(g0 + (d + p) * kXgTileN, a_lo0 + (d * nkt + kt) * 1024 + 2 * k16, b_lo + p * 512 + 2 * k16, idesc,
                                  p == 0 ? (uint32_t)((kt | k16) != 0) : 1u);
              }
            }
            tc::mma_commit(&empty[s]);
            if (kt == nkt - 1) {
              tc::mma_commit(&tfull[set]);
              if (item + 1 >= t1 || (item + 1) / nct != rt) tc::mma_commit(aempty);   // last item of this row tile
            }
          }
          __syncwarp();
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ================= epilogue: thread = row =================
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    int it = 0;
    for (long long item = t0; item < t1; ++item, ++it) {
      const long long rt = item / nct;
      const int ct = (int)(item % nct);
      const int b = (int)(rt / tpc), r0 = (int)(rt % tpc) * 128;
      const int set = it & 1, use = it >> 1;
      const float sc = __ldg(a.a_scale + b) * __ldg(a.b_scale + (a.Bb > 1 ? b : 0));
      const int row = r0 + warp * 32 + lane;
      tc::mbar_wait(&tfull[set], use & 1);
      tc::tc_fence_after();
      const uint32_t tb = tmem + lane_base + set * kXgDigits * kXgTileN;
      float mx = -INFINITY, sum = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < kXgTileN; c0 += 32) {
        float v[32], w[32];
        tc::tmem_ld32(tb + (kXgDigits - 1) * kXgTileN + c0, v);           // lowest weight first
#pragma unroll
        for (int g = kXgDigits - 2; g >= 0; --g) {
          tc::tmem_ld32(tb + g * kXgTileN + c0, w);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaf(v[i], 0.00390625f, w[i]);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= sc;
        const int j0 = ct * kXgTileN + c0;
        if (EPI == 1) {
          if (j0 < a.Rb) {                                    // warp-uniform
            float cm = -INFINITY;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              v[i] = (j0 + i < a.Rb) ? v[i] * a.logit_mul : -INFINITY;
              cm = fmaxf(cm, v[i]);
            }
            const float mn = fmaxf(mx, cm);
            float part = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) part += __expf(v[i] - mn);
            sum = sum * __expf(mx - mn) + part;
            mx = mn;
          }
        } else {
          if (j0 >= a.Rb) continue;                           // warp-uniform
          if (a.amax_out) {
            float m = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) m = fmaxf(m, (j0 + i < a.Rb && row < a.Ra) ? fabsf(v[i]) : 0.f);
            const unsigned mb = __reduce_max_sync(kFull, __float_as_uint(m));
            const int ng = (a.Rb + a.amax_group - 1) / a.amax_group;
            if (lane == 0 && mb) atomicMax(a.amax_out + (long long)b * ng + j0 / a.amax_group, mb);
          }
          // smem-staged, coalesced row-major store: phase 1 thread = row dumps 32 columns, phase 2 eight lanes per row
          float* sl = slab + warp * 32 * kXgSlabLd;
#pragma unroll
          for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(sl + lane * kXgSlabLd + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          __syncwarp();
          const int rq = lane >> 3, cq = (lane & 7) * 4;
          const bool vec = a.ldo % 4 == 0 && reinterpret_cast<uintptr_t>(a.out) % 16 == 0 && j0 + cq + 4 <= a.Rb;
#pragma unroll
          for (int r8 = 0; r8 < 8; ++r8) {
            const int r = r8 * 4 + rq;
            const int grow = r0 + warp * 32 + r;
            if (grow >= a.Ra || j0 + cq >= a.Rb) continue;
            const float4 t = *reinterpret_cast<const float4*>(sl + r * kXgSlabLd + cq);
            float* op = a.out + ((long long)b * a.Ra + grow) * a.ldo + j0 + cq;
            if (vec) {
              *reinterpret_cast<float4*>(op) = t;
            } else {
              op[0] = t.x;
              if (j0 + cq + 1 < a.Rb) op[1] = t.y;
              if (j0 + cq + 2 < a.Rb) op[2] = t.z;
              if (j0 + cq + 3 < a.Rb) op[3] = t.w;
            }
          }
          __syncwarp();
        }
      }
      if (EPI == 1 && row < a.Ra) a.stat_out[((long long)b * a.Ra + row) * nct + ct] = make_float2(mx, sum);
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty[set]);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// Integer restatement (tests): one thread per output, int32 digit sums, the same two fused roundings.
__global__ void __launch_bounds__(256) xgemm_ref_kernel(const __nv_bfloat16* __restrict__ ap, long long a_plane,
                                                        const __nv_bfloat16* __restrict__ bp, long long b_plane, XgArgs a) {
  const long long total = (long long)a.Ba * a.Ra * a.Rb;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(o % a.Rb);
    const long long bi = o / a.Rb;                            // b*Ra + i
    const int b = (int)(bi / a.Ra);
    const int bb = a.Bb > 1 ? b : 0;
    const __nv_bfloat16* ar = ap + bi * a.Cp;
    const __nv_bfloat16* br = bp + ((long long)bb * a.Rb + j) * a.Cp;
    int g[kXgDigits];
    for (int t = 0; t < kXgDigits; ++t) g[t] = 0;
    for (int c = 0; c < a.Cp; ++c) {
      int da[kXgDigits], db[kXgDigits];
      for (int t = 0; t < kXgDigits; ++t) da[t] = (int)__bfloat162float(ar[t * a_plane + c]), db[t] = (int)__bfloat162float(br[t * b_plane + c]);
      for (int i = 0; i < kXgDigits; ++i)
        for (int j = 0; i + j < kXgDigits; ++j) g[i + j] += da[i] * db[j];
    }
    const float sc = a.a_scale[b] * a.b_scale[bb];
    float v = (float)g[kXgDigits - 1];
    for (int t = kXgDigits - 2; t >= 0; --t) v = fmaf(v, 0.00390625f, (float)g[t]);
    a.out[bi * a.ldo + j] = v * sc;
  }
}

static size_t xg_smem_bytes(int nkt, int epi) {
  return (size_t)kXgDigits * nkt * 16384 + (size_t)(epi ? xg_stages<1>() : xg_stages<0>()) * kXgStageBytes + (epi == 0 ? kXgSlabBytes : 0) + 512 + 1024;
}

static int launch_xgemm(const __nv_bfloat16* ap, const __nv_bfloat16* bp, const XgArgs& a, cudaStream_t st) {
  const long long a_plane = (long long)a.Ba * a.Ra * a.Cp, b_plane = (long long)a.Bb * a.Rb * a.Cp;
  alignas(64) XgMaps maps;
  for (int p = 0; p < kXgDigits; ++p) {
    if (int e = make_tile_map(&maps.a[p], ap + p * a_plane, a.Cp, a.Cp, a.Ra, a.Ba, 128, 2)) return e;
    if (int e = make_tile_map(&maps.b[p], bp + p * b_plane, a.Cp, a.Cp, a.Rb, a.Bb, kXgTileN, 2)) return e;
  }
  const int epi = a.stat_out ? 1 : 0;
  const size_t smem = xg_smem_bytes(a.Cp >> 6, epi);
  auto kern = epi ? (a.Cp == 64 ? xgemm_kernel<1, 1> : xgemm_kernel<1, 2>) : (a.Cp == 64 ? xgemm_kernel<0, 1> : xgemm_kernel<0, 2>);
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return check_launch("xgemm smem attribute");
  const long long total = (long long)a.Ba * ceil_div(a.Ra, 128) * ceil_div(a.Rb, kXgTileN);
  const int grid = (int)(total < 148 ? total : 148);
  SAMBLE_PRE(st);
  kern<<<grid, kXgThreads, smem, st>>>(maps, a);
  SAMBLE_LAUNCHED(epi ? "xgemm_rowstat_kernel" : "xgemm_store_kernel");
  return SAMBLE_OK;
}

// linear_tc.cu
int launch_ds_rowstats_finalize(const float2* part, int ntiles, const float* q, long long ldq, const float* k_tok, int M, int D,
                                int nb, float* rowmax, float* rowsum, float* token_logits, cudaStream_t st);

}  // namespace samble

using namespace samble;

extern "C" size_t samble_digits_bytes(int B, int R, int C) {
  if (B <= 0 || R <= 0 || C <= 0) return 0;
  return (size_t)kXgDigits * B * R * align_up(C, 64) * sizeof(__nv_bfloat16);
}

extern "C" int samble_digits(const float* x, long long ld, long long batch_stride, int B, int R, int C, const unsigned* amax_in,
                             int amax_stride, void* planes, float* scale, unsigned* amax_scratch, samble_stream_t stream) {
  SAMBLE_REQUIRE(x && planes && scale, "samble_digits: null pointer");
  SAMBLE_REQUIRE(amax_in || amax_scratch, "samble_digits: either amax_in or amax_scratch (B words) is required");
  SAMBLE_REQUIRE(B > 0 && R > 0 && C > 0 && B <= 65535, "samble_digits: bad shape");
  SAMBLE_REQUIRE(C % 4 == 0 && ld % 4 == 0 && batch_stride % 4 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)planes % 16 == 0,
                 "samble_digits: rows must be 16-byte aligned, C a multiple of 4");
  cudaStream_t st = (cudaStream_t)stream;
  const int Cp = (int)align_up(C, 64);
  const long long work = (long long)R * (Cp / 8);
  const int gx = (int)std::min<long long>((work + 255) / 256, 148 * 4);
  if (!amax_in) {
    if (cudaMemsetAsync(amax_scratch, 0, (size_t)B * sizeof(unsigned), st) != cudaSuccess) return check_launch("samble_digits memset");
    count_launch();
    SAMBLE_PRE(st);
    xg_absmax_kernel<<<dim3(gx, B), 256, 0, st>>>(x, ld, batch_stride, R, C, amax_scratch);
    SAMBLE_LAUNCHED("xg_absmax_kernel");
    amax_in = amax_scratch;
    amax_stride = 1;
  }
  SAMBLE_PRE(st);
  xg_slice_kernel<<<dim3(gx, B), 256, 0, st>>>(x, ld, batch_stride, R, C, Cp, amax_in, amax_stride, (__nv_bfloat16*)planes,
                                               (long long)B * R * Cp, scale);
  SAMBLE_LAUNCHED("xg_slice_kernel");
  return SAMBLE_OK;
}

static int xg_check(const void* ap, const float* as, int Ba, int Ra, const void* bp, const float* bs, int Bb, int Rb, int C) {
  SAMBLE_REQUIRE(ap && as && bp && bs, "samble_xgemm: null pointer");
  SAMBLE_REQUIRE(Ba > 0 && Ra > 0 && Rb > 0 && (Bb == Ba || Bb == 1), "samble_xgemm: bad shape (Bb must be Ba or 1)");
  SAMBLE_REQUIRE(C > 0 && C <= 128, "samble_xgemm: C=%d outside (0,128]", C);
  SAMBLE_REQUIRE(((uintptr_t)ap | (uintptr_t)bp) % 16 == 0, "samble_xgemm: digit planes must be 16-byte aligned");
  return SAMBLE_OK;
}

extern "C" int samble_xgemm(const void* a_planes, const float* a_scale, int Ba, int Ra, const void* b_planes, const float* b_scale,
                            int Bb, int Rb, int C, float* out, long long ldo, unsigned* amax_out, int amax_group, int reference,
                            samble_stream_t stream) {
  if (int e = xg_check(a_planes, a_scale, Ba, Ra, b_planes, b_scale, Bb, Rb, C)) return e;
  SAMBLE_REQUIRE(out && ldo >= Rb, "samble_xgemm: out/ldo");
  SAMBLE_REQUIRE(!amax_out || (amax_group > 0 && amax_group % 32 == 0), "samble_xgemm: amax_group must be a multiple of 32");
  cudaStream_t st = (cudaStream_t)stream;
  const int Cp = (int)align_up(C, 64);
  XgArgs a{a_scale, b_scale, Ba, Ra, Bb, Rb, Cp, out, ldo, amax_out, amax_group, nullptr, 1.f};
  if (amax_out) {
    if (cudaMemsetAsync(amax_out, 0, (size_t)Ba * ceil_div(Rb, amax_group) * sizeof(unsigned), st) != cudaSuccess)
      return check_launch("samble_xgemm memset");
    count_launch();
  }
  if (reference) {
    SAMBLE_REQUIRE(!amax_out, "samble_xgemm: the reference kernel has no amax output");
    const long long total = (long long)Ba * Ra * Rb;
    SAMBLE_PRE(st);
    xgemm_ref_kernel<<<(int)std::min<long long>((total + 255) / 256, 148 * 8), 256, 0, st>>>(
        (const __nv_bfloat16*)a_planes, (long long)Ba * Ra * Cp, (const __nv_bfloat16*)b_planes, (long long)Bb * Rb * Cp, a);
    SAMBLE_LAUNCHED("xgemm_ref_kernel");
    return SAMBLE_OK;
  }
  return launch_xgemm((const __nv_bfloat16*)a_planes, (const __nv_bfloat16*)b_planes, a, st);
}

extern "C" size_t samble_ds_row_stats_exact_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  return align_up((size_t)B * N * ceil_div(N, kXgTileN) * sizeof(float2), 256) + 256;
}

extern "C" int samble_ds_row_stats_exact(const void* q_planes, const float* q_scale, const void* k_planes, const float* k_scale,
                                         const float* q, long long ldq, const float* k_tok, int B, int N, int D, int nb,
                                         float* rowmax, float* rowsum, float* token_logits, void* ws, size_t ws_bytes,
                                         samble_stream_t stream) {
  if (int e = xg_check(q_planes, q_scale, B, N, k_planes, k_scale, B, N, D)) return e;
  SAMBLE_REQUIRE(q && rowmax && rowsum && ws, "samble_ds_row_stats_exact: null pointer");
  SAMBLE_REQUIRE(nb == 0 || (k_tok && token_logits), "samble_ds_row_stats_exact: token pointers required when nb > 0");
  SAMBLE_REQUIRE(nb >= 0 && nb <= 32 && D % 4 == 0 && ldq % 4 == 0 && (uintptr_t)q % 16 == 0, "samble_ds_row_stats_exact: bad q / nb");
  SAMBLE_REQUIRE(ws_bytes >= samble_ds_row_stats_exact_workspace_bytes(B, N), "samble_ds_row_stats_exact: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace w(ws, ws_bytes);
  const int nct = ceil_div(N, kXgTileN);
  float2* part = w.take<float2>((size_t)B * N * nct);
  XgArgs a{q_scale, k_scale, B, N, B, N, (int)align_up(D, 64), nullptr, 0, nullptr, 0, part, 1.f / sqrtf((float)D)};
  if (int e = launch_xgemm((const __nv_bfloat16*)q_planes, (const __nv_bfloat16*)k_planes, a, st)) return e;
  return launch_ds_rowstats_finalize(part, nct, q, ldq, k_tok, B * N, D, nb, rowmax, rowsum, token_logits, st);
}
