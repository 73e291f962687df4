// Neighbor2Point attention core (reference models/attention.py:165-185, 207-250; scalar_dot, asm "dot").
//
// The reference gathers (B,C,N,K) neighbour differences and pushes them through the bias-free
// k/v 1x1 convolutions: 2 x 33.5 MB of activations per cloud-layer at N=2048.  Both convolutions
// are linear, so  W(x_j - x_i) = W x_j - W x_i :  the caller projects the N points once and this
// kernel only gathers projected rows.  The -q_i.(Wk x_i) term is constant over j and cancels in the
// softmax; the value term contributes -(Wv x_i) because the weights sum to one.
//
// One warp per point, lanes across channels (float4 per lane), heads = groups of (C/H)/4 lanes.
// Traffic per point: K rows of k and v (L2 hits: one cloud's projections are <= 2 MB).
#include "common.cuh"

namespace samble {

template <class I, int KMAX>
__global__ void __launch_bounds__(256) n2p_attend_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                         const float* __restrict__ v, long long ld,
                                                         const I* __restrict__ idx, int N, int C, int K, int lph,
                                                         float sqrt_d, const float* __restrict__ residual, long long ld_res,
                                                         const float* __restrict__ scale, const float* __restrict__ shift,
                                                         float* __restrict__ out, long long ld_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int n = blockIdx.x * 8 + warp;
  if (n >= N) return;
  const bool on = lane * 4 < C;                       // lanes beyond the channel count idle
  const long long row = (long long)b * N + n;
  const int my = lane < K ? ld_idx(idx, row * K + lane) : 0;
  float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (on) q4 = *reinterpret_cast<const float4*>(q + row * ld + lane * 4);

  float lg[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) {
    lg[j] = -INFINITY;
    if (j < K) {
      const int nj = __shfl_sync(kFull, my, j);
      float part = 0.f;
      if (on) {
        const float4 k4 = __ldg(reinterpret_cast<const float4*>(k + ((long long)b * N + nj) * ld + lane * 4));
        part = q4.x * k4.x;
        part = fmaf(q4.y, k4.y, part);
        part = fmaf(q4.z, k4.z, part);
        part = fmaf(q4.w, k4.w, part);
      }
      for (int o = lph >> 1; o > 0; o >>= 1) part += __shfl_xor_sync(kFull, part, o);
      lg[j] = part / sqrt_d;
    }
  }
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < KMAX; ++j) m = fmaxf(m, lg[j]);
  float s = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int j = 0; j < KMAX; ++j) {
    if (j < K) {
      const int nj = __shfl_sync(kFull, my, j);
      const float p = expf(lg[j] - m);
      s += p;
      if (on) {
        const float4 v4 = __ldg(reinterpret_cast<const float4*>(v + ((long long)b * N + nj) * ld + lane * 4));
        acc.x = fmaf(p, v4.x, acc.x);
        acc.y = fmaf(p, v4.y, acc.y);
        acc.z = fmaf(p, v4.z, acc.z);
        acc.w = fmaf(p, v4.w, acc.w);
      }
    }
  }
  if (on) {
    const float4 vi = *reinterpret_cast<const float4*>(v + row * ld + lane * 4);
    float4 o;
    o.x = acc.x / s - vi.x;
    o.y = acc.y / s - vi.y;
    o.z = acc.z / s - vi.z;
    o.w = acc.w / s - vi.w;
    if (residual) {   // fused  bn1(x + attention)  of attention.py:187 (eval-mode BN folded to scale/shift)
      const float4 r = *reinterpret_cast<const float4*>(residual + row * ld_res + lane * 4);
      o.x += r.x, o.y += r.y, o.z += r.z, o.w += r.w;
    }
    if (scale) {
      const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + lane * 4));
      const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + lane * 4));
      o.x = fmaf(o.x, sc.x, sh.x), o.y = fmaf(o.y, sc.y, sh.y), o.z = fmaf(o.z, sc.z, sh.z), o.w = fmaf(o.w, sc.w, sh.w);
    }
    *reinterpret_cast<float4*>(out + row * ld_out + lane * 4) = o;
  }
}


// ---- eight lanes per point, C = 128, four heads (the shipped configs) -- round 2 ---------------------------------------
// ncu on the warp-per-point kernel above (profiles/r2_step_full.md): 3131 warp instructions per point, issue slots 59 %
// busy -- instruction-bound on per-neighbour bookkeeping (index shuffle, address arithmetic, three shuffle+add steps for
// FOUR useful FMAs per lane).  A first eight-lane version gave each lane 16 CONTIGUOUS channels: 5x fewer instructions
// but no faster -- every LDG.128 then touched 32 half-used sectors and L1/TEX throughput went from 32 % to 93 %
// (profiles/r2_n2p_attend.md).  Here the eight lanes of a point interleave: lane `sub` owns float4 number c*8 + sub of the
// row (c = 0..3), so one load instruction of a group covers 128 contiguous bytes = 4 whole sectors, and a warp works on 4
// points at once.  With D = 32 channels per head, float4 number f belongs to head f / 8 = c: every lane holds one partial
// dot product PER HEAD, and a transposing butterfly (2 + 1 + 1 shuffles) leaves head (sub >> 1)'s logit in lane `sub`.
template <class I>
__global__ void __launch_bounds__(256, 3) n2p_attend8_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                              const float* __restrict__ v, long long ld,
                                                              const I* __restrict__ idx, int N, int K, float sqrt_d,
                                                              const float* __restrict__ residual, long long ld_res,
                                                              const float* __restrict__ scale, const float* __restrict__ shift,
                                                              float* __restrict__ out, long long ld_out) {
  const int lane = threadIdx.x & 31, sub = lane & 7;
  const int b = blockIdx.y;
  const int n_raw = blockIdx.x * 32 + (threadIdx.x >> 3);
  const bool live = n_raw < N;
  const int n = live ? n_raw : N - 1;                 // dead groups shadow the last point (shuffles stay converged), no store
  const long long row = (long long)b * N + n;
  int my[4];                                          // lane `sub` of a group holds neighbours 4*sub .. 4*sub+3
#pragma unroll
  for (int t = 0; t < 4; ++t) my[t] = (sub * 4 + t) < K ? ld_idx(idx, row * K + sub * 4 + t) : n;
  float4 qv[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) qv[c] = reinterpret_cast<const float4*>(q + row * ld)[c * 8 + sub];
  const float* kb = k + (long long)b * N * ld;
  const float* vb = v + (long long)b * N * ld;
  const bool hi4 = (sub & 4) != 0, hi2 = (sub & 2) != 0;

  float lg[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    lg[j] = -INFINITY;
    if (j < K) {
      const int nj = __shfl_sync(kFull, my[j & 3], j >> 2, 8);
      const float4* kp = reinterpret_cast<const float4*>(kb + (long long)nj * ld) + sub;
      float pr[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 k4 = __ldg(kp + c * 8);
        pr[c] = fmaf(qv[c].w, k4.w, fmaf(qv[c].z, k4.z, fmaf(qv[c].y, k4.y, qv[c].x * k4.x)));
      }
      // transposing butterfly over the 8 lanes: 4 partials per lane -> the full sum of head (sub >> 1) in lane sub
      float a0 = hi4 ? pr[2] : pr[0], a1 = hi4 ? pr[3] : pr[1];
      a0 += __shfl_xor_sync(kFull, hi4 ? pr[0] : pr[2], 4);
      a1 += __shfl_xor_sync(kFull, hi4 ? pr[1] : pr[3], 4);
      float t = hi2 ? a1 : a0;
      t += __shfl_xor_sync(kFull, hi2 ? a0 : a1, 2);
      t += __shfl_xor_sync(kFull, t, 1);
      lg[j] = t / sqrt_d;
    }
  }
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < 32; ++j) m = fmaxf(m, lg[j]);
  float s = 0.f;
  float4 acc[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j < K) {
      const int nj = __shfl_sync(kFull, my[j & 3], j >> 2, 8);
      const float4* vp = reinterpret_cast<const float4*>(vb + (long long)nj * ld) + sub;
      const float p = expf(lg[j] - m);                // head sub >> 1
      s += p;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float pc = __shfl_sync(kFull, p, 2 * c, 8);       // head c's probability for this neighbour
        const float4 v4 = __ldg(vp + c * 8);
        acc[c].x = fmaf(pc, v4.x, acc[c].x);
        acc[c].y = fmaf(pc, v4.y, acc[c].y);
        acc[c].z = fmaf(pc, v4.z, acc[c].z);
        acc[c].w = fmaf(pc, v4.w, acc[c].w);
      }
    }
  }
  float sc4[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) sc4[c] = __shfl_sync(kFull, s, 2 * c, 8);
  if (!live) return;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int f = c * 8 + sub;
    const float4 vi = reinterpret_cast<const float4*>(v + row * ld)[f];
    float4 o;
    o.x = acc[c].x / sc4[c] - vi.x;
    o.y = acc[c].y / sc4[c] - vi.y;
    o.z = acc[c].z / sc4[c] - vi.z;
    o.w = acc[c].w / sc4[c] - vi.w;
    if (residual) {
      const float4 r = reinterpret_cast<const float4*>(residual + row * ld_res)[f];
      o.x += r.x, o.y += r.y, o.z += r.z, o.w += r.w;
    }
    if (scale) {
      const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + f);
      const float4 sh = __ldg(reinterpret_cast<const float4*>(shift) + f);
      o.x = fmaf(o.x, sc.x, sh.x), o.y = fmaf(o.y, sc.y, sh.y), o.z = fmaf(o.z, sc.z, sh.z), o.w = fmaf(o.w, sc.w, sh.w);
    }
    reinterpret_cast<float4*>(out + row * ld_out)[f] = o;
  }
}

static int g_n2p_mode = 0;   // 0 auto (eight lanes per point when the shape allows), 1 warp-per-point kernel only

}  // namespace samble

using namespace samble;

extern "C" void samble_set_n2p_mode(int mode) { samble::g_n2p_mode = mode; }

extern "C" int samble_n2p_attend(const float* q, const float* k, const float* v, long long ld, const void* idx,
                                 int idx_bits, int B, int N, int C, int K, int heads, const float* residual,
                                 long long ld_res, const float* scale, const float* shift, float* out, long long ld_out,
                                 samble_stream_t stream) {
  SAMBLE_REQUIRE(q && k && v && idx && out, "samble_n2p_attend: null pointer");
  SAMBLE_REQUIRE(B > 0 && N > 0 && K > 0, "samble_n2p_attend: bad shape");
  SAMBLE_REQUIRE(K <= 32, "samble_n2p_attend: K=%d > 32", K);
  SAMBLE_REQUIRE(C > 0 && C % 4 == 0 && C <= 128, "samble_n2p_attend: C=%d must be a multiple of 4, <= 128", C);
  SAMBLE_REQUIRE(heads > 0 && C % heads == 0 && (C / heads) % 4 == 0, "samble_n2p_attend: C/heads must be a multiple of 4");
  const int lph = C / heads / 4;
  SAMBLE_REQUIRE((lph & (lph - 1)) == 0, "samble_n2p_attend: (C/heads)/4 = %d must be a power of two", lph);
  SAMBLE_REQUIRE(ld % 4 == 0 && ld_out % 4 == 0, "samble_n2p_attend: leading dimensions must be multiples of 4");
  SAMBLE_REQUIRE(((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) % 16 == 0, "samble_n2p_attend: 16-byte alignment required");
  SAMBLE_REQUIRE(idx_bits == 32 || idx_bits == 64, "samble_n2p_attend: idx_bits must be 32 or 64");
  SAMBLE_REQUIRE((scale == nullptr) == (shift == nullptr), "samble_n2p_attend: scale and shift go together");
  SAMBLE_REQUIRE(!residual || (ld_res % 4 == 0 && (uintptr_t)residual % 16 == 0), "samble_n2p_attend: residual alignment");
  SAMBLE_REQUIRE(!scale || ((uintptr_t)scale | (uintptr_t)shift) % 16 == 0, "samble_n2p_attend: scale/shift alignment");
  cudaStream_t st = (cudaStream_t)stream;
  const float sqrt_d = sqrtf((float)(C / heads));
  // eight lanes per point: the shipped shape (C = 128, four heads of 32 channels)
  if (g_n2p_mode == 0 && C == 128 && heads == 4) {
    dim3 grid8(ceil_div(N, 32), B);
    SAMBLE_PRE(st);
    if (idx_bits == 64)
      n2p_attend8_kernel<long long><<<grid8, 256, 0, st>>>(q, k, v, ld, (const long long*)idx, N, K, sqrt_d, residual, ld_res, scale, shift, out, ld_out);
    else
      n2p_attend8_kernel<int><<<grid8, 256, 0, st>>>(q, k, v, ld, (const int*)idx, N, K, sqrt_d, residual, ld_res, scale, shift, out, ld_out);
    SAMBLE_LAUNCHED("n2p_attend8_kernel");
    return SAMBLE_OK;
  }
  dim3 grid(ceil_div(N, 8), B);
  SAMBLE_PRE(st);
  if (idx_bits == 64)
    n2p_attend_kernel<long long, 32><<<grid, 256, 0, st>>>(q, k, v, ld, (const long long*)idx, N, C, K, lph, sqrt_d, residual, ld_res, scale, shift, out, ld_out);
  else
    n2p_attend_kernel<int, 32><<<grid, 256, 0, st>>>(q, k, v, ld, (const int*)idx, N, C, K, lph, sqrt_d, residual, ld_res, scale, shift, out, ld_out);
  SAMBLE_LAUNCHED("n2p_attend_kernel");
  return SAMBLE_OK;
}
