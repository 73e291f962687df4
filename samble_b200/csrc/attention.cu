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

}  // namespace samble

using namespace samble;

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
  dim3 grid(ceil_div(N, 8), B);
  const float sqrt_d = sqrtf((float)(C / heads));
  SAMBLE_PRE(st);
  if (idx_bits == 64)
    n2p_attend_kernel<long long, 32><<<grid, 256, 0, st>>>(q, k, v, ld, (const long long*)idx, N, C, K, lph, sqrt_d, residual, ld_res, scale, shift, out, ld_out);
  else
    n2p_attend_kernel<int, 32><<<grid, 256, 0, st>>>(q, k, v, ld, (const int*)idx, N, C, K, lph, sqrt_d, residual, ld_res, scale, shift, out, ld_out);
  SAMBLE_LAUNCHED("n2p_attend_kernel");
  return SAMBLE_OK;
}
