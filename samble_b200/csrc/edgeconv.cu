// Fused EdgeConv core (reference models/embedding.py:29-39): group -> conv1+BN+LeakyReLU -> conv2+BN+
// LeakyReLU -> max over the K neighbours, without ever writing the (B,2C,N,K) grouped tensor or the two
// (B,64,N,K) activations (the reference moves ~100 MB per cloud-layer through HBM for them).
//
// Algebra (eval mode; everything bias-free and linear up to the first LeakyReLU):
//   conv1([x_i ; x_j - x_i]) = (W1a - W1b) x_i + W1b x_j         -> per-POINT projections P_i, R_j
//   BN1 folds into P, R (scale) and P (shift):  h_ij = lrelu(P'_i + R'_j)
//   BN2 folds into W2 (row scale) and a bias:   z_ij = W2' h_ij + b2
//   max_j lrelu(z_ij) = lrelu(max_j z_ij)        (monotone)
// The caller supplies PR = [P' | R'] (B,N,2*C1) point-major (one small library GEMM over the N points),
// W2' (C2,C1) and b2 (C2).  This kernel does the gather, the 32 x C1 x C2 product per point and the max.
//
// CTA = 8 points; warp = one point; lane = C2/32 output channels for all 32 edges of its point, so the
// max over K is thread-local.  h tile (8 x K x C1) and W2' live in shared memory; the inner product
// reads h as warp-broadcast LDS.128 and W2' as conflict-free LDS.128 rows (stride C1+4).
#include "common.cuh"

namespace samble {

constexpr int kEcPoints = 8;     // points per CTA (= warps)
constexpr int kEcK = 32;         // max neighbours

template <class I, int CPT>      // CPT = output channels per lane (C2 = 32*CPT)
__global__ void __launch_bounds__(256, CPT <= 2 ? 2 : 1)
    edge_mlp_max_kernel(const float* __restrict__ pr, long long ld_pr, const I* __restrict__ idx,
                        const float* __restrict__ w2, const float* __restrict__ b2, int N, int K, int C1,
                        float* __restrict__ out /* (B, C2, N), or point-major rows when out_ld > 0 */, long long out_ld) {
  extern __shared__ __align__(16) float sm[];
  const int C2 = 32 * CPT;
  const int ldw = C1 + 4;
  float* W = sm;                                    // [C2][C1+4]
  float* H = W + (size_t)C2 * ldw;                  // [8][K][C1]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, n0 = blockIdx.x * kEcPoints;

  for (int t = threadIdx.x; t < C2 * (C1 / 4); t += blockDim.x) {
    const int c = t / (C1 / 4), k4 = (t % (C1 / 4)) * 4;
    *reinterpret_cast<float4*>(W + c * ldw + k4) = __ldg(reinterpret_cast<const float4*>(w2 + (size_t)c * C1 + k4));
  }
  // h tile: each warp builds the K x C1 block of its own point
  const int n = n0 + warp;
  const bool live = n < N;
  float* Hp = H + (size_t)warp * K * C1;
  if (live) {
    const long long row = (long long)b * N + n;
    const int my = lane < K ? ld_idx(idx, row * K + lane) : 0;
    const int f4 = C1 / 4;                          // float4 per row
    for (int t0 = 0; t0 < K * f4; t0 += 32) {       // uniform trip count: every lane takes part in the shuffle
      const int t = t0 + lane;
      const int e = min(t / f4, K - 1), c4 = (t % f4) * 4;
      const int j = __shfl_sync(kFull, my, e);      // each lane pulls the neighbour index of ITS edge
      if (t >= K * f4) continue;
      const float4 p = __ldg(reinterpret_cast<const float4*>(pr + row * ld_pr + c4));
      const float4 r = __ldg(reinterpret_cast<const float4*>(pr + ((long long)b * N + j) * ld_pr + C1 + c4));
      float4 h;
      h.x = p.x + r.x, h.y = p.y + r.y, h.z = p.z + r.z, h.w = p.w + r.w;
      h.x = h.x > 0.f ? h.x : 0.2f * h.x;
      h.y = h.y > 0.f ? h.y : 0.2f * h.y;
      h.z = h.z > 0.f ? h.z : 0.2f * h.z;
      h.w = h.w > 0.f ? h.w : 0.2f * h.w;
      *reinterpret_cast<float4*>(Hp + e * C1 + c4) = h;
    }
  }
  __syncthreads();
  if (!live) return;

  float acc[kEcK][CPT];
#pragma unroll
  for (int e = 0; e < kEcK; ++e)
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[e][c] = 0.f;
  for (int k4 = 0; k4 < C1; k4 += 4) {
    float4 w[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) w[c] = *reinterpret_cast<const float4*>(W + (lane + 32 * c) * ldw + k4);
#pragma unroll
    for (int e = 0; e < kEcK; ++e) {
      if (e < K) {
        const float4 h = *reinterpret_cast<const float4*>(Hp + e * C1 + k4);
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          float v = acc[e][c];
          v = fmaf(h.x, w[c].x, v);
          v = fmaf(h.y, w[c].y, v);
          v = fmaf(h.z, w[c].z, v);
          v = fmaf(h.w, w[c].w, v);
          acc[e][c] = v;
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    float m = -INFINITY;
#pragma unroll
    for (int e = 0; e < kEcK; ++e)
      if (e < K) m = fmaxf(m, acc[e][c]);
    const int ch = lane + 32 * c;
    m += __ldg(b2 + ch);
    out[out_ld ? ((long long)b * N + n) * out_ld + ch : ((long long)b * C2 + ch) * N + n] = m > 0.f ? m : 0.2f * m;
  }
}

template <class I, int CPT>
static int launch_edge_mlp(const float* pr, long long ld_pr, const I* idx, const float* w2, const float* b2, int B,
                           int N, int K, int C1, float* out, long long out_ld, cudaStream_t st) {
  const int C2 = 32 * CPT;
  size_t smem = ((size_t)C2 * (C1 + 4) + (size_t)kEcPoints * K * C1) * sizeof(float);
  auto kern = edge_mlp_max_kernel<I, CPT>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return check_launch("edge_mlp_max smem attribute");
  SAMBLE_PRE(st);
  kern<<<dim3(ceil_div(N, kEcPoints), B), 256, smem, st>>>(pr, ld_pr, idx, w2, b2, N, K, C1, out, out_ld);
  SAMBLE_LAUNCHED("edge_mlp_max_kernel");
  return SAMBLE_OK;
}

// tensor-core form (edgeconv_tc.cu)
bool edge_tc_eligible(int K, int C1, int C2);
template <class I>
int edge_mlp_tc(const float* pr, long long ld_pr, const I* idx, const float* w2, const float* b2, int B, int N, int K, int C1,
                int C2, float* out, long long out_ld, cudaStream_t st);
static int g_edge_mode = 0;   // 0 auto (tcgen05 when eligible), 1 FFMA kernel only

}  // namespace samble

using namespace samble;

extern "C" void samble_set_edge_mode(int mode) { g_edge_mode = mode; }

extern "C" int samble_edge_mlp_max(const float* pr, long long ld_pr, const void* idx, int idx_bits, const float* w2,
                                   const float* b2, int B, int N, int K, int C1, int C2, float* out, long long out_ld,
                                   samble_stream_t stream) {
  SAMBLE_REQUIRE(pr && idx && w2 && b2 && out, "samble_edge_mlp_max: null pointer");
  SAMBLE_REQUIRE(B > 0 && N > 0 && B <= 65535, "samble_edge_mlp_max: bad shape");
  SAMBLE_REQUIRE(K >= 1 && K <= kEcK, "samble_edge_mlp_max: K=%d outside [1,%d]", K, kEcK);
  SAMBLE_REQUIRE(C1 > 0 && C1 % 4 == 0 && C1 <= 128, "samble_edge_mlp_max: C1=%d must be a multiple of 4, <= 128", C1);
  SAMBLE_REQUIRE(C2 == 32 || C2 == 64 || C2 == 128, "samble_edge_mlp_max: C2=%d must be 32, 64 or 128", C2);
  SAMBLE_REQUIRE(ld_pr % 4 == 0 && ld_pr >= 2 * C1 && ((uintptr_t)pr | (uintptr_t)w2) % 16 == 0,
                 "samble_edge_mlp_max: PR/W2 need 16-byte aligned rows");
  SAMBLE_REQUIRE(idx_bits == 32 || idx_bits == 64, "samble_edge_mlp_max: idx_bits must be 32 or 64");
  SAMBLE_REQUIRE(out_ld == 0 || out_ld >= C2, "samble_edge_mlp_max: out_ld=%lld < C2", out_ld);
  cudaStream_t st = (cudaStream_t)stream;
  if (g_edge_mode == 0 && edge_tc_eligible(K, C1, C2)) {
    if (idx_bits == 64) return edge_mlp_tc<long long>(pr, ld_pr, (const long long*)idx, w2, b2, B, N, K, C1, C2, out, out_ld, st);
    return edge_mlp_tc<int>(pr, ld_pr, (const int*)idx, w2, b2, B, N, K, C1, C2, out, out_ld, st);
  }
#define EC_DISPATCH(CPT)                                                                                              \
  (idx_bits == 64 ? launch_edge_mlp<long long, CPT>(pr, ld_pr, (const long long*)idx, w2, b2, B, N, K, C1, out, out_ld, st)   \
                  : launch_edge_mlp<int, CPT>(pr, ld_pr, (const int*)idx, w2, b2, B, N, K, C1, out, out_ld, st))
  if (C2 == 32) return EC_DISPATCH(1);
  if (C2 == 64) return EC_DISPATCH(2);
  return EC_DISPATCH(4);
#undef EC_DISPATCH
}
