// DownSampleToken attention-map scoring (reference models/downsample.py:124-153, 300-344).
//
// Reference: energy (B,1,N,N+nb) -> /sqrt(D) -> softmax -> * dense kNN mask -> column sum / indeg^2:
// five N x N fp32 temporaries per cloud.  Here:
//   pass 1  ds_row_stats : q k^T tiles (shared FFMA tile engine) with an online max / sum-of-exp per
//                          row; only (B,N) statistics and the nb pre-softmax token columns are written.
//   pass 2  ds_edge_score: only the N*K kNN edges are re-evaluated in fp32 FFMA and reduced per destination column
//                          in a FIXED order (per-warp private partial columns, no float atomics) => deterministic.
//                          (pass 1 normally runs on the tensor cores -- ds_rowstats_tc.cu / linear_tma.cu; the FFMA
//                          kernel below is the cross-check, samble_set_ds_mode(1).)
#include "common.cuh"
#include "gemm_tile.cuh"

namespace samble {

template <class Cfg>
struct RowStatsEpilogue {
  float* rm;   // smem [TQ] running max
  float* rs;   // smem [TQ] running sum of exp(l - max)
  int q0, Nq, Nr;
  float scale;

  __device__ __forceinline__ void tile(const float* S, int ldS, int n0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int rr = 0; rr < Cfg::RPW; ++rr) {
      const int row = warp * Cfg::RPW + rr;
      if (q0 + row >= Nq) break;
      float l[4], tmax = -INFINITY;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = n0 + lane + 32 * u;
        l[u] = j < Nr ? __fdiv_rn(S[row * ldS + lane + 32 * u], scale) : -INFINITY;
        tmax = fmaxf(tmax, l[u]);
      }
      tmax = warp_max(tmax);
      const float m_old = rm[row], m_new = fmaxf(m_old, tmax);
      float ps = 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) ps += expf(l[u] - m_new);   // exp(-inf) = 0 for masked columns
      ps = warp_sum(ps);
      __syncwarp();
      if (lane == 0) {
        rs[row] = rs[row] * expf(m_old - m_new) + ps;
        rm[row] = m_new;
      }
      __syncwarp();
    }
  }
};

template <class Cfg>
__global__ void __launch_bounds__(256, 2)
    ds_row_stats_kernel(const float* __restrict__ q, long long ldq, const float* __restrict__ k, long long ldk,
                        const float* __restrict__ k_tok, int N, int D, int nb, float scale,
                        float* __restrict__ rowmax, float* __restrict__ rowsum, float* __restrict__ token_logits) {
  extern __shared__ __align__(16) float smem[];
  const int b = blockIdx.y, q0 = blockIdx.x * Cfg::TQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* rm = smem + Cfg::smem_floats(D);
  float* rs = rm + Cfg::TQ;
  for (int r = threadIdx.x; r < Cfg::TQ; r += blockDim.x) rm[r] = -INFINITY, rs[r] = 0.f;
  __syncthreads();
  RowStatsEpilogue<Cfg> epi{rm, rs, q0, N, N, scale};
  dot_tiles<Cfg>(q + (long long)b * N * ldq, ldq, q0, N, k + (long long)b * N * ldk, ldk, N, D, D, smem, epi);
  __syncthreads();
  // token columns (downsample.py:116-118,149-152): nb extra keys shared by the whole batch
  const float* As = smem;
  const int lda = Cfg::lda(D);
  for (int rr = 0; rr < Cfg::RPW; ++rr) {
    const int row = warp * Cfg::RPW + rr, i = q0 + row;
    if (i >= N) break;
    float lt = -INFINITY;
    if (lane < nb) {
      float acc = 0.f;
      for (int c = 0; c < D; ++c) acc = fmaf(As[row * lda + c], __ldg(k_tok + lane * D + c), acc);
      lt = __fdiv_rn(acc, scale);
      token_logits[((long long)b * N + i) * nb + lane] = lt;
    }
    const float tmax = warp_max(lt);
    const float m_old = rm[row], m_new = fmaxf(m_old, tmax);
    const float ps = warp_sum(lane < nb ? expf(lt - m_new) : 0.f);
    if (lane == 0) {
      rowsum[(long long)b * N + i] = rs[row] * expf(m_old - m_new) + ps;
      rowmax[(long long)b * N + i] = m_new;
    }
  }
}

// ---- pass 2 -------------------------------------------------------------------------------
// grid (P, B); CTA = W warps; warp w of chunk p owns rows [ (p*W + w)*RW, +RW ) and walks them in order.
// Per row the warp fetches the K neighbour key rows with coalesced 16-byte loads (lane = 4-channel chunk, all K loads
// in flight at once), multiplies by its chunk of q_i, and reduces the 32 x 32 partial products with a halving
// butterfly (31 shuffles) that leaves neighbour e's logit in lane e.  Lane e then adds p_ie to ITS column of the
// warp-private column sums (the K neighbours of a row are distinct => race-free, no float atomics, fixed order).
template <class I>
__global__ void __launch_bounds__(256) ds_edge_partial_kernel(const float* __restrict__ q, long long ldq,
                                                              const float* __restrict__ k, long long ldk,
                                                              const float* __restrict__ rowmax,
                                                              const float* __restrict__ rowsum,
                                                              const I* __restrict__ idx, int N, int D, int K, int RW,
                                                              float scale, float* __restrict__ part,
                                                              int* __restrict__ indeg) {
  extern __shared__ __align__(16) float sm[];
  const int W = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, p = blockIdx.x, P = gridDim.x;
  float* col = sm + (size_t)warp * N;                 // this warp's private partial column sums
  for (int j = lane; j < N; j += 32) col[j] = 0.f;
  __syncwarp();
  const int nch = D >> 2;                             // 16-byte chunks per row
  const int r0 = (p * W + warp) * RW;
  for (int i = r0; i < min(r0 + RW, N); ++i) {
    const long long row = (long long)b * N + i;
    const float rmax = rowmax[row], rsum = rowsum[row];
    for (int e0 = 0; e0 < K; e0 += 32) {
      const int e = e0 + lane;
      const int j_mine = e < K ? (int)ld_idx(idx, row * K + e) : 0;
      float acc[32];
      int jt[32];                                     // the 32 neighbour indices, broadcast while the warp is converged
#pragma unroll
      for (int t = 0; t < 32; ++t) acc[t] = 0.f, jt[t] = __shfl_sync(kFull, j_mine, t);
      for (int c = lane; c < nch; c += 32) {          // lanes beyond the chunk count (D < 128) skip: no shuffles inside
        const float4 qv = __ldg(reinterpret_cast<const float4*>(q + row * ldq) + c);
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          const int j = jt[t];
          const float4 kv = __ldg(reinterpret_cast<const float4*>(k + ((long long)b * N + j) * ldk) + c);
          acc[t] = fmaf(qv.x, kv.x, acc[t]);
          acc[t] = fmaf(qv.y, kv.y, acc[t]);
          acc[t] = fmaf(qv.z, kv.z, acc[t]);
          acc[t] = fmaf(qv.w, kv.w, acc[t]);
        }
      }
#pragma unroll
      for (int w = 16; w >= 1; w >>= 1) {
        const bool upper = (lane & w) != 0;
#pragma unroll
        for (int t = 0; t < w; ++t) {
          const float keep = upper ? acc[t + w] : acc[t], send = upper ? acc[t] : acc[t + w];
          acc[t] = keep + __shfl_xor_sync(kFull, send, w);
        }
      }
      if (e < K) {                                    // acc[0] = logit numerator of neighbour e0 + lane
        const float pr = __fdiv_rn(expf(__fdiv_rn(acc[0], scale) - rmax), rsum);
        col[j_mine] += pr;
        atomicAdd(indeg + (long long)b * N + j_mine, 1);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  float* dst = part + ((long long)b * P + p) * N;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < W; ++w) s += sm[(size_t)w * N + j];
    dst[j] = s;
  }
}

__global__ void __launch_bounds__(256) ds_edge_finalize_kernel(const float* __restrict__ part,
                                                               const int* __restrict__ indeg, int N, int P,
                                                               float* __restrict__ score) {
  const int b = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  float s = 0.f;
  for (int p = 0; p < P; ++p) s += part[((long long)b * P + p) * N + j];
  // sparse_num = mask.sum(-2) + 1e-8 (downsample.py:311); score = colsum / num / num (:335-338)
  const float num = __fadd_rn((float)indeg[(long long)b * N + j], 1e-8f);
  float v = __fdiv_rn(__fdiv_rn(s, num), num);
  score[(long long)b * N + j] = (v != v) ? 0.f : v;      // NaN -> 0 (:342)
}

struct EdgePlan {
  int W, RW, P;
  size_t smem, bytes;
};
static EdgePlan edge_plan(int B, int N, int D) {
  EdgePlan e;
  size_t budget = 200 * 1024;
  int W = (int)((budget - 8 * (size_t)D * 4) / ((size_t)N * 4));
  e.W = W > 8 ? 8 : W;
  if (e.W < 1) e.W = 0;
  int rows_per_cta_min = e.W ? e.W : 1;
  // at most 32 row chunks per cloud (bounds the partial buffer), at least 8 rows per warp
  int RW = ceil_div(N, 32 * rows_per_cta_min);
  e.RW = RW < 8 ? 8 : RW;
  e.P = e.W ? ceil_div(N, e.W * e.RW) : 0;
  e.smem = ((size_t)e.W * N + (size_t)e.W * D) * sizeof(float);
  e.bytes = align_up((size_t)B * e.P * N * sizeof(float), 256) + align_up((size_t)B * N * sizeof(int), 256) + 512;
  return e;
}

// tensor-core form (ds_rowstats_tc.cu)
bool ds_row_stats_tc_eligible(int D, int nb, long long ldq, long long ldk);
int launch_ds_row_stats_tc(const float* q, long long ldq, const float* k, long long ldk, const float* k_tok, int B, int N, int D,
                           int nb, float* rowmax, float* rowsum, float* token_logits, cudaStream_t st);
static int g_ds_mode = 0;   // 0 auto (tcgen05 when eligible), 1 exact FFMA tile kernel only

}  // namespace samble

using namespace samble;

extern "C" void samble_set_ds_mode(int mode) { g_ds_mode = mode; }

extern "C" int samble_ds_row_stats(const float* q, long long ldq, const float* k, long long ldk, const float* k_tok,
                                   int B, int N, int D, int nb, float* rowmax, float* rowsum, float* token_logits,
                                   samble_stream_t stream) {
  SAMBLE_REQUIRE(q && k && rowmax && rowsum, "samble_ds_row_stats: null pointer");
  SAMBLE_REQUIRE(nb == 0 || (k_tok && token_logits), "samble_ds_row_stats: token pointers required when nb > 0");
  SAMBLE_REQUIRE(B > 0 && N > 0 && B <= 65535, "samble_ds_row_stats: bad shape");
  SAMBLE_REQUIRE(D > 0 && D % 16 == 0 && D <= 256, "samble_ds_row_stats: D=%d must be a multiple of 16, <= 256", D);
  SAMBLE_REQUIRE(nb >= 0 && nb <= 32, "samble_ds_row_stats: nb=%d outside [0,32]", nb);
  SAMBLE_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ((uintptr_t)q | (uintptr_t)k) % 16 == 0,
                 "samble_ds_row_stats: q/k need 16-byte aligned rows");
  if (g_ds_mode == 0 && ds_row_stats_tc_eligible(D, nb, ldq, ldk))
    return launch_ds_row_stats_tc(q, ldq, k, ldk, k_tok, B, N, D, nb, rowmax, rowsum, token_logits, (cudaStream_t)stream);
  using Cfg = DotTileCfg<4>;
  size_t smem = (Cfg::smem_floats(D) + 2 * Cfg::TQ) * sizeof(float);
  auto kern = ds_row_stats_kernel<Cfg>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return check_launch("ds_row_stats smem attribute");
  const float scale = sqrtf((float)D);
  SAMBLE_PRE((cudaStream_t)stream);
  kern<<<dim3(ceil_div(N, Cfg::TQ), B), 256, smem, (cudaStream_t)stream>>>(q, ldq, k, ldk, k_tok, N, D, nb, scale, rowmax,
                                                                          rowsum, token_logits);
  SAMBLE_LAUNCHED("ds_row_stats_kernel");
  return SAMBLE_OK;
}

extern "C" size_t samble_ds_edge_score_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  return edge_plan(B, N, 256).bytes;
}

extern "C" int samble_ds_edge_score(const float* q, long long ldq, const float* k, long long ldk, const float* rowmax,
                                    const float* rowsum, const void* idx, int idx_bits, int B, int N, int D, int K,
                                    float* score, void* ws, size_t ws_bytes, samble_stream_t stream) {
  SAMBLE_REQUIRE(q && k && rowmax && rowsum && idx && score && ws, "samble_ds_edge_score: null pointer");
  SAMBLE_REQUIRE(B > 0 && N > 0 && K > 0 && B <= 65535, "samble_ds_edge_score: bad shape");
  SAMBLE_REQUIRE(D > 0 && D % 4 == 0 && D <= 256, "samble_ds_edge_score: D=%d must be a multiple of 4, <= 256", D);
  SAMBLE_REQUIRE(ldk % 4 == 0 && (uintptr_t)k % 16 == 0 && ldq % 4 == 0 && (uintptr_t)q % 16 == 0,
                 "samble_ds_edge_score: q and k need 16-byte aligned rows");
  SAMBLE_REQUIRE(idx_bits == 32 || idx_bits == 64, "samble_ds_edge_score: idx_bits must be 32 or 64");
  EdgePlan e = edge_plan(B, N, 256);
  SAMBLE_REQUIRE(e.W >= 1, "samble_ds_edge_score: N=%d too large for one shared-memory column buffer", N);
  SAMBLE_REQUIRE(ws_bytes >= e.bytes, "samble_ds_edge_score: workspace %zu < %zu bytes", ws_bytes, e.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  Workspace w(ws, ws_bytes);
  float* part = w.take<float>((size_t)B * e.P * N);
  int* indeg = w.take<int>((size_t)B * N);
  if (cudaMemsetAsync(indeg, 0, (size_t)B * N * sizeof(int), st) != cudaSuccess) return check_launch("memset indeg");
  count_launch();
  const float scale = sqrtf((float)D);
  size_t smem = ((size_t)e.W * N + (size_t)e.W * D) * sizeof(float);
  if (idx_bits == 64) {
    auto kern = ds_edge_partial_kernel<long long>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SAMBLE_PRE(st);
    kern<<<dim3(e.P, B), e.W * 32, smem, st>>>(q, ldq, k, ldk, rowmax, rowsum, (const long long*)idx, N, D, K, e.RW, scale, part, indeg);
  } else {
    auto kern = ds_edge_partial_kernel<int>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SAMBLE_PRE(st);
    kern<<<dim3(e.P, B), e.W * 32, smem, st>>>(q, ldq, k, ldk, rowmax, rowsum, (const int*)idx, N, D, K, e.RW, scale, part, indeg);
  }
  SAMBLE_LAUNCHED("ds_edge_partial_kernel");
  SAMBLE_PRE(st);
  ds_edge_finalize_kernel<<<dim3(ceil_div(N, 256), B), 256, 0, st>>>(part, indeg, N, e.P, score);
  SAMBLE_LAUNCHED("ds_edge_finalize_kernel");
  return SAMBLE_OK;
}

// ---- the M selected rows (models/downsample.py:242-252): one pass gathers what the two per-cloud GEMMs need ----
// warp per selected row: q row -> q_sel; its softmax statistics -> m_sel / s_sel; and the nb token columns' share of
// the output, tok_mix = sum_t softmax(row)[N + t] * v_tok[t], which the second GEMM takes as its residual.
__global__ void __launch_bounds__(256) ds_select_rows_kernel(const float* __restrict__ q, long long ldq,
                                                             const float* __restrict__ rowmax, const float* __restrict__ rowsum,
                                                             const float* __restrict__ token_logits,
                                                             const float* __restrict__ v_tok, const long long* __restrict__ idx,
                                                             int N, int M, int D, int nb, int C, float* __restrict__ q_sel,
                                                             float* __restrict__ m_sel, float* __restrict__ s_sel,
                                                             float* __restrict__ tok_mix) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, i = blockIdx.x * 8 + warp;
  if (i >= M) return;
  const long long o = (long long)b * M + i;
  const long long src = (long long)b * N + idx[o];
  for (int c = lane * 4; c < D; c += 128)
    *reinterpret_cast<float4*>(q_sel + o * D + c) = __ldg(reinterpret_cast<const float4*>(q + src * ldq + c));
  const float mx = rowmax[src], sm = rowsum[src];
  if (lane == 0) m_sel[o] = mx, s_sel[o] = sm;
  for (int c = lane * 4; c < C; c += 128) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < nb; ++t) {
      const float p = __fdiv_rn(expf(token_logits[src * nb + t] - mx), sm);
      const float4 v = __ldg(reinterpret_cast<const float4*>(v_tok + (long long)t * C + c));
      acc.x = fmaf(p, v.x, acc.x), acc.y = fmaf(p, v.y, acc.y), acc.z = fmaf(p, v.z, acc.z), acc.w = fmaf(p, v.w, acc.w);
    }
    *reinterpret_cast<float4*>(tok_mix + o * C + c) = acc;
  }
}

extern "C" int samble_ds_select_rows(const float* q, long long ldq, const float* rowmax, const float* rowsum,
                                     const float* token_logits, const float* v_tok, const long long* idx, int B, int N, int M,
                                     int D, int nb, int C, float* q_sel, float* m_sel, float* s_sel, float* tok_mix,
                                     samble_stream_t stream) {
  SAMBLE_REQUIRE(q && rowmax && rowsum && idx && q_sel && m_sel && s_sel && tok_mix, "samble_ds_select_rows: null pointer");
  SAMBLE_REQUIRE(nb == 0 || (token_logits && v_tok), "samble_ds_select_rows: token pointers required when nb > 0");
  SAMBLE_REQUIRE(B > 0 && N > 0 && M > 0 && B <= 65535, "samble_ds_select_rows: bad shape");
  SAMBLE_REQUIRE(D % 4 == 0 && C % 4 == 0 && ldq % 4 == 0 && ((uintptr_t)q | (uintptr_t)q_sel | (uintptr_t)tok_mix | (uintptr_t)v_tok) % 16 == 0,
                 "samble_ds_select_rows: 16-byte aligned rows, D and C multiples of 4");
  cudaStream_t st = (cudaStream_t)stream;
  SAMBLE_PRE(st);
  ds_select_rows_kernel<<<dim3(ceil_div(M, 8), B), 256, 0, st>>>(q, ldq, rowmax, rowsum, token_logits, v_tok, idx, N, M, D, nb, C,
                                                                 q_sel, m_sel, s_sel, tok_mix);
  SAMBLE_LAUNCHED("ds_select_rows_kernel");
  return SAMBLE_OK;
}

