// kNN for SAMBLE clouds: replaces reference utils/ops.py:17-44 (mean/std normalisation, torch.cdist,
// topk) with   stats -> normalise -> fused (distance tile + k-selection)   kernels.
// The (B,Nq,Nr) distance matrix is never written to HBM.
#include <cuda_bf16.h>

#include "common.cuh"
#include "gemm_tile.cuh"

namespace samble {

// ------------------------------------------------------------------------------------------
// per-cloud, per-channel mean and unbiased std of the QUERY cloud (ops.py:23,27).
// one warp per channel; fp64 one-pass sums (exactly-rounded-quality results, order-independent
// to well below fp32 resolution).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) knn_stats_kernel(const float* __restrict__ a, long long sb, long long sn,
                                                        long long sc, int N, int C, float* __restrict__ mean,
                                                        float* __restrict__ stdv) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + warp, b = blockIdx.y;
  if (c >= C) return;
  const float* p = a + b * sb + c * sc;
  // 4 independent accumulator pairs: the loads of one trip do not wait on the previous trip's adds
  double sa[4] = {0.0, 0.0, 0.0, 0.0}, qa[4] = {0.0, 0.0, 0.0, 0.0};
  int n = lane;
  for (; n + 96 < N; n += 128) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = p[(long long)(n + 32 * u) * sn];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      sa[u] += (double)v[u];
      qa[u] += (double)v[u] * (double)v[u];
    }
  }
  for (; n < N; n += 32) {
    const double v = (double)p[(long long)n * sn];
    sa[0] += v;
    qa[0] += v * v;
  }
  double s = warp_sum((sa[0] + sa[1]) + (sa[2] + sa[3]));
  double s2 = warp_sum((qa[0] + qa[1]) + (qa[2] + qa[3]));
  if (lane == 0) {
    double m = s / N;
    double var = (s2 - s * m) / (double)(N - 1);
    mean[b * C + c] = (float)m;
    stdv[b * C + c] = (float)sqrt(var > 0.0 ? var : 0.0);
  }
}

// Point-major input (unit channel stride): lane = channel, so a warp reads 128 contiguous bytes of one point; the 16
// warps of the CTA take points w, w+16, ... and their fp64 partial sums are combined in warp order (deterministic).
__global__ void __launch_bounds__(512) knn_stats_pm_kernel(const float* __restrict__ a, long long sb, long long sn, int N,
                                                           int C, double* __restrict__ part, unsigned* __restrict__ zero_me) {
  __shared__ double ps[16][32], pq[16][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 32 + lane, b = blockIdx.y, z = blockIdx.z, Z = gridDim.z;
  // (the per-cloud max-norm word the next kernel of the neighbour search accumulates into with atomicMax: cleared here
  // instead of by a memset node of its own)
  if (zero_me && threadIdx.x == 0 && blockIdx.x == 0 && z == 0) zero_me[b] = 0u;
  const bool live = c < C;
  const float* p = a + b * sb + (live ? c : 0);
  // this CTA's slice of the points: [n_lo, n_hi)
  const int per = (N + Z - 1) / Z, n_lo = z * per, n_hi = min(N, n_lo + per);
  double sa[4] = {0.0, 0.0, 0.0, 0.0}, qa[4] = {0.0, 0.0, 0.0, 0.0};
  int n = n_lo + warp;
  for (; n + 48 < n_hi; n += 64) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = p[(long long)(n + 16 * u) * sn];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      sa[u] += (double)v[u];
      qa[u] += (double)v[u] * (double)v[u];
    }
  }
  for (; n < n_hi; n += 16) {
    const double v = (double)p[(long long)n * sn];
    sa[0] += v;
    qa[0] += v * v;
  }
  ps[warp][lane] = (sa[0] + sa[1]) + (sa[2] + sa[3]);
  pq[warp][lane] = (qa[0] + qa[1]) + (qa[2] + qa[3]);
  __syncthreads();
  if (warp == 0 && live) {
    double s = 0.0, s2 = 0.0;
    for (int w = 0; w < 16; ++w) {
      s += ps[w][lane];
      s2 += pq[w][lane];
    }
    double* o = part + (((long long)b * Z + z) * C + c) * 2;
    o[0] = s, o[1] = s2;
  }
}

// slices combined in slice order (deterministic)
__global__ void knn_stats_pm_combine_kernel(const double* __restrict__ part, int Z, int N, int C, float* __restrict__ mean,
                                            float* __restrict__ stdv) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (c >= C) return;
  double s = 0.0, s2 = 0.0;
  for (int z = 0; z < Z; ++z) {
    const double* o = part + (((long long)b * Z + z) * C + c) * 2;
    s += o[0], s2 += o[1];
  }
  const double m = s / N;
  const double var = (s2 - s * m) / (double)(N - 1);
  mean[b * C + c] = (float)m;
  stdv[b * C + c] = (float)sqrt(var > 0.0 ? var : 0.0);
}

// sigma = mean over channels of the per-channel std (ops.py:27), summed in channel order.
__device__ __forceinline__ float cloud_sigma(const float* stdv, int C) {
  float s = 0.f;
  for (int c = 0; c < C; ++c) s = __fadd_rn(s, stdv[c]);
  return __fdiv_rn(s, (float)C);
}

// xyz path (C <= 3): one float4 per point = (x', y', z', |p'|^2), primes = normalised (ops.py:24-29).
__global__ void __launch_bounds__(256) knn_prep_xyz_kernel(const float* __restrict__ x, long long sb, long long sn,
                                                           long long sc, int N, int C, const float* __restrict__ mean,
                                                           const float* __restrict__ stdv, float4* __restrict__ out) {
  const int b = blockIdx.y, n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float sigma = cloud_sigma(stdv + b * C, C);
  float v[3] = {0.f, 0.f, 0.f};
  for (int c = 0; c < C; ++c) v[c] = __fdiv_rn(__fsub_rn(x[b * sb + n * sn + c * sc], mean[b * C + c]), sigma);
  // x.pow(2).sum(-1) of at::_euclidean_dist: products rounded separately, summed in channel order
  float nn = __fmul_rn(v[0], v[0]);
  nn = __fadd_rn(nn, __fmul_rn(v[1], v[1]));
  nn = __fadd_rn(nn, __fmul_rn(v[2], v[2]));
  out[(long long)b * N + n] = make_float4(v[0], v[1], v[2], nn);
}

// Tensor-core operand planes (knn_tc.cu): x = hi + lo + eps with hi = bf16(x), lo = bf16(x - hi), |eps| <= 2^-18 |x|.
__device__ __forceinline__ void bf16_split(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// Norm slice of the tensor-core path: candidate n contributes the K=16 operand row (h1, h2, h3, 0, ..., 0),
// h = -|b|^2/2 split into three bf16 terms (24 bits), stored per (cloud, tile of 128) in the compact no-swizzle operand
// layout [chunk][row][16 B]; only chunk 0 is non-zero.
__device__ __forceinline__ void knn_store_ext(float* ext, int b, int N, int n, float h) {
  const int ntiles = (N + 127) >> 7;
  if (n >= ntiles * 128) return;
  const __nv_bfloat16 h1 = __float2bfloat16_rn(h);
  const float r1 = h - __bfloat162float(h1);
  const __nv_bfloat16 h2 = __float2bfloat16_rn(r1);
  const __nv_bfloat16 h3 = __float2bfloat16_rn(r1 - __bfloat162float(h2));
  float* blk = ext + ((size_t)b * ntiles + (n >> 7)) * 1024;
  uint4 row;
  row.x = (uint32_t)__bfloat16_as_ushort(h1) | ((uint32_t)__bfloat16_as_ushort(h2) << 16);
  row.y = (uint32_t)__bfloat16_as_ushort(h3);
  row.z = 0u, row.w = 0u;
  *reinterpret_cast<uint4*>(blk + (n & 127) * 4) = row;
  *reinterpret_cast<uint4*>(blk + 512 + (n & 127) * 4) = make_uint4(0u, 0u, 0u, 0u);
}

// feature path: point-major normalised copy (B,N,Cp) (channels >= C zero) + squared norms (B,N).
// 32x32 smem transpose so both the channel-major read and the point-major write coalesce.
__global__ void __launch_bounds__(256) knn_prep_feat_kernel(const float* __restrict__ x, long long sb, long long sn,
                                                            long long sc, int N, int C, int Cp,
                                                            const float* __restrict__ mean,
                                                            const float* __restrict__ stdv, float* __restrict__ out,
                                                            float* __restrict__ norms, unsigned* __restrict__ maxnorm,
                                                            float* __restrict__ ext, __nv_bfloat16* __restrict__ out_hi,
                                                            __nv_bfloat16* __restrict__ out_lo) {
  __shared__ float tile[32][33];
  const int b = blockIdx.y, n0 = blockIdx.x * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (n0 >= N) {                                        // padding rows of the last 128-candidate tile (ext only)
    if (ty == 0) knn_store_ext(ext, b, N, n0 + tx, -1e30f);
    return;
  }
  const float sigma = cloud_sigma(stdv + b * C, C);
  float nn = 0.f;
  for (int c0 = 0; c0 < Cp; c0 += 32) {
    if (sc == 1) {               // point-major input: lanes across channels (coalesced), tile[c][n] filled column-wise
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        int c = c0 + tx, n = n0 + ty + 8 * r;
        float v = 0.f;
        if (c < C && n < N) v = __fdiv_rn(__fsub_rn(x[b * sb + n * sn + c], mean[b * C + c]), sigma);
        tile[tx][ty + 8 * r] = v;
      }
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        int c = c0 + ty + 8 * r, n = n0 + tx;
        float v = 0.f;
        if (c < C && n < N) v = __fdiv_rn(__fsub_rn(x[b * sb + n * sn + c * sc], mean[b * C + c]), sigma);
        tile[ty + 8 * r][tx] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int n = n0 + ty + 8 * r, c = c0 + tx;
      if (n < N && c < Cp) {
        const float v = tile[tx][ty + 8 * r];
        out[((long long)b * N + n) * Cp + c] = v;
        if (out_hi) {              // tensor-core operand planes
          __nv_bfloat16 h, l;
          bf16_split(v, h, l);
          out_hi[((long long)b * N + n) * Cp + c] = h;
          out_lo[((long long)b * N + n) * Cp + c] = l;
        }
      }
    }
    if (ty == 0) {
      int cmax = min(32, Cp - c0);
      for (int c = 0; c < cmax; ++c) nn = __fadd_rn(nn, __fmul_rn(tile[c][tx], tile[c][tx]));
    }
    __syncthreads();
  }
  if (ty == 0) {
    if (n0 + tx < N) norms[(long long)b * N + n0 + tx] = nn;
    if (ext) knn_store_ext(ext, b, N, n0 + tx, n0 + tx < N ? -0.5f * nn : -1e30f);   // past N: can never win
    if (maxnorm) {                                    // per-cloud max |p'|^2 (>= 0: uint order == float order)
      const unsigned m = __reduce_max_sync(kFull, n0 + tx < N ? __float_as_uint(nn) : 0u);
      if (tx == 0) atomicMax(maxnorm + b, m);
    }
  }
}

// Point-major input (the blocks' own activations): no transpose to do, so every thread moves one 16-byte chunk per
// 32-channel block -- one LDG.128, one STG.128 (fp32 copy) and two 8-byte stores (bf16 hi / lo planes) -- and only the squared norms go through
// shared memory, to be summed per point in channel order exactly as above.
__global__ void __launch_bounds__(256) knn_prep_feat_pm_kernel(const float* __restrict__ x, long long sb, long long sn, int N,
                                                               int C, int Cp, const float* __restrict__ mean,
                                                               const float* __restrict__ stdv, float* __restrict__ out,
                                                               float* __restrict__ norms, unsigned* __restrict__ maxnorm,
                                                               float* __restrict__ ext, __nv_bfloat16* __restrict__ out_hi,
                                                               __nv_bfloat16* __restrict__ out_lo,
                                                               const double* __restrict__ part, int Z, int Nstat,
                                                               float* __restrict__ mean_out, float* __restrict__ stdv_out) {
  __shared__ float tile[32][33];                        // [channel of the block][point]
  __shared__ __align__(16) float s_mean[512];
  __shared__ float s_std[512];
  __shared__ float s_sigma;
  const int b = blockIdx.y, n0 = blockIdx.x * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (n0 >= N) {
    if (ty == 0) knn_store_ext(ext, b, N, n0 + tx, -1e30f);
    return;
  }
  float sigma;
  if (part) {
    // the statistics arrive as per-slice fp64 partial sums (knn_stats_pm_kernel): every CTA combines them itself, in slice
    // order with the arithmetic of knn_stats_pm_combine_kernel (same bits), instead of waiting for one more launch
    for (int c = threadIdx.x; c < C; c += 256) {
      double s1 = 0.0, s2 = 0.0;
      for (int z = 0; z < Z; ++z) {
        const double* o = part + (((long long)b * Z + z) * C + c) * 2;
        s1 += o[0], s2 += o[1];
      }
      const double m = s1 / Nstat;
      const double var = (s2 - s1 * m) / (double)(Nstat - 1);
      s_mean[c] = (float)m;
      s_std[c] = (float)sqrt(var > 0.0 ? var : 0.0);
      if (blockIdx.x == 0 && mean_out) mean_out[b * C + c] = s_mean[c], stdv_out[b * C + c] = s_std[c];
    }
    __syncthreads();
    if (threadIdx.x == 0) s_sigma = cloud_sigma(s_std, C);
    __syncthreads();
    sigma = s_sigma;
  } else {
    sigma = cloud_sigma(stdv + b * C, C);
  }
  const int row = threadIdx.x >> 3, f4 = threadIdx.x & 7;
  const int n = n0 + row;
  float nn = 0.f;
  for (int c0 = 0; c0 < Cp; c0 += 32) {
    const int c = c0 + f4 * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N && c < C) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + b * sb + n * sn + c));
      const float4 mv = part ? *reinterpret_cast<const float4*>(s_mean + c) : __ldg(reinterpret_cast<const float4*>(mean + b * C + c));
      v.x = __fdiv_rn(__fsub_rn(xv.x, mv.x), sigma);
      v.y = __fdiv_rn(__fsub_rn(xv.y, mv.y), sigma);
      v.z = __fdiv_rn(__fsub_rn(xv.z, mv.z), sigma);
      v.w = __fdiv_rn(__fsub_rn(xv.w, mv.w), sigma);
    }
    if (n < N && c < Cp) {
      const long long o = ((long long)b * N + n) * Cp + c;
      *reinterpret_cast<float4*>(out + o) = v;
      if (out_hi) {
        __nv_bfloat16 h[4], l[4];
        bf16_split(v.x, h[0], l[0]);
        bf16_split(v.y, h[1], l[1]);
        bf16_split(v.z, h[2], l[2]);
        bf16_split(v.w, h[3], l[3]);
        uint2 ph, pl;
        ph.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
        ph.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
        pl.x = (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16);
        pl.y = (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16);
        *reinterpret_cast<uint2*>(out_hi + o) = ph;
        *reinterpret_cast<uint2*>(out_lo + o) = pl;
      }
    }
    tile[f4 * 4 + 0][row] = v.x;
    tile[f4 * 4 + 1][row] = v.y;
    tile[f4 * 4 + 2][row] = v.z;
    tile[f4 * 4 + 3][row] = v.w;
    __syncthreads();
    if (ty == 0) {
      const int cmax = min(32, Cp - c0);
      for (int cc = 0; cc < cmax; ++cc) nn = __fadd_rn(nn, __fmul_rn(tile[cc][tx], tile[cc][tx]));
    }
    __syncthreads();
  }
  if (ty == 0) {
    if (n0 + tx < N) norms[(long long)b * N + n0 + tx] = nn;
    if (ext) knn_store_ext(ext, b, N, n0 + tx, n0 + tx < N ? -0.5f * nn : -1e30f);
    if (maxnorm) {
      const unsigned m = __reduce_max_sync(kFull, n0 + tx < N ? __float_as_uint(nn) : 0u);
      if (tx == 0) atomicMax(maxnorm + b, m);
    }
  }
}

// ------------------------------------------------------------------------------------------
// K1: xyz kNN.  One warp per query (QW queries per warp in flight), lanes across candidates
// staged in shared memory; the running k-set lives one entry per lane (LaneTopK).
// Squared distance follows the 5-term GEMM row of at::_euclidean_dist in k order:
//   ((( (-2x)x' + (-2y)y' ) + (-2z)z' ) + |q|^2 ) + |p|^2 , clamp >= 0
// which reproduces torch.cdist on CPU bit for bit (probed).
// ------------------------------------------------------------------------------------------
constexpr int kXyzChunk = 2048;
constexpr int kXyzQW = 4;

template <class I>
__global__ void __launch_bounds__(256) knn_xyz_kernel(const float4* __restrict__ qp, const float4* __restrict__ rp,
                                                      int Nq, int Nr, int k, I* __restrict__ idx_out,
                                                      float* __restrict__ dist_out) {
  __shared__ float4 cand[kXyzChunk];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, b = blockIdx.y;
  const int qbase = (blockIdx.x * 8 + warp) * kXyzQW;
  float ax[kXyzQW], ay[kXyzQW], az[kXyzQW], aw[kXyzQW];
  LaneTopK t[kXyzQW];
#pragma unroll
  for (int i = 0; i < kXyzQW; ++i) {
    float4 q = qp[(long long)b * Nq + min(qbase + i, Nq - 1)];
    ax[i] = -2.f * q.x, ay[i] = -2.f * q.y, az[i] = -2.f * q.z, aw[i] = q.w;
    t[i].init(lane, k);
  }
  for (int c0 = 0; c0 < Nr; c0 += kXyzChunk) {
    const int n = min(kXyzChunk, Nr - c0);
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) cand[j] = rp[(long long)b * Nr + c0 + j];
    __syncthreads();
    if (qbase >= Nq) continue;
    for (int j0 = 0; j0 < n; j0 += 32) {
      const int j = j0 + lane;
      const bool valid = j < n;
      const float4 p = cand[valid ? j : 0];
#pragma unroll
      for (int i = 0; i < kXyzQW; ++i) {
        float acc = __fmul_rn(ax[i], p.x);
        acc = __fmaf_rn(ay[i], p.y, acc);
        acc = __fmaf_rn(az[i], p.z, acc);
        acc = __fadd_rn(acc, aw[i]);
        acc = __fadd_rn(acc, p.w);
        t[i].offer(dist_bits(acc), c0 + j, valid);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kXyzQW; ++i) {
    const int q = qbase + i;
    if (q >= Nq) break;
    const int r = t[i].rank();
    if (t[i].active) {
      long long o = ((long long)b * Nq + q) * k + r;
      idx_out[o] = (I)t[i].i;
      if (dist_out) dist_out[o] = -sqrtf(__uint_as_float(t[i].d));
    }
  }
}

// ------------------------------------------------------------------------------------------
// K1b: xyz kNN in two passes -- same distances, same answer, ~5x fewer instructions (profiles/r1_knn_xyz.md: K1 spends
// 90% of its issue slots on k-set insertions, ~165 per query at k=32, N=2048).
//   pass 1  each lane keeps the minimum distance of two interleaved candidate groups (64 groups per query); the k-th
//           smallest group minimum T bounds the k-th nearest distance (k distinct candidates reach it)
//   pass 2  distances are recomputed (bit-identical) and the few candidates with d <= T (~45 of 2048) are appended to a
//           per-query list as (distance bits, index) keys
//   final   rank counting over the list: the key with r smaller keys is the r-th neighbour (distance, then index --
//           the order of K1's LaneTopK); a list that overflows sends the whole CTA back to K1's algorithm.
// ------------------------------------------------------------------------------------------
constexpr int kXyzCap = 128;          // list entries per query

// the 5-term GEMM row of torch.cdist in k order, before the clamp at zero
__device__ __forceinline__ float xyz_dist_raw(float ax, float ay, float az, float aw, const float4 p) {
  float acc = __fmul_rn(ax, p.x);
  acc = __fmaf_rn(ay, p.y, acc);
  acc = __fmaf_rn(az, p.z, acc);
  acc = __fadd_rn(acc, aw);
  acc = __fadd_rn(acc, p.w);
  return acc;
}
__device__ __forceinline__ unsigned xyz_dist_bits(float ax, float ay, float az, float aw, const float4 p) {
  return dist_bits(xyz_dist_raw(ax, ay, az, aw, p));
}

size_t knn_xyz2_smem() { return (size_t)kXyzChunk * sizeof(float4) + (size_t)8 * kXyzQW * kXyzCap * sizeof(unsigned long long); }

template <class I>
__global__ void __launch_bounds__(256) knn_xyz2_kernel(const float4* __restrict__ qp, const float4* __restrict__ rp,
                                                       int Nq, int Nr, int k, I* __restrict__ idx_out,
                                                       float* __restrict__ dist_out) {
  extern __shared__ __align__(16) unsigned char xyz_sm[];
  float4* cand = reinterpret_cast<float4*>(xyz_sm);
  unsigned long long* lists = reinterpret_cast<unsigned long long*>(cand + kXyzChunk);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, b = blockIdx.y;
  const int qbase = (blockIdx.x * 8 + warp) * kXyzQW;
  const bool warp_live = qbase < Nq;
  unsigned long long* mylist = lists + (size_t)warp * kXyzQW * kXyzCap;
  float ax[kXyzQW], ay[kXyzQW], az[kXyzQW], aw[kXyzQW];
#pragma unroll
  for (int i = 0; i < kXyzQW; ++i) {
    float4 q = qp[(long long)b * Nq + min(qbase + i, Nq - 1)];
    ax[i] = -2.f * q.x, ay[i] = -2.f * q.y, az[i] = -2.f * q.z, aw[i] = q.w;
  }
  // ---- pass 1: 64 group minima per query ----
  // (round 2, ncu: the kernel issues on the ALU pipe 57 % / FMA pipe 30 % of the time -- the clamp, validity selects and
  // integer minima cost more than the five float operations of a distance.  Minima are now taken on the RAW
  // accumulators with one FMNMX and clamped once at the end -- min_j max(x_j, 0) = max(min_j x_j, 0) -- and the tail of
  // a chunk is padded with sentinel candidates at distance NaN instead of per-pair validity selects.)
  float m0[kXyzQW], m1[kXyzQW];
#pragma unroll
  for (int i = 0; i < kXyzQW; ++i) m0[i] = m1[i] = INFINITY;
  // NaN distance: ignored by fminf in pass 1, never <= T in pass 2 (even if T is +inf)
  const float4 sentinel = make_float4(0.f, 0.f, 0.f, __int_as_float(0x7fc00000));
  for (int c0 = 0; c0 < Nr; c0 += kXyzChunk) {
    const int n = min(kXyzChunk, Nr - c0), npad = (n + 63) & ~63;
    __syncthreads();
    for (int j = threadIdx.x; j < npad; j += blockDim.x) cand[j] = j < n ? rp[(long long)b * Nr + c0 + j] : sentinel;
    __syncthreads();
    if (!warp_live) continue;
    for (int j0 = 0; j0 < npad; j0 += 64) {
      const float4 pa = cand[j0 + lane], pb = cand[j0 + 32 + lane];
#pragma unroll
      for (int i = 0; i < kXyzQW; ++i) {
        m0[i] = fminf(m0[i], xyz_dist_raw(ax[i], ay[i], az[i], aw[i], pa));
        m1[i] = fminf(m1[i], xyz_dist_raw(ax[i], ay[i], az[i], aw[i], pb));
      }
    }
  }
  // ---- thresholds: k-th smallest of the 64 minima (bitonic sort, element p = u*32 + lane) ----
  unsigned T[kXyzQW];
#pragma unroll
  for (int i = 0; i < kXyzQW; ++i) {
    unsigned x[2] = {dist_bits(m0[i]), dist_bits(m1[i])};
#pragma unroll
    for (int size = 2; size <= 64; size <<= 1) {
#pragma unroll
      for (int stride = size >> 1; stride >= 1; stride >>= 1) {
        if (stride == 32) {
          const unsigned lo = min(x[0], x[1]), hi = max(x[0], x[1]);
          x[0] = lo, x[1] = hi;
        } else {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const unsigned other = __shfl_xor_sync(kFull, x[u], stride);
            const bool up = ((u * 32 + lane) & size) == 0;
            const bool lower = (lane & stride) == 0;
            x[u] = (lower == up) ? min(x[u], other) : max(x[u], other);
          }
        }
      }
    }
    T[i] = __shfl_sync(kFull, x[0], k - 1);             // k <= 32: the k smallest sit in x[0]
  }
  // ---- pass 2: collect d <= T ----
  int cnt[kXyzQW];
  float Tf[kXyzQW];                                     // T >= 0, so  max(x, 0) <= T  <=>  x <= T  on the raw accumulator
#pragma unroll
  for (int i = 0; i < kXyzQW; ++i) cnt[i] = 0, Tf[i] = __uint_as_float(T[i]);
  const unsigned lt = (1u << lane) - 1u;
  for (int c0 = 0; c0 < Nr; c0 += kXyzChunk) {
    const int n = min(kXyzChunk, Nr - c0), npad = (n + 63) & ~63;
    if (Nr > kXyzChunk) {                               // single-chunk clouds keep pass 1's copy
      __syncthreads();
      for (int j = threadIdx.x; j < npad; j += blockDim.x) cand[j] = j < n ? rp[(long long)b * Nr + c0 + j] : sentinel;
      __syncthreads();
    }
    if (!warp_live) continue;
    for (int j0 = 0; j0 < npad; j0 += 32) {
      const int j = j0 + lane;
      const float4 p = cand[j];
#pragma unroll
      for (int i = 0; i < kXyzQW; ++i) {
        const float raw = xyz_dist_raw(ax[i], ay[i], az[i], aw[i], p);
        const bool hit = raw <= Tf[i];                  // (false for the NaN sentinels)
        const unsigned m = __ballot_sync(kFull, hit);
        if (m) {
          const int pos = cnt[i] + __popc(m & lt);
          if (hit && pos < kXyzCap) mylist[i * kXyzCap + pos] = ((unsigned long long)dist_bits(raw) << 32) | (unsigned)(c0 + j);
          cnt[i] += __popc(m);
        }
      }
    }
  }
  // ---- final: rank counting; overflow -> the whole CTA redoes its queries with the k-set algorithm ----
  bool over = false;
#pragma unroll
  for (int i = 0; i < kXyzQW; ++i) over |= warp_live && qbase + i < Nq && cnt[i] > kXyzCap;
  if (__syncthreads_or(over ? 1 : 0)) {
    LaneTopK t[kXyzQW];
#pragma unroll
    for (int i = 0; i < kXyzQW; ++i) t[i].init(lane, k);
    for (int c0 = 0; c0 < Nr; c0 += kXyzChunk) {
      const int n = min(kXyzChunk, Nr - c0);
      __syncthreads();
      for (int j = threadIdx.x; j < n; j += blockDim.x) cand[j] = rp[(long long)b * Nr + c0 + j];
      __syncthreads();
      if (!warp_live) continue;
      for (int j0 = 0; j0 < n; j0 += 32) {
        const int j = j0 + lane;
        const bool valid = j < n;
        const float4 p = cand[valid ? j : 0];
#pragma unroll
        for (int i = 0; i < kXyzQW; ++i) t[i].offer(xyz_dist_bits(ax[i], ay[i], az[i], aw[i], p), c0 + j, valid);
      }
    }
#pragma unroll
    for (int i = 0; i < kXyzQW; ++i) {
      const int q = qbase + i;
      if (q >= Nq) break;
      const int r = t[i].rank();
      if (t[i].active) {
        long long o = ((long long)b * Nq + q) * k + r;
        idx_out[o] = (I)t[i].i;
        if (dist_out) dist_out[o] = -sqrtf(__uint_as_float(t[i].d));
      }
    }
    return;
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < kXyzQW; ++i) {
    const int q = qbase + i;
    if (q >= Nq) break;
    const unsigned long long* L = mylist + i * kXyzCap;
    const int c = cnt[i];
    for (int t = lane; t < c; t += 32) {
      const unsigned long long mine = L[t];
      int rank = 0;
      for (int u = 0; u < c; ++u) rank += L[u] < mine ? 1 : 0;
      if (rank < k) {
        const long long o = ((long long)b * Nq + q) * k + rank;
        idx_out[o] = (I)(unsigned)(mine & 0xffffffffu);
        if (dist_out) dist_out[o] = -sqrtf(__uint_as_float((unsigned)(mine >> 32)));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// K2 (exact fp32 form): feature-space kNN.  FFMA dot tiles + per-row k-selection epilogue.
//   d2 = fma(-2, <a,b>, |a|^2 + |b|^2), clamp >= 0.
// ------------------------------------------------------------------------------------------
template <class Cfg>
struct KnnEpilogue {
  unsigned* listd;      // smem [TQ][32]
  int* listi;           // smem [TQ][32]
  const float* qq;      // smem [TQ]
  const float* bnorm;   // global, this cloud's candidate norms
  int q0, Nq, Nr, k;

  __device__ __forceinline__ void tile(const float* S, int ldS, int n0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float bb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int j = n0 + lane + 32 * u;
      bb[u] = j < Nr ? bnorm[j] : 0.f;
    }
    for (int rr = 0; rr < Cfg::RPW; ++rr) {
      const int row = warp * Cfg::RPW + rr;
      if (q0 + row >= Nq) break;
      LaneTopK t;
      t.load(lane, k, listd[row * 32 + lane], listi[row * 32 + lane]);
      const float aa = qq[row];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = n0 + lane + 32 * u;
        const float dot = S[row * ldS + lane + 32 * u];
        const float d2 = __fmaf_rn(-2.f, dot, __fadd_rn(aa, bb[u]));
        t.offer(dist_bits(d2), j, j < Nr);
      }
      listd[row * 32 + lane] = t.d;
      listi[row * 32 + lane] = t.i;
    }
  }
};

template <class Cfg, class I>
__global__ void __launch_bounds__(256, Cfg::MQ == 4 ? 2 : 1)
    knn_feat_kernel(const float* __restrict__ an, const float* __restrict__ anorm, const float* __restrict__ bn,
                    const float* __restrict__ bnorm, int Nq, int Nr, int C, int Cp, int k, I* __restrict__ idx_out,
                    float* __restrict__ dist_out, const int* __restrict__ row_flags) {
  extern __shared__ __align__(16) float smem[];
  const int b = blockIdx.y, q0 = blockIdx.x * Cfg::TQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (row_flags) {   // repair mode: only tiles holding a row the tensor-core path could not finish
    if (!row_flags[(long long)gridDim.y * Nq]) return;          // nothing flagged anywhere: the usual case
    int any = 0;
    for (int r = threadIdx.x; r < Cfg::TQ; r += blockDim.x) any |= (q0 + r < Nq) ? row_flags[(long long)b * Nq + q0 + r] : 0;
    if (!__syncthreads_or(any)) return;
  }
  float* tail = smem + Cfg::smem_floats(Cp);
  unsigned* listd = reinterpret_cast<unsigned*>(tail);
  int* listi = reinterpret_cast<int*>(tail + Cfg::TQ * 32);
  float* qq = tail + 2 * Cfg::TQ * 32;
  for (int r = warp; r < Cfg::TQ; r += 8) {
    LaneTopK t;
    t.init(lane, k);
    listd[r * 32 + lane] = t.d;
    listi[r * 32 + lane] = t.i;
  }
  for (int r = threadIdx.x; r < Cfg::TQ; r += blockDim.x) qq[r] = (q0 + r) < Nq ? anorm[(long long)b * Nq + q0 + r] : 0.f;
  __syncthreads();
  KnnEpilogue<Cfg> epi{listd, listi, qq, bnorm + (long long)b * Nr, q0, Nq, Nr, k};
  dot_tiles<Cfg>(an + (long long)b * Nq * Cp, Cp, q0, Nq, bn + (long long)b * Nr * Cp, Cp, Nr, Cp, Cp, smem, epi);
  __syncthreads();
  for (int rr = 0; rr < Cfg::RPW; ++rr) {
    const int row = warp * Cfg::RPW + rr, q = q0 + row;
    if (q >= Nq) break;
    if (row_flags && !row_flags[(long long)b * Nq + q]) continue;
    LaneTopK t;
    t.load(lane, k, listd[row * 32 + lane], listi[row * 32 + lane]);
    const int r = t.rank();
    if (t.active) {
      long long o = ((long long)b * Nq + q) * k + r;
      idx_out[o] = (I)t.i;
      if (dist_out) dist_out[o] = -sqrtf(__uint_as_float(t.d));
    }
  }
}

template <class Cfg>
static size_t knn_feat_smem(int Cp) {
  return (Cfg::smem_floats(Cp) + 2 * Cfg::TQ * 32 + Cfg::TQ) * sizeof(float);
}

template <class Cfg, class I>
static int launch_knn_feat(const float* an, const float* anorm, const float* bn, const float* bnorm, int B, int Nq,
                           int Nr, int C, int Cp, int k, I* idx, float* dist, const int* row_flags, cudaStream_t st) {
  size_t smem = knn_feat_smem<Cfg>(Cp);
  auto kern = knn_feat_kernel<Cfg, I>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return check_launch("knn_feat smem attribute");
  dim3 grid(ceil_div(Nq, Cfg::TQ), B);
  SAMBLE_PRE(st);
  kern<<<grid, 256, smem, st>>>(an, anorm, bn, bnorm, Nq, Nr, C, Cp, k, idx, dist, row_flags);
  SAMBLE_LAUNCHED(row_flags ? "knn_feat_repair_kernel" : "knn_feat_kernel");
  return SAMBLE_OK;
}

// shared with upsample.cu
constexpr int kStatsSlices = 4;     // point slices per (cloud, 32-channel block) in the point-major form
size_t knn_stats_scratch_bytes(int B, int C) { return (size_t)B * kStatsSlices * C * 2 * sizeof(double); }

bool knn_stats_point_major(long long sc, int C) { return sc == 1 && C >= 8; }
static int launch_knn_stats_ex(const float* a, long long sb, long long sn, long long sc, int B, int N, int C, float* mean,
                               float* stdv, cudaStream_t st, double* scratch, bool defer_combine, unsigned* zero_me) {
  if (knn_stats_point_major(sc, C) && scratch) {     // point-major rows (the blocks' own activations)
    SAMBLE_PRE(st);
    knn_stats_pm_kernel<<<dim3(ceil_div(C, 32), B, kStatsSlices), 512, 0, st>>>(a, sb, sn, N, C, scratch, zero_me);
    SAMBLE_LAUNCHED("knn_stats_kernel");
    if (defer_combine) return SAMBLE_OK;    // the consumer (knn_prep_feat_pm_kernel) combines the slices itself
    SAMBLE_PRE(st);
    knn_stats_pm_combine_kernel<<<dim3(ceil_div(C, 128), B), 128, 0, st>>>(scratch, kStatsSlices, N, C, mean, stdv);
    SAMBLE_LAUNCHED("knn_stats_combine_kernel");
    return SAMBLE_OK;
  }
  SAMBLE_PRE(st);
  knn_stats_kernel<<<dim3(ceil_div(C, 8), B), 256, 0, st>>>(a, sb, sn, sc, N, C, mean, stdv);
  SAMBLE_LAUNCHED("knn_stats_kernel");
  return SAMBLE_OK;
}
int launch_knn_stats(const float* a, long long sb, long long sn, long long sc, int B, int N, int C, float* mean,
                     float* stdv, cudaStream_t st, double* scratch) {
  return launch_knn_stats_ex(a, sb, sn, sc, B, N, C, mean, stdv, st, scratch, false, nullptr);
}
int launch_knn_prep_xyz(const float* x, long long sb, long long sn, long long sc, int B, int N, int C,
                        const float* mean, const float* stdv, float4* out, cudaStream_t st) {
  SAMBLE_PRE(st);
  knn_prep_xyz_kernel<<<dim3(ceil_div(N, 256), B), 256, 0, st>>>(x, sb, sn, sc, N, C, mean, stdv, out);
  SAMBLE_LAUNCHED("knn_prep_xyz_kernel");
  return SAMBLE_OK;
}

// tensor-core path (knn_tc.cu)
bool knn_tc_eligible(int Nq, int Nr, int C, int k);
size_t knn_tc_workspace_bytes(int B, int Nq, int Nr);
size_t knn_tc_ext_floats(int B, int Nr);
template <class I>
int launch_knn_tc(const float* an, const float* anorm, const float* bn, const float* bnorm, const __nv_bfloat16* a_hi,
                  const __nv_bfloat16* a_lo, const __nv_bfloat16* b_hi, const __nv_bfloat16* b_lo, const float* bext,
                  const unsigned* bbmax, int B, int Nq, int Nr, int Cp, int k, float* thr, uint32_t* cand, int* cnt, bool ordered,
                  I* idx, float* dist, int* row_flags, cudaStream_t st);

static int g_knn_mode = 0;   // 0 auto, 1 exact FFMA kernel only, 2 tensor-core path wherever eligible

struct KnnPlan {
  int Cp;
  bool xyz;
  size_t bytes;
};
static KnnPlan knn_plan(int B, int Nq, int Nr, int C) {
  KnnPlan p;
  p.xyz = C <= 3;
  p.Cp = p.xyz ? 4 : (int)align_up(C, 32);   // 32-channel K-blocks for the tensor-core path
  size_t per_pt = p.xyz ? sizeof(float4) : (size_t)(p.Cp + 1) * sizeof(float);
  p.bytes = 2 * align_up((size_t)B * C * sizeof(float), 256) + align_up((size_t)B * Nq * per_pt, 256) +
            align_up((size_t)B * Nr * per_pt, 256) + 2 * align_up((size_t)B * Nq * sizeof(float), 256) +
            align_up((size_t)B * sizeof(unsigned), 256) + (p.xyz ? 0 : knn_tc_workspace_bytes(B, Nq, Nr)) +
            (p.xyz ? 0 : align_up((size_t)B * Nq * per_pt, 256) + align_up((size_t)B * Nr * per_pt, 256) + 4 * 256) /* bf16 planes */ +
            align_up((size_t)B * 4 * C * 2 * sizeof(double), 256) /* stats slices */ + 14 * 256;
  return p;
}

static bool prep_point_major(const float* x, long long sb, long long sn, long long sc, int C) {
  return sc == 1 && C % 4 == 0 && sn % 4 == 0 && sb % 4 == 0 && (uintptr_t)x % 16 == 0;
}
static void launch_prep_feat(dim3 grid, cudaStream_t st, const float* x, long long sb, long long sn, long long sc, int N, int C,
                             int Cp, const float* mean, const float* stdv, float* out, float* norms, unsigned* maxnorm,
                             float* ext, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, const double* part = nullptr, int Nstat = 0) {
  const bool pm = prep_point_major(x, sb, sn, sc, C);
  if (pm)
    knn_prep_feat_pm_kernel<<<grid, 256, 0, st>>>(x, sb, sn, N, C, Cp, mean, stdv, out, norms, maxnorm, ext, out_hi, out_lo, part,
                                                  kStatsSlices, Nstat, const_cast<float*>(mean), const_cast<float*>(stdv));
  else
    knn_prep_feat_kernel<<<grid, 256, 0, st>>>(x, sb, sn, sc, N, C, Cp, mean, stdv, out, norms, maxnorm, ext, out_hi, out_lo);
}

template <class I>
static int knn_impl(const float* a, long long a_sb, long long a_sn, long long a_sc, const float* b, long long b_sb,
                    long long b_sn, long long b_sc, int B, int Nq, int Nr, int C, int k, I* idx_out, float* dist_out,
                    bool ordered, void* ws, size_t ws_bytes, cudaStream_t st) {
  KnnPlan plan = knn_plan(B, Nq, Nr, C);
  SAMBLE_REQUIRE(ws_bytes >= plan.bytes, "samble_knn: workspace %zu < %zu bytes", ws_bytes, plan.bytes);
  const bool self = (a == b) && a_sb == b_sb && a_sn == b_sn && a_sc == b_sc && Nq == Nr;
  Workspace w(ws, ws_bytes);
  float* mean = w.take<float>((size_t)B * C);
  float* stdv = w.take<float>((size_t)B * C);
  double* stats_scratch = w.take<double>(knn_stats_scratch_bytes(B, C) / sizeof(double));
  unsigned* bbmax = w.take<unsigned>((size_t)B);
  // feature path on point-major activations: the prep kernels combine the statistics slices themselves (one launch less on the
  // critical path of every neighbour search)
  const bool fuse_stats = !plan.xyz && C <= 512 && knn_stats_point_major(a_sc, C) && prep_point_major(a, a_sb, a_sn, a_sc, C) &&
                          (self || prep_point_major(b, b_sb, b_sn, b_sc, C));
  if (int e = launch_knn_stats_ex(a, a_sb, a_sn, a_sc, B, Nq, C, mean, stdv, st, stats_scratch, fuse_stats, fuse_stats ? bbmax : nullptr)) return e;
  const double* part = fuse_stats ? stats_scratch : nullptr;
  if (plan.xyz) {
    float4* qa = w.take<float4>((size_t)B * Nq);
    float4* qb = self ? qa : w.take<float4>((size_t)B * Nr);
    if (int e = launch_knn_prep_xyz(a, a_sb, a_sn, a_sc, B, Nq, C, mean, stdv, qa, st)) return e;
    if (!self)
      if (int e = launch_knn_prep_xyz(b, b_sb, b_sn, b_sc, B, Nr, C, mean, stdv, qb, st)) return e;
    dim3 grid(ceil_div(Nq, 8 * kXyzQW), B);
    if (g_knn_mode != 1 && Nr >= 256) {        // two-pass form; small clouds gain nothing from the threshold
      auto kern = knn_xyz2_kernel<I>;
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)knn_xyz2_smem()) != cudaSuccess)
        return check_launch("knn_xyz2 smem attribute");
      SAMBLE_PRE(st);
      kern<<<grid, 256, knn_xyz2_smem(), st>>>(qa, qb, Nq, Nr, k, idx_out, dist_out);
      SAMBLE_LAUNCHED("knn_xyz2_kernel");
      return SAMBLE_OK;
    }
    SAMBLE_PRE(st);
    knn_xyz_kernel<I><<<grid, 256, 0, st>>>(qa, qb, Nq, Nr, k, idx_out, dist_out);
    SAMBLE_LAUNCHED("knn_xyz_kernel");
    return SAMBLE_OK;
  }
  const int Cp = plan.Cp;
  float* an = w.take<float>((size_t)B * Nq * Cp);
  float* anorm = w.take<float>((size_t)B * Nq);
  float* bn = self ? an : w.take<float>((size_t)B * Nr * Cp);
  float* bnorm = self ? anorm : w.take<float>((size_t)B * Nr);
  SAMBLE_PRE(st);
  float* thr = w.take<float>((size_t)B * Nq);
  int* row_flags = w.take<int>((size_t)B * Nq + 1);          // + one "any row flagged" word
  uint32_t* cand = w.take<uint32_t>((size_t)B * Nq * 128);
  int* cand_cnt = w.take<int>((size_t)B * Nq);
  float* bext = w.take<float>(knn_tc_ext_floats(B, Nr));
  // bf16 hi/lo operand planes for the tensor-core passes (together the bytes of one fp32 copy)
  __nv_bfloat16* a_hi = w.take<__nv_bfloat16>((size_t)B * Nq * Cp);
  __nv_bfloat16* a_lo = w.take<__nv_bfloat16>((size_t)B * Nq * Cp);
  __nv_bfloat16* b_hi = self ? a_hi : w.take<__nv_bfloat16>((size_t)B * Nr * Cp);
  __nv_bfloat16* b_lo = self ? a_lo : w.take<__nv_bfloat16>((size_t)B * Nr * Cp);
  const bool use_tc = g_knn_mode != 1 && knn_tc_eligible(Nq, Nr, C, k);
  if (use_tc && !fuse_stats) {   // per-cloud maximum candidate norm: an atomicMax target of the prep kernel (else cleared by the stats kernel)
    if (cudaMemsetAsync(bbmax, 0, (size_t)B * sizeof(unsigned), st) != cudaSuccess) return check_launch("memset knn bbmax");
    count_launch();
  }
  SAMBLE_PRE(st);
  // the candidate-side launch also covers the padding rows of the last 128-candidate tile (norm slice only)
  const bool a_is_cand = use_tc && self;
  launch_prep_feat(dim3(ceil_div(a_is_cand ? (int)align_up(Nq, 128) : Nq, 32), B), st, a, a_sb, a_sn, a_sc, Nq, C, Cp, mean, stdv,
                   an, anorm, a_is_cand ? bbmax : nullptr, a_is_cand ? bext : nullptr, use_tc ? a_hi : nullptr, use_tc ? a_lo : nullptr,
                   part, Nq);
  SAMBLE_LAUNCHED("knn_prep_feat_kernel");
  if (!self) {
    SAMBLE_PRE(st);
    launch_prep_feat(dim3(ceil_div(use_tc ? (int)align_up(Nr, 128) : Nr, 32), B), st, b, b_sb, b_sn, b_sc, Nr, C, Cp, mean, stdv, bn,
                     bnorm, use_tc ? bbmax : nullptr, use_tc ? bext : nullptr, use_tc ? b_hi : nullptr, use_tc ? b_lo : nullptr, part, Nq);
    SAMBLE_LAUNCHED("knn_prep_feat_kernel");
  }
  if (use_tc) {
    if (int e = launch_knn_tc<I>(an, anorm, bn, bnorm, a_hi, a_lo, b_hi, b_lo, bext, bbmax, B, Nq, Nr, Cp, k, thr, cand, cand_cnt, ordered, idx_out,
                                 dist_out, row_flags, st))
      return e;
    // (rows whose candidate list saturated were searched exactly by their own warp inside knn_select / knn_rerank)
    return SAMBLE_OK;
  }
  // 128-row CTAs when they still give every SM at least ~2 CTAs of work, else 64-row CTAs
  const bool big = (long long)ceil_div(Nq, 128) * B >= 2 * 148 && knn_feat_smem<DotTileCfg<8>>(Cp) <= 200 * 1024;
  if (big) return launch_knn_feat<DotTileCfg<8>, I>(an, anorm, bn, bnorm, B, Nq, Nr, C, Cp, k, idx_out, dist_out, nullptr, st);
  return launch_knn_feat<DotTileCfg<4>, I>(an, anorm, bn, bnorm, B, Nq, Nr, C, Cp, k, idx_out, dist_out, nullptr, st);
}

}  // namespace samble

using namespace samble;

extern "C" void samble_set_knn_mode(int mode) { g_knn_mode = mode; }

extern "C" size_t samble_knn_workspace_bytes(int B, int Nq, int Nr, int C) {
  if (B <= 0 || Nq <= 0 || Nr <= 0 || C <= 0) return 0;
  return knn_plan(B, Nq, Nr, C).bytes;
}

extern "C" int samble_knn(const float* a, long long a_sb, long long a_sn, long long a_sc, const float* b,
                          long long b_sb, long long b_sn, long long b_sc, int B, int Nq, int Nr, int C, int k,
                          void* idx_out, int idx_bits, float* dist_out, int flags, void* ws, size_t ws_bytes,
                          samble_stream_t stream) {
  SAMBLE_REQUIRE(a && b && idx_out && ws, "samble_knn: null pointer");
  SAMBLE_REQUIRE(B > 0 && Nq > 0 && Nr > 0 && C > 0, "samble_knn: empty shape B=%d Nq=%d Nr=%d C=%d", B, Nq, Nr, C);
  SAMBLE_REQUIRE(C <= 512, "samble_knn: C=%d > 512", C);
  SAMBLE_REQUIRE(k >= 1 && k <= 32, "samble_knn: k=%d outside [1,32]", k);
  SAMBLE_REQUIRE(k <= Nr, "samble_knn: k=%d > number of candidates %d", k, Nr);
  SAMBLE_REQUIRE(idx_bits == 32 || idx_bits == 64, "samble_knn: idx_bits must be 32 or 64");
  SAMBLE_REQUIRE(B <= 65535, "samble_knn: B=%d > 65535", B);
  SAMBLE_REQUIRE((flags & ~SAMBLE_KNN_ANY_ORDER) == 0, "samble_knn: unknown flags 0x%x", flags);
  SAMBLE_REQUIRE(!(flags & SAMBLE_KNN_ANY_ORDER) || !dist_out, "samble_knn: SAMBLE_KNN_ANY_ORDER returns indices only");
  const bool ordered = !(flags & SAMBLE_KNN_ANY_ORDER);
  cudaStream_t st = (cudaStream_t)stream;
  if (idx_bits == 64)
    return knn_impl<long long>(a, a_sb, a_sn, a_sc, b, b_sb, b_sn, b_sc, B, Nq, Nr, C, k, (long long*)idx_out,
                               dist_out, ordered, ws, ws_bytes, st);
  return knn_impl<int>(a, a_sb, a_sn, a_sc, b, b_sb, b_sn, b_sc, B, Nq, Nr, C, k, (int*)idx_out, dist_out, ordered,
                       ws, ws_bytes, st);
}
