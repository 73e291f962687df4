// UpSampleInterpolation core (reference models/upsample.py:194-212 + utils/ops.py:68-80):
// 3-NN of every "up" point among the selected points (xyz, normalised by the up cloud's statistics),
// inverse-distance weights, weighted sum of the selected features -- one fused kernel, no (B,N,M)
// distance matrix and no (B,C,N,3) gathered tensor.  HBM-bound: reads C*M + 3(N+M), writes C*N floats.
#include "common.cuh"

namespace samble {

int launch_knn_stats(const float* a, long long sb, long long sn, long long sc, int B, int N, int C, float* mean,
                     float* stdv, cudaStream_t st, double* scratch);
int launch_knn_prep_xyz(const float* x, long long sb, long long sn, long long sc, int B, int N, int C,
                        const float* mean, const float* stdv, float4* out, cudaStream_t st);

constexpr int kUpChunk = 2048;

// Four threads per up-point (round 2: one thread per point left 2 warps per scheduler scanning 1024 candidates each through a
// dependent compare chain -- 10 % of the warp slots busy, profiles/r2_step_full.md).  Thread `sub` of a point scans the
// candidates j = sub (mod 4) in ascending order into its own top 3; the four triples are merged by (distance, index),
// which is exactly the order the sequential scan's strict `<` produces (ties keep the lower index).
struct Top3 {
  float d0, d1, d2;
  int i0, i1, i2;
};
__device__ __forceinline__ bool before(float da, int ia, float db, int ib) { return da < db || (da == db && ia < ib); }
__device__ __forceinline__ void top3_insert(Top3& t, float dd, int ii) {           // general insert by (distance, index)
  if (before(dd, ii, t.d2, t.i2)) {
    if (before(dd, ii, t.d1, t.i1)) {
      t.d2 = t.d1, t.i2 = t.i1;
      if (before(dd, ii, t.d0, t.i0)) t.d1 = t.d0, t.i1 = t.i0, t.d0 = dd, t.i0 = ii;
      else t.d1 = dd, t.i1 = ii;
    } else {
      t.d2 = dd, t.i2 = ii;
    }
  }
}

__global__ void __launch_bounds__(128) interpolate3_kernel(const float4* __restrict__ up, const float4* __restrict__ sel,
                                                           const float* __restrict__ feat, int N, int M, int C,
                                                           float* __restrict__ out, long long* __restrict__ idx_out,
                                                           float* __restrict__ dist_out, int* __restrict__ nn_idx,
                                                           float* __restrict__ nn_w) {
  __shared__ float4 cand[kUpChunk];
  const int sub = threadIdx.x & 3;
  const int b = blockIdx.y, n = blockIdx.x * 32 + (threadIdx.x >> 2);
  const bool live = n < N;
  const float4 q = up[(long long)b * N + (live ? n : N - 1)];
  const float ax = -2.f * q.x, ay = -2.f * q.y, az = -2.f * q.z;
  Top3 t{INFINITY, INFINITY, INFINITY, 0x7ffffffd, 0x7ffffffe, 0x7fffffff};
  for (int c0 = 0; c0 < M; c0 += kUpChunk) {
    const int cn = min(kUpChunk, M - c0);
    __syncthreads();
    for (int j = threadIdx.x; j < cn; j += blockDim.x) cand[j] = sel[(long long)b * M + c0 + j];
    __syncthreads();
    // four candidates of this thread's residue class per trip: independent distance chains, inserts in ascending j
    for (int j = sub; j < cn; j += 16) {
      float d[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 p = cand[min(j + 4 * u, cn - 1)];
        // same 5-term row as the xyz kNN (knn.cu)
        float acc = __fmul_rn(ax, p.x);
        acc = __fmaf_rn(ay, p.y, acc);
        acc = __fmaf_rn(az, p.z, acc);
        acc = __fadd_rn(acc, q.w);
        acc = __fadd_rn(acc, p.w);
        d[u] = (j + 4 * u < cn) ? (acc > 0.f ? acc : 0.f) : INFINITY;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float dd = d[u];
        if (dd < t.d2) {                      // (ascending j within a thread: strict < keeps the lower index on ties)
          const int ii = c0 + j + 4 * u;
          if (dd < t.d1) {
            t.d2 = t.d1, t.i2 = t.i1;
            if (dd < t.d0) t.d1 = t.d0, t.i1 = t.i0, t.d0 = dd, t.i0 = ii;
            else t.d1 = dd, t.i1 = ii;
          } else {
            t.d2 = dd, t.i2 = ii;
          }
        }
      }
    }
  }
  // merge the four partial triples (lanes sub = 0..3 of the point) -- after two exchange rounds every lane holds the result
#pragma unroll
  for (int w = 1; w <= 2; w <<= 1) {
    const float od0 = __shfl_xor_sync(kFull, t.d0, w), od1 = __shfl_xor_sync(kFull, t.d1, w), od2 = __shfl_xor_sync(kFull, t.d2, w);
    const int oi0 = __shfl_xor_sync(kFull, t.i0, w), oi1 = __shfl_xor_sync(kFull, t.i1, w), oi2 = __shfl_xor_sync(kFull, t.i2, w);
    top3_insert(t, od0, oi0);
    top3_insert(t, od1, oi1);
    top3_insert(t, od2, oi2);
  }
  const float d0 = t.d0, d1 = t.d1, d2 = t.d2;
  const int i0 = t.i0, i1 = t.i1, i2 = t.i2;
  if (!live) return;
  const float e0 = sqrtf(d0), e1 = sqrtf(d1), e2 = sqrtf(d2);
  // weights = 1/(d+1e-8), normalised by their left-to-right sum (upsample.py:206-209)
  float w0 = __fdiv_rn(1.0f, __fadd_rn(e0, 1e-8f));
  float w1 = __fdiv_rn(1.0f, __fadd_rn(e1, 1e-8f));
  float w2 = __fdiv_rn(1.0f, __fadd_rn(e2, 1e-8f));
  const float ws = __fadd_rn(__fadd_rn(w0, w1), w2);
  w0 = __fdiv_rn(w0, ws), w1 = __fdiv_rn(w1, ws), w2 = __fdiv_rn(w2, ws);
  if (idx_out && sub == 0) {
    long long* o = idx_out + ((long long)b * N + n) * 3;
    o[0] = i0, o[1] = i1, o[2] = i2;
  }
  if (dist_out && sub == 0) {
    float* o = dist_out + ((long long)b * N + n) * 3;
    o[0] = e0, o[1] = e1, o[2] = e2;
  }
  if (nn_idx && sub == 0) {                           // point-major form: the gather runs in interp3_rows_kernel
    int* o = nn_idx + ((long long)b * N + n) * 3;
    float* ow = nn_w + ((long long)b * N + n) * 3;
    o[0] = i0, o[1] = i1, o[2] = i2;
    ow[0] = w0, ow[1] = w1, ow[2] = w2;
  }
  if (!feat) return;
  const float* f = feat + (long long)b * C * M;
  float* y = out + (long long)b * C * N + n;
  for (int c = sub; c < C; c += 4) {                  // the point's four threads share its channels
    const float* fr = f + (long long)c * M;
    // sum over the 3 neighbours of (feature * weight), products rounded separately (upsample.py:210-212)
    float v = __fmul_rn(__ldg(fr + i0), w0);
    v = __fadd_rn(v, __fmul_rn(__ldg(fr + i1), w1));
    v = __fadd_rn(v, __fmul_rn(__ldg(fr + i2), w2));
    y[(long long)c * N] = v;
  }
}

// point-major features: warp per up-point, lanes across 16-byte channel chunks -- three coalesced row reads, one
// coalesced row write (possibly into a column slice of a wider buffer: ld_out), same per-element arithmetic
__global__ void __launch_bounds__(256) interp3_rows_kernel(const int* __restrict__ nn_idx, const float* __restrict__ nn_w,
                                                           const float* __restrict__ feat, long long ld_feat, int N, int M,
                                                           int C, float* __restrict__ out, long long ld_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, n = blockIdx.x * 8 + warp;
  if (n >= N) return;
  const long long p = (long long)b * N + n;
  const int i0 = nn_idx[p * 3], i1 = nn_idx[p * 3 + 1], i2 = nn_idx[p * 3 + 2];
  const float w0 = nn_w[p * 3], w1 = nn_w[p * 3 + 1], w2 = nn_w[p * 3 + 2];
  const float* f = feat + (long long)b * M * ld_feat;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(f + i0 * ld_feat + c));
    const float4 bq = __ldg(reinterpret_cast<const float4*>(f + i1 * ld_feat + c));
    const float4 cq = __ldg(reinterpret_cast<const float4*>(f + i2 * ld_feat + c));
    float4 v;
    v.x = __fadd_rn(__fadd_rn(__fmul_rn(a.x, w0), __fmul_rn(bq.x, w1)), __fmul_rn(cq.x, w2));
    v.y = __fadd_rn(__fadd_rn(__fmul_rn(a.y, w0), __fmul_rn(bq.y, w1)), __fmul_rn(cq.y, w2));
    v.z = __fadd_rn(__fadd_rn(__fmul_rn(a.z, w0), __fmul_rn(bq.z, w1)), __fmul_rn(cq.z, w2));
    v.w = __fadd_rn(__fadd_rn(__fmul_rn(a.w, w0), __fmul_rn(bq.w, w1)), __fmul_rn(cq.w, w2));
    *reinterpret_cast<float4*>(out + p * ld_out + c) = v;
  }
}

static size_t interp_bytes(int B, int N, int M) {
  return 2 * align_up((size_t)B * 3 * sizeof(float), 256) + align_up((size_t)B * N * sizeof(float4), 256) +
         align_up((size_t)B * M * sizeof(float4), 256) + 2 * align_up((size_t)B * N * 3 * sizeof(float), 256) + 1024;
}

}  // namespace samble

using namespace samble;

extern "C" size_t samble_interpolate3_workspace_bytes(int B, int N, int M) {
  if (B <= 0 || N <= 0 || M <= 0) return 0;
  return interp_bytes(B, N, M);
}

extern "C" int samble_interpolate3(const float* xyz_up, const float* xyz_sel, const float* feat, int B, int N, int M,
                                   int C, float* out, long long* idx_out, float* dist_out, void* ws, size_t ws_bytes,
                                   samble_stream_t stream) {
  SAMBLE_REQUIRE(xyz_up && xyz_sel && feat && out && ws, "samble_interpolate3: null pointer");
  SAMBLE_REQUIRE(B > 0 && N > 0 && C > 0 && B <= 65535, "samble_interpolate3: bad shape");
  SAMBLE_REQUIRE(M >= 3, "samble_interpolate3: need at least 3 selected points, got %d", M);
  SAMBLE_REQUIRE(ws_bytes >= interp_bytes(B, N, M), "samble_interpolate3: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace w(ws, ws_bytes);
  float* mean = w.take<float>((size_t)B * 3);
  float* stdv = w.take<float>((size_t)B * 3);
  float4* up = w.take<float4>((size_t)B * N);
  float4* sel = w.take<float4>((size_t)B * M);
  // channel-major (B,3,N): strides (3N, 1, N)
  if (int e = launch_knn_stats(xyz_up, 3LL * N, 1, N, B, N, 3, mean, stdv, st, nullptr)) return e;
  if (int e = launch_knn_prep_xyz(xyz_up, 3LL * N, 1, N, B, N, 3, mean, stdv, up, st)) return e;
  if (int e = launch_knn_prep_xyz(xyz_sel, 3LL * M, 1, M, B, M, 3, mean, stdv, sel, st)) return e;
  SAMBLE_PRE(st);
  interpolate3_kernel<<<dim3(ceil_div(N, 32), B), 128, 0, st>>>(up, sel, feat, N, M, C, out, idx_out, dist_out, nullptr,
                                                                  nullptr);
  SAMBLE_LAUNCHED("interpolate3_kernel");
  return SAMBLE_OK;
}

extern "C" int samble_interpolate3_rows(const float* xyz_up, const float* xyz_sel, const float* feat, long long ld_feat, int B,
                                        int N, int M, int C, float* out, long long ld_out, void* ws, size_t ws_bytes,
                                        samble_stream_t stream) {
  SAMBLE_REQUIRE(xyz_up && xyz_sel && feat && out && ws, "samble_interpolate3_rows: null pointer");
  SAMBLE_REQUIRE(B > 0 && N > 0 && C > 0 && B <= 65535, "samble_interpolate3_rows: bad shape");
  SAMBLE_REQUIRE(M >= 3, "samble_interpolate3_rows: need at least 3 selected points, got %d", M);
  SAMBLE_REQUIRE(C % 4 == 0 && ld_feat % 4 == 0 && ld_out % 4 == 0 && ld_feat >= C && ld_out >= C &&
                     ((uintptr_t)feat | (uintptr_t)out) % 16 == 0,
                 "samble_interpolate3_rows: rows must be 16-byte aligned, C a multiple of 4");
  SAMBLE_REQUIRE(ws_bytes >= interp_bytes(B, N, M), "samble_interpolate3_rows: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace w(ws, ws_bytes);
  float* mean = w.take<float>((size_t)B * 3);
  float* stdv = w.take<float>((size_t)B * 3);
  float4* up = w.take<float4>((size_t)B * N);
  float4* sel = w.take<float4>((size_t)B * M);
  int* nn_idx = w.take<int>((size_t)B * N * 3);
  float* nn_w = w.take<float>((size_t)B * N * 3);
  if (int e = launch_knn_stats(xyz_up, 3LL * N, 1, N, B, N, 3, mean, stdv, st, nullptr)) return e;
  if (int e = launch_knn_prep_xyz(xyz_up, 3LL * N, 1, N, B, N, 3, mean, stdv, up, st)) return e;
  if (int e = launch_knn_prep_xyz(xyz_sel, 3LL * M, 1, M, B, M, 3, mean, stdv, sel, st)) return e;
  SAMBLE_PRE(st);
  interpolate3_kernel<<<dim3(ceil_div(N, 32), B), 128, 0, st>>>(up, sel, nullptr, N, M, C, nullptr, nullptr, nullptr, nn_idx,
                                                                  nn_w);
  SAMBLE_LAUNCHED("interpolate3_kernel");
  SAMBLE_PRE(st);
  interp3_rows_kernel<<<dim3(ceil_div(N, 8), B), 256, 0, st>>>(nn_idx, nn_w, feat, ld_feat, N, M, C, out, ld_out);
  SAMBLE_LAUNCHED("interp3_rows_kernel");
  return SAMBLE_OK;
}


// The two halves of samble_interpolate3_rows as separate calls, so that the 3-NN search (xyz only) can run as a parallel
// graph branch beside the convolution that produces `feat` (samble_b200.ops.fork).
extern "C" int samble_interpolate3_search(const float* xyz_up, const float* xyz_sel, int B, int N, int M, int* nn_idx, float* nn_w,
                                          void* ws, size_t ws_bytes, samble_stream_t stream) {
  SAMBLE_REQUIRE(xyz_up && xyz_sel && nn_idx && nn_w && ws, "samble_interpolate3_search: null pointer");
  SAMBLE_REQUIRE(B > 0 && N > 0 && B <= 65535, "samble_interpolate3_search: bad shape");
  SAMBLE_REQUIRE(M >= 3, "samble_interpolate3_search: need at least 3 selected points, got %d", M);
  SAMBLE_REQUIRE(ws_bytes >= interp_bytes(B, N, M), "samble_interpolate3_search: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace w(ws, ws_bytes);
  float* mean = w.take<float>((size_t)B * 3);
  float* stdv = w.take<float>((size_t)B * 3);
  float4* up = w.take<float4>((size_t)B * N);
  float4* sel = w.take<float4>((size_t)B * M);
  if (int e = launch_knn_stats(xyz_up, 3LL * N, 1, N, B, N, 3, mean, stdv, st, nullptr)) return e;
  if (int e = launch_knn_prep_xyz(xyz_up, 3LL * N, 1, N, B, N, 3, mean, stdv, up, st)) return e;
  if (int e = launch_knn_prep_xyz(xyz_sel, 3LL * M, 1, M, B, M, 3, mean, stdv, sel, st)) return e;
  SAMBLE_PRE(st);
  interpolate3_kernel<<<dim3(ceil_div(N, 32), B), 128, 0, st>>>(up, sel, nullptr, N, M, 0, nullptr, nullptr, nullptr, nn_idx, nn_w);
  SAMBLE_LAUNCHED("interpolate3_kernel");
  return SAMBLE_OK;
}

extern "C" int samble_interpolate3_gather_rows(const int* nn_idx, const float* nn_w, const float* feat, long long ld_feat, int B, int N,
                                               int M, int C, float* out, long long ld_out, samble_stream_t stream) {
  SAMBLE_REQUIRE(nn_idx && nn_w && feat && out, "samble_interpolate3_gather_rows: null pointer");
  SAMBLE_REQUIRE(B > 0 && N > 0 && M >= 3 && C > 0 && B <= 65535, "samble_interpolate3_gather_rows: bad shape");
  SAMBLE_REQUIRE(C % 4 == 0 && ld_feat % 4 == 0 && ld_out % 4 == 0 && ld_feat >= C && ld_out >= C &&
                     ((uintptr_t)feat | (uintptr_t)out) % 16 == 0,
                 "samble_interpolate3_gather_rows: rows must be 16-byte aligned multiples of 4 channels");
  cudaStream_t st = (cudaStream_t)stream;
  SAMBLE_PRE(st);
  interp3_rows_kernel<<<dim3(ceil_div(N, 8), B), 256, 0, st>>>(nn_idx, nn_w, feat, ld_feat, N, M, C, out, ld_out);
  SAMBLE_LAUNCHED("interp3_rows_kernel");
  return SAMBLE_OK;
}
