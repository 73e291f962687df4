// Neighbour gather / grouping kernels: reference utils/ops.py:5-14 (index_points), :47-65,83-112
// (select_neighbors / group after the kNN), :125-133 (neighbor_mask), :136-145 (gather_by_idx).
// All are HBM-write-bound copies; the job is to keep every store a full 128 B line and every
// gather an L1/L2 hit.
#include "common.cuh"
#include "tc_common.cuh"

namespace samble {

// out[b, r, :] = points[b, idx[b, r], :]     points (B,N,C) point-major.
template <class I>
__global__ void __launch_bounds__(256) index_points_kernel(const float* __restrict__ pts, const I* __restrict__ idx,
                                                           int N, int C, long long R, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int lanes = C % 4 == 0 ? C / 4 : C;          // float4 lanes when the row allows it
  const long long total = R * lanes;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long r = t / lanes;
    const int l = (int)(t % lanes);
    const int j = ld_idx(idx, (long long)b * R + r);
    if (C % 4 == 0) {
      const float4* src = reinterpret_cast<const float4*>(pts + ((long long)b * N + j) * C);
      reinterpret_cast<float4*>(out + ((long long)b * R + r) * C)[l] = __ldg(src + l);
    } else {
      out[((long long)b * R + r) * C + l] = __ldg(pts + ((long long)b * N + j) * C + l);
    }
  }
}

// The same gather with the copy engine (north star: "neighbour-gather kernels with TMA-staged feature tiles").  A warp owns
// a run of G consecutive output rows: every lane issues ONE bulk copy (cp.async.bulk, UBLKCP) of its source row into the
// warp's shared-memory tile, the byte count lands on an mbarrier, and one lane hands the whole tile -- G rows, contiguous
// in the output -- back to the copy engine as a single bulk store.  No data passes through registers; loads of run r+1
// are issued while the store of run r drains (two tiles per warp).  Needs 16-byte aligned rows (C % 4 == 0).
constexpr int kIpWarps = 8;
template <class I>
__global__ void __launch_bounds__(kIpWarps * 32) index_points_bulk_kernel(const float* __restrict__ pts, const I* __restrict__ idx,
                                                                          int N, int C, long long R, float* __restrict__ out, int G) {
  extern __shared__ __align__(128) uint8_t ip_smem[];
  __shared__ uint64_t bars[kIpWarps][2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, b = blockIdx.y;
  const uint32_t row_bytes = (uint32_t)C * 4u;
  uint8_t* tile = ip_smem + (size_t)warp * 2 * G * row_bytes;
  if (lane == 0) {
    tc::mbar_init(&bars[warp][0], 1);
    tc::mbar_init(&bars[warp][1], 1);
    tc::mbar_init_fence();
  }
  __syncwarp();
  const long long runs = (R + G - 1) / G;
  int use = 0;
  for (long long run = (long long)blockIdx.x * kIpWarps + warp; run < runs; run += (long long)gridDim.x * kIpWarps, ++use) {
    const int buf = use & 1;
    const long long r0 = run * G;
    const int g = (int)min((long long)G, R - r0);
    uint8_t* dst = tile + (size_t)buf * G * row_bytes;
    if (lane == 0) {
      tc::bulk_wait_read<1>();                           // the store issued from this buffer two runs ago has read it
      tc::mbar_arrive_expect_tx(&bars[warp][buf], (uint32_t)g * row_bytes);
    }
    __syncwarp();
    for (int t = lane; t < g; t += 32) {
      const int j = ld_idx(idx, (long long)b * R + r0 + t);
      tc::bulk_load_1d(dst + (size_t)t * row_bytes, pts + ((long long)b * N + j) * C, row_bytes, &bars[warp][buf]);
    }
    tc::mbar_wait(&bars[warp][buf], (use >> 1) & 1);
    if (lane == 0) {
      tc::bulk_store_1d(out + ((long long)b * R + r0) * C, dst, (uint32_t)g * row_bytes);
      tc::bulk_commit();
    }
  }
  if (lane == 0) tc::bulk_wait<0>();
}

// neighbor / diff from a CHANNEL-major cloud into (B,N,K,C): a 32-point x 32-channel tile is
// transposed through shared memory once, then each (n,k) row is written as contiguous channels.
// The gather source is the channel-major tensor itself (L2-resident: one cloud is <= 1 MB).
template <class I, bool DIFF>
__global__ void __launch_bounds__(256) group_rows_kernel(const float* __restrict__ pcd, const I* __restrict__ idx,
                                                         int C, int N, int K, float* __restrict__ out) {
  // thread = (channel lane, row); rows = (n,k) pairs; consecutive threads walk channels so stores coalesce.
  const int b = blockIdx.y;
  const long long rows = (long long)N * K;
  const int cl = threadIdx.x % 32, rl = threadIdx.x / 32;
  for (long long r0 = (long long)blockIdx.x * 8; r0 < rows; r0 += (long long)gridDim.x * 8) {
    const long long r = r0 + rl;
    if (r >= rows) continue;
    const int n = (int)(r / K);
    const int j = ld_idx(idx, (long long)b * rows + r);
    const float* base = pcd + (long long)b * C * N;
    for (int c = cl; c < C; c += 32) {
      float v = __ldg(base + (long long)c * N + j);
      if (DIFF) v = __fsub_rn(v, __ldg(base + (long long)c * N + n));
      out[((long long)b * rows + r) * C + c] = v;
    }
  }
}

// center_* : out (B,2C,N,K).  A CTA owns one (cloud, channel) plane pair: the channel's N values are
// staged in shared memory, then out[c][n][k] = x[n] and out[C+c][n][k] = x[idx[n][k]] (- x[n]) stream
// out as fully coalesced 128 B rows.
template <class I, bool DIFF>
__global__ void __launch_bounds__(256) group_center_kernel(const float* __restrict__ pcd, const I* __restrict__ idx,
                                                           int C, int N, int K, float* __restrict__ out) {
  extern __shared__ float plane[];
  const int c = blockIdx.x, b = blockIdx.y;
  const float* src = pcd + ((long long)b * C + c) * N;
  for (int n = threadIdx.x; n < N; n += blockDim.x) plane[n] = src[n];
  __syncthreads();
  const long long rows = (long long)N * K;
  float* o_center = out + ((long long)b * 2 * C + c) * rows;
  float* o_nbr = out + ((long long)b * 2 * C + C + c) * rows;
  const I* ib = idx + (long long)b * rows;
  // split the plane over gridDim.z CTAs
  const long long per = (rows + gridDim.z - 1) / gridDim.z;
  const long long lo = blockIdx.z * per, hi = min(rows, lo + per);
  for (long long r = lo + threadIdx.x; r < hi; r += blockDim.x) {
    const float ctr = plane[r / K];
    float v = plane[ld_idx(ib, r)];
    if (DIFF) v = __fsub_rn(v, ctr);
    o_center[r] = ctr;
    o_nbr[r] = v;
  }
}

template <class I>
__global__ void __launch_bounds__(256) gather_by_idx_kernel(const float* __restrict__ pcd, const I* __restrict__ idx,
                                                            int C, int N, int M, float* __restrict__ out) {
  const int b = blockIdx.y;
  const long long total = (long long)C * M;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t / M), m = (int)(t % M);
    out[(long long)b * total + t] = __ldg(pcd + ((long long)b * C + c) * N + ld_idx(idx, (long long)b * M + m));
  }
}

template <class I>
__global__ void __launch_bounds__(256) mask_scatter_kernel(const I* __restrict__ idx, int N, long long rows_k,
                                                           float* __restrict__ out, int K) {
  const int b = blockIdx.y;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < rows_k; t += (long long)gridDim.x * blockDim.x) {
    const long long n = t / K;
    out[((long long)b * N + n) * N + ld_idx(idx, (long long)b * rows_k + t)] = 1.0f;
  }
}

static int grid_for(long long work, int per_block, int cap = 148 * 16) {
  long long g = (work + per_block - 1) / per_block;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

template <class I>
static int group_impl(const float* pcd, const I* idx, int B, int C, int N, int K, int type, float* out, cudaStream_t st) {
  const long long rows = (long long)N * K;
  if (type == SAMBLE_GROUP_NEIGHBOR || type == SAMBLE_GROUP_DIFF) {
    dim3 grid(grid_for(rows, 8), B);
    SAMBLE_PRE(st);
    if (type == SAMBLE_GROUP_DIFF)
      group_rows_kernel<I, true><<<grid, 256, 0, st>>>(pcd, idx, C, N, K, out);
    else
      group_rows_kernel<I, false><<<grid, 256, 0, st>>>(pcd, idx, C, N, K, out);
    SAMBLE_LAUNCHED("group_rows_kernel");
    return SAMBLE_OK;
  }
  SAMBLE_REQUIRE((size_t)N * sizeof(float) <= 200 * 1024, "samble_group: N=%d too large for the center_* plane kernel", N);
  // enough CTAs per plane that B*C*z covers the machine a few times over
  int z = (int)((4LL * 148 + (long long)B * C - 1) / ((long long)B * C));
  z = z < 1 ? 1 : (z > 16 ? 16 : z);
  dim3 grid(C, B, z);
  size_t smem = (size_t)N * sizeof(float);
  auto kd = group_center_kernel<I, true>;
  auto kn = group_center_kernel<I, false>;
  if (smem > 48 * 1024) {
    cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  SAMBLE_PRE(st);
  if (type == SAMBLE_GROUP_CENTER_DIFF)
    kd<<<grid, 256, smem, st>>>(pcd, idx, C, N, K, out);
  else
    kn<<<grid, 256, smem, st>>>(pcd, idx, C, N, K, out);
  SAMBLE_LAUNCHED("group_center_kernel");
  return SAMBLE_OK;
}

}  // namespace samble

using namespace samble;

static int g_gather_mode = 0;   // 0: copy-engine (bulk) row gather where rows are 16-byte aligned, 1: thread-copy kernel only
extern "C" void samble_set_gather_mode(int mode) { g_gather_mode = mode; }

#define IDX_DISPATCH(bits, CALL64, CALL32) ((bits) == 64 ? (CALL64) : (CALL32))

extern "C" int samble_index_points(const float* points, const void* idx, int idx_bits, int B, int N, int C, int R,
                                   float* out, samble_stream_t stream) {
  SAMBLE_REQUIRE(points && idx && out, "samble_index_points: null pointer");
  SAMBLE_REQUIRE(B > 0 && N > 0 && C > 0 && R >= 0, "samble_index_points: bad shape");
  SAMBLE_REQUIRE(idx_bits == 32 || idx_bits == 64, "samble_index_points: idx_bits must be 32 or 64");
  if (R == 0) return SAMBLE_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (g_gather_mode == 0 && C % 4 == 0 && C >= 16 && (size_t)C * 4 <= 4096 && ((uintptr_t)points | (uintptr_t)out) % 16 == 0) {
    // copy-engine path: tiles of G rows (<= 32, one bulk load per lane) up to 16 KB, two per warp
    int G = (int)(8192 / ((size_t)C * 4));
    G = G > 32 ? 32 : (G < 1 ? 1 : G);
    const size_t smem = (size_t)kIpWarps * 2 * G * C * 4;
    auto k64 = index_points_bulk_kernel<long long>;
    auto k32 = index_points_bulk_kernel<int>;
    if (cudaFuncSetAttribute(k64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(k32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return check_launch("index_points_bulk smem attribute");
    const long long runs = ((long long)R + G - 1) / G;
    long long gx = (runs + kIpWarps - 1) / kIpWarps;
    const long long cap = (148LL * 8 + B - 1) / B;
    dim3 gridb((unsigned)(gx < 1 ? 1 : (gx > cap ? cap : gx)), B);
    SAMBLE_PRE(st);
    if (idx_bits == 64)
      k64<<<gridb, kIpWarps * 32, smem, st>>>(points, (const long long*)idx, N, C, R, out, G);
    else
      k32<<<gridb, kIpWarps * 32, smem, st>>>(points, (const int*)idx, N, C, R, out, G);
    SAMBLE_LAUNCHED("index_points_bulk_kernel");
    return SAMBLE_OK;
  }
  dim3 grid(grid_for((long long)R * (C % 4 == 0 ? C / 4 : C), 256), B);
  SAMBLE_PRE(st);
  if (idx_bits == 64)
    index_points_kernel<long long><<<grid, 256, 0, st>>>(points, (const long long*)idx, N, C, R, out);
  else
    index_points_kernel<int><<<grid, 256, 0, st>>>(points, (const int*)idx, N, C, R, out);
  SAMBLE_LAUNCHED("index_points_kernel");
  return SAMBLE_OK;
}

extern "C" int samble_group(const float* pcd, const void* idx, int idx_bits, int B, int C, int N, int K, int type,
                            float* out, samble_stream_t stream) {
  SAMBLE_REQUIRE(pcd && idx && out, "samble_group: null pointer");
  SAMBLE_REQUIRE(B > 0 && C > 0 && N > 0 && K > 0, "samble_group: bad shape");
  SAMBLE_REQUIRE(type >= 0 && type <= 3, "samble_group: unknown group type %d", type);
  SAMBLE_REQUIRE(idx_bits == 32 || idx_bits == 64, "samble_group: idx_bits must be 32 or 64");
  cudaStream_t st = (cudaStream_t)stream;
  if (idx_bits == 64) return group_impl<long long>(pcd, (const long long*)idx, B, C, N, K, type, out, st);
  return group_impl<int>(pcd, (const int*)idx, B, C, N, K, type, out, st);
}

extern "C" int samble_gather_by_idx(const float* pcd, const void* idx, int idx_bits, int B, int C, int N, int M,
                                    float* out, samble_stream_t stream) {
  SAMBLE_REQUIRE(pcd && idx && out, "samble_gather_by_idx: null pointer");
  SAMBLE_REQUIRE(B > 0 && C > 0 && N > 0 && M > 0, "samble_gather_by_idx: bad shape");
  SAMBLE_REQUIRE(idx_bits == 32 || idx_bits == 64, "samble_gather_by_idx: idx_bits must be 32 or 64");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(grid_for((long long)C * M, 256), B);
  SAMBLE_PRE(st);
  if (idx_bits == 64)
    gather_by_idx_kernel<long long><<<grid, 256, 0, st>>>(pcd, (const long long*)idx, C, N, M, out);
  else
    gather_by_idx_kernel<int><<<grid, 256, 0, st>>>(pcd, (const int*)idx, C, N, M, out);
  SAMBLE_LAUNCHED("gather_by_idx_kernel");
  return SAMBLE_OK;
}

extern "C" int samble_neighbor_mask(const void* idx, int idx_bits, int B, int N, int K, float* out,
                                    samble_stream_t stream) {
  SAMBLE_REQUIRE(idx && out, "samble_neighbor_mask: null pointer");
  SAMBLE_REQUIRE(B > 0 && N > 0 && K > 0, "samble_neighbor_mask: bad shape");
  SAMBLE_REQUIRE(idx_bits == 32 || idx_bits == 64, "samble_neighbor_mask: idx_bits must be 32 or 64");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(out, 0, (size_t)B * N * N * sizeof(float), st) != cudaSuccess) return check_launch("memset mask");
  count_launch();
  dim3 grid(grid_for((long long)N * K, 256), B);
  SAMBLE_PRE(st);
  if (idx_bits == 64)
    mask_scatter_kernel<long long><<<grid, 256, 0, st>>>((const long long*)idx, N, (long long)N * K, out, K);
  else
    mask_scatter_kernel<int><<<grid, 256, 0, st>>>((const int*)idx, N, (long long)N * K, out, K);
  SAMBLE_LAUNCHED("mask_scatter_kernel");
  return SAMBLE_OK;
}

// ---- layout change between the reference's channel-major clouds (B,C,N) and the point-major rows (B,N,C) the
// gathers and GEMM operand tiles want.  32x32 tiles through padded shared memory: both sides coalesced.
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, long long in_sb, long long in_ld, int R,
                                                        int Cc, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* src = in + (long long)b * in_sb;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    if (r < R && c < Cc) tile[ty + 8 * i][tx] = src[(long long)r * in_ld + c];
  }
  __syncthreads();
  float* dst = out + (long long)b * R * Cc;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, r = r0 + tx;
    if (r < R && c < Cc) dst[(long long)c * R + r] = tile[tx][ty + 8 * i];
  }
}

extern "C" int samble_transpose(const float* in, long long in_batch_stride, long long in_row_stride, int B, int R, int C,
                                float* out, samble_stream_t stream) {
  SAMBLE_REQUIRE(in && out, "samble_transpose: null pointer");
  SAMBLE_REQUIRE(B > 0 && R > 0 && C > 0 && B <= 65535 && (R + 31) / 32 <= 65535 && in_row_stride >= C,
                 "samble_transpose: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  SAMBLE_PRE(st);
  transpose_kernel<<<dim3((C + 31) / 32, (R + 31) / 32, B), 256, 0, st>>>(in, in_batch_stride, in_row_stride, R, C, out);
  SAMBLE_LAUNCHED("transpose_kernel");
  return SAMBLE_OK;
}

