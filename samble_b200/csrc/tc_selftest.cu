// Hardware self-test of the tcgen05 plumbing in tc_common.cuh: one CTA computes D = A * B^T
// (A, B: 128 x K fp32, K-major; kind::tf32; fp32 accumulate in TMEM) and writes D (128 x 128).
// It exercises exactly what the production kernels rely on: the 128-byte-swizzled operand layout,
// the shared-memory and instruction descriptors, K-advance inside a swizzle atom, commit -> mbarrier,
// and the TMEM lane/column mapping of tcgen05.ld.32x32b.
#include "tc_common.cuh"

namespace samble {

// With ext != 0 one more K=8 step is taken from compact NON-swizzled [128 x 32 B] slices (Ax, Bx: 128 x 8 fp32),
// the layout the kNN kernel uses to carry the candidate norms through the GEMM.
__global__ void __launch_bounds__(128) tc_gemm_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                               int K, float* __restrict__ D, const float* __restrict__ Ax,
                                                               const float* __restrict__ Bx) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = tc::smem_align1024(smem_raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkb = K / 32;
  uint8_t* sA = base;
  uint8_t* sB = base + (size_t)nkb * 16384;
  uint8_t* sAx = sB + (size_t)nkb * 16384;
  uint8_t* sBx = sAx + 4096;
  if (Ax)
    for (int p = tid; p < 256; p += 128) {
      const int row = p >> 1, ch = p & 1;
      *reinterpret_cast<float4*>(sAx + tc::nosw_offset(row, ch, 128)) = *reinterpret_cast<const float4*>(Ax + row * 8 + ch * 4);
      *reinterpret_cast<float4*>(sBx + tc::nosw_offset(row, ch, 128)) = *reinterpret_cast<const float4*>(Bx + row * 8 + ch * 4);
    }
  for (int kb = 0; kb < nkb; ++kb)
    for (int p = tid; p < 1024; p += 128) {
      const int row = p >> 3, ch = p & 7;
      const float4 a = *reinterpret_cast<const float4*>(A + (size_t)row * K + kb * 32 + ch * 4);
      const float4 b = *reinterpret_cast<const float4*>(B + (size_t)row * K + kb * 32 + ch * 4);
      *reinterpret_cast<float4*>(sA + (size_t)kb * 16384 + tc::sw128_offset(row, ch)) = a;
      *reinterpret_cast<float4*>(sB + (size_t)kb * 16384 + tc::sw128_offset(row, ch)) = b;
    }
  tc::fence_proxy_async();
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 128);
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_init_fence();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = tc::instr_desc(2, 128, 128);
    for (int kb = 0; kb < nkb; ++kb) {
      const uint64_t ad = tc::smem_desc_sw128(tc::smem_u32(sA + (size_t)kb * 16384));
      const uint64_t bd = tc::smem_desc_sw128(tc::smem_u32(sB + (size_t)kb * 16384));
      for (int k8 = 0; k8 < 4; ++k8) tc::mma_tf32(tmem, ad + 2 * k8, bd + 2 * k8, idesc, (kb | k8) != 0);
    }
    if (Ax) tc::mma_tf32(tmem, tc::smem_desc_nosw(tc::smem_u32(sAx), 128), tc::smem_desc_nosw(tc::smem_u32(sBx), 128), idesc, 1);
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < 128; c0 += 32) {
    float v[32];
    tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) D[(size_t)row * 128 + c0 + i] = v[i];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

}  // namespace samble

using namespace samble;

extern "C" int samble_selftest_tc_gemm(const float* A, const float* B, int K, float* D, const float* Ax, const float* Bx,
                                       samble_stream_t stream) {
  SAMBLE_REQUIRE(A && B && D, "samble_selftest_tc_gemm: null pointer");
  SAMBLE_REQUIRE(K > 0 && K % 32 == 0 && K <= 192, "samble_selftest_tc_gemm: K=%d must be a multiple of 32, <= 192", K);
  SAMBLE_REQUIRE((Ax == nullptr) == (Bx == nullptr), "samble_selftest_tc_gemm: Ax and Bx go together");
  size_t smem = (size_t)(K / 32) * 2 * 16384 + 8192 + 1024;
  if (cudaFuncSetAttribute(tc_gemm_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return check_launch("tc_gemm_selftest smem attribute");
  cudaStream_t st = (cudaStream_t)stream;
  SAMBLE_PRE(st);
  tc_gemm_selftest_kernel<<<1, 128, smem, st>>>(A, B, K, D, Ax, Bx);
  SAMBLE_LAUNCHED("tc_gemm_selftest_kernel");
  return SAMBLE_OK;
}

// ---- MMA issue-rate probe: every CTA issues `iters` x 4 back-to-back kind::tf32 128xNx8 MMAs on fixed smem tiles ----
namespace samble {
template <int NT>
__global__ void __launch_bounds__(128) tc_mma_rate_kernel(int iters, long long* cycles_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = tc::smem_align1024(smem_raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int p = tid; p < (16384 + NT * 128) / 16; p += 128) reinterpret_cast<float4*>(base)[p] = make_float4(1.f, 0.5f, 0.25f, 2.f);
  tc::fence_proxy_async();
  if (warp == 0) tc::tmem_alloc(&tmem_slot, NT < 32 ? 32 : NT);
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_init_fence();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = tc::instr_desc(2, 128, NT);
    const uint64_t ad = tc::smem_desc_sw128(tc::smem_u32(base)), bd = tc::smem_desc_sw128(tc::smem_u32(base + 16384));
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int k8 = 0; k8 < 4; ++k8) tc::mma_tf32(tmem, ad + 2 * k8, bd + 2 * k8, idesc, 1);
    tc::mma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    cycles_out[blockIdx.x] = clock64() - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, NT < 32 ? 32 : NT);
}
}  // namespace samble

extern "C" int samble_selftest_mma_rate(int n_tile, int iters, int ctas, long long* cycles_out, samble_stream_t stream) {
  SAMBLE_REQUIRE(cycles_out && iters > 0 && ctas > 0 && (n_tile == 64 || n_tile == 128 || n_tile == 256), "samble_selftest_mma_rate: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  size_t smem = 16384 + (size_t)n_tile * 128 + 1024;
  SAMBLE_PRE(st);
  if (n_tile == 64) {
    cudaFuncSetAttribute(tc_mma_rate_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tc_mma_rate_kernel<64><<<ctas, 128, smem, st>>>(iters, cycles_out);
  } else if (n_tile == 128) {
    cudaFuncSetAttribute(tc_mma_rate_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tc_mma_rate_kernel<128><<<ctas, 128, smem, st>>>(iters, cycles_out);
  } else {
    cudaFuncSetAttribute(tc_mma_rate_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tc_mma_rate_kernel<256><<<ctas, 128, smem, st>>>(iters, cycles_out);
  }
  SAMBLE_LAUNCHED("tc_mma_rate_kernel");
  return SAMBLE_OK;
}

// ---- extended rate probe: bf16 MMAs of 128 x NT x 16, rotating over NACC accumulators and NA distinct A tiles (the
// exact-product GEMM pattern: one B box against several A digit planes into several accumulators).  Everything is a
// compile-time pattern so that the descriptors stay in uniform registers and the loop is not issue-bound. ----
namespace samble {
template <int NT, int NACC, int NA>
__global__ void __launch_bounds__(128) tc_mma_rate_ex_kernel(int iters, long long* cycles_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = tc::smem_align1024(smem_raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int p = tid; p < (4 * 16384 + NT * 128) / 16; p += 128) reinterpret_cast<uint4*>(base)[p] = make_uint4(0x3f803f80u, 0x3f003f00u, 0x3e803e80u, 0x40004000u);
  tc::fence_proxy_async();
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_init_fence();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = tc::instr_desc(1, 128, NT);
    const uint64_t ad = tc::smem_desc_sw128(tc::smem_u32(base)), bd = tc::smem_desc_sw128(tc::smem_u32(base + 4 * 16384));
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        tc::mma_bf16(tmem + (k % NACC) * NT, ad + (k % NA) * (16384 >> 4) + 2 * k, bd + 2 * k, idesc, 1);
    tc::mma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    cycles_out[blockIdx.x] = clock64() - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

template <int NT, int NACC, int NA>
static void launch_rate_ex(int iters, int ctas, long long* out, cudaStream_t st) {
  const size_t smem = 4 * 16384 + (size_t)NT * 128 + 2048;
  cudaFuncSetAttribute(tc_mma_rate_ex_kernel<NT, NACC, NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tc_mma_rate_ex_kernel<NT, NACC, NA><<<ctas, 128, smem, st>>>(iters, out);
}
}  // namespace samble

extern "C" int samble_selftest_mma_rate_ex(int kind, int n_tile, int n_acc, int n_a, int iters, int ctas, long long* cycles_out,
                                           samble_stream_t stream) {
  SAMBLE_REQUIRE(cycles_out && iters > 0 && ctas > 0 && kind == 1, "samble_selftest_mma_rate_ex: bad arguments (kind 1 = bf16 only)");
  cudaStream_t st = (cudaStream_t)stream;
  const int key = n_tile * 100 + n_acc * 10 + n_a;
  SAMBLE_PRE(st);
  switch (key) {
    case 6411: launch_rate_ex<64, 1, 1>(iters, ctas, cycles_out, st); break;
    case 6441: launch_rate_ex<64, 4, 1>(iters, ctas, cycles_out, st); break;
    case 6444: launch_rate_ex<64, 4, 4>(iters, ctas, cycles_out, st); break;
    case 6414: launch_rate_ex<64, 1, 4>(iters, ctas, cycles_out, st); break;
    case 12811: launch_rate_ex<128, 1, 1>(iters, ctas, cycles_out, st); break;
    case 12841: launch_rate_ex<128, 4, 1>(iters, ctas, cycles_out, st); break;
    case 12844: launch_rate_ex<128, 4, 4>(iters, ctas, cycles_out, st); break;
    case 12814: launch_rate_ex<128, 1, 4>(iters, ctas, cycles_out, st); break;
    case 25611: launch_rate_ex<256, 1, 1>(iters, ctas, cycles_out, st); break;
    case 25621: launch_rate_ex<256, 2, 1>(iters, ctas, cycles_out, st); break;
    case 25614: launch_rate_ex<256, 1, 4>(iters, ctas, cycles_out, st); break;
    default: set_error("samble_selftest_mma_rate_ex: unsupported (n_tile, n_acc, n_a) = (%d, %d, %d)", n_tile, n_acc, n_a); return SAMBLE_E_INVALID;
  }
  SAMBLE_LAUNCHED("tc_mma_rate_ex_kernel");
  return SAMBLE_OK;
}

// ---- A operand from tensor memory (the hidden activations of csrc/mlp2.cu): every thread parks its row of A in TMEM
// columns [128, 128+K) with tcgen05.st; D = A * B^T then reads A from there.  With iters > 0 the MMA sequence is repeated
// (accumulating) and timed: the issue rate of the TMEM-A form. ----
namespace samble {
__global__ void __launch_bounds__(128) tc_gemm_ts_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, int K,
                                                                  float* __restrict__ D, int iters, long long* cycles_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = tc::smem_align1024(smem_raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkb = K / 32;
  uint8_t* sB = base;
  for (int kb = 0; kb < nkb; ++kb)
    for (int p = tid; p < 1024; p += 128) {
      const int row = p >> 3, ch = p & 7;
      *reinterpret_cast<float4*>(sB + (size_t)kb * 16384 + tc::sw128_offset(row, ch)) =
          *reinterpret_cast<const float4*>(B + (size_t)row * K + kb * 32 + ch * 4);
    }
  tc::fence_proxy_async();
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_init_fence();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < K; c0 += 32) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = A[(size_t)row * K + c0 + i];
    tc::tmem_st32(tmem + lane_base + 128 + c0, v);
  }
  tc::tmem_st_wait();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc = tc::instr_desc(2, 128, 128);
    const long long t0 = clock64();
    for (int it = 0; it < (iters > 0 ? iters : 1); ++it)
      for (int kb = 0; kb < nkb; ++kb) {
        const uint32_t bl = tc::smem_desc_sw128_lo(tc::smem_u32(sB + (size_t)kb * 16384));
#pragma unroll
        for (int k8 = 0; k8 < 4; ++k8) tc::mma_tf32_ts(tmem, tmem + 128 + kb * 32 + k8 * 8, bl + 2 * k8, idesc, (it | kb | k8) != 0);
      }
    tc::mma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    if (cycles_out) cycles_out[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  if (blockIdx.x == 0)
    for (int c0 = 0; c0 < 128; c0 += 32) {
      float v[32];
      tc::tmem_ld32(tmem + lane_base + c0, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) D[(size_t)row * 128 + c0 + i] = v[i];
    }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}
}  // namespace samble

extern "C" int samble_selftest_tc_gemm_ts(const float* A, const float* B, int K, float* D, int iters, int ctas, long long* cycles_out,
                                          samble_stream_t stream) {
  SAMBLE_REQUIRE(A && B && D, "samble_selftest_tc_gemm_ts: null pointer");
  SAMBLE_REQUIRE(K > 0 && K % 32 == 0 && K <= 256, "samble_selftest_tc_gemm_ts: K=%d must be a multiple of 32, <= 256", K);
  SAMBLE_REQUIRE(ctas >= 1 && iters >= 0, "samble_selftest_tc_gemm_ts: bad iters / ctas");
  size_t smem = (size_t)(K / 32) * 16384 + 1024;
  if (cudaFuncSetAttribute(tc_gemm_ts_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return check_launch("tc_gemm_ts_selftest smem attribute");
  cudaStream_t st = (cudaStream_t)stream;
  SAMBLE_PRE(st);
  tc_gemm_ts_selftest_kernel<<<ctas, 128, smem, st>>>(A, B, K, D, iters, cycles_out);
  SAMBLE_LAUNCHED("tc_gemm_ts_selftest_kernel");
  return SAMBLE_OK;
}

// ---- bf16 (kind::f16) with the A operand in tensor memory: two K elements per 32-bit column.  A, B arrive as fp32 and are rounded to
// bf16 here; B goes to shared memory as [128 x 64 bf16] SW128 K-tiles, A to TMEM columns [128, 128 + K/2). ----
#include <cuda_bf16.h>
namespace samble {
template <bool A_SMEM>
__global__ void __launch_bounds__(128) tc_gemm_ts_bf16_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, int K,
                                                                       float* __restrict__ D, int iters, long long* cycles_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = tc::smem_align1024(smem_raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkt = K / 64;
  uint8_t* sB = base;
  uint8_t* sA = base + (size_t)nkt * 16384;
  auto pack = [](float lo, float hi) {
    return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(lo)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(hi)) << 16);
  };
  for (int kt = 0; kt < nkt; ++kt)
    for (int p = tid; p < 1024; p += 128) {
      const int row = p >> 3, ch = p & 7;                                // 16-byte chunk = 8 bf16
      const float* src = B + (size_t)row * K + kt * 64 + ch * 8;
      *reinterpret_cast<uint4*>(sB + (size_t)kt * 16384 + tc::sw128_offset(row, ch)) =
          make_uint4(pack(src[0], src[1]), pack(src[2], src[3]), pack(src[4], src[5]), pack(src[6], src[7]));
      const float* sa = A + (size_t)row * K + kt * 64 + ch * 8;
      *reinterpret_cast<uint4*>(sA + (size_t)kt * 16384 + tc::sw128_offset(row, ch)) =
          make_uint4(pack(sa[0], sa[1]), pack(sa[2], sa[3]), pack(sa[4], sa[5]), pack(sa[6], sa[7]));
    }
  tc::fence_proxy_async();
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_init_fence();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < K / 2; c0 += 32) {
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = pack(A[(size_t)row * K + 2 * (c0 + i)], A[(size_t)row * K + 2 * (c0 + i) + 1]);
    tc::tmem_st32u(tmem + lane_base + 128 + c0, v);
  }
  tc::tmem_st_wait();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc = tc::instr_desc(1, 128, 128);
    const long long t0 = clock64();
    for (int it = 0; it < (iters > 0 ? iters : 1); ++it)
      for (int kt = 0; kt < nkt; ++kt) {
        const uint32_t bl = tc::smem_desc_sw128_lo(tc::smem_u32(sB + (size_t)kt * 16384));
        const uint32_t al = tc::smem_desc_sw128_lo(tc::smem_u32(sA + (size_t)kt * 16384));
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16) {
          if (A_SMEM) tc::mma_bf16_lo(tmem, al + 2 * k16, bl + 2 * k16, idesc, (it | kt | k16) != 0);
          else tc::mma_bf16_ts(tmem, tmem + 128 + kt * 32 + k16 * 8, bl + 2 * k16, idesc, (it | kt | k16) != 0);
        }
      }
    tc::mma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    if (cycles_out) cycles_out[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  if (blockIdx.x == 0)
    for (int c0 = 0; c0 < 128; c0 += 32) {
      float v[32];
      tc::tmem_ld32(tmem + lane_base + c0, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) D[(size_t)row * 128 + c0 + i] = v[i];
    }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}
}  // namespace samble

extern "C" int samble_selftest_tc_gemm_ts_bf16(const float* A, const float* B, int K, float* D, int iters, int ctas, long long* cycles_out,
                                               int a_in_smem, samble_stream_t stream) {
  SAMBLE_REQUIRE(A && B && D, "samble_selftest_tc_gemm_ts_bf16: null pointer");
  SAMBLE_REQUIRE(K > 0 && K % 64 == 0 && K <= 256, "samble_selftest_tc_gemm_ts_bf16: K=%d must be a multiple of 64, <= 256", K);
  SAMBLE_REQUIRE(ctas >= 1 && iters >= 0, "samble_selftest_tc_gemm_ts_bf16: bad iters / ctas");
  size_t smem = (size_t)(K / 64) * 2 * 16384 + 1024;
  auto kern = a_in_smem ? tc_gemm_ts_bf16_selftest_kernel<true> : tc_gemm_ts_bf16_selftest_kernel<false>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return check_launch("tc_gemm_ts_bf16_selftest smem attribute");
  cudaStream_t st = (cudaStream_t)stream;
  SAMBLE_PRE(st);
  kern<<<ctas, 128, smem, st>>>(A, B, K, D, iters, cycles_out);
  SAMBLE_LAUNCHED("tc_gemm_ts_bf16_selftest_kernel");
  return SAMBLE_OK;
}
