// Point-wise linear layers on the tensor cores with fp32-class accuracy (3xTF32).
//
// The reference's 1x1 Conv1d/Conv2d/Linear layers over the N points of a cloud (q/k/v projections
// models/attention.py:171-181, downsample.py:124-137; feed-forward attention.py:189-191; Conv+BN+LeakyReLU
// blocks upsample.py:150-160, seg_model.py:141-160) are GEMMs  Y[M x Nout] = X[M x K] W[Nout x K]^T  with
// M = B*N.  Sampled indices downstream depend on the last bits of these activations, so plain TF32 (2^-11
// operand rounding) is not acceptable and cuBLAS fp32 runs on the SIMT pipe (~44 TFLOP/s measured).  Here every
// operand is split x = hi + lo (hi = the 19 bits the tensor core reads, lo = x - hi, exact) and
//     X W^T  ~=  hi*hi + lo*hi + hi*lo        (3 tcgen05.mma kind::tf32, fp32 accumulate in TMEM)
// whose error (~2^-21 relative per product) is of the order of fp32 rounding itself.
//
// CTA = 128 rows x NT output columns.  Per 32-channel K-block one smem stage holds X_hi|X_lo|W_hi|W_lo in the
// 128-byte-swizzled K-major layout; 4 loader warps produce it (LDG -> split -> STS), one thread issues the 12 MMAs,
// 4 epilogue warps (thread = row) apply   y = acc*scale[c] + shift[c] -> LeakyReLU -> + residual   and store
// row-major or channel-major.  mbarrier full/empty ring as in knn_tc.cu.
#include "common.cuh"
#include "tc_common.cuh"

namespace samble {

constexpr int kLinThreads = 288;
constexpr int kLinStages = 2;
// The tensor core adds into its fp32 accumulator with truncation, a bias that grows with the length of the
// accumulation chain (measured: 9e-5 abs at K=1024 vs 2e-5 at K=128 on O(10) outputs).  Chains are therefore cut
// every kLinChain K-blocks: each chunk gets its own TMEM accumulator and the epilogue adds the chunks in fp32.
constexpr int kLinChain = 8;

template <int NT>
struct LinCfg {
  static constexpr int kStageBytes = 2 * 16384 + 2 * NT * 128;   // X_hi, X_lo, W_hi, W_lo
  static constexpr size_t smem = (size_t)kLinStages * kStageBytes + 1024 + 256;
};

struct LinArgs {
  const float* X; long long ldx;      // row-major: X[m*ldx + k];  channel-major (x_cm): X[(b*K + k)*npc + n], m = b*npc + n
  const float* W; long long ldw;      // Nout x K
  const float* scale;                 // [Nout] or null (=1)
  const float* shift;                 // [Nout] (+ b*shift_ldb) or null (=0)
  const float* residual; long long ldr;   // same indexing as out, or null
  float* out; long long ldo;          // row-major: out[m*ldo + c];  channel-major (out_cm): out[(b*Nout + c)*npc + n]
  int M, K, Nout, npc;                // npc = points per cloud (needed by either channel-major side and by shift_ldb)
  int lrelu, x_cm, out_cm, res_first; // res_first: y = (acc + res)*scale + shift  (else residual is added last)
  int res_cm;                         // residual layout (row-major with ldr, or channel-major), independent of out's
  long long shift_ldb;                // per-cloud shift stride (0 = shared)
};

__device__ __forceinline__ void split_store(uint8_t* hi_tile, uint8_t* lo_tile, int row, int ch, float4 v) {
  float4 lo;
  lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
  lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
  lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
  lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
  const uint32_t off = tc::sw128_offset(row, ch);
  *reinterpret_cast<float4*>(hi_tile + off) = v;      // the MMA itself ignores the low 13 mantissa bits
  *reinterpret_cast<float4*>(lo_tile + off) = lo;
}

template <int NT>
__global__ void __launch_bounds__(kLinThreads, 1) linear_tc_kernel(LinArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  using Cfg = LinCfg<NT>;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)kLinStages * Cfg::kStageBytes);
  uint64_t* full = bars;                  // [kLinStages] 128 loader arrivals
  uint64_t* empty = bars + kLinStages;    // [kLinStages] tcgen05.commit
  uint64_t* done = empty + kLinStages;    // accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * NT;
  const int nkb = (a.K + 31) / 32;

  if (tid == 0) {
    for (int s = 0; s < kLinStages; ++s) {
      tc::mbar_init(&full[s], 128);
      tc::mbar_init(&empty[s], 1);
    }
    tc::mbar_init(done, 1);
    tc::mbar_init_fence();
  }
  const int nacc = (nkb + kLinChain - 1) / kLinChain;                 // host guarantees nacc * NT <= 512
  uint32_t tcols = 32;
  while (tcols < (uint32_t)(nacc * NT)) tcols <<= 1;
  if (warp == 0) tc::tmem_alloc(tmem_slot, tcols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp >= 5) {
    // ================= loaders =================
    const int lt = tid - 160;
    const bool vec = (a.ldx % 4 == 0) && (a.ldw % 4 == 0) && (a.K % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(a.X) | reinterpret_cast<uintptr_t>(a.W)) % 16 == 0);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % kLinStages, ph = (kb / kLinStages) & 1;
      tc::mbar_wait(&empty[s], ph ^ 1);
      uint8_t* st = base + (size_t)s * Cfg::kStageBytes;
      uint8_t *xh = st, *xl = st + 16384, *wh = st + 32768, *wl = st + 32768 + NT * 128;
      const int k0 = kb * 32;
      auto load4 = [&](const float* src, long long ld, int row, int rows_valid, int ch) -> float4 {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int k = k0 + ch * 4;
        if (row < rows_valid && k < a.K) {
          const float* p = src + (long long)row * ld + k;
          if (vec) v = __ldg(reinterpret_cast<const float4*>(p));
          else {
            v.x = __ldg(p);
            if (k + 1 < a.K) v.y = __ldg(p + 1);
            if (k + 2 < a.K) v.z = __ldg(p + 2);
            if (k + 3 < a.K) v.w = __ldg(p + 3);
          }
        }
        return v;
      };
      // X K-block: 128 rows x 8 chunks
      if (!a.x_cm) {
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int p = lt + 128 * i;
          v[i] = load4(a.X + (long long)m0 * a.ldx, a.ldx, p >> 3, a.M - m0, p & 7);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int p = lt + 128 * i;
          split_store(xh, xl, p >> 3, p & 7, v[i]);
        }
      } else {
        // channel-major input: thread = row; for each channel the 128 rows of the tile are consecutive points
        const int m = m0 + lt;
        const bool ok = m < a.M;
        const long long bq = ok ? m / a.npc : 0, nq = ok ? m % a.npc : 0;
        const float* src = a.X + (bq * a.K + k0) * a.npc + nq;
        float v[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = (ok && k0 + c < a.K) ? __ldg(src + (long long)c * a.npc) : 0.f;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) split_store(xh, xl, lt, ch, make_float4(v[4 * ch], v[4 * ch + 1], v[4 * ch + 2], v[4 * ch + 3]));
      }
      // W K-block: NT rows x 8 chunks
#pragma unroll
      for (int j0 = 0; j0 < NT * 8; j0 += 128 * 8) {
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int p = j0 + lt + 128 * i;
          v[i] = p < NT * 8 ? load4(a.W + (long long)n0 * a.ldw, a.ldw, p >> 3, a.Nout - n0, p & 7) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int p = j0 + lt + 128 * i;
          if (p < NT * 8) split_store(wh, wl, p >> 3, p & 7, v[i]);
        }
      }
      tc::fence_proxy_async();
      tc::mbar_arrive(&full[s]);
    }
  } else if (warp == 4) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = tc::instr_desc(2, 128, NT);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kLinStages;
        tc::mbar_wait(&full[s], (kb / kLinStages) & 1);
        tc::tc_fence_after();
        const uint32_t st = tc::smem_u32(base + (size_t)s * Cfg::kStageBytes);
        const uint64_t xh = tc::smem_desc_sw128(st), xl = tc::smem_desc_sw128(st + 16384);
        const uint64_t wh = tc::smem_desc_sw128(st + 32768), wl = tc::smem_desc_sw128(st + 32768 + NT * 128);
        const uint32_t acc = tmem + (kb / kLinChain) * NT;
#pragma unroll
        for (int k8 = 0; k8 < 4; ++k8) {
          tc::mma_tf32(acc, xh + 2 * k8, wh + 2 * k8, idesc, ((kb % kLinChain) | k8) != 0);
          tc::mma_tf32(acc, xl + 2 * k8, wh + 2 * k8, idesc, 1);
          tc::mma_tf32(acc, xh + 2 * k8, wl + 2 * k8, idesc, 1);
        }
        tc::mma_commit(&empty[s]);
      }
      tc::mma_commit(done);
    }
    __syncwarp();
  } else {
    // ================= epilogue: thread = output row =================
    tc::mbar_wait(done, 0);
    tc::tc_fence_after();
    const int m = m0 + warp * 32 + lane;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const bool live = m < a.M;
    long long ob = 0, on = 0;
    if (a.npc > 0) {
      ob = (live ? m : 0) / a.npc;
      on = (live ? m : 0) % a.npc;
    }
    const float* shift = a.shift ? a.shift + ob * a.shift_ldb : nullptr;
#pragma unroll 1
    for (int c0 = 0; c0 < NT; c0 += 32) {
      float v[32];
      tc::tmem_ld32(tmem + lane_base + c0, v);      // warp-collective: every lane takes part, stores are predicated
      for (int ac = 1; ac < nacc; ++ac) {           // add the accumulation chunks in fp32
        float w[32];
        tc::tmem_ld32(tmem + lane_base + ac * NT + c0, w);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += w[i];
      }
      if (!live || n0 + c0 >= a.Nout) continue;
      const bool full32 = n0 + c0 + 32 <= a.Nout;
      float r[32];
      if (a.residual) {
        if (a.res_cm) {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = (n0 + c0 + i < a.Nout) ? __ldg(a.residual + (ob * a.Nout + n0 + c0 + i) * a.npc + on) : 0.f;
        } else {
          const float* rrow = a.residual + (long long)m * a.ldr + n0 + c0;
          if (full32 && a.ldr % 4 == 0 && reinterpret_cast<uintptr_t>(a.residual) % 16 == 0) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(rrow + i));
              r[i] = t.x, r[i + 1] = t.y, r[i + 2] = t.z, r[i + 3] = t.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = (n0 + c0 + i < a.Nout) ? __ldg(rrow + i) : 0.f;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int c = min(n0 + c0 + i, a.Nout - 1);
        float y = v[i];
        if (a.residual && a.res_first) y += r[i];
        if (a.scale) y *= __ldg(a.scale + c);
        if (shift) y += __ldg(shift + c);
        if (a.lrelu) y = y > 0.f ? y : 0.2f * y;
        if (a.residual && !a.res_first) y += r[i];
        v[i] = y;
      }
      if (a.out_cm) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (n0 + c0 + i < a.Nout) a.out[(ob * a.Nout + n0 + c0 + i) * a.npc + on] = v[i];
      } else {
        float* orow = a.out + (long long)m * a.ldo + n0 + c0;
        if (full32 && a.ldo % 4 == 0 && reinterpret_cast<uintptr_t>(a.out) % 16 == 0) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(orow + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (n0 + c0 + i < a.Nout) orow[i] = v[i];
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, tcols);
}

template <int NT>
static int launch_linear(const LinArgs& a, cudaStream_t st) {
  auto kern = linear_tc_kernel<NT>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LinCfg<NT>::smem) != cudaSuccess)
    return check_launch("linear_tc smem attribute");
  dim3 grid(ceil_div(a.M, 128), ceil_div(a.Nout, NT));
  SAMBLE_PRE(st);
  kern<<<grid, kLinThreads, LinCfg<NT>::smem, st>>>(a);
  SAMBLE_LAUNCHED("linear_tc_kernel");
  return SAMBLE_OK;
}

}  // namespace samble

using namespace samble;

extern "C" int samble_linear(const float* X, long long ldx, int x_channel_major, const float* W, long long ldw,
                             const float* scale, const float* shift, long long shift_cloud_stride, int lrelu,
                             const float* residual, long long ldr, int residual_channel_major, int residual_first,
                             float* out, long long ldo, int out_channel_major, int M, int K, int Nout,
                             int points_per_cloud, samble_stream_t stream) {
  SAMBLE_REQUIRE(X && W && out, "samble_linear: null pointer");
  SAMBLE_REQUIRE(M > 0 && K > 0 && Nout > 0, "samble_linear: bad shape M=%d K=%d Nout=%d", M, K, Nout);
  const bool need_npc = x_channel_major || out_channel_major || shift_cloud_stride != 0 || (residual && residual_channel_major);
  SAMBLE_REQUIRE(!need_npc || (points_per_cloud > 0 && M % points_per_cloud == 0),
                 "samble_linear: M=%d is not a whole number of clouds of %d points", M, points_per_cloud);
  SAMBLE_REQUIRE(ceil_div(Nout, 64) <= 65535, "samble_linear: Nout too large");
  LinArgs a{X, ldx, W, ldw, scale, shift, residual, ldr, out, ldo, M, K, Nout, need_npc ? points_per_cloud : 0,
            lrelu, x_channel_major, out_channel_major, residual_first, residual_channel_major, shift_cloud_stride};
  cudaStream_t st = (cudaStream_t)stream;
  const int nacc = ceil_div(ceil_div(K, 32), kLinChain);
  SAMBLE_REQUIRE(nacc * 64 <= 512, "samble_linear: K=%d too large (max %d)", K, 8 * kLinChain * 32);
  if (Nout > 128 && nacc * 256 <= 512) return launch_linear<256>(a, st);
  if (Nout > 64 && nacc * 128 <= 512) return launch_linear<128>(a, st);
  return launch_linear<64>(a, st);
}
