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
#include "linear_common.cuh"

namespace samble {

template <int NT>
struct LinCfg {
  static constexpr int kStageBytes = 2 * 16384 + 2 * NT * 128;   // X_hi, X_lo, W_hi, W_lo
  static constexpr size_t smem = (size_t)kLinStages * kStageBytes + 1024 + 256;
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

// Persistent: CTA c walks output tiles c, c+grid, ... (tile = (m-tile, n-tile), n fastest so that neighbouring CTAs
// share the X tile in L2).  The loader streams K-block stages continuously across tile boundaries, the MMA issuer
// alternates between two TMEM accumulator sets, and the epilogue of tile i overlaps the MMAs of tile i+1.
template <int NT>
__global__ void __launch_bounds__(kLinThreads, 1) linear_tc_kernel(LinArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = tc::smem_align1024(smem_raw);
  using Cfg = LinCfg<NT>;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)kLinStages * Cfg::kStageBytes);
  uint64_t* full = bars;                  // [kLinStages] 128 loader arrivals
  uint64_t* empty = bars + kLinStages;    // [kLinStages] tcgen05.commit
  uint64_t* tfull = empty + kLinStages;   // [2] accumulator set complete (tcgen05.commit)
  uint64_t* tempty = tfull + 2;           // [2] accumulator set drained (128 epilogue arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkb = (a.K + 31) / 32;
  const int mtiles = (a.M + 127) / 128, ntiles = (a.Nout + NT - 1) / NT;
  const int total = mtiles * ntiles;
  const int nacc = (nkb + kLinChain - 1) / kLinChain;      // accumulation chunks per tile (host: nacc * NT <= 512)
  const int nsets = (2 * nacc * NT <= 512) ? 2 : 1;        // double-buffer the accumulators when TMEM allows
  uint32_t tcols = 32;
  while (tcols < (uint32_t)(nsets * nacc * NT)) tcols <<= 1;

  if (tid == 0) {
    for (int s = 0; s < kLinStages; ++s) {
      tc::mbar_init(&full[s], 4);        // one arrival per loader warp
      tc::mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&tfull[i], 1);
      tc::mbar_init(&tempty[i], 4);      // one arrival per epilogue warp
    }
    tc::mbar_init_fence();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, tcols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp >= 5) {
    // ================= loaders =================
    // Raw operand tiles ARE the "hi" tiles (the MMA ignores the low 13 mantissa bits), so they are plain async copies.
    // Row-major X: cp.async kLinAhead stages ahead, then this thread re-reads ITS OWN chunks from smem, writes
    // lo = x - trunc(x), fences to the async proxy and arrives.  Channel-major X: coalesced scalar loads, one stage
    // prefetched in registers.  W and W_lo: cp.async (W_lo was precomputed).
    const int lt = tid - 160;
    const int ntl = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA
    const int G = ntl * nkb;
    float pre[32];
    auto tile_of = [&](int g, int& m0, int& n0, int& kb) {
      const int tile = blockIdx.x + (g / nkb) * gridDim.x;
      m0 = (tile / ntiles) * 128;
      n0 = (tile % ntiles) * NT;
      kb = g % nkb;
    };
    auto prefetch_cm = [&](int g) {
      int m0, n0, kb;
      tile_of(g, m0, n0, kb);
      const int m = m0 + lt;
      const bool ok = m < a.M;
      const float* src = a.X + ((ok ? m / a.npc : 0) * (long long)a.K) * a.npc + (ok ? m % a.npc : 0);
#pragma unroll
      for (int c = 0; c < 32; ++c) pre[c] = (ok && kb * 32 + c < a.K) ? __ldg(src + (long long)(kb * 32 + c) * a.npc) : 0.f;
    };
    auto issue = [&](int g) {
      int m0, n0, kb;
      tile_of(g, m0, n0, kb);
      uint8_t* st = base + (size_t)(g % kLinStages) * Cfg::kStageBytes;
      uint8_t *xh = st, *wh = st + 32768, *wl = st + 32768 + NT * 128;
      const int k0 = kb * 32;
      if (!a.x_cm) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int p = lt + 128 * i, row = p >> 3, ch = p & 7;
          const bool ok = (m0 + row) < a.M && (k0 + ch * 4) < a.K;
          cp_async16(xh + tc::sw128_offset(row, ch), a.X + (long long)(ok ? m0 + row : 0) * a.ldx + (ok ? k0 + ch * 4 : 0), ok);
        }
      }
      const float* Wt = a.W + (long long)n0 * a.ldw;
      const float* Wl = a.Wlo + (long long)n0 * a.ldw;
#pragma unroll
      for (int i = 0; i < NT / 16; ++i) {
        const int p = lt + 128 * i, row = p >> 3, ch = p & 7;
        const bool ok = (n0 + row) < a.Nout && (k0 + ch * 4) < a.K;
        const long long off = (long long)(ok ? row : 0) * a.ldw + (ok ? k0 + ch * 4 : 0);
        cp_async16(wh + tc::sw128_offset(row, ch), Wt + off, ok);
        cp_async16(wl + tc::sw128_offset(row, ch), Wl + off, ok);
      }
      cp_async_commit();
    };
    auto finish = [&](int g) {     // this thread's cp.async groups up to stage g have landed
      const int s = g % kLinStages;
      uint8_t* st = base + (size_t)s * Cfg::kStageBytes;
      uint8_t *xh = st, *xl = st + 16384;
      if (!a.x_cm) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int p = lt + 128 * i;
          const uint32_t off = tc::sw128_offset(p >> 3, p & 7);
          const float4 v = *reinterpret_cast<const float4*>(xh + off);
          float4 lo;
          lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
          lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
          lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
          lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
          *reinterpret_cast<float4*>(xl + off) = lo;
        }
      } else {
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) split_store(xh, xl, lt, ch, make_float4(pre[4 * ch], pre[4 * ch + 1], pre[4 * ch + 2], pre[4 * ch + 3]));
      }
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&full[s]);
    };
    if (a.x_cm && G > 0) prefetch_cm(0);
    // Software pipeline: stage g-kLinAhead is completed (split + arrive) BEFORE we block on a free slot for stage g,
    // so the MMAs of the older stage run while the newer loads are being issued.
    for (int g = 0; g < G + kLinAhead; ++g) {
      const int fb = g - kLinAhead;
      if (fb >= 0) {
        cp_async_wait<kLinAhead - 1>();     // groups issued so far: 0..g-1  ->  group fb has landed
        finish(fb);
        if (a.x_cm && fb + 1 < G) prefetch_cm(fb + 1);   // in flight while we wait for the next free stage
      }
      if (g < G) {
        tc::mbar_wait(&empty[g % kLinStages], ((g / kLinStages) & 1) ^ 1);
        issue(g);
      } else {
        cp_async_commit();           // keep the group count in step
      }
    }
  } else if (warp == 4) {
    // ================= MMA issuer =================
    if (tc::elect_one()) {
      const uint32_t idesc = tc::instr_desc(2, 128, NT);
      int g = 0, it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
        const int set = (nsets == 2) ? (it & 1) : 0;
        const int use = (nsets == 2) ? (it >> 1) : it;           // how often this set was used before
        tc::mbar_wait(&tempty[set], (use & 1) ^ 1);
        tc::tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const int s = g % kLinStages;
          tc::mbar_wait(&full[s], (g / kLinStages) & 1);
          tc::tc_fence_after();
          const uint32_t st = tc::smem_u32(base + (size_t)s * Cfg::kStageBytes);
          const uint64_t xh = tc::smem_desc_sw128(st), xl = tc::smem_desc_sw128(st + 16384);
          const uint64_t wh = tc::smem_desc_sw128(st + 32768), wl = tc::smem_desc_sw128(st + 32768 + NT * 128);
          const uint32_t acc = tmem + (set * nacc + kb / kLinChain) * NT;
#pragma unroll
          for (int k8 = 0; k8 < 4; ++k8) {
            tc::mma_tf32(acc, xh + 2 * k8, wh + 2 * k8, idesc, ((kb % kLinChain) | k8) != 0);
            tc::mma_tf32(acc, xl + 2 * k8, wh + 2 * k8, idesc, 1);
            tc::mma_tf32(acc, xh + 2 * k8, wl + 2 * k8, idesc, 1);
          }
          tc::mma_commit(&empty[s]);
        }
        tc::mma_commit(&tfull[set]);
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue: thread = output row =================
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int set = (nsets == 2) ? (it & 1) : 0;
      const int use = (nsets == 2) ? (it >> 1) : it;
      const int m0 = (tile / ntiles) * 128, n0 = (tile % ntiles) * NT;
      tc::mbar_wait(&tfull[set], use & 1);
      tc::tc_fence_after();
      linear_epilogue_tile<NT>(a, tmem, set, nacc, m0, n0, warp, lane);
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty[set]);
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
  const long long total = (long long)ceil_div(a.M, 128) * ceil_div(a.Nout, NT);
  const int grid = (int)(total < 148 ? total : 148);       // one persistent CTA per SM
  SAMBLE_PRE(st);
  kern<<<grid, kLinThreads, LinCfg<NT>::smem, st>>>(a);
  SAMBLE_LAUNCHED("linear_tc_kernel");
  return SAMBLE_OK;
}

int launch_linear_tma_auto(const LinArgs& a, int nacc, cudaStream_t st);   // linear_tma.cu
extern int g_lt_debug;
template <int NT>
int launch_linear_tma(const LinArgs& a, cudaStream_t st);

}  // namespace samble

using namespace samble;

__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ lo, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    lo[i] = v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
  }
}

extern "C" int samble_split_tf32(const float* x, float* lo, long long n, samble_stream_t stream) {
  SAMBLE_REQUIRE(x && lo && n > 0, "samble_split_tf32: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  long long g = (n + 255) / 256;
  SAMBLE_PRE(st);
  split_tf32_kernel<<<(int)(g > 1184 ? 1184 : g), 256, 0, st>>>(x, lo, n);
  SAMBLE_LAUNCHED("split_tf32_kernel");
  return SAMBLE_OK;
}

extern "C" int samble_linear(const float* X, long long ldx, int x_channel_major, const float* W, const float* W_lo, long long ldw,
                             const float* scale, const float* shift, long long shift_cloud_stride, int lrelu,
                             const float* residual, long long ldr, int residual_channel_major, int residual_first,
                             float* out, long long ldo, int out_channel_major, int M, int K, int Nout,
                             int points_per_cloud, samble_stream_t stream) {
  SAMBLE_REQUIRE(X && W && W_lo && out, "samble_linear: null pointer");
  SAMBLE_REQUIRE(ldw % 4 == 0 && ldw >= ((K + 3) / 4) * 4 && ((uintptr_t)W | (uintptr_t)W_lo) % 16 == 0,
                 "samble_linear: weight rows must be 16-byte aligned and zero-padded to a multiple of 4 columns");
  SAMBLE_REQUIRE(x_channel_major || (ldx % 4 == 0 && ldx >= ((K + 3) / 4) * 4 && (uintptr_t)X % 16 == 0),
                 "samble_linear: row-major X needs 16-byte aligned rows, zero-padded to a multiple of 4 columns");
  SAMBLE_REQUIRE(M > 0 && K > 0 && Nout > 0, "samble_linear: bad shape M=%d K=%d Nout=%d", M, K, Nout);
  const bool need_npc = x_channel_major || out_channel_major || shift_cloud_stride != 0 || (residual && residual_channel_major);
  SAMBLE_REQUIRE(!need_npc || (points_per_cloud > 0 && M % points_per_cloud == 0),
                 "samble_linear: M=%d is not a whole number of clouds of %d points", M, points_per_cloud);
  SAMBLE_REQUIRE(ceil_div(Nout, 64) <= 65535, "samble_linear: Nout too large");
  LinArgs a{X, ldx, W, ldw, W_lo, scale, shift, residual, ldr, out, ldo, M, K, Nout, need_npc ? points_per_cloud : 0,
            lrelu, x_channel_major, out_channel_major, residual_first, residual_channel_major, shift_cloud_stride,
            nullptr, nullptr, 0, nullptr, nullptr, 1.f, 0, nullptr};
  cudaStream_t st = (cudaStream_t)stream;
  const int nacc = ceil_div(ceil_div(K, 32), kLinChain);
  SAMBLE_REQUIRE(nacc * 64 <= 512, "samble_linear: K=%d too large (max %d)", K, 8 * kLinChain * 32);
  const bool wide = Nout > 64 && nacc * 128 <= 512;
  if (!x_channel_major) return launch_linear_tma_auto(a, nacc, st);
  return wide ? launch_linear<128>(a, st) : launch_linear<64>(a, st);
}

// ---- pooled linear: max / mean over the points of each cloud of  lrelu(X W^T * scale + shift), y never stored ----
__global__ void linear_pool_finalize_kernel(const float* __restrict__ pmax, const float* __restrict__ psum, int groups_per_cloud,
                                            int Nout, int npc, float* __restrict__ out_max, float* __restrict__ out_mean) {
  const int b = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Nout) return;
  float m = -INFINITY, s = 0.f;
  for (int g = 0; g < groups_per_cloud; ++g) {        // fixed order: deterministic
    const long long o = ((long long)b * groups_per_cloud + g) * Nout + c;
    m = fmaxf(m, pmax[o]);
    if (psum) s += psum[o];
  }
  if (out_max) out_max[(long long)b * Nout + c] = m;
  if (out_mean) out_mean[(long long)b * Nout + c] = s / (float)npc;
}

extern "C" size_t samble_linear_pool_workspace_bytes(int M, int Nout) {
  if (M <= 0 || Nout <= 0) return 0;
  return 2 * align_up((size_t)ceil_div(M, 32) * Nout * sizeof(float), 256) + 256;
}

extern "C" int samble_linear_pool(const float* X, long long ldx, const float* W, const float* W_lo, long long ldw,
                                  const float* scale, const float* shift, long long shift_cloud_stride, int lrelu, int M,
                                  int K, int Nout, int points_per_cloud, float* out_max, float* out_mean, void* ws,
                                  size_t ws_bytes, samble_stream_t stream) {
  SAMBLE_REQUIRE(X && W && W_lo && ws && (out_max || out_mean), "samble_linear_pool: null pointer");
  SAMBLE_REQUIRE(ldw % 4 == 0 && ldw >= ((K + 3) / 4) * 4 && ((uintptr_t)W | (uintptr_t)W_lo) % 16 == 0,
                 "samble_linear_pool: weight rows must be 16-byte aligned and zero-padded to a multiple of 4 columns");
  SAMBLE_REQUIRE(ldx % 4 == 0 && ldx >= ((K + 3) / 4) * 4 && (uintptr_t)X % 16 == 0,
                 "samble_linear_pool: X needs 16-byte aligned rows, zero-padded to a multiple of 4 columns");
  SAMBLE_REQUIRE(M > 0 && K > 0 && Nout > 0, "samble_linear_pool: bad shape M=%d K=%d Nout=%d", M, K, Nout);
  SAMBLE_REQUIRE(points_per_cloud > 0 && points_per_cloud % 32 == 0 && M % points_per_cloud == 0,
                 "samble_linear_pool: clouds of %d points (must be a multiple of 32 dividing M=%d)", points_per_cloud, M);
  SAMBLE_REQUIRE(ws_bytes >= samble_linear_pool_workspace_bytes(M, Nout), "samble_linear_pool: workspace too small");
  const int nacc = ceil_div(ceil_div(K, 32), kLinChain);
  SAMBLE_REQUIRE(nacc * 64 <= 512, "samble_linear_pool: K=%d too large (max %d)", K, 8 * kLinChain * 32);
  Workspace w(ws, ws_bytes);
  float* pmax = w.take<float>((size_t)ceil_div(M, 32) * Nout);
  float* psum = w.take<float>((size_t)ceil_div(M, 32) * Nout);
  LinArgs a{X, ldx, W, ldw, W_lo, scale, shift, nullptr, 0, nullptr, 0, M, K, Nout, points_per_cloud,
            lrelu, 0, 0, 0, 0, shift_cloud_stride, pmax, out_mean ? psum : nullptr, 0, nullptr, nullptr, 1.f, 0, nullptr};
  cudaStream_t st = (cudaStream_t)stream;
  // clouds of whole 128-row tiles and 128-channel tiles: swapped orientation, the reduction over the points runs down each
  // thread's own accumulator columns (linear_common.cuh); one partial per (tile, channel) instead of per (32 rows, channel)
  const bool wide = Nout > 64 && nacc * 128 <= 512;
  a.pool_rows = (points_per_cloud % 128 == 0 && wide && !(g_lt_debug & (64 | 128))) ? 128 : 32;
  if (int e = launch_linear_tma_auto(a, nacc, st)) return e;
  SAMBLE_PRE(st);
  linear_pool_finalize_kernel<<<dim3(ceil_div(Nout, 128), M / points_per_cloud), 128, 0, st>>>(
      pmax, out_mean ? psum : nullptr, points_per_cloud / a.pool_rows, Nout, points_per_cloud, out_max, out_mean);
  SAMBLE_LAUNCHED("linear_pool_finalize_kernel");
  return SAMBLE_OK;
}

// ---- per-cloud products: out[b] = X[b] W[b]^T, optionally turned into softmax rows with known statistics ----
extern "C" int samble_cloud_matmul(const float* X, long long ldx, const float* W, const float* W_lo, long long ldw, int M, int K,
                                   int Nout, int rows_per_cloud, const float* row_max, const float* row_sum, float logit_div,
                                   const float* residual, long long ldr, float* out, long long ldo, samble_stream_t stream) {
  SAMBLE_REQUIRE(X && W && W_lo && out, "samble_cloud_matmul: null pointer");
  SAMBLE_REQUIRE((row_max == nullptr) == (row_sum == nullptr), "samble_cloud_matmul: row_max and row_sum come together");
  SAMBLE_REQUIRE(ldw % 4 == 0 && ldw >= ((K + 3) / 4) * 4 && ((uintptr_t)W | (uintptr_t)W_lo) % 16 == 0,
                 "samble_cloud_matmul: weight rows must be 16-byte aligned and zero-padded to a multiple of 4 columns");
  SAMBLE_REQUIRE(ldx % 4 == 0 && ldx >= ((K + 3) / 4) * 4 && (uintptr_t)X % 16 == 0,
                 "samble_cloud_matmul: X needs 16-byte aligned rows, zero-padded to a multiple of 4 columns");
  SAMBLE_REQUIRE(M > 0 && K > 0 && Nout > 0, "samble_cloud_matmul: bad shape M=%d K=%d Nout=%d", M, K, Nout);
  SAMBLE_REQUIRE(rows_per_cloud > 0 && rows_per_cloud % 128 == 0 && M % rows_per_cloud == 0,
                 "samble_cloud_matmul: %d rows per cloud (must be a multiple of 128 dividing M=%d)", rows_per_cloud, M);
  SAMBLE_REQUIRE(ldo >= Nout, "samble_cloud_matmul: ldo < Nout");
  // long contractions (attention rows x V: K = number of keys, outputs are convex combinations of O(1) values) run
  // 16-block chains so that four accumulators of a 128-wide tile still fit TMEM; the truncation bias of the longer
  // chain (~2e-5 abs) is far inside the activation tolerance
  const int chain = K > 1024 ? 16 : kLinChain;
  const int nacc = ceil_div(ceil_div(K, 32), chain);
  SAMBLE_REQUIRE(nacc * 64 <= 512, "samble_cloud_matmul: K=%d too large (max %d)", K, 8 * 16 * 32);
  SAMBLE_REQUIRE(!residual || ldr >= Nout, "samble_cloud_matmul: ldr < Nout");
  LinArgs a{X, ldx, W, ldw, W_lo, nullptr, nullptr, residual, ldr, out, ldo, M, K, Nout, rows_per_cloud,
            0, 0, 0, 0, 0, 0, nullptr, nullptr, 1, row_max, row_sum, logit_div, chain, nullptr};
  return launch_linear_tma_auto(a, nacc, (cudaStream_t)stream);
}

// ---- DownSampleToken pass 1 on the same kernel: row statistics of softmax(q [k | k_tok]^T / sqrt(D)) ----
// The point columns run as a per-cloud product (W[b] = k[b]) with the row-statistics epilogue: one (max, sum) pair per
// row and 128-key tile.  The finalize kernel merges the pairs in tile order, adds the nb token columns (exact fp32
// dot products, written out as token_logits) and leaves rowmax / rowsum.
// One warp finalises 8 rows.  Phase 1 merges each row's (max, sum) tile pairs (lanes across tiles, butterfly).  Phase 2
// runs the nb token dot products with lane = (row, token): every lane owns one sequential 128-term FMA chain (ascending
// channels, one accumulator = the exact FFMA kernel's bits) fed by 16-byte shared-memory reads -- the first version
// used 4 active lanes and scalar reads per row and saturated the shared-memory pipe (85 us; ncu: mio_throttle 25).
constexpr int kFinRows = 8;          // rows per warp
constexpr int kFinPad = 4;           // floats of padding per staged row (16-byte aligned, conflict-free)

__global__ void __launch_bounds__(256) ds_rowstats_finalize_kernel(const float2* __restrict__ part, int ntiles,
                                                                   const float* __restrict__ q, long long ldq,
                                                                   const float* __restrict__ k_tok, int M, int D, int nb,
                                                                   float scale, float* __restrict__ rowmax,
                                                                   float* __restrict__ rowsum, float* __restrict__ token_logits) {
  extern __shared__ __align__(16) float fsm[];         // [nb][D+pad] tokens | [8 warps][8 rows][D+pad] query rows
  const int ld = D + kFinPad;
  float* s_tok = fsm;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s_q = fsm + (size_t)nb * ld + (size_t)warp * kFinRows * ld;
  for (int i = threadIdx.x; i < nb * D; i += blockDim.x) s_tok[(i / D) * ld + (i % D)] = k_tok[i];
  const long long m_base = ((long long)blockIdx.x * 8 + warp) * kFinRows;
  for (int r = 0; r < kFinRows; ++r) {
    const long long m = m_base + r;
    if (m < M)
      for (int c = lane; c < D; c += 32) s_q[r * ld + c] = q[m * ldq + c];          // coalesced
  }
  __syncthreads();
  if (m_base >= M) return;
  // ---- phase 1: merged point-column statistics of the 8 rows (every lane ends up with all eight pairs) ----
  float mxr[kFinRows], sumr[kFinRows];
#pragma unroll
  for (int r = 0; r < kFinRows; ++r) {
    const long long m = m_base + r;
    float mx = -INFINITY, sum = 0.f;
    if (m < M) {
      for (int t = lane; t < ntiles; t += 32) {
        const float2 p = part[m * ntiles + t];
        const float mn = fmaxf(mx, p.x);
        sum = sum * expf(mx - mn) + p.y * expf(p.x - mn);
        mx = mn;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(kFull, mx, o), os = __shfl_xor_sync(kFull, sum, o);
      const float mn = fmaxf(mx, om);
      sum = (mx == -INFINITY ? 0.f : sum * expf(mx - mn)) + (om == -INFINITY ? 0.f : os * expf(om - mn));
      mx = mn;
    }
    mxr[r] = mx, sumr[r] = sum;
  }
  // ---- phase 2: lane = (row r, token j of the current group of 4) ----
  const int r = lane >> 2, jj = lane & 3;
  const long long m = m_base + r;
  float my_mx = mxr[0], my_sum = sumr[0];
#pragma unroll
  for (int i = 1; i < kFinRows; ++i)
    if (r == i) my_mx = mxr[i], my_sum = sumr[i];
  float lt[8];                                        // up to 8 groups of 4 tokens (nb <= 32)
  float tmax = -INFINITY;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    lt[g] = -INFINITY;
    const int j = g * 4 + jj;
    if (g * 4 < nb) {                                 // warp-uniform
      if (j < nb && m < M) {
        const float4* qr = reinterpret_cast<const float4*>(s_q + r * ld);
        const float4* kt = reinterpret_cast<const float4*>(s_tok + j * ld);
        float acc = 0.f;
        for (int c4 = 0; c4 < D / 4; ++c4) {
          const float4 a = qr[c4], b = kt[c4];
          acc = fmaf(a.x, b.x, acc);
          acc = fmaf(a.y, b.y, acc);
          acc = fmaf(a.z, b.z, acc);
          acc = fmaf(a.w, b.w, acc);
        }
        lt[g] = __fdiv_rn(acc, scale);
        token_logits[m * nb + j] = lt[g];
      }
      tmax = fmaxf(tmax, lt[g]);
    }
  }
  tmax = fmaxf(tmax, __shfl_xor_sync(kFull, tmax, 1));        // over the row's 4 lanes
  tmax = fmaxf(tmax, __shfl_xor_sync(kFull, tmax, 2));
  const float m_new = fmaxf(my_mx, tmax);
  float ps = 0.f;
#pragma unroll
  for (int g = 0; g < 8; ++g)
    if (g * 4 < nb && lt[g] != -INFINITY) ps += expf(lt[g] - m_new);
  ps += __shfl_xor_sync(kFull, ps, 1);
  ps += __shfl_xor_sync(kFull, ps, 2);
  if (jj == 0 && m < M) {
    rowsum[m] = (my_mx == -INFINITY ? 0.f : my_sum * expf(my_mx - m_new)) + ps;
    rowmax[m] = m_new;
  }
}

namespace samble {
// shared with xgemm.cu (samble_ds_row_stats_exact)
int launch_ds_rowstats_finalize(const float2* part, int ntiles, const float* q, long long ldq, const float* k_tok, int M, int D,
                                int nb, float* rowmax, float* rowsum, float* token_logits, cudaStream_t st) {
  SAMBLE_PRE(st);
  const size_t fsmem = (size_t)(nb + 8 * kFinRows) * (D + kFinPad) * sizeof(float);
  if (cudaFuncSetAttribute(ds_rowstats_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem) != cudaSuccess)
    return check_launch("ds_rowstats_finalize smem attribute");
  ds_rowstats_finalize_kernel<<<ceil_div(M, 8 * kFinRows), 256, fsmem, st>>>(part, ntiles, q, ldq, k_tok, M, D, nb, sqrtf((float)D),
                                                                              rowmax, rowsum, token_logits);
  SAMBLE_LAUNCHED("ds_rowstats_finalize_kernel");
  return SAMBLE_OK;
}
}  // namespace samble

extern "C" size_t samble_ds_row_stats_fast_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  return align_up((size_t)B * N * ceil_div(N, 128) * sizeof(float2), 256) + 256;
}

extern "C" int samble_ds_row_stats_fast(const float* q, long long ldq, const float* k, const float* k_lo, long long ldk,
                                        const float* k_tok, int B, int N, int D, int nb, float* rowmax, float* rowsum,
                                        float* token_logits, void* ws, size_t ws_bytes, samble_stream_t stream) {
  SAMBLE_REQUIRE(q && k && k_lo && rowmax && rowsum && ws, "samble_ds_row_stats_fast: null pointer");
  SAMBLE_REQUIRE(nb == 0 || (k_tok && token_logits), "samble_ds_row_stats_fast: token pointers required when nb > 0");
  SAMBLE_REQUIRE(B > 0 && N > 0 && N % 128 == 0, "samble_ds_row_stats_fast: N=%d must be a multiple of 128", N);
  SAMBLE_REQUIRE(D > 0 && D % 4 == 0 && D <= 256, "samble_ds_row_stats_fast: D=%d must be a multiple of 4, <= 256", D);
  SAMBLE_REQUIRE(nb >= 0 && nb <= 32, "samble_ds_row_stats_fast: nb=%d outside [0,32]", nb);
  SAMBLE_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ldq >= D && ldk >= D && ((uintptr_t)q | (uintptr_t)k | (uintptr_t)k_lo) % 16 == 0,
                 "samble_ds_row_stats_fast: q/k need 16-byte aligned rows");
  SAMBLE_REQUIRE(ws_bytes >= samble_ds_row_stats_fast_workspace_bytes(B, N), "samble_ds_row_stats_fast: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace w(ws, ws_bytes);
  const int ntiles = ceil_div(N, 128);
  float2* part = w.take<float2>((size_t)B * N * ntiles);
  const float scale = sqrtf((float)D);
  LinArgs a{q, ldq, k, ldk, k_lo, nullptr, nullptr, nullptr, 0, nullptr, 0, B * N, D, N, N,
            0, 0, 0, 0, 0, 0, nullptr, nullptr, 1, nullptr, nullptr, scale, 2 /* short chains: sharp logits, see ds_rowstats_tc.cu */,
            part};
  if (int e = launch_linear_tma<128>(a, st)) return e;
  return launch_ds_rowstats_finalize(part, ntiles, q, ldq, k_tok, B * N, D, nb, rowmax, rowsum, token_logits, st);
}

