// DownSampleToken attention-map row statistics on the tensor cores (reference models/downsample.py:139-153).
//   logits_ij = <q_i, k_j> / sqrt(D)  over all N keys (+ nb token keys);  per row: max and sum of exp(l - max).
// Same pipeline as knn_tc.cu (128 query rows per CTA, key tiles of 128, K-blocks of 32 channels, 128B-swizzled smem,
// two TMEM accumulator sets) with the 3xTF32 operand split of linear_tc.cu, because the sampled indices depend on
// these sums: q tile (hi, lo) resident, key K-blocks streamed by cp.async and split in shared memory.  The
// accumulation chain is cut in two (first / second half of the channels, separate TMEM columns, added in fp32) to
// halve the tensor core's truncation bias.  Thread-per-row online softmax epilogue: ~7 instructions per pair, no
// cross-lane traffic.  The nb token columns are evaluated in exact fp32 (they are an output: :149-152).
#include "common.cuh"
#include "tc_common.cuh"

namespace samble {

constexpr int kRsThreads = 288;
constexpr int kRsStages = 3;
constexpr int kRsAhead = 2;

__global__ void __launch_bounds__(kRsThreads, 1)
    ds_row_stats_tc_kernel(const float* __restrict__ q, long long ldq, const float* __restrict__ k, long long ldk,
                           const float* __restrict__ k_tok, int N, int D, int nb, float scale, float* __restrict__ rowmax,
                           float* __restrict__ rowsum, float* __restrict__ token_logits) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = tc::smem_align1024(smem_raw);
  const int nkb = D / 32;
  uint8_t* sQh = base;                                   // [nkb][16 KB]
  uint8_t* sQl = sQh + (size_t)nkb * 16384;
  uint8_t* sK = sQl + (size_t)nkb * 16384;               // ring: [stage][hi 16 KB | lo 16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sK + (size_t)kRsStages * 32768);
  uint64_t* full = bars;
  uint64_t* empty = bars + kRsStages;
  uint64_t* tfull = empty + kRsStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, q0 = blockIdx.x * 128;
  const float* Q_g = q + (size_t)b * N * ldq;
  const float* K_g = k + (size_t)b * N * ldk;
  const int ntiles = (N + 127) / 128;
  const int G = ntiles * nkb;
  const int nch = nkb >= 2 ? 2 : 1;                       // accumulation chains
  const int half = (nkb + nch - 1) / nch;

  for (int p = tid; p < nkb * 1024; p += kRsThreads) {
    const int kb = p >> 10, row = (p >> 3) & 127, ch = p & 7;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + row < N) v = __ldg(reinterpret_cast<const float4*>(Q_g + (size_t)(q0 + row) * ldq + kb * 32 + ch * 4));
    float4 lo;
    lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
    lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
    lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
    lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
    const uint32_t off = (uint32_t)kb * 16384 + tc::sw128_offset(row, ch);
    *reinterpret_cast<float4*>(sQh + off) = v;
    *reinterpret_cast<float4*>(sQl + off) = lo;
  }
  tc::fence_proxy_async();
  if (tid == 0) {
    for (int s = 0; s < kRsStages; ++s) {
      tc::mbar_init(&full[s], 4);
      tc::mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&tfull[i], 1);
      tc::mbar_init(&tempty[i], 4);
    }
    tc::mbar_init_fence();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp >= 5) {
    // ================= loaders: raw key K-block by cp.async (= hi), lo computed from smem =================
    const int lt = tid - 160;
    auto issue = [&](int g) {
      const int t = g / nkb, kb = g % nkb;
      uint8_t* dst = sK + (size_t)(g % kRsStages) * 32768;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int p = lt + 128 * i, row = p >> 3, ch = p & 7;
        const int n = t * 128 + row;
        const bool ok = n < N;
        cp_async16(dst + tc::sw128_offset(row, ch), K_g + (size_t)(ok ? n : 0) * ldk + kb * 32 + ch * 4, ok);
      }
      cp_async_commit();
    };
    auto finish = [&](int g) {
      uint8_t* hi = sK + (size_t)(g % kRsStages) * 32768;
      uint8_t* lo_t = hi + 16384;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int p = lt + 128 * i;
        const uint32_t off = tc::sw128_offset(p >> 3, p & 7);
        const float4 v = *reinterpret_cast<const float4*>(hi + off);
        float4 lo;
        lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
        lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
        lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
        lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
        *reinterpret_cast<float4*>(lo_t + off) = lo;
      }
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&full[g % kRsStages]);
    };
    for (int g = 0; g < G + kRsAhead; ++g) {
      const int fb = g - kRsAhead;
      if (fb >= 0) {
        cp_async_wait<kRsAhead - 1>();
        finish(fb);
      }
      if (g < G) {
        tc::mbar_wait(&empty[g % kRsStages], ((g / kRsStages) & 1) ^ 1);
        issue(g);
      } else {
        cp_async_commit();
      }
    }
  } else if (warp == 4) {
    // ================= MMA issuer =================
    if (tc::elect_one()) {
      const uint32_t idesc = tc::instr_desc(2, 128, 128);
      int g = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int set = t & 1;
        tc::mbar_wait(&tempty[set], ((t >> 1) & 1) ^ 1);
        tc::tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const int s = g % kRsStages;
          tc::mbar_wait(&full[s], (g / kRsStages) & 1);
          tc::tc_fence_after();
          const uint32_t st = tc::smem_u32(sK + (size_t)s * 32768);
          const uint64_t kh = tc::smem_desc_sw128(st), kl = tc::smem_desc_sw128(st + 16384);
          const uint64_t qh = tc::smem_desc_sw128(tc::smem_u32(sQh + (size_t)kb * 16384));
          const uint64_t ql = tc::smem_desc_sw128(tc::smem_u32(sQl + (size_t)kb * 16384));
          const uint32_t acc = tmem + (set * 2 + kb / half) * 128;
#pragma unroll
          for (int k8 = 0; k8 < 4; ++k8) {
            tc::mma_tf32(acc, qh + 2 * k8, kh + 2 * k8, idesc, ((kb % half) | k8) != 0);
            tc::mma_tf32(acc, ql + 2 * k8, kh + 2 * k8, idesc, 1);
            tc::mma_tf32(acc, qh + 2 * k8, kl + 2 * k8, idesc, 1);
          }
          tc::mma_commit(&empty[s]);
        }
        tc::mma_commit(&tfull[set]);
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue: thread = query row, online softmax over the key tiles =================
    const int row = warp * 32 + lane;
    const int i_q = q0 + row;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const float inv_scale = 1.0f / scale;
    float m = -INFINITY, s = 0.f;
    for (int t = 0; t < ntiles; ++t) {
      const int set = t & 1;
      tc::mbar_wait(&tfull[set], (t >> 1) & 1);
      tc::tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 64) {
        float v[64];
        tc::tmem_ld64(tmem + lane_base + (set * 2) * 128 + c0, v);
        if (nch == 2) {
          float w[64];
          tc::tmem_ld64(tmem + lane_base + (set * 2 + 1) * 128 + c0, w);
#pragma unroll
          for (int i = 0; i < 64; ++i) v[i] += w[i];
        }
        const int jbase = t * 128 + c0;
        float cmax = -INFINITY;
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          v[i] = (jbase + i < N) ? v[i] * inv_scale : -INFINITY;
          cmax = fmaxf(cmax, v[i]);
        }
        const float m_new = fmaxf(m, cmax);            // finite: every chunk that is processed holds a valid column
        float ps = 0.f;
#pragma unroll
        for (int i = 0; i < 64; ++i) ps += __expf(v[i] - m_new);
        s = s * __expf(m - m_new) + ps;
        m = m_new;
        if (jbase + 64 >= N) break;                    // rest of the tile is padding
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty[set]);
    }
    if (i_q < N) {
      // token columns in exact fp32 (ascending channels, one accumulator: same value as the FFMA kernel)
      const float4* qr = reinterpret_cast<const float4*>(Q_g + (size_t)i_q * ldq);
      float lt_[8];
      float tmax = -INFINITY;
      for (int tk = 0; tk < nb; ++tk) {
        float acc = 0.f;
        const float4* kr = reinterpret_cast<const float4*>(k_tok + (size_t)tk * D);
        for (int c4 = 0; c4 < D / 4; ++c4) {
          const float4 a4 = __ldg(qr + c4), b4 = __ldg(kr + c4);
          acc = fmaf(a4.x, b4.x, acc);
          acc = fmaf(a4.y, b4.y, acc);
          acc = fmaf(a4.z, b4.z, acc);
          acc = fmaf(a4.w, b4.w, acc);
        }
        const float l = __fdiv_rn(acc, scale);
        token_logits[((size_t)b * N + i_q) * nb + tk] = l;
        if (tk < 8) lt_[tk] = l;
        tmax = fmaxf(tmax, l);
      }
      const float m_new = fmaxf(m, tmax);
      float ps = 0.f;
      for (int tk = 0; tk < nb && tk < 8; ++tk) ps += expf(lt_[tk] - m_new);
      rowsum[(size_t)b * N + i_q] = s * expf(m - m_new) + ps;
      rowmax[(size_t)b * N + i_q] = m_new;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

bool ds_row_stats_tc_eligible(int D, int nb, long long ldq, long long ldk) {
  return D % 32 == 0 && D <= 128 && nb <= 8 && ldq % 4 == 0 && ldk % 4 == 0;
}

int launch_ds_row_stats_tc(const float* q, long long ldq, const float* k, long long ldk, const float* k_tok, int B, int N, int D,
                           int nb, float* rowmax, float* rowsum, float* token_logits, cudaStream_t st) {
  const int nkb = D / 32;
  size_t smem = (size_t)2 * nkb * 16384 + (size_t)kRsStages * 32768 + 1024 + 256;
  if (cudaFuncSetAttribute(ds_row_stats_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return check_launch("ds_row_stats_tc smem attribute");
  SAMBLE_PRE(st);
  ds_row_stats_tc_kernel<<<dim3(ceil_div(N, 128), B), kRsThreads, smem, st>>>(q, ldq, k, ldk, k_tok, N, D, nb, sqrtf((float)D),
                                                                                rowmax, rowsum, token_logits);
  SAMBLE_LAUNCHED("ds_row_stats_tc_kernel");
  return SAMBLE_OK;
}

}  // namespace samble
