// Fused EdgeConv core on the tensor cores (reference models/embedding.py:29-39; same algebra as edgeconv.cu):
//   out[c][n] = lrelu( max_k ( W2' . lrelu(P'_n + R'_{idx[n,k]}) )[c] + b2[c] )
// The per-edge 64 -> C2 product is a GEMM with M = N*K edge rows.  It runs as tcgen05 kind::tf32 MMAs with the
// 3xTF32 operand split (fp32-class accuracy, see linear_tc.cu): per tile of 4 points (128 edge rows) the loader
// warps gather R' rows, add P', apply LeakyReLU, split hi/lo and write the 128-byte-swizzled A operand; W2' (hi and
// lo) stays resident in shared memory; the accumulator (128 edge rows x C2) lives in TMEM, double buffered, and the
// epilogue takes the max over the 32 lanes (= the 32 edges of one point) of every column with one REDUX.
#include "common.cuh"
#include "tc_common.cuh"

namespace samble {

constexpr int kEtThreads = 416;      // 4 epilogue warps, MMA issuer, 2 x 4 loader warps
constexpr int kEtStages = 4;

__device__ __forceinline__ unsigned f2ord(float f) {       // order-preserving float -> uint
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

template <class I, int C2>
__global__ void __launch_bounds__(kEtThreads, 1)
    edge_mlp_tc_kernel(const float* __restrict__ pr, long long ld_pr, const I* __restrict__ idx, const float* __restrict__ w2,
                       const float* __restrict__ b2, int B, int N, int K, int C1, float* __restrict__ out, long long out_ld, int dbg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = tc::smem_align1024(smem_raw);
  const int nkb = C1 / 32;
  constexpr int kWBlock = C2 * 128;                          // one K-block of W2': C2 rows x 128 B
  uint8_t* sWh = base;                                       // [nkb][C2 x 128 B]
  uint8_t* sWl = sWh + (size_t)nkb * kWBlock;
  uint8_t* sH = sWl + (size_t)nkb * kWBlock;                 // ring: [stage][hi 16 KB | lo 16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sH + (size_t)kEtStages * 32768);
  uint64_t* full = bars;                 // [kEtStages] 4 loader-warp arrivals
  uint64_t* empty = bars + kEtStages;    // [kEtStages] tcgen05.commit
  uint64_t* tfull = empty + kEtStages;   // [2]
  uint64_t* tempty = tfull + 2;          // [2] 4 epilogue-warp arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_per_cloud = (N + 3) / 4;
  const int total = B * tiles_per_cloud;

  // resident W2' (hi = raw, lo = w - trunc(w)), K-major swizzled
  for (int p = tid; p < nkb * C2 * 8; p += kEtThreads) {
    const int kb = p / (C2 * 8), row = (p / 8) % C2, ch = p & 7;
    const float4 v = __ldg(reinterpret_cast<const float4*>(w2 + (size_t)row * C1 + kb * 32 + ch * 4));
    float4 lo;
    lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
    lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
    lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
    lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
    const uint32_t off = (uint32_t)kb * kWBlock + tc::sw128_offset(row, ch);
    *reinterpret_cast<float4*>(sWh + off) = v;
    *reinterpret_cast<float4*>(sWl + off) = lo;
  }
  tc::fence_proxy_async();
  if (tid == 0) {
    for (int s = 0; s < kEtStages; ++s) {
      tc::mbar_init(&full[s], 4);
      tc::mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&tfull[i], 1);
      tc::mbar_init(&tempty[i], 4);
    }
    tc::mbar_init_fence();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 2 * C2);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp >= 5) {
    // ================= loaders: warp = point lw of the tile, its 32 edge rows built cooperatively =================
    // Two groups of four warps take alternate stages, so one group's gathers (an L2 round trip that registers cannot
    // prefetch more than one stage deep) are in flight while the other group builds its stage.
    // Eight lanes share one gathered row: a load instruction covers 4 rows x 128 contiguous bytes = 16 whole sectors.
    // (Round 1 had thread = edge row, i.e. 32 half-used sectors per instruction and the centre row P' re-read by every
    // lane: L1/TEX throughput 85 %, profiles/r2_step_full.md.)
    const int grp = (warp - 5) >> 2, lw = (warp - 5) & 3;
    const int sub = lane & 7, rsel = lane >> 3;
    const int ntl = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA
    const int G = ntl * nkb;
    float4 pv, rv[8];
    auto fetch = [&](int g) {                       // loads for stage g into registers
      const int tile = blockIdx.x + (g / nkb) * gridDim.x, kb = g % nkb;
      const int b = tile / tiles_per_cloud, n = (tile % tiles_per_cloud) * 4 + lw;
      const bool ok = n < N;
      const long long prow = (long long)b * N + (ok ? n : 0);
      const int j = ok ? ld_idx(idx, prow * K + min(lane, K - 1)) : 0;     // lanes >= K repeat the last edge
      pv = ok ? __ldg(reinterpret_cast<const float4*>(pr + prow * ld_pr + kb * 32) + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float* rbase = pr + (long long)b * N * ld_pr + C1 + kb * 32;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int jr = __shfl_sync(kFull, j, t * 4 + rsel);                // edge row t*4 + rsel of this point
        rv[t] = ok ? __ldg(reinterpret_cast<const float4*>(rbase + (long long)jr * ld_pr) + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    if (grp < G && !(dbg & 1)) fetch(grp);
    for (int g = grp; g < G; g += 2) {
      const int s = g % kEtStages;
      tc::mbar_wait(&empty[s], ((g / kEtStages) & 1) ^ 1);
      uint8_t* hh = sH + (size_t)s * 32768;
      uint8_t* hl = hh + 16384;
      if (!(dbg & 8))
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        float4 h;
        h.x = pv.x + rv[t].x, h.y = pv.y + rv[t].y, h.z = pv.z + rv[t].z, h.w = pv.w + rv[t].w;
        h.x = h.x > 0.f ? h.x : 0.2f * h.x;
        h.y = h.y > 0.f ? h.y : 0.2f * h.y;
        h.z = h.z > 0.f ? h.z : 0.2f * h.z;
        h.w = h.w > 0.f ? h.w : 0.2f * h.w;
        float4 lo;
        lo.x = h.x - __uint_as_float(__float_as_uint(h.x) & 0xffffe000u);
        lo.y = h.y - __uint_as_float(__float_as_uint(h.y) & 0xffffe000u);
        lo.z = h.z - __uint_as_float(__float_as_uint(h.z) & 0xffffe000u);
        lo.w = h.w - __uint_as_float(__float_as_uint(h.w) & 0xffffe000u);
        const uint32_t off = tc::sw128_offset(lw * 32 + t * 4 + rsel, sub);
        *reinterpret_cast<float4*>(hh + off) = h;
        *reinterpret_cast<float4*>(hl + off) = lo;
      }
      // this group's next stage: its gathers fly while the fence / arrive / next barrier wait happen
      if (g + 2 < G && !(dbg & 1)) fetch(g + 2);
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&full[s]);
    }
  } else if (warp == 4) {
    // ================= MMA issuer =================
    if (tc::elect_one()) {
      const uint32_t idesc = tc::instr_desc(2, 128, C2);
      int g = 0, it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
        const int set = it & 1;
        tc::mbar_wait(&tempty[set], ((it >> 1) & 1) ^ 1);
        tc::tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const int s = g % kEtStages;
          tc::mbar_wait(&full[s], (g / kEtStages) & 1);
          tc::tc_fence_after();
          const uint32_t st = tc::smem_u32(sH + (size_t)s * 32768);
          const uint64_t hh = tc::smem_desc_sw128(st), hl = tc::smem_desc_sw128(st + 16384);
          const uint64_t wh = tc::smem_desc_sw128(tc::smem_u32(sWh + (size_t)kb * kWBlock));
          const uint64_t wl = tc::smem_desc_sw128(tc::smem_u32(sWl + (size_t)kb * kWBlock));
          const uint32_t acc = tmem + set * C2;
          if (!(dbg & 2))
#pragma unroll
          for (int k8 = 0; k8 < 4; ++k8) {
            tc::mma_tf32(acc, hh + 2 * k8, wh + 2 * k8, idesc, (kb | k8) != 0);
            tc::mma_tf32(acc, hl + 2 * k8, wh + 2 * k8, idesc, 1);
            tc::mma_tf32(acc, hh + 2 * k8, wl + 2 * k8, idesc, 1);
          }
          tc::mma_commit(&empty[s]);
        }
        tc::mma_commit(&tfull[set]);
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue: warp = point, lane = edge row; max over lanes per column =================
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int set = it & 1;
      const int b = tile / tiles_per_cloud, n = (tile % tiles_per_cloud) * 4 + warp;
      tc::mbar_wait(&tfull[set], (it >> 1) & 1);
      tc::tc_fence_after();
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + set * C2;
      if (!(dbg & 4))
#pragma unroll 1
      for (int c0 = 0; c0 < C2; c0 += 32) {
        float v[32];
        tc::tmem_ld32(taddr + c0, v);
        // max over the 32 lanes (= the 32 edges of this point) of each of the 32 columns, by a halving butterfly: at
        // every step a lane keeps the half of its columns selected by one of its lane bits and hands the other half to
        // its partner -- 31 shuffles + 31 max, after which lane l holds column l.  (One REDUX per column, round 1's
        // form, needs an order-preserving float->uint map each way: ~300 instructions per block instead of ~125; the
        // epilogue was 40 % of the kernel at C2 = 128, tools/probe_edge.py.)
#pragma unroll
        for (int w = 16; w >= 1; w >>= 1) {
          const bool upper = (lane & w) != 0;
#pragma unroll
          for (int i = 0; i < w; ++i) {
            const float keep = upper ? v[i + w] : v[i], send = upper ? v[i] : v[i + w];
            v[i] = fmaxf(keep, __shfl_xor_sync(kFull, send, w));
          }
        }
        const float mine = v[0];
        if (n < N) {
          const int c = c0 + lane;
          const float y = mine + __ldg(b2 + c);
          // point-major rows (out_ld > 0): lane = channel -> one 128-byte line per warp store; channel-major: stride N
          out[out_ld ? ((long long)b * N + n) * out_ld + c : ((long long)b * C2 + c) * N + n] = y > 0.f ? y : 0.2f * y;
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty[set]);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 2 * C2);
}

// measurement switches (tools/probe_edge.py): 1 = no gathers, 8 = no stage build, 2 = no MMAs, 4 = idle epilogue (garbage results)
int g_edge_debug = 0;

template <class I, int C2>
static int launch_edge_tc(const float* pr, long long ld_pr, const I* idx, const float* w2, const float* b2, int B, int N,
                          int K, int C1, float* out, long long out_ld, cudaStream_t st) {
  const int nkb = C1 / 32;
  size_t smem = (size_t)2 * nkb * C2 * 128 + (size_t)kEtStages * 32768 + 1024 + 256;
  auto kern = edge_mlp_tc_kernel<I, C2>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return check_launch("edge_mlp_tc smem attribute");
  const long long total = (long long)B * ((N + 3) / 4);
  const int grid = (int)(total < 148 ? total : 148);
  SAMBLE_PRE(st);
  kern<<<grid, kEtThreads, smem, st>>>(pr, ld_pr, idx, w2, b2, B, N, K, C1, out, out_ld, g_edge_debug);
  SAMBLE_LAUNCHED("edge_mlp_tc_kernel");
  return SAMBLE_OK;
}

bool edge_tc_eligible(int K, int C1, int C2) {
  if (!(K <= 32 && C1 % 32 == 0 && C1 <= 128 && (C2 == 64 || C2 == 128))) return false;
  return (size_t)2 * (C1 / 32) * C2 * 128 + (size_t)kEtStages * 32768 + 1280 <= 225 * 1024;
}

template <class I>
int edge_mlp_tc(const float* pr, long long ld_pr, const I* idx, const float* w2, const float* b2, int B, int N, int K, int C1,
                int C2, float* out, long long out_ld, cudaStream_t st) {
  if (C2 == 64) return launch_edge_tc<I, 64>(pr, ld_pr, idx, w2, b2, B, N, K, C1, out, out_ld, st);
  return launch_edge_tc<I, 128>(pr, ld_pr, idx, w2, b2, B, N, K, C1, out, out_ld, st);
}
template int edge_mlp_tc<int>(const float*, long long, const int*, const float*, const float*, int, int, int, int, int, float*, long long, cudaStream_t);
template int edge_mlp_tc<long long>(const float*, long long, const long long*, const float*, const float*, int, int, int, int, int, float*, long long, cudaStream_t);

}  // namespace samble

extern "C" void samble_set_edge_debug(int bits) { samble::g_edge_debug = bits; }
