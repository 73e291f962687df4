// K2 (tensor-core form): feature-space kNN with tcgen05.  reference utils/ops.py:35-43 (cdist + topk).
//
// Two passes over the SAME tf32 contraction, each with a thread-per-row TMEM epilogue that costs a
// handful of instructions per (query, candidate) pair and no cross-lane traffic, then an exact fp32
// re-rank of a small candidate set:
//
//   pass A  approx d~_ij = |b_j|^2 - 2<a_i,b_j>_tf32 ; per row keep the minimum of each of 64 interleaved
//           column groups; the k-th smallest of those 64 minima bounds the k-th nearest approx distance
//           (k distinct candidates lie below it).                            -> thr_i = bound + 2e_i
//   pass B  same contraction; every candidate with d~_ij <= thr_i is appended to the row's buffer
//           (shared memory, <= kCap entries, ascending index).  Then, in the same kernel, the exact
//           fp32 distance of the buffered candidates (the FFMA formula and accumulation order of the
//           exact kernel in knn.cu) and the k smallest by (distance, index) -> idx/dist.
//
// e_i bounds |d~ - d_fp32| rigorously (tf32 truncation 2^-10 per operand; see knn_margin()), so the
// buffer is a superset of the exact kernel's answer and the output is IDENTICAL to knn_feat_kernel's.
// Rows whose buffer would overflow are flagged and recomputed by the exact kernel (knn.cu).
//
// CTA = 128 query rows (UMMA M=128), candidate tiles of 128 (UMMA N=128), K-blocks of 32 channels
// (128-byte rows, SWIZZLE_128B).  Warps 0-3: epilogue (thread = TMEM lane = query row); warp 4: MMA
// issuer; warps 5-8: cp.async loaders.  smem ring of K-block stages (full/empty mbarriers), two TMEM
// accumulators (tmem_full/tmem_empty mbarriers) so the epilogue of tile t overlaps the MMAs of t+1.
#include "common.cuh"
#include "tc_common.cuh"

namespace samble {

constexpr int kTcRows = 128;      // query rows per CTA
constexpr int kTcTile = 128;      // candidates per tile
constexpr int kTcStages = 6;      // smem ring depth (16 KB each)
constexpr int kTcCap = 64;        // candidate buffer entries per row
constexpr int kTcThreads = 288;   // 4 epilogue + 1 mma + 4 loader warps

struct TcSmem {                   // offsets into the 1024-aligned dynamic smem block
  static __host__ __device__ size_t a_bytes(int nkb) { return (size_t)nkb * 16384; }
  static __host__ __device__ size_t total(int nkb, bool pass_b) {
    return a_bytes(nkb) + (size_t)kTcStages * 16384 + (pass_b ? (size_t)kTcRows * kTcCap * 2 + kTcRows * 8 : 0) + 1024 + 256;
  }
};

// |d~ - d_fp32| <= e:  2 * |<a,b>_tf32 - <a,b>| <= 2 * (2^-10 + 2^-10 + 2^-20) |a||b|  (operand truncation)
// + tensor-core fp32 accumulation slop + the fp32 rounding of the exact formula itself.
__device__ __forceinline__ float knn_margin(float aa, float bbmax) {
  const float s = sqrtf(aa * bbmax);
  return (0.00390625f + 0.00012207031f) * s + 3.0517578e-05f * (aa + bbmax);
}

// in-register bitonic sort of 64 floats (ascending); all indices are compile-time after unrolling
__device__ __forceinline__ void sort64(float (&v)[64]) {
#pragma unroll
  for (int size = 2; size <= 64; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
      for (int t = 0; t < 32; ++t) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const float a = v[lo], b = v[hi];
        v[lo] = up ? fminf(a, b) : fmaxf(a, b);
        v[hi] = up ? fmaxf(a, b) : fminf(a, b);
      }
    }
  }
}

template <bool PASS_B, class I>
__global__ void __launch_bounds__(kTcThreads, 1)
    knn_tc_kernel(const float* __restrict__ an, const float* __restrict__ anorm, const float* __restrict__ bn,
                  const float* __restrict__ bnorm, const unsigned* __restrict__ bbmax_bits, int Nq, int Nr, int Cp,
                  int k, float* __restrict__ thr, I* __restrict__ idx_out, float* __restrict__ dist_out,
                  int* __restrict__ row_flags) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int nkb = Cp / 32;
  uint8_t* sA = base;
  uint8_t* sB = sA + TcSmem::a_bytes(nkb);
  uint8_t* tail = sB + (size_t)kTcStages * 16384;
  unsigned short* cand = reinterpret_cast<unsigned short*>(tail);                    // [128][kTcCap]   (pass B)
  int* cand_cnt = reinterpret_cast<int*>(tail + (PASS_B ? kTcRows * kTcCap * 2 : 0));  // [128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + (PASS_B ? kTcRows * kTcCap * 2 + kTcRows * 8 : 0));
  uint64_t* full = bars;                      // [kTcStages]
  uint64_t* empty = bars + kTcStages;         // [kTcStages]
  uint64_t* tfull = bars + 2 * kTcStages;     // [2]
  uint64_t* tempty = tfull + 2;               // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, q0 = blockIdx.x * kTcRows;
  const float* A_g = an + (size_t)b * Nq * Cp;
  const float* B_g = bn + (size_t)b * Nr * Cp;
  const float* bnorm_g = bnorm + (size_t)b * Nr;
  const int ntiles = (Nr + kTcTile - 1) / kTcTile;
  const int G = ntiles * nkb;

  // ---- one-time setup: resident query tile (swizzled), barriers, TMEM ----
  for (int p = tid; p < nkb * 1024; p += kTcThreads) {
    const int kb = p >> 10, row = (p >> 3) & 127, ch = p & 7;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + row < Nq) v = __ldg(reinterpret_cast<const float4*>(A_g + (size_t)(q0 + row) * Cp + kb * 32 + ch * 4));
    *reinterpret_cast<float4*>(sA + (size_t)kb * 16384 + tc::sw128_offset(row, ch)) = v;
  }
  tc::fence_proxy_async();
  if (tid == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      tc::mbar_init(&full[s], 128);
      tc::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      tc::mbar_init(&tfull[a], 1);
      tc::mbar_init(&tempty[a], 128);
    }
    tc::mbar_init_fence();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 256);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp >= 5) {
    // ================= loaders: candidate K-blocks -> swizzled smem ring =================
    const int lt = tid - 160;                         // 0..127
    for (int g = 0; g < G; ++g) {
      const int s = g % kTcStages, ph = (g / kTcStages) & 1;
      tc::mbar_wait(&empty[s], ph ^ 1);
      const int t = g / nkb, kb = g % nkb;
      uint8_t* dst = sB + (size_t)s * 16384;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int p = lt + 128 * i, row = p >> 3, ch = p & 7;
        const int n = t * kTcTile + row;
        const bool ok = n < Nr;
        cp_async16(dst + tc::sw128_offset(row, ch), B_g + (size_t)(ok ? n : 0) * Cp + kb * 32 + ch * 4, ok);
      }
      cp_async_commit();
      if (g > 0) {
        cp_async_wait<1>();
        tc::fence_proxy_async();
        tc::mbar_arrive(&full[(g - 1) % kTcStages]);
      }
    }
    cp_async_wait<0>();
    tc::fence_proxy_async();
    tc::mbar_arrive(&full[(G - 1) % kTcStages]);
  } else if (warp == 4) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = tc::instr_desc(2, kTcRows, kTcTile);
      for (int t = 0; t < ntiles; ++t) {
        const int acc = t & 1;
        tc::mbar_wait(&tempty[acc], ((t >> 1) & 1) ^ 1);
        tc::tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          const int g = t * nkb + kb, s = g % kTcStages;
          tc::mbar_wait(&full[s], (g / kTcStages) & 1);
          tc::tc_fence_after();
          const uint64_t ad = tc::smem_desc_sw128(tc::smem_u32(sA + (size_t)kb * 16384));
          const uint64_t bd = tc::smem_desc_sw128(tc::smem_u32(sB + (size_t)s * 16384));
#pragma unroll
          for (int k8 = 0; k8 < 4; ++k8) tc::mma_tf32(tmem + acc * kTcTile, ad + 2 * k8, bd + 2 * k8, idesc, (kb | k8) != 0);
          tc::mma_commit(&empty[s]);                // smem stage reusable once these MMAs retire
        }
        tc::mma_commit(&tfull[acc]);                // accumulator complete
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue: thread = query row =================
    const int row = warp * 32 + lane;
    const int q = q0 + row;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float gmin[64];
    float my_thr = 0.f;
    int cnt = 0;
    if (!PASS_B) {
#pragma unroll
      for (int i = 0; i < 64; ++i) gmin[i] = INFINITY;
    } else {
      my_thr = q < Nq ? thr[(size_t)b * Nq + q] : -INFINITY;
    }
    for (int t = 0; t < ntiles; ++t) {
      const int acc = t & 1;
      tc::mbar_wait(&tfull[acc], (t >> 1) & 1);
      tc::tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < kTcTile; c0 += 32) {
        float v[32];
        tc::tmem_ld32(tmem + lane_base + acc * kTcTile + c0, v);
        const int jbase = t * kTcTile + c0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int j = jbase + i;
          const float bb = j < Nr ? __ldg(bnorm_g + j) : INFINITY;
          const float d = fmaf(-2.f, v[i], bb);          // + |a|^2 is constant per row: added at the end
          if (!PASS_B) {
            const int gi = (c0 & 32) + i;                 // 64 interleaved groups: column mod 64
            gmin[gi] = fminf(gmin[gi], d);
          } else if (d <= my_thr) {
            if (cnt < kTcCap) cand[row * kTcCap + cnt] = (unsigned short)j;
            ++cnt;
          }
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&tempty[acc]);
    }
    if (!PASS_B) {
      if (q < Nq) {
        sort64(gmin);
        // k-th smallest group minimum: k distinct candidates have approx distance <= it
        float kth = gmin[0];
#pragma unroll
        for (int i = 1; i < 32; ++i) kth = (i == k - 1) ? gmin[i] : kth;
        const float aa = anorm[(size_t)b * Nq + q];
        const float e = knn_margin(aa, __uint_as_float(bbmax_bits[b]));
        // in "d - |a|^2" units, like d above.  The max() covers k or more candidates sitting at (clamped)
        // distance zero, e.g. duplicated points: everything within e of zero must then be collected.
        thr[(size_t)b * Nq + q] = fmaxf(kth + 2.f * e, e - aa);
      }
    } else {
      cand_cnt[row] = cnt;
    }
  }

  if (PASS_B) {
    // ================= exact fp32 re-rank of the buffered candidates: warp per row =================
    __syncthreads();
    const int nwarps = kTcThreads / 32;
    for (int row = warp; row < kTcRows; row += nwarps) {
      const int q = q0 + row;
      if (q >= Nq) break;
      const int cnt = cand_cnt[row];
      if (cnt > kTcCap) {                                 // overflow: hand the row to the exact kernel
        if (lane == 0) row_flags[(size_t)b * Nq + q] = 1;
        continue;
      }
      const float aa = anorm[(size_t)b * Nq + q];
      LaneTopK top;
      top.init(lane, k);
#pragma unroll 1
      for (int r = 0; r < kTcCap; r += 32) {
        if (r >= cnt) break;
        const int e = r + lane;
        const bool have = e < cnt;
        const int j = have ? (int)cand[row * kTcCap + e] : 0;
        float acc = 0.f;
        if (have) {
          const float4* br = reinterpret_cast<const float4*>(B_g + (size_t)j * Cp);
          for (int kb = 0; kb < nkb; ++kb) {
            const uint8_t* ak = sA + (size_t)kb * 16384;
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
              const float4 a4 = *reinterpret_cast<const float4*>(ak + tc::sw128_offset(row, ch));
              const float4 b4 = __ldg(br + kb * 8 + ch);
              acc = fmaf(a4.x, b4.x, acc);
              acc = fmaf(a4.y, b4.y, acc);
              acc = fmaf(a4.z, b4.z, acc);
              acc = fmaf(a4.w, b4.w, acc);
            }
          }
        }
        const float d2 = __fmaf_rn(-2.f, acc, __fadd_rn(aa, have ? __ldg(bnorm_g + j) : 0.f));
        top.offer(dist_bits(d2), j, have);
      }
      const int rk = top.rank();
      if (top.active) {
        const long long o = ((long long)b * Nq + q) * k + rk;
        idx_out[o] = (I)top.i;
        if (dist_out) dist_out[o] = -sqrtf(__uint_as_float(top.d));
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// ---- host side ----
bool knn_tc_eligible(int Nq, int Nr, int C, int k) {
  const int Cp = (int)align_up(C, 32);
  return C >= 4 && Cp <= 128 && k <= 32 && Nr <= 65535 && Nr >= k;
}

template <class I>
int launch_knn_tc(const float* an, const float* anorm, const float* bn, const float* bnorm, const unsigned* bbmax, int B,
                  int Nq, int Nr, int Cp, int k, float* thr, I* idx, float* dist, int* row_flags, cudaStream_t st) {
  const int nkb = Cp / 32;
  dim3 grid(ceil_div(Nq, kTcRows), B);
  {
    size_t smem = TcSmem::total(nkb, false);
    auto kern = knn_tc_kernel<false, I>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return check_launch("knn_tc pass A smem attribute");
    SAMBLE_PRE(st);
    kern<<<grid, kTcThreads, smem, st>>>(an, anorm, bn, bnorm, bbmax, Nq, Nr, Cp, k, thr, idx, dist, row_flags);
    SAMBLE_LAUNCHED("knn_tc_threshold_kernel");
  }
  {
    size_t smem = TcSmem::total(nkb, true);
    auto kern = knn_tc_kernel<true, I>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return check_launch("knn_tc pass B smem attribute");
    SAMBLE_PRE(st);
    kern<<<grid, kTcThreads, smem, st>>>(an, anorm, bn, bnorm, bbmax, Nq, Nr, Cp, k, thr, idx, dist, row_flags);
    SAMBLE_LAUNCHED("knn_tc_select_kernel");
  }
  return SAMBLE_OK;
}

template int launch_knn_tc<int>(const float*, const float*, const float*, const float*, const unsigned*, int, int, int, int,
                                int, float*, int*, float*, int*, cudaStream_t);
template int launch_knn_tc<long long>(const float*, const float*, const float*, const float*, const unsigned*, int, int, int,
                                      int, int, float*, long long*, float*, int*, cudaStream_t);

}  // namespace samble
