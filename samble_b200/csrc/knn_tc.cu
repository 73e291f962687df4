// K2 (tensor-core form): feature-space kNN with tcgen05.  reference utils/ops.py:35-43 (cdist + topk).
//
// Two passes over the SAME tf32 contraction with a thread-per-row TMEM epilogue that costs one or two
// instructions per (query, candidate) pair, then an exact fp32 re-rank of a small candidate set:
//
//   The squared norm of the candidate rides in the GEMM: an extra K=8 MMA per tile multiplies the constant
//   query-side row (1,1,0..) with (-|b|^2/2 split into tf32 hi+lo), so the accumulator holds
//       s_ij = <a_i,b_j>_tf32 - |b_j|^2/2        and      d~_ij = |a_i|^2 - 2 s_ij.
//   pass A  per row keep the MAXIMUM of s over each of 64 interleaved column groups; the k-th largest of
//           those 64 maxima bounds the k-th nearest approx distance (k distinct candidates reach it).
//           -> T_i = that bound, loosened by the rigorous tf32 error e_i (knn_margin)
//   pass B  every candidate with s_ij >= T_i is appended to the row's list (ascending index)
//   pass C  (knn_rerank_kernel) exact fp32 distance of the listed candidates -- the FFMA formula and
//           accumulation order of the exact kernel in knn.cu -- and the k smallest by (distance, index).
//
// The list is a provable superset of the exact kernel's answer, so idx/dist are IDENTICAL to
// knn_feat_kernel's.  Rows whose list overflows are flagged and redone by the exact kernel.
//
// CTA = 128 query rows (UMMA M=128), candidate tiles of 128 (UMMA N=128), K-blocks of 32 channels
// (128-byte rows, SWIZZLE_128B).  Warps 0-3: epilogue (thread = TMEM lane = query row); warp 4: MMA issuer;
// warps 5-8: cp.async loaders.  smem ring of K-block stages (full/empty mbarriers), two TMEM accumulators
// (tmem_full/tmem_empty mbarriers) so the epilogue of tile t overlaps the MMAs of tile t+1.
#include "common.cuh"
#include "tc_common.cuh"

namespace samble {

constexpr int kTcRows = 128;      // query rows per CTA
constexpr int kTcTile = 128;      // candidates per tile
constexpr int kTcStages = 5;      // smem ring depth (16 KB each)
constexpr int kExt = 4096;        // compact non-swizzled [128 x 32 B] slice (tc::smem_desc_nosw)
constexpr int kTcCap = 128;       // candidate list entries per row
constexpr int kTcThreads = 288;   // 4 epilogue + 1 mma + 4 loader warps

static size_t tc_smem_bytes(int nkb, bool pass_b) {
  return (size_t)nkb * 16384 + (size_t)kTcStages * 16384 /* A + B ring */ + 3 * kExt /* A_ext, B_ext x2 */
         + (pass_b ? (size_t)kTcRows * kTcCap * 2 : 0) + 1024 + 256;
}

// |d~ - d_fp32| <= e:  2 * |<a,b>_tf32 - <a,b>| <= 2 * (2^-10 + 2^-10 + 2^-20) |a||b|  (operand truncation)
// + tensor-core fp32 accumulation slop + the tf32 split of |b|^2/2 + the fp32 rounding of the exact formula.
__device__ __forceinline__ float knn_margin(float aa, float bbmax) {
  const float s = sqrtf(aa * bbmax);
  return (0.00390625f + 0.00012207031f) * s + 3.0517578e-05f * (aa + bbmax);
}

// in-register bitonic sort of 64 floats (ascending).  Canonical counted loops so that everything unrolls and
// every index is a compile-time constant (otherwise the array drops to local memory).
__device__ __forceinline__ void sort64(float (&v)[64]) {
#pragma unroll
  for (int ls = 1; ls <= 6; ++ls) {
#pragma unroll
    for (int lt = ls - 1; lt >= 0; --lt) {
      const int size = 1 << ls, stride = 1 << lt;
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

template <bool PASS_B>
__global__ void __launch_bounds__(kTcThreads, 1)
    knn_tc_kernel(const float* __restrict__ an, const float* __restrict__ anorm, const float* __restrict__ bn,
                  const float* __restrict__ bnorm, const unsigned* __restrict__ bbmax_bits, int Nq, int Nr, int Cp,
                  int k, float* __restrict__ thr, unsigned short* __restrict__ cand_out, int* __restrict__ cnt_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int nkb = Cp / 32;
  uint8_t* sA = base;                                       // nkb K-blocks of the query tile
  uint8_t* sB = sA + (size_t)nkb * 16384;                   // ring
  uint8_t* sAx = sB + (size_t)kTcStages * 16384;            // query-side norm slice: (1,1,0,...) per row
  uint8_t* sBx = sAx + kExt;                                // candidate-side norm slice, double buffered per tile
  uint8_t* tail = sBx + 2 * kExt;
  unsigned short* cand = reinterpret_cast<unsigned short*>(tail);                    // [128][kTcCap]   (pass B)
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + (PASS_B ? kTcRows * kTcCap * 2 : 0));
  uint64_t* full = bars;                      // [kTcStages]
  uint64_t* empty = bars + kTcStages;         // [kTcStages]
  uint64_t* tfull = bars + 2 * kTcStages;     // [2]
  uint64_t* tempty = tfull + 2;               // [2]
  uint64_t* xempty = tempty + 2;              // [2]  norm slice of tile t may be overwritten (its MMA retired)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xempty + 2);

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
  for (int p = tid; p < 256; p += kTcThreads) {
    const int row = p >> 1, ch = p & 1;
    *reinterpret_cast<float4*>(sAx + tc::nosw_offset(row, ch, 128)) = ch == 0 ? make_float4(1.f, 1.f, 0.f, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
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
      tc::mbar_init(&xempty[a], 1);
    }
    tc::mbar_init_fence();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 256);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp >= 5) {
    // ================= loaders: candidate K-blocks (+ norm block with K-block 0) -> swizzled smem =================
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
      if (kb == 0) {
        // norm slice row = (-|b|^2/2 as tf32 hi, lo, 0, ...); buffer t&1 is free once the norm MMA of tile t-2 retired.
        // These generic stores precede this thread's fence.proxy.async + arrive on full[] of (t, kb=0).
        tc::mbar_wait(&xempty[t & 1], ((t >> 1) & 1) ^ 1);
        const int n = t * kTcTile + lt;
        const float h = n < Nr ? -0.5f * __ldg(bnorm_g + n) : -1e30f;      // out-of-range candidates can never win
        const float hi = __uint_as_float(__float_as_uint(h) & 0xffffe000u);
        uint8_t* xb = sBx + (size_t)(t & 1) * kExt;
        *reinterpret_cast<float4*>(xb + tc::nosw_offset(lt, 0, 128)) = make_float4(hi, h - hi, 0.f, 0.f);
        *reinterpret_cast<float4*>(xb + tc::nosw_offset(lt, 1, 128)) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
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
      const uint64_t axd = tc::smem_desc_nosw(tc::smem_u32(sAx), 128);
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
          if (kb == nkb - 1) {
            // norm slice of tile t: stored before the loaders' arrive on full[] of (t, kb=0), which this thread waited on
            tc::mma_tf32(tmem + acc * kTcTile, axd, tc::smem_desc_nosw(tc::smem_u32(sBx + (size_t)(t & 1) * kExt), 128), idesc, 1);
            tc::mma_commit(&xempty[t & 1]);
          }
          tc::mma_commit(&empty[s]);                // smem stage reusable once these MMAs retire
        }
        tc::mma_commit(&tfull[acc]);                // accumulator complete
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue: thread = query row; s = <a,b> - |b|^2/2 straight from TMEM =================
    const int row = warp * 32 + lane;
    const int q = q0 + row;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float gmax[64];
    float my_thr = 0.f;
    int cnt = 0;
    if (!PASS_B) {
#pragma unroll
      for (int i = 0; i < 64; ++i) gmax[i] = -INFINITY;
    } else {
      my_thr = q < Nq ? thr[(size_t)b * Nq + q] : INFINITY;
    }
    for (int t = 0; t < ntiles; ++t) {
      const int acc = t & 1;
      tc::mbar_wait(&tfull[acc], (t >> 1) & 1);
      tc::tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < kTcTile; c0 += 64) {
        float v[64];
        tc::tmem_ld64(tmem + lane_base + acc * kTcTile + c0, v);
        if (!PASS_B) {
#pragma unroll
          for (int i = 0; i < 64; ++i) gmax[i] = fmaxf(gmax[i], v[i]);     // group = column mod 64
        } else {
          const int jbase = t * kTcTile + c0;
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            const bool hit = v[i] >= my_thr;
            if (hit && cnt < kTcCap) cand[row * kTcCap + cnt] = (unsigned short)(jbase + i);
            cnt += hit ? 1 : 0;
          }
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&tempty[acc]);
    }
    if (!PASS_B) {
      if (q < Nq) {
        sort64(gmax);
        // k-th LARGEST group maximum: k distinct candidates have s >= it
        // (= the smallest of the top k of the ascending array; written as a min so that no dynamic
        // register-array index appears, which would push gmax[] into local memory)
        float kth = INFINITY;
#pragma unroll
        for (int i = 32; i < 64; ++i) kth = fminf(kth, i >= 64 - k ? gmax[i] : INFINITY);
        const float aa = anorm[(size_t)b * Nq + q];
        const float e = knn_margin(aa, __uint_as_float(bbmax_bits[b]));
        // collect d~ <= d~_kth + 2e  <=>  s >= s_kth - e.  The min() covers k or more candidates sitting at (clamped)
        // distance zero, e.g. duplicated points: then everything with d~ <= e, i.e. s >= (|a|^2 - e)/2, is needed.
        thr[(size_t)b * Nq + q] = fminf(kth - e, 0.5f * (aa - e));
      }
    } else if (q < Nq) {
      cnt_out[(size_t)b * Nq + q] = cnt;
    }
  }

  if (PASS_B) {
    // candidate lists -> global, 256 B per row, coalesced
    __syncthreads();
    for (int row = warp; row < kTcRows; row += kTcThreads / 32) {
      const int q = q0 + row;
      if (q >= Nq) break;
      reinterpret_cast<uint2*>(cand_out + ((size_t)b * Nq + q) * kTcCap)[lane] = reinterpret_cast<const uint2*>(cand + row * kTcCap)[lane];
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// ---- pass C: exact fp32 re-rank.  One warp per query row, lanes across its candidates. ----
template <class I>
__global__ void __launch_bounds__(256) knn_rerank_kernel(const float* __restrict__ an, const float* __restrict__ anorm,
                                                         const float* __restrict__ bn, const float* __restrict__ bnorm,
                                                         const unsigned short* __restrict__ cand,
                                                         const int* __restrict__ cnt_in, int Nq, int Nr, int Cp, int k,
                                                         I* __restrict__ idx_out, float* __restrict__ dist_out,
                                                         int* __restrict__ row_flags) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, q = blockIdx.x * 8 + warp;
  if (q >= Nq) return;
  const size_t rowi = (size_t)b * Nq + q;
  const int cnt = cnt_in[rowi];
  if (cnt > kTcCap) {                                   // overflow: hand the row to the exact kernel
    if (lane == 0) row_flags[rowi] = 1;
    return;
  }
  const float aa = anorm[rowi];
  const float4* ar = reinterpret_cast<const float4*>(an + rowi * Cp);
  const float* B_g = bn + (size_t)b * Nr * Cp;
  const float* bnorm_g = bnorm + (size_t)b * Nr;
  LaneTopK top;
  top.init(lane, k);
  for (int r = 0; r < cnt; r += 32) {
    const int e = r + lane;
    const bool have = e < cnt;
    const int j = have ? (int)cand[rowi * kTcCap + e] : 0;
    const float4* br = reinterpret_cast<const float4*>(B_g + (size_t)j * Cp);
    float acc = 0.f;
#pragma unroll 4
    for (int c4 = 0; c4 < Cp / 4; ++c4) {              // ascending channels, one accumulator: knn.cu's order
      const float4 a4 = __ldg(ar + c4);
      const float4 b4 = __ldg(br + c4);
      acc = fmaf(a4.x, b4.x, acc);
      acc = fmaf(a4.y, b4.y, acc);
      acc = fmaf(a4.z, b4.z, acc);
      acc = fmaf(a4.w, b4.w, acc);
    }
    const float d2 = __fmaf_rn(-2.f, acc, __fadd_rn(aa, __ldg(bnorm_g + j)));
    const unsigned db = dist_bits(d2);
    if (r == 0 && cnt >= k) {                          // the first k listed candidates seed the set directly
      top.fill(db, j, have && lane < k);
      top.offer(db, j, have && lane >= k);
    } else {
      top.offer(db, j, have);
    }
  }
  const int rk = top.rank();
  if (top.active) {
    const long long o = (long long)rowi * k + rk;
    idx_out[o] = (I)top.i;
    if (dist_out) dist_out[o] = -sqrtf(__uint_as_float(top.d));
  }
}

// ---- host side ----
bool knn_tc_eligible(int Nq, int Nr, int C, int k) {
  const int Cp = (int)align_up(C, 32);
  return C >= 4 && Cp <= 128 && k <= 32 && Nr <= 65535 && Nr >= k;
}

size_t knn_tc_workspace_bytes(int B, int Nq) {
  return align_up((size_t)B * Nq * kTcCap * sizeof(unsigned short), 256) + align_up((size_t)B * Nq * sizeof(int), 256);
}

template <class I>
int launch_knn_tc(const float* an, const float* anorm, const float* bn, const float* bnorm, const unsigned* bbmax, int B,
                  int Nq, int Nr, int Cp, int k, float* thr, unsigned short* cand, int* cnt, I* idx, float* dist,
                  int* row_flags, cudaStream_t st) {
  const int nkb = Cp / 32;
  dim3 grid(ceil_div(Nq, kTcRows), B);
  {
    size_t smem = tc_smem_bytes(nkb, false);
    auto kern = knn_tc_kernel<false>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return check_launch("knn_tc pass A smem attribute");
    SAMBLE_PRE(st);
    kern<<<grid, kTcThreads, smem, st>>>(an, anorm, bn, bnorm, bbmax, Nq, Nr, Cp, k, thr, cand, cnt);
    SAMBLE_LAUNCHED("knn_tc_threshold_kernel");
  }
  {
    size_t smem = tc_smem_bytes(nkb, true);
    auto kern = knn_tc_kernel<true>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return check_launch("knn_tc pass B smem attribute");
    SAMBLE_PRE(st);
    kern<<<grid, kTcThreads, smem, st>>>(an, anorm, bn, bnorm, bbmax, Nq, Nr, Cp, k, thr, cand, cnt);
    SAMBLE_LAUNCHED("knn_tc_collect_kernel");
  }
  SAMBLE_PRE(st);
  knn_rerank_kernel<I><<<dim3(ceil_div(Nq, 8), B), 256, 0, st>>>(an, anorm, bn, bnorm, cand, cnt, Nq, Nr, Cp, k, idx, dist, row_flags);
  SAMBLE_LAUNCHED("knn_rerank_kernel");
  return SAMBLE_OK;
}

template int launch_knn_tc<int>(const float*, const float*, const float*, const float*, const unsigned*, int, int, int, int,
                                int, float*, unsigned short*, int*, int*, float*, int*, cudaStream_t);
template int launch_knn_tc<long long>(const float*, const float*, const float*, const float*, const unsigned*, int, int, int,
                                      int, int, float*, unsigned short*, int*, long long*, float*, int*, cudaStream_t);

}  // namespace samble
