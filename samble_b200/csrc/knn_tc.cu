// K2 (tensor-core form): feature-space kNN with tcgen05.  reference utils/ops.py:35-43 (cdist + topk).
//
// Two passes over the SAME tensor-core contraction with a thread-per-row TMEM epilogue that costs one or two
// instructions per (query, candidate) pair, then an exact fp32 re-rank of a small candidate set:
//
//   The squared norm of the candidate rides in the GEMM: an extra K=16 MMA per tile multiplies the constant
//   query-side row (1,1,1,0..) with (-|b|^2/2 split into three bf16 terms), so the accumulator holds
//       s_ij = <a_i,b_j>_bf16x2 - |b_j|^2/2      and      d~_ij = |a_i|^2 - 2 s_ij,
//   operands split x = hi + lo into two bf16 planes (knn.cu prep), <a,b> ~= ah.bh + al.bh + ah.bl on kind::f16 MMAs
//   (twice the tf32 issue rate, same operand bytes as one tf32 copy, error 2^-16 instead of 2^-10).
//   pass A  per row keep the MAXIMUM of s over each of 64 interleaved column groups; the k-th largest of
//           those 64 maxima bounds the k-th nearest approx distance (k distinct candidates reach it).
//           -> T_i = that bound, loosened by the rigorous contraction error e_i (knn_margin)
//   pass B  every candidate with s_ij >= T_i is appended to the row's list (ascending index)
//   pass C  (knn_rerank_kernel) exact fp32 distance of the listed candidates -- the FFMA formula and
//           accumulation order of the exact kernel in knn.cu -- and the k smallest by (distance, index).
//
// The list is a provable superset of the exact kernel's answer, so idx/dist are IDENTICAL to
// knn_feat_kernel's.  Rows whose list overflows are flagged and redone by the exact kernel.
//
// CTA = 128 query rows (UMMA M=128), candidate tiles of 128 (UMMA N=128), K-tiles of 64 bf16 channels
// (128-byte rows, SWIZZLE_128B), one per operand plane.  Warps 0-3: epilogue (thread = TMEM lane = query row); warp 4: MMA issuer;
// warp 5: TMA producer (one thread: tensor-map box loads of the K-blocks, a bulk copy of the tile's norm slice).
// smem ring of K-block stages (full/empty mbarriers, full[] counted in bytes by the TMA unit), two TMEM accumulators
// (tmem_full/tmem_empty mbarriers) so the epilogue of tile t overlaps the MMAs of tile t+1.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace samble {

constexpr int kTcRows = 128;      // query rows per CTA
constexpr int kTcTile = 128;      // candidates per tile
constexpr int kTcStages = 5;      // smem ring depth (16 KB each), collect pass
// The threshold pass needs neither the query-side lo plane nor the candidate lists: their room goes to the ring.  The
// candidate stream is LATENCY-bound by the bytes in flight per SM, not by L2 bandwidth (tools/probe_knn.py: loads alone
// take 27 of 37 us, and halving the L2 traffic with cluster multicast changes nothing), so depth is what buys time.
constexpr int kTcStagesA = 10;
constexpr int kExt = 4096;        // compact non-swizzled [128 x 32 B] slice (tc::smem_desc_nosw)
constexpr int kTcCap = 120;       // usable candidate-list entries per row (a row that reaches it is redone exactly)
constexpr int kTcListLd = 128;    // list row pitch in global memory (32-bit entries)
constexpr int kTcListSm = 129;    // ... and in shared memory: odd, so the 32 rows of a warp hit 32 banks
constexpr int kTcThreads = 192;   // 4 epilogue warps + MMA issuer + TMA producer

static size_t tc_smem_bytes(int nkt, bool pass_b) {
  return (size_t)(pass_b ? 2 : 1) * nkt * 16384 + (size_t)(pass_b ? kTcStages : kTcStagesA) * 16384 /* A hi[+lo] + B ring */
         + 3 * kExt /* A_ext, B_ext x2 */ + (pass_b ? (size_t)kTcRows * kTcListSm * 4 : 0) + 1024 + 256;
}

// |d~ - d_fp32| <= e.  Operands: x = hi + lo + eps, |eps| <= 2^-18 |x| (two bf16 roundings), and the al.bl product is
// dropped (<= 2^-18 |a||b|): |<a,b> - (ah.bh + al.bh + ah.bl)| <= 3 * 2^-18 |a||b|, i.e. 2^-15.4 |a||b| on d = ... - 2<a,b>.
// The bf16 x bf16 products are exact in fp32; the tensor core's fp32 accumulation truncates: <= 2^-23 per update of an
// accumulator bounded by |a||b| + |b|^2/2, ~25 updates per tile -> budgeted 2^-13 |a||b| (measured chains: DESIGN.md).
// Plus the three-term bf16 split of |b|^2/2 (2^-25) and the fp32 rounding of the exact formula: 2^-15 (|a|^2 + |b|^2).
__device__ __forceinline__ float knn_margin(float aa, float bbmax) {
  const float s = sqrtf(aa * bbmax);
  return (3.0517578e-05f + 0.00012207031f) * s + 3.0517578e-05f * (aa + bbmax);
}

// Pass A only has to bound the k-th distance, so it runs the single product ah.bh (half the operand bytes, a third of
// the MMAs): |<a,b> - ah.bh| <= (2^-9 + 2^-9 + 2^-18) |a||b|, i.e. 2^-7 |a||b| on d, plus the same accumulation / rounding
// budget.  Its threshold is loosened by (e_A + e_B)/2 in score units (derivation at the pass-A epilogue).
__device__ __forceinline__ float knn_margin_coarse(float aa, float bbmax) {
  const float s = sqrtf(aa * bbmax);
  return (0.0078125f + 3.0517578e-05f + 0.00012207031f) * s + 3.0517578e-05f * (aa + bbmax);
}

// A list entry is one 32-bit word: candidate index in the low half, and in the high half the upper 16 bits (sign,
// exponent, 7 mantissa bits) of delta = s - T >= 0, the score's offset above the row's collection threshold T.
// Words therefore order by score, and  delta_lo <= delta <= delta_lo * (1 + 2^-7)  with delta_lo = word & 0xffff0000.
__device__ __forceinline__ uint32_t knn_pack(float delta, int j) { return (__float_as_uint(delta) & 0xffff0000u) | (uint32_t)j; }
__device__ __forceinline__ float knn_delta_lo(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ float knn_delta_hi(float lo) { return lo * 1.0079f; }

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

// CS = CTAs per cluster (1, 2 or 4 consecutive query tiles of one cloud).  Every CTA streams ALL candidate tiles, so with
// CS = 1 the L2 -> SM traffic is (Nq / 128) x the candidate planes; phase ablation (tools/probe_knn.py) showed the loads
// alone taking 27 of 37 us (threshold) and 37 of 71 us (collect) at N = 2048, C = 128 -- the kernel is bound by that stream.
// In a cluster each stage of the candidate ring is fetched by ONE CTA (round robin) and multicast into all of them; a stage
// is released when the MMAs of every CTA have retired (multicast tcgen05.commit onto every CTA's empty[] barrier).
template <bool PASS_B, int CS, bool PROF = false>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(kTcThreads, 1)
    knn_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                  const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                  const float* __restrict__ anorm, const float* __restrict__ bext, const unsigned* __restrict__ bbmax_bits,
                  int Nq, int Nr, int Cp, int k, float* __restrict__ thr, uint32_t* __restrict__ cand_out,
                  int* __restrict__ cnt_out, int dbg, long long* __restrict__ prof_out) {
  // cycle counters of the warp roles exist only in the PROF instantiation (tools/probe_knn_roles.py): even a predicated-off clock read per
  // stage in the MMA-issuing thread slowed the production kernels by 7-13 %
  long long* const prof = PROF ? prof_out : nullptr;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = tc::smem_align1024(smem_raw);
  const int nkt = (Cp + 63) / 64;                           // K-tiles of 64 bf16 (128-byte rows) per plane
  constexpr int kPlanes = PASS_B ? 2 : 1;                   // planes per operand (pass A: hi only)
  constexpr int kStages = PASS_B ? kTcStages : kTcStagesA;
  uint8_t* sA = base;                                       // query tile: [hi: nkt tiles][lo: nkt tiles (pass B)]
  uint8_t* sB = sA + (size_t)kPlanes * nkt * 16384;         // ring
  uint8_t* sAx = sB + (size_t)kStages * 16384;              // query-side norm slice: (1,1,0,...) per row
  uint8_t* sBx = sAx + kExt;                                // candidate-side norm slice, double buffered per tile
  uint8_t* tail = sBx + 2 * kExt;
  uint32_t* cand = reinterpret_cast<uint32_t*>(tail);                                 // [128][kTcListSm]   (pass B)
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + (PASS_B ? kTcRows * kTcListSm * 4 : 0));
  uint64_t* full = bars;                      // [kStages]
  uint64_t* empty = bars + kStages;           // [kStages]
  uint64_t* tfull = bars + 2 * kStages;       // [2]
  uint64_t* tempty = tfull + 2;               // [2]
  uint64_t* afull = tempty + 2;               // query tile landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(afull + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, q0 = blockIdx.x * kTcRows;
  const int ntiles = (Nr + kTcTile - 1) / kTcTile;
  const uint32_t crank = CS > 1 ? tc::cluster_ctarank() : 0u;
  constexpr uint16_t kMask = (uint16_t)((1u << CS) - 1u);

  // ---- one-time setup: query-side norm slice, barriers, TMEM ----
  for (int p = tid; p < 256; p += kTcThreads) {
    const int row = p >> 1, ch = p & 1;
    // bf16 (1, 1, 1, 0, ...): picks up the three bf16 terms of -|b|^2/2
    *reinterpret_cast<uint4*>(sAx + tc::nosw_offset(row, ch, 128)) = ch == 0 ? make_uint4(0x3f803f80u, 0x00003f80u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
  }
  tc::fence_proxy_async();
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], CS);              // one multicast commit per CTA of the cluster
    }
    for (int a = 0; a < 2; ++a) {
      tc::mbar_init(&tfull[a], 1);
      tc::mbar_init(&tempty[a], 128);
    }
    tc::mbar_init(afull, 1);
    tc::mbar_init_fence();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 256);
  tc::tc_fence_before();
  __syncthreads();
  if (CS > 1) tc::cluster_sync();               // every CTA's barriers exist before a peer multicasts into them
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 5) {
    // ================= TMA producer =================
    if (tc::elect_one()) {
      tc::tma_prefetch_desc(&map_a_hi);
      tc::tma_prefetch_desc(&map_b_hi);
      tc::tma_prefetch_desc(&map_b_lo);
      // resident query tile, both planes: boxes of [128 rows x 64 bf16]; rows past Nq / channels past Cp read as zero
      tc::mbar_arrive_expect_tx(afull, (uint32_t)(kPlanes * nkt) * 16384u);
      for (int kt = 0; kt < nkt; ++kt) {
        tc::tma_load_3d(sA + (size_t)kt * 16384, &map_a_hi, afull, kt * 64, q0, b);
        if (PASS_B) tc::tma_load_3d(sA + (size_t)(nkt + kt) * 16384, &map_a_lo, afull, kt * 64, q0, b);
      }
      const float* ext_g = bext + (size_t)b * ntiles * (kExt / 4);
      int s = 0, ph = 0;
      uint32_t seq = 0;                                     // stage sequence number: CTA (seq mod CS) fetches it for the cluster
      long long w_empty = 0;                                // (measurement, tools/probe_knn_roles.py)
      const long long p_t0 = PROF ? clock64() : 0;
      for (int t = 0; t < ntiles; ++t) {
        for (int g = 0; g < kPlanes * nkt; ++g, ++seq) {    // stage order per tile: (kt 0: hi[, lo]), (kt 1: hi[, lo]), ...
          const int kt = PASS_B ? g >> 1 : g;
          const long long p_t = (PROF && prof) ? clock64() : 0;
          tc::mbar_wait(&empty[s], ph ^ 1);                 // the MMAs of EVERY CTA of the cluster have retired from stage s
          if (PROF && prof) w_empty += clock64() - p_t;
          if (g == 0) {
            // the tile's norm slice rides on the barrier of its first stage; buffer t&1 is free once the norm MMA of tile t-2
            // retired -- which is what tfull[t&1] of that tile announces (a tcgen05.commit costs the issuing thread ~300 cycles,
            // tools/probe_knn_roles.py: 96 of them were 30 k of the MMA thread's 55 k cycles, so the slice has no barrier of its
            // own any more; tile t's own completion of tfull[t&1] cannot precede this wait, it needs the loads issued below)
            if (t >= 2) tc::mbar_wait(&tfull[t & 1], ((t - 2) >> 1) & 1);
            tc::mbar_arrive_expect_tx(&full[s], 16384u + kExt);
            tc::bulk_load_1d(sBx + (size_t)(t & 1) * kExt, ext_g + (size_t)t * (kExt / 4), kExt, &full[s]);
          } else {
            tc::mbar_arrive_expect_tx(&full[s], 16384u);
          }
          const CUtensorMap* mp = (PASS_B && (g & 1)) ? &map_b_lo : &map_b_hi;
          if (CS == 1)
            tc::tma_load_3d(sB + (size_t)s * 16384, mp, &full[s], kt * 64, t * kTcTile, b);
          else if (seq % CS == crank)
            tc::tma_load_3d_multicast(sB + (size_t)s * 16384, mp, &full[s], kt * 64, t * kTcTile, b, kMask);
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
      if (PROF && prof) {
        long long* o = prof + 8 * (blockIdx.y * gridDim.x + blockIdx.x);
        o[5] = clock64() - p_t0, o[6] = w_empty;
      }
    }
    __syncwarp();
  } else if (warp == 4) {
    // ================= MMA issuer =================
    if (tc::elect_one()) {
      const uint32_t idesc = tc::instr_desc(1, kTcRows, kTcTile);          // bf16 x bf16 -> fp32
      const uint64_t axd = tc::smem_desc_nosw(tc::smem_u32(sAx), 128);
      long long w_full = 0, w_tempty = 0, w_a = 0, w_issue = 0;          // (measurement, tools/probe_knn_roles.py)
      const long long m_t0 = PROF ? clock64() : 0;
      tc::mbar_wait(afull, 0);
      if (PROF && prof) w_a = clock64() - m_t0;
      int s = 0, ph = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int acc = t & 1;
        long long m_t = (PROF && prof) ? clock64() : 0;
        tc::mbar_wait(&tempty[acc], ((t >> 1) & 1) ^ 1);
        if (PROF && prof) w_tempty += clock64() - m_t;
        tc::tc_fence_after();
        for (int g = 0; g < kPlanes * nkt; ++g) {
          const int kt = PASS_B ? g >> 1 : g;
          m_t = (PROF && prof) ? clock64() : 0;
          tc::mbar_wait(&full[s], ph);
          if (PROF && prof) w_full += clock64() - m_t;
          tc::tc_fence_after();
          const long long i_t = (PROF && prof) ? clock64() : 0;
          const uint64_t ah = tc::smem_desc_sw128(tc::smem_u32(sA + (size_t)kt * 16384));
          const uint64_t al = tc::smem_desc_sw128(tc::smem_u32(sA + (size_t)(nkt + kt) * 16384));
          const uint64_t bd = tc::smem_desc_sw128(tc::smem_u32(sB + (size_t)s * 16384));
          if (dbg & 1) {
            // (measurement: no MMAs)
          } else if (!PASS_B || (g & 1) == 0) {                    // B_hi tile: ah.bh (+ al.bh in pass B)
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16) tc::mma_bf16(tmem + acc * kTcTile, ah + 2 * k16, bd + 2 * k16, idesc, (g | k16) != 0);
            if (PASS_B) {
#pragma unroll
              for (int k16 = 0; k16 < 4; ++k16) tc::mma_bf16(tmem + acc * kTcTile, al + 2 * k16, bd + 2 * k16, idesc, 1);
            }
          } else {                                          // B_lo tile: ah.bl
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16) tc::mma_bf16(tmem + acc * kTcTile, ah + 2 * k16, bd + 2 * k16, idesc, 1);
          }
          if (PROF && prof) w_issue += clock64() - i_t;
          if (g == kPlanes * nkt - 1) {
            // norm slice of tile t: landed with full[] of the tile's first stage, which this thread waited on
            if (!(dbg & 1)) tc::mma_bf16(tmem + acc * kTcTile, axd, tc::smem_desc_nosw(tc::smem_u32(sBx + (size_t)(t & 1) * kExt), 128), idesc, 1);
          }
          if (CS == 1) tc::mma_commit(&empty[s]);   // smem stage reusable once these MMAs retire ...
          else tc::mma_commit_multicast(&empty[s], kMask);          // ... in every CTA of the cluster
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
        tc::mma_commit(&tfull[acc]);                // accumulator complete
      }
      if (PROF && prof) {
        long long* o = prof + 8 * (blockIdx.y * gridDim.x + blockIdx.x);
        o[0] = clock64() - m_t0, o[1] = w_full, o[2] = w_tempty, o[3] = w_a, o[7] = w_issue;
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
    const uint32_t lbeg = tc::smem_u32(cand + row * kTcListSm);      // list cursor as a 32-bit shared address
    uint32_t lptr = lbeg;
    if (!PASS_B) {
#pragma unroll
      for (int i = 0; i < 64; ++i) gmax[i] = -INFINITY;
    } else {
      my_thr = q < Nq ? thr[(size_t)b * Nq + q] : INFINITY;
    }
    long long w_tfull = 0;
    for (int t = 0; t < ntiles; ++t) {
      const int acc = t & 1;
      const long long e_t = (PROF && prof) ? clock64() : 0;
      tc::mbar_wait(&tfull[acc], (t >> 1) & 1);
      if (PROF && prof) w_tfull += clock64() - e_t;
      tc::tc_fence_after();
      if (!(dbg & 2))
#pragma unroll
      for (int c0 = 0; c0 < kTcTile; c0 += 64) {
        float v[64];
        tc::tmem_ld64(tmem + lane_base + acc * kTcTile + c0, v);
        if (!PASS_B) {
#pragma unroll
          for (int i = 0; i < 64; ++i) gmax[i] = fmaxf(gmax[i], v[i]);     // group = column mod 64
        } else {
          // branch-free compaction: every element is stored at the cursor, the cursor moves on a hit.  The cursor
          // is clamped every 8 elements (the row has 8 words of slack), so the list saturates at kTcCap.
          const int jbase = t * kTcTile + c0;
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            const float delta = v[i] - my_thr;
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(lptr), "r"(knn_pack(delta, jbase + i)) : "memory");
            lptr += delta >= 0.f ? 4u : 0u;
            if ((i & 7) == 7) lptr = min(lptr, lbeg + 4u * kTcCap);
          }
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&tempty[acc]);
    }
    if (PROF && prof && tid == 0) prof[8 * (blockIdx.y * gridDim.x + blockIdx.x) + 4] = w_tfull;
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
        const float bbm = __uint_as_float(bbmax_bits[b]);
        const float e = knn_margin(aa, bbm), ea = knn_margin_coarse(aa, bbm);
        // k candidates have pass-A score >= kth, hence exact d <= |a|^2 - 2 kth + e_A, so the k-th exact distance is at
        // most that; a member j of the exact answer then has pass-B score s_j >= (|a|^2 - d_j - e)/2 >= kth - (e_A + e)/2.
        // The min() covers k or more candidates sitting at (clamped) distance zero, e.g. duplicated points: then
        // everything with pass-B d~ <= e, i.e. s >= (|a|^2 - e)/2, is needed.
        thr[(size_t)b * Nq + q] = fminf(kth - 0.5f * (ea + e), 0.5f * (aa - e));
      }
    } else if (q < Nq) {
      cnt_out[(size_t)b * Nq + q] = (int)(min(lptr, lbeg + 4u * kTcCap) - lbeg) >> 2;   // == kTcCap: saturated
    }
  }

  if (PASS_B) {
    // candidate lists -> global, 512 B per row, coalesced
    __syncthreads();
    for (int row = warp; row < kTcRows; row += kTcThreads / 32) {
      const int q = q0 + row;
      if (q >= Nq) break;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        cand_out[((size_t)b * Nq + q) * kTcListLd + u * 32 + lane] = cand[row * kTcListSm + u * 32 + lane];
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (CS > 1) tc::cluster_sync();               // no CTA leaves while a peer can still arrive on its barriers
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// A row the tensor-core passes could not finish (list saturated: more than kTcCap candidates above the threshold, e.g. many
// near-duplicate points) is searched exactly by ITS OWN warp: every candidate's fp32 distance with the exact kernel's
// arithmetic (ascending channels, one accumulator, d2 = fma(-2, dot, |a|^2 + |b|^2)) through the running k-set.  Round 1
// sent such rows to the exact TILE kernel, which recomputes all 128 rows of the row's tile against every candidate on one
// SM: ~0.25 ms for a single flagged row -- the whole of the unexplained rank skew of the multi-GPU runs
// (profiles/r2_scaling.md: two of eight batch slices hold a few such rows).  Here it costs that warp ~40 us.
template <class I>
__device__ __forceinline__ void knn_exact_row(const float* __restrict__ arow, float aa, const float* __restrict__ B_g,
                                              const float* __restrict__ bnorm_g, int Nr, int Cp, int k, int lane,
                                              I* __restrict__ idx_row, float* __restrict__ dist_row, bool ordered) {
  const float4* ar = reinterpret_cast<const float4*>(arow);
  LaneTopK top;
  top.init(lane, k);
  for (int j0 = 0; j0 < Nr; j0 += 128) {                 // four candidates per lane and trip: independent FMA chains
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float4* br[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) br[u] = reinterpret_cast<const float4*>(B_g + (size_t)min(j0 + 32 * u + lane, Nr - 1) * Cp);
#pragma unroll 2
    for (int c4 = 0; c4 < Cp / 4; ++c4) {
      const float4 a4 = __ldg(ar + c4);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 b4 = __ldg(br[u] + c4);
        acc[u] = fmaf(a4.x, b4.x, acc[u]);
        acc[u] = fmaf(a4.y, b4.y, acc[u]);
        acc[u] = fmaf(a4.z, b4.z, acc[u]);
        acc[u] = fmaf(a4.w, b4.w, acc[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {                        // offered in ascending index order: LaneTopK's contract
      const int j = j0 + 32 * u + lane;
      const bool have = j < Nr;
      const float d2 = __fmaf_rn(-2.f, acc[u], __fadd_rn(aa, __ldg(bnorm_g + min(j, Nr - 1))));
      top.offer(dist_bits(d2), j, have);
    }
  }
  const int pos = ordered ? top.rank() : lane;
  if (top.active) {
    idx_row[pos] = (I)top.i;
    if (dist_row) dist_row[pos] = -sqrtf(__uint_as_float(top.d));
  }
}

// ---- pass C: exact fp32 re-rank.  One warp per query row, lanes across its candidates. ----
template <class I>
__global__ void __launch_bounds__(256) knn_rerank_kernel(const float* __restrict__ an, const float* __restrict__ anorm,
                                                         const float* __restrict__ bn, const float* __restrict__ bnorm,
                                                         const uint32_t* __restrict__ cand,
                                                         const int* __restrict__ cnt_in, int Nq, int Nr, int Cp, int k,
                                                         I* __restrict__ idx_out, float* __restrict__ dist_out,
                                                         int* __restrict__ row_flags) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, q = blockIdx.x * 8 + warp;
  if (q >= Nq) return;
  const size_t rowi = (size_t)b * Nq + q;
  const int cnt = cnt_in[rowi];
  const float aa = anorm[rowi];
  const float4* ar = reinterpret_cast<const float4*>(an + rowi * Cp);
  const float* B_g = bn + (size_t)b * Nr * Cp;
  const float* bnorm_g = bnorm + (size_t)b * Nr;
  if (cnt >= kTcCap) {                                  // saturated list: this warp searches the row exactly
    knn_exact_row<I>(an + rowi * Cp, aa, B_g, bnorm_g, Nr, Cp, k, lane, idx_out + rowi * k, dist_out ? dist_out + rowi * k : nullptr, true);
    return;
  }
  LaneTopK top;
  top.init(lane, k);
  for (int r = 0; r < cnt; r += 32) {
    const int e = r + lane;
    const bool have = e < cnt;
    const int j = have ? (int)(cand[rowi * kTcListLd + e] & 0xffffu) : 0;
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

// ---- pass C', indices only, any order: exact work only where the approximate scores cannot decide. ----
// With |d~ - d| <= e for every candidate (knn_margin) and d~(k), d~(k+1) the k-th / (k+1)-th smallest approx distances
// of the row (both are in the list, which holds everything up to d~(k) + 2e):
//   d~_j < d~(k+1) - 2e  and  d~(k+1) > e   =>  fewer than k candidates can beat j  -> j is in the exact answer
//   d~_j > d~(k)   + 2e  and  d~_j    > e   =>  at least k candidates beat j         -> j is not
// (the "> e" clauses keep the argument valid under the exact kernel's clamp of negative d^2 to 0, where
// near-duplicates of the query tie and the index decides).  Only the band in between -- a handful of candidates --
// gets exact fp32 distances, and the best (k - #in) of it by (distance, index) completes the SET the exact kernel
// returns.  In s = <a,b> - |b|^2/2 units (d~ = |a|^2 - 2s):  in: s_j > s(k+1) + e,  out: s_j < s(k) - e, evaluated
// on the list's truncated offsets delta = s - T with their interval [delta_lo, delta_hi] (knn_pack).
// Classification step of knn_select_kernel for a list of up to 32*NPL entries (element p = u*32 + lane):
// bitonic-sort the packed words (descending score), read the k-th and (k+1)-th, then split the list into
// certain-in (written to idx_row), certain-out (dropped) and ambiguous (compacted into amb[], list order kept).
template <int NPL, class I>
__device__ __forceinline__ void select_classify(const uint32_t* __restrict__ list, int cnt, int k, float aa, float e,
                                                float thr_row, int lane, I* __restrict__ idx_row,
                                                unsigned short* __restrict__ amb, int& n_in, int& n_amb) {
  constexpr int kN = 32 * NPL;
  uint32_t w[NPL], x[NPL];
#pragma unroll
  for (int u = 0; u < NPL; ++u) {
    const int t = u * 32 + lane;
    w[u] = t < cnt ? __ldg(list + t) : 0u;
    x[u] = ~w[u];                                       // ascending x = descending score; padding (all ones) last
  }
#pragma unroll
  for (int size = 2; size <= kN; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride >= 1; stride >>= 1) {
      if (stride >= 32) {
        const int du = stride >> 5;
#pragma unroll
        for (int u = 0; u < NPL; ++u) {
          if ((u & du) == 0) {
            const bool up = ((u * 32) & size) == 0;
            const uint32_t lo = min(x[u], x[u | du]), hi = max(x[u], x[u | du]);
            x[u] = up ? lo : hi;
            x[u | du] = up ? hi : lo;
          }
        }
      } else {
#pragma unroll
        for (int u = 0; u < NPL; ++u) {
          const uint32_t other = __shfl_xor_sync(kFull, x[u], stride);
          const bool up = ((u * 32 + lane) & size) == 0;
          const bool lower = (lane & stride) == 0;
          x[u] = (lower == up) ? min(x[u], other) : max(x[u], other);
        }
      }
    }
  }
  uint32_t xk = 0xffffffffu, xk1 = 0xffffffffu;         // sorted[k-1], sorted[k]  (k <= 32 <= kN)
#pragma unroll
  for (int u = 0; u < NPL; ++u) {
    const uint32_t a0 = __shfl_sync(kFull, x[u], (k - 1) & 31);
    const uint32_t a1 = __shfl_sync(kFull, x[u], k & 31);
    if (((k - 1) >> 5) == u) xk = a0;
    if ((k >> 5) == u) xk1 = a1;
  }
  const float dk = knn_delta_lo(~xk);                                   // k-th largest offset (lower end)
  const float dk1 = cnt > k ? knn_delta_lo(~xk1) : -INFINITY;           // (k+1)-th; none if the list holds exactly k
  // offsets at or above n0 may belong to d~ <= e, where the exact kernel's clamp makes the index decide
  const float n0 = (0.5f * (aa - e) - thr_row) * 0.999f;
  const bool in_ok = knn_delta_hi(dk1) < n0 || cnt == k;
  const float in_thr = knn_delta_hi(dk1) + e;           // in:  delta_lo_j > this
  const float out_thr = dk - e;                         // out: delta_hi_j < this
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int u = 0; u < NPL; ++u) {
    const bool have = u * 32 + lane < cnt;
    const float dlo = knn_delta_lo(w[u]), dhi = knn_delta_hi(dlo);
    const int j = (int)(w[u] & 0xffffu);
    const bool is_in = have && in_ok && dlo > in_thr;
    const bool is_out = have && dhi < out_thr && dhi < n0;
    const bool is_amb = have && !is_in && !is_out;
    const unsigned m_in = __ballot_sync(kFull, is_in), m_amb = __ballot_sync(kFull, is_amb);
    if (is_in) idx_row[n_in + __popc(m_in & lt)] = (I)j;
    if (is_amb) amb[n_amb + __popc(m_amb & lt)] = (unsigned short)j;
    n_in += __popc(m_in);
    n_amb += __popc(m_amb);
  }
}

constexpr int kSelBatch = 4;          // candidate rows fetched per round
constexpr int kSelRowLd = 132;        // floats per parked row (pad 4: the 8 evaluating lanes hit distinct banks)

template <class I>
__global__ void __launch_bounds__(256, 5) knn_select_kernel(const float* __restrict__ an, const float* __restrict__ anorm,
                                                         const float* __restrict__ bn, const float* __restrict__ bnorm,
                                                         const unsigned* __restrict__ bbmax_bits,
                                                         const uint32_t* __restrict__ cand,
                                                         const float* __restrict__ thr_in, const int* __restrict__ cnt_in,
                                                         int Nq, int Nr, int Cp, int k, I* __restrict__ idx_out,
                                                         int* __restrict__ row_flags) {
  __shared__ unsigned short amb[8][128];
  __shared__ __align__(16) float s_arow[8][128];                       // query row
  __shared__ __align__(16) float s_rows[8][kSelBatch * kSelRowLd];     // candidate rows of the current round
  __shared__ unsigned long long s_keys[8][128];                        // (distance bits, index) of the ambiguous ones
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, q = blockIdx.x * 8 + warp;
  if (q >= Nq) return;
  const size_t rowi = (size_t)b * Nq + q;
  const int cnt = cnt_in[rowi];
  const float aa = anorm[rowi];
  if (cnt >= kTcCap) {                                  // saturated list: this warp searches the row exactly
    knn_exact_row<I>(an + rowi * Cp, aa, bn + (size_t)b * Nr * Cp, bnorm + (size_t)b * Nr, Nr, Cp, k, lane, idx_out + rowi * k, nullptr, false);
    return;
  }
  const float e = knn_margin(aa, __uint_as_float(bbmax_bits[b])) * 1.001f;
  int n_in = 0, n_amb = 0;
  const uint32_t* list = cand + rowi * kTcListLd;
  const float thr_row = __ldg(thr_in + rowi);
  // the list is sorted only to read off its k-th and (k+1)-th scores: 1, 2 or 4 entries per lane as its length needs
  if (cnt <= 32)
    select_classify<1, I>(list, cnt, k, aa, e, thr_row, lane, idx_out + rowi * k, amb[warp], n_in, n_amb);
  else if (cnt <= 64)
    select_classify<2, I>(list, cnt, k, aa, e, thr_row, lane, idx_out + rowi * k, amb[warp], n_in, n_amb);
  else
    select_classify<4, I>(list, cnt, k, aa, e, thr_row, lane, idx_out + rowi * k, amb[warp], n_in, n_amb);
  const int need = k - n_in;
  if (need <= 0) return;                                // (n_in <= k always: an "in" entry has at most k-1 rivals)
  if (n_amb < need) {                                   // cannot happen while the margin holds; stay safe
    __syncwarp();
    knn_exact_row<I>(an + rowi * Cp, aa, bn + (size_t)b * Nr * Cp, bnorm + (size_t)b * Nr, Nr, Cp, k, lane, idx_out + rowi * k, nullptr, false);
    return;
  }
  __syncwarp();
  // Exact fp32 distances of the ambiguous candidates.  Rows are fetched by the whole warp (lane = 16-byte chunk, up to
  // 8 rows in flight per round: coalesced 512-byte reads instead of one serial gather per lane) and parked in shared
  // memory; lane u then runs candidate u's dot product there -- ascending channels, one accumulator, i.e. exactly the
  // arithmetic of the exact kernel in knn.cu.
  const int nch = Cp >> 2;
  const float4* ar = reinterpret_cast<const float4*>(an + rowi * Cp);
  const float* B_g = bn + (size_t)b * Nr * Cp;
  const float* bnorm_g = bnorm + (size_t)b * Nr;
  if (lane < nch) reinterpret_cast<float4*>(s_arow[warp])[lane] = __ldg(ar + lane);
  for (int r = 0; r < n_amb; r += kSelBatch) {
    const int nb = min(kSelBatch, n_amb - r);
    float4 tmp[kSelBatch];
#pragma unroll
    for (int u = 0; u < kSelBatch; ++u)
      if (u < nb && lane < nch) tmp[u] = __ldg(reinterpret_cast<const float4*>(B_g + (size_t)amb[warp][r + u] * Cp) + lane);
#pragma unroll
    for (int u = 0; u < kSelBatch; ++u)
      if (u < nb && lane < nch) reinterpret_cast<float4*>(s_rows[warp] + u * kSelRowLd)[lane] = tmp[u];
    __syncwarp();
    if (lane < nb) {
      const int j = (int)amb[warp][r + lane];
      const float4* br = reinterpret_cast<const float4*>(s_rows[warp] + lane * kSelRowLd);
      const float4* aq = reinterpret_cast<const float4*>(s_arow[warp]);
      float acc = 0.f;
#pragma unroll 4
      for (int c4 = 0; c4 < nch; ++c4) {
        const float4 a4 = aq[c4];
        const float4 b4 = br[c4];
        acc = fmaf(a4.x, b4.x, acc);
        acc = fmaf(a4.y, b4.y, acc);
        acc = fmaf(a4.z, b4.z, acc);
        acc = fmaf(a4.w, b4.w, acc);
      }
      const float d2 = __fmaf_rn(-2.f, acc, __fadd_rn(aa, __ldg(bnorm_g + j)));
      s_keys[warp][r + lane] = ((unsigned long long)dist_bits(d2) << 32) | (unsigned)j;
    }
    __syncwarp();
  }
  // the `need` smallest (distance, index) keys, by rank counting (keys are distinct: the indices are)
  for (int t = lane; t < n_amb; t += 32) {
    const unsigned long long mine = s_keys[warp][t];
    int rank = 0;
    for (int u = 0; u < n_amb; ++u) rank += s_keys[warp][u] < mine ? 1 : 0;
    if (rank < need) idx_out[rowi * k + n_in + rank] = (I)(unsigned)(mine & 0xffffffffu);
  }
}

// measurement switches (tools/probe_knn.py): 1 = no MMAs, 2 = idle epilogue.  Results are garbage while set.
int g_knn_tc_debug = 0;
// CTAs per cluster: -1 = automatic (2 when the number of query tiles is even), 1 / 2 / 4 forced (4 needs a multiple of 4 tiles)
int g_knn_cluster = -1;
// measurement (tools/probe_knn_roles.py): per CTA [MMA thread total, wait operands, wait accumulator drain, wait query tile, epilogue thread's wait for
// accumulators, producer total, producer's wait for free stages, MMA thread's descriptor + MMA issue time] cycles; written by the NEXT tensor-core launches while non-null
long long* g_knn_prof = nullptr;

// ---- host side ----
size_t knn_tc_ext_floats(int B, int Nr);
bool knn_tc_eligible(int Nq, int Nr, int C, int k) {
  const int Cp = (int)align_up(C, 32);
  return C >= 4 && Cp <= 128 && k <= 32 && Nr <= 65535 && Nr >= k;
}

size_t knn_tc_workspace_bytes(int B, int Nq, int Nr) {
  return align_up((size_t)B * Nq * kTcListLd * sizeof(uint32_t), 256) + align_up((size_t)B * Nq * sizeof(int), 256) +
         align_up(knn_tc_ext_floats(B, Nr) * sizeof(float), 256);
}
// candidate-side norm slices, one 4 KB block per (cloud, tile of 128 candidates), already in the compact
// no-swizzle operand layout (tc::nosw_offset) so that the kernel fetches a tile's slice with one bulk copy
size_t knn_tc_ext_floats(int B, int Nr) { return (size_t)B * ceil_div(Nr, kTcTile) * (kExt / 4); }

template <class I>
int launch_knn_tc(const float* an, const float* anorm, const float* bn, const float* bnorm, const __nv_bfloat16* a_hi,
                  const __nv_bfloat16* a_lo, const __nv_bfloat16* b_hi, const __nv_bfloat16* b_lo, const float* bext,
                  const unsigned* bbmax, int B, int Nq, int Nr, int Cp, int k, float* thr, uint32_t* cand, int* cnt, bool ordered,
                  I* idx, float* dist, int* row_flags, cudaStream_t st) {
  const int nkb = (Cp + 63) / 64;                           // K-tiles of 64 bf16 per plane
  alignas(64) CUtensorMap map_a, map_al, map_b, map_bl;
  if (int e = make_tile_map(&map_a, a_hi, Cp, Cp, Nq, B, kTcRows, 2)) return e;
  if (int e = make_tile_map(&map_al, a_lo, Cp, Cp, Nq, B, kTcRows, 2)) return e;
  if (int e = make_tile_map(&map_b, b_hi, Cp, Cp, Nr, B, kTcTile, 2)) return e;
  if (int e = make_tile_map(&map_bl, b_lo, Cp, Cp, Nr, B, kTcTile, 2)) return e;
  dim3 grid(ceil_div(Nq, kTcRows), B);
  const int cs = g_knn_cluster > 0 ? g_knn_cluster : 1;     // clusters sharing the candidate stream: measured no gain (the stream is latency-, not bandwidth-bound), kept switchable
  {
    size_t smem = tc_smem_bytes(nkb, false);
    auto kern = cs == 2 ? knn_tc_kernel<false, 2> : (cs == 4 ? knn_tc_kernel<false, 4> : (g_knn_prof ? knn_tc_kernel<false, 1, true> : knn_tc_kernel<false, 1>));
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return check_launch("knn_tc pass A smem attribute");
    SAMBLE_PRE(st);
    kern<<<grid, kTcThreads, smem, st>>>(map_a, map_al, map_b, map_bl, anorm, bext, bbmax, Nq, Nr, Cp, k, thr, cand, cnt, g_knn_tc_debug, g_knn_prof);
    SAMBLE_LAUNCHED("knn_tc_threshold_kernel");
  }
  {
    size_t smem = tc_smem_bytes(nkb, true);
    auto kern = cs == 2 ? knn_tc_kernel<true, 2> : (cs == 4 ? knn_tc_kernel<true, 4> : (g_knn_prof ? knn_tc_kernel<true, 1, true> : knn_tc_kernel<true, 1>));
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return check_launch("knn_tc pass B smem attribute");
    SAMBLE_PRE(st);
    kern<<<grid, kTcThreads, smem, st>>>(map_a, map_al, map_b, map_bl, anorm, bext, bbmax, Nq, Nr, Cp, k, thr, cand, cnt, g_knn_tc_debug, g_knn_prof);
    SAMBLE_LAUNCHED("knn_tc_collect_kernel");
  }
  SAMBLE_PRE(st);
  if (ordered) {
    knn_rerank_kernel<I><<<dim3(ceil_div(Nq, 8), B), 256, 0, st>>>(an, anorm, bn, bnorm, cand, cnt, Nq, Nr, Cp, k, idx, dist, row_flags);
    SAMBLE_LAUNCHED("knn_rerank_kernel");
  } else {
    knn_select_kernel<I><<<dim3(ceil_div(Nq, 8), B), 256, 0, st>>>(an, anorm, bn, bnorm, bbmax, cand, thr, cnt, Nq, Nr, Cp, k, idx, row_flags);
    SAMBLE_LAUNCHED("knn_select_kernel");
  }
  return SAMBLE_OK;
}

template int launch_knn_tc<int>(const float*, const float*, const float*, const float*, const __nv_bfloat16*, const __nv_bfloat16*,
                                const __nv_bfloat16*, const __nv_bfloat16*, const float*, const unsigned*, int, int, int, int, int,
                                float*, uint32_t*, int*, bool, int*, float*, int*, cudaStream_t);
template int launch_knn_tc<long long>(const float*, const float*, const float*, const float*, const __nv_bfloat16*,
                                      const __nv_bfloat16*, const __nv_bfloat16*, const __nv_bfloat16*, const float*, const unsigned*,
                                      int, int, int, int, int, float*, uint32_t*, int*, bool, long long*, float*, int*,
                                      cudaStream_t);

}  // namespace samble

extern "C" void samble_set_knn_probe(long long* cycles) { samble::g_knn_prof = cycles; }
extern "C" void samble_set_knn_debug(int bits) { samble::g_knn_tc_debug = bits & 0xff; samble::g_knn_cluster = (bits >> 8) ? (bits >> 8) : -1; }
