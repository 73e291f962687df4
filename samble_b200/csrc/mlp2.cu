// Two point-wise linear layers back to back, the hidden activation never leaves the SM:
//   out = epi2( lrelu( (X W1^T) * scale1 + shift1 ) W2^T )
// Neighbor2PointAttention's feed-forward (models/attention.py:187-192: Conv1d C->4C, LeakyReLU, Conv1d 4C->C, then
// bn2(x + ff(x))) and the segmentation head's conv2 -> conv3 (models/seg_model.py:205-214).  As two linear_tma launches the
// (M x Hd) hidden tensor is written to and re-read from memory (436 MB per seg step for the feed-forwards, 268 MB for the
// head) and each launch pays its own operand stream and epilogue.
//
// One CTA owns a 128-row tile of X (resident in shared memory, hi + lo planes) and walks the hidden width in chunks of 128:
//   G1  acc1[128 x 128]  = X W1_j^T                  (3xTF32, both operands from shared memory, like linear_tma.cu)
//   cv  eight warps turn acc1 into the A operand of the next product IN TENSOR MEMORY: scale/shift/LeakyReLU, the value
//       goes back to its own columns (the tensor core reads its top 19 bits = hi), hi's remainder `lo` to the next 128
//   G2  Y[128 x N2]     += [hi|lo] W2_j^T            (A operand read from TMEM: tcgen05.mma [d], [a], b-desc)
// so per chunk only the two weight slices stream through a 3-stage TMA ring; X is loaded once per tile and the hidden
// activation exists only as TMEM columns.
//   N2 = 128: [0,256) two acc1/hi buffers, [256,384) lo, [384,512) Y.  The first product runs one chunk ahead of the
//             second (G1_0 G1_1 G2_0 G1_2 G2_1 ...): chunk j is converted while the tensor pipe executes G1_{j+1}.
//   N2 = 256: [0,128) acc1/hi, [128,256) lo, [256,512) Y: no room for a second buffer, G1_j -> conversion -> G2_j in turn.
// Y is ONE accumulation chain (Hd / 32 K-blocks) in both layouts.
//
// Warps 0-7 conversion + output epilogue (warp & 3 = TMEM lane quadrant, warp >> 2 = column half), 8 MMA issuer, 9 weight
// producer, 10-11 X loader / splitters (X_lo = X - trunc_tf32(X)).  12 warps: 168 registers per thread (no spills without first-layer
// scale / shift; ~0.5 KB of spill traffic in the CONSTS instantiations, whose 64 constants are prefetched across the accumulator wait).
#include "linear_common.cuh"

namespace samble {

constexpr int kM2Threads = 12 * 32;
constexpr int kM2Stages = 3;
// X hi/lo (128 KB) + weight ring (96 KB) + barriers (256 B) + the second layer's scale / shift (2 * N2 floats).  No alignment
// slack: the dynamic window of a kernel without static shared memory starts 1024-byte aligned (checked, traps otherwise).
constexpr size_t m2_smem(int n2) { return 8 * 16384 + kM2Stages * 32768 + 256 + 2 * (size_t)n2 * 4; }

long long* g_m2_wait_cycles = nullptr;
int g_m2_debug = 0;   // measurement switches (tools/probe_mlp2.py): 1 = no conversion math / stores, 2 = no output epilogue, 4 = no MMAs

struct Mlp2Args {
  LinArgs l2;                 // the second layer's epilogue: scale, shift, residual, out, M, Nout = N2, npc, lrelu, res_first
  const float* scale1;        // [Hd] or null
  const float* shift1;        // [Hd] (+ cloud * shift1_ldb) or null
  long long shift1_ldb;
  int lrelu1, Hd, nkb1, dbg;
  long long* wait_cycles;     // measurement (tools/probe_mlp2.py): per CTA [total, wfull, hready, yempty, xready] cycles of the MMA thread
};

// LA = 1 (N2 = 128): the first product runs one chunk AHEAD of the second (two acc1/hi buffers, G1_{j+1} is issued before
// G2_j), so the conversion of chunk j happens while the tensor pipe executes G1_{j+1}; LA = 0 (N2 = 256, no TMEM left for a
// second buffer): G1_j, conversion, G2_j in turn.
template <int N2, int LA, bool CONSTS, bool PROF = false>
__global__ void __launch_bounds__(kM2Threads, 1)
    mlp2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
                const __grid_constant__ CUtensorMap map_w1lo, const __grid_constant__ CUtensorMap map_w2,
                const __grid_constant__ CUtensorMap map_w2lo, const Mlp2Args a) {
  constexpr int NH = N2 / 128;              // 128-column halves of the output
  constexpr int NB = LA + 1;                // acc1 / hi buffers
  constexpr uint32_t kLoCol = NB * 128, kYCol = kLoCol + 128;
  constexpr int NACC2 = (512 - (int)kYCol) / N2;      // accumulation chains of the second product that fit in TMEM
  static_assert(NACC2 >= 1, "TMEM budget");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw;
  if ((tc::smem_u32(base) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("samble: mlp2 shared-memory window is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* xhi = base;                      // [4 K-blocks][128 rows x 128 B]
  uint8_t* xlo = base + 4 * 16384;
  uint8_t* ring = base + 8 * 16384;         // [stage][hi 16 KB | lo 16 KB] one K-block of a 128-row weight slice
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + kM2Stages * 32768);
  uint64_t* xland = bars;                   // [4] X K-block landed
  uint64_t* xready = bars + 4;              // [4] ... and its lo plane written (2 splitter warps)
  uint64_t* xfree = bars + 8;               // every MMA reading this tile's X retired
  uint64_t* wfull = bars + 9;               // [3]
  uint64_t* wempty = bars + 12;             // [3]
  uint64_t* hfull = bars + 15;              // [2] acc1 buffer complete
  uint64_t* hready = bars + 17;             // [2 buffers][4] 32-column block of [hi|lo] written (4 warps each): = one K-block of G2
  uint64_t* g2done = bars + 25;             // second product of a chunk retired: the lo plane may be rewritten (LA)
  uint64_t* yfull = bars + 26;              // Y of the tile complete
  uint64_t* yempty = bars + 27;             // Y drained (8 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);
  float* s_scale2 = reinterpret_cast<float*>(bars + 32);       // [N2] second layer's per-column scale (1 when absent)
  float* s_shift2 = s_scale2 + N2;                             // [N2] ... and shift (0 when absent or per cloud)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mtiles = (a.l2.M + 127) / 128;
  const int nchunks = a.Hd / 128;
  // chunks per accumulation chain: 2 (= the 8-K-block chains of linear_tma.cu) whenever the chains fit in TMEM
  const int cpc = (nchunks + NACC2 - 1) / NACC2 > 2 ? (nchunks + NACC2 - 1) / NACC2 : 2;
  const int nkb1 = a.nkb1;

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      tc::mbar_init(&xland[i], 1);
      tc::mbar_init(&xready[i], 2);
    }
    for (int i = 0; i < 8; ++i) tc::mbar_init(&hready[i], 4);
    tc::mbar_init(xfree, 1);
    for (int i = 0; i < kM2Stages; ++i) {
      tc::mbar_init(&wfull[i], 1);
      tc::mbar_init(&wempty[i], 1);
    }
    tc::mbar_init(&hfull[0], 1);
    tc::mbar_init(&hfull[1], 1);
    tc::mbar_init(g2done, 1);
    tc::mbar_init(yfull, 1);
    tc::mbar_init(yempty, 8);
    tc::mbar_init_fence();
  }
  for (int c = tid; c < N2; c += kM2Threads) {
    s_scale2[c] = a.l2.scale ? __ldg(a.l2.scale + c) : 1.f;
    s_shift2[c] = (a.l2.shift && a.l2.shift_ldb == 0) ? __ldg(a.l2.shift + c) : 0.f;
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp >= 10) {
    // ================= X loader + splitters =================
    const int lt = tid - 320;
    int it = 0;
    for (int tile = blockIdx.x; tile < mtiles; tile += gridDim.x, ++it) {
      if (warp == 10) {
        if (tc::elect_one()) {
          if (it == 0) tc::tma_prefetch_desc(&map_x);
          tc::mbar_wait(xfree, (it & 1) ^ 1);                   // the previous tile's first products are done with X
          for (int kb = 0; kb < nkb1; ++kb) {
            tc::mbar_arrive_expect_tx(&xland[kb], 16384u);
            tc::tma_load_3d(xhi + kb * 16384, &map_x, &xland[kb], kb * 32, tile * 128, 0);
          }
        }
        __syncwarp();
      }
      for (int kb = 0; kb < nkb1; ++kb) {
        tc::mbar_wait(&xland[kb], it & 1);
        const uint8_t* h = xhi + kb * 16384;
        uint8_t* l = xlo + kb * 16384;
#pragma unroll 8
        for (int i = 0; i < 16; ++i) {
          const uint32_t off = (uint32_t)(lt + 64 * i) * 16u;
          const float4 v = *reinterpret_cast<const float4*>(h + off);
          float4 lo;
          lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
          lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
          lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
          lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
          *reinterpret_cast<float4*>(l + off) = lo;
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&xready[kb]);
      }
    }
  } else if (warp == 9) {
    // ================= weight producer: the stages in the order the MMA thread consumes them =================
    if (tc::elect_one()) {
      tc::tma_prefetch_desc(&map_w1);
      tc::tma_prefetch_desc(&map_w1lo);
      tc::tma_prefetch_desc(&map_w2);
      tc::tma_prefetch_desc(&map_w2lo);
      int s = 0, ph = 0;
      auto stage = [&](const CUtensorMap* mh, const CUtensorMap* ml, int c0, int c1) {
        tc::mbar_wait(&wempty[s], ph ^ 1);
        uint8_t* st = ring + s * 32768;
        tc::mbar_arrive_expect_tx(&wfull[s], 32768u);
        tc::tma_load_3d(st, mh, &wfull[s], c0, c1, 0);
        tc::tma_load_3d(st + 16384, ml, &wfull[s], c0, c1, 0);
        if (++s == kM2Stages) { s = 0; ph ^= 1; }
      };
      auto g1 = [&](int j) {
        for (int kb = 0; kb < nkb1; ++kb) stage(&map_w1, &map_w1lo, kb * 32, j * 128);
      };
      auto g2 = [&](int j) {
        for (int h = 0; h < NH; ++h)
          for (int kb = 0; kb < 4; ++kb) stage(&map_w2, &map_w2lo, j * 128 + kb * 32, h * 128);
      };
      for (int tile = blockIdx.x; tile < mtiles; tile += gridDim.x) {
        if (LA) g1(0);
        for (int j = 0; j < nchunks; ++j) {
          if (!LA) g1(j);
          else if (j + 1 < nchunks) g1(j + 1);
          g2(j);
        }
      }
    }
    __syncwarp();
  } else if (warp == 8) {
    // ================= MMA issuer =================
    if (tc::elect_one()) {
      const uint32_t idesc = tc::instr_desc(2, 128, 128);
      const uint32_t xh0 = tc::smem_desc_sw128_lo(tc::smem_u32(xhi)), xl0 = tc::smem_desc_sw128_lo(tc::smem_u32(xlo));
      const uint32_t r0 = tc::smem_desc_sw128_lo(tc::smem_u32(ring));
      const bool mma_on = !(a.dbg & 4);
      int s = 0, ph = 0, it = 0, c1 = 0, c2 = 0;               // c1 / c2: chunks whose first / second product has been issued
      long long tw[4] = {0, 0, 0, 0};
      // (cycle counters only in the PROF instantiation, tools/probe_mlp2.py: the production issue loop carries no clock reads)
      const bool prof = PROF && a.wait_cycles != nullptr;
      const long long t_start = PROF ? clock64() : 0;
#define M2_WAIT(slot, ...)                                   \
  do {                                                       \
    const long long _t = (PROF && prof) ? clock64() : 0;     \
    __VA_ARGS__;                                             \
    if (PROF && prof) tw[slot] += clock64() - _t;            \
  } while (0)
      // G1: acc1[buffer] = X W1_j^T.  The tensor pipe executes in issue order, so the second product that last read this
      // buffer as its A operand (issued earlier) is done with it before these MMAs overwrite it.
      auto g1 = [&](int j) {
        const uint32_t acc1 = tmem + (c1 % NB) * 128;
        for (int kb = 0; kb < nkb1; ++kb) {
          if (j == 0) M2_WAIT(3, tc::mbar_wait(&xready[kb], it & 1));
          M2_WAIT(0, tc::mbar_wait(&wfull[s], ph));
          tc::tc_fence_after();
          const uint32_t xh = xh0 + kb * (16384 >> 4), xl = xl0 + kb * (16384 >> 4);
          const uint32_t wh = r0 + s * (32768 >> 4), wl = wh + (16384 >> 4);
          if (mma_on)
#pragma unroll
          for (int k8 = 0; k8 < 4; ++k8) {
            tc::mma_tf32_lo(acc1, xh + 2 * k8, wh + 2 * k8, idesc, (kb | k8) != 0);
            tc::mma_tf32_lo(acc1, xl + 2 * k8, wh + 2 * k8, idesc, 1);
            tc::mma_tf32_lo(acc1, xh + 2 * k8, wl + 2 * k8, idesc, 1);
          }
          tc::mma_commit(&wempty[s]);
          if (++s == kM2Stages) { s = 0; ph ^= 1; }
        }
        tc::mma_commit(&hfull[c1 % NB]);
        if (j == nchunks - 1) tc::mma_commit(xfree);
        ++c1;
      };
      // G2: Y (+)= [hi|lo] W2_j^T, A operand from tensor memory
      auto g2 = [&](int j) {
        if (j == 0) {
          M2_WAIT(2, tc::mbar_wait(yempty, (it & 1) ^ 1));
          tc::tc_fence_after();
        }
        const int buf = c2 % NB;
        const uint32_t hi = tmem + buf * 128, lo = tmem + kLoCol;
        const uint32_t acc = tmem + kYCol + (j / cpc) * N2;
        const bool first = (j % cpc) == 0;
        for (int h = 0; h < NH; ++h)
          for (int kb = 0; kb < 4; ++kb) {
            if (h == 0) {                       // K-block kb of this chunk = 32-column block kb of the converted accumulator
              M2_WAIT(1, tc::mbar_wait(&hready[buf * 4 + kb], (c2 / NB) & 1));
              tc::tc_fence_after();
            }
            M2_WAIT(0, tc::mbar_wait(&wfull[s], ph));
            tc::tc_fence_after();
            const uint32_t wh = r0 + s * (32768 >> 4), wl = wh + (16384 >> 4);
            if (mma_on)
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8) {
              const uint32_t col = kb * 32 + k8 * 8;
              tc::mma_tf32_ts(acc + h * 128, hi + col, wh + 2 * k8, idesc, !(first && kb == 0 && k8 == 0));
              tc::mma_tf32_ts(acc + h * 128, lo + col, wh + 2 * k8, idesc, 1);
              tc::mma_tf32_ts(acc + h * 128, hi + col, wl + 2 * k8, idesc, 1);
            }
            tc::mma_commit(&wempty[s]);
            if (++s == kM2Stages) { s = 0; ph ^= 1; }
          }
        if (LA) tc::mma_commit(g2done);
        ++c2;
      };
      for (int tile = blockIdx.x; tile < mtiles; tile += gridDim.x, ++it) {
        if (LA) g1(0);
        for (int j = 0; j < nchunks; ++j) {
          if (!LA) g1(j);
          else if (j + 1 < nchunks) g1(j + 1);
          g2(j);
        }
        tc::mma_commit(yfull);
      }
#undef M2_WAIT
      if (PROF && prof) {
        long long* o = a.wait_cycles + 5 * blockIdx.x;
        o[0] = clock64() - t_start, o[1] = tw[0], o[2] = tw[1], o[3] = tw[2], o[4] = tw[3];
      }
    }
    __syncwarp();
  } else {
    // ================= conversion (acc1 -> [hi|lo] A operand) + output epilogue =================
    const int q = warp & 3, half = warp >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const bool consts = CONSTS && !(a.dbg & 8);     // CONSTS: scale1 or shift1 present (their 64 registers exist only then)
    int it = 0, cc = 0;
    for (int tile = blockIdx.x; tile < mtiles; tile += gridDim.x, ++it) {
      const int m0 = tile * 128;
      const float* shift1 = a.shift1 ? a.shift1 + (a.shift1_ldb ? (long long)(m0 / a.l2.npc) * a.shift1_ldb : 0) : nullptr;
      for (int j = 0; j < nchunks; ++j, ++cc) {
        const int buf = cc % NB;
        const uint32_t hi = tmem + lane_base + buf * 128, lo_t = tmem + lane_base + kLoCol;
        // per-column constants of this warp's first 32 columns: requested BEFORE the wait, so their latency is hidden
        float sc[32], sh[32];
        auto load_consts = [&](int hc) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f), h4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (consts && a.scale1) s4 = __ldg(reinterpret_cast<const float4*>(a.scale1 + hc + i));
            if (consts && shift1) h4 = __ldg(reinterpret_cast<const float4*>(shift1 + hc + i));
            sc[i] = s4.x, sc[i + 1] = s4.y, sc[i + 2] = s4.z, sc[i + 3] = s4.w;
            sh[i] = h4.x, sh[i + 1] = h4.y, sh[i + 2] = h4.z, sh[i + 3] = h4.w;
          }
        };
        if (consts) load_consts(j * 128 + half * 64);
        tc::mbar_wait(&hfull[buf], (cc / NB) & 1);
        tc::tc_fence_after();
        if (!(a.dbg & 1))
#pragma unroll 1
        for (int r = 0; r < 2; ++r) {
          const int c0 = half * 64 + r * 32;
          float v[32];
          tc::tmem_ld32(hi + c0, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float y = v[i];
            if (consts && a.scale1) y *= sc[i];
            if (consts && shift1) y += sh[i];
            if (a.lrelu1) y = y > 0.f ? y : 0.2f * y;
            v[i] = y;
          }
          if (consts && r == 0) load_consts(j * 128 + half * 64 + 32);     // in flight during the stores below
          tc::tmem_st32(hi + c0, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] -= __uint_as_float(__float_as_uint(v[i]) & 0xffffe000u);     // hi's remainder
          if (LA && r == 0 && cc > 0) {
            // the single lo plane is still the A operand of the previous chunk's second product until that retires
            tc::mbar_wait(g2done, (cc - 1) & 1);
            tc::tc_fence_after();
          }
          tc::tmem_st32(lo_t + c0, v);
          // each 32-column block is published on its own: the second product starts on K-block 0 while the rest converts
          tc::tmem_st_wait();
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&hready[buf * 4 + half * 2 + r]);
        }
        else {
          if (LA && cc > 0) tc::mbar_wait(g2done, (cc - 1) & 1);
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&hready[buf * 4 + half * 2]), tc::mbar_arrive(&hready[buf * 4 + half * 2 + 1]);
        }
      }
      // ---- output epilogue: this warp's 32 rows x N2/2 columns.  Scale / shift come from shared memory; the residual
      //      block of the NEXT 32 columns is requested before the current one is finished (the first before the wait).
      const int m = m0 + q * 32 + lane;
      const bool live = m < a.l2.M;
      const float* rrow = a.l2.residual ? a.l2.residual + (long long)(live ? m : 0) * a.l2.ldr : nullptr;
      const float* cshift = (a.l2.shift && a.l2.shift_ldb) ? a.l2.shift + (long long)(m0 / a.l2.npc) * a.l2.shift_ldb : nullptr;
      float* orow = a.l2.out + (long long)(live ? m : 0) * a.l2.ldo;
      const bool vec = a.l2.ldo % 8 == 0 && reinterpret_cast<uintptr_t>(a.l2.out) % 32 == 0 &&
                       (!rrow || (a.l2.ldr % 8 == 0 && reinterpret_cast<uintptr_t>(a.l2.residual) % 32 == 0));
      const int cb = half * (N2 / 2), ce = cb + N2 / 2;
      float rn[32];
      auto load_res = [&](int c0) {
        if (vec) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) tc::ldg256(rrow + c0 + i, rn + i);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) rn[i] = __ldg(rrow + c0 + i);
        }
      };
      if (rrow && !(a.dbg & 2)) load_res(cb);
      tc::mbar_wait(yfull, it & 1);
      tc::tc_fence_after();
      if (!(a.dbg & 2))
#pragma unroll 1
      for (int c0 = cb; c0 < ce; c0 += 32) {
        float v[32], r[32];
        tc::tmem_ld32(tmem + lane_base + kYCol + c0, v);
        if (rrow) {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = rn[i];
          if (c0 + 32 < ce) load_res(c0 + 32);
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 s4 = *reinterpret_cast<const float4*>(s_scale2 + c0 + i), h4 = *reinterpret_cast<const float4*>(s_shift2 + c0 + i);
          const float sc[4] = {s4.x, s4.y, s4.z, s4.w}, sh[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            float y = v[i + t];
            if (rrow && a.l2.res_first) y += r[i + t];
            y = __fadd_rn(__fmul_rn(y, sc[t]), sh[t]);           // separate multiply and add: the rounding of linear_tma's epilogue
            if (cshift) y += __ldg(cshift + c0 + i + t);
            if (a.l2.lrelu) y = y > 0.f ? y : 0.2f * y;
            if (rrow && !a.l2.res_first) y += r[i + t];
            v[i + t] = y;
          }
        }
        if (live) {
          if (vec) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) tc::stg256(orow + c0 + i, v + i);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) orow[c0 + i] = v[i];
          }
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(yempty);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

template <int N2, int LA, bool CONSTS>
static int launch_mlp2(const Mlp2Args& a, const float* X, long long ldx, int K1, const float* W1, const float* W1lo, long long ldw1,
                       const float* W2, const float* W2lo, long long ldw2, cudaStream_t st) {
  alignas(64) CUtensorMap mx, mw1, mw1l, mw2, mw2l;
  const int k4 = (K1 + 3) / 4 * 4;
  if (int e = make_tile_map(&mx, X, k4, ldx, a.l2.M, 1, 128)) return e;
  if (int e = make_tile_map(&mw1, W1, k4, ldw1, a.Hd, 1, 128)) return e;
  if (int e = make_tile_map(&mw1l, W1lo, k4, ldw1, a.Hd, 1, 128)) return e;
  if (int e = make_tile_map(&mw2, W2, a.Hd, ldw2, N2, 1, 128)) return e;
  if (int e = make_tile_map(&mw2l, W2lo, a.Hd, ldw2, N2, 1, 128)) return e;
  auto kern = a.wait_cycles ? mlp2_kernel<N2, LA, CONSTS, true> : mlp2_kernel<N2, LA, CONSTS, false>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m2_smem(N2)) != cudaSuccess)
    return check_launch("mlp2 smem attribute");
  const int mtiles = ceil_div(a.l2.M, 128);
  const int grid = mtiles < 148 ? mtiles : 148;
  SAMBLE_PRE(st);
  kern<<<grid, kM2Threads, m2_smem(N2), st>>>(mx, mw1, mw1l, mw2, mw2l, a);
  SAMBLE_LAUNCHED("mlp2_kernel");
  return SAMBLE_OK;
}

}  // namespace samble

using namespace samble;

extern "C" void samble_set_mlp2_debug(int bits) { samble::g_m2_debug = bits; }
extern "C" void samble_set_mlp2_probe(long long* wait_cycles) { samble::g_m2_wait_cycles = wait_cycles; }

extern "C" int samble_mlp2(const float* X, long long ldx, int M, int K1, const float* W1, const float* W1_lo, long long ldw1, int Hd,
                           const float* scale1, const float* shift1, long long shift1_cloud_stride, int lrelu1, const float* W2,
                           const float* W2_lo, long long ldw2, int N2, const float* scale2, const float* shift2,
                           long long shift2_cloud_stride, int lrelu2, const float* residual, long long ldr, int residual_first,
                           float* out, long long ldo, int points_per_cloud, samble_stream_t stream) {
  SAMBLE_REQUIRE(X && W1 && W1_lo && W2 && W2_lo && out, "samble_mlp2: null pointer");
  SAMBLE_REQUIRE(M > 0 && K1 > 0 && K1 <= 128, "samble_mlp2: M=%d, K1=%d (the input width must be <= 128)", M, K1);
  SAMBLE_REQUIRE(Hd >= 128 && Hd % 128 == 0, "samble_mlp2: hidden width %d must be a multiple of 128", Hd);
  SAMBLE_REQUIRE(N2 == 128 || N2 == 256, "samble_mlp2: output width %d must be 128 or 256", N2);
  SAMBLE_REQUIRE(ldx % 4 == 0 && ldx >= (K1 + 3) / 4 * 4 && (uintptr_t)X % 16 == 0,
                 "samble_mlp2: X needs 16-byte aligned rows, zero-padded to a multiple of 4 columns");
  SAMBLE_REQUIRE(ldw1 % 4 == 0 && ldw1 >= (K1 + 3) / 4 * 4 && ((uintptr_t)W1 | (uintptr_t)W1_lo) % 16 == 0 && ldw2 % 4 == 0 &&
                     ldw2 >= Hd && ((uintptr_t)W2 | (uintptr_t)W2_lo) % 16 == 0,
                 "samble_mlp2: weight rows must be 16-byte aligned (zero-padded to a multiple of 4 columns)");
  SAMBLE_REQUIRE((!scale1 || (uintptr_t)scale1 % 16 == 0) && (!shift1 || ((uintptr_t)shift1 % 16 == 0 && shift1_cloud_stride % 4 == 0)),
                 "samble_mlp2: scale1 / shift1 must be 16-byte aligned");
  const bool need_npc = shift1_cloud_stride != 0 || shift2_cloud_stride != 0;
  SAMBLE_REQUIRE(!need_npc || (points_per_cloud > 0 && points_per_cloud % 128 == 0 && M % points_per_cloud == 0),
                 "samble_mlp2: per-cloud shifts need clouds of a multiple of 128 points (%d) dividing M=%d", points_per_cloud, M);
  SAMBLE_REQUIRE(ldo >= N2 && (!residual || ldr >= N2), "samble_mlp2: output / residual pitch below the output width");
  Mlp2Args a;
  a.l2 = LinArgs{nullptr, 0, nullptr, 0, nullptr, scale2, shift2, residual, ldr, out, ldo, M, Hd, N2, need_npc ? points_per_cloud : 0,
                 lrelu2, 0, 0, residual_first, 0, shift2_cloud_stride, nullptr, nullptr, 0, nullptr, nullptr, 1.f, 0, nullptr, 0};
  a.scale1 = scale1;
  a.shift1 = shift1;
  a.shift1_ldb = shift1_cloud_stride;
  a.lrelu1 = lrelu1;
  a.Hd = Hd;
  a.nkb1 = (K1 + 31) / 32;
  a.dbg = g_m2_debug;
  a.wait_cycles = g_m2_wait_cycles;
  cudaStream_t st = (cudaStream_t)stream;
  const bool consts = scale1 || shift1;
  if (N2 == 128)
    return consts ? launch_mlp2<128, 1, true>(a, X, ldx, K1, W1, W1_lo, ldw1, W2, W2_lo, ldw2, st)
                  : launch_mlp2<128, 1, false>(a, X, ldx, K1, W1, W1_lo, ldw1, W2, W2_lo, ldw2, st);
  return consts ? launch_mlp2<256, 0, true>(a, X, ldx, K1, W1, W1_lo, ldw1, W2, W2_lo, ldw2, st)
                : launch_mlp2<256, 0, false>(a, X, ldx, K1, W1, W1_lo, ldw1, W2, W2_lo, ldw2, st);
}
