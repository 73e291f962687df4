// Point-wise linear layers, row-major activations: 3xTF32 on tcgen05 with TMA operand loads.
//
// Same math and epilogue as linear_tc.cu (see there for the why of 3xTF32 and of the chunked accumulation); what
// differs is how operands reach shared memory.  Measured on B200 the cp.async loader of linear_tc.cu executes ~1200
// instructions per 32-channel stage and was the kernel's limit (profiles/r1_linear_tma.md); here
//   * one thread issues TMA box loads (tensor maps with 128-byte swizzle = the UMMA K-major operand layout); the raw
//     fp32 tile IS the "hi" operand (the tensor core reads only the top 19 bits);
//   * four "splitter" warps derive X_lo = X - trunc_tf32(X) from the landed tile (LDS.128 -> 8 ALU -> STS.128),
//     publish it to the async proxy and arrive on ready[];
//   * when the whole [NT x K] weight slice (hi and pre-split lo) fits next to the X ring (K <= 128), it is loaded
//     ONCE per run of m-tiles and stays resident: a CTA owns a contiguous run of m-tiles of one n-tile, so the
//     L2->SM traffic per output tile drops from X + 2W to X.  Larger K streams W_hi/W_lo with each stage.
//
// Warps 0-3 epilogue (thread = output row = TMEM lane), 4 MMA issuer, 5 TMA producer, 6-9 splitters.
#include "linear_common.cuh"

namespace samble {

constexpr int kLtThreads = 320;

// measurement switches (tools/probe_linear.py): 1 = splitters idle, 2 = epilogue idle, 4 = no MMAs, 8 / 16 = no residual loads /
// no stores in the direct epilogue (results are garbage), 64 = 96-wide tiles + TMA stores for resident weights (K = 128).
int g_lt_debug = 0;

struct LtPlan {
  int stages;       // X (or X+W) ring depth
  int w_res;        // weight slice resident
  int stage_bytes;
  int staged;       // epilogue staging slab present
  size_t smem;
};

constexpr size_t kLtFixed = 1024 /* alignment slack */ + 512 /* barriers */;
constexpr size_t kLtLimit = 227 * 1024;

// Resident weights need a ring of at least 3 X stages next to them to hide the TMA latency.  The epilogue staging
// slab (coalesced row-major stores, measured ~10% on the streaming shapes) is taken only when it costs no stage.
template <int NT>
static bool lt_resident_ok(int K) {
  const int nkb = (K + 31) / 32;
  return (size_t)2 * nkb * NT * 128 + (size_t)3 * 32768 + kLtFixed <= kLtLimit;
}

template <int NT>
static LtPlan lt_plan(int K) {
  const int nkb = (K + 31) / 32;
  LtPlan p;
  const size_t wres = lt_resident_ok<NT>(K) ? (size_t)2 * nkb * NT * 128 : 0;
  p.w_res = wres != 0;
  p.stage_bytes = p.w_res ? 32768 : 32768 + 2 * NT * 128;
  const size_t room = kLtLimit - kLtFixed - wres;
  p.stages = (int)(room / p.stage_bytes);
  if (p.stages > 4) p.stages = 4;
  p.staged = room - (size_t)p.stages * p.stage_bytes >= (size_t)kLinStoreBytes;
  if (!p.staged && p.stages == 4) {          // 4 -> 3 stages still hides the latency
    p.stages = 3;
    p.staged = 1;
  }
  p.smem = wres + (size_t)p.stages * p.stage_bytes + (p.staged ? kLinStoreBytes : 0) + kLtFixed;
  return p;
}

// CTA c owns tiles [c*T/G, (c+1)*T/G) of the n-major tile order (tile = n_tile*mtiles + m_tile): contiguous m-tiles
// of one n-tile, at most a couple of n-tile changes per CTA.
__device__ __forceinline__ void lt_range(int total, int& t0, int& t1) {
  t0 = (int)((long long)blockIdx.x * total / gridDim.x);
  t1 = (int)((long long)(blockIdx.x + 1) * total / gridDim.x);
}

// EPI selects the one epilogue compiled into an instantiation (0 direct, 1 smem-staged row-major, 2 pooling, 3 row
// statistics, 5 pooling in the SWAPPED orientation: weight tile as the A operand, accumulator = transposed tile, see
// linear_common.cuh): with all of them in one function ptxas sized the kernel for their union and spilled.
template <int NT, int EPI>
__global__ void __launch_bounds__(kLtThreads, 1)
    linear_tma_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                      const __grid_constant__ CUtensorMap map_wlo, const __grid_constant__ CUtensorMap map_out, LinArgs a,
                      int stages, int w_res, int stage_bytes, int staged, int dbg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = tc::smem_align1024(smem_raw);
  const int nkb = (a.K + 31) / 32;
  const int wtile = NT * 128;                                   // one K-block of the weight slice
  uint8_t* wres = base;                                         // [hi: nkb tiles][lo: nkb tiles]   (w_res only)
  uint8_t* ring = base + (w_res ? (size_t)2 * nkb * wtile : 0);
  uint8_t* stage_out = ring + (size_t)stages * stage_bytes;                          // TMA-store staging (4 warps x 2 x 4 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_out + (staged ? kLinStoreBytes : 0));
  uint64_t* landed = bars;            // [4] TMA bytes of the stage arrived
  uint64_t* ready = bars + 4;         // [4] ... and X_lo written (4 splitter-warp arrivals)
  uint64_t* empty = bars + 8;         // [4] tcgen05.commit: stage consumed
  uint64_t* tfull = bars + 12;        // [2] accumulator set complete
  uint64_t* tempty = bars + 14;       // [2] accumulator set drained (4 epilogue-warp arrivals)
  uint64_t* wfull = bars + 16;        // resident weight slice landed
  uint64_t* wempty = bars + 17;       // ... no longer read by any MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mtiles = (a.M + 127) / 128, ntiles = (a.Nout + NT - 1) / NT;
  const int total = mtiles * ntiles;
  const int nclouds = a.w_batched ? a.M / a.npc : 1;
  int t0, t1;
  lt_range(total, t0, t1);
  const int chain = a.chain > 0 ? a.chain : kLinChain;
  const int nacc = (nkb + chain - 1) / chain;
  const int nsets = (2 * nacc * NT <= 512) ? 2 : 1;
  uint32_t tcols = 32;
  while (tcols < (uint32_t)(nsets * nacc * NT)) tcols <<= 1;

  if (tid == 0) {
    for (int s = 0; s < 4; ++s) {
      tc::mbar_init(&landed[s], 1);
      tc::mbar_init(&ready[s], 4);
      tc::mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&tfull[i], 1);
      tc::mbar_init(&tempty[i], 4);
    }
    tc::mbar_init(wfull, 1);
    tc::mbar_init(wempty, 1);
    tc::mbar_init_fence();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, tcols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 5) {
    // ================= TMA producer =================
    if (tc::elect_one()) {
      tc::tma_prefetch_desc(&map_x);
      tc::tma_prefetch_desc(&map_w);
      tc::tma_prefetch_desc(&map_wlo);
      int s = 0, ph = 0, cur_n = -1, wruns = 0;
      for (int tile = t0; tile < t1; ++tile) {
        const int nt = tile / mtiles, m0 = (tile % mtiles) * 128, n0 = nt * NT;
        const int cloud = a.w_batched ? m0 / a.npc : 0;       // weight slice of this tile (npc % 128 == 0)
        const int wkey = nt * nclouds + cloud;
        if (w_res && wkey != cur_n) {
          tc::mbar_wait(wempty, (wruns & 1) ^ 1);             // MMAs of the previous run retired
          tc::mbar_arrive_expect_tx(wfull, (uint32_t)(2 * nkb * wtile));
          for (int kb = 0; kb < nkb; ++kb) {
            tc::tma_load_3d(wres + (size_t)kb * wtile, &map_w, wfull, kb * 32, n0, cloud);
            tc::tma_load_3d(wres + (size_t)(nkb + kb) * wtile, &map_wlo, wfull, kb * 32, n0, cloud);
          }
          cur_n = wkey;
          ++wruns;
        }
        for (int kb = 0; kb < nkb; ++kb) {
          tc::mbar_wait(&empty[s], ph ^ 1);
          uint8_t* st = ring + (size_t)s * stage_bytes;
          tc::mbar_arrive_expect_tx(&landed[s], 16384u + (w_res ? 0u : 2u * wtile));
          tc::tma_load_3d(st, &map_x, &landed[s], kb * 32, m0, 0);
          if (!w_res) {
            tc::tma_load_3d(st + 32768, &map_w, &landed[s], kb * 32, n0, cloud);
            tc::tma_load_3d(st + 32768 + wtile, &map_wlo, &landed[s], kb * 32, n0, cloud);
          }
          if (++s == stages) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 6) {
    // ================= splitters: X_lo = X - trunc_tf32(X) =================
    const int lt = tid - 192;                           // 0..127
    int s = 0, ph = 0;
    const int G = (t1 - t0) * nkb;
    for (int g = 0; g < G; ++g) {
      tc::mbar_wait(&landed[s], ph);
      uint8_t* xh = ring + (size_t)s * stage_bytes;
      uint8_t* xl = xh + 16384;
      if (!(dbg & 1))
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t off = (uint32_t)(lt + 128 * i) * 16u;  // element-wise op: any chunk -> same chunk, no swizzle math
        const float4 v = *reinterpret_cast<const float4*>(xh + off);
        float4 lo;
        lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
        lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
        lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
        lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
        *reinterpret_cast<float4*>(xl + off) = lo;
      }
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&ready[s]);
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 4) {
    // ================= MMA issuer =================
    if (tc::elect_one()) {
      const uint32_t idesc = tc::instr_desc(2, 128, NT);
      int s = 0, ph = 0, it = 0, cur_n = -1, wruns = 0;
      const uint32_t wbase = tc::smem_u32(wres);
      for (int tile = t0; tile < t1; ++tile, ++it) {
        const int nt = tile / mtiles;
        const int wkey = nt * nclouds + (a.w_batched ? ((tile % mtiles) * 128) / a.npc : 0);
        const int set = (nsets == 2) ? (it & 1) : 0;
        const int use = (nsets == 2) ? (it >> 1) : it;
        if (w_res && wkey != cur_n) {
          tc::mbar_wait(wfull, wruns & 1);
          cur_n = wkey;
          ++wruns;
        }
        tc::mbar_wait(&tempty[set], (use & 1) ^ 1);
        tc::tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          tc::mbar_wait(&ready[s], ph);
          tc::tc_fence_after();
          const uint32_t st = tc::smem_u32(ring + (size_t)s * stage_bytes);
          const uint64_t xh = tc::smem_desc_sw128(st), xl = tc::smem_desc_sw128(st + 16384);
          const uint64_t wh = tc::smem_desc_sw128(w_res ? wbase + kb * wtile : st + 32768);
          const uint64_t wl = tc::smem_desc_sw128(w_res ? wbase + (nkb + kb) * wtile : st + 32768 + wtile);
          const uint32_t acc = tmem + (set * nacc + kb / chain) * NT;
          if (!(dbg & 4))
#pragma unroll
          for (int k8 = 0; k8 < 4; ++k8) {
            if (EPI == 5) {                                      // swapped: D^T = W X^T (same products, same order)
              tc::mma_tf32(acc, wh + 2 * k8, xh + 2 * k8, idesc, ((kb % chain) | k8) != 0);
              tc::mma_tf32(acc, wh + 2 * k8, xl + 2 * k8, idesc, 1);
              tc::mma_tf32(acc, wl + 2 * k8, xh + 2 * k8, idesc, 1);
            } else {
              tc::mma_tf32(acc, xh + 2 * k8, wh + 2 * k8, idesc, ((kb % chain) | k8) != 0);
              tc::mma_tf32(acc, xl + 2 * k8, wh + 2 * k8, idesc, 1);
              tc::mma_tf32(acc, xh + 2 * k8, wl + 2 * k8, idesc, 1);
            }
          }
          tc::mma_commit(&empty[s]);
          if (++s == stages) { s = 0; ph ^= 1; }
        }
        tc::mma_commit(&tfull[set]);
        // last tile of this weight run: the resident slice may be replaced once these MMAs retire
        if (w_res) {
          const int nxt = tile + 1;
          const int nkey = nxt < t1 ? (nxt / mtiles) * nclouds + (a.w_batched ? ((nxt % mtiles) * 128) / a.npc : 0) : -1;
          if (nkey != wkey) tc::mma_commit(wempty);
        }
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue: thread = output row =================
    int it = 0, parity = 0;
    for (int tile = t0; tile < t1; ++tile, ++it) {
      const int set = (nsets == 2) ? (it & 1) : 0;
      const int use = (nsets == 2) ? (it >> 1) : it;
      const int m0 = (tile % mtiles) * 128, n0 = (tile / mtiles) * NT;
      tc::mbar_wait(&tfull[set], use & 1);
      tc::tc_fence_after();
      if (dbg & 2) {
        // (measurement switch: accumulators are drained but nothing is computed or stored)
      } else if (EPI == 3)
        linear_epilogue_tile_rowstat<NT>(a, tmem, set, nacc, m0, n0, warp, lane, ntiles);
      else if (EPI == 2)
        linear_epilogue_tile_pool<NT>(a, tmem, set, nacc, m0, n0, warp, lane);
      else if (EPI == 0)
        linear_epilogue_tile<NT>(a, tmem, set, nacc, m0, n0, warp, lane);
      else if (EPI == 5)
        linear_epilogue_tile_pool_swapped<NT>(a, tmem, set, nacc, m0, n0, warp, lane);
      else
        linear_epilogue_tile_tma<NT>(a, &map_out, tmem, set, nacc, m0, n0, warp, lane, stage_out, parity);
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty[set]);
    }
    if (EPI == 1 && lane == 0) tc::bulk_wait<0>();              // this warp's stores have left shared memory and landed
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, tcols);
}


template <int NT>
int launch_linear_tma(const LinArgs& a_in, cudaStream_t st) {
  const LtPlan p = lt_plan<NT>(a_in.K);
  LinArgs a = a_in;
  a.dbg = g_lt_debug;
  alignas(64) CUtensorMap mx, mw, mwl;
  // inner extent = K rounded up to 4 (the zero padding the ABI asks for); the rest of a 32-channel box reads as zero
  const int k4 = (a.K + 3) / 4 * 4;
  if (int e = make_tile_map(&mx, a.X, k4, a.ldx, a.M, 1, 128)) return e;
  const int wb = a.w_batched ? a.M / a.npc : 1;
  if (int e = make_tile_map(&mw, a.W, k4, a.ldw, a.Nout, wb, NT)) return e;
  if (int e = make_tile_map(&mwl, a.Wlo, k4, a.ldw, a.Nout, wb, NT)) return e;
  const bool tma_out = p.staged && !a.out_cm && !(a.residual && a.res_cm) && a.out && a.ldo % 4 == 0 &&
                       reinterpret_cast<uintptr_t>(a.out) % 16 == 0 && !a.stat_out && !a.pool_max;
  // pooling in the swapped orientation (linear_common.cuh): 128-channel tiles, whole 128-row tiles inside one cloud.
  // samble_set_linear_debug(128) turns it off (the caller then asks for 32-row groups).
  const bool swap_pool = NT == 128 && !(g_lt_debug & 128) && a.pool_max && a.pool_rows == 128;
  if (a.pool_max && a.pool_rows == 128 && !swap_pool) {
    set_error("linear_tma: 128-row pooling groups need the swapped 128-channel kernel");
    return SAMBLE_E_INVALID;
  }
  const int epi = a.stat_out ? 3 : (a.pool_max ? (swap_pool ? 5 : 2) : (tma_out ? 1 : 0));
  alignas(64) CUtensorMap mo;
  if (tma_out) {
    if (int e = make_tile_map(&mo, a.out, a.Nout, a.ldo, a.M, 1, 32)) return e;
  } else {
    mo = mx;                                                   // (unused by the other epilogues)
  }
  auto kern = epi == 3 ? linear_tma_kernel<NT, 3>
                       : (epi == 2 ? linear_tma_kernel<NT, 2> : (epi == 1 ? linear_tma_kernel<NT, 1> : linear_tma_kernel<NT, 0>));
  if (NT == 128 && epi == 5) kern = linear_tma_kernel<128, 5>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem) != cudaSuccess)
    return check_launch("linear_tma smem attribute");
  const long long total = (long long)ceil_div(a.M, 128) * ceil_div(a.Nout, NT);
  const int grid = (int)(total < 148 ? total : 148);
  SAMBLE_PRE(st);
  kern<<<grid, kLtThreads, p.smem, st>>>(mx, mw, mwl, mo, a, p.stages, p.w_res, p.stage_bytes, p.staged, g_lt_debug);
  SAMBLE_LAUNCHED("linear_tma_kernel");
  return SAMBLE_OK;
}

template int launch_linear_tma<128>(const LinArgs&, cudaStream_t);
template int launch_linear_tma<96>(const LinArgs&, cudaStream_t);
template int launch_linear_tma<64>(const LinArgs&, cudaStream_t);

int launch_linear_tma_auto(const LinArgs& a, int nacc, cudaStream_t st) {
  const bool wide_ok = a.Nout > 64 && nacc * 128 <= 512;
  if (!wide_ok) return launch_linear_tma<64>(a, st);
  // Row-major outputs leave through TMA stores where 32 KB of staging fit next to the operand ring (all streaming shapes,
  // K > 128: 13 % faster than the slab-staged thread stores they replace).  With a resident 128-wide weight slice (K = 128)
  // there is no room; a 96-wide slice would make it, but measured slower there (more n-tiles re-read X, and 4 KB TMA boxes
  // drain no faster than 256-bit thread stores: tools/probe_linear.py, bit 64), so those shapes keep the direct epilogue.
  if ((g_lt_debug & 64) && lt_plan<128>(a.K).w_res && !lt_plan<128>(a.K).staged && !a.out_cm && !a.stat_out && !a.pool_max)
    return launch_linear_tma<96>(a, st);
  return launch_linear_tma<128>(a, st);
}

}  // namespace samble

extern "C" void samble_set_linear_debug(int bits) { samble::g_lt_debug = bits; }
