// Blackwell (sm_100a) tensor-core plumbing used by the tcgen05 kernels: mbarriers, proxy fences, TMEM
// allocation, UMMA shared-memory / instruction descriptors, tcgen05.mma / commit / ld.
// Descriptor bit layouts follow cute/arch/mma_sm100_desc.hpp (vendored CUTLASS headers were read for the
// field positions only; nothing is included from them).
#pragma once
#include <cuda.h>   // CUtensorMap (types only; the encoder is resolved at run time, libcuda is not linked)

#include "common.cuh"

namespace samble {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of the (converged) warp.  Unlike `lane == 0`, elect.sync tells the compiler that exactly one thread
// runs the guarded region, so descriptors and barrier addresses go straight to uniform registers instead of
// through a per-lane broadcast loop (3x fewer instructions in the MMA issue loop).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (launch error reported to the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(32);                        // back off: hundreds of threads poll a handful of barriers
    if (clock64() - t0 > 4000000000LL) {   // ~2 s
      printf("samble: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
      __trap();
    }
  }
}

// ---- TMA (bulk async copies; completion is counted in bytes on an mbarrier) ----
// this thread's arrival + "expect `bytes` more" on the barrier
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// box of a 3-D tensor map -> shared memory (layout/swizzle as encoded in the map); coordinates innermost first
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// ---- thread-block clusters: one candidate tile fetched from L2 once and written into every CTA of the cluster ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the box lands at the same shared-memory offset in every CTA of `mask`, and each of their mbarriers (same offset) gets
// the complete_tx for its copy
__device__ __forceinline__ void tma_load_3d_multicast(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                                      uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5, %6}], [%2], %3;" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// arrives on the mbarrier at this offset in every CTA of `mask` once all MMAs issued so far by this thread completed
__device__ __forceinline__ void mma_commit_multicast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}

// shared -> global tile store (bulk async group of the issuing thread); the tensor map clips rows / columns past the end
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// at most N of this thread's bulk groups still READING their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// contiguous global -> shared copy, 16-byte aligned, size a multiple of 16
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// contiguous shared -> global copy (bulk async group of the issuing thread), 16-byte aligned, size a multiple of 16
__device__ __forceinline__ void bulk_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM ----
// whole warp; writes the base address (lane 0, column c) to *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}

// ---- descriptors ----
// K-major operand tile, rows of exactly 128 bytes, 128-byte swizzle (Swizzle<3,4,3>): 16-byte chunk c of
// row r lives at  base + r*128 + ((c ^ (r & 7)) << 4);  base is 1024-byte aligned; 8-row groups are
// 1024 bytes apart (SBO).  One descriptor addresses a [rows x 32 B] K-slice; advancing K by 32 bytes
// inside the 128-byte row adds 2 to the (addr >> 4) start field.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);        // start address        [0,14)
  d |= (uint64_t)1 << 16;                            // leading byte offset  [16,30)  (unused for SW128 K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset   [32,46)
  d |= (uint64_t)1 << 46;                            // descriptor version 1 [46,48)
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B         [61,64)
  return d;
}
// K-major operand slice WITHOUT swizzle, compact: [rows x 32 B] (one UMMA K step) stored as 8x16-byte core
// matrices: byte offset(row, chunk) = chunk*lbo + (row/8)*128 + (row%8)*16, lbo = rows*16.  16-byte aligned.
__device__ __forceinline__ uint64_t smem_desc_nosw(uint32_t smem_addr, uint32_t rows) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((rows * 16) >> 4) << 16;           // leading byte offset: next 16-byte chunk along K
  d |= (uint64_t)(128 >> 4) << 32;                   // stride byte offset: next 8-row group
  d |= (uint64_t)1 << 46;
  return d;                                          // layout type 0 = SWIZZLE_NONE
}
__device__ __forceinline__ uint32_t nosw_offset(int row, int chunk, int rows) { return chunk * rows * 16 + (row >> 3) * 128 + (row & 7) * 16; }

// byte offset of element (row, 16-byte chunk) inside a SW128 K-block tile
__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

// instruction descriptor, kind::tf32 / kind::f16 with fp32 accumulate, both operands K-major
// a/b format: 0 = f16, 1 = bf16, 2 = tf32.
__host__ __device__ constexpr uint32_t instr_desc(int fmt, int M, int N) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; one elected thread issues.
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with bf16 operands (kind::f16, 128 x N x 16 per instruction: twice the tf32 rate)
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Lean issue path.  Measured (tools/probe_mma_rate.py, tools/probe_xgemm.py): a 128 x 64 x 16 MMA takes 48 cycles
// (shared-memory operand reads bound it), but an issuing warp that rebuilds 64-bit descriptors with ~15 uniform-ALU
// instructions per MMA, or moves them through R2UR, issues one MMA per ~75 cycles and paces the kernel.  The high
// word of a SW128 K-major descriptor is a constant and the low word is (address >> 4) | LBO, so an operand tile at a
// compile-time offset from a base is ONE 32-bit add away: lo = lo_base + (byte offset >> 4).
constexpr uint32_t kDescSw128Hi = 0x40004040u;     // SBO = 1024 B, descriptor version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t smem_desc_sw128_lo(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ void mma_bf16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n .reg .b64 da, db;\n setp.ne.b32 p, %4, 0;\n mov.b64 da, {%1, %5};\n mov.b64 db, {%2, %5};\n"
      " tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescSw128Hi)
      : "memory");
}
__device__ __forceinline__ void mma_tf32_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n .reg .b64 da, db;\n setp.ne.b32 p, %4, 0;\n mov.b64 da, {%1, %5};\n mov.b64 db, {%2, %5};\n"
      " tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescSw128Hi)
      : "memory");
}
// 256-bit global accesses (sm_100: LDG.E.ENL2.256 / STG.E.ENL2.256): a thread that owns a whole row moves one full
// 32-byte sector per instruction instead of two half sectors (thread = row epilogues).  32-byte aligned addresses only.
__device__ __forceinline__ void ldg256(const float* p, float* v) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
               "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// 1024-byte aligned start of the dynamic shared-memory window, derived by POINTER arithmetic: the compiler keeps the
// shared address space and emits LDS/STS.  (Rounding through uintptr_t, as every kernel here did until round 2, turns each
// access behind the pointer into a generic LD.E/ST.E: the operand splitters of linear_tma_kernel spent ~4000 cycles per
// 16 KB stage in eight serialised generic round trips and paced the whole kernel, profiles/r2_linear_tma.md.)
__device__ __forceinline__ uint8_t* smem_align1024(uint8_t* p) { return p + ((1024u - (smem_u32(p) & 1023u)) & 1023u); }
// mbarrier arrives when every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// TMEM -> registers: this thread's lane (= accumulator row), consecutive fp32 columns starting at taddr's column.
// Load and wait::ld live in ONE asm statement, so no consumer of the outputs can be scheduled before the wait.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float (&v)[64]) {
  uint32_t r[64];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];\n"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(r[i]);
}


// registers -> TMEM: this thread's lane, 32 consecutive 32-bit columns (the inverse of tmem_ld32); the wait makes the
// data visible to later tcgen05 operations once followed by tc_fence_before + a barrier.
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(v[i]);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr), "r"(r[0]),"r"(r[1]),"r"(r[2]),"r"(r[3]),"r"(r[4]),"r"(r[5]),"r"(r[6]),"r"(r[7]),"r"(r[8]),"r"(r[9]),"r"(r[10]),"r"(r[11]),"r"(r[12]),"r"(r[13]),"r"(r[14]),"r"(r[15]),"r"(r[16]),"r"(r[17]),"r"(r[18]),"r"(r[19]),"r"(r[20]),"r"(r[21]),"r"(r[22]),"r"(r[23]),"r"(r[24]),"r"(r[25]),"r"(r[26]),"r"(r[27]),"r"(r[28]),"r"(r[29]),"r"(r[30]),"r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]^T : the A operand is read from tensor memory (lane = row, one 32-bit column per tf32
// K element, 8 columns per instruction), so an activation produced by a previous MMA can feed the next one without a
// round trip through shared memory (csrc/mlp2.cu).
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n .reg .b64 db;\n setp.ne.b32 p, %4, 0;\n mov.b64 db, {%2, %5};\n"
      " tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %3, p;\n}" ::"r"(tmem_d),
      "r"(tmem_a), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescSw128Hi)
      : "memory");
}

// the same with bf16 operands (kind::f16): A in tensor memory holds TWO K elements per 32-bit column (low half = the even one),
// 8 columns per 128 x N x 16 instruction.  Without the A tile the MMA reads half the shared-memory bytes: a 128 x 128 x 16
// product runs at the tensor pipe's ~66 cycles instead of the ~100 of the smem-smem form (tools/probe_tmem_a.py).
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n .reg .b64 db;\n setp.ne.b32 p, %4, 0;\n mov.b64 db, {%2, %5};\n"
      " tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n}" ::"r"(tmem_d),
      "r"(tmem_a), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescSw128Hi)
      : "memory");
}
__device__ __forceinline__ void tmem_st32u(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr), "r"(r[0]),"r"(r[1]),"r"(r[2]),"r"(r[3]),"r"(r[4]),"r"(r[5]),"r"(r[6]),"r"(r[7]),"r"(r[8]),"r"(r[9]),"r"(r[10]),"r"(r[11]),"r"(r[12]),"r"(r[13]),"r"(r[14]),"r"(r[15]),"r"(r[16]),"r"(r[17]),"r"(r[18]),"r"(r[19]),"r"(r[20]),"r"(r[21]),"r"(r[22]),"r"(r[23]),"r"(r[24]),"r"(r[25]),"r"(r[26]),"r"(r[27]),"r"(r[28]),"r"(r[29]),"r"(r[30]),"r"(r[31])
      : "memory");
}

}  // namespace tc

// host: rank-3 fp32 tensor map over a (batch, rows, inner) array with row pitch `ld` floats (16-byte multiple) and
// batch pitch rows*ld, box = (1, box_rows, 32 floats = 128 B), 128-byte swizzle -- i.e. exactly the K-major SW128
// operand tile smem_desc_sw128 describes; out-of-range rows and channels read as zero.
// Returns SAMBLE_OK or sets the error text.
// elem_bytes 4 = fp32 (box 32 wide), 2 = bf16 (box 64 wide): the box is always one 128-byte swizzled row segment.
int make_tile_map(CUtensorMap* map, const void* base, int inner, long long ld, int rows, int batch, int box_rows,
                  int elem_bytes = 4);

}  // namespace samble
