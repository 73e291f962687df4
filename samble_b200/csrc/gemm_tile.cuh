// FP32 (FFMA) tile engine shared by the feature-space kNN and the DownSample row-statistics pass.
//
// A CTA owns TQ query rows (point-major, C channels) that stay resident in shared memory and
// streams candidate rows in tiles of TN through a 2-stage cp.async ring, KC channels at a time.
// Every (query, candidate) dot product is accumulated by ONE thread in ONE register with FFMAs
// in ascending channel order, so a dot product recomputed elsewhere with the same order
// (ds_edge_score) is bit-identical.  After each candidate tile the TQ x TN dots are parked in a
// shared-memory tile and handed to an epilogue functor (k-selection / online softmax statistics)
// with rows spread over warps and candidates over lanes.  The full Nq x Nr matrix never exists.
#pragma once
#include "common.cuh"

namespace samble {

template <int MQ_>
struct DotTileCfg {
  static constexpr int MQ = MQ_;            // rows per thread (4 or 8)
  static constexpr int MN = 8;              // candidates per thread
  static constexpr int NT = 256;            // threads
  static constexpr int TQ = 16 * MQ;        // query rows per CTA (64 or 128)
  static constexpr int TN = 16 * MN;        // candidates per tile (128)
  static constexpr int KC = 16;             // channels per pipeline stage
  static constexpr int LDB = KC + 4;        // padded row strides (floats): conflict-free LDS.128
  static constexpr int LDS_ = TN + 4;
  static constexpr int RPW = TQ / 8;        // rows per warp in the epilogue

  static __host__ __device__ int lda(int Cp) { return Cp + 4; }
  static __host__ __device__ size_t smem_floats(int Cp) {
    return (size_t)TQ * lda(Cp) + 2 * TN * LDB + (size_t)TQ * LDS_;
  }
};

// A: (Nq rows) x C, row stride lda_g; B: (Nr rows) x C, row stride ldb_g.  C % 4 == 0, Cp = roundup(C, KC).
// epi.tile(S, ldS, n0) is called by all threads once per candidate tile; S[row][col] holds the
// raw dot of query q0+row with candidate n0+col (garbage-free zeros where out of range).
template <class Cfg, class Epi>
__device__ __forceinline__ void dot_tiles(const float* __restrict__ A_g, long long lda_g, int q0, int Nq,
                                          const float* __restrict__ B_g, long long ldb_g, int Nr,
                                          int C, int Cp, float* smem, Epi& epi) {
  constexpr int MQ = Cfg::MQ, MN = Cfg::MN, NT = Cfg::NT, TQ = Cfg::TQ, TN = Cfg::TN, KC = Cfg::KC;
  constexpr int LDB = Cfg::LDB, LDSS = Cfg::LDS_;
  const int lda = Cfg::lda(Cp);
  float* As = smem;
  float* Bs = As + (size_t)TQ * lda;
  float* S = Bs + 2 * TN * LDB;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int nchunks = Cp / KC;
  const int ntiles = (Nr + TN - 1) / TN;
  const int G = ntiles * nchunks;

  // resident query tile
  for (int p = tid; p < TQ * (Cp / 4); p += NT) {
    int row = p / (Cp / 4), c = (p % (Cp / 4)) * 4;
    bool ok = (q0 + row) < Nq && c < C;
    const float* src = A_g + (long long)(ok ? q0 + row : 0) * lda_g + (ok ? c : 0);
    cp_async16(As + (size_t)row * lda + c, src, ok);
  }
  auto load_B = [&](int g) {
    int nt = g / nchunks, kc = g % nchunks;
    float* dst = Bs + (g & 1) * TN * LDB;
#pragma unroll
    for (int p = tid; p < TN * (KC / 4); p += NT) {
      int row = p / (KC / 4), c = (p % (KC / 4)) * 4;
      int n = nt * TN + row, cg = kc * KC + c;
      bool ok = n < Nr && cg < C;
      const float* src = B_g + (long long)(ok ? n : 0) * ldb_g + (ok ? cg : 0);
      cp_async16(dst + row * LDB + c, src, ok);
    }
  };
  load_B(0);
  cp_async_commit();

  float acc[MQ][MN];
#pragma unroll
  for (int r = 0; r < MQ; ++r)
#pragma unroll
    for (int m = 0; m < MN; ++m) acc[r][m] = 0.f;

  for (int g = 0; g < G; ++g) {
    if (g + 1 < G) load_B(g + 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const int kc = g % nchunks;
    const float* Bst = Bs + (g & 1) * TN * LDB;
    const float* Ak = As + kc * KC;
#pragma unroll
    for (int c = 0; c < KC; c += 4) {
      float4 a4[MQ], b4[MN];
#pragma unroll
      for (int r = 0; r < MQ; ++r) a4[r] = *reinterpret_cast<const float4*>(Ak + (size_t)(ty + 16 * r) * lda + c);
#pragma unroll
      for (int m = 0; m < MN; ++m) b4[m] = *reinterpret_cast<const float4*>(Bst + (tx + 16 * m) * LDB + c);
#pragma unroll
      for (int r = 0; r < MQ; ++r)
#pragma unroll
        for (int m = 0; m < MN; ++m) {
          float v = acc[r][m];
          v = fmaf(a4[r].x, b4[m].x, v);
          v = fmaf(a4[r].y, b4[m].y, v);
          v = fmaf(a4[r].z, b4[m].z, v);
          v = fmaf(a4[r].w, b4[m].w, v);
          acc[r][m] = v;
        }
    }
    if (kc == nchunks - 1) {
#pragma unroll
      for (int r = 0; r < MQ; ++r)
#pragma unroll
        for (int m = 0; m < MN; ++m) {
          S[(ty + 16 * r) * LDSS + tx + 16 * m] = acc[r][m];
          acc[r][m] = 0.f;
        }
      __syncthreads();
      epi.tile(S, LDSS, (g / nchunks) * TN);
    }
    __syncthreads();
  }
  cp_async_wait<0>();
}

}  // namespace samble
