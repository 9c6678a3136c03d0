// Pieces shared by the CTA-pair backward kernels (tc_kernels.cu: tc_block_grad2_kernel; tc_bwd_persist.cu: the persistent
// variant): launch-shape constants, the G = f(S) epilogue arithmetic, and the host-side preparation of the column
// statistics / f16 copy of Y.  Internal to the library; nothing here is part of the C ABI.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tc_host.cuh"

namespace mclip {

// column statistics (prep_ly2_kernel) and the f16 view / copy of Y that one backward launch reads
struct Bwd2Prep {
  float* ly2; float* bcol; float* stepmm; float* mu0;
  const void* y16; int64_t ld16;
  bool has_col;
};
// Launches the preparation on a.stream: stat_ws takes 2 * n_pad + 2 * (n_pad / 256) + 1 floats, y16_ws N * D halves (bf16
// inputs without a.y16 only).  Defined in tc_kernels.cu.
int tc_bwd2_prepare(const BlockGradArgs& a, int64_t n_pad, float* stat_ws, void* y16_ws, Bwd2Prep* out);
int tc_dbg_flags();     // option "dbg" (MCLIP_DBG at load)
int tc_pair_slots();    // CTA-pair slots of the current device (SM count / 2)

// persistent CTA-pair backward (tc_bwd_persist.cu), selected by the option "bwd_persist"
size_t tc_block_grad2p_ws(int64_t N, int64_t D, int64_t n_pad);
int tc_block_grad2p(const BlockGradArgs& a);

namespace {

using namespace ptx;

// In-kernel cycle accounting (clock64 around the mbarrier waits + printf from a few CTAs) is compiled in only with
// -DMCLIP_PROFILE (python -m mamba_clip_b200.build --profile) and enabled at run time with MCLIP_DBG=16; release
// builds carry no printf and no extra registers.
#ifdef MCLIP_PROFILE
constexpr bool kProfile = true;
#else
constexpr bool kProfile = false;
#endif

constexpr int kThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kEpiThreads = 256;
constexpr uint32_t kChunkBytes = 128 * 64 * 2;  // [128 rows x 64 k] 16-bit
constexpr uint32_t kSmemMax = 232448;           // 227 KB opt-in limit per CTA
constexpr uint32_t kMiscBytes = 2048;           // barriers + small staging
constexpr uint32_t kAlignSlack = 1024;
constexpr uint32_t kMaxStages = 8;
constexpr int kMaxKch = 12;  // D <= 768

__device__ __forceinline__ uint32_t align1024(uint32_t a) { return (a + 1023u) & ~1023u; }

constexpr uint32_t kTile8K = 64 * 64 * 2;     // [64 rows x 64 k]
constexpr uint32_t kStage2 = 2 * kChunkBytes; // ring stage: two [128 x 64] tiles
constexpr int kRing2 = 4;
constexpr float kGScale = 4096.f;             // G is stored as G * 2^12 in f16

template <bool kMasked, bool kCol>
__device__ __forceinline__ void bwd2_chunk(const uint32_t (&v)[32], uint32_t (&g)[16], float k2, float lx2,
                                           const float* __restrict__ ly2, float w_diag_s, int64_t col0, int64_t N,
                                           int64_t jd, float& rd) {
  float ly[32];
  if (kCol) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(ly2 + j));
      ly[j] = t.x; ly[j + 1] = t.y; ly[j + 2] = t.z; ly[j + 3] = t.w;
    }
  }
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    float gv[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float c = __uint_as_float(v[j + e]);
      float p_row = ex2_approx(fmaf(c, k2, -lx2));
      float gg = p_row;
      if (kCol) gg += ex2_approx(fmaf(c, k2, -ly[j + e]));
      if (kMasked) {
        if (col0 + j + e == jd) gg -= w_diag_s;
        if (col0 + j + e >= N) { gg = 0.f; p_row = 0.f; }
      }
      rd = fmaf(p_row, c, rd);
      gv[e] = gg;
    }
    g[j >> 1] = pack_f16x2(gv[0], gv[1]);
  }
}

// One exponential per element: P^col_ij = P^row_ij * 2^(lx2_i - ly2_j) = P^row_ij * a_i * b_j, so
// G = P^row (1 + a_i b_j).  Only used when the caller has checked that a_i, b_j and a_i * b_j stay finite
// (|lx2_i - ly2_j| <= 100 over the tile); otherwise bwd2_chunk evaluates both exponentials.
template <bool kMasked>
__device__ __forceinline__ void bwd2_chunk_fast(const uint32_t (&v)[32], uint32_t (&g)[16], float k2, float lx2, float a_i,
                                                const float* __restrict__ bcol, float w_diag_s, int64_t col0, int64_t N,
                                                int64_t jd, float& rd) {
  float b[32];
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(bcol + j));
    b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
  }
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    float gv[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float c = __uint_as_float(v[j + e]);
      float p_row = ex2_approx(fmaf(c, k2, -lx2));
      float gg = p_row * fmaf(a_i, b[j + e], 1.f);
      if (kMasked) {
        if (col0 + j + e == jd) gg -= w_diag_s;
        if (col0 + j + e >= N) { gg = 0.f; p_row = 0.f; }
      }
      rd = fmaf(p_row, c, rd);
      gv[e] = gg;
    }
    g[j >> 1] = pack_f16x2(gv[0], gv[1]);
  }
}


}  // namespace

}  // namespace mclip
