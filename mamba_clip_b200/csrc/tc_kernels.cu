// Tensor-core path of the contrastive-loss kernels for B200 (sm_100a):
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> shared memory -> tcgen05.mma (bf16/f16, f32 accumulate
//   in TMEM) -> tcgen05.ld epilogue warps.  The logits block only ever exists as 128 x BN f32 tiles in TMEM.
//
// Two kernels, both "one-sided" (rows of X against all rows of Y), see DESIGN.md:
//   tc_row_lse_kernel     partial row (max, sum-exp) of ls * X Y^T           (forward)
//   tc_block_grad_kernel  dX = alpha * G Y with G recomputed tile by tile     (backward)
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4..11 = epilogue (two warpgroups; warp w reads TMEM lanes 32*(w%4).. and one column half).
#include <cuda.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace mclip {

namespace {

using namespace ptx;

constexpr int kThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kEpiThreads = 256;
constexpr uint32_t kChunkBytes = 128 * 64 * 2;  // [128 rows x 64 k] 16-bit
constexpr uint32_t kSmemMax = 232448;           // 227 KB opt-in limit per CTA
constexpr uint32_t kMiscBytes = 2048;           // barriers + small staging
constexpr uint32_t kAlignSlack = 1024;
constexpr uint32_t kMaxStages = 8;
constexpr int kMaxKch = 12;  // D <= 768

struct FwdParams {
  int64_t M, N;
  int kch;            // ceil(D / 64)
  int stages;
  int tiles_total;    // ceil(N / BN)
  int tiles_per_split;
  int64_t diag_off;
  const float* ls;
  float* part_m2;
  float* part_s;
  float* diag;
  int bf16;
};

struct BwdParams {
  int64_t M, N, D;
  int kch;
  int stages;
  int steps_total;      // ceil(N / 128)
  int steps_per_split;
  int nsplit;
  int64_t diag_off;
  const float* ls;
  const float* go;
  const float* lse_x;
  const float* lse_y;
  float w_row, w_col, w_diag, inv_2n;
  void* dX;
  int64_t lddx;
  float* acc_ws;        // f32 [M, D] when nsplit > 1
  float* rowdot;
};

__device__ __forceinline__ uint32_t align1024(uint32_t a) { return (a + 1023u) & ~1023u; }

// =================================================================================================
// forward
// =================================================================================================
template <bool kMasked>
__device__ __forceinline__ void fwd_chunk(const uint32_t (&v)[32], float k2, int64_t col0, int64_t N, int64_t jd,
                                          float& m2, float& sum, float& diag_val) {
  float x[32];
  float cmax = -INFINITY;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float c = __uint_as_float(v[j]);
    float xv = c * k2;
    if (kMasked) {
      if (col0 + j >= N) xv = -INFINITY;
      if (col0 + j == jd) diag_val = c;
    }
    x[j] = xv;
    cmax = fmaxf(cmax, xv);
  }
  const float m_new = fmaxf(m2, cmax);
  if (m_new == -INFINITY) return;  // nothing valid yet (only possible in masked tail chunks)
  float s = sum * ex2_approx(m2 - m_new);
#pragma unroll
  for (int j = 0; j < 32; ++j) s += ex2_approx(x[j] - m_new);
  sum = s;
  m2 = m_new;
}

template <int BN, bool XRES>
__global__ void __launch_bounds__(kThreads, 1)
tc_row_lse_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = align1024(smem_u32(smem_raw));
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  constexpr uint32_t kYStage = BN * 128;
  constexpr uint32_t kStageBytes = kYStage + (XRES ? 0u : kChunkBytes);
  const uint32_t x_bytes = XRES ? (uint32_t)p.kch * kChunkBytes : 0u;
  const uint32_t ring_base = smem_base + x_bytes;
  const uint32_t misc_base = ring_base + (uint32_t)p.stages * kStageBytes;
  // misc layout: [0,1024) merge scratch (float2 x 128), then barriers
  uint8_t* misc_gen = smem_gen + (misc_base - smem_base);
  float2* merge = reinterpret_cast<float2*>(misc_gen);
  const uint32_t bar_base = misc_base + 1024;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  const uint32_t xfull_bar = bar_base + 8u * (2 * kMaxStages);
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * kMaxStages + 1 + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * kMaxStages + 3 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 5);
  uint32_t* tmem_slot_gen = reinterpret_cast<uint32_t*>(misc_gen + 1024 + 8u * (2 * kMaxStages + 5));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m0 = (int64_t)blockIdx.x * 128;
  const int t0 = blockIdx.y * p.tiles_per_split;
  const int t1 = min(p.tiles_total, t0 + p.tiles_per_split);
  const int ntiles = t1 - t0;
  constexpr uint32_t kTmemCols = 2 * BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(xfull_bar, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), kEpiThreads / 32); }
    fence_barrier_init();
  } else if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      if (XRES) {
        mbar_expect_tx(xfull_bar, x_bytes);
        for (int c = 0; c < p.kch; ++c) tma_load_2d(smem_base + c * kChunkBytes, &tmX, c * 64, (int32_t)m0, xfull_bar);
      }
      uint32_t it = 0;
      for (int t = t0; t < t1; ++t) {
        const int32_t n0 = t * BN;
        for (int c = 0; c < p.kch; ++c, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          mbar_expect_tx(full_bar(s), kStageBytes);
          const uint32_t dst = ring_base + s * kStageBytes;
          tma_load_2d(dst, &tmY, c * 64, n0, full_bar(s));
          if (!XRES) tma_load_2d(dst + kYStage, &tmX, c * 64, (int32_t)m0, full_bar(s));
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      const uint32_t idesc = make_idesc_f16(p.bf16 != 0, 128, BN, false, false);
      if (XRES) mbar_wait(xfull_bar, 0);
      uint32_t it = 0;
      for (int lt = 0; lt < ntiles; ++lt) {
        const int buf = lt & 1;
        const uint32_t bph = (lt >> 1) & 1;
        mbar_wait(tempty_bar(buf), bph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int c = 0; c < p.kch; ++c, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t b_addr = ring_base + s * kStageBytes;
          const uint32_t a_addr = XRES ? (smem_base + c * kChunkBytes) : (b_addr + kYStage);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = make_smem_desc_sw128(a_addr + k * 32, 0, 1024);
            const uint64_t bd = make_smem_desc_sw128(b_addr + k * 32, 0, 1024);
            mma_ss(d_tmem, ad, bd, idesc, (c | k) != 0);
          }
          mma_commit(empty_bar(s));
        }
        mma_commit(tfull_bar(buf));
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ---------------- epilogue: online row log-sum-exp ----------------
    const int ew = warp - kEpiWarp0;
    const int q = warp & 3;           // TMEM sub-partition this warp may access
    const int half = ew >> 2;         // which half of the tile's columns
    const int row_in_tile = q * 32 + lane;
    const int64_t row = m0 + row_in_tile;
    const float k2 = p.ls[0] * kLog2e;
    const int64_t jd = row + p.diag_off;  // this row's positive column
    float m2 = -INFINITY, sum = 0.f, diag_val = 0.f;
    constexpr int kHalfCols = BN / 2;
    for (int lt = 0; lt < ntiles; ++lt) {
      const int buf = lt & 1;
      const uint32_t bph = (lt >> 1) & 1;
      const int64_t n0 = (int64_t)(t0 + lt) * BN;
      mbar_wait(tfull_bar(buf), bph);
      tc_fence_after();
      // tile needs the masked variant if it has columns past N or may hold a diagonal element of this block
      const bool special = (n0 + BN > p.N) || (p.diag != nullptr && n0 < m0 + p.diag_off + 128 && n0 + BN > m0 + p.diag_off);
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + half * kHalfCols;
#pragma unroll 1
      for (int cc = 0; cc < kHalfCols / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32(t_addr + cc * 32, v);
        tmem_ld_wait();
        const int64_t col0 = n0 + half * kHalfCols + cc * 32;
        if (special) {
          if (col0 < p.N) fwd_chunk<true>(v, k2, col0, p.N, jd, m2, sum, diag_val);
        } else {
          fwd_chunk<false>(v, k2, col0, p.N, jd, m2, sum, diag_val);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(buf));
    }
    // merge the two column halves, then write the split's partial
    if (half == 1) merge[row_in_tile] = make_float2(m2, sum);
    named_bar_sync(1, kEpiThreads);
    if (half == 0) {
      const float2 o = merge[row_in_tile];
      const float mm = fmaxf(m2, o.x);
      float s = 0.f;
      if (m2 > -INFINITY) s += sum * exp2f(m2 - mm);
      if (o.x > -INFINITY) s += o.y * exp2f(o.x - mm);
      if (row < p.M) {
        p.part_m2[(int64_t)blockIdx.y * p.M + row] = mm;
        p.part_s[(int64_t)blockIdx.y * p.M + row] = s;
      }
    }
    if (p.diag != nullptr && row < p.M && jd >= 0 && jd < p.N) {
      // exactly one (split, half) owns column jd
      const int64_t c_lo = (int64_t)t0 * BN, c_hi = (int64_t)t1 * BN;
      if (jd >= c_lo && jd < c_hi && (int)((jd % BN) / kHalfCols) == half) p.diag[row] = diag_val;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// =================================================================================================
// backward
// =================================================================================================
// TMEM columns: S/G buffers [0,128) and [128,256); dX accumulator [256, 512).
// G (16-bit, two per column) overwrites its own S buffer: warpgroup 0 owns S columns [0,64) -> G columns
// [0,32); warpgroup 1 owns S columns [64,128) -> G columns [64,96).
template <bool kBF16, bool kMasked, bool kCol>
__device__ __forceinline__ void bwd_chunk(const uint32_t (&v)[32], uint32_t (&g)[16], float k2, float lx2,
                                          const float* __restrict__ ly2, float lw_diag, int64_t col0, int64_t N,
                                          int64_t jd, float& rd) {
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    float gv[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float c = __uint_as_float(v[j + e]);
      float p_row = ex2_approx(fmaf(c, k2, -lx2));
      float gg = p_row;
      if (kCol) gg += ex2_approx(fmaf(c, k2, -ly2[j + e]));
      if (kMasked) {
        if (col0 + j + e == jd) gg -= lw_diag;
        if (col0 + j + e >= N) { gg = 0.f; p_row = 0.f; }
      }
      rd = fmaf(p_row, c, rd);
      gv[e] = gg;
    }
    g[j >> 1] = kBF16 ? pack_bf16x2(gv[0], gv[1]) : pack_f16x2(gv[0], gv[1]);
  }
}

template <bool kBF16, bool XRES>
__global__ void __launch_bounds__(kThreads, 1)
tc_block_grad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = align1024(smem_u32(smem_raw));
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  constexpr uint32_t kStageBytes = kChunkBytes + (XRES ? 0u : kChunkBytes);
  const uint32_t x_bytes = XRES ? (uint32_t)p.kch * kChunkBytes : 0u;
  const int dc0 = blockIdx.y * 4;                       // first 64-wide d chunk of this CTA
  const int ndc = min(4, p.kch - dc0);                  // d chunks handled here (N of the dX MMA = 64 * ndc)
  const uint32_t yd_base = smem_base + x_bytes;         // [ndc][128 y][64 d], one buffer
  const uint32_t ring_base = yd_base + 4 * kChunkBytes;
  const uint32_t misc_base = ring_base + (uint32_t)p.stages * kStageBytes;
  uint8_t* misc_gen = smem_gen + (misc_base - smem_base);
  float* ly2_s = reinterpret_cast<float*>(misc_gen);  // [2][128]
  const uint32_t bar_base = misc_base + 1024;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  const uint32_t xfull_bar = bar_base + 8u * (2 * kMaxStages);
  auto sfull_bar = [&](int b) { return bar_base + 8u * (2 * kMaxStages + 1 + b); };
  auto sempty_bar = [&](int b) { return bar_base + 8u * (2 * kMaxStages + 3 + b); };
  auto gfull_bar = [&](int b) { return bar_base + 8u * (2 * kMaxStages + 5 + b); };
  const uint32_t ydfull_bar = bar_base + 8u * (2 * kMaxStages + 7);
  const uint32_t ydempty_bar = bar_base + 8u * (2 * kMaxStages + 8);
  const uint32_t dxfull_bar = bar_base + 8u * (2 * kMaxStages + 9);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 10);
  uint32_t* tmem_slot_gen = reinterpret_cast<uint32_t*>(misc_gen + 1024 + 8u * (2 * kMaxStages + 10));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m0 = (int64_t)blockIdx.x * 128;
  const int s0 = blockIdx.z * p.steps_per_split;
  const int s1 = min(p.steps_total, s0 + p.steps_per_split);
  const int nsteps = s1 - s0;
  constexpr uint32_t kTmemCols = 512;
  constexpr uint32_t kDxCol = 256;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(xfull_bar, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(sfull_bar(b), 1);
      mbar_init(sempty_bar(b), 1);
      mbar_init(gfull_bar(b), kEpiThreads / 32);
    }
    mbar_init(ydfull_bar, 1);
    mbar_init(ydempty_bar, 1);
    mbar_init(dxfull_bar, 1);
    fence_barrier_init();
  } else if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      if (XRES) {
        mbar_expect_tx(xfull_bar, x_bytes);
        for (int c = 0; c < p.kch; ++c) tma_load_2d(smem_base + c * kChunkBytes, &tmX, c * 64, (int32_t)m0, xfull_bar);
      }
      uint32_t it = 0;
      for (int ls_ = 0; ls_ < nsteps; ++ls_) {
        const int32_t n0 = (s0 + ls_) * 128;
        for (int c = 0; c < p.kch; ++c, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          mbar_expect_tx(full_bar(s), kStageBytes);
          const uint32_t dst = ring_base + s * kStageBytes;
          tma_load_2d(dst, &tmY, c * 64, n0, full_bar(s));
          if (!XRES) tma_load_2d(dst + kChunkBytes, &tmX, c * 64, (int32_t)m0, full_bar(s));
        }
        // Y rows of this step again, as the [K = y][N = d] operand of the dX MMA (single buffer)
        mbar_wait(ydempty_bar, (ls_ & 1) ^ 1);
        mbar_expect_tx(ydfull_bar, (uint32_t)ndc * kChunkBytes);
        for (int qd = 0; qd < ndc; ++qd) tma_load_2d(yd_base + qd * kChunkBytes, &tmY, (dc0 + qd) * 64, n0, ydfull_bar);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      const uint32_t idesc_s = make_idesc_f16(kBF16, 128, 128, false, false);
      const uint32_t idesc_dx = make_idesc_f16(kBF16, 128, (uint32_t)ndc * 64, false, true);
      if (XRES) mbar_wait(xfull_bar, 0);
      uint32_t it = 0;
      auto issue_s = [&](int ls_) {
        const int buf = ls_ & 1;
        const uint32_t bph = (ls_ >> 1) & 1;
        mbar_wait(sempty_bar(buf), bph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * 128;
        for (int c = 0; c < p.kch; ++c, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t b_addr = ring_base + s * kStageBytes;
          const uint32_t a_addr = XRES ? (smem_base + c * kChunkBytes) : (b_addr + kChunkBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = make_smem_desc_sw128(a_addr + k * 32, 0, 1024);
            const uint64_t bd = make_smem_desc_sw128(b_addr + k * 32, 0, 1024);
            mma_ss(d_tmem, ad, bd, idesc_s, (c | k) != 0);
          }
          mma_commit(empty_bar(s));
        }
        mma_commit(sfull_bar(buf));
      };
      auto issue_dx = [&](int ls_) {
        const int buf = ls_ & 1;
        const uint32_t bph = (ls_ >> 1) & 1;
        mbar_wait(gfull_bar(buf), bph);
        mbar_wait(ydfull_bar, ls_ & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // A = G[128 x 16] from TMEM (8 packed columns per k-step); B = Y[16 y][64*ndc d], MN-major:
          // 64-wide d atoms are kChunkBytes apart (LBO), 8-row y groups 1024 B apart (SBO).
          const uint32_t a_tmem = tmem_base + buf * 128 + (k < 4 ? k * 8 : 64 + (k - 4) * 8);
          const uint64_t bd = make_smem_desc_sw128(yd_base + k * 2048, kChunkBytes, 1024);
          mma_ts(tmem_base + kDxCol, a_tmem, bd, idesc_dx, (ls_ | k) != 0);
        }
        mma_commit(ydempty_bar);
        mma_commit(sempty_bar(buf));
      };
      if (nsteps > 0) issue_s(0);
      for (int ls_ = 0; ls_ < nsteps; ++ls_) {
        if (ls_ + 1 < nsteps) issue_s(ls_ + 1);
        issue_dx(ls_);
      }
      mma_commit(dxfull_bar);
    }
  } else if (warp >= kEpiWarp0) {
    // ---------------- epilogue: S -> G (16-bit, back into TMEM), then dX out ----------------
    const int ew = warp - kEpiWarp0;
    const int q = warp & 3;
    const int half = ew >> 2;
    const int et = threadIdx.x - kEpiWarp0 * 32;  // 0..255
    const int row_in_tile = q * 32 + lane;
    const int64_t row = m0 + row_in_tile;
    const float ls = p.ls[0];
    const float k2 = ls * kLog2e;
    const bool has_col = p.w_col != 0.f;
    // weights folded into the exponents: w * 2^a = 2^(a + log2 w)
    const float lw_row = log2f(p.w_row);
    const float lw_col = has_col ? log2f(p.w_col) : 0.f;
    const float lx2 = (row < p.M ? p.lse_x[row] * kLog2e : 0.f) - lw_row;
    const int64_t jd = row + p.diag_off;
    float rd = 0.f;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int ls_ = 0; ls_ < nsteps; ++ls_) {
      const int buf = ls_ & 1;
      const uint32_t bph = (ls_ >> 1) & 1;
      const int64_t n0 = (int64_t)(s0 + ls_) * 128;
      // stage lse_y (log2 units, weight folded) for this step; +inf for columns past N
      if (has_col && et < 128) {
        const int64_t col = n0 + et;
        ly2_s[buf * 128 + et] = col < p.N ? p.lse_y[col] * kLog2e - lw_col : INFINITY;
      }
      named_bar_sync(1, kEpiThreads);
      mbar_wait(sfull_bar(buf), bph);
      tc_fence_after();
      const bool special = (n0 + 128 > p.N) || (n0 < m0 + p.diag_off + 128 && n0 + 128 > m0 + p.diag_off);
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t v[32];
        uint32_t g[16];
        const int cbase = half * 64 + cc * 32;  // S column of this chunk
        tmem_ld32(lane_addr + buf * 128 + cbase, v);
        tmem_ld_wait();
        const float* ly2 = ly2_s + buf * 128 + cbase;
        const int64_t col0 = n0 + cbase;
        if (special) {
          if (has_col) bwd_chunk<kBF16, true, true>(v, g, k2, lx2, ly2, p.w_diag, col0, p.N, jd, rd);
          else bwd_chunk<kBF16, true, false>(v, g, k2, lx2, ly2, p.w_diag, col0, p.N, jd, rd);
        } else {
          if (has_col) bwd_chunk<kBF16, false, true>(v, g, k2, lx2, ly2, p.w_diag, col0, p.N, jd, rd);
          else bwd_chunk<kBF16, false, false>(v, g, k2, lx2, ly2, p.w_diag, col0, p.N, jd, rd);
        }
        tmem_st16(lane_addr + buf * 128 + half * 64 + cc * 16, g);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(gfull_bar(buf));
    }
    // ---- dX accumulator -> global ----
    mbar_wait(dxfull_bar, 0);
    tc_fence_after();
    const float alpha = (p.go ? p.go[0] : 1.f) * ls * p.inv_2n;
    const int ncc = ndc;  // 32-column chunks per warpgroup: (ndc * 64 / 2) / 32
    for (int cc = 0; cc < ncc; ++cc) {
      uint32_t v[32];
      const int cbase = half * (ndc * 32) + cc * 32;
      tmem_ld32(lane_addr + kDxCol + cbase, v);
      tmem_ld_wait();
      const int64_t d0 = (int64_t)dc0 * 64 + cbase;
      if (row < p.M && nsteps > 0) {
        if (p.nsplit > 1) {
          float* dst = p.acc_ws + row * p.D + d0;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (d0 + j < p.D) atomicAdd(dst + j, __uint_as_float(v[j]));
        } else if (d0 + 32 <= p.D) {
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.dX) + row * p.lddx + d0);
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 o;
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j + e]) * alpha;
            o.x = kBF16 ? pack_bf16x2(f[0], f[1]) : pack_f16x2(f[0], f[1]);
            o.y = kBF16 ? pack_bf16x2(f[2], f[3]) : pack_f16x2(f[2], f[3]);
            o.z = kBF16 ? pack_bf16x2(f[4], f[5]) : pack_f16x2(f[4], f[5]);
            o.w = kBF16 ? pack_bf16x2(f[6], f[7]) : pack_f16x2(f[6], f[7]);
            dst[j >> 3] = o;
          }
        } else {
          uint16_t* dst = reinterpret_cast<uint16_t*>(p.dX) + row * p.lddx + d0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (d0 + j < p.D) {
              const uint32_t pk = kBF16 ? pack_bf16x2(__uint_as_float(v[j]) * alpha, 0.f)
                                        : pack_f16x2(__uint_as_float(v[j]) * alpha, 0.f);
              dst[j] = (uint16_t)(pk & 0xFFFFu);
            }
          }
        }
      }
    }
    if (p.rowdot != nullptr && blockIdx.y == 0 && row < p.M) atomicAdd(p.rowdot + row, rd);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// f32 accumulation workspace -> dX (split-y mode)
template <typename T>
__global__ void acc_to_dx_kernel(const float* __restrict__ acc, int64_t M, int64_t D, const float* __restrict__ ls,
                                 const float* __restrict__ go, float inv_2n, T* __restrict__ dX, int64_t lddx) {
  const float alpha = (go ? go[0] : 1.f) * ls[0] * inv_2n;
  const int64_t n = M * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / D, d = i - r * D;
    dX[r * lddx + d] = from_f32<T>(acc[i] * alpha);
  }
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn* out) {
  static EncodeTiledFn cached = nullptr;
  if (!cached) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    MCLIP_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || fn == nullptr) {
      set_error("cuTensorMapEncodeTiled not available from the driver");
      return MCLIP_ERR_CUDA;
    }
    cached = reinterpret_cast<EncodeTiledFn>(fn);
  }
  *out = cached;
  return MCLIP_OK;
}

// [rows, D] row-major 16-bit matrix, box = [box_rows x 64 elements], 128-byte swizzle, zero fill.
int make_tmap(CUtensorMap* map, const void* base, int64_t rows, int64_t D, int64_t ld, int dtype, uint32_t box_rows) {
  EncodeTiledFn enc;
  int rc = get_encode_fn(&enc);
  if (rc) return rc;
  const cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = dtype == MCLIP_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = enc(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld D=%lld ld=%lld", (int)r, (long long)rows, (long long)D, (long long)ld);
    return MCLIP_ERR_CUDA;
  }
  return MCLIP_OK;
}

struct FwdPlan { bool xres; int bn; int stages; int kch; int tiles_total; int nsplit; int tiles_per_split; uint32_t smem; };

FwdPlan plan_fwd(int64_t M, int64_t N, int64_t D) {
  FwdPlan f;
  f.kch = (int)ceil_div(D, 64);
  f.xres = f.kch <= 8;
  f.bn = 256;
  const uint32_t avail = kSmemMax - kAlignSlack - kMiscBytes;
  const uint32_t x_bytes = f.xres ? f.kch * kChunkBytes : 0;
  const uint32_t stage = f.bn * 128 + (f.xres ? 0 : kChunkBytes);
  int st = (int)((avail - x_bytes) / stage);
  f.stages = st > (int)kMaxStages ? (int)kMaxStages : st;
  f.tiles_total = (int)ceil_div(N, f.bn);
  const int64_t m_tiles = ceil_div(M, 128);
  // splits: fill the 148 SMs, prefer wave counts that quantise well, never more than the tiles
  int best = 1;
  double best_cost = 1e30;
  const int max_split = f.tiles_total < 64 ? f.tiles_total : 64;
  for (int s = 1; s <= max_split; ++s) {
    const int tps = (int)ceil_div(f.tiles_total, s);
    const int real = (int)ceil_div(f.tiles_total, tps);
    if (real != s) continue;
    const int64_t ctas = m_tiles * s;
    const double waves = (double)ceil_div(ctas, 148);
    const double cost = waves * (tps + 1.5);  // +1.5 tile-times of per-CTA prologue (X load, TMEM alloc, drain)
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  f.nsplit = best;
  f.tiles_per_split = (int)ceil_div(f.tiles_total, best);
  f.smem = kAlignSlack + x_bytes + f.stages * stage + kMiscBytes;
  return f;
}

struct BwdPlan { bool xres; int stages; int kch; int dchunks; int steps_total; int nsplit; int steps_per_split; uint32_t smem; };

BwdPlan plan_bwd(int64_t M, int64_t N, int64_t D) {
  BwdPlan b;
  b.kch = (int)ceil_div(D, 64);
  b.xres = b.kch <= 8;
  b.dchunks = (int)ceil_div(b.kch, 4);
  const uint32_t avail = kSmemMax - kAlignSlack - kMiscBytes;
  const uint32_t x_bytes = b.xres ? b.kch * kChunkBytes : 0;
  const uint32_t stage = kChunkBytes + (b.xres ? 0 : kChunkBytes);
  int st = (int)((avail - x_bytes - 4 * kChunkBytes) / stage);
  b.stages = st > (int)kMaxStages ? (int)kMaxStages : st;
  b.steps_total = (int)ceil_div(N, 128);
  const int64_t items = ceil_div(M, 128) * b.dchunks;
  int best = 1;
  double best_cost = 1e30;
  const int max_split = b.steps_total < 32 ? b.steps_total : 32;
  for (int s = 1; s <= max_split; ++s) {
    const int sps = (int)ceil_div(b.steps_total, s);
    const int real = (int)ceil_div(b.steps_total, sps);
    if (real != s) continue;
    const double waves = (double)ceil_div(items * s, 148);
    const double cost = waves * (sps + 3.0) + (s > 1 ? 0.02 * b.steps_total : 0.0);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  b.nsplit = best;
  b.steps_per_split = (int)ceil_div(b.steps_total, best);
  b.smem = kAlignSlack + x_bytes + 4 * kChunkBytes + b.stages * stage + kMiscBytes;
  return b;
}

template <typename K>
int set_smem(K kernel, uint32_t bytes) {
  MCLIP_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return MCLIP_OK;
}

}  // namespace

bool tc_supported(int64_t M, int64_t N, int64_t D, int64_t ldx, int64_t ldy, int dtype, int op) {
  (void)M; (void)N; (void)op;
  if (dtype != MCLIP_DTYPE_BF16 && dtype != MCLIP_DTYPE_F16) return false;
  if (D % 8 != 0 || D > 64 * kMaxKch) return false;
  if (ldx % 8 != 0 || ldy % 8 != 0) return false;
  return true;
}

size_t tc_row_lse_ws(int64_t M, int64_t N, int64_t D) {
  const FwdPlan f = plan_fwd(M, N, D);
  return align_up((size_t)f.nsplit * M * 2 * sizeof(float), 256);
}

size_t tc_block_grad_ws(int64_t M, int64_t N, int64_t D) {
  const BwdPlan b = plan_bwd(M, N, D);
  return b.nsplit > 1 ? align_up((size_t)M * D * sizeof(float), 256) : 0;
}

int tc_row_lse(const RowLseArgs& a) {
  if (((uintptr_t)a.X | (uintptr_t)a.Y) & 15) { set_error("row_lse(tcgen05): X/Y must be 16-byte aligned"); return MCLIP_ERR_INVALID; }
  const FwdPlan f = plan_fwd(a.M, a.N, a.D);
  if (f.stages < 2) { set_error("row_lse(tcgen05): not enough shared memory for D=%lld", (long long)a.D); return MCLIP_ERR_UNSUPPORTED; }
  CUtensorMap tmX, tmY;
  int rc = make_tmap(&tmX, a.X, a.M, a.D, a.ldx, a.dtype, 128);
  if (rc) return rc;
  rc = make_tmap(&tmY, a.Y, a.N, a.D, a.ldy, a.dtype, (uint32_t)f.bn);
  if (rc) return rc;
  FwdParams p;
  p.M = a.M; p.N = a.N; p.kch = f.kch; p.stages = f.stages; p.tiles_total = f.tiles_total;
  p.tiles_per_split = f.tiles_per_split; p.diag_off = a.diag_off; p.ls = a.logit_scale;
  p.part_m2 = reinterpret_cast<float*>(a.ws);
  p.part_s = p.part_m2 + (size_t)f.nsplit * a.M;
  p.diag = a.diag; p.bf16 = a.dtype == MCLIP_DTYPE_BF16;
  if (a.diag) MCLIP_CUDA_OK(cudaMemsetAsync(a.diag, 0, sizeof(float) * a.M, a.stream));
  dim3 grid((unsigned)ceil_div(a.M, 128), (unsigned)f.nsplit);
  if (f.xres) {
    rc = set_smem(tc_row_lse_kernel<256, true>, f.smem);
    if (rc) return rc;
    tc_row_lse_kernel<256, true><<<grid, kThreads, f.smem, a.stream>>>(tmX, tmY, p);
  } else {
    rc = set_smem(tc_row_lse_kernel<256, false>, f.smem);
    if (rc) return rc;
    tc_row_lse_kernel<256, false><<<grid, kThreads, f.smem, a.stream>>>(tmX, tmY, p);
  }
  count_launch();
  MCLIP_CUDA_OK(cudaGetLastError());
  return launch_lse_merge(p.part_m2, p.part_s, f.nsplit, a.M, a.lse, a.stream);
}

int tc_block_grad(const BlockGradArgs& a) {
  if (((uintptr_t)a.X | (uintptr_t)a.Y | (uintptr_t)a.dX) & 15) { set_error("block_grad(tcgen05): X/Y/dX must be 16-byte aligned"); return MCLIP_ERR_INVALID; }
  if (!(a.w_row > 0.f) || a.w_col < 0.f) { set_error("block_grad(tcgen05): needs w_row > 0 and w_col >= 0"); return MCLIP_ERR_INVALID; }
  const BwdPlan b = plan_bwd(a.M, a.N, a.D);
  if (b.stages < 2) { set_error("block_grad(tcgen05): not enough shared memory for D=%lld", (long long)a.D); return MCLIP_ERR_UNSUPPORTED; }
  CUtensorMap tmX, tmY;
  int rc = make_tmap(&tmX, a.X, a.M, a.D, a.ldx, a.dtype, 128);
  if (rc) return rc;
  rc = make_tmap(&tmY, a.Y, a.N, a.D, a.ldy, a.dtype, 128);
  if (rc) return rc;
  BwdParams p;
  p.M = a.M; p.N = a.N; p.D = a.D; p.kch = b.kch; p.stages = b.stages; p.steps_total = b.steps_total;
  p.steps_per_split = b.steps_per_split; p.nsplit = b.nsplit; p.diag_off = a.diag_off; p.ls = a.logit_scale;
  p.go = a.grad_out; p.lse_x = a.lse_x; p.lse_y = a.lse_y; p.w_row = a.w_row; p.w_col = a.w_col;
  p.w_diag = a.w_diag; p.inv_2n = a.inv_2n; p.dX = a.dX; p.lddx = a.lddx;
  p.acc_ws = reinterpret_cast<float*>(a.ws); p.rowdot = a.rowdot;
  if (b.nsplit > 1) MCLIP_CUDA_OK(cudaMemsetAsync(a.ws, 0, sizeof(float) * a.M * a.D, a.stream));
  if (a.rowdot) MCLIP_CUDA_OK(cudaMemsetAsync(a.rowdot, 0, sizeof(float) * a.M, a.stream));
  dim3 grid((unsigned)ceil_div(a.M, 128), (unsigned)b.dchunks, (unsigned)b.nsplit);
  const bool bf = a.dtype == MCLIP_DTYPE_BF16;
#define MCLIP_LAUNCH_BWD(BF, XR)                                                              \
  do {                                                                                        \
    rc = set_smem(tc_block_grad_kernel<BF, XR>, b.smem);                                      \
    if (rc) return rc;                                                                        \
    tc_block_grad_kernel<BF, XR><<<grid, kThreads, b.smem, a.stream>>>(tmX, tmY, p);          \
  } while (0)
  if (bf && b.xres) MCLIP_LAUNCH_BWD(true, true);
  else if (bf) MCLIP_LAUNCH_BWD(true, false);
  else if (b.xres) MCLIP_LAUNCH_BWD(false, true);
  else MCLIP_LAUNCH_BWD(false, false);
#undef MCLIP_LAUNCH_BWD
  count_launch();
  MCLIP_CUDA_OK(cudaGetLastError());
  if (b.nsplit > 1) {
    const int64_t n = a.M * a.D;
    const unsigned blocks = (unsigned)(ceil_div(n, 256) < 148 * 8 ? ceil_div(n, 256) : 148 * 8);
    if (bf)
      acc_to_dx_kernel<__nv_bfloat16><<<blocks, 256, 0, a.stream>>>(p.acc_ws, a.M, a.D, a.logit_scale, a.grad_out,
                                                                   a.inv_2n, reinterpret_cast<__nv_bfloat16*>(a.dX), a.lddx);
    else
      acc_to_dx_kernel<__half><<<blocks, 256, 0, a.stream>>>(p.acc_ws, a.M, a.D, a.logit_scale, a.grad_out, a.inv_2n,
                                                            reinterpret_cast<__half*>(a.dX), a.lddx);
    count_launch();
    MCLIP_CUDA_OK(cudaGetLastError());
  }
  return MCLIP_OK;
}

}  // namespace mclip
