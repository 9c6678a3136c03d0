// Tensor-core path of the contrastive-loss kernels for B200 (sm_100a):
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> shared memory -> tcgen05.mma (bf16/f16, f32 accumulate
//   in TMEM) -> tcgen05.ld epilogue warps.  The logits block only ever exists as f32 tiles in TMEM.
//
// "One-sided" kernels (rows of X against all rows of Y), see DESIGN.md section 3:
//   tc_row_lse_kernel / tc_row_lse2_kernel   partial row (max, sum-exp) of ls * X Y^T  (forward: predicated fallback,
//                                            local_loss=False rows; the two-sided forward lives in tc_pair_lse.cu)
//   tc_block_grad2_kernel                    dX = alpha * G Y with G recomputed tile by tile, CTA pairs (backward);
//                                            kStoreG also hands G to the TN GEMM of tc_gemm_tn.cu (mclip_fused_grad)
//   tc_block_grad2p_kernel                   the same as a persistent kernel over (item, step) units (option bwd_persist)
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4..11 = epilogue (two warpgroups; warp w reads TMEM lanes 32*(w%4).. and one column half).
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tc_bwd_common.cuh"
#include "tc_host.cuh"

namespace mclip {

namespace {

using namespace ptx;


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
  float* part_c;      // sum of 2^(x-m) * <x_i, y_j> per split (null: not wanted)
  float* diag;
  int bf16;
  int dbg;            // development switch (MCLIP_DBG & 16): print barrier-wait cycle counts of a few CTAs
  const int* run_if;  // device flag (null = always run): 0 makes the whole grid exit before touching anything
};


// =================================================================================================
// forward
// =================================================================================================
template <bool kMasked, bool kDot>
__device__ __forceinline__ void fwd_chunk(const uint32_t (&v)[32], float k2, int64_t col0, int64_t N, int64_t jd,
                                          float& m2, float& sum, float& sc, float& diag_val) {
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
  const float rescale = ex2_approx(m2 - m_new);
  float s = sum * rescale;
  float c = kDot ? sc * rescale : 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float e = ex2_approx(x[j] - m_new);
    s += e;
    if (kDot) c = fmaf(e, __uint_as_float(v[j]), c);
  }
  sum = s;
  if (kDot) sc = c;
  m2 = m_new;
}

template <int BN, bool XRES>
__global__ void __launch_bounds__(kThreads, 1)
tc_row_lse_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  if (p.run_if != nullptr && *p.run_if == 0) return;
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
  float* merge_c = reinterpret_cast<float*>(misc_gen + 1536);   // [128]
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
    {
      // ---------------- MMA issuer: the whole warp waits, one elected lane issues ----------------
      const bool elected = elect_one();
      const uint32_t idesc = make_idesc_f16(p.bf16 != 0, p.bf16 != 0, 128, BN, false, false);
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
          if (elected) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = make_smem_desc_sw128(a_addr + k * 32, 0, 1024);
              const uint64_t bd = make_smem_desc_sw128(b_addr + k * 32, 0, 1024);
              mma_ss(d_tmem, ad, bd, idesc, (c | k) != 0);
            }
            mma_commit(empty_bar(s));
            if (c == p.kch - 1) mma_commit(tfull_bar(buf));
          }
          __syncwarp();
        }
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
    float m2 = -INFINITY, sum = 0.f, sc = 0.f, diag_val = 0.f;
    const bool want_dot = p.part_c != nullptr;
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
          if (col0 < p.N) {
            if (want_dot) fwd_chunk<true, true>(v, k2, col0, p.N, jd, m2, sum, sc, diag_val);
            else fwd_chunk<true, false>(v, k2, col0, p.N, jd, m2, sum, sc, diag_val);
          }
        } else {
          if (want_dot) fwd_chunk<false, true>(v, k2, col0, p.N, jd, m2, sum, sc, diag_val);
          else fwd_chunk<false, false>(v, k2, col0, p.N, jd, m2, sum, sc, diag_val);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(buf));
    }
    // merge the two column halves, then write the split's partial
    if (half == 1) { merge[row_in_tile] = make_float2(m2, sum); merge_c[row_in_tile] = sc; }
    named_bar_sync(1, kEpiThreads);
    if (half == 0) {
      const float2 o = merge[row_in_tile];
      const float mm = fmaxf(m2, o.x);
      float s = 0.f, c = 0.f;
      if (m2 > -INFINITY) { const float w = exp2f(m2 - mm); s += sum * w; c += sc * w; }
      if (o.x > -INFINITY) { const float w = exp2f(o.x - mm); s += o.y * w; c += merge_c[row_in_tile] * w; }
      if (row < p.M) {
        p.part_m2[(int64_t)blockIdx.y * p.M + row] = mm;
        p.part_s[(int64_t)blockIdx.y * p.M + row] = s;
        if (want_dot) p.part_c[(int64_t)blockIdx.y * p.M + row] = c;
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
// forward, CTA-pair version (D <= 512): cta_group::2 MMAs with M = 256 (128 rows per CTA), N = 256.
// Each CTA streams only its half of the Y tile (the pair shares the N operand), which halves the bytes an SM has
// to ingest per flop -- the single-CTA kernel needs ~62 B/clk/SM, right at the measured L2->SM limit.
// =================================================================================================
constexpr int kFwd2Stages = 6;

template <bool kUnused = true>
__global__ void __launch_bounds__(kThreads, 1)
tc_row_lse2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int BN = 256;
  if (p.run_if != nullptr && *p.run_if == 0) return;   // uniform over the grid: no CTA reaches the cluster barrier
  const uint32_t smem_base = align1024(smem_u32(smem_raw));
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t x_bytes = 8 * kChunkBytes;                       // [8][128 rows][64 k]
  const uint32_t ring_base = smem_base + x_bytes;                 // [kFwd2Stages][128 y][64 k]
  const uint32_t misc_base = ring_base + kFwd2Stages * kChunkBytes;
  uint8_t* misc_gen = smem_gen + (misc_base - smem_base);
  float2* merge = reinterpret_cast<float2*>(misc_gen);
  float* merge_c = reinterpret_cast<float*>(misc_gen + 1536);   // [128]
  const uint32_t bar_base = misc_base + 1024;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                 // leader
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };  // per CTA (multicast commit)
  const uint32_t xfull_bar = bar_base + 8u * (2 * kMaxStages);               // leader
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * kMaxStages + 1 + b); };   // per CTA (multicast commit)
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * kMaxStages + 3 + b); };  // leader: 16 epilogue warps
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 5);
  uint32_t* tmem_slot_gen = reinterpret_cast<uint32_t*>(misc_gen + 1024 + 8u * (2 * kMaxStages + 5));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int64_t m0 = (int64_t)(blockIdx.x >> 1) * 256 + 128 * rank;
  const int t0 = blockIdx.y * p.tiles_per_split;
  const int t1 = min(p.tiles_total, t0 + p.tiles_per_split);
  const int ntiles = t1 - t0;
  constexpr uint32_t kTmemCols = 512;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
    for (int s = 0; s < kFwd2Stages; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
    mbar_init(xfull_bar, 2);
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 2 * (kEpiThreads / 32)); }
    fence_barrier_init();
  } else if (warp == 2) {
    tmem_alloc_cg2(tmem_slot, kTmemCols);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    if (lane == 0) {
      if (leader) mbar_expect_tx(xfull_bar, 2 * x_bytes); else mbar_arrive_cluster(xfull_bar, 0);
      for (int c = 0; c < 8; ++c) tma_load_2d_cg2(smem_base + c * kChunkBytes, &tmX, c * 64, (int32_t)m0, xfull_bar);
      uint32_t it = 0;
      for (int t = t0; t < t1; ++t) {
        const int32_t y0 = t * BN + 128 * (int32_t)rank;     // this CTA's half of the tile's Y rows
        for (int c = 0; c < p.kch; ++c, ++it) {
          const int s = it % kFwd2Stages;
          const uint32_t ph = (it / kFwd2Stages) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          if (leader) mbar_expect_tx(full_bar(s), 2 * kChunkBytes); else mbar_arrive_cluster(full_bar(s), 0);
          tma_load_2d_cg2(ring_base + s * kChunkBytes, &tmY, c * 64, y0, full_bar(s));
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      const bool elected = elect_one();
      const uint32_t idesc = make_idesc_f16(p.bf16 != 0, p.bf16 != 0, 256, BN, false, false);
      mbar_wait(xfull_bar, 0);
      uint32_t it = 0;
      const bool prof = kProfile && (p.dbg & 16) != 0 && elected;
      long long t_full = 0, t_tempty = 0, t_begin = clock64();
      for (int lt = 0; lt < ntiles; ++lt) {
        const int buf = lt & 1;
        const uint32_t bph = (lt >> 1) & 1;
        {
          const long long t0 = prof ? clock64() : 0;
          mbar_wait(tempty_bar(buf), bph ^ 1);
          if (prof) t_tempty += clock64() - t0;
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int c = 0; c < p.kch; ++c, ++it) {
          const int s = it % kFwd2Stages;
          const uint32_t ph = (it / kFwd2Stages) & 1;
          {
            const long long t0 = prof ? clock64() : 0;
            mbar_wait(full_bar(s), ph);
            if (prof) t_full += clock64() - t0;
          }
          tc_fence_after();
          const uint32_t b_addr = ring_base + s * kChunkBytes;
          const uint32_t a_addr = smem_base + c * kChunkBytes;
          if (elected) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = make_smem_desc_sw128(a_addr + k * 32, 0, 1024);
              const uint64_t bd = make_smem_desc_sw128(b_addr + k * 32, 0, 1024);
              mma_ss_cg2(d_tmem, ad, bd, idesc, (c | k) != 0);
            }
            mma_commit_cg2(empty_bar(s), 3);
            if (c == p.kch - 1) mma_commit_cg2(tfull_bar(buf), 3);
          }
          __syncwarp();
        }
      }
      if (prof && blockIdx.x < 4 && blockIdx.y == 0)
        printf("[fwd mma cta %d] tiles=%d total=%lld clk wait_full=%lld wait_tempty=%lld (per tile: total %lld full %lld tempty %lld)\n",
               (int)blockIdx.x, ntiles, clock64() - t_begin, t_full, t_tempty, (clock64() - t_begin) / max(ntiles, 1),
               t_full / max(ntiles, 1), t_tempty / max(ntiles, 1));
    }
  } else if (warp >= kEpiWarp0) {
    const int ew = warp - kEpiWarp0;
    const int q = warp & 3;
    const int half = ew >> 2;
    const int row_in_tile = q * 32 + lane;
    const int64_t row = m0 + row_in_tile;
    const float k2 = p.ls[0] * kLog2e;
    const int64_t jd = row + p.diag_off;
    float m2 = -INFINITY, sum = 0.f, sc = 0.f, diag_val = 0.f;
    const bool want_dot = p.part_c != nullptr;
    constexpr int kHalfCols = BN / 2;
    const bool eprof = kProfile && (p.dbg & 16) != 0 && warp == kEpiWarp0 && lane == 0;
    long long e_wait = 0, e_begin = clock64();
    for (int lt = 0; lt < ntiles; ++lt) {
      const int buf = lt & 1;
      const uint32_t bph = (lt >> 1) & 1;
      const int64_t n0 = (int64_t)(t0 + lt) * BN;
      {
        const long long tt = eprof ? clock64() : 0;
        mbar_wait(tfull_bar(buf), bph);
        if (eprof) e_wait += clock64() - tt;
      }
      tc_fence_after();
      const bool special = (n0 + BN > p.N) || (p.diag != nullptr && n0 < m0 + p.diag_off + 128 && n0 + BN > m0 + p.diag_off);
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + half * kHalfCols;
#pragma unroll 1
      for (int cc = 0; cc < kHalfCols / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32(t_addr + cc * 32, v);
        tmem_ld_wait();
        const int64_t col0 = n0 + half * kHalfCols + cc * 32;
        if (special) {
          if (col0 < p.N) {
            if (want_dot) fwd_chunk<true, true>(v, k2, col0, p.N, jd, m2, sum, sc, diag_val);
            else fwd_chunk<true, false>(v, k2, col0, p.N, jd, m2, sum, sc, diag_val);
          }
        } else {
          if (want_dot) fwd_chunk<false, true>(v, k2, col0, p.N, jd, m2, sum, sc, diag_val);
          else fwd_chunk<false, false>(v, k2, col0, p.N, jd, m2, sum, sc, diag_val);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(tempty_bar(buf)); else mbar_arrive_cluster(tempty_bar(buf), 0);
      }
    }
    if (eprof && blockIdx.x < 4 && blockIdx.y == 0)
      printf("[fwd epi cta %d] total=%lld clk wait_tfull=%lld (per tile: total %lld wait %lld)\n", (int)blockIdx.x,
             clock64() - e_begin, e_wait, (clock64() - e_begin) / max(ntiles, 1), e_wait / max(ntiles, 1));
    if (half == 1) { merge[row_in_tile] = make_float2(m2, sum); merge_c[row_in_tile] = sc; }
    named_bar_sync(1, kEpiThreads);
    if (half == 0) {
      const float2 o = merge[row_in_tile];
      const float mm = fmaxf(m2, o.x);
      float s = 0.f, c = 0.f;
      if (m2 > -INFINITY) { const float w = exp2f(m2 - mm); s += sum * w; c += sc * w; }
      if (o.x > -INFINITY) { const float w = exp2f(o.x - mm); s += o.y * w; c += merge_c[row_in_tile] * w; }
      if (row < p.M) {
        p.part_m2[(int64_t)blockIdx.y * p.M + row] = mm;
        p.part_s[(int64_t)blockIdx.y * p.M + row] = s;
        if (want_dot) p.part_c[(int64_t)blockIdx.y * p.M + row] = c;
      }
    }
    if (p.diag != nullptr && row < p.M && jd >= 0 && jd < p.N) {
      const int64_t c_lo = (int64_t)t0 * BN, c_hi = (int64_t)t1 * BN;
      if (jd >= c_lo && jd < c_hi && (int)((jd % BN) / kHalfCols) == half) p.diag[row] = diag_val;
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, kTmemCols);
  }
}

// per-split f32 partials -> dX (and rowdot), summed in split order: deterministic, no atomics
template <typename T>
__global__ void acc_to_dx_kernel(const float* __restrict__ acc, const float* __restrict__ rd_ws, int nsplit, int64_t M,
                                 int64_t D, const float* __restrict__ ls, const float* __restrict__ go, float scale,
                                 T* __restrict__ dX, int64_t lddx, float* __restrict__ rowdot) {
  const float alpha = (go ? go[0] : 1.f) * ls[0] * scale;
  const int64_t n = M * D;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid * 4; i < n; i += nth * 4) {   // D % 8 == 0 on this path, so a float4 never straddles rows
    float4 a = *reinterpret_cast<const float4*>(acc + i);
    for (int s = 1; s < nsplit; ++s) {
      const float4 b = *reinterpret_cast<const float4*>(acc + (int64_t)s * n + i);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    const int64_t r = i / D, d = i - r * D;
    T* o = dX + r * lddx + d;
    o[0] = from_f32<T>(a.x * alpha); o[1] = from_f32<T>(a.y * alpha);
    o[2] = from_f32<T>(a.z * alpha); o[3] = from_f32<T>(a.w * alpha);
  }
  if (rowdot != nullptr) {
    for (int64_t i = tid; i < M; i += nth) {
      float t = rd_ws[i];
      for (int s = 1; s < nsplit; ++s) t += rd_ws[(int64_t)s * M + i];
      rowdot[i] = t;
    }
  }
}


// =================================================================================================
// backward, CTA-pair version (D <= 512): no second S recompute
// =================================================================================================
// A cluster of two CTAs owns 128 rows of X (64 per CTA) and issues cta_group::2 MMAs with M = 128.  In that
// shape each CTA's accumulator holds 64 rows with the N columns folded over the two lane halves (lanes 0-63:
// columns [0, N/2), lanes 64-127: [N/2, N)), so the whole dX block [64 x 512] f32 takes only 256 TMEM columns
// per CTA and still leaves room for two S buffers [64 x 256] (128 columns each).  Per 256-row step of Y:
//   S  = X Y^T        M=128 N=256 K=D      A: resident X tiles (K-major), B: Y tiles, N split over the pair
//   G  = f(S)         epilogue warps, f16 * 2^12, written to shared memory as the next A operand (K-major, SW128)
//   dX += G Y16       M=128 N=256 (x2 halves of D) K=256   B: Y16 tiles read MN-major, N split over the pair
// All Y traffic streams through one ring of 32 KB stages (two [128 x 64] tiles each).
struct Bwd2Params {
  int64_t M, N, D;
  int kpairs;           // ceil(ceil(D/64) / 2): S stages per step
  int ndh;              // ceil(D / 256): 256-wide halves of D
  int steps_total;      // ceil(N / 256)
  int steps_per_split;
  int nsplit;
  int64_t diag_off;
  const float* ls;
  const float* go;
  const float* lse_x;
  const float* ly2;     // [steps_total * 256] lse_y in log2 units, weight and G scale folded, +inf padded; null if w_col == 0
  const float* bcol;    // [steps_total * 256] 2^(mu0 - ly2[j]) (0 in the padding): column factor of the one-exp path
  const float* stepmm;  // [steps_total][2] min / max of ly2 over each 256-column step (valid columns only)
  const float* mu0;     // [1] midpoint of the ly2 range
  float w_row, w_diag, inv_2n;
  int has_col;
  void* dX;
  int64_t lddx;
  float* acc_ws;
  float* rd_ws;
  float* rowdot;
  int dbg;              // development switches (MCLIP_DBG): 1 = epilogue without math, 2 = no S MMAs, 4 = no dX MMAs, 8 = always two exps
  int g_tiles_per_row;  // kStoreG: 64-column tiles per 64-row block of the tile-major G scratch (= n_pad / 64)
  void* g_ptr;          // kStoreG: the scratch, [row blocks of 64][g_tiles_per_row][64 x 64 f16 in SWIZZLE_128B byte order]
};

// kStoreG: every G tile (f16 * 2^12, exactly what the dX MMA consumes) is also copied to global memory straight out of
// the shared-memory operand buffer (warps 2 and 3; the buffer is released by the dX commit AND the copies' reads).  The panel-wise shared-recompute backward (tc_fused_grad) feeds dY = G^T X from it, so
// that one S recompute serves both gradients.
// kNX = 8: D <= 512 -- X block = 8 tiles (64 KB), ring of 4 stages, TMEM = 2 S buffers (2 x 128 columns) + dX (256): the S
//          MMAs run two steps ahead of the dX MMAs.
// kNX = 12: 512 < D <= 768 -- X block = 12 tiles (96 KB), ring of 3 stages, TMEM = 1 S buffer (128) + dX (384): S runs one
//          step ahead; S(st+1) is issued as soon as the epilogue has pulled S(st) out of TMEM (`sread`), behind dX(st-1)
//          in the tensor pipe, so the single buffer costs no bubble.  Same per-flop operand traffic as D = 512.
template <bool kBF16, bool kStoreG, int kNX>
__global__ void __launch_bounds__(kThreads, 1)
tc_block_grad2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                      const __grid_constant__ CUtensorMap tmY16, const Bwd2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = align1024(smem_u32(smem_raw));
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  constexpr bool kBig = kNX > 8;
  constexpr int kRing = kBig ? 3 : kRing2;
  const uint32_t x_base = smem_base;                              // [kNX][64 rows][64 k]  64 / 96 KB
  const uint32_t g_base = x_base + kNX * kTile8K;                 // [4][64 rows][64 y]    32 KB (single buffer)
  const uint32_t ring_base = g_base + 4 * kTile8K;                // [kRing][2][128][64]   128 / 96 KB
  const uint32_t misc_base = ring_base + kRing * kStage2;
  uint8_t* misc_gen = smem_gen + (misc_base - smem_base);
  const uint32_t bar_base = misc_base;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };             // leader: both CTAs' TMA bytes
  auto empty_bar = [&](int s) { return bar_base + 8u * (4 + s); };      // per CTA, MMA commit multicast
  const uint32_t xfull_bar = bar_base + 8u * 8;                          // leader
  auto sfull_bar = [&](int b) { return bar_base + 8u * (9 + b); };      // per CTA, multicast
  const uint32_t gfull_bar = bar_base + 8u * 11;                         // leader: 16 epilogue warps
  const uint32_t gempty_bar = bar_base + 8u * 13;                        // per CTA, multicast
  const uint32_t dxfull_bar = bar_base + 8u * 15;                        // per CTA, multicast
  auto sread_bar = [&](int b) { return bar_base + 8u * (17 + b); };     // leader: 16 epilogue warps have loaded S(b)
  const uint32_t tmem_slot = bar_base + 8u * 19;
  uint32_t* tmem_slot_gen = reinterpret_cast<uint32_t*>(misc_gen + 8u * 19);
  const uint32_t gstore_bar = bar_base + 8u * 20;                        // per CTA: its 8 epilogue warps have written G
  // The S MMAs run two steps ahead of the dX MMAs (S(st+2) is issued as soon as the epilogue has pulled S(st)
  // out of TMEM), which gives the epilogue two S durations instead of one before dX(st) needs G(st).
  float* rd_scratch = reinterpret_cast<float*>(misc_gen + 256);          // [4][64]
  float* range_scratch = reinterpret_cast<float*>(misc_gen + 1280);      // [8][2]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();           // role inside the pair
  constexpr uint32_t leader_rank = 0;
  const bool leader = rank == 0;
  constexpr uint16_t kAllMask = 3, pair_mask = 3;
  const int64_t pair_row0 = (int64_t)(blockIdx.x / 2) * 128;
  const int64_t m0 = pair_row0 + 64 * rank;          // this CTA's 64 rows
  const int s0 = blockIdx.y * p.steps_per_split;
  const int s1 = min(p.steps_total, s0 + p.steps_per_split);
  const int nsteps = s1 - s0;
  constexpr uint32_t kTmemCols = 512;
  constexpr uint32_t kDxCol = kBig ? 128 : 256;
  // S buffer / barrier slot and mbarrier phase of step st
  auto sbuf = [](int st) { return kBig ? 0 : (st & 1); };
  auto sphase = [](int st) { return (uint32_t)(kBig ? (st & 1) : ((st >> 1) & 1)); };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
    tma_prefetch_desc(&tmY16);
    for (int s = 0; s < kRing; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
    mbar_init(xfull_bar, 2);
    for (int b = 0; b < 2; ++b) { mbar_init(sfull_bar(b), 1); mbar_init(sread_bar(b), 2 * (kEpiThreads / 32)); }
    mbar_init(gfull_bar, 2 * (kEpiThreads / 32));
    mbar_init(gempty_bar, kStoreG ? 3 : 1);     // dX commit (+ the two G-store warps)
    mbar_init(gstore_bar, kEpiThreads / 32);
    mbar_init(dxfull_bar, 1);
    fence_barrier_init();
  } else if (warp == 2) {
    tmem_alloc_cg2(tmem_slot, kTmemCols);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer (both CTAs; bytes are credited to the leader's barriers) ----------------
      if (leader) mbar_expect_tx(xfull_bar, 2 * kNX * kTile8K); else mbar_arrive_cluster(xfull_bar, leader_rank);
      for (int c = 0; c < kNX; ++c) tma_load_2d_cg2(x_base + c * kTile8K, &tmX, c * 64, (int32_t)m0, xfull_bar);
      uint32_t it = 0;
      const bool pprof = kProfile && (p.dbg & 16) != 0;
      long long p_empty = 0, p_begin = clock64();
      auto stage_begin = [&]() -> uint32_t {
        const int s = it % kRing;
        const uint32_t ph = (it / kRing) & 1;
        const long long t0 = pprof ? clock64() : 0;
        mbar_wait(empty_bar(s), ph ^ 1);
        if (pprof) p_empty += clock64() - t0;
        if (leader) mbar_expect_tx(full_bar(s), 2 * kStage2); else mbar_arrive_cluster(full_bar(s), leader_rank);
        ++it;
        return (uint32_t)s;
      };
      // one ring stage = two boxes
      auto load_box = [&](uint32_t dst, const CUtensorMap* tm, int e, int32_t c_inner, int32_t c_outer, uint32_t bar) {
        tma_load_2d_cg2(dst + e * kChunkBytes, tm, c_inner, c_outer, bar);
      };
      auto load_s = [&](int st) {   // Y tiles of step st as the N operand of S: this CTA's 128 rows, all of D
        const int32_t y0 = (s0 + st) * 256 + 128 * (int32_t)rank;
        for (int i = 0; i < p.kpairs; ++i) {
          const uint32_t s = stage_begin();
          const uint32_t dst = ring_base + s * kStage2;
          load_box(dst, &tmY, 0, (2 * i) * 64, y0, full_bar(s));
          load_box(dst, &tmY, 1, (2 * i + 1) * 64, y0, full_bar(s));
        }
      };
      auto load_dx = [&](int st) {  // Y16 tiles of step st as the [K = y][N = d] operand of dX: all 256 rows, this CTA's d
        for (int yh = 0; yh < 2; ++yh) {
          const int32_t y0 = (s0 + st) * 256 + 128 * yh;
          for (int h = 0; h < p.ndh; ++h) {
            const uint32_t s = stage_begin();
            const uint32_t dst = ring_base + s * kStage2;
            const int32_t dcol = (4 * h + 2 * (int32_t)rank) * 64;
            load_box(dst, &tmY16, 0, dcol, y0, full_bar(s));
            load_box(dst, &tmY16, 1, dcol + 64, y0, full_bar(s));
          }
        }
      };
      constexpr int kAhead = kBig ? 1 : 2;       // how many steps S runs ahead of dX
      for (int st = 0; st < kAhead && st < nsteps; ++st) load_s(st);
      for (int st = 0; st < nsteps; ++st) {
        if (st + kAhead < nsteps) load_s(st + kAhead);
        load_dx(st);
      }
      if (pprof && blockIdx.x < 4 && blockIdx.y == 0)
        printf("[tma cta %d] total=%lld clk  wait_empty=%lld (per step: total %lld empty %lld)\n", (int)blockIdx.x,
               clock64() - p_begin, p_empty, (clock64() - p_begin) / max(nsteps, 1), p_empty / max(nsteps, 1));
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------- MMA issuer (leader CTA only): the whole warp waits, one elected lane issues ----------------
      const bool elected = elect_one();
      const uint32_t idesc_s = make_idesc_f16(kBF16, kBF16, 128, 256, false, false);
      const uint32_t idesc_dx = make_idesc_f16(false, false, 128, 256, false, true);
      mbar_wait(xfull_bar, 0);
      uint32_t it = 0;
      const bool prof = kProfile && (p.dbg & 16) != 0 && elected;
      long long t_full = 0, t_gfull = 0, t_begin = clock64();
      auto stage_wait = [&]() -> uint32_t {
        const int s = it % kRing;
        const uint32_t ph = (it / kRing) & 1;
        const long long t0 = prof ? clock64() : 0;
        mbar_wait(full_bar(s), ph);
        if (prof) t_full += clock64() - t0;
        tc_fence_after();
        ++it;
        return (uint32_t)s;
      };
      auto issue_s = [&](int st) {
        const int buf = sbuf(st);
        // S buffer `buf` was last read by the epilogue of step st-2, whose gfull arrival the dX issue of that step
        // has already waited for; nothing more to wait on here.
        const uint32_t d_tmem = tmem_base + buf * 128;
        for (int i = 0; i < p.kpairs; ++i) {
          const uint32_t s = stage_wait();
          const uint32_t b_addr = ring_base + s * kStage2;
          if (elected) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t ad = make_smem_desc_sw128(x_base + (2 * i + e) * kTile8K + k * 32, 0, 1024);
                const uint64_t bd = make_smem_desc_sw128(b_addr + e * kChunkBytes + k * 32, 0, 1024);
                if (!(p.dbg & 2)) mma_ss_cg2(d_tmem, ad, bd, idesc_s, (i | e | k) != 0);
              }
            }
            mma_commit_cg2(empty_bar(s), kAllMask);
            if (i == p.kpairs - 1) mma_commit_cg2(sfull_bar(buf), pair_mask);
          }
          __syncwarp();
        }
      };
      auto issue_dx = [&](int st) {
        {
          const long long t0 = prof ? clock64() : 0;
          mbar_wait(gfull_bar, st & 1);
          if (prof) t_gfull += clock64() - t0;
        }
        tc_fence_after();
        for (int yh = 0; yh < 2; ++yh) {
          for (int h = 0; h < p.ndh; ++h) {
            const uint32_t s = stage_wait();
            const uint32_t b_addr = ring_base + s * kStage2;
            if (elected) {
#pragma unroll
              for (int kk = 0; kk < 8; ++kk) {
                // A = G[64 rows x 16 y] of K-chunk (2*yh + kk/4); B = Y16[16 y][128 d per CTA], MN-major
                const uint64_t bd = make_smem_desc_sw128(b_addr + kk * 2048, kChunkBytes, 1024);
                const uint64_t ad = make_smem_desc_sw128(g_base + (2 * yh + (kk >> 2)) * kTile8K + (kk & 3) * 32, 0, 1024);
                if (!(p.dbg & 4)) mma_ss_cg2(tmem_base + kDxCol + h * 128, ad, bd, idesc_dx, (st | yh | kk) != 0);
              }
              mma_commit_cg2(empty_bar(s), kAllMask);
              if (yh == 1 && h == p.ndh - 1) {
                mma_commit_cg2(gempty_bar, pair_mask);
                if (st == nsteps - 1) mma_commit_cg2(dxfull_bar, pair_mask);
              }
            }
            __syncwarp();
          }
        }
      };
      constexpr int kAhead = kBig ? 1 : 2;
      for (int st = 0; st < kAhead && st < nsteps; ++st) issue_s(st);
      for (int st = 0; st < nsteps; ++st) {
        if (st + kAhead < nsteps) {
          mbar_wait(sread_bar(sbuf(st)), sphase(st));   // S(st) is in registers: its TMEM buffer may be overwritten
          tc_fence_after();
          issue_s(st + kAhead);
        }
        issue_dx(st);
      }
      if (prof && blockIdx.x < 4 && blockIdx.y == 0)
        printf("[mma cta %d] steps=%d total=%lld clk  wait_full=%lld  wait_gfull=%lld  (per step: total %lld full %lld gfull %lld)\n",
               (int)blockIdx.x, nsteps, clock64() - t_begin, t_full, t_gfull, (clock64() - t_begin) / max(nsteps, 1),
               t_full / max(nsteps, 1), t_gfull / max(nsteps, 1));
    }
  } else if (kStoreG && (warp == 2 || warp == 3)) {
    // ---------------- G store (both CTAs, two warps): shared-memory operand tiles -> global tile-major scratch ----------------
    // Plain coalesced ld.shared / st.global copies, NOT TMA stores: the TMA unit of an SM is the resource this kernel is
    // bound by (64 B/clk of operand loads); 32 KB of TMA stores per step queued behind those loads cost ~1300 clk/step
    // (measured, profiles/r2_summary.md).  The four [64 x 64] f16 tiles of a step are contiguous both in shared memory
    // and in the scratch (tile-major), and are copied byte for byte, i.e. in their SWIZZLE_128B order: the consumer
    // (tc_gemm_tn_kernel) loads them back with an unswizzled TMA box and uses the same descriptors.
    const int sw = warp - 2;
    const bool sprof = kProfile && (p.dbg & 16) != 0 && lane == 0;
    long long s_wait = 0, s_copy = 0, s_begin = clock64();
    for (int st = 0; st < nsteps; ++st) {
      long long t0 = sprof ? clock64() : 0;
      mbar_wait(gstore_bar, st & 1);            // this CTA's 8 epilogue warps wrote G(st)
      if (sprof) { s_wait += clock64() - t0; t0 = clock64(); }
      uint8_t* gdst = reinterpret_cast<uint8_t*>(p.g_ptr) +
                      (((size_t)(m0 >> 6) * p.g_tiles_per_row + (size_t)(s0 + st) * 4) << 13) + lane * 16;
      const uint32_t gsrc = g_base + lane * 16;
#pragma unroll 8
      for (int i = sw; i < 64; i += 2) {
        uint32_t a, b, c, d;
        ld_shared_v4(gsrc + i * 512, a, b, c, d);
        asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(gdst + i * 512), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
      }
      if (sprof) s_copy += clock64() - t0;
      __syncwarp();
      if (lane == 0) mbar_arrive(gempty_bar);   // the tiles have been read out of shared memory
    }
    if (sprof && sw == 1 && blockIdx.x < 4 && blockIdx.y == 0)
      printf("[gst cta %d] total=%lld clk  wait_gstore=%lld  copy=%lld (per step: total %lld gstore %lld copy %lld)\n",
             (int)blockIdx.x, clock64() - s_begin, s_wait, s_copy, (clock64() - s_begin) / max(nsteps, 1), s_wait / max(nsteps, 1),
             s_copy / max(nsteps, 1));
  } else if (warp >= kEpiWarp0) {
    // ---------------- epilogue: S -> G (f16 * 2^12) into shared memory; finally dX out ----------------
    const int ew = warp - kEpiWarp0;
    const int q = warp & 3;                   // TMEM lane quadrant
    const int half = ew >> 2;                 // which 64 of this lane-half's 128 columns
    const int r = (q & 1) * 32 + lane;        // row within this CTA's 64
    const int cS = (q >> 1) * 128 + half * 64;   // first S column (of 256) this thread handles
    const int kc = cS >> 6;                   // K-chunk of G it fills
    const int64_t row = m0 + r;
    const float ls = p.ls[0];
    const float k2 = ls * kLog2e;
    const bool has_col = p.has_col != 0;
    // rows past M reuse the last valid row's LSE so that they do not widen the range check below
    const float lx2 = p.lse_x[row < p.M ? row : p.M - 1] * kLog2e - (log2f(p.w_row) + 12.f);
    const float w_diag_s = p.w_diag * kGScale;
    const int64_t jd = row + p.diag_off;
    float rd = 0.f;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int64_t blk_lo = pair_row0 + p.diag_off;   // diagonal columns of the pair's rows
    // one-exp path set-up: a_i and the range of lx2 over this CTA's rows
    float mu0 = 0.f, a_i = 0.f, lx_min = 0.f, lx_max = 0.f;
    bool a_ok = false;
    if (has_col) {
      mu0 = p.mu0[0];
      a_i = ex2_approx(lx2 - mu0);
      float mn = lx2, mx = lx2;
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      }
      if (lane == 0) { range_scratch[2 * ew] = mn; range_scratch[2 * ew + 1] = mx; }
      named_bar_sync(1, kEpiThreads);
      lx_min = range_scratch[0]; lx_max = range_scratch[1];
#pragma unroll
      for (int w = 1; w < 8; ++w) {
        lx_min = fminf(lx_min, range_scratch[2 * w]);
        lx_max = fmaxf(lx_max, range_scratch[2 * w + 1]);
      }
      a_ok = fabsf(lx_max - mu0) <= 120.f && fabsf(lx_min - mu0) <= 120.f;
    }
    const bool eprof = kProfile && (p.dbg & 16) != 0 && warp == kEpiWarp0 && lane == 0;
    long long e_sfull = 0, e_gempty = 0, e_begin = clock64();
    for (int st = 0; st < nsteps; ++st) {
      const int buf = sbuf(st);
      const uint32_t bph = sphase(st);
      const int64_t n0 = (int64_t)(s0 + st) * 256;
      {
        const long long t0 = eprof ? clock64() : 0;
        mbar_wait(sfull_bar(buf), bph);
        if (eprof) e_sfull += clock64() - t0;
      }
      tc_fence_after();
      const bool special = (n0 + 256 > p.N) || (n0 < blk_lo + 128 && n0 + 256 > blk_lo);
      bool fast = false;
      if (has_col && a_ok && !(p.dbg & 8)) {
        const float2 mm = __ldg(reinterpret_cast<const float2*>(p.stepmm) + (s0 + st));   // (min, max) of ly2 in this step
        fast = (lx_max - mm.x <= 100.f) && (mm.y - lx_min <= 100.f) &&
               (!(mm.x <= mm.y) || (fabsf(mm.x - mu0) <= 120.f && fabsf(mm.y - mu0) <= 120.f));
      }
      const uint32_t g_row = g_base + kc * kTile8K + r * 128;
      uint32_t g[2][16];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t v[32];
        tmem_ld32(lane_addr + buf * 128 + half * 64 + cc * 32, v);
        tmem_ld_wait();
        if (cc == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(sread_bar(buf)); else mbar_arrive_cluster(sread_bar(buf), leader_rank);
          }
        }
        const int64_t col0 = n0 + cS + cc * 32;
        const float* ly2 = has_col ? p.ly2 + col0 : nullptr;
        if (p.dbg & 1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) g[cc][j] = v[2 * j] & 0x3c003c00u;
        } else if (fast) {
          if (special) bwd2_chunk_fast<true>(v, g[cc], k2, lx2, a_i, p.bcol + col0, w_diag_s, col0, p.N, jd, rd);
          else bwd2_chunk_fast<false>(v, g[cc], k2, lx2, a_i, p.bcol + col0, w_diag_s, col0, p.N, jd, rd);
        } else if (special) {
          if (has_col) bwd2_chunk<true, true>(v, g[cc], k2, lx2, ly2, w_diag_s, col0, p.N, jd, rd);
          else bwd2_chunk<true, false>(v, g[cc], k2, lx2, ly2, w_diag_s, col0, p.N, jd, rd);
        } else {
          if (has_col) bwd2_chunk<false, true>(v, g[cc], k2, lx2, ly2, w_diag_s, col0, p.N, jd, rd);
          else bwd2_chunk<false, false>(v, g[cc], k2, lx2, ly2, w_diag_s, col0, p.N, jd, rd);
        }
      }
      tc_fence_before();          // TMEM reads of S are complete
      // G is single-buffered: the dX MMAs of the previous step must have finished reading it.  The values are
      // already in registers, so this wait overlaps with the S MMAs of the next step on the tensor pipe.
      {
        const long long t0 = eprof ? clock64() : 0;
        mbar_wait(gempty_bar, (st & 1) ^ 1);
        if (eprof) e_gempty += clock64() - t0;
      }
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        // 32 f16 = four 16-byte pieces (cc*4 .. cc*4+3) of this row's 128-byte line, 128B-swizzled
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) {
          const uint32_t piece = (uint32_t)((cc * 4 + pc) ^ (r & 7));
          st_shared_v4(g_row + piece * 16, g[cc][4 * pc], g[cc][4 * pc + 1], g[cc][4 * pc + 2], g[cc][4 * pc + 3]);
        }
      }
      fence_proxy_async_smem();   // G visible to the tensor-core / TMA (async) proxy
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(gfull_bar); else mbar_arrive_cluster(gfull_bar, leader_rank);
        if (kStoreG) mbar_arrive(gstore_bar);
      }
    }
    if (eprof && blockIdx.x < 4 && blockIdx.y == 0)
      printf("[epi cta %d] total=%lld clk  wait_sfull=%lld  wait_gempty=%lld (per step: total %lld sfull %lld gempty %lld)\n",
             (int)blockIdx.x, clock64() - e_begin, e_sfull, e_gempty, (clock64() - e_begin) / max(nsteps, 1),
             e_sfull / max(nsteps, 1), e_gempty / max(nsteps, 1));
    // ---- dX accumulator -> global ----
    mbar_wait(dxfull_bar, 0);
    tc_fence_after();
    const float alpha = (p.go ? p.go[0] : 1.f) * ls * p.inv_2n * (1.f / kGScale);
    for (int h = 0; h < p.ndh; ++h) {
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t v[32];
        tmem_ld32(lane_addr + kDxCol + h * 128 + half * 64 + cc * 32, v);
        tmem_ld_wait();
        const int64_t d0 = (int64_t)h * 256 + (q >> 1) * 128 + half * 64 + cc * 32;
        if (row < p.M && nsteps > 0 && d0 < p.D) {
          if (p.nsplit > 1) {
            float* dst = p.acc_ws + ((int64_t)blockIdx.y * p.M + row) * p.D + d0;
            if (d0 + 32 <= p.D) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<uint4*>(dst + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (d0 + j < p.D) dst[j] = __uint_as_float(v[j]);
            }
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
    }
    // rowdot: four threads share a row (2 lane halves x 2 column halves); add in fixed order via shared memory
    if (p.rowdot != nullptr) {
      rd_scratch[((q >> 1) * 2 + half) * 64 + r] = rd;
      named_bar_sync(1, kEpiThreads);
      if ((q >> 1) == 0 && half == 0 && row < p.M) {
        const float tot = ((rd_scratch[r] + rd_scratch[64 + r]) + (rd_scratch[128 + r] + rd_scratch[192 + r])) * (1.f / kGScale);
        if (p.nsplit > 1) p.rd_ws[(int64_t)blockIdx.y * p.M + row] = tot;
        else p.rowdot[row] = tot;
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, kTmemCols);
  }
}

// Column statistics for the backward kernel.  Every block first reduces the global (min, max) of ly2 itself (N floats
// from L2 -- cheaper than a second launch or a grid-wide barrier), then handles its share of the 256-column steps:
//   ly2[j]  = lse_y[j] * log2e - lw_col              (+inf in the padding up to a multiple of 256)
//   mu0     = midpoint of the ly2 range;  bcol[j] = 2^(mu0 - ly2[j])  (0 in the padding)
//   stepmm  = (min, max) of ly2 over each 256-column step
__global__ void __launch_bounds__(1024)
prep_ly2_kernel(const float* __restrict__ lse_y, int64_t N, int64_t n_pad, float lw_col, float* __restrict__ ly2,
                float* __restrict__ bcol, float* __restrict__ stepmm, float* __restrict__ mu0_out) {
  __shared__ float red_min[32], red_max[32];
  __shared__ float mu_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float mn = INFINITY, mx = -INFINITY;
#pragma unroll 8
  for (int64_t j = tid; j < N; j += 1024) {
    const float v = lse_y[j] * kLog2e - lw_col;
    mn = fminf(mn, v); mx = fmaxf(mx, v);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (lane == 0) { red_min[warp] = mn; red_max[warp] = mx; }
  __syncthreads();
  if (warp == 0) {
    mn = red_min[lane]; mx = red_max[lane];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) { mu_s = 0.5f * mn + 0.5f * mx; if (blockIdx.x == 0) mu0_out[0] = mu_s; }
  }
  __syncthreads();
  const float mu = mu_s;
  const int64_t nsteps = n_pad / 256;
  for (int64_t st = (int64_t)blockIdx.x * 32 + warp; st < nsteps; st += (int64_t)gridDim.x * 32) {
    float smn = INFINITY, smx = -INFINITY;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int64_t j = st * 256 + e * 32 + lane;
      const bool ok = j < N;
      const float v = ok ? lse_y[j] * kLog2e - lw_col : INFINITY;
      ly2[j] = v;
      bcol[j] = ok ? exp2f(mu - v) : 0.f;
      if (ok) { smn = fminf(smn, v); smx = fmaxf(smx, v); }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      smn = fminf(smn, __shfl_xor_sync(0xffffffffu, smn, o));
      smx = fmaxf(smx, __shfl_xor_sync(0xffffffffu, smx, o));
    }
    if (lane == 0) { stepmm[2 * st] = smn; stepmm[2 * st + 1] = smx; }
  }
}

// bf16 -> f16 copy of Y (exact for 6.1e-5 <= |v| <= 65504; saturating above, f16-subnormal below)
__global__ void bf16_to_f16_kernel(const __nv_bfloat16* __restrict__ src, int64_t rows, int64_t D, int64_t ld,
                                   __half* __restrict__ dst) {
  const int64_t n8 = rows * (D / 8);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (D / 8), c = (i - r * (D / 8)) * 8;
    const uint4 in = *reinterpret_cast<const uint4*>(src + r * ld + c);
    const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&in);
    uint4 out;
    __half2* h = reinterpret_cast<__half2*>(&out);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 f = __bfloat1622float2(b[e]);
      f.x = fminf(fmaxf(f.x, -65504.f), 65504.f);
      f.y = fminf(fmaxf(f.y, -65504.f), 65504.f);
      h[e] = __floats2half2_rn(f.x, f.y);
    }
    *reinterpret_cast<uint4*>(dst + r * D + c) = out;
  }
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Development switches are read from the environment ONCE, when the library is loaded (never on the launch path), and
// can be changed afterwards through mclip_set_option (tests do that instead of re-reading the environment).
struct Options {
  int bwd_persist;   // MCLIP_BWD_PERSIST: persistent CTA-pair backward kernel: -1 = by shape (default), 0 = never, 1 = always
  int dbg;           // MCLIP_DBG: development masks; only honoured by -DMCLIP_PROFILE builds
  int fused_bwd;     // MCLIP_FUSED_BWD: shared-recompute backward (one S recompute for dX and dY) where it applies
  Options() {
    auto geti = [](const char* k, int dflt) { const char* e = getenv(k); return e ? atoi(e) : dflt; };
    bwd_persist = geti("MCLIP_BWD_PERSIST", -1);
    dbg = kProfile ? geti("MCLIP_DBG", 0) : 0;
    fused_bwd = geti("MCLIP_FUSED_BWD", 1);
  }
};
Options& options() {
  static Options o;
  return o;
}

// SM count of the current device (cudaDeviceProp, cached per device); a CTA pair occupies two SMs.
int sm_count() {
  static std::mutex mu;
  static std::unordered_map<int, int> cache;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(dev);
  if (it != cache.end()) return it->second;
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 2) { (void)cudaGetLastError(); n = 148; }
  cache[dev] = n;
  return n;
}
int pair_slots() { return sm_count() / 2; }
unsigned prep_blocks(int64_t n_pad) {   // prep_ly2_kernel: one warp per 256-column step, ~8 of its 32 warps busy per block
  const int64_t want = ceil_div(n_pad / 256, 8);
  return (unsigned)(want < 1 ? 1 : (want > 64 ? 64 : want));
}
unsigned ew_blocks(int64_t work_items, int per_block) {   // grid of an elementwise helper: at most 8 blocks per SM
  const int64_t want = ceil_div(work_items, per_block), cap = (int64_t)sm_count() * 8;
  return (unsigned)(want < cap ? want : cap);
}

// Driver-API calls (cuTensorMapEncodeTiled) need the primary context bound to the calling thread; autograd's
// backward threads may not have touched the runtime yet.  cudaFree(0) binds it (no-op otherwise).
int bind_context() {
  static thread_local bool bound = false;
  if (!bound) {
    MCLIP_CUDA_OK(cudaFree(0));
    bound = true;
  }
  return MCLIP_OK;
}

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
  int rc = bind_context();
  if (rc) return rc;
  rc = get_encode_fn(&enc);
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
  // splits: fill the SMs, prefer wave counts that quantise well, never more than the tiles
  int best = 1;
  double best_cost = 1e30;
  const int max_split = f.tiles_total < 64 ? f.tiles_total : 64;
  for (int s = 1; s <= max_split; ++s) {
    const int tps = (int)ceil_div(f.tiles_total, s);
    const int real = (int)ceil_div(f.tiles_total, tps);
    if (real != s) continue;
    const int64_t ctas = m_tiles * s;
    const double waves = (double)ceil_div(ctas, sm_count());
    const double cost = waves * (tps + 1.5);  // +1.5 tile-times of per-CTA prologue (X load, TMEM alloc, drain)
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  f.nsplit = best;
  f.tiles_per_split = (int)ceil_div(f.tiles_total, best);
  f.smem = kAlignSlack + x_bytes + f.stages * stage + kMiscBytes;
  return f;
}

bool fwd_uses_pair(int64_t D) { return D <= 512; }

FwdPlan plan_fwd2(int64_t M, int64_t N, int64_t D) {
  FwdPlan f;
  f.kch = (int)ceil_div(D, 64);
  f.xres = true;
  f.bn = 256;
  f.stages = kFwd2Stages;
  f.tiles_total = (int)ceil_div(N, f.bn);
  const int64_t pairs = ceil_div(M, 256);
  int best = 1;
  double best_cost = 1e30;
  const int max_split = f.tiles_total < 64 ? f.tiles_total : 64;
  for (int s = 1; s <= max_split; ++s) {
    const int tps = (int)ceil_div(f.tiles_total, s);
    const int real = (int)ceil_div(f.tiles_total, tps);
    if (real != s) continue;
    const double waves = (double)ceil_div(pairs * s, pair_slots());
    const double cost = waves * (tps + 1.5);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  f.nsplit = best;
  f.tiles_per_split = (int)ceil_div(f.tiles_total, best);
  f.smem = kAlignSlack + 8 * kChunkBytes + kFwd2Stages * kChunkBytes + kMiscBytes;
  return f;
}

struct Bwd2Plan { int kch; int kpairs; int ndh; int steps_total; int nsplit; int steps_per_split; uint32_t smem; double cost; };

// (Variants measured and removed in round 2, all numerically identical -- numbers in profiles/r1_ncu_summary.md section 3:
// 4-CTA clusters with multicast Y tiles 1.95 vs 1.85 ms; G handed over through TMEM 1.61 vs 1.53 ms; transposed pair
// kernel with a DSMEM exchange 1.98 vs 1.65 ms; single-CTA kernel 4.00 ms at D = 512 and 68.5 vs 25.1 ms per C4 step at D = 768.)
Bwd2Plan plan_bwd2(int64_t M, int64_t N, int64_t D) {
  Bwd2Plan b;
  b.kch = (int)ceil_div(D, 64);
  b.kpairs = (b.kch + 1) / 2;
  b.ndh = (int)ceil_div(D, 256);
  b.steps_total = (int)ceil_div(N, 256);
  const int64_t pairs = ceil_div(M, 128);                  // clusters
  const int per_wave = pair_slots();                       // clusters resident at once
  int best = 1;
  double best_cost = 1e30;
  const int max_split = b.steps_total < 32 ? b.steps_total : 32;
  for (int s = 1; s <= max_split; ++s) {
    const int sps = (int)ceil_div(b.steps_total, s);
    const int real = (int)ceil_div(b.steps_total, sps);
    if (real != s) continue;
    const double waves = (double)ceil_div(pairs * s, per_wave);
    const double cost = waves * (sps + 2.0) + (s > 1 ? 0.02 * b.steps_total : 0.0);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  b.nsplit = best;
  b.cost = best_cost;
  b.steps_per_split = (int)ceil_div(b.steps_total, best);
  b.smem = D <= 512 ? kAlignSlack + 8 * kTile8K + 4 * kTile8K + kRing2 * kStage2 + 1536     // 64 KB X + 32 KB G + 4 x 32 KB ring
                    : kAlignSlack + 12 * kTile8K + 4 * kTile8K + 3 * kStage2 + 1536;        // 96 KB X + 32 KB G + 3 x 32 KB ring
  return b;
}

// Persistent pair kernel (tc_bwd_persist.cu): P pairs walk U = row_blocks x steps units, so the launch has no wave
// quantisation, at the price of a per-pair item switch (X reload, accumulator drain) and f32 partials + a fix-up pass for
// the row blocks a range boundary cuts.  Option bwd_persist: 1 = always, 0 = never, -1 (default) = by shape, taken when
//   (a) the best split grid leaves more than 8 % of its pair-slot time idle,
//   (b) a pair's range is long (>= 96 steps) and
//   (c) a row block is not shared by many pairs (steps per row block <= 1.5 x steps per pair).
// Measured whole calls, split grid vs persistent (B200, bf16, D = 512, us; profiles/r2_summary.md section 11):
//   M x N          32768x32768  16384x32768  20480x32768  10240x32768  8192x32768  16384x16384 | 4096x32768  4096x65536
//   split grid        1470          780         1003          578         460          452     |    232         442
//   persistent        1500          774          968          522         384          403     |    260         526
//   rule picks        split        split      persistent   persistent  persistent  persistent  |   split       split
// (8192 x 32768 is a rank of C3 on 4 GPUs: 64 row blocks for 74 pair slots.  The two shapes on the right fail (c): with 32
// row blocks every one is cut two or three times and the persistent kernel loses.)
double plan_bwd2_cost(int64_t M, int64_t N, int64_t D) { return plan_bwd2(M, N, D).cost; }
bool use_persistent_bwd(int64_t M, int64_t N, int64_t D) {
  const int mode = options().bwd_persist;
  if (mode >= 0) return mode != 0;
  if (D > 512) return false;
  const double units = (double)ceil_div(M, 128) * (double)ceil_div(N, 256);
  const double per_pair = units / pair_slots();
  const double steps = (double)ceil_div(N, 256);
  return per_pair >= 96.0 && steps <= 1.5 * per_pair && plan_bwd2_cost(M, N, D) > 1.08 * (per_pair + 4.0);
}

// workspace carve-up shared by the size query and the launcher
struct BwdWs { size_t acc, rd, ly2, y16, total; };
BwdWs bwd_ws_layout(int nsplit, int64_t M, int64_t N, int64_t D, int64_t n_pad, bool need_y16) {
  BwdWs w;
  size_t off = 0;
  w.acc = off; off += nsplit > 1 ? align_up((size_t)nsplit * M * D * sizeof(float), 256) : 0;
  w.rd = off;  off += nsplit > 1 ? align_up((size_t)nsplit * M * sizeof(float), 256) : 0;
  w.ly2 = off; off += align_up(((size_t)2 * n_pad + 2 * (n_pad / 256) + 64) * sizeof(float), 256);  // ly2, bcol, stepmm, mu0
  w.y16 = off; off += need_y16 ? align_up((size_t)N * D * 2, 256) : 0;
  w.total = off;
  return w;
}

// cudaFuncSetAttribute once per (kernel, device) and size high-water mark instead of on every launch
template <typename K>
int set_smem(K kernel, uint32_t bytes) {
  static std::mutex mu;
  static std::unordered_map<uint64_t, uint32_t> done;
  int dev = 0;
  MCLIP_CUDA_OK(cudaGetDevice(&dev));
  const uint64_t key = (uint64_t)reinterpret_cast<uintptr_t>(reinterpret_cast<const void*>(kernel)) ^ ((uint64_t)dev << 56);
  std::lock_guard<std::mutex> lock(mu);
  auto it = done.find(key);
  if (it != done.end() && it->second >= bytes) return MCLIP_OK;
  MCLIP_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  done[key] = bytes;
  return MCLIP_OK;
}

}  // namespace

// ---- optional in-situ timing of the dominant kernel (bench.py's roofline leg) --------------------------------
// While enabled, every launch of the pair backward kernel is bracketed by two CUDA events on its own stream; reading
// synchronises those events and returns the summed duration.  Off by default: the product path records nothing.
namespace {
std::mutex g_timing_mu;
bool g_timing_on = false;
std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_timing_events;
}  // namespace

void kernel_timing_enable(bool on) {
  std::lock_guard<std::mutex> lock(g_timing_mu);
  g_timing_on = on;
}

int kernel_timing_read(float* total_ms, int* count) {
  std::lock_guard<std::mutex> lock(g_timing_mu);
  float tot = 0.f;
  int n = 0;
  for (auto& ev : g_timing_events) {
    float ms = 0.f;
    if (cudaEventSynchronize(ev.second) == cudaSuccess && cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) {
      tot += ms;
      ++n;
    }
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  g_timing_events.clear();
  (void)cudaGetLastError();
  if (total_ms) *total_ms = tot;
  if (count) *count = n;
  return MCLIP_OK;
}

namespace {
thread_local int g_timer_suppress = 0;   // > 0 inside a bracket that already covers the launches (mclip_fused_grad)
struct TimerSuppress {
  TimerSuppress() { ++g_timer_suppress; }
  ~TimerSuppress() { --g_timer_suppress; }
};
struct ScopedKernelTimer {
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t stream;
  bool active = false;
  explicit ScopedKernelTimer(cudaStream_t s) : stream(s) {
    std::lock_guard<std::mutex> lock(g_timing_mu);
    if (!g_timing_on || g_timer_suppress > 0) return;
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) return;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
    active = cudaEventRecord(a, s) == cudaSuccess;
  }
  ~ScopedKernelTimer() {
    if (!active) return;
    cudaEventRecord(b, stream);
    std::lock_guard<std::mutex> lock(g_timing_mu);
    g_timing_events.emplace_back(a, b);
  }
};
}  // namespace

int tc_make_tmap(CUtensorMap* map, const void* base, int64_t rows, int64_t D, int64_t ld, int dtype, uint32_t box_rows) {
  return make_tmap(map, base, rows, D, ld, dtype, box_rows);
}

int tc_make_tmap_f32(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, uint32_t box_cols, uint32_t box_rows) {
  EncodeTiledFn enc;
  int rc = bind_context();
  if (rc) return rc;
  rc = get_encode_fn(&enc);
  if (rc) return rc;
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(f32) failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, (long long)rows, (long long)cols, (long long)ld);
    return MCLIP_ERR_CUDA;
  }
  return MCLIP_OK;
}

int tc_make_tmap_tiles(CUtensorMap* map, const void* base, int64_t ntiles) {
  EncodeTiledFn enc;
  int rc = bind_context();
  if (rc) return rc;
  rc = get_encode_fn(&enc);
  if (rc) return rc;
  const cuuint64_t gdim[3] = {64, 64, (cuuint64_t)ntiles};
  const cuuint64_t gstride[2] = {128, 8192};
  const cuuint32_t box[3] = {64, 64, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(tiles) failed (%d) ntiles=%lld", (int)r, (long long)ntiles);
    return MCLIP_ERR_CUDA;
  }
  return MCLIP_OK;
}

int tc_set_smem(const void* kernel, uint32_t bytes) { return set_smem(kernel, bytes); }

int tc_set_option(const char* name, int value) {
  if (!name) return MCLIP_ERR_INVALID;
  Options& o = options();
  const std::string k(name);
  if (k == "bwd_persist") o.bwd_persist = value;
  else if (k == "dbg") o.dbg = kProfile ? value : 0;
  else if (k == "fused_bwd") o.fused_bwd = value;
  else { set_error("unknown option '%s'", name); return MCLIP_ERR_INVALID; }
  return MCLIP_OK;
}
int tc_get_option(const char* name, int* value) {
  if (!name || !value) return MCLIP_ERR_INVALID;
  const Options& o = options();
  const std::string k(name);
  if (k == "bwd_persist") *value = o.bwd_persist;
  else if (k == "dbg") *value = o.dbg;
  else if (k == "fused_bwd") *value = o.fused_bwd;
  else { set_error("unknown option '%s'", name); return MCLIP_ERR_INVALID; }
  return MCLIP_OK;
}

bool tc_supported(int64_t M, int64_t N, int64_t D, int64_t ldx, int64_t ldy, int dtype, int op) {
  (void)M; (void)N; (void)op;
  if (dtype != MCLIP_DTYPE_BF16 && dtype != MCLIP_DTYPE_F16) return false;
  if (D % 8 != 0 || D > 64 * kMaxKch) return false;
  if (ldx % 8 != 0 || ldy % 8 != 0) return false;
  return true;
}

size_t tc_row_lse_ws(int64_t M, int64_t N, int64_t D) {
  const FwdPlan f = fwd_uses_pair(D) ? plan_fwd2(M, N, D) : plan_fwd(M, N, D);
  return align_up((size_t)f.nsplit * M * 3 * sizeof(float), 256);
}

size_t tc_block_grad_ws(int64_t M, int64_t N, int64_t D) {
  // dtype is not known here: always reserve room for the f16 copy of Y; cover both pair kernels
  const Bwd2Plan b = plan_bwd2(M, N, D);
  const int64_t n_pad = (int64_t)b.steps_total * 256;
  const size_t v2 = bwd_ws_layout(b.nsplit, M, N, D, n_pad, true).total;
  if (D > 512) return v2;
  const size_t vp = tc_block_grad2p_ws(N, D, n_pad);
  return v2 > vp ? v2 : vp;
}

int tc_row_lse(const RowLseArgs& a) {
  if (((uintptr_t)a.X | (uintptr_t)a.Y) & 15) { set_error("row_lse(tcgen05): X/Y must be 16-byte aligned"); return MCLIP_ERR_INVALID; }
  const bool pair = fwd_uses_pair(a.D);
  FwdPlan f = pair ? plan_fwd2(a.M, a.N, a.D) : plan_fwd(a.M, a.N, a.D);
  if (a.run_if) {   // chained fallback: normally exits at once, so keep the grid (= its launch cost) small
    f.nsplit = 1;
    f.tiles_per_split = f.tiles_total;
  }
  if (f.stages < 2) { set_error("row_lse(tcgen05): not enough shared memory for D=%lld", (long long)a.D); return MCLIP_ERR_UNSUPPORTED; }
  CUtensorMap tmX, tmY;
  int rc = make_tmap(&tmX, a.X, a.M, a.D, a.ldx, a.dtype, 128);
  if (rc) return rc;
  rc = make_tmap(&tmY, a.Y, a.N, a.D, a.ldy, a.dtype, pair ? 128u : (uint32_t)f.bn);
  if (rc) return rc;
  FwdParams p;
  p.M = a.M; p.N = a.N; p.kch = f.kch; p.stages = f.stages; p.tiles_total = f.tiles_total;
  p.tiles_per_split = f.tiles_per_split; p.diag_off = a.diag_off; p.ls = a.logit_scale;
  p.part_m2 = reinterpret_cast<float*>(a.ws);
  p.part_s = p.part_m2 + (size_t)f.nsplit * a.M;
  p.part_c = a.rowdot ? p.part_s + (size_t)f.nsplit * a.M : nullptr;
  p.diag = a.diag; p.bf16 = a.dtype == MCLIP_DTYPE_BF16;
  p.run_if = a.run_if;
  if (a.run_if && a.diag) { set_error("row_lse(tcgen05): a predicated call cannot write diag"); return MCLIP_ERR_INVALID; }
  p.dbg = options().dbg;
  if (a.diag) MCLIP_CUDA_OK(cudaMemsetAsync(a.diag, 0, sizeof(float) * a.M, a.stream));
  dim3 grid((unsigned)ceil_div(a.M, 128), (unsigned)f.nsplit);
  if (pair) {
    rc = set_smem(tc_row_lse2_kernel<true>, f.smem);
    if (rc) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * ceil_div(a.M, 256)), (unsigned)f.nsplit);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = f.smem;
    cfg.stream = a.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MCLIP_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_row_lse2_kernel<true>, tmX, tmY, p));
  } else if (f.xres) {
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
  return launch_lse_merge(p.part_m2, p.part_s, p.part_c, f.nsplit, a.M, a.lse, a.rowdot, a.stream, a.run_if);
}

namespace {
// ---- CTA-pair backward (D <= 512): preparation and launch, shared by mclip_block_grad and mclip_fused_grad ----
// column statistics of lse_y + the f16 copy of Y (bf16 inputs) into the caller's scratch
int bwd2_prepare(const BlockGradArgs& a, int64_t n_pad, float* stat_ws, void* y16_ws, Bwd2Prep* out) {
  Bwd2Prep r;
  r.ly2 = stat_ws;
  r.bcol = r.ly2 + n_pad;
  r.stepmm = r.bcol + n_pad;
  r.mu0 = r.stepmm + 2 * (n_pad / 256);
  r.has_col = a.w_col != 0.f;
  if (r.has_col) {
    prep_ly2_kernel<<<prep_blocks(n_pad), 1024, 0, a.stream>>>(a.lse_y, a.N, n_pad, log2f(a.w_col) + 12.f, r.ly2, r.bcol, r.stepmm, r.mu0);
    count_launch();
    MCLIP_CUDA_OK(cudaGetLastError());
  }
  r.y16 = a.Y;
  r.ld16 = a.ldy;
  if (a.dtype == MCLIP_DTYPE_BF16 && a.y16 != nullptr) {        // converted ahead of time by the caller (mclip_convert_f16)
    r.y16 = a.y16;
    r.ld16 = a.D;
  } else if (a.dtype == MCLIP_DTYPE_BF16) {
    __half* dst = reinterpret_cast<__half*>(y16_ws);
    const int64_t n8 = a.N * (a.D / 8);
    bf16_to_f16_kernel<<<ew_blocks(n8, 256), 256, 0, a.stream>>>(reinterpret_cast<const __nv_bfloat16*>(a.Y), a.N, a.D, a.ldy, dst);
    count_launch();
    MCLIP_CUDA_OK(cudaGetLastError());
    r.y16 = dst;
    r.ld16 = a.D;
  }
  *out = r;
  return MCLIP_OK;
}

// One launch of tc_block_grad2_kernel over rows [0, a.M) of a.X against all of Y.  `G` (may be null): tile-major f16
// scratch [ceil(a.M / 128) * 2][n_pad / 64][64][64] that receives every G tile (kStoreG).  With `force_partials` the
// accumulators always go to acc_ws (f32).
int bwd2_launch(const BlockGradArgs& a, const Bwd2Plan& b, const Bwd2Prep& pr, float* acc_ws, float* rd_ws, void* G,
                int64_t n_pad, bool force_partials) {
  const bool bf = a.dtype == MCLIP_DTYPE_BF16;
  CUtensorMap tmX, tmY, tmY16;
  int rc = make_tmap(&tmX, a.X, a.M, a.D, a.ldx, a.dtype, 64);
  if (rc) return rc;
  rc = make_tmap(&tmY, a.Y, a.N, a.D, a.ldy, a.dtype, 128);
  if (rc) return rc;
  rc = make_tmap(&tmY16, pr.y16, a.N, a.D, pr.ld16, MCLIP_DTYPE_F16, 128);
  if (rc) return rc;
  Bwd2Params p;
  p.M = a.M; p.N = a.N; p.D = a.D; p.kpairs = b.kpairs; p.ndh = b.ndh; p.steps_total = b.steps_total;
  p.steps_per_split = b.steps_per_split; p.nsplit = force_partials ? 2 : b.nsplit;   // > 1 only selects the partial-store epilogue
  p.diag_off = a.diag_off; p.ls = a.logit_scale;
  p.go = a.grad_out; p.lse_x = a.lse_x; p.ly2 = pr.has_col ? pr.ly2 : nullptr; p.bcol = pr.bcol; p.stepmm = pr.stepmm;
  p.mu0 = pr.mu0; p.w_row = a.w_row; p.w_diag = a.w_diag;
  p.inv_2n = a.inv_2n; p.has_col = pr.has_col ? 1 : 0; p.dX = a.dX; p.lddx = a.lddx;
  p.acc_ws = acc_ws; p.rd_ws = rd_ws; p.rowdot = a.rowdot;
  p.dbg = options().dbg;
  p.g_tiles_per_row = (int)(n_pad / 64);
  p.g_ptr = G;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * ceil_div(a.M, 128)), (unsigned)b.nsplit);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = b.smem;
  cfg.stream = a.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#define MCLIP_LAUNCH_BWD2(BF, SG, NX)                                                                         \
  do {                                                                                                        \
    rc = set_smem(tc_block_grad2_kernel<BF, SG, NX>, b.smem);                                                 \
    if (rc) return rc;                                                                                        \
    MCLIP_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_block_grad2_kernel<BF, SG, NX>, tmX, tmY, tmY16, p));           \
  } while (0)
  {
    ScopedKernelTimer timer(a.stream);   // no-op unless bench.py asked for in-situ timing; brackets only this launch
    if (a.D > 512) {                     // 512 < D <= 768: 12 X tiles, one S buffer (no G store: the dY GEMM stops at D = 512)
      if (G) { set_error("block_grad(tcgen05): the G store needs D <= 512"); return MCLIP_ERR_UNSUPPORTED; }
      if (bf) MCLIP_LAUNCH_BWD2(true, false, 12); else MCLIP_LAUNCH_BWD2(false, false, 12);
    } else if (G) { if (bf) MCLIP_LAUNCH_BWD2(true, true, 8); else MCLIP_LAUNCH_BWD2(false, true, 8); }
    else { if (bf) MCLIP_LAUNCH_BWD2(true, false, 8); else MCLIP_LAUNCH_BWD2(false, false, 8); }
  }
#undef MCLIP_LAUNCH_BWD2
  count_launch();
  MCLIP_CUDA_OK(cudaGetLastError());
  return MCLIP_OK;
}

template <typename T>
int launch_acc_to_dx(const float* acc_ws, const float* rd_ws, int nsplit, const BlockGradArgs& a) {
  const int64_t n = a.M * a.D;
  acc_to_dx_kernel<T><<<ew_blocks(n, 1024), 256, 0, a.stream>>>(acc_ws, rd_ws, nsplit, a.M, a.D, a.logit_scale, a.grad_out,
                                                                 a.inv_2n * (1.f / kGScale), reinterpret_cast<T*>(a.dX), a.lddx,
                                                                 a.rowdot);
  count_launch();
  MCLIP_CUDA_OK(cudaGetLastError());
  return MCLIP_OK;
}

int tc_block_grad2(const BlockGradArgs& a) {
  const Bwd2Plan b = plan_bwd2(a.M, a.N, a.D);
  const bool bf = a.dtype == MCLIP_DTYPE_BF16;
  const int64_t n_pad = (int64_t)b.steps_total * 256;
  const BwdWs w = bwd_ws_layout(b.nsplit, a.M, a.N, a.D, n_pad, bf);
  if (w.total > a.ws_bytes) { set_error("block_grad(tcgen05): workspace %zu < %zu", a.ws_bytes, w.total); return MCLIP_ERR_WORKSPACE; }
  uint8_t* ws = reinterpret_cast<uint8_t*>(a.ws);
  Bwd2Prep pr;
  int rc = bwd2_prepare(a, n_pad, reinterpret_cast<float*>(ws + w.ly2), ws + w.y16, &pr);
  if (rc) return rc;
  float* acc_ws = reinterpret_cast<float*>(ws + w.acc);
  float* rd_ws = reinterpret_cast<float*>(ws + w.rd);
  rc = bwd2_launch(a, b, pr, acc_ws, rd_ws, nullptr, 0, false);
  if (rc) return rc;
  if (b.nsplit > 1) return bf ? launch_acc_to_dx<__nv_bfloat16>(acc_ws, rd_ws, b.nsplit, a) : launch_acc_to_dx<__half>(acc_ws, rd_ws, b.nsplit, a);
  return MCLIP_OK;
}

// =================================================================================================
// shared-recompute backward: ONE recompute of S per logits tile feeds both gradients (4 GEMM units per step, not 5)
// =================================================================================================
// For row panels of X (sized so that one launch of the pair kernel fills every SM pair exactly once):
//   main stream:  tc_block_grad2_kernel<.., kStoreG>   dX[panel] partials + G[panel, :] (f16 * 2^12) -> global scratch
//                 acc_to_dx_dot_kernel                 dX[panel] (input dtype) and xdot[i] = <X_i, dX_i acc> (for d logit_scale)
//   side stream:  tc_gemm_tn_kernel                    dY_acc += G[panel, :]^T X16[panel]            (overlaps the next panel)
// and finally dY = alpha * dY_acc.  Only one [panel x N] strip of G (two buffers) ever exists; the N x N matrix does not.
// d(logit_scale) comes from Euler's identity: the loss depends on X only through logit_scale * X Y^T, hence
// sum_i <X_i, dL/dX_i> = logit_scale * dL/d(logit_scale): t = sum_ij G_ij C_ij = sum_i <X_i, (G Y)_i>.
struct FusedPlan { int64_t panel_rows; int panels; Bwd2Plan b; int64_t n_pad; };

FusedPlan plan_fused(int64_t M, int64_t N, int64_t D) {
  FusedPlan f;
  const int slots = pair_slots();
  const int steps_total = (int)ceil_div(N, 256);
  // column splits x row blocks per panel: a panel launch should fill the pair slots once; pick the split count that
  // minimises panels x (steps per CTA + ~3 step-times of per-CTA prologue and accumulator drain)
  const int64_t rbs = ceil_div(M, 128);
  int nsplit = 1;
  double best = 1e30;
  for (int s = 1; s <= 8 && s <= slots && s <= steps_total; ++s) {
    const int64_t rb_s = (slots / s) < rbs ? (slots / s) : rbs;
    const double cost = (double)ceil_div(rbs, rb_s) * ((double)ceil_div(steps_total, s) + 3.0);
    if (cost < best - 1e-9) { best = cost; nsplit = s; }
  }
  int64_t rb = slots / nsplit;                       // row blocks per panel
  if (rb > rbs) rb = rbs;
  f.panels = (int)ceil_div(rbs, rb);
  rb = ceil_div(rbs, f.panels);                      // equalise the panels
  f.panel_rows = rb * 128;
  f.b = plan_bwd2(f.panel_rows, N, D);
  f.b.nsplit = nsplit;
  f.b.steps_per_split = (int)ceil_div(steps_total, nsplit);
  f.b.nsplit = (int)ceil_div(steps_total, f.b.steps_per_split);
  f.n_pad = (int64_t)steps_total * 256;
  return f;
}

// G strips / partial-sum buffers in flight.  Two suffice: a third one (so that the recompute launch of panel p+2 need not wait
// for the dY GEMM of panel p and can fill the SMs that GEMM's second round leaves idle) was measured and changes nothing
// (2.64 vs 2.65 ms per call at C3) -- the call is bound by total tensor / L2-port work, not by launch-level gaps.
constexpr int kFusedBufs = 2;
struct FusedWs { size_t stat, y16, x16, acc_y, acc_x, g, total; };
FusedWs fused_ws_layout(const FusedPlan& f, int64_t M, int64_t N, int64_t D, bool bf) {
  FusedWs w;
  size_t off = 0;
  w.stat = off;  off += align_up(((size_t)2 * f.n_pad + 2 * (f.n_pad / 256) + 64) * sizeof(float), 256);
  w.y16 = off;   off += bf ? align_up((size_t)N * D * 2, 256) : 0;
  w.x16 = off;   off += bf ? align_up((size_t)M * D * 2, 256) : 0;
  w.acc_y = off; off += align_up((size_t)N * D * sizeof(float), 256);
  w.acc_x = off; off += kFusedBufs * align_up((size_t)f.b.nsplit * f.panel_rows * D * sizeof(float), 256);
  w.g = off;     off += kFusedBufs * align_up((size_t)f.panel_rows * f.n_pad * 2, 1024);
  w.total = off;
  return w;
}

// per-device side stream + fork/join events (created once, never destroyed: they live as long as the library)
struct FusedStreams { cudaStream_t side = nullptr; std::vector<cudaEvent_t> ev; };
int fused_streams(int nev, FusedStreams** out) {
  static std::mutex mu;
  static std::unordered_map<int, FusedStreams> cache;
  int dev = 0;
  MCLIP_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  FusedStreams& fs = cache[dev];
  if (!fs.side) MCLIP_CUDA_OK(cudaStreamCreateWithFlags(&fs.side, cudaStreamNonBlocking));
  while ((int)fs.ev.size() < nev) {
    cudaEvent_t e;
    MCLIP_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    fs.ev.push_back(e);
  }
  *out = &fs;
  return MCLIP_OK;
}

// sum of the column-split partials -> dX (input dtype) and xdot[row] = <X[row], sum> / 2^12; one warp per row, 8 elements
// per lane and iteration (two float4 per partial, one 16-byte load of X, one 16-byte store of dX)
template <typename T>
__global__ void __launch_bounds__(256)
acc_to_dx_dot_kernel(const float* __restrict__ acc, int nsplit, int64_t M, int64_t D, const T* __restrict__ X, int64_t ldx,
                     const float* __restrict__ ls, const float* __restrict__ go, float scale, T* __restrict__ dX, int64_t lddx,
                     float* __restrict__ xdot) {
  const float alpha = (go ? go[0] : 1.f) * ls[0] * scale;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n = M * D;
  for (int64_t r = warp; r < M; r += nwarps) {
    float dot = 0.f;
    for (int64_t d = lane * 8; d < D; d += 256) {     // D % 8 == 0 and 16-byte aligned rows on this path
      const int64_t i = r * D + d;
      float4 a0 = *reinterpret_cast<const float4*>(acc + i), a1 = *reinterpret_cast<const float4*>(acc + i + 4);
      for (int s = 1; s < nsplit; ++s) {
        const float4 b0 = *reinterpret_cast<const float4*>(acc + (int64_t)s * n + i);
        const float4 b1 = *reinterpret_cast<const float4*>(acc + (int64_t)s * n + i + 4);
        a0.x += b0.x; a0.y += b0.y; a0.z += b0.z; a0.w += b0.w;
        a1.x += b1.x; a1.y += b1.y; a1.z += b1.z; a1.w += b1.w;
      }
      const float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const uint4 xin = *reinterpret_cast<const uint4*>(X + r * ldx + d);
      const T* xe = reinterpret_cast<const T*>(&xin);
      uint4 out;
      T* oe = reinterpret_cast<T*>(&out);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        dot = fmaf(v[e], to_f32<T>(xe[e]), dot);
        oe[e] = from_f32<T>(v[e] * alpha);
      }
      *reinterpret_cast<uint4*>(dX + r * lddx + d) = out;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (lane == 0) xdot[r] = dot * (1.f / kGScale);
  }
}

template <typename T>
int fused_grad_impl(const FusedGradArgs& a) {
  const bool bf = a.dtype == MCLIP_DTYPE_BF16;
  const FusedPlan f = plan_fused(a.M, a.N, a.D);
  const FusedWs w = fused_ws_layout(f, a.M, a.N, a.D, bf);
  if (w.total > a.ws_bytes) { set_error("fused_grad: workspace %zu < %zu", a.ws_bytes, w.total); return MCLIP_ERR_WORKSPACE; }
  uint8_t* ws = reinterpret_cast<uint8_t*>(a.ws);
  // bench.py's in-situ bracket covers the whole call here (its two kernels overlap on two streams; the side stream has
  // been joined back into a.stream when the closing event is recorded); the per-launch brackets inside are suppressed
  ScopedKernelTimer timer(a.stream);
  TimerSuppress no_inner;
  BlockGradArgs g;
  g.X = a.X; g.Y = a.Y; g.M = a.M; g.N = a.N; g.D = a.D; g.ldx = a.ldx; g.ldy = a.ldy; g.dtype = a.dtype;
  g.logit_scale = a.logit_scale; g.grad_out = a.grad_out; g.lse_x = a.lse_x; g.lse_y = a.lse_y; g.diag_off = a.diag_off;
  g.w_row = 1.f; g.w_col = 1.f; g.w_diag = 2.f; g.inv_2n = a.inv_2n; g.dX = a.dX; g.lddx = a.lddx; g.rowdot = nullptr;
  g.ws = nullptr; g.ws_bytes = 0; g.stream = a.stream;
  FusedStreams* fs = nullptr;
  int rc = fused_streams(2 * f.panels + 2, &fs);
  if (rc) return rc;
  // fork: everything only the dY product needs (the f16 copy of X, the partial sums of each panel, the GEMM itself) runs on
  // the side stream, so that the main stream is nothing but prep + back-to-back recompute launches
  cudaEvent_t ev_fork = fs->ev[2 * f.panels];
  MCLIP_CUDA_OK(cudaEventRecord(ev_fork, a.stream));
  MCLIP_CUDA_OK(cudaStreamWaitEvent(fs->side, ev_fork, 0));
  const void* x16 = a.X;
  int64_t ldx16 = a.ldx;
  if (bf) {
    __half* dst = reinterpret_cast<__half*>(ws + w.x16);
    const int64_t n8 = a.M * (a.D / 8);
    bf16_to_f16_kernel<<<ew_blocks(n8, 256), 256, 0, fs->side>>>(reinterpret_cast<const __nv_bfloat16*>(a.X), a.M, a.D, a.ldx, dst);
    count_launch();
    MCLIP_CUDA_OK(cudaGetLastError());
    x16 = dst;
    ldx16 = a.D;
  }
  Bwd2Prep pr;
  rc = bwd2_prepare(g, f.n_pad, reinterpret_cast<float*>(ws + w.stat), ws + w.y16, &pr);
  if (rc) return rc;
  float* acc_y = reinterpret_cast<float*>(ws + w.acc_y);
  const size_t accx_stride = align_up((size_t)f.b.nsplit * f.panel_rows * a.D * sizeof(float), 256);
  const size_t g_stride = align_up((size_t)f.panel_rows * f.n_pad * 2, 1024);
  const int slots = pair_slots();
  for (int pi = 0; pi < f.panels; ++pi) {
    const int64_t r0 = (int64_t)pi * f.panel_rows;
    const int64_t rows = (a.M - r0 < f.panel_rows) ? a.M - r0 : f.panel_rows;
    void* gbuf = ws + w.g + (size_t)(pi % kFusedBufs) * g_stride;
    float* acc_x = reinterpret_cast<float*>(ws + w.acc_x + (size_t)(pi % kFusedBufs) * accx_stride);
    if (pi >= kFusedBufs)       // G and partial buffers of panel pi - kFusedBufs are free again once its dY GEMM is done
      MCLIP_CUDA_OK(cudaStreamWaitEvent(a.stream, fs->ev[2 * (pi - kFusedBufs) + 1], 0));
    BlockGradArgs gp = g;
    gp.X = reinterpret_cast<const T*>(a.X) + r0 * a.ldx;
    gp.M = rows;
    gp.lse_x = a.lse_x + r0;
    gp.diag_off = a.diag_off + r0;
    gp.dX = reinterpret_cast<T*>(a.dX) + r0 * a.lddx;
    rc = bwd2_launch(gp, f.b, pr, acc_x, nullptr, gbuf, f.n_pad, true);
    if (rc) return rc;
    MCLIP_CUDA_OK(cudaEventRecord(fs->ev[2 * pi], a.stream));
    MCLIP_CUDA_OK(cudaStreamWaitEvent(fs->side, fs->ev[2 * pi], 0));
    acc_to_dx_dot_kernel<T><<<ew_blocks(rows * 32, 256), 256, 0, fs->side>>>(
        acc_x, f.b.nsplit, rows, a.D, reinterpret_cast<const T*>(gp.X), a.ldx, a.logit_scale, a.grad_out,
        a.inv_2n * (1.f / kGScale), reinterpret_cast<T*>(gp.dX), a.lddx, a.xdot + r0);
    count_launch();
    MCLIP_CUDA_OK(cudaGetLastError());
    rc = launch_gemm_tn(gbuf, f.n_pad / 64, reinterpret_cast<const __half*>(x16) + r0 * ldx16, ldx16, acc_y, a.D, rows, a.N, a.D,
                        slots, options().dbg, pi == 0, fs->side);
    if (rc) return rc;
    MCLIP_CUDA_OK(cudaEventRecord(fs->ev[2 * pi + 1], fs->side));
  }
  MCLIP_CUDA_OK(cudaStreamWaitEvent(a.stream, fs->ev[2 * (f.panels - 1) + 1], 0));
  {
    const int64_t n = a.N * a.D;
    acc_to_dx_kernel<T><<<ew_blocks(n, 1024), 256, 0, a.stream>>>(acc_y, nullptr, 1, a.N, a.D, a.logit_scale, a.grad_out,
                                                                   a.inv_2n * (1.f / kGScale), reinterpret_cast<T*>(a.dY), a.lddy,
                                                                   nullptr);
    count_launch();
    MCLIP_CUDA_OK(cudaGetLastError());
  }
  return MCLIP_OK;
}

}  // namespace

int tc_bwd2_prepare(const BlockGradArgs& a, int64_t n_pad, float* stat_ws, void* y16_ws, Bwd2Prep* out) {
  return bwd2_prepare(a, n_pad, stat_ws, y16_ws, out);
}
int tc_dbg_flags() { return options().dbg; }
int tc_pair_slots() { return pair_slots(); }

int launch_convert_f16(const void* src, int64_t rows, int64_t D, int64_t ld, void* dst, cudaStream_t stream) {
  const int64_t n8 = rows * (D / 8);
  bf16_to_f16_kernel<<<ew_blocks(n8, 256), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(src), rows, D, ld,
                                                             reinterpret_cast<__half*>(dst));
  count_launch();
  MCLIP_CUDA_OK(cudaGetLastError());
  return MCLIP_OK;
}

bool tc_fused_supported(int64_t M, int64_t N, int64_t D, int64_t ldx, int64_t ldy, int dtype) {
  if (!options().fused_bwd) return false;
  if (!tc_supported(M, N, D, ldx, ldy, dtype, MCLIP_OP_BLOCK_GRAD) || D > 512) return false;
  // pays off once the panel launches are long enough to hide their prologue / drain and the strip round trip; measured
  // (tools/one_fused.py, square problems): B = 8192: 0.243 ms vs 0.221 ms for two recompute launches (loses), B = 16384:
  // 0.694 vs 0.868 ms, B = 32768: 2.65 vs 2.88 ms
  return N >= 16384 && M >= 4096;
}

size_t tc_fused_grad_ws(int64_t M, int64_t N, int64_t D) {
  return fused_ws_layout(plan_fused(M, N, D), M, N, D, true).total;
}

int tc_fused_grad(const FusedGradArgs& a) {
  if (((uintptr_t)a.X | (uintptr_t)a.Y | (uintptr_t)a.dX | (uintptr_t)a.dY) & 15) { set_error("fused_grad: X/Y/dX/dY must be 16-byte aligned"); return MCLIP_ERR_INVALID; }
  if (a.D > 512) { set_error("fused_grad: D=%lld > 512", (long long)a.D); return MCLIP_ERR_UNSUPPORTED; }
  return a.dtype == MCLIP_DTYPE_BF16 ? fused_grad_impl<__nv_bfloat16>(a) : fused_grad_impl<__half>(a);
}

int tc_block_grad(const BlockGradArgs& a) {
  if (((uintptr_t)a.X | (uintptr_t)a.Y | (uintptr_t)a.dX) & 15) { set_error("block_grad(tcgen05): X/Y/dX must be 16-byte aligned"); return MCLIP_ERR_INVALID; }
  if (!(a.w_row > 0.f) || a.w_col < 0.f) { set_error("block_grad(tcgen05): needs w_row > 0 and w_col >= 0"); return MCLIP_ERR_INVALID; }
  // D <= 512: both S buffers in TMEM; 512 < D <= 768: one S buffer, X tiles 8..11 resident too (kNX = 12)
  return (a.D <= 512 && use_persistent_bwd(a.M, a.N, a.D)) ? tc_block_grad2p(a) : tc_block_grad2(a);
}

}  // namespace mclip
