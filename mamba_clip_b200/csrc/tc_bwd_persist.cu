// Persistent CTA-pair backward kernel (option "bwd_persist"; D <= 512): the stream-K style variant of
// tc_block_grad2_kernel (tc_kernels.cu).  Not the default: the measurements are quoted at use_persistent_bwd() in
// tc_kernels.cu and in DESIGN.md section 3.2.  It passes the same parity tests (tests/test_gpu_kernels.py).
#include <cuda.h>

#include <cstdio>
#include <mutex>
#include <unordered_map>

#include "tc_bwd_common.cuh"

namespace mclip {

namespace {

// =================================================================================================
// backward, persistent CTA-pair version (D <= 512): opt-in (MCLIP_BWD_PERSIST=1), see use_persistent_bwd()
// =================================================================================================
// Same per-step pipeline as tc_block_grad2_kernel (S two steps ahead of dX, G through shared memory, one exponential
// per element), but a CTA pair no longer owns one (row block, column split): the launch has one pair per pair slot
// of the GPU and pair p walks the contiguous range [floor(p U / P), floor((p+1) U / P)) of the U = row_blocks x steps
// step units, crossing row-block boundaries on the way.  That removes the wave quantisation of the split grid (at
// W = 8 a rank has 32 row blocks for 74 pair slots), most of the per-CTA prologue / drain bubbles, and all but the
// boundary partial sums: a row block that falls entirely inside one pair's range is written straight to dX; only
// the row blocks cut by a range boundary go through f32 partials (fixed piece order: deterministic).
struct Bwd2PParams {
  int64_t M, N, D;
  int kpairs, ndh;
  int steps_total;      // S = ceil(N / 256)
  int64_t units;        // U = ceil(M / 128) * S
  int npairs;           // P (grid = 2 P CTAs)
  int64_t diag_off;
  const float* ls;
  const float* go;
  const float* lse_x;
  const float* ly2;
  const float* bcol;
  const float* stepmm;
  const float* mu0;
  float w_row, w_diag, inv_2n;
  int has_col;
  void* dX;
  int64_t lddx;
  float* acc_ws;        // [pair][2][128][D] f32 partial dX (unscaled): slot 0 = the pair's first item, 1 = any later one
  float* rd_ws;         // [pair][2][128]    f32 partial rowdot
  float* rowdot;
  int dbg;              // MCLIP_DBG & 16 (profile builds only): print barrier-wait cycle counts of a few pairs
};

__host__ __device__ __forceinline__ int64_t unit_lo(int64_t p, int64_t U, int64_t P) { return p * U / P; }
// first / last pair whose unit range intersects row block rb (units [rb S, (rb+1) S))
__host__ __device__ __forceinline__ int pair_of_unit(int64_t u, int64_t U, int64_t P) {
  int64_t p = u * P / U;
  while (p + 1 < P && unit_lo(p + 1, U, P) <= u) ++p;
  while (p > 0 && unit_lo(p, U, P) > u) --p;
  return (int)p;
}

template <bool kBF16>
__global__ void __launch_bounds__(kThreads, 1)
tc_block_grad2p_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                       const __grid_constant__ CUtensorMap tmY16, const Bwd2PParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = align1024(smem_u32(smem_raw));
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t x_base = smem_base;                              // [8][64 rows][64 k]   64 KB
  const uint32_t g_base = x_base + 8 * kTile8K;                   // [4][64 rows][64 y]    32 KB (single buffer)
  const uint32_t ring_base = g_base + 4 * kTile8K;                // [kRing2][2][128][64]  128 KB
  const uint32_t misc_base = ring_base + kRing2 * kStage2;
  uint8_t* misc_gen = smem_gen + (misc_base - smem_base);
  const uint32_t bar_base = misc_base;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };             // leader: both CTAs' TMA bytes
  auto empty_bar = [&](int s) { return bar_base + 8u * (4 + s); };      // per CTA, MMA commit multicast
  const uint32_t xfull_bar = bar_base + 8u * 8;                          // leader
  auto sfull_bar = [&](int b) { return bar_base + 8u * (9 + b); };      // per CTA, multicast
  const uint32_t gfull_bar = bar_base + 8u * 11;                         // leader: 16 epilogue warps
  const uint32_t xempty_bar = bar_base + 8u * 12;                        // per CTA, multicast: S MMAs of an item done
  const uint32_t gempty_bar = bar_base + 8u * 13;                        // per CTA, multicast
  const uint32_t dxempty_bar = bar_base + 8u * 14;                       // leader: 16 epilogue warps drained dX
  const uint32_t dxfull_bar = bar_base + 8u * 15;                        // per CTA, multicast
  auto sread_bar = [&](int b) { return bar_base + 8u * (17 + b); };     // leader: 16 epilogue warps have loaded S(b)
  const uint32_t tmem_slot = bar_base + 8u * 19;
  uint32_t* tmem_slot_gen = reinterpret_cast<uint32_t*>(misc_gen + 8u * 19);
  float* rd_scratch = reinterpret_cast<float*>(misc_gen + 256);          // [4][64]
  float* range_scratch = reinterpret_cast<float*>(misc_gen + 1280);      // [8][2]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int64_t pair = blockIdx.x >> 1;
  const int S = p.steps_total;
  const int64_t u0 = unit_lo(pair, p.units, p.npairs), u1 = unit_lo(pair + 1, p.units, p.npairs);
  const int G = (int)(u1 - u0);                      // steps this pair executes
  constexpr uint32_t kTmemCols = 512;
  constexpr uint32_t kDxCol = 256;
  // step g of this pair is unit u0 + g = (row block (u0 + g) / S, y step (u0 + g) % S); every role walks its own
  // cursors (no divisions inside the loops: the epilogue has no slack for them)
  const int rb0 = (int)(u0 / S), ys0 = (int)(u0 - (int64_t)rb0 * S);
  struct Cursor {
    int g, rb, ys;
    __device__ __forceinline__ void next(int S_) { ++g; if (++ys == S_) { ys = 0; ++rb; } }
  };
  auto first_of_item = [&](const Cursor& c) { return c.g == 0 || c.ys == 0; };
  auto last_of_item = [&](const Cursor& c) { return c.g == G - 1 || c.ys == S - 1; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
    tma_prefetch_desc(&tmY16);
    for (int s = 0; s < kRing2; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
    mbar_init(xfull_bar, 2);
    mbar_init(xempty_bar, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(sfull_bar(b), 1); mbar_init(sread_bar(b), 2 * (kEpiThreads / 32)); }
    mbar_init(gfull_bar, 2 * (kEpiThreads / 32));
    mbar_init(gempty_bar, 1);
    mbar_init(dxfull_bar, 1);
    mbar_init(dxempty_bar, 2 * (kEpiThreads / 32));
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
    if (lane == 0 && G > 0) {
      // ---------------- TMA producer (both CTAs; bytes are credited to the leader's barriers) ----------------
      uint32_t it = 0, items = 0;
      auto stage_begin = [&]() -> uint32_t {
        const int s = it % kRing2;
        const uint32_t ph = (it / kRing2) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        if (leader) mbar_expect_tx(full_bar(s), 2 * kStage2); else mbar_arrive_cluster(full_bar(s), 0);
        ++it;
        return (uint32_t)s;
      };
      Cursor cs{0, rb0, ys0}, cd{0, rb0, ys0};
      auto load_s = [&]() {   // Y tiles of step cs.g as the N operand of S: this CTA's 128 rows, all of D
        if (first_of_item(cs)) {
          // new row block: the resident X tiles are replaced once the S MMAs of the previous item have read them
          if (items > 0) mbar_wait(xempty_bar, (items - 1) & 1);
          ++items;
          const int32_t m0 = cs.rb * 128 + 64 * (int32_t)rank;
          if (leader) mbar_expect_tx(xfull_bar, 2 * 8 * kTile8K); else mbar_arrive_cluster(xfull_bar, 0);
          for (int c = 0; c < 8; ++c) tma_load_2d_cg2(x_base + c * kTile8K, &tmX, c * 64, m0, xfull_bar);
        }
        const int32_t y0 = cs.ys * 256 + 128 * (int32_t)rank;
        for (int i = 0; i < p.kpairs; ++i) {
          const uint32_t s = stage_begin();
          const uint32_t dst = ring_base + s * kStage2;
          tma_load_2d_cg2(dst, &tmY, (2 * i) * 64, y0, full_bar(s));
          tma_load_2d_cg2(dst + kChunkBytes, &tmY, (2 * i + 1) * 64, y0, full_bar(s));
        }
        cs.next(S);
      };
      auto load_dx = [&]() {  // Y16 tiles of step cd.g as the [K = y][N = d] operand of dX: all 256 rows, this CTA's d
        for (int yh = 0; yh < 2; ++yh) {
          const int32_t y0 = cd.ys * 256 + 128 * yh;
          for (int h = 0; h < p.ndh; ++h) {
            const uint32_t s = stage_begin();
            const uint32_t dst = ring_base + s * kStage2;
            const int32_t dcol = (4 * h + 2 * (int32_t)rank) * 64;
            tma_load_2d_cg2(dst, &tmY16, dcol, y0, full_bar(s));
            tma_load_2d_cg2(dst + kChunkBytes, &tmY16, dcol + 64, y0, full_bar(s));
          }
        }
        cd.next(S);
      };
      load_s();
      if (G > 1) load_s();
      for (int g = 0; g < G; ++g) {
        if (g + 2 < G) load_s();
        load_dx();
      }
    }
  } else if (warp == 1) {
    if (leader && G > 0) {
      // ---------------- MMA issuer (leader CTA only): the whole warp waits, one elected lane issues ----------------
      const bool elected = elect_one();
      const uint32_t idesc_s = make_idesc_f16(kBF16, kBF16, 128, 256, false, false);
      const uint32_t idesc_dx = make_idesc_f16(false, false, 128, 256, false, true);
      uint32_t it = 0, s_items = 0, dx_items = 0;
      const bool prof = kProfile && (p.dbg & 16) != 0 && elected;
      long long t_full = 0, t_gfull = 0, t_xfull = 0, t_dxempty = 0, t_sread = 0, t_begin = clock64();
      auto stage_wait = [&]() -> uint32_t {
        const int s = it % kRing2;
        const uint32_t ph = (it / kRing2) & 1;
        const long long t0 = prof ? clock64() : 0;
        mbar_wait(full_bar(s), ph);
        if (prof) t_full += clock64() - t0;
        tc_fence_after();
        ++it;
        return (uint32_t)s;
      };
      Cursor cs{0, rb0, ys0}, cd{0, rb0, ys0};
      auto issue_s = [&]() {
        const int buf = cs.g & 1;
        if (first_of_item(cs)) {
          const long long t0 = prof ? clock64() : 0;
          mbar_wait(xfull_bar, s_items & 1);
          if (prof) t_xfull += clock64() - t0;
          tc_fence_after();
          ++s_items;
        }
        const bool last = last_of_item(cs);
        cs.next(S);
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
                mma_ss_cg2(d_tmem, ad, bd, idesc_s, (i | e | k) != 0);
              }
            }
            mma_commit_cg2(empty_bar(s), 3);
            if (i == p.kpairs - 1) {
              mma_commit_cg2(sfull_bar(buf), 3);
              if (last) mma_commit_cg2(xempty_bar, 3);
            }
          }
          __syncwarp();
        }
      };
      auto issue_dx = [&]() {
        {
          const long long t0 = prof ? clock64() : 0;
          mbar_wait(gfull_bar, cd.g & 1);
          if (prof) t_gfull += clock64() - t0;
        }
        tc_fence_after();
        const bool first = first_of_item(cd), last = last_of_item(cd);
        cd.next(S);
        if (first) {
          // the accumulator of the previous item must have been drained before accumulate = 0 overwrites it
          if (dx_items > 0) {
            const long long t0 = prof ? clock64() : 0;
            mbar_wait(dxempty_bar, (dx_items - 1) & 1);
            if (prof) t_dxempty += clock64() - t0;
            tc_fence_after();
          }
          ++dx_items;
        }
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
                mma_ss_cg2(tmem_base + kDxCol + h * 128, ad, bd, idesc_dx, !(first && yh == 0 && kk == 0));
              }
              mma_commit_cg2(empty_bar(s), 3);
              if (yh == 1 && h == p.ndh - 1) {
                mma_commit_cg2(gempty_bar, 3);
                if (last) mma_commit_cg2(dxfull_bar, 3);
              }
            }
            __syncwarp();
          }
        }
      };
      issue_s();
      if (G > 1) issue_s();
      for (int g = 0; g < G; ++g) {
        if (g + 2 < G) {
          const long long t0 = prof ? clock64() : 0;
          mbar_wait(sread_bar(g & 1), (g >> 1) & 1);   // S(g) is in registers: its TMEM buffer may be overwritten
          if (prof) t_sread += clock64() - t0;
          tc_fence_after();
          issue_s();
        }
        issue_dx();
      }
      if (prof && (pair < 3 || pair == p.npairs - 1))
        printf("[bwd2p mma pair %d] steps=%d items=%u total=%lld clk (per step %lld)  wait: full %lld gfull %lld sread %lld xfull %lld dxempty %lld\n",
               (int)pair, G, dx_items, clock64() - t_begin, (clock64() - t_begin) / max(G, 1), t_full, t_gfull, t_sread, t_xfull, t_dxempty);
    }
  } else if (warp >= kEpiWarp0) {
    // ---------------- epilogue: S -> G (f16 * 2^12) into shared memory; dX out at the end of every item ----------------
    const int ew = warp - kEpiWarp0;
    const int q = warp & 3;                   // TMEM lane quadrant
    const int half = ew >> 2;                 // which 64 of this lane-half's 128 columns
    const int r = (q & 1) * 32 + lane;        // row within this CTA's 64
    const int cS = (q >> 1) * 128 + half * 64;   // first S column (of 256) this thread handles
    const int kc = cS >> 6;                   // K-chunk of G it fills
    const float ls = p.ls[0];
    const float k2 = ls * kLog2e;
    const bool has_col = p.has_col != 0;
    const float w_diag_s = p.w_diag * kGScale;
    const float mu0 = has_col ? p.mu0[0] : 0.f;
    const float alpha = (p.go ? p.go[0] : 1.f) * ls * p.inv_2n * (1.f / kGScale);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t g_row = g_base + kc * kTile8K + r * 128;
    // per-item state
    int64_t row = 0, jd = 0, blk_lo = 0;
    int item_first_ys = 0, item_first_g = 0, items = 0;
    Cursor cur{0, rb0, ys0};
    float lx2 = 0.f, a_i = 0.f, lx_min = 0.f, lx_max = 0.f, rd = 0.f;
    bool a_ok = false;
    const bool eprof = kProfile && (p.dbg & 16) != 0 && warp == kEpiWarp0 && lane == 0;
    long long e_sfull = 0, e_gempty = 0, e_dxfull = 0, e_drain = 0, e_setup = 0, e_begin = clock64();
    for (int g = 0; g < G; ++g, cur.next(S)) {
      const int ys = cur.ys;
      if (first_of_item(cur)) {
        const long long ts0 = eprof ? clock64() : 0;
        row = (int64_t)cur.rb * 128 + 64 * rank + r;
        jd = row + p.diag_off;
        blk_lo = (int64_t)cur.rb * 128 + p.diag_off;
        item_first_ys = ys;
        item_first_g = g;
        rd = 0.f;
        // rows past M reuse the last valid row's LSE so that they do not widen the range check below
        lx2 = p.lse_x[row < p.M ? row : p.M - 1] * kLog2e - (log2f(p.w_row) + 12.f);
        if (has_col) {
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
        if (eprof) e_setup += clock64() - ts0;
      }
      const int buf = g & 1;
      const uint32_t bph = (g >> 1) & 1;
      const int64_t n0 = (int64_t)ys * 256;
      {
        const long long t0 = eprof ? clock64() : 0;
        mbar_wait(sfull_bar(buf), bph);
        if (eprof) e_sfull += clock64() - t0;
      }
      tc_fence_after();
      const bool special = (n0 + 256 > p.N) || (n0 < blk_lo + 128 && n0 + 256 > blk_lo);
      bool fast = false;
      if (has_col && a_ok) {
        const float2 mm = __ldg(reinterpret_cast<const float2*>(p.stepmm) + ys);   // (min, max) of ly2 in this step
        fast = (lx_max - mm.x <= 100.f) && (mm.y - lx_min <= 100.f) &&
               (!(mm.x <= mm.y) || (fabsf(mm.x - mu0) <= 120.f && fabsf(mm.y - mu0) <= 120.f));
      }
      uint32_t gq[2][16];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t v[32];
        tmem_ld32(lane_addr + buf * 128 + half * 64 + cc * 32, v);
        tmem_ld_wait();
        if (cc == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(sread_bar(buf)); else mbar_arrive_cluster(sread_bar(buf), 0);
          }
        }
        const int64_t col0 = n0 + cS + cc * 32;
        const float* ly2 = has_col ? p.ly2 + col0 : nullptr;
        if (fast) {
          if (special) bwd2_chunk_fast<true>(v, gq[cc], k2, lx2, a_i, p.bcol + col0, w_diag_s, col0, p.N, jd, rd);
          else bwd2_chunk_fast<false>(v, gq[cc], k2, lx2, a_i, p.bcol + col0, w_diag_s, col0, p.N, jd, rd);
        } else if (special) {
          if (has_col) bwd2_chunk<true, true>(v, gq[cc], k2, lx2, ly2, w_diag_s, col0, p.N, jd, rd);
          else bwd2_chunk<true, false>(v, gq[cc], k2, lx2, ly2, w_diag_s, col0, p.N, jd, rd);
        } else {
          if (has_col) bwd2_chunk<false, true>(v, gq[cc], k2, lx2, ly2, w_diag_s, col0, p.N, jd, rd);
          else bwd2_chunk<false, false>(v, gq[cc], k2, lx2, ly2, w_diag_s, col0, p.N, jd, rd);
        }
      }
      tc_fence_before();          // TMEM reads of S are complete
      // G is single-buffered: the dX MMAs of the previous step must have finished reading it.  The values are
      // already in registers, so this wait overlaps with the S MMAs of the next step on the tensor pipe.
      {
        const long long t0 = eprof ? clock64() : 0;
        mbar_wait(gempty_bar, (g & 1) ^ 1);
        if (eprof) e_gempty += clock64() - t0;
      }
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) {
          const uint32_t piece = (uint32_t)((cc * 4 + pc) ^ (r & 7));
          st_shared_v4(g_row + piece * 16, gq[cc][4 * pc], gq[cc][4 * pc + 1], gq[cc][4 * pc + 2], gq[cc][4 * pc + 3]);
        }
      }
      fence_proxy_async_smem();   // G visible to the tensor-core (async) proxy
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(gfull_bar); else mbar_arrive_cluster(gfull_bar, 0);
      }

      if (last_of_item(cur)) {
        // ---- end of the item: dX accumulator -> global (directly, or as the f32 partial of a cut row block) ----
        const bool whole = item_first_ys == 0 && ys == S - 1;
        const int64_t slot = 2 * pair + (item_first_g == 0 ? 0 : 1);
        const long long td0 = eprof ? clock64() : 0;
        mbar_wait(dxfull_bar, items & 1);
        if (eprof) e_dxfull += clock64() - td0;
        ++items;
        tc_fence_after();
        for (int h = 0; h < p.ndh; ++h) {
#pragma unroll 1
          for (int cc = 0; cc < 2; ++cc) {
            uint32_t v[32];
            tmem_ld32(lane_addr + kDxCol + h * 128 + half * 64 + cc * 32, v);
            tmem_ld_wait();
            const int64_t d0 = (int64_t)h * 256 + (q >> 1) * 128 + half * 64 + cc * 32;
            if (row < p.M && d0 < p.D) {
              if (!whole) {
                float* dst = p.acc_ws + ((slot * 128 + 64 * rank + r) * p.D + d0);
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
        if (eprof) e_drain += clock64() - td0;
        // the accumulator is in registers / memory: the next item may overwrite it
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(dxempty_bar); else mbar_arrive_cluster(dxempty_bar, 0);
        }
        // rowdot: four threads share a row (2 lane halves x 2 column halves); add in fixed order via shared memory
        if (p.rowdot != nullptr) {
          rd_scratch[((q >> 1) * 2 + half) * 64 + r] = rd;
          named_bar_sync(2, kEpiThreads);
          if ((q >> 1) == 0 && half == 0 && row < p.M) {
            const float tot = ((rd_scratch[r] + rd_scratch[64 + r]) + (rd_scratch[128 + r] + rd_scratch[192 + r])) * (1.f / kGScale);
            if (!whole) p.rd_ws[slot * 128 + 64 * rank + r] = tot;
            else p.rowdot[row] = tot;
          }
          named_bar_sync(2, kEpiThreads);   // rd_scratch may be rewritten by the next item
        }
      }
    }
    if (eprof && (pair < 3 || pair == p.npairs - 1))
      printf("[bwd2p epi pair %d cta %u] steps=%d items=%d total=%lld clk  wait: sfull %lld gempty %lld dxfull %lld  drain(incl. dxfull) %lld  item setup %lld\n",
             (int)pair, rank, G, items, clock64() - e_begin, e_sfull, e_gempty, e_dxfull, e_drain, e_setup);
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, kTmemCols);
  }
}

// Row blocks cut by a unit-range boundary: sum their f32 partials in pair order -> dX (and rowdot).
// grid (row blocks, 4): blockIdx.y takes a quarter of the 128 rows.
template <typename T>
__global__ void __launch_bounds__(256)
bwd2p_fixup_kernel(const float* __restrict__ acc_ws, const float* __restrict__ rd_ws, int64_t M, int64_t D, int64_t S,
                   int64_t U, int npairs, const float* __restrict__ ls, const float* __restrict__ go,
                   float scale, T* __restrict__ dX, int64_t lddx, float* __restrict__ rowdot) {
  const int64_t rb = blockIdx.x;
  const int pf = pair_of_unit(rb * S, U, npairs), pl = pair_of_unit((rb + 1) * S - 1, U, npairs);
  if (pf == pl) return;            // the row block lies inside one pair's range: written directly by the kernel
  const float alpha = (go ? go[0] : 1.f) * ls[0] * scale;
  const int64_t r_lo = rb * 128 + 32 * blockIdx.y;
  const int64_t r_hi = (r_lo + 32 < M) ? r_lo + 32 : M;
  const int64_t d4 = D / 4;
  // pair q's piece of this row block is its first item iff its range starts inside the row block
  auto slot_of = [&](int q) -> int64_t { return 2 * (int64_t)q + (unit_lo(q, U, npairs) >= rb * S ? 0 : 1); };
  for (int64_t idx = threadIdx.x; idx < (r_hi - r_lo) * d4; idx += blockDim.x) {
    const int64_t rr = idx / d4, d = (idx - rr * d4) * 4;
    const int64_t row = r_lo + rr;
    const size_t off = (size_t)(row - rb * 128) * D + d;
    float4 a = *reinterpret_cast<const float4*>(acc_ws + (size_t)slot_of(pf) * 128 * D + off);
    for (int q = pf + 1; q <= pl; ++q) {
      const float4 b = *reinterpret_cast<const float4*>(acc_ws + (size_t)slot_of(q) * 128 * D + off);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    T* o = dX + row * lddx + d;
    o[0] = from_f32<T>(a.x * alpha); o[1] = from_f32<T>(a.y * alpha);
    o[2] = from_f32<T>(a.z * alpha); o[3] = from_f32<T>(a.w * alpha);
  }
  if (rowdot != nullptr && threadIdx.x < 32) {
    const int64_t row = r_lo + threadIdx.x;
    if (row < M) {
      float t = 0.f;
      for (int q = pf; q <= pl; ++q) t += rd_ws[(size_t)slot_of(q) * 128 + (row - rb * 128)];
      rowdot[row] = t;
    }
  }
}

constexpr int kMaxPairSlots = 128;  // upper bound used for sizing only; the launch uses pair_slots()
struct Bwd2PPlan { int kch; int kpairs; int ndh; int steps_total; int64_t units; int npairs; bool cut; uint32_t smem; };

Bwd2PPlan plan_bwd2p(int64_t M, int64_t N, int64_t D, int pair_slots) {
  Bwd2PPlan b;
  b.kch = (int)ceil_div(D, 64);
  b.kpairs = (b.kch + 1) / 2;
  b.ndh = (int)ceil_div(D, 256);
  b.steps_total = (int)ceil_div(N, 256);
  b.units = ceil_div(M, 128) * b.steps_total;
  if (pair_slots > kMaxPairSlots) pair_slots = kMaxPairSlots;
  if (pair_slots < 1) pair_slots = 1;
  // at least two steps per pair (a pair costs ~1.5 step times of prologue), never more pairs than slots
  int64_t want = b.units / 2;
  if (want < 1) want = 1;
  if (want > pair_slots) want = pair_slots;
  // prefer whole row blocks per pair when that loses nothing (no partial sums, no fix-up launch)
  const int64_t rbs = ceil_div(M, 128);
  if (rbs <= want && rbs * 2 > want) want = rbs;
  b.npairs = (int)want;
  b.cut = false;
  for (int q = 1; q < b.npairs; ++q)
    if (unit_lo(q, b.units, b.npairs) % b.steps_total != 0) { b.cut = true; break; }
  b.smem = kAlignSlack + 8 * kTile8K + 4 * kTile8K + kRing2 * kStage2 + 1536;
  return b;
}

struct Bwd2PWs { size_t acc, rd, ly2, y16, total; };
Bwd2PWs bwd2p_ws_layout(int npairs, int64_t N, int64_t D, int64_t n_pad, bool need_y16) {
  Bwd2PWs w;
  size_t off = 0;
  w.acc = off; off += align_up((size_t)2 * npairs * 128 * D * sizeof(float), 256);
  w.rd = off;  off += align_up((size_t)2 * npairs * 128 * sizeof(float), 256);
  w.ly2 = off; off += align_up(((size_t)2 * n_pad + 2 * (n_pad / 256) + 64) * sizeof(float), 256);  // ly2, bcol, stepmm, mu0
  w.y16 = off; off += need_y16 ? align_up((size_t)N * D * 2, 256) : 0;
  w.total = off;
  return w;
}

// CTA-pair backward (D <= 512)
// pair slots of the current device for this kernel (clusters of 2, 1 CTA per SM), queried once per device
template <typename K>
int pair_slots_for(K kernel, uint32_t smem, int* out) {
  static std::mutex mu;
  static std::unordered_map<int, int> cache;
  int dev = 0;
  MCLIP_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(dev);
  if (it != cache.end()) { *out = it->second; return MCLIP_OK; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * tc_pair_slots());
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kernel, &cfg);
  if (e != cudaSuccess || n < 1) { (void)cudaGetLastError(); n = tc_pair_slots(); }
  if (n > tc_pair_slots()) n = tc_pair_slots();
  cache[dev] = n;
  *out = n;
  return MCLIP_OK;
}

}  // namespace

size_t tc_block_grad2p_ws(int64_t N, int64_t D, int64_t n_pad) {
  return bwd2p_ws_layout(tc_pair_slots(), N, D, n_pad, true).total;
}

int tc_block_grad2p(const BlockGradArgs& a) {
  const bool bf = a.dtype == MCLIP_DTYPE_BF16;
  int rc;
  int slots = tc_pair_slots();
  {
    const uint32_t smem = kAlignSlack + 8 * kTile8K + 4 * kTile8K + kRing2 * kStage2 + 1536;
    rc = bf ? tc_set_smem(reinterpret_cast<const void*>(tc_block_grad2p_kernel<true>), smem) : tc_set_smem(reinterpret_cast<const void*>(tc_block_grad2p_kernel<false>), smem);
    if (rc) return rc;
    rc = bf ? pair_slots_for(tc_block_grad2p_kernel<true>, smem, &slots) : pair_slots_for(tc_block_grad2p_kernel<false>, smem, &slots);
    if (rc) return rc;
  }
  const Bwd2PPlan b = plan_bwd2p(a.M, a.N, a.D, slots);
  const int64_t n_pad = (int64_t)b.steps_total * 256;
  const Bwd2PWs w = bwd2p_ws_layout(b.npairs, a.N, a.D, n_pad, bf);
  if (w.total > a.ws_bytes) { set_error("block_grad(tcgen05): workspace %zu < %zu", a.ws_bytes, w.total); return MCLIP_ERR_WORKSPACE; }
  uint8_t* ws = reinterpret_cast<uint8_t*>(a.ws);
  Bwd2Prep pr;
  rc = tc_bwd2_prepare(a, n_pad, reinterpret_cast<float*>(ws + w.ly2), ws + w.y16, &pr);
  if (rc) return rc;
  CUtensorMap tmX, tmY, tmY16;
  rc = tc_make_tmap(&tmX, a.X, a.M, a.D, a.ldx, a.dtype, 64);
  if (rc) return rc;
  rc = tc_make_tmap(&tmY, a.Y, a.N, a.D, a.ldy, a.dtype, 128);
  if (rc) return rc;
  rc = tc_make_tmap(&tmY16, pr.y16, a.N, a.D, pr.ld16, MCLIP_DTYPE_F16, 128);
  if (rc) return rc;
  Bwd2PParams p;
  p.M = a.M; p.N = a.N; p.D = a.D; p.kpairs = b.kpairs; p.ndh = b.ndh; p.steps_total = b.steps_total;
  p.units = b.units; p.npairs = b.npairs; p.diag_off = a.diag_off; p.ls = a.logit_scale; p.go = a.grad_out;
  p.lse_x = a.lse_x; p.ly2 = pr.has_col ? pr.ly2 : nullptr; p.bcol = pr.bcol; p.stepmm = pr.stepmm; p.mu0 = pr.mu0;
  p.w_row = a.w_row; p.w_diag = a.w_diag; p.inv_2n = a.inv_2n; p.has_col = pr.has_col ? 1 : 0;
  p.dX = a.dX; p.lddx = a.lddx;
  p.acc_ws = reinterpret_cast<float*>(ws + w.acc); p.rd_ws = reinterpret_cast<float*>(ws + w.rd); p.rowdot = a.rowdot;
  p.dbg = tc_dbg_flags();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * b.npairs));
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
  if (bf) MCLIP_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_block_grad2p_kernel<true>, tmX, tmY, tmY16, p));
  else MCLIP_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_block_grad2p_kernel<false>, tmX, tmY, tmY16, p));
  count_launch();
  MCLIP_CUDA_OK(cudaGetLastError());
  if (b.cut) {
    const dim3 grid((unsigned)ceil_div(a.M, 128), 4);
    const float scale = a.inv_2n * (1.f / kGScale);
    if (bf)
      bwd2p_fixup_kernel<__nv_bfloat16><<<grid, 256, 0, a.stream>>>(p.acc_ws, p.rd_ws, a.M, a.D, b.steps_total, b.units, b.npairs,
                                                                   a.logit_scale, a.grad_out, scale,
                                                                   reinterpret_cast<__nv_bfloat16*>(a.dX), a.lddx, a.rowdot);
    else
      bwd2p_fixup_kernel<__half><<<grid, 256, 0, a.stream>>>(p.acc_ws, p.rd_ws, a.M, a.D, b.steps_total, b.units, b.npairs,
                                                            a.logit_scale, a.grad_out, scale, reinterpret_cast<__half*>(a.dX),
                                                            a.lddx, a.rowdot);
    count_launch();
    MCLIP_CUDA_OK(cudaGetLastError());
  }
  return MCLIP_OK;
}


}  // namespace mclip
