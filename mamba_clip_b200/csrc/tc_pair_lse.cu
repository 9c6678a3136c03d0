// Two-sided forward of the contrastive loss on tcgen05 (sm_100a): ONE pass over S = ls * X Y^T yields the row
// log-sum-exps (image -> text direction) AND the column sums (text -> image direction), so the forward executes the
// logits GEMM once instead of once per direction (reference loss.py:102-111 builds both logits blocks, :142-145 runs
// F.cross_entropy on each).
//
// Every exponential is taken against ONE uniform reference c0 (log2 units): e_ij = 2^(k2 * <x_i, y_j> - c0).  Row
// sums and column sums of the same e_ij are then plain sums -- no running maximum, no rescaling, partial sums of
// different CTAs / ranks simply add.  c0 is picked from the positive-pair logits (max_i k2 * <x_i, y_i+off> - 15);
// whether that choice was good enough is checked on the RESULT: every row / column total must lie in
// [2^-75, 2^120].  Inside that window terms flushed to zero by f32 (< 2^-126) are below 2^-30 of the total even for
// 2^20 of them and nothing overflowed.  Outside it the call raises a device-side status flag and the caller's
// predicated one-sided kernels (tc_row_lse with `run_if`) redo the work with per-row running maxima.
//
// Structure: the CTA-pair pipeline of tc_row_lse2_kernel (cta_group::2, M = 256, N = 256, X resident, Y streamed
// through a TMA ring, two TMEM accumulators).  The epilogue reads TMEM with tcgen05.ld.16x256b: a thread then owns
// 4 rows x 8 columns of a 32 x 32 block instead of 1 row x 32 columns, so the column reduction over the 32 lanes of a
// TMEM quadrant needs 3 recursive-halving shuffle stages on 8 values (7 SHFL) instead of 5 stages on 32 (31 SHFL).
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tc_host.cuh"

namespace mclip {

namespace {

using namespace ptx;

constexpr int kThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kEpiThreads = 256;
constexpr uint32_t kChunkBytes = 128 * 64 * 2;   // [128 rows x 64 k] 16-bit
constexpr int kStages = 5;                       // one stage less than tc_row_lse2_kernel: room for the column buffers
constexpr uint32_t kColBufBytes = 2 * 2 * 4 * 128 * 4;   // [parity][half][quadrant][128 columns] f32
constexpr uint32_t kMiscBytes = 2048;
constexpr uint32_t kAlignSlack = 1024;
// c0 = max positive-pair logit (log2 units) - kRefMargin.  The exponent t = s*k2 - c0 is rounded once (FFMA) relative to
// |t|, and for a saturated row that rounding goes straight into its LSE: with 15 the dominant terms have |t| < 16
// (ulp 9.5e-7).  Window arithmetic: logits up to ~88 log2 units above the largest positive pair and row / column maxima
// down to ~90 below it stay inside [2^-75, 2^120] for N = 2^17.
constexpr float kRefMargin = 15.f;
constexpr float kSumLo = 2.6469779601696886e-23f;   // 2^-75
constexpr float kSumHi = 1.329227995784916e36f;     // 2^120

struct PairParams {
  int64_t M, N;
  int kch;              // ceil(D / 64)
  int tiles_total;      // ceil(N / 256)
  int tiles_per_split;
  int64_t n_pad;        // tiles_total * 256: row stride of part_cs
  const float* ls;
  const float* ref;     // [1] c0 in log2 units
  float* part_rs;       // [nsplit][M] row sums of e
  float* part_rc;       // [nsplit][M] row sums of e * <x_i, y_j>
  float* part_cs;       // [ceil(M / 128)][n_pad] column sums of e over each 128-row block
  float* diag;          // [M] or null: overwritten with the accumulator's own <x_i, y_i+off> (see pair_chunk)
  int64_t diag_off;
  int bf16;
};

__device__ __forceinline__ uint32_t align1024(uint32_t a) { return (a + 1023u) & ~1023u; }

// 16 TMEM lanes x 32 columns; thread t: register j -> lane t/4 + 8*((j%4)/2), column 8*(j/4) + 2*(t%4) + (j%2)
// (layout verified on the GPU by tools/tmem_layout_probe.cu)
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// One 32 x 32 block of S held as 4 rows (r + 8*ri) x 8 columns (8*(ci/2) + 2*qd + ci%2) per thread.
// v[0..15]: lanes +0..15 of the quadrant, v[16..31]: lanes +16..31.
// kMasked (tail tile in N, or a tile crossing the diagonal): columns >= N contribute nothing, and the positive-pair
// dot is taken from the accumulator itself -- the loss subtracts ls * diag from an LSE that is dominated by the very
// same product when the softmax is saturated, so both must carry the same tensor-core rounding.
template <bool kMasked>
__device__ __forceinline__ void pair_chunk(const uint32_t (&v)[32], float k2, const float (&bias)[4], int64_t colq,
                                           int64_t N, float (&rs)[4], float (&rc)[4], float (&cs)[8],
                                           const int64_t (&jd)[4], float* __restrict__ diag_rows) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int h = j >> 4, jj = j & 15, rep = jj >> 2;
    const int ri = 2 * h + ((jj & 3) >> 1), ci = 2 * rep + (jj & 1);
    const float c = __uint_as_float(v[j]);
    float e = ex2_approx(fmaf(c, k2, bias[ri]));
    if (kMasked) {
      const int64_t col = colq + 8 * rep + (jj & 1);
      if (col >= N) e = 0.f;
      else if (col == jd[ri]) diag_rows[8 * ri] = c;
    }
    rs[ri] += e;
    rc[ri] = fmaf(e, c, rc[ri]);
    cs[ci] = (ri == 0) ? e : cs[ci] + e;
  }
}

// Recursive halving over the 8 row groups (lane bits 4, 3, 2): afterwards the lane holds the sum over the quadrant's
// 32 rows of column 16*b4 + 8*b3 + 2*(lane%4) + b2 of the 32-column chunk.
__device__ __forceinline__ float col_halving(const float (&cs)[8], int lane) {
  const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
  float a[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float send = b4 ? cs[k] : cs[k + 4];
    const float keep = b4 ? cs[k + 4] : cs[k];
    a[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  float b[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float send = b3 ? a[k] : a[k + 2];
    const float keep = b3 ? a[k + 2] : a[k];
    b[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  const float send = b2 ? b[0] : b[1];
  const float keep = b2 ? b[1] : b[0];
  return keep + __shfl_xor_sync(0xffffffffu, send, 4);
}

__global__ void __launch_bounds__(kThreads, 1)
tc_pair_lse2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const PairParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int BN = 256;
  const uint32_t smem_base = align1024(smem_u32(smem_raw));
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t x_bytes = 8 * kChunkBytes;                       // [8][128 rows][64 k]
  const uint32_t ring_base = smem_base + x_bytes;                 // [kStages][128 y][64 k]
  const uint32_t colbuf_base = ring_base + kStages * kChunkBytes;
  float* colbuf = reinterpret_cast<float*>(smem_gen + (colbuf_base - smem_base));
  const uint32_t misc_base = colbuf_base + kColBufBytes;
  uint8_t* misc_gen = smem_gen + (misc_base - smem_base);
  float2* merge = reinterpret_cast<float2*>(misc_gen);            // [128] (row sum, row dot) of column half 1
  const uint32_t bar_base = misc_base + 1024;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                 // leader
  auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };          // per CTA (multicast commit)
  const uint32_t xfull_bar = bar_base + 8u * 16;                            // leader
  auto tfull_bar = [&](int b) { return bar_base + 8u * (17 + b); };         // per CTA (multicast commit)
  auto tempty_bar = [&](int b) { return bar_base + 8u * (19 + b); };        // leader: 16 epilogue warps
  const uint32_t tmem_slot = bar_base + 8u * 21;
  uint32_t* tmem_slot_gen = reinterpret_cast<uint32_t*>(misc_gen + 1024 + 8u * 21);

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
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
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
        for (int c = 0; c < p.kch; ++c) {
          if (c >= 8) {
            // 512 < D <= 768: only the first 8 k-chunks of the X block are resident (128 KB); chunks 8..11 travel through
            // the ring in front of their Y chunk, once per column tile (+33 % ring traffic: 42 B/clk/SM, under the
            // ingest limit that the streamed-everything single-CTA kernel sits on)
            const int s = it % kStages;
            const uint32_t ph = (it / kStages) & 1;
            mbar_wait(empty_bar(s), ph ^ 1);
            if (leader) mbar_expect_tx(full_bar(s), 2 * kChunkBytes); else mbar_arrive_cluster(full_bar(s), 0);
            tma_load_2d_cg2(ring_base + s * kChunkBytes, &tmX, c * 64, (int32_t)m0, full_bar(s));
            ++it;
          }
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          if (leader) mbar_expect_tx(full_bar(s), 2 * kChunkBytes); else mbar_arrive_cluster(full_bar(s), 0);
          tma_load_2d_cg2(ring_base + s * kChunkBytes, &tmY, c * 64, y0, full_bar(s));
          ++it;
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      const bool elected = elect_one();
      const uint32_t idesc = make_idesc_f16(p.bf16 != 0, p.bf16 != 0, 256, BN, false, false);
      mbar_wait(xfull_bar, 0);
      uint32_t it = 0;
      for (int lt = 0; lt < ntiles; ++lt) {
        const int buf = lt & 1;
        const uint32_t bph = (lt >> 1) & 1;
        mbar_wait(tempty_bar(buf), bph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int c = 0; c < p.kch; ++c) {
          uint32_t a_addr = smem_base + c * kChunkBytes;
          int sx = -1;
          if (c >= 8) {                      // streamed X chunk (see the producer)
            sx = it % kStages;
            mbar_wait(full_bar(sx), (it / kStages) & 1);
            a_addr = ring_base + sx * kChunkBytes;
            ++it;
          }
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          ++it;
          const uint32_t b_addr = ring_base + s * kChunkBytes;
          if (elected) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = make_smem_desc_sw128(a_addr + k * 32, 0, 1024);
              const uint64_t bd = make_smem_desc_sw128(b_addr + k * 32, 0, 1024);
              mma_ss_cg2(d_tmem, ad, bd, idesc, (c | k) != 0);
            }
            if (sx >= 0) mma_commit_cg2(empty_bar(sx), 3);
            mma_commit_cg2(empty_bar(s), 3);
            if (c == p.kch - 1) mma_commit_cg2(tfull_bar(buf), 3);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    const int ew = warp - kEpiWarp0;
    const int q = warp & 3;            // TMEM lane quadrant of this warp
    const int half = ew >> 2;          // column half of the tile
    const int r = lane >> 2, qd = lane & 3;
    const float k2 = p.ls[0] * kLog2e;
    const float c0 = p.ref[0];
    float bias[4], rs[4], rc[4];
    int64_t jd[4];
#pragma unroll
    for (int ri = 0; ri < 4; ++ri) {
      const int64_t row = m0 + q * 32 + r + 8 * ri;
      bias[ri] = row < p.M ? -c0 : -INFINITY;   // rows past M hold TMA zero fill: e = 2^-inf = 0 in every sum
      jd[ri] = (p.diag != nullptr && row < p.M) ? row + p.diag_off : -1;
      rs[ri] = 0.f;
      rc[ri] = 0.f;
    }
    float* diag_rows = p.diag + (m0 + q * 32 + r);   // only dereferenced where jd matched (row < M)
    // column this lane owns after col_halving, inside a 32-column chunk
    const int own_col = 16 * ((lane >> 4) & 1) + 8 * ((lane >> 3) & 1) + 2 * qd + ((lane >> 2) & 1);
    float* part_cs_row = p.part_cs + (size_t)(m0 / 128) * p.n_pad;
    constexpr int kHalfCols = BN / 2;
    for (int lt = 0; lt < ntiles; ++lt) {
      const int buf = lt & 1;
      const uint32_t bph = (lt >> 1) & 1;
      const int64_t n0 = (int64_t)(t0 + lt) * BN;
      mbar_wait(tfull_bar(buf), bph);
      tc_fence_after();
      const bool tail = (n0 + BN > p.N) ||
                        (p.diag != nullptr && n0 < m0 + p.diag_off + 128 && n0 + BN > m0 + p.diag_off);
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + half * kHalfCols;
      float* cb = colbuf + ((lt & 1) * 2 + half) * 512 + q * 128;
#pragma unroll 1
      for (int cc = 0; cc < kHalfCols / 32; ++cc) {
        uint32_t v[32];
        tmem_ld_16x256b_x4(t_addr + cc * 32, v);
        tmem_ld_16x256b_x4(t_addr + (16u << 16) + cc * 32, v + 16);
        tmem_ld_wait();
        if (cc == kHalfCols / 32 - 1) {
          // the accumulator is in registers: hand the TMEM buffer back before the math of the last chunk
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(tempty_bar(buf)); else mbar_arrive_cluster(tempty_bar(buf), 0);
          }
        }
        float cs[8];
        const int64_t colq = n0 + half * kHalfCols + cc * 32 + 2 * qd;
        if (tail) pair_chunk<true>(v, k2, bias, colq, p.N, rs, rc, cs, jd, diag_rows);
        else pair_chunk<false>(v, k2, bias, colq, p.N, rs, rc, cs, jd, diag_rows);
        cb[cc * 32 + own_col] = col_halving(cs, lane);
      }
      // sum the four quadrants of this column half and write the 128-row-block partial
      named_bar_sync(2 + half, 128);
      {
        const float* cq = colbuf + ((lt & 1) * 2 + half) * 512 + q * 32 + lane;
        const float tot = (cq[0] + cq[128]) + (cq[256] + cq[384]);
        part_cs_row[n0 + half * kHalfCols + q * 32 + lane] = tot;
      }
    }
    // rows: combine the 4 lanes of a quad, then the two column halves
#pragma unroll
    for (int ri = 0; ri < 4; ++ri) {
      rs[ri] += __shfl_xor_sync(0xffffffffu, rs[ri], 1);
      rs[ri] += __shfl_xor_sync(0xffffffffu, rs[ri], 2);
      rc[ri] += __shfl_xor_sync(0xffffffffu, rc[ri], 1);
      rc[ri] += __shfl_xor_sync(0xffffffffu, rc[ri], 2);
    }
    const float my_rs = qd == 0 ? rs[0] : qd == 1 ? rs[1] : qd == 2 ? rs[2] : rs[3];
    const float my_rc = qd == 0 ? rc[0] : qd == 1 ? rc[1] : qd == 2 ? rc[2] : rc[3];
    const int row_in_tile = q * 32 + r + 8 * qd;
    const int64_t row = m0 + row_in_tile;
    if (half == 1) merge[row_in_tile] = make_float2(my_rs, my_rc);
    named_bar_sync(1, kEpiThreads);
    if (half == 0 && row < p.M) {
      const float2 o = merge[row_in_tile];
      p.part_rs[(int64_t)blockIdx.y * p.M + row] = my_rs + o.x;
      p.part_rc[(int64_t)blockIdx.y * p.M + row] = my_rc + o.y;
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, kTmemCols);
  }
}

// ---- O(B) helpers ---------------------------------------------------------------------------------------------
// diag[i] = <X[i], Y[i + off]> (raw, f32 accumulate; 0 outside [0, N)); per-block max / min of the valid ones.
// One warp per row; kVec: 16-byte loads (8 x 16-bit) when D, the leading dimensions and the bases allow it.
template <typename T> __device__ __forceinline__ float dot8(const uint4& a, const uint4& b);
template <> __device__ __forceinline__ float dot8<__nv_bfloat16>(const uint4& a, const uint4& b) {
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {   // bf16 -> f32 is a 16-bit shift
    acc = fmaf(__uint_as_float(aw[k] << 16), __uint_as_float(bw[k] << 16), acc);
    acc = fmaf(__uint_as_float(aw[k] & 0xffff0000u), __uint_as_float(bw[k] & 0xffff0000u), acc);
  }
  return acc;
}
template <> __device__ __forceinline__ float dot8<__half>(const uint4& a, const uint4& b) {
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&aw[k]));
    const float2 fb = __half22float2(*reinterpret_cast<const __half2*>(&bw[k]));
    acc = fmaf(fa.x, fb.x, acc);
    acc = fmaf(fa.y, fb.y, acc);
  }
  return acc;
}
template <> __device__ __forceinline__ float dot8<float>(const uint4&, const uint4&) { return 0.f; }

template <typename T, bool kVec>
__global__ void __launch_bounds__(256)
diag_dots_kernel(const T* __restrict__ X, const T* __restrict__ Y, int64_t M, int64_t N, int64_t D, int64_t ldx,
                 int64_t ldy, int64_t off, float* __restrict__ diag, float* __restrict__ blk_max,
                 float* __restrict__ blk_min) {
  __shared__ float smax[8], smin[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float vmax = -INFINITY, vmin = INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * 8 + warp; i < M; i += (int64_t)gridDim.x * 8) {
    const int64_t j = i + off;
    float acc = 0.f;
    const bool valid = j >= 0 && j < N;
    if (valid) {
      const T* x = X + i * ldx;
      const T* y = Y + j * ldy;
      if (kVec) {
        const uint4* x4 = reinterpret_cast<const uint4*>(x);
        const uint4* y4 = reinterpret_cast<const uint4*>(y);
        for (int64_t d = lane; d < D / 8; d += 32) acc += dot8<T>(__ldg(x4 + d), __ldg(y4 + d));
      } else {
        for (int64_t d = lane; d < D; d += 32) acc = fmaf(to_f32<T>(x[d]), to_f32<T>(y[d]), acc);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      vmax = fmaxf(vmax, acc);
      vmin = fminf(vmin, acc);
    }
    if (lane == 0) diag[i] = acc;
  }
  if (lane == 0) { smax[warp] = vmax; smin[warp] = vmin; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { vmax = fmaxf(vmax, smax[w]); vmin = fminf(vmin, smin[w]); }
    blk_max[blockIdx.x] = vmax;
    blk_min[blockIdx.x] = vmin;
  }
}

// ref[0] = c0 = max over the positive pairs of k2 * dot - margin (0 when there is no positive pair); status = 0.
__global__ void pair_ref_kernel(const float* __restrict__ blk_max, const float* __restrict__ blk_min, int nblk,
                                const float* __restrict__ ls, float* __restrict__ ref, int* __restrict__ status) {
  __shared__ float smax[32], smin[32];
  float vmax = -INFINITY, vmin = INFINITY;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) { vmax = fmaxf(vmax, blk_max[i]); vmin = fminf(vmin, blk_min[i]); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
  }
  if ((threadIdx.x & 31) == 0) { smax[threadIdx.x >> 5] = vmax; smin[threadIdx.x >> 5] = vmin; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { vmax = fmaxf(vmax, smax[w]); vmin = fminf(vmin, smin[w]); }
    const float k2 = ls[0] * kLog2e;
    float c0 = 0.f;
    if (vmax >= vmin) c0 = fmaxf(k2 * vmax, k2 * vmin) - kRefMargin;
    if (!(fabsf(c0) < 1e30f)) c0 = 0.f;
    ref[0] = c0;
    if (status) *status = 0;
  }
}

// rows: lse = ln2 * (c0 + log2(sum of the split partials)), rowdot = sum(e c) / sum(e); totals outside the window
// raise the status flag.
__global__ void pair_rows_finalize_kernel(const float* __restrict__ part_rs, const float* __restrict__ part_rc, int nsplit,
                                          int64_t M, const float* __restrict__ ref, float* __restrict__ lse,
                                          float* __restrict__ rowdot, int* __restrict__ status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  float s = 0.f, c = 0.f;
  for (int k = 0; k < nsplit; ++k) { s += part_rs[(int64_t)k * M + i]; c += part_rc[(int64_t)k * M + i]; }
  lse[i] = (float)(((double)ref[0] + log2((double)s)) * 0.6931471805599453);   // one rounding, at the very end
  if (rowdot != nullptr) rowdot[i] = c / s;
  if (!(s >= kSumLo && s <= kSumHi)) atomicOr(status, 1);
}

// columns: sum the per-row-block partials in fixed order.  mode 0: out = column lse; mode 1: out = raw column sum
// (still relative to c0: partial sums of other ranks add to it).  64 columns x 4 row-block groups per CTA.
__global__ void __launch_bounds__(256)
pair_cols_finalize_kernel(const float* __restrict__ part_cs, int nrb, int64_t n_pad, int64_t N, const float* __restrict__ ref,
                          int mode, float* __restrict__ out, int* __restrict__ status) {
  __shared__ float sm[4][64];
  const int cx = threadIdx.x & 63, g = threadIdx.x >> 6;
  const int64_t j = (int64_t)blockIdx.x * 64 + cx;
  float s = 0.f;
  if (j < N) {
    const int per = (nrb + 3) / 4;
    const int b0 = g * per, b1 = min(nrb, b0 + per);
    for (int b = b0; b < b1; ++b) s += part_cs[(size_t)b * n_pad + j];
  }
  sm[g][cx] = s;
  __syncthreads();
  if (g == 0 && j < N) {
    const float tot = (sm[0][cx] + sm[1][cx]) + (sm[2][cx] + sm[3][cx]);
    if (mode == 0) {
      out[j] = (float)(((double)ref[0] + log2((double)tot)) * 0.6931471805599453);
      if (!(tot >= kSumLo && tot <= kSumHi)) atomicOr(status, 2);
    } else {
      out[j] = tot;
      if (!(tot <= kSumHi)) atomicOr(status, 2);   // the lower bound is checked on the cross-rank total
    }
  }
}

__global__ void lse_from_sum_kernel(const float* __restrict__ sum, int64_t n, const float* __restrict__ ref,
                                    float* __restrict__ lse, int* __restrict__ status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s = sum[i];
  lse[i] = (float)(((double)ref[0] + log2((double)s)) * 0.6931471805599453);
  if (!(s >= kSumLo && s <= kSumHi)) atomicOr(status, 2);
}

// col_mode 1: the reference and the status word travel with the raw column sums (slots N and N + 1 of col_out), so
// one all-gather hands every rank everything it needs to finish its own columns.
__global__ void pair_pack_kernel(const float* __restrict__ ref, const int* __restrict__ status, float* __restrict__ tail) {
  tail[0] = ref[0];
  tail[1] = __int_as_float(*status);
}

// Column LSEs of this rank's columns [col0, col0 + n) from the W gathered partial-sum vectors (each relative to its
// own reference): lse = ln2 * (m + log2(sum_q part_q * 2^(c_q - m))), m = max_q c_q.  ORs every rank's status word
// into *status (a rank whose rows overflowed poisons everybody's columns) plus 2 for an out-of-window total.
__global__ void merge_col_sums_kernel(const float* __restrict__ parts, int W, int64_t stride, int64_t n_total, int64_t col0,
                                      int64_t n, float* __restrict__ lse, int* __restrict__ status) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float m = -INFINITY;
  int bits = 0;
  for (int q = 0; q < W; ++q) {
    m = fmaxf(m, parts[(size_t)q * stride + n_total]);
    bits |= __float_as_int(parts[(size_t)q * stride + n_total + 1]);
  }
  if (j == 0 && bits != 0) atomicOr(status, bits);
  if (j >= n) return;
  float tot = 0.f;
  for (int q = 0; q < W; ++q) {
    const float cq = parts[(size_t)q * stride + n_total];
    tot += parts[(size_t)q * stride + col0 + j] * exp2f(cq - m);
  }
  lse[j] = (float)(((double)m + log2((double)tot)) * 0.6931471805599453);
  if (!(tot >= kSumLo && tot <= kSumHi)) atomicOr(status, 2);
}

struct PairPlan { int kch; int tiles_total; int nsplit; int tiles_per_split; uint32_t smem; int nrb; int64_t n_pad; };

PairPlan plan_pair(int64_t M, int64_t N, int64_t D) {
  PairPlan f;
  f.kch = (int)ceil_div(D, 64);
  f.tiles_total = (int)ceil_div(N, 256);
  const int64_t pairs = ceil_div(M, 256);
  int best = 1;
  double best_cost = 1e30;
  const int max_split = f.tiles_total < 64 ? f.tiles_total : 64;
  for (int s = 1; s <= max_split; ++s) {
    const int tps = (int)ceil_div(f.tiles_total, s);
    const int real = (int)ceil_div(f.tiles_total, tps);
    if (real != s) continue;
    const double waves = (double)ceil_div(pairs * s, 74);
    const double cost = waves * (tps + 1.5);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  f.nsplit = best;
  f.tiles_per_split = (int)ceil_div(f.tiles_total, best);
  f.smem = kAlignSlack + 8 * kChunkBytes + kStages * kChunkBytes + kColBufBytes + kMiscBytes;
  f.nrb = (int)(2 * ceil_div(M, 256));
  f.n_pad = (int64_t)f.tiles_total * 256;
  return f;
}

constexpr int kDiagBlocks = 592;   // 4 x 148

struct PairWs { size_t rs, rc, cs, total; };
PairWs pair_ws_layout(const PairPlan& f, int64_t M) {
  PairWs w;
  size_t off = 0;
  w.rs = off; off += align_up((size_t)f.nsplit * M * sizeof(float), 256);
  w.rc = off; off += align_up((size_t)f.nsplit * M * sizeof(float), 256);
  w.cs = off; off += align_up((size_t)f.nrb * f.n_pad * sizeof(float), 256);
  w.total = off;
  return w;
}

}  // namespace

bool tc_pair_supported(int64_t M, int64_t N, int64_t D, int64_t ldx, int64_t ldy, int dtype) {
  (void)M; (void)N;
  if (dtype != MCLIP_DTYPE_BF16 && dtype != MCLIP_DTYPE_F16) return false;
  if (D % 8 != 0 || D > 768) return false;
  if (ldx % 8 != 0 || ldy % 8 != 0) return false;
  return true;
}

size_t tc_pair_lse_ws(int64_t M, int64_t N, int64_t D) { return pair_ws_layout(plan_pair(M, N, D), M).total; }

size_t pair_ref_ws() { return align_up((size_t)2 * kDiagBlocks * sizeof(float), 256); }

template <typename T, bool kVec>
void launch_diag_dots(const PairRefArgs& a, int nblk, float* blk_max, float* blk_min) {
  diag_dots_kernel<T, kVec><<<nblk, 256, 0, a.stream>>>(reinterpret_cast<const T*>(a.X), reinterpret_cast<const T*>(a.Y), a.M,
                                                       a.N, a.D, a.ldx, a.ldy, a.diag_off, a.diag, blk_max, blk_min);
}

int launch_pair_ref(const PairRefArgs& a) {
  float* blk_max = reinterpret_cast<float*>(a.ws);
  float* blk_min = blk_max + kDiagBlocks;
  const int nblk = (int)(ceil_div(a.M, 8) < kDiagBlocks ? ceil_div(a.M, 8) : kDiagBlocks);
  const bool vec = a.dtype != MCLIP_DTYPE_F32 && a.D % 8 == 0 && a.ldx % 8 == 0 && a.ldy % 8 == 0 &&
                   (((uintptr_t)a.X | (uintptr_t)a.Y) & 15) == 0;
  switch (a.dtype) {
    case MCLIP_DTYPE_F32: launch_diag_dots<float, false>(a, nblk, blk_max, blk_min); break;
    case MCLIP_DTYPE_BF16:
      if (vec) launch_diag_dots<__nv_bfloat16, true>(a, nblk, blk_max, blk_min);
      else launch_diag_dots<__nv_bfloat16, false>(a, nblk, blk_max, blk_min);
      break;
    default:
      if (vec) launch_diag_dots<__half, true>(a, nblk, blk_max, blk_min);
      else launch_diag_dots<__half, false>(a, nblk, blk_max, blk_min);
  }
  count_launch();
  MCLIP_CUDA_OK(cudaGetLastError());
  pair_ref_kernel<<<1, 256, 0, a.stream>>>(blk_max, blk_min, nblk, a.logit_scale, a.ref, a.status);
  count_launch();
  MCLIP_CUDA_OK(cudaGetLastError());
  return MCLIP_OK;
}

int launch_lse_from_sum(const float* sum, int64_t n, const float* ref, float* lse, int* status, cudaStream_t stream) {
  lse_from_sum_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(sum, n, ref, lse, status);
  count_launch();
  MCLIP_CUDA_OK(cudaGetLastError());
  return MCLIP_OK;
}

int tc_pair_lse(const PairLseArgs& a) {
  if (((uintptr_t)a.X | (uintptr_t)a.Y) & 15) { set_error("pair_lse: X/Y must be 16-byte aligned"); return MCLIP_ERR_INVALID; }
  const PairPlan f = plan_pair(a.M, a.N, a.D);
  const PairWs w = pair_ws_layout(f, a.M);
  CUtensorMap tmX, tmY;
  int rc = tc_make_tmap(&tmX, a.X, a.M, a.D, a.ldx, a.dtype, 128);
  if (rc) return rc;
  rc = tc_make_tmap(&tmY, a.Y, a.N, a.D, a.ldy, a.dtype, 128);
  if (rc) return rc;
  uint8_t* ws = reinterpret_cast<uint8_t*>(a.ws);
  PairParams p;
  p.M = a.M; p.N = a.N; p.kch = f.kch; p.tiles_total = f.tiles_total; p.tiles_per_split = f.tiles_per_split;
  p.n_pad = f.n_pad; p.ls = a.logit_scale; p.ref = a.ref;
  p.part_rs = reinterpret_cast<float*>(ws + w.rs);
  p.part_rc = reinterpret_cast<float*>(ws + w.rc);
  p.part_cs = reinterpret_cast<float*>(ws + w.cs);
  p.diag = a.diag; p.diag_off = a.diag_off;
  p.bf16 = a.dtype == MCLIP_DTYPE_BF16;
  rc = tc_set_smem(reinterpret_cast<const void*>(tc_pair_lse2_kernel), f.smem);
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
  MCLIP_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_pair_lse2_kernel, tmX, tmY, p));
  count_launch();
  MCLIP_CUDA_OK(cudaGetLastError());
  pair_rows_finalize_kernel<<<(unsigned)ceil_div(a.M, 256), 256, 0, a.stream>>>(p.part_rs, p.part_rc, f.nsplit, a.M, a.ref,
                                                                               a.row_lse, a.rowdot, a.status);
  count_launch();
  MCLIP_CUDA_OK(cudaGetLastError());
  pair_cols_finalize_kernel<<<(unsigned)ceil_div(a.N, 64), 256, 0, a.stream>>>(p.part_cs, f.nrb, f.n_pad, a.N, a.ref,
                                                                              a.col_mode, a.col_out, a.status);
  count_launch();
  MCLIP_CUDA_OK(cudaGetLastError());
  if (a.col_mode == 1) {
    pair_pack_kernel<<<1, 1, 0, a.stream>>>(a.ref, a.status, a.col_out + a.N);
    count_launch();
    MCLIP_CUDA_OK(cudaGetLastError());
  }
  return MCLIP_OK;
}

int launch_merge_col_sums(const float* parts, int W, int64_t stride, int64_t n_total, int64_t col0, int64_t n, float* lse,
                          int* status, cudaStream_t stream) {
  merge_col_sums_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(parts, W, stride, n_total, col0, n, lse, status);
  count_launch();
  MCLIP_CUDA_OK(cudaGetLastError());
  return MCLIP_OK;
}

}  // namespace mclip
