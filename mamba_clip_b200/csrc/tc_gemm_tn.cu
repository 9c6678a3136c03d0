// dY_acc[N, D] (f32) += G[K, N]^T @ X16[K, D]     -- the second half of the shared-recompute backward (tc_fused_grad).
//
// G is one row panel of the f16 * 2^12 gradient-of-logits matrix, written by tc_block_grad2_kernel<.., kStoreG = true>
// (K = panel rows of X, N = all rows of Y), X16 the f16 copy of the same X panel.  Replaces, together with that kernel,
// the MmBackward pair `dT = G^T I` of reference loss.py:102-111 without a second recompute of S.
//
// Persistent CTA-pair kernel, one pair per SM pair, output tiles of 256 rows x D (<= 512) columns:
//   tcgen05.mma.cta_group::2, M = 256 (128 output rows per CTA), N = 256 per instruction (D in 256-wide halves), K = 16.
//   A = G^T: M (= y index) is the contiguous dimension of a G tile -> "MN-major" A operand, boxes [64 k-rows x 64 y] of
//       8 KB -- the tile-major scratch holds exactly these boxes as contiguous blocks, in the swizzled byte order the
//       producing kernel had them in shared memory (both sides use the same SWIZZLE_128B box);
//   B = X16: N (= d) contiguous -> MN-major B operand, boxes [64 k-rows x 64 d]; each CTA supplies its 128-wide half.
//   Ring of 4 stages x 48 KB (16 KB of A + 32 KB of B per CTA and 64 k-rows) = 8 MMAs = 1024 clk per stage: ~47 B/clk/SM.
//   The accumulator [128 x 512] f32 fills TMEM; the epilogue warps move it out through 4 KB swizzled staging tiles and
//   cp.reduce.async.bulk.tensor (.add.f32) into the global accumulator -- every output element is owned by exactly one
//   CTA per launch and launches are stream-ordered, so the sum order is fixed (deterministic).
#include <cuda.h>

#include <cstdio>

#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tc_host.cuh"

namespace mclip {

namespace {

using namespace ptx;

constexpr int kThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 8;
constexpr uint32_t kBox = 64 * 64 * 2;          // [64 k-rows x 64 elements] f16
constexpr uint32_t kABytes = 2 * kBox;          // 128 y per CTA
constexpr uint32_t kBBytes = 4 * kBox;          // 2 halves of D x 128 d per CTA
constexpr uint32_t kStage = kABytes + kBBytes;  // 48 KB
constexpr int kStages = 4;
constexpr uint32_t kStageTile = 32 * 32 * 4;    // epilogue staging tile [32 rows x 32 f32]
constexpr uint32_t kStagingBytes = kEpiWarps * kStageTile;       // 32 KB: one staging tile per warp
constexpr uint32_t kSmemBytes = 1024 + kStages * kStage + kStagingBytes + 1024;

#ifdef MCLIP_PROFILE
constexpr bool kProfile = true;
#else
constexpr bool kProfile = false;
#endif

struct GemmTnParams {
  int dbg;
  int overwrite;         // 1: acc = product (first panel: no memset, no read-modify-write); 0: acc += product
  int g_tiles_per_row;   // 64-column tiles per 64-row block of the tile-major G scratch
  int kchunks;   // ceil(K / 64)
  int tiles;     // ceil(N / 256)
  int ndh;       // ceil(D / 256)
};

__device__ __forceinline__ uint32_t align1024(uint32_t a) { return (a + 1023u) & ~1023u; }

__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_tn_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX16,
                  const __grid_constant__ CUtensorMap tmAcc, const GemmTnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = align1024(smem_u32(smem_raw));
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t ring_base = smem_base;
  const uint32_t staging_base = ring_base + kStages * kStage;
  const uint32_t bar_base = staging_base + kStagingBytes;
  uint8_t* misc_gen = smem_gen + (bar_base - smem_base);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };            // leader: both CTAs' TMA bytes
  auto empty_bar = [&](int s) { return bar_base + 8u * (4 + s); };     // per CTA, MMA commit multicast
  const uint32_t tfull_bar = bar_base + 8u * 8;                         // per CTA, multicast: accumulator complete
  const uint32_t tempty_bar = bar_base + 8u * 9;                        // leader: 16 epilogue warps drained it
  const uint32_t tmem_slot = bar_base + 8u * 10;
  uint32_t* tmem_slot_gen = reinterpret_cast<uint32_t*>(misc_gen + 8u * 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const uint32_t stage_bytes = kABytes + (uint32_t)p.ndh * 2 * kBox;   // per CTA

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmG);
    tma_prefetch_desc(&tmX16);
    tma_prefetch_desc(&tmAcc);
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 2 * kEpiWarps);
    fence_barrier_init();
  } else if (warp == 2) {
    tmem_alloc_cg2(tmem_slot, 512);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer (both CTAs) ----------------
      uint32_t it = 0;
      for (int t = pair; t < p.tiles; t += npairs) {
        const int32_t y0 = t * 256 + 128 * (int32_t)rank;
        for (int kc = 0; kc < p.kchunks; ++kc, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          if (leader) mbar_expect_tx(full_bar(s), 2 * stage_bytes); else mbar_arrive_cluster(full_bar(s), 0);
          const uint32_t dst = ring_base + s * kStage;
          const int32_t gt = kc * p.g_tiles_per_row + (y0 >> 6);      // two adjacent 8 KB tiles: 16 KB contiguous
          tma_load_3d_cg2(dst, &tmG, 0, 0, gt, full_bar(s));
          tma_load_3d_cg2(dst + kBox, &tmG, 0, 0, gt + 1, full_bar(s));
          for (int h = 0; h < p.ndh; ++h) {
            const int32_t d0 = 256 * h + 128 * (int32_t)rank;
            tma_load_2d_cg2(dst + kABytes + (2 * h) * kBox, &tmX16, d0, kc * 64, full_bar(s));
            tma_load_2d_cg2(dst + kABytes + (2 * h + 1) * kBox, &tmX16, d0 + 64, kc * 64, full_bar(s));
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------- MMA issuer: the whole warp waits, one elected lane issues ----------------
      const bool elected = elect_one();
      const uint32_t idesc = make_idesc_f16(false, false, 256, 256, true, true);
      uint32_t it = 0, lt = 0;
      const bool prof = kProfile && (p.dbg & 16) != 0 && elected;
      long long w_full = 0, w_tempty = 0, t_begin = clock64();
      for (int t = pair; t < p.tiles; t += npairs, ++lt) {
        long long t0 = prof ? clock64() : 0;
        mbar_wait(tempty_bar, (lt & 1) ^ 1);
        if (prof) w_tempty += clock64() - t0;
        tc_fence_after();
        for (int kc = 0; kc < p.kchunks; ++kc, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          t0 = prof ? clock64() : 0;
          mbar_wait(full_bar(s), ph);
          if (prof) w_full += clock64() - t0;
          tc_fence_after();
          const uint32_t a_addr = ring_base + s * kStage;
          const uint32_t b_addr = a_addr + kABytes;
          if (elected) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t ad = make_smem_desc_sw128(a_addr + kk * 2048, kBox, 1024);
              for (int h = 0; h < p.ndh; ++h) {
                const uint64_t bd = make_smem_desc_sw128(b_addr + (2 * h) * kBox + kk * 2048, kBox, 1024);
                mma_ss_cg2(tmem_base + 256 * h, ad, bd, idesc, (kc | kk) != 0);
              }
            }
            mma_commit_cg2(empty_bar(s), 3);
            if (kc == p.kchunks - 1) mma_commit_cg2(tfull_bar, 3);
          }
          __syncwarp();
        }
      }
      if (prof && blockIdx.x < 8)
        printf("[gemm_tn mma cta %d] tiles=%u chunks=%u total=%lld clk  wait_full=%lld  wait_tempty=%lld  (per chunk: total %lld full %lld)\n",
               (int)blockIdx.x, lt, it, clock64() - t_begin, w_full, w_tempty, (clock64() - t_begin) / max(it, 1u), w_full / max(it, 1u));
    }
  } else if (warp >= kEpiWarp0) {
    // ---------------- epilogue: TMEM -> swizzled staging tile -> TMA reduce-add into the global accumulator ----------------
    const int ew = warp - kEpiWarp0;
    const int q = warp & 3;                 // TMEM lane quadrant: rows 32q .. 32q+31 of this CTA's 128
    const int half = ew >> 2;               // 256-wide half of D
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t stg = staging_base + (uint32_t)ew * kStageTile;
    uint32_t nstore = 0, lt = 0;
    for (int t = pair; t < p.tiles; t += npairs, ++lt) {
      mbar_wait(tfull_bar, lt & 1);
      tc_fence_after();
      const int32_t row0 = t * 256 + 128 * (int32_t)rank + 32 * q;
      if (half < p.ndh) {
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          uint32_t v[32];
          tmem_ld32(lane_addr + 256 * half + 32 * c, v);
          tmem_ld_wait();
          const uint32_t buf = stg;
          if (nstore >= 1) {                 // the reduce issued from the staging tile one chunk ago has read it
            if (lane == 0) tma_store_wait_read0();
            __syncwarp();
          }
          const uint32_t rowp = buf + lane * 128;
#pragma unroll
          for (int pc = 0; pc < 8; ++pc)
            st_shared_v4(rowp + ((uint32_t)(pc ^ (lane & 7)) << 4), v[4 * pc], v[4 * pc + 1], v[4 * pc + 2], v[4 * pc + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (p.overwrite) tma_store_2d(&tmAcc, buf, 256 * half + 32 * c, row0);
            else tma_reduce_add_2d(&tmAcc, buf, 256 * half + 32 * c, row0);
            tma_store_commit();
          }
          ++nstore;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(tempty_bar); else mbar_arrive_cluster(tempty_bar, 0);
      }
    }
    if (lane == 0) tma_store_wait_all0();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, 512);
  }
}

}  // namespace

size_t gemm_tn_smem_bytes() { return kSmemBytes; }

// G: tile-major f16 scratch [ceil(K / 64)][tiles_per_row][64 k-rows][64 y] as written by tc_block_grad2_kernel<.., true>
// (tiles_per_row * 64 >= ceil(N / 256) * 256; rows >= K of the last tile hold finite values that meet zero-filled X16
// rows), X16: [K, ldx16] f16, acc: [N, ldacc] f32.
int launch_gemm_tn(const void* G, int64_t tiles_per_row, const void* X16, int64_t ldx16, float* acc, int64_t ldacc, int64_t K,
                   int64_t N, int64_t D, int pair_slots, int dbg, bool overwrite, cudaStream_t stream) {
  if (D > 512 || D % 8 != 0 || ldx16 % 8 != 0 || ldacc % 4 != 0 || tiles_per_row * 64 < ceil_div(N, 256) * 256) {
    set_error("gemm_tn: unsupported shape D=%lld tiles_per_row=%lld ldx=%lld", (long long)D, (long long)tiles_per_row, (long long)ldx16);
    return MCLIP_ERR_UNSUPPORTED;
  }
  CUtensorMap tmG, tmX, tmA;
  int rc = tc_make_tmap_tiles(&tmG, G, ceil_div(K, 64) * tiles_per_row);
  if (rc) return rc;
  rc = tc_make_tmap(&tmX, X16, K, D, ldx16, MCLIP_DTYPE_F16, 64);
  if (rc) return rc;
  rc = tc_make_tmap_f32(&tmA, acc, N, D, ldacc, 32, 32);
  if (rc) return rc;
  GemmTnParams p;
  p.dbg = dbg;
  p.overwrite = overwrite ? 1 : 0;
  p.g_tiles_per_row = (int)tiles_per_row;
  p.kchunks = (int)ceil_div(K, 64);
  p.tiles = (int)ceil_div(N, 256);
  p.ndh = (int)ceil_div(D, 256);
  rc = tc_set_smem(reinterpret_cast<const void*>(tc_gemm_tn_kernel), kSmemBytes);
  if (rc) return rc;
  int pairs = p.tiles < pair_slots ? p.tiles : pair_slots;
  if (pairs < 1) pairs = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MCLIP_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_gemm_tn_kernel, tmG, tmX, tmA, p));
  count_launch();
  MCLIP_CUDA_OK(cudaGetLastError());
  return MCLIP_OK;
}

}  // namespace mclip
