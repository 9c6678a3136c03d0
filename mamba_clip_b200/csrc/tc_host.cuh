// Host-side helpers shared by the tcgen05 translation units (defined in tc_kernels.cu).
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace mclip {

// [rows, D] row-major 16-bit matrix, TMA box = [box_rows x 64 elements], 128-byte swizzle, zero fill.
int tc_make_tmap(CUtensorMap* map, const void* base, int64_t rows, int64_t D, int64_t ld, int dtype, uint32_t box_rows);
// [rows, cols] row-major f32 matrix, TMA box = [box_rows x box_cols] (box_cols * 4 <= 128), 128-byte swizzle.
int tc_make_tmap_f32(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, uint32_t box_cols, uint32_t box_rows);
// tile-major 16-bit scratch [ntiles][64 rows][64 elements] (8 KB per tile, contiguous), box = one tile, NO swizzle: the
// tiles are stored in the byte order they have in shared memory (already SWIZZLE_128B) and are loaded back verbatim.
int tc_make_tmap_tiles(CUtensorMap* map, const void* base, int64_t ntiles);
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device) and size high-water mark.
int tc_set_smem(const void* kernel, uint32_t bytes);

// dY_acc[N, D] (f32) (+)= G[K, N]^T X16[K, D] (tc_gemm_tn.cu): the dY half of the shared-recompute backward.
size_t gemm_tn_smem_bytes();
int launch_gemm_tn(const void* G, int64_t tiles_per_row, const void* X16, int64_t ldx16, float* acc, int64_t ldacc, int64_t K,
                   int64_t N, int64_t D, int pair_slots, int dbg, bool overwrite, cudaStream_t stream);

}  // namespace mclip
