// Host-side helpers shared by the tcgen05 translation units (defined in tc_kernels.cu).
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace mclip {

// [rows, D] row-major 16-bit matrix, TMA box = [box_rows x 64 elements], 128-byte swizzle, zero fill.
int tc_make_tmap(CUtensorMap* map, const void* base, int64_t rows, int64_t D, int64_t ld, int dtype, uint32_t box_rows);
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device) and size high-water mark.
int tc_set_smem(const void* kernel, uint32_t bytes);

}  // namespace mclip
