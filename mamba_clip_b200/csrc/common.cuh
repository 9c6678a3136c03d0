// Shared declarations for libmclip_b200 (internal; the public ABI is include/mclip_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/mclip_b200.h"

namespace mclip {

// ---- error plumbing (thread-local message, never abort) ------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch(int n = 1);

#define MCLIP_CUDA_OK(expr)                                              \
  do {                                                                   \
    cudaError_t _e = (expr);                                             \
    if (_e != cudaSuccess) return ::mclip::cuda_fail(_e, #expr);         \
  } while (0)

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Problem descriptors passed from the ABI layer to the path-specific launchers.
struct RowLseArgs {
  const void* X; const void* Y;
  int64_t M, N, D, ldx, ldy;
  int dtype;
  const float* logit_scale;
  int64_t diag_off;
  float* lse;
  float* diag;       // may be null
  float* rowdot;     // may be null
  void* ws; size_t ws_bytes;
  cudaStream_t stream;
  const int* run_if = nullptr;   // device flag: when non-null and *run_if == 0 every launch of the call exits at once
};

// two-sided forward (tc_pair_lse.cu)
struct PairRefArgs {
  const void* X; const void* Y;
  int64_t M, N, D, ldx, ldy;
  int dtype;
  const float* logit_scale;
  int64_t diag_off;
  float* diag;       // [M] raw positive-pair dots
  float* ref;        // [1] uniform exponent reference c0 (log2 units)
  int* status;       // [1] zeroed here; may be null
  void* ws;
  cudaStream_t stream;
};

struct PairLseArgs {
  const void* X; const void* Y;
  int64_t M, N, D, ldx, ldy;
  int dtype;
  const float* logit_scale;
  const float* ref;
  int64_t diag_off;
  float* diag;       // may be null: [M], pre-filled by launch_pair_ref, overwritten with the accumulator's own dots
  float* row_lse;
  float* rowdot;     // may be null
  float* col_out;    // [N] column lse (col_mode 0) or [N + 2] raw column sums relative to ref, ref, status (col_mode 1)
  int col_mode;
  int* status;
  void* ws; size_t ws_bytes;
  cudaStream_t stream;
};

struct BlockGradArgs {
  const void* X; const void* Y;
  int64_t M, N, D, ldx, ldy;
  int dtype;
  const float* logit_scale;
  const float* grad_out;  // may be null (== 1)
  const float* lse_x;
  const float* lse_y;     // may be null iff w_col == 0
  int64_t diag_off;
  float w_row, w_col, w_diag, inv_2n;
  void* dX; int64_t lddx;
  float* rowdot;          // may be null
  void* ws; size_t ws_bytes;
  cudaStream_t stream;
  const void* y16 = nullptr;   // optional f16 copy of Y ([N, D] contiguous) made ahead of time: bf16 inputs skip their own conversion
};

// shared-recompute backward: both feature gradients from ONE recompute of the logits block (tc_kernels.cu)
struct FusedGradArgs {
  const void* X; const void* Y;
  int64_t M, N, D, ldx, ldy;
  int dtype;
  const float* logit_scale;
  const float* grad_out;  // may be null (== 1)
  const float* lse_x;     // [M] row LSEs
  const float* lse_y;     // [N] column LSEs
  int64_t diag_off;
  float inv_2n;
  void* dX; int64_t lddx;
  void* dY; int64_t lddy;
  float* xdot;            // [M] <X_i, (G Y)_i>: sums to logit_scale-free t of mclip_dls_finalize
  void* ws; size_t ws_bytes;
  cudaStream_t stream;
};

// latency path for small global batches (small_kernels.cu): one forward and one backward kernel per step
struct SmallArgs {
  const void* A; const void* B;   // image / text rows in the blocked layout: row g at base + (g / Bl) * blk_stride + (g % Bl) * D
  int64_t Bl, Bg, D, blk_stride;
  int dtype;
  const float* logit_scale;
  int64_t lo, hi;                 // forward: rows whose loss / t terms are summed
  float* stats;                   // [5 * Bg + 2]: row_lse, col_lse, diag, u, v, then loss, t
  // backward only
  int64_t off;                    // first global row of this rank
  const float* grad_out;
  float w_row, w_col, w_diag, inv_2n, dls_scale;
  void* dA; void* dB; float* dls_out;
  void* ws; size_t ws_bytes;
  unsigned* counters;             // zero-initialised by the caller once; every launch leaves it zeroed
  cudaStream_t stream;
};
bool small_supported(int64_t Bl, int64_t Bg, int64_t D);
size_t small_ws_bytes(int64_t Bl, int64_t Bg, int64_t D);
int small_counter_words(int64_t Bl, int64_t Bg);
int small_forward(const SmallArgs& a);
int small_backward(const SmallArgs& a);
int small_pack(const void* a, const void* b, int64_t n, int in_dtype, int out_dtype, void* out, cudaStream_t stream);

// SIMT (FFMA, fp32-exact) path -- simt_kernels.cu
size_t simt_row_lse_ws(int64_t M, int64_t N, int64_t D);
int simt_row_lse(const RowLseArgs& a);
size_t simt_block_grad_ws(int64_t M, int64_t N, int64_t D);
int simt_block_grad(const BlockGradArgs& a);

// tcgen05 / TMEM / TMA path -- tc_kernels.cu
bool tc_supported(int64_t M, int64_t N, int64_t D, int64_t ldx, int64_t ldy, int dtype, int op);
size_t tc_row_lse_ws(int64_t M, int64_t N, int64_t D);
int tc_row_lse(const RowLseArgs& a);
size_t tc_block_grad_ws(int64_t M, int64_t N, int64_t D);
int tc_block_grad(const BlockGradArgs& a);
bool tc_fused_supported(int64_t M, int64_t N, int64_t D, int64_t ldx, int64_t ldy, int dtype);
size_t tc_fused_grad_ws(int64_t M, int64_t N, int64_t D);
int tc_fused_grad(const FusedGradArgs& a);
int launch_convert_f16(const void* src, int64_t rows, int64_t D, int64_t ld, void* dst, cudaStream_t stream);
int tc_set_option(const char* name, int value);
int tc_get_option(const char* name, int* value);

// two-sided forward -- tc_pair_lse.cu
bool tc_pair_supported(int64_t M, int64_t N, int64_t D, int64_t ldx, int64_t ldy, int dtype);
size_t tc_pair_lse_ws(int64_t M, int64_t N, int64_t D);
size_t pair_ref_ws();
int launch_pair_ref(const PairRefArgs& a);
int tc_pair_lse(const PairLseArgs& a);
int launch_lse_from_sum(const float* sum, int64_t n, const float* ref, float* lse, int* status, cudaStream_t stream);
int launch_merge_col_sums(const float* parts, int W, int64_t stride, int64_t n_total, int64_t col0, int64_t n, float* lse,
                          int* status, cudaStream_t stream);

// in-situ timing of the dominant kernel (tc_kernels.cu); used by bench.py's roofline leg only
void kernel_timing_enable(bool on);
int kernel_timing_read(float* total_ms, int* count);

// shared small kernels -- simt_kernels.cu
// merge `nsplit` partial (max2, sum, sum*c) triples per row into a natural-log LSE (+ rowdot if asked for).
int launch_lse_merge(const float* part_m2, const float* part_s, const float* part_c, int nsplit, int64_t M, float* lse,
                     float* rowdot, cudaStream_t stream, const int* run_if = nullptr);
int launch_loss_finalize(const float* row_lse, const float* col_lse, const float* diag, int64_t n,
                         const float* logit_scale, float* loss, cudaStream_t stream);
int launch_dls_finalize(const float* u, const float* v, const float* diag, int64_t n, const float* grad_out,
                        float scale, float* t_out, float* dls_out, cudaStream_t stream);

// producer epilogue (normalise + cast) -- producer_kernels.cu
int launch_normalize_rows(const float* x, int64_t M, int64_t D, int64_t ldx, float eps, int out_dtype, void* y, int64_t ldy,
                          float* inv_norm, cudaStream_t stream);
int launch_normalize_rows_bwd(const float* x, const void* g, int64_t M, int64_t D, int64_t ldx, int64_t ldg, int g_dtype,
                              float eps, float* dx, int64_t lddx, cudaStream_t stream);

// ---- dtype helpers -------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

}  // namespace mclip
