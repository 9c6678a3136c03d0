/*
 * mclip_b200.h -- C ABI of libmclip_b200.so: the B200 (sm_100a) contrastive-loss hot path.
 *
 * The reference (psmyth94/mamba-clip) has no FFI: its boundary is the Python class
 * `ClipLoss` (src/mamba_clip/loss.py:56-147).  `mamba_clip_b200.loss.ClipLoss` keeps that class
 * API and binds the entry points below with ctypes.  All pointers are raw
 * DEVICE pointers owned by the caller (PyTorch allocates everything); the library never allocates
 * or frees device memory, never synchronises the host, and launches only on the stream handed in.
 * Every entry point returns 0 on success or an MCLIP_ERR_* code; `mclip_last_error()` returns a
 * thread-local message for the last failure.  Nothing aborts.
 *
 * Arithmetic: inputs are f32, bf16 or f16 row-major matrices; all statistics, the loss and
 * d(logit_scale) are f32; products are accumulated in f32 (tcgen05 tensor cores for bf16/f16
 * inputs, FFMA for f32 inputs).  The B x B logits matrix is never written to memory.
 */
#ifndef MCLIP_B200_H_
#define MCLIP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCLIP_ABI_VERSION 6

enum { MCLIP_DTYPE_F32 = 0, MCLIP_DTYPE_BF16 = 1, MCLIP_DTYPE_F16 = 2 };
enum { MCLIP_PATH_AUTO = 0, MCLIP_PATH_SIMT = 1, MCLIP_PATH_TCGEN05 = 2 };
enum { MCLIP_OP_ROW_LSE = 0, MCLIP_OP_BLOCK_GRAD = 1, MCLIP_OP_PAIR_LSE = 2, MCLIP_OP_PAIR_REF = 3, MCLIP_OP_FUSED_GRAD = 4, MCLIP_OP_SMALL = 5 };
enum {
  MCLIP_OK = 0,
  MCLIP_ERR_INVALID = 1,      /* bad shape / pointer / alignment / enum */
  MCLIP_ERR_CUDA = 2,         /* a CUDA runtime/driver call failed (message has the CUDA error) */
  MCLIP_ERR_UNSUPPORTED = 3,  /* device is not sm_100, or the forced path cannot run this shape */
  MCLIP_ERR_WORKSPACE = 4     /* workspace smaller than mclip_workspace_bytes() */
};

/* ABI version of the loaded library (== MCLIP_ABI_VERSION of the header it was built from). */
int mclip_abi_version(void);

/* Thread-local, NUL-terminated description of the last error on this thread ("" if none). */
const char* mclip_last_error(void);

/* 0 if `device` can run the kernels (compute capability 10.x); writes major*10+minor to *sm. */
int mclip_device_supported(int device, int* sm);

/* Which path MCLIP_PATH_AUTO would take for this problem (MCLIP_PATH_SIMT / MCLIP_PATH_TCGEN05). */
int mclip_select_path(int64_t M, int64_t N, int64_t D, int64_t ldx, int64_t ldy, int dtype, int op);

/* Bytes of scratch `ws` the given op needs for this problem (256-byte aligned pointer expected). */
int mclip_workspace_bytes(int64_t M, int64_t N, int64_t D, int dtype, int op, int path, size_t* bytes);

/*
 * Row log-sum-exp of one logits block, without materialising it.
 *   S = logit_scale * X @ Y^T            X: [M, D] (ldx), Y: [N, D] (ldy), row-major
 *   lse[i]  = log sum_j exp(S[i, j])                          (natural log, f32)
 *   diag[i] = <X[i], Y[i + diag_off]>  (raw dot product, 0 when i + diag_off is outside [0, N))
 *   rowdot[i] = sum_j softmax_j(S[i, :]) * <X[i], Y[j]>       (f32; feeds d logit_scale, see mclip_dls_finalize)
 * Replaces, for one side of the loss, reference loss.py:102-111 (`logit_scale * a @ b.T`) fused with
 * the log_softmax half of F.cross_entropy at loss.py:143-144; `diag_off` is the label offset of
 * loss.py:80-81 (`labels + num_logits * rank`).  One-sided form (the forward of fp32 inputs / D > 512, the
 * predicated fallback behind mclip_pair_lse): called for (image rows, all text) and
 * (text rows, all image).  `logit_scale` is a device scalar (no host sync).  `diag` and `rowdot` may be NULL.
 * `run_if` (device int, may be NULL): when non-NULL and *run_if == 0 at execution time, every kernel of the call
 * exits immediately and no output is written -- this is how the robust one-sided path is chained behind
 * mclip_pair_lse's status flag without a host synchronisation.  A predicated call must pass diag == NULL.
 */
int mclip_row_lse(const void* X, const void* Y, int64_t M, int64_t N, int64_t D, int64_t ldx, int64_t ldy,
                  int dtype, const float* logit_scale, int64_t diag_off, float* lse, float* diag, float* rowdot,
                  const int* run_if, void* ws, size_t ws_bytes, int path, void* cuda_stream);

/*
 * Two-sided forward: one pass over S = logit_scale * X @ Y^T produces BOTH directions of the loss
 * (reference loss.py:102-111 builds logits_per_image and logits_per_text, :142-145 runs cross_entropy on each):
 *   row_lse[i] = log sum_j exp(S[i, j])                     rowdot[i] = sum_j softmax_j(S[i, :]) <X[i], Y[j]>
 *   col_out[j] = log sum_i exp(S[i, j])  (col_mode 0)   or   sum_i 2^(S[i, j] log2(e) - ref[0])  (col_mode 1)
 * All exponentials use the single reference ref[0] (log2 units) produced by mclip_pair_ref, so partial column sums
 * of different row blocks simply add.  In col_mode 1 (row block of one rank, M < all rows) col_out has N + 2 slots:
 * the raw sums, then ref[0], then the status word; all-gather these vectors and finish the rank's own columns with
 * mclip_merge_col_sums (or a single vector with mclip_lse_from_sum).  Validity is checked on the result: if any row / column total leaves
 * [2^-75, 2^120] the call ORs a non-zero bit into *status (device int) and the outputs must be recomputed by
 * mclip_row_lse(..., run_if = status).  tcgen05 path only: mclip_pair_supported() says whether a problem qualifies
 * (bf16/f16, D % 8 == 0, D <= 768, leading dimensions % 8 == 0; for 512 < D the k-chunks past 512 of the row block are streamed).
 * `diag` (may be NULL; otherwise the vector mclip_pair_ref filled, same diag_off) is overwritten with the tensor-core
 * accumulator's own <X[i], Y[i + diag_off]>: with a saturated softmax the loss subtracts logit_scale * diag from an LSE
 * dominated by that same product, so both must carry identical rounding.
 */
int mclip_pair_supported(int64_t M, int64_t N, int64_t D, int64_t ldx, int64_t ldy, int dtype);

/*
 * diag[i] = <X[i], Y[i + diag_off]> (raw dot, 0 when the index is outside [0, N)) -- the positive-pair logits the
 * loss needs anyway (labels of loss.py:76-87) -- and ref[0] = max_i(logit_scale log2(e) diag[i]) - 15, the uniform
 * exponent reference of mclip_pair_lse.  Also zeroes *status (may be NULL).  Workspace: MCLIP_OP_PAIR_REF.
 */
int mclip_pair_ref(const void* X, const void* Y, int64_t M, int64_t N, int64_t D, int64_t ldx, int64_t ldy, int dtype,
                   const float* logit_scale, int64_t diag_off, float* diag, float* ref, int* status, void* ws,
                   size_t ws_bytes, void* cuda_stream);

int mclip_pair_lse(const void* X, const void* Y, int64_t M, int64_t N, int64_t D, int64_t ldx, int64_t ldy, int dtype,
                   const float* logit_scale, const float* ref, int64_t diag_off, float* diag, float* row_lse, float* rowdot,
                   float* col_out, int col_mode, int* status, void* ws, size_t ws_bytes, void* cuda_stream);

/*
 * Column LSEs of columns [col0, col0 + n) from W gathered col_mode-1 vectors (`parts`, W rows of `stride` >= n_total + 2
 * floats, n_total = N of the producing calls): lse[j] = ln2 * (m + log2(sum_q parts[q][col0 + j] 2^(ref_q - m))),
 * m = max_q ref_q.  ORs every producer's status word, and 2 for an out-of-window total, into *status.
 * This is the only cross-rank step of the text->image direction: it replaces the second logits block of
 * loss.py:102-111 at W > 1.
 */
int mclip_merge_col_sums(const float* parts, int W, int64_t stride, int64_t n_total, int64_t col0, int64_t n, float* lse,
                         int* status, void* cuda_stream);

/* lse[i] = ln(2) * (ref[0] + log2(sum[i])); ORs 2 into *status if a sum is outside [2^-75, 2^120]. */
int mclip_lse_from_sum(const float* sum, int64_t n, const float* ref, float* lse, int* status, void* cuda_stream);

/*
 * Gradient of one logits block w.r.t. its row operand, recomputing the block (never stored):
 *   s_ij  = logit_scale * <X[i], Y[j]>
 *   G_ij  = w_row * exp(s_ij - lse_x[i]) + w_col * exp(s_ij - lse_y[j]) - w_diag * [j == i + diag_off]
 *   dX    = (grad_out * logit_scale * inv_2n) * G @ Y          written in the input dtype, [M, D] (lddx)
 *   rowdot[i] = sum_j exp(s_ij - lse_x[i]) * <X[i], Y[j]>      (f32; same quantity mclip_row_lse can emit)
 * Replaces the implicit autograd graph of loss.py:102-111,142-145 (MmBackward + LogSoftmaxBackward +
 * NllLossBackward) for one side.  `lse_y` may be NULL iff w_col == 0 (local_loss without
 * gather_with_grad: own-row terms only).  `grad_out` is a device scalar (GradScaler's scale reaches
 * the loss through it, reference train.py:59-63) or NULL for 1.  `rowdot` may be NULL.
 * `Y16` (may be NULL; only read for bf16 inputs on the tcgen05 path): an f16 copy of Y, [N, D] contiguous, made ahead of
 * time with mclip_convert_f16 -- e.g. on a side stream next to the forward kernel -- so that the call does not spend its
 * own pass over Y on it (G is f16 * 2^12 and the dX MMA needs both operands in f16).
 */
int mclip_block_grad(const void* X, const void* Y, int64_t M, int64_t N, int64_t D, int64_t ldx, int64_t ldy,
                     int dtype, const float* logit_scale, const float* grad_out, const float* lse_x,
                     const float* lse_y, int64_t diag_off, float w_row, float w_col, float w_diag,
                     float inv_2n, void* dX, int64_t lddx, float* rowdot, const void* Y16, void* ws, size_t ws_bytes,
                     int path, void* cuda_stream);

/* dst[rows, D] (f16, contiguous) = src[rows, D] (bf16, leading dimension ld): exact for 6.1e-5 <= |v| <= 65504, saturating
 * above.  D % 8 == 0, 16-byte aligned pointers. */
int mclip_convert_f16(const void* src, int64_t rows, int64_t D, int64_t ld, void* dst, void* cuda_stream);

/*
 * Both feature gradients of one logits block from ONE recompute of it (the backward of reference loss.py:102-111,
 * 142-145 at world_size 1: MmBackward x4 + LogSoftmaxBackward + NllLossBackward, 3 GEMM units instead of 4):
 *   G_ij = exp(s_ij - lse_x[i]) + exp(s_ij - lse_y[j]) - 2 [j == i + diag_off]        s_ij = logit_scale <X[i], Y[j]>
 *   dX   = (grad_out * logit_scale * inv_2n) * G   @ Y        [M, D] (lddx), input dtype
 *   dY   = (grad_out * logit_scale * inv_2n) * G^T @ X        [N, D] (lddy), input dtype
 *   xdot[i] = sum_j G_ij <X[i], Y[j]>   (f32; sum_i xdot[i] is the `t` of mclip_dls_finalize: by Euler's identity
 *             sum_i <X[i], dL/dX[i]> = logit_scale * dL/d logit_scale, so no second statistic pass is needed)
 * G is formed tile by tile in f16 * 2^12 exactly as in mclip_block_grad and handed from the dX pass to the dY pass
 * through a scratch strip of `panel x N` elements inside `ws` (two strips in flight; panel ~ 4.7k rows on a B200):
 * the M x N matrix never exists.  tcgen05 path only: bf16/f16, D % 8 == 0, D <= 512 (mclip_fused_grad_supported says
 * whether a problem qualifies AND is large enough to profit).  Workspace: MCLIP_OP_FUSED_GRAD.  The dY pass runs on a
 * library-owned side stream that is forked from and joined back into `cuda_stream` with events (capturable).
 */
int mclip_fused_grad_supported(int64_t M, int64_t N, int64_t D, int64_t ldx, int64_t ldy, int dtype);
int mclip_fused_grad(const void* X, const void* Y, int64_t M, int64_t N, int64_t D, int64_t ldx, int64_t ldy, int dtype,
                     const float* logit_scale, const float* grad_out, const float* lse_x, const float* lse_y,
                     int64_t diag_off, float inv_2n, void* dX, int64_t lddx, void* dY, int64_t lddy, float* xdot, void* ws,
                     size_t ws_bytes, void* cuda_stream);

/*
 * loss = (1 / (2 n)) * sum_i (row_lse[i] + col_lse[i] - 2 * logit_scale * diag[i])
 * i.e. (CE(logits_per_image) + CE(logits_per_text)) / 2 of loss.py:142-145 for the n local samples.
 */
int mclip_loss_finalize(const float* row_lse, const float* col_lse, const float* diag, int64_t n,
                        const float* logit_scale, float* loss, void* cuda_stream);

/*
 * t   = sum_i (u[i] + v[i] - 2 * diag[i])            (the rank's partial of sum_ij G_ij C_ij * 2n)
 * dls = grad_out * scale * t                         (scale = 1/(2 n_ls); grad_out NULL -> 1)
 * Replaces the d(logit_scale) branch of autograd through `logit_scale * features` (loss.py:102-111).
 * Writes t_out[0] = t and dls_out[0] = dls.  `v` and `diag` may be NULL (taken as 0): with u = xdot of mclip_fused_grad
 * this finishes d(logit_scale) of the shared-recompute backward.
 */
int mclip_dls_finalize(const float* u, const float* v, const float* diag, int64_t n, const float* grad_out,
                       float scale, float* t_out, float* dls_out, void* cuda_stream);

/*
 * Producer epilogue (next row of the scope table): y = cast(x / max(||x||_2, eps)) per row, f32 in, `out_dtype` out.
 * Replaces F.normalize(features, dim=-1) of reference model.py:1011-1017 plus the rounding autocast applies in front
 * of the logits matmul.  `inv_norm` ([M] f32, may be NULL) receives 1 / max(||x||, eps).
 */
int mclip_normalize_rows(const float* x, int64_t M, int64_t D, int64_t ldx, float eps, int out_dtype, void* y, int64_t ldy,
                         float* inv_norm, void* cuda_stream);

/* Backward of mclip_normalize_rows: dx = (g - n <n, g>) / ||x|| with n = x / ||x|| (g / eps where the norm was clamped). */
int mclip_normalize_rows_bwd(const float* x, const void* g, int64_t M, int64_t D, int64_t ldx, int64_t ldg, int g_dtype,
                             float eps, float* dx, int64_t lddx, void* cuda_stream);

/*
 * Measurement hook (bench.py's roofline leg; never used by the loss itself).  enable = 1: from now on every launch of
 * the dominant kernel (the CTA-pair backward kernel behind mclip_block_grad) is bracketed by two CUDA events on its own
 * stream.  enable = 0 / 2: stop (0) or keep going (2), synchronise the recorded events and return their summed duration
 * in *total_ms and their number in *count (either may be NULL); the record list is cleared.
 */
int mclip_kernel_timing(int enable, float* total_ms, int* count);

/*
 * Latency path for small global batches (B_g <= 1024, D <= 512, D % 8 == 0, B_l * B_g <= 128 Ki; any of the three
 * dtypes): the whole step is ONE forward and ONE backward kernel.  Every rank evaluates the full B_g x B_g problem, as the
 * reference does for local_loss=False (loss.py:104-108), so only the feature gather crosses ranks: no statistics
 * exchange, no scalar all-reduce.  fp32 FFMA arithmetic (fp32 inputs keep the 1e-5 bar).
 *   A / B: image / text rows of ALL ranks in a blocked layout -- global row g lives at base + (g / Bl) * blk_stride +
 *          (g % Bl) * D elements -- which is how the all-gather of [image shard; text shard] lands ([W][2][Bl][D]:
 *          A = recv, B = recv + Bl * D, blk_stride = 2 * Bl * D); at world_size 1 A / B are the inputs themselves.
 *   mclip_small_forward : stats[0..Bg) row LSEs, [Bg..2Bg) column LSEs, [2Bg..3Bg) positive-pair dots, [3Bg..4Bg) /
 *          [4Bg..5Bg) softmax-weighted dots u / v, stats[5Bg] = loss = (1/(2n)) sum_{i in [lo,hi)} (row_lse + col_lse
 *          - 2 ls diag)  (loss.py:142-145), stats[5Bg+1] = t = sum_{i in [lo,hi)} (u + v - 2 diag).
 *   mclip_small_backward: for the rank's rows [off, off + Bl): dA = alpha G_A @ B_all, dB = alpha G_B @ A_all with
 *          G = w_row P^row + w_col P^col - w_diag E and alpha = grad_out * logit_scale * inv_2n (the weights / n of the
 *          mode table, as mclip_block_grad), and dls_out[0] = grad_out * dls_scale * t.
 *   mclip_small_pack    : out = [a; b] converted to out_dtype (the all-gather's contiguous send buffer), n elements each.
 * `counters`: mclip_small_counter_words() 32-bit words, zeroed once by the caller; every launch leaves them zeroed
 * (last-CTA-done reductions inside the kernels).  Workspace: MCLIP_OP_SMALL with M = Bl, N = Bg.
 */
int mclip_small_supported(int64_t Bl, int64_t Bg, int64_t D, int dtype);
int mclip_small_counter_words(int64_t Bl, int64_t Bg);
int mclip_small_forward(const void* A, const void* B, int64_t Bl, int64_t Bg, int64_t D, int64_t blk_stride, int dtype,
                        const float* logit_scale, int64_t lo, int64_t hi, float* stats, void* ws, size_t ws_bytes,
                        unsigned* counters, void* cuda_stream);
int mclip_small_backward(const void* A, const void* B, int64_t Bl, int64_t Bg, int64_t D, int64_t blk_stride, int dtype,
                         const float* logit_scale, const float* grad_out, const float* stats, int64_t off, float w_row,
                         float w_col, float w_diag, float inv_2n, float dls_scale, void* dA, void* dB, float* dls_out, void* ws,
                         size_t ws_bytes, unsigned* counters, void* cuda_stream);
int mclip_small_pack(const void* a, const void* b, int64_t n, int in_dtype, int out_dtype, void* out, void* cuda_stream);

/*
 * Development / measurement switches.  Read from the environment once at load time (MCLIP_BWD_PERSIST, MCLIP_FUSED_BWD,
 * MCLIP_DBG), never on the launch path; these calls change them afterwards.  Names: "bwd_persist" (persistent variant
 * of the CTA-pair backward kernel: 1 = always, 0 = never, -1 = chosen per launch shape, the default), "fused_bwd" (0 makes
 * mclip_fused_grad_supported answer 0), "dbg" (profiling builds).
 */
int mclip_set_option(const char* name, int value);
int mclip_get_option(const char* name, int* value);

/* Number of kernel launches issued by this library on the calling thread since load (bench.py's
 * "gpu_launches" counter). */
int64_t mclip_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MCLIP_B200_H_ */
