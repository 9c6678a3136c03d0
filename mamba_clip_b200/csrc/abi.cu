// extern "C" boundary of libmclip_b200.so (see include/mclip_b200.h for the contract).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace mclip {

static thread_local char g_err[512] = {0};
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return MCLIP_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static bool valid_dtype(int d) { return d == MCLIP_DTYPE_F32 || d == MCLIP_DTYPE_BF16 || d == MCLIP_DTYPE_F16; }

static int check_common(const void* X, const void* Y, int64_t M, int64_t N, int64_t D, int64_t ldx,
                        int64_t ldy, int dtype, const float* ls, int path, const char* op) {
  if (!X || !Y || !ls) { set_error("%s: null pointer (X=%p Y=%p logit_scale=%p)", op, X, Y, (const void*)ls); return MCLIP_ERR_INVALID; }
  if (M <= 0 || N <= 0 || D <= 0) { set_error("%s: empty problem M=%lld N=%lld D=%lld", op, (long long)M, (long long)N, (long long)D); return MCLIP_ERR_INVALID; }
  if (M > (1ll << 30) || N > (1ll << 30) || D > (1ll << 20)) { set_error("%s: problem too large", op); return MCLIP_ERR_INVALID; }
  if (ldx < D || ldy < D) { set_error("%s: leading dimension smaller than D (ldx=%lld ldy=%lld D=%lld)", op, (long long)ldx, (long long)ldy, (long long)D); return MCLIP_ERR_INVALID; }
  if (!valid_dtype(dtype)) { set_error("%s: bad dtype %d", op, dtype); return MCLIP_ERR_INVALID; }
  if (path != MCLIP_PATH_AUTO && path != MCLIP_PATH_SIMT && path != MCLIP_PATH_TCGEN05) { set_error("%s: bad path %d", op, path); return MCLIP_ERR_INVALID; }
  return MCLIP_OK;
}

static int resolve_path(int64_t M, int64_t N, int64_t D, int64_t ldx, int64_t ldy, int dtype, int op, int path,
                        const char* name, int* out) {
  const bool tc = tc_supported(M, N, D, ldx, ldy, dtype, op);
  if (path == MCLIP_PATH_TCGEN05 && !tc) {
    set_error("%s: tcgen05 path cannot run this problem (needs bf16/f16, D %% 8 == 0, D <= 768, ld %% 8 == 0)", name);
    return MCLIP_ERR_UNSUPPORTED;
  }
  *out = (path == MCLIP_PATH_AUTO) ? (tc ? MCLIP_PATH_TCGEN05 : MCLIP_PATH_SIMT) : path;
  return MCLIP_OK;
}

}  // namespace mclip

using namespace mclip;

extern "C" {

int mclip_abi_version(void) { return MCLIP_ABI_VERSION; }

const char* mclip_last_error(void) { return g_err; }

int64_t mclip_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int mclip_device_supported(int device, int* sm) {
  cudaDeviceProp p;
  MCLIP_CUDA_OK(cudaGetDeviceProperties(&p, device));
  if (sm) *sm = p.major * 10 + p.minor;
  if (p.major != 10) {
    set_error("device %d is sm_%d%d; libmclip_b200 is built for sm_100a only", device, p.major, p.minor);
    return MCLIP_ERR_UNSUPPORTED;
  }
  return MCLIP_OK;
}

int mclip_select_path(int64_t M, int64_t N, int64_t D, int64_t ldx, int64_t ldy, int dtype, int op) {
  return tc_supported(M, N, D, ldx, ldy, dtype, op) ? MCLIP_PATH_TCGEN05 : MCLIP_PATH_SIMT;
}

int mclip_workspace_bytes(int64_t M, int64_t N, int64_t D, int dtype, int op, int path, size_t* bytes) {
  if (!bytes || M <= 0 || N <= 0 || D <= 0 || !valid_dtype(dtype)) { set_error("workspace_bytes: invalid argument"); return MCLIP_ERR_INVALID; }
  // AUTO callers may not know leading dims yet: size for the larger of the two paths.
  size_t simt = 0, tc = 0;
  if (op == MCLIP_OP_ROW_LSE) { simt = simt_row_lse_ws(M, N, D); tc = tc_row_lse_ws(M, N, D); }
  else if (op == MCLIP_OP_BLOCK_GRAD) { simt = simt_block_grad_ws(M, N, D); tc = tc_block_grad_ws(M, N, D); }
  else if (op == MCLIP_OP_PAIR_LSE) { *bytes = tc_pair_lse_ws(M, N, D); return MCLIP_OK; }
  else if (op == MCLIP_OP_PAIR_REF) { *bytes = pair_ref_ws(); return MCLIP_OK; }
  else if (op == MCLIP_OP_FUSED_GRAD) { *bytes = tc_fused_grad_ws(M, N, D); return MCLIP_OK; }
  else if (op == MCLIP_OP_SMALL) { *bytes = small_ws_bytes(M, N, D); return MCLIP_OK; }
  else { set_error("workspace_bytes: bad op %d", op); return MCLIP_ERR_INVALID; }
  if (path == MCLIP_PATH_SIMT) *bytes = simt;
  else if (path == MCLIP_PATH_TCGEN05) *bytes = tc;
  else *bytes = simt > tc ? simt : tc;
  return MCLIP_OK;
}

int mclip_row_lse(const void* X, const void* Y, int64_t M, int64_t N, int64_t D, int64_t ldx, int64_t ldy,
                  int dtype, const float* logit_scale, int64_t diag_off, float* lse, float* diag, float* rowdot,
                  const int* run_if, void* ws, size_t ws_bytes, int path, void* cuda_stream) {
  int rc = check_common(X, Y, M, N, D, ldx, ldy, dtype, logit_scale, path, "row_lse");
  if (rc) return rc;
  if (!lse) { set_error("row_lse: lse is null"); return MCLIP_ERR_INVALID; }
  int p;
  rc = resolve_path(M, N, D, ldx, ldy, dtype, MCLIP_OP_ROW_LSE, path, "row_lse", &p);
  if (rc) return rc;
  const size_t need = (p == MCLIP_PATH_TCGEN05) ? tc_row_lse_ws(M, N, D) : simt_row_lse_ws(M, N, D);
  if (need > 0 && (!ws || ws_bytes < need)) { set_error("row_lse: workspace %zu < %zu bytes", ws_bytes, need); return MCLIP_ERR_WORKSPACE; }
  if (run_if && diag) { set_error("row_lse: a predicated call (run_if) cannot write diag"); return MCLIP_ERR_INVALID; }
  RowLseArgs a{X, Y, M, N, D, ldx, ldy, dtype, logit_scale, diag_off, lse, diag, rowdot, ws, ws_bytes, (cudaStream_t)cuda_stream, run_if};
  return (p == MCLIP_PATH_TCGEN05) ? tc_row_lse(a) : simt_row_lse(a);
}

int mclip_pair_supported(int64_t M, int64_t N, int64_t D, int64_t ldx, int64_t ldy, int dtype) {
  return (M > 0 && N > 0 && D > 0 && tc_pair_supported(M, N, D, ldx, ldy, dtype)) ? 1 : 0;
}

int mclip_pair_ref(const void* X, const void* Y, int64_t M, int64_t N, int64_t D, int64_t ldx, int64_t ldy, int dtype,
                   const float* logit_scale, int64_t diag_off, float* diag, float* ref, int* status, void* ws,
                   size_t ws_bytes, void* cuda_stream) {
  int rc = check_common(X, Y, M, N, D, ldx, ldy, dtype, logit_scale, MCLIP_PATH_AUTO, "pair_ref");
  if (rc) return rc;
  if (!diag || !ref) { set_error("pair_ref: null diag/ref"); return MCLIP_ERR_INVALID; }
  if (!ws || ws_bytes < pair_ref_ws()) { set_error("pair_ref: workspace %zu < %zu bytes", ws_bytes, pair_ref_ws()); return MCLIP_ERR_WORKSPACE; }
  PairRefArgs a{X, Y, M, N, D, ldx, ldy, dtype, logit_scale, diag_off, diag, ref, status, ws, (cudaStream_t)cuda_stream};
  return launch_pair_ref(a);
}

int mclip_pair_lse(const void* X, const void* Y, int64_t M, int64_t N, int64_t D, int64_t ldx, int64_t ldy, int dtype,
                   const float* logit_scale, const float* ref, int64_t diag_off, float* diag, float* row_lse, float* rowdot,
                   float* col_out, int col_mode, int* status, void* ws, size_t ws_bytes, void* cuda_stream) {
  int rc = check_common(X, Y, M, N, D, ldx, ldy, dtype, logit_scale, MCLIP_PATH_TCGEN05, "pair_lse");
  if (rc) return rc;
  if (!ref || !row_lse || !col_out || !status) { set_error("pair_lse: null ref/row_lse/col_out/status"); return MCLIP_ERR_INVALID; }
  if (col_mode != 0 && col_mode != 1) { set_error("pair_lse: bad col_mode %d", col_mode); return MCLIP_ERR_INVALID; }
  if (!tc_pair_supported(M, N, D, ldx, ldy, dtype)) {
    set_error("pair_lse: needs bf16/f16, D %% 8 == 0, D <= 768, ld %% 8 == 0");
    return MCLIP_ERR_UNSUPPORTED;
  }
  const size_t need = tc_pair_lse_ws(M, N, D);
  if (!ws || ws_bytes < need) { set_error("pair_lse: workspace %zu < %zu bytes", ws_bytes, need); return MCLIP_ERR_WORKSPACE; }
  PairLseArgs a{X, Y, M, N, D, ldx, ldy, dtype, logit_scale, ref, diag_off, diag, row_lse, rowdot, col_out, col_mode, status, ws, ws_bytes,
                (cudaStream_t)cuda_stream};
  return tc_pair_lse(a);
}

int mclip_merge_col_sums(const float* parts, int W, int64_t stride, int64_t n_total, int64_t col0, int64_t n, float* lse,
                         int* status, void* cuda_stream) {
  if (!parts || !lse || !status || W <= 0 || n <= 0 || col0 < 0 || col0 + n > n_total || stride < n_total + 2) {
    set_error("merge_col_sums: invalid argument");
    return MCLIP_ERR_INVALID;
  }
  return launch_merge_col_sums(parts, W, stride, n_total, col0, n, lse, status, (cudaStream_t)cuda_stream);
}

int mclip_lse_from_sum(const float* sum, int64_t n, const float* ref, float* lse, int* status, void* cuda_stream) {
  if (!sum || !ref || !lse || !status || n <= 0) { set_error("lse_from_sum: invalid argument"); return MCLIP_ERR_INVALID; }
  return launch_lse_from_sum(sum, n, ref, lse, status, (cudaStream_t)cuda_stream);
}

int mclip_block_grad(const void* X, const void* Y, int64_t M, int64_t N, int64_t D, int64_t ldx, int64_t ldy,
                     int dtype, const float* logit_scale, const float* grad_out, const float* lse_x,
                     const float* lse_y, int64_t diag_off, float w_row, float w_col, float w_diag, float inv_2n,
                     void* dX, int64_t lddx, float* rowdot, const void* Y16, void* ws, size_t ws_bytes, int path,
                     void* cuda_stream) {
  int rc = check_common(X, Y, M, N, D, ldx, ldy, dtype, logit_scale, path, "block_grad");
  if (rc) return rc;
  if (!lse_x || !dX) { set_error("block_grad: null lse_x/dX"); return MCLIP_ERR_INVALID; }
  if (w_col != 0.f && !lse_y) { set_error("block_grad: w_col != 0 needs lse_y"); return MCLIP_ERR_INVALID; }
  if (lddx < D) { set_error("block_grad: lddx < D"); return MCLIP_ERR_INVALID; }
  int p;
  // dX shares X's alignment requirements on the tcgen05 path
  rc = resolve_path(M, N, D, ldx | lddx, ldy, dtype, MCLIP_OP_BLOCK_GRAD, path, "block_grad", &p);
  if (rc) return rc;
  const size_t need = (p == MCLIP_PATH_TCGEN05) ? tc_block_grad_ws(M, N, D) : simt_block_grad_ws(M, N, D);
  if (need > 0 && (!ws || ws_bytes < need)) { set_error("block_grad: workspace %zu < %zu bytes", ws_bytes, need); return MCLIP_ERR_WORKSPACE; }
  if (Y16 && (((uintptr_t)Y16) & 15)) { set_error("block_grad: Y16 must be 16-byte aligned"); return MCLIP_ERR_INVALID; }
  BlockGradArgs a{X, Y, M, N, D, ldx, ldy, dtype, logit_scale, grad_out, lse_x, lse_y, diag_off,
                  w_row, w_col, w_diag, inv_2n, dX, lddx, rowdot, ws, ws_bytes, (cudaStream_t)cuda_stream};
  a.y16 = (dtype == MCLIP_DTYPE_BF16) ? Y16 : nullptr;
  return (p == MCLIP_PATH_TCGEN05) ? tc_block_grad(a) : simt_block_grad(a);
}

int mclip_loss_finalize(const float* row_lse, const float* col_lse, const float* diag, int64_t n,
                        const float* logit_scale, float* loss, void* cuda_stream) {
  if (!row_lse || !col_lse || !diag || !logit_scale || !loss || n <= 0) { set_error("loss_finalize: invalid argument"); return MCLIP_ERR_INVALID; }
  return launch_loss_finalize(row_lse, col_lse, diag, n, logit_scale, loss, (cudaStream_t)cuda_stream);
}

int mclip_fused_grad_supported(int64_t M, int64_t N, int64_t D, int64_t ldx, int64_t ldy, int dtype) {
  return (M > 0 && N > 0 && D > 0 && tc_fused_supported(M, N, D, ldx, ldy, dtype)) ? 1 : 0;
}

int mclip_fused_grad(const void* X, const void* Y, int64_t M, int64_t N, int64_t D, int64_t ldx, int64_t ldy, int dtype,
                     const float* logit_scale, const float* grad_out, const float* lse_x, const float* lse_y,
                     int64_t diag_off, float inv_2n, void* dX, int64_t lddx, void* dY, int64_t lddy, float* xdot, void* ws,
                     size_t ws_bytes, void* cuda_stream) {
  int rc = check_common(X, Y, M, N, D, ldx, ldy, dtype, logit_scale, MCLIP_PATH_TCGEN05, "fused_grad");
  if (rc) return rc;
  if (!lse_x || !lse_y || !dX || !dY || !xdot) { set_error("fused_grad: null lse_x/lse_y/dX/dY/xdot"); return MCLIP_ERR_INVALID; }
  if (lddx < D || lddy < D) { set_error("fused_grad: lddx/lddy < D"); return MCLIP_ERR_INVALID; }
  if (!tc_supported(M, N, D, ldx | lddx, ldy | lddy, dtype, MCLIP_OP_BLOCK_GRAD) || D > 512) {
    set_error("fused_grad: needs bf16/f16, D %% 8 == 0, D <= 512, leading dimensions %% 8 == 0");
    return MCLIP_ERR_UNSUPPORTED;
  }
  const size_t need = tc_fused_grad_ws(M, N, D);
  if (!ws || ws_bytes < need) { set_error("fused_grad: workspace %zu < %zu bytes", ws_bytes, need); return MCLIP_ERR_WORKSPACE; }
  FusedGradArgs a{X, Y, M, N, D, ldx, ldy, dtype, logit_scale, grad_out, lse_x, lse_y, diag_off, inv_2n, dX, lddx, dY, lddy, xdot,
                  ws, ws_bytes, (cudaStream_t)cuda_stream};
  return tc_fused_grad(a);
}

int mclip_small_supported(int64_t Bl, int64_t Bg, int64_t D, int dtype) {
  return (valid_dtype(dtype) && small_supported(Bl, Bg, D)) ? 1 : 0;
}

int mclip_small_counter_words(int64_t Bl, int64_t Bg) { return small_counter_words(Bl, Bg); }

static int check_small(const void* A, const void* B, int64_t Bl, int64_t Bg, int64_t D, int64_t blk_stride, int dtype,
                       const float* ls, const float* stats, const void* ws, size_t ws_bytes, const unsigned* counters, const char* op) {
  if (!A || !B || !ls || !stats || !counters) { set_error("%s: null pointer", op); return MCLIP_ERR_INVALID; }
  if (!valid_dtype(dtype) || !small_supported(Bl, Bg, D)) { set_error("%s: unsupported problem Bl=%lld Bg=%lld D=%lld", op, (long long)Bl, (long long)Bg, (long long)D); return MCLIP_ERR_UNSUPPORTED; }
  if (Bg > Bl && blk_stride < Bl * D) { set_error("%s: blk_stride smaller than one shard", op); return MCLIP_ERR_INVALID; }
  if ((((uintptr_t)A) | ((uintptr_t)B)) & 15) { set_error("%s: A/B must be 16-byte aligned", op); return MCLIP_ERR_INVALID; }
  const size_t need = small_ws_bytes(Bl, Bg, D);
  if (!ws || ws_bytes < need) { set_error("%s: workspace %zu < %zu bytes", op, ws_bytes, need); return MCLIP_ERR_WORKSPACE; }
  return MCLIP_OK;
}

int mclip_small_forward(const void* A, const void* B, int64_t Bl, int64_t Bg, int64_t D, int64_t blk_stride, int dtype,
                        const float* logit_scale, int64_t lo, int64_t hi, float* stats, void* ws, size_t ws_bytes,
                        unsigned* counters, void* cuda_stream) {
  int rc = check_small(A, B, Bl, Bg, D, blk_stride, dtype, logit_scale, stats, ws, ws_bytes, counters, "small_forward");
  if (rc) return rc;
  if (lo < 0 || hi > Bg || lo >= hi) { set_error("small_forward: bad row range [%lld, %lld)", (long long)lo, (long long)hi); return MCLIP_ERR_INVALID; }
  SmallArgs a{};
  a.A = A; a.B = B; a.Bl = Bl; a.Bg = Bg; a.D = D; a.blk_stride = blk_stride; a.dtype = dtype; a.logit_scale = logit_scale;
  a.lo = lo; a.hi = hi; a.stats = stats; a.ws = ws; a.ws_bytes = ws_bytes; a.counters = counters; a.stream = (cudaStream_t)cuda_stream;
  return small_forward(a);
}

int mclip_small_backward(const void* A, const void* B, int64_t Bl, int64_t Bg, int64_t D, int64_t blk_stride, int dtype,
                         const float* logit_scale, const float* grad_out, const float* stats, int64_t off, float w_row,
                         float w_col, float w_diag, float inv_2n, float dls_scale, void* dA, void* dB, float* dls_out, void* ws,
                         size_t ws_bytes, unsigned* counters, void* cuda_stream) {
  int rc = check_small(A, B, Bl, Bg, D, blk_stride, dtype, logit_scale, stats, ws, ws_bytes, counters, "small_backward");
  if (rc) return rc;
  if (!dA || !dB || off < 0 || off + Bl > Bg) { set_error("small_backward: null outputs or bad row offset"); return MCLIP_ERR_INVALID; }
  SmallArgs a{};
  a.A = A; a.B = B; a.Bl = Bl; a.Bg = Bg; a.D = D; a.blk_stride = blk_stride; a.dtype = dtype; a.logit_scale = logit_scale;
  a.stats = const_cast<float*>(stats); a.off = off; a.grad_out = grad_out; a.w_row = w_row; a.w_col = w_col; a.w_diag = w_diag;
  a.inv_2n = inv_2n; a.dls_scale = dls_scale; a.dA = dA; a.dB = dB; a.dls_out = dls_out; a.ws = ws; a.ws_bytes = ws_bytes;
  a.counters = counters; a.stream = (cudaStream_t)cuda_stream;
  return small_backward(a);
}

int mclip_small_pack(const void* a, const void* b, int64_t n, int in_dtype, int out_dtype, void* out, void* cuda_stream) {
  if (!a || !b || !out || n <= 0 || !valid_dtype(in_dtype) || !valid_dtype(out_dtype)) { set_error("small_pack: invalid argument"); return MCLIP_ERR_INVALID; }
  return small_pack(a, b, n, in_dtype, out_dtype, out, (cudaStream_t)cuda_stream);
}

int mclip_convert_f16(const void* src, int64_t rows, int64_t D, int64_t ld, void* dst, void* cuda_stream) {
  if (!src || !dst || rows <= 0 || D <= 0 || D % 8 != 0 || ld < D || ld % 8 != 0 || ((((uintptr_t)src) | ((uintptr_t)dst)) & 15)) {
    set_error("convert_f16: needs 16-byte aligned pointers, D %% 8 == 0, ld %% 8 == 0");
    return MCLIP_ERR_INVALID;
  }
  return launch_convert_f16(src, rows, D, ld, dst, (cudaStream_t)cuda_stream);
}

int mclip_set_option(const char* name, int value) { return tc_set_option(name, value); }
int mclip_get_option(const char* name, int* value) { return tc_get_option(name, value); }

int mclip_dls_finalize(const float* u, const float* v, const float* diag, int64_t n, const float* grad_out,
                       float scale, float* t_out, float* dls_out, void* cuda_stream) {
  if (!u || !t_out || !dls_out || n <= 0) { set_error("dls_finalize: invalid argument"); return MCLIP_ERR_INVALID; }
  return launch_dls_finalize(u, v, diag, n, grad_out, scale, t_out, dls_out, (cudaStream_t)cuda_stream);
}

int mclip_kernel_timing(int enable, float* total_ms, int* count) {
  if (enable == 1) { kernel_timing_enable(true); return MCLIP_OK; }
  if (enable == 0) kernel_timing_enable(false);
  return kernel_timing_read(total_ms, count);
}

int mclip_normalize_rows(const float* x, int64_t M, int64_t D, int64_t ldx, float eps, int out_dtype, void* y, int64_t ldy,
                         float* inv_norm, void* cuda_stream) {
  if (!x || !y || M <= 0 || D <= 0 || ldx < D || ldy < D || !valid_dtype(out_dtype) || !(eps > 0.f)) { set_error("normalize_rows: invalid argument"); return MCLIP_ERR_INVALID; }
  return launch_normalize_rows(x, M, D, ldx, eps, out_dtype, y, ldy, inv_norm, (cudaStream_t)cuda_stream);
}

int mclip_normalize_rows_bwd(const float* x, const void* g, int64_t M, int64_t D, int64_t ldx, int64_t ldg, int g_dtype,
                             float eps, float* dx, int64_t lddx, void* cuda_stream) {
  if (!x || !g || !dx || M <= 0 || D <= 0 || ldx < D || ldg < D || lddx < D || !valid_dtype(g_dtype) || !(eps > 0.f)) { set_error("normalize_rows_bwd: invalid argument"); return MCLIP_ERR_INVALID; }
  return launch_normalize_rows_bwd(x, g, M, D, ldx, ldg, g_dtype, eps, dx, lddx, (cudaStream_t)cuda_stream);
}

}  // extern "C"
