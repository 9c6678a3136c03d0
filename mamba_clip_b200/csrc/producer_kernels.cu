// Producer epilogue of the loss: reference model.py:1011-1017 applies
// F.normalize(features, dim=-1) to the towers' fp32 projections and autocast then rounds them to bf16/f16 in
// front of the logits matmul.  These two HBM-bound kernels do normalise + cast in one pass over the rows (and
// the matching backward), so the loss kernels read 16-bit unit-norm features that were written exactly once.
#include "common.cuh"

namespace mclip {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T> __device__ __forceinline__ void store_out(T* p, float v);
template <> __device__ __forceinline__ void store_out<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void store_out<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ void store_out<__half>(__half* p, float v) { *p = __float2half_rn(v); }

// one warp per row; y = x / max(||x||, eps)
template <typename T>
__global__ void __launch_bounds__(256)
normalize_rows_kernel(const float* __restrict__ x, int64_t M, int64_t D, int64_t ldx, float eps, T* __restrict__ y,
                      int64_t ldy, float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float* xr = x + row * ldx;
  float ss = 0.f;
  for (int64_t d = lane; d < D; d += 32) { const float v = xr[d]; ss = fmaf(v, v, ss); }
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), eps);
  if (lane == 0 && inv_norm != nullptr) inv_norm[row] = inv;
  T* yr = y + row * ldy;
  for (int64_t d = lane; d < D; d += 32) store_out<T>(yr + d, xr[d] * inv);
}

// dx = inv * (g - n <n, g>) with n = x * inv   (inv * g when the norm was clamped to eps)
template <typename T>
__global__ void __launch_bounds__(256)
normalize_rows_bwd_kernel(const float* __restrict__ x, const T* __restrict__ g, int64_t M, int64_t D, int64_t ldx,
                          int64_t ldg, float eps, float* __restrict__ dx, int64_t lddx) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float* xr = x + row * ldx;
  const T* gr = g + row * ldg;
  float ss = 0.f, xg = 0.f;
  for (int64_t d = lane; d < D; d += 32) {
    const float v = xr[d], gg = to_f32<T>(gr[d]);
    ss = fmaf(v, v, ss);
    xg = fmaf(v, gg, xg);
  }
  ss = warp_sum(ss);
  xg = warp_sum(xg);
  const float nrm = sqrtf(ss);
  const bool clamped = nrm <= eps;
  const float inv = 1.f / fmaxf(nrm, eps);
  const float proj = clamped ? 0.f : xg * inv * inv;      // <n, g> / ||x|| * ... folded: dx = inv*g - x * (<x,g> inv^3)
  float* dr = dx + row * lddx;
  for (int64_t d = lane; d < D; d += 32) dr[d] = inv * (to_f32<T>(gr[d]) - xr[d] * proj);
}

}  // namespace

int launch_normalize_rows(const float* x, int64_t M, int64_t D, int64_t ldx, float eps, int out_dtype, void* y, int64_t ldy,
                          float* inv_norm, cudaStream_t stream) {
  const unsigned blocks = (unsigned)ceil_div(M, 8);
  switch (out_dtype) {
    case MCLIP_DTYPE_F32: normalize_rows_kernel<float><<<blocks, 256, 0, stream>>>(x, M, D, ldx, eps, (float*)y, ldy, inv_norm); break;
    case MCLIP_DTYPE_BF16: normalize_rows_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(x, M, D, ldx, eps, (__nv_bfloat16*)y, ldy, inv_norm); break;
    case MCLIP_DTYPE_F16: normalize_rows_kernel<__half><<<blocks, 256, 0, stream>>>(x, M, D, ldx, eps, (__half*)y, ldy, inv_norm); break;
    default: set_error("normalize_rows: bad dtype %d", out_dtype); return MCLIP_ERR_INVALID;
  }
  count_launch();
  MCLIP_CUDA_OK(cudaGetLastError());
  return MCLIP_OK;
}

int launch_normalize_rows_bwd(const float* x, const void* g, int64_t M, int64_t D, int64_t ldx, int64_t ldg, int g_dtype,
                              float eps, float* dx, int64_t lddx, cudaStream_t stream) {
  const unsigned blocks = (unsigned)ceil_div(M, 8);
  switch (g_dtype) {
    case MCLIP_DTYPE_F32: normalize_rows_bwd_kernel<float><<<blocks, 256, 0, stream>>>(x, (const float*)g, M, D, ldx, ldg, eps, dx, lddx); break;
    case MCLIP_DTYPE_BF16: normalize_rows_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(x, (const __nv_bfloat16*)g, M, D, ldx, ldg, eps, dx, lddx); break;
    case MCLIP_DTYPE_F16: normalize_rows_bwd_kernel<__half><<<blocks, 256, 0, stream>>>(x, (const __half*)g, M, D, ldx, ldg, eps, dx, lddx); break;
    default: set_error("normalize_rows_bwd: bad dtype %d", g_dtype); return MCLIP_ERR_INVALID;
  }
  count_launch();
  MCLIP_CUDA_OK(cudaGetLastError());
  return MCLIP_OK;
}

}  // namespace mclip
