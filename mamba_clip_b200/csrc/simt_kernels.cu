// FFMA (fp32-exact) kernels of the contrastive-loss path, plus the small merge/finalize kernels that
// both paths share.  These serve fp32 inputs (1e-5 parity bar: tensor cores cannot meet it) and the
// shapes the tcgen05 path does not take (D not a multiple of 8, D > 768, unaligned leading dims).
// Same flash-style structure as the tensor-core path: the logits block only ever exists as a
// 64x64 (forward) or 32x32 (backward) register/shared-memory tile.
#include "common.cuh"

namespace mclip {

namespace {

constexpr int kFwdBM = 64, kFwdBN = 64, kBK = 32;
constexpr int kBwdBM = 32, kBwdBN = 32, kBwdDC = 512;

__device__ __forceinline__ float warp16_max(float v) {
#pragma unroll
  for (int o = 8; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp16_sum(float v) {
#pragma unroll
  for (int o = 8; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// forward: partial row (max2, sum) of exp2(log2e * ls * <x_i, y_j>) over one column split
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
simt_row_lse_kernel(const T* __restrict__ X, const T* __restrict__ Y, int64_t M, int64_t N, int64_t D,
                    int64_t ldx, int64_t ldy, const float* __restrict__ ls_ptr, int64_t diag_off,
                    int64_t cols_per_split, float* __restrict__ part_m2, float* __restrict__ part_s,
                    float* __restrict__ part_c, float* __restrict__ diag, const int* __restrict__ run_if) {
  if (run_if != nullptr && *run_if == 0) return;   // predicated call (mclip_row_lse): nothing is read or written
  __shared__ float Xs[kBK][kFwdBM + 1];
  __shared__ float Ys[kBK][kFwdBN + 1];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int64_t row0 = (int64_t)blockIdx.x * kFwdBM;
  const int64_t c_begin = (int64_t)blockIdx.y * cols_per_split;
  const int64_t c_end = min(N, c_begin + cols_per_split);
  const float k2 = ls_ptr[0] * kLog2e;

  float m2[4], sum[4], sc[4];   // running max (log2 units), sum of 2^(x-m), sum of 2^(x-m) * <x_i,y_j>
#pragma unroll
  for (int a = 0; a < 4; ++a) { m2[a] = -INFINITY; sum[a] = 0.f; sc[a] = 0.f; }

  for (int64_t col0 = c_begin; col0 < c_end; col0 += kFwdBN) {
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

    for (int64_t k0 = 0; k0 < D; k0 += kBK) {
#pragma unroll
      for (int i = 0; i < (kFwdBM * kBK) / 256; ++i) {
        const int lin = tid + 256 * i, r = lin >> 5, k = lin & 31;
        const int64_t gr = row0 + r, gc = col0 + r, gk = k0 + k;
        Xs[k][r] = (gr < M && gk < D) ? to_f32<T>(X[gr * ldx + gk]) : 0.f;
        Ys[k][r] = (gc < c_end && gk < D) ? to_f32<T>(Y[gc * ldy + gk]) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kBK; ++k) {
        float xa[4], yb[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) xa[a] = Xs[k][ty * 4 + a];
#pragma unroll
        for (int b = 0; b < 4; ++b) yb[b] = Ys[k][tx * 4 + b];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(xa[a], yb[b], acc[a][b]);
      }
      __syncthreads();
    }

#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int64_t row = row0 + ty * 4 + a;
      float x[4], tmax = -INFINITY;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int64_t col = col0 + tx * 4 + b;
        const bool ok = col < c_end;
        x[b] = ok ? acc[a][b] * k2 : -INFINITY;
        tmax = fmaxf(tmax, x[b]);
        if (ok && diag != nullptr && row < M && col == row + diag_off) diag[row] = acc[a][b];
      }
      tmax = warp16_max(tmax);
      const float m_new = fmaxf(m2[a], tmax);
      if (m_new > -INFINITY) {
        const float rescale = (m2[a] == -INFINITY) ? 0.f : exp2f(m2[a] - m_new);
        float s = sum[a] * rescale, c = sc[a] * rescale;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const float e = exp2f(x[b] - m_new);     // 0 for masked columns (x = -inf)
          s += e;
          c = fmaf(e, acc[a][b], c);
        }
        sum[a] = s;
        sc[a] = c;
        m2[a] = m_new;
      }
    }
  }

#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const float s = warp16_sum(sum[a]);
    const float c = warp16_sum(sc[a]);
    const int64_t row = row0 + ty * 4 + a;
    if (tx == 0 && row < M) {
      part_m2[(int64_t)blockIdx.y * M + row] = m2[a];
      part_s[(int64_t)blockIdx.y * M + row] = s;
      part_c[(int64_t)blockIdx.y * M + row] = c;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// merge per-split partials:  lse = ln2 * (m + log2(sum_k s_k 2^(m_k - m)))
// ---------------------------------------------------------------------------------------------
__global__ void lse_merge_kernel(const float* __restrict__ part_m2, const float* __restrict__ part_s,
                                 const float* __restrict__ part_c, int nsplit, int64_t M, float* __restrict__ lse,
                                 float* __restrict__ rowdot, const int* __restrict__ run_if) {
  if (run_if != nullptr && *run_if == 0) return;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  float m = -INFINITY;
  for (int k = 0; k < nsplit; ++k) m = fmaxf(m, part_m2[(int64_t)k * M + i]);
  float s = 0.f, c = 0.f;
  for (int k = 0; k < nsplit; ++k) {
    const float mk = part_m2[(int64_t)k * M + i];
    if (mk > -INFINITY) {
      const float w = exp2f(mk - m);
      s += part_s[(int64_t)k * M + i] * w;
      if (rowdot != nullptr) c += part_c[(int64_t)k * M + i] * w;
    }
  }
  lse[i] = (m + log2f(s)) * kLn2;
  if (rowdot != nullptr) rowdot[i] = c / s;
}

// ---------------------------------------------------------------------------------------------
// backward: dX = alpha * G @ Y for one 32-row block and one 512-wide slice of D
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
simt_block_grad_kernel(const T* __restrict__ X, const T* __restrict__ Y, int64_t M, int64_t N, int64_t D,
                       int64_t ldx, int64_t ldy, const float* __restrict__ ls_ptr,
                       const float* __restrict__ go_ptr, const float* __restrict__ lse_x,
                       const float* __restrict__ lse_y, int64_t diag_off, float w_row, float w_col,
                       float w_diag, float inv_2n, T* __restrict__ dX, int64_t lddx,
                       float* __restrict__ rowdot) {
  __shared__ float Xs[kBK][kBwdBM + 1];
  __shared__ float Ys[kBK][kBwdBN + 1];
  __shared__ __align__(16) float Gs[kBwdBN][kBwdBM];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int64_t row0 = (int64_t)blockIdx.x * kBwdBM;
  const int64_t d0 = (int64_t)blockIdx.y * kBwdDC;
  const float ls = ls_ptr[0];
  const float alpha = (go_ptr ? go_ptr[0] : 1.f) * ls * inv_2n;

  float acc0[kBwdBM], acc1[kBwdBM];
#pragma unroll
  for (int r = 0; r < kBwdBM; ++r) { acc0[r] = 0.f; acc1[r] = 0.f; }
  float rd[2] = {0.f, 0.f};
  float lx[2];
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int64_t row = row0 + ty * 2 + a;
    lx[a] = row < M ? lse_x[row] : 0.f;
  }
  const int64_t dcol0 = d0 + tid, dcol1 = d0 + tid + 256;

  for (int64_t col0 = 0; col0 < N; col0 += kBwdBN) {
    float c[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    for (int64_t k0 = 0; k0 < D; k0 += kBK) {
#pragma unroll
      for (int i = 0; i < (kBwdBM * kBK) / 256; ++i) {
        const int lin = tid + 256 * i, r = lin >> 5, k = lin & 31;
        const int64_t gr = row0 + r, gc = col0 + r, gk = k0 + k;
        Xs[k][r] = (gr < M && gk < D) ? to_f32<T>(X[gr * ldx + gk]) : 0.f;
        Ys[k][r] = (gc < N && gk < D) ? to_f32<T>(Y[gc * ldy + gk]) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kBK; ++k) {
        const float x0 = Xs[k][ty * 2], x1 = Xs[k][ty * 2 + 1];
        const float y0 = Ys[k][tx * 2], y1 = Ys[k][tx * 2 + 1];
        c[0][0] = fmaf(x0, y0, c[0][0]); c[0][1] = fmaf(x0, y1, c[0][1]);
        c[1][0] = fmaf(x1, y0, c[1][0]); c[1][1] = fmaf(x1, y1, c[1][1]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int64_t row = row0 + ty * 2 + a;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int64_t col = col0 + tx * 2 + b;
        float g = 0.f;
        if (col < N && row < M) {
          const float s = c[a][b] * ls;
          const float p_row = expf(s - lx[a]);
          rd[a] = fmaf(p_row, c[a][b], rd[a]);
          g = w_row * p_row;
          if (w_col != 0.f) g = fmaf(w_col, expf(s - lse_y[col]), g);
          if (col == row + diag_off) g -= w_diag;
        }
        Gs[tx * 2 + b][ty * 2 + a] = g;
      }
    }
    __syncthreads();
    const int jmax = (int)min((int64_t)kBwdBN, N - col0);
    for (int j = 0; j < jmax; ++j) {
      const float y0 = dcol0 < D ? to_f32<T>(Y[(col0 + j) * ldy + dcol0]) : 0.f;
      const float y1 = dcol1 < D ? to_f32<T>(Y[(col0 + j) * ldy + dcol1]) : 0.f;
      const float4* g4 = reinterpret_cast<const float4*>(&Gs[j][0]);
#pragma unroll
      for (int q = 0; q < kBwdBM / 4; ++q) {
        const float4 g = g4[q];
        acc0[4 * q + 0] = fmaf(g.x, y0, acc0[4 * q + 0]); acc1[4 * q + 0] = fmaf(g.x, y1, acc1[4 * q + 0]);
        acc0[4 * q + 1] = fmaf(g.y, y0, acc0[4 * q + 1]); acc1[4 * q + 1] = fmaf(g.y, y1, acc1[4 * q + 1]);
        acc0[4 * q + 2] = fmaf(g.z, y0, acc0[4 * q + 2]); acc1[4 * q + 2] = fmaf(g.z, y1, acc1[4 * q + 2]);
        acc0[4 * q + 3] = fmaf(g.w, y0, acc0[4 * q + 3]); acc1[4 * q + 3] = fmaf(g.w, y1, acc1[4 * q + 3]);
      }
    }
    __syncthreads();
  }

#pragma unroll
  for (int r = 0; r < kBwdBM; ++r) {
    const int64_t row = row0 + r;
    if (row < M) {
      if (dcol0 < D) dX[row * lddx + dcol0] = from_f32<T>(acc0[r] * alpha);
      if (dcol1 < D) dX[row * lddx + dcol1] = from_f32<T>(acc1[r] * alpha);
    }
  }
  if (rowdot != nullptr && blockIdx.y == 0) {
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const float s = warp16_sum(rd[a]);
      const int64_t row = row0 + ty * 2 + a;
      if (tx == 0 && row < M) rowdot[row] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// scalar finalizers (single block, deterministic tree reduction)
// ---------------------------------------------------------------------------------------------
__device__ float block_sum_1024(float v) {
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
  if (warp == 0) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;  // valid in thread 0
}

__global__ void __launch_bounds__(1024)
loss_finalize_kernel(const float* __restrict__ row_lse, const float* __restrict__ col_lse,
                     const float* __restrict__ diag, int64_t n, const float* __restrict__ ls_ptr,
                     float* __restrict__ loss) {
  const float ls = ls_ptr[0];
  float acc = 0.f;
#pragma unroll 8
  for (int64_t i = threadIdx.x; i < n; i += 1024)     // unrolled: the loads of 8 iterations are in flight together
    acc += (row_lse[i] - ls * diag[i]) + (col_lse[i] - ls * diag[i]);
  const float tot = block_sum_1024(acc);
  if (threadIdx.x == 0) loss[0] = tot * (0.5f / (float)n);
}

__global__ void __launch_bounds__(1024)
dls_finalize_kernel(const float* __restrict__ u, const float* __restrict__ v, const float* __restrict__ diag,
                    int64_t n, const float* __restrict__ go_ptr, float scale, float* __restrict__ t_out,
                    float* __restrict__ dls_out) {
  float acc = 0.f;
#pragma unroll 8
  for (int64_t i = threadIdx.x; i < n; i += 1024) {
    const float d = diag ? diag[i] : 0.f;
    acc += (u[i] - d) + ((v ? v[i] : 0.f) - d);
  }
  const float tot = block_sum_1024(acc);
  if (threadIdx.x == 0) {
    t_out[0] = tot;
    dls_out[0] = (go_ptr ? go_ptr[0] : 1.f) * scale * tot;
  }
}

int pick_splits(int64_t row_tiles, int64_t col_tiles) {
  // enough CTAs for ~2 waves of the SMs, never more splits than column tiles
  static const int sms = [] {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) {
      (void)cudaGetLastError();
      n = 148;
    }
    return n;
  }();
  int64_t want = ceil_div(2 * sms, row_tiles);
  if (want < 1) want = 1;
  if (want > col_tiles) want = col_tiles;
  if (want > 64) want = 64;
  return (int)want;
}

template <typename T>
int run_row_lse(const RowLseArgs& a) {
  const int64_t row_tiles = ceil_div(a.M, kFwdBM), col_tiles = ceil_div(a.N, kFwdBN);
  const int nsplit = pick_splits(row_tiles, col_tiles);
  const int64_t cols_per_split = ceil_div(col_tiles, nsplit) * kFwdBN;
  const int real_splits = (int)ceil_div(a.N, cols_per_split);
  float* part_m2 = reinterpret_cast<float*>(a.ws);
  float* part_s = part_m2 + (size_t)real_splits * a.M;
  float* part_c = part_s + (size_t)real_splits * a.M;
  if ((size_t)real_splits * a.M * 3 * sizeof(float) > a.ws_bytes) {
    set_error("row_lse(simt): workspace too small");
    return MCLIP_ERR_WORKSPACE;
  }
  if (a.diag) MCLIP_CUDA_OK(cudaMemsetAsync(a.diag, 0, sizeof(float) * a.M, a.stream));
  dim3 grid((unsigned)row_tiles, (unsigned)real_splits);
  simt_row_lse_kernel<T><<<grid, 256, 0, a.stream>>>(
      reinterpret_cast<const T*>(a.X), reinterpret_cast<const T*>(a.Y), a.M, a.N, a.D, a.ldx, a.ldy,
      a.logit_scale, a.diag_off, cols_per_split, part_m2, part_s, part_c, a.diag, a.run_if);
  count_launch();
  MCLIP_CUDA_OK(cudaGetLastError());
  return launch_lse_merge(part_m2, part_s, part_c, real_splits, a.M, a.lse, a.rowdot, a.stream, a.run_if);
}

template <typename T>
int run_block_grad(const BlockGradArgs& a) {
  dim3 grid((unsigned)ceil_div(a.M, kBwdBM), (unsigned)ceil_div(a.D, kBwdDC));
  simt_block_grad_kernel<T><<<grid, 256, 0, a.stream>>>(
      reinterpret_cast<const T*>(a.X), reinterpret_cast<const T*>(a.Y), a.M, a.N, a.D, a.ldx, a.ldy,
      a.logit_scale, a.grad_out, a.lse_x, a.lse_y, a.diag_off, a.w_row, a.w_col, a.w_diag, a.inv_2n,
      reinterpret_cast<T*>(a.dX), a.lddx, a.rowdot);
  count_launch();
  MCLIP_CUDA_OK(cudaGetLastError());
  return MCLIP_OK;
}

}  // namespace

size_t simt_row_lse_ws(int64_t M, int64_t N, int64_t) {
  const int nsplit = pick_splits(ceil_div(M, kFwdBM), ceil_div(N, kFwdBN));
  return align_up((size_t)nsplit * M * 3 * sizeof(float), 256);
}
size_t simt_block_grad_ws(int64_t, int64_t, int64_t) { return 0; }

int simt_row_lse(const RowLseArgs& a) {
  switch (a.dtype) {
    case MCLIP_DTYPE_F32: return run_row_lse<float>(a);
    case MCLIP_DTYPE_BF16: return run_row_lse<__nv_bfloat16>(a);
    case MCLIP_DTYPE_F16: return run_row_lse<__half>(a);
  }
  set_error("row_lse: bad dtype %d", a.dtype);
  return MCLIP_ERR_INVALID;
}

int simt_block_grad(const BlockGradArgs& a) {
  switch (a.dtype) {
    case MCLIP_DTYPE_F32: return run_block_grad<float>(a);
    case MCLIP_DTYPE_BF16: return run_block_grad<__nv_bfloat16>(a);
    case MCLIP_DTYPE_F16: return run_block_grad<__half>(a);
  }
  set_error("block_grad: bad dtype %d", a.dtype);
  return MCLIP_ERR_INVALID;
}

int launch_lse_merge(const float* part_m2, const float* part_s, const float* part_c, int nsplit, int64_t M, float* lse,
                     float* rowdot, cudaStream_t stream, const int* run_if) {
  lse_merge_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, stream>>>(part_m2, part_s, part_c, nsplit, M, lse, rowdot, run_if);
  count_launch();
  MCLIP_CUDA_OK(cudaGetLastError());
  return MCLIP_OK;
}

int launch_loss_finalize(const float* row_lse, const float* col_lse, const float* diag, int64_t n,
                         const float* logit_scale, float* loss, cudaStream_t stream) {
  loss_finalize_kernel<<<1, 1024, 0, stream>>>(row_lse, col_lse, diag, n, logit_scale, loss);
  count_launch();
  MCLIP_CUDA_OK(cudaGetLastError());
  return MCLIP_OK;
}

int launch_dls_finalize(const float* u, const float* v, const float* diag, int64_t n, const float* grad_out,
                        float scale, float* t_out, float* dls_out, cudaStream_t stream) {
  dls_finalize_kernel<<<1, 1024, 0, stream>>>(u, v, diag, n, grad_out, scale, t_out, dls_out);
  count_launch();
  MCLIP_CUDA_OK(cudaGetLastError());
  return MCLIP_OK;
}

}  // namespace mclip
