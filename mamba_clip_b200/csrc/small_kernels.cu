// Latency path for small global batches (B_g <= 1024, D <= 512): ONE forward kernel and ONE backward kernel per step.
//
// This is BASELINE.json's C1 / C2 regime (README batch 64; 8 ranks x 64 rows, local_loss=False): 0.8 GFLOP per step, where
// the general path's ~20 dependent launches (reference exponent, tensor-core tiles, per-split merges, finalizers, predicated
// fallbacks) cost far more than the arithmetic.  Here every rank evaluates the WHOLE B_g x B_g problem redundantly, exactly
// as the reference does for local_loss=False (loss.py:104-108), so nothing but the feature gather crosses ranks:
//   small_fwd_kernel   S tiles (64 x 64, fp32 FFMA) for both directions -> per-tile (max, sum, sum*c) partials; the last
//                      CTA of a row tile merges them into LSE / softmax-weighted dots; the last CTA overall reduces the loss
//                      and t = sum_ij G_ij C_ij over the requested row range.  (loss.py:102-111, 142-145)
//   small_bwd_kernel   for the rank's own rows, both directions: S recompute -> G = w_row P^row + w_col P^col - w_diag E in
//                      shared memory -> dX += G Y; column splits meet in f32 partials that the last CTA of a row tile sums
//                      in fixed order (deterministic) and writes in the input dtype; d(logit_scale) = go * scale * t.
// fp32 arithmetic throughout, so fp32 inputs keep the 1e-5 bar and bf16/f16 inputs are exact products.
// Rows are addressed through a blocked layout (row g lives at base + (g / Bl) * blk_stride + (g % Bl) * D) so that the
// kernels read the all-gather receive buffer [W][image shard; text shard] directly.
#include "common.cuh"
#include "tc_host.cuh"

namespace mclip {

namespace {

constexpr int kT = 64;          // forward tile (rows and columns)
constexpr int kBK = 32;         // forward k chunk
constexpr int kBR = 32;         // backward rows per CTA
constexpr int kBC = 64;         // backward columns per chunk

struct SmallFwdParams {
  const void* A; const void* B;   // image rows, text rows
  int Bg, D, Bl;
  int64_t blk_stride;             // elements between consecutive rank blocks
  int lo, hi;                     // rows whose loss / t terms are summed
  const float* ls;
  float* stats;                   // [5][Bg] row_lse, col_lse, diag, u, v; then loss, t
  float* part;                    // [2][ct][Bg][3]
  unsigned* counters;             // [2 * rt + 1]: zero on entry, zero again on exit
};

struct SmallBwdParams {
  const void* A; const void* B;
  int Bg, D, Bl;
  int64_t blk_stride;
  int off;                        // first global row of this rank
  const float* ls; const float* go;
  const float* stats;
  float w_row, w_col, w_diag, inv_2n, dls_scale;
  void* dA; void* dB;             // [Bl, D] in the input dtype
  float* dls_out;                 // may be null
  float* part;                    // [2][rtiles][cs][kBR * D]
  unsigned* counters;             // [2 * rtiles]
  int cs, cols_per_split;
};

template <typename T>
__device__ __forceinline__ const T* row_ptr(const void* base, int g, int Bl, int64_t blk_stride, int D) {
  return reinterpret_cast<const T*>(base) + (int64_t)(g / Bl) * blk_stride + (int64_t)(g % Bl) * D;
}

// 8 consecutive elements -> f32 (D % 8 == 0 on this path)
template <typename T> __device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(h[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
}
template <> __device__ __forceinline__ void load8<__half>(const __half* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(h[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
}

__device__ __forceinline__ float half_warp_max(float v) {
#pragma unroll
  for (int o = 8; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int o = 8; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float block_sum_256(float v, float* red) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
  }
  return t;   // valid on thread 0
}

// grid (ct, rt, 2): blockIdx.z = 0: rows of A against columns of B (row statistics of S); 1: rows of B against A (column statistics)
template <typename T>
__global__ void __launch_bounds__(256)
small_fwd_kernel(const SmallFwdParams p) {
  // k-major tiles with a 16-byte aligned pitch: a thread's 4 rows / 4 columns at one k are ONE ld.shared.v4, and the 16
  // lanes that share rows read the same address (broadcast): ~3 shared-memory wavefronts per k and warp instead of 8+
  __shared__ __align__(16) float Xs[kBK][kT + 4];
  __shared__ __align__(16) float Ys[kBK][kT + 4];
  __shared__ float red[8];
  __shared__ unsigned flag;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int side = blockIdx.z;
  const void* X = side == 0 ? p.A : p.B;
  const void* Y = side == 0 ? p.B : p.A;
  const int row0 = blockIdx.y * kT, col0 = blockIdx.x * kT;
  const int ct = gridDim.x, rt = gridDim.y;
  const float ls = p.ls[0];
  const float k2 = ls * kLog2e;

  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  // staging: thread (r = tid / 4, kq = tid % 4) moves 8 consecutive k of row r of both tiles (k-major in shared memory)
  const int sr = tid >> 2, skq = tid & 3;
  const bool xr_ok = row0 + sr < p.Bg, yr_ok = col0 + sr < p.Bg;
  const T* xrow = row_ptr<T>(X, xr_ok ? row0 + sr : 0, p.Bl, p.blk_stride, p.D);
  const T* yrow = row_ptr<T>(Y, yr_ok ? col0 + sr : 0, p.Bl, p.blk_stride, p.D);
  // register prefetch: the global loads of chunk k0 + 32 are in flight while chunk k0 is multiplied
  float xv[8], yv[8];
  auto fetch = [&](int k0) {
    const int gk = k0 + skq * 8;
    if (xr_ok && gk < p.D) load8<T>(xrow + gk, xv); else { for (int e = 0; e < 8; ++e) xv[e] = 0.f; }
    if (yr_ok && gk < p.D) load8<T>(yrow + gk, yv); else { for (int e = 0; e < 8; ++e) yv[e] = 0.f; }
  };
  fetch(0);
  for (int k0 = 0; k0 < p.D; k0 += kBK) {
#pragma unroll
    for (int e = 0; e < 8; ++e) { Xs[skq * 8 + e][sr] = xv[e]; Ys[skq * 8 + e][sr] = yv[e]; }
    __syncthreads();
    if (k0 + kBK < p.D) fetch(k0 + kBK);
#pragma unroll
    for (int k = 0; k < kBK; ++k) {
      const float4 xq = *reinterpret_cast<const float4*>(&Xs[k][ty * 4]);
      const float4 yq = *reinterpret_cast<const float4*>(&Ys[k][tx * 4]);
      const float xa[4] = {xq.x, xq.y, xq.z, xq.w}, yb[4] = {yq.x, yq.y, yq.z, yq.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(xa[a], yb[b], acc[a][b]);
    }
    __syncthreads();
  }

  // per-row partial over this tile's 64 columns
  float* part = p.part + ((size_t)(side * ct + blockIdx.x) * p.Bg) * 3;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int row = row0 + ty * 4 + a;
    float x[4], tmax = -INFINITY;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int col = col0 + tx * 4 + b;
      const bool ok = col < p.Bg;
      x[b] = ok ? acc[a][b] * k2 : -INFINITY;
      tmax = fmaxf(tmax, x[b]);
      if (side == 0 && ok && row < p.Bg && col == row) p.stats[2 * p.Bg + row] = acc[a][b];   // positive-pair dot
    }
    tmax = half_warp_max(tmax);
    float s = 0.f, c = 0.f;
    if (tmax > -INFINITY) {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const float e = exp2f(x[b] - tmax);      // 0 for masked columns
        s += e;
        c = fmaf(e, acc[a][b], c);
      }
    }
    s = half_warp_sum(s);
    c = half_warp_sum(c);
    if (tx == 0 && row < p.Bg) { part[row * 3] = tmax; part[row * 3 + 1] = s; part[row * 3 + 2] = c; }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) flag = atomicAdd(&p.counters[side * rt + blockIdx.y], 1u) == (unsigned)(ct - 1);
  __syncthreads();
  if (!flag) return;

  // ---- last CTA of this (side, row tile): merge the ct partials of its 64 rows ----
  __threadfence();
  if (tid < kT) {
    const int row = row0 + tid;
    if (row < p.Bg) {
      float m = -INFINITY;
      for (int k = 0; k < ct; ++k) m = fmaxf(m, __ldcg(p.part + ((size_t)(side * ct + k) * p.Bg + row) * 3));
      float s = 0.f, c = 0.f;
      for (int k = 0; k < ct; ++k) {
        const float* q = p.part + ((size_t)(side * ct + k) * p.Bg + row) * 3;
        const float mk = __ldcg(q);
        if (mk > -INFINITY) { const float w = exp2f(mk - m); s = fmaf(__ldcg(q + 1), w, s); c = fmaf(__ldcg(q + 2), w, c); }
      }
      p.stats[side * p.Bg + row] = kLn2 * (m + log2f(s));
      p.stats[(3 + side) * p.Bg + row] = c / s;
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    p.counters[side * rt + blockIdx.y] = 0;
    flag = atomicAdd(&p.counters[2 * rt], 1u) == (unsigned)(2 * rt - 1);
  }
  __syncthreads();
  if (!flag) return;

  // ---- last CTA overall: loss and t over rows [lo, hi) ----
  __threadfence();
  float la = 0.f, ta = 0.f;
  for (int i = p.lo + tid; i < p.hi; i += 256) {
    const float rl = __ldcg(p.stats + i), cl = __ldcg(p.stats + p.Bg + i), d = __ldcg(p.stats + 2 * p.Bg + i);
    const float u = __ldcg(p.stats + 3 * p.Bg + i), v = __ldcg(p.stats + 4 * p.Bg + i);
    la += (rl - ls * d) + (cl - ls * d);
    ta += (u - d) + (v - d);
  }
  const float lsum = block_sum_256(la, red);
  __syncthreads();
  const float tsum = block_sum_256(ta, red);
  if (tid == 0) {
    p.stats[5 * p.Bg] = lsum / (2.f * (float)(p.hi - p.lo));
    p.stats[5 * p.Bg + 1] = tsum;
    p.counters[2 * rt] = 0;
  }
}

// grid (cs, rtiles, 2): blockIdx.z = 0: dA of the rank's image rows (columns = all text rows); 1: dB of its text rows
template <typename T>
__global__ void __launch_bounds__(256)
small_bwd_kernel(const SmallBwdParams p) {
  extern __shared__ float sm[];
  const int pitch = p.D + 4;
  float* Xs = sm;                               // [kBR][pitch]
  float* Ys = Xs + kBR * pitch;                 // [kBC][pitch]
  float* Gs = Ys + kBC * pitch;                 // [kBR][kBC + 1]
  __shared__ unsigned flag;
  const int tid = threadIdx.x;
  const int side = blockIdx.z;
  const void* X = side == 0 ? p.A : p.B;
  const void* Y = side == 0 ? p.B : p.A;
  const float* lse_x = p.stats + (side == 0 ? 0 : p.Bg);
  const float* lse_y = p.stats + (side == 0 ? p.Bg : 0);
  const int r0 = blockIdx.y * kBR;                              // local row
  const int c_begin = blockIdx.x * p.cols_per_split;
  const int c_end = min(p.Bg, c_begin + p.cols_per_split);
  const float ls = p.ls[0];
  const int d8 = p.D / 8;

  if (blockIdx.x == 0 && blockIdx.y == 0 && side == 0 && tid == 0 && p.dls_out != nullptr)
    p.dls_out[0] = (p.go ? p.go[0] : 1.f) * p.dls_scale * p.stats[5 * p.Bg + 1];

  // stage the CTA's X rows as f32
#pragma unroll 4
  for (int idx = tid; idx < kBR * d8; idx += 256) {
    const int r = idx / d8, k = (idx - r * d8) * 8;
    float v[8];
    if (r0 + r < p.Bl) load8<T>(row_ptr<T>(X, p.off + r0 + r, p.Bl, p.blk_stride, p.D) + k, v);
    else { for (int e = 0; e < 8; ++e) v[e] = 0.f; }
    *reinterpret_cast<float4*>(Xs + r * pitch + k) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(Xs + r * pitch + k + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }

  // dX accumulators: thread = 4 rows (rg) x 16 d (4 float4 at d = dg*4 + 128 q)
  const int rg = tid >> 5, dg = tid & 31;
  float acc[4][16];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[a][e] = 0.f;
  // S tile threads (first 128): 4 rows (sy) x 4 columns (sx + 16 b)
  const int sy = tid >> 4, sx = tid & 15;

  for (int c0 = c_begin; c0 < c_end; c0 += kBC) {
    __syncthreads();                              // previous chunk's Ys / Gs fully consumed (and Xs staged)
#pragma unroll 4
    for (int idx = tid; idx < kBC * d8; idx += 256) {
      const int r = idx / d8, k = (idx - r * d8) * 8;
      float v[8];
      if (c0 + r < c_end) load8<T>(row_ptr<T>(Y, c0 + r, p.Bl, p.blk_stride, p.D) + k, v);
      else { for (int e = 0; e < 8; ++e) v[e] = 0.f; }
      *reinterpret_cast<float4*>(Ys + r * pitch + k) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(Ys + r * pitch + k + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    __syncthreads();
    if (tid < 128) {
      float s[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) s[a][b] = 0.f;
      for (int k = 0; k < p.D; k += 4) {
        float4 xa[4], yb[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) xa[a] = *reinterpret_cast<const float4*>(Xs + (sy * 4 + a) * pitch + k);
#pragma unroll
        for (int b = 0; b < 4; ++b) yb[b] = *reinterpret_cast<const float4*>(Ys + (sx + 16 * b) * pitch + k);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            s[a][b] = fmaf(xa[a].x, yb[b].x, s[a][b]);
            s[a][b] = fmaf(xa[a].y, yb[b].y, s[a][b]);
            s[a][b] = fmaf(xa[a].z, yb[b].z, s[a][b]);
            s[a][b] = fmaf(xa[a].w, yb[b].w, s[a][b]);
          }
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int lr = r0 + sy * 4 + a;                           // local row
        const float lx = lr < p.Bl ? lse_x[p.off + lr] : 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int col = c0 + sx + 16 * b;
          float g = 0.f;
          if (lr < p.Bl && col < c_end) {
            const float sv = ls * s[a][b];
            g = p.w_row * __expf(sv - lx);
            if (p.w_col != 0.f) g += p.w_col * __expf(sv - lse_y[col]);
            if (col == p.off + lr) g -= p.w_diag;
          }
          Gs[(sy * 4 + a) * (kBC + 1) + sx + 16 * b] = g;
        }
      }
    }
    __syncthreads();
    const int ncol = min(kBC, c_end - c0);
    for (int j = 0; j < ncol; ++j) {
      float g[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) g[a] = Gs[(rg * 4 + a) * (kBC + 1) + j];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int d = dg * 4 + 128 * q;
        if (d < p.D) {
          const float4 y = *reinterpret_cast<const float4*>(Ys + j * pitch + d);
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            acc[a][4 * q] = fmaf(g[a], y.x, acc[a][4 * q]);
            acc[a][4 * q + 1] = fmaf(g[a], y.y, acc[a][4 * q + 1]);
            acc[a][4 * q + 2] = fmaf(g[a], y.z, acc[a][4 * q + 2]);
            acc[a][4 * q + 3] = fmaf(g[a], y.w, acc[a][4 * q + 3]);
          }
        }
      }
    }
  }

  const float alpha = (p.go ? p.go[0] : 1.f) * ls * p.inv_2n;
  T* out = reinterpret_cast<T*>(side == 0 ? p.dA : p.dB);
  const int rtiles = gridDim.y, cs = gridDim.x;
  if (cs == 1) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int lr = r0 + rg * 4 + a;
      if (lr >= p.Bl) continue;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int d = dg * 4 + 128 * q;
        if (d < p.D) {
#pragma unroll
          for (int e = 0; e < 4; ++e) out[(int64_t)lr * p.D + d + e] = from_f32<T>(acc[a][4 * q + e] * alpha);
        }
      }
    }
    return;
  }
  // column splits: f32 partials, summed in split order by the last CTA of this (side, row tile)
  float* mine = p.part + ((size_t)(side * rtiles + blockIdx.y) * cs + blockIdx.x) * (kBR * p.D);
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int d = dg * 4 + 128 * q;
      if (d < p.D)
        *reinterpret_cast<float4*>(mine + (rg * 4 + a) * p.D + d) =
            make_float4(acc[a][4 * q], acc[a][4 * q + 1], acc[a][4 * q + 2], acc[a][4 * q + 3]);
    }
  __threadfence();
  __syncthreads();
  if (tid == 0) flag = atomicAdd(&p.counters[side * rtiles + blockIdx.y], 1u) == (unsigned)(cs - 1);
  __syncthreads();
  if (!flag) return;
  __threadfence();
  const float* base = p.part + ((size_t)(side * rtiles + blockIdx.y) * cs) * (kBR * p.D);
  const int d4 = p.D / 4;
  for (int idx = tid; idx < kBR * d4; idx += 256) {
    const int r = idx / d4, d = (idx - r * d4) * 4;
    const int lr = r0 + r;
    if (lr >= p.Bl) continue;
    float4 t = __ldcg(reinterpret_cast<const float4*>(base + r * p.D + d));
    for (int k = 1; k < cs; ++k) {
      const float4 u = __ldcg(reinterpret_cast<const float4*>(base + (size_t)k * (kBR * p.D) + r * p.D + d));
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    T* o = out + (int64_t)lr * p.D + d;
    o[0] = from_f32<T>(t.x * alpha); o[1] = from_f32<T>(t.y * alpha);
    o[2] = from_f32<T>(t.z * alpha); o[3] = from_f32<T>(t.w * alpha);
  }
  if (tid == 0) p.counters[side * rtiles + blockIdx.y] = 0;
}

// [image shard; text shard] -> one contiguous send buffer of the compute dtype (the all-gather's input)
template <typename TI, typename TO>
__global__ void small_pack_kernel(const TI* __restrict__ a, const TI* __restrict__ b, int64_t n, TO* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = from_f32<TO>(to_f32<TI>(i < n ? a[i] : b[i - n]));
}

int bwd_splits(int Bg, int rtiles) {
  const int chunks = (Bg + kBC - 1) / kBC;
  int cs = 160 / (2 * rtiles);
  if (cs < 1) cs = 1;
  if (cs > chunks) cs = chunks;
  return cs;
}

}  // namespace

bool small_supported(int64_t Bl, int64_t Bg, int64_t D) {
  return Bg >= 1 && Bg <= 1024 && Bl >= 1 && Bl <= Bg && Bg % Bl == 0 && D >= 8 && D <= 512 && D % 8 == 0 && Bl * Bg <= 128 * 1024;
}

size_t small_ws_bytes(int64_t Bl, int64_t Bg, int64_t D) {
  const int ct = (int)ceil_div(Bg, kT);
  const int rtiles = (int)ceil_div(Bl, kBR);
  const size_t fwd = (size_t)2 * ct * Bg * 3 * sizeof(float);
  const size_t bwd = (size_t)2 * rtiles * bwd_splits((int)Bg, rtiles) * kBR * D * sizeof(float);
  return align_up(fwd > bwd ? fwd : bwd, 256);
}

int small_counter_words(int64_t Bl, int64_t Bg) {
  const int rt = (int)ceil_div(Bg, kT), rtiles = (int)ceil_div(Bl, kBR);
  const int a = 2 * rt + 1, b = 2 * rtiles;
  return a > b ? a : b;
}

template <typename T>
static int small_forward_t(const SmallArgs& a) {
  SmallFwdParams p;
  p.A = a.A; p.B = a.B; p.Bg = (int)a.Bg; p.D = (int)a.D; p.Bl = (int)a.Bl; p.blk_stride = a.blk_stride;
  p.lo = (int)a.lo; p.hi = (int)a.hi; p.ls = a.logit_scale; p.stats = a.stats; p.part = reinterpret_cast<float*>(a.ws);
  p.counters = a.counters;
  const dim3 grid((unsigned)ceil_div(a.Bg, kT), (unsigned)ceil_div(a.Bg, kT), 2);
  small_fwd_kernel<T><<<grid, 256, 0, a.stream>>>(p);
  count_launch();
  MCLIP_CUDA_OK(cudaGetLastError());
  return MCLIP_OK;
}

int small_forward(const SmallArgs& a) {
  if (a.dtype == MCLIP_DTYPE_F32) return small_forward_t<float>(a);
  if (a.dtype == MCLIP_DTYPE_BF16) return small_forward_t<__nv_bfloat16>(a);
  return small_forward_t<__half>(a);
}

template <typename T>
static int small_backward_t(const SmallArgs& a) {
  SmallBwdParams p;
  p.A = a.A; p.B = a.B; p.Bg = (int)a.Bg; p.D = (int)a.D; p.Bl = (int)a.Bl; p.blk_stride = a.blk_stride; p.off = (int)a.off;
  p.ls = a.logit_scale; p.go = a.grad_out; p.stats = a.stats; p.w_row = a.w_row; p.w_col = a.w_col; p.w_diag = a.w_diag;
  p.inv_2n = a.inv_2n; p.dls_scale = a.dls_scale; p.dA = a.dA; p.dB = a.dB; p.dls_out = a.dls_out;
  p.part = reinterpret_cast<float*>(a.ws); p.counters = a.counters;
  const int rtiles = (int)ceil_div(a.Bl, kBR);
  p.cs = bwd_splits((int)a.Bg, rtiles);
  p.cols_per_split = (int)(ceil_div(ceil_div(a.Bg, kBC), p.cs) * kBC);
  p.cs = (int)ceil_div(a.Bg, p.cols_per_split);
  const size_t smem = ((size_t)(kBR + kBC) * (a.D + 4) + kBR * (kBC + 1)) * sizeof(float);
  int rc = tc_set_smem(reinterpret_cast<const void*>(small_bwd_kernel<T>), (uint32_t)smem);   // once per (kernel, device)
  if (rc) return rc;
  const dim3 grid((unsigned)p.cs, (unsigned)rtiles, 2);
  small_bwd_kernel<T><<<grid, 256, smem, a.stream>>>(p);
  count_launch();
  MCLIP_CUDA_OK(cudaGetLastError());
  return MCLIP_OK;
}

int small_backward(const SmallArgs& a) {
  if (a.dtype == MCLIP_DTYPE_F32) return small_backward_t<float>(a);
  if (a.dtype == MCLIP_DTYPE_BF16) return small_backward_t<__nv_bfloat16>(a);
  return small_backward_t<__half>(a);
}

template <typename TI>
static int small_pack_t(const void* a, const void* b, int64_t n, int out_dtype, void* out, cudaStream_t stream) {
  const unsigned blocks = (unsigned)(ceil_div(2 * n, 256) < 1184 ? ceil_div(2 * n, 256) : 1184);
  const TI* pa = reinterpret_cast<const TI*>(a);
  const TI* pb = reinterpret_cast<const TI*>(b);
  if (out_dtype == MCLIP_DTYPE_F32) small_pack_kernel<TI, float><<<blocks, 256, 0, stream>>>(pa, pb, n, reinterpret_cast<float*>(out));
  else if (out_dtype == MCLIP_DTYPE_BF16) small_pack_kernel<TI, __nv_bfloat16><<<blocks, 256, 0, stream>>>(pa, pb, n, reinterpret_cast<__nv_bfloat16*>(out));
  else small_pack_kernel<TI, __half><<<blocks, 256, 0, stream>>>(pa, pb, n, reinterpret_cast<__half*>(out));
  count_launch();
  MCLIP_CUDA_OK(cudaGetLastError());
  return MCLIP_OK;
}

int small_pack(const void* a, const void* b, int64_t n, int in_dtype, int out_dtype, void* out, cudaStream_t stream) {
  if (in_dtype == MCLIP_DTYPE_F32) return small_pack_t<float>(a, b, n, out_dtype, out, stream);
  if (in_dtype == MCLIP_DTYPE_BF16) return small_pack_t<__nv_bfloat16>(a, b, n, out_dtype, out, stream);
  return small_pack_t<__half>(a, b, n, out_dtype, out, stream);
}

}  // namespace mclip
