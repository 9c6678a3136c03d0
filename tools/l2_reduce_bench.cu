// Microbenchmarks behind DESIGN.md section 6 ("can one S recompute feed dI and dT inside one kernel at D = 512?"):
//   (1) cp.reduce.async.bulk .add.f32 from shared memory into a global f32 buffer (L2-resident and not),
//       all SMs at once: the rate at which partial dT tiles could be flushed through L2 reductions;
//   (2) red.global.add.v4.f32 from registers (same question without TMA);
//   (3) st.shared::cluster.v4 pushes into the peer CTA of a cluster of 2 (DSMEM): the rate at which G tiles could be
//       handed to a neighbouring SM.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/l2_reduce_bench tools/l2_reduce_bench.cu
// Run:    tools/l2_reduce_bench            (prints one line per experiment; numbers go to profiles/)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// (1) each CTA owns `smem_bytes` of f32 in shared memory and pushes `iters` x (smem_bytes / chunk) bulk reductions
__global__ void __launch_bounds__(128) tma_reduce_kernel(float* dst, size_t dst_floats, int chunk_bytes, int smem_bytes, int iters) {
  extern __shared__ __align__(128) uint8_t sm[];
  float* s = reinterpret_cast<float*>(sm);
  for (int i = threadIdx.x; i < smem_bytes / 4; i += blockDim.x) s[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nchunk = smem_bytes / chunk_bytes;
    const size_t chunk_floats = chunk_bytes / 4;
    const size_t nslots = dst_floats / chunk_floats;
    size_t slot = ((size_t)blockIdx.x * 2654435761u) % nslots;
    for (int it = 0; it < iters; ++it) {
      for (int c = 0; c < nchunk; ++c) {
        float* g = dst + slot * chunk_floats;
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                     ::"l"(g), "r"(smem_u32(sm + c * chunk_bytes)), "r"(chunk_bytes) : "memory");
        slot += gridDim.x * 7 + 1;          // walk the whole buffer, different CTAs on different lines
        if (slot >= nslots) slot %= nslots;
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // at most 2 groups in flight per CTA
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// (2) red.global.add.v4.f32: every thread adds 16 B per instruction, a warp covers 512 contiguous bytes
__global__ void __launch_bounds__(256) red_v4_kernel(float* dst, size_t dst_floats, int iters) {
  const size_t nvec = dst_floats / 4;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int it = 0; it < iters; ++it) {
    float* g = dst + (idx % nvec) * 4;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(g), "f"(1.0f) : "memory");
    idx += stride;
  }
}

// (3) DSMEM push: each CTA of a cluster of 2 streams `bytes_per_iter` into its peer's shared memory with 16-byte stores
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256) dsmem_push_kernel(int bytes_per_iter, int iters, int* sink) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t peer = rank ^ 1;
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(sm)), "r"(peer));
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  for (int it = 0; it < iters; ++it) {
    for (int off = threadIdx.x * 16; off < bytes_per_iter; off += blockDim.x * 16) {
      asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(remote + off), "r"(it) : "memory");
    }
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x == 0 && sink) sink[blockIdx.x] = reinterpret_cast<int*>(sm)[0];
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main() {
  int dev = 0, sms = 0, clk = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev));
  printf("device SMs=%d max_clock=%d MHz\n", sms, clk / 1000);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const size_t big = (size_t)1 << 30;      // 1 GiB of f32 (beyond L2)
  float* buf;
  CK(cudaMalloc(&buf, big));
  CK(cudaMemset(buf, 0, big));
  int* sink;
  CK(cudaMalloc(&sink, 4096 * sizeof(int)));

  // ---- (1) TMA bulk reduce ----
  const int smem_bytes = 64 * 1024;
  CK(cudaFuncSetAttribute(tma_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const size_t targets[3] = {(size_t)32 << 20, (size_t)64 << 20, (size_t)1 << 30};
  const int chunks[4] = {1024, 4096, 16384, 65536};
  for (int t = 0; t < 3; ++t) {
    for (int c = 0; c < 4; ++c) {
      for (int ctas_per_sm = 1; ctas_per_sm <= 2; ++ctas_per_sm) {
        const int grid = sms * ctas_per_sm;
        const int iters = 200;
        tma_reduce_kernel<<<grid, 128, smem_bytes>>>(buf, targets[t] / 4, chunks[c], smem_bytes, 20);   // warm
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        tma_reduce_kernel<<<grid, 128, smem_bytes>>>(buf, targets[t] / 4, chunks[c], smem_bytes, iters);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        const double bytes = (double)grid * iters * smem_bytes;
        const float ms = time_ms(e0, e1);
        printf("tma_reduce_add_f32  target=%4zu MiB  chunk=%5d B  ctas/SM=%d : %8.1f GB/s  (%.3f ms, %.1f B/clk/SM at %d MHz)\n",
               targets[t] >> 20, chunks[c], ctas_per_sm, bytes / ms * 1e-6, ms, bytes / ms * 1e-3 / sms / (clk / 1000.0) , clk / 1000);
      }
    }
  }
  // ---- (2) red.global.add.v4.f32 ----
  for (int t = 0; t < 3; ++t) {
    const int grid = sms * 8, iters = 2000;
    red_v4_kernel<<<grid, 256>>>(buf, targets[t] / 4, 50);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    red_v4_kernel<<<grid, 256>>>(buf, targets[t] / 4, iters);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    const double bytes = (double)grid * 256 * iters * 16;
    const float ms = time_ms(e0, e1);
    printf("red.global.add.v4.f32  target=%4zu MiB : %8.1f GB/s  (%.3f ms)\n", targets[t] >> 20, bytes / ms * 1e-6, ms);
  }
  // ---- (3) DSMEM push ----
  {
    const int bytes_per_iter = 64 * 1024, iters = 2000;
    CK(cudaFuncSetAttribute(dsmem_push_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes_per_iter));
    const int grid = (sms / 2) * 2;
    dsmem_push_kernel<<<grid, 256, bytes_per_iter>>>(bytes_per_iter, 10, sink);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    dsmem_push_kernel<<<grid, 256, bytes_per_iter>>>(bytes_per_iter, iters, sink);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    const double bytes = (double)grid * iters * bytes_per_iter;
    const float ms = time_ms(e0, e1);
    printf("dsmem st.shared::cluster.v4 push (cluster of 2, 256 thr): %8.1f GB/s total, %.1f GB/s per SM (%.3f ms)\n",
           bytes / ms * 1e-6, bytes / ms * 1e-6 / grid, ms);
  }
  CK(cudaFree(buf));
  CK(cudaFree(sink));
  return 0;
}
