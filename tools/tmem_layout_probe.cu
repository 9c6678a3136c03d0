// Probe of the tcgen05.ld.16x256b.x8 register layout: TMEM is filled through 32x32b stores with value = lane*1000 + col,
// read back with 16x256b.x8, and every thread's 32 registers are dumped.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void probe(uint32_t* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  // each of the 4 warps fills its 32 lanes x 64 columns
  for (int c0 = 0; c0 < 64; c0 += 8) {
    uint32_t v[8];
    for (int j = 0; j < 8; ++j) v[j] = (uint32_t)((warp * 32 + lane) * 1000 + c0 + j);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(base + ((uint32_t)(warp * 32) << 16) + c0),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 1) {   // quadrant 1: lanes 32..63
    for (int halfsel = 0; halfsel < 2; ++halfsel) {
      uint32_t r[32];
      const uint32_t addr = base + ((uint32_t)(32 + 16 * halfsel) << 16);
      asm volatile(
          "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
            "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
            "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(addr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) out[(halfsel * 32 + lane) * 32 + j] = r[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(base) : "memory");
}
int main() {
  uint32_t* d; cudaMalloc(&d, 2 * 32 * 32 * 4);
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  static uint32_t h[2 * 32 * 32];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int hs = 0; hs < 2; ++hs)
    for (int t = 0; t < 32; ++t)
      for (int j = 0; j < 32; ++j) {
        const int rep = j / 4, q = j % 4;
        const int row = 32 + 16 * hs + t / 4 + 8 * (q / 2), col = 8 * rep + 2 * (t % 4) + (q % 2);
        const uint32_t want = row * 1000 + col, got = h[(hs * 32 + t) * 32 + j];
        if (want != got && bad < 10) { printf("mismatch hs=%d t=%d j=%d want lane %d col %d got lane %u col %u\n", hs, t, j, row, col, got / 1000, got % 1000); }
        bad += want != got;
      }
  printf("mismatches: %d (hypothesis: reg j -> rep=j/4, row=t/4+8*((j%%4)/2), col=8*rep+2*(t%%4)+(j%%2))\n", bad);
  for (int j = 0; j < 8; ++j) printf("t=5 reg %d -> lane %u col %u\n", j, h[5 * 32 + j] / 1000, h[5 * 32 + j] % 1000);
  return 0;
}
