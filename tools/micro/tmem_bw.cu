// TMEM read bandwidth: W warps per SM issue tcgen05.ld.32x32b.x{16,32} back to back.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, int warps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < warps) {
    uint32_t r[32];
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        ld32(tm + c * 32, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= r[j];
      }
    }
    t1 = clock64();
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)acc;
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1 << 20] = (float)(t1 - t0);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}
int main() {
  float* out;
  cudaMalloc(&out, (1 << 22) + 64);
  const int iters = 2000;
  for (int warps : {1, 2, 4, 8, 12, 16}) {
    for (int rep = 0; rep < 2; ++rep) {
      k<<<148, 512>>>(out, iters, warps);
      cudaDeviceSynchronize();
    }
    float cyc;
    cudaMemcpy(&cyc, out + (1 << 20), 4, cudaMemcpyDeviceToHost);
    const double bytes = (double)iters * 4 * 32 * 32 * 4 * warps;
    printf("%2d warps: %.0f cycles, %.1f B/clk/SM TMEM read (per warp %.1f B/clk; one x32 load+wait = %.0f cycles)\n", warps, cyc,
           bytes / cyc, bytes / cyc / warps, cyc / (iters * 4));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
