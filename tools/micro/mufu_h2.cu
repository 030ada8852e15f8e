// Microbenchmark: does ex2.approx.f16x2 deliver two exponentials per MUFU issue slot on sm_100a?
// Mode 0: f32 ex2 (reference, 16 / clk / SM).  Mode 1: f16x2 ex2, counted in ELEMENTS per clk per SM.
// Mode 2: f16x2 softmax mix: HSUB2 (s - m), ex2.f16x2, HADD2 row-sum, result kept packed (it IS the P operand).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_h2 mufu_h2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t hsub2(uint32_t a, uint32_t b) { uint32_t y; asm volatile("sub.rn.f16x2 %0, %1, %2;" : "=r"(y) : "r"(a), "r"(b)); return y; }
__device__ __forceinline__ uint32_t hadd2(uint32_t a, uint32_t b) { uint32_t y; asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(y) : "r"(a), "r"(b)); return y; }

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float s[32];
  uint32_t h[16];
#pragma unroll
  for (int j = 0; j < 32; ++j) s[j] = seed * (j + threadIdx.x);
#pragma unroll
  for (int j = 0; j < 16; ++j) { __half2 v = __floats2half2_rn(s[2 * j], s[2 * j + 1]); h[j] = *reinterpret_cast<uint32_t*>(&v); }
  uint32_t racc[4] = {0, 0, 0, 0};
  __half2 mm = __floats2half2_rn(seed, seed);
  const uint32_t m2 = *reinterpret_cast<uint32_t*>(&mm);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (MODE == 0) {
        s[2 * j] = ex2(s[2 * j]);
        s[2 * j + 1] = ex2(s[2 * j + 1]);
      } else if (MODE == 1) {
        h[j] = ex2h2(h[j]);
      } else {
        const uint32_t p = ex2h2(hsub2(h[j], m2));
        racc[j & 3] = hadd2(racc[j & 3], p);
        h[j] = p;
      }
    }
  }
  long long t1 = clock64();
  float acc = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) acc += s[j];
  uint32_t x = racc[0] ^ racc[1] ^ racc[2] ^ racc[3];
#pragma unroll
  for (int j = 0; j < 16; ++j) x ^= h[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __uint_as_float(x);
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}

int main() {
  float* out;
  cudaMalloc(&out, 1 << 24);
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int wps = 1; wps <= 4; ++wps) {
      const int threads = 128 * wps;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, threads>>>(out, iters, 0.001f);
        if (mode == 1) k<1><<<148, threads>>>(out, iters, 0.001f);
        if (mode == 2) k<2><<<148, threads>>>(out, iters, 0.001f);
        cudaDeviceSynchronize();
      }
      float cyc;
      cudaMemcpy(&cyc, out, 4, cudaMemcpyDeviceToHost);
      printf("mode %d warps/sched %d: %.0f cycles, %.2f exp2 ELEMENTS/clk/SM (f32 MUFU peak 16)\n", mode, wps, cyc,
             (double)iters * 32 * threads / cyc);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
