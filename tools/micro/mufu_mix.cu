// Microbenchmark: attainable MUFU.EX2 rate per SM for the flash-softmax instruction mix, as a function of
// warps per scheduler.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_mix mufu_mix.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t pack(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(uint64_t v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint32_t cvt2(float lo, float hi) { uint32_t y; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo)); return y; }

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float s[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) s[j] = seed * (j + threadIdx.x);
  uint64_t racc[4] = {0, 0, 0, 0};
  uint32_t hacc = 0;
  const uint64_t nm2 = pack(-seed, -seed);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (MODE == 0) {  // MUFU only
        s[2 * j] = ex2(s[2 * j]);
        s[2 * j + 1] = ex2(s[2 * j + 1]);
      } else {  // full mix: FADD2, 2 MUFU, FADD2, F2FP
        float x0, x1;
        unpack(fadd2(pack(s[2 * j], s[2 * j + 1]), nm2), x0, x1);
        float p0 = ex2(x0), p1 = ex2(x1);
        racc[j & 3] = fadd2(racc[j & 3], pack(p0, p1));
        hacc ^= cvt2(p0, p1);
        if (MODE == 2) { s[2 * j] = p0; s[2 * j + 1] = p1; }
      }
    }
  }
  long long t1 = clock64();
  float acc = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) acc += s[j];
  float a, b;
  unpack(fadd2(fadd2(racc[0], racc[1]), fadd2(racc[2], racc[3])), a, b);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + a + b + __uint_as_float(hacc);
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}

int main() {
  float* out;
  cudaMalloc(&out, 1 << 24);
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int wps = 1; wps <= 4; ++wps) {  // warps per scheduler
      const int threads = 128 * wps;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, threads>>>(out, iters, 0.001f);
        if (mode == 1) k<1><<<148, threads>>>(out, iters, 0.001f);
        if (mode == 2) k<2><<<148, threads>>>(out, iters, 0.001f);
        cudaDeviceSynchronize();
      }
      float cyc;
      cudaMemcpy(&cyc, out, 4, cudaMemcpyDeviceToHost);
      const double mufu_per_clk_sm = (double)iters * 32 * threads / cyc;
      printf("mode %d warps/sched %d: %.0f cycles, %.2f MUFU/clk/SM (peak 16)\n", mode, wps, cyc, mufu_per_clk_sm);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
