// How does the softmax exp2 stream (FADD2, MUFU on 3/4 of the pairs, FMA-pipe polynomial on 1/4, row sum, fp16 pack,
// STS.128) scale with resident warps per scheduler?  Decides whether more softmax groups per SM can pay off.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t pack(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(uint64_t v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t fsub2(uint64_t a, uint64_t b) { uint64_t d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint32_t cvt2(float lo, float hi) { uint32_t y; asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo)); return y; }
__device__ __forceinline__ float fmax3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ uint64_t exp2_poly2(uint64_t x) {
  float a, b; unpack(x, a, b); a = fmaxf(a, -24.f); b = fmaxf(b, -24.f); x = pack(a, b);
  const uint64_t magic = pack(12582912.f, 12582912.f);
  const uint64_t t = fadd2(x, magic);
  const uint64_t f = fsub2(x, fsub2(t, magic));
  uint64_t p = pack(5.516747385e-02f, 5.516747385e-02f);
  p = ffma2(p, f, pack(2.426107377e-01f, 2.426107377e-01f));
  p = ffma2(p, f, pack(6.932617426e-01f, 6.932617426e-01f));
  p = ffma2(p, f, pack(9.999281168e-01f, 9.999281168e-01f));
  float ta, tb, pa, pb; unpack(t, ta, tb); unpack(p, pa, pb);
  pa = __int_as_float(__float_as_int(pa) + (__float_as_int(ta) << 23));
  pb = __int_as_float(__float_as_int(pb) + (__float_as_int(tb) << 23));
  return pack(pa, pb);
}
template <int POLY_MASK>  // pair j uses the polynomial when (j & POLY_MASK) == POLY_MASK (POLY_MASK < 0: never)
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, float seed) {
  extern __shared__ uint8_t sm[];
  const int t = threadIdx.x;
  float s[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) s[j] = -seed * (j + t);
  uint64_t racc[4] = {0, 0, 0, 0};
  float rm[2] = {-1e30f, -1e30f};
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint64_t nm2 = pack(-seed * it, -seed * it);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t ph[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = s[c * 16 + 2 * j], b = s[c * 16 + 2 * j + 1];
        rm[j & 1] = fmax3(rm[j & 1], a, b);
        uint64_t y = fadd2(pack(a, b), nm2);
        if (POLY_MASK >= 0 && (j & POLY_MASK) == POLY_MASK) y = exp2_poly2(y);
        else { float ya, yb; unpack(y, ya, yb); y = pack(ex2(ya), ex2(yb)); }
        racc[j & 3] = fadd2(racc[j & 3], y);
        float ya, yb; unpack(y, ya, yb);
        ph[j] = cvt2(ya, yb);
      }
#pragma unroll
      for (int q = 0; q < 2; ++q)
        *reinterpret_cast<uint4*>(sm + t * 128 + (((2 * c + q) ^ (t & 7)) << 4)) = make_uint4(ph[4 * q], ph[4 * q + 1], ph[4 * q + 2], ph[4 * q + 3]);
    }
  }
  long long t1 = clock64();
  float a, b;
  unpack(fadd2(fadd2(racc[0], racc[1]), fadd2(racc[2], racc[3])), a, b);
  out[blockIdx.x * blockDim.x + t] = a + b + rm[0] + rm[1];
  if (t == 0 && blockIdx.x == 0) out[1 << 20] = (float)(t1 - t0) / iters;
}
int main() {
  float* out; cudaMalloc(&out, (1 << 22) + 64);
  cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(k<-1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int variant = 0; variant < 3; ++variant)
    for (int wps = 1; wps <= 4; ++wps) {
      for (int rep = 0; rep < 2; ++rep) {
        if (variant == 0) k<-1><<<148, 128 * wps, 65536>>>(out, 2000, 0.001f);
        if (variant == 1) k<3><<<148, 128 * wps, 65536>>>(out, 2000, 0.001f);
        if (variant == 2) k<1><<<148, 128 * wps, 65536>>>(out, 2000, 0.001f);
        cudaDeviceSynchronize();
      }
      float cyc; cudaMemcpy(&cyc, out + (1 << 20), 4, cudaMemcpyDeviceToHost);
      printf("poly %s, %d warps/scheduler: %.0f cycles per 64-column item per warp, %.1f columns/clk/SM (MUFU-only ceiling 16)\n",
             variant == 0 ? "0  " : (variant == 1 ? "1/4" : "1/2"), wps, cyc, 64.0 * 128 * wps / cyc);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
