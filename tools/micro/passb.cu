// Microbenchmark of the flash pass-B instruction stream with ONE working warp per scheduler:
// what does a lone warp reach, and what do STS / fence.proxy.async / a spinning partner warp cost?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o passb passb.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t pack(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(uint64_t v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint32_t cvt2(float lo, float hi) { uint32_t y; asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo)); return y; }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

__device__ __forceinline__ void chunk(const float (&s)[32], uint64_t nm2, uint64_t (&racc)[4], uint8_t* sP, int t, int k0, bool sts) {
  uint32_t ph[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float x0, x1;
    unpack(fadd2(pack(s[2 * j], s[2 * j + 1]), nm2), x0, x1);
    const float p0 = ex2(x0), p1 = ex2(x1);
    racc[j & 3] = fadd2(racc[j & 3], pack(p0, p1));
    ph[j] = cvt2(p0, p1);
  }
  if (sts) {
    uint8_t* blk = sP + (k0 >> 6) * 16384;
    const int ch0 = (k0 & 63) >> 3;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      *reinterpret_cast<uint4*>(blk + t * 128 + (((ch0 + c) ^ (t & 7)) << 4)) = make_uint4(ph[4 * c], ph[4 * c + 1], ph[4 * c + 2], ph[4 * c + 3]);
  } else {
    uint32_t x = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) x ^= ph[j];
    if (x == 0x12345678u) sP[0] = 1;
  }
}

__device__ __forceinline__ uint64_t fadd2v(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ex2v(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t cvt2v(float lo, float hi) { uint32_t y; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo)); return y; }
__device__ __forceinline__ void sts128v(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// manually software-pipelined order (all volatile): M o M o o, consumers 4 pairs behind their MUFUs
template <int D>
__device__ __forceinline__ void chunk_sched(const float (&s)[32], uint64_t nm2, uint64_t (&racc)[4], uint32_t sP_row, int t, int k0) {
  uint64_t x[16];
  float p[32];
  uint32_t ph[16];
  const uint32_t blk = sP_row + (k0 >> 6) * 16384;
  const int ch0 = (k0 & 63) >> 3;
#pragma unroll
  for (int j = 0; j < D; ++j) x[j] = fadd2v(pack(s[2 * j], s[2 * j + 1]), nm2);
#pragma unroll
  for (int j = 0; j < 16 + D; ++j) {
    float a = 0.f, b = 0.f;
    if (j < 16) {
      unpack(x[j], a, b);
      p[2 * j] = ex2v(a);
      if (j + D < 16) x[j + D] = fadd2v(pack(s[2 * (j + D)], s[2 * (j + D) + 1]), nm2);
      p[2 * j + 1] = ex2v(b);
    }
    if (j >= D) {
      const int i = j - D;
      racc[i & 3] = fadd2v(racc[i & 3], pack(p[2 * i], p[2 * i + 1]));
      ph[i] = cvt2v(p[2 * i], p[2 * i + 1]);
      if ((i & 3) == 3) {
        const int c = i >> 2;
        sts128v(blk + ((((ch0 + c) ^ (t & 7))) << 4), ph[i - 3], ph[i - 2], ph[i - 1], ph[i]);
      }
    }
  }
}

// mode bit0: STS, bit1: fence.proxy.async per half, bit2: spinner warps (warps 4-7 spin on an mbarrier)
__global__ void __launch_bounds__(256, 1) k(float* out, int iters, float seed, int mode) {
  extern __shared__ uint8_t smraw[];
  uint8_t* sP = smraw + 1024;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smraw);
  const int t = threadIdx.x & 127;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(128) : "memory");
  }
  __syncthreads();
  if (threadIdx.x >= 128) {
    if (mode & 4) {
      while (!try_wait(bar, 0)) {}
    }
    return;
  }
  float s[4][32];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int j = 0; j < 32; ++j) s[c][j] = -seed * (j + c * 32 + threadIdx.x);
  uint64_t racc[4] = {0, 0, 0, 0};
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint64_t nm2 = pack(-seed * it, -seed * it);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (mode & 8) chunk_sched<4>(s[c], nm2, racc, smem_u32(sP) + t * 128, t, c * 32);
      else if (mode & 16) chunk_sched<6>(s[c], nm2, racc, smem_u32(sP) + t * 128, t, c * 32);
      else chunk(s[c], nm2, racc, sP, t, c * 32, mode & 1);
      if ((mode & 2) && (c & 1)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
  }
  long long t1 = clock64();
  float a, b;
  unpack(fadd2(fadd2(racc[0], racc[1]), fadd2(racc[2], racc[3])), a, b);
  out[blockIdx.x * 128 + t] = a + b;
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
  if (threadIdx.x == 0 && blockIdx.x == 0) out[148 * 128] = (float)(t1 - t0) / iters;
}

int main() {
  float* out;
  cudaMalloc(&out, 1 << 20);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  for (int mode : {1, 3, 8, 10, 16, 18}) {
    for (int rep = 0; rep < 2; ++rep) {
      k<<<148, 256, 40000>>>(out, 2000, 0.001f, mode);
      cudaDeviceSynchronize();
    }
    float cyc;
    cudaMemcpy(&cyc, out + 148 * 128, 4, cudaMemcpyDeviceToHost);
    printf("mode %d (sts %d fence %d spinner %d): %.0f cycles per 128-column pass (MUFU floor 1024)\n", mode, mode & 1, (mode >> 1) & 1, (mode >> 2) & 1, cyc);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
