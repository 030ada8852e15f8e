// Math helpers shared by the triangle-attention kernels: packed fp32x2 arithmetic (FADD2 / FFMA2), 3-input max
// (FMNMX3), exp2 on the FMA pipe, TMEM store.
#pragma once
#include "prd_common.cuh"

namespace prd {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kMaskFillLog2 = -32768.0f * kLog2e;  // modules.py:177,220 in the exp2 domain

// ---- packed fp32x2 / 3-input max helpers (sm_100: FADD2, FMNMX3) ----
__device__ __forceinline__ uint64_t pack_f2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack_f2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fsub2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
// exp2 of two non-positive arguments on the FMA pipe (no MUFU): Cody-Waite split x = n + f, |f| <= 1/2,
// degree-3 minimax polynomial for 2^f (max relative error 7.5e-5, well below the fp16 rounding of P that
// follows), n added into the exponent field.  Arguments are clamped at -24 (2^-24 is below fp16 range).
// A quarter of the exponentials of the all-valid path go through here: the MUFU pipe (16 exp2/clk/SM) is the
// nominal bound of this kernel, the FMA pipe has room.
__device__ __forceinline__ uint64_t exp2_poly2(uint64_t x) {
  float a, b;
  unpack_f2(x, a, b);
  a = fmaxf(a, -24.f);
  b = fmaxf(b, -24.f);
  x = pack_f2(a, b);
  const uint64_t magic = pack_f2(12582912.f, 12582912.f);  // 1.5 * 2^23: the sum's low mantissa bits = round(x)
  const uint64_t t = fadd2(x, magic);
  const uint64_t f = fsub2(x, fsub2(t, magic));
  uint64_t p = pack_f2(5.516747385e-02f, 5.516747385e-02f);
  p = ffma2(p, f, pack_f2(2.426107377e-01f, 2.426107377e-01f));
  p = ffma2(p, f, pack_f2(6.932617426e-01f, 6.932617426e-01f));
  p = ffma2(p, f, pack_f2(9.999281168e-01f, 9.999281168e-01f));
  float ta, tb, pa, pb;
  unpack_f2(t, ta, tb);
  unpack_f2(p, pa, pb);
  pa = __int_as_float(__float_as_int(pa) + (__float_as_int(ta) << 23));
  pb = __int_as_float(__float_as_int(pb) + (__float_as_int(tb) << 23));
  return pack_f2(pa, pb);
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ uint32_t cvt_f16x2(float lo, float hi) {
  uint32_t y;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo));
  return y;
}
__device__ __forceinline__ void tmem_ld_wait32(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
                 "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
                 "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// Per-key softmax terms in the exp2 domain: t_j = s_j * mul_j + add_j
//   valid key   : mul = 1, add = 0
//   masked key  : mul = 0, add = -2^15 * log2(e)   (the reference's finite fill value)
//   j >= N (pad): mul = 0, add = -inf               (does not exist: p = 0)
// A key tile whose 128 keys are all valid takes a fast path without any per-key loads.

// registers -> TMEM: 32 lanes (this warp's quarter) x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  __syncwarp();
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

}  // namespace prd
