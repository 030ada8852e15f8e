// Head-dim-16 attention of the backward pass on the warp-level tensor cores (mma.sync m16n8k8, tf32 operands rounded to
// nearest, fp32 accumulate): forward recompute with log-sum-exp, dq, and dk / dv.  Reference math: modules.py:185-225
// (TriangleAttention rows / columns and FoldingBlock.single_attn with its pair bias), differentiated by hand.
//
// Why mma.sync and not tcgen05 here: with 16-channel heads every product is K = 16 (scores) or N = 16 (outputs); the
// work per score element is one exp and five multiply-adds on the SIMT pipes against 7 x 32 tensor flops, so the kernels
// are bound by the elementwise softmax algebra, not by the tensor pipe -- register-resident fragments (no TMEM round
// trip between the score GEMM and the GEMM that consumes P / dS) are the cheapest way to feed it.  The score tile comes
// out of the first GEMM in the accumulator layout (row g: columns 2t, 2t+1) and goes into the second GEMM as the A
// operand (row g: k-slots t, t+4) WITHOUT a shuffle: the k-slots of the second GEMM are a permutation of the keys
// (slot t <-> key 2t, slot t+4 <-> key 2t+1) and its B operand is read from shared memory with the same permutation.
//
//   forward:  S = (q / 4) k^T (+ bias), masked_fill(-32768), P = softmax, O = P v, lse = log sum exp
//   dq kernel (CTA = 128 queries, loops over keys):   dP = dO v^T, dS = P (dP - D), dq = dS k / 4, optional dbias = dS
//   dkv kernel (CTA = 128 keys, loops over queries):  the same tiles transposed (S^T = k q^T), dv = P^T dO, dk = dS^T q / 4
//
// Scores are kept in log2 units (q carries scale * log2 e, the bias is multiplied by log2 e on the fly, lse is stored as
// log2 sum 2^s): one FADD + one ex2 per probability.  Operands produced in registers (P, dS) are rounded to nearest by
// adding half a tf32 ulp to the bit pattern -- the tensor core drops the 13 low bits itself.  A key's state is one float
// (+inf valid, -32768 log2 e masked_fill, -inf beyond the sequence): min(score, state) applies masked_fill and the
// padding in one instruction.
//
// Shared-memory rows are 16 floats at a stride of 20: both fragment access patterns ((g, t) and (2t, g)) then touch 32
// distinct banks.
#include "prd_bwd.h"

#include <stdlib.h>

#include "prd_common.cuh"

namespace prd {

namespace {
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kMaskFill2 = -32768.0f * kLog2e;  // modules.py:177,220 in log2 units
constexpr int kLd = 20;                 // shared-memory row stride (floats)
constexpr int kChunk = 128;             // keys (queries) staged per pass

struct AttnDev {
  int B, N, H, mode;
  const float* mask;
  const float* bias;
  float scale;
};
__device__ __forceinline__ long long attn_row(const AttnDev& a, long long s, int t) {
  if (a.mode == 1) {
    const long long b = s / a.N;
    return b * a.N * a.N + (long long)t * a.N + (s - b * a.N);
  }
  return s * a.N + t;
}
__device__ __forceinline__ float attn_seq_mask(const AttnDev& a, long long s) { return a.mode == 2 ? 1.0f : a.mask[s]; }
__device__ __forceinline__ long long attn_batch(const AttnDev& a, long long s) { return a.mode == 2 ? s : s / a.N; }

// operand bits of a tf32 MMA, rounded to nearest once the tensor core has dropped the 13 low bits
__device__ __forceinline__ uint32_t tf32_bits(float x) { return __float_as_uint(x) + 0x1000u; }
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// state of key kt of sequence s (see the header)
__device__ __forceinline__ float key_state(const AttnDev& a, long long b, float ms, int kt) {
  return kt < a.N ? ((ms * a.mask[b * a.N + kt] >= 0.5f) ? INFINITY : kMaskFill2) : -INFINITY;
}
// d += a (16 x 8, row) * b (8 x 8, col)
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}

// A-operand fragments of one 16-row tile of a [rows, 16] matrix held in global memory (rows past N read as zero):
// frag[ks] = {(g, 8 ks + t), (g + 8, 8 ks + t), (g, 8 ks + t + 4), (g + 8, 8 ks + t + 4)}
__device__ __forceinline__ void load_a_frag(const AttnDev& a, long long s, int r0, const float* __restrict__ base, long long ld,
                                            int col0, float mul, int g, int t, uint32_t (&frag)[2][4]) {
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    const int r = r0 + g + 8 * hf;
    const float* p = r < a.N ? base + attn_row(a, s, r) * ld + col0 : nullptr;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      frag[ks][hf] = __float_as_uint(round_tf32(p ? p[8 * ks + t] * mul : 0.f));
      frag[ks][2 + hf] = __float_as_uint(round_tf32(p ? p[8 * ks + t + 4] * mul : 0.f));
    }
  }
}

// One thread stages row (r0 + threadIdx.x) of a 16-wide column block into shared memory, rounded to tf32 (zeros past N)
__device__ __forceinline__ void stage_row(const AttnDev& a, long long s, int r, const float* __restrict__ base, long long ld,
                                          int col0, float mul, float* dst) {
  float4 v[4];
  if (r < a.N) {
    const float4* p = reinterpret_cast<const float4*>(base + attn_row(a, s, r) * ld + col0);
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = p[c];
  } else {
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float4 o;
    o.x = round_tf32(v[c].x * mul); o.y = round_tf32(v[c].y * mul); o.z = round_tf32(v[c].z * mul); o.w = round_tf32(v[c].w * mul);
    *reinterpret_cast<float4*>(dst + 4 * c) = o;
  }
}

// -----------------------------------------------------------------------------------------------------------------
// forward: grid (nseq * H, ceil(N / 128)), 4 warps x 32 queries
// -----------------------------------------------------------------------------------------------------------------
template <bool kBias>
__global__ void __launch_bounds__(128) attn_tc_fwd_kernel(AttnDev a, const float* __restrict__ qkvg, long long ld,
                                                          float* __restrict__ O, float* __restrict__ lse) {
  __shared__ __align__(16) float sK[kChunk * kLd];
  __shared__ __align__(16) float sV[kChunk * kLd];
  __shared__ float sFlag[kChunk];
  const int h = blockIdx.x % a.H;
  const long long s = blockIdx.x / a.H;
  const long long b = attn_batch(a, s);
  const float ms = attn_seq_mask(a, s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.y * 128 + warp * 32;
  const bool active = q0 < a.N;  // warp-uniform
  uint32_t qa[2][2][4];
  float acc[2][2][4], m[2][2], l[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    load_a_frag(a, s, q0 + 16 * mt, qkvg, ld, h * 16, a.scale * kLog2e, g, t, qa[mt]);
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      m[mt][hf] = -INFINITY;
      l[mt][hf] = 0.f;
    }
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
  }
  for (int k0 = 0; k0 < a.N; k0 += kChunk) {
    __syncthreads();
    {
      const int kt = k0 + threadIdx.x;
      stage_row(a, s, kt, qkvg, ld, 64 + h * 16, 1.f, sK + threadIdx.x * kLd);
      stage_row(a, s, kt, qkvg, ld, 128 + h * 16, 1.f, sV + threadIdx.x * kLd);
      sFlag[threadIdx.x] = key_state(a, b, ms, kt);
    }
    __syncthreads();
    if (!active) continue;
    const int kn = a.N - k0 < kChunk ? a.N - k0 : kChunk;
    for (int kb = 0; kb < kn; kb += 32) {
      float sc[2][4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int key = kb + 8 * nt;
        const float* kp = sK + (key + g) * kLd + t;
        const uint32_t b00 = __float_as_uint(kp[0]), b01 = __float_as_uint(kp[4]);
        const uint32_t b10 = __float_as_uint(kp[8]), b11 = __float_as_uint(kp[12]);
        const float f0 = sFlag[key + 2 * t], f1 = sFlag[key + 2 * t + 1];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
          for (int e = 0; e < 4; ++e) sc[mt][nt][e] = 0.f;
          mma_tf32(sc[mt][nt], qa[mt][0], b00, b01);
          mma_tf32(sc[mt][nt], qa[mt][1], b10, b11);
          if (kBias) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              const int q = q0 + 16 * mt + g + 8 * hf;
              if (q < a.N) {
                const float* bp = a.bias + ((b * a.H + h) * a.N + q) * (long long)a.N + k0 + key + 2 * t;
                if (k0 + key + 2 * t < a.N) sc[mt][nt][2 * hf] = fmaf(bp[0], kLog2e, sc[mt][nt][2 * hf]);
                if (k0 + key + 2 * t + 1 < a.N) sc[mt][nt][2 * hf + 1] = fmaf(bp[1], kLog2e, sc[mt][nt][2 * hf + 1]);
              }
            }
          }
          sc[mt][nt][0] = fminf(sc[mt][nt][0], f0);
          sc[mt][nt][1] = fminf(sc[mt][nt][1], f1);
          sc[mt][nt][2] = fminf(sc[mt][nt][2], f0);
          sc[mt][nt][3] = fminf(sc[mt][nt][3], f1);
        }
      }
      // online softmax: rows (mt, hf) = q0 + 16 mt + g + 8 hf; the first key of every group is a real key, so the
      // running maximum is finite from the first group on
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          float mx = -INFINITY;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mx = fmaxf(mx, fmaxf(sc[mt][nt][2 * hf], sc[mt][nt][2 * hf + 1]));
          mx = quad_max(mx);
          const float mnew = fmaxf(m[mt][hf], mx);
          const float corr = ex2f(m[mt][hf] - mnew);
          m[mt][hf] = mnew;
          float part = 0.f;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const float p0 = ex2f(sc[mt][nt][2 * hf] - mnew), p1 = ex2f(sc[mt][nt][2 * hf + 1] - mnew);
            part += p0 + p1;
            sc[mt][nt][2 * hf] = p0;
            sc[mt][nt][2 * hf + 1] = p1;
          }
          l[mt][hf] = l[mt][hf] * corr + part;
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            acc[mt][nt][2 * hf] *= corr;
            acc[mt][nt][2 * hf + 1] *= corr;
          }
        }
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int key = kb + 8 * nt;
        const float* vp = sV + (key + 2 * t) * kLd + g;
        const uint32_t v00 = __float_as_uint(vp[0]), v01 = __float_as_uint(vp[kLd]);
        const uint32_t v10 = __float_as_uint(vp[8]), v11 = __float_as_uint(vp[kLd + 8]);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const uint32_t pa[4] = {tf32_bits(sc[mt][nt][0]), tf32_bits(sc[mt][nt][2]), tf32_bits(sc[mt][nt][1]), tf32_bits(sc[mt][nt][3])};
          mma_tf32(acc[mt][0], pa, v00, v01);
          mma_tf32(acc[mt][1], pa, v10, v11);
        }
      }
    }
  }
  if (!active) return;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int q = q0 + 16 * mt + g + 8 * hf;
      const float lt = quad_sum(l[mt][hf]);
      if (q >= a.N) continue;
      const float inv = 1.0f / lt;
      float* op = O + attn_row(a, s, q) * 64 + h * 16 + 2 * t;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
        *reinterpret_cast<float2*>(op + 8 * nt) = make_float2(acc[mt][nt][2 * hf] * inv, acc[mt][nt][2 * hf + 1] * inv);
      if (t == 0) lse[(s * a.H + h) * a.N + q] = m[mt][hf] + log2f(lt);  // log2 units
    }
  }
}

// -----------------------------------------------------------------------------------------------------------------
// dq (and D = <dO, O>, optional dbias): grid (nseq * H, ceil(N / 128)), 4 warps x 32 queries, loop over keys
// -----------------------------------------------------------------------------------------------------------------
template <bool kBias>
__global__ void __launch_bounds__(128) attn_tc_dq_kernel(AttnDev a, const float* __restrict__ qkvg, long long ld,
                                                         const float* __restrict__ O, const float* __restrict__ lse,
                                                         const float* __restrict__ dO, float* __restrict__ Dbuf,
                                                         float* __restrict__ dqkvg, long long ldd, float* __restrict__ dbias) {
  __shared__ __align__(16) float sK[kChunk * kLd];
  __shared__ __align__(16) float sV[kChunk * kLd];
  __shared__ float sFlag[kChunk];
  const int h = blockIdx.x % a.H;
  const long long s = blockIdx.x / a.H;
  const long long b = attn_batch(a, s);
  const float ms = attn_seq_mask(a, s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.y * 128 + warp * 32;
  const bool active = q0 < a.N;
  uint32_t qa[2][2][4], da[2][2][4];
  float dq[2][2][4], L[2][2], D[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    load_a_frag(a, s, q0 + 16 * mt, qkvg, ld, h * 16, a.scale * kLog2e, g, t, qa[mt]);
    load_a_frag(a, s, q0 + 16 * mt, dO, 64, h * 16, 1.f, g, t, da[mt]);
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int q = q0 + 16 * mt + g + 8 * hf;
      float part = 0.f;
      if (q < a.N) {
        const long long row = attn_row(a, s, q);
        const float* op = O + row * 64 + h * 16 + t;
        const float* dp = dO + row * 64 + h * 16 + t;
#pragma unroll
        for (int c = 0; c < 4; ++c) part = fmaf(dp[4 * c], op[4 * c], part);
      }
      const float Dq = quad_sum(part);
      D[mt][hf] = Dq;
      // rows past N: exp(score - inf) = 0 everywhere
      L[mt][hf] = q < a.N ? lse[(s * a.H + h) * a.N + q] : INFINITY;
      if (q < a.N && t == 0) Dbuf[(s * a.H + h) * a.N + q] = Dq;
    }
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) dq[mt][nt][e] = 0.f;
  }
  for (int k0 = 0; k0 < a.N; k0 += kChunk) {
    __syncthreads();
    {
      const int kt = k0 + threadIdx.x;
      stage_row(a, s, kt, qkvg, ld, 64 + h * 16, 1.f, sK + threadIdx.x * kLd);
      stage_row(a, s, kt, qkvg, ld, 128 + h * 16, 1.f, sV + threadIdx.x * kLd);
      sFlag[threadIdx.x] = key_state(a, b, ms, kt);
    }
    __syncthreads();
    if (!active) continue;
    const int kn = a.N - k0 < kChunk ? a.N - k0 : kChunk;
#pragma unroll 2
    for (int key = 0; key < kn; key += 8) {
      const float* kp = sK + (key + g) * kLd + t;
      const float* vp = sV + (key + g) * kLd + t;
      const uint32_t kb00 = __float_as_uint(kp[0]), kb01 = __float_as_uint(kp[4]);
      const uint32_t kb10 = __float_as_uint(kp[8]), kb11 = __float_as_uint(kp[12]);
      const uint32_t vb00 = __float_as_uint(vp[0]), vb01 = __float_as_uint(vp[4]);
      const uint32_t vb10 = __float_as_uint(vp[8]), vb11 = __float_as_uint(vp[12]);
      const float* k2 = sK + (key + 2 * t) * kLd + g;
      const uint32_t kc00 = __float_as_uint(k2[0]), kc01 = __float_as_uint(k2[kLd]);
      const uint32_t kc10 = __float_as_uint(k2[8]), kc11 = __float_as_uint(k2[kLd + 8]);
      // only valid keys carry a gradient: masked_fill cuts it (modules.py:220), keys past the sequence do not exist
      const bool v0 = sFlag[key + 2 * t] > 0.f, v1 = sFlag[key + 2 * t + 1] > 0.f;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        float sc[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
        mma_tf32(sc, qa[mt][0], kb00, kb01);
        mma_tf32(sc, qa[mt][1], kb10, kb11);
        mma_tf32(dp, da[mt][0], vb00, vb01);
        mma_tf32(dp, da[mt][1], vb10, vb11);
        float ds[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int hf = e >> 1;
          float x = sc[e];
          if (kBias) {
            const int q = q0 + 16 * mt + g + 8 * hf, kk = k0 + key + 2 * t + (e & 1);
            if (q < a.N && kk < a.N) x = fmaf(a.bias[((b * a.H + h) * a.N + q) * (long long)a.N + kk], kLog2e, x);
          }
          const float p = ex2f(x - L[mt][hf]);
          ds[e] = ((e & 1) ? v1 : v0) ? p * (dp[e] - D[mt][hf]) : 0.f;
          if (kBias) {
            const int q = q0 + 16 * mt + g + 8 * hf, kk = k0 + key + 2 * t + (e & 1);
            if (dbias != nullptr && q < a.N && kk < a.N) dbias[((b * a.H + h) * a.N + q) * (long long)a.N + kk] = ds[e];
          }
        }
        const uint32_t sa[4] = {tf32_bits(ds[0]), tf32_bits(ds[2]), tf32_bits(ds[1]), tf32_bits(ds[3])};
        mma_tf32(dq[mt][0], sa, kc00, kc01);
        mma_tf32(dq[mt][1], sa, kc10, kc11);
      }
    }
  }
  if (!active) return;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int q = q0 + 16 * mt + g + 8 * hf;
      if (q >= a.N) continue;
      float* o = dqkvg + attn_row(a, s, q) * ldd + h * 16 + 2 * t;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
        *reinterpret_cast<float2*>(o + 8 * nt) =
            make_float2(round_tf32(dq[mt][nt][2 * hf] * a.scale), round_tf32(dq[mt][nt][2 * hf + 1] * a.scale));
    }
}

// -----------------------------------------------------------------------------------------------------------------
// dk, dv: grid (nseq * H, ceil(N / 128)), 4 warps x 32 keys, loop over queries; the tiles are the transposes of the dq
// kernel's (rows = keys, columns = queries)
// -----------------------------------------------------------------------------------------------------------------
template <bool kBias>
__global__ void __launch_bounds__(128) attn_tc_dkv_kernel(AttnDev a, const float* __restrict__ qkvg, long long ld,
                                                          const float* __restrict__ lse, const float* __restrict__ dO,
                                                          const float* __restrict__ Dbuf, float* __restrict__ dqkvg,
                                                          long long ldd) {
  __shared__ __align__(16) float sQ[kChunk * kLd];
  __shared__ __align__(16) float sdO[kChunk * kLd];
  __shared__ float sL[kChunk], sD[kChunk];
  const int h = blockIdx.x % a.H;
  const long long s = blockIdx.x / a.H;
  const long long b = attn_batch(a, s);
  const float ms = attn_seq_mask(a, s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int kw0 = blockIdx.y * 128 + warp * 32;
  const bool active = kw0 < a.N;
  uint32_t ka[2][2][4], va[2][2][4];
  float dk[2][2][4], dv[2][2][4], kf[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    load_a_frag(a, s, kw0 + 16 * mt, qkvg, ld, 64 + h * 16, 1.f, g, t, ka[mt]);
    load_a_frag(a, s, kw0 + 16 * mt, qkvg, ld, 128 + h * 16, 1.f, g, t, va[mt]);
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      // rows past the sequence are never stored, so only valid / masked_fill matters here
      kf[mt][hf] = key_state(a, b, ms, kw0 + 16 * mt + g + 8 * hf);
    }
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) dk[mt][nt][e] = dv[mt][nt][e] = 0.f;
  }
  for (int qc = 0; qc < a.N; qc += kChunk) {
    __syncthreads();
    {
      const int qt = qc + threadIdx.x;
      stage_row(a, s, qt, qkvg, ld, h * 16, a.scale * kLog2e, sQ + threadIdx.x * kLd);
      stage_row(a, s, qt, dO, 64, h * 16, 1.f, sdO + threadIdx.x * kLd);
      // queries past N: exp(score - inf) = 0
      sL[threadIdx.x] = qt < a.N ? lse[(s * a.H + h) * a.N + qt] : INFINITY;
      sD[threadIdx.x] = qt < a.N ? Dbuf[(s * a.H + h) * a.N + qt] : 0.f;
    }
    __syncthreads();
    if (!active) continue;
    const int qn = a.N - qc < kChunk ? a.N - qc : kChunk;
#pragma unroll 2
    for (int q = 0; q < qn; q += 8) {
      const float* qp = sQ + (q + g) * kLd + t;
      const float* dp_ = sdO + (q + g) * kLd + t;
      const uint32_t qb00 = __float_as_uint(qp[0]), qb01 = __float_as_uint(qp[4]);
      const uint32_t qb10 = __float_as_uint(qp[8]), qb11 = __float_as_uint(qp[12]);
      const uint32_t db00 = __float_as_uint(dp_[0]), db01 = __float_as_uint(dp_[4]);
      const uint32_t db10 = __float_as_uint(dp_[8]), db11 = __float_as_uint(dp_[12]);
      const float* q2 = sQ + (q + 2 * t) * kLd + g;
      const float* d2 = sdO + (q + 2 * t) * kLd + g;
      const uint32_t qc00 = __float_as_uint(q2[0]), qc01 = __float_as_uint(q2[kLd]);
      const uint32_t qc10 = __float_as_uint(q2[8]), qc11 = __float_as_uint(q2[kLd + 8]);
      const uint32_t dc00 = __float_as_uint(d2[0]), dc01 = __float_as_uint(d2[kLd]);
      const uint32_t dc10 = __float_as_uint(d2[8]), dc11 = __float_as_uint(d2[kLd + 8]);
      const float L0 = sL[q + 2 * t], L1 = sL[q + 2 * t + 1];
      const float D0 = sD[q + 2 * t], D1 = sD[q + 2 * t + 1];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        float sc[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
        mma_tf32(sc, ka[mt][0], qb00, qb01);
        mma_tf32(sc, ka[mt][1], qb10, qb11);
        mma_tf32(dp, va[mt][0], db00, db01);
        mma_tf32(dp, va[mt][1], db10, db11);
        float p[4], ds[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int hf = e >> 1;
          const float f = kf[mt][hf];
          float x = sc[e];
          if (kBias) {
            const int qq = qc + q + 2 * t + (e & 1), kt = kw0 + 16 * mt + g + 8 * hf;
            if (qq < a.N && kt < a.N) x = fmaf(a.bias[((b * a.H + h) * a.N + qq) * (long long)a.N + kt], kLog2e, x);
          }
          p[e] = ex2f(fminf(x, f) - ((e & 1) ? L1 : L0));
          ds[e] = f > 0.f ? p[e] * (dp[e] - ((e & 1) ? D1 : D0)) : 0.f;
        }
        const uint32_t pa[4] = {tf32_bits(p[0]), tf32_bits(p[2]), tf32_bits(p[1]), tf32_bits(p[3])};
        const uint32_t sa[4] = {tf32_bits(ds[0]), tf32_bits(ds[2]), tf32_bits(ds[1]), tf32_bits(ds[3])};
        mma_tf32(dv[mt][0], pa, dc00, dc01);
        mma_tf32(dv[mt][1], pa, dc10, dc11);
        mma_tf32(dk[mt][0], sa, qc00, qc01);  // sQ carries scale * log2 e
        mma_tf32(dk[mt][1], sa, qc10, qc11);
      }
    }
  }
  if (!active) return;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int kt = kw0 + 16 * mt + g + 8 * hf;
      if (kt >= a.N) continue;
      float* o = dqkvg + attn_row(a, s, kt) * ldd + h * 16 + 2 * t;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        *reinterpret_cast<float2*>(o + 64 + 8 * nt) =
            make_float2(round_tf32(dk[mt][nt][2 * hf] * (1.0f / kLog2e)), round_tf32(dk[mt][nt][2 * hf + 1] * (1.0f / kLog2e)));
        *reinterpret_cast<float2*>(o + 128 + 8 * nt) = make_float2(round_tf32(dv[mt][nt][2 * hf]), round_tf32(dv[mt][nt][2 * hf + 1]));
      }
    }
}

AttnDev attn_dev(const AttnGeom& g) { return AttnDev{g.B, g.N, g.H, g.mode, g.mask, g.bias, g.scale}; }
long long attn_nseq(const AttnGeom& g) { return g.mode == 2 ? g.B : (long long)g.B * g.N; }
}  // namespace

bool bw_attn_tc_enabled() {
  const char* e = getenv("PRD_ATTN_SIMT");
  return !(e && e[0] == '1');
}

int bw_attn_tc_fwd(const AttnGeom& g, const float* qkvg, long long ld, float* O, float* lse, cudaStream_t s) {
  PRD_REQUIRE(ld % 4 == 0, "attn: row stride must be a multiple of 4 floats");
  const long long nsh = attn_nseq(g) * g.H;
  PRD_REQUIRE(nsh < 2147483647LL, "attn: too many (sequence, head) pairs");
  const dim3 grid((unsigned)nsh, (g.N + 127) / 128);
  if (g.bias) attn_tc_fwd_kernel<true><<<grid, 128, 0, s>>>(attn_dev(g), qkvg, ld, O, lse);
  else attn_tc_fwd_kernel<false><<<grid, 128, 0, s>>>(attn_dev(g), qkvg, ld, O, lse);
  PRD_LAUNCHED();
  return 0;
}

int bw_attn_tc_bwd(const AttnGeom& g, const float* qkvg, long long ld, const float* O, const float* lse, const float* dO,
                   float* Dbuf, float* dqkvg, long long ldd, float* dbias, cudaStream_t s) {
  PRD_REQUIRE(ld % 4 == 0 && ldd % 2 == 0, "attn: row strides must be multiples of 4 / 2 floats");
  PRD_REQUIRE(dbias == nullptr || g.bias != nullptr, "attn: dbias without a bias");
  const long long nsh = attn_nseq(g) * g.H;
  PRD_REQUIRE(nsh < 2147483647LL, "attn: too many (sequence, head) pairs");
  const dim3 grid((unsigned)nsh, (g.N + 127) / 128);
  if (g.bias) {
    attn_tc_dq_kernel<true><<<grid, 128, 0, s>>>(attn_dev(g), qkvg, ld, O, lse, dO, Dbuf, dqkvg, ldd, dbias);
    PRD_LAUNCHED();
    attn_tc_dkv_kernel<true><<<grid, 128, 0, s>>>(attn_dev(g), qkvg, ld, lse, dO, Dbuf, dqkvg, ldd);
    PRD_LAUNCHED();
  } else {
    attn_tc_dq_kernel<false><<<grid, 128, 0, s>>>(attn_dev(g), qkvg, ld, O, lse, dO, Dbuf, dqkvg, ldd, dbias);
    PRD_LAUNCHED();
    attn_tc_dkv_kernel<false><<<grid, 128, 0, s>>>(attn_dev(g), qkvg, ld, lse, dO, Dbuf, dqkvg, ldd);
    PRD_LAUNCHED();
  }
  return 0;
}

}  // namespace prd
