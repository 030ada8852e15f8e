// Batched fp16 x fp16 -> fp32 GEMM on tcgen05 tensor cores with TMA-staged operands.
//
//   C[b][m][n] = epilogue( alpha * sum_k A[b][m][k] * B[b][n][k] )        (both operands K-major)
//
// One CTA computes one 128 x BN output tile: warp 0 = TMA producer, warp 1 = UMMA issuer,
// warps 2..3 idle (they only complete the first warpgroup, which gives its registers away), warps 4..11 = epilogue
// (TMEM -> registers -> global; two warps per TMEM lane quarter, alternating 32-column chunks; fp32 results leave as
// TMA tile stores when the C layout allows a tensor map).  Operand tiles are [rows x 64] halves
// in SWIZZLE_128B shared memory, STAGES-deep mbarrier ring; the accumulator lives in TMEM.
//
// Used for: the triangle-multiplication contraction (per (b, channel) NxNxN GEMMs), every
// single-representation linear layer, SPAttention logits / PV.  See prd_denoiser.h.
#include <stdlib.h>

#include "prd_common.cuh"
#include "prd_kernels.h"

namespace prd {

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int kABytes = 128 * 64 * 2;
  static constexpr int kBBytes = BN * 64 * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiBytes = 8 * 4096;  // one 4 KB store-coalescing slice per epilogue warp
  static constexpr int kTotal = STAGES * kStageBytes + kEpiBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;  // two accumulators
};

struct GemmEpilogue {
  int M, N;
  float alpha;
  const float* bias;                          // [N]
  int act;                                    // 0 none, 1 relu, 2 sigmoid
  const float* rowscale;                      // [M] per batch
  long long rs_bs1, rs_bs2;
  const float* mul;                           // [M, N] fp32, elementwise factor
  long long ldmul, mul_bs1, mul_bs2;
  const float* add;                           // [M, N] fp32, elementwise addend (residual / 2-D bias)
  long long ldadd, add_bs1, add_bs2;
  void* C;
  long long ldc, c_bs1, c_bs2;
  int c_fp16;
  int mul_step;   // the mul operand gates instead of scaling: v = mul > 0 ? v : 0 (ReLU backward)
  int round_tf32; // fp32 C rounded to nearest tf32 (the result feeds another tf32 GEMM as an operand)
  int c_tma;      // fp32 C with 16-byte aligned rows / batches: 32 x 32 boxes leave through map_c (clipped at M, N)
  int c_b1, c_b2; // 1 when that batch index is a coordinate of map_c
};

struct GemmTiling {
  int tiles_n, tiles_m, nb1;
  long long total;
  __device__ __forceinline__ void decode(long long tile, int& n0, int& m0, int& i1, int& i2, int bn) const {
    const int tn = static_cast<int>(tile % tiles_n);
    long long r = tile / tiles_n;
    const int tm = static_cast<int>(r % tiles_m);
    r /= tiles_m;
    i1 = static_cast<int>(r % nb1);
    i2 = static_cast<int>(r / nb1);
    n0 = tn * bn;
    m0 = tm * 128;
  }
};

// Persistent: every CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  The accumulator is
// double buffered in TMEM, so the epilogue of tile i (TMEM -> registers -> global) overlaps the
// TMA / UMMA main loop of tile i+1.
template <int BN, int STAGES, bool TF32>
__global__ void __launch_bounds__(384, 1)
gemm_f16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const __grid_constant__ CUtensorMap map_c, int num_kb,
                int a_wrap, int b_wrap, GemmTiling tl, int a_b1, int a_b2, int b_b1, int b_b2, GemmEpilogue ep) {
  // TF32: operands are fp32 in memory (kind::tf32 reads the upper 19 bits: producers round to nearest first), one K-block =
  // 32 elements = the same 128-byte swizzle row, one UMMA = K 8 = the same 32-byte descriptor step as 16 halves.
  // a_wrap / b_wrap: K-blocks after which that operand's K coordinate wraps to 0.  With a weight
  // stored as [hi | lo] along K (fp16 pair, removes the systematic weight rounding) the activation
  // operand is simply read twice: sum_k x_k (w_hi + w_lo)_k.
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  using L = GemmSmem<BN, STAGES>;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* epi = smem + STAGES * L::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi + L::kEpiBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;       // [2]
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 256);
    }
    fence_barrier_init();
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (ep.c_tma) tma_prefetch_desc(&map_c);
  }
  if (warp == 0) tmem_alloc(tmem_slot, L::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // everything above touched only weights / shared memory; the predecessor's output is read below
  const uint32_t tmem = *tmem_slot;
  // 384 threads x 168 registers = 128 x 56 + 256 x 224: the epilogue keeps a chunk of the accumulator and of both [M, N]
  // epilogue operands in registers
  // (setmaxnreg at the head of each role's branch)

  if (warp == 0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (lane == 0) {
      long long it = 0;  // running k-block counter across tiles
      for (long long tile = blockIdx.x; tile < tl.total; tile += gridDim.x) {
        int n0, m0, i1, i2;
        tl.decode(tile, n0, m0, i1, i2, BN);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = static_cast<int>(it % STAGES);
          const uint32_t ph = static_cast<uint32_t>((it / STAGES) & 1);
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], L::kStageBytes);
          uint8_t* sa = smem + s * L::kStageBytes;
          constexpr int kKbElems = TF32 ? 32 : 64;
          tma_load_4d(sa, &map_a, &full[s], (kb % a_wrap) * kKbElems, m0, i1 * a_b1, i2 * a_b2);
          tma_load_4d(sa + L::kABytes, &map_b, &full[s], (kb % b_wrap) * kKbElems, n0, i1 * b_b1, i2 * b_b2);
        }
      }
    }
  } else if (warp == 1) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (lane == 0) {
      constexpr uint32_t idesc = TF32 ? umma_idesc_tf32(128, BN) : umma_idesc_f16(128, BN);
      long long it = 0;
      int lt = 0;  // local tile counter
      for (long long tile = blockIdx.x; tile < tl.total; tile += gridDim.x, ++lt) {
        const int buf = lt & 1;
        mbar_wait(&tmem_empty[buf], ((lt >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = static_cast<int>(it % STAGES);
          const uint32_t ph = static_cast<uint32_t>((it / STAGES) & 1);
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t sa = base + s * L::kStageBytes;
          if constexpr (TF32) umma_kblock_tf32(tmem + buf * BN, sa, sa + L::kABytes, idesc, kb > 0);
          else umma_kblock(tmem + buf * BN, sa, sa + L::kABytes, idesc, kb > 0);
          umma_commit(&empty[s]);
        }
        umma_commit(&tmem_full[buf]);
      }
    }
  } else if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // epilogue: warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32); the two warps of a lane quarter take
    // alternating column chunks
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    uint8_t* slice = epi + (warp - 4) * 4096;
    int lt = 0;
    for (long long tile = blockIdx.x; tile < tl.total; tile += gridDim.x, ++lt) {
      int n0, m0, i1, i2;
      tl.decode(tile, n0, m0, i1, i2, BN);
      const int buf = lt & 1;
      const int row = m0 + q * 32 + lane;
      mbar_wait(&tmem_full[buf], (lt >> 1) & 1);
      tc_fence_after();
      const long long boff_c = (long long)i1 * ep.c_bs1 + (long long)i2 * ep.c_bs2;
      const float rs = (ep.rowscale != nullptr && row < ep.M)
                           ? ep.rowscale[(long long)i1 * ep.rs_bs1 + (long long)i2 * ep.rs_bs2 + row]
                           : 1.0f;
      const float* mulp = ep.mul ? ep.mul + (long long)i1 * ep.mul_bs1 + (long long)i2 * ep.mul_bs2 + (long long)row * ep.ldmul : nullptr;
      const float* addp = ep.add ? ep.add + (long long)i1 * ep.add_bs1 + (long long)i2 * ep.add_bs2 + (long long)row * ep.ldadd : nullptr;
      const bool plain = ep.bias == nullptr && ep.act == 0 && ep.rowscale == nullptr && mulp == nullptr &&
                         addp == nullptr && ep.alpha == 1.0f;
      // warp-uniform: fp32 C whose full tile width exists and whose rows are 16-byte aligned
      const bool coalesce = !ep.c_fp16 && (n0 + BN <= ep.N) && ((ep.ldc & 3) == 0) &&
                            ((reinterpret_cast<uintptr_t>(reinterpret_cast<float*>(ep.C) + boff_c + n0) & 15) == 0);
      // warp-uniform: plain fp16 C (the tri-mul contraction result), full tile width, 16-byte aligned rows: 64-column chunks
      // (128 bytes per row) leave as full lines
      // (bias / activation / alpha are applied in registers; row scale, mul and add operands keep the general path)
      const bool simple = ep.rowscale == nullptr && mulp == nullptr && addp == nullptr;
      const bool coalesce16 = ep.c_fp16 && simple && BN >= 64 && (n0 + BN <= ep.N) && ((ep.ldc & 7) == 0) && (m0 + 128 <= ep.M) &&
                              ((reinterpret_cast<uintptr_t>(reinterpret_cast<__half*>(ep.C) + boff_c + n0) & 15) == 0);
      if (coalesce16) {
#pragma unroll 1
        for (int c = half; c < BN / 64; c += 2) {
          uint32_t r0[32], r1[32];
          uint4 ov[8];
          tmem_ld32(tmem + buf * BN + (static_cast<uint32_t>(q * 32) << 16) + c * 64, r0);
          tmem_ld32(tmem + buf * BN + (static_cast<uint32_t>(q * 32) << 16) + c * 64 + 32, r1);
          tmem_ld_wait();
          if (!plain) {
            const int colb = n0 + c * 64;
            const bool bvec = ep.bias != nullptr && ((reinterpret_cast<uintptr_t>(ep.bias + colb) & 15) == 0);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
              if (bvec) {
                b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + colb) + j4);
                b1 = __ldg(reinterpret_cast<const float4*>(ep.bias + colb + 32) + j4);
              } else if (ep.bias != nullptr) {
                b0 = make_float4(__ldg(ep.bias + colb + 4 * j4), __ldg(ep.bias + colb + 4 * j4 + 1), __ldg(ep.bias + colb + 4 * j4 + 2),
                                 __ldg(ep.bias + colb + 4 * j4 + 3));
                b1 = make_float4(__ldg(ep.bias + colb + 32 + 4 * j4), __ldg(ep.bias + colb + 32 + 4 * j4 + 1),
                                 __ldg(ep.bias + colb + 32 + 4 * j4 + 2), __ldg(ep.bias + colb + 32 + 4 * j4 + 3));
              }
              const float bb0[4] = {b0.x, b0.y, b0.z, b0.w}, bb1[4] = {b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int j = 4 * j4 + e;
                float x0 = fmaf(ep.alpha, __uint_as_float(r0[j]), bb0[e]), x1 = fmaf(ep.alpha, __uint_as_float(r1[j]), bb1[e]);
                if (ep.act == 1) {
                  x0 = fmaxf(x0, 0.0f);
                  x1 = fmaxf(x1, 0.0f);
                } else if (ep.act == 2) {
                  x0 = 1.0f / (1.0f + __expf(-x0));
                  x1 = 1.0f / (1.0f + __expf(-x1));
                }
                r0[j] = __float_as_uint(x0);
                r1[j] = __float_as_uint(x1);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            ov[j] = make_uint4(pack_half2(__uint_as_float(r0[8 * j]), __uint_as_float(r0[8 * j + 1])),
                               pack_half2(__uint_as_float(r0[8 * j + 2]), __uint_as_float(r0[8 * j + 3])),
                               pack_half2(__uint_as_float(r0[8 * j + 4]), __uint_as_float(r0[8 * j + 5])),
                               pack_half2(__uint_as_float(r0[8 * j + 6]), __uint_as_float(r0[8 * j + 7])));
            ov[4 + j] = make_uint4(pack_half2(__uint_as_float(r1[8 * j]), __uint_as_float(r1[8 * j + 1])),
                                   pack_half2(__uint_as_float(r1[8 * j + 2]), __uint_as_float(r1[8 * j + 3])),
                                   pack_half2(__uint_as_float(r1[8 * j + 4]), __uint_as_float(r1[8 * j + 5])),
                                   pack_half2(__uint_as_float(r1[8 * j + 6]), __uint_as_float(r1[8 * j + 7])));
          }
          __half* cb = reinterpret_cast<__half*>(ep.C) + boff_c + (long long)(m0 + q * 32) * ep.ldc + n0 + c * 64;
          warp_store_rows128(slice, lane, ov, cb, (long long)ep.ldc * 2, 32);
        }
        tc_fence_before();
        mbar_arrive(&tmem_empty[buf]);
        continue;
      }
#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        uint32_t r[32];
        uint4 ov[8];
        tmem_ld32(tmem + buf * BN + (static_cast<uint32_t>(q * 32) << 16) + c * 32, r);
        const int col0 = n0 + c * 32;
        const int rows_left = ep.M - (m0 + q * 32);
        // the [M, N] epilogue operands of this chunk: coalesced loads issued before the wait on the accumulator
        const float* mb = ep.mul ? ep.mul + (long long)i1 * ep.mul_bs1 + (long long)i2 * ep.mul_bs2 + (long long)(m0 + q * 32) * ep.ldmul + col0 : nullptr;
        const float* ab = ep.add ? ep.add + (long long)i1 * ep.add_bs1 + (long long)i2 * ep.add_bs2 + (long long)(m0 + q * 32) * ep.ldadd + col0 : nullptr;
        const bool mul_fast = mb != nullptr && (col0 + 32 <= ep.N) && ((ep.ldmul & 3) == 0) && ((reinterpret_cast<uintptr_t>(mb) & 15) == 0);
        const bool add_fast = ab != nullptr && (col0 + 32 <= ep.N) && ((ep.ldadd & 3) == 0) && ((reinterpret_cast<uintptr_t>(ab) & 15) == 0);
        uint4 mreg[8], areg[8];
        if (mul_fast) warp_load_rows128_issue(lane, mreg, mb, (long long)ep.ldmul * 4, rows_left);
        if (add_fast) warp_load_rows128_issue(lane, areg, ab, (long long)ep.ldadd * 4, rows_left);
        tmem_ld_wait();
        if (ep.c_tma && (mul_fast || add_fast)) {  // the slice may still be read by this warp's previous tile store
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
        }
        // Epilogue operands: per-element scalar loads (32 bias loads and, per lane, 32 strided mul / add loads per chunk) made
        // the four epilogue warps the bottleneck of every single-representation GEMM (ncu: tensor pipe 3-7 % on the
        // SPAttention logits / PV / gate GEMMs).  Full chunks now take the bias as 8 broadcast float4 loads and the
        // [M, N] operands as coalesced 128-byte rows through the warp's slice.
        const bool chunk_full = (col0 + 32 <= ep.N);  // warp-uniform
        const bool active = row < ep.M && col0 < ep.N;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (!plain) {
          const bool bias_vec = ep.bias != nullptr && chunk_full && ((reinterpret_cast<uintptr_t>(ep.bias + col0) & 15) == 0);
          if (bias_vec) {
            const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bb = __ldg(b4 + j);
              v[4 * j] = fmaf(ep.alpha, v[4 * j], bb.x);
              v[4 * j + 1] = fmaf(ep.alpha, v[4 * j + 1], bb.y);
              v[4 * j + 2] = fmaf(ep.alpha, v[4 * j + 2], bb.z);
              v[4 * j + 3] = fmaf(ep.alpha, v[4 * j + 3], bb.w);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              v[j] *= ep.alpha;
              if (ep.bias != nullptr && col0 + j < ep.N) v[j] += __ldg(ep.bias + col0 + j);
            }
          }
          if (ep.act == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
          } else if (ep.act == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 1.0f / (1.0f + __expf(-v[j]));
          }
          if (ep.rowscale != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= rs;
          }
          if (ep.mul != nullptr) {
            if (mul_fast) {
              uint4 t4[8];
              warp_load_rows128_finish(slice, lane, mreg, t4);
              if (ep.mul_step) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  v[4 * j] = __uint_as_float(t4[j].x) > 0.f ? v[4 * j] : 0.f; v[4 * j + 1] = __uint_as_float(t4[j].y) > 0.f ? v[4 * j + 1] : 0.f;
                  v[4 * j + 2] = __uint_as_float(t4[j].z) > 0.f ? v[4 * j + 2] : 0.f; v[4 * j + 3] = __uint_as_float(t4[j].w) > 0.f ? v[4 * j + 3] : 0.f;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  v[4 * j] *= __uint_as_float(t4[j].x); v[4 * j + 1] *= __uint_as_float(t4[j].y);
                  v[4 * j + 2] *= __uint_as_float(t4[j].z); v[4 * j + 3] *= __uint_as_float(t4[j].w);
                }
              }
            } else if (active) {
              // (static indices only: a dynamically indexed v[] would live in local memory for the whole epilogue)
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < ep.N) v[j] = ep.mul_step ? (mulp[col0 + j] > 0.f ? v[j] : 0.f) : v[j] * mulp[col0 + j];
            }
          }
          if (ep.add != nullptr) {
            if (add_fast) {
              uint4 t4[8];
              warp_load_rows128_finish(slice, lane, areg, t4);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                v[4 * j] += __uint_as_float(t4[j].x); v[4 * j + 1] += __uint_as_float(t4[j].y);
                v[4 * j + 2] += __uint_as_float(t4[j].z); v[4 * j + 3] += __uint_as_float(t4[j].w);
              }
            } else if (active) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < ep.N) v[j] += addp[col0 + j];
            }
          }
        }
        if (ep.round_tf32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = round_tf32(v[j]);
        }
        if (ep.c_tma) {
          // fp32 C through the tensor map: the warp's 32 x 32 box goes to its swizzled slice and leaves as one tile store
          // (rows past M / columns past N are clipped by the TMA unit)
          if (rows_left > 0 && col0 < ep.N) {  // warp-uniform
            if (lane == 0) bulk_wait_read0();  // the previous box of this warp has been read out of the slice
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(slice + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                  make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&map_c, slice, col0, m0 + q * 32, i1 * ep.c_b1, i2 * ep.c_b2);
              bulk_commit();
            }
          }
          continue;
        }
        if (active) {
          const bool full_chunk = (col0 + 32 <= ep.N);
          if (ep.c_fp16) {
            __half* cp = reinterpret_cast<__half*>(ep.C) + boff_c + (long long)row * ep.ldc + col0;
            if (full_chunk && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 o;
                o.x = pack_half2(v[j], v[j + 1]);
                o.y = pack_half2(v[j + 2], v[j + 3]);
                o.z = pack_half2(v[j + 4], v[j + 5]);
                o.w = pack_half2(v[j + 6], v[j + 7]);
                *reinterpret_cast<uint4*>(cp + j) = o;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < ep.N) cp[j] = __float2half_rn(v[j]);
            }
          } else if (!coalesce) {
            float* cp = reinterpret_cast<float*>(ep.C) + boff_c + (long long)row * ep.ldc + col0;
            if (full_chunk && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(cp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < ep.N) cp[j] = v[j];
            }
          }
          if (coalesce) {
#pragma unroll
            for (int j = 0; j < 8; ++j) ov[j] = make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                                                           __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
          }
        }
        if (coalesce && col0 < ep.N) {
          // fp32 output, every 32-column chunk of the tile complete and 16-byte aligned (warp-uniform): the
          // warp's 32 rows x 128 bytes leave as full lines
          float* cb = reinterpret_cast<float*>(ep.C) + boff_c + (long long)(m0 + q * 32) * ep.ldc + col0;
          warp_store_rows128(slice, lane, ov, cb, (long long)ep.ldc * 4, ep.M - (m0 + q * 32));
        }
      }
      // this thread's TMEM reads of the accumulator are complete: hand it back to the MMA warp
      tc_fence_before();
      mbar_arrive(&tmem_empty[buf]);
    }
    if (ep.c_tma && lane == 0) bulk_wait_read0();  // shared memory must outlive the last tile store's read
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, L::kTmemCols);
}

template <int BN, int STAGES, bool TF32>
static int launch_gemm(const GemmArgs& g, cudaStream_t stream) {
  using L = GemmSmem<BN, STAGES>;
  CUtensorMap map_a, map_b;
  TmaDims da, db;
  constexpr int kEl = TF32 ? 4 : 2;          // operand element bytes
  constexpr int kKb = TF32 ? 32 : 64;        // elements per 128-byte K-block
  constexpr int kAl = 16 / kEl;              // elements per 16 bytes (TMA stride granularity)
  const int nb1 = g.nb1 > 0 ? g.nb1 : 1, nb2 = g.nb2 > 0 ? g.nb2 : 1;
  const int kb_half = (g.K + kKb - 1) / kKb;
  if (g.split != 0) PRD_REQUIRE(g.K % kKb == 0, "gemm: split weights need K %% %d == 0 (K=%d)", kKb, g.K);
  PRD_REQUIRE(!(TF32 && g.c_fp16), "gemm: tf32 operands write fp32 results");
  auto fill = [&](TmaDims& d, long long rows, long long ld, long long bs1, long long bs2, int box_rows, int k_copies) {
    d.size[0] = (uint64_t)k_copies * g.K;
    d.size[1] = (uint64_t)rows;
    d.size[2] = bs1 != 0 ? (uint64_t)nb1 : 1;
    d.size[3] = bs2 != 0 ? (uint64_t)nb2 : 1;
    d.stride[0] = (uint64_t)ld * kEl;
    d.stride[1] = (uint64_t)(bs1 != 0 ? bs1 : ld * rows) * kEl;
    d.stride[2] = (uint64_t)(bs2 != 0 ? bs2 : ld * rows) * kEl;
    // dims of size 1 still need a 16-byte-multiple stride
    d.stride[1] = (d.stride[1] + 15) & ~15ull;
    d.stride[2] = (d.stride[2] + 15) & ~15ull;
    d.box[0] = kKb;
    d.box[1] = (uint32_t)box_rows;
    d.box[2] = 1;
    d.box[3] = 1;
  };
  PRD_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "gemm: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
  PRD_REQUIRE(g.lda % kAl == 0 && g.ldb % kAl == 0, "gemm: lda/ldb must be multiples of 16 bytes (lda=%lld ldb=%lld)", g.lda, g.ldb);
  PRD_REQUIRE((g.a_bs1 % kAl == 0) && (g.a_bs2 % kAl == 0) && (g.b_bs1 % kAl == 0) && (g.b_bs2 % kAl == 0), "gemm: batch strides must be multiples of 16 bytes");
  fill(da, g.M, g.lda, g.a_bs1, g.a_bs2, 128, (g.split == 2 || g.split == 3) ? 2 : 1);
  fill(db, g.N, g.ldb, g.b_bs1, g.b_bs2, BN, g.split == 1 ? 2 : (g.split == 3 ? 3 : 1));
  if (make_tensor_map(&map_a, g.A, kEl, 4, da, true)) return 1;
  if (make_tensor_map(&map_b, g.B, kEl, 4, db, true)) return 1;
  GemmEpilogue ep;
  ep.M = g.M; ep.N = g.N; ep.alpha = g.alpha; ep.bias = g.bias; ep.act = g.act;
  ep.rowscale = g.rowscale; ep.rs_bs1 = g.rs_bs1; ep.rs_bs2 = g.rs_bs2;
  ep.mul = g.mul; ep.ldmul = g.ldmul; ep.mul_bs1 = g.mul_bs1; ep.mul_bs2 = g.mul_bs2;
  ep.add = g.add; ep.ldadd = g.ldadd; ep.add_bs1 = g.add_bs1; ep.add_bs2 = g.add_bs2;
  ep.C = g.C; ep.ldc = g.ldc; ep.c_bs1 = g.c_bs1; ep.c_bs2 = g.c_bs2; ep.c_fp16 = g.c_fp16;
  ep.mul_step = g.mul_step; ep.round_tf32 = g.round_tf32;
  // fp32 C whose rows and batch strides are 16-byte multiples: tile stores through a tensor map (PRD_GEMM_TMA_STORE=0: off)
  static const bool tma_store_on = !(getenv("PRD_GEMM_TMA_STORE") && getenv("PRD_GEMM_TMA_STORE")[0] == '0');
  CUtensorMap map_c = map_a;
  ep.c_tma = 0; ep.c_b1 = g.c_bs1 != 0 ? 1 : 0; ep.c_b2 = g.c_bs2 != 0 ? 1 : 0;
  if (tma_store_on && !g.c_fp16 && g.ldc % 4 == 0 && g.c_bs1 % 4 == 0 && g.c_bs2 % 4 == 0 &&
      (reinterpret_cast<uintptr_t>(g.C) & 15) == 0) {
    TmaDims dc;
    dc.size[0] = (uint64_t)g.N; dc.size[1] = (uint64_t)g.M;
    dc.size[2] = g.c_bs1 != 0 ? (uint64_t)nb1 : 1; dc.size[3] = g.c_bs2 != 0 ? (uint64_t)nb2 : 1;
    dc.stride[0] = (uint64_t)g.ldc * 4;
    dc.stride[1] = ((uint64_t)(g.c_bs1 != 0 ? g.c_bs1 : g.ldc * (long long)g.M) * 4 + 15) & ~15ull;
    dc.stride[2] = ((uint64_t)(g.c_bs2 != 0 ? g.c_bs2 : g.ldc * (long long)g.M) * 4 + 15) & ~15ull;
    dc.box[0] = 32; dc.box[1] = 32; dc.box[2] = 1; dc.box[3] = 1;
    if (make_tensor_map(&map_c, g.C, 4, 4, dc, true)) return 1;
    ep.c_tma = 1;
  }
  auto kern = gemm_f16_kernel<BN, STAGES, TF32>;
  static bool attr_set = false;
  if (!attr_set) {
    PRD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  GemmTiling tl;
  tl.tiles_n = (g.N + BN - 1) / BN;
  tl.tiles_m = (g.M + 127) / 128;
  tl.nb1 = nb1;
  tl.total = (long long)tl.tiles_n * tl.tiles_m * nb1 * nb2;
  const int grid = (int)(tl.total < kNumSMs ? tl.total : kNumSMs);
  const int num_kb = g.split == 3 ? 3 * kb_half : (g.split != 0 ? 2 * kb_half : kb_half);
  const int a_wrap = g.split == 1 ? kb_half : (g.split == 3 ? 2 * kb_half : num_kb);  // split 3: hi, lo, hi again
  const int b_wrap = g.split == 2 ? kb_half : num_kb;
  PRD_CUDA_OK(launch_pdl(kern, grid, 384, L::kTotal, stream, map_a, map_b, map_c, num_kb, a_wrap, b_wrap, tl, g.a_bs1 != 0 ? 1 : 0,
                         g.a_bs2 != 0 ? 1 : 0, g.b_bs1 != 0 ? 1 : 0, g.b_bs2 != 0 ? 1 : 0, ep));
  PRD_LAUNCHED();
  return 0;
}

int gemm_f16(const GemmArgs& g, cudaStream_t stream) {
  if (g.tf32) {  // fp32 operands on kind::tf32 (the backward pass): same tiles, same pipeline
    if (g.N <= 64) return launch_gemm<64, 6, true>(g, stream);
    return launch_gemm<128, 5, true>(g, stream);
  }
  if (g.N <= 64) return launch_gemm<64, 6, false>(g, stream);
  if (g.N % 256 == 0 && (long long)((g.M + 127) / 128) * (g.N / 256) * (g.nb1 > 0 ? g.nb1 : 1) * (g.nb2 > 0 ? g.nb2 : 1) >= kNumSMs)
    return launch_gemm<256, 4, false>(g, stream);
  return launch_gemm<128, 5, false>(g, stream);
}

}  // namespace prd
