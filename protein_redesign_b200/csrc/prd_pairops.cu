// Pair-stack kernels that mix single-representation operands into the pair tensor, and the
// coordinate head:
//   outer_linear        OuterLinear (modules.py:283-287) in bilinear form
//   pair_embed_dynamic  per-step pair embedding: RBF distance projection (modules.py:73-82,
//                       model.py:359-361) + time embedding + OuterProductUpdate epilogue
//                       (AF2_modules.py:519-543, modules.py:395-397)
//   coord_head          symmetrise (modules.py:403) + weight_radial + equivariant sum (model.py:364-372)
#include "prd_kernels.h"
#include <stdlib.h>
#include "prd_rowtile.cuh"

namespace prd {

// =========================================================================================
// OuterLinear:  dst[b,i,j,z] = [pair +] sum_d W1[z,d] x_i[d] x_j[d] + u[b,i,z] - u[b,j,z] + bias[z]
// with x = LN(single), u = x W2^T (precomputed), W = [W1 | W2] = linear.weight[:, :c_s | c_s:].
// CTA = (j-tile of 128 tokens, chunk of i, b); A = x[b, j-tile, :] stays in shared memory (c_s/64 K-blocks,
// TMA).  Warp-specialised, no CTA-wide barrier in the row loop:
//   warps 0-7   builders: W1 lives in their REGISTERS (each thread owns a fixed set of 16-byte chunks); for
//               every i the B operand (W1 * x_i) is rebuilt K-block by K-block into a 4-slot ring
//               (mbarriers full[] / empty[])
//   warps 8-11  epilogue: thread = token j (TMEM lane): accumulator + u_i - u_j + bias + residual, pair rows
//               loaded and stored as full lines through warp-private shared-memory slices
//   warp 12     UMMA issue (c_s/16 UMMAs per row into one of two [128 x c_z] accumulators) and the A-tile TMA
// The only per-row synchronisation of the builders is one named barrier (x_i staged in shared memory).
// =========================================================================================
// Shared-memory split per configuration: the B-operand ring depth and whether the residual rows get a TMA stage.
// ncu: the builders wait for free ring slots 70 % of the time with a 4-slot ring (a slot is free only when the UMMAs
// that read it have COMPLETED, ~1000 cycles after issue), so depth goes first.
template <int CZ, int KBS>
struct OlCfg {
  static constexpr int kRing = 4;  // (a ring of 8 without the residual stage measures 0.39 ms against 0.31)
  static constexpr bool kResTma = true;
  static constexpr int kSmem = 1024 + KBS * 16384 + kRing * CZ * 128 + (kResTma ? (CZ * 4 / 128) * 16384 : 0) + 4 * 4096 +
                               (2 * KBS * 64 + 5 * CZ) * 4 + 256;
};
constexpr int kOlThreads = 512;  // 4 warpgroups: builders, builders, epilogue, {UMMA/TMA warp + 3 idle warps that only donate registers}

template <int CZ, int KBS>
__global__ void __launch_bounds__(kOlThreads, 1)
outer_linear_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_pair, const float* pair,
                    float* dst, int residual,
                    const __half* __restrict__ xn16, const __half* __restrict__ w1, const float* __restrict__ u,
                    const float* __restrict__ bias, int N, int ilen) {
  extern __shared__ uint8_t raw[];
  pdl_trigger();
  constexpr int CS = KBS * 64;
  constexpr int NCH = KBS * CZ * 8 / 256;  // 16-byte W1 chunks per builder thread
  constexpr int CPK = NCH / KBS;           // ... per K-block
  constexpr int RB = OlCfg<CZ, KBS>::kRing;        // ring of B-operand K-blocks
  constexpr bool kResTma = OlCfg<CZ, KBS>::kResTma;  // residual rows prefetched by TMA (when shared memory allows)
  constexpr int RP = CZ * 4 / 128;         // 128-byte pieces per pair row
  uint8_t* sm = smem_align1024(raw);
  uint8_t* sA = sm;                        // KBS x 16 KB
  uint8_t* sB = sA + KBS * 16384;          // RB x [CZ x 64]
  uint8_t* sRes = sB + RB * CZ * 128;      // residual pair rows of the next i: RP swizzled TMA boxes [128 j x 32 floats]
  uint8_t* sSl = sRes + (kResTma ? RP * 16384 : 0);  // 4 epilogue warps x 4 KB
  float* sXi = reinterpret_cast<float*>(sSl + 4 * 4096);  // [2][CS]
  float* sUi = sXi + 2 * CS;               // [4][CZ]
  float* sBias = sUi + 4 * CZ;
  uint64_t* bar_a = reinterpret_cast<uint64_t*>(sBias + CZ);
  uint64_t* full = bar_a + 1;              // [RB] K-block built (256 builder arrivals)
  uint64_t* empty = full + RB;             // [RB] K-block consumed (UMMA commit)
  uint64_t* acc_full = empty + RB;         // [2]
  uint64_t* acc_empty = acc_full + 2;      // [2] (128 epilogue arrivals)
  uint64_t* res_full = acc_empty + 2;      // residual rows of row i landed (TMA transaction bytes)
  uint64_t* res_empty = res_full + 1;      // ... and were copied to registers (128 epilogue arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_empty + 1);
  constexpr int TCOLS = 2 * CZ < 32 ? 32 : 2 * CZ;

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int jt = blockIdx.x, b = blockIdx.z;
  const int i0 = blockIdx.y * ilen;
  const int i1 = min(N, i0 + ilen);
  if (t == 0) {
    mbar_init(bar_a, 1);
    for (int q = 0; q < RB; ++q) {
      mbar_init(&full[q], 256);
      mbar_init(&empty[q], 1);
    }
    for (int q = 0; q < 2; ++q) {
      mbar_init(&acc_full[q], 1);
      mbar_init(&acc_empty[q], 128);
    }
    mbar_init(res_full, 1);
    mbar_init(res_empty, 128);
    fence_barrier_init();
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_pair);
  }
  if (warp == 0) tmem_alloc(tmem_slot, TCOLS);
  for (int i = t; i < CZ; i += kOlThreads) sBias[i] = bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // everything above touched only weights / shared memory; the predecessor's output is read below
  const uint32_t tmem = *tmem_slot;

  if (warp < 8) {
    // ------------------------------------------------------------------ builders
    // 64 registers of W1 + working set: stays at the launch allocation of 128 (these warps wait for ring slots most
    // of the time; the registers go to the UMMA warp, whose issue loop is the critical path)
    // this thread's W1 chunks (fixed for the whole kernel): chunk n covers K-block n / CPK,
    // row z = (t >> 3) + 32 * (n % CPK), 16-byte column chunk ch = t & 7
    const int ch = t & 7;
    uint4 wreg[NCH];
#pragma unroll
    for (int n = 0; n < NCH; ++n) {
      const int kb = n / CPK, z = (t >> 3) + 32 * (n % CPK);
      wreg[n] = __ldg(reinterpret_cast<const uint4*>(w1 + (long long)z * CS + kb * 64 + ch * 8));
    }
    // x_i (fp16, the same rounding as the A operand x_j) / u_i are fetched one row ahead into registers so
    // their global latency is off the critical path.  B' = W1 * x_i is formed with HMUL2 straight from the
    // packed W1 registers: no fp32 temporaries (a loop-invariant half->float expansion of W1 would spill).
    constexpr int XPT = CS / 512;  // half2 of x_i per thread
    static_assert(XPT == 1 || CS == 256, "x_i staging assumes c_s in {256, 512}");
    uint32_t xnext = 0;
    float unext = 0.f;
    auto fetch_row = [&](int i) {
      if (i < i1) {
        const uint32_t* xi = reinterpret_cast<const uint32_t*>(xn16 + ((long long)b * N + i) * CS);
        if (t < CS / 2) xnext = __ldg(xi + t);
        if (t < CZ) unext = __ldg(u + ((long long)b * N + i) * CZ + t);
      }
    };
    fetch_row(i0);
    uint32_t ring_it = 0;  // K-blocks generated so far (ring position)
    for (int i = i0; i < i1; ++i) {
      uint32_t* sx = reinterpret_cast<uint32_t*>(sXi) + (i & 1) * (CS / 2);
      if (t < CS / 2) sx[t] = xnext;
      if (t < CZ) sUi[(i & 3) * CZ + t] = unext;  // read by the epilogue of row i (ordered through full[] -> acc_full[])
      fetch_row(i + 1);
      asm volatile("bar.sync 1, 256;" ::: "memory");  // x_i complete; everyone is done with the buffer of row i-2
#pragma unroll
      for (int kb = 0; kb < KBS; ++kb, ++ring_it) {
        const uint32_t slot = ring_it % RB;
        if (ring_it >= RB) mbar_wait(&empty[slot], ((ring_it / RB) - 1) & 1);
        const uint4 xv = *reinterpret_cast<const uint4*>(sx + (kb * 64 + ch * 8) / 2);
        const __half2* x2 = reinterpret_cast<const __half2*>(&xv);
        uint8_t* sBs = sB + slot * (CZ * 128);
#pragma unroll
        for (int n = 0; n < CPK; ++n) {
          const int z = (t >> 3) + 32 * n;
          const __half2* w2 = reinterpret_cast<const __half2*>(&wreg[kb * CPK + n]);
          uint4 o;
          __half2* o2 = reinterpret_cast<__half2*>(&o);
          o2[0] = __hmul2(w2[0], x2[0]);
          o2[1] = __hmul2(w2[1], x2[1]);
          o2[2] = __hmul2(w2[2], x2[2]);
          o2[3] = __hmul2(w2[3], x2[3]);
          *reinterpret_cast<uint4*>(sBs + sw128_offset(z, ch)) = o;
        }
        fence_proxy_async_smem();
        mbar_arrive(&full[slot]);
      }
    }
  } else if (warp < 12) {
    // ------------------------------------------------------------------ epilogue
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    const int q4 = warp & 3;  // TMEM lane quarter
    const int j0 = jt * 128 + q4 * 32;
    const int j = j0 + lane;
    const int rows_valid = N - j0;
    const uint32_t tm_lane = tmem + (static_cast<uint32_t>(q4 * 32) << 16);
    uint8_t* slice = sSl + q4 * 4096;
    float ujb[CZ];  // bias - u_j
    if (j < N) {
      const float4* up = reinterpret_cast<const float4*>(u + ((long long)b * N + j) * CZ);
#pragma unroll
      for (int c = 0; c < CZ / 4; ++c) {
        const float4 v = __ldg(up + c);
        ujb[c * 4] = sBias[c * 4] - v.x;
        ujb[c * 4 + 1] = sBias[c * 4 + 1] - v.y;
        ujb[c * 4 + 2] = sBias[c * 4 + 2] - v.z;
        ujb[c * 4 + 3] = sBias[c * 4 + 3] - v.w;
      }
    } else {
#pragma unroll
      for (int c = 0; c < CZ; ++c) ujb[c] = 0.f;
    }
    for (int i = i0; i < i1; ++i) {
      const long long rowoff = (((long long)b * N + i) * N + j0) * CZ;  // first of this warp's 32 pair rows
      uint4 res[RP][8];
      if (residual && !kResTma) {
#pragma unroll
        for (int p = 0; p < RP; ++p) warp_load_rows128(slice, lane, res[p], pair + rowoff + p * 32, CZ * 4, rows_valid);
      } else if (residual) {
        // this row's residual tile was fetched by the loader warp while the previous row was in flight (a synchronous
        // load here left one row of HBM latency exposed per i: long-scoreboard was 7 of 13 stall cycles per issue)
        mbar_wait(res_full, (i - i0) & 1);
        const int tl = q4 * 32 + lane;
#pragma unroll
        for (int p = 0; p < RP; ++p)
#pragma unroll
          for (int c = 0; c < 8; ++c)
            res[p][c] = *reinterpret_cast<const uint4*>(sRes + p * 16384 + tl * 128 + ((c ^ (tl & 7)) << 4));
        // generic-proxy reads of the stage must be ordered before the async-proxy (TMA) write that refills it: without
        // this fence a few rows per launch picked up the NEXT row's residual (seen only with a warm allocator / L2)
        fence_proxy_async_smem();
        mbar_arrive(res_empty);
      } else {
#pragma unroll
        for (int p = 0; p < RP; ++p)
#pragma unroll
          for (int c = 0; c < 8; ++c) res[p][c] = make_uint4(0, 0, 0, 0);
      }
      mbar_wait(&acc_full[i & 1], ((i - i0) >> 1) & 1);
      tc_fence_after();
      const float* ui = sUi + (i & 3) * CZ;
#pragma unroll
      for (int p = 0; p < RP; ++p) {
        uint32_t acc[32];
        tmem_ld32(tm_lane + (i & 1) * CZ + p * 32, acc);
        tmem_ld_wait();
        if (p == RP - 1) {  // all TMEM reads of this accumulator are done: hand it back
          tc_fence_before();
          mbar_arrive(&acc_empty[i & 1]);
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 uv = *reinterpret_cast<const float4*>(ui + p * 32 + c * 4);
          uint4& r = res[p][c];
          r.x = __float_as_uint(__uint_as_float(r.x) + __uint_as_float(acc[c * 4 + 0]) + uv.x + ujb[p * 32 + c * 4 + 0]);
          r.y = __float_as_uint(__uint_as_float(r.y) + __uint_as_float(acc[c * 4 + 1]) + uv.y + ujb[p * 32 + c * 4 + 1]);
          r.z = __float_as_uint(__uint_as_float(r.z) + __uint_as_float(acc[c * 4 + 2]) + uv.z + ujb[p * 32 + c * 4 + 2]);
          r.w = __float_as_uint(__uint_as_float(r.w) + __uint_as_float(acc[c * 4 + 3]) + uv.w + ujb[p * 32 + c * 4 + 3]);
        }
        warp_store_rows128(slice, lane, res[p], dst + rowoff + p * 32, CZ * 4, rows_valid);
      }
    }
  } else {
    // ------------------------------------------------------------------ TMA + UMMA warp (12), residual loader (13); 14-15 idle
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");  // 8 x 128 + 4 x 200 + 4 x 56 = 16 x 128 (24 made the issue loop spill)
    if (kResTma && warp == 13 && residual) {
      for (int i = i0; i < i1; ++i) {
        if (i > i0) mbar_wait(res_empty, (i - i0 - 1) & 1);  // the epilogue holds row i-1 in registers
        if (elect_one()) {
          mbar_expect_tx(res_full, RP * 16384);
          for (int p = 0; p < RP; ++p) tma_load_4d(sRes + p * 16384, &map_pair, res_full, p * 32, jt * 128, i, b);
        }
        __syncwarp();
      }
    }
    if (warp == 12) {
      if (elect_one()) {
        mbar_expect_tx(bar_a, KBS * 16384);
        for (int kb = 0; kb < KBS; ++kb) tma_load_3d(sA + kb * 16384, &map_x, bar_a, kb * 64, jt * 128, b);
      }
      __syncwarp();
      mbar_wait(bar_a, 0);
      // The issue loop is this kernel's critical path (ncu: the warp is never parked, ~90 dependent instructions and
      // three local-memory reloads per K-block = 940 cycles against 290 of tensor time).  K-blocks fully unrolled: the
      // ring slot (kb % RB) and, for KBS a multiple of 2 RB, the ring phase are compile-time, the descriptors are
      // base + constant in uniform registers.
      static_assert(KBS % 4 == 0 && RB % 4 == 0, "K-blocks are issued in unrolled groups of four ring slots");
      const uint64_t da0 = umma_desc_sw128(smem_u32(sA));
      const uint64_t db0 = umma_desc_sw128(smem_u32(sB));
      const uint32_t idesc = umma_idesc_f16(128, CZ);
      uint32_t kb_total = 0;  // K-blocks issued so far
      for (int i = i0; i < i1; ++i) {
        const int buf = i & 1;
        const uint32_t tacc = tmem + buf * CZ;
        if (i - i0 >= 2) mbar_wait(&acc_empty[buf], (((i - i0) >> 1) - 1) & 1);  // epilogue of row i-2 drained it
#pragma unroll 1
        for (int q = 0; q < KBS / 4; ++q, kb_total += 4) {
          const uint32_t slot0 = kb_total % RB;
          const uint32_t ph = (kb_total / RB) & 1;
          const uint64_t da_q = da0 + static_cast<uint64_t>(q * 4 * (16384 >> 4));
          const uint64_t db_q = db0 + static_cast<uint64_t>(slot0 * ((CZ * 128) >> 4));
          uint64_t* fq = full + slot0;
          uint64_t* eq = empty + slot0;
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4) {
            mbar_wait_spin(fq + s4, ph);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t da = da_q + static_cast<uint64_t>(s4 * (16384 >> 4));
              const uint64_t db = db_q + static_cast<uint64_t>(s4 * ((CZ * 128) >> 4));
#pragma unroll
              for (uint32_t k = 0; k < 4; ++k) umma_f16(tacc, da + 2 * k, db + 2 * k, idesc, (q > 0 || s4 > 0 || k > 0) ? 1u : 0u);
              umma_commit(eq + s4);
              if (s4 == 3 && q == KBS / 4 - 1) umma_commit(&acc_full[buf]);
            }
            __syncwarp();
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

template <int CZ, int KBS>
static int launch_outer_linear(const CUtensorMap& mx, const CUtensorMap& mp, dim3 grid, const float* pair, float* dst, int residual,
                               const __half* xn16, const __half* w1, const float* u, const float* bias, int N, int ilen,
                               cudaStream_t s) {
  constexpr int smem = OlCfg<CZ, KBS>::kSmem;
  static_assert(smem <= 227 * 1024, "outer_linear shared memory budget");
  auto kern = outer_linear_kernel<CZ, KBS>;
  PRD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  PRD_CUDA_OK(launch_pdl(kern, grid, kOlThreads, smem, s, mx, mp, pair, dst, residual, xn16, w1, u, bias, N, ilen));
  PRD_LAUNCHED();
  return 0;
}

int outer_linear(const PairDims& d, int CS, const float* pair, float* dst, int residual, const __half* xn16,
                 const float* xn32, const __half* w1, const float* u, const float* bias, cudaStream_t s) {
  PRD_REQUIRE(CS == 512 || CS == 256, "outer_linear: single_dim %d unsupported (built: 512, 256)", CS);
  PRD_REQUIRE(d.CZ == 64 || d.CZ == 32, "outer_linear: pair_dim %d unsupported (built: 64, 32)", d.CZ);
  const int N = d.N;
  CUtensorMap mx;
  TmaDims t;
  t.size[0] = (uint64_t)CS; t.size[1] = (uint64_t)N; t.size[2] = (uint64_t)d.B; t.size[3] = 1;
  t.stride[0] = (uint64_t)CS * 2; t.stride[1] = (uint64_t)N * CS * 2; t.stride[2] = 0;
  t.box[0] = 64; t.box[1] = 128; t.box[2] = 1; t.box[3] = 1;
  if (make_tensor_map(&mx, xn16, 2, 3, t, true)) return 1;
  // residual rows: pair [B][N][N][CZ] fp32 as (c, j, i, b); one box = [128 j x 32 floats] (128-byte swizzle rows)
  CUtensorMap mp;
  {
    TmaDims r;
    r.size[0] = (uint64_t)d.CZ; r.size[1] = (uint64_t)N; r.size[2] = (uint64_t)N; r.size[3] = (uint64_t)d.B;
    r.stride[0] = (uint64_t)d.CZ * 4; r.stride[1] = (uint64_t)N * d.CZ * 4; r.stride[2] = (uint64_t)N * N * d.CZ * 4;
    r.box[0] = 32; r.box[1] = 128; r.box[2] = 1; r.box[3] = 1;
    if (make_tensor_map(&mp, (residual && pair) ? pair : dst, 4, 4, r, true)) return 1;  // unused without the residual
  }
  const int jtiles = (N + 127) / 128;
  // every CTA re-uses its A tile and its W1 registers for `ilen` rows; one CTA per SM, so pick the number of
  // i-chunks that fills whole waves (least idle SM time in the last wave), at least 8 rows per CTA
  int ichunks = 1;
  {
    const long long per = (long long)jtiles * d.B;
    double best = 1e30;
    for (int c = 1; c <= (N + 7) / 8; ++c) {
      const int len = (N + c - 1) / c;
      const long long ctas = per * ((N + len - 1) / len);
      const long long waves = (ctas + kNumSMs - 1) / kNumSMs;
      const double cost = (double)waves * (len + 3.0);  // rows per wave + ~3 rows of prologue (A tile, W1)
      if (cost < best - 1e-9) {
        best = cost;
        ichunks = c;
      }
    }
  }
  int ilen = (N + ichunks - 1) / ichunks;
  ichunks = (N + ilen - 1) / ilen;
  dim3 grid(jtiles, ichunks, d.B);
  if (d.CZ == 64 && CS == 512) return launch_outer_linear<64, 8>(mx, mp, grid, pair, dst, residual, xn16, w1, u, bias, N, ilen, s);
  if (d.CZ == 64 && CS == 256) return launch_outer_linear<64, 4>(mx, mp, grid, pair, dst, residual, xn16, w1, u, bias, N, ilen, s);
  if (d.CZ == 32 && CS == 512) return launch_outer_linear<32, 8>(mx, mp, grid, pair, dst, residual, xn16, w1, u, bias, N, ilen, s);
  return launch_outer_linear<32, 4>(mx, mp, grid, pair, dst, residual, xn16, w1, u, bias, N, ilen, s);
}

// =========================================================================================
// Per-step pair embedding + OuterProductUpdate.  Rows = flattened (b,i,j).
//   pair = static + m2 * (W_dist rbf(|z_i - z_j|) + beta[b]) + m2 * (W_o (a_i * b_j) + b_o) / (m2 + 1e-3)
// rbf_k(d) = exp(-scale (d - center_k)^2) is generated thread-locally straight into the fp16 A
// operand (dist_dim/64 K-blocks); a_i * b_j (opm_dim/64 K-blocks) likewise.  Two accumulators.
// =========================================================================================
template <int CZ, bool LUT>
__global__ void __launch_bounds__(256, LUT ? 2 : 1)
pair_embed_kernel(const float* __restrict__ pstatic, float* __restrict__ pair, const float* __restrict__ z,
                  const float* __restrict__ mask, const float* __restrict__ beta, const __half* __restrict__ w_dist,
                  int DD, const float* __restrict__ centers, float rbf_scale, const float* __restrict__ opm_a,
                  const float* __restrict__ opm_b, int OD, const __half* __restrict__ w_opm,
                  const float* __restrict__ b_opm, int N, long long R, int flags, const float* __restrict__ lut) {
  // flags: 1 = OuterProductUpdate term only (no distance / time embedding); 2 = do not multiply the
  // OPM term by mask_2d (stand-alone OuterProductUpdate.forward).  pstatic may be NULL (= zeros).
  // lut != NULL: the distance embedding d -> W_dist rbf(d) (a smooth function of ONE scalar) is read from a table
  // (rbf_lut_kernel) and interpolated linearly instead of being rebuilt as 256 exponentials + a K = 256 GEMM per row.
  extern __shared__ uint8_t raw[];
  const bool with_dist = (flags & 1) == 0;
  const bool use_lut = LUT && with_dist && lut != nullptr;  // LUT variant: ~85 KB of shared memory, two CTAs per SM
  const int KBD = (with_dist && !use_lut) ? DD / 64 : 0, KBO = OD / 64;
  const float lut_inv_h = use_lut ? lut[0] : 0.f;
  const float lut_m = use_lut ? lut[1] : 0.f;
  uint8_t* sm = smem_align1024(raw);
  uint8_t* sA1 = sm;                         // rbf, KBD x 16 KB
  uint8_t* sA2 = sA1 + KBD * 16384;          // a_i*b_j, KBO x 16 KB
  uint8_t* sW1 = sA2 + KBO * 16384;          // W_dist hi then lo, each KBD x [CZ x 64]
  uint8_t* sW2 = sW1 + 2 * KBD * CZ * 128;   // W_opm,  KBO x [CZ x 64]
  uint8_t* sSt = sW2 + KBO * CZ * 128;       // one padded row stage
  float* sC = reinterpret_cast<float*>(sSt + RowStage<CZ>::kBytes);  // centers [DD]
  float* sBo = sC + DD;                                            // b_opm [CZ]
  uint64_t* full = reinterpret_cast<uint64_t*>(sBo + CZ);
  uint64_t* mma_bar = full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);
  constexpr int TCOLS = 2 * CZ;

  // two threads per row: thread (t, half) generates half of the radial-basis / outer-product columns of row t and
  // handles half of the output channels (the kernel is instruction bound: ~1500 instructions per row)
  const int tid = threadIdx.x, t = tid & 127, half = tid >> 7, warp = t >> 5;
  if (tid == 0) {
    mbar_init(full, kTileRows);
    mbar_init(mma_bar, 1);
    fence_barrier_init();
  }
  if (tid < 32) tmem_alloc(tmem_slot, TCOLS);
  if (KBD > 0) {
    load_weight_kblocks(sW1, w_dist, CZ, DD, DD, tid, 256);
    load_weight_kblocks(sW1 + KBD * CZ * 128, w_dist + CZ * DD, CZ, DD, DD, tid, 256);
  }
  load_weight_kblocks(sW2, w_opm, CZ, OD, OD, tid, 256);
  if (KBD > 0)
    for (int i = tid; i < DD; i += 256) sC[i] = centers[i];
  for (int i = tid; i < CZ; i += 256) sBo[i] = b_opm[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const long long NN = (long long)N * N;

  const long long num_tiles = (R + kTileRows - 1) / kTileRows;
  uint32_t mma_phase = 0;
  int it = 0;
  for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
    const long long r = tile * kTileRows + t;
    const bool valid = r < R;
    if (half == 0) bulk_wait_read0();  // previous tile's bulk store has finished reading the stage
    if (half == 0) issue_row_load<CZ>(sSt, t, pstatic + r * CZ, valid && pstatic != nullptr, full);
    int b = 0, i = 0, j = 0;
    if (valid) {
      b = static_cast<int>(r / NN);
      const int rem = static_cast<int>(r - (long long)b * NN);
      i = rem / N;
      j = rem - i * N;
    }
    const float* zi = z + ((long long)b * N + i) * 3;
    const float* zj = z + ((long long)b * N + j) * 3;
    const float dx = zi[0] - zj[0], dy = zi[1] - zj[1], dz = zi[2] - zj[2];
    const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
    const float m2 = valid ? mask[(long long)b * N + i] * mask[(long long)b * N + j] : 0.f;
    // table rows m, m+1 around dist / h (this thread's 32 output channels); issued now, interpolated in the epilogue
    const bool my_out = (CZ / 32 == 2) || (half == 0);
    float lv[LUT ? 32 : 1];
    if (LUT && use_lut && my_out) {
      const float u = fminf(dist * lut_inv_h, lut_m);
      const int mi = min(static_cast<int>(u), static_cast<int>(lut_m) - 1);
      const float lfrac = u - static_cast<float>(mi);
      const float4* r0 = reinterpret_cast<const float4*>(lut + (long long)(mi + 1) * CZ + (CZ / 32 == 2 ? 32 * half : 0));
      const float4* r1 = r0 + CZ / 4;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 a = __ldg(r0 + q), c = __ldg(r1 + q);
        lv[4 * q + 0] = fmaf(lfrac, c.x - a.x, a.x);
        lv[4 * q + 1] = fmaf(lfrac, c.y - a.y, a.y);
        lv[4 * q + 2] = fmaf(lfrac, c.z - a.z, a.z);
        lv[4 * q + 3] = fmaf(lfrac, c.w - a.w, a.w);
      }
    }
    // radial basis -> A1
#pragma unroll 1
    for (int k0 = half * KBD * 32; k0 < (half + 1) * KBD * 32; k0 += 32) {
      float v[32];
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const float dd = dist - sC[k0 + q];
        v[q] = __expf(-rbf_scale * dd * dd);
      }
      store_a_cols32(sA1, t, k0, v);
    }
    // outer product a_i * b_j -> A2
    const float* ap = opm_a + ((long long)b * N + i) * OD;
    const float* bp = opm_b + ((long long)b * N + j) * OD;
#pragma unroll 1
    for (int k0 = half * (OD / 2); k0 < (half + 1) * (OD / 2); k0 += 32) {
      float v[32];
#pragma unroll
      for (int q = 0; q < 32; q += 4) {
        const float4 av = __ldg(reinterpret_cast<const float4*>(ap + k0 + q));
        const float4 bv = __ldg(reinterpret_cast<const float4*>(bp + k0 + q));
        v[q] = av.x * bv.x; v[q + 1] = av.y * bv.y; v[q + 2] = av.z * bv.z; v[q + 3] = av.w * bv.w;
      }
      store_a_cols32(sA2, t, k0, v);
    }
    sync_before_mma();
    if (tid < 32) {  // warp-uniform issue: UMMA operands stay in uniform registers
      tc_fence_after();
      if (elect_one()) {
        if (KBD > 0) {
          umma_multi(tmem, smem_u32(sA1), smem_u32(sW1), KBD, CZ * 128, umma_idesc_f16(128, CZ), false);
          umma_multi(tmem, smem_u32(sA1), smem_u32(sW1 + KBD * CZ * 128), KBD, CZ * 128, umma_idesc_f16(128, CZ), true);
        }
        umma_multi(tmem + CZ, smem_u32(sA2), smem_u32(sW2), KBO, CZ * 128, umma_idesc_f16(128, CZ), false);
        umma_commit(mma_bar);
      }
      __syncwarp();
    }
    mbar_wait(full, it & 1);
    mbar_wait(mma_bar, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    float* my = stage_row<CZ>(sSt, t);
    const float* bt = beta + (long long)b * CZ;
    const float inv_norm = 1.0f / (m2 + 1e-3f);
    const float m2o = (flags & 2) ? 1.0f : m2;
    const bool have_static = pstatic != nullptr;
    // output channels [32 half, 32 half + 32) (pair_dim 64) / all channels by half 0 (pair_dim 32)
#pragma unroll
    for (int c = 0; c < CZ / 32; ++c) {
      if (CZ / 32 == 2 ? (c != half) : (half != 0)) continue;
      uint32_t a1[32], a2[32];
      if (KBD > 0) tmem_ld32(tm_lane + c * 32, a1);
      tmem_ld32(tm_lane + CZ + c * 32, a2);
      tmem_ld_wait();
      if (LUT && use_lut) {
#pragma unroll
        for (int q = 0; q < 32; ++q) a1[q] = __float_as_uint(lv[LUT ? q : 0]);
      }
#pragma unroll
      for (int q = 0; q < 32; q += 4) {
        float4 x = have_static ? *reinterpret_cast<float4*>(my + c * 32 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int cc = c * 32 + q + e;
          o[e] = m2o * ((__uint_as_float(a2[q + e]) + sBo[cc]) * inv_norm);
          if (with_dist) o[e] += m2 * (__uint_as_float(a1[q + e]) + __ldg(bt + cc));
        }
        x.x += o[0]; x.y += o[1]; x.z += o[2]; x.w += o[3];
        *reinterpret_cast<float4*>(my + c * 32 + q) = x;
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();  // both halves of every row are staged; TMEM / A tiles are free for the next tile
    if (half == 0) {
      if (valid) bulk_s2g(pair + r * CZ, my, CZ * 4);
      bulk_commit();
    }
  }
  bulk_wait0();
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, TCOLS);
}

// -----------------------------------------------------------------------------------------
// pair_dim 64, OPM hidden 128, N % 128 == 0, table-driven distance embedding: the remaining cost of the kernel above is load
// latency (ncu: issue slots 20 % busy, long-scoreboard 12 cycles per issue on the per-thread opm_a / opm_b / z / table loads).
// Here a tile is 128 consecutive j of one (b, i) row and a CTA keeps seeing the SAME (b, j-tile) for long runs (tiles are
// strided by the grid size), so the opm_b rows of the j-tile (64 KB fp32, four swizzled TMA boxes), z_j and mask_j are kept
// resident and re-fetched only when (b, j-tile) changes; a_i of the NEXT tile arrives by a 512-byte bulk copy a tile ahead;
// the table rows are requested at tile start and consumed in the epilogue.  One accumulator (OPM), 256 threads = two per row.
// -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
pair_embed_lut_kernel(const __grid_constant__ CUtensorMap map_b, const float* __restrict__ pstatic, float* __restrict__ pair,
                      const float* __restrict__ z, const float* __restrict__ mask, const float* __restrict__ beta,
                      const float* __restrict__ opm_a, const __half* __restrict__ w_opm, const float* __restrict__ b_opm,
                      int N, long long num_tiles, const float* __restrict__ lut,
                      const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out) {
  // The static rows of a tile arrive as ONE TMA tile (two swizzled boxes [128 rows][32 channels], thread (t, half) owns row t
  // of box half) and the finished rows leave as two tile stores, all issued by thread 0: 256-byte bulk copies per thread
  // serialise lane by lane on the uniform datapath.
  constexpr int CZ = 64, OD = 128, KBO = 2;
  extern __shared__ uint8_t raw[];
  uint8_t* sm = smem_align1024(raw);
  uint8_t* sA2 = sm;                         // a_i * b_j, 2 x 16 KB
  uint8_t* sBt = sA2 + KBO * 16384;          // opm_b rows of the j-tile: 4 swizzled boxes [128 x 32 floats]
  uint8_t* sW2 = sBt + 4 * 16384;            // W_opm, 2 x [CZ x 64]
  uint8_t* sSt = sW2 + KBO * CZ * 128;       // padded row stage (static rows in, output rows out)
  float* sAi = reinterpret_cast<float*>(sSt + RowStage<CZ>::kBytes);  // [2][OD]
  float* sBo = sAi + 2 * OD;                                          // b_opm [CZ]
  uint64_t* full = reinterpret_cast<uint64_t*>(sBo + CZ);
  uint64_t* mma_bar = full + 1;
  uint64_t* bt_bar = full + 2;
  uint64_t* ai_bar = full + 3;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 5);

  const int tid = threadIdx.x, t = tid & 127, half = tid >> 7, warp = t >> 5;
  if (tid == 0) {
    mbar_init(full, 1);
    mbar_init(mma_bar, 1);
    mbar_init(bt_bar, 1);
    mbar_init(&ai_bar[0], 1);
    mbar_init(&ai_bar[1], 1);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_in);
    tma_prefetch_desc(&map_out);
    fence_barrier_init();
  }
  if (tid < 32) tmem_alloc(tmem_slot, CZ);
  load_weight_kblocks(sW2, w_opm, CZ, OD, OD, tid, 256);
  for (int i = tid; i < CZ; i += 256) sBo[i] = b_opm[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const float lut_inv_h = lut[0], lut_m = lut[1];
  const int tpr = N / kTileRows;  // tiles per (b, i) row
  auto decode = [&](long long tl, int& b, int& i, int& jt) {
    const long long bi = tl / tpr;
    jt = static_cast<int>(tl - bi * tpr);
    b = static_cast<int>(bi / N);
    i = static_cast<int>(bi - (long long)b * N);
  };
  int cur_b = -1, cur_jt = -1;
  uint32_t bt_phase = 0, mma_phase = 0;
  float zj0 = 0.f, zj1 = 0.f, zj2 = 0.f, mj = 0.f;
  int it = 0;
  if (blockIdx.x < num_tiles && tid == 0) {  // a_i of the first tile
    int b, i, jt;
    decode(blockIdx.x, b, i, jt);
    mbar_expect_tx(&ai_bar[0], OD * 4);
    bulk_g2s(sAi, opm_a + ((long long)b * N + i) * OD, OD * 4, &ai_bar[0]);
  }
  for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
    int b, i, jt;
    decode(tile, b, i, jt);
    const long long r = tile * kTileRows + t;
    if (tid == 0) {
      bulk_wait_read0();  // the previous tile's stores have finished reading the stage (CTA barriers order the rest behind this)
      if (pstatic != nullptr) {
        mbar_expect_tx(full, 32768);
        tma_load_2d(sSt, &map_in, full, 0, static_cast<int>(tile * kTileRows));
        tma_load_2d(sSt + 16384, &map_in, full, 32, static_cast<int>(tile * kTileRows));
      }
    }
    if (tid == 0 && tile + gridDim.x < num_tiles) {  // a_i of the next tile (its buffer was last read two tiles ago)
      int nb, ni, njt;
      decode(tile + gridDim.x, nb, ni, njt);
      mbar_expect_tx(&ai_bar[(it + 1) & 1], OD * 4);
      bulk_g2s(sAi + ((it + 1) & 1) * OD, opm_a + ((long long)nb * N + ni) * OD, OD * 4, &ai_bar[(it + 1) & 1]);
    }
    if (b != cur_b || jt != cur_jt) {  // CTA-uniform: new j-tile -> resident operands (rare)
      cur_b = b;
      cur_jt = jt;
      // every thread is past the previous tile's reads of sBt (end-of-tile barrier, which follows a proxy fence)
      if (tid == 0) {
        mbar_expect_tx(bt_bar, 4 * 16384);
        for (int q = 0; q < 4; ++q) tma_load_3d(sBt + q * 16384, &map_b, bt_bar, q * 32, jt * kTileRows, b);
      }
      const float* zj = z + ((long long)b * N + jt * kTileRows + t) * 3;
      zj0 = zj[0];
      zj1 = zj[1];
      zj2 = zj[2];
      mj = mask[(long long)b * N + jt * kTileRows + t];
      mbar_wait(bt_bar, bt_phase);
      bt_phase ^= 1;
    }
    const float* zi = z + ((long long)b * N + i) * 3;
    const float dx = __ldg(zi) - zj0, dy = __ldg(zi + 1) - zj1, dz = __ldg(zi + 2) - zj2;
    const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
    const float m2 = __ldg(mask + (long long)b * N + i) * mj;
    // table rows around dist / h for this thread's 32 output channels: requested now, interpolated in the epilogue
    float4 l0[8], l1[8];
    const float u = fminf(dist * lut_inv_h, lut_m);
    const int mi = min(static_cast<int>(u), static_cast<int>(lut_m) - 1);
    const float lfrac = u - static_cast<float>(mi);
    {
      const float4* r0 = reinterpret_cast<const float4*>(lut + (long long)(mi + 1) * CZ + 32 * half);
      const float4* r1 = r0 + CZ / 4;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        l0[q] = __ldg(r0 + q);
        l1[q] = __ldg(r1 + q);
      }
    }
    // outer product a_i * b_j -> A2: this thread's K-block (= half): 64 products
    mbar_wait(&ai_bar[it & 1], (it >> 1) & 1);
    {
      const float* ai = sAi + (it & 1) * OD + half * 64;
#pragma unroll
      for (int k0 = 0; k0 < 64; k0 += 32) {
        float v[32];
        const uint8_t* box = sBt + (half * 2 + (k0 >> 5)) * 16384 + t * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 bv = *reinterpret_cast<const float4*>(box + ((c ^ (t & 7)) << 4));
          const float4 av = *reinterpret_cast<const float4*>(ai + k0 + 4 * c);
          v[4 * c] = av.x * bv.x; v[4 * c + 1] = av.y * bv.y; v[4 * c + 2] = av.z * bv.z; v[4 * c + 3] = av.w * bv.w;
        }
        store_a_cols32(sA2, t, half * 64 + k0, v);
      }
    }
    sync_before_mma();
    if (tid < 32) {  // warp-uniform issue: UMMA operands stay in uniform registers
      tc_fence_after();
      if (elect_one()) {
        umma_multi(tmem, smem_u32(sA2), smem_u32(sW2), KBO, CZ * 128, umma_idesc_f16(128, CZ), false);
        umma_commit(mma_bar);
      }
      __syncwarp();
    }
    if (pstatic != nullptr) mbar_wait(full, it & 1);
    mbar_wait(mma_bar, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    uint8_t* my = sSt + half * 16384 + t * 128;  // row t of box `half`; 16-byte chunk q sits at (q ^ (t & 7)) << 4
    const float* bt = beta + (long long)b * CZ + 32 * half;
    const float* bo = sBo + 32 * half;
    const float inv_norm = m2 / (m2 + 1e-3f);  // mask_2d * (.) / (mask_2d + 1e-3)  (AF2_modules.py:539-543, modules.py:395)
    const bool have_static = pstatic != nullptr;
    {
      uint32_t a2[32];
      tmem_ld32(tm_lane + 32 * half, a2);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4* px = reinterpret_cast<float4*>(my + ((q ^ (t & 7)) << 4));
        float4 x = have_static ? *px : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 be = __ldg(reinterpret_cast<const float4*>(bt + 4 * q));
        x.x += (__uint_as_float(a2[4 * q + 0]) + bo[4 * q + 0]) * inv_norm + m2 * (fmaf(lfrac, l1[q].x - l0[q].x, l0[q].x) + be.x);
        x.y += (__uint_as_float(a2[4 * q + 1]) + bo[4 * q + 1]) * inv_norm + m2 * (fmaf(lfrac, l1[q].y - l0[q].y, l0[q].y) + be.y);
        x.z += (__uint_as_float(a2[4 * q + 2]) + bo[4 * q + 2]) * inv_norm + m2 * (fmaf(lfrac, l1[q].z - l0[q].z, l0[q].z) + be.z);
        x.w += (__uint_as_float(a2[4 * q + 3]) + bo[4 * q + 3]) * inv_norm + m2 * (fmaf(lfrac, l1[q].w - l0[q].w, l0[q].w) + be.w);
        *px = x;
      }
    }
    fence_proxy_async_smem();  // output rows -> tile store; also orders this tile's reads of sAi / sBt before later refills
    tc_fence_before();
    __syncthreads();  // both halves of every row are staged; TMEM / A tiles are free for the next tile
    if (tid == 0) {
      tma_store_2d(&map_out, sSt, 0, static_cast<int>(tile * kTileRows));
      tma_store_2d(&map_out, sSt + 16384, 32, static_cast<int>(tile * kTileRows));
      bulk_commit();
    }
  }
  bulk_wait0();
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, CZ);
}

int pair_embed_dynamic(const PairDims& d, const float* pair_static, float* pair, const float* z, const float* mask,
                       const float* beta, const __half* w_dist, int dist_dim, const float* centers, float rbf_scale,
                       const float* opm_a, const float* opm_b, int opm_dim, const __half* w_opm, const float* b_opm,
                       int flags, const float* lut, cudaStream_t s) {
  PRD_REQUIRE(dist_dim % 64 == 0 && opm_dim % 64 == 0, "pair_embed: dist_dim %d / opm hidden %d must be multiples of 64",
              dist_dim, opm_dim);
  const long long R = (long long)d.B * d.N * d.N;
  const long long tiles = (R + kTileRows - 1) / kTileRows;
  const int KBD = ((flags & 1) == 0 && lut == nullptr) ? dist_dim / 64 : 0, KBO = opm_dim / 64;
  const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
  static const bool resident_off = getenv("PRD_PAIR_EMBED_RESIDENT") && getenv("PRD_PAIR_EMBED_RESIDENT")[0] == '0';  // A/B timing
  if (d.CZ == 64 && opm_dim == 128 && flags == 0 && lut != nullptr && d.N % kTileRows == 0 && !resident_off && R < 0x7fffffffLL) {
    // resident-operand kernel: opm_b [B][N][128] fp32 as (k, j, b); box = [32 k][128 j] with the 128-byte swizzle
    CUtensorMap mb;
    TmaDims t;
    t.size[0] = 128; t.size[1] = (uint64_t)d.N; t.size[2] = (uint64_t)d.B; t.size[3] = 1;
    t.stride[0] = 128 * 4; t.stride[1] = (uint64_t)d.N * 128 * 4; t.stride[2] = 0;
    t.box[0] = 32; t.box[1] = 128; t.box[2] = 1; t.box[3] = 1;
    if (make_tensor_map(&mb, opm_b, 4, 3, t, true)) return 1;
    // static rows in / finished rows out: the 2-D tensors [R, 64], box = [32 channels][128 rows]
    CUtensorMap m_in, m_out;
    {
      TmaDims tr;
      tr.size[0] = 64; tr.size[1] = (uint64_t)R; tr.size[2] = 1; tr.size[3] = 1;
      tr.stride[0] = 256; tr.stride[1] = (uint64_t)R * 256; tr.stride[2] = tr.stride[1];
      tr.box[0] = 32; tr.box[1] = kTileRows; tr.box[2] = 1; tr.box[3] = 1;
      if (make_tensor_map(&m_out, pair, 4, 2, tr, true)) return 1;
      if (make_tensor_map(&m_in, pair_static != nullptr ? pair_static : pair, 4, 2, tr, true)) return 1;
    }
    constexpr int smem = 1024 + 2 * 16384 + 4 * 16384 + 2 * 64 * 128 + RowStage<64>::kBytes + (2 * 128 + 64) * 4 + 64;
    PRD_CUDA_OK(cudaFuncSetAttribute(pair_embed_lut_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    pair_embed_lut_kernel<<<grid, 256, smem, s>>>(mb, pair_static, pair, z, mask, beta, opm_a, w_opm, b_opm, d.N, tiles, lut, m_in, m_out);
    PRD_LAUNCHED();
    return 0;
  }
  if (d.CZ == 64) {
    constexpr int CZ = 64;
    const int smem = 1024 + (KBD + KBO) * 16384 + (2 * KBD + KBO) * CZ * 128 + RowStage<CZ>::kBytes + (dist_dim + CZ) * 4 + 64;
    PRD_REQUIRE(smem <= 227 * 1024, "pair_embed: shared memory %d B exceeds 227 KB", smem);
    const bool lut_variant = (flags & 1) == 0 && lut != nullptr;
    auto kern = lut_variant ? pair_embed_kernel<CZ, true> : pair_embed_kernel<CZ, false>;
    PRD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<lut_variant ? (int)(tiles < 2 * kNumSMs ? tiles : 2 * kNumSMs) : grid, 256, smem, s>>>(pair_static, pair, z, mask, beta, w_dist, dist_dim, centers, rbf_scale, opm_a, opm_b,
                                 opm_dim, w_opm, b_opm, d.N, R, flags, lut);
  } else if (d.CZ == 32) {
    constexpr int CZ = 32;
    const int smem = 1024 + (KBD + KBO) * 16384 + (2 * KBD + KBO) * CZ * 128 + RowStage<CZ>::kBytes + (dist_dim + CZ) * 4 + 64;
    PRD_REQUIRE(smem <= 227 * 1024, "pair_embed: shared memory %d B exceeds 227 KB", smem);
    const bool lut_variant = (flags & 1) == 0 && lut != nullptr;
    auto kern = lut_variant ? pair_embed_kernel<CZ, true> : pair_embed_kernel<CZ, false>;
    PRD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<lut_variant ? (int)(tiles < 2 * kNumSMs ? tiles : 2 * kNumSMs) : grid, 256, smem, s>>>(pair_static, pair, z, mask, beta, w_dist, dist_dim, centers, rbf_scale, opm_a, opm_b,
                                 opm_dim, w_opm, b_opm, d.N, R, flags, lut);
  } else {
    set_error("pair_embed: unsupported pair_dim %d", d.CZ);
    return 1;
  }
  PRD_LAUNCHED();
  return 0;
}

// Table of the distance embedding: lut[(1 + m) * CZ + c] = sum_k W_dist[c][k] exp(-scale (m h - center_k)^2), m = 0 .. M, in
// fp32 from the fp32 weights; header row: lut[0] = 1 / h, lut[1] = M.  Linear interpolation between rows is exact to
// h^2 / 8 * |f''| ~ 3e-6 relative at M = 8192 over [0, max center + 0.52] (beyond that every term is below 1e-15).
__global__ void rbf_lut_kernel(const float* __restrict__ w_dist, const float* __restrict__ centers, float scale, int CZ,
                               int DD, float h, int M, float* __restrict__ lut) {
  extern __shared__ float sE[];  // [DD]
  const int m = blockIdx.x;
  const float d = h * static_cast<float>(m);
  for (int k = threadIdx.x; k < DD; k += blockDim.x) {
    const float dd = d - centers[k];
    sE[k] = expf(-scale * dd * dd);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < CZ; c += blockDim.x) {
    const float* wr = w_dist + (long long)c * DD;
    float acc = 0.f;
    for (int k = 0; k < DD; ++k) acc = fmaf(wr[k], sE[k], acc);
    lut[(long long)(1 + m) * CZ + c] = acc;
    if (m == 0) lut[c] = c == 0 ? 1.0f / h : (c == 1 ? static_cast<float>(M) : 0.f);
  }
}

int rbf_lut_build(int CZ, int DD, const float* w_dist, const float* centers, float scale, float d_max, int M, float* lut,
                  cudaStream_t s) {
  PRD_REQUIRE(M >= 2 && d_max > 0.f && CZ >= 2, "rbf_lut_build: bad arguments");
  rbf_lut_kernel<<<M + 1, 64, DD * sizeof(float), s>>>(w_dist, centers, scale, CZ, DD, d_max / M, M, lut);
  PRD_LAUNCHED();
  return 0;
}

// =========================================================================================
// Coordinate head.  CTA = (b, i); loops over j-tiles.
//   p = LN(0.5 (pair[b,i,j] + pair[b,j,i]));  w = w2 . relu(W1 p + b1)
//   eps_raw[b,i,:] = sum_j m_i m_j w (z_i - z_j) rsqrt(|z_i - z_j|^2 + 1e-4)
// =========================================================================================
template <int CZ>
__global__ void __launch_bounds__(128, 2)
coord_head_kernel(const float* __restrict__ pair, const float* __restrict__ z, const float* __restrict__ mask,
                  const __half* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                  float* __restrict__ eps_raw, int N, const __grid_constant__ CUtensorMap map_rows,
                  const __grid_constant__ CUtensorMap map_cols, int row_tma) {
  // row_tma (pair_dim 64): the 128 rows (i, j0..) and the 128 transposed rows (j0.., i) of a tile arrive as two TMA tiles
  // issued by one thread instead of 256 per-thread bulk copies (which serialise lane by lane on the uniform datapath)
  extern __shared__ uint8_t raw[];
  uint8_t* sm = smem_align1024(raw);
  uint8_t* sA = sm;
  uint8_t* sW = sA + 16384;          // W1 hi then lo
  uint8_t* sSt = sW + 2 * CZ * 128;  // two stages: row (i,j) and transposed row (j,i)
  float* sB1 = reinterpret_cast<float*>(sSt + 2 * RowStage<CZ>::kBytes);
  float* sW2 = sB1 + CZ;
  float* sRed = sW2 + CZ;  // [4 warps][3]
  uint64_t* full = reinterpret_cast<uint64_t*>(sRed + 16);
  uint64_t* mma_bar = full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);
  constexpr int TCOLS = CZ < 32 ? 32 : CZ;

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int b = blockIdx.x / N, i = blockIdx.x % N;
  if (t == 0) {
    mbar_init(full, row_tma ? 1 : 2 * kTileRows);
    mbar_init(mma_bar, 1);
    if (row_tma) {
      tma_prefetch_desc(&map_rows);
      tma_prefetch_desc(&map_cols);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, TCOLS);
  load_weight_kblocks(sW, w1, CZ, CZ, CZ, t, 128);
  load_weight_kblocks(sW + CZ * 128, w1 + CZ * CZ, CZ, CZ, CZ, t, 128);
  for (int q = t; q < CZ; q += 128) {
    sB1[q] = b1[q];
    sW2[q] = w2[q];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);

  const float mi = mask[(long long)b * N + i];
  const float zix = z[((long long)b * N + i) * 3 + 0], ziy = z[((long long)b * N + i) * 3 + 1],
              ziz = z[((long long)b * N + i) * 3 + 2];
  float ax = 0.f, ay = 0.f, az = 0.f;
  uint32_t mma_phase = 0;
  const int ntiles = (N + 127) / 128;
  for (int jt = 0; jt < ntiles; ++jt) {
    const int j = jt * 128 + t;
    const bool valid = j < N;
    if (row_tma) {
      if (t == 0) {  // every thread is past its reads of the previous tile (end-of-tile barrier)
        mbar_expect_tx(full, 65536);
        const int r0 = static_cast<int>(((long long)b * N + i) * N + jt * 128);  // rows past the tensor's end are zero-filled
        uint8_t* sT = sSt + RowStage<CZ>::kBytes;
        for (int hh = 0; hh < 2; ++hh) {
          tma_load_2d(sSt + hh * 16384, &map_rows, full, hh * 32, r0);
          tma_load_4d(sT + hh * 16384, &map_cols, full, hh * 32, i, jt * 128, b);
        }
      }
    } else {
      issue_row_load<CZ>(sSt, t, pair + (((long long)b * N + i) * N + j) * CZ, valid, full);
      issue_row_load<CZ>(sSt + RowStage<CZ>::kBytes, t, pair + (((long long)b * N + j) * N + i) * CZ, valid, full);
    }
    mbar_wait(full, jt & 1);
    {
      float x[CZ], y[CZ];
      if (valid && row_tma) {
        if constexpr (CZ == 64) {
          read_row_tma64(sSt, t, x);
          read_row_tma64(sSt + RowStage<CZ>::kBytes, t, y);
        }
#pragma unroll
        for (int q = 0; q < CZ; ++q) x[q] = 0.5f * (x[q] + y[q]);
      } else if (valid) {
        read_row<CZ>(stage_row<CZ>(sSt, t), x);
        read_row<CZ>(stage_row<CZ>(sSt + RowStage<CZ>::kBytes, t), y);
#pragma unroll
        for (int q = 0; q < CZ; ++q) x[q] = 0.5f * (x[q] + y[q]);
      } else {
#pragma unroll
        for (int q = 0; q < CZ; ++q) x[q] = 0.f;
      }
      layernorm_inplace<CZ>(x);
      store_a_row<CZ>(sA, t, x);
    }
    sync_before_mma();
    if (t < 32) {  // warp-uniform issue: UMMA operands stay in uniform registers
      tc_fence_after();
      if (elect_one()) {
        umma_multi(tmem, smem_u32(sA), smem_u32(sW), 1, CZ * 128, umma_idesc_f16(128, CZ), false);
        umma_multi(tmem, smem_u32(sA), smem_u32(sW + CZ * 128), 1, CZ * 128, umma_idesc_f16(128, CZ), true);
        umma_commit(mma_bar);
      }
      __syncwarp();
    }
    mbar_wait(mma_bar, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    float w = 0.f;
#pragma unroll
    for (int c = 0; c < CZ / 32; ++c) {
      uint32_t acc[32];
      tmem_ld32(tm_lane + c * 32, acc);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 32; ++q) w += fmaxf(__uint_as_float(acc[q]) + sB1[c * 32 + q], 0.f) * sW2[c * 32 + q];
    }
    if (valid) {
      const float* zj = z + ((long long)b * N + j) * 3;
      const float dx = zix - zj[0], dy = ziy - zj[1], dz = ziz - zj[2];
      const float rn = rsqrtf(dx * dx + dy * dy + dz * dz + 1e-4f);
      const float f = mi * mask[(long long)b * N + j] * w * rn;
      ax += f * dx;
      ay += f * dy;
      az += f * dz;
    }
    tc_fence_before();
    __syncthreads();
  }
  ax = warp_sum(ax);
  ay = warp_sum(ay);
  az = warp_sum(az);
  if (lane == 0) {
    sRed[warp * 3 + 0] = ax;
    sRed[warp * 3 + 1] = ay;
    sRed[warp * 3 + 2] = az;
  }
  __syncthreads();
  if (t < 3) eps_raw[((long long)b * N + i) * 3 + t] = sRed[t] + sRed[3 + t] + sRed[6 + t] + sRed[9 + t];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

int coord_head(const PairDims& d, const float* pair, const float* z, const float* mask, const __half* w1,
               const float* b1, const float* w2, float* eps_raw, cudaStream_t s) {
  const int grid = d.B * d.N;
  if (d.CZ == 64) {
    constexpr int CZ = 64;
    constexpr int smem = 1024 + 16384 + 2 * CZ * 128 + 2 * RowStage<CZ>::kBytes + (2 * CZ + 16) * 4 + 64;
    static_assert(RowStage<CZ>::kBytes % 1024 == 0 && RowStage<CZ>::kBytes >= 32768, "coord_head: the row stages double as TMA tiles");
    auto kern = coord_head_kernel<CZ>;
    PRD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    // rows (b, i, j0..): the 2-D tensor [R, 64]; transposed rows (b, j0.., i): (channel, i, j, b) with the box along j
    const long long R = (long long)d.B * d.N * d.N;
    const int row_tma = (R < 0x7fffffffLL && (reinterpret_cast<uintptr_t>(pair) & 15) == 0) ? 1 : 0;
    CUtensorMap m_rows, m_cols;
    {
      TmaDims tr;
      tr.size[0] = 64; tr.size[1] = (uint64_t)R; tr.size[2] = 1; tr.size[3] = 1;
      tr.stride[0] = 256; tr.stride[1] = (uint64_t)R * 256; tr.stride[2] = tr.stride[1];
      tr.box[0] = 32; tr.box[1] = kTileRows; tr.box[2] = 1; tr.box[3] = 1;
      if (make_tensor_map(&m_rows, pair, 4, 2, tr, true)) return 1;
      tr.size[1] = (uint64_t)d.N; tr.size[2] = (uint64_t)d.N; tr.size[3] = (uint64_t)d.B;
      tr.stride[1] = (uint64_t)d.N * 256; tr.stride[2] = (uint64_t)d.N * d.N * 256;
      tr.box[1] = 1; tr.box[2] = kTileRows;
      if (make_tensor_map(&m_cols, pair, 4, 4, tr, true)) return 1;
    }
    kern<<<grid, 128, smem, s>>>(pair, z, mask, w1, b1, w2, eps_raw, d.N, m_rows, m_cols, row_tma);
  } else if (d.CZ == 32) {
    constexpr int CZ = 32;
    constexpr int smem = 1024 + 16384 + 2 * CZ * 128 + 2 * RowStage<CZ>::kBytes + (2 * CZ + 16) * 4 + 64;
    auto kern = coord_head_kernel<CZ>;
    PRD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUtensorMap none{};
    kern<<<grid, 128, smem, s>>>(pair, z, mask, w1, b1, w2, eps_raw, d.N, none, none, 0);
  } else {
    set_error("coord_head: unsupported pair_dim %d", d.CZ);
    return 1;
  }
  PRD_LAUNCHED();
  return 0;
}

}  // namespace prd
