// Triangle attention core (modules.py:185-225 applied to every row / column of the pair tensor,
// modules.py:236-243): flash-style gated attention, 4 heads x 16 channels, key mask with the
// reference's finite fill value (-2^15), online softmax, S = QK^T and O = PV on tcgen05.
//
// Persistent kernel, one CTA per SM, 12 warps:
//   warps 0-3 / 4-7   softmax groups A / B: each group owns one work unit (sequence, 128-query tile) at a
//                     time; thread t of a group owns query row t (= TMEM lane t)
//   warps 8 / 9       UMMA issue for group A / B        warps 10 / 11   TMA loads for group A / B
// The exp2 (MUFU) pipe is the binding unit of this kernel (16 exp2/clk/SM against 128 scores per row per
// item).  The two groups hand a per-scheduler token back and forth (mbarriers tok[g][w]) so that on every
// SM sub-partition exactly one warp is in its exp2 pass while its partner does everything else (row max,
// O read-back, epilogue, waiting for UMMAs): without the token the two drift into lock step, share the
// pipe during the exp2 pass and leave it idle during the rest (measured: 55 % MUFU utilisation).
//
// An "item" is (key tile kt, head h), G = global item counter of the group across its units:
//   S_G   = Q_h K_h^T        one UMMA, M=128 N=128 K=16 -> TMEM buffer G&1 (two 128-column buffers / group)
//   pass A  row max of S_G (FMNMX3), exp2 domain: q is pre-scaled by log2(e)/sqrt(c) in triattn_proj
//   pass B  P = exp2(S_G - m) (FADD2, MUFU, packed row sum) -> fp16 A operand in shared memory
//           (SWIZZLE_128B, 2 K-blocks of 64 keys)
//   O_G   = P V_h            2 x 4 UMMAs, M=128 N=16 K=16, issued per 64-key half as soon as it is
//                            written; the product lands in columns [0,16) of S_G's own (now consumed)
//                            TMEM buffer, is read back by the threads during item G+1 and
//                            rescaled/accumulated in registers (4 x 16 fp32 per row).
// S_{G+1} is issued into the other buffer as soon as every thread has read O_{G-1} out of it, i.e. one
// full exp2 pass before it is needed: UMMA latency is off the critical path.  Threads synchronise only
// through mbarriers (P half ready, P.V done, O read, token), never with a CTA-wide barrier.  The TMA warp
// runs ahead across unit boundaries (next unit's Q / K / V are loaded while the current one finishes).
//
// Inputs come from triattn_proj (prd_rowtile.cu): q (x log2(e)/sqrt(c)), k, g=sigmoid(gate) as
// [B*N seq][N tok][64] fp16 and vt [B*N seq][64][plane_ld(N)] fp16.
// Output og [B*N*N][64] fp16 = g * softmax(..) v, consumed by triattn_out.
#include "prd_kernels.h"
#include "prd_rowtile.cuh"
#include "prd_flash_math.cuh"
#include <stdlib.h>

#include <algorithm>

namespace prd {

constexpr int kFlashThreads = 384;
constexpr bool kFlashPolyHalf = true;  // every fourth column pair: exp2 on the FMA pipe (exp2_poly2); measured at B=8 N=512:
                                       // none 1.86 ms, 1/8 1.81, 1/4 1.77 -> shipped, 1/2 1.90 (instruction bound)
constexpr bool kFlashToken = false;  // strict per-scheduler MUFU ping-pong between the two groups; measured slower (2.32 vs 1.78 ms):
                                     // a lone warp in its exp2 pass is latency bound, two overlapping passes fill each other's bubbles

// kTrace: debug instantiation that records clock64() at phase boundaries (PRD_FLASH_TRACE=<file>,
// tools/flash_trace.py).
constexpr int kTraceCtas = 8, kTraceEvents = 512;
#define FLASH_TRACE(ev)                                                                  \
  do {                                                                                   \
    if (kTrace && tr != nullptr && (threadIdx.x & 31) == 0 && tr_n < kTraceEvents) {     \
      tr[tr_n++] = (clock64() << 8) | (ev);                                              \
    }                                                                                    \
  } while (0)

// pass A: running row max over one 128-key S tile.
template <bool kAllValid>
__device__ __forceinline__ float flash_row_max(uint32_t tS, const float2* keyp, float mx) {
  if (kAllValid) {
    float rm[4] = {mx, -INFINITY, -INFINITY, -INFINITY};  // independent chains (ILP)
    uint32_t sa[32], sb[32];
    // tcgen05.wait::ld waits for ALL outstanding loads, so the next chunk's load is issued right after
    // the wait and is in flight while the current chunk is reduced
    tmem_ld32(tS, sa);
    tmem_ld_wait32(sa);
    tmem_ld32(tS + 32, sb);
#pragma unroll
    for (int j = 0; j < 16; ++j) rm[j & 3] = fmax3(rm[j & 3], __uint_as_float(sa[2 * j]), __uint_as_float(sa[2 * j + 1]));
    tmem_ld_wait32(sb);
    tmem_ld32(tS + 64, sa);
#pragma unroll
    for (int j = 0; j < 16; ++j) rm[j & 3] = fmax3(rm[j & 3], __uint_as_float(sb[2 * j]), __uint_as_float(sb[2 * j + 1]));
    tmem_ld_wait32(sa);
    tmem_ld32(tS + 96, sb);
#pragma unroll
    for (int j = 0; j < 16; ++j) rm[j & 3] = fmax3(rm[j & 3], __uint_as_float(sa[2 * j]), __uint_as_float(sa[2 * j + 1]));
    tmem_ld_wait32(sb);
#pragma unroll
    for (int j = 0; j < 16; ++j) rm[j & 3] = fmax3(rm[j & 3], __uint_as_float(sb[2 * j]), __uint_as_float(sb[2 * j + 1]));
    return fmaxf(fmaxf(rm[0], rm[1]), fmaxf(rm[2], rm[3]));
  } else {
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t sv[32];
      tmem_ld32(tS + c * 32, sv);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float2 e = keyp[c * 32 + j];
        mx = fmaxf(mx, fmaf(__uint_as_float(sv[j]), e.x, e.y));
      }
    }
    return mx;
  }
}

// volatile flavours: keep the hand-written software pipeline below in program order up to ptxas
__device__ __forceinline__ uint64_t fadd2v(uint64_t a, uint64_t b) {
  uint64_t d;
  asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float ex2v(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// in-place flavours ("+" ties input and output to the same register: no temporaries for ptxas to funnel
// every MUFU operand through, which serialises the stream on write-after-read scoreboards)
__device__ __forceinline__ void fadd2_ip(uint64_t& x, uint64_t b) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(b)); }
__device__ __forceinline__ void ex2_ip(float& x) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x)); }
__device__ __forceinline__ uint32_t cvt_f16x2v(float lo, float hi) {
  uint32_t y;
  asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo));
  return y;
}
__device__ __forceinline__ void sts128v(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// pass B: P = exp2(t - m) as the fp16 A operand, row sum returned.  Each 64-key half is announced on
// bar_pr[half] (count 128) as soon as this thread has written it.
// The caller has already issued the TMEM load of the first 16 columns into b0 (before waiting for its
// turn on the MUFU pipe).
//
// All-valid path: one flat software pipeline over the 64 column pairs of the item, in 8 chunks of 16
// columns held in three rotating 16-register buffers:
//   stage L  tcgen05.ld of chunk c+2            (in flight during chunk c)
//   stage A  x = s - m (FADD2) of chunk c+1     (interleaved with the MUFUs of chunk c)
//   stage M  p = exp2(x) (2 MUFU per pair) of chunk c
//   stage C  row sum (FADD2), fp16 pack (F2FP), 16-byte store of chunk c, 3 pairs behind stage M
// so the MUFU stream never drains at a chunk boundary and every MUFU operand has its own register.
template <bool kAllValid, bool kTrace>
__device__ __forceinline__ float flash_row_exp(uint32_t tS, const float2* keyp, float mx, uint8_t* sP, int t,
                                               uint64_t* bar_pr, uint32_t (&b0)[16], long long* tr, int& tr_n) {
  if (kAllValid) {
    // straightforward per-chunk code (ptxas schedules it better than a hand-ordered volatile pipeline: 154 vs 195
    // cycles per 16-column chunk for a lone warp, tools/micro/warps_scaling.cu); the TMEM load of chunk c+1 is in
    // flight while chunk c is processed
    const uint64_t nm2 = pack_f2(-mx, -mx);
    uint64_t racc[4] = {0ull, 0ull, 0ull, 0ull};
    const uint32_t sP_row = smem_u32(sP) + t * 128;
    uint32_t b1[16];
    auto chunk = [&](const uint32_t(&s)[16], int c) {
      uint32_t ph[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint64_t y = fadd2(pack_f2(__uint_as_float(s[2 * j]), __uint_as_float(s[2 * j + 1])), nm2);
        if (kFlashPolyHalf && (j & 3) == 3) {
          y = exp2_poly2(y);
        } else {
          float ya, yb;
          unpack_f2(y, ya, yb);
          y = pack_f2(ex2_approx(ya), ex2_approx(yb));
        }
        racc[j & 3] = fadd2(racc[j & 3], y);
        float ya, yb;
        unpack_f2(y, ya, yb);
        ph[j] = cvt_f16x2(ya, yb);
      }
      const int k0 = c * 16;
      const uint32_t blk = sP_row + (k0 >> 6) * (kTileRows * 128);
#pragma unroll
      for (int q = 0; q < 2; ++q)
        sts128v(blk + (((((k0 & 63) >> 3) + q) ^ (t & 7)) << 4), ph[4 * q], ph[4 * q + 1], ph[4 * q + 2], ph[4 * q + 3]);
    };
#pragma unroll
    for (int c = 0; c < 8; c += 2) {
      FLASH_TRACE(40 + c);
      tmem_ld_wait16(b0);
      tmem_ld16(tS + (c + 1) * 16, b1);
      chunk(b0, c);
      tmem_ld_wait16(b1);
      if (c + 2 < 8) tmem_ld16(tS + (c + 2) * 16, b0);
      chunk(b1, c + 1);
      if (c == 2 || c == 6) {
        tc_fence_before();  // all TMEM reads of this half of the S buffer are done (P.V overwrites columns 0-15)
        fence_proxy_async_smem();
        mbar_arrive(&bar_pr[c >> 2]);
      }
    }
    float r0, r1, r2, r3;
    unpack_f2(fadd2(racc[0], racc[1]), r0, r1);
    unpack_f2(fadd2(racc[2], racc[3]), r2, r3);
    return (r0 + r1) + (r2 + r3);
  } else {
    float rs = 0.f;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t sv[32];
      tmem_ld32(tS + c * 32, sv);
      tmem_ld_wait();
      float p[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float2 e = keyp[c * 32 + j];
        p[j] = ex2_approx(fmaf(__uint_as_float(sv[j]), e.x, e.y) - mx);
        rs += p[j];
      }
      store_a_cols32(sP, t, c * 32, p);
      if (c & 1) {
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(&bar_pr[c >> 1]);
      }
    }
    return rs;
  }
}

constexpr int kFlashGroupFixed = 16384 + 32768 + 16384 + 32768;  // Q, 2 K, V^T, P

template <bool kTrace>
__global__ void __launch_bounds__(kFlashThreads, 1)
triattn_flash_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_vt, const float* __restrict__ mask,
                     const __half* __restrict__ g_gate, __half* __restrict__ og, int N, int total_units,
                     long long* trace) {
  long long* tr = nullptr;
  int tr_n = 0;
  if (kTrace && blockIdx.x < kTraceCtas) tr = trace + ((long long)blockIdx.x * 12 + (threadIdx.x >> 5)) * kTraceEvents;

  extern __shared__ uint8_t raw[];
  const int nkt = (N + 127) / 128;
  const int n_items = nkt * 4;
  const int group_bytes = (kFlashGroupFixed + nkt * 1024 + 64 + 1023) & ~1023;
  uint8_t* sm = smem_align1024(raw);
  uint64_t* bars_all = reinterpret_cast<uint64_t*>(sm + 2 * group_bytes);
  uint64_t* tok = bars_all + 32;  // [2 groups][4 warps]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tok + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = warp < 8 ? (warp >> 2) : (warp & 1);  // softmax warps 0-3 / 4-7, control warps 8,10 / 9,11
  uint8_t* gsm = sm + grp * group_bytes;
  uint8_t* sQ = gsm;                  // [128 x 64] halves, 16 KB
  uint8_t* sK = sQ + 16384;           // 2 buffers of [128 keys x 64], 16 KB each
  uint8_t* sVt = sK + 32768;          // 2 K-blocks (64 keys each) of [64 (h,c) rows x 64 keys], 8 KB each
  uint8_t* sP = sVt + 16384;          // 2 K-blocks of [128 x 64 keys], 32 KB
  float2* sKey = reinterpret_cast<float2*>(sP + 32768);    // (mul, add) per key, nkt * 128 entries
  int* sWarpValid = reinterpret_cast<int*>(sKey + nkt * 128);  // [nkt][4]: the warp's 32 keys of the tile are all valid
  uint64_t* bars = bars_all + grp * 16;
  uint64_t* bar_q = bars;        // Q tile loaded
  uint64_t* bar_k = bars + 1;    // [2] K tile loaded
  uint64_t* bar_v = bars + 3;    // [4] V^T slice of head h loaded
  uint64_t* bar_s = bars + 7;    // [2] S tile complete in TMEM buffer
  uint64_t* bar_pr = bars + 9;   // [2] P half written by all 128 threads
  uint64_t* bar_pv = bars + 11;  // P.V of the item complete (P and V_h free, O partial readable)
  uint64_t* bar_or = bars + 12;  // all 128 threads have read the previous item's O partial

  if (threadIdx.x == 0) {
    for (int gg = 0; gg < 2; ++gg) {
      uint64_t* b = bars_all + gg * 16;
      for (int i = 0; i < 9; ++i) mbar_init(&b[i], 1);  // q, k[2], v[4], s[2]
      mbar_init(&b[9], 128);
      mbar_init(&b[10], 128);
      mbar_init(&b[11], 1);
      mbar_init(&b[12], 128);
    }
    for (int i = 0; i < 8; ++i) mbar_init(&tok[i], 1);
    fence_barrier_init();
    for (int w = 0; w < 4; ++w) mbar_arrive(&tok[w]);  // group A goes first
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_vt);
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot + grp * 256;

  // work units: unit u = (sequence u / nkt, query tile u % nkt); CTA c takes unit pairs c, c + grid, ...;
  // group g takes unit 2 * pair + g.  The q-tiles of one sequence are processed close together in time
  // (K / V stay in L2).
  const int npairs = (total_units + 1) >> 1;
  const int rounds = ((int)blockIdx.x < npairs) ? (npairs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  auto unit_of = [&](int r) { return 2 * ((int)blockIdx.x + r * (int)gridDim.x) + grp; };
  const int nu = (rounds > 0 && unit_of(rounds - 1) >= total_units) ? rounds - 1 : rounds;  // this group's units
  const int total_items = nu * n_items;
  FLASH_TRACE(1);

  if (warp >= 10) {
    // ------------------------------------------------------------------ TMA warp of group grp
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    if (nu > 0) {
      auto load_q = [&](int r) {
        if (r >= nu) return;
        const int u = unit_of(r);
        mbar_expect_tx(bar_q, 16384);
        tma_load_3d(sQ, &map_q, bar_q, 0, (u % nkt) * 128, u / nkt);
      };
      auto load_k = [&](int gk) {  // gk = r * nkt + kt
        const int r = gk / nkt;
        if (r >= nu) return;
        const int u = unit_of(r);
        uint64_t* bar = &bar_k[gk & 1];
        mbar_expect_tx(bar, 16384);
        tma_load_3d(sK + (gk & 1) * 16384, &map_k, bar, 0, (gk - r * nkt) * 128, u / nkt);
      };
      auto load_v = [&](int gk, int h) {
        const int r = gk / nkt;
        if (r >= nu) return;
        const int u = unit_of(r);
        const int key0 = (gk - r * nkt) * 128;
        mbar_expect_tx(&bar_v[h], 4096);
        tma_load_3d(sVt + h * 2048, &map_vt, &bar_v[h], key0, h * 16, u / nkt);
        tma_load_3d(sVt + 8192 + h * 2048, &map_vt, &bar_v[h], key0 + 64, h * 16, u / nkt);
      };
      if (elect_one()) {
        load_q(0);
        load_k(0);
        load_k(1);
        for (int h = 0; h < 4; ++h) load_v(0, h);
      }
      int r = 0, it = 0;
      for (int G = 0; G < total_items; ++G) {
        const int h = it & 3, gk = r * nkt + (it >> 2);
        // P.V_G complete (and with it every UMMA issued before: S_G, S_{G+1}): V slice h, after the last head
        // the K buffer, and after the unit's last S the Q tile may be overwritten
        mbar_wait(bar_pv, G & 1);
        if (elect_one()) {
          load_v(gk + 1, h);
          if (h == 3) load_k(gk + 2);
          if (it == n_items - 2) load_q(r + 1);
        }
        __syncwarp();
        if (++it == n_items) {
          it = 0;
          ++r;
        }
      }
    }
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ UMMA warp of group grp
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    if (nu > 0) {
      const uint64_t dq = umma_desc_sw128(smem_u32(sQ));
      const uint64_t dk0 = umma_desc_sw128(smem_u32(sK));
      const uint64_t dp0 = umma_desc_sw128(smem_u32(sP)), dp1 = umma_desc_sw128(smem_u32(sP) + 16384);
      const uint64_t dv0 = umma_desc_sw128(smem_u32(sVt)), dv1 = umma_desc_sw128(smem_u32(sVt) + 8192);
      const uint32_t idesc_s = umma_idesc_f16(128, 128), idesc_o = umma_idesc_f16(128, 16);
      mbar_wait(bar_q, 0);
      mbar_wait(&bar_k[0], 0);
      tc_fence_after();
      if (elect_one()) {
        umma_f16(tmem, dq, dk0, idesc_s, 0u);
        umma_commit(&bar_s[0]);
      }
      __syncwarp();
      int r = 0, it = 0;
      for (int G = 0; G < total_items; ++G) {
        const int h = it & 3, gk = r * nkt + (it >> 2);
        int r1 = r, it1 = it + 1;
        if (it1 == n_items) {
          it1 = 0;
          ++r1;
        }
        // every thread has read O_{G-1}: the other S buffer is free
        mbar_wait(bar_or, G & 1);
        FLASH_TRACE(10);
        if (G + 1 < total_items) {
          const int h1 = it1 & 3, gk1 = r1 * nkt + (it1 >> 2);
          if (it1 == 0) mbar_wait(bar_q, r1 & 1);
          if (h1 == 0) mbar_wait(&bar_k[gk1 & 1], (gk1 >> 1) & 1);
          tc_fence_after();
          if (elect_one()) {
            umma_f16(tmem + ((G + 1) & 1) * 128, dq + 2 * h1, dk0 + (gk1 & 1) * (16384 >> 4) + 2 * h1, idesc_s, 0u);
            umma_commit(&bar_s[(G + 1) & 1]);
          }
          __syncwarp();
        }
        FLASH_TRACE(11);
        mbar_wait(&bar_v[h], gk & 1);
        const uint32_t tO = tmem + (G & 1) * 128;
        mbar_wait(&bar_pr[0], G & 1);
        tc_fence_after();
        FLASH_TRACE(12);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tO, dp0 + 2 * k, dv0 + h * (2048 >> 4) + 2 * k, idesc_o, k > 0 ? 1u : 0u);
        }
        __syncwarp();
        mbar_wait(&bar_pr[1], G & 1);
        tc_fence_after();
        FLASH_TRACE(13);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tO, dp1 + 2 * k, dv1 + h * (2048 >> 4) + 2 * k, idesc_o, 1u);
          umma_commit(bar_pv);
        }
        __syncwarp();
        FLASH_TRACE(14);
        r = r1;
        it = it1;
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax group grp
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int t = threadIdx.x & 127, w = warp & 3;
    const uint32_t tm_lane = tmem + (static_cast<uint32_t>(w * 32) << 16);
    uint64_t* tok_mine = &tok[grp * 4 + w];
    uint64_t* tok_other = &tok[(grp ^ 1) * 4 + w];
    int prev_seq = -1;
    int G = 0;
    if (nu > 0) mbar_arrive(bar_or);  // "O_{-1} has been read"
    for (int r = 0; r < nu; ++r) {
      const int u = unit_of(r);
      const int seq = u / nkt, qt = u - seq * nkt;
      const bool partner = kFlashToken && (u ^ 1) < total_units;  // the other group runs a unit in this round: take turns on the MUFU
      if (seq != prev_seq) {
        // key mask = m[b,seq_pos] * m[b,key]  (mask_2d row / column; symmetric, one formula for both modes).
        // Every thread of the group is past its last read of the previous table (it has seen P.V of the last item).
        prev_seq = seq;
        const int b = seq / N;
        const float ms = mask[seq];
        for (int kt = 0; kt < nkt; ++kt) {
          const int j = kt * 128 + t;
          float2 e;
          if (j >= N) e = make_float2(0.f, -INFINITY);
          else if (ms * mask[(long long)b * N + j] < 0.5f) e = make_float2(0.f, kMaskFillLog2);
          else e = make_float2(1.f, 0.f);
          sKey[j] = e;
          const bool all = __all_sync(0xffffffffu, e.x != 0.f);
          if (lane == 0) sWarpValid[kt * 4 + w] = all ? 1 : 0;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
      }
      float o[4][16];
      float mrow[4], lrow[4], alpha[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        mrow[h] = -INFINITY;
        lrow[h] = 0.f;
        alpha[h] = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) o[h][c] = 0.f;
      }
      for (int kt = 0; kt < nkt; ++kt) {
        const int4 wv = *reinterpret_cast<const int4*>(sWarpValid + kt * 4);
        const bool all_valid = (wv.x & wv.y & wv.z & wv.w) != 0;
        const float2* keyp = sKey + kt * 128;
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const uint32_t tS = tm_lane + (G & 1) * 128;
          FLASH_TRACE(20);
          mbar_wait(&bar_s[G & 1], (G >> 1) & 1);
          tc_fence_after();
          FLASH_TRACE(21);
          const float mx = all_valid ? flash_row_max<true>(tS, keyp, mrow[h]) : flash_row_max<false>(tS, keyp, mrow[h]);
          alpha[h] = ex2_approx(mrow[h] - mx);
          mrow[h] = mx;
          FLASH_TRACE(22);
          if (kt > 0 || h > 0) {
            // O partial of the previous item: columns [0,16) of the other S buffer
            const int hp = (h + 3) & 3;  // static after unrolling
            mbar_wait(bar_pv, (G - 1) & 1);
            tc_fence_after();
            uint32_t ov[16];
            tmem_ld16(tm_lane + ((G - 1) & 1) * 128, ov);
            tmem_ld_wait16(ov);
            tc_fence_before();
            mbar_arrive(bar_or);
#pragma unroll
            for (int c = 0; c < 16; ++c) o[hp][c] = fmaf(o[hp][c], alpha[hp], __uint_as_float(ov[c]));
          }
          FLASH_TRACE(23);
          uint32_t sa[16];
          if (all_valid) tmem_ld16(tS, sa);  // in flight while this warp waits for its turn
          if (partner) mbar_wait(tok_mine, G & 1);
          FLASH_TRACE(24);
          const float rs = all_valid ? flash_row_exp<true, kTrace>(tS, keyp, mx, sP, t, bar_pr, sa, tr, tr_n)
                                     : flash_row_exp<false, kTrace>(tS, keyp, mx, sP, t, bar_pr, sa, tr, tr_n);
          if (partner && lane == 0) mbar_arrive(tok_other);
          lrow[h] = fmaf(lrow[h], alpha[h], rs);
          FLASH_TRACE(25);
          ++G;
        }
      }
      // ---- unit epilogue: last O partial, normalise, gate, store
      {
        mbar_wait(bar_pv, (G - 1) & 1);
        tc_fence_after();
        uint32_t ov[16];
        tmem_ld16(tm_lane + ((G - 1) & 1) * 128, ov);
        tmem_ld_wait16(ov);
        tc_fence_before();
        mbar_arrive(bar_or);  // for the first item of the next unit
#pragma unroll
        for (int c = 0; c < 16; ++c) o[3][c] = fmaf(o[3][c], alpha[3], __uint_as_float(ov[c]));
      }
      FLASH_TRACE(30);
      {
        // rows [32 w, 32 w + 32) of P K-block 0 are private to this warp and idle (the last P.V is complete):
        // 4 KB slice for the coalesced load of the gate rows / store of the output rows
        uint8_t* slice = sP + w * 4096;
        const int row0 = qt * 128 + w * 32;
        const long long grow = ((long long)seq * N + row0) * 64;
        uint4 gv[8];
        warp_load_rows128(slice, lane, gv, g_gate + grow, 128, N - row0);
        uint4 ovv[8];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const float inv = 1.0f / lrow[h];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const __half2* g2 = reinterpret_cast<const __half2*>(&gv[h * 2 + half]);
            uint32_t* o32 = reinterpret_cast<uint32_t*>(&ovv[h * 2 + half]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 gf = __half22float2(g2[e]);
              o32[e] = pack_half2(o[h][half * 8 + 2 * e] * inv * gf.x, o[h][half * 8 + 2 * e + 1] * inv * gf.y);
            }
          }
        }
        warp_store_rows128(slice, lane, ovv, og + grow, 128, N - row0);
        __syncwarp();  // the slice is P again from here on
      }
      FLASH_TRACE(31);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(*tmem_slot, 512);
}

int triattn_flash(const PairDims& d, const float* mask, const __half* q, const __half* k, const __half* g,
                  const __half* vt, __half* og, cudaStream_t s) {
  if (triattn_flash_g4_applies(d)) return triattn_flash_g4(d, mask, q, k, g, vt, og, s);
  const int N = d.N, Np = plane_ld(N);
  const long long nseq = (long long)d.B * N;
  CUtensorMap mq, mk, mv;
  TmaDims t;
  // q / k: [seq][tok][64]
  t.size[0] = 64; t.size[1] = (uint64_t)N; t.size[2] = (uint64_t)nseq; t.size[3] = 1;
  t.stride[0] = 128; t.stride[1] = (uint64_t)N * 128; t.stride[2] = 0;
  t.box[0] = 64; t.box[1] = 128; t.box[2] = 1; t.box[3] = 1;
  if (make_tensor_map(&mq, q, 2, 3, t, true)) return 1;
  if (make_tensor_map(&mk, k, 2, 3, t, true)) return 1;
  // vt: [seq][64 (h,c)][tok], tok contiguous, row stride Np; one box = one head's 16 rows x 64 keys
  t.size[0] = (uint64_t)N; t.size[1] = 64; t.size[2] = (uint64_t)nseq;
  t.stride[0] = (uint64_t)Np * 2; t.stride[1] = (uint64_t)Np * 2 * 64;
  t.box[0] = 64; t.box[1] = 16; t.box[2] = 1;
  if (make_tensor_map(&mv, vt, 2, 3, t, true)) return 1;
  const int nkt = (N + 127) / 128;
  const int group_bytes = (kFlashGroupFixed + nkt * 1024 + 64 + 1023) & ~1023;
  const int smem = 1024 + 2 * group_bytes + 48 * 8 + 16;
  PRD_REQUIRE(smem <= 227 * 1024, "triattn_flash: N=%d needs %d B of shared memory", N, smem);
  PRD_REQUIRE(nseq * nkt <= 2147483647LL, "triattn_flash: too many work units");
  const int total_units = (int)(nseq * nkt);
  const int grid = std::min((total_units + 1) / 2, kNumSMs);
  if (const char* path = getenv("PRD_FLASH_TRACE")) {
    // debug: phase timeline of the first kTraceCtas CTAs written to `path` (tools/flash_trace.py)
    static long long* dtrace = nullptr;
    const size_t n = (size_t)kTraceCtas * 12 * kTraceEvents;
    if (!dtrace) PRD_CUDA_OK(cudaMalloc(&dtrace, n * 8));
    PRD_CUDA_OK(cudaMemsetAsync(dtrace, 0, n * 8, s));
    PRD_CUDA_OK(cudaFuncSetAttribute(triattn_flash_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    triattn_flash_kernel<true><<<grid, kFlashThreads, smem, s>>>(mq, mk, mv, mask, g, og, N, total_units, dtrace);
    PRD_LAUNCHED();
    PRD_CUDA_OK(cudaStreamSynchronize(s));
    long long* h = (long long*)malloc(n * 8);
    PRD_CUDA_OK(cudaMemcpy(h, dtrace, n * 8, cudaMemcpyDeviceToHost));
    if (FILE* f = fopen(path, "wb")) {
      fwrite(h, 8, n, f);
      fclose(f);
    }
    free(h);
    return 0;
  }
  PRD_CUDA_OK(cudaFuncSetAttribute(triattn_flash_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  triattn_flash_kernel<false><<<grid, kFlashThreads, smem, s>>>(mq, mk, mv, mask, g, og, N, total_units, nullptr);
  PRD_LAUNCHED();
  return 0;
}

}  // namespace prd
