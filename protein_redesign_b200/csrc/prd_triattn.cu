// Triangle attention core (modules.py:185-225 applied to every row / column of the pair tensor,
// modules.py:236-243): flash-style gated attention, 4 heads x 16 channels, key mask with the
// reference's finite fill value (-2^15), online softmax, S = QK^T and O = PV on tcgen05.
//
// One CTA = one (sequence, 128-query tile).  Thread t owns query row t (TMEM lane t):
//   S_h  = Q_h K_h^T   one UMMA, M=128 N=128 K=16 (head h = 32-byte K-slice of the 128-byte rows)
//   P_h  = exp2(S_h - m)  -> fp16 A operand in shared memory (SWIZZLE_128B, 2 K-blocks)
//   O_h += P_h V_h     8 UMMAs, M=128 N=16 K=16, B = V^T tile [16 x 128 keys] (K-major)
// Running max / sum / output (4 x 16 fp32) stay in registers; the per-key-tile partial product is
// read back from TMEM and rescaled there, so no TMEM "correction" pass is needed.
//
// Inputs come from triattn_proj (prd_rowtile.cu): q (x 1/sqrt(c)), k, g=sigmoid(gate) as
// [B*N seq][N tok][64] fp16 and vt [B*N seq][64][plane_ld(N)] fp16.
// Output og [B*N*N][64] fp16 = g * softmax(..) v, consumed by triattn_out.
#include "prd_kernels.h"
#include "prd_rowtile.cuh"

namespace prd {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kMaskFillLog2 = -32768.0f * kLog2e;  // modules.py:177,220 in the exp2 domain

// Per-key softmax terms in the exp2 domain: t_j = s_j * mul_j + add_j
//   valid key   : mul = log2(e), add = 0
//   masked key  : mul = 0,       add = -2^15 * log2(e)   (the reference's finite fill value)
//   j >= N (pad): mul = 0,       add = -inf               (does not exist: p = 0)
// A key tile whose 128 keys are all valid takes a fast path without any per-key loads.
//
// One softmax group (128 threads) per CTA, two CTAs per SM: the two CTAs drift apart, so one is in
// its exp2 loop while the other waits for an S tile.  (A warp-specialised single-CTA variant with two
// softmax groups sharing K/V, a TMA warp and a UMMA warp was measured 20 % slower -- the groups ran in
// lockstep -- see profiles/r01_flash_variants.md.)
__global__ void __launch_bounds__(128, 2)
triattn_flash_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_vt, const float* __restrict__ mask,
                     const __half* __restrict__ g, __half* __restrict__ og, int N) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = smem_align1024(raw);
  uint8_t* sQ = sm;                 // [128 x 64] halves, 16 KB
  uint8_t* sK = sQ + 16384;         // 2 buffers of [128 keys x 64], 16 KB each (next tile prefetched)
  uint8_t* sVt = sK + 32768;        // 2 boxes of [64 rows x 64 keys], 8 KB each
  uint8_t* sP = sVt + 16384;        // 2 K-blocks of [128 x 64 keys], 32 KB
  const int nkt = (N + 127) / 128;
  float2* sKey = reinterpret_cast<float2*>(sP + 32768);  // (mul, add) per key, nkt * 128 entries
  int* sAllValid = reinterpret_cast<int*>(sKey + nkt * 128);  // per key tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(sAllValid + ((nkt + 1) & ~1));
  uint64_t* bar_q = bars;
  uint64_t* bar_k = bars + 1;  // [2]
  uint64_t* bar_v = bars + 3;
  uint64_t* bar_s = bars + 4;
  uint64_t* bar_o = bars + 5;
  uint64_t* bar_p = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  const int t = threadIdx.x, warp = t >> 5;
  const int qt = blockIdx.x % nkt;   // the q-tiles of one sequence are adjacent CTAs: K/V stay in L2
  const int seq = blockIdx.x / nkt;  // b * N + s
  const int b = seq / N;
  if (t == 0) {
    mbar_init(bar_q, 1);
    mbar_init(&bar_k[0], 1);
    mbar_init(&bar_k[1], 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    mbar_init(bar_p, 1);
    fence_barrier_init();
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_vt);
    mbar_expect_tx(bar_q, 16384);
    tma_load_3d(sQ, &map_q, bar_q, 0, qt * 128, seq);
    mbar_expect_tx(&bar_k[0], 16384);
    tma_load_3d(sK, &map_k, &bar_k[0], 0, 0, seq);
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  {
    // key mask = m[b,seq_pos] * m[b,key]  (mask_2d row / column; symmetric, one formula for both modes)
    const float ms = mask[seq];
    for (int j = t; j < nkt * 128; j += 128) {
      float2 e;
      if (j >= N) e = make_float2(0.f, -INFINITY);
      else if (ms * mask[(long long)b * N + j] < 0.5f) e = make_float2(0.f, kMaskFillLog2);
      else e = make_float2(kLog2e, 0.f);
      sKey[j] = e;
    }
  }
  __syncthreads();
  if (t < nkt) {
    int all = 1;
    for (int j = 0; j < 128; ++j) all &= (sKey[t * 128 + j].x != 0.f) ? 1 : 0;
    sAllValid[t] = all;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const uint32_t tm_o = 128;  // column offset of the four 16-column O chunks

  float o[4][16];
  float mrow[4], lrow[4];
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    mrow[h] = -INFINITY;
    lrow[h] = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) o[h][c] = 0.f;
  }
  uint32_t ph_s = 0, n_p = 0;
  const uint64_t dq = umma_desc_sw128(smem_u32(sQ));

  for (int kt = 0; kt < nkt; ++kt) {
    uint8_t* sKc = sK + (kt & 1) * 16384;
    if (t == 0) {
      // V^T of this tile (its buffer was released by the end-of-tile barrier) and K of the next tile
      mbar_expect_tx(bar_v, 16384);
      tma_load_3d(sVt, &map_vt, bar_v, kt * 128, 0, seq);
      tma_load_3d(sVt + 8192, &map_vt, bar_v, kt * 128 + 64, 0, seq);
      if (kt + 1 < nkt) {
        mbar_expect_tx(&bar_k[(kt + 1) & 1], 16384);
        tma_load_3d(sK + ((kt + 1) & 1) * 16384, &map_k, &bar_k[(kt + 1) & 1], 0, (kt + 1) * 128, seq);
      }
      if (kt == 0) {
        mbar_wait(bar_q, 0);
        mbar_wait(&bar_k[0], 0);
        tc_fence_after();
        umma_f16(tmem, dq, umma_desc_sw128(smem_u32(sKc)), umma_idesc_f16(128, 128), 0u);  // S for head 0
        umma_commit(bar_s);
      }
    }
    float alpha[4];
    const bool all_valid = sAllValid[kt] != 0;
    const float2* keyp = sKey + kt * 128;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      mbar_wait(bar_s, ph_s);
      ph_s ^= 1;
      tc_fence_after();
      // pass A: row max over this key tile (exp2 domain)
      float mx = mrow[h];
      if (all_valid) {
        float rm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // independent chains (ILP)
        // 16-column chunks, double buffered: the TMEM load of chunk c+1 is in flight while chunk c is reduced
        uint32_t sa[16], sb[16];
        tmem_ld16(tm_lane, sa);
        tmem_ld_wait16(sa);
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
          tmem_ld16(tm_lane + (c + 1) * 16, sb);
#pragma unroll
          for (int j = 0; j < 16; ++j) rm[j & 3] = fmaxf(rm[j & 3], __uint_as_float(sa[j]));
          tmem_ld_wait16(sb);
          if (c + 2 < 8) tmem_ld16(tm_lane + (c + 2) * 16, sa);
#pragma unroll
          for (int j = 0; j < 16; ++j) rm[j & 3] = fmaxf(rm[j & 3], __uint_as_float(sb[j]));
          if (c + 2 < 8) tmem_ld_wait16(sa);
        }
        mx = fmaxf(mx, fmaxf(fmaxf(rm[0], rm[1]), fmaxf(rm[2], rm[3])) * kLog2e);
      } else {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t sv[32];
          tmem_ld32(tm_lane + c * 32, sv);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float2 e = keyp[c * 32 + j];
            mx = fmaxf(mx, fmaf(__uint_as_float(sv[j]), e.x, e.y));
          }
        }
      }
      alpha[h] = ex2_approx(mrow[h] - mx);
      mrow[h] = mx;
      // the P buffer is free once the P.V UMMAs of the previous head have completed
      if (n_p > 0) mbar_wait(bar_p, (n_p - 1) & 1);
      // pass B: p = exp2(t - m), row sum, fp16 P tile
      float rs = 0.f;
      if (all_valid) {
        const float nmx = -mx;
        float r4[4] = {0.f, 0.f, 0.f, 0.f};  // independent chains (ILP)
        uint32_t sa[16], sb[16];
        tmem_ld16(tm_lane, sa);
        tmem_ld_wait16(sa);
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
          tmem_ld16(tm_lane + (c + 1) * 16, sb);
          {
            float p[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              p[j] = ex2_approx(fmaf(__uint_as_float(sa[j]), kLog2e, nmx));
              r4[j & 3] += p[j];
            }
            store_a_cols16(sP, t, c * 16, p);
          }
          tmem_ld_wait16(sb);
          if (c + 2 < 8) tmem_ld16(tm_lane + (c + 2) * 16, sa);
          {
            float p[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              p[j] = ex2_approx(fmaf(__uint_as_float(sb[j]), kLog2e, nmx));
              r4[j & 3] += p[j];
            }
            store_a_cols16(sP, t, (c + 1) * 16, p);
          }
          if (c + 2 < 8) tmem_ld_wait16(sa);
        }
        rs = (r4[0] + r4[1]) + (r4[2] + r4[3]);
      } else {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t sv[32];
          tmem_ld32(tm_lane + c * 32, sv);
          tmem_ld_wait();
          float p[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float2 e = keyp[c * 32 + j];
            p[j] = ex2_approx(fmaf(__uint_as_float(sv[j]), e.x, e.y) - mx);
            rs += p[j];
          }
          store_a_cols32(sP, t, c * 32, p);
        }
      }
      lrow[h] = lrow[h] * alpha[h] + rs;
      ++n_p;
      sync_before_mma();  // P visible to the tensor core; every thread is done reading S
      if (t == 0) {
        tc_fence_after();
        // the S tile every thread waits for next goes first: next head, or head 0 of the next key tile
        if (h < 3) {
          umma_f16(tmem, dq + 2 * (h + 1), umma_desc_sw128(smem_u32(sKc)) + 2 * (h + 1), umma_idesc_f16(128, 128), 0u);
          umma_commit(bar_s);
        } else if (kt + 1 < nkt) {
          mbar_wait(&bar_k[(kt + 1) & 1], ((kt + 1) >> 1) & 1);
          tc_fence_after();
          umma_f16(tmem, dq, umma_desc_sw128(smem_u32(sK + ((kt + 1) & 1) * 16384)), umma_idesc_f16(128, 128), 0u);
          umma_commit(bar_s);
        }
        if (h == 0) mbar_wait(bar_v, kt & 1);
        const uint32_t idesc = umma_idesc_f16(128, 16);
        umma_kblock(tmem + tm_o + 16 * h, smem_u32(sP), smem_u32(sVt) + h * 2048, idesc, false);
        umma_kblock(tmem + tm_o + 16 * h, smem_u32(sP) + 16384, smem_u32(sVt) + 8192 + h * 2048, idesc, true);
        umma_commit(bar_p);
        if (h == 3) umma_commit(bar_o);
      }
    }
    // the four P.V products of this key tile: read back once, rescale in registers
    mbar_wait(bar_o, kt & 1);
    tc_fence_after();
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      uint32_t ov[16];
      tmem_ld16(tm_lane + tm_o + 16 * h, ov);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 16; ++c) o[h][c] = o[h][c] * alpha[h] + __uint_as_float(ov[c]);
    }
    // V^T / P buffers and the O columns are reused by the next key tile
    tc_fence_before();
    __syncthreads();
  }

  const int tok = qt * 128 + t;
  if (tok < N) {
    const long long r = (long long)seq * N + tok;
    const uint4* gp = reinterpret_cast<const uint4*>(g + r * 64);
    uint4* op = reinterpret_cast<uint4*>(og + r * 64);
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const float inv = 1.0f / lrow[h];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint4 gv = __ldg(gp + h * 2 + half);
        const __half2* g2 = reinterpret_cast<const __half2*>(&gv);
        uint4 ovv;
        uint32_t* o32 = reinterpret_cast<uint32_t*>(&ovv);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 gf = __half22float2(g2[e]);
          o32[e] = pack_half2(o[h][half * 8 + 2 * e] * inv * gf.x, o[h][half * 8 + 2 * e + 1] * inv * gf.y);
        }
        op[h * 2 + half] = ovv;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

int triattn_flash(const PairDims& d, const float* mask, const __half* q, const __half* k, const __half* g,
                  const __half* vt, __half* og, cudaStream_t s) {
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
  // vt: [seq][64 (h,c)][tok], tok contiguous, row stride Np
  t.size[0] = (uint64_t)N; t.size[1] = 64; t.size[2] = (uint64_t)nseq;
  t.stride[0] = (uint64_t)Np * 2; t.stride[1] = (uint64_t)Np * 2 * 64;
  t.box[0] = 64; t.box[1] = 64; t.box[2] = 1;
  if (make_tensor_map(&mv, vt, 2, 3, t, true)) return 1;
  const int nkt = (N + 127) / 128;
  const int smem = 1024 + 16384 * 4 + 32768 + nkt * 128 * 8 + ((nkt + 1) & ~1) * 4 + 128;
  PRD_REQUIRE(smem <= 113 * 1024, "triattn_flash: N=%d needs %d B of shared memory", N, smem);
  PRD_CUDA_OK(cudaFuncSetAttribute(triattn_flash_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  PRD_REQUIRE(nseq * nkt <= 2147483647LL, "triattn_flash: grid overflow");
  triattn_flash_kernel<<<(unsigned)(nseq * nkt), 128, smem, s>>>(mq, mk, mv, mask, g, og, N);
  PRD_LAUNCHED();
  return 0;
}

}  // namespace prd
