// Triangle attention core, four-group variant (same math and interface as prd_triattn.cu).
//
// The exp2 stream of the softmax is the binding resource; with two softmax groups per SM (prd_triattn.cu) every
// item leaves the MUFU / FMA pipes idle while its rows wait for UMMAs and barriers (~35 % of the time, phase
// trace in profiles/r01_flash_variants.md).  TMEM (512 columns) caps the groups per SM, so this variant shrinks
// the per-group footprint to 128 columns and runs FOUR groups = 16 softmax warps per SM:
//   * item = (64-key tile, head): S = Q_h K_h^T, one UMMA M=128 N=64 K=16 into the group's single S buffer
//     (columns [0,64)); O_h += P V_h accumulates in TMEM (columns [64,128)) over all key tiles, nothing is
//     read back until the unit ends; P goes through one 16 KB shared-memory tile per group
//   * lazy running max: m_h is fixed by the first key tile; a later tile keeps it while no score exceeds it by
//     2^14 (P still fits fp16), otherwise the exact path rescales O_h in TMEM
//   * the four groups of a CTA take the q-tiles 4 r .. 4 r + 3 of ONE sequence: K / V tiles are loaded once per CTA
//     into 3-slot rings shared by the groups (full / empty mbarriers, empty counts one arrival per group)
//   * warps 0-15 softmax (thread = query row), 16-19 UMMA issue (one per group), 20 K/V TMA, 21 Q TMA, 22-23 idle
// A group's S / P.V latency (single buffers) is hidden by the other three groups.  Measured at B=8 N=512: 1.56 ms against
// 1.63 ms of the two-group kernel (softmax warps still wait ~27 % of the time for S / P.V of their own group; staggering
// the groups' start does not change it).  Used when the number of query tiles per sequence is a multiple of four.
#include "prd_kernels.h"
#include "prd_rowtile.cuh"
#include "prd_flash_math.cuh"

#include <stdlib.h>

#include <algorithm>
#include <type_traits>

namespace prd {

namespace {

constexpr int kG4Threads = 768;
constexpr int kG4Ring = 3;
constexpr float kG4LazyBound = 14.0f;  // log2: P <= 2^14 < fp16 max

// exp2 of one 16-column chunk of scores (already in registers): running max tracking, P = exp2(s - m) (MUFU on 3/4 of
// the pairs, FMA-pipe polynomial on 1/4), packed row sum, fp16 pack, two 16-byte stores into the P row.
// kTrackMax: keep the running maximum of the scores in rm (the exact path needs it; the fast path detects an overflow of
// the lazy bound through the row SUM instead -- any P > 2^14 makes the fp32 sum exceed 2^14 -- which removes one FMNMX3
// per pair, ~12 % of the fast path's instructions, from a loop that is issue bound)
template <int kPolyMask, bool kTrackMax = true>  // bit j set: pair j of the chunk (0..7) takes the FMA-pipe polynomial instead of MUFU
__device__ __forceinline__ void g4_chunk(const uint32_t (&s)[16], uint64_t nm2, uint64_t (&racc)[2], float (&rm)[2],
                                         uint32_t sP_row, int t, int c, uint64_t* p_free = nullptr, uint32_t p_parity = 0) {
  uint32_t ph[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float a = __uint_as_float(s[2 * j]), b = __uint_as_float(s[2 * j + 1]);
    if (kTrackMax) rm[j & 1] = fmax3(rm[j & 1], a, b);
    uint64_t y = fadd2(pack_f2(a, b), nm2);
    if ((kPolyMask >> j) & 1) {
      y = exp2_poly2(y);
    } else {
      float ya, yb;
      unpack_f2(y, ya, yb);
      y = pack_f2(ex2_approx(ya), ex2_approx(yb));
    }
    racc[j & 1] = fadd2(racc[j & 1], y);
    float ya, yb;
    unpack_f2(y, ya, yb);
    ph[j] = cvt_f16x2(ya, yb);
  }
  // the P tile is free once P.V of the previous item is complete: waited for here, behind the chunk's arithmetic
  if (p_free != nullptr) mbar_wait(p_free, p_parity);
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sP_row + (((2 * c + q) ^ (t & 7)) << 4)), "r"(ph[4 * q]),
                 "r"(ph[4 * q + 1]), "r"(ph[4 * q + 2]), "r"(ph[4 * q + 3])
                 : "memory");
  }
}

// scores of chunk c with the key mask applied (masked key -> fill value, non-existent key -> -inf)
__device__ __forceinline__ void g4_mask_chunk(uint32_t (&s)[16], const float* keyp, int c) {
#pragma unroll
  for (int j = 0; j < 16; j += 4) {
    const float4 e = *reinterpret_cast<const float4*>(keyp + c * 16 + j);
    if (e.x != 0.f) s[j] = __float_as_uint(e.x);
    if (e.y != 0.f) s[j + 1] = __float_as_uint(e.y);
    if (e.z != 0.f) s[j + 2] = __float_as_uint(e.z);
    if (e.w != 0.f) s[j + 3] = __float_as_uint(e.w);
  }
}

struct G4Smem {
  static constexpr int kK = 0;                         // kG4Ring x 8 KB
  static constexpr int kV = kK + kG4Ring * 8192;       // kG4Ring x 8 KB
  static constexpr int kQ = kV + kG4Ring * 8192;       // 4 x 16 KB
  static constexpr int kP = kQ + 4 * 16384;            // 4 x 16 KB
  static constexpr int kWo = kP + 4 * 16384;           // fused out-projection: W_o hi, lo: 2 x [64 x 64] halves (1024-aligned)
  static constexpr int kBo = kWo + 16384;              // b_o [64]
  static constexpr int kKey = kBo + 256;               // 4 x (nkt * 64 floats + nkt * 16 bytes), sized at run time
};

// kFused: the output projection of the attention module (modules.py:222-225: out_proj, + residual) runs in the unit
// epilogue: the gated rows become a [128 x 64] fp16 A tile (the idle P tile), one UMMA pair against W_o (hi + lo)
// accumulates into the (already read) O columns, and the pair rows are updated in place with full-line loads /
// stores.  The kernel is exp2 bound with HBM idle, so the 2 P of traffic of the former triattn_out kernel are free
// here, and the og round trip (1 P) disappears.
struct G4OutProj {
  const float* pair;   // residual source (may equal dst)
  float* dst;
  const __half* w_o;   // fp16 pair [hi; lo], each [64 x 64]
  const float* b_o;    // [64]
  int residual;
  int transposed;      // 0: row (b, seq, tok) = pair[b, seq, tok]; 1: pair[b, tok, seq]
};

// Key tiles of a sequence that can contribute to its softmax.  The last valid key depends on the batch row only, so every
// CTA tabulates it once per launch (sNkEff, one warp per batch row) instead of scanning the mask per sequence.  Ragged batches:
//   * a VALID sequence (m_s = 1): keys of a padded token carry the logit -2^15, i.e. exp2(fill - max) == 0 exactly in fp32
//     next to at least one valid key, so every 64-key tile AFTER the last valid key adds exactly nothing to P, l and O and
//     is skipped by all roles (K / V loads, both UMMAs, the softmax pass);
//   * a PADDED sequence (m_s = 0): every logit is the same constant, the softmax is uniform over all N tokens.  The same
//     result comes out of the fast path when Q is zero (S = 0, P = exp2(0) = 1, l = N, O = sum_k V / N): the Q loader
//     zero-fills the tile instead of fetching it and the key table stays all-valid, so a padded sequence costs what a
//     valid one costs instead of taking the two-pass masked path on every tile.
// Both are exact restatements of the reference's masked_fill(-2^15) + softmax (modules.py:216-222).
constexpr int kG4MaxBatch = 256;  // batch rows whose effective tile count is tabulated (beyond that: no tile skipping)
template <bool kRagged>
__device__ __forceinline__ int g4_eff_tiles(const float* __restrict__ mask, const int* sNkEff, int seq, int N, int nkt,
                                            bool& padded_seq) {
  if (!kRagged) {  // all-valid variant: the tile count is the launch constant, nothing is read per sequence
    padded_seq = false;
    return nkt;
  }
  padded_seq = mask[seq] < 0.5f;
  const int bb = seq / N;
  return (padded_seq || bb >= kG4MaxBatch) ? nkt : sNkEff[bb];
}

template <bool kFused, int kPM = 0x88, bool kRagged = true, bool kSumCheck = true>
__global__ void __launch_bounds__(kG4Threads, 1)
triattn_flash_g4_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                        const __grid_constant__ CUtensorMap map_vt, const float* __restrict__ mask,
                        const __half* __restrict__ g_gate, __half* __restrict__ og, int N, int nseq, G4OutProj op) {
  extern __shared__ uint8_t raw[];
  pdl_trigger();
  const int nqt = (N + 127) / 128;  // query tiles per sequence
  const int nkt = (N + 63) / 64;    // 64-key tiles per sequence
  const int nrr = (nqt + 3) / 4;    // rounds of four query tiles per sequence
  const int n_items = nkt * 4;
  const int key_bytes = (nkt * 64 * 4 + nkt * 16 + 127) & ~127;
  uint8_t* sm = smem_align1024(raw);
  uint8_t* sK = sm + G4Smem::kK;
  uint8_t* sV = sm + G4Smem::kV;
  uint8_t* sWo = sm + G4Smem::kWo;
  float* sBo = reinterpret_cast<float*>(sm + G4Smem::kBo);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + G4Smem::kKey + 4 * key_bytes);
  uint64_t* k_full = bars;               // [ring]
  uint64_t* k_empty = bars + kG4Ring;    // [ring] one arrival per group
  uint64_t* v_full = bars + 2 * kG4Ring;
  uint64_t* v_empty = bars + 3 * kG4Ring;
  uint64_t* gb = bars + 4 * kG4Ring;     // per group: 8 barriers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gb + 4 * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kG4Ring; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 4);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 4);
    }
    for (int g = 0; g < 4; ++g) {
      uint64_t* b = gb + g * 8;
      mbar_init(&b[0], 1);    // q_full
      mbar_init(&b[1], 1);    // q_empty (UMMA commit after the unit's last S)
      mbar_init(&b[2], 1);    // s_full
      mbar_init(&b[3], 128);  // pr: P written, S consumed
      mbar_init(&b[4], 1);    // pv: P.V complete, P free
      mbar_init(&b[5], 128);  // o_read: O region free
      mbar_init(&b[6], 1);    // proj: output projection complete (kFused)
      mbar_init(&b[7], 128);  // sc: S consumed (its TMEM buffer may be overwritten), ahead of pr
    }
    fence_barrier_init();
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_vt);
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  if (kFused) {
    load_weight_kblocks(sWo, op.w_o, 64, 64, 64, threadIdx.x, kG4Threads);
    load_weight_kblocks(sWo + 8192, op.w_o + 64 * 64, 64, 64, 64, threadIdx.x, kG4Threads);
    if (threadIdx.x < 64) sBo[threadIdx.x] = op.b_o[threadIdx.x];
    fence_proxy_async_smem();
  }
  // effective key tiles per batch row (see g4_eff_tiles): tiles up to the last valid key; the token mask is an input of
  // the whole step, not a product of the previous kernel
  __shared__ int sNkEffBuf[kG4MaxBatch];
  const int* sNkEff = sNkEffBuf;
  if (kRagged) {
    const int nb = nseq / N < kG4MaxBatch ? nseq / N : kG4MaxBatch;
    for (int bb = warp; bb < nb; bb += kG4Threads / 32) {
      int last = -1;
      for (int j0 = 0; j0 < N; j0 += 32 * 8) {
        float mv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = j0 + u * 32 + lane;
          mv[u] = j < N ? mask[(long long)bb * N + j] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (mv[u] >= 0.5f) last = j0 + u * 32 + lane;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
      if (lane == 0) sNkEffBuf[bb] = last < 0 ? nkt : (last >> 6) + 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // everything above touched only weights / shared memory; the predecessor's output is read below

  // sequences of this CTA: seq = blockIdx.x + k * gridDim.x; round R = k * nrr + rr; group g takes q-tile 4 rr + g
  const int nseq_cta = ((int)blockIdx.x < nseq) ? (nseq - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int total_rounds = nseq_cta * nrr;

  if (warp < 16) {
    // ------------------------------------------------------------------ softmax group g
    asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
    const int g = warp >> 2, w = warp & 3, t = threadIdx.x & 127;
    uint64_t* b = gb + g * 8;
    uint64_t* s_full = &b[2];
    uint64_t* pr = &b[3];
    uint64_t* pv = &b[4];
    uint64_t* o_read = &b[5];
    uint64_t* sc = &b[7];
    const uint32_t tmem = *tmem_slot + g * 128;
    const uint32_t tS = tmem + (static_cast<uint32_t>(w * 32) << 16);
    const uint32_t tO = tS + 64;
    uint8_t* sP = sm + G4Smem::kP + g * 16384;
    const uint32_t sP_row = smem_u32(sP) + t * 128;
    float* sKey = reinterpret_cast<float*>(sm + G4Smem::kKey + g * key_bytes);
    int* sWarpValid = reinterpret_cast<int*>(sKey + nkt * 64);  // [ceil(nkt/2)][4]
    int G = 0;  // items of this group so far (barrier parities)
    int R = 0;
    int Ug = 0;  // units of this group so far
    for (int ks = 0; ks < nseq_cta; ++ks) {
      const int seq = (int)blockIdx.x + ks * (int)gridDim.x;
      bool table = false;
      bool padded_seq;
      const int nkt_s = g4_eff_tiles<kRagged>(mask, sNkEff, seq, N, nkt, padded_seq);
      for (int rr = 0; rr < nrr; ++rr, ++R) {
        const int qt = rr * 4 + g;
        if (qt >= nqt) continue;  // (uniform per group) no unit in this round
        if (!table) {
          // key mask = m[b,seq_pos] * m[b,key]; one private copy per group, rebuilt once per sequence.  Every thread
          // of the group is past its last read of the previous table (the unit epilogue waits for the last P.V).
          table = true;
          const int bb = seq / N;
          const float ms = mask[seq];
          for (int i = 0; i * 128 < nkt * 64; ++i) {
            const int j = i * 128 + t;
            float e = 0.f;
            if (j >= N) e = -INFINITY;
            else if (!padded_seq && ms * mask[(long long)bb * N + j] < 0.5f) e = kMaskFillLog2;
            if (j < nkt * 64) sKey[j] = e;
            const bool all = __all_sync(0xffffffffu, e == 0.f);
            if (lane == 0) sWarpValid[i * 4 + w] = all ? 1 : 0;
          }
          asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");
        }
        float mrow[4], lrow[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          mrow[h] = -INFINITY;
          lrow[h] = 0.f;
        }
        for (int kt = 0; kt < nkt_s; ++kt) {
          const int2 wv = *reinterpret_cast<const int2*>(sWarpValid + (kt >> 1) * 4 + (kt & 1) * 2);
          const bool all_valid = (wv.x & wv.y) != 0;
          const float* keyp = sKey + kt * 64;
#pragma unroll
          for (int h = 0; h < 4; ++h, ++G) {
            mbar_wait(s_full, G & 1);
            tc_fence_after();
            // Every item starts on the fast path (one pass over the scores against a lazily kept reference max); tiles with
            // masked keys apply the key table to each chunk first and keep all their exponentials on MUFU (the polynomial's
            // clamp at -24 would weight a masked key with 2^-24 instead of 0).  The two-pass exact path is the fallback for
            // a row whose scores outgrow the reference max by more than the lazy bound.
            bool exact = false;
            auto fast_item = [&](auto masked_c) {
              constexpr bool kMasked = decltype(masked_c)::value;
              constexpr int PM = kMasked ? 0 : kPM;
              uint32_t sa[16], sb[16];
              tmem_ld16(tS, sa);
              tmem_ld_wait16(sa);
              tmem_ld16(tS + 16, sb);
              if (kMasked) g4_mask_chunk(sa, keyp, 0);
              if (kt == 0) {
                // first tile of a unit: the reference max comes from the row's first 16 scores instead of a full max pass
                // (the exact path cost ~1.5 fast items on 4 of the 32 items of a unit)
                float m0 = -INFINITY;
#pragma unroll
                for (int j = 0; j < 16; j += 2) m0 = fmax3(m0, __uint_as_float(sa[j]), __uint_as_float(sa[j + 1]));
                mrow[h] = m0;
              }
              const uint64_t nm2 = pack_f2(-mrow[h], -mrow[h]);
              uint64_t racc[2] = {0ull, 0ull};
              float rm[2] = {-INFINITY, -INFINITY};
              g4_chunk<PM, !kSumCheck>(sa, nm2, racc, rm, sP_row, t, 0, G >= 1 ? pv : nullptr, (G - 1) & 1);
              tmem_ld_wait16(sb);
              tmem_ld16(tS + 32, sa);
              if (kMasked) g4_mask_chunk(sb, keyp, 1);
              g4_chunk<PM, !kSumCheck>(sb, nm2, racc, rm, sP_row, t, 1);
              tmem_ld_wait16(sa);
              tmem_ld16(tS + 48, sb);
              if (kMasked) g4_mask_chunk(sa, keyp, 2);
              g4_chunk<PM, !kSumCheck>(sa, nm2, racc, rm, sP_row, t, 2);
              tmem_ld_wait16(sb);
              if (kMasked) g4_mask_chunk(sb, keyp, 3);
              // the last chunk is in registers: decide NOW whether a score pushes P past 2^14 (then the item is redone on
              // the exact path; warp-uniform, the TMEM rescale there is warp-collective) -- otherwise S is released a
              // quarter of an item before P is complete, so S_{G+1} is ready when this item ends.  Chunks 0-2: through
              // their row sum (a P above 2^14 makes the sum exceed it); last chunk: through its raw scores.
              bool redo;
              {
                float p0, p1;
                unpack_f2(fadd2(racc[0], racc[1]), p0, p1);
                float m3 = kSumCheck ? -INFINITY : fmaxf(rm[0], rm[1]);
#pragma unroll
                for (int j = 0; j < 16; j += 2) m3 = fmax3(m3, __uint_as_float(sb[j]), __uint_as_float(sb[j + 1]));
                redo = __any_sync(0xffffffffu, (m3 - mrow[h] > kG4LazyBound) || (kSumCheck && !(p0 + p1 <= 16384.0f)));
              }
              if (!redo) {
                tc_fence_before();
                mbar_arrive(sc);
                g4_chunk<PM, !kSumCheck>(sb, nm2, racc, rm, sP_row, t, 3);
                float r0, r1;
                unpack_f2(fadd2(racc[0], racc[1]), r0, r1);
                lrow[h] += r0 + r1;
              }
              return redo;
            };
            if (all_valid) exact = fast_item(std::false_type{});
            else exact = fast_item(std::true_type{});
            if (exact) {
              // pass 1: row max of the (masked) scores
              float tmax = -INFINITY;
#pragma unroll 1
              for (int c = 0; c < 4; ++c) {
                uint32_t s[16];
                tmem_ld16(tS + c * 16, s);
                tmem_ld_wait16(s);
                if (!all_valid) g4_mask_chunk(s, keyp, c);
#pragma unroll
                for (int j = 0; j < 8; ++j) tmax = fmax3(tmax, __uint_as_float(s[2 * j]), __uint_as_float(s[2 * j + 1]));
              }
              const float m_new = fmaxf(mrow[h], tmax);
              if (kt > 0) {
                const float alpha = ex2_approx(mrow[h] - m_new);
                lrow[h] *= alpha;
                // O_h of this row lives in this thread's TMEM lane; every earlier P.V of this group is complete (see
                // the wait on pv above)
                uint32_t ov[16];
                tmem_ld16(tO + h * 16, ov);
                tmem_ld_wait16(ov);
#pragma unroll
                for (int c = 0; c < 16; ++c) ov[c] = __float_as_uint(__uint_as_float(ov[c]) * alpha);
                tmem_st16(tO + h * 16, ov);
                tmem_st_wait();
              }
              mrow[h] = m_new;
              // pass 2: P = exp2(t - m_new) (MUFU only: the polynomial's clamp would weight masked keys with 2^-24)
              const uint64_t nm2 = pack_f2(-m_new, -m_new);
              uint64_t racc[2] = {0ull, 0ull};
              float rm[2] = {-INFINITY, -INFINITY};
#pragma unroll 1
              for (int c = 0; c < 4; ++c) {
                uint32_t s[16];
                tmem_ld16(tS + c * 16, s);
                tmem_ld_wait16(s);
                if (!all_valid) g4_mask_chunk(s, keyp, c);
                g4_chunk<0>(s, nm2, racc, rm, sP_row, t, c);
              }
              float r0, r1;
              unpack_f2(fadd2(racc[0], racc[1]), r0, r1);
              lrow[h] += r0 + r1;
              tc_fence_before();
              mbar_arrive(sc);
            }
            tc_fence_before();  // S reads (and a possible O rescale) are done before the UMMA warp moves on
            fence_proxy_async_smem();
            mbar_arrive(pr);
          }
        }
        // ---- unit epilogue: O (4 x 16 fp32, complete once the last P.V is), normalise, gate, store
        mbar_wait(pv, (G - 1) & 1);
        tc_fence_after();
        {
          // rows [32 w, 32 w + 32) of the P tile are private to this warp and idle: 4 KB slice for the coalesced load
          // of the gate rows / store of the output rows
          uint8_t* slice = sP + w * 4096;
          const int row0 = qt * 128 + w * 32;
          const long long grow = ((long long)seq * N + row0) * 64;
          uint4 gv[8];
          warp_load_rows128(slice, lane, gv, g_gate + grow, 128, N - row0);
          uint4 ovv[8];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            uint32_t ov[16];
            tmem_ld16(tO + h * 16, ov);
            tmem_ld_wait16(ov);
            const float inv = 1.0f / lrow[h];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const __half2* g2 = reinterpret_cast<const __half2*>(&gv[h * 2 + half]);
              uint32_t* o32 = reinterpret_cast<uint32_t*>(&ovv[h * 2 + half]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 gf = __half22float2(g2[e]);
                o32[e] = pack_half2(__uint_as_float(ov[half * 8 + 2 * e]) * inv * gf.x,
                                    __uint_as_float(ov[half * 8 + 2 * e + 1]) * inv * gf.y);
              }
            }
          }
          if (!kFused) {
            tc_fence_before();
            mbar_arrive(o_read);  // the O region may be overwritten by the next unit
            warp_store_rows128(slice, lane, ovv, og + grow, 128, N - row0);
          } else {
            // gated rows -> A tile (this thread's row of the idle P tile), out-projection on the tensor core into the
            // O columns (every thread of the group has read its O: group barrier), pair rows updated with full lines
            __syncwarp();  // the slice (= this warp's rows of the P tile) is done as a load stage
#pragma unroll
            for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(sP + sw128_offset(t, c)) = ovv[c];
            fence_proxy_async_smem();
            tc_fence_before();
            asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");
            if (w == 0) {
              tc_fence_after();
              if (elect_one()) {
                const uint32_t idesc = umma_idesc_f16(128, 64);
                umma_kblock(tmem + 64, smem_u32(sP), smem_u32(sWo), idesc, false);
                umma_kblock(tmem + 64, smem_u32(sP), smem_u32(sWo) + 8192, idesc, true);
                umma_commit(&b[6]);
              }
              __syncwarp();
            }
            // this warp's 32 pair rows (tokens row0 ..): contiguous rows in "starting" mode, stride N rows in "ending"
            const int bb = seq / N, ss = seq - bb * N;
            const long long prow = op.transposed ? (((long long)bb * N + row0) * N + ss) : ((long long)seq * N + row0);
            const long long pitch = (op.transposed ? (long long)N : 1LL) * 64 * 4;
            mbar_wait(&b[6], Ug & 1);
            tc_fence_after();
#pragma unroll
            for (int p = 0; p < 2; ++p) {
              uint4 res[8];
              if (op.residual) {
                warp_load_rows128(slice, lane, res, op.pair + prow * 64 + p * 32, pitch, N - row0);
              } else {
#pragma unroll
                for (int c = 0; c < 8; ++c) res[c] = make_uint4(0, 0, 0, 0);
              }
              uint32_t acc[32];
              tmem_ld32(tO + p * 32, acc);
              tmem_ld_wait();
              if (p == 1) {
                tc_fence_before();
                mbar_arrive(o_read);  // the O columns may be overwritten by the next unit
              }
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                uint4& r = res[c];
                r.x = __float_as_uint(__uint_as_float(r.x) + __uint_as_float(acc[4 * c + 0]) + sBo[p * 32 + 4 * c + 0]);
                r.y = __float_as_uint(__uint_as_float(r.y) + __uint_as_float(acc[4 * c + 1]) + sBo[p * 32 + 4 * c + 1]);
                r.z = __float_as_uint(__uint_as_float(r.z) + __uint_as_float(acc[4 * c + 2]) + sBo[p * 32 + 4 * c + 2]);
                r.w = __float_as_uint(__uint_as_float(r.w) + __uint_as_float(acc[4 * c + 3]) + sBo[p * 32 + 4 * c + 3]);
              }
              warp_store_rows128(slice, lane, res, op.dst + prow * 64 + p * 32, pitch, N - row0);
            }
            ++Ug;
          }
          __syncwarp();  // the slice is P again from here on
        }
      }
    }
  } else if (warp < 20) {
    // ------------------------------------------------------------------ UMMA warp of group g
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    const int g = warp - 16;
    uint64_t* b = gb + g * 8;
    uint64_t* q_full = &b[0];
    uint64_t* q_empty = &b[1];
    uint64_t* s_full = &b[2];
    uint64_t* pr = &b[3];
    uint64_t* pv = &b[4];
    uint64_t* o_read = &b[5];
    uint64_t* sc = &b[7];
    const uint32_t tmem = *tmem_slot + g * 128;
    const uint64_t dq = umma_desc_sw128(smem_u32(sm + G4Smem::kQ + g * 16384));
    const uint64_t dp = umma_desc_sw128(smem_u32(sm + G4Smem::kP + g * 16384));
    const uint64_t dk0 = umma_desc_sw128(smem_u32(sK));
    const uint64_t dv0 = umma_desc_sw128(smem_u32(sV));
    const uint32_t idesc_s = umma_idesc_f16(128, 64), idesc_o = umma_idesc_f16(128, 16);
    int G = 0, U = 0;  // items / units of this group so far
    int gt_base = 0;   // K / V tiles loaded before this round (the ring position)
    for (int ks = 0; ks < nseq_cta; ++ks) {
      bool padded_seq;
      const int nkt_s = g4_eff_tiles<kRagged>(mask, sNkEff, (int)blockIdx.x + ks * (int)gridDim.x, N, nkt, padded_seq);
      const int n_items_s = nkt_s * 4;
      for (int rr = 0; rr < nrr; ++rr, gt_base += nkt_s) {
        const int gt0 = gt_base;
        if (rr * 4 + g >= nqt) {
          // no unit: still release the ring slots, paced by the loads (an arrival must land in the slot's current phase)
          for (int kt = 0; kt < nkt_s; ++kt) {
            const int gt = gt0 + kt, slot = gt % kG4Ring;
            mbar_wait(&k_full[slot], (gt / kG4Ring) & 1);
            mbar_wait(&v_full[slot], (gt / kG4Ring) & 1);
            if (elect_one()) {
              mbar_arrive(&k_empty[slot]);
              mbar_arrive(&v_empty[slot]);
            }
            __syncwarp();
          }
          continue;
        }
        mbar_wait(q_full, U & 1);
        // first S of the unit (its S buffer was released by the previous unit's last pass, or never used)
        if (G >= 1) mbar_wait(sc, (G - 1) & 1);
        mbar_wait(&k_full[gt0 % kG4Ring], (gt0 / kG4Ring) & 1);
        tc_fence_after();
        if (elect_one()) {
          umma_f16(tmem, dq, dk0 + (gt0 % kG4Ring) * (8192 >> 4), idesc_s, 0u);
          umma_commit(s_full);
        }
        __syncwarp();
        for (int it = 0; it < n_items_s; ++it, ++G) {
          const int h = it & 3, kt = it >> 2, gt = gt0 + kt, slot = gt % kG4Ring;
          // ---- everything that does not depend on P_G is prepared BEFORE waiting for it: the softmax threads of this
          // group stall from their last arrival until S_{G+1} / P.V_G are issued, so the path behind the wait is short
          const bool have_next = it + 1 < n_items_s;
          const int h1 = (it + 1) & 3, gt1 = gt0 + ((it + 1) >> 2), slot1 = gt1 % kG4Ring;
          const uint64_t dq1 = dq + 2 * h1, dk1 = dk0 + slot1 * (8192 >> 4) + 2 * h1;
          const uint32_t tO = tmem + 64 + h * 16;
          const uint64_t dv = dv0 + slot * (8192 >> 4) + h * (2048 >> 4);
          const bool last_s = (it + 1 == n_items_s - 1);
          if (have_next && h1 == 0) mbar_wait(&k_full[slot1], (gt1 / kG4Ring) & 1);
          // a unit's first tile overwrites the O region (every thread has read the previous unit's O)
          if (it == 0 && U >= 1) mbar_wait(o_read, (U - 1) & 1);
          if (h == 0) mbar_wait(&v_full[slot], (gt / kG4Ring) & 1);
          mbar_wait(sc, G & 1);  // S_G consumed (a quarter of an item before P_G is complete)
          tc_fence_after();
          if (have_next && elect_one()) {  // S of the next item: ready when the softmax threads finish this one
            umma_f16(tmem, dq1, dk1, idesc_s, 0u);
            umma_commit(s_full);
            if (h1 == 3) umma_commit(&k_empty[slot1]);  // last read of this K tile by this group
            if (last_s) umma_commit(q_empty);           // last read of the Q tile
          }
          __syncwarp();
          mbar_wait(pr, G & 1);  // P_G written
          tc_fence_after();
          if (elect_one()) {
            // O_h (+)= P_G V_h
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(tO, dp + 2 * k, dv + 2 * k, idesc_o, (kt > 0 || k > 0) ? 1u : 0u);
            umma_commit(pv);
            if (h == 3) umma_commit(&v_empty[slot]);
          }
          __syncwarp();
        }
        ++U;
      }
    }
  } else if (warp == 20) {
    // ------------------------------------------------------------------ K / V ring TMA warp
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    int gt = 0;
    for (int ks = 0; ks < nseq_cta; ++ks) {
      const int seq = (int)blockIdx.x + ks * (int)gridDim.x;
      bool padded_seq;
      const int nkt_s = g4_eff_tiles<kRagged>(mask, sNkEff, seq, N, nkt, padded_seq);
      for (int rr = 0; rr < nrr; ++rr) {
        for (int kt = 0; kt < nkt_s; ++kt, ++gt) {
          const int slot = gt % kG4Ring;
          if (gt >= kG4Ring) {
            mbar_wait(&k_empty[slot], ((gt / kG4Ring) - 1) & 1);
            mbar_wait(&v_empty[slot], ((gt / kG4Ring) - 1) & 1);
          }
          if (elect_one()) {
            mbar_expect_tx(&k_full[slot], 8192);
            tma_load_3d(sK + slot * 8192, &map_k, &k_full[slot], 0, kt * 64, seq);
            mbar_expect_tx(&v_full[slot], 8192);
            tma_load_3d(sV + slot * 8192, &map_vt, &v_full[slot], kt * 64, 0, seq);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 21) {
    // ------------------------------------------------------------------ Q TMA warp (all four groups)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    int U[4] = {0, 0, 0, 0};
    for (int ks = 0; ks < nseq_cta; ++ks) {
      const int seq = (int)blockIdx.x + ks * (int)gridDim.x;
      const bool padded_seq = kRagged && mask[seq] < 0.5f;
      for (int rr = 0; rr < nrr; ++rr) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int qt = rr * 4 + g;
          if (qt >= nqt) continue;
          uint64_t* b = gb + g * 8;
          if (U[g] >= 1) mbar_wait(&b[1], (U[g] - 1) & 1);  // the previous unit's last S is complete
          if (padded_seq) {
            // uniform softmax of a padded sequence = the fast path on Q = 0 (see g4_eff_tiles): zero the tile by hand
            uint4* qz = reinterpret_cast<uint4*>(sm + G4Smem::kQ + g * 16384);
            for (int i = lane; i < 1024; i += 32) qz[i] = make_uint4(0, 0, 0, 0);
            fence_proxy_async_smem();
            __syncwarp();
            if (elect_one()) mbar_arrive(&b[0]);
          } else if (elect_one()) {
            mbar_expect_tx(&b[0], 16384);
            tma_load_3d(sm + G4Smem::kQ + g * 16384, &map_q, &b[0], 0, qt * 128, seq);
          }
          __syncwarp();
          ++U[g];
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(*tmem_slot, 512);
}

static int g4_launch(const PairDims& d, const float* mask, const __half* q, const __half* k, const __half* g,
                     const __half* vt, __half* og, const G4OutProj* op, cudaStream_t s) {
  const int N = d.N, Np = plane_ld(N);
  const long long nseq = (long long)d.B * N;
  CUtensorMap mq, mk, mv;
  TmaDims t;
  // q: [seq][tok][64], box 128 tokens; k: box 64 keys
  t.size[0] = 64; t.size[1] = (uint64_t)N; t.size[2] = (uint64_t)nseq; t.size[3] = 1;
  t.stride[0] = 128; t.stride[1] = (uint64_t)N * 128; t.stride[2] = 0;
  t.box[0] = 64; t.box[1] = 128; t.box[2] = 1; t.box[3] = 1;
  if (make_tensor_map(&mq, q, 2, 3, t, true)) return 1;
  t.box[1] = 64;
  if (make_tensor_map(&mk, k, 2, 3, t, true)) return 1;
  // vt: [seq][64 (h,c)][tok], tok contiguous, row stride Np; one box = all 64 rows x 64 keys
  t.size[0] = (uint64_t)N; t.size[1] = 64; t.size[2] = (uint64_t)nseq;
  t.stride[0] = (uint64_t)Np * 2; t.stride[1] = (uint64_t)Np * 2 * 64;
  t.box[0] = 64; t.box[1] = 64; t.box[2] = 1;
  if (make_tensor_map(&mv, vt, 2, 3, t, true)) return 1;
  const int nkt = (N + 63) / 64;
  const int key_bytes = (nkt * 64 * 4 + nkt * 16 + 127) & ~127;
  const int smem = 1024 + G4Smem::kKey + 4 * key_bytes + (4 * kG4Ring + 32) * 8 + 16;
  PRD_REQUIRE(smem <= 227 * 1024, "triattn_flash_g4: N=%d needs %d B of shared memory", N, smem);
  PRD_REQUIRE(nseq <= 2147483647LL / 64, "triattn_flash_g4: too many sequences");
  const int grid = (int)std::min<long long>(nseq, kNumSMs);
  if (op) {
    PRD_REQUIRE(d.CZ == 64, "triattn_flash_g4: fused output projection needs pair_dim 64");
    PRD_CUDA_OK(cudaFuncSetAttribute(triattn_flash_g4_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    triattn_flash_g4_kernel<true><<<grid, kG4Threads, smem, s>>>(mq, mk, mv, mask, g, og, N, (int)nseq, *op);
  } else {
    // share of the exponentials on the FMA-pipe polynomial: 1/4 (default), PRD_FLASH_POLY=1 -> 3/8, =2 -> 1/2 (A/B timing)
    static const int poly = getenv("PRD_FLASH_POLY") ? atoi(getenv("PRD_FLASH_POLY")) : 0;
    auto kern = poly == 1 ? triattn_flash_g4_kernel<false, 0xA8> : poly == 2 ? triattn_flash_g4_kernel<false, 0xAA>
                                                                             : triattn_flash_g4_kernel<false, 0x88>;
    PRD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    // kRagged = false: the variant without per-sequence tile counts (all-valid batches; correct for any mask, it just
    // takes the general masked path on every tile of a ragged one).  PRD_FLASH_RAGGED=0 forces it (A/B timing).
    static const int ragged_env = !(getenv("PRD_FLASH_RAGGED") && getenv("PRD_FLASH_RAGGED")[0] == '0');
    static const int sumcheck = !(getenv("PRD_FLASH_SUMCHECK") && getenv("PRD_FLASH_SUMCHECK")[0] == '0');
    if ((!ragged_env || d.all_valid) && !sumcheck) {
      auto kern0 = triattn_flash_g4_kernel<false, 0x88, false, false>;
      PRD_CUDA_OK(cudaFuncSetAttribute(kern0, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      PRD_CUDA_OK(launch_pdl(kern0, grid, kG4Threads, smem, s, mq, mk, mv, mask, g, og, N, (int)nseq, G4OutProj{}));
    } else if (!ragged_env || d.all_valid) {
      auto kern0 = triattn_flash_g4_kernel<false, 0x88, false>;
      PRD_CUDA_OK(cudaFuncSetAttribute(kern0, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      PRD_CUDA_OK(launch_pdl(kern0, grid, kG4Threads, smem, s, mq, mk, mv, mask, g, og, N, (int)nseq, G4OutProj{}));
    } else {
      PRD_CUDA_OK(launch_pdl(kern, grid, kG4Threads, smem, s, mq, mk, mv, mask, g, og, N, (int)nseq, G4OutProj{}));
    }
  }
  PRD_LAUNCHED();
  return 0;
}

}  // namespace

int triattn_flash_g4(const PairDims& d, const float* mask, const __half* q, const __half* k, const __half* g,
                     const __half* vt, __half* og, cudaStream_t s) {
  return g4_launch(d, mask, q, k, g, vt, og, nullptr, s);
}

// Attention core + output projection + residual in one kernel (replaces triattn_flash + triattn_out).
int triattn_flash_out_g4(const PairDims& d, const float* mask, const __half* q, const __half* k, const __half* g,
                         const __half* vt, const float* pair, float* dst, int residual, int mode, const __half* w_o,
                         const float* b_o, cudaStream_t s) {
  G4OutProj op{pair, dst, w_o, b_o, residual, mode};
  return g4_launch(d, mask, q, k, g, vt, nullptr, &op, s);
}

// The four-group kernel keeps all groups busy when the query tiles per sequence are a multiple of four, but since its
// first-tile / masked-tile fast paths (round 2) it also beats the two-group kernel when one or two groups idle: N = 300
// 3.01 -> 2.72 ms per step, N = 140 (README dims) 1.91 -> 1.77 ms (bench.py --workload config2 / config1, PRD_FLASH_G4=0 A/B).
bool triattn_flash_g4_applies(const PairDims& d) {
  const char* force = getenv("PRD_FLASH_G4");
  return force ? (force[0] == '1') : d.N <= 2048;
}

}  // namespace prd
