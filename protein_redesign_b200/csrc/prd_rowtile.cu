// Pair-row tile kernels: every op of the pair stack whose contraction dim is the channel dim.
//   pair_transition   FoldingBlock.pair_fc                       (modules.py:321-326,342)
//   trimul_in/out     TriangleMultiplication projections/gating  (modules.py:262-274)
//   triattn_proj/out  TriangleAttention q,k,v,gate / out_proj    (modules.py:185-225,236-243)
// See prd_rowtile.cuh for the execution model.
#include "prd_kernels.h"
#include <stdlib.h>
#include <string.h>
#include "prd_rowtile.cuh"

namespace prd {

namespace {

struct RowMap {
  int N;
  long long NN;
  int transposed;  // 0: logical row (b,s,t) reads pair[b,s,t]; 1: reads pair[b,t,s]
  __device__ __forceinline__ void decompose(long long r, int& b, int& s, int& t) const {
    int rem;
    if (r < 0x7fffffffLL && NN < 0x7fffffffLL) {  // 32-bit division (the 64-bit one costs ~100 instructions)
      const unsigned r32 = static_cast<unsigned>(r), nn32 = static_cast<unsigned>(NN);
      b = static_cast<int>(r32 / nn32);
      rem = static_cast<int>(r32 - static_cast<unsigned>(b) * nn32);
    } else {
      b = static_cast<int>(r / NN);
      rem = static_cast<int>(r - (long long)b * NN);
    }
    s = rem / N;
    t = rem - s * N;
  }
  __device__ __forceinline__ long long src_row(int b, int s, int t) const {
    return transposed ? ((long long)b * NN + (long long)t * N + s) : ((long long)b * NN + (long long)s * N + t);
  }
};

template <typename Kern>
int set_smem(Kern k, int bytes) {
  return check_cuda(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes), "cudaFuncSetAttribute");
}

// Tensor map over the fp32 pair tensor [B][N][N][64] for tiles of 128 tokens of one sequence:
//   mode 0 (row (b,s,tok) = pair[b,s,tok]): dims (c, tok, s, b), box (32, 128, 1, 1)  -> coords (32 h, tok0, s, b)
//   mode 1 (row (b,s,tok) = pair[b,tok,s]): dims (c, s, tok, b), box (32, 1, 128, 1)  -> coords (32 h, s, tok0, b)
inline int make_pair_tile_map(CUtensorMap* m, const float* pair, int B, int N, int mode) {
  TmaDims t;
  t.size[0] = 64; t.size[1] = (uint64_t)N; t.size[2] = (uint64_t)N; t.size[3] = (uint64_t)B;
  t.stride[0] = 256; t.stride[1] = (uint64_t)N * 256; t.stride[2] = (uint64_t)N * N * 256;
  t.box[0] = 32; t.box[1] = mode ? 1 : 128; t.box[2] = mode ? 128 : 1; t.box[3] = 1;
  return make_tensor_map(m, pair, 4, 4, t, true);
}

inline int grid_for(long long tiles, int ctas_per_sm) {
  long long g = (long long)kNumSMs * ctas_per_sm;
  return static_cast<int>(tiles < g ? tiles : g);
}

}  // namespace

// =========================================================================================
// pair_fc: dst = [pair +] W2 relu(W1 LN(pair) + b1) + b2
// Both weights are fp16 hi + lo pairs (w1: [2][HID][CZ], w2: [2][CZ][HID]); the hidden layer is
// processed in two halves so that A, the hidden tile, both split weights and one row stage fit
// in shared memory.  The row itself stays in registers for the residual; the single stage is
// re-armed for the next tile as soon as the row has been read (prefetch), and the hidden-tile
// buffer doubles as the output staging area.
// =========================================================================================
template <int CZ, int HID>
struct PairFcSmem {
  static constexpr int HH = HID / 2;        // hidden columns per half
  static constexpr int KBHH = HH / 64;      // K-blocks per half
  static constexpr int KBH = HID / 64;
  static constexpr int kHBytes = (KBHH * 16384 > RowStage<CZ>::kBytes ? KBHH * 16384 : RowStage<CZ>::kBytes + 1023) / 1024 * 1024;
  static constexpr int kA = 0;
  static constexpr int kH = kA + 16384;
  static constexpr int kW1 = kH + kHBytes;                 // hi, lo: 2 * HID * 128
  static constexpr int kW2 = kW1 + 2 * HID * 128;          // hi, lo: 2 * KBH * CZ * 128
  static constexpr int kStage = kW2 + 2 * KBH * CZ * 128;
  static constexpr int kBias = kStage + RowStage<CZ>::kBytes;
  static constexpr int kBars = kBias + (HID + CZ) * 4;
  static constexpr int kTotal = kBars + 64 + 1024;
  static constexpr int kTmemCols = (HID + CZ) <= 256 ? 256 : 512;
};

template <int CZ, int HID>
__global__ void __launch_bounds__(128, 1)
pair_transition_kernel(const float* pair, float* dst, int residual, long long R, const __half* __restrict__ w1,
                       const float* __restrict__ b1, const __half* __restrict__ w2, const float* __restrict__ b2) {
  extern __shared__ uint8_t raw[];
  using L = PairFcSmem<CZ, HID>;
  uint8_t* sm = smem_align1024(raw);
  uint8_t* sA = sm + L::kA;
  uint8_t* sH = sm + L::kH;
  uint8_t* sW1 = sm + L::kW1;
  uint8_t* sW2 = sm + L::kW2;
  uint8_t* sSt = sm + L::kStage;
  float* sB1 = reinterpret_cast<float*>(sm + L::kBias);
  float* sB2 = sB1 + HID;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + L::kBars);
  uint64_t* mma_bar = full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);

  const int t = threadIdx.x, warp = t >> 5;
  if (t == 0) {
    mbar_init(full, kTileRows);
    mbar_init(mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, L::kTmemCols);
  load_weight_kblocks(sW1, w1, HID, CZ, CZ, t, 128);
  load_weight_kblocks(sW1 + HID * 128, w1 + HID * CZ, HID, CZ, CZ, t, 128);
  load_weight_kblocks(sW2, w2, CZ, HID, HID, t, 128);
  load_weight_kblocks(sW2 + L::KBH * CZ * 128, w2 + CZ * HID, CZ, HID, HID, t, 128);
  for (int i = t; i < HID; i += 128) sB1[i] = b1[i];
  for (int i = t; i < CZ; i += 128) sB2[i] = b2[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const uint32_t tm_out = HID;  // accumulator columns of the second GEMM

  const long long num_tiles = (R + kTileRows - 1) / kTileRows;
  uint32_t mma_phase = 0;
  long long tile = blockIdx.x;
  if (tile < num_tiles) {
    const long long r = tile * kTileRows + t;
    issue_row_load<CZ>(sSt, t, pair + r * CZ, r < R, full);
  }
  for (int it = 0; tile < num_tiles; tile += gridDim.x, ++it) {
    mbar_wait(full, it & 1);
    const long long r = tile * kTileRows + t;
    const bool valid = r < R;
    float x[CZ];
    if (valid) {
      read_row<CZ>(stage_row<CZ>(sSt, t), x);
    } else {
#pragma unroll
      for (int i = 0; i < CZ; ++i) x[i] = 0.f;
    }
    {  // the stage slot of this thread is free again: prefetch the next tile's row
      const long long next = tile + gridDim.x;
      if (next < num_tiles) {
        const long long rn = next * kTileRows + t;
        issue_row_load<CZ>(sSt, t, pair + rn * CZ, rn < R, full);
      }
    }
    {
      float y[CZ];
#pragma unroll
      for (int i = 0; i < CZ; ++i) y[i] = x[i];
      layernorm_inplace<CZ>(y);
      store_a_row<CZ>(sA, t, y);
    }
    bulk_wait_read0();  // previous tile's output rows (staged in sH) have been read by the bulk store
    sync_before_mma();
    if (t < 32) {  // warp-uniform issue: UMMA operands stay in uniform registers
      tc_fence_after();
      if (elect_one()) {
        umma_multi(tmem, smem_u32(sA), smem_u32(sW1), 1, HID * 128, umma_idesc_f16(128, HID), false);
        umma_multi(tmem, smem_u32(sA), smem_u32(sW1 + HID * 128), 1, HID * 128, umma_idesc_f16(128, HID), true);
        umma_commit(mma_bar);
      }
      __syncwarp();
    }
    mbar_wait(mma_bar, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
#pragma unroll 1
      for (int c = 0; c < L::HH / 32; ++c) {
        uint32_t acc[32];
        tmem_ld32(tm_lane + half * L::HH + c * 32, acc);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(__uint_as_float(acc[j]) + sB1[half * L::HH + c * 32 + j], 0.f);
        store_a_cols32(sH, t, c * 32, v);
      }
      sync_before_mma();
      if (t < 32) {  // warp-uniform issue: UMMA operands stay in uniform registers
        tc_fence_after();
        if (elect_one()) {
          const uint32_t idesc = umma_idesc_f16(128, CZ);
          for (int kb = 0; kb < L::KBHH; ++kb)
            umma_kblock(tmem + tm_out, smem_u32(sH) + kb * 16384, smem_u32(sW2) + (half * L::KBHH + kb) * CZ * 128, idesc,
                        half > 0 || kb > 0);
          for (int kb = 0; kb < L::KBHH; ++kb)
            umma_kblock(tmem + tm_out, smem_u32(sH) + kb * 16384,
                        smem_u32(sW2) + (L::KBH + half * L::KBHH + kb) * CZ * 128, idesc, true);
          umma_commit(mma_bar);
        }
        __syncwarp();
      }
      mbar_wait(mma_bar, mma_phase);  // the hidden tile is rewritten next: its UMMAs must be done
      mma_phase ^= 1;
      tc_fence_after();
    }
    float* my = stage_row<CZ>(sH, t);  // output staging aliases the hidden tile (its UMMAs are complete)
#pragma unroll
    for (int c = 0; c < CZ / 32; ++c) {
      uint32_t acc[32];
      tmem_ld32(tm_lane + tm_out + c * 32, acc);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 o;
        o.x = __uint_as_float(acc[j + 0]) + sB2[c * 32 + j + 0];
        o.y = __uint_as_float(acc[j + 1]) + sB2[c * 32 + j + 1];
        o.z = __uint_as_float(acc[j + 2]) + sB2[c * 32 + j + 2];
        o.w = __uint_as_float(acc[j + 3]) + sB2[c * 32 + j + 3];
        if (residual) {
          o.x += x[c * 32 + j + 0];
          o.y += x[c * 32 + j + 1];
          o.z += x[c * 32 + j + 2];
          o.w += x[c * 32 + j + 3];
        }
        *reinterpret_cast<float4*>(my + c * 32 + j) = o;
      }
    }
    fence_proxy_async_smem();
    if (valid) bulk_s2g(dst + r * CZ, my, CZ * 4);
    bulk_commit();
    // TMEM columns and sA are rewritten by the next tile only after its sync_before_mma
  }
  bulk_wait0();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, L::kTmemCols);
}

template <int CZ, int HID>
static int launch_pair_transition(const PairDims& d, const float* pair, float* dst, int residual, const __half* w1,
                                  const float* b1, const __half* w2, const float* b2, cudaStream_t s) {
  using L = PairFcSmem<CZ, HID>;
  static_assert(L::kTotal <= 227 * 1024, "pair_fc shared memory budget");
  auto kern = pair_transition_kernel<CZ, HID>;
  if (set_smem(kern, L::kTotal)) return 1;
  const long long R = (long long)d.B * d.N * d.N;
  const long long tiles = (R + kTileRows - 1) / kTileRows;
  kern<<<grid_for(tiles, 1), 128, L::kTotal, s>>>(pair, dst, residual, R, w1, b1, w2, b2);
  PRD_LAUNCHED();
  return 0;
}

// -----------------------------------------------------------------------------------------
// pair_dim 64 / hidden 256: warp-specialised pipeline (the kernel above serialises LayerNorm, two GEMMs and two
// epilogues per tile on four warps).  Persistent CTA, one per SM, 16 warps:
//   warps 0-3    row warps: thread = pair row.  Rows arrive by TMA (two swizzled boxes of [128 rows x 32 floats], issued
//                a tile ahead by warp 14), LayerNorm -> A tile (2-slot ring); later the output epilogue of the same tile
//                (accumulator + b2 + residual from registers), full-line stores.  LayerNorm of tile i+1 is done BEFORE
//                the output of tile i, so the tensor core never waits for it.
//   warps 4-11   mid warps: hidden quarter q (64 columns): D1_q + b1 -> ReLU -> fp16 -> H tile (2-slot ring);
//                warps 4-7 take the even quarters, 8-11 the odd ones
//   warps 12/13  UMMA issue of the first / second GEMM (M1_q: D1[q&1] = A W1_q^T with the hi and lo weight halves side by
//                side along N; M2_q: D2 += H_q W2_q^T), all hand-offs through mbarriers
//   warp 14      TMA loader of the row stage (one 32 KB stage: it is free again as soon as the row warps hold the rows in
//                registers, i.e. the load of tile i+1 overlaps everything tile i still has to do)
//   warp 15      idle (donates registers)
// The second GEMM uses the hi half of W2 only: the hidden activations are already rounded to fp16 and W2's lo half
// contributes 2e-5 of the update at N = 512, 1.5e-4 on the worst known case (tools/precision_pairfc.py) against the 1e-3
// budget; dropping it frees exactly the 32 KB the row stage needs and a quarter of the second GEMM.
// TMEM: D1 quarters 2 x 128 columns (hi | lo), D2 2 x 64 columns.
// -----------------------------------------------------------------------------------------
namespace {
constexpr int kPtThreads = 512;
struct PtSmem {
  static constexpr int kW1 = 0;                        // hi, lo: 2 x [256 x 64] = 64 KB
  static constexpr int kW2 = kW1 + 2 * 256 * 128;      // hi: 4 K-blocks of [64 x 64] = 32 KB
  static constexpr int kA = kW2 + 4 * 64 * 128;        // 2 x 16 KB
  static constexpr int kH = kA + 2 * 16384;            // 2 x 16 KB
  static constexpr int kSl = kH + 2 * 16384;           // 4 row warps x 4 KB
  static constexpr int kStage = kSl + 4 * 4096;        // row stage: 2 boxes x 16 KB
  static constexpr int kBias = kStage + 2 * 16384;     // b1[256], b2[64]
  static constexpr int kWb = kBias + 320 * 4;          // fused attn_bias: W_b [4][64], {sum_c W_b[h][c]}[4], b_b[4]
  static constexpr int kBars = kWb + (4 * 64 + 8) * 4;
  static constexpr int kTotal = kBars + 20 * 8 + 16 + 1024;
};
}  // namespace

__global__ void __launch_bounds__(kPtThreads, 1)
pair_transition_ws_kernel(const __grid_constant__ CUtensorMap map_rows, float* dst, int residual, long long R,
                          const __half* __restrict__ w1, const float* __restrict__ b1, const __half* __restrict__ w2,
                          const float* __restrict__ b2, const float* __restrict__ w_bias, const float* __restrict__ b_bias,
                          float* __restrict__ bias_out, unsigned NN) {
  constexpr int CZ = 64, HID = 256;
  extern __shared__ uint8_t raw[];
  pdl_trigger();
  using L = PtSmem;
  uint8_t* sm = smem_align1024(raw);
  uint8_t* sW1 = sm + L::kW1;
  uint8_t* sW2 = sm + L::kW2;
  uint8_t* sA = sm + L::kA;
  uint8_t* sH = sm + L::kH;
  uint8_t* sStage = sm + L::kStage;
  float* sB1 = reinterpret_cast<float*>(sm + L::kBias);
  float* sB2 = sB1 + HID;
  float* sWb = reinterpret_cast<float*>(sm + L::kWb);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L::kBars);
  uint64_t* a_full = bars;        // [2] 128 row-thread arrivals
  uint64_t* a_empty = bars + 2;   // [2] UMMA commit
  uint64_t* d1_full = bars + 4;   // [2] UMMA commit
  uint64_t* d1_empty = bars + 6;  // [2] 128 mid-thread arrivals (mid set Q&1)
  uint64_t* h_full = bars + 8;    // [2] 128 mid-thread arrivals
  uint64_t* h_empty = bars + 10;  // [2] UMMA commit
  uint64_t* d2_full = bars + 12;  // [2] UMMA commit
  uint64_t* d2_empty = bars + 14; // [2] 128 row-thread arrivals
  uint64_t* st_full = bars + 16;  // TMA complete_tx
  uint64_t* st_empty = bars + 17; // 128 row-thread arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 128);
      mbar_init(&a_empty[i], 1);
      mbar_init(&d1_full[i], 1);
      mbar_init(&d1_empty[i], 128);
      mbar_init(&h_full[i], 128);
      mbar_init(&h_empty[i], 1);
      mbar_init(&d2_full[i], 1);
      mbar_init(&d2_empty[i], 128);
    }
    mbar_init(st_full, 1);
    mbar_init(st_empty, 128);
    fence_barrier_init();
    tma_prefetch_desc(&map_rows);
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  // B operands of the first GEMM, 128 rows per hidden quarter q: rows [0,64) = hi weights, [64,128) = lo weights, so one
  // UMMA (N = 128) computes both halves; the mid warps add the two 64-column accumulator halves
  for (int q = 0; q < 4; ++q) {
    load_weight_kblocks(sW1 + q * 16384, w1 + q * 64 * CZ, 64, CZ, CZ, threadIdx.x, kPtThreads);
    load_weight_kblocks(sW1 + q * 16384 + 8192, w1 + HID * CZ + q * 64 * CZ, 64, CZ, CZ, threadIdx.x, kPtThreads);
    load_weight_kblocks(sW2 + q * 8192, w2 + q * 64, CZ, 64, HID, threadIdx.x, kPtThreads);  // hi half only
  }
  for (int i = threadIdx.x; i < HID; i += kPtThreads) sB1[i] = b1[i];
  for (int i = threadIdx.x; i < CZ; i += kPtThreads) sB2[i] = b2[i];
  if (bias_out != nullptr) {
    for (int i = threadIdx.x; i < 4 * CZ; i += kPtThreads) sWb[i] = w_bias[i];
    if (threadIdx.x < 4) {
      float sw = 0.f;
      for (int c = 0; c < CZ; ++c) sw += w_bias[threadIdx.x * CZ + c];
      sWb[4 * CZ + threadIdx.x] = sw;
      sWb[4 * CZ + 4 + threadIdx.x] = b_bias ? b_bias[threadIdx.x] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // everything above touched only weights / shared memory; the predecessor's output is read below
  const uint32_t tmem = *tmem_slot;
  const long long num_tiles = (R + kTileRows - 1) / kTileRows;
  const int my_tiles = (blockIdx.x < num_tiles) ? static_cast<int>((num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

  if (warp < 4) {
    // ------------------------------------------------------------------ row warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    const int t = threadIdx.x;  // row of the tile = TMEM lane
    uint8_t* slice = sm + L::kSl + warp * 4096;
    const uint32_t tm_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    float xc[CZ], xn[CZ];
    auto load_ln = [&](int il, float (&x)[CZ]) {  // rows of local tile il -> x (kept for the residual), LN -> A slot
      mbar_wait(st_full, il & 1);
      read_row_tma64(sStage, t, x);
      fence_proxy_async_smem();  // generic reads of the stage before the async-proxy refill
      mbar_arrive(st_empty);
      // LayerNorm statistics (x itself is kept for the residual)
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < CZ; ++i) s4[i & 3] += x[i];
      const float mean = ((s4[0] + s4[1]) + (s4[2] + s4[3])) * (1.0f / CZ);
      float v4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < CZ; ++i) {
        const float dlt = x[i] - mean;
        v4[i & 3] = fmaf(dlt, dlt, v4[i & 3]);
      }
      const float rstd = rsqrtf(((v4[0] + v4[1]) + (v4[2] + v4[3])) * (1.0f / CZ) + kLnEps);
      const float nmr = -mean * rstd;
      if (il >= 2) mbar_wait(&a_empty[il & 1], ((il >> 1) - 1) & 1);
      uint8_t* a_tile = sA + (il & 1) * 16384;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint4 o;
        o.x = pack_half2(fmaf(x[ch * 8 + 0], rstd, nmr), fmaf(x[ch * 8 + 1], rstd, nmr));
        o.y = pack_half2(fmaf(x[ch * 8 + 2], rstd, nmr), fmaf(x[ch * 8 + 3], rstd, nmr));
        o.z = pack_half2(fmaf(x[ch * 8 + 4], rstd, nmr), fmaf(x[ch * 8 + 5], rstd, nmr));
        o.w = pack_half2(fmaf(x[ch * 8 + 6], rstd, nmr), fmaf(x[ch * 8 + 7], rstd, nmr));
        *reinterpret_cast<uint4*>(a_tile + sw128_offset(t, ch)) = o;
      }
      fence_proxy_async_smem();
      mbar_arrive(&a_full[il & 1]);
    };
    if (my_tiles > 0) load_ln(0, xc);
    for (int il = 0; il < my_tiles; ++il) {
      if (il + 1 < my_tiles) load_ln(il + 1, xn);
      // output epilogue of tile il
      const long long row0 = ((long long)blockIdx.x + (long long)il * gridDim.x) * kTileRows + warp * 32;
      const long long left = R - row0;
      const int rows_valid = left > 32 ? 32 : (left < 0 ? 0 : static_cast<int>(left));
      mbar_wait(&d2_full[il & 1], (il >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        uint32_t acc[32];
        tmem_ld32(tm_lane + 256 + (il & 1) * 64 + p * 32, acc);
        tmem_ld_wait();
        if (p == 1) {
          tc_fence_before();
          mbar_arrive(&d2_empty[il & 1]);
        }
        uint4 o[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            v[e] = __uint_as_float(acc[4 * c + e]) + sB2[p * 32 + 4 * c + e];
            if (residual) v[e] += xc[p * 32 + 4 * c + e];
            xc[p * 32 + 4 * c + e] = v[e];  // the updated row (the input row is not needed any more)
          }
          o[c] = make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
        }
        warp_store_rows128(slice, lane, o, dst + row0 * CZ + p * 32, CZ * 4, rows_valid);
      }
      if (bias_out != nullptr) {
        // the NEXT block's attention bias of this row (FoldingBlock.attn_bias, modules.py:300-304): LayerNorm without
        // affine + H x c_z projection, from the updated row in registers -- the pair tensor is not read again for it.
        // bias_h = rstd (W_h . x - mean sum(W_h)) + b_h
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < CZ; ++i) s4[i & 3] += xc[i];
        const float mean = ((s4[0] + s4[1]) + (s4[2] + s4[3])) * (1.0f / CZ);
        float v4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < CZ; ++i) {
          const float dlt = xc[i] - mean;
          v4[i & 3] = fmaf(dlt, dlt, v4[i & 3]);
        }
        const float rstd = rsqrtf(((v4[0] + v4[1]) + (v4[2] + v4[3])) * (1.0f / CZ) + kLnEps);
        float hb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < CZ; c += 4) {
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const float4 wv = *reinterpret_cast<const float4*>(sWb + h * CZ + c);  // broadcast
            hb[h] = fmaf(xc[c], wv.x, fmaf(xc[c + 1], wv.y, fmaf(xc[c + 2], wv.z, fmaf(xc[c + 3], wv.w, hb[h]))));
          }
        }
        if (lane < rows_valid) {
          const unsigned r = static_cast<unsigned>(row0) + lane;
          const unsigned b = r / NN, ij = r - b * NN;
#pragma unroll
          for (int h = 0; h < 4; ++h)
            bias_out[(static_cast<size_t>(b) * 4 + h) * NN + ij] = rstd * (hb[h] - mean * sWb[4 * CZ + h]) + sWb[4 * CZ + 4 + h];
        }
      }
#pragma unroll
      for (int i = 0; i < CZ; ++i) xc[i] = xn[i];
    }
  } else if (warp < 12) {
    // ------------------------------------------------------------------ mid warps: set 0 (warps 4-7) takes the even
    // hidden quarters, set 1 (warps 8-11) the odd ones, so two quarters are in flight at once
    asm volatile("setmaxnreg.dec.sync.aligned.u32 112;");
    const int m = warp - 4;
    const int t = (m & 3) * 32 + lane;  // tile row = TMEM lane
    const int set = m >> 2;
    const uint32_t tm_lane = tmem + (static_cast<uint32_t>((m & 3) * 32) << 16);
    for (int il = 0; il < my_tiles; ++il) {
#pragma unroll 1
      for (int qq = 0; qq < 2; ++qq) {
        const int q = 2 * qq + set, Q = 4 * il + q;  // Q & 1 == set
        mbar_wait(&d1_full[set], (Q >> 1) & 1);
        tc_fence_after();
        if (Q >= 2) mbar_wait(&h_empty[set], ((Q >> 1) - 1) & 1);
#pragma unroll 1
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t acc[32], acl[32];
          tmem_ld32(tm_lane + set * 128 + hf * 32, acc);
          tmem_ld32(tm_lane + set * 128 + 64 + hf * 32, acl);
          tmem_ld_wait();
          if (hf == 1) {
            tc_fence_before();
            mbar_arrive(&d1_empty[set]);
          }
          float v[32];
          const float* bq = sB1 + q * 64 + hf * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf((__uint_as_float(acc[j]) + __uint_as_float(acl[j])) + bq[j], 0.f);
          store_a_cols32(sH + set * 16384, t, hf * 32, v);
        }
        fence_proxy_async_smem();
        mbar_arrive(&h_full[set]);
      }
    }
  } else {
    // ------------------------------------------------------------------ UMMA warps: 12 issues the first GEMM (gated by
    // A tiles and free D1 buffers), 13 the second (gated by H tiles and free D2 buffers); 14 loads rows; 15 idle
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 12) {
      const uint32_t idesc = umma_idesc_f16(128, 128);
      int Q = 0;
      for (int il = 0; il < my_tiles; ++il) {
        mbar_wait(&a_full[il & 1], (il >> 1) & 1);
        for (int q = 0; q < 4; ++q, ++Q) {
          if (Q >= 2) mbar_wait(&d1_empty[Q & 1], ((Q >> 1) - 1) & 1);
          tc_fence_after();
          if (elect_one()) {
            umma_kblock(tmem + (Q & 1) * 128, smem_u32(sA) + (il & 1) * 16384, smem_u32(sW1) + q * 16384, idesc, false);
            umma_commit(&d1_full[Q & 1]);
            if (q == 3) umma_commit(&a_empty[il & 1]);
          }
          __syncwarp();
        }
      }
    } else if (warp == 13) {
      const uint32_t idesc = umma_idesc_f16(128, 64);
      int Q = 0;
      for (int il = 0; il < my_tiles; ++il) {
        if (il >= 2) mbar_wait(&d2_empty[il & 1], ((il >> 1) - 1) & 1);
        for (int q = 0; q < 4; ++q, ++Q) {
          mbar_wait(&h_full[Q & 1], (Q >> 1) & 1);
          tc_fence_after();
          if (elect_one()) {
            umma_kblock(tmem + 256 + (il & 1) * 64, smem_u32(sH) + (Q & 1) * 16384, smem_u32(sW2) + q * 8192, idesc, q > 0);
            umma_commit(&h_empty[Q & 1]);
            if (q == 3) umma_commit(&d2_full[il & 1]);
          }
          __syncwarp();
        }
      }
    } else if (warp == 14) {
      for (int il = 0; il < my_tiles; ++il) {
        if (il >= 1) mbar_wait(st_empty, (il - 1) & 1);
        if (elect_one()) {
          const long long row0 = ((long long)blockIdx.x + (long long)il * gridDim.x) * kTileRows;
          mbar_expect_tx(st_full, 2 * 16384);
          tma_load_2d(sStage, &map_rows, st_full, 0, static_cast<int>(row0));
          tma_load_2d(sStage + 16384, &map_rows, st_full, 32, static_cast<int>(row0));
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static int launch_pair_transition_ws(const PairDims& d, const float* pair, float* dst, int residual, const __half* w1,
                                     const float* b1, const __half* w2, const float* b2, const float* w_bias,
                                     const float* b_bias, float* bias_out, cudaStream_t s) {
  static_assert(PtSmem::kTotal <= 227 * 1024, "pair_fc shared memory budget");
  if (set_smem(pair_transition_ws_kernel, PtSmem::kTotal)) return 1;
  const long long R = (long long)d.B * d.N * d.N;
  const long long tiles = (R + kTileRows - 1) / kTileRows;
  PRD_REQUIRE(R < 0x7fffffffLL, "pair_transition: %lld pair rows exceed the TMA coordinate range", R);
  CUtensorMap map_rows;  // the pair tensor as [R rows][64 floats]: one tile = two boxes of [128 rows x 32 floats]
  TmaDims td;
  td.size[0] = 64; td.size[1] = (uint64_t)R;
  td.stride[0] = 256;
  td.box[0] = 32; td.box[1] = 128;
  if (make_tensor_map(&map_rows, pair, 4, 2, td, true)) return 1;
  PRD_CUDA_OK(launch_pdl(pair_transition_ws_kernel, grid_for(tiles, 1), kPtThreads, PtSmem::kTotal, s, map_rows, dst, residual, R, w1, b1,
                         w2, b2, w_bias, b_bias, bias_out, (unsigned)((long long)d.N * d.N)));
  PRD_LAUNCHED();
  return 0;
}

int pair_transition(const PairDims& d, const float* pair, float* dst, int residual, const __half* w1, const float* b1,
                    const __half* w2, const float* b2, int hidden, const float* w_bias, const float* b_bias, float* bias_out,
                    cudaStream_t s) {
  if (d.CZ == 64 && hidden == 256)
    return launch_pair_transition_ws(d, pair, dst, residual, w1, b1, w2, b2, w_bias, b_bias, bias_out, s);
  if (d.CZ == 32 && hidden == 128) {
    if (launch_pair_transition<32, 128>(d, pair, dst, residual, w1, b1, w2, b2, s)) return 1;
    // the narrow variant has no fused bias epilogue: separate stream kernel on the updated rows
    return bias_out ? pair_bias_proj(d, 4, dst, nullptr, nullptr, w_bias, b_bias, bias_out, s) : 0;
  }
  set_error("pair_transition: unsupported pair_dim %d / hidden %d (built: 64/256, 32/128)", d.CZ, hidden);
  return 1;
}

// =========================================================================================
// Triangle multiplication, input side: [a|b] = m2 * sigmoid(G p + bg) * (W p + bw), p = LN(pair),
// written as fp16 channel planes ab[which][b][d][i][k] (k contiguous, row stride plane_ld(N)) so
// that both einsum modes become plain K-major NT GEMMs  x_d[i,j] = sum_k a_d[i,k] b_d[j,k]:
//   outgoing: a_d[i,k] = a[b,i,k,d]      (logical row (b,i,k) reads pair[b,i,k])
//   incoming: a_d[i,k] = a[b,k,i,d]      (logical row (b,i,k) reads pair[b,k,i])
// w_in: fp16 pair [hi; lo], each [4CZ x CZ] with rows [0,2CZ) = ab_proj, [2CZ,4CZ) = ab_gate.
// 256 threads = two compute groups (see Group).
// =========================================================================================
template <int CZ>
__global__ void __launch_bounds__(256, 1)
trimul_in_kernel(const float* __restrict__ pair, const float* __restrict__ mask, RowMap map, int B, long long R,
                 const __half* __restrict__ w_in, const float* __restrict__ b_in, __half* __restrict__ ab, int Np) {
  extern __shared__ uint8_t raw[];
  constexpr int NOUT = 4 * CZ;
  constexpr int kGroupBytes = 16384 + (RowStage<CZ>::kBytes + 1023) / 1024 * 1024;
  uint8_t* sm = smem_align1024(raw);
  uint8_t* sW = sm;                  // hi
  uint8_t* sWl = sW + NOUT * 128;    // lo
  uint8_t* sG = sWl + NOUT * 128;    // per group: A tile, row stage
  float* sB = reinterpret_cast<float*>(sG + 2 * kGroupBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + NOUT);  // full[2], mma[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const Group g;
  uint8_t* sA = sG + g.grp * kGroupBytes;
  uint8_t* sSt = sA + 16384;
  uint64_t* full = bars + g.grp;
  uint64_t* mma_bar = bars + 2 + g.grp;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], kTileRows);
    mbar_init(&bars[1], kTileRows);
    mbar_init(&bars[2], 1);
    mbar_init(&bars[3], 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc(tmem_slot, 2 * NOUT);
  load_weight_kblocks(sW, w_in, NOUT, CZ, CZ, threadIdx.x, 256);
  load_weight_kblocks(sWl, w_in + NOUT * CZ, NOUT, CZ, CZ, threadIdx.x, 256);
  for (int i = threadIdx.x; i < NOUT; i += 256) sB[i] = b_in[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot + g.grp * NOUT;
  const uint32_t tm_lane = tmem + (static_cast<uint32_t>(g.warp * 32) << 16);
  const long long plane = (long long)map.N * Np;
  const int t = g.t;

  const long long num_tiles = (R + kTileRows - 1) / kTileRows;
  const long long stride = (long long)gridDim.x * 2;
  uint32_t mma_phase = 0;
  long long tile = (long long)blockIdx.x * 2 + g.grp;
  auto issue = [&](long long tl) {
    const long long r = tl * kTileRows + t;
    int b = 0, s = 0, k = 0;
    if (r < R) map.decompose(r, b, s, k);
    issue_row_load<CZ>(sSt, t, pair + map.src_row(b, s, k) * CZ, r < R, full);
  };
  if (tile < num_tiles) issue(tile);
  for (int it = 0; tile < num_tiles; tile += stride, ++it) {
    mbar_wait(full, it & 1);
    const long long r = tile * kTileRows + t;
    const bool valid = r < R;
    int b = 0, i = 0, k = 0;
    if (valid) map.decompose(r, b, i, k);
    {
      float x[CZ];
      if (valid) {
        read_row<CZ>(stage_row<CZ>(sSt, t), x);
      } else {
#pragma unroll
        for (int q = 0; q < CZ; ++q) x[q] = 0.f;
      }
      if (tile + stride < num_tiles) issue(tile + stride);  // this thread's stage slot is free again
      layernorm_inplace<CZ>(x);
      store_a_row<CZ>(sA, t, x);
    }
    g.sync_before_mma();
    if (t < 32) {  // warp-uniform issue: UMMA operands stay in uniform registers
      tc_fence_after();
      if (elect_one()) {
        umma_multi(tmem, smem_u32(sA), smem_u32(sW), 1, NOUT * 128, umma_idesc_f16(128, NOUT), false);
        umma_multi(tmem, smem_u32(sA), smem_u32(sWl), 1, NOUT * 128, umma_idesc_f16(128, NOUT), true);
        umma_commit(mma_bar);
      }
      __syncwarp();
    }
    // mask_2d for the *source* element: m[b,i]*m[b,k] is symmetric, same for both modes
    const float m2 = valid ? mask[(long long)b * map.N + i] * mask[(long long)b * map.N + k] : 0.f;
    __half* dst0 = ab + ((long long)b * CZ * map.N + i) * Np + k;
    mbar_wait(mma_bar, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    // channel cc in [0, 2CZ): a = [0,CZ) -> planes of tensor 0, b = [CZ,2CZ) -> planes of tensor 1.
    // One running pointer per 32-channel chunk, advanced by the plane stride: no 64-bit multiplies.
#pragma unroll 1
    for (int c = 0; c < 2 * CZ / 32; ++c) {
      uint32_t pr[32], ga[32];
      tmem_ld32(tm_lane + c * 32, pr);
      tmem_ld32(tm_lane + 2 * CZ + c * 32, ga);
      tmem_ld_wait();
      if (valid) {
        const int cc0 = c * 32;
        const long long pl0 = (cc0 < CZ) ? cc0 : (cc0 - CZ) + (long long)B * CZ;
        __half* dp = dst0 + pl0 * plane;
        const float* bp = sB + cc0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float gt = sigmoidf_fast(__uint_as_float(ga[j]) + bp[2 * CZ + j]);
          const float v = m2 * gt * (__uint_as_float(pr[j]) + bp[j]);
          *dp = __float2half_rn(v);
          dp += plane;
        }
      }
    }
    // the next tile's UMMA overwrites these TMEM columns only after the group barrier in sync_before_mma
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(*tmem_slot, 2 * NOUT);
}

// -----------------------------------------------------------------------------------------
// pair_dim 64: the projections are computed TRANSPOSED, D^T[channel][k] = W[channel][:] . X[k][:]^T (the weight
// tile is the UMMA A operand, the LayerNorm tile the B operand), so TMEM lane = output channel and a thread
// owns 64 consecutive k of one channel plane: 128 contiguous bytes instead of 128 two-byte stores with a
// plane stride.  A tile is 128 consecutive k of ONE (b, i) row; two threads per TMEM lane (Group2): half h
// takes columns [64 h, 64 h + 64).  512 threads = two compute groups.
// -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1)
trimul_in_t_kernel(const __grid_constant__ CUtensorMap map_pair, const float* __restrict__ mask, RowMap map, int B,
                   const __half* __restrict__ w_in, const float* __restrict__ b_in, __half* __restrict__ ab, int Np) {
  constexpr int CZ = 64, NOUT = 256;
  extern __shared__ uint8_t raw[];
  pdl_trigger();
  constexpr int kStage = (RowStage<CZ>::kBytes + 1023) / 1024 * 1024;
  constexpr int kGroupBytes = 32768 + kStage;  // X tile (store slices of warps 0-3 afterwards), slices of warps 4-7, row stage
  uint8_t* sm = smem_align1024(raw);
  uint8_t* sW = sm;                  // hi: proj rows [0,128), gate rows [128,256): two [128 x 64] A tiles
  uint8_t* sWl = sW + NOUT * 128;    // lo
  uint8_t* sG = sWl + NOUT * 128;
  float* sM2 = reinterpret_cast<float*>(sG + 2 * kGroupBytes);  // [2][128] mask product per tile row
  uint64_t* bars = reinterpret_cast<uint64_t*>(sM2 + 256);      // full[2], mma[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const Group2 g;
  uint8_t* sA = sG + g.grp * kGroupBytes;
  uint8_t* sSl = sA + (g.half ? 16384 : 0) + g.warp * 4096;  // this warp's store slice
  uint8_t* sSt = sA + 32768;
  float* sMask = sM2 + g.grp * 128;
  uint64_t* full = bars + g.grp;
  uint64_t* mma_bar = bars + 2 + g.grp;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    tma_prefetch_desc(&map_pair);
    mbar_init(&bars[2], 1);
    mbar_init(&bars[3], 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc(tmem_slot, 2 * NOUT);
  load_weight_kblocks(sW, w_in, NOUT, CZ, CZ, threadIdx.x, 512);
  load_weight_kblocks(sWl, w_in + NOUT * CZ, NOUT, CZ, CZ, threadIdx.x, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // everything above touched only weights / shared memory; the predecessor's output is read below
  const uint32_t tmem = *tmem_slot + g.grp * NOUT;
  const uint32_t tm_lane = tmem + (static_cast<uint32_t>(g.warp * 32) << 16);
  const int t = g.t, half = g.half, lane = t & 31;
  const int N = map.N;
  // this thread's output channel: lane t = channel of [a | b]; biases in the exp2 / plain domain
  const float bias_p = b_in[t];
  const float bias_g = -1.4426950408889634f * b_in[2 * CZ + t];
  const int tps = (N + kTileRows - 1) / kTileRows;  // k-tiles per (b, i) row
  const long long num_tiles = (long long)B * N * tps;
  const long long stride = (long long)gridDim.x * 2;
  const long long plane = (long long)N * Np;
  uint32_t mma_phase = 0;
  long long tile = (long long)blockIdx.x * 2 + g.grp;
  auto issue = [&](long long tl) {  // rows (b, i, k0 .. k0 + 127): one TMA tile (rows past N are zero-filled)
    if (g.tt != 0) return;
    const long long bi = tl / tps;
    const int kk0 = static_cast<int>(tl - bi * tps) * kTileRows;
    const int b = static_cast<int>(bi / N), i = static_cast<int>(bi - (long long)b * N);
    mbar_expect_tx(full, 32768);
    for (int h = 0; h < 2; ++h) {
      if (map.transposed) tma_load_4d(sSt + h * 16384, &map_pair, full, h * 32, i, kk0, b);
      else tma_load_4d(sSt + h * 16384, &map_pair, full, h * 32, kk0, i, b);
    }
  };
  if (tile < num_tiles) issue(tile);
  for (int it = 0; tile < num_tiles; tile += stride, ++it) {
    mbar_wait(full, it & 1);
    const long long bi = tile / tps;
    const int k0 = static_cast<int>(tile - bi * tps) * kTileRows;
    const int b = static_cast<int>(bi / N), i = static_cast<int>(bi - (long long)b * N);
    const bool valid = k0 + t < N;
    {
      // the warps of half 0 normalise the row and write all of it, the warps of half 1 fetch the mask product: with both
      // threads of a row normalising it (round 1) the kernel, which is issue bound, ran 11 % more instructions
      // (step 26.54 -> 26.13 ms together with triattn_proj)
      if (half == 0) {
        float x[CZ];
        read_row_tma64(sSt, t, x);
        layernorm_inplace<CZ>(x);
        store_a_row<CZ>(sA, t, x);
      } else {
        // mask_2d of the *source* element: m[b,i]*m[b,k] is symmetric, same for both modes
        sMask[t] = valid ? mask[(long long)b * N + i] * mask[(long long)b * N + k0 + t] : 0.f;
      }
    }
    g.sync_before_mma();
    if (tile + stride < num_tiles) issue(tile + stride);
    if (g.tt < 32) {  // warp-uniform issue: UMMA operands stay in uniform registers
      tc_fence_after();
      if (elect_one()) {
        const uint32_t idesc = umma_idesc_f16(128, 128);
        umma_kblock(tmem, smem_u32(sW), smem_u32(sA), idesc, false);                  // proj, hi
        umma_kblock(tmem, smem_u32(sWl), smem_u32(sA), idesc, true);                  // proj, lo
        umma_kblock(tmem + 128, smem_u32(sW) + 16384, smem_u32(sA), idesc, false);    // gate, hi
        umma_kblock(tmem + 128, smem_u32(sWl) + 16384, smem_u32(sA), idesc, true);    // gate, lo
        umma_commit(mma_bar);
      }
      __syncwarp();
    }
    mbar_wait(mma_bar, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    // ---- epilogue: channel t, columns k0 + 64 half + [0, 64)
    uint4 ov[8];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t pr[32], ga[32];
      tmem_ld32(tm_lane + half * 64 + c * 32, pr);
      tmem_ld32(tm_lane + 128 + half * 64 + c * 32, ga);
      tmem_ld_wait();
      const float4* mp = reinterpret_cast<const float4*>(sMask + half * 64 + c * 32);
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        const float4 m0 = mp[j >> 2], m1 = mp[(j >> 2) + 1];
        const float mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float z = fmaf(__uint_as_float(ga[j + e]), -1.4426950408889634f, bias_g);
          const float sg = __fdividef(mm[e], 1.0f + ex2_approx(z));  // m2 * sigmoid
          v[e] = sg * (__uint_as_float(pr[j + e]) + bias_p);
        }
        ov[c * 4 + (j >> 3)] = make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]),
                                          pack_half2(v[6], v[7]));
      }
    }
    tc_fence_before();
    if (k0 + 64 * half < Np) {  // (warp-uniform) the 64-column half lies inside the padded plane row
      // warp = 32 consecutive channels of tensor a (lanes 0-63) or b (lanes 64-127)
      const int ch0 = g.warp * 32;
      const long long pl0 = (ch0 < CZ) ? ch0 : (ch0 - CZ) + (long long)B * CZ;
      __half* gb = ab + ((long long)b * CZ + pl0) * plane + (long long)i * Np + k0 + 64 * half;
      warp_store_rows128(sSl, lane, ov, gb, plane * 2, 32);
    }
    g.bar();  // the X tile / store slices are rewritten by the next tile
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(*tmem_slot, 2 * NOUT);
}

template <int CZ>
static int launch_trimul_in(const PairDims& d, const float* pair, const float* mask, int mode, const __half* w_in,
                            const float* b_in, __half* ab, cudaStream_t s) {
  RowMap map{d.N, (long long)d.N * d.N, mode};
  if (CZ == 64) {
    const long long tiles = (long long)d.B * d.N * ((d.N + kTileRows - 1) / kTileRows);
    constexpr int kStage = (RowStage<64>::kBytes + 1023) / 1024 * 1024;
    constexpr int smem = 1024 + 2 * 256 * 128 + 2 * (32768 + kStage) + 256 * 4 + 64;
    if (set_smem(trimul_in_t_kernel, smem)) return 1;
    CUtensorMap mp;
    if (make_pair_tile_map(&mp, pair, d.B, d.N, mode)) return 1;
    PRD_CUDA_OK(launch_pdl(trimul_in_t_kernel, grid_for((tiles + 1) / 2, 1), 512, smem, s, mp, mask, map, d.B, w_in, b_in, ab, plane_ld(d.N)));
    PRD_LAUNCHED();
    return 0;
  }
  const long long R = (long long)d.B * d.N * d.N;
  const long long tiles = (R + kTileRows - 1) / kTileRows;
  constexpr int kGroupBytes = 16384 + (RowStage<CZ>::kBytes + 1023) / 1024 * 1024;
  constexpr int smem = 1024 + 2 * 4 * CZ * 128 + 2 * kGroupBytes + 4 * CZ * 4 + 64;
  auto kern = trimul_in_kernel<CZ>;
  if (set_smem(kern, smem)) return 1;
  kern<<<grid_for((tiles + 1) / 2, 1), 256, smem, s>>>(pair, mask, map, d.B, R, w_in, b_in, ab, plane_ld(d.N));
  PRD_LAUNCHED();
  return 0;
}

int trimul_in(const PairDims& d, const float* pair, const float* mask, int mode, const __half* w_in, const float* b_in,
              __half* ab, cudaStream_t s) {
  if (d.CZ == 64) return launch_trimul_in<64>(d, pair, mask, mode, w_in, b_in, ab, s);
  if (d.CZ == 32) return launch_trimul_in<32>(d, pair, mask, mode, w_in, b_in, ab, s);
  set_error("trimul_in: unsupported pair_dim %d", d.CZ);
  return 1;
}

// =========================================================================================
// Triangle multiplication, output side: dst = [pair +] sigmoid(Go p + bgo) * (Wo LN(x) + bo)
// x: fp16 planes [B][CZ][N][Nx] from the contraction GEMM (fp32 accumulate, one rounding).  w_out: fp16 pair, four [CZ x CZ] tiles
// in the order out_gate_hi, out_proj_hi, out_gate_lo, out_proj_lo.  Two compute groups; the pair row
// stays in registers for the residual, the output is staged over the (then idle) A tiles.
// =========================================================================================
template <int CZ>
__global__ void __launch_bounds__(256, 1)
trimul_out_kernel(const float* pair, float* dst, int residual, const __half* __restrict__ xpl,
                  const __grid_constant__ CUtensorMap map_x, int use_tma, int N, int Nx, long long R,
                  const __half* __restrict__ w_out, const float* __restrict__ b_out,
                  const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out, int row_tma) {
  // row_tma (pair_dim 64): the 128 pair rows of a tile arrive as ONE TMA tile (two swizzled boxes) and the output rows leave
  // as two tile stores per warp.  Per-thread 256-byte bulk copies serialise lane by lane on the uniform datapath
  // (ELECT / R2UR / UBLKCP loops: 26 % of this kernel's instructions in ncu, profiles/r02_launches.md).
  extern __shared__ uint8_t raw[];
  pdl_trigger();
  constexpr int kStage = (RowStage<CZ>::kBytes + 1023) / 1024 * 1024;
  constexpr int kAO = kStage > 32768 ? kStage : 32768;  // A_p | A_x, re-used as the output stage
  constexpr int kXBytes = CZ * kTileRows * 2;           // contraction-result tile [CZ planes][128 j] fp16
  constexpr int kGroupBytes = kAO + kStage + kXBytes;
  uint8_t* sm = smem_align1024(raw);
  uint8_t* sW = sm;  // four B tiles of [CZ x 64]
  uint8_t* sG = sW + 4 * CZ * 128;
  float* sB = reinterpret_cast<float*>(sG + 2 * kGroupBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 2 * CZ);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  constexpr int TCOLS = 2 * CZ;

  const Group g;
  uint8_t* sAp = sG + g.grp * kGroupBytes;
  uint8_t* sAx = sAp + 16384;
  uint8_t* sSt = sAp + kAO;
  const __half* sX = reinterpret_cast<const __half*>(sSt + kStage);
  uint64_t* full = bars + g.grp;
  uint64_t* mma_bar = bars + 2 + g.grp;
  uint64_t* xfull = bars + 4 + g.grp;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], row_tma ? 1 : kTileRows);
    mbar_init(&bars[1], row_tma ? 1 : kTileRows);
    mbar_init(&bars[2], 1);
    mbar_init(&bars[3], 1);
    mbar_init(&bars[4], 1);
    mbar_init(&bars[5], 1);
    if (use_tma) tma_prefetch_desc(&map_x);
    if (row_tma) {
      tma_prefetch_desc(&map_in);
      tma_prefetch_desc(&map_out);
    }
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc(tmem_slot, 2 * TCOLS);
  for (int q = 0; q < 4; ++q) load_weight_kblocks(sW + q * CZ * 128, w_out + q * CZ * CZ, CZ, CZ, CZ, threadIdx.x, 256);
  for (int i = threadIdx.x; i < 2 * CZ; i += 256) sB[i] = b_out[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // everything above touched only weights / shared memory; the predecessor's output is read below
  const uint32_t tmem = *tmem_slot + g.grp * TCOLS;
  const uint32_t tm_lane = tmem + (static_cast<uint32_t>(g.warp * 32) << 16);
  const long long NN = (long long)N * N;
  const long long xplane = (long long)N * Nx;
  const int t = g.t;

  const long long num_tiles = (R + kTileRows - 1) / kTileRows;
  const long long stride = (long long)gridDim.x * 2;
  uint32_t mma_phase = 0;
  long long tile = (long long)blockIdx.x * 2 + g.grp;
  // N % 128 == 0: a tile is 128 consecutive j of one (b, i) row, so its contraction results are ONE TMA box
  // [CZ planes][128 j] fetched a tile ahead by one thread (the per-thread 64-plane gather was this kernel's top stall)
  const int tiles_per_row = N / kTileRows;
  auto issue_x = [&](long long tl) {
    const long long bi = tl / tiles_per_row;
    const int j0 = static_cast<int>(tl - bi * tiles_per_row) * kTileRows;
    const int b = static_cast<int>(bi / N), i = static_cast<int>(bi - (long long)b * N);
    mbar_expect_tx(xfull, kXBytes);
    tma_load_4d(const_cast<__half*>(sX), &map_x, xfull, j0, i, 0, b);
  };
  auto issue_rows = [&](long long tl) {  // one thread: the tile's 128 rows (rows past R are zero-filled)
    mbar_expect_tx(full, 32768);
    const int r0 = static_cast<int>(tl * kTileRows);
    tma_load_2d(sSt, &map_in, full, 0, r0);
    tma_load_2d(sSt + 16384, &map_in, full, 32, r0);
  };
  if (tile < num_tiles) {
    const long long r = tile * kTileRows + t;
    if (row_tma) {
      if (t == 0) issue_rows(tile);
    } else {
      issue_row_load<CZ>(sSt, t, pair + r * CZ, r < R, full);
    }
    if (use_tma && t == 0) issue_x(tile);
  }
  for (int it = 0; tile < num_tiles; tile += stride, ++it) {
    const long long r = tile * kTileRows + t;
    const bool valid = r < R;
    // the previous tile's output rows were staged over the A tiles: its bulk stores must have read them
    bulk_wait_read0();
    g.bar();
    // contraction result for this (b,i,j): one value per channel plane.  TMA path: column t of the staged tile;
    // otherwise a gather, coalesced across lanes, issued first and consumed last.
    float x[CZ];
    if (use_tma) {
      mbar_wait(xfull, it & 1);
#pragma unroll
      for (int dch = 0; dch < CZ; ++dch) x[dch] = __half2float(sX[dch * kTileRows + t]);
    } else if (valid) {
      int b, rem;
      if (r < 0x7fffffffLL && NN < 0x7fffffffLL) {
        b = static_cast<int>(static_cast<unsigned>(r) / static_cast<unsigned>(NN));
        rem = static_cast<int>(static_cast<unsigned>(r) - static_cast<unsigned>(b) * static_cast<unsigned>(NN));
      } else {
        b = static_cast<int>(r / NN);
        rem = static_cast<int>(r - (long long)b * NN);
      }
      const int i = rem / N, j = rem - i * N;
      const __half* xp = xpl + (long long)b * CZ * xplane + (long long)i * Nx + j;
#pragma unroll
      for (int dch = 0; dch < CZ; ++dch) {
        x[dch] = __half2float(__ldg(xp));
        xp += xplane;
      }
    } else {
#pragma unroll
      for (int q = 0; q < CZ; ++q) x[q] = 0.f;
    }
    mbar_wait(full, it & 1);
    float xr[CZ];
    if (row_tma) {
      if constexpr (CZ == 64) read_row_tma64(sSt, t, xr);
    } else if (valid) {
      read_row<CZ>(stage_row<CZ>(sSt, t), xr);
    } else {
#pragma unroll
      for (int q = 0; q < CZ; ++q) xr[q] = 0.f;
    }
    if (!row_tma && tile + stride < num_tiles) {
      const long long rn = (tile + stride) * kTileRows + t;
      issue_row_load<CZ>(sSt, t, pair + rn * CZ, rn < R, full);
    }
    {
      float y[CZ];
#pragma unroll
      for (int q = 0; q < CZ; ++q) y[q] = xr[q];
      layernorm_inplace<CZ>(y);
      store_a_row<CZ>(sAp, t, y);
    }
    layernorm_inplace<CZ>(x);
    store_a_row<CZ>(sAx, t, x);
    g.sync_before_mma();
    if (use_tma && t == 64 && tile + stride < num_tiles) issue_x(tile + stride);  // every thread has read the x tile
    if (row_tma && t == 96 && tile + stride < num_tiles) issue_rows(tile + stride);  // ... and its pair row
    if (t < 32) {  // warp-uniform issue: UMMA operands stay in uniform registers
      tc_fence_after();
      if (elect_one()) {
        umma_multi(tmem, smem_u32(sAp), smem_u32(sW), 1, CZ * 128, umma_idesc_f16(128, CZ), false);
        umma_multi(tmem, smem_u32(sAp), smem_u32(sW + 2 * CZ * 128), 1, CZ * 128, umma_idesc_f16(128, CZ), true);
        umma_multi(tmem + CZ, smem_u32(sAx), smem_u32(sW + CZ * 128), 1, CZ * 128, umma_idesc_f16(128, CZ), false);
        umma_multi(tmem + CZ, smem_u32(sAx), smem_u32(sW + 3 * CZ * 128), 1, CZ * 128, umma_idesc_f16(128, CZ), true);
        umma_commit(mma_bar);
      }
      __syncwarp();
    }
    mbar_wait(mma_bar, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    float* my = stage_row<CZ>(sAp, t);  // output stage over the A tiles (their UMMAs are complete)
#pragma unroll
    for (int c = 0; c < CZ / 32; ++c) {
      uint32_t ga[32], pr[32];
      tmem_ld32(tm_lane + c * 32, ga);
      tmem_ld32(tm_lane + CZ + c * 32, pr);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int cc = c * 32 + j + e;
          o[e] = sigmoidf_fast(__uint_as_float(ga[j + e]) + sB[cc]) * (__uint_as_float(pr[j + e]) + sB[CZ + cc]);
          if (residual) o[e] += xr[cc];
        }
        if (row_tma)  // swizzled [2 boxes][128 rows][128 bytes] tile: box c, row t, 16-byte chunk (j / 4) ^ (t & 7)
          *reinterpret_cast<float4*>(sAp + c * 16384 + t * 128 + ((((j >> 2)) ^ (t & 7)) << 4)) = make_float4(o[0], o[1], o[2], o[3]);
        else
          *reinterpret_cast<float4*>(my + c * 32 + j) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
    fence_proxy_async_smem();
    if (row_tma) {
      __syncwarp();
      if ((t & 31) == 0 && tile * kTileRows + g.warp * 32 < R) {  // this warp's 32 rows x 2 boxes (clipped at R by the TMA unit)
        const int r0 = static_cast<int>(tile * kTileRows) + g.warp * 32;
        tma_store_2d(&map_out, sAp + g.warp * 4096, 0, r0);
        tma_store_2d(&map_out, sAp + 16384 + g.warp * 4096, 32, r0);
      }
    } else if (valid) {
      bulk_s2g(dst + r * CZ, my, CZ * 4);
    }
    bulk_commit();
    tc_fence_before();
  }
  bulk_wait0();
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(*tmem_slot, 2 * TCOLS);
}

template <int CZ>
static int launch_trimul_out(const PairDims& d, const float* pair, float* dst, int residual, const __half* x,
                             const __half* w_out, const float* b_out, cudaStream_t s) {
  const long long R = (long long)d.B * d.N * d.N;
  const long long tiles = (R + kTileRows - 1) / kTileRows;
  constexpr int kStage = (RowStage<CZ>::kBytes + 1023) / 1024 * 1024;
  constexpr int kAO = kStage > 32768 ? kStage : 32768;
  constexpr int smem = 1024 + 4 * CZ * 128 + 2 * (kAO + kStage + CZ * kTileRows * 2) + 2 * CZ * 4 + 64;
  auto kern = trimul_out_kernel<CZ>;
  if (set_smem(kern, smem)) return 1;
  // x planes [B][CZ][N][Nx] fp16 as a 4-D map (j, i, channel, b); box = [128 j][1][CZ][1]
  const int Nx = xplane_ld(d.N);
  const int use_tma = (d.N % kTileRows == 0) ? 1 : 0;
  CUtensorMap mx;
  {
    TmaDims t;
    t.size[0] = (uint64_t)d.N; t.size[1] = (uint64_t)d.N; t.size[2] = (uint64_t)CZ; t.size[3] = (uint64_t)d.B;
    t.stride[0] = (uint64_t)Nx * 2; t.stride[1] = (uint64_t)d.N * Nx * 2; t.stride[2] = (uint64_t)CZ * d.N * Nx * 2;
    t.box[0] = use_tma ? kTileRows : 8; t.box[1] = 1; t.box[2] = CZ; t.box[3] = 1;
    if (make_tensor_map(&mx, x, 2, 4, t, false)) return 1;
  }
  // pair rows as a 2-D tensor [R, CZ]: load box = [32 channels][128 rows], store box = [32 channels][32 rows] (one warp)
  CUtensorMap m_in = mx, m_out = mx;
  static const bool row_tma_off = getenv("PRD_ROW_TMA") && getenv("PRD_ROW_TMA")[0] == '0';  // A/B switch
  const int row_tma = (!row_tma_off && CZ == 64 && R < 0x7fffffffLL && (reinterpret_cast<uintptr_t>(pair) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(dst) & 15) == 0) ? 1 : 0;
  if (row_tma) {
    TmaDims t;
    t.size[0] = (uint64_t)CZ; t.size[1] = (uint64_t)R; t.size[2] = 1; t.size[3] = 1;
    t.stride[0] = (uint64_t)CZ * 4; t.stride[1] = (uint64_t)CZ * 4 * R; t.stride[2] = t.stride[1];
    t.box[0] = 32; t.box[1] = kTileRows; t.box[2] = 1; t.box[3] = 1;
    if (make_tensor_map(&m_in, pair, 4, 2, t, true)) return 1;
    t.box[1] = 32;
    if (make_tensor_map(&m_out, dst, 4, 2, t, true)) return 1;
  }
  PRD_CUDA_OK(launch_pdl(kern, grid_for((tiles + 1) / 2, 1), 256, smem, s, pair, dst, residual, x, mx, use_tma, d.N, Nx, R, w_out, b_out,
                         m_in, m_out, row_tma));
  PRD_LAUNCHED();
  return 0;
}

int trimul_out(const PairDims& d, const float* pair, float* dst, int residual, const __half* x, const __half* w_out,
               const float* b_out, cudaStream_t s) {
  if (d.CZ == 64) return launch_trimul_out<64>(d, pair, dst, residual, x, w_out, b_out, s);
  if (d.CZ == 32) return launch_trimul_out<32>(d, pair, dst, residual, x, w_out, b_out, s);
  set_error("trimul_out: unsupported pair_dim %d", d.CZ);
  return 1;
}

// =========================================================================================
// Triangle attention projections.  Logical row (b, seq, tok): "starting" reads pair[b,seq,tok],
// "ending" reads pair[b,tok,seq].  w: fp16 pair [hi; lo], each [256 x CZ] with rows [0,64) q,
// [64,128) k, [128,192) v, [192,256) gate.
// Outputs (fp16): q (pre-scaled by log2(e)/sqrt(c): the flash kernel works in the exp2 domain), k,
// g = sigmoid(gate) as [rows][64];
// v transposed per sequence: vt[(b*N+seq)][h*16+c][tok] (tok contiguous, row stride plane_ld(N)),
// i.e. the K-major B operand of the P.V product.
// A tile is 128 consecutive tokens of ONE sequence (ceil(N/128) tiles per sequence), so the v tile can be
// transposed through shared memory (the A tile, idle after the UMMA) and leave as 256-byte rows with
// 16-byte stores instead of 64 two-byte stores per thread.  Two compute groups.
// =========================================================================================
template <int CZ>
__global__ void __launch_bounds__(512, 1)
triattn_proj_kernel(const __grid_constant__ CUtensorMap map_pair, const float* __restrict__ pair, RowMap map, int B, const __half* __restrict__ w,
                    const float* __restrict__ b_gate, __half* __restrict__ q, __half* __restrict__ k,
                    __half* __restrict__ gout, __half* __restrict__ vt, int Np) {
  extern __shared__ uint8_t raw[];
  pdl_trigger();
  constexpr int NOUT = 256;
  constexpr bool kTma = (CZ == 64);  // pair_dim 64: the row tile arrives by TMA (32 KB swizzled stage)
  constexpr int kGroupBytes = 32768 + (RowStage<CZ>::kBytes + 1023) / 1024 * 1024;  // A tile, output stage, row stage
  uint8_t* sm = smem_align1024(raw);
  uint8_t* sW = sm;
  uint8_t* sWl = sW + NOUT * 128;
  uint8_t* sG = sWl + NOUT * 128;
  float* sB = reinterpret_cast<float*>(sG + 2 * kGroupBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 64);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const Group2 g;
  uint8_t* sA = sG + g.grp * kGroupBytes;
  uint8_t* sO = sA + 16384;  // gate rows (warp-private 4 KB slices), then the transposed v tile
  uint8_t* sSt = sA + 32768;
  uint64_t* full = bars + g.grp;
  uint64_t* mma_bar = bars + 2 + g.grp;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], kTma ? 1 : kTileRows);
    mbar_init(&bars[1], kTma ? 1 : kTileRows);
    if (kTma) tma_prefetch_desc(&map_pair);
    mbar_init(&bars[2], 1);
    mbar_init(&bars[3], 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc(tmem_slot, 2 * NOUT);
  load_weight_kblocks(sW, w, NOUT, CZ, CZ, threadIdx.x, 512);
  load_weight_kblocks(sWl, w + NOUT * CZ, NOUT, CZ, CZ, threadIdx.x, 512);
  if (threadIdx.x < 64) sB[threadIdx.x] = -1.4426950408889634f * b_gate[threadIdx.x];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // everything above touched only weights / shared memory; the predecessor's output is read below
  const uint32_t tmem = *tmem_slot + g.grp * NOUT;
  const uint32_t tm_lane = tmem + (static_cast<uint32_t>(g.warp * 32) << 16);
  const int t = g.t, half = g.half;
  const int N = map.N;
  const int tps = (N + kTileRows - 1) / kTileRows;  // tiles per sequence
  const long long num_tiles = (long long)B * N * tps;
  const long long stride = (long long)gridDim.x * 2;
  uint32_t mma_phase = 0;
  long long tile = (long long)blockIdx.x * 2 + g.grp;
  // tile -> (sequence = b * N + s, first token); the row of thread pair t is loaded by its half-0 thread
  auto issue = [&](long long tl) {
    const long long seq = tl / tps;
    const int tok0 = static_cast<int>(tl - seq * tps) * kTileRows;
    const int b = static_cast<int>(seq / N), s = static_cast<int>(seq - (long long)b * N);
    if (kTma) {
      if (g.tt == 0) {  // rows past N are zero-filled by the TMA unit
        mbar_expect_tx(full, 32768);
        for (int h = 0; h < 2; ++h) {
          if (map.transposed) tma_load_4d(sSt + h * 16384, &map_pair, full, h * 32, s, tok0, b);
          else tma_load_4d(sSt + h * 16384, &map_pair, full, h * 32, tok0, s, b);
        }
      }
    } else if (half == 0) {
      const int tok = tok0 + t;
      issue_row_load<CZ>(sSt, t, pair + map.src_row(b, s, tok < N ? tok : 0) * CZ, tok < N, full);
    }
  };
  if (tile < num_tiles) issue(tile);
  for (int it = 0; tile < num_tiles; tile += stride, ++it) {
    mbar_wait(full, it & 1);
    const long long seq = tile / tps;
    const int tok0 = static_cast<int>(tile - seq * tps) * kTileRows;
    const bool valid = tok0 + t < N;
    const long long r0 = seq * N + tok0;  // first row of the tile
    if (half == 0) {  // warp-uniform: the warps of half 0 normalise the row and write all of it (see trimul_in_t_kernel)
      float x[CZ];
      if (kTma) {
        if constexpr (CZ == 64) read_row_tma64(sSt, t, x);
      } else if (valid) {
        read_row<CZ>(stage_row<CZ>(sSt, t), x);
      } else {
#pragma unroll
        for (int i = 0; i < CZ; ++i) x[i] = 0.f;
      }
      layernorm_inplace<CZ>(x);
      store_a_row<CZ>(sA, t, x);
    }
    g.sync_before_mma();  // also: both threads have read the row, its stage slot may be refilled
    if (tile + stride < num_tiles) issue(tile + stride);
    if (g.tt < 32) {  // warp-uniform issue: UMMA operands stay in uniform registers
      tc_fence_after();
      if (elect_one()) {
        umma_multi(tmem, smem_u32(sA), smem_u32(sW), 1, NOUT * 128, umma_idesc_f16(128, NOUT), false);
        umma_multi(tmem, smem_u32(sA), smem_u32(sWl), 1, NOUT * 128, umma_idesc_f16(128, NOUT), true);
        umma_commit(mma_bar);
      }
      __syncwarp();
    }
    mbar_wait(mma_bar, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    if (half == 0) {
      // ---- q (scaled), k: row-major [rows][64]
#pragma unroll
      for (int part = 0; part < 2; ++part) {
        uint32_t acc[2][32];
        tmem_ld32(tm_lane + part * 64, acc[0]);
        tmem_ld32(tm_lane + part * 64 + 32, acc[1]);
        tmem_ld_wait();
        const float sc = part == 0 ? 0.25f * 1.4426950408889634f : 1.0f;
        uint4 o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t* a = &acc[j >> 2][(j & 3) * 8];
          o[j].x = pack_half2(__uint_as_float(a[0]) * sc, __uint_as_float(a[1]) * sc);
          o[j].y = pack_half2(__uint_as_float(a[2]) * sc, __uint_as_float(a[3]) * sc);
          o[j].z = pack_half2(__uint_as_float(a[4]) * sc, __uint_as_float(a[5]) * sc);
          o[j].w = pack_half2(__uint_as_float(a[6]) * sc, __uint_as_float(a[7]) * sc);
        }
        // the A tile is idle after the UMMA: warp-private 4 KB slices
        warp_store_rows128(sA + g.warp * 4096, t & 31, o, (part == 0 ? q : k) + (r0 + g.warp * 32) * 64, 128,
                           N - (tok0 + g.warp * 32));
      }
    } else {
      // ---- gate: sigmoid(a + b) = 1 / (1 + exp2(-(a + b) log2 e)), row-major
      {
        uint32_t acc[2][32];
        tmem_ld32(tm_lane + 192, acc[0]);
        tmem_ld32(tm_lane + 224, acc[1]);
        tmem_ld_wait();
        uint4 o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float gsig[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int col = j * 8 + e;
            const float z = fmaf(__uint_as_float(acc[col >> 5][col & 31]), -1.4426950408889634f, sB[col]);
            gsig[e] = __fdividef(1.0f, 1.0f + ex2_approx(z));
          }
          o[j].x = pack_half2(gsig[0], gsig[1]);
          o[j].y = pack_half2(gsig[2], gsig[3]);
          o[j].z = pack_half2(gsig[4], gsig[5]);
          o[j].w = pack_half2(gsig[6], gsig[7]);
        }
        warp_store_rows128(sO + g.warp * 4096, t & 31, o, gout + (r0 + g.warp * 32) * 64, 128, N - (tok0 + g.warp * 32));
      }
      // ---- v: transpose through shared memory: sO[ch][tok] halves, 256 B per channel (the four gate warps
      // are done with their slices of sO)
      asm volatile("bar.sync %0, 128;" ::"r"(g.grp + 3) : "memory");
      {
        uint32_t acc[2][32];
        tmem_ld32(tm_lane + 128, acc[0]);
        tmem_ld32(tm_lane + 160, acc[1]);
        tmem_ld_wait();
        __half* sV = reinterpret_cast<__half*>(sO);
#pragma unroll
        for (int c = 0; c < 64; ++c) sV[c * kTileRows + t] = __float2half_rn(__uint_as_float(acc[c >> 5][c & 31]));
      }
    }
    tc_fence_before();
    g.bar();
    {
      // thread -> 16-byte vector (tt & 15) of channel rows (tt >> 4) + 16 i: a warp writes two full 256-byte rows
      const int vec = g.tt & 15;
      __half* vrow = vt + seq * 64 * Np + tok0 + vec * 8;
      const int n_ok = N - (tok0 + vec * 8);  // tokens of this vector that exist
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ch = (g.tt >> 4) + 16 * i;
        const uint4 v = *reinterpret_cast<const uint4*>(sO + ch * 256 + vec * 16);
        __half* dp = vrow + (long long)ch * Np;
        if (n_ok >= 8) {
          *reinterpret_cast<uint4*>(dp) = v;
        } else if (n_ok > 0) {  // ragged tail: static register indices (a dynamic one would spill v to local memory)
          const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (e < n_ok) reinterpret_cast<unsigned short*>(dp)[e] = static_cast<unsigned short>(w4[e >> 1] >> ((e & 1) * 16));
        }
      }
    }
    g.bar();  // the A tile and the output stage are rewritten by the next tile
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(*tmem_slot, 2 * NOUT);
}

template <int CZ>
static int launch_triattn_proj(const PairDims& d, const float* pair, int mode, const __half* w_qkvg, const float* b_gate,
                               __half* q, __half* k, __half* g, __half* vt, cudaStream_t s) {
  const long long tiles = (long long)d.B * d.N * ((d.N + kTileRows - 1) / kTileRows);
  RowMap map{d.N, (long long)d.N * d.N, mode};
  constexpr bool kTma = (CZ == 64);  // pair_dim 64: the row tile arrives by TMA (32 KB swizzled stage)
  constexpr int kGroupBytes = 32768 + (RowStage<CZ>::kBytes + 1023) / 1024 * 1024;  // A tile, output stage, row stage
  constexpr int smem = 1024 + 2 * 256 * 128 + 2 * kGroupBytes + 64 * 4 + 64;
  auto kern = triattn_proj_kernel<CZ>;
  if (set_smem(kern, smem)) return 1;
  CUtensorMap mp;
  if (CZ == 64) {
    if (make_pair_tile_map(&mp, pair, d.B, d.N, mode)) return 1;
  } else {
    memset(&mp, 0, sizeof(mp));
  }
  PRD_CUDA_OK(launch_pdl(kern, grid_for((tiles + 1) / 2, 1), 512, smem, s, mp, pair, map, d.B, w_qkvg, b_gate, q, k, g, vt, plane_ld(d.N)));
  PRD_LAUNCHED();
  return 0;
}

int triattn_proj(const PairDims& d, const float* pair, int mode, const __half* w_qkvg, const float* b_gate, __half* q,
                 __half* k, __half* g, __half* vt, cudaStream_t s) {
  if (d.CZ == 64) return launch_triattn_proj<64>(d, pair, mode, w_qkvg, b_gate, q, k, g, vt, s);
  if (d.CZ == 32) return launch_triattn_proj<32>(d, pair, mode, w_qkvg, b_gate, q, k, g, vt, s);
  set_error("triattn_proj: unsupported pair_dim %d", d.CZ);
  return 1;
}

// =========================================================================================
// Triangle attention output projection + residual: dst[src(b,seq,tok)] = [pair +] Wo og[(b,seq,tok)] + bo
// og: [rows][64] fp16 gated attention output (logical row order).  w_o: fp16 pair [hi; lo], [CZ x 64]
// each.  Two compute groups, one stage per group used for both the residual row and the output.
// =========================================================================================
template <int CZ>
__global__ void __launch_bounds__(256, 1)
triattn_out_kernel(const __grid_constant__ CUtensorMap map_og, const float* pair, float* dst, int residual, RowMap map,
                   long long R, const __half* __restrict__ w_o, const float* __restrict__ b_o,
                   const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out, int row_tma) {
  // row_tma (pair_dim 64; "ending" needs N % 128 == 0 so that a tile is one sequence): the residual rows of a tile arrive as
  // ONE TMA tile (two swizzled boxes) and the output rows leave as two tile stores per warp -- the per-thread 256-byte bulk
  // copies serialise lane by lane on the uniform datapath and were 58 % of this kernel's instructions (ncu).
  // Both inputs of a tile are in flight a tile ahead: the og tile [128 rows x 64 halves] is one swizzled TMA box that IS the
  // UMMA A operand (issued as soon as the previous tile's UMMAs have completed), the residual rows go to the other of two row
  // stages (the stage a tile was read from also carries its output rows to the bulk store).
  extern __shared__ uint8_t raw[];
  pdl_trigger();
  constexpr int kStage = (RowStage<CZ>::kBytes + 1023) / 1024 * 1024;
  constexpr int kGroupBytes = 16384 + 2 * kStage;
  uint8_t* sm = smem_align1024(raw);
  uint8_t* sW = sm;
  uint8_t* sWl = sW + CZ * 128;
  uint8_t* sG = sm + ((2 * CZ * 128 + 1023) / 1024) * 1024;
  float* sB = reinterpret_cast<float*>(sG + 2 * kGroupBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + CZ);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  constexpr int TCOLS = CZ < 32 ? 32 : CZ;

  const Group g;
  uint8_t* sA = sG + g.grp * kGroupBytes;
  uint8_t* sSt = sA + 16384;  // two stages, kStage apart
  uint64_t* full = bars + g.grp * 2;  // [2]
  uint64_t* mma_bar = bars + 4 + g.grp;
  uint64_t* afull = bars + 6 + g.grp;
  if (threadIdx.x == 0) {
    for (int q = 0; q < 4; ++q) mbar_init(&bars[q], row_tma ? 1 : kTileRows);
    for (int q = 4; q < 8; ++q) mbar_init(&bars[q], 1);
    tma_prefetch_desc(&map_og);
    if (row_tma) {
      tma_prefetch_desc(&map_in);
      tma_prefetch_desc(&map_out);
    }
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc(tmem_slot, 2 * TCOLS);
  load_weight_kblocks(sW, w_o, CZ, 64, 64, threadIdx.x, 256);
  load_weight_kblocks(sWl, w_o + CZ * 64, CZ, 64, 64, threadIdx.x, 256);
  for (int i = threadIdx.x; i < CZ; i += 256) sB[i] = b_o[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // everything above touched only weights / shared memory; the predecessor's output is read below
  const uint32_t tmem = *tmem_slot + g.grp * TCOLS;
  const uint32_t tm_lane = tmem + (static_cast<uint32_t>(g.warp * 32) << 16);
  const int t = g.t;

  const long long num_tiles = (R + kTileRows - 1) / kTileRows;
  const long long stride = (long long)gridDim.x * 2;
  uint32_t mma_phase = 0;
  long long tile = (long long)blockIdx.x * 2 + g.grp;
  auto src_of = [&](long long tl, bool& valid) {
    const long long r = tl * kTileRows + t;
    valid = r < R;
    long long src = 0;
    if (valid) {
      int b, s, tk;
      map.decompose(r, b, s, tk);
      src = map.src_row(b, s, tk);
    }
    return src;
  };
  auto issue_og = [&](long long tl) {  // one thread; rows past R are zero-filled by the TMA unit
    mbar_expect_tx(afull, 16384);
    tma_load_3d(sA, &map_og, afull, 0, static_cast<int>(tl * kTileRows), 0);
  };
  // coordinates of the tile's first row in map_in / map_out: (row, 0, 0) for "starting" (the rows are consecutive in memory),
  // (seq, tok, b) for "ending"
  auto tile_coords = [&](long long tl, int row_off, int& c1, int& c2, int& c3) {
    const long long r0 = tl * kTileRows + row_off;
    if (map.transposed) {
      int b, sq, tk;
      map.decompose(r0, b, sq, tk);
      c1 = sq; c2 = tk; c3 = b;
    } else {
      c1 = static_cast<int>(r0); c2 = 0; c3 = 0;
    }
  };
  auto issue_rows = [&](long long tl, uint8_t* stage, uint64_t* bar) {  // one thread; rows past R are zero-filled
    int c1, c2, c3;
    tile_coords(tl, 0, c1, c2, c3);
    mbar_expect_tx(bar, 32768);
    tma_load_4d(stage, &map_in, bar, 0, c1, c2, c3);
    tma_load_4d(stage + 16384, &map_in, bar, 32, c1, c2, c3);
  };
  bool valid = false;
  long long src = 0;
  if (tile < num_tiles) {
    if (row_tma) {
      if (t == 32 && residual) issue_rows(tile, sSt, full);
    } else {
      src = src_of(tile, valid);
      issue_row_load<CZ>(sSt, t, pair + src * CZ, valid && residual, full);
    }
    if (t == 0) issue_og(tile);
  }
  for (int it = 0; tile < num_tiles; tile += stride, ++it) {
    const int cur = it & 1;
    const bool has_next = tile + stride < num_tiles;
    bool valid_n = false;
    long long src_n = 0;
    // the other stage carried the previous tile's output row of this thread: its bulk store must have read it
    bulk_wait_read0();
    if (has_next && !row_tma) {
      src_n = src_of(tile + stride, valid_n);
      issue_row_load<CZ>(sSt + (cur ^ 1) * kStage, t, pair + src_n * CZ, valid_n && residual, full + (cur ^ 1));
    }
    // every thread of the group has finished the previous tile's TMEM reads
    tc_fence_before();
    g.bar();
    // (row_tma) ... and every warp's tile stores out of the other stage have been read: it can take the next tile's rows
    if (row_tma && has_next && residual && t == 32) issue_rows(tile + stride, sSt + (cur ^ 1) * kStage, full + (cur ^ 1));
    if (t < 32) {  // warp-uniform issue: UMMA operands stay in uniform registers
      mbar_wait(afull, it & 1);
      tc_fence_after();
      if (elect_one()) {
        umma_multi(tmem, smem_u32(sA), smem_u32(sW), 1, CZ * 128, umma_idesc_f16(128, CZ), false);
        umma_multi(tmem, smem_u32(sA), smem_u32(sWl), 1, CZ * 128, umma_idesc_f16(128, CZ), true);
        umma_commit(mma_bar);
      }
      __syncwarp();
    }
    if (!row_tma || residual) mbar_wait(full + cur, (it >> 1) & 1);
    mbar_wait(mma_bar, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    if (t == 64 && has_next) issue_og(tile + stride);  // the A tile is free: this tile's UMMAs have completed
    float* my = stage_row<CZ>(sSt + cur * kStage, t);
    uint8_t* tstage = sSt + cur * kStage;  // row_tma: swizzled [2 boxes][128 rows][128 bytes]
#pragma unroll
    for (int c = 0; c < CZ / 32; ++c) {
      uint32_t acc[32];
      tmem_ld32(tm_lane + c * 32, acc);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4* px = row_tma ? reinterpret_cast<float4*>(tstage + c * 16384 + t * 128 + (((j >> 2) ^ (t & 7)) << 4))
                             : reinterpret_cast<float4*>(my + c * 32 + j);
        float4 x = (residual && (row_tma || valid)) ? *px : make_float4(0.f, 0.f, 0.f, 0.f);
        x.x += __uint_as_float(acc[j + 0]) + sB[c * 32 + j + 0];
        x.y += __uint_as_float(acc[j + 1]) + sB[c * 32 + j + 1];
        x.z += __uint_as_float(acc[j + 2]) + sB[c * 32 + j + 2];
        x.w += __uint_as_float(acc[j + 3]) + sB[c * 32 + j + 3];
        *px = x;
      }
    }
    fence_proxy_async_smem();
    if (row_tma) {
      __syncwarp();
      if ((t & 31) == 0 && tile * kTileRows + g.warp * 32 < R) {  // this warp's 32 rows x 2 boxes, clipped at R by the TMA unit
        int c1, c2, c3;
        tile_coords(tile, g.warp * 32, c1, c2, c3);
        tma_store_4d(&map_out, tstage + g.warp * 4096, 0, c1, c2, c3);
        tma_store_4d(&map_out, tstage + 16384 + g.warp * 4096, 32, c1, c2, c3);
      }
    } else if (valid) {
      bulk_s2g(dst + src * CZ, my, CZ * 4);
    }
    bulk_commit();
    valid = valid_n;
    src = src_n;
  }
  bulk_wait0();
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(*tmem_slot, 2 * TCOLS);
}

template <int CZ>
static int launch_triattn_out(const PairDims& d, const float* pair, float* dst, int residual, int mode, const __half* og,
                              const __half* w_o, const float* b_o, cudaStream_t s) {
  const long long R = (long long)d.B * d.N * d.N;
  const long long tiles = (R + kTileRows - 1) / kTileRows;
  RowMap map{d.N, (long long)d.N * d.N, mode};
  constexpr int kStage = (RowStage<CZ>::kBytes + 1023) / 1024 * 1024;
  constexpr int kGroupBytes = 16384 + 2 * kStage;
  constexpr int smem = 1024 + ((2 * CZ * 128 + 1023) / 1024) * 1024 + 2 * kGroupBytes + CZ * 4 + 128;
  auto kern = triattn_out_kernel<CZ>;
  if (set_smem(kern, smem)) return 1;
  // og [R rows][64 halves] as (channel, row, 1); box = [64][128 rows] with the 128-byte swizzle = one UMMA A K-block
  CUtensorMap mo;
  {
    TmaDims t;
    t.size[0] = 64; t.size[1] = (uint64_t)R; t.size[2] = 1; t.size[3] = 1;
    t.stride[0] = 128; t.stride[1] = (uint64_t)R * 128; t.stride[2] = 0;
    t.box[0] = 64; t.box[1] = 128; t.box[2] = 1; t.box[3] = 1;
    if (make_tensor_map(&mo, og, 2, 3, t, true)) return 1;
  }
  // pair rows: "starting" = the 2-D tensor [R, CZ] (consecutive rows); "ending" = (channel, seq, tok, b) of [B][tok][seq][CZ]
  // with the tile along tok.  Load box = 128 rows, store box = 32 rows (one warp).
  static const bool row_tma_off = getenv("PRD_ROW_TMA") && getenv("PRD_ROW_TMA")[0] == '0';  // A/B switch
  const int row_tma = (!row_tma_off && CZ == 64 && R < 0x7fffffffLL && (mode == 0 || d.N % kTileRows == 0) &&
                       (reinterpret_cast<uintptr_t>(pair) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) ? 1 : 0;
  CUtensorMap m_in = mo, m_out = mo;
  if (row_tma) {
    TmaDims t;
    for (int pass = 0; pass < 2; ++pass) {
      const uint32_t rows = pass == 0 ? kTileRows : 32;
      if (mode == 0) {
        t.size[0] = (uint64_t)CZ; t.size[1] = (uint64_t)R; t.size[2] = 1; t.size[3] = 1;
        t.stride[0] = (uint64_t)CZ * 4; t.stride[1] = (uint64_t)CZ * 4 * R; t.stride[2] = t.stride[1];
        t.box[0] = 32; t.box[1] = rows; t.box[2] = 1; t.box[3] = 1;
      } else {
        t.size[0] = (uint64_t)CZ; t.size[1] = (uint64_t)d.N; t.size[2] = (uint64_t)d.N; t.size[3] = (uint64_t)d.B;
        t.stride[0] = (uint64_t)CZ * 4; t.stride[1] = (uint64_t)d.N * CZ * 4; t.stride[2] = (uint64_t)d.N * d.N * CZ * 4;
        t.box[0] = 32; t.box[1] = 1; t.box[2] = rows; t.box[3] = 1;
      }
      if (make_tensor_map(pass == 0 ? &m_in : &m_out, pass == 0 ? pair : dst, 4, 4, t, true)) return 1;
    }
  }
  PRD_CUDA_OK(launch_pdl(kern, grid_for((tiles + 1) / 2, 1), 256, smem, s, mo, pair, dst, residual, map, R, w_o, b_o, m_in, m_out,
                         row_tma));
  PRD_LAUNCHED();
  return 0;
}

int triattn_out(const PairDims& d, const float* pair, float* dst, int residual, int mode, const __half* og,
                const __half* w_o, const float* b_o, cudaStream_t s) {
  if (d.CZ == 64) return launch_triattn_out<64>(d, pair, dst, residual, mode, og, w_o, b_o, s);
  if (d.CZ == 32) return launch_triattn_out<32>(d, pair, dst, residual, mode, og, w_o, b_o, s);
  set_error("triattn_out: unsupported pair_dim %d", d.CZ);
  return 1;
}

}  // namespace prd
