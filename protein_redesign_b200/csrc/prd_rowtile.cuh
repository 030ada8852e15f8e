// Shared device helpers for the "pair-row tile" kernels.
//
// Execution model (all kernels in prd_rowtile.cu / prd_pairembed.cu / prd_coord.cu):
//   * a tile = 128 pair elements ("rows"), one thread per row, thread t <-> TMEM lane t;
//   * each thread bulk-copies its own row (C_Z fp32) global -> shared into a padded staging
//     slot (row stride C_Z*4+16 bytes, so same-column reads by the 32 lanes hit 8 distinct
//     16-byte bank groups: conflict free with static register indices);
//   * LayerNorm runs thread-locally in registers, the fp16 result is written as a K-major
//     SWIZZLE_128B UMMA A-operand tile; weights sit in shared memory as B operands;
//   * thread 0 issues tcgen05.mma, all threads read their accumulator row back from TMEM.
#pragma once
#include "prd_common.cuh"

namespace prd {

constexpr int kTileRows = 128;
constexpr float kLnEps = 1e-5f;

template <int CZ>
struct RowStage {
  static constexpr int kRowBytes = CZ * 4 + 16;
  static constexpr int kBytes = kTileRows * kRowBytes;
};

// Row r of a staging buffer.
template <int CZ>
__device__ __forceinline__ float* stage_row(uint8_t* stage, int t) {
  return reinterpret_cast<float*>(stage + t * RowStage<CZ>::kRowBytes);
}

// Thread-private async row load: arrive on `bar` (count = 128) expecting CZ*4 bytes, or a plain
// arrive for rows past the end of the problem.
template <int CZ>
__device__ __forceinline__ void issue_row_load(uint8_t* stage, int t, const float* gsrc, bool valid, uint64_t* bar) {
  if (valid) {
    // the slot was read (ld.shared, generic proxy) by this thread: order those reads before the async-proxy refill
    fence_proxy_async_smem();
    mbar_expect_tx(bar, CZ * 4);
    bulk_g2s(stage_row<CZ>(stage, t), gsrc, CZ * 4, bar);
  } else {
    mbar_arrive(bar);
  }
}

template <int CZ>
__device__ __forceinline__ void read_row(const float* row, float (&x)[CZ]) {
#pragma unroll
  for (int c = 0; c < CZ / 4; ++c) {
    const float4 v = *reinterpret_cast<const float4*>(row + c * 4);
    x[c * 4 + 0] = v.x;
    x[c * 4 + 1] = v.y;
    x[c * 4 + 2] = v.z;
    x[c * 4 + 3] = v.w;
  }
}

template <int CZ>
__device__ __forceinline__ void write_row(float* row, const float (&x)[CZ]) {
#pragma unroll
  for (int c = 0; c < CZ / 4; ++c)
    *reinterpret_cast<float4*>(row + c * 4) = make_float4(x[c * 4], x[c * 4 + 1], x[c * 4 + 2], x[c * 4 + 3]);
}

// Pair-row tile staged by TMA: 128 rows x 64 fp32 as two SWIZZLE_128B boxes of [128 rows x 32 floats] (16 KB each).
// Row t, 16-byte chunk cc (0..15) lives at  (cc >> 3) * 16384 + t * 128 + (((cc & 7) ^ (t & 7)) << 4): a thread reading
// its own row is bank-conflict free, and ONE thread issues the whole tile (two cp.async.bulk.tensor) instead of 128
// per-thread row copies (which serialise lane by lane on the uniform datapath, ~8 % of a row kernel's time).
__device__ __forceinline__ void read_row_tma64(const uint8_t* stage, int t, float (&x)[64]) {
#pragma unroll
  for (int cc = 0; cc < 16; ++cc) {
    const float4 v = *reinterpret_cast<const float4*>(stage + (cc >> 3) * 16384 + t * 128 + (((cc & 7) ^ (t & 7)) << 4));
    x[cc * 4 + 0] = v.x;
    x[cc * 4 + 1] = v.y;
    x[cc * 4 + 2] = v.z;
    x[cc * 4 + 3] = v.w;
  }
}

// In-place LayerNorm without affine (nn.LayerNorm(elementwise_affine=False), eps 1e-5).
template <int CZ>
__device__ __forceinline__ void layernorm_inplace(float (&x)[CZ]) {
  // four independent partial sums: a single 64-deep dependent FADD chain costs ~250 cycles of pure latency
  float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < CZ; ++i) s4[i & 3] += x[i];
  const float mean = ((s4[0] + s4[1]) + (s4[2] + s4[3])) * (1.0f / CZ);
  float v4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < CZ; ++i) {
    x[i] -= mean;
    v4[i & 3] = fmaf(x[i], x[i], v4[i & 3]);
  }
  const float rstd = rsqrtf(((v4[0] + v4[1]) + (v4[2] + v4[3])) * (1.0f / CZ) + kLnEps);
#pragma unroll
  for (int i = 0; i < CZ; ++i) x[i] *= rstd;
}

// Write one row (CZ <= 64 values, zero padded to 64) of a [128 x 64] fp16 K-block.
template <int CZ>
__device__ __forceinline__ void store_a_row(uint8_t* a_tile, int t, const float (&y)[CZ]) {
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    uint4 o = make_uint4(0, 0, 0, 0);
    if (ch * 8 < CZ) {
      o.x = pack_half2(y[ch * 8 + 0], y[ch * 8 + 1]);
      o.y = pack_half2(y[ch * 8 + 2], y[ch * 8 + 3]);
      o.z = pack_half2(y[ch * 8 + 4], y[ch * 8 + 5]);
      o.w = pack_half2(y[ch * 8 + 6], y[ch * 8 + 7]);
    }
    *reinterpret_cast<uint4*>(a_tile + sw128_offset(t, ch)) = o;
  }
}

// Write 32 consecutive K-values (columns k0 .. k0+31, k0 a multiple of 32) of row t into a
// multi-K-block A operand whose K-blocks ([128 x 64] halves, 16 KB) are stored back to back.
__device__ __forceinline__ void store_a_cols32(uint8_t* a_tiles, int t, int k0, const float (&v)[32]) {
  uint8_t* blk = a_tiles + (k0 >> 6) * (kTileRows * 128);
  const int ch0 = (k0 & 63) >> 3;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 o;
    o.x = pack_half2(v[c * 8 + 0], v[c * 8 + 1]);
    o.y = pack_half2(v[c * 8 + 2], v[c * 8 + 3]);
    o.z = pack_half2(v[c * 8 + 4], v[c * 8 + 5]);
    o.w = pack_half2(v[c * 8 + 6], v[c * 8 + 7]);
    *reinterpret_cast<uint4*>(blk + sw128_offset(t, ch0 + c)) = o;
  }
}

// Same for 16 consecutive K-values (k0 a multiple of 16).
__device__ __forceinline__ void store_a_cols16(uint8_t* a_tiles, int t, int k0, const float (&v)[16]) {
  uint8_t* blk = a_tiles + (k0 >> 6) * (kTileRows * 128);
  const int ch0 = (k0 & 63) >> 3;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint4 o;
    o.x = pack_half2(v[c * 8 + 0], v[c * 8 + 1]);
    o.y = pack_half2(v[c * 8 + 2], v[c * 8 + 3]);
    o.z = pack_half2(v[c * 8 + 4], v[c * 8 + 5]);
    o.w = pack_half2(v[c * 8 + 6], v[c * 8 + 7]);
    *reinterpret_cast<uint4*>(blk + sw128_offset(t, ch0 + c)) = o;
  }
}

// Cooperative copy of a row-major fp16 weight W[rows][K] (leading dim ld halves) into K-major
// SWIZZLE_128B B-operand K-blocks: block kb holds columns [64 kb, 64 kb + 64) of all rows,
// zero padded past K.  Block size = rows * 128 bytes (rows must be a multiple of 8).
__device__ __forceinline__ void load_weight_kblocks(uint8_t* dst, const __half* W, int rows, int K, int ld, int tid,
                                                    int nthreads) {
  const int kblocks = (K + 63) >> 6;
  const int total = kblocks * rows * 8;
  for (int idx = tid; idx < total; idx += nthreads) {
    const int ch = idx & 7;
    const int r = (idx >> 3) % rows;
    const int kb = (idx >> 3) / rows;
    const int col = kb * 64 + ch * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (col + 8 <= K) {
      v = *reinterpret_cast<const uint4*>(W + (long long)r * ld + col);
    } else if (col < K) {
      __half tmp[8];
      for (int e = 0; e < 8; ++e) tmp[e] = (col + e < K) ? W[(long long)r * ld + col + e] : __float2half(0.f);
      v = *reinterpret_cast<uint4*>(tmp);
    }
    *reinterpret_cast<uint4*>(dst + kb * rows * 128 + sw128_offset(r, ch)) = v;
  }
}

// Issue UMMAs over `kblocks` K-blocks: A K-blocks are [128 x 64] (16 KB apart), B K-blocks are
// [n_rows x 64] (n_rows*128 bytes apart).  Called by one thread.
__device__ __forceinline__ void umma_multi(uint32_t tmem_d, uint32_t a_smem, uint32_t b_smem, int kblocks,
                                           uint32_t b_block_bytes, uint32_t idesc, bool accumulate) {
  for (int kb = 0; kb < kblocks; ++kb)
    umma_kblock(tmem_d, a_smem + kb * (kTileRows * 128), b_smem + kb * b_block_bytes, idesc, accumulate || kb > 0);
}

// smem carve-up helper: returns a 1024-byte aligned base inside the dynamic smem window.
__device__ __forceinline__ uint8_t* smem_align1024(uint8_t* raw) {
  const uint32_t a = smem_u32(raw);
  return raw + (((a + 1023u) & ~1023u) - a);
}

// Two independent 128-thread compute groups per CTA (warps 0-3 / 4-7): each group owns a tile, its
// own A operand, row stage, mbarriers and TMEM columns, and shares only the weight tiles.  With one
// warp per scheduler a row kernel is latency bound (every dependent instruction stalls); a second
// group lets the schedulers overlap one tile's LayerNorm / epilogue with the other's UMMA wait.
struct Group {
  int grp;   // 0 or 1
  int t;     // thread index inside the group, 0..127 (= tile row = TMEM lane)
  int warp;  // warp index inside the group, 0..3 (= TMEM lane quarter)
  __device__ __forceinline__ Group() {
    grp = threadIdx.x >> 7;
    t = threadIdx.x & 127;
    warp = t >> 5;
  }
  __device__ __forceinline__ void bar() const { asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory"); }
  // make generic-proxy smem writes visible to the tensor core, then sync the group
  __device__ __forceinline__ void sync_before_mma() const {
    fence_proxy_async_smem();
    tc_fence_before();
    bar();
  }
};

// Two compute groups of 256 threads (warps 0-7 / 8-15), TWO threads per tile row: thread (row, half) with
// row = TMEM lane, half = which half of the work of that row (channels of the LayerNorm output, columns of
// the accumulator).  Warps w and w+4 of a group share a TMEM lane quarter, which tcgen05.ld allows (the
// quarter is warp_id % 4).  Twice the warps per scheduler and half the serial work per thread of `Group`.
struct Group2 {
  int grp;   // 0 or 1
  int tt;    // thread index inside the group, 0..255
  int t;     // tile row = TMEM lane, 0..127
  int half;  // 0 or 1
  int warp;  // lane quarter, 0..3
  __device__ __forceinline__ Group2() {
    grp = threadIdx.x >> 8;
    tt = threadIdx.x & 255;
    t = tt & 127;
    half = tt >> 7;
    warp = t >> 5;
  }
  __device__ __forceinline__ void bar() const { asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory"); }
  __device__ __forceinline__ void sync_before_mma() const {
    fence_proxy_async_smem();
    tc_fence_before();
    bar();
  }
};

// All threads: make generic-proxy smem writes visible to the tensor core, then sync the CTA.
__device__ __forceinline__ void sync_before_mma() {
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
}

}  // namespace prd
