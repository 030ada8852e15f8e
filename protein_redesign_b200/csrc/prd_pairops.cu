// Pair-stack kernels that mix single-representation operands into the pair tensor, and the
// coordinate head:
//   outer_linear        OuterLinear (modules.py:283-287) in bilinear form
//   pair_embed_dynamic  per-step pair embedding: RBF distance projection (modules.py:73-82,
//                       model.py:359-361) + time embedding + OuterProductUpdate epilogue
//                       (AF2_modules.py:519-543, modules.py:395-397)
//   coord_head          symmetrise (modules.py:403) + weight_radial + equivariant sum (model.py:364-372)
#include "prd_kernels.h"
#include "prd_rowtile.cuh"

namespace prd {

// =========================================================================================
// OuterLinear:  dst[b,i,j,z] = [pair +] sum_d W1[z,d] x_i[d] x_j[d] + u[b,i,z] - u[b,j,z] + bias[z]
// with x = LN(single), u = x W2^T (precomputed), W = [W1 | W2] = linear.weight[:, :c_s | c_s:].
// CTA (256 threads) = (j-tile of 128 tokens, chunk of i, b).  A = x[b, j-tile, :] stays in shared
// memory (c_s/64 K-blocks, TMA).  W1 lives in REGISTERS (each thread owns a fixed set of 16-byte
// chunks), so for every i the B operand (W1 * x_i) is rebuilt with pure register x smem math;
// c_s/16 UMMAs then produce one [128 x c_z] accumulator.  Two accumulators alternate: the UMMAs of
// row i run while the epilogue of row i-1 (TMEM -> registers -> global) is in flight, and the
// residual pair values of row i-1 are prefetched before the B operand is rebuilt.
// =========================================================================================
template <int CZ, int KBS>
__global__ void __launch_bounds__(256, 1)
outer_linear_kernel(const __grid_constant__ CUtensorMap map_x, const float* pair, float* dst, int residual,
                    const float* __restrict__ xn32, const __half* __restrict__ w1, const float* __restrict__ u,
                    const float* __restrict__ bias, int N, int ilen) {
  extern __shared__ uint8_t raw[];
  constexpr int CS = KBS * 64;
  constexpr int NCH = KBS * CZ * 8 / 256;  // 16-byte W1 chunks per thread
  constexpr int CPK = NCH / KBS;           // ... per K-block
  constexpr int HC = CZ / 2;               // accumulator columns per thread (two warp groups split them)
  constexpr int RB = 4;                    // ring of B-operand K-blocks
  uint8_t* sm = smem_align1024(raw);
  uint8_t* sA = sm;                        // KBS x 16 KB
  uint8_t* sB = sA + KBS * 16384;          // RB x [CZ x 64]
  float* sXi = reinterpret_cast<float*>(sB + RB * CZ * 128);
  float* sUi = sXi + CS;                   // [2][CZ]
  float* sBias = sUi + 2 * CZ;
  uint64_t* bar_a = reinterpret_cast<uint64_t*>(sBias + CZ);
  uint64_t* ring_free = bar_a + 1;         // [RB]
  uint64_t* acc_full = ring_free + RB;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 2);
  constexpr int TCOLS = 2 * CZ < 32 ? 32 : 2 * CZ;

  const int t = threadIdx.x, warp = t >> 5;
  const int grp = warp >> 2;               // 0: columns [0, HC), 1: columns [HC, CZ)
  const int lane_row = (warp & 3) * 32 + (t & 31);
  const int jt = blockIdx.x, b = blockIdx.z;
  const int i0 = blockIdx.y * ilen;
  const int i1 = min(N, i0 + ilen);
  if (t == 0) {
    mbar_init(bar_a, 1);
    for (int q = 0; q < RB; ++q) mbar_init(&ring_free[q], 1);
    mbar_init(&acc_full[0], 1);
    mbar_init(&acc_full[1], 1);
    fence_barrier_init();
    tma_prefetch_desc(&map_x);
  }
  if (warp == 0) tmem_alloc(tmem_slot, TCOLS);
  for (int i = t; i < CZ; i += 256) sBias[i] = bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_lane = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  if (t == 0) {
    mbar_expect_tx(bar_a, KBS * 16384);
    for (int kb = 0; kb < KBS; ++kb) tma_load_3d(sA + kb * 16384, &map_x, bar_a, kb * 64, jt * 128, b);
  }
  // this thread's W1 chunks (fixed for the whole kernel): chunk n covers K-block n / CPK,
  // row z = (t >> 3) + 32 * (n % CPK), 16-byte column chunk ch = t & 7
  const int ch = t & 7;
  uint4 wreg[NCH];
#pragma unroll
  for (int n = 0; n < NCH; ++n) {
    const int kb = n / CPK, z = (t >> 3) + 32 * (n % CPK);
    wreg[n] = __ldg(reinterpret_cast<const uint4*>(w1 + (long long)z * CS + kb * 64 + ch * 8));
  }
  const int j = jt * 128 + lane_row;
  const bool valid = j < N;
  float uj[HC];
  if (valid) {
    const float4* up = reinterpret_cast<const float4*>(u + ((long long)b * N + j) * CZ + grp * HC);
#pragma unroll
    for (int c = 0; c < HC / 4; ++c) {
      const float4 v = __ldg(up + c);
      uj[c * 4] = v.x; uj[c * 4 + 1] = v.y; uj[c * 4 + 2] = v.z; uj[c * 4 + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int c = 0; c < HC; ++c) uj[c] = 0.f;
  }
  mbar_wait(bar_a, 0);

  // epilogue of row `ie` from accumulator `ie & 1`; `res` holds the prefetched residual values
  auto epilogue = [&](int ie, const float4 (&res)[HC / 4]) {
    float* drow = dst + (((long long)b * N + ie) * N + j) * CZ + grp * HC;
    const float* ui = sUi + (ie & 1) * CZ + grp * HC;
    const float* bs = sBias + grp * HC;
#pragma unroll
    for (int c = 0; c < HC / 16; ++c) {
      uint32_t acc[16];
      tmem_ld16(tm_lane + (ie & 1) * CZ + grp * HC + c * 16, acc);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int q = 0; q < 16; q += 4) {
          float4 x = res[c * 4 + q / 4];
          x.x += __uint_as_float(acc[q + 0]) + ui[c * 16 + q + 0] - uj[c * 16 + q + 0] + bs[c * 16 + q + 0];
          x.y += __uint_as_float(acc[q + 1]) + ui[c * 16 + q + 1] - uj[c * 16 + q + 1] + bs[c * 16 + q + 1];
          x.z += __uint_as_float(acc[q + 2]) + ui[c * 16 + q + 2] - uj[c * 16 + q + 2] + bs[c * 16 + q + 2];
          x.w += __uint_as_float(acc[q + 3]) + ui[c * 16 + q + 3] - uj[c * 16 + q + 3] + bs[c * 16 + q + 3];
          *reinterpret_cast<float4*>(drow + c * 16 + q) = x;
        }
      }
    }
  };

  // x_i / u_i are fetched one row ahead into registers so their global latency is off the critical path
  constexpr int XPT = CS / 256;  // floats of x_i per thread
  float xnext[XPT];
  float unext = 0.f;
  auto fetch_row = [&](int i) {
    if (i < i1) {
      const float* xi = xn32 + ((long long)b * N + i) * CS;
#pragma unroll
      for (int q = 0; q < XPT; ++q) xnext[q] = __ldg(xi + t + 256 * q);
      if (t < CZ) unext = __ldg(u + ((long long)b * N + i) * CZ + t);
    }
  };
  fetch_row(i0);
  uint32_t ring_it = 0;  // K-blocks generated so far (ring position)
  for (int i = i0; i <= i1; ++i) {
    // (1) prefetch the residual of row i-1 (consumed by its epilogue at the end of this iteration)
    float4 res[HC / 4];
#pragma unroll
    for (int c = 0; c < HC / 4; ++c) res[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i > i0 && residual && valid) {
      const float4* pr = reinterpret_cast<const float4*>(pair + (((long long)b * N + (i - 1)) * N + j) * CZ + grp * HC);
#pragma unroll
      for (int c = 0; c < HC / 4; ++c) res[c] = pr[c];
    }
    const float ucur = unext;
    if (i < i1) {
      // (2) x_i -> shared (registers filled one iteration ago); every thread finished reading the previous
      // x in the K-block barriers of the previous row
#pragma unroll
      for (int q = 0; q < XPT; ++q) sXi[t + 256 * q] = xnext[q];
      // u_i is read by the epilogue of row i (next iteration); its slot was last read by the epilogue of
      // row i-2, i.e. before the barriers of row i-1
      if (t < CZ) sUi[(i & 1) * CZ + t] = ucur;
    }
    fetch_row(i + 1);
    __syncthreads();
    if (i < i1) {
      // (3) per K-block: B'[z][d] = W1[z][d] * x_i[d] into a ring slot, then its four UMMAs are issued while
      // the next K-block is being built
#pragma unroll
      for (int kb = 0; kb < KBS; ++kb, ++ring_it) {
        const uint32_t slot = ring_it % RB;
        if (ring_it >= RB) mbar_wait(&ring_free[slot], ((ring_it / RB) - 1) & 1);
        const int d0 = kb * 64 + ch * 8;
        const float4 xa = *reinterpret_cast<const float4*>(sXi + d0);
        const float4 xb = *reinterpret_cast<const float4*>(sXi + d0 + 4);
        uint8_t* sBs = sB + slot * (CZ * 128);
#pragma unroll
        for (int n = 0; n < CPK; ++n) {
          const int z = (t >> 3) + 32 * n;
          const __half2* w2 = reinterpret_cast<const __half2*>(&wreg[kb * CPK + n]);
          const float2 f0 = __half22float2(w2[0]), f1 = __half22float2(w2[1]);
          const float2 f2 = __half22float2(w2[2]), f3 = __half22float2(w2[3]);
          uint4 o;
          o.x = pack_half2(f0.x * xa.x, f0.y * xa.y);
          o.y = pack_half2(f1.x * xa.z, f1.y * xa.w);
          o.z = pack_half2(f2.x * xb.x, f2.y * xb.y);
          o.w = pack_half2(f3.x * xb.z, f3.y * xb.w);
          *reinterpret_cast<uint4*>(sBs + sw128_offset(z, ch)) = o;
        }
        sync_before_mma();
        if (t < 32) {  // warp-uniform issue: UMMA operands stay in uniform registers
          tc_fence_after();
          if (elect_one()) {
            umma_kblock(tmem + (i & 1) * CZ, smem_u32(sA) + kb * 16384, smem_u32(sBs), umma_idesc_f16(128, CZ), kb > 0);
            umma_commit(&ring_free[slot]);
            if (kb == KBS - 1) umma_commit(&acc_full[i & 1]);
          }
          __syncwarp();
        }
      }
    }
    // (4) epilogue of row i-1 while the UMMAs of row i drain
    if (i > i0) {
      mbar_wait(&acc_full[(i - 1) & 1], ((i - 1 - i0) >> 1) & 1);
      tc_fence_after();
      epilogue(i - 1, res);
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

template <int CZ, int KBS>
static int launch_outer_linear(const CUtensorMap& mx, dim3 grid, const float* pair, float* dst, int residual,
                               const float* xn32, const __half* w1, const float* u, const float* bias, int N, int ilen,
                               cudaStream_t s) {
  constexpr int smem = 1024 + KBS * 16384 + 4 * CZ * 128 + (KBS * 64 + 3 * CZ) * 4 + 128;
  static_assert(smem <= 227 * 1024, "outer_linear shared memory budget");
  auto kern = outer_linear_kernel<CZ, KBS>;
  PRD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<grid, 256, smem, s>>>(mx, pair, dst, residual, xn32, w1, u, bias, N, ilen);
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
  const int jtiles = (N + 127) / 128;
  // enough CTAs for a few waves; every CTA re-uses its A tile and its W1 registers for `ilen` rows
  int ichunks = (4 * kNumSMs + jtiles * d.B - 1) / (jtiles * d.B);
  if (ichunks > N) ichunks = N;
  if (ichunks < 1) ichunks = 1;
  const int ilen = (N + ichunks - 1) / ichunks;
  ichunks = (N + ilen - 1) / ilen;
  dim3 grid(jtiles, ichunks, d.B);
  if (d.CZ == 64 && CS == 512) return launch_outer_linear<64, 8>(mx, grid, pair, dst, residual, xn32, w1, u, bias, N, ilen, s);
  if (d.CZ == 64 && CS == 256) return launch_outer_linear<64, 4>(mx, grid, pair, dst, residual, xn32, w1, u, bias, N, ilen, s);
  if (d.CZ == 32 && CS == 512) return launch_outer_linear<32, 8>(mx, grid, pair, dst, residual, xn32, w1, u, bias, N, ilen, s);
  return launch_outer_linear<32, 4>(mx, grid, pair, dst, residual, xn32, w1, u, bias, N, ilen, s);
}

// =========================================================================================
// Per-step pair embedding + OuterProductUpdate.  Rows = flattened (b,i,j).
//   pair = static + m2 * (W_dist rbf(|z_i - z_j|) + beta[b]) + m2 * (W_o (a_i * b_j) + b_o) / (m2 + 1e-3)
// rbf_k(d) = exp(-scale (d - center_k)^2) is generated thread-locally straight into the fp16 A
// operand (dist_dim/64 K-blocks); a_i * b_j (opm_dim/64 K-blocks) likewise.  Two accumulators.
// =========================================================================================
template <int CZ>
__global__ void __launch_bounds__(128, 1)
pair_embed_kernel(const float* __restrict__ pstatic, float* __restrict__ pair, const float* __restrict__ z,
                  const float* __restrict__ mask, const float* __restrict__ beta, const __half* __restrict__ w_dist,
                  int DD, const float* __restrict__ centers, float rbf_scale, const float* __restrict__ opm_a,
                  const float* __restrict__ opm_b, int OD, const __half* __restrict__ w_opm,
                  const float* __restrict__ b_opm, int N, long long R, int flags) {
  // flags: 1 = OuterProductUpdate term only (no distance / time embedding); 2 = do not multiply the
  // OPM term by mask_2d (stand-alone OuterProductUpdate.forward).  pstatic may be NULL (= zeros).
  extern __shared__ uint8_t raw[];
  const bool with_dist = (flags & 1) == 0;
  const int KBD = with_dist ? DD / 64 : 0, KBO = OD / 64;
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

  const int t = threadIdx.x, warp = t >> 5;
  if (t == 0) {
    mbar_init(full, kTileRows);
    mbar_init(mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, TCOLS);
  if (with_dist) {
    load_weight_kblocks(sW1, w_dist, CZ, DD, DD, t, 128);
    load_weight_kblocks(sW1 + KBD * CZ * 128, w_dist + CZ * DD, CZ, DD, DD, t, 128);
  }
  load_weight_kblocks(sW2, w_opm, CZ, OD, OD, t, 128);
  if (with_dist)
    for (int i = t; i < DD; i += 128) sC[i] = centers[i];
  for (int i = t; i < CZ; i += 128) sBo[i] = b_opm[i];
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
    bulk_wait_read0();  // previous tile's bulk store has finished reading the stage
    issue_row_load<CZ>(sSt, t, pstatic + r * CZ, valid && pstatic != nullptr, full);
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
    // radial basis -> A1
#pragma unroll 1
    for (int k0 = 0; k0 < KBD * 64; k0 += 32) {
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
    for (int k0 = 0; k0 < OD; k0 += 32) {
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
    if (t < 32) {  // warp-uniform issue: UMMA operands stay in uniform registers
      tc_fence_after();
      if (elect_one()) {
        if (with_dist) {
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
#pragma unroll
    for (int c = 0; c < CZ / 32; ++c) {
      uint32_t a1[32], a2[32];
      tmem_ld32(tm_lane + c * 32, a1);
      tmem_ld32(tm_lane + CZ + c * 32, a2);
      tmem_ld_wait();
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
    if (valid) bulk_s2g(pair + r * CZ, my, CZ * 4);
    bulk_commit();
    tc_fence_before();
    __syncthreads();
  }
  bulk_wait0();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

int pair_embed_dynamic(const PairDims& d, const float* pair_static, float* pair, const float* z, const float* mask,
                       const float* beta, const __half* w_dist, int dist_dim, const float* centers, float rbf_scale,
                       const float* opm_a, const float* opm_b, int opm_dim, const __half* w_opm, const float* b_opm,
                       int flags, cudaStream_t s) {
  PRD_REQUIRE(dist_dim % 64 == 0 && opm_dim % 64 == 0, "pair_embed: dist_dim %d / opm hidden %d must be multiples of 64",
              dist_dim, opm_dim);
  const long long R = (long long)d.B * d.N * d.N;
  const long long tiles = (R + kTileRows - 1) / kTileRows;
  const int KBD = dist_dim / 64, KBO = opm_dim / 64;
  const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
  if (d.CZ == 64) {
    constexpr int CZ = 64;
    const int smem = 1024 + (KBD + KBO) * 16384 + (2 * KBD + KBO) * CZ * 128 + RowStage<CZ>::kBytes + (dist_dim + CZ) * 4 + 64;
    PRD_REQUIRE(smem <= 227 * 1024, "pair_embed: shared memory %d B exceeds 227 KB", smem);
    auto kern = pair_embed_kernel<CZ>;
    PRD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, 128, smem, s>>>(pair_static, pair, z, mask, beta, w_dist, dist_dim, centers, rbf_scale, opm_a, opm_b,
                                 opm_dim, w_opm, b_opm, d.N, R, flags);
  } else if (d.CZ == 32) {
    constexpr int CZ = 32;
    const int smem = 1024 + (KBD + KBO) * 16384 + (2 * KBD + KBO) * CZ * 128 + RowStage<CZ>::kBytes + (dist_dim + CZ) * 4 + 64;
    PRD_REQUIRE(smem <= 227 * 1024, "pair_embed: shared memory %d B exceeds 227 KB", smem);
    auto kern = pair_embed_kernel<CZ>;
    PRD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, 128, smem, s>>>(pair_static, pair, z, mask, beta, w_dist, dist_dim, centers, rbf_scale, opm_a, opm_b,
                                 opm_dim, w_opm, b_opm, d.N, R, flags);
  } else {
    set_error("pair_embed: unsupported pair_dim %d", d.CZ);
    return 1;
  }
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
                  float* __restrict__ eps_raw, int N) {
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
    mbar_init(full, 2 * kTileRows);
    mbar_init(mma_bar, 1);
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
    issue_row_load<CZ>(sSt, t, pair + (((long long)b * N + i) * N + j) * CZ, valid, full);
    issue_row_load<CZ>(sSt + RowStage<CZ>::kBytes, t, pair + (((long long)b * N + j) * N + i) * CZ, valid, full);
    mbar_wait(full, jt & 1);
    {
      float x[CZ], y[CZ];
      if (valid) {
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
    auto kern = coord_head_kernel<CZ>;
    PRD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, 128, smem, s>>>(pair, z, mask, w1, b1, w2, eps_raw, d.N);
  } else if (d.CZ == 32) {
    constexpr int CZ = 32;
    constexpr int smem = 1024 + 16384 + 2 * CZ * 128 + 2 * RowStage<CZ>::kBytes + (2 * CZ + 16) * 4 + 64;
    auto kern = coord_head_kernel<CZ>;
    PRD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, 128, smem, s>>>(pair, z, mask, w1, b1, w2, eps_raw, d.N);
  } else {
    set_error("coord_head: unsupported pair_dim %d", d.CZ);
    return 1;
  }
  PRD_LAUNCHED();
  return 0;
}

}  // namespace prd
