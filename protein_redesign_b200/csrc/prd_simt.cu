// Bandwidth- and latency-bound kernels (no tensor cores): LayerNorm of single-representation rows,
// the pair-bias projection stream (LN + c_z -> H), row softmax, the small-head single attention,
// symmetrise, remove_mean, embeddings and the sampler update.
#include "prd_kernels.h"
#include "prd_common.cuh"

namespace prd {

// -----------------------------------------------------------------------------------------
// LayerNorm over the last dim of [rows, C] fp32 (warp per row), optional affine; writes fp16
// (GEMM A operand) and/or fp32.  nn.LayerNorm semantics, eps 1e-5.
// -----------------------------------------------------------------------------------------
__global__ void layernorm_rows_kernel(const float* __restrict__ x, int rows, int C, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, __half* __restrict__ out16,
                                      float* __restrict__ out32) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* xr = x + (long long)warp * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
  const float mean = warp_sum(s) / C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = xr[c] - mean;
    v += d * d;
  }
  const float rstd = rsqrtf(warp_sum(v) / C + 1e-5f);
  for (int c = lane; c < C; c += 32) {
    float y = (xr[c] - mean) * rstd;
    if (gamma) y = y * gamma[c] + beta[c];
    if (out16) out16[(long long)warp * C + c] = __float2half_rn(y);
    if (out32) out32[(long long)warp * C + c] = y;
  }
}

int layernorm_rows(const float* x, int rows, int C, const float* gamma, const float* beta, __half* out16, float* out32,
                   cudaStream_t s) {
  const int threads = 256;
  const int blocks = (rows * 32 + threads - 1) / threads;
  layernorm_rows_kernel<<<blocks, threads, 0, s>>>(x, rows, C, gamma, beta, out16, out32);
  PRD_LAUNCHED();
  return 0;
}

// -----------------------------------------------------------------------------------------
// Pair-bias projection stream: bias[b,h,i,j] = LN(pair[b,i,j,:]) . w[h,:] (+ bvec[h]).
// HBM-bound: reads the pair tensor once, writes 4/c_z of it.
// -----------------------------------------------------------------------------------------
// Streaming design: one thread per pair element.  Every thread bulk-copies its own 4*CZ-byte row
// (cp.async.bulk + mbarrier) into a padded slot of a 3-stage shared-memory ring, so three rows per
// thread (~200 KB per SM) are in flight no matter what the warps are doing; LayerNorm and the four
// head dot products then run thread-locally in registers (no shuffles), and consecutive threads
// write consecutive positions of the four head planes (full 128-byte lines).
// kTwo: a second projection (own LayerNorm affine, weights, bias) of the SAME rows in the same pass: SPAttention's bias and
// the first FoldingBlock's attn_bias are both functions of the pair tensor the embedding leaves (SPAttention only updates
// the single representation), so one read of P serves both.
template <int CZ, bool kTwo>
__global__ void __launch_bounds__(256, 1)
pair_bias_kernel(const float* __restrict__ pair, long long R, long long NN, const float* __restrict__ ln_w,
                 const float* __restrict__ ln_b, const float* __restrict__ w, const float* __restrict__ bvec,
                 float* __restrict__ out, const float* __restrict__ ln_w2, const float* __restrict__ ln_b2,
                 const float* __restrict__ w2, const float* __restrict__ bvec2, float* __restrict__ out2) {
  constexpr int ROWS = 256;                 // elements per stage = threads per CTA
  constexpr int STAGES = 3;
  constexpr int kRowBytes = CZ * 4 + 16;    // padded: same-column reads of 32 lanes hit distinct banks
  constexpr int kStageBytes = ROWS * kRowBytes;
  extern __shared__ __align__(128) uint8_t smem_pb[];
  float* sW = reinterpret_cast<float*>(smem_pb + STAGES * kStageBytes);  // [4][CZ] folded weights, [4] folded bias (16-byte aligned)
  float* sW2 = sW + 4 * CZ + 4;                                          // the same for the second projection
  uint64_t* full = reinterpret_cast<uint64_t*>(sW2 + 4 * CZ + 4);
  const int t = threadIdx.x;
  pdl_trigger();
  // fold the affine LayerNorm into the projection:  (y*g + b) . w_h = y . (g*w_h) + b . w_h
  for (int i = t; i < 4 * CZ; i += 256) sW[i] = w[i] * (ln_w ? ln_w[i % CZ] : 1.0f);
  if (t < 4) {
    float acc = bvec ? bvec[t] : 0.f;
    if (ln_b)
      for (int c = 0; c < CZ; ++c) acc += ln_b[c] * w[t * CZ + c];
    sW[4 * CZ + t] = acc;
  }
  if (kTwo) {
    for (int i = t; i < 4 * CZ; i += 256) sW2[i] = w2[i] * (ln_w2 ? ln_w2[i % CZ] : 1.0f);
    if (t < 4) {
      float acc = bvec2 ? bvec2[t] : 0.f;
      if (ln_b2)
        for (int c = 0; c < CZ; ++c) acc += ln_b2[c] * w2[t * CZ + c];
      sW2[4 * CZ + t] = acc;
    }
  }
  if (t == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], ROWS);
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();  // everything above touched only weights / shared memory; the predecessor's output is read below
  const long long nchunks = (R + ROWS - 1) / ROWS;
  auto load = [&](long long chunk, int s) {
    const long long e = chunk * ROWS + t;
    uint8_t* dst = smem_pb + s * kStageBytes + t * kRowBytes;
    if (e < R) {
      mbar_expect_tx(&full[s], CZ * 4);
      bulk_g2s(dst, pair + e * CZ, CZ * 4, &full[s]);
    } else {
      mbar_arrive(&full[s]);
    }
  };
  for (int s = 0; s < STAGES; ++s) {
    const long long c = (long long)blockIdx.x + (long long)s * gridDim.x;
    if (c < nchunks) load(c, s);
  }
  int it = 0;
  for (long long chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x, ++it) {
    const int s = it % STAGES;
    mbar_wait(&full[s], (it / STAGES) & 1);
    const float* row = reinterpret_cast<const float*>(smem_pb + s * kStageBytes + t * kRowBytes);
    float x[CZ];
#pragma unroll
    for (int c = 0; c < CZ / 4; ++c) {
      const float4 v = *reinterpret_cast<const float4*>(row + c * 4);
      x[c * 4] = v.x; x[c * 4 + 1] = v.y; x[c * 4 + 2] = v.z; x[c * 4 + 3] = v.w;
    }
    // this thread's slot is free again: refill it for the chunk three rounds ahead
    {
      const long long cn = chunk + (long long)STAGES * gridDim.x;
      if (cn < nchunks) load(cn, s);
    }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < CZ; ++c) sum += x[c];
    const float mean = sum * (1.0f / CZ);
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < CZ; ++c) {
      x[c] -= mean;
      var += x[c] * x[c];
    }
    const float rstd = rsqrtf(var * (1.0f / CZ) + 1e-5f);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < CZ; c += 4) {
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const float4 wv = *reinterpret_cast<const float4*>(sW + h * CZ + c);  // broadcast
        acc[h] += x[c] * wv.x + x[c + 1] * wv.y + x[c + 2] * wv.z + x[c + 3] * wv.w;
      }
    }
    float acc2[4] = {0.f, 0.f, 0.f, 0.f};
    if (kTwo) {
#pragma unroll
      for (int c = 0; c < CZ; c += 4) {
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const float4 wv = *reinterpret_cast<const float4*>(sW2 + h * CZ + c);  // broadcast
          acc2[h] += x[c] * wv.x + x[c + 1] * wv.y + x[c + 2] * wv.z + x[c + 3] * wv.w;
        }
      }
    }
    const long long e = chunk * ROWS + t;
    if (e < R) {
      const long long b = e / NN;
      const long long ij = e - b * NN;
#pragma unroll
      for (int h = 0; h < 4; ++h) out[(b * 4 + h) * NN + ij] = acc[h] * rstd + sW[4 * CZ + h];
      if (kTwo) {
#pragma unroll
        for (int h = 0; h < 4; ++h) out2[(b * 4 + h) * NN + ij] = acc2[h] * rstd + sW2[4 * CZ + h];
      }
    }
  }
}

int pair_bias_proj2(const PairDims& d, int H, const float* pair, const float* ln_w, const float* ln_b, const float* w,
                    const float* bvec, float* bias_out, const float* ln_w2, const float* ln_b2, const float* w2,
                    const float* bvec2, float* bias_out2, cudaStream_t s) {
  PRD_REQUIRE(H == 4, "pair_bias_proj: num_heads %d unsupported (built for 4)", H);
  const long long NN = (long long)d.N * d.N, R = NN * d.B;
  const long long nchunks = (R + 255) / 256;
  const int blocks = (int)(nchunks < kNumSMs ? nchunks : kNumSMs);
  const bool two = bias_out2 != nullptr;
  if (d.CZ == 64) {
    constexpr int smem = 3 * 256 * (64 * 4 + 16) + 64 + 2 * (4 * 64 + 4) * 4;
    auto kern = two ? pair_bias_kernel<64, true> : pair_bias_kernel<64, false>;
    PRD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    PRD_CUDA_OK(launch_pdl(kern, blocks, 256, smem, s, pair, R, NN, ln_w, ln_b, w, bvec, bias_out, ln_w2, ln_b2, w2, bvec2, bias_out2));
  } else if (d.CZ == 32) {
    constexpr int smem = 3 * 256 * (32 * 4 + 16) + 64 + 2 * (4 * 32 + 4) * 4;
    auto kern = two ? pair_bias_kernel<32, true> : pair_bias_kernel<32, false>;
    PRD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    PRD_CUDA_OK(launch_pdl(kern, blocks, 256, smem, s, pair, R, NN, ln_w, ln_b, w, bvec, bias_out, ln_w2, ln_b2, w2, bvec2, bias_out2));
  } else {
    set_error("pair_bias_proj: unsupported pair_dim %d", d.CZ);
    return 1;
  }
  PRD_LAUNCHED();
  return 0;
}
int pair_bias_proj(const PairDims& d, int H, const float* pair, const float* ln_w, const float* ln_b, const float* w,
                   const float* bvec, float* bias_out, cudaStream_t s) {
  return pair_bias_proj2(d, H, pair, ln_w, ln_b, w, bvec, bias_out, nullptr, nullptr, nullptr, nullptr, nullptr, s);
}

// -----------------------------------------------------------------------------------------
// Row softmax: logits fp32 [rows, n] -> probs fp16 [rows, n (ld_out)], pad columns zeroed.
// -----------------------------------------------------------------------------------------
__global__ void softmax_rows_kernel(const float* __restrict__ logits, __half* __restrict__ probs, long long rows, int n,
                                    int ld_in, int ld_out) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* x = logits + row * ld_in;
  float m = -INFINITY;
  for (int c = lane; c < n; c += 32) m = fmaxf(m, x[c]);
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane; c < n; c += 32) s += __expf(x[c] - m);
  s = warp_sum(s);
  const float inv = 1.0f / s;
  __half* p = probs + row * ld_out;
  for (int c = lane; c < ld_out; c += 32) p[c] = __float2half_rn(c < n ? __expf(x[c] - m) * inv : 0.f);
}

int softmax_rows(float* logits, __half* probs, long long rows, int n, int ld_in, int ld_out, cudaStream_t s) {
  const int threads = 256;
  const long long blocks = (rows * 32 + threads - 1) / threads;
  softmax_rows_kernel<<<(unsigned)blocks, threads, 0, s>>>(logits, probs, rows, n, ld_in, ld_out);
  PRD_LAUNCHED();
  return 0;
}

// -----------------------------------------------------------------------------------------
// FoldingBlock.single_attn core (modules.py:185-225): CTA = 32 queries of one (b, head); warp w takes a quarter of the keys
// for ALL 32 queries (lane = query), so every K / V row is a warp-uniform shared-memory read (one broadcast wavefront;
// with four threads per query the four key rows of a warp conflicted and short-scoreboard was 5.7 cycles per issue).  Scores
// of 8 keys at a time (independent dot products), one running-max update per 8 keys; the bias row of a query is read
// straight from global memory (row-contiguous: 32 bytes per lane per 8 keys, the next 8 prefetched).  The four partial
// softmax states of a query are merged through shared memory; each thread then writes 4 of the 16 gated channels.
// K/V of (b,h) in shared memory.  qkvg [B*N, 4*H*16] fp32 (q|k|v|gate pre-activation), bias [B,H,N,N], mask [B,N].
// -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
single_attention_kernel(int N, int H, const float* __restrict__ qkvg, const float* __restrict__ bias,
                        const float* __restrict__ mask, __half* __restrict__ og) {
  extern __shared__ float sm[];
  constexpr int C = 16, QT = 32;
  float* sK = sm;               // [N][16]
  float* sV = sK + (size_t)N * C;
  float* sMask = sV + (size_t)N * C;  // [N]
  float* sPart = sMask + ((N + 3) & ~3);  // [4 warps][32 queries][18]: m, l, o[16]
  const int t = threadIdx.x;
  const int h = blockIdx.y, b = blockIdx.z;
  const int warp = t >> 5, lane = t & 31;
  const int i = blockIdx.x * QT + lane;
  const int ld = 4 * H * C;
  // K / V of this (b, head): 16 floats per token and tensor, staged with 16-byte loads (4 per row and tensor)
  for (int idx = t; idx < N * (C / 4); idx += 128) {
    const int j = idx >> 2, c4 = idx & 3;
    const float* row = qkvg + ((long long)b * N + j) * ld + h * C + c4 * 4;
    reinterpret_cast<float4*>(sK)[idx] = __ldg(reinterpret_cast<const float4*>(row + H * C));
    reinterpret_cast<float4*>(sV)[idx] = __ldg(reinterpret_cast<const float4*>(row + 2 * H * C));
  }
  for (int j = t; j < N; j += 128) sMask[j] = mask[(long long)b * N + j];
  float q[C];
  if (i < N) {
    const float4* row = reinterpret_cast<const float4*>(qkvg + ((long long)b * N + i) * ld + h * C);
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
      const float4 v = __ldg(row + c4);
      q[4 * c4] = 0.25f * v.x; q[4 * c4 + 1] = 0.25f * v.y; q[4 * c4 + 2] = 0.25f * v.z; q[4 * c4 + 3] = 0.25f * v.w;
    }
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) q[c] = 0.f;
  }
  float m = -INFINITY, l = 0.f, o[C];
#pragma unroll
  for (int c = 0; c < C; ++c) o[c] = 0.f;
  // this warp's key range: a quarter of the keys, in whole groups of 8
  const int chunk = (((N + 3) / 4) + 7) & ~7;
  const int jbeg = warp * chunk, jend = min(N, jbeg + chunk);
  const float* brow = bias + (((long long)b * H + h) * N + min(i, N - 1)) * N;
  const bool vec = (N & 3) == 0;  // 16-byte aligned bias rows
  float bn[8];
  auto fetch_bias = [&](int j0) {
    if (vec && j0 + 8 <= jend) {
      const float4 x0 = __ldg(reinterpret_cast<const float4*>(brow + j0)), x1 = __ldg(reinterpret_cast<const float4*>(brow + j0 + 4));
      bn[0] = x0.x; bn[1] = x0.y; bn[2] = x0.z; bn[3] = x0.w; bn[4] = x1.x; bn[5] = x1.y; bn[6] = x1.z; bn[7] = x1.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) bn[e] = (j0 + e < jend) ? __ldg(brow + j0 + e) : 0.f;
    }
  };
  if (jbeg < jend) fetch_bias(jbeg);
  __syncthreads();  // K / V / mask staged
  for (int j0 = jbeg; j0 < jend; j0 += 8) {
    float bc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) bc[e] = bn[e];
    if (j0 + 8 < jend) fetch_bias(j0 + 8);
    float sc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int j = j0 + e;
      if (j < jend) {  // warp-uniform
        const float4* kr = reinterpret_cast<const float4*>(sK + j * C);
        const float4 k0 = kr[0], k1 = kr[1], k2 = kr[2], k3 = kr[3];
        const float p0 = q[0] * k0.x + q[1] * k0.y + q[2] * k0.z + q[3] * k0.w;
        const float p1 = q[4] * k1.x + q[5] * k1.y + q[6] * k1.z + q[7] * k1.w;
        const float p2 = q[8] * k2.x + q[9] * k2.y + q[10] * k2.z + q[11] * k2.w;
        const float p3 = q[12] * k3.x + q[13] * k3.y + q[14] * k3.z + q[15] * k3.w;
        float sv = ((p0 + p1) + (p2 + p3)) + bc[e];
        if (sMask[j] < 0.5f) sv = -32768.0f;
        sc[e] = sv;
      } else {
        sc[e] = -INFINITY;
      }
    }
    float mn = m;
#pragma unroll
    for (int e = 0; e < 8; ++e) mn = fmaxf(mn, sc[e]);
    const float a = __expf(m - mn);  // m = -inf (first group) -> 0; mn is finite: the group holds at least one key
    l *= a;
#pragma unroll
    for (int c = 0; c < C; ++c) o[c] *= a;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int j = j0 + e;
      if (j < jend) {
        const float pe = __expf(sc[e] - mn);
        l += pe;
        const float4* vr = reinterpret_cast<const float4*>(sV + j * C);
        const float4 v0 = vr[0], v1 = vr[1], v2 = vr[2], v3 = vr[3];
        o[0] += pe * v0.x; o[1] += pe * v0.y; o[2] += pe * v0.z; o[3] += pe * v0.w;
        o[4] += pe * v1.x; o[5] += pe * v1.y; o[6] += pe * v1.z; o[7] += pe * v1.w;
        o[8] += pe * v2.x; o[9] += pe * v2.y; o[10] += pe * v2.z; o[11] += pe * v2.w;
        o[12] += pe * v3.x; o[13] += pe * v3.y; o[14] += pe * v3.z; o[15] += pe * v3.w;
      }
    }
    m = mn;
  }
  // partial state of (warp, query) -> shared memory; thread (w, lane) then merges the four states of query `lane` and
  // writes channels [4 w, 4 w + 4)
  {
    float* p = sPart + (warp * QT + lane) * 18;
    p[0] = m;
    p[1] = l;
#pragma unroll
    for (int c = 0; c < C; ++c) p[2 + c] = o[c];
  }
  __syncthreads();
  if (i < N) {
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < 4; ++w) mm = fmaxf(mm, sPart[(w * QT + lane) * 18]);
    float lsum = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float* p = sPart + (w * QT + lane) * 18;
      const float sc = (p[0] == -INFINITY) ? 0.f : __expf(p[0] - mm);
      lsum += p[1] * sc;
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[e] += p[2 + 4 * warp + e] * sc;
    }
    const float inv = 1.0f / lsum;
    const float* row = qkvg + ((long long)b * N + i) * ld + 3 * H * C + h * C + 4 * warp;
    __half* dst = og + ((long long)b * N + i) * (H * C) + h * C + 4 * warp;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gate = 1.0f / (1.0f + __expf(-row[e]));
      dst[e] = __float2half_rn(gate * acc[e] * inv);
    }
  }
}

int single_attention(int B, int N, int H, int c, const float* qkvg, const float* bias, const float* mask, __half* og,
                     cudaStream_t s) {
  PRD_REQUIRE(c == 16, "single_attention: head_dim %d unsupported (built for 16)", c);
  const int smem = (2 * N * 16 + ((N + 3) & ~3) + 4 * 32 * 18) * 4;
  PRD_REQUIRE(smem <= 227 * 1024, "single_attention: N=%d needs %d B of shared memory", N, smem);
  PRD_CUDA_OK(cudaFuncSetAttribute(single_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid((N + 31) / 32, H, B);
  single_attention_kernel<<<grid, 128, smem, s>>>(N, H, qkvg, bias, mask, og);
  PRD_LAUNCHED();
  return 0;
}

// -----------------------------------------------------------------------------------------
// pair <- 0.5 (pair + pair^T)   (modules.py:403), in place; one thread per float4 of an (i<j) pair
// -----------------------------------------------------------------------------------------
__global__ void symmetrize_kernel(float* __restrict__ pair, int N, int CZ4, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % CZ4);
  const long long e = idx / CZ4;
  const int j = static_cast<int>(e % N);
  const long long bi = e / N;
  const int i = static_cast<int>(bi % N);
  if (j <= i) return;
  const long long b = bi / N;
  float4* pij = reinterpret_cast<float4*>(pair) + ((b * N + i) * N + j) * CZ4 + c;
  float4* pji = reinterpret_cast<float4*>(pair) + ((b * N + j) * N + i) * CZ4 + c;
  const float4 a = *pij, d = *pji;
  const float4 r = make_float4(0.5f * (a.x + d.x), 0.5f * (a.y + d.y), 0.5f * (a.z + d.z), 0.5f * (a.w + d.w));
  *pij = r;
  *pji = r;
}

int symmetrize_pair(const PairDims& d, float* pair, cudaStream_t s) {
  const int CZ4 = d.CZ / 4;
  const long long total = (long long)d.B * d.N * d.N * CZ4;
  symmetrize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(pair, d.N, CZ4, total);
  PRD_LAUNCHED();
  return 0;
}

// -----------------------------------------------------------------------------------------
// remove_mean (utils.py:32-36): x[b,i,:] -= mask[b,i] * sum_i(mask*x) / sum_i(mask); one CTA per b
// -----------------------------------------------------------------------------------------
__global__ void remove_mean_kernel(int N, int C, float* __restrict__ x, const float* __restrict__ mask, int mask_rows) {
  __shared__ float sSum[32];
  __shared__ float sCnt;
  const int b = blockIdx.x, t = threadIdx.x;
  float* xb = x + (long long)b * N * C;
  const float* mb = mask + (long long)(b % mask_rows) * N;
  if (t < 32) sSum[t] = 0.f;
  if (t == 0) sCnt = 0.f;
  __syncthreads();
  // one warp per channel keeps the summation order fixed (deterministic)
  const int warp = t >> 5, lane = t & 31, nwarps = blockDim.x >> 5;
  for (int c = warp; c < C; c += nwarps) {
    float s = 0.f;
    for (int i = lane; i < N; i += 32) s += mb[i] * xb[(long long)i * C + c];
    s = warp_sum(s);
    if (lane == 0) sSum[c] = s;
  }
  if (warp == 0) {
    float n = 0.f;
    for (int i = lane; i < N; i += 32) n += mb[i];
    n = warp_sum(n);
    if (lane == 0) sCnt = n;
  }
  __syncthreads();
  for (int idx = t; idx < N * C; idx += blockDim.x) {
    const int i = idx / C, c = idx % C;
    xb[idx] -= mb[i] * sSum[c] / sCnt;
  }
}

int remove_mean3(int B, int N, int C, float* x, const float* mask, int mask_rows, cudaStream_t s) {
  PRD_REQUIRE(C >= 1 && C <= 32, "remove_mean: channel count %d not in [1,32]", C);
  PRD_REQUIRE(mask_rows >= 1, "remove_mean: mask_rows %d", mask_rows);
  remove_mean_kernel<<<B, 256, 0, s>>>(N, C, x, mask, mask_rows);
  PRD_LAUNCHED();
  return 0;
}

// -----------------------------------------------------------------------------------------
// Step-invariant pair embedding (model.py:348-358): bond features / bond distance / relative
// position gathers.  One thread per float4 of channels.
// -----------------------------------------------------------------------------------------
struct BondTables {
  const float* t[3];
};

__global__ void embed_pair_static_kernel(int N, int CZ4, long long total, const float* __restrict__ atom_mask,
                                         const float* __restrict__ residue_mask, const float* __restrict__ bond_mask,
                                         const int64_t* __restrict__ bond_feats, const int64_t* __restrict__ bond_distance,
                                         const int64_t* __restrict__ residue_index, const int64_t* __restrict__ chain_index,
                                         BondTables bt, const float* __restrict__ bdist_table, int max_bd,
                                         const float* __restrict__ relpos_table, int max_rel, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % CZ4);
  const long long e = idx / CZ4;
  const int j = static_cast<int>(e % N);
  const long long bi = e / N;
  const int i = static_cast<int>(bi % N);
  const long long b = bi / N;
  const float am2 = atom_mask[b * N + i] * atom_mask[b * N + j];
  const float rm2 = residue_mask[b * N + i] * residue_mask[b * N + j];
  auto row4 = [&](const float* table, long long r) { return reinterpret_cast<const float4*>(table)[r * CZ4 + c]; };
  const float scale = 0.57735026918962584f;  // 1/sqrt(3) bond features (modules.py:63)
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  {
    float4 bond = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      // nn.Embedding raises on an out-of-range index; the kernel clamps into the table (features.py:52-60 vocabularies)
      constexpr int kBondVocab[3] = {5, 6, 2};
      long long bf = bond_feats[e * 3 + f];
      bf = bf < 0 ? 0 : (bf >= kBondVocab[f] ? kBondVocab[f] - 1 : bf);
      const float4 v = row4(bt.t[f], bf);
      bond.x += scale * v.x; bond.y += scale * v.y; bond.z += scale * v.z; bond.w += scale * v.w;
    }
    const float bm = bond_mask[e];
    long long bd = bond_distance[e];
    if (bd > max_bd) bd = max_bd;
    if (bd < 0) bd = 0;
    const float4 dv = row4(bdist_table, bd);
    acc.x = am2 * (bm * bond.x + dv.x); acc.y = am2 * (bm * bond.y + dv.y);
    acc.z = am2 * (bm * bond.z + dv.z); acc.w = am2 * (bm * bond.w + dv.w);
  }
  {
    long long rel = residue_index[b * N + i] - residue_index[b * N + j];
    if (rel < -max_rel) rel = -max_rel;
    if (rel > max_rel) rel = max_rel;
    const float same = (chain_index[b * N + i] == chain_index[b * N + j]) ? 1.f : 0.f;
    const float4 rv = row4(relpos_table, max_rel + rel);
    acc.x += rm2 * (same * rv.x); acc.y += rm2 * (same * rv.y);
    acc.z += rm2 * (same * rv.z); acc.w += rm2 * (same * rv.w);
  }
  reinterpret_cast<float4*>(out)[idx] = acc;
}

int embed_pair_static(const PairDims& d, const float* atom_mask, const float* residue_mask, const float* bond_mask,
                      const int64_t* bond_feats, const int64_t* bond_distance, const int64_t* residue_index,
                      const int64_t* chain_index, const float* const* bond_tables, const int* bond_vocab,
                      const float* bdist_table, int max_bond_distance, const float* relpos_table, int max_relpos,
                      float* out, cudaStream_t s) {
  (void)bond_vocab;
  const int CZ4 = d.CZ / 4;
  const long long total = (long long)d.B * d.N * d.N * CZ4;
  BondTables bt{{bond_tables[0], bond_tables[1], bond_tables[2]}};
  embed_pair_static_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
      d.N, CZ4, total, atom_mask, residue_mask, bond_mask, bond_feats, bond_distance, residue_index, chain_index, bt,
      bdist_table, max_bond_distance, relpos_table, max_relpos, out);
  PRD_LAUNCHED();
  return 0;
}

}  // namespace prd
