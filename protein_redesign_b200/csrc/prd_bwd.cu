// Backward-pass SIMT kernels (see prd_bwd.h for the conventions).  All of them are bandwidth- or FFMA-bound helpers
// around the tf32 tensor-core GEMM; reference math: ProteinReDiff/modules.py, models/AF2_modules.py, model.py
// (cited per kernel), differentiated by hand.
#include "prd_bwd.h"

#include <stdlib.h>

#include <initializer_list>

#include "prd_common.cuh"

namespace prd {

int bw_colsum(const float* dY, long long ldy, long long R, int Nout, float* db, float alpha, cudaStream_t s);

namespace {
constexpr float kLnEps = 1e-5f;
constexpr float kMaskFill = -32768.0f;  // modules.py:177,220

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
// 2 ulp of ex2.approx and of the approximate division: far below the tf32 rounding (2^-11) every user applies next
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 round4(float4 v) { return make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w)); }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
// sum over the LPR-lane group this lane belongs to (LPR a power of two, groups aligned)
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
inline unsigned grid_for(long long n, int per_block, long long cap = 148LL * 32) {
  long long g = (n + per_block - 1) / per_block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}
}  // namespace

// ------------------------------------------------------------------------------------------------------------
// LayerNorm forward / backward, one warp per row (rows are 21 .. 2048 wide)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bw_ln_fwd_kernel(const float* __restrict__ x, long long R, int C,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float* __restrict__ out, float* __restrict__ out_lo, long long ldo) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = warp0; r < R; r += nw) {
    const float* xr = x + r * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += xr[c];
    const float mean = warp_sum(s) / C;
    float v = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float d = xr[c] - mean;
      v += d * d;
    }
    const float rstd = rsqrtf(warp_sum(v) / C + kLnEps);
    for (int c = lane; c < C; c += 32) {
      float y = (xr[c] - mean) * rstd;
      if (gamma) y = y * gamma[c] + (beta ? beta[c] : 0.f);
      const float hi = round_tf32(y);
      out[r * ldo + c] = hi;
      if (out_lo) out_lo[r * ldo + c] = round_tf32(y - hi);
    }
  }
}
// Rows of C = 4 * LPR floats (32 / 64 / 128: every pair-side LayerNorm): LPR lanes x float4 per row, 32 / LPR rows per warp
// and load instruction, two such row groups in flight per warp.  The one-warp-per-row kernel above keeps a single 256-byte
// row in flight per warp and reached 2.1 TB/s.
template <int LPR>
__global__ void __launch_bounds__(256) bw_ln_fwd_vec_kernel(const float* __restrict__ x, long long R,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* __restrict__ out, float* __restrict__ out_lo, long long ldo) {
  constexpr int C = 4 * LPR, RPW = 32 / LPR, U = 2;
  const int lane = threadIdx.x & 31, sub = lane / LPR, l = lane % LPR;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  float4 gm = make_float4(1.f, 1.f, 1.f, 1.f), bt = make_float4(0.f, 0.f, 0.f, 0.f);
  if (gamma) gm = ld4(gamma + 4 * l);
  if (gamma && beta) bt = ld4(beta + 4 * l);
  for (long long r0 = warp0 * (RPW * U); r0 < R; r0 += nw * (RPW * U)) {
    float4 v[U];
    long long r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      r[u] = r0 + u * RPW + sub;
      v[u] = r[u] < R ? ld4(x + r[u] * C + 4 * l) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float mean = group_sum<LPR>(v[u].x + v[u].y + v[u].z + v[u].w) * (1.0f / C);
      const float4 d = make_float4(v[u].x - mean, v[u].y - mean, v[u].z - mean, v[u].w - mean);
      const float rstd = rsqrtf(group_sum<LPR>(d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w) * (1.0f / C) + kLnEps);
      const float4 y = make_float4(fmaf(d.x * rstd, gm.x, bt.x), fmaf(d.y * rstd, gm.y, bt.y), fmaf(d.z * rstd, gm.z, bt.z),
                                   fmaf(d.w * rstd, gm.w, bt.w));
      const float4 hi = round4(y);
      if (r[u] < R) {
        st4(out + r[u] * ldo + 4 * l, hi);
        if (out_lo) st4(out_lo + r[u] * ldo + 4 * l, round4(make_float4(y.x - hi.x, y.y - hi.y, y.z - hi.z, y.w - hi.w)));
      }
    }
  }
}
template <int LPR>
static void launch_ln_fwd_vec(const float* x, long long R, const float* gamma, const float* beta, float* out, float* out_lo,
                              long long ldo, cudaStream_t s) {
  bw_ln_fwd_vec_kernel<LPR><<<grid_for(R, 8 * (32 / LPR) * 2), 256, 0, s>>>(x, R, gamma, beta, out, out_lo, ldo);
}
int bw_ln_fwd(const float* x, long long R, int C, const float* gamma, const float* beta, float* out, cudaStream_t s,
              float* out_lo, long long ldo) {
  if (ldo == 0) ldo = C;
  const bool vec = aligned16(x) && aligned16(out) && (out_lo == nullptr || aligned16(out_lo)) && ldo % 4 == 0 &&
                   (gamma == nullptr || aligned16(gamma)) && (beta == nullptr || aligned16(beta));
  if (vec && (C == 32 || C == 64 || C == 128)) {
    if (C == 32) launch_ln_fwd_vec<8>(x, R, gamma, beta, out, out_lo, ldo, s);
    else if (C == 64) launch_ln_fwd_vec<16>(x, R, gamma, beta, out, out_lo, ldo, s);
    else launch_ln_fwd_vec<32>(x, R, gamma, beta, out, out_lo, ldo, s);
    PRD_LAUNCHED();
    return 0;
  }
  bw_ln_fwd_kernel<<<grid_for(R, 8), 256, 0, s>>>(x, R, C, gamma, beta, out, out_lo, ldo);
  PRD_LAUNCHED();
  return 0;
}

// dgamma / dbeta are accumulated per warp in shared memory (C <= 2048 floats each would be 16 KB x 8 warps: too much),
// so they go straight to global atomics per row chunk instead: each warp keeps its lane's columns in registers for
// C <= 1024 (kMaxE = 32) -- every affine LayerNorm of the model is 64 or 512 wide.
template <int kMaxE>
__global__ void __launch_bounds__(256) bw_ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, long long R,
                                                        int C, const float* __restrict__ gamma, float* __restrict__ dx_io,
                                                        int accumulate, float* __restrict__ dgamma,
                                                        float* __restrict__ dbeta) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  float accg[kMaxE], accb[kMaxE];
#pragma unroll
  for (int e = 0; e < kMaxE; ++e) accg[e] = accb[e] = 0.f;
  for (long long r = warp0; r < R; r += nw) {
    const float* xr = x + r * C;
    const float* gr = g + r * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += xr[c];
    const float mean = warp_sum(s) / C;
    float v = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float d = xr[c] - mean;
      v += d * d;
    }
    const float rstd = rsqrtf(warp_sum(v) / C + kLnEps);
    float s1 = 0.f, s2 = 0.f;
    int e = 0;
    for (int c = lane; c < C; c += 32, ++e) {
      const float xh = (xr[c] - mean) * rstd;
      const float gy = gr[c];
      const float gg = gamma ? gy * gamma[c] : gy;
      s1 += gg;
      s2 += gg * xh;
      if (kMaxE > 1 && e < kMaxE) {
        accg[e] += gy * xh;
        accb[e] += gy;
      }
    }
    const float m1 = warp_sum(s1) / C, m2 = warp_sum(s2) / C;
    for (int c = lane; c < C; c += 32) {
      const float xh = (xr[c] - mean) * rstd;
      const float gg = gamma ? gr[c] * gamma[c] : gr[c];
      const float d = rstd * (gg - m1 - xh * m2);
      const long long o = r * C + c;
      dx_io[o] = round_tf32(accumulate ? dx_io[o] + d : d);
    }
  }
  if (kMaxE > 1) {
    int e = 0;
    for (int c = lane; c < C && e < kMaxE; c += 32, ++e) {
      if (dgamma) atomicAdd(dgamma + c, accg[e]);
      if (dbeta) atomicAdd(dbeta + c, accb[e]);
    }
  }
}
// The non-affine backward on rows of C = 4 * LPR floats, laid out like bw_ln_fwd_vec_kernel.
template <int LPR>
__global__ void __launch_bounds__(256) bw_ln_bwd_vec_kernel(const float* __restrict__ x, const float* __restrict__ g, long long R,
                                                            float* __restrict__ dx_io, int accumulate) {
  constexpr int C = 4 * LPR, RPW = 32 / LPR, U = 2;
  const int lane = threadIdx.x & 31, sub = lane / LPR, l = lane % LPR;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r0 = warp0 * (RPW * U); r0 < R; r0 += nw * (RPW * U)) {
    float4 v[U], gy[U], acc[U];
    long long r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      r[u] = r0 + u * RPW + sub;
      const bool ok = r[u] < R;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      v[u] = ok ? ld4(x + r[u] * C + 4 * l) : z;
      gy[u] = ok ? ld4(g + r[u] * C + 4 * l) : z;
      acc[u] = (ok && accumulate) ? ld4(dx_io + r[u] * C + 4 * l) : z;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float mean = group_sum<LPR>(v[u].x + v[u].y + v[u].z + v[u].w) * (1.0f / C);
      const float4 d = make_float4(v[u].x - mean, v[u].y - mean, v[u].z - mean, v[u].w - mean);
      const float rstd = rsqrtf(group_sum<LPR>(d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w) * (1.0f / C) + kLnEps);
      const float4 xh = make_float4(d.x * rstd, d.y * rstd, d.z * rstd, d.w * rstd);
      const float m1 = group_sum<LPR>(gy[u].x + gy[u].y + gy[u].z + gy[u].w) * (1.0f / C);
      const float m2 = group_sum<LPR>(gy[u].x * xh.x + gy[u].y * xh.y + gy[u].z * xh.z + gy[u].w * xh.w) * (1.0f / C);
      const float4 o = make_float4(acc[u].x + rstd * (gy[u].x - m1 - xh.x * m2), acc[u].y + rstd * (gy[u].y - m1 - xh.y * m2),
                                   acc[u].z + rstd * (gy[u].z - m1 - xh.z * m2), acc[u].w + rstd * (gy[u].w - m1 - xh.w * m2));
      if (r[u] < R) st4(dx_io + r[u] * C + 4 * l, round4(o));
    }
  }
}
int bw_ln_bwd(const float* x, const float* g, long long R, int C, const float* gamma, float* dx_io, int accumulate,
              float* dgamma, float* dbeta, cudaStream_t s) {
  if (!dgamma && !dbeta && !gamma && (C == 32 || C == 64 || C == 128) && aligned16(x) && aligned16(g) && aligned16(dx_io)) {
    const unsigned grid = grid_for(R, 8 * (128 / C) * 2);
    if (C == 32) bw_ln_bwd_vec_kernel<8><<<grid, 256, 0, s>>>(x, g, R, dx_io, accumulate);
    else if (C == 64) bw_ln_bwd_vec_kernel<16><<<grid, 256, 0, s>>>(x, g, R, dx_io, accumulate);
    else bw_ln_bwd_vec_kernel<32><<<grid, 256, 0, s>>>(x, g, R, dx_io, accumulate);
    PRD_LAUNCHED();
    return 0;
  }
  if (dgamma || dbeta) {
    PRD_REQUIRE(C <= 1024, "ln_bwd: affine LayerNorm wider than 1024 (%d)", C);
    bw_ln_bwd_kernel<32><<<grid_for(R, 8, 148 * 4), 256, 0, s>>>(x, g, R, C, gamma, dx_io, accumulate, dgamma, dbeta);
  } else {
    bw_ln_bwd_kernel<1><<<grid_for(R, 8), 256, 0, s>>>(x, g, R, C, gamma, dx_io, accumulate, nullptr, nullptr);
  }
  PRD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// dW[n,k] += alpha sum_r dY[r,n] X[r,k]: 64 x 64 output tile per CTA over a chunk of rows, 4 x 4 per thread, exact fp32
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bw_dw_acc_kernel(const float* __restrict__ dY, long long ldy, const float* __restrict__ X,
                                                        long long ldx, long long R, int Nout, int K, float* __restrict__ dW,
                                                        long long ldw, float* __restrict__ db, float alpha,
                                                        long long rows_per_cta) {
  __shared__ __align__(16) float sA[32][64];
  __shared__ __align__(16) float sX[32][64];
  const int k0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  const long long r_begin = (long long)blockIdx.z * rows_per_cta;
  long long r_end = r_begin + rows_per_cta;
  if (r_end > R) r_end = R;
  const int tid = threadIdx.x, tn = tid >> 4, tk = tid & 15;
  float acc[4][4];
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  const bool do_bias = db != nullptr && blockIdx.x == 0 && tk == 0;
  for (long long r0 = r_begin; r0 < r_end; r0 += 32) {
    __syncthreads();
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int idx = it * 256 + tid, rr = idx >> 6, cc = idx & 63;
      const long long r = r0 + rr;
      const bool rv = r < r_end;
      sA[rr][cc] = (rv && n0 + cc < Nout) ? dY[r * ldy + n0 + cc] : 0.f;
      sX[rr][cc] = (rv && k0 + cc < K) ? X[r * ldx + k0 + cc] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < 32; ++rr) {
      const float4 a4 = *reinterpret_cast<const float4*>(&sA[rr][tn * 4]);
      const float4 x4 = *reinterpret_cast<const float4*>(&sX[rr][tk * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], xv[b], acc[a][b]);
        bsum[a] += av[a];
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int n = n0 + tn * 4 + a;
    if (n >= Nout) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int k = k0 + tk * 4 + b;
      if (k < K) atomicAdd(dW + (long long)n * ldw + k, alpha * acc[a][b]);
    }
    if (do_bias) atomicAdd(db + n, alpha * bsum[a]);
  }
}
int bw_dw_acc(const float* dY, long long ldy, const float* X, long long ldx, long long R, int Nout, int K, float* dW,
              long long ldw, float* db, float alpha, cudaStream_t s) {
  if (R <= 0) return 0;
  static const bool simt_only = getenv("PRD_DW_SIMT") && getenv("PRD_DW_SIMT")[0] == '1';  // A/B switch
  if (!simt_only && bw_dw_tc_applies(dY, ldy, X, ldx, R)) {
    return bw_dw_tc(dY, ldy, X, ldx, R, Nout, K, dW, ldw, alpha, s, db);
  }
  const int gx = (K + 63) / 64, gy = (Nout + 63) / 64;
  long long want = (148LL * 4 + gx * gy - 1) / (gx * gy);  // ~4 CTAs per SM in total
  long long chunks = (R + 255) / 256;
  if (chunks > want) chunks = want;
  if (chunks < 1) chunks = 1;
  long long rows_per_cta = ((R + chunks - 1) / chunks + 31) / 32 * 32;
  chunks = (R + rows_per_cta - 1) / rows_per_cta;
  bw_dw_acc_kernel<<<dim3(gx, gy, (unsigned)chunks), 256, 0, s>>>(dY, ldy, X, ldx, R, Nout, K, dW, ldw, db, alpha, rows_per_cta);
  PRD_LAUNCHED();
  return 0;
}

// db[n] += alpha sum_r dY[r, n]
__global__ void bw_colsum_kernel(const float* __restrict__ dY, long long ldy, long long R, int Nout, float* __restrict__ db,
                                 float alpha) {
  // block (32, 8): x = column within a 32-wide strip, y = row lane
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (n < Nout)
    for (long long r = (long long)blockIdx.y * 8 + threadIdx.y; r < R; r += (long long)gridDim.y * 8) acc += dY[r * ldy + n];
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && n < Nout) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    atomicAdd(db + n, alpha * t);
  }
}
int bw_colsum(const float* dY, long long ldy, long long R, int Nout, float* db, float alpha, cudaStream_t s) {
  long long gy = (R + 63) / 64;
  if (gy > 148 * 4) gy = 148 * 4;
  if (gy < 1) gy = 1;
  bw_colsum_kernel<<<dim3((Nout + 31) / 32, (unsigned)gy), dim3(32, 8), 0, s>>>(dY, ldy, R, Nout, db, alpha);
  PRD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// small data-movement kernels
// ------------------------------------------------------------------------------------------------------------
__global__ void bw_transpose_kernel(const float* __restrict__ src, long long lds, long long src_bs, float* __restrict__ dst,
                                    long long ldd, long long dst_bs, int rows, int cols, float alpha) {
  __shared__ float tile[32][33];
  const float* sp = src + (long long)blockIdx.z * src_bs;
  float* dp = dst + (long long)blockIdx.z * dst_bs;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? sp[(long long)r * lds + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) dp[(long long)c * ldd + r] = round_tf32(alpha * tile[threadIdx.x][i]);
  }
}
int bw_transpose(const float* src, long long lds, long long src_bs, float* dst, long long ldd, long long dst_bs, int rows,
                 int cols, int batch, float alpha, cudaStream_t s) {
  // gridDim.z is limited to 65535: loop over batch slabs
  for (int b0 = 0; b0 < batch; b0 += 32768) {
    const int nb = batch - b0 < 32768 ? batch - b0 : 32768;
    bw_transpose_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32, nb), dim3(32, 8), 0, s>>>(
        src + (long long)b0 * src_bs, lds, src_bs, dst + (long long)b0 * dst_bs, ldd, dst_bs, rows, cols, alpha);
    PRD_LAUNCHED();
  }
  return 0;
}

__global__ void bw_copy2d_kernel(const float* src, long long lds, float* dst, long long ldd,
                                 long long R, int cols) {
  const long long total = R * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    dst[r * ldd + c] = round_tf32(src[r * lds + c]);
  }
}
int bw_copy2d(const float* src, long long lds, float* dst, long long ldd, long long R, int cols, cudaStream_t s) {
  bw_copy2d_kernel<<<grid_for(R * cols, 256), 256, 0, s>>>(src, lds, dst, ldd, R, cols);
  PRD_LAUNCHED();
  return 0;
}
// hi = round(src), lo = round(src - hi): src ~= hi + lo to 21 mantissa bits
__global__ void bw_split2d_kernel(const float* __restrict__ src, long long lds, float* __restrict__ hi, float* __restrict__ lo,
                                  long long ldd, long long R, int cols) {
  const long long total = R * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    const float v = src[r * lds + c];
    const float h = round_tf32(v);
    hi[r * ldd + c] = h;
    lo[r * ldd + c] = round_tf32(v - h);
  }
}
int bw_split2d(const float* src, long long lds, float* hi, float* lo, long long ldd, long long R, int cols, cudaStream_t s) {
  bw_split2d_kernel<<<grid_for(R * cols, 256), 256, 0, s>>>(src, lds, hi, lo, ldd, R, cols);
  PRD_LAUNCHED();
  return 0;
}
struct PrepJobs {
  PrepJob j[PrepBatch::kMax];
};
__global__ void __launch_bounds__(256) bw_prep_kernel(const __grid_constant__ PrepJobs jobs) {
  const PrepJob& jb = jobs.j[blockIdx.y];
  const int total = jb.drows * jb.dcols;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int i = idx / jb.dcols, j = idx - i * jb.dcols;
    const int sr = jb.transpose ? j : i, sc = jb.transpose ? i : j;
    float v = 0.f;
    if (jb.src != nullptr && sr < jb.rows && sc < jb.cols) v = jb.alpha * jb.src[(long long)sr * jb.lds + sc];
    const float hi = jb.round ? round_tf32(v) : v;
    jb.dst[(long long)i * jb.ldd + j] = hi;
    if (jb.dst_lo != nullptr) jb.dst_lo[(long long)i * jb.ldd + j] = round_tf32(v - hi);
  }
}
int bw_prep(const PrepBatch& b, cudaStream_t s) {
  PRD_REQUIRE(b.err == 0, "prep: more than %d jobs in one batch", PrepBatch::kMax);
  if (b.n == 0) return 0;
  PrepJobs jobs;
  for (int i = 0; i < b.n; ++i) jobs.j[i] = b.jobs[i];
  for (int i = b.n; i < PrepBatch::kMax; ++i) jobs.j[i] = PrepJob{nullptr, nullptr, nullptr, 0, 0, 0, 0, 0, 0, 0, 0, 0.f};
  bw_prep_kernel<<<dim3(16, b.n), 256, 0, s>>>(jobs);
  PRD_LAUNCHED();
  return 0;
}

__global__ void bw_relu_kernel(float* x, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = round_tf32(fmaxf(x[i], 0.f));
}
__global__ void bw_relu_vec_kernel(float* x, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = ld4(x + 4 * i);
    st4(x + 4 * i, round4(make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f))));
  }
}
int bw_relu_inplace(float* x, long long n, cudaStream_t s) {
  if (n % 4 == 0 && aligned16(x)) {
    bw_relu_vec_kernel<<<grid_for(n / 4, 256), 256, 0, s>>>(x, n / 4);
    PRD_LAUNCHED();
    return 0;
  }
  bw_relu_kernel<<<grid_for(n, 256), 256, 0, s>>>(x, n);
  PRD_LAUNCHED();
  return 0;
}
int bw_zero(float* p, long long n, cudaStream_t s) {
  PRD_CUDA_OK(cudaMemsetAsync(p, 0, (size_t)n * sizeof(float), s));
  return 0;
}

__global__ void bw_gate_fwd_kernel(const float* __restrict__ gpre, long long ldg, const float* __restrict__ o, long long ldo,
                                   float* __restrict__ og, long long ldog, long long R, int W) {
  const long long total = R * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / W;
    const int c = (int)(i - r * W);
    og[r * ldog + c] = round_tf32(sigmoid_acc(gpre[r * ldg + c]) * o[r * ldo + c]);
  }
}
// float4 versions of the gating kernels (rows of W floats, W % 4 == 0, fewer than 2^32 float4 groups): one 32-bit
// division per four elements instead of a 64-bit one per element, which made the scalar kernels instruction bound
__global__ void bw_gate_fwd_vec_kernel(const float* __restrict__ gpre, long long ldg, const float* __restrict__ o, long long ldo,
                                       float* __restrict__ og, long long ldog, unsigned total, unsigned W4) {
  for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
    const unsigned r = q / W4, c = (q - r * W4) * 4;
    const float4 gp = ld4(gpre + (long long)r * ldg + c), ov = ld4(o + (long long)r * ldo + c);
    st4(og + (long long)r * ldog + c, round4(make_float4(sigmoid_fast(gp.x) * ov.x, sigmoid_fast(gp.y) * ov.y,
                                                        sigmoid_fast(gp.z) * ov.z, sigmoid_fast(gp.w) * ov.w)));
  }
}
__global__ void bw_gate_bwd_vec_kernel(const float* __restrict__ d_og, long long ld1, const float* __restrict__ gpre, long long ldg,
                                       const float* __restrict__ o, long long ldo, float* __restrict__ d_o, long long ld2,
                                       float* __restrict__ d_gpre, long long ld3, unsigned total, unsigned W4) {
  for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
    const unsigned r = q / W4, c = (q - r * W4) * 4;
    const float4 gp = ld4(gpre + (long long)r * ldg + c), d = ld4(d_og + (long long)r * ld1 + c), ov = ld4(o + (long long)r * ldo + c);
    const float4 g = make_float4(sigmoid_fast(gp.x), sigmoid_fast(gp.y), sigmoid_fast(gp.z), sigmoid_fast(gp.w));
    st4(d_o + (long long)r * ld2 + c, round4(make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w)));
    st4(d_gpre + (long long)r * ld3 + c, round4(make_float4(d.x * ov.x * g.x * (1.f - g.x), d.y * ov.y * g.y * (1.f - g.y),
                                                           d.z * ov.z * g.z * (1.f - g.z), d.w * ov.w * g.w * (1.f - g.w))));
  }
}
static bool vec_rows_ok(long long R, int W, std::initializer_list<const void*> ptrs, std::initializer_list<long long> lds) {
  if (W % 4 != 0 || R * (W / 4) >= 4294967295LL || R >= 4294967295LL) return false;
  for (const void* p : ptrs)
    if (!aligned16(p)) return false;
  for (long long l : lds)
    if (l % 4 != 0) return false;
  return true;
}
int bw_gate_fwd(const float* gpre, long long ldg, const float* o, long long ldo, float* og, long long ldog, long long R, int W,
                cudaStream_t s) {
  if (vec_rows_ok(R, W, {gpre, o, og}, {ldg, ldo, ldog})) {
    bw_gate_fwd_vec_kernel<<<grid_for(R * (W / 4), 256), 256, 0, s>>>(gpre, ldg, o, ldo, og, ldog, (unsigned)(R * (W / 4)), (unsigned)(W / 4));
    PRD_LAUNCHED();
    return 0;
  }
  bw_gate_fwd_kernel<<<grid_for(R * W, 256), 256, 0, s>>>(gpre, ldg, o, ldo, og, ldog, R, W);
  PRD_LAUNCHED();
  return 0;
}
__global__ void bw_gate_bwd_kernel(const float* __restrict__ d_og, long long ld1, const float* __restrict__ gpre, long long ldg,
                                   const float* __restrict__ o, long long ldo, float* __restrict__ d_o, long long ld2,
                                   float* __restrict__ d_gpre, long long ld3, long long R, int W) {
  const long long total = R * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / W;
    const int c = (int)(i - r * W);
    const float g = sigmoid_acc(gpre[r * ldg + c]);
    const float d = d_og[r * ld1 + c];
    const float ov = o[r * ldo + c];
    d_o[r * ld2 + c] = round_tf32(d * g);
    d_gpre[r * ld3 + c] = round_tf32(d * ov * g * (1.f - g));
  }
}
int bw_gate_bwd(const float* d_og, long long ld1, const float* gpre, long long ldg, const float* o, long long ldo, float* d_o,
                long long ld2, float* d_gpre, long long ld3, long long R, int W, cudaStream_t s) {
  if (vec_rows_ok(R, W, {d_og, gpre, o, d_o, d_gpre}, {ld1, ldg, ldo, ld2, ld3})) {
    bw_gate_bwd_vec_kernel<<<grid_for(R * (W / 4), 256), 256, 0, s>>>(d_og, ld1, gpre, ldg, o, ldo, d_o, ld2, d_gpre, ld3,
                                                                      (unsigned)(R * (W / 4)), (unsigned)(W / 4));
    PRD_LAUNCHED();
    return 0;
  }
  bw_gate_bwd_kernel<<<grid_for(R * W, 256), 256, 0, s>>>(d_og, ld1, gpre, ldg, o, ldo, d_o, ld2, d_gpre, ld3, R, W);
  PRD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// attention with 16-channel heads (modules.py:185-225), exact fp32 on the FFMA pipe
// ------------------------------------------------------------------------------------------------------------
namespace {
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
}  // namespace

// grid (nseq * H, ceil(N / 128)), 128 threads = 128 queries; K / V of 128 keys at a time in shared memory
__global__ void __launch_bounds__(128) bw_attn_fwd_kernel(AttnDev a, const float* __restrict__ qkvg, long long ld,
                                                          float* __restrict__ O, float* __restrict__ lse) {
  __shared__ __align__(16) float sK[128][16];
  __shared__ __align__(16) float sV[128][16];
  __shared__ float sValid[128];
  const int h = blockIdx.x % a.H;
  const long long s = blockIdx.x / a.H;
  const long long b = attn_batch(a, s);
  const float ms = attn_seq_mask(a, s);
  const int qi = blockIdx.y * 128 + threadIdx.x;
  const bool qv = qi < a.N;
  float q[16], acc[16];
  float m = -INFINITY, l = 0.f;
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = 0.f;
  if (qv) {
    const float4* qp = reinterpret_cast<const float4*>(qkvg + attn_row(a, s, qi) * ld + h * 16);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float4 t = qp[c];
      q[4 * c] = t.x * a.scale; q[4 * c + 1] = t.y * a.scale; q[4 * c + 2] = t.z * a.scale; q[4 * c + 3] = t.w * a.scale;
    }
  }
  const float* brow = (a.bias && qv) ? a.bias + ((b * a.H + h) * a.N + qi) * (long long)a.N : nullptr;
  for (int k0 = 0; k0 < a.N; k0 += 128) {
    __syncthreads();
    const int kt = k0 + threadIdx.x;
    if (kt < a.N) {
      const float* kp = qkvg + attn_row(a, s, kt) * ld + h * 16;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        *reinterpret_cast<float4*>(&sK[threadIdx.x][4 * c]) = *reinterpret_cast<const float4*>(kp + 64 + 4 * c);
        *reinterpret_cast<float4*>(&sV[threadIdx.x][4 * c]) = *reinterpret_cast<const float4*>(kp + 128 + 4 * c);
      }
      sValid[threadIdx.x] = (ms * a.mask[b * a.N + kt] >= 0.5f) ? 1.f : 0.f;
    }
    __syncthreads();
    if (!qv) continue;
    const int kn = a.N - k0 < 128 ? a.N - k0 : 128;
    for (int kk = 0; kk < kn; kk += 8) {
      float sc[8];
      float cmax = -INFINITY;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (kk + u < kn) {
          float d = 0.f;
#pragma unroll
          for (int c = 0; c < 16; ++c) d = fmaf(q[c], sK[kk + u][c], d);
          if (brow) d += brow[k0 + kk + u];
          sc[u] = sValid[kk + u] > 0.5f ? d : kMaskFill;
          cmax = fmaxf(cmax, sc[u]);
        } else {
          sc[u] = -INFINITY;
        }
      }
      const float mnew = fmaxf(m, cmax);
      const float corr = __expf(m - mnew);  // m = -inf on the first group: exp(-inf) = 0
      l *= corr;
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[c] *= corr;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (kk + u < kn) {
          const float p = __expf(sc[u] - mnew);
          l += p;
#pragma unroll
          for (int c = 0; c < 16; ++c) acc[c] = fmaf(p, sV[kk + u][c], acc[c]);
        }
      }
      m = mnew;
    }
  }
  if (qv) {
    const float inv = 1.0f / l;
    float* op = O + attn_row(a, s, qi) * 64 + h * 16;
#pragma unroll
    for (int c = 0; c < 16; ++c) op[c] = acc[c] * inv;
    lse[(s * a.H + h) * a.N + qi] = m + logf(l);
  }
}

// thread = query: dq, D = <dO, O>, optional dbias
__global__ void __launch_bounds__(128) bw_attn_dq_kernel(AttnDev a, const float* __restrict__ qkvg, long long ld,
                                                         const float* __restrict__ O, const float* __restrict__ lse,
                                                         const float* __restrict__ dO, float* __restrict__ Dbuf,
                                                         float* __restrict__ dqkvg, long long ldd, float* __restrict__ dbias) {
  __shared__ __align__(16) float sK[128][16];
  __shared__ __align__(16) float sV[128][16];
  __shared__ float sValid[128];
  const int h = blockIdx.x % a.H;
  const long long s = blockIdx.x / a.H;
  const long long b = attn_batch(a, s);
  const float ms = attn_seq_mask(a, s);
  const int qi = blockIdx.y * 128 + threadIdx.x;
  const bool qv = qi < a.N;
  float q[16], dq[16], dOq[16];
  float D = 0.f, L = 0.f;
  long long row = 0;
#pragma unroll
  for (int c = 0; c < 16; ++c) dq[c] = 0.f;
  if (qv) {
    row = attn_row(a, s, qi);
    const float* qp = qkvg + row * ld + h * 16;
    const float* op = O + row * 64 + h * 16;
    const float* dp = dO + row * 64 + h * 16;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      q[c] = qp[c] * a.scale;
      dOq[c] = dp[c];
      D = fmaf(dOq[c], op[c], D);
    }
    L = lse[(s * a.H + h) * a.N + qi];
    Dbuf[(s * a.H + h) * a.N + qi] = D;
  }
  const long long boff = qv && (a.bias || dbias) ? ((b * a.H + h) * a.N + qi) * (long long)a.N : 0;
  for (int k0 = 0; k0 < a.N; k0 += 128) {
    __syncthreads();
    const int kt = k0 + threadIdx.x;
    if (kt < a.N) {
      const float* kp = qkvg + attn_row(a, s, kt) * ld + h * 16;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        *reinterpret_cast<float4*>(&sK[threadIdx.x][4 * c]) = *reinterpret_cast<const float4*>(kp + 64 + 4 * c);
        *reinterpret_cast<float4*>(&sV[threadIdx.x][4 * c]) = *reinterpret_cast<const float4*>(kp + 128 + 4 * c);
      }
      sValid[threadIdx.x] = (ms * a.mask[b * a.N + kt] >= 0.5f) ? 1.f : 0.f;
    }
    __syncthreads();
    if (!qv) continue;
    const int kn = a.N - k0 < 128 ? a.N - k0 : 128;
    for (int kk = 0; kk < kn; ++kk) {
      float d = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        d = fmaf(q[c], sK[kk][c], d);
        dp = fmaf(dOq[c], sV[kk][c], dp);
      }
      if (a.bias) d += a.bias[boff + k0 + kk];
      const bool valid = sValid[kk] > 0.5f;
      const float p = __expf((valid ? d : kMaskFill) - L);
      const float ds = valid ? p * (dp - D) : 0.f;  // masked_fill cuts the gradient (modules.py:220)
#pragma unroll
      for (int c = 0; c < 16; ++c) dq[c] = fmaf(ds, sK[kk][c], dq[c]);
      if (dbias) dbias[boff + k0 + kk] = ds;
    }
  }
  if (qv) {
    float* o = dqkvg + row * ldd + h * 16;
#pragma unroll
    for (int c = 0; c < 16; ++c) o[c] = round_tf32(dq[c] * a.scale);
  }
}

// thread = key: dk, dv; queries (q, dO, lse, D) of 128 at a time in shared memory
__global__ void __launch_bounds__(128) bw_attn_dkv_kernel(AttnDev a, const float* __restrict__ qkvg, long long ld,
                                                          const float* __restrict__ lse, const float* __restrict__ dO,
                                                          const float* __restrict__ Dbuf, float* __restrict__ dqkvg,
                                                          long long ldd) {
  __shared__ __align__(16) float sQ[128][16];
  __shared__ __align__(16) float sdO[128][16];
  __shared__ float sL[128], sD[128];
  const int h = blockIdx.x % a.H;
  const long long s = blockIdx.x / a.H;
  const long long b = attn_batch(a, s);
  const float ms = attn_seq_mask(a, s);
  const int ki = blockIdx.y * 128 + threadIdx.x;
  const bool kv = ki < a.N;
  float k[16], v[16], dk[16], dv[16];
  long long row = 0;
  bool valid = false;
#pragma unroll
  for (int c = 0; c < 16; ++c) dk[c] = dv[c] = 0.f;
  if (kv) {
    row = attn_row(a, s, ki);
    const float* kp = qkvg + row * ld + h * 16;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      k[c] = kp[64 + c];
      v[c] = kp[128 + c];
    }
    valid = ms * a.mask[b * a.N + ki] >= 0.5f;
  }
  for (int q0 = 0; q0 < a.N; q0 += 128) {
    __syncthreads();
    const int qt = q0 + threadIdx.x;
    if (qt < a.N) {
      const long long qr = attn_row(a, s, qt);
      const float* qp = qkvg + qr * ld + h * 16;
      const float* dp = dO + qr * 64 + h * 16;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float4 t = *reinterpret_cast<const float4*>(qp + 4 * c);
        t.x *= a.scale; t.y *= a.scale; t.z *= a.scale; t.w *= a.scale;
        *reinterpret_cast<float4*>(&sQ[threadIdx.x][4 * c]) = t;
        *reinterpret_cast<float4*>(&sdO[threadIdx.x][4 * c]) = *reinterpret_cast<const float4*>(dp + 4 * c);
      }
      sL[threadIdx.x] = lse[(s * a.H + h) * a.N + qt];
      sD[threadIdx.x] = Dbuf[(s * a.H + h) * a.N + qt];
    }
    __syncthreads();
    if (!kv) continue;
    const int qn = a.N - q0 < 128 ? a.N - q0 : 128;
    for (int qq = 0; qq < qn; ++qq) {
      float d = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        d = fmaf(sQ[qq][c], k[c], d);
        dp = fmaf(sdO[qq][c], v[c], dp);
      }
      if (a.bias) d += a.bias[((b * a.H + h) * a.N + q0 + qq) * (long long)a.N + ki];
      const float p = __expf((valid ? d : kMaskFill) - sL[qq]);
#pragma unroll
      for (int c = 0; c < 16; ++c) dv[c] = fmaf(p, sdO[qq][c], dv[c]);
      if (valid) {
        const float ds = p * (dp - sD[qq]);
#pragma unroll
        for (int c = 0; c < 16; ++c) dk[c] = fmaf(ds, sQ[qq][c], dk[c]);  // sQ carries the scale already
      }
    }
  }
  if (kv) {
    float* o = dqkvg + row * ldd + h * 16;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      o[64 + c] = round_tf32(dk[c]);
      o[128 + c] = round_tf32(dv[c]);
    }
  }
}

static AttnDev attn_dev(const AttnGeom& g) { return AttnDev{g.B, g.N, g.H, g.mode, g.mask, g.bias, g.scale}; }
static long long attn_nseq(const AttnGeom& g) { return g.mode == 2 ? g.B : (long long)g.B * g.N; }

int bw_attn_fwd(const AttnGeom& g, const float* qkvg, long long ld, float* O, float* lse, cudaStream_t s) {
  if (bw_attn_tc_enabled()) return bw_attn_tc_fwd(g, qkvg, ld, O, lse, s);
  PRD_REQUIRE(ld % 4 == 0, "attn: row stride must be a multiple of 4 floats");
  const long long nsh = attn_nseq(g) * g.H;
  PRD_REQUIRE(nsh < 2147483647LL, "attn: too many (sequence, head) pairs");
  bw_attn_fwd_kernel<<<dim3((unsigned)nsh, (g.N + 127) / 128), 128, 0, s>>>(attn_dev(g), qkvg, ld, O, lse);
  PRD_LAUNCHED();
  return 0;
}
int bw_attn_bwd(const AttnGeom& g, const float* qkvg, long long ld, const float* O, const float* lse, const float* dO,
                float* Dbuf, float* dqkvg, long long ldd, float* dbias, cudaStream_t s) {
  if (bw_attn_tc_enabled()) return bw_attn_tc_bwd(g, qkvg, ld, O, lse, dO, Dbuf, dqkvg, ldd, dbias, s);
  PRD_REQUIRE(ld % 4 == 0, "attn: row stride must be a multiple of 4 floats");
  const long long nsh = attn_nseq(g) * g.H;
  const dim3 grid((unsigned)nsh, (g.N + 127) / 128);
  bw_attn_dq_kernel<<<grid, 128, 0, s>>>(attn_dev(g), qkvg, ld, O, lse, dO, Dbuf, dqkvg, ldd, dbias);
  PRD_LAUNCHED();
  bw_attn_dkv_kernel<<<grid, 128, 0, s>>>(attn_dev(g), qkvg, ld, lse, dO, Dbuf, dqkvg, ldd);
  PRD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// pair-bias projection backward (modules.py:300-304, AF2_modules.py:454-459): warp per pair row, lane = 2 channels
// ------------------------------------------------------------------------------------------------------------
template <int CPL>  // channels per lane (CZ = 32 * CPL)
__global__ void __launch_bounds__(256) bw_pair_bias_bwd_kernel(int B, int N, int H, int ldb, const float* __restrict__ pair,
                                                               const float* __restrict__ dbias, const float* __restrict__ W,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               float* __restrict__ d_pair, float* __restrict__ dW,
                                                               float* __restrict__ dbvec, float* __restrict__ dgamma,
                                                               float* __restrict__ dbeta) {
  constexpr int CZ = 32 * CPL;
  const int lane = threadIdx.x & 31;
  const long long R = (long long)B * N * N;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  float w[4][CPL], aW[4][CPL], aG[CPL], aB[CPL], ab[4] = {0.f, 0.f, 0.f, 0.f};
  float gm[CPL], bt[CPL];
#pragma unroll
  for (int e = 0; e < CPL; ++e) {
    const int c = lane + 32 * e;
    gm[e] = gamma ? gamma[c] : 1.f;
    bt[e] = beta ? beta[c] : 0.f;
    aG[e] = aB[e] = 0.f;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      w[h][e] = h < H ? W[h * CZ + c] : 0.f;
      aW[h][e] = 0.f;
    }
  }
  const long long NN = (long long)N * N;
  for (long long r = warp0; r < R; r += nw) {
    const long long b = r / NN, ij = r - b * NN;
    float x[CPL], s = 0.f;
#pragma unroll
    for (int e = 0; e < CPL; ++e) {
      x[e] = pair[r * CZ + lane + 32 * e];
      s += x[e];
    }
    const float mean = warp_sum(s) / CZ;
    float v = 0.f;
#pragma unroll
    for (int e = 0; e < CPL; ++e) v += (x[e] - mean) * (x[e] - mean);
    const float rstd = rsqrtf(warp_sum(v) / CZ + kLnEps);
    float db[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) db[h] = h < H ? dbias[((b * H + h) * N + ij / N) * (long long)ldb + ij % N] : 0.f;
    float s1 = 0.f, s2 = 0.f, gg[CPL], xh[CPL];
#pragma unroll
    for (int e = 0; e < CPL; ++e) {
      xh[e] = (x[e] - mean) * rstd;
      const float y = xh[e] * gm[e] + bt[e];
      float dy = 0.f;
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        dy = fmaf(db[h], w[h][e], dy);
        aW[h][e] = fmaf(db[h], y, aW[h][e]);
      }
      aG[e] = fmaf(dy, xh[e], aG[e]);
      aB[e] += dy;
      gg[e] = dy * gm[e];
      s1 += gg[e];
      s2 += gg[e] * xh[e];
    }
#pragma unroll
    for (int h = 0; h < 4; ++h) ab[h] += db[h];
    const float m1 = warp_sum(s1) / CZ, m2 = warp_sum(s2) / CZ;
#pragma unroll
    for (int e = 0; e < CPL; ++e) {
      const long long o = r * CZ + lane + 32 * e;
      d_pair[o] = round_tf32(d_pair[o] + rstd * (gg[e] - m1 - xh[e] * m2));
    }
  }
#pragma unroll
  for (int e = 0; e < CPL; ++e) {
    const int c = lane + 32 * e;
#pragma unroll
    for (int h = 0; h < 4; ++h)
      if (h < H) atomicAdd(dW + h * CZ + c, aW[h][e]);
    if (dgamma) atomicAdd(dgamma + c, aG[e]);
    if (dbeta) atomicAdd(dbeta + c, aB[e]);
  }
  if (dbvec && lane == 0) {
#pragma unroll
    for (int h = 0; h < 4; ++h)
      if (h < H) atomicAdd(dbvec + h, ab[h]);
  }
}
int bw_pair_bias_bwd(int B, int N, int CZ, int H, const float* pair, const float* dbias, int ldb, const float* W, const float* gamma,
                     const float* beta, float* d_pair, float* dW, float* dbvec, float* dgamma, float* dbeta, cudaStream_t s) {
  PRD_REQUIRE(H <= 4 && (CZ == 64 || CZ == 32), "pair_bias_bwd: built for <= 4 heads and c_z 32 / 64");
  const unsigned grid = grid_for((long long)B * N * N, 8, 148 * 8);
  if (CZ == 64)
    bw_pair_bias_bwd_kernel<2><<<grid, 256, 0, s>>>(B, N, H, ldb, pair, dbias, W, gamma, beta, d_pair, dW, dbvec, dgamma, dbeta);
  else
    bw_pair_bias_bwd_kernel<1><<<grid, 256, 0, s>>>(B, N, H, ldb, pair, dbias, W, gamma, beta, d_pair, dW, dbvec, dgamma, dbeta);
  PRD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// rows <-> channel planes: per (b, i) a [N tokens, C channels] <-> [C, N] transpose through shared memory
// ------------------------------------------------------------------------------------------------------------
// transposed: planes[(b, c)][j][i] instead of [i][j] -- the block then walks i for a fixed j (rows N * ld apart, still one
// full 128-byte line per row), so the transposed planes need no second pass over the natural ones
__global__ void bw_rows_to_planes_kernel(const float* __restrict__ rows, long long ld, int col0, int N, int C, int Np,
                                         float* __restrict__ planes, int transposed) {
  __shared__ float tile[32][33];
  const long long bi = blockIdx.z;  // b * N + i  (transposed: b * N + j)
  const long long b = bi / N, i = bi - b * N;
  const int j0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int t = threadIdx.y; t < 32; t += 8) {
    const int j = j0 + t, c = c0 + threadIdx.x;
    const long long row = transposed ? (b * N + j) * N + i : bi * N + j;
    tile[t][threadIdx.x] = (j < N && c < C) ? rows[row * ld + col0 + c] : 0.f;
  }
  __syncthreads();
  for (int t = threadIdx.y; t < 32; t += 8) {
    const int c = c0 + t, j = j0 + threadIdx.x;
    if (c < C && j < N) planes[((b * C + c) * N + i) * (long long)Np + j] = round_tf32(tile[threadIdx.x][t]);
  }
}
__global__ void bw_planes_to_rows_kernel(const float* __restrict__ planes, int N, int C, int Np, float* __restrict__ rows,
                                         long long ld, int col0) {
  __shared__ float tile[32][33];
  const long long bi = blockIdx.z;
  const long long b = bi / N, i = bi - b * N;
  const int j0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int t = threadIdx.y; t < 32; t += 8) {
    const int c = c0 + t, j = j0 + threadIdx.x;
    tile[t][threadIdx.x] = (c < C && j < N) ? planes[((b * C + c) * N + i) * (long long)Np + j] : 0.f;
  }
  __syncthreads();
  for (int t = threadIdx.y; t < 32; t += 8) {
    const int j = j0 + t, c = c0 + threadIdx.x;
    if (j < N && c < C) rows[(bi * N + j) * ld + col0 + c] = round_tf32(tile[threadIdx.x][t]);
  }
}
// 64 tokens x 64 channels per block: float4 along the channels on the row side (a row's 64 channels are one or two full
// lines), 128-byte runs along the tokens on the plane side.  The 32 x 32 scalar tiles above moved 2.5 TB/s.
__global__ void __launch_bounds__(256) bw_rows_to_planes64_kernel(const float* __restrict__ rows, long long ld, int col0, int N, int C,
                                                                  int Np, float* __restrict__ planes, int transposed) {
  __shared__ float tile[64][65];
  const long long bi = blockIdx.z;
  const long long b = bi / N, i = bi - b * N;
  const int j0 = blockIdx.x * 64, c0 = blockIdx.y * 64, tid = threadIdx.x;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int idx = tid + 256 * k, jr = idx >> 4, c = 4 * (idx & 15);
    const int j = j0 + jr;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < N && c0 + c < C) {
      const long long row = transposed ? (b * N + j) * N + i : bi * N + j;
      v = ld4(rows + row * ld + col0 + c0 + c);
    }
    tile[jr][c] = v.x; tile[jr][c + 1] = v.y; tile[jr][c + 2] = v.z; tile[jr][c + 3] = v.w;
  }
  __syncthreads();
#pragma unroll 4
  for (int k = 0; k < 16; ++k) {
    const int idx = tid + 256 * k, c = idx >> 6, jj = idx & 63;
    if (c0 + c < C && j0 + jj < N) planes[((b * C + c0 + c) * N + i) * (long long)Np + j0 + jj] = round_tf32(tile[jj][c]);
  }
}
__global__ void __launch_bounds__(256) bw_planes_to_rows64_kernel(const float* __restrict__ planes, int N, int C, int Np,
                                                                  float* __restrict__ rows, long long ld, int col0) {
  __shared__ float tile[64][65];
  const long long bi = blockIdx.z;
  const long long b = bi / N, i = bi - b * N;
  const int j0 = blockIdx.x * 64, c0 = blockIdx.y * 64, tid = threadIdx.x;
#pragma unroll 4
  for (int k = 0; k < 16; ++k) {
    const int idx = tid + 256 * k, c = idx >> 6, jj = idx & 63;
    tile[jj][c] = (c0 + c < C && j0 + jj < N) ? planes[((b * C + c0 + c) * N + i) * (long long)Np + j0 + jj] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int idx = tid + 256 * k, jr = idx >> 4, c = 4 * (idx & 15);
    const int j = j0 + jr;
    if (j < N && c0 + c < C)
      st4(rows + (bi * N + j) * ld + col0 + c0 + c,
          round4(make_float4(tile[jr][c], tile[jr][c + 1], tile[jr][c + 2], tile[jr][c + 3])));
  }
}
int bw_rows_to_planes(const float* rows, long long ld, int col0, int B, int N, int C, int Np, float* planes, cudaStream_t s,
                      int transposed) {
  PRD_REQUIRE((long long)B * N <= 65535, "rows_to_planes: B*N = %lld exceeds the grid limit", (long long)B * N);
  if (C % 4 == 0 && col0 % 4 == 0 && ld % 4 == 0 && aligned16(rows)) {
    bw_rows_to_planes64_kernel<<<dim3((N + 63) / 64, (C + 63) / 64, B * N), 256, 0, s>>>(rows, ld, col0, N, C, Np, planes, transposed);
    PRD_LAUNCHED();
    return 0;
  }
  bw_rows_to_planes_kernel<<<dim3((N + 31) / 32, (C + 31) / 32, B * N), dim3(32, 8), 0, s>>>(rows, ld, col0, N, C, Np, planes, transposed);
  PRD_LAUNCHED();
  return 0;
}
int bw_planes_to_rows(const float* planes, int B, int N, int C, int Np, float* rows, long long ld, int col0, cudaStream_t s) {
  PRD_REQUIRE((long long)B * N <= 65535, "planes_to_rows: B*N = %lld exceeds the grid limit", (long long)B * N);
  if (C % 4 == 0 && col0 % 4 == 0 && ld % 4 == 0 && aligned16(rows)) {
    bw_planes_to_rows64_kernel<<<dim3((N + 63) / 64, (C + 63) / 64, B * N), 256, 0, s>>>(planes, N, C, Np, rows, ld, col0);
    PRD_LAUNCHED();
    return 0;
  }
  bw_planes_to_rows_kernel<<<dim3((N + 31) / 32, (C + 31) / 32, B * N), dim3(32, 8), 0, s>>>(planes, N, C, Np, rows, ld, col0);
  PRD_LAUNCHED();
  return 0;
}

__global__ void bw_trimul_ab_kernel(const float* __restrict__ pre, long long ld, const float* __restrict__ mask, int N, int C2,
                                    long long R, float* __restrict__ ab) {
  const long long total = R * C2;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / C2;
    const int c = (int)(idx - r * C2);
    const long long bi = r / N;            // b * N + i
    const long long b = bi / N;
    const int j = (int)(r - bi * N);
    const float m2 = mask[bi] * mask[b * N + j];
    ab[idx] = round_tf32(m2 * sigmoid_acc(pre[r * ld + C2 + c]) * pre[r * ld + c]);
  }
}
__global__ void bw_trimul_ab_vec_kernel(const float* __restrict__ pre, long long ld, const float* __restrict__ mask, unsigned N,
                                        unsigned C2, unsigned total, float* __restrict__ ab) {
  const unsigned W4 = C2 / 4;
  for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
    const unsigned r = q / W4, c = (q - r * W4) * 4;
    const unsigned bi = r / N, j = r - bi * N, b = bi / N;
    const float m2 = mask[bi] * mask[b * N + j];
    const float4 p = ld4(pre + (long long)r * ld + c), gp = ld4(pre + (long long)r * ld + C2 + c);
    st4(ab + (long long)r * C2 + c, round4(make_float4(m2 * sigmoid_fast(gp.x) * p.x, m2 * sigmoid_fast(gp.y) * p.y,
                                                      m2 * sigmoid_fast(gp.z) * p.z, m2 * sigmoid_fast(gp.w) * p.w)));
  }
}
__global__ void bw_trimul_ab_bwd_vec_kernel(const float* __restrict__ pre, long long ld, const float* __restrict__ mask, unsigned N,
                                            unsigned C2, unsigned total, const float* __restrict__ dab, float* __restrict__ dpre,
                                            long long ldd) {
  const unsigned W4 = C2 / 4;
  for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
    const unsigned r = q / W4, c = (q - r * W4) * 4;
    const unsigned bi = r / N, j = r - bi * N, b = bi / N;
    const float m2 = mask[bi] * mask[b * N + j];
    const float4 p = ld4(pre + (long long)r * ld + c), gp = ld4(pre + (long long)r * ld + C2 + c);
    const float4 dd = ld4(dab + (long long)r * C2 + c);
    const float4 g = make_float4(sigmoid_fast(gp.x), sigmoid_fast(gp.y), sigmoid_fast(gp.z), sigmoid_fast(gp.w));
    const float4 d = make_float4(dd.x * m2, dd.y * m2, dd.z * m2, dd.w * m2);
    st4(dpre + (long long)r * ldd + c, round4(make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w)));
    st4(dpre + (long long)r * ldd + C2 + c, round4(make_float4(d.x * p.x * g.x * (1.f - g.x), d.y * p.y * g.y * (1.f - g.y),
                                                               d.z * p.z * g.z * (1.f - g.z), d.w * p.w * g.w * (1.f - g.w))));
  }
}
int bw_trimul_ab(const float* pre, long long ld, const float* mask, int B, int N, int C2, float* ab, cudaStream_t s) {
  const long long R = (long long)B * N * N;
  if (vec_rows_ok(R, C2, {pre, ab}, {ld})) {
    bw_trimul_ab_vec_kernel<<<grid_for(R * (C2 / 4), 256), 256, 0, s>>>(pre, ld, mask, (unsigned)N, (unsigned)C2, (unsigned)(R * (C2 / 4)), ab);
    PRD_LAUNCHED();
    return 0;
  }
  bw_trimul_ab_kernel<<<grid_for(R * C2, 256), 256, 0, s>>>(pre, ld, mask, N, C2, R, ab);
  PRD_LAUNCHED();
  return 0;
}
__global__ void bw_trimul_ab_bwd_kernel(const float* __restrict__ pre, long long ld, const float* __restrict__ mask, int N,
                                        int C2, long long R, const float* __restrict__ dab, float* __restrict__ dpre,
                                        long long ldd) {
  const long long total = R * C2;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / C2;
    const int c = (int)(idx - r * C2);
    const long long bi = r / N;
    const long long b = bi / N;
    const int j = (int)(r - bi * N);
    const float m2 = mask[bi] * mask[b * N + j];
    const float g = sigmoid_acc(pre[r * ld + C2 + c]);
    const float p = pre[r * ld + c];
    const float d = dab[idx] * m2;
    dpre[r * ldd + c] = round_tf32(d * g);
    dpre[r * ldd + C2 + c] = round_tf32(d * p * g * (1.f - g));
  }
}
int bw_trimul_ab_bwd(const float* pre, long long ld, const float* mask, int B, int N, int C2, const float* dab, float* dpre,
                     long long ldd, cudaStream_t s) {
  const long long R = (long long)B * N * N;
  if (vec_rows_ok(R, C2, {pre, dab, dpre}, {ld, ldd})) {
    bw_trimul_ab_bwd_vec_kernel<<<grid_for(R * (C2 / 4), 256), 256, 0, s>>>(pre, ld, mask, (unsigned)N, (unsigned)C2,
                                                                            (unsigned)(R * (C2 / 4)), dab, dpre, ldd);
    PRD_LAUNCHED();
    return 0;
  }
  bw_trimul_ab_bwd_kernel<<<grid_for(R * C2, 256), 256, 0, s>>>(pre, ld, mask, N, C2, R, dab, dpre, ldd);
  PRD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// materialised softmax (SPAttention: head width c_s, no mask), warp per row
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bw_softmax_rows_kernel(const float* __restrict__ logits, float* __restrict__ probs,
                                                              long long rows, int n, int ld) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* lp = logits + r * ld;
  float m = -INFINITY;
  for (int c = lane; c < n; c += 32) m = fmaxf(m, lp[c]);
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane; c < n; c += 32) s += expf(lp[c] - m);
  const float inv = 1.0f / warp_sum(s);
  for (int c = lane; c < n; c += 32) probs[r * ld + c] = round_tf32(expf(lp[c] - m) * inv);
}
int bw_softmax_rows(const float* logits, float* probs, long long rows, int n, int ld, cudaStream_t s) {
  bw_softmax_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(logits, probs, rows, n, ld);
  PRD_LAUNCHED();
  return 0;
}
__global__ void __launch_bounds__(256) bw_softmax_bwd_rows_kernel(const float* __restrict__ probs, const float* __restrict__ dprobs,
                                                                  float* __restrict__ dlogits, long long rows, int n, int ld) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  float s = 0.f;
  for (int c = lane; c < n; c += 32) s = fmaf(probs[r * ld + c], dprobs[r * ld + c], s);
  s = warp_sum(s);
  for (int c = lane; c < n; c += 32) dlogits[r * ld + c] = round_tf32(probs[r * ld + c] * (dprobs[r * ld + c] - s));
}
int bw_softmax_bwd_rows(const float* probs, const float* dprobs, float* dlogits, long long rows, int n, int ld, cudaStream_t s) {
  bw_softmax_bwd_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(probs, dprobs, dlogits, rows, n, ld);
  PRD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// bilinear pair terms (OuterLinear product term modules.py:283-287, OuterProductUpdate AF2_modules.py:532-537)
// ------------------------------------------------------------------------------------------------------------
// ET[b, n, z, m] = E[b, n, m, z] (transpose_ij = 0) or E[b, m, n, z] (transpose_ij = 1), plus the other one if add_transposed
__global__ void bw_pair_to_izj_kernel(const float* __restrict__ E, int N, int CZ, int Np, int transpose_ij, int add_transposed,
                                      float* __restrict__ ET) {
  __shared__ float tile[32][33];
  const long long bn = blockIdx.z;
  const long long b = bn / N, n = bn - b * N;
  const int m0 = blockIdx.x * 32, z0 = blockIdx.y * 32;
  const float* Eb = E + b * (long long)N * N * CZ;
  for (int t = threadIdx.y; t < 32; t += 8) {
    const int m = m0 + t, z = z0 + threadIdx.x;
    float v = 0.f;
    if (m < N && z < CZ) {
      const float direct = Eb[(n * N + m) * CZ + z], swapped = Eb[((long long)m * N + n) * CZ + z];
      v = transpose_ij ? swapped : direct;
      if (add_transposed) v = direct + swapped;
    }
    tile[t][threadIdx.x] = v;
  }
  __syncthreads();
  for (int t = threadIdx.y; t < 32; t += 8) {
    const int z = z0 + t, m = m0 + threadIdx.x;
    if (z < CZ && m < N) ET[(bn * CZ + z) * (long long)Np + m] = round_tf32(tile[threadIdx.x][t]);
  }
}
int bw_pair_to_izj(const float* E, int B, int N, int CZ, int Np, int transpose_ij, int add_transposed, float* ET, cudaStream_t s) {
  PRD_REQUIRE((long long)B * N <= 65535, "pair_to_izj: B*N too large");
  bw_pair_to_izj_kernel<<<dim3((N + 31) / 32, (CZ + 31) / 32, B * N), dim3(32, 8), 0, s>>>(E, N, CZ, Np, transpose_ij, add_transposed, ET);
  PRD_LAUNCHED();
  return 0;
}

// block per (b, n) slab of T [CZ, C]; thread = channel c
__global__ void bw_bilinear_reduce_kernel(const float* __restrict__ T, const float* __restrict__ W, long long ldw,
                                          const float* __restrict__ self, int CZ, int C, float* __restrict__ d_self,
                                          int accumulate, float* __restrict__ dW, int BN, float alpha, float alpha_dw) {
  // dW needs a reduction over (b, n): each block walks a strided set of slabs and keeps dW partials in registers
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float w[64], aw[64];
    for (int z = 0; z < CZ; ++z) {
      w[z] = W[z * ldw + c];
      aw[z] = 0.f;
    }
    for (int bn = blockIdx.x; bn < BN; bn += gridDim.x) {
      const float* Tp = T + (long long)bn * CZ * C + c;
      const float sv = self ? self[(long long)bn * C + c] : 0.f;
      float acc = 0.f;
      for (int z = 0; z < CZ; ++z) {
        const float t = Tp[(long long)z * C];
        acc = fmaf(w[z], t, acc);
        aw[z] = fmaf(sv, t, aw[z]);
      }
      const long long o = (long long)bn * C + c;
      d_self[o] = round_tf32(accumulate ? d_self[o] + alpha * acc : alpha * acc);
    }
    if (dW)
      for (int z = 0; z < CZ; ++z) atomicAdd(dW + z * ldw + c, alpha_dw * aw[z]);
  }
}
int bw_bilinear_reduce(const float* T, const float* W, long long ldw, const float* self, int B, int N, int CZ, int C,
                       float* d_self, int accumulate, float* dW, float alpha, float alpha_dw, cudaStream_t s) {
  PRD_REQUIRE(CZ <= 64, "bilinear_reduce: c_z <= 64");
  const int BN = B * N;
  const int threads = C < 256 ? ((C + 31) / 32 * 32) : 256;
  bw_bilinear_reduce_kernel<<<BN < 148 * 2 ? BN : 148 * 2, threads, 0, s>>>(T, W, ldw, self, CZ, C, d_self, accumulate, dW, BN, alpha, alpha_dw);
  PRD_LAUNCHED();
  return 0;
}

// U[b,n,z] = sum_j E[b,n,j,z] - sum_i E[b,i,n,z]
__global__ void bw_pair_rowcol_diff_kernel(const float* __restrict__ E, int N, int CZ, float* __restrict__ U) {
  const long long bn = blockIdx.x;
  const long long b = bn / N, n = bn - b * N;
  const float* Eb = E + b * (long long)N * N * CZ;
  for (int z = threadIdx.x; z < CZ; z += blockDim.x) {
    float acc = 0.f;
    for (int m = 0; m < N; ++m) acc += Eb[(n * N + m) * CZ + z] - Eb[((long long)m * N + n) * CZ + z];
    U[bn * CZ + z] = round_tf32(acc);
  }
}
int bw_pair_rowcol_diff(const float* E, int B, int N, int CZ, float* U, cudaStream_t s) {
  bw_pair_rowcol_diff_kernel<<<B * N, 64, 0, s>>>(E, N, CZ, U);
  PRD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// elementwise helpers
// ------------------------------------------------------------------------------------------------------------
__global__ void bw_scale_rows_kernel(const float* __restrict__ a, long long lda, const float* __restrict__ rowscale, float alpha,
                                     float* __restrict__ out, long long ldo, long long R, int W) {
  const long long total = R * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / W;
    const int c = (int)(i - r * W);
    out[r * ldo + c] = round_tf32(alpha * a[r * lda + c] * (rowscale ? rowscale[r] : 1.f));
  }
}
int bw_scale_rows(const float* a, long long lda, const float* rowscale, float alpha, float* out, long long ldo, long long R,
                  int W, cudaStream_t s) {
  bw_scale_rows_kernel<<<grid_for(R * W, 256), 256, 0, s>>>(a, lda, rowscale, alpha, out, ldo, R, W);
  PRD_LAUNCHED();
  return 0;
}
__global__ void bw_mask_pair_kernel(const float* __restrict__ E, const float* __restrict__ mask, int N, int CZ, long long R,
                                    float alpha, float* __restrict__ out) {
  const long long total = R * CZ;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / CZ;
    const long long bi = r / N;
    const long long b = bi / N;
    const int j = (int)(r - bi * N);
    out[idx] = round_tf32(alpha * mask[bi] * mask[b * N + j] * E[idx]);
  }
}
int bw_mask_pair(const float* E, const float* mask, int B, int N, int CZ, float alpha, float* out, cudaStream_t s) {
  const long long R = (long long)B * N * N;
  bw_mask_pair_kernel<<<grid_for(R * CZ, 256), 256, 0, s>>>(E, mask, N, CZ, R, alpha, out);
  PRD_LAUNCHED();
  return 0;
}
__global__ void bw_add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = round_tf32(a[i] + b[i]);
}
int bw_add(const float* a, const float* b, float* dst, long long n, cudaStream_t s) {
  bw_add_kernel<<<grid_for(n, 256), 256, 0, s>>>(a, b, dst, n);
  PRD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// coordinate head (model.py:364-373, utils.py:32-36)
// ------------------------------------------------------------------------------------------------------------
__global__ void bw_remove_mean_adj_kernel(const float* __restrict__ d_out, const float* __restrict__ mask, int N,
                                          float* __restrict__ d_eps) {
  __shared__ float red[4][128];
  const int b = blockIdx.x;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float m = mask[b * N + i];
    s[0] += m * d_out[(b * (long long)N + i) * 3 + 0];
    s[1] += m * d_out[(b * (long long)N + i) * 3 + 1];
    s[2] += m * d_out[(b * (long long)N + i) * 3 + 2];
    s[3] += m;
  }
  for (int k = 0; k < 4; ++k) red[k][threadIdx.x] = s[k];
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int k = 0; k < 4; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + o];
    __syncthreads();
  }
  const float n = red[3][0];
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float m = mask[b * N + i];
    for (int k = 0; k < 3; ++k) {
      const long long o = (b * (long long)N + i) * 3 + k;
      d_eps[o] = d_out[o] - m * red[k][0] / n;
    }
  }
}
int bw_remove_mean_adj(const float* d_out, const float* mask, int B, int N, float* d_eps, cudaStream_t s) {
  bw_remove_mean_adj_kernel<<<B, 128, 0, s>>>(d_out, mask, N, d_eps);
  PRD_LAUNCHED();
  return 0;
}

template <int CPL>
__global__ void __launch_bounds__(256) bw_coord_dh_kernel(const float* __restrict__ h, const float* __restrict__ z,
                                                          const float* __restrict__ mask, const float* __restrict__ d_eps,
                                                          const float* __restrict__ w2, int B, int N, float* __restrict__ dh,
                                                          float* __restrict__ dw2) {
  constexpr int CZ = 32 * CPL;
  const int lane = threadIdx.x & 31;
  const long long R = (long long)B * N * N;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  float w[CPL], aw[CPL];
#pragma unroll
  for (int e = 0; e < CPL; ++e) {
    w[e] = w2[lane + 32 * e];
    aw[e] = 0.f;
  }
  for (long long r = warp0; r < R; r += nw) {
    const long long bi = r / N;
    const long long b = bi / N;
    const int j = (int)(r - bi * N);
    const long long bj = b * N + j;
    const float m2 = mask[bi] * mask[bj];
    const float dx = z[bi * 3] - z[bj * 3], dy = z[bi * 3 + 1] - z[bj * 3 + 1], dz = z[bi * 3 + 2] - z[bj * 3 + 2];
    const float inv = rsqrtf(dx * dx + dy * dy + dz * dz + 1e-4f);
    const float dw = m2 * inv * (d_eps[bi * 3] * dx + d_eps[bi * 3 + 1] * dy + d_eps[bi * 3 + 2] * dz);
#pragma unroll
    for (int e = 0; e < CPL; ++e) {
      const long long o = r * CZ + lane + 32 * e;
      const float hv = h[o];
      dh[o] = round_tf32(hv > 0.f ? dw * w[e] : 0.f);
      aw[e] = fmaf(dw, hv, aw[e]);
    }
  }
#pragma unroll
  for (int e = 0; e < CPL; ++e) atomicAdd(dw2 + lane + 32 * e, aw[e]);
}
int bw_coord_dh(const float* h, const float* z, const float* mask, const float* d_eps, const float* w2, int B, int N, int CZ,
                float* dh, float* dw2, cudaStream_t s) {
  PRD_REQUIRE(CZ == 64 || CZ == 32, "coord_dh: c_z 32 / 64");
  const unsigned grid = grid_for((long long)B * N * N, 8, 148 * 8);
  if (CZ == 64) bw_coord_dh_kernel<2><<<grid, 256, 0, s>>>(h, z, mask, d_eps, w2, B, N, dh, dw2);
  else bw_coord_dh_kernel<1><<<grid, 256, 0, s>>>(h, z, mask, d_eps, w2, B, N, dh, dw2);
  PRD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// embeddings (model.py:332-361)
// ------------------------------------------------------------------------------------------------------------
__global__ void bw_rbf_rows_kernel(const float* __restrict__ z, const float* __restrict__ centers, float scale, int N, int DD,
                                   long long R, float* __restrict__ rbf) {
  const long long total = R * DD;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / DD;
    const int k = (int)(idx - r * DD);
    const long long bi = r / N;
    const long long b = bi / N;
    const long long bj = b * N + (r - bi * N);
    const float dx = z[bi * 3] - z[bj * 3], dy = z[bi * 3 + 1] - z[bj * 3 + 1], dz = z[bi * 3 + 2] - z[bj * 3 + 2];
    const float d = sqrtf(dx * dx + dy * dy + dz * dz) - centers[k];
    rbf[idx] = round_tf32(expf(-scale * d * d));
  }
}
int bw_rbf_rows(const float* z, const float* centers, float scale, int B, int N, int DD, float* rbf, cudaStream_t s) {
  const long long R = (long long)B * N * N;
  bw_rbf_rows_kernel<<<grid_for(R * DD, 256), 256, 0, s>>>(z, centers, scale, N, DD, R, rbf);
  PRD_LAUNCHED();
  return 0;
}

// colsum[b, z] = sum_ij Dm[b,i,j,z], then dW_beta[z, k] += sum_b colsum[b, z] sincos_k(t_b / T)
__global__ void bw_pair_colsum_kernel(const float* __restrict__ Dm, long long NN, int CZ, float* __restrict__ colsum) {
  // grid (B, chunks); block 256: thread = (row lane, channel)
  const int b = blockIdx.x;
  const int z = threadIdx.x % CZ, rl = threadIdx.x / CZ, rstep = blockDim.x / CZ;
  float acc = 0.f;
  for (long long r = (long long)blockIdx.y * rstep + rl; r < NN; r += (long long)gridDim.y * rstep) acc += Dm[(b * NN + r) * CZ + z];
  atomicAdd(colsum + b * CZ + z, acc);
}
__global__ void bw_time_embed_bwd_kernel(const float* __restrict__ colsum, const int64_t* __restrict__ t, int num_steps,
                                         const float* __restrict__ freq, int B, int CZ, int TD, float* __restrict__ dW_beta) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= CZ * TD) return;
  const int zc = idx / TD, k = idx - zc * TD, half = TD / 2;
  float acc = 0.f;
  for (int b = 0; b < B; ++b) {
    long long tb = t[b];
    const float scaled = static_cast<float>(tb) / static_cast<float>(num_steps);
    const float wx = freq[k < half ? k : k - half] * scaled;
    acc += colsum[b * CZ + zc] * (k < half ? sinf(wx) : cosf(wx));
  }
  atomicAdd(dW_beta + idx, acc);
}
int bw_time_embed_bwd(const float* Dm, const int64_t* t, int num_steps, const float* freq, int B, int N, int CZ, int TD,
                      float* colsum_scratch, float* dW_beta, cudaStream_t s) {
  PRD_REQUIRE(256 % CZ == 0, "time_embed_bwd: c_z must divide 256");
  if (bw_zero(colsum_scratch, (long long)B * CZ, s)) return 1;
  bw_pair_colsum_kernel<<<dim3(B, 148), 256, 0, s>>>(Dm, (long long)N * N, CZ, colsum_scratch);
  PRD_LAUNCHED();
  bw_time_embed_bwd_kernel<<<(CZ * TD + 255) / 256, 256, 0, s>>>(colsum_scratch, t, num_steps, freq, B, CZ, TD, dW_beta);
  PRD_LAUNCHED();
  return 0;
}

// scatter into the five small tables through a shared-memory copy per CTA (86 rows x c_z)
__global__ void __launch_bounds__(256) bw_pair_static_bwd_kernel(
    const float* __restrict__ d_pair, const float* __restrict__ atom_mask, const float* __restrict__ residue_mask,
    const float* __restrict__ bond_mask, const int64_t* __restrict__ bond_feats, const int64_t* __restrict__ bond_distance,
    const int64_t* __restrict__ residue_index, const int64_t* __restrict__ chain_index, int N, int CZ, long long R, int max_bd,
    int max_rel, float* __restrict__ d_bond0, float* __restrict__ d_bond1, float* __restrict__ d_bond2,
    float* __restrict__ d_bdist, float* __restrict__ d_relpos) {
  extern __shared__ float tab[];  // rows: bond0 5 | bond1 6 | bond2 2 | bdist (max_bd + 1) | relpos (2 max_rel + 1)
  const int rows_bd = max_bd + 1, rows_rel = 2 * max_rel + 1;
  const int off1 = 5, off2 = 11, offd = 13, offr = 13 + rows_bd, total_rows = offr + rows_rel;
  for (int i = threadIdx.x; i < total_rows * CZ; i += blockDim.x) tab[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const float scale = 0.57735026918962584f;
  for (long long r = (long long)blockIdx.x * nwarp + warp; r < R; r += (long long)gridDim.x * nwarp) {
    const long long bi = r / N;
    const long long b = bi / N;
    const long long bj = b * N + (r - bi * N);
    const float am2 = atom_mask[bi] * atom_mask[bj];
    const float rm2 = residue_mask[bi] * residue_mask[bj];
    if (am2 == 0.f && rm2 == 0.f) continue;
    const float same = chain_index[bi] == chain_index[bj] ? 1.f : 0.f;
    int i0 = 0, i1 = 0, i2 = 0, ibd = 0, irel = 0;
    float bm = 0.f;
    if (am2 != 0.f) {
      constexpr int kVocab[3] = {5, 6, 2};
      long long f0 = bond_feats[r * 3], f1 = bond_feats[r * 3 + 1], f2 = bond_feats[r * 3 + 2];
      i0 = (int)(f0 < 0 ? 0 : (f0 >= kVocab[0] ? kVocab[0] - 1 : f0));
      i1 = (int)(f1 < 0 ? 0 : (f1 >= kVocab[1] ? kVocab[1] - 1 : f1));
      i2 = (int)(f2 < 0 ? 0 : (f2 >= kVocab[2] ? kVocab[2] - 1 : f2));
      long long bd = bond_distance[r];
      ibd = (int)(bd > max_bd ? max_bd : (bd < 0 ? 0 : bd));
      bm = bond_mask[r];
    }
    if (rm2 != 0.f) {
      long long rel = residue_index[bi] - residue_index[bj];
      rel = rel < -max_rel ? -max_rel : (rel > max_rel ? max_rel : rel);
      irel = max_rel + (int)rel;
    }
    for (int c = lane; c < CZ; c += 32) {
      const float d = d_pair[r * CZ + c];
      if (am2 != 0.f) {
        const float db = am2 * bm * scale * d;
        if (db != 0.f) {
          atomicAdd(&tab[i0 * CZ + c], db);
          atomicAdd(&tab[(off1 + i1) * CZ + c], db);
          atomicAdd(&tab[(off2 + i2) * CZ + c], db);
        }
        atomicAdd(&tab[(offd + ibd) * CZ + c], am2 * d);
      }
      if (rm2 != 0.f && same != 0.f) atomicAdd(&tab[(offr + irel) * CZ + c], rm2 * d);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < total_rows * CZ; i += blockDim.x) {
    const float v = tab[i];
    if (v == 0.f) continue;
    const int row = i / CZ, c = i - row * CZ;
    if (row < off1) atomicAdd(d_bond0 + row * CZ + c, v);
    else if (row < off2) atomicAdd(d_bond1 + (row - off1) * CZ + c, v);
    else if (row < offd) atomicAdd(d_bond2 + (row - off2) * CZ + c, v);
    else if (row < offr) atomicAdd(d_bdist + (row - offd) * CZ + c, v);
    else atomicAdd(d_relpos + (row - offr) * CZ + c, v);
  }
}
int bw_pair_static_bwd(const float* d_pair, const float* atom_mask, const float* residue_mask, const float* bond_mask,
                       const int64_t* bond_feats, const int64_t* bond_distance, const int64_t* residue_index,
                       const int64_t* chain_index, int B, int N, int CZ, int max_bd, int max_rel, float* d_bond0,
                       float* d_bond1, float* d_bond2, float* d_bdist, float* d_relpos, cudaStream_t s) {
  const long long R = (long long)B * N * N;
  const size_t smem = (size_t)(13 + max_bd + 1 + 2 * max_rel + 1) * CZ * sizeof(float);
  PRD_REQUIRE(smem <= 48 * 1024, "pair_static_bwd: tables do not fit shared memory");
  bw_pair_static_bwd_kernel<<<grid_for(R, 8, 148 * 2), 256, smem, s>>>(d_pair, atom_mask, residue_mask, bond_mask, bond_feats,
                                                                     bond_distance, residue_index, chain_index, N, CZ, R, max_bd,
                                                                     max_rel, d_bond0, d_bond1, d_bond2, d_bdist, d_relpos);
  PRD_LAUNCHED();
  return 0;
}

__global__ void __launch_bounds__(128) bw_single_embed_bwd_kernel(const float* __restrict__ d_single,
                                                                  const int64_t* __restrict__ atom_feats,
                                                                  const float* __restrict__ atom_mask,
                                                                  const float* __restrict__ residue_mask,
                                                                  const float* __restrict__ seq_t, const float* __restrict__ w_type,
                                                                  int CS, AtomGradTables tabs, float* __restrict__ d_ty,
                                                                  float* __restrict__ d_esm, float* __restrict__ lnseq) {
  __shared__ float sLn[21];
  __shared__ long long sIdx[9];
  const long long tok = blockIdx.x;
  const int t = threadIdx.x;
  constexpr int kVocab[9] = {119, 4, 12, 12, 10, 6, 6, 2, 2};
  if (t < 9) {
    const long long v = atom_feats[tok * 9 + t];
    sIdx[t] = v < 0 ? 0 : (v >= kVocab[t] ? kVocab[t] - 1 : v);
  }
  if (t == 0) {
    const float* sq = seq_t + tok * 21;
    float mean = 0.f;
    for (int k = 0; k < 21; ++k) mean += sq[k];
    mean /= 21.f;
    float var = 0.f;
    for (int k = 0; k < 21; ++k) var += (sq[k] - mean) * (sq[k] - mean);
    const float rstd = rsqrtf(var / 21.f + kLnEps);
    for (int k = 0; k < 21; ++k) {
      sLn[k] = (sq[k] - mean) * rstd;
      lnseq[tok * 21 + k] = sLn[k];
    }
  }
  __syncthreads();
  const float am = atom_mask[tok], rm = residue_mask[tok];
  for (int c = t; c < CS; c += 128) {
    const float d = d_single[tok * CS + c];
    if (am != 0.f) {
      const float da = am * (1.0f / 3.0f) * d;
#pragma unroll
      for (int f = 0; f < 9; ++f) atomicAdd(tabs.t[f] + sIdx[f] * CS + c, da);
    }
    float ty = 0.f;
    const float* wr = w_type + (long long)c * 21;
#pragma unroll
    for (int k = 0; k < 21; ++k) ty += wr[k] * sLn[k];
    d_ty[tok * CS + c] = ty > 0.f ? round_tf32(rm * d) : 0.f;
    d_esm[tok * CS + c] = round_tf32(rm * d);
  }
}
int bw_single_embed_bwd(const float* d_single, const int64_t* atom_feats, const float* atom_mask, const float* residue_mask,
                        const float* seq_t, const float* w_type, int B, int N, int CS, AtomGradTables tabs, float* d_ty,
                        float* d_esm, float* lnseq, cudaStream_t s) {
  bw_single_embed_bwd_kernel<<<B * N, 128, 0, s>>>(d_single, atom_feats, atom_mask, residue_mask, seq_t, w_type, CS, tabs, d_ty,
                                                   d_esm, lnseq);
  PRD_LAUNCHED();
  return 0;
}

}  // namespace prd
