// Internal launcher declarations (host side).  Public C-ABI: include/prd_denoiser.h.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace prd {

// ---- batched fp16 GEMM (prd_gemm.cu) ---------------------------------------------------
struct GemmArgs {
  int M = 0, N = 0, K = 0;
  int nb1 = 1, nb2 = 1;  // two batch levels; blockIdx.z = i2 * nb1 + i1
  const void* A = nullptr;  // [M, K] row-major (K contiguous), leading dim lda (elements: halves, or floats with tf32)
  long long lda = 0, a_bs1 = 0, a_bs2 = 0;  // batch strides in elements; 0 = broadcast
  const void* B = nullptr;  // [N, K] row-major ("weight" layout)
  long long ldb = 0, b_bs1 = 0, b_bs2 = 0;
  float alpha = 1.0f;
  const float* bias = nullptr;      // [N]
  int act = 0;                      // 0 none, 1 relu, 2 sigmoid
  const float* rowscale = nullptr;  // [M] (per batch with rs_bs*)
  long long rs_bs1 = 0, rs_bs2 = 0;
  const float* mul = nullptr;  // [M, N] fp32
  long long ldmul = 0, mul_bs1 = 0, mul_bs2 = 0;
  const float* add = nullptr;  // [M, N] fp32
  long long ldadd = 0, add_bs1 = 0, add_bs2 = 0;
  void* C = nullptr;  // fp32 or fp16 [M, N]
  long long ldc = 0, c_bs1 = 0, c_bs2 = 0;
  int c_fp16 = 0;
  // 0: plain.  1: B is a weight stored as [hi | lo] along K (ldb >= 2K).  2: same for A.
  // 3: three-term split product in ONE accumulator: A = [a_hi | a_lo] (lda >= 2K), B = [b_hi | b_hi | b_lo] (ldb >= 3K):
  //    a_hi b_hi + a_lo b_hi + a_hi b_lo (the lo x lo term is below fp32 resolution)
  int split = 0;
  int tf32 = 0;        // A, B are fp32 in memory, multiplied on kind::tf32 (producers round to nearest tf32); C fp32
  int mul_step = 0;    // `mul` gates instead of scaling: v = mul > 0 ? v : 0
  int round_tf32 = 0;  // round the fp32 result to nearest tf32 (it becomes a tf32 operand next)
};
// v = alpha*acc (+bias) -> act -> *rowscale -> *mul -> +add
int gemm_f16(const GemmArgs& g, cudaStream_t stream);

// ---- pair-row tile kernels (prd_rowtile.cu): 128 pair elements per tile, thread per row ----
struct PairDims {
  int B, N, CZ;
  int all_valid = 0;  // caller's promise that every token of the batch is valid (mask all ones): selects the attention core
                      // without per-sequence ragged handling; 0 = unknown (always correct)
};
inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
inline int plane_ld(int N) { return round_up(N, 64); }  // row stride (halves) of fp16 channel planes
inline int xplane_ld(int N) { return round_up(N, 8); }  // row stride (halves) of the fp16 contraction-result planes

// All residual ops: dst = (residual ? src : 0) + update; dst may alias src.
// bias_out (optional, [B,4,N,N]): the NEXT block's attention bias of the updated rows, LN(dst row) . w_bias^T + b_bias
// (modules.py:300-304), emitted from the output epilogue so that the pair tensor is not read again for it
int pair_transition(const PairDims& d, const float* pair, float* dst, int residual, const __half* w1, const float* b1,
                    const __half* w2, const float* b2, int hidden, const float* w_bias, const float* b_bias, float* bias_out,
                    cudaStream_t s);
// mode 0 = outgoing, 1 = incoming.  ab: [2][B][CZ][N][plane_ld(N)] fp16 channel planes.
int trimul_in(const PairDims& d, const float* pair, const float* mask, int mode, const __half* w_in, const float* b_in,
              __half* ab, cudaStream_t s);
// x: [B][CZ][N][xplane_ld(N)] fp16 contraction result.
int trimul_out(const PairDims& d, const float* pair, float* dst, int residual, const __half* x, const __half* w_out,
               const float* b_out, cudaStream_t s);
// mode 0 = starting, 1 = ending.  q,k,g: [B*N*N][64] fp16 (logical row = (b, seq, tok));
// vt: [B*N][64][plane_ld(N)] fp16.
int triattn_proj(const PairDims& d, const float* pair, int mode, const __half* w_qkvg, const float* b_gate, __half* q,
                 __half* k, __half* g, __half* vt, cudaStream_t s);
int triattn_flash_g4(const PairDims& d, const float* mask, const __half* q, const __half* k, const __half* g,
                     const __half* vt, __half* og, cudaStream_t s);  // four softmax groups per SM (prd_triattn4.cu)
int triattn_flash_out_g4(const PairDims& d, const float* mask, const __half* q, const __half* k, const __half* g,
                         const __half* vt, const float* pair, float* dst, int residual, int mode, const __half* w_o,
                         const float* b_o, cudaStream_t s);  // attention core + out_proj + residual fused
bool triattn_flash_g4_applies(const PairDims& d);
int triattn_flash(const PairDims& d, const float* mask, const __half* q, const __half* k, const __half* g,
                  const __half* vt, __half* og, cudaStream_t s);
int triattn_out(const PairDims& d, const float* pair, float* dst, int residual, int mode, const __half* og,
                const __half* w_o, const float* b_o, cudaStream_t s);
// eps_raw[b,i,:] = sum_j m_i m_j w_ij r_ij on the symmetrised pair (mean not removed yet)
int coord_head(const PairDims& d, const float* pair, const float* z, const float* mask, const __half* w1,
               const float* b1, const float* w2, float* eps_raw, cudaStream_t s);
// pair = static + m2*(W_dist rbf(d) + beta[b]) + m2*(W_o (a_i*b_j) + b_o)/(m2+1e-3)
int pair_embed_dynamic(const PairDims& d, const float* pair_static, float* pair, const float* z, const float* mask,
                       const float* beta, const __half* w_dist, int dist_dim, const float* centers, float rbf_scale,
                       const float* opm_a, const float* opm_b, int opm_dim, const __half* w_opm, const float* b_opm,
                       int flags, const float* rbf_lut, cudaStream_t s);
// rbf_lut: [(M + 2)][CZ] floats = header row {1/h, M} + M + 1 table rows of d -> W_dist rbf(d) at d = m h, h = d_max / M
int rbf_lut_build(int CZ, int DD, const float* w_dist, const float* centers, float scale, float d_max, int M, float* lut,
                  cudaStream_t s);
// pair[b,i,j,:] += W1 . (x_i * x_j) + u_i - u_j + bias   (OuterLinear, bilinear form)
int outer_linear(const PairDims& d, int CS, const float* pair, float* dst, int residual, const __half* xn16,
                 const float* xn32, const __half* w1, const float* u, const float* bias, cudaStream_t s);

// ---- SIMT / bandwidth kernels (prd_simt.cu) ----------------------------------------------
int layernorm_rows(const float* x, int rows, int C, const float* gamma, const float* beta, __half* out16,
                   float* out32, cudaStream_t s);
// bias_out[b,h,i,j] = (LN(pair[b,i,j,:]) (affine optional)) . w[h,:] + bvec[h]
int pair_bias_proj(const PairDims& d, int H, const float* pair, const float* ln_w, const float* ln_b, const float* w,
                   const float* bvec, float* bias_out, cudaStream_t s);
// the same with an optional second projection of the same rows (bias_out2 NULL = one projection)
int pair_bias_proj2(const PairDims& d, int H, const float* pair, const float* ln_w, const float* ln_b, const float* w,
                    const float* bvec, float* bias_out, const float* ln_w2, const float* ln_b2, const float* w2,
                    const float* bvec2, float* bias_out2, cudaStream_t s);
int softmax_rows(float* logits, __half* probs, long long rows, int n, int ld_in, int ld_out, cudaStream_t s);
// FoldingBlock.single_attn core: qkvg [B*N, 4*H*c] fp32 (q|k|v|gate pre-activation, gate bias added),
// bias [B,H,N,N], mask [B,N] -> og [B*N, H*c] fp16 (gated attention output)
int single_attention(int B, int N, int H, int c, const float* qkvg, const float* bias, const float* mask, __half* og,
                     cudaStream_t s);
int symmetrize_pair(const PairDims& d, float* pair, cudaStream_t s);
int remove_mean3(int B, int N, int C, float* x, const float* mask, int mask_rows, cudaStream_t s);
int embed_pair_static(const PairDims& d, const float* atom_mask, const float* residue_mask, const float* bond_mask,
                      const int64_t* bond_feats, const int64_t* bond_distance, const int64_t* residue_index,
                      const int64_t* chain_index, const float* const* bond_tables, const int* bond_vocab,
                      const float* bdist_table, int max_bond_distance, const float* relpos_table, int max_relpos,
                      float* out, cudaStream_t s);

}  // namespace prd
