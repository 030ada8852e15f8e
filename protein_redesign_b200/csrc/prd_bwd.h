// Backward-pass building blocks (prd_bwd.cu): fp32 SIMT kernels around the tf32 tensor-core GEMM (prd_gemm.cu).
//
// Conventions of the backward pass
//   * every activation / gradient buffer is fp32; whatever becomes an operand of a tf32 GEMM is written ROUNDED TO
//     NEAREST tf32 by its producer (the tensor core truncates, which would otherwise bias every product by -2^-11);
//   * weight gradients are accumulated (+=) in fp32 by exact FFMA reductions (bw_dw_acc), never through tf32;
//   * nothing is stored by the forward pass except block-boundary checkpoints: each op's backward recomputes its own
//     intermediates from the op's input (the reference checkpoints per block, modules.py:399-401).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace prd {

// out[r, :] = round_tf32(LN(x[r, :]) * gamma + beta)   (gamma / beta may be NULL); eps 1e-5
// out_lo (optional): round_tf32(y - out), so that out + out_lo carries 21 mantissa bits (split-operand GEMMs)
// ldo: row stride of out / out_lo (0 = C); with out_lo = out + C and ldo = 2 C the rows are [hi | lo], the A operand of the
// three-term split GEMM (GemmArgs::split = 3)
int bw_ln_fwd(const float* x, long long R, int C, const float* gamma, const float* beta, float* out, cudaStream_t s,
              float* out_lo = nullptr, long long ldo = 0);
// dx_io[r, :] = round_tf32((accumulate ? dx_io[r, :] : 0) + dLN/dx . g[r, :]);  with gamma: g is first multiplied by gamma,
// and dgamma += sum_r g * xhat, dbeta += sum_r g (either may be NULL)
int bw_ln_bwd(const float* x, const float* g, long long R, int C, const float* gamma, float* dx_io, int accumulate,
              float* dgamma, float* dbeta, cudaStream_t s);
// dW[n, k] += alpha * sum_r dY[r, n] X[r, k]  (n < Nout, k < K);  db[n] += alpha * sum_r dY[r, n] (db may be NULL)
int bw_dw_acc(const float* dY, long long ldy, const float* X, long long ldx, long long R, int Nout, int K, float* dW,
              long long ldw, float* db, float alpha, cudaStream_t s);
// the same reduction on the tensor cores (prd_bwd_dw.cu: tcgen05 kind::tf32, both operands MN-major, split over CTAs);
// bw_dw_acc dispatches to it when bw_dw_tc_applies and PRD_DW_SIMT is not set
bool bw_dw_tc_applies(const float* dY, long long ldy, const float* X, long long ldx, long long R);
// db (optional): db[n] += alpha * sum_r dY[r, n], added up from the dY tiles the kernel stages anyway
int bw_dw_tc(const float* dY, long long ldy, const float* X, long long ldx, long long R, int Nout, int K, float* dW,
             long long ldw, float alpha, cudaStream_t s, float* db = nullptr);
// dst[b][c][r] = round_tf32(alpha * src[b][r][c]),  r < rows, c < cols
int bw_transpose(const float* src, long long lds, long long src_bs, float* dst, long long ldd, long long dst_bs, int rows,
                 int cols, int batch, float alpha, cudaStream_t s);
// dst[r, dcol + c] = round_tf32(src[r, scol + c]) for c < cols (strided 2-D copy, also used to pack weights row-wise)
int bw_copy2d(const float* src, long long lds, float* dst, long long ldd, long long R, int cols, cudaStream_t s);
// Several small weight-preparation jobs (rounded copies / transposes / zero fills / hi-lo splits of matrices of a few
// hundred rows) in ONE launch: every backward op starts with 4 - 9 of them, and a step is ~1000 dependent launches whose
// gaps (2 - 3 us each, also inside a CUDA graph) are not free.  A job is defined on its DESTINATION [drows, dcols]:
// element (i, j) = src(i, j) (copy) or src(j, i) (transpose) where that lies inside the source [rows, cols], else 0 --
// zero padding needs no separate fill, and jobs of one batch must write disjoint memory.
struct PrepJob {
  const float* src;  // NULL: zero fill
  float* dst;
  float* dst_lo;     // split: dst = round(src), dst_lo = round(src - dst)
  long long lds, ldd;
  int rows, cols;    // source extent
  int drows, dcols;  // destination extent
  int transpose, round;
  float alpha;
};
struct PrepBatch {
  static constexpr int kMax = 12;
  PrepJob jobs[kMax];
  int n = 0;
  int err = 0;
  void add(const PrepJob& j) {
    if (n < kMax) jobs[n++] = j;
    else err = 1;
  }
  void copy(const float* src, long long lds, float* dst, long long ldd, int rows, int cols, int round = 1) {
    add(PrepJob{src, dst, nullptr, lds, ldd, rows, cols, rows, cols, 0, round, 1.f});
  }
  // dst [cols (padded to dcols rows... ), ldd]: dst[c][r] = alpha * src[r][c]; the destination extent may exceed the source (zeros)
  void transpose(const float* src, long long lds, float* dst, long long ldd, int rows, int cols, int drows, int dcols, float alpha = 1.f) {
    add(PrepJob{src, dst, nullptr, lds, ldd, rows, cols, drows, dcols, 1, 1, alpha});
  }
  void zero(float* dst, int n_) { add(PrepJob{nullptr, dst, nullptr, 0, n_, 0, 0, 1, n_, 0, 0, 0.f}); }
  void split(const float* src, long long lds, float* hi, float* lo, long long ldd, int rows, int cols) {
    add(PrepJob{src, hi, lo, lds, ldd, rows, cols, rows, cols, 0, 1, 1.f});
  }
};
int bw_prep(const PrepBatch& b, cudaStream_t s);
// hi = round_tf32(src), lo = round_tf32(src - hi)
int bw_split2d(const float* src, long long lds, float* hi, float* lo, long long ldd, long long R, int cols, cudaStream_t s);
// x = round_tf32(max(x, 0))
int bw_relu_inplace(float* x, long long n, cudaStream_t s);
int bw_zero(float* p, long long n, cudaStream_t s);
// db[n] += alpha * sum_r dY[r, n]
int bw_colsum(const float* dY, long long ldy, long long R, int Nout, float* db, float alpha, cudaStream_t s);

// ---- gating ------------------------------------------------------------------------------------------------
// og[r, c] = round(sigmoid(gpre[r, c]) * o[r, c])
int bw_gate_fwd(const float* gpre, long long ldg, const float* o, long long ldo, float* og, long long ldog, long long R,
                int W, cudaStream_t s);
// d_o[r, c] = round(d_og * g);  d_gpre[r, c] = round(d_og * o * g (1 - g))
int bw_gate_bwd(const float* d_og, long long ld1, const float* gpre, long long ldg, const float* o, long long ldo,
                float* d_o, long long ld2, float* d_gpre, long long ld3, long long R, int W, cudaStream_t s);

// ---- head-dim-16 attention (TriangleAttention rows / columns, FoldingBlock.single_attn) -------------------------
struct AttnGeom {
  int B, N, H;       // H heads of 16 channels
  int mode;          // 0: pair rows (starting), 1: pair columns (ending), 2: single representation (one sequence per b)
  const float* mask; // [B, N] token mask
  const float* bias; // [B, H, N, N] (mode 2) or NULL
  float scale;       // 1 / sqrt(16)
};
// qkvg: [rows, ld] with q at column 0, k at 64, v at 128 (gate pre-activation at 192 is not read here)
int bw_attn_fwd(const AttnGeom& g, const float* qkvg, long long ld, float* O, float* lse, cudaStream_t s);
// dqkvg columns 0..191 (dq | dk | dv) are written (rounded); Dbuf: scratch [nseq * H * N]; dbias: [B,H,N,N] or NULL
int bw_attn_bwd(const AttnGeom& g, const float* qkvg, long long ld, const float* O, const float* lse, const float* dO,
                float* Dbuf, float* dqkvg, long long ldd, float* dbias, cudaStream_t s);

// the same two steps on the warp-level tensor cores (prd_bwd_attn.cu: mma.sync tf32, operands rounded to nearest);
// bw_attn_fwd / bw_attn_bwd dispatch to them unless PRD_ATTN_SIMT=1 (then: exact fp32 on the FFMA pipe)
bool bw_attn_tc_enabled();
int bw_attn_tc_fwd(const AttnGeom& g, const float* qkvg, long long ld, float* O, float* lse, cudaStream_t s);
int bw_attn_tc_bwd(const AttnGeom& g, const float* qkvg, long long ld, const float* O, const float* lse, const float* dO,
                   float* Dbuf, float* dqkvg, long long ldd, float* dbias, cudaStream_t s);

// ---- pair-bias projection backward (FoldingBlock.attn_bias, SPAttention.linear_z) ---------------------------------
// bias[b,h,i,j] = sum_c W[h,c] (LN(pair[b,i,j,:]) gamma + beta)_c (+ bvec[h]).  d_pair += dLN(...), dW, dbvec, dgamma, dbeta +=
// dbias[((b*H + h)*N + i)*ldb + j]
int bw_pair_bias_bwd(int B, int N, int CZ, int H, const float* pair, const float* dbias, int ldb, const float* W, const float* gamma,
                     const float* beta, float* d_pair, float* dW, float* dbvec, float* dgamma, float* dbeta, cudaStream_t s);

// ---- channel-last rows <-> channel planes (triangle multiplication) -----------------------------------------------
// rows [B*N*N, ld] (channel c at column col0 + c) -> planes[(b*C + c)][i][j] with row stride Np (natural orientation), rounded
// transposed = 1: planes[(b*C + c)][j][i] (the planes of the transposed pair tensor) straight from the rows
int bw_rows_to_planes(const float* rows, long long ld, int col0, int B, int N, int C, int Np, float* planes, cudaStream_t s,
                      int transposed = 0);
int bw_planes_to_rows(const float* planes, int B, int N, int C, int Np, float* rows, long long ld, int col0, cudaStream_t s);
// ab[r, 0:2C] = round(m_i m_j sigmoid(pre[r, 2C + c]) pre[r, c])          (pre: [R, ld], proj at 0, gate at 2C)
int bw_trimul_ab(const float* pre, long long ld, const float* mask, int B, int N, int C2, float* ab, cudaStream_t s);
// dpre[r, c] = round(dab m2 g), dpre[r, C2 + c] = round(dab m2 p g (1 - g))
int bw_trimul_ab_bwd(const float* pre, long long ld, const float* mask, int B, int N, int C2, const float* dab, float* dpre,
                     long long ldd, cudaStream_t s);
// y = sigmoid(gpre) * o : d_o = round(dy g), d_gpre = round(dy o g (1-g))  -> bw_gate_bwd

// ---- softmax over the last dim (SPAttention, materialised) ----------------------------------------------------------
int bw_softmax_rows(const float* logits, float* probs, long long rows, int n, int ld, cudaStream_t s);  // probs rounded
int bw_softmax_bwd_rows(const float* probs, const float* dprobs, float* dlogits, long long rows, int n, int ld, cudaStream_t s);

// ---- bilinear pair terms: y[b,i,j,z] = sum_c W[z,c] l[b,i,c] r[b,j,c]   (OuterLinear product term, OuterProductUpdate) ----
// E: [B,N,N,CZ] incoming gradient (already scaled / masked by the caller).
// ET[b,i,z,j] = E[b,i,j,z] (+ E[b,j,i,z] if symmetric), row stride Np;  rounded
int bw_pair_to_izj(const float* E, int B, int N, int CZ, int Np, int transpose_ij, int add_transposed, float* ET, cudaStream_t s);
// T: [B,N,CZ,C] = sum_j ET[b,n,z,j] other[b,j,c] (from the GEMM).  d_this[b,n,c] (+)= sum_z W[z,c] T[b,n,z,c];
// dW[z,c] += sum_{b,n} this[b,n,c] T[b,n,z,c]  (dW may be NULL)
// W / dW have row stride ldw; d_self gets alpha * (...), dW gets alpha_dw * (...)
int bw_bilinear_reduce(const float* T, const float* W, long long ldw, const float* self, int B, int N, int CZ, int C,
                       float* d_self, int accumulate, float* dW, float alpha, float alpha_dw, cudaStream_t s);
// U[b,n,z] = sum_j E[b,n,j,z] - sum_i E[b,i,n,z]   (OuterLinear difference term), rounded
int bw_pair_rowcol_diff(const float* E, int B, int N, int CZ, float* U, cudaStream_t s);

// ---- elementwise helpers ---------------------------------------------------------------------------------------
// out[r, c] = round(alpha * a[r, c] * rowscale[r])          (rowscale may be NULL)
int bw_scale_rows(const float* a, long long lda, const float* rowscale, float alpha, float* out, long long ldo, long long R,
                  int W, cudaStream_t s);
// out[b,i,j,:] = round(m_i m_j E[b,i,j,:] * alpha)
int bw_mask_pair(const float* E, const float* mask, int B, int N, int CZ, float alpha, float* out, cudaStream_t s);
// dst = round(a + b)
int bw_add(const float* a, const float* b, float* dst, long long n, cudaStream_t s);

// ---- coordinate head (model.py:364-373) ------------------------------------------------------------------------------
// d_eps = d_out - m * sum_i(m_i d_out_i) / sum(m)   (adjoint of remove_mean)
int bw_remove_mean_adj(const float* d_out, const float* mask, int B, int N, float* d_eps, cudaStream_t s);
// per pair row: dw = m_i m_j <d_eps_i, r_ij>;  dh[r, :] = round(dw * w2 * [h > 0]);  dw2 += sum_r dw h[r, :]
int bw_coord_dh(const float* h, const float* z, const float* mask, const float* d_eps, const float* w2, int B, int N, int CZ,
                float* dh, float* dw2, cudaStream_t s);

// ---- embeddings -----------------------------------------------------------------------------------------------------
// rbf[r, k] = round(exp(-scale (|z_i - z_j| - center_k)^2))      [B*N*N, DD]
int bw_rbf_rows(const float* z, const float* centers, float scale, int B, int N, int DD, float* rbf, cudaStream_t s);
// dW_beta[z, k] += sum_b (sum_ij Dm[b,i,j,z]) sincos_k(t_b / T)
int bw_time_embed_bwd(const float* Dm, const int64_t* t, int num_steps, const float* freq, int B, int N, int CZ, int TD,
                      float* colsum_scratch, float* dW_beta, cudaStream_t s);
// scatter of d_pair into the step-invariant tables (bond features x3, bond distance, relpos)
int bw_pair_static_bwd(const float* d_pair, const float* atom_mask, const float* residue_mask, const float* bond_mask,
                       const int64_t* bond_feats, const int64_t* bond_distance, const int64_t* residue_index,
                       const int64_t* chain_index, int B, int N, int CZ, int max_bd, int max_rel, float* d_bond0,
                       float* d_bond1, float* d_bond2, float* d_bdist, float* d_relpos, cudaStream_t s);
// single embedding: atom tables scatter; d_ty[tok, c] = round(rm d_single [W_type LN(seq_t) > 0]); d_esm_in[tok,c] = round(rm d_single);
// lnseq[tok, 0:21] = LN(seq_t)
struct AtomGradTables { float* t[9]; };
int bw_single_embed_bwd(const float* d_single, const int64_t* atom_feats, const float* atom_mask, const float* residue_mask,
                        const float* seq_t, const float* w_type, int B, int N, int CS, AtomGradTables tabs, float* d_ty,
                        float* d_esm, float* lnseq, cudaStream_t s);

}  // namespace prd
