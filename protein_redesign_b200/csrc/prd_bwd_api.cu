// extern "C" backward entry points (include/prd_denoiser.h, "backward pass"): one prd_<op>_bwd per forward op.
//
// Each op recomputes its own intermediates from the op's INPUT (the reference checkpoints per block and recomputes,
// modules.py:399-401), in unfused fp32 form: LayerNorm / gating / softmax on SIMT kernels (prd_bwd.cu), every
// activation product on the tcgen05 GEMM with tf32 operands (prd_gemm.cu), every weight gradient as an exact fp32
// reduction.  Gradient buffers are in/out (`out[0]`: d(output) on entry, d(input) on exit for the residual ops);
// weight gradients are ACCUMULATED (+=) into caller-owned fp32 buffers shaped like the reference parameters.
// `weights` are the RAW fp32 reference parameters ([out, in] nn.Linear layout), not the packed fp16 forward copies.
#include "../../include/prd_denoiser.h"
#include "prd_bwd.h"
#include "prd_common.cuh"
#include "prd_kernels.h"

using namespace prd;

namespace {

struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* b) : base(static_cast<char*>(b)) {}
  float* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    float* p = base ? reinterpret_cast<float*>(base + off) : nullptr;
    off += n * sizeof(float);
    return p;
  }
  size_t total() const { return (off + 255) & ~size_t(255); }
};

inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }
inline const float* inf(const void* const* a, int i) { return static_cast<const float*>(a[i]); }
inline float* outf(void* const* a, int i) { return static_cast<float*>(a[i]); }
inline int up4(int x) { return (x + 3) & ~3; }

// every helper below is a no-op in "dry" mode (workspace sizing): only the carve sequence runs
#define RUN(expr)                  \
  do {                             \
    if (!dry && (expr)) return 1;  \
  } while (0)

GemmArgs tfg(int M, int N, int K, const float* A, long long lda, const float* B, long long ldb, float* C, long long ldc) {
  GemmArgs g;
  g.M = M; g.N = N; g.K = K;
  g.A = A; g.lda = lda;
  g.B = B; g.ldb = ldb;
  g.C = C; g.ldc = ldc;
  g.tf32 = 1;
  g.round_tf32 = 1;
  return g;
}

// W [rows, cols] (row stride lds) -> rounded copy [rows, up4(cols)]
float* rounded(bool dry, Carver& c, cudaStream_t st, const float* W, int rows, int cols, long long lds, int* rc) {
  float* r = c.take((size_t)rows * up4(cols));
  if (!dry && bw_copy2d(W, lds, r, up4(cols), rows, cols, st)) *rc = 1;
  return r;
}
// W [rows, cols] -> W^T [cols, up4(rows)] rounded
float* transposed(bool dry, Carver& c, cudaStream_t st, const float* W, int rows, int cols, long long lds, int* rc, float alpha = 1.f) {
  float* r = c.take((size_t)cols * up4(rows));
  if (!dry && bw_transpose(W, lds, 0, r, up4(rows), 0, rows, cols, 1, alpha, st)) *rc = 1;
  return r;
}

// h = relu(LN(x) W^T + b) with split operands (x_hi + x_lo)(W_hi + W_lo): the ReLU mask [h > 0] of the backward pass must
// agree with an fp32 evaluation -- with single tf32 operands ~3e-4 of the pre-activations change sign, and every flipped
// element costs a full-size error in dh (relative L2 of dW ~ sqrt(3e-4) = 1.7e-2, measured 1.3e-2).  The three products
// x_hi W_hi + x_lo W_hi + x_hi W_lo run as ONE tf32 GEMM over K = 3 Cin into one accumulator (GemmArgs::split = 3: LayerNorm
// writes rows [hi | lo], the weight is laid out [W_hi | W_hi | W_lo]) with bias, ReLU and the tf32 rounding in its epilogue;
// as three GEMMs chained through memory plus a ReLU pass the same math moved 3.3 x the bytes.
// *xop / *ldx: LN(x) rounded (the hi half of the rows), the operand of the weight-gradient reductions.
// prep: jobs of the caller that go into the same weight-preparation launch as the hi / lo split of W
int relu_layer_fwd(bool dry, Carver& c, cudaStream_t st, long long R, int Cin, int Chid, const float* x, const float* W,
                   const float* b, const float** xop, long long* ldx, float* h, PrepBatch* prep = nullptr) {
  float* xcat = c.take((size_t)R * 2 * Cin);      // rows [LN(x)_hi | LN(x)_lo]
  float* Wcat = c.take((size_t)Chid * 3 * Cin);   // rows [W_hi | W_hi | W_lo]
  *xop = xcat;
  *ldx = 2 * Cin;
  if (dry) return 0;
  PRD_REQUIRE(Cin % 32 == 0, "relu_layer_fwd: input width %d must be a multiple of 32", Cin);
  PrepBatch own;
  PrepBatch& pb = prep ? *prep : own;
  pb.split(W, Cin, Wcat, Wcat + 2 * Cin, 3 * Cin, Chid, Cin);
  pb.copy(W, Cin, Wcat + Cin, 3 * Cin, Chid, Cin);
  if (bw_prep(pb, st)) return 1;
  if (bw_ln_fwd(x, R, Cin, nullptr, nullptr, xcat, st, xcat + Cin, 2 * Cin)) return 1;
  GemmArgs g = tfg((int)R, Chid, Cin, xcat, 2 * Cin, Wcat, 3 * Cin, h, Chid);
  g.split = 3; g.bias = b; g.act = 1;
  return gemm_f16(g, st);
}

// ---------------------------------------------------------------------------------------------------------------
// LN -> Linear(Cin -> Chid) -> ReLU -> Linear(Chid -> Cout) (+ residual): single_fc / pair_fc (modules.py:306-326),
// seq_mlp (model.py:117-122).  dx_io: d(out) on entry when residual, d(in) on exit.
// ---------------------------------------------------------------------------------------------------------------
int mlp_bwd(bool dry, Carver& c, cudaStream_t st, long long R, int Cin, int Chid, int Cout, const float* x, const float* dy,
            const float* W1, const float* b1, const float* W2, float* dx_io, int residual, float* dW1, float* db1, float* dW2,
            float* db2) {
  int rc = 0;
  const int Cop = up4(Cout) < 32 && Cout % 4 != 0 ? 32 : up4(Cout);  // padded width of dy as a GEMM operand
  const float* xh = nullptr;
  long long ldxh = 0;
  float* h = c.take((size_t)R * Chid);
  float* dh = c.take((size_t)R * Chid);
  float* dxh = c.take((size_t)R * Cin);
  float* W1T = c.take((size_t)Cin * up4(Chid));                   // [Cin, Chid]
  float* W2T = c.take((size_t)Chid * Cop);                        // [Chid, Cop] (zero padded)
  PrepBatch pb;
  if (!dry) {
    pb.transpose(W1, Cin, W1T, up4(Chid), Chid, Cin, Cin, Chid);
    pb.transpose(W2, Chid, W2T, Cop, Cout, Chid, Chid, Cop);
  }
  if (relu_layer_fwd(dry, c, st, R, Cin, Chid, x, W1, b1, &xh, &ldxh, h, &pb)) return 1;
  const float* dyop = dy;
  float* dypad = nullptr;
  if (Cop != Cout) dypad = c.take((size_t)R * Cop);
  if (dry) return 0;
  if (rc) return 1;
  if (dypad) {
    if (bw_zero(dypad, R * Cop, st)) return 1;
    if (bw_copy2d(dy, Cout, dypad, Cop, R, Cout, st)) return 1;
    dyop = dypad;
  }
  if (bw_dw_acc(dy, Cout, h, Chid, R, Cout, Chid, dW2, Chid, db2, 1.f, st)) return 1;
  {
    GemmArgs g = tfg((int)R, Chid, Cop, dyop, Cop, W2T, Cop, dh, Chid);
    g.mul = h; g.ldmul = Chid; g.mul_step = 1;
    if (gemm_f16(g, st)) return 1;
  }
  if (bw_dw_acc(dh, Chid, xh, ldxh, R, Chid, Cin, dW1, Cin, db1, 1.f, st)) return 1;
  {
    GemmArgs g = tfg((int)R, Cin, Chid, dh, Chid, W1T, up4(Chid), dxh, Cin);
    if (gemm_f16(g, st)) return 1;
  }
  return bw_ln_bwd(x, dxh, R, Cin, nullptr, dx_io, residual, nullptr, nullptr, st);
}

// ---------------------------------------------------------------------------------------------------------------
// gated attention with 4 x 16 heads (modules.py:185-225) on rows of width Cin: TriangleAttention (mode 0 / 1) and
// FoldingBlock.single_attn (mode 2, with pair bias).  dx_io: d(out) on entry, d(in) = d(out) + branch on exit.
// ---------------------------------------------------------------------------------------------------------------
struct AttnW {
  const float *Wq, *Wk, *Wv, *Wg, *bg, *Wo, *bo;
  float *dWq, *dWk, *dWv, *dWg, *dbg, *dWo, *dbo;
};
int gated_attn_bwd(bool dry, Carver& c, cudaStream_t st, const AttnGeom& geom, long long R, int Cin, const float* x, float* dx_io,
                   const AttnW& w, float* dbias) {
  int rc = 0;
  const long long nseq = geom.mode == 2 ? geom.B : (long long)geom.B * geom.N;
  float* xh = c.take((size_t)R * Cin);
  float* qkvg = c.take((size_t)R * 256);
  float* O = c.take((size_t)R * 64);
  float* og = c.take((size_t)R * 64);
  float* d_og = c.take((size_t)R * 64);
  float* dO = c.take((size_t)R * 64);
  float* dqkvg = c.take((size_t)R * 256);
  float* dxh = c.take((size_t)R * Cin);
  float* lse = c.take((size_t)nseq * geom.H * geom.N);
  float* Dbuf = c.take((size_t)nseq * geom.H * geom.N);
  float* Wcat = c.take((size_t)256 * Cin);
  float* bcat = c.take(256);
  float* WcatT = c.take((size_t)Cin * 256);
  float* WoT = c.take((size_t)64 * Cin);
  if (dry) return 0;
  (void)rc;
  {
    const float* ws[4] = {w.Wq, w.Wk, w.Wv, w.Wg};
    PrepBatch pb;
    for (int k = 0; k < 4; ++k) {
      pb.copy(ws[k], Cin, Wcat + (size_t)k * 64 * Cin, Cin, 64, Cin);
      pb.transpose(ws[k], Cin, WcatT + 64 * k, 256, 64, Cin, Cin, 64);  // columns [64 k, 64 k + 64) of [Cin, 256]
    }
    pb.zero(bcat, 192);
    pb.copy(w.bg, 64, bcat + 192, 64, 1, 64, 0);
    pb.transpose(w.Wo, 64, WoT, Cin, Cin, 64, 64, Cin);  // Wo [Cin, 64] -> [64, Cin]
    if (bw_prep(pb, st)) return 1;
  }
  if (bw_ln_fwd(x, R, Cin, nullptr, nullptr, xh, st)) return 1;
  {
    GemmArgs g = tfg((int)R, 256, Cin, xh, Cin, Wcat, Cin, qkvg, 256);
    g.bias = bcat; g.round_tf32 = 0;
    if (gemm_f16(g, st)) return 1;
  }
  if (bw_attn_fwd(geom, qkvg, 256, O, lse, st)) return 1;
  if (bw_gate_fwd(qkvg + 192, 256, O, 64, og, 64, R, 64, st)) return 1;
  // out = x + og Wo^T + bo
  if (bw_dw_acc(dx_io, Cin, og, 64, R, Cin, 64, w.dWo, 64, w.dbo, 1.f, st)) return 1;
  {
    GemmArgs g = tfg((int)R, 64, Cin, dx_io, Cin, WoT, Cin, d_og, 64);
    if (gemm_f16(g, st)) return 1;
  }
  if (bw_gate_bwd(d_og, 64, qkvg + 192, 256, O, 64, dO, 64, dqkvg + 192, 256, R, 64, st)) return 1;
  if (bw_attn_bwd(geom, qkvg, 256, O, lse, dO, Dbuf, dqkvg, 256, dbias, st)) return 1;
  float* dws[4] = {w.dWq, w.dWk, w.dWv, w.dWg};
  const size_t wsz = (size_t)64 * Cin;
  if (w.dWk == w.dWq + wsz && w.dWv == w.dWq + 2 * wsz && w.dWg == w.dWq + 3 * wsz) {
    // the four gradients are adjacent (one flat gradient bucket in parameter order, autograd.FlatGrads): one reduction
    // [R, 256]^T [R, Cin] instead of four that each re-read xh
    if (bw_dw_acc(dqkvg, 256, xh, Cin, R, 256, Cin, w.dWq, Cin, nullptr, 1.f, st)) return 1;
    if (bw_colsum(dqkvg + 192, 256, R, 64, w.dbg, 1.f, st)) return 1;
  } else {
    for (int k = 0; k < 4; ++k)
      if (bw_dw_acc(dqkvg + 64 * k, 256, xh, Cin, R, 64, Cin, dws[k], Cin, k == 3 ? w.dbg : nullptr, 1.f, st)) return 1;
  }
  {
    GemmArgs g = tfg((int)R, Cin, 256, dqkvg, 256, WcatT, 256, dxh, Cin);
    if (gemm_f16(g, st)) return 1;
  }
  return bw_ln_bwd(x, dxh, R, Cin, nullptr, dx_io, 1, nullptr, nullptr, st);
}

}  // namespace

extern "C" {

// Test hook: dW[n, k] += alpha * sum_r dY[r, n] X[r, k]  (the weight-gradient reduction; mode 0 = dispatch like the
// backward ops do, 1 = force the SIMT kernel, 2 = force the tensor-core kernel)
int prd_dw_acc(const float* dY, long long ldy, const float* X, long long ldx, long long R, int Nout, int K, float* dW,
               long long ldw, float* db, float alpha, int mode, void* stream) {
  if (prd_device_check()) return 1;
  if (mode == 2) {
    PRD_REQUIRE(bw_dw_tc_applies(dY, ldy, X, ldx, R), "dw_acc: operands do not qualify for the tensor-core kernel");
    return bw_dw_tc(dY, ldy, X, ldx, R, Nout, K, dW, ldw, alpha, S(stream), db);
  }
  return bw_dw_acc(dY, ldy, X, ldx, R, Nout, K, dW, ldw, db, alpha, S(stream));
}

#define PRD_BWD_OP(name)                                                                                              \
  static int name##_impl(bool dry, const PrdDims* d, const void* const* in, void* const* out, const void* const* w,   \
                         Carver& c, cudaStream_t st);                                                                 \
  size_t prd_##name##_bwd_workspace_bytes(const PrdDims* d) {                                                         \
    Carver c(nullptr);                                                                                                \
    name##_impl(true, d, nullptr, nullptr, nullptr, c, nullptr);                                                      \
    return c.total();                                                                                                 \
  }                                                                                                                   \
  int prd_##name##_bwd(const PrdDims* d, const void* const* in, void* const* out, const void* const* w,               \
                       void* workspace, size_t workspace_bytes, void* stream) {                                       \
    if (prd_device_check()) return 1;                                                                                 \
    PRD_REQUIRE(workspace_bytes >= prd_##name##_bwd_workspace_bytes(d), "%s: workspace too small", __func__);         \
    Carver c(workspace);                                                                                              \
    return name##_impl(false, d, in, out, w, c, S(stream));                                                           \
  }                                                                                                                   \
  static int name##_impl(bool dry, const PrdDims* d, const void* const* in, void* const* out, const void* const* w,   \
                         Carver& c, cudaStream_t st)

#define IN(i) (dry ? nullptr : inf(in, i))
#define OUT(i) (dry ? nullptr : outf(out, i))
#define WT(i) (dry ? nullptr : inf(w, i))

// modules.py:321-326,342 pair_fc.   in: [pair_in]   out: [d_pair io | dW1 | db1 | dW2 | db2]   weights: [W1 | b1 | W2 | b2]
PRD_BWD_OP(pair_transition) {
  const long long R = (long long)d->B * d->N * d->N;
  return mlp_bwd(dry, c, st, R, d->c_z, d->c_z * d->tf, d->c_z, IN(0), OUT(0), WT(0), WT(1), WT(2), OUT(0), 1, OUT(1), OUT(2),
                 OUT(3), OUT(4));
}

// modules.py:306-311,336 single_fc.  in: [single_in]  out: [d_single io | dW1 | db1 | dW2 | db2]  weights: [W1 | b1 | W2 | b2]
PRD_BWD_OP(single_transition) {
  const long long R = (long long)d->B * d->N;
  return mlp_bwd(dry, c, st, R, d->c_s, d->c_s * d->tf, d->c_s, IN(0), OUT(0), WT(0), WT(1), WT(2), OUT(0), 1, OUT(1), OUT(2),
                 OUT(3), OUT(4));
}

// model.py:374 seq_mlp.  in: [single | d_seq_pred B,N,21]  out: [d_single (written) | dW1 | db1 | dW2]  weights: [W1 | b1 | W2]
PRD_BWD_OP(seq_head) {
  const long long R = (long long)d->B * d->N;
  return mlp_bwd(dry, c, st, R, d->c_s, d->c_s, 21, IN(0), IN(1), WT(0), WT(1), WT(2), OUT(0), 0, OUT(1), OUT(2), OUT(3), nullptr);
}

// modules.py:236-243 TriangleAttention (d->mode 0 starting / 1 ending).
// in: [pair_in | mask]   out: [d_pair io | dWq | dWk | dWv | dWg | dbg | dWo | dbo]   weights: [Wq | Wk | Wv | Wg | bg | Wo | bo]
PRD_BWD_OP(triangle_attention) {
  const long long R = (long long)d->B * d->N * d->N;
  AttnGeom g{d->B, d->N, d->H, d->mode, IN(1), nullptr, 0.25f};
  AttnW a{WT(0), WT(1), WT(2), WT(3), WT(4), WT(5), WT(6), OUT(1), OUT(2), OUT(3), OUT(4), OUT(5), OUT(6), OUT(7)};
  if (!dry) PRD_REQUIRE(d->H == 4 && d->c == 16, "triangle_attention_bwd: built for 4 heads x 16 channels");
  return gated_attn_bwd(dry, c, st, g, R, d->c_z, IN(0), OUT(0), a, nullptr);
}

// modules.py:300-304 attn_bias + :185-225 single_attn (+ residual :335).
// in: [single_in | pair | mask]
// out: [d_single io | d_pair io (+=) | dWb | dbb | dWq | dWk | dWv | dWg | dbg | dWo | dbo]
// weights: [Wb H x c_z | bb | Wq | Wk | Wv | Wg | bg | Wo | bo]
PRD_BWD_OP(single_attention) {
  const long long R = (long long)d->B * d->N;
  float* bias = c.take((size_t)d->B * d->H * d->N * d->N);
  float* dbias = c.take((size_t)d->B * d->H * d->N * d->N);
  if (!dry) {
    PRD_REQUIRE(d->H == 4 && d->c == 16, "single_attention_bwd: built for 4 heads x 16 channels");
    if (pair_bias_proj(PairDims{d->B, d->N, d->c_z}, d->H, IN(1), nullptr, nullptr, WT(0), WT(1), bias, st)) return 1;
  }
  AttnGeom g{d->B, d->N, d->H, 2, IN(2), bias, 0.25f};
  AttnW a{WT(2), WT(3), WT(4), WT(5), WT(6), WT(7), WT(8), OUT(4), OUT(5), OUT(6), OUT(7), OUT(8), OUT(9), OUT(10)};
  if (gated_attn_bwd(dry, c, st, g, R, d->c_s, IN(0), OUT(0), a, dbias)) return 1;
  if (dry) return 0;
  return bw_pair_bias_bwd(d->B, d->N, d->c_z, d->H, IN(1), dbias, d->N, WT(0), nullptr, nullptr, OUT(1), OUT(2), OUT(3), nullptr,
                          nullptr, st);
}

// modules.py:262-274 TriangleMultiplication (d->mode 0 outgoing / 1 incoming).
// in: [pair_in | mask]
// out: [d_pair io | dW_ab | db_ab | dG_ab | dbg_ab | dW_o | db_o | dG_o | dbg_o]
// weights: [ab_proj.weight 2c_z x c_z | ab_proj.bias | ab_gate.weight | ab_gate.bias | out_proj.weight | out_proj.bias |
//           out_gate.weight | out_gate.bias]
PRD_BWD_OP(triangle_multiplication) {
  const int B = d->B, N = d->N, CZ = d->c_z, C2 = 2 * CZ, CP = 5 * CZ, Np = up4(N);
  const long long R = (long long)B * N * N;
  const size_t PL = (size_t)B * CZ * N * Np;  // one set of channel planes
  float* p = c.take((size_t)R * CZ);
  float* pre = c.take((size_t)R * CP);
  float* ab = c.take((size_t)R * C2);
  float* pa = c.take(PL);
  float* pb = c.take(PL);
  float* paT = c.take(PL);
  float* pbT = c.take(PL);
  float* xp = c.take(PL);
  float* xc = c.take((size_t)R * CZ);
  float* xn = c.take((size_t)R * CZ);
  float* o = c.take((size_t)R * CZ);
  float* d_o = c.take((size_t)R * CZ);
  float* dxn = c.take((size_t)R * CZ);
  float* dxc = c.take((size_t)R * CZ);
  float* dxp = c.take(PL);
  float* dxpT = c.take(PL);
  float* dap = c.take(PL);
  float* dbp = c.take(PL);
  float* dab = c.take((size_t)R * C2);
  float* dpre = c.take((size_t)R * CP);
  float* dp = c.take((size_t)R * CZ);
  float* Wcat = c.take((size_t)CP * CZ);
  float* bcat = c.take(CP);
  float* WcatT = c.take((size_t)CZ * up4(CP));
  float* Wor = c.take((size_t)CZ * CZ);
  float* WoT = c.take((size_t)CZ * CZ);
  if (dry) return 0;
  PRD_REQUIRE(d->mode == 0 || d->mode == 1, "triangle_multiplication_bwd: invalid mode %d", d->mode);
  const float *x = IN(0), *mask = IN(1);
  float* dy = OUT(0);
  // packed projection [ab_proj ; ab_gate ; out_gate] : [5 c_z, c_z]
  {
    PrepBatch pb;
    const float* wsrc[3] = {WT(0), WT(2), WT(6)};
    const float* bsrc[3] = {WT(1), WT(3), WT(7)};
    const int wrows[3] = {C2, C2, CZ}, woff[3] = {0, C2, 2 * C2};
    for (int k = 0; k < 3; ++k) {
      pb.copy(wsrc[k], CZ, Wcat + (size_t)woff[k] * CZ, CZ, wrows[k], CZ);
      pb.transpose(wsrc[k], CZ, WcatT + woff[k], up4(CP), wrows[k], CZ, CZ, wrows[k]);  // columns of [CZ, up4(CP)]
      pb.copy(bsrc[k], wrows[k], bcat + woff[k], wrows[k], 1, wrows[k], 0);
    }
    pb.copy(WT(4), CZ, Wor, CZ, CZ, CZ);
    pb.transpose(WT(4), CZ, WoT, CZ, CZ, CZ, CZ, CZ);
    if (bw_prep(pb, st)) return 1;
  }
  // ---- recompute forward ----
  if (bw_ln_fwd(x, R, CZ, nullptr, nullptr, p, st)) return 1;
  {
    GemmArgs g = tfg((int)R, CP, CZ, p, CZ, Wcat, CZ, pre, CP);
    g.bias = bcat; g.round_tf32 = 0;
    if (gemm_f16(g, st)) return 1;
  }
  if (bw_trimul_ab(pre, CP, mask, B, N, C2, ab, st)) return 1;
  if (bw_rows_to_planes(ab, C2, 0, B, N, CZ, Np, pa, st)) return 1;
  if (bw_rows_to_planes(ab, C2, CZ, B, N, CZ, Np, pb, st)) return 1;
  const long long ps = (long long)N * Np;  // plane stride
  if (bw_rows_to_planes(ab, C2, 0, B, N, CZ, Np, paT, st, 1)) return 1;
  if (bw_rows_to_planes(ab, C2, CZ, B, N, CZ, Np, pbT, st, 1)) return 1;
  auto plane_gemm = [&](const float* A, const float* Bm, float* C) {
    GemmArgs g = tfg(N, N, N, A, Np, Bm, Np, C, Np);
    g.nb1 = B * CZ; g.a_bs1 = ps; g.b_bs1 = ps; g.c_bs1 = ps;
    return gemm_f16(g, st);
  };
  // outgoing x[i][j] = sum_k a[i][k] b[j][k];  incoming x[i][j] = sum_k a[k][i] b[k][j] = aT[i][k] bT[j][k]
  if (d->mode == 0 ? plane_gemm(pa, pb, xp) : plane_gemm(paT, pbT, xp)) return 1;
  if (bw_planes_to_rows(xp, B, N, CZ, Np, xc, CZ, 0, st)) return 1;
  if (bw_ln_fwd(xc, R, CZ, nullptr, nullptr, xn, st)) return 1;
  {
    GemmArgs g = tfg((int)R, CZ, CZ, xn, CZ, Wor, CZ, o, CZ);
    g.bias = WT(5); g.round_tf32 = 0;
    if (gemm_f16(g, st)) return 1;
  }
  // ---- backward: out = x + sigmoid(gate_o) * o ----
  if (bw_gate_bwd(dy, CZ, pre + 2 * C2, CP, o, CZ, d_o, CZ, dpre + 2 * C2, CP, R, CZ, st)) return 1;
  if (bw_dw_acc(d_o, CZ, xn, CZ, R, CZ, CZ, OUT(5), CZ, OUT(6), 1.f, st)) return 1;
  {
    GemmArgs g = tfg((int)R, CZ, CZ, d_o, CZ, WoT, CZ, dxn, CZ);
    if (gemm_f16(g, st)) return 1;
  }
  if (bw_ln_bwd(xc, dxn, R, CZ, nullptr, dxc, 0, nullptr, nullptr, st)) return 1;
  if (bw_rows_to_planes(dxc, CZ, 0, B, N, CZ, Np, dxp, st)) return 1;
  if (bw_rows_to_planes(dxc, CZ, 0, B, N, CZ, Np, dxpT, st, 1)) return 1;
  if (d->mode == 0) {
    // da[i][k] = sum_j dx[i][j] b[j][k] = dx . (bT)^T ;  db[j][k] = sum_i dx[i][j] a[i][k] = dxT . (aT)^T
    if (plane_gemm(dxp, pbT, dap)) return 1;
    if (plane_gemm(dxpT, paT, dbp)) return 1;
  } else {
    // da[k][i] = sum_j b[k][j] dx[i][j] = b . dx^T ;  db[k][j] = sum_i a[k][i] dx[i][j] = a . (dxT)^T
    if (plane_gemm(pb, dxp, dap)) return 1;
    if (plane_gemm(pa, dxpT, dbp)) return 1;
  }
  if (bw_planes_to_rows(dap, B, N, CZ, Np, dab, C2, 0, st)) return 1;
  if (bw_planes_to_rows(dbp, B, N, CZ, Np, dab, C2, CZ, st)) return 1;
  if (bw_trimul_ab_bwd(pre, CP, mask, B, N, C2, dab, dpre, CP, st)) return 1;
  if (bw_dw_acc(dpre, CP, p, CZ, R, C2, CZ, OUT(1), CZ, OUT(2), 1.f, st)) return 1;
  if (bw_dw_acc(dpre + C2, CP, p, CZ, R, C2, CZ, OUT(3), CZ, OUT(4), 1.f, st)) return 1;
  if (bw_dw_acc(dpre + 2 * C2, CP, p, CZ, R, CZ, CZ, OUT(7), CZ, OUT(8), 1.f, st)) return 1;
  {
    GemmArgs g = tfg((int)R, CZ, CP, dpre, CP, WcatT, up4(CP), dp, CZ);
    if (gemm_f16(g, st)) return 1;
  }
  return bw_ln_bwd(x, dp, R, CZ, nullptr, dy, 1, nullptr, nullptr, st);
}

// modules.py:283-287,337 OuterLinear.  pair_out = pair + W [x_i * x_j | x_i - x_j] + b with x = LN(single).
// in: [single_in | d_pair (read only)]   out: [d_single io (+=) | dW c_z x 2c_s | db]   weights: [linear.weight c_z x 2c_s]
PRD_BWD_OP(outer_linear) {
  const int B = d->B, N = d->N, CZ = d->c_z, CS = d->c_s, Np = up4(N), M = B * N;
  float* s = c.take((size_t)M * CS);
  float* sT = c.take((size_t)B * CS * Np);
  float* ET = c.take((size_t)M * CZ * Np);
  float* T = c.take((size_t)M * CZ * CS);
  float* U = c.take((size_t)M * CZ);
  float* ds = c.take((size_t)M * CS);
  float* W2T = c.take((size_t)CS * CZ);
  if (dry) return 0;
  const float *single = IN(0), *dy = IN(1), *W = WT(0);
  if (bw_ln_fwd(single, M, CS, nullptr, nullptr, s, st)) return 1;
  if (bw_colsum(dy, CZ, (long long)M * N, CZ, OUT(2), 1.f, st)) return 1;
  if (bw_transpose(s, CS, (long long)N * CS, sT, Np, (long long)CS * Np, N, CS, B, 1.f, st)) return 1;
  if (bw_pair_to_izj(dy, B, N, CZ, Np, 0, 1, ET, st)) return 1;  // dy + dy^T(i <-> j)
  {
    GemmArgs g = tfg(CZ, CS, N, ET, Np, sT, Np, T, CS);
    g.nb1 = N; g.nb2 = B;
    g.a_bs1 = (long long)CZ * Np; g.a_bs2 = (long long)N * CZ * Np;
    g.b_bs1 = 0; g.b_bs2 = (long long)CS * Np;
    g.c_bs1 = (long long)CZ * CS; g.c_bs2 = (long long)N * CZ * CS;
    g.round_tf32 = 0;
    if (gemm_f16(g, st)) return 1;
  }
  // product term: W[:, :CS];  the symmetrised T counts every (i, j) pair twice for dW
  if (bw_bilinear_reduce(T, W, 2 * CS, s, B, N, CZ, CS, ds, 0, OUT(1), 1.f, 0.5f, st)) return 1;
  // difference term: W[:, CS:]
  if (bw_pair_rowcol_diff(dy, B, N, CZ, U, st)) return 1;
  if (bw_transpose(W + CS, 2 * CS, 0, W2T, CZ, 0, CZ, CS, 1, 1.f, st)) return 1;  // [CS, CZ]
  {
    GemmArgs g = tfg(M, CS, CZ, U, CZ, W2T, CZ, ds, CS);
    g.add = ds; g.ldadd = CS;
    if (gemm_f16(g, st)) return 1;
  }
  if (bw_dw_acc(U, CZ, s, CS, M, CZ, CS, OUT(1) + CS, 2 * CS, nullptr, 1.f, st)) return 1;
  return bw_ln_bwd(single, ds, M, CS, nullptr, OUT(0), 1, nullptr, nullptr, st);
}

// model.py:364-373 coordinate head (+ modules.py:403 symmetrisation, utils.py:32-36 remove_mean).
// in: [pair (before symmetrisation) | z | mask | d_noise_pred B,N,3]   out: [d_pair (written) | dW1 | db1 | dw2 c_z]
// weights: [weight_radial.1.weight | weight_radial.1.bias | weight_radial.3.weight (c_z)]
PRD_BWD_OP(coord_head) {
  const int B = d->B, N = d->N, CZ = d->c_z;
  const long long R = (long long)B * N * N;
  int rc = 0;
  float* ps = c.take((size_t)R * CZ);
  const float* xh = nullptr;
  long long ldxh = 0;
  float* h = c.take((size_t)R * CZ);
  float* dh = c.take((size_t)R * CZ);
  float* dxh = c.take((size_t)R * CZ);
  float* d_eps = c.take((size_t)B * N * 3);
  float* W1T = transposed(dry, c, st, WT(0), CZ, CZ, CZ, &rc);
  if (!dry) {
    if (rc) return 1;
    PRD_CUDA_OK(cudaMemcpyAsync(ps, IN(0), (size_t)R * CZ * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (symmetrize_pair(PairDims{B, N, CZ}, ps, st)) return 1;
  }
  if (relu_layer_fwd(dry, c, st, R, CZ, CZ, ps, WT(0), WT(1), &xh, &ldxh, h)) return 1;
  if (dry) return 0;
  if (bw_remove_mean_adj(IN(3), IN(2), B, N, d_eps, st)) return 1;
  if (bw_coord_dh(h, IN(1), IN(2), d_eps, WT(2), B, N, CZ, dh, OUT(3), st)) return 1;
  if (bw_dw_acc(dh, CZ, xh, ldxh, R, CZ, CZ, OUT(1), CZ, OUT(2), 1.f, st)) return 1;
  {
    GemmArgs g = tfg((int)R, CZ, CZ, dh, CZ, W1T, CZ, dxh, CZ);
    if (gemm_f16(g, st)) return 1;
  }
  float* dpair = OUT(0);
  if (bw_ln_bwd(ps, dxh, R, CZ, nullptr, dpair, 0, nullptr, nullptr, st)) return 1;
  if (symmetrize_pair(PairDims{B, N, CZ}, dpair, st)) return 1;   // adjoint of 0.5 (p + p^T) is itself
  return bw_copy2d(dpair, CZ, dpair, CZ, R, CZ, st);              // rounded: it is a tf32 operand next
}

// AF2_modules.py:421-473 SPAttention: out = LN_a(single) + mha(LN_a(single), bias(pair)).
// in: [single_in | pair]
// out: [d_single io (d out -> d in) | d_pair (+=) | d ln_m.w | d ln_m.b | d ln_z.w | d ln_z.b | dWz | dWq | dWk | dWv | dWg | dbg | dWo | dbo]
// weights: [ln_m.w | ln_m.b | ln_z.w | ln_z.b | Wz H x c_z | Wq | Wk | Wv | Wg | bg | Wo | bo]
PRD_BWD_OP(spattention) {
  const int B = d->B, N = d->N, CS = d->c_s, H = d->H, HC = H * CS, M = B * N, Np = up4(N);
  int rc = 0;
  const size_t MH = (size_t)M * HC, SS = (size_t)B * H * N * Np, TT = (size_t)B * HC * Np;
  float* xln = c.take((size_t)M * CS);
  float* q = c.take(MH);
  float* k = c.take(MH);
  float* v = c.take(MH);
  float* gpre = c.take(MH);
  float* bias = c.take((size_t)B * H * N * N);
  float* Sm = c.take(SS);
  float* P = c.take(SS);
  float* PT = c.take(SS);
  float* dP = c.take(SS);
  float* dS = c.take(SS);
  float* dST = c.take(SS);
  float* vT = c.take(TT);
  float* kT = c.take(TT);
  float* qT = c.take(TT);
  float* dOT = c.take(TT);
  float* O = c.take(MH);
  float* og = c.take(MH);
  float* d_og = c.take(MH);
  float* dO = c.take(MH);
  float* dgpre = c.take(MH);
  float* dq = c.take(MH);
  float* dk = c.take(MH);
  float* dv = c.take(MH);
  float* dxln = c.take((size_t)M * CS);
  float* Wr[4];
  float* WTt[4];
  for (int i = 0; i < 4; ++i) {
    Wr[i] = rounded(dry, c, st, WT(5 + i), HC, CS, CS, &rc);
    WTt[i] = transposed(dry, c, st, WT(5 + i), HC, CS, CS, &rc);  // [CS, HC]
  }
  float* WoT = transposed(dry, c, st, WT(10), CS, HC, HC, &rc);   // Wo [CS, HC] -> [HC, CS]
  if (dry) return 0;
  if (rc) return 1;
  const float *single = IN(0), *pair = IN(1);
  const float alpha = 1.0f / sqrtf((float)CS);
  if (bw_ln_fwd(single, M, CS, WT(0), WT(1), xln, st)) return 1;
  float* outs[4] = {q, k, v, gpre};
  for (int i = 0; i < 4; ++i) {
    GemmArgs g = tfg(M, HC, CS, xln, CS, Wr[i], CS, outs[i], HC);
    if (i == 0) g.alpha = alpha;
    if (i == 3) { g.bias = WT(9); g.round_tf32 = 0; }
    if (gemm_f16(g, st)) return 1;
  }
  if (pair_bias_proj(PairDims{B, N, d->c_z}, H, pair, WT(2), WT(3), WT(4), nullptr, bias, st)) return 1;
  auto head_batches = [&](GemmArgs& g, long long a1, long long a2, long long b1, long long b2, long long c1, long long c2) {
    g.nb1 = H; g.nb2 = B;
    g.a_bs1 = a1; g.a_bs2 = a2; g.b_bs1 = b1; g.b_bs2 = b2; g.c_bs1 = c1; g.c_bs2 = c2;
  };
  const long long sNN = (long long)N * Np;
  {  // S[b,h] = q_h k_h^T + bias
    GemmArgs g = tfg(N, N, CS, q, HC, k, HC, Sm, Np);
    head_batches(g, CS, (long long)N * HC, CS, (long long)N * HC, sNN, (long long)H * sNN);
    g.add = bias; g.ldadd = N; g.add_bs1 = (long long)N * N; g.add_bs2 = (long long)H * N * N;
    g.round_tf32 = 0;
    if (gemm_f16(g, st)) return 1;
  }
  if (bw_softmax_rows(Sm, P, (long long)B * H * N, N, Np, st)) return 1;
  if (bw_transpose(v, HC, (long long)N * HC, vT, Np, (long long)HC * Np, N, HC, B, 1.f, st)) return 1;
  {  // O[b,:,h,:] = P[b,h] V[b,h]
    GemmArgs g = tfg(N, CS, N, P, Np, vT, Np, O, HC);
    head_batches(g, sNN, (long long)H * sNN, (long long)CS * Np, (long long)HC * Np, CS, (long long)N * HC);
    g.round_tf32 = 0;
    if (gemm_f16(g, st)) return 1;
  }
  if (bw_gate_fwd(gpre, HC, O, HC, og, HC, M, HC, st)) return 1;
  // ---- backward ----
  float* dy = OUT(0);
  if (bw_dw_acc(dy, CS, og, HC, M, CS, HC, OUT(12), HC, OUT(13), 1.f, st)) return 1;
  {
    GemmArgs g = tfg(M, HC, CS, dy, CS, WoT, CS, d_og, HC);
    if (gemm_f16(g, st)) return 1;
  }
  if (bw_gate_bwd(d_og, HC, gpre, HC, O, HC, dO, HC, dgpre, HC, M, HC, st)) return 1;
  {  // dP[b,h] = dO_h V_h^T
    GemmArgs g = tfg(N, N, CS, dO, HC, v, HC, dP, Np);
    head_batches(g, CS, (long long)N * HC, CS, (long long)N * HC, sNN, (long long)H * sNN);
    g.round_tf32 = 0;
    if (gemm_f16(g, st)) return 1;
  }
  if (bw_softmax_bwd_rows(P, dP, dS, (long long)B * H * N, N, Np, st)) return 1;
  if (bw_transpose(P, Np, sNN, PT, Np, sNN, N, N, B * H, 1.f, st)) return 1;
  if (bw_transpose(dS, Np, sNN, dST, Np, sNN, N, N, B * H, 1.f, st)) return 1;
  if (bw_transpose(dO, HC, (long long)N * HC, dOT, Np, (long long)HC * Np, N, HC, B, 1.f, st)) return 1;
  if (bw_transpose(k, HC, (long long)N * HC, kT, Np, (long long)HC * Np, N, HC, B, 1.f, st)) return 1;
  if (bw_transpose(q, HC, (long long)N * HC, qT, Np, (long long)HC * Np, N, HC, B, 1.f, st)) return 1;
  {  // dV_h = P_h^T dO_h
    GemmArgs g = tfg(N, CS, N, PT, Np, dOT, Np, dv, HC);
    head_batches(g, sNN, (long long)H * sNN, (long long)CS * Np, (long long)HC * Np, CS, (long long)N * HC);
    if (gemm_f16(g, st)) return 1;
  }
  {  // d(x Wq^T) = alpha dS_h K_h
    GemmArgs g = tfg(N, CS, N, dS, Np, kT, Np, dq, HC);
    head_batches(g, sNN, (long long)H * sNN, (long long)CS * Np, (long long)HC * Np, CS, (long long)N * HC);
    g.alpha = alpha;
    if (gemm_f16(g, st)) return 1;
  }
  {  // dK_h = dS_h^T Q_h (Q already carries alpha)
    GemmArgs g = tfg(N, CS, N, dST, Np, qT, Np, dk, HC);
    head_batches(g, sNN, (long long)H * sNN, (long long)CS * Np, (long long)HC * Np, CS, (long long)N * HC);
    if (gemm_f16(g, st)) return 1;
  }
  float* dacts[4] = {dq, dk, dv, dgpre};
  for (int i = 0; i < 4; ++i) {
    if (bw_dw_acc(dacts[i], HC, xln, CS, M, HC, CS, OUT(7 + i), CS, i == 3 ? OUT(11) : nullptr, 1.f, st)) return 1;
    GemmArgs g = tfg(M, CS, HC, dacts[i], HC, WTt[i], HC, dxln, CS);
    g.add = i == 0 ? dy : dxln; g.ldadd = CS;   // residual on LN_a(single): dxln starts from dy
    if (gemm_f16(g, st)) return 1;
  }
  if (bw_ln_bwd(single, dxln, M, CS, WT(0), OUT(0), 0, OUT(2), OUT(3), st)) return 1;
  return bw_pair_bias_bwd(B, N, d->c_z, H, pair, dS, Np, WT(4), WT(2), WT(3), OUT(1), OUT(6), nullptr, OUT(4), OUT(5), st);
}

// AF2_modules.py:519-530 OuterProductUpdate projections a = mask (W1 LN_a(s) + b1), b likewise.
// in: [single_in | mask | d_a B,N,c_s/4 | d_b]   out: [d_single io (+=) | d ln.w | d ln.b | dW1 | db1 | dW2 | db2]
// weights: [ln.w | ln.b | W1 | b1 | W2 | b2]
PRD_BWD_OP(opm_project) {
  const int M = d->B * d->N, CS = d->c_s, OD = d->c_s / 4;
  int rc = 0;
  float* xln = c.take((size_t)M * CS);
  float* dam = c.take((size_t)M * OD);
  float* dbm = c.take((size_t)M * OD);
  float* dxln = c.take((size_t)M * CS);
  float* W1T = transposed(dry, c, st, WT(2), OD, CS, CS, &rc);  // [CS, OD]
  float* W2T = transposed(dry, c, st, WT(4), OD, CS, CS, &rc);
  if (dry) return 0;
  if (rc) return 1;
  if (bw_ln_fwd(IN(0), M, CS, WT(0), WT(1), xln, st)) return 1;
  if (bw_scale_rows(IN(2), OD, IN(1), 1.f, dam, OD, M, OD, st)) return 1;
  if (bw_scale_rows(IN(3), OD, IN(1), 1.f, dbm, OD, M, OD, st)) return 1;
  if (bw_dw_acc(dam, OD, xln, CS, M, OD, CS, OUT(3), CS, OUT(4), 1.f, st)) return 1;
  if (bw_dw_acc(dbm, OD, xln, CS, M, OD, CS, OUT(5), CS, OUT(6), 1.f, st)) return 1;
  {
    GemmArgs g = tfg(M, CS, OD, dam, OD, W1T, OD, dxln, CS);
    if (gemm_f16(g, st)) return 1;
    g = tfg(M, CS, OD, dbm, OD, W2T, OD, dxln, CS);
    g.add = dxln; g.ldadd = CS;
    if (gemm_f16(g, st)) return 1;
  }
  return bw_ln_bwd(IN(0), dxln, M, CS, WT(0), OUT(0), 1, OUT(1), OUT(2), st);
}

// model.py:348-361 + AF2_modules.py:532-543 + modules.py:395-397: everything that writes the initial pair tensor.
// in: [d_pair (read only) | z | mask | t i64 B | opm_a | opm_b | atom_mask | residue_mask | bond_mask | bond_feats i64 |
//      bond_distance i64 | residue_index i64 | residue_chain_index i64]
// out: [d_opm_a (written) | d_opm_b (written) | dW_opm c_z x c_s/4 | db_opm | dW_dist c_z x dist | dW_beta c_z x time |
//       d bond table 0 | 1 | 2 | d bond-distance table | d relpos table]
// weights: [W_opm (linear_out.weight) | centers | freq]
PRD_BWD_OP(pair_embed) {
  const int B = d->B, N = d->N, CZ = d->c_z, OD = d->c_s / 4, Np = up4(N), M = B * N, DD = d->dist_dim;
  const long long R = (long long)B * N * N;
  float* Dm = c.take((size_t)R * CZ);
  float* ET = c.take((size_t)M * CZ * Np);
  float* oT = c.take((size_t)B * OD * Np);
  float* T = c.take((size_t)M * CZ * OD);
  float* rbf = c.take((size_t)R * DD);
  float* colsum = c.take((size_t)B * CZ);
  if (dry) return 0;
  const float *dpair = IN(0), *mask = IN(2);
  const float inv = 1.0f / 1.001f;  // m2 / (m2 + 1e-3) on valid pairs (AF2_modules.py:539-543, modules.py:395)
  if (bw_mask_pair(dpair, mask, B, N, CZ, 1.f, Dm, st)) return 1;
  if (bw_colsum(Dm, CZ, R, CZ, OUT(3), inv, st)) return 1;
  auto bilinear = [&](int transpose_ij, const float* other, const float* self, float* d_self, float* dW) {
    if (bw_pair_to_izj(Dm, B, N, CZ, Np, transpose_ij, 0, ET, st)) return 1;
    if (bw_transpose(other, OD, (long long)N * OD, oT, Np, (long long)OD * Np, N, OD, B, 1.f, st)) return 1;
    GemmArgs g = tfg(CZ, OD, N, ET, Np, oT, Np, T, OD);
    g.nb1 = N; g.nb2 = B;
    g.a_bs1 = (long long)CZ * Np; g.a_bs2 = (long long)N * CZ * Np;
    g.b_bs1 = 0; g.b_bs2 = (long long)OD * Np;
    g.c_bs1 = (long long)CZ * OD; g.c_bs2 = (long long)N * CZ * OD;
    g.round_tf32 = 0;
    if (gemm_f16(g, st)) return 1;
    return bw_bilinear_reduce(T, inf(w, 0), OD, self, B, N, CZ, OD, d_self, 0, dW, inv, inv, st);
  };
  // y_ij = W (a_i * b_j):  d a_i = sum_j (W^T E_ij) * b_j  (and dW);  d b_j = sum_i (W^T E_ij) * a_i
  if (bilinear(0, IN(5), IN(4), OUT(0), OUT(2))) return 1;
  if (bilinear(1, IN(4), nullptr, OUT(1), nullptr)) return 1;
  // distance embedding weight: dW_dist[z, k] += sum m2 d_pair[z] rbf_k(|z_i - z_j|)
  if (bw_rbf_rows(IN(1), WT(1), (DD - 1) / 2.0f, B, N, DD, rbf, st)) return 1;
  if (bw_dw_acc(Dm, CZ, rbf, DD, R, CZ, DD, OUT(4), DD, nullptr, 1.f, st)) return 1;
  if (bw_time_embed_bwd(Dm, static_cast<const int64_t*>(in[3]), d->num_steps, WT(2), B, N, CZ, d->time_dim, colsum, OUT(5), st))
    return 1;
  return bw_pair_static_bwd(dpair, IN(6), IN(7), IN(8), static_cast<const int64_t*>(in[9]), static_cast<const int64_t*>(in[10]),
                            static_cast<const int64_t*>(in[11]), static_cast<const int64_t*>(in[12]), B, N, CZ,
                            d->max_bond_distance, d->max_relpos, OUT(6), OUT(7), OUT(8), OUT(9), OUT(10), st);
}

// model.py:342-346 single embedding (+ :99-102 ESM projection).
// in: [d_single | atom_feats i64 | atom_mask | residue_mask | seq_t | residue_esm]
// out: [d atom table 0..8 | dW_type c_s x 21 | dW_esm c_s x esm]     weights: [w_type c_s x 21]
PRD_BWD_OP(single_embed) {
  const int M = d->B * d->N, CS = d->c_s, ED = d->esm_dim;
  float* d_ty = c.take((size_t)M * CS);
  float* d_esm = c.take((size_t)M * CS);
  float* lnseq = c.take((size_t)M * 21);
  float* esm_ln = c.take((size_t)M * ED);
  if (dry) return 0;
  AtomGradTables tabs;
  for (int f = 0; f < 9; ++f) tabs.t[f] = outf(out, f);
  if (bw_single_embed_bwd(IN(0), static_cast<const int64_t*>(in[1]), IN(2), IN(3), IN(4), WT(0), d->B, d->N, CS, tabs, d_ty, d_esm,
                          lnseq, st))
    return 1;
  if (bw_dw_acc(d_ty, CS, lnseq, 21, M, CS, 21, OUT(9), 21, nullptr, 1.f, st)) return 1;
  if (bw_ln_fwd(IN(5), M, ED, nullptr, nullptr, esm_ln, st)) return 1;
  return bw_dw_acc(d_esm, CS, esm_ln, ED, M, CS, ED, OUT(10), ED, nullptr, 1.f, st);
}

}  // extern "C"
