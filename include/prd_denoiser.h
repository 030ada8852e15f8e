/* libprd_sm100 -- C ABI of the B200-native ProteinReDiff denoiser hot path.
 *
 * The reference (HySonLab/Protein_Redesign) has no FFI: its "operator API" is the Python
 * nn.Module surface of ProteinReDiff/modules.py, ProteinReDiff/models/AF2_modules.py and
 * ProteinReDiff/model.py.  Every entry point below replaces the forward of one of those
 * modules (file:line cited per function); the Python mirror in protein_redesign_b200/ binds
 * them with ctypes.
 *
 * Conventions
 *   - all data pointers are DEVICE pointers (tensor.data_ptr()), contiguous row-major with the
 *     reference's shapes; fp32 activations, int64 indices, fp16 ("_h") packed weights;
 *   - the caller owns every buffer including the workspace; nothing is allocated, freed or
 *     synchronised inside, every launch goes to `stream` (a cudaStream_t), so a sequence of
 *     calls is CUDA-graph capturable;
 *   - return 0 on success, non-zero on failure with a thread-local message in prd_last_error();
 *   - uniform signature for the module-level ops:
 *       int prd_<op>_fwd(const PrdDims*, const void* const* in, void* const* out,
 *                        const void* const* weights, void* workspace, size_t workspace_bytes,
 *                        void* stream);
 *       size_t prd_<op>_workspace_bytes(const PrdDims*);
 *     the pointer-array order is documented per function;
 *   - there is no CPU fallback: without an sm_100 device every op fails (prd_device_check());
 *   - contract limits (the reference's defaults; anything else returns an error, never a silent fallback):
 *       attention heads are H = 4 heads of c = 16 channels (modules.py:148-149 defaults) in triangle_attention and
 *       single_attention; pair_dim c_z is 64 or 32; single_dim c_s is a multiple of 64; SPAttention uses c_hidden = c_s
 *       (modules.py:366-371) and OuterProductUpdate c_hidden = c_s / 4 (modules.py:372-374); RadialBasisProjection spans
 *       [0, 2] nm (modules.py:73-82);
 *     pair masks: the pair-stack ops take the TOKEN mask m [B, N]; the reference's mask_2d is always m (x) m
 *     (modules.py:334,393) and the Python mirror rejects any other [B, N, N] mask (modules._mask_from_2d).
 */
#ifndef PRD_DENOISER_H_
#define PRD_DENOISER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PRD_VERSION 2

typedef struct PrdDims {
  int32_t B;                 /* batch rows                                  */
  int32_t N;                 /* tokens per row (ligand atoms + residues + pad) */
  int32_t c_s;               /* single_dim  (model.py:145)                  */
  int32_t c_z;               /* pair_dim    (model.py:146); built: 64, 32   */
  int32_t H;                 /* num_heads   (model.py:148); built: 4        */
  int32_t c;                 /* head_dim    (model.py:147); built: 16       */
  int32_t tf;                /* transition_factor (model.py:149)            */
  int32_t esm_dim;           /* model.py:142                                */
  int32_t time_dim;          /* model.py:143                                */
  int32_t dist_dim;          /* model.py:144                                */
  int32_t max_bond_distance; /* model.py:151                                */
  int32_t max_relpos;        /* model.py:152                                */
  int32_t num_steps;         /* model.py:153                                */
  int32_t mode;              /* op specific: 0 = outgoing/starting, 1 = incoming/ending */
  int32_t residual;          /* residual ops: 1 -> dst = src + update (dst may alias src, the fused
                                FoldingBlock form, modules.py:335-342); 0 -> dst = update only (what the
                                reference sub-module's forward returns) */
} PrdDims;

int prd_version(void);
const char* prd_last_error(void);
int prd_device_check(void); /* 0 iff the current device is sm_100 */
long long prd_launch_count(void); /* kernels launched by this library so far in this process */

/* Plain batched GEMM on the tcgen05 path (building block + test hook):
 * C[b] = alpha * A[b] (M x K, fp16) * B[b]^T (N x K, fp16), fp32 accumulate, optional epilogue. */
typedef struct PrdGemm {
  int32_t M, N, K, nb1, nb2;
  const void* A; int64_t lda, a_bs1, a_bs2;
  const void* B; int64_t ldb, b_bs1, b_bs2;
  float alpha; int32_t act;                    /* 0 none, 1 relu, 2 sigmoid */
  const float* bias;                           /* [N] or NULL */
  const float* rowscale; int64_t rs_bs1, rs_bs2;
  const float* mul; int64_t ldmul, mul_bs1, mul_bs2;
  const float* add; int64_t ldadd, add_bs1, add_bs2;
  void* C; int64_t ldc, c_bs1, c_bs2; int32_t c_fp16;
  int32_t tf32;       /* 1: A, B are fp32 in memory and multiplied on kind::tf32 (lda / ldb / strides in floats); C fp32 */
  int32_t mul_step;   /* 1: `mul` gates instead of scaling: v = mul > 0 ? v : 0 (ReLU backward) */
  int32_t round_tf32; /* 1: fp32 result rounded to nearest tf32 (it is the operand of a following tf32 GEMM) */
} PrdGemm;
int prd_gemm_f16(const PrdGemm* g, void* stream);

#define PRD_DECLARE_OP(name)                                                                      \
  int prd_##name##_fwd(const PrdDims* d, const void* const* in, void* const* out,                 \
                       const void* const* weights, void* workspace, size_t workspace_bytes,       \
                       void* stream);                                                             \
  size_t prd_##name##_workspace_bytes(const PrdDims* d);

/* --- input embeddings -------------------------------------------------------------------- */
/* model.py:99-102 embed_residue_esm (LN(esm_dim) -> Linear no bias); step invariant.
 * in: [residue_esm f32 B,N,esm]   out: [esm_emb f32 B,N,c_s]   weights: [w_esm_h c_s x esm] */
PRD_DECLARE_OP(esm_embed)
/* model.py:342-346 (+ modules.py:35-51, model.py:89-93).
 * in: [atom_feats i64 B,N,9 | atom_mask | residue_mask | seq_t f32 B,N,21 | esm_emb]
 * out: [single f32 B,N,c_s]   weights: [atom table 0..8 f32 | w_type f32 c_s x 21] */
PRD_DECLARE_OP(single_embed)
/* model.py:348-358 step-invariant pair terms (modules.py:54-70 BondEmbedding, bond distance, relpos).
 * in: [atom_mask | residue_mask | bond_mask B,N,N | bond_feats i64 B,N,N,3 | bond_distance i64 B,N,N |
 *      residue_index i64 B,N | residue_chain_index i64 B,N]
 * out: [pair_static f32 B,N,N,c_z]   weights: [bond table 0..2 | bond_distance table | relpos table] */
PRD_DECLARE_OP(pair_embed_static)
/* AF2_modules.py:519-530 OuterProductUpdate projections a = mask*(W1 LN_a(s)+b1), b likewise.
 * in: [single | mask B,N]  out: [opm_a f32 B,N,c_s/4 | opm_b]
 * weights: [ln_w | ln_b | w1_h | b1 | w2_h | b2] */
PRD_DECLARE_OP(opm_project)
/* model.py:337-341,359-361 + modules.py:73-97 + AF2_modules.py:532-543 + modules.py:395-397:
 * pair = pair_static + m2 (W_dist rbf(|z_i-z_j|) + W_beta sincos(t/T)) + m2 (W_o(a_i*b_j)+b_o)/(m2+1e-3).
 * in: [pair_static (NULL = zeros) | z f32 B,N,3 | mask | t i64 B (or NULL with sampler_state) | opm_a | opm_b | sampler_state]
 * d->mode bit 0: OuterProductUpdate term only (Denoiser.forward's `pair += mask_2d * opm(...)`, modules.py:395-397);
 *         bit 1: OPM term not multiplied by mask_2d (stand-alone OuterProductUpdate.forward, AF2_modules.py:503-545).
 * out: [pair]   weights: [freq | w_beta f32 c_z x time | w_dist_h c_z x dist | centers | w_opm_h c_z x c_s/4 | b_opm |
 *                        rbf_lut f32 (or NULL)]  -- ALWAYS seven entries.
 * rbf_lut (built by prd_rbf_lut_build): W_dist rbf(d) is a function of one scalar, so it may be tabulated once per weight
 * version (fp32, PRD_RBF_LUT_POINTS + 1 rows over [0, d_max]) and interpolated linearly (3e-6 relative) instead of being
 * rebuilt per pair as dist_dim exponentials + a K = dist_dim GEMM; NULL keeps the GEMM form (modules.py:73-82 literally). */
PRD_DECLARE_OP(pair_embed)
#define PRD_RBF_LUT_POINTS 8192
size_t prd_rbf_lut_floats(const PrdDims* d); /* (PRD_RBF_LUT_POINTS + 2) * c_z: header row {1/h, M} + table rows */
/* w_dist f32 [c_z x dist_dim] (embed_dist.1.weight), centers f32 [dist_dim] (embed_dist.0.center); d_max: distances
 * beyond it embed to zero (use max center + 0.52 nm: every term is below 1e-15 there). */
int prd_rbf_lut_build(const PrdDims* d, const float* w_dist, const float* centers, float d_max, float* lut, void* stream);

/* --- Denoiser trunk (modules.py:391-404) ------------------------------------------------- */
/* The pair-bias projections of the single-representation attentions as ONE stream over the pair tensor:
 * bias_a[b,h,i,j] = (LN(pair[b,i,j,:]) * ln_w_a + ln_b_a) . w_a[h,:] + bvec_a[h]  (SPAttention.linear_z, AF2_modules.py:454-459) and,
 * optionally, bias_b likewise (the first FoldingBlock's attn_bias, modules.py:300-304: both read the pair tensor the embedding
 * leaves, SPAttention only updates the single representation).  NULL affine / bias vectors mean identity / zero.
 * in: [pair]   out: [bias_a f32 B,H,N,N | bias_b (or NULL)]
 * weights: [ln_w_a | ln_b_a | w_a f32 H x c_z | bvec_a | ln_w_b | ln_b_b | w_b | bvec_b]  -- ALWAYS eight entries */
PRD_DECLARE_OP(pair_bias)
/* AF2_modules.py:421-473 SPAttention (+ :251-367, :613-627): single <- LN_a(single) + mha(...).
 * in: [single | pair | bias f32 B,H,N,N precomputed by prd_pair_bias_fwd (or NULL: projected here)] -- ALWAYS three entries
 * out: [single_out (may alias single)]
 * weights: [ln_m_w | ln_m_b | ln_z_w | ln_z_b | w_z f32 H x c_z | w_q_h | w_k_h | w_v_h | w_g_h | b_g | w_o_h | b_o] */
PRD_DECLARE_OP(spattention)
/* modules.py:300-304 attn_bias + modules.py:185-225 Attention on the single rep + residual (:335).
 * in: [single | pair (or NULL) | mask | attn_bias f32 B,H,N,N (used when pair is NULL; NULL = no bias)]  out: [single_out]
 * weights: [w_bias f32 H x c_z | b_bias | w_qkvg_h 4Hc x c_s | b_qkvg f32 4Hc (zeros | gate bias) | w_o_h c_s x Hc | b_o] */
PRD_DECLARE_OP(single_attention)
/* modules.py:306-311,336 single_fc + residual.  in: [single]  out: [single_out]  weights: [w1_h | b1 | w2_h | b2] */
PRD_DECLARE_OP(single_transition)
/* modules.py:283-287,337 OuterLinear + residual.  in: [single | pair]  out: [pair_out]
 * weights: [w_mul_h c_z x c_s (= weight[:, :c_s]) | w_sub_h c_z x c_s (= weight[:, c_s:]) | bias] */
PRD_DECLARE_OP(outer_linear)
/* modules.py:262-274,338-339 TriangleMultiplication + residual (d->mode: 0 outgoing, 1 incoming).
 * in: [pair | mask]  out: [pair_out]
 * weights: [w_in_h 4c_z x c_z (ab_proj ; ab_gate) | b_in | w_out_h 2c_z x c_z (out_gate ; out_proj) | b_out] */
PRD_DECLARE_OP(triangle_multiplication)
/* modules.py:236-243 (+185-225),340-341 TriangleAttention + residual (d->mode bit 0: 0 starting, 1 ending; bit 1: the
 * caller promises that the token mask is all ones -- a performance hint that selects the attention core without the
 * per-sequence handling of ragged batches; results are identical either way).
 * in: [pair | mask]  out: [pair_out]
 * weights: [w_qkvg_h 4Hc x c_z (q;k;v;gate) | b_gate f32 Hc | w_o_h c_z x Hc | b_o] */
PRD_DECLARE_OP(triangle_attention)
/* modules.py:321-326,342 pair_fc + residual.  in: [pair]
 * out: [pair_out | bias_next f32 B,H,N,N (or NULL)] -- ALWAYS two entries; bias_next = the NEXT FoldingBlock's attn_bias
 *      (modules.py:300-304) of the updated rows, emitted from the output epilogue (pair_dim 64) so that the pair tensor is not read
 *      again for it; pass it to prd_single_attention_fwd as in[3] with pair = NULL
 * weights: [w1_h | b1 | w2_h | b2 | w_bias f32 H x c_z (or NULL) | b_bias f32 H (or NULL)] -- ALWAYS six entries */
PRD_DECLARE_OP(pair_transition)
/* modules.py:403  pair <- 0.5 (pair + pair^T).  out: [pair] */
PRD_DECLARE_OP(symmetrize)

/* --- heads (model.py:364-374) -------------------------------------------------------------- */
/* model.py:364-373: symmetrise + weight_radial + sum_j m2 w r + remove_mean.
 * in: [pair | z | mask]  out: [noise_pred f32 B,N,3]  weights: [w1_h c_z x c_z | b1 | w2 f32 c_z] */
PRD_DECLARE_OP(coord_head)
/* model.py:374 seq_mlp.  in: [single]  out: [seq_pred f32 B,N,21]  weights: [w1_h | b1 | w2_h 21 x c_s] */
PRD_DECLARE_OP(seq_head)

/* --- sampler (model.py:395-420) ------------------------------------------------------------ */
/* utils.py:32-36 remove_mean over rows; d->mode = channel count (3 or 21); x has shape [B, N, mode].
 * in: [mask (B rows, broadcast cyclically if x has k*B rows; d->B = rows of x, d->H = rows of mask)]  out: [x] */
PRD_DECLARE_OP(remove_mean)
/* model.py:407-420 one reverse step on device (see csrc/prd_embed.cu).
 * in: [noise_pred | seq_pred | noise f32 steps,B,N,3 | coef f32 T,3]  out: [z | seq_t | sampler_state int32[2]] */
PRD_DECLARE_OP(sampler_update)

/* --- training objective (model.py:471-549) ----------------------------------- */
/* model.py:471-488 q(): z_t = sa[t] x + s1[t] noise_z; seq_t = keep*seq + drop*(sa[t] seq + s1[t] noise_seq);
 * seq_t1 = sa[t1] seq + s1[t1] noise_seq with t1 = max(t-1, 0).  d->num_steps = rows of sched.
 * in: [x f32 B,N,3 | seq (residue_one_hot) f32 B,N,21 | t i64 B | noise_z | noise_seq | keep (residue_extra_mask) B,N |
 *      drop (residue_inv_extra_mask) B,N | sched f32 T,2 {sqrt_alphas_cumprod, sqrt_one_minus_alphas_cumprod}]
 * out: [z_t | seq_t | seq_t1] */
PRD_DECLARE_OP(diffusion_q)
/* model.py:499-526 (the three loss terms after the network call) + :538-541 (loss = mean(diff_loss / num_nodes)),
 * and d loss / d noise_pred, d loss / d seq_pred (what autograd hands to the network's backward).
 * in: [noise_pred | seq_pred | noise_z | noise_seq | seq_t1 | mask (residue_and_atom_mask) | residue_mask |
 *      residue_type i64 B,N | t i64 B | sched]
 * out: [loss f32 1 | diff_loss f32 B | terms f32 B+2 {mse_b.., KL, CE} (or NULL) | d_noise_pred (or NULL) |
 *       d_seq_pred (or NULL)] */
PRD_DECLARE_OP(diffusion_loss)

/* --- output post-processing -------------------------------------------------------- */
/* generate.py:76-91 predict_seq / update_seq: tokens = argmax softmax(logits) (0 = 'X' at masked positions).
 * in: [logits f32 B,N,21 | residue_mask B,N (or NULL)]   out: [tokens i64 B,N] */
PRD_DECLARE_OP(decode_argmax)
/* generate.py:176-195 + tmalign.py:23-49: rigid superposition of every sample onto a reference (d->mode = rows of the
 * reference: 1 or B), RMSD and TM-score (normalised by the reference length) under the identity residue correspondence,
 * for the sample and for its mirror image; aligned = t + pos @ R as in the reference.
 * in: [pos f32 B,N,3 | ref f32 (1|B),N,3 | mask B,N]   out: [tm f32 B,2 | rmsd f32 B,2 | R f32 B,2,3,3 | t f32 B,2,3] */
PRD_DECLARE_OP(kabsch)

/* --- backward pass (reference: Lightning's backward through model.py:528-549 with per-block
 * checkpointing, modules.py:399-401) --------------------------------------------------------------------------
 * One prd_<op>_bwd per forward op, same uniform signature.  Conventions (csrc/prd_bwd_api.cu):
 *   - `in`  : the op's forward INPUTS (the op recomputes its intermediates: nothing else is stored by the forward);
 *   - `out` : out[0] is the gradient buffer of the op's main operand, IN/OUT for the residual ops (d output on entry,
 *             d input on exit, may not alias `in`); further gradient buffers marked (+=) are accumulated; weight
 *             gradients are fp32 buffers shaped like the REFERENCE parameters and are accumulated (+=);
 *   - `weights` : the RAW fp32 reference parameters (nn.Linear layout [out, in]), not the packed fp16 forward copies;
 *   - arithmetic: fp32 activations, products on kind::tf32 tensor cores (operands rounded to nearest), weight gradients
 *     as exact fp32 reductions.  Pointer orders are documented above each op in csrc/prd_bwd_api.cu. */
#define PRD_DECLARE_BWD(name)                                                                     \
  int prd_##name##_bwd(const PrdDims* d, const void* const* in, void* const* out,                 \
                       const void* const* weights, void* workspace, size_t workspace_bytes,       \
                       void* stream);                                                             \
  size_t prd_##name##_bwd_workspace_bytes(const PrdDims* d);
PRD_DECLARE_BWD(pair_transition)          /* modules.py:321-326,342 */
PRD_DECLARE_BWD(single_transition)        /* modules.py:306-311,336 */
PRD_DECLARE_BWD(seq_head)                 /* model.py:374 */
PRD_DECLARE_BWD(coord_head)               /* model.py:364-373 + modules.py:403 */
PRD_DECLARE_BWD(triangle_attention)       /* modules.py:236-243 */
PRD_DECLARE_BWD(triangle_multiplication)  /* modules.py:262-274 */
PRD_DECLARE_BWD(outer_linear)             /* modules.py:283-287 */
PRD_DECLARE_BWD(single_attention)         /* modules.py:300-304,185-225 */
PRD_DECLARE_BWD(spattention)              /* AF2_modules.py:421-473 */
PRD_DECLARE_BWD(opm_project)              /* AF2_modules.py:519-530 */
PRD_DECLARE_BWD(pair_embed)               /* model.py:348-361, AF2_modules.py:532-543 */
PRD_DECLARE_BWD(single_embed)             /* model.py:342-346,99-102 */
#undef PRD_DECLARE_BWD

/* Test hook of the weight-gradient reduction dW[n,k] += alpha sum_r dY[r,n] X[r,k] (db[n] += alpha sum_r dY[r,n], or NULL):
 * mode 0 = dispatch as the backward ops do, 1 = never used, 2 = force the tcgen05 kernel (both operands MN-major tf32). */
int prd_dw_acc(const float* dY, long long ldy, const float* X, long long ldx, long long R, int Nout, int K, float* dW,
               long long ldw, float* db, float alpha, int mode, void* stream);

/* Profiling hook used by bench.py: average duration (ms) of ONE named kernel ("triattn_flash",
 * "trimul_gemm", "pair_bias") over `iters` launches on the data a previous full op left in the
 * workspace; CUDA events on `stream`.  aux: mask (triattn_flash) / pair (pair_bias). */
int prd_profile_kernel(const char* name, const PrdDims* d, void* workspace, size_t workspace_bytes,
                       const void* aux, int iters, float* ms_out, void* stream);

#undef PRD_DECLARE_OP

#ifdef __cplusplus
}
#endif
#endif /* PRD_DENOISER_H_ */
