// Small latency-bound kernels: single-representation input embedding (model.py:342-346), the
// time embedding vector (modules.py:85-97 + model.py:86-87) and the DDPM sampler update
// (model.py:403-420).
#include "prd_kernels.h"
#include "prd_embed.h"
#include "prd_common.cuh"

namespace prd {

// single[b,n,:] = atom_mask * (1/3) sum_f E_f[atom_feats[b,n,f]] +
//                 residue_mask * (relu(W_type LN(seq_t[b,n,:21])) + esm_emb[b,n,:])
__global__ void __launch_bounds__(128)
single_embed_kernel(int CS, const int64_t* __restrict__ atom_feats, const float* __restrict__ atom_mask,
                    const float* __restrict__ residue_mask, const float* __restrict__ seq_t,
                    const float* __restrict__ esm_emb, AtomTables tabs, const float* __restrict__ w_type,
                    float* __restrict__ single) {
  __shared__ float sLn[21];
  __shared__ long long sIdx[9];
  const long long tok = blockIdx.x;
  const int t = threadIdx.x;
  // nn.Embedding raises on an out-of-range index; a kernel cannot, so indices are clamped into the table
  // (features.py:31-60 vocabulary sizes) instead of reading out of bounds
  constexpr int kVocab[9] = {119, 4, 12, 12, 10, 6, 6, 2, 2};
  if (t < 9) {
    const long long v = atom_feats[tok * 9 + t];
    sIdx[t] = v < 0 ? 0 : (v >= kVocab[t] ? kVocab[t] - 1 : v);
  }
  if (t == 0) {
    const float* s = seq_t + tok * 21;
    float mean = 0.f;
    for (int k = 0; k < 21; ++k) mean += s[k];
    mean /= 21.f;
    float var = 0.f;
    for (int k = 0; k < 21; ++k) var += (s[k] - mean) * (s[k] - mean);
    const float rstd = rsqrtf(var / 21.f + 1e-5f);
    for (int k = 0; k < 21; ++k) sLn[k] = (s[k] - mean) * rstd;
  }
  __syncthreads();
  const float am = atom_mask[tok], rm = residue_mask[tok];
  const float scale = 1.0f / 3.0f;  // 1/sqrt(9) (modules.py:45)
  for (int c = t; c < CS; c += 128) {
    float atom = 0.f;
#pragma unroll
    for (int f = 0; f < 9; ++f) atom += scale * tabs.t[f][sIdx[f] * CS + c];
    float ty = 0.f;
    const float* wr = w_type + (long long)c * 21;
#pragma unroll
    for (int k = 0; k < 21; ++k) ty += wr[k] * sLn[k];
    ty = fmaxf(ty, 0.f);
    single[tok * CS + c] = am * atom + rm * (ty + esm_emb[tok * CS + c]);
  }
}

int single_embed(int B, int N, int CS, const int64_t* atom_feats, const float* atom_mask, const float* residue_mask,
                 const float* seq_t, const float* esm_emb, const AtomTables& tabs, const float* w_type, float* single,
                 cudaStream_t s) {
  single_embed_kernel<<<B * N, 128, 0, s>>>(CS, atom_feats, atom_mask, residue_mask, seq_t, esm_emb, tabs, w_type, single);
  PRD_LAUNCHED();
  return 0;
}

// beta[b,:] = W_beta [sin(freq * t_b/T), cos(freq * t_b/T)];  t comes either from a per-row int64
// tensor (forward / sample_step signature) or from the device-side sampler state (graph replay).
__global__ void time_embed_kernel(int CZ, int TD, const int64_t* __restrict__ t, const SamplerState* __restrict__ st,
                                  int num_steps, const float* __restrict__ freq, const float* __restrict__ w_beta,
                                  float* __restrict__ beta) {
  extern __shared__ float sF[];  // [TD]
  const int b = blockIdx.x;
  long long tb = st ? (long long)st->t_cur : t[b];
  tb = tb < 0 ? 0 : (tb >= num_steps ? num_steps - 1 : tb);  // a stale / over-replayed state must not index out of range
  const float scaled = static_cast<float>(tb) / static_cast<float>(num_steps);
  const int half = TD / 2;
  for (int k = threadIdx.x; k < half; k += blockDim.x) {
    const float wx = freq[k] * scaled;
    sF[k] = sinf(wx);
    sF[half + k] = cosf(wx);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < CZ; c += blockDim.x) {
    const float* wr = w_beta + (long long)c * TD;
    float acc = 0.f;
    for (int k = 0; k < TD; ++k) acc += wr[k] * sF[k];
    beta[(long long)b * CZ + c] = acc;
  }
}

int time_embed(int B, int CZ, int TD, const int64_t* t, const SamplerState* st, int num_steps, const float* freq,
               const float* w_beta, float* beta, cudaStream_t s) {
  time_embed_kernel<<<B, 64, TD * sizeof(float), s>>>(CZ, TD, t, st, num_steps, freq, w_beta, beta);
  PRD_LAUNCHED();
  return 0;
}

// One reverse-diffusion update (model.py:407-420) for the whole batch, all on device:
//   mean  = (z - (1-alpha_t)/sqrt(1-abar_t) * eps) / sqrt(alpha_t)
//   z     = mean                      if t == 0
//         = mean + sqrt(beta_t) * n   otherwise          (n already mean-removed)
//   seq_t = 2 softmax(seq_pred) - 1
// coef[t] = {1/sqrt(alpha_t), (1-alpha_t)/sqrt(1-abar_t), sqrt(beta_t)}.  The current time index
// and the step counter live in *st (device memory) so that one captured CUDA graph serves every
// step; noise is the pre-generated, mean-removed tensor [steps][B*N][3] indexed by st->step.
// A second one-thread kernel advances the state after every block has read it.
__global__ void sampler_update_kernel(long long n_tok, int T, const float* __restrict__ eps, const float* __restrict__ seq_pred,
                                      const float* __restrict__ noise, const float* __restrict__ coef,
                                      const SamplerState* __restrict__ st, float* __restrict__ z,
                                      float* __restrict__ seq_t) {
  const long long tok = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tok >= n_tok) return;
  // clamped: replaying the captured step more than T times must not read past coef[T][3] / noise[T-1][..]
  const int tc = min(max(st->t_cur, 0), T - 1);
  const int sidx = min(max(st->step, 0), max(T - 2, 0));
  const float c1 = coef[tc * 3 + 0], c2 = coef[tc * 3 + 1], sd = coef[tc * 3 + 2];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float mean = c1 * (z[tok * 3 + k] - c2 * eps[tok * 3 + k]);
    z[tok * 3 + k] = (tc == 0) ? mean : mean + sd * noise[((long long)sidx * n_tok + tok) * 3 + k];
  }
  const float* sp = seq_pred + tok * 21;
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < 21; ++k) m = fmaxf(m, sp[k]);
  float e[21], s = 0.f;
#pragma unroll
  for (int k = 0; k < 21; ++k) {
    e[k] = __expf(sp[k] - m);
    s += e[k];
  }
  const float inv = 1.0f / s;
#pragma unroll
  for (int k = 0; k < 21; ++k) seq_t[tok * 21 + k] = e[k] * inv * 2.0f - 1.0f;
}

// After the last step (t_cur == 0) the state wraps to the start of a new trajectory, so a replay loop of any length
// (bench.py) stays inside the schedule and the noise tensor.
__global__ void sampler_advance_kernel(SamplerState* st, int T) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (st->t_cur <= 0) {
      st->t_cur = T - 1;
      st->step = 0;
    } else {
      st->t_cur -= 1;
      st->step += 1;
    }
  }
}

int sampler_update(int B, int N, int T, const float* eps, const float* seq_pred, const float* noise, const float* coef,
                   SamplerState* st, float* z, float* seq_t, cudaStream_t s) {
  const long long n_tok = (long long)B * N;
  sampler_update_kernel<<<(unsigned)((n_tok + 127) / 128), 128, 0, s>>>(n_tok, T, eps, seq_pred, noise, coef, st, z, seq_t);
  PRD_LAUNCHED();
  sampler_advance_kernel<<<1, 32, 0, s>>>(st, T);
  PRD_LAUNCHED();
  return 0;
}

}  // namespace prd
