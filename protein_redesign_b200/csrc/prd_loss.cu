// Training objective of ProteinReDiff on device: the forward noising q() (model.py:471-488), the three
// loss terms of diffusion_loss after the network call (model.py:499-526), loss = mean(diff_loss / num_nodes)
// (model.py:538-541) and the gradient of that loss with respect to the two network outputs (the seed of the backward
// pass).  Small warp-shuffle kernels: one thread per token for q(), one warp per token (lane = residue class) for the
// loss terms, one block for the fixed-order final reduction (no atomics: results are run-to-run deterministic).
#include "prd_common.cuh"
#include "prd_loss.h"

namespace prd {

namespace {
constexpr int kClasses = 21;  // len(RESIDUE_TYPES) + 1 (protein.py:28-31, model.py:433)

// log_softmax over the 21 class lanes of a warp; lanes >= 21 pass -inf and get -inf back
__device__ __forceinline__ float warp_log_softmax(float v, bool live) {
  const float m = warp_max(live ? v : -INFINITY);
  const float e = live ? expf(v - m) : 0.f;
  const float s = warp_sum(e);
  return live ? (v - m) - logf(s) : -INFINITY;
}
}  // namespace

// z_t = sa[t] x + s1[t] nz;  seq_t = keep*seq + drop*(sa[t] seq + s1[t] ns);  seq_t1 = sa[t1] seq + s1[t1] ns, t1 = max(t-1, 0)
__global__ void diffusion_q_kernel(int N, long long n_tok, int T, const float* __restrict__ x, const float* __restrict__ seq,
                                   const int64_t* __restrict__ t, const float* __restrict__ nz,
                                   const float* __restrict__ ns, const float* __restrict__ keep,
                                   const float* __restrict__ drop, const float* __restrict__ sched,
                                   float* __restrict__ z_t, float* __restrict__ seq_t, float* __restrict__ seq_t1) {
  const long long tok = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tok >= n_tok) return;
  const int b = static_cast<int>(tok / N);
  long long tb = t[b];
  tb = tb < 0 ? 0 : (tb >= T ? T - 1 : tb);
  const long long t1 = tb > 0 ? tb - 1 : 0;
  const float sa = sched[tb * 2], s1 = sched[tb * 2 + 1], sa1 = sched[t1 * 2], s11 = sched[t1 * 2 + 1];
#pragma unroll
  for (int k = 0; k < 3; ++k) z_t[tok * 3 + k] = sa * x[tok * 3 + k] + s1 * nz[tok * 3 + k];
  const float kp = keep[tok], dr = drop[tok];
#pragma unroll
  for (int k = 0; k < kClasses; ++k) {
    const float s = seq[tok * kClasses + k], n = ns[tok * kClasses + k];
    seq_t[tok * kClasses + k] = kp * s + dr * (sa * s + s1 * n);
    seq_t1[tok * kClasses + k] = sa1 * s + s11 * n;
  }
}

int diffusion_q(int B, int N, int T, const float* x, const float* seq, const int64_t* t, const float* nz, const float* ns,
                const float* keep, const float* drop, const float* sched, float* z_t, float* seq_t, float* seq_t1,
                cudaStream_t s) {
  const long long n_tok = (long long)B * N;
  diffusion_q_kernel<<<(unsigned)((n_tok + 127) / 128), 128, 0, s>>>(N, n_tok, T, x, seq, t, nz, ns, keep, drop, sched, z_t,
                                                                     seq_t, seq_t1);
  PRD_LAUNCHED();
  return 0;
}

// row_w[b] = 1 / (B * num_nodes_b), num_nodes_b = #(mask > 0.5)   (model.py:538,541: mean over rows of diff_loss / num_nodes)
__global__ void loss_row_weight_kernel(int B, int N, const float* __restrict__ mask, float* __restrict__ row_w) {
  __shared__ int part[32];
  const int b = blockIdx.x;
  int n = 0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) n += mask[(long long)b * N + i] > 0.5f ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = n;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) tot += part[w];
    row_w[b] = 1.0f / (static_cast<float>(B) * static_cast<float>(tot));  // inf on an empty row, as the reference
  }
}

// One warp per token.  partial[tok] = {mask * |noise_pred - noise_z|^2, KL term, CE term}; optional gradients of
// loss = sum_b row_w[b] * (mse_b + KL + CE) with respect to noise_pred and seq_pred.
__global__ void loss_tokens_kernel(int B, int N, long long n_tok, int T, const float* __restrict__ noise_pred,
                                   const float* __restrict__ seq_pred, const float* __restrict__ noise_z,
                                   const float* __restrict__ noise_seq, const float* __restrict__ seq_t1,
                                   const float* __restrict__ mask, const float* __restrict__ residue_mask,
                                   const int64_t* __restrict__ residue_type, const int64_t* __restrict__ t,
                                   const float* __restrict__ sched, const float* __restrict__ row_w,
                                   float* __restrict__ partial, float* __restrict__ d_noise,
                                   float* __restrict__ d_seq) {
  const int lane = threadIdx.x & 31;
  const long long tok = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (tok >= n_tok) return;
  const int b = static_cast<int>(tok / N);
  long long tb = t[b];
  tb = tb < 0 ? 0 : (tb >= T ? T - 1 : tb);
  const long long t1 = tb > 0 ? tb - 1 : 0;
  const float sa1 = sched[t1 * 2], s11 = sched[t1 * 2 + 1];
  const float m = mask[tok], rm = residue_mask[tok];
  const float wb = row_w[b];
  float wsum = 0.f;  // sum_b row_w[b]: the scalar KL / CE terms are added to every row (model.py:512-525)
  for (int r = 0; r < B; ++r) wsum += row_w[r];

  // noise MSE (model.py:504-511)
  float sq = 0.f;
  if (lane < 3) {
    const float d = noise_pred[tok * 3 + lane] - noise_z[tok * 3 + lane];
    sq = m * d * d;
    if (d_noise) d_noise[tok * 3 + lane] = wb * 2.0f * m * d;
  }
  sq = warp_sum(sq);

  const bool live = lane < kClasses;
  const float sp = live ? seq_pred[tok * kClasses + lane] : 0.f;
  // KL(softmax(seq_{t-1}) * rm || .) with input log_softmax(seq_pred_{t-1}) * rm, reduction none, summed (model.py:512-518):
  // term = xlogy(target, target) - target * input
  const float u = live ? sa1 * sp + s11 * noise_seq[tok * kClasses + lane] : 0.f;
  const float logq = warp_log_softmax(u, live);
  const float logp = warp_log_softmax(live ? seq_t1[tok * kClasses + lane] : 0.f, live);
  const float tg = live ? expf(logp) * rm : 0.f;
  float kl = 0.f;
  if (live && tg > 0.f) kl = tg * ((rm == 1.0f ? logp : logf(tg)) - logq * rm);
  kl = warp_sum(kl);
  const float tg_sum = warp_sum(tg);

  // CE((seq_pred + 1) / 2, residue_type, ignore_index 0) * mask (model.py:520-525)
  const long long ty = residue_type[tok];
  const float lsm = warp_log_softmax(0.5f * (sp + 1.0f), live);
  const float picked = __shfl_sync(0xffffffffu, lsm, static_cast<int>(ty) & 31);
  const bool counted = ty > 0 && ty < kClasses;
  const float ce = counted ? -picked * m : 0.f;

  if (lane == 0) {
    partial[tok * 3 + 0] = sq;
    partial[tok * 3 + 1] = kl;
    partial[tok * 3 + 2] = ce;
  }
  if (d_seq && live) {
    const float q = expf(logq);
    float g = wsum * sa1 * rm * (q * tg_sum - tg);
    if (counted) g += wsum * m * 0.5f * (expf(lsm) - (lane == ty ? 1.0f : 0.0f));
    d_seq[tok * kClasses + lane] = g;
  }
}

// One block: diff_loss[b] = mse_b + KL + CE, loss = sum_b row_w[b] * diff_loss[b]; terms = {mse_0..mse_{B-1}, KL, CE}.
// Fixed summation order (thread-strided partial sums in double, then a shared-memory tree).
__global__ void loss_finish_kernel(int B, int N, const float* __restrict__ partial, const float* __restrict__ row_w,
                                   float* __restrict__ loss, float* __restrict__ diff_loss, float* __restrict__ terms) {
  extern __shared__ double red[];  // [blockDim.x]
  __shared__ double s_kl, s_ce;
  auto block_sum = [&](double v) {
    red[threadIdx.x] = v;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    const double r = red[0];
    __syncthreads();
    return r;
  };
  const long long n_tok = (long long)B * N;
  double kl = 0.0, ce = 0.0;
  for (long long i = threadIdx.x; i < n_tok; i += blockDim.x) {
    kl += partial[i * 3 + 1];
    ce += partial[i * 3 + 2];
  }
  kl = block_sum(kl);
  ce = block_sum(ce);
  if (threadIdx.x == 0) {
    s_kl = kl;
    s_ce = ce;
    if (terms) {
      terms[B] = static_cast<float>(kl);
      terms[B + 1] = static_cast<float>(ce);
    }
  }
  double total = 0.0;
  for (int b = 0; b < B; ++b) {
    double mse = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) mse += partial[((long long)b * N + i) * 3];
    mse = block_sum(mse);
    if (threadIdx.x == 0) {
      const float dl = static_cast<float>(mse) + static_cast<float>(s_kl) + static_cast<float>(s_ce);
      diff_loss[b] = dl;
      if (terms) terms[b] = static_cast<float>(mse);
      total += static_cast<double>(dl) * static_cast<double>(row_w[b]);
    }
  }
  if (threadIdx.x == 0) loss[0] = static_cast<float>(total);
}

int diffusion_loss(int B, int N, int T, const float* noise_pred, const float* seq_pred, const float* noise_z,
                   const float* noise_seq, const float* seq_t1, const float* mask, const float* residue_mask,
                   const int64_t* residue_type, const int64_t* t, const float* sched, float* row_w, float* partial,
                   float* loss, float* diff_loss, float* terms, float* d_noise, float* d_seq, cudaStream_t s) {
  const long long n_tok = (long long)B * N;
  loss_row_weight_kernel<<<B, 128, 0, s>>>(B, N, mask, row_w);
  PRD_LAUNCHED();
  loss_tokens_kernel<<<(unsigned)((n_tok * 32 + 255) / 256), 256, 0, s>>>(B, N, n_tok, T, noise_pred, seq_pred, noise_z, noise_seq,
                                                                         seq_t1, mask, residue_mask, residue_type, t, sched,
                                                                         row_w, partial, d_noise, d_seq);
  PRD_LAUNCHED();
  loss_finish_kernel<<<1, 256, 256 * sizeof(double), s>>>(B, N, partial, row_w, loss, diff_loss, terms);
  PRD_LAUNCHED();
  return 0;
}

}  // namespace prd
