// Declarations for prd_loss.cu (forward noising q(), loss terms and output gradients).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace prd {

// sched: [T][2] = {sqrt_alphas_cumprod, sqrt_one_minus_alphas_cumprod} (model.py:181-189)
int diffusion_q(int B, int N, int T, const float* x, const float* seq, const int64_t* t, const float* noise_z,
                const float* noise_seq, const float* keep, const float* drop, const float* sched, float* z_t, float* seq_t,
                float* seq_t1, cudaStream_t s);
// row_w [B] and partial [B*N*3] are scratch; terms [B+2] (per-row MSE, KL, CE), d_noise, d_seq may be NULL.
int diffusion_loss(int B, int N, int T, const float* noise_pred, const float* seq_pred, const float* noise_z,
                   const float* noise_seq, const float* seq_t1, const float* mask, const float* residue_mask,
                   const int64_t* residue_type, const int64_t* t, const float* sched, float* row_w, float* partial,
                   float* loss, float* diff_loss, float* terms, float* d_noise, float* d_seq, cudaStream_t s);

}  // namespace prd
