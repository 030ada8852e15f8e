// Declarations for prd_embed.cu (single embedding, time embedding, sampler update).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace prd {

struct AtomTables {
  const float* t[9];
};

// Device-resident sampler state, so one captured CUDA graph can be replayed for every step.
struct SamplerState {
  int t_cur;  // current diffusion time index (num_steps-1 ... 0)
  int step;   // number of completed reverse steps (indexes the pre-generated noise)
};

int single_embed(int B, int N, int CS, const int64_t* atom_feats, const float* atom_mask, const float* residue_mask,
                 const float* seq_t, const float* esm_emb, const AtomTables& tabs, const float* w_type, float* single,
                 cudaStream_t s);
int time_embed(int B, int CZ, int TD, const int64_t* t, const SamplerState* st, int num_steps, const float* freq,
               const float* w_beta, float* beta, cudaStream_t s);
int sampler_update(int B, int N, int T, const float* eps, const float* seq_pred, const float* noise, const float* coef,
                   SamplerState* st, float* z, float* seq_t, cudaStream_t s);

}  // namespace prd
