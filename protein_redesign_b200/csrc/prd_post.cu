// Output post-processing on the GPU: sequence decode (reference generate.py:76-91: argmax of the
// softmax of the sampled logits) and a batched rigid superposition of every sample onto a reference structure with RMSD
// and TM-score under the identity residue correspondence -- the role the TMalign subprocess plays in
// generate.py:176-195 (ProteinReDiff/tmalign.py:23-49), including its mirror-image variant.
#include "../../include/prd_denoiser.h"
#include "prd_common.cuh"

namespace prd {

// tokens[b, n] = argmax_k logits[b, n, k] (first maximum, like torch.argmax); 0 where the residue mask is 0
__global__ void decode_argmax_kernel(const float* __restrict__ logits, const float* __restrict__ residue_mask, long long n_tok,
                                     int K, int64_t* __restrict__ tokens) {
  const long long tok = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tok >= n_tok) return;
  const float* lp = logits + tok * K;
  int best = 0;
  float bv = lp[0];
  for (int k = 1; k < K; ++k)
    if (lp[k] > bv) {
      bv = lp[k];
      best = k;
    }
  tokens[tok] = (residue_mask == nullptr || residue_mask[tok] > 0.5f) ? best : 0;
}

// Largest eigenpair of a symmetric 4 x 4 matrix by cyclic Jacobi rotations (double precision, one thread).
__device__ void jacobi4(double a[4][4], double v[4][4]) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) v[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 24; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < 4; ++i)
      for (int j = i + 1; j < 4; ++j) off += a[i][j] * a[i][j];
    if (off < 1e-30) break;
    for (int p = 0; p < 4; ++p)
      for (int q = p + 1; q < 4; ++q) {
        if (fabs(a[p][q]) < 1e-300) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 4; ++k) {
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 4; ++k) {
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 4; ++k) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
  }
}

// grid (B, 2): blockIdx.y = 1 superposes the mirror image (z -> -z) of the sample.  Row-vector convention of the
// reference (generate.py:186): aligned = t + pos @ R.
__global__ void __launch_bounds__(128) kabsch_kernel(const float* __restrict__ pos, const float* __restrict__ ref,
                                                     const float* __restrict__ mask, int N, int ref_rows,
                                                     float* __restrict__ tm, float* __restrict__ rmsd, float* __restrict__ Rout,
                                                     float* __restrict__ tout) {
  __shared__ double red[128];
  __shared__ double stat[16];   // n, cp[3], cq[3]
  __shared__ double M[9];
  __shared__ double Rs[9], ts[3];
  const int b = blockIdx.x, mirror = blockIdx.y, tid = threadIdx.x;
  const float* P = pos + (long long)b * N * 3;
  const float* Q = ref + (long long)(ref_rows == 1 ? 0 : b) * N * 3;
  const float* W = mask + (long long)b * N;
  const double zs = mirror ? -1.0 : 1.0;
  auto block_sum = [&](double v) {
    red[tid] = v;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
      if (tid < o) red[tid] += red[tid + o];
      __syncthreads();
    }
    const double r = red[0];
    __syncthreads();
    return r;
  };
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int i = tid; i < N; i += 128) {
    const double w = W[i] > 0.5f ? 1.0 : 0.0;
    acc[0] += w;
    acc[1] += w * P[i * 3]; acc[2] += w * P[i * 3 + 1]; acc[3] += w * zs * P[i * 3 + 2];
    acc[4] += w * Q[i * 3]; acc[5] += w * Q[i * 3 + 1]; acc[6] += w * Q[i * 3 + 2];
  }
  for (int k = 0; k < 7; ++k) {
    const double s = block_sum(acc[k]);
    if (tid == 0) stat[k] = s;
  }
  __syncthreads();
  const double n = stat[0] > 0 ? stat[0] : 1.0;
  const double cp[3] = {stat[1] / n, stat[2] / n, stat[3] / n}, cq[3] = {stat[4] / n, stat[5] / n, stat[6] / n};
  double m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, gp = 0, gq = 0;
  for (int i = tid; i < N; i += 128) {
    if (W[i] <= 0.5f) continue;
    const double p[3] = {P[i * 3] - cp[0], P[i * 3 + 1] - cp[1], zs * P[i * 3 + 2] - cp[2]};
    const double q[3] = {Q[i * 3] - cq[0], Q[i * 3 + 1] - cq[1], Q[i * 3 + 2] - cq[2]};
    for (int x = 0; x < 3; ++x) {
      gp += p[x] * p[x];
      gq += q[x] * q[x];
      for (int y = 0; y < 3; ++y) m[x * 3 + y] += p[x] * q[y];
    }
  }
  for (int k = 0; k < 9; ++k) {
    const double s = block_sum(m[k]);
    if (tid == 0) M[k] = s;
  }
  const double GP = block_sum(gp), GQ = block_sum(gq);
  if (tid == 0) {
    const double Sxx = M[0], Sxy = M[1], Sxz = M[2], Syx = M[3], Syy = M[4], Syz = M[5], Szx = M[6], Szy = M[7], Szz = M[8];
    double a[4][4] = {{Sxx + Syy + Szz, Syz - Szy, Szx - Sxz, Sxy - Syx},
                      {Syz - Szy, Sxx - Syy - Szz, Sxy + Syx, Szx + Sxz},
                      {Szx - Sxz, Sxy + Syx, -Sxx + Syy - Szz, Syz + Szy},
                      {Sxy - Syx, Szx + Sxz, Syz + Szy, -Sxx - Syy + Szz}};
    double v[4][4];
    jacobi4(a, v);
    int best = 0;
    for (int k = 1; k < 4; ++k)
      if (a[k][k] > a[best][best]) best = k;
    const double lam = a[best][best];
    double q0 = v[0][best], qx = v[1][best], qy = v[2][best], qz = v[3][best];
    const double qn = sqrt(q0 * q0 + qx * qx + qy * qy + qz * qz);
    q0 /= qn; qx /= qn; qy /= qn; qz /= qn;
    // column-vector rotation Rc (q ~= Rc p); the reference uses rows: aligned = p @ R with R = Rc^T
    const double Rc[9] = {q0 * q0 + qx * qx - qy * qy - qz * qz, 2 * (qx * qy - q0 * qz), 2 * (qx * qz + q0 * qy),
                          2 * (qy * qx + q0 * qz), q0 * q0 - qx * qx + qy * qy - qz * qz, 2 * (qy * qz - q0 * qx),
                          2 * (qz * qx - q0 * qy), 2 * (qz * qy + q0 * qx), q0 * q0 - qx * qx - qy * qy + qz * qz};
    for (int x = 0; x < 3; ++x)
      for (int y = 0; y < 3; ++y) Rs[x * 3 + y] = Rc[y * 3 + x];
    for (int y = 0; y < 3; ++y) ts[y] = cq[y] - (cp[0] * Rs[0 * 3 + y] + cp[1] * Rs[1 * 3 + y] + cp[2] * Rs[2 * 3 + y]);
    double r2 = (GP + GQ - 2.0 * lam) / n;
    rmsd[b * 2 + mirror] = (float)sqrt(r2 > 0 ? r2 : 0.0);
    // hand R back in terms of the UN-mirrored sample: p_mirrored = p diag(1, 1, -1)  =>  R_out = diag(1, 1, -1) R
    for (int x = 0; x < 3; ++x)
      for (int y = 0; y < 3; ++y) Rout[((b * 2 + mirror) * 3 + x) * 3 + y] = (float)((x == 2 ? zs : 1.0) * Rs[x * 3 + y]);
    for (int y = 0; y < 3; ++y) tout[(b * 2 + mirror) * 3 + y] = (float)ts[y];
  }
  __syncthreads();
  // TM-score normalised by the reference length (TM-align's "TM2", tmalign.py:41), identity correspondence
  const double L = stat[0];
  double d0 = L > 21.0 ? 1.24 * cbrt(L - 15.0) - 1.8 : 0.5;
  if (d0 < 0.5) d0 = 0.5;
  double s = 0.0;
  for (int i = tid; i < N; i += 128) {
    if (W[i] <= 0.5f) continue;
    const double p[3] = {P[i * 3], P[i * 3 + 1], zs * P[i * 3 + 2]};
    double d2 = 0.0;
    for (int y = 0; y < 3; ++y) {
      const double a = ts[y] + p[0] * Rs[0 * 3 + y] + p[1] * Rs[1 * 3 + y] + p[2] * Rs[2 * 3 + y] - Q[i * 3 + y];
      d2 += a * a;
    }
    s += 1.0 / (1.0 + d2 / (d0 * d0));
  }
  const double S = block_sum(s);
  if (tid == 0) tm[b * 2 + mirror] = (float)(L > 0 ? S / L : 0.0);
}

}  // namespace prd

using namespace prd;

extern "C" {

// generate.py:76-91.  in: [logits f32 B,N,21 | residue_mask B,N (or NULL)]   out: [tokens i64 B,N]
size_t prd_decode_argmax_workspace_bytes(const PrdDims*) { return 256; }
int prd_decode_argmax_fwd(const PrdDims* d, const void* const* in, void* const* out, const void* const*, void*, size_t,
                          void* stream) {
  if (prd_device_check()) return 1;
  const long long n = (long long)d->B * d->N;
  decode_argmax_kernel<<<(unsigned)((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float*>(in[0]), static_cast<const float*>(in[1]), n, 21, static_cast<int64_t*>(out[0]));
  PRD_LAUNCHED();
  return 0;
}

// tmalign.py:23-49 / generate.py:176-195.  in: [pos f32 B,N,3 | ref f32 (B or 1),N,3 | mask B,N]; d->mode = rows of ref.
// out: [tm f32 B,2 | rmsd f32 B,2 | R f32 B,2,3,3 | t f32 B,2,3]   (index 1 of the size-2 axis: mirror image)
size_t prd_kabsch_workspace_bytes(const PrdDims*) { return 256; }
int prd_kabsch_fwd(const PrdDims* d, const void* const* in, void* const* out, const void* const*, void*, size_t,
                   void* stream) {
  if (prd_device_check()) return 1;
  PRD_REQUIRE(d->mode == 1 || d->mode == d->B, "kabsch: the reference must have 1 or B rows (got %d)", d->mode);
  kabsch_kernel<<<dim3(d->B, 2), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float*>(in[0]), static_cast<const float*>(in[1]), static_cast<const float*>(in[2]), d->N, d->mode,
      static_cast<float*>(out[0]), static_cast<float*>(out[1]), static_cast<float*>(out[2]), static_cast<float*>(out[3]));
  PRD_LAUNCHED();
  return 0;
}

}  // extern "C"
