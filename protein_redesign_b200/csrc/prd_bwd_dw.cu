// Weight gradients on the tensor cores:  dW[n, k] += alpha * sum_r dY[r, n] X[r, k]   (tcgen05, kind::tf32).
//
// The contraction runs over the ROWS of two row-major activations, i.e. both UMMA operands are MN-major: A = dY^T
// ([Nout x R], Nout contiguous in memory), B = X^T ([K x R], K contiguous).  MN-major operands of a 32-bit type use the
// "128B swizzle with 32-byte atoms" (UMMA layout type 1, SWIZZLE_128B_BASE32B; TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B):
// rows of 128 bytes = 32 fp32 along MN, the four 32-byte chunks of a row XOR-ed with (row mod 4), an atom = 4 rows along
// the contraction (cute::UMMA::Layout_MN_SW128_32B_Atom).  A TMA box of [32 columns x 32 rows] therefore lands as eight
// such atoms back to back: SBO = 512 B between 4-row atoms, LBO = 4096 B between 32-column chunks (one box); one UMMA
// consumes 8 rows (K = 8 for tf32), so a k-step advances the start address by 1024 B.  Each CTA owns a [128 x BN] tile of dW and a slice of the rows, accumulates in TMEM and adds its
// partial sum to global memory with fp32 atomics (split-K over CTAs).  Operands are expected rounded to tf32 by their
// producers (prd_bwd.h conventions); products of two tf32 values are exact in the fp32 accumulator.
#include "prd_bwd.h"

#include <stdlib.h>

#include "prd_common.cuh"

namespace prd {

namespace {

constexpr int kDwStages = 4;
constexpr int kRowsPerBlock = 32;  // rows of the activations per pipeline stage = 4 UMMAs of K = 8

template <int BN>
struct DwSmem {
  static constexpr int kABytes = 4 * 4096;         // 128 dY columns: four boxes of [32 cols x 32 rows] fp32
  static constexpr int kBBytes = (BN / 32) * 4096; // BN X columns
  static constexpr int kStage = kABytes + kBBytes;
  static constexpr int kTotal = kDwStages * kStage + 1024 + 256;
};

// MN-major SWIZZLE_128B_BASE32B descriptor: LBO (bits 16..29) = 4096 B between 32-column chunks, SBO (bits 32..45) =
// 512 B between 4-row atoms, version 1, layout type 1.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(4096 >> 4) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(1) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_tf32_mn(uint32_t M, uint32_t N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

template <int BN>
__global__ void __launch_bounds__(192, 1)
bw_dw_tc_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, long long R,
                long long rows_per_cta, int Nout, int K, float* __restrict__ dW, long long ldw, float alpha,
                float* __restrict__ db) {
  extern __shared__ uint8_t smem_raw[];
  using L = DwSmem<BN>;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDwStages * L::kStage);
  uint64_t* full = bars;
  uint64_t* empty = bars + kDwStages;
  uint64_t* done = bars + 2 * kDwStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kDwStages + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * 128, k0 = blockIdx.y * BN;
  const long long r_begin = (long long)blockIdx.z * rows_per_cta;
  long long r_end = r_begin + rows_per_cta;
  if (r_end > R) r_end = R;
  const int nblk = r_end > r_begin ? static_cast<int>((r_end - r_begin + kRowsPerBlock - 1) / kRowsPerBlock) : 0;
  // db (the bias gradient, = the column sums of dY): the four epilogue warps are idle during the main loop and the dY tile
  // sits in shared memory anyway -- they add the 32 rows of every stage for their column (the CTAs of the first k-tile
  // only), which replaces a separate pass over dY
  const bool colsum = db != nullptr && blockIdx.y == 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kDwStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], colsum ? 5 : 1);
    }
    mbar_init(done, 1);
    fence_barrier_init();
    tma_prefetch_desc(&map_dy);
    tma_prefetch_desc(&map_x);
  }
  if (warp == 0) tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (nblk > 0) {
    if (warp == 0) {
      if (lane == 0) {
        for (int it = 0; it < nblk; ++it) {
          const int s = it % kDwStages;
          mbar_wait(&empty[s], ((it / kDwStages) & 1) ^ 1);
          mbar_expect_tx(&full[s], L::kStage);
          uint8_t* sa = smem + s * L::kStage;
          const int r0 = static_cast<int>(r_begin + (long long)it * kRowsPerBlock);
#pragma unroll
          for (int c = 0; c < 4; ++c) tma_load_2d(sa + c * 4096, &map_dy, &full[s], n0 + 32 * c, r0);
#pragma unroll
          for (int c = 0; c < BN / 32; ++c) tma_load_2d(sa + L::kABytes + c * 4096, &map_x, &full[s], k0 + 32 * c, r0);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = umma_idesc_tf32_mn(128, BN);
        for (int it = 0; it < nblk; ++it) {
          const int s = it % kDwStages;
          mbar_wait(&full[s], (it / kDwStages) & 1);
          tc_fence_after();
          const uint32_t sa = base + s * L::kStage;
          const uint64_t da = umma_desc_mn_sw128(sa), db = umma_desc_mn_sw128(sa + L::kABytes);
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks)  // 8 rows per UMMA: one 1024-byte atom along the contraction
            umma_tf32(tmem, da + ks * (1024 >> 4), db + ks * (1024 >> 4), idesc, (it > 0 || ks > 0) ? 1u : 0u);
          umma_commit(&empty[s]);
        }
        umma_commit(done);
      }
    } else {
      const int q = warp & 3;
      if (colsum) {
        const int col = q * 32 + lane;
        float csum = 0.f;
        for (int it = 0; it < nblk; ++it) {
          const int s = it % kDwStages;
          mbar_wait(&full[s], (it / kDwStages) & 1);
          if (n0 + col < Nout) {
            // box q = columns [32 q, 32 q + 32): 32 rows of 128 bytes, the 32-byte chunks of row r XOR-ed with (r mod 4)
            const uint8_t* box = smem + s * L::kStage + q * 4096 + ((lane & 7) << 2);
            const int ch = lane >> 3;
#pragma unroll 8
            for (int r = 0; r < kRowsPerBlock; ++r) csum += *reinterpret_cast<const float*>(box + r * 128 + ((ch ^ (r & 3)) << 5));
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[s]);
        }
        if (n0 + col < Nout) atomicAdd(db + n0 + col, alpha * csum);
      }
      mbar_wait(done, 0);
      tc_fence_after();
      // TMEM lane = row n of dW: adding straight from the registers makes every RED instruction touch 32 different lines
      // (32 L2 atomic transactions; 148+ CTAs x 4096 of them were ~20 us of every launch).  Each warp transposes its
      // 32 x 32 chunk through shared memory (the pipeline stages are free now) and adds whole 128-byte rows instead.
      float* tile = reinterpret_cast<float*>(smem) + q * (32 * 33);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem + (static_cast<uint32_t>(q * 32) << 16) + c * 32, r);
        tmem_ld_wait();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = alpha * __uint_as_float(r[j]);
        __syncwarp();
        const int k = k0 + c * 32 + lane;
        if (k < K) {
#pragma unroll 8
          for (int i = 0; i < 32; ++i) {
            const int n = n0 + q * 32 + i;
            if (n < Nout) atomicAdd(dW + (long long)n * ldw + k, tile[i * 33 + lane]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, BN < 32 ? 32 : BN);
}

template <int BN>
int launch_dw(const float* dY, long long ldy, const float* X, long long ldx, long long R, int Nout, int K, float* dW,
              long long ldw, float alpha, float* db, cudaStream_t s) {
  using L = DwSmem<BN>;
  CUtensorMap map_dy, map_x;
  TmaDims d;
  d.size[0] = (uint64_t)Nout; d.size[1] = (uint64_t)R;
  d.stride[0] = (uint64_t)ldy * 4;
  d.box[0] = 32; d.box[1] = kRowsPerBlock;
  if (make_tensor_map_mode(&map_dy, dY, 4, 2, d, 2)) return 1;
  d.size[0] = (uint64_t)K;
  d.stride[0] = (uint64_t)ldx * 4;
  if (make_tensor_map_mode(&map_x, X, 4, 2, d, 2)) return 1;
  static bool attr_set = false;
  if (!attr_set) {
    PRD_CUDA_OK(cudaFuncSetAttribute(bw_dw_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  const int gx = (Nout + 127) / 128, gy = (K + BN - 1) / BN;
  // CTAs per SM in total (PRD_DW_CPS: tuning switch): two pipelines per SM hide each other's prologue / epilogue; the widest
  // tile's stages leave room for one
  static const int cps_env = getenv("PRD_DW_CPS") ? atoi(getenv("PRD_DW_CPS")) : 0;
  const int cps = cps_env > 0 ? cps_env : (BN == 256 ? 1 : 2);
  long long chunks = ((long long)cps * kNumSMs + gx * gy - 1) / (gx * gy);
  const long long max_chunks = (R + 8 * kRowsPerBlock - 1) / (8 * kRowsPerBlock);  // at least 8 pipeline blocks per CTA
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  long long rows_per_cta = ((R + chunks - 1) / chunks + kRowsPerBlock - 1) / kRowsPerBlock * kRowsPerBlock;
  chunks = (R + rows_per_cta - 1) / rows_per_cta;
  bw_dw_tc_kernel<BN><<<dim3(gx, gy, (unsigned)chunks), 192, L::kTotal, s>>>(map_dy, map_x, R, rows_per_cta, Nout, K, dW, ldw, alpha, db);
  PRD_LAUNCHED();
  return 0;
}

}  // namespace

// Tensor-core path of bw_dw_acc: needs 16-byte aligned operands with row strides that are multiples of 4 floats.
bool bw_dw_tc_applies(const float* dY, long long ldy, const float* X, long long ldx, long long R) {
  return (ldy % 4 == 0) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(dY) & 15) == 0) &&
         ((reinterpret_cast<uintptr_t>(X) & 15) == 0) && R >= 256 && R < 0x7fffffffLL;
}
int bw_dw_tc(const float* dY, long long ldy, const float* X, long long ldx, long long R, int Nout, int K, float* dW,
             long long ldw, float alpha, cudaStream_t s, float* db) {
  if (K <= 64) return launch_dw<64>(dY, ldy, X, ldx, R, Nout, K, dW, ldw, alpha, db, s);
  if (K <= 128) return launch_dw<128>(dY, ldy, X, ldx, R, Nout, K, dW, ldw, alpha, db, s);
  return launch_dw<256>(dY, ldy, X, ldx, R, Nout, K, dW, ldw, alpha, db, s);
}

}  // namespace prd
