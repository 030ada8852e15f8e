// Common device-side building blocks for the sm_100a kernels of the ProteinReDiff denoiser.
//
// Everything here is thin inline PTX over the Blackwell primitives we use:
//   * mbarrier (with a watchdog so a mis-programmed pipeline traps instead of hanging the box)
//   * cp.async.bulk (1-D bulk copies global<->shared) and cp.async.bulk.tensor (TMA tiles)
//   * tcgen05: TMEM alloc/dealloc, UMMA smem/instruction descriptors, mma, commit, ld, fences
//   * the 128-byte-swizzled K-major shared-memory tile layout shared by TMA and UMMA
//
// Conventions: fp16 operands, fp32 accumulation, K-major operands, SWIZZLE_128B everywhere
// (one smem "K-block" = rows x 64 halves = rows x 128 bytes, 8-row atoms of 1024 bytes).
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

namespace prd {

// ---------------------------------------------------------------------------------------
// host-side error plumbing (thread-local message returned by prd_last_error())
// ---------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define PRD_CUDA_OK(expr)                                   \
  do {                                                      \
    if (::prd::check_cuda((expr), #expr) != 0) return 1;    \
  } while (0)
#define PRD_REQUIRE(cond, ...)                              \
  do {                                                      \
    if (!(cond)) {                                          \
      ::prd::set_error(__VA_ARGS__);                        \
      return 1;                                             \
    }                                                       \
  } while (0)

// count of kernels launched by this library in this process (prd_launch_count())
extern std::atomic<long long> g_launches;
#define PRD_LAUNCHED()                       \
  do {                                       \
    ++::prd::g_launches;                     \
    PRD_CUDA_OK(cudaGetLastError());         \
  } while (0)

constexpr int kNumSMs = 148;

// ---------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  A kernel launched through launch_pdl() may start while its predecessor in the
// stream is still draining: its CTAs become resident as the predecessor's CTAs exit, run their prologue (barrier init,
// TMEM allocation, weight tiles -> shared memory: nothing the predecessor produces) and then block in pdl_wait() until
// the predecessor has completed and its writes are visible.  Every kernel launched this way calls pdl_trigger() first
// (lets ITS successor do the same) and pdl_wait() before the first access to anything a previous kernel wrote; both are
// no-ops for an ordinary launch.  PRD_PDL=0 turns the launch attribute off (A/B timing).
// ---------------------------------------------------------------------------------------
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ float sigmoidf_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Byte offset of (row, 16-byte chunk) inside a K-major SWIZZLE_128B tile whose rows are 128 bytes
// (64 halves).  Swizzle<3,4,3>: chunk index (addr bits 4..6) ^= row-in-atom (addr bits 7..9).
// The tile base must be 1024-byte aligned.
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t chunk) {
  return row * 128u + ((chunk ^ (row & 7u)) << 4);
}

// Coalesced store of 32 rows of 128 bytes owned one-per-lane (v[c] = 16-byte chunk c of this lane's row;
// row i of the warp goes to gbase + i * pitch_bytes).  A direct uint4 store per lane hits 32 different
// lines with 16 bytes each per instruction (partial sectors, 32 requests); staged through a warp-private
// 4 KB swizzled shared-memory slice, every store instruction writes four full 128-byte lines.
// rows_valid: rows of this warp that exist (the rest is not written).
__device__ __forceinline__ void warp_store_rows128(uint8_t* slice, int lane, const uint4 (&v)[8], void* gbase,
                                                   long long pitch_bytes, int rows_valid) {
  __syncwarp();  // the slice may still be read by the previous call
#pragma unroll
  for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(slice + lane * 128 + ((c ^ (lane & 7)) << 4)) = v[c];
  __syncwarp();
  uint8_t* gp = reinterpret_cast<uint8_t*>(gbase) + (lane & 7) * 16;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 4 + (lane >> 3), c = lane & 7;
    const uint4 o = *reinterpret_cast<const uint4*>(slice + row * 128 + ((c ^ (row & 7)) << 4));
    if (row < rows_valid) *reinterpret_cast<uint4*>(gp + row * pitch_bytes) = o;
  }
}

// The matching load: 32 rows of 128 bytes (row i at gbase + i * pitch_bytes) arrive one-per-lane in v[],
// fetched with fully coalesced 16-byte loads.  Rows >= rows_valid read as zero.
__device__ __forceinline__ void warp_load_rows128(uint8_t* slice, int lane, uint4 (&v)[8], const void* gbase,
                                                  long long pitch_bytes, int rows_valid) {
  __syncwarp();
  const uint8_t* gp = reinterpret_cast<const uint8_t*>(gbase) + (lane & 7) * 16;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 4 + (lane >> 3), c = lane & 7;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (row < rows_valid) o = __ldg(reinterpret_cast<const uint4*>(gp + row * pitch_bytes));
    *reinterpret_cast<uint4*>(slice + row * 128 + ((c ^ (row & 7)) << 4)) = o;
  }
  __syncwarp();
#pragma unroll
  for (int c = 0; c < 8; ++c) v[c] = *reinterpret_cast<const uint4*>(slice + lane * 128 + ((c ^ (lane & 7)) << 4));
}

// The same load in two halves, so that the global loads can be in flight while the caller waits for something else
// (the GEMM epilogue issues them before tcgen05.wait::ld): issue = coalesced 16-byte loads into registers,
// finish = transpose through the slice into one row per lane.
__device__ __forceinline__ void warp_load_rows128_issue(int lane, uint4 (&t)[8], const void* gbase, long long pitch_bytes,
                                                        int rows_valid) {
  const uint8_t* gp = reinterpret_cast<const uint8_t*>(gbase) + (lane & 7) * 16;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 4 + (lane >> 3);
    t[i] = make_uint4(0, 0, 0, 0);
    if (row < rows_valid) t[i] = __ldg(reinterpret_cast<const uint4*>(gp + row * pitch_bytes));
  }
}
__device__ __forceinline__ void warp_load_rows128_finish(uint8_t* slice, int lane, const uint4 (&t)[8], uint4 (&v)[8]) {
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 4 + (lane >> 3), c = lane & 7;
    *reinterpret_cast<uint4*>(slice + row * 128 + ((c ^ (row & 7)) << 4)) = t[i];
  }
  __syncwarp();
#pragma unroll
  for (int c = 0; c < 8; ++c) v[c] = *reinterpret_cast<const uint4*>(slice + lane * 128 + ((c ^ (lane & 7)) << 4));
}

// ---------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Wait with a watchdog: a pipeline bug traps (visible as a launch failure) instead of hanging.  A warp whose barrier is
// not ready yet sleeps a few tens of nanoseconds between polls: a polling warp is always "ready" to the scheduler, and in
// the warp-specialised row kernels the polls (try_wait + 64-bit clock compare + branch, 8 instructions) were 48 - 55 % of
// all executed instructions (ncu: pair_transition_ws, outer_linear), issued in competition with the warps that had work.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t polls = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(32);
    if (++polls > (1u << 26)) {  // > 2 s of sleeping alone
      printf("prd: mbarrier watchdog (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}

// Lean wait for hot single-warp issue loops: the watchdog version costs ~10 registers and a printf call site per use,
// which makes a 40-register UMMA warp spill.  Use only where the same protocol is already covered by mbar_wait elsewhere.
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "SPIN_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra SPIN_DONE;\n\t"
      "bra SPIN_WAIT;\n\t"
      "SPIN_DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// generic-proxy writes (st.shared) -> visible to the async proxy (TMA store / UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------------------
// bulk copies (no tensor map): global -> shared with mbarrier completion, shared -> global
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------------
// TMA tiled loads (tensor maps are built on the host, see prd_tma.cu)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// shared -> global tile store (bulk async group; the box is clipped at the tensor's bounds)
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(m),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m), "r"(smem_u32(smem_src)),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ---------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------
// Allocate `ncols` (power of two >= 32) TMEM columns; executed by ONE full warp.  The base
// address is written to *smem_slot.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor for a K-major SWIZZLE_128B operand tile (rows x 64 halves per
// K-block, 8-row atoms of 1024 B).  Field layout: cute::UMMA::SmemDescriptor (mma_sm100_desc.hpp):
//   [0,14) start>>4, [16,30) LBO>>4 (=1, unused for swizzled K-major), [32,46) SBO>>4 (=64: 1024 B
//   between 8-row groups), [46,48) version=1, [61,64) layout type (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// One elected lane of a fully converged warp (warp-uniform control flow around it keeps UMMA / TMA operands
// in uniform registers; an `if (lane == 0)` branch costs ~15 instructions per tcgen05.mma instead).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
// Instruction descriptor, kind::f16: fp16 A/B (format 0), fp32 accumulate (c_format 1), both
// operands K-major, M = 128.  Layout: cute::UMMA::InstrDescriptor.
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t M, uint32_t N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// kind::tf32: fp32 words in shared memory, of which the tensor core reads sign + exponent + 10 mantissa bits (format 2).
__host__ __device__ constexpr uint32_t umma_idesc_tf32(uint32_t M, uint32_t N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// Round to nearest tf32 (the tensor core truncates: an unrounded operand carries a systematic -2^-11 bias).
// cvt.rna.tf32.f32 (nearest, ties away from zero) is emulated on sm_100a with ~6 instructions (SASS: FSETP |x| < inf,
// LOP3, IADD3, select); the same result for every finite input in two: add half an ulp of the 10-bit mantissa to the
// magnitude, clear the 13 low bits (inf stays inf, NaN stays NaN).
__device__ __forceinline__ float round_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

// D[tmem] (+)= A[smem] * B[smem]^T, one UMMA (K = 16 halves); issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued UMMAs of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// Issue the UMMAs for one [128 x 64] K-block of A against one [N x 64] K-block of B.
__device__ __forceinline__ void umma_kblock(uint32_t tmem_d, uint32_t a_smem, uint32_t b_smem, uint32_t idesc,
                                            bool accumulate_first) {
  const uint64_t da = umma_desc_sw128(a_smem);
  const uint64_t db = umma_desc_sw128(b_smem);
#pragma unroll
  for (uint32_t k = 0; k < 4; ++k) {
    // advancing 16 halves (32 B) along K inside the 128 B swizzle row = +2 in the (addr >> 4) field
    umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (accumulate_first || k > 0) ? 1u : 0u);
  }
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One [128 x 32] fp32 K-block of A against one [N x 32] K-block of B: four UMMAs of K = 8 (32 bytes each).
__device__ __forceinline__ void umma_kblock_tf32(uint32_t tmem_d, uint32_t a_smem, uint32_t b_smem, uint32_t idesc,
                                                 bool accumulate_first) {
  const uint64_t da = umma_desc_sw128(a_smem);
  const uint64_t db = umma_desc_sw128(b_smem);
#pragma unroll
  for (uint32_t k = 0; k < 4; ++k) umma_tf32(tmem_d, da + 2 * k, db + 2 * k, idesc, (accumulate_first || k > 0) ? 1u : 0u);
}

// TMEM -> registers: 32 lanes (this warp's quarter) x 32 consecutive 32-bit columns.
// taddr = base + (lane_base << 16) + column.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  __syncwarp();  // .sync.aligned: the whole warp must issue together
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  __syncwarp();
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Wait for outstanding TMEM loads and tie the wait to the destination registers ("+r"), so the compiler
// cannot schedule uses of r[] ahead of the wait when loads are software pipelined.
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// ---------------------------------------------------------------------------------------
// host: TMA tensor-map construction (prd_tma.cu).  fp16 or fp32 elements, up to 4 dims,
// innermost box = 128 bytes with SWIZZLE_128B (matches umma_desc_sw128 / sw128_offset).
// ---------------------------------------------------------------------------------------
struct TmaDims {
  uint64_t size[4];    // elements per dim, dim 0 innermost (contiguous)
  uint64_t stride[3];  // BYTES between consecutive indices of dims 1..3 (multiples of 16)
  uint32_t box[4];     // box extent per dim (elements)
};
int make_tensor_map(CUtensorMap* out, const void* base, int elem_bytes, int rank, const TmaDims& d, bool swizzle128);
// mode: 0 none, 1 SWIZZLE_128B, 2 SWIZZLE_128B_ATOM_32B (32-byte swizzle chunks: MN-major 32-bit UMMA operands)
int make_tensor_map_mode(CUtensorMap* out, const void* base, int elem_bytes, int rank, const TmaDims& d, int mode);

}  // namespace prd
