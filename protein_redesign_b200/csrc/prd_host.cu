// Host-side plumbing shared by every entry point: thread-local error string, version query,
// and TMA tensor-map construction through the driver entry point (resolved at run time with
// cudaGetDriverEntryPoint, so the library has no link-time dependency on libcuda).
#include <stdarg.h>
#include <atomic>
#include <mutex>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/prd_denoiser.h"
#include "prd_common.cuh"

namespace prd {

static thread_local char g_error[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static const bool on = getenv("PRD_PDL") && getenv("PRD_PDL")[0] == '1';  // measured: no gain under graph replay (profiles/r02_pdl.md), so opt-in
  return on;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("CUDA error %s (%s) at %s", cudaGetErrorName(e), cudaGetErrorString(e), what);
  return 1;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

int make_tensor_map(CUtensorMap* out, const void* base, int elem_bytes, int rank, const TmaDims& d, bool swizzle128) {
  return make_tensor_map_mode(out, base, elem_bytes, rank, d, swizzle128 ? 1 : 0);
}

// mode: 0 none, 1 SWIZZLE_128B (16-byte chunks), 2 SWIZZLE_128B_ATOM_32B (32-byte chunks: MN-major 32-bit UMMA operands)
int make_tensor_map_mode(CUtensorMap* out, const void* base, int elem_bytes, int rank, const TmaDims& d, int mode) {
  const bool swizzle128 = mode != 0;
  EncodeTiledFn fn = get_encode_fn();
  PRD_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  PRD_REQUIRE(rank >= 2 && rank <= 4, "tensor map rank %d unsupported", rank);
  PRD_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map base not 16-byte aligned");
  cuuint64_t size[4];
  cuuint64_t stride[3];
  cuuint32_t box[4];
  cuuint32_t estr[4];
  for (int i = 0; i < rank; ++i) {
    size[i] = d.size[i];
    box[i] = d.box[i];
    estr[i] = 1;
    PRD_REQUIRE(d.size[i] > 0 && d.box[i] > 0 && d.box[i] <= 256, "bad tensor map dim %d (size %llu box %u)", i,
                (unsigned long long)d.size[i], d.box[i]);
  }
  for (int i = 0; i + 1 < rank; ++i) {
    stride[i] = d.stride[i];
    PRD_REQUIRE((d.stride[i] & 15) == 0, "tensor map stride %d (%llu B) not a multiple of 16", i,
                (unsigned long long)d.stride[i]);
  }
  if (swizzle128) PRD_REQUIRE(d.box[0] * elem_bytes == 128, "swizzle-128 box must span 128 bytes");
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), size, stride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mode == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : (mode == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PRD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

}  // namespace prd

extern "C" {

int prd_version(void) { return PRD_VERSION; }

const char* prd_last_error(void) { return prd::g_error; }

long long prd_launch_count(void) { return prd::g_launches.load(); }

int prd_device_check(void) {
  static std::atomic<int> cached[64];  // per device: 0 unknown, 1 ok (zero-initialised: static storage)
  int dev = 0;
  if (prd::check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return 1;
  if (dev >= 0 && dev < 64 && cached[dev] == 1) return 0;
  int major = 0, minor = 0;
  if (prd::check_cuda(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev), "cudaDeviceGetAttribute")) return 1;
  if (prd::check_cuda(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev), "cudaDeviceGetAttribute")) return 1;
  if (major != 10) {
    prd::set_error("libprd_sm100 needs an sm_100 (B200) device, found sm_%d%d", major, minor);
    return 1;
  }
  if (dev >= 0 && dev < 64) cached[dev] = 1;
  return 0;
}

}  // extern "C"
