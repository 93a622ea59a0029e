// Host-side helpers shared by all translation units of libosufusion_sm100.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include <atomic>

#include "../../include/osufusion_b200.h"

namespace ofx {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launch_count;
inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

int device_sm_count();

// cuTensorMapEncodeTiled resolved at run time through cudart (no link-time libcuda dependency, so the
// library loads on a GPU-less build box).
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled();

// bf16 tensor map, SWIZZLE_128B, zero OOB fill. rank in [3,4]. dims/strides innermost first; strides[i] is the
// byte stride of dim i+1.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const unsigned long long* dims,
                   const unsigned long long* strides_bytes, const unsigned* box);

// Launch with programmatic stream serialisation (PDL): the kernel may be scheduled while the previous kernel in the stream is
// still draining; it must execute griddepcontrol.wait (ptx.cuh: pdl_wait) before touching global memory.  OF_PDL=0 disables it.
// OF_PDL = 0: off; 1 (default): tensor-core kernels (GEMM, attention: they have a real prologue to hide); 2: every kernel.
int pdl_level();
template <int kLevel = 2, typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_level() >= kLevel ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define OF_CHECK_CUDA(expr)                                                                \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ofx::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return OF_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

#define OF_REQUIRE(cond, ...)         \
  do {                                \
    if (!(cond)) {                    \
      ofx::set_error(__VA_ARGS__);    \
      return OF_ERR_INVALID;          \
    }                                 \
  } while (0)

}  // namespace ofx
