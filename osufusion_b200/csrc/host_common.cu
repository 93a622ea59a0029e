#include <stdlib.h>
#include "host_common.h"

#include <string.h>

#include <mutex>

namespace ofx {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launch_count{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int pdl_level() {
  static const int level = [] {
    const char* e = getenv("OF_PDL");
    return (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
  }();
  return level;
}

static std::atomic<int> g_sm_limit{0};

// Persistent kernels (the GEMM) size their grids from this.  When NCCL collectives overlap the backward pass their CTAs occupy
// SMs for a millisecond at a time; a persistent grid that assumes all 148 SMs then runs as TWO waves (the CTAs that found no SM
// start only when the first ones exit).  The data-parallel wrapper therefore reserves SMs for the collective through this limit.
extern "C" int of_set_sm_limit(int n) {
  g_sm_limit.store(n < 0 ? 0 : n, std::memory_order_relaxed);
  return OF_OK;
}

int device_sm_count() {
  const int lim = g_sm_limit.load(std::memory_order_relaxed);
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  if (lim > 0 && lim < cached[dev]) return lim;
  return cached[dev];
}

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const unsigned long long* dims,
                   const unsigned long long* strides_bytes, const unsigned* box) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return OF_ERR_CUDA;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    if (gstr[i] % 16 != 0) {
      set_error("TMA stride %d = %llu bytes is not a multiple of 16", i, (unsigned long long)gstr[i]);
      return OF_ERR_INVALID;
    }
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_error("TMA base pointer %p is not 16-byte aligned", base);
    return OF_ERR_INVALID;
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu,%llu box %u,%u,%u)", (int)r,
              rank, dims[0], dims[1], dims[2], box[0], box[1], box[2]);
    return OF_ERR_CUDA;
  }
  return OF_OK;
}

}  // namespace ofx

extern "C" {
const char* of_last_error(void) { return ofx::g_err; }
int of_version(void) { return 1; }
long long of_launch_count(void) { return ofx::g_launch_count.load(); }
void of_reset_launch_count(void) { ofx::g_launch_count.store(0); }
}
