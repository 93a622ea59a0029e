// LoRA / DoRA adapter kernels (reference: osu_fusion/modules/lora_layers.py:15-96,292-328 for Conv1d; peft 0.12.0
// DoraLinearLayer for nn.Linear).  The adapted layer is evaluated as ONE GEMM with the effective weight
//     W_eff[co, e] = s[co] * (W[co, e] + scaling * sum_r B[co, r] * A[r, e]),   s = magnitude / ||W + scaling*B*A||_2 (detached)
// (e runs over (ci, tap)), which is algebraically the reference's  base(x) + (s-1)*conv(x, W) + s*scaling*B(A(x)).
// Backward: the weight-gradient GEMM produces dW_eff; of_dora_grad projects it onto dA, dB and d(magnitude).
#include "host_common.h"
#include "rowops.cuh"

namespace ofx {

constexpr int kDoraE = 256;   // e-columns per CTA (one per thread)
constexpr int kDoraCo = 16;   // output channels per CTA

struct DoraArgs {
  const float* W;   // (Cout, E)   E = Cin*k, torch layout (ci-major, tap-minor)
  const float* A;   // (r, E)
  const float* B;   // (Cout, r)
  const float* mag; // (Cout) or nullptr (plain LoRA: s = 1)
  float scaling;
  int Cout, Cin, k, r, Cin_pad;
};

__device__ __forceinline__ void dora_load_tiles(const DoraArgs& a, int e0, int co0, float* sA, float* sB) {
  const int E = a.Cin * a.k;
  for (int i = threadIdx.x; i < a.r * kDoraE; i += blockDim.x) {
    const int r = i / kDoraE, e = e0 + (i - r * kDoraE);
    sA[i] = e < E ? a.A[(long long)r * E + e] : 0.f;
  }
  for (int i = threadIdx.x; i < kDoraCo * a.r; i += blockDim.x) {
    const int c = i / a.r, co = co0 + c;
    sB[i] = co < a.Cout ? a.B[(long long)co * a.r + (i - c * a.r)] : 0.f;
  }
  __syncthreads();
}

// n2[co] += sum_e (W + scaling*BA)^2
__global__ void __launch_bounds__(kDoraE) dora_norm_kernel(const DoraArgs a, float* __restrict__ n2) {
  extern __shared__ float sm[];
  float* sA = sm;
  float* sB = sA + a.r * kDoraE;
  float* sN = sB + kDoraCo * a.r;
  const int E = a.Cin * a.k;
  const int e0 = blockIdx.x * kDoraE, co0 = blockIdx.y * kDoraCo;
  if (threadIdx.x < kDoraCo) sN[threadIdx.x] = 0.f;
  dora_load_tiles(a, e0, co0, sA, sB);
  const int e = e0 + threadIdx.x;
  for (int c = 0; c < kDoraCo; ++c) {
    const int co = co0 + c;
    float sq = 0.f;
    if (co < a.Cout && e < E) {
      float d = 0.f;
      for (int r = 0; r < a.r; ++r) d += sB[c * a.r + r] * sA[r * kDoraE + threadIdx.x];
      const float v = a.W[(long long)co * E + e] + a.scaling * d;
      sq = v * v;
    }
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sN[c], sq);
  }
  __syncthreads();
  if (threadIdx.x < kDoraCo && co0 + threadIdx.x < a.Cout) atomicAdd(n2 + co0 + threadIdx.x, sN[threadIdx.x]);
}

// packed[t][co][ci] = bf16(s[co] * (W + scaling*BA)[co, ci, t]);  s_out[co] = s[co]
__global__ void __launch_bounds__(kDoraE) dora_merge_kernel(const DoraArgs a, const float* __restrict__ n2,
                                                            __nv_bfloat16* __restrict__ packed, long long tap_stride,
                                                            float* __restrict__ s_out) {
  extern __shared__ float sm[];
  float* sA = sm;
  float* sB = sA + a.r * kDoraE;
  const int E = a.Cin * a.k;
  const int e0 = blockIdx.x * kDoraE, co0 = blockIdx.y * kDoraCo;
  dora_load_tiles(a, e0, co0, sA, sB);
  const int e = e0 + threadIdx.x;
  const int ci = e / a.k, t = e - ci * a.k;
  for (int c = 0; c < kDoraCo; ++c) {
    const int co = co0 + c;
    if (co >= a.Cout) break;
    const float s = a.mag ? a.mag[co] * rsqrtf(n2[co]) : 1.0f;
    if (blockIdx.x == 0 && threadIdx.x == 0 && s_out) s_out[co] = s;
    if (e < E) {
      float d = 0.f;
      for (int r = 0; r < a.r; ++r) d += sB[c * a.r + r] * sA[r * kDoraE + threadIdx.x];
      const float v = a.W[(long long)co * E + e] + a.scaling * d;
      packed[(long long)t * tap_stride + (long long)co * a.Cin_pad + ci] = __float2bfloat16_rn(s * v);
    }
  }
}

// From dWp = d(loss)/d(W_eff) in the packed [t][co][ci] fp32 layout:
//   dB[co, r] += scaling * s[co] * sum_e dWp[co,e] * A[r,e]
//   dA[r, e]  += scaling * sum_co B[co,r] * s[co] * dWp[co,e]
//   dmag[co]  += sum_e dWp[co,e] * (W + scaling*BA)[co,e] / n[co]          (n = ||W + scaling*BA||, detached)
__global__ void __launch_bounds__(kDoraE) dora_grad_kernel(const DoraArgs a, const float* __restrict__ n2,
                                                           const float* __restrict__ dWp, long long tap_stride,
                                                           float* __restrict__ dA, float* __restrict__ dB,
                                                           float* __restrict__ dmag) {
  extern __shared__ float sm[];
  float* sA = sm;                           // [r][kDoraE]
  float* sB = sA + a.r * kDoraE;            // [kDoraCo][r]
  float* sG = sB + kDoraCo * a.r;           // [kDoraCo][kDoraE]  G = scaling * s * dWp
  float* sM = sG + kDoraCo * kDoraE;        // [kDoraCo]
  const int E = a.Cin * a.k;
  const int e0 = blockIdx.x * kDoraE, co0 = blockIdx.y * kDoraCo;
  if (threadIdx.x < kDoraCo) sM[threadIdx.x] = 0.f;
  dora_load_tiles(a, e0, co0, sA, sB);
  const int e = e0 + threadIdx.x;
  const int ci = e / a.k, t = e - ci * a.k;
  for (int c = 0; c < kDoraCo; ++c) {
    const int co = co0 + c;
    float g = 0.f, gv = 0.f, s = 1.f;
    if (co < a.Cout && e < E) {
      g = dWp[(long long)t * tap_stride + (long long)co * a.Cin_pad + ci];
      if (a.mag) {
        const float inv_n = rsqrtf(n2[co]);
        s = a.mag[co] * inv_n;
        float d = 0.f;
        for (int r = 0; r < a.r; ++r) d += sB[c * a.r + r] * sA[r * kDoraE + threadIdx.x];
        gv = g * (a.W[(long long)co * E + e] + a.scaling * d) * inv_n;
      }
    }
    sG[c * kDoraE + threadIdx.x] = a.scaling * s * g;
    if (a.mag) {
      gv = warp_sum(gv);
      if ((threadIdx.x & 31) == 0) atomicAdd(&sM[c], gv);
    }
  }
  __syncthreads();
  // dA: thread e, loop r, sum over the CTA's output channels
  if (e < E) {
    for (int r = 0; r < a.r; ++r) {
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < kDoraCo; ++c) acc += sB[c * a.r + r] * sG[c * kDoraE + threadIdx.x];
      atomicAdd(dA + (long long)r * E + e, acc);
    }
  }
  // dB: kDoraCo x r outputs, each a dot product over the CTA's e-columns
  for (int o = threadIdx.x; o < kDoraCo * a.r; o += blockDim.x) {
    const int c = o / a.r, r = o - c * a.r;
    if (co0 + c >= a.Cout) continue;
    float acc = 0.f;
    for (int j = 0; j < kDoraE; ++j) acc += sG[c * kDoraE + j] * sA[r * kDoraE + j];
    atomicAdd(dB + (long long)(co0 + c) * a.r + r, acc);
  }
  if (a.mag && threadIdx.x < kDoraCo && co0 + threadIdx.x < a.Cout) atomicAdd(dmag + co0 + threadIdx.x, sM[threadIdx.x]);
}

static int dora_check(const float* W, const float* A, const float* B, int Cout, int Cin, int k, int r, const char* who) {
  OF_REQUIRE(W && A && B, "%s: null pointer", who);
  OF_REQUIRE(Cout >= 1 && Cin >= 1 && k >= 1 && r >= 1 && r <= 128, "%s: bad sizes (r=%d)", who, r);
  return OF_OK;
}

}  // namespace ofx

using namespace ofx;

extern "C" int of_dora_merge(const float* W, const float* A, const float* B, const float* mag, float scaling, int Cout, int Cin,
                             int k, int r, float* n2_ws, void* packed_bf16, int Cin_pad, long long tap_stride, float* s_out,
                             void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = dora_check(W, A, B, Cout, Cin, k, r, "of_dora_merge");
  if (rc) return rc;
  OF_REQUIRE(packed_bf16 && n2_ws && Cin_pad >= Cin, "of_dora_merge: bad outputs");
  DoraArgs a{W, A, B, mag, scaling, Cout, Cin, k, r, Cin_pad};
  const int E = Cin * k;
  dim3 grid((E + kDoraE - 1) / kDoraE, (Cout + kDoraCo - 1) / kDoraCo);
  size_t smem = ((size_t)r * kDoraE + (size_t)kDoraCo * r + kDoraCo) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    OF_CHECK_CUDA(cudaFuncSetAttribute(dora_norm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    OF_CHECK_CUDA(cudaFuncSetAttribute(dora_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    OF_CHECK_CUDA(cudaFuncSetAttribute(dora_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  if (mag) {
    OF_CHECK_CUDA(cudaMemsetAsync(n2_ws, 0, (size_t)Cout * sizeof(float), stream));
    dora_norm_kernel<<<grid, kDoraE, smem, stream>>>(a, n2_ws);
    OF_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }
  dora_merge_kernel<<<grid, kDoraE, smem, stream>>>(a, n2_ws, reinterpret_cast<__nv_bfloat16*>(packed_bf16), tap_stride, s_out);
  OF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OF_OK;
}

extern "C" int of_dora_grad(const float* W, const float* A, const float* B, const float* mag, float scaling, int Cout, int Cin,
                            int k, int r, const float* n2, const float* dW_packed, int Cin_pad, long long tap_stride, float* dA,
                            float* dB, float* dmag, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = dora_check(W, A, B, Cout, Cin, k, r, "of_dora_grad");
  if (rc) return rc;
  OF_REQUIRE(dW_packed && dA && dB && (!mag || (dmag && n2)), "of_dora_grad: null pointer");
  DoraArgs a{W, A, B, mag, scaling, Cout, Cin, k, r, Cin_pad};
  const int E = Cin * k;
  dim3 grid((E + kDoraE - 1) / kDoraE, (Cout + kDoraCo - 1) / kDoraCo);
  size_t smem = ((size_t)r * kDoraE + (size_t)kDoraCo * r + (size_t)kDoraCo * kDoraE + kDoraCo) * sizeof(float);
  dora_grad_kernel<<<grid, kDoraE, smem, stream>>>(a, n2, dW_packed, tap_stride, dA, dB, dmag);
  OF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OF_OK;
}
