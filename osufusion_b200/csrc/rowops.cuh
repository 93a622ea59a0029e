// Device helpers for the bandwidth-bound channels-last kernels: 16-byte vector access (8 bf16 / 4 fp32),
// block/warp reductions, GroupNorm(1,C)+FiLM+SiLU recompute.
#pragma once
#include "ptx.cuh"

namespace ofx {

struct V8 {
  float v[8];
};

__device__ __forceinline__ V8 ld_bf16x8(const __nv_bfloat16* p) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  V8 r;
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y; r.v[4] = c.x; r.v[5] = c.y; r.v[6] = d.x; r.v[7] = d.y;
  return r;
}
__device__ __forceinline__ void st_bf16x8(__nv_bfloat16* p, const V8& r) {
  uint4 u;
  u.x = pack_bf16x2(r.v[0], r.v[1]); u.y = pack_bf16x2(r.v[2], r.v[3]);
  u.z = pack_bf16x2(r.v[4], r.v[5]); u.w = pack_bf16x2(r.v[6], r.v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ V8 ld_f32x8(const float* p) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  V8 r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void st_f32x8(float* p, const V8& r) {
  *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
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
// Sum over the whole block (<= 1024 threads); result valid in every thread. `sm` must hold 32 floats.
__device__ __forceinline__ float block_sum(float v, float* sm) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) sm[w] = v;
  __syncthreads();
  float r = (l < nw) ? sm[l] : 0.f;
  return warp_sum(r);
}
__device__ __forceinline__ float block_max(float v, float* sm) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) sm[w] = v;
  __syncthreads();
  float r = (l < nw) ? sm[l] : -INFINITY;
  return warp_max(r);
}

// Accurate SiLU for the bandwidth kernels (expf, not __expf: these feed parity-critical tensors).
__device__ __forceinline__ float silu_acc(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float dsilu_acc(float x) {
  float s = 1.0f / (1.0f + expf(-x));
  return s * (1.0f + x * (1.0f - s));
}

// GroupNorm(1, C) statistics of one sample from the (sum, sumsq) pair accumulated by the conv epilogue.
__device__ __forceinline__ void gn_mean_rstd(const double* stats, int b, double n, float eps, float& mean, float& rstd) {
  double s1 = stats[2 * b], s2 = stats[2 * b + 1];
  double m = s1 / n;
  double var = s2 / n - m * m;
  if (var < 0) var = 0;
  mean = (float)m;
  rstd = (float)(1.0 / sqrt(var + (double)eps));
}

}  // namespace ofx
