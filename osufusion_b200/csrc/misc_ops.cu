// Remaining bandwidth-bound / tiny kernels of the denoiser hot path.  Contracts and reference citations are in
// include/osufusion_b200.h next to each of_* declaration.
#include <stdlib.h>

#include "host_common.h"
#include "rowops.cuh"

namespace ofx {

// =================================================================================================== LayerNorm
// warp per row; lane owns 16-byte vectors lane + 32*i (i < NV); row kept in registers.
template <int NV>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, long long x_ld, int rows, int C,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float eps, float* __restrict__ out_f32,
                                                            __nv_bfloat16* __restrict__ out_bf16, long long out_ld,
                                                            float* __restrict__ mean_rstd) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const int vecs = C >> 3;
  V8 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < vecs) {
      v[i] = ld_f32x8(x + (long long)row * x_ld + vi * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i].v[j];
    }
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < vecs) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float d = v[i].v[j] - mean;
        q += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
  if (lane == 0 && mean_rstd) {
    mean_rstd[2 * row] = mean;
    mean_rstd[2 * row + 1] = rstd;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < vecs) {
      V8 g = ld_f32x8(gamma + vi * 8), bt = ld_f32x8(beta + vi * 8), o;
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] = (v[i].v[j] - mean) * rstd * g.v[j] + bt.v[j];
      if (out_f32) st_f32x8(out_f32 + (long long)row * out_ld + vi * 8, o);
      if (out_bf16) st_bf16x8(out_bf16 + (long long)row * out_ld + vi * 8, o);
    }
  }
}

template <int NV>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, long long dy_ld,
                                                            const float* __restrict__ x, long long x_ld, int rows, int C,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ mean_rstd, float* __restrict__ dx_f32,
                                                            __nv_bfloat16* __restrict__ dx_bf16, long long dx_ld,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                            int rows_per_warp) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int vecs = C >> 3;
  V8 g[NV], dg[NV], db[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < vecs) g[i] = ld_f32x8(gamma + vi * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) dg[i].v[j] = db[i].v[j] = 0.f;
  }
  const int row0 = (blockIdx.x * 8 + warp) * rows_per_warp;
  constexpr bool kPrefetch = NV <= 2;   // C <= 512: the next row's operands are loaded while this row is reduced
  V8 xv[NV], dv[NV];
  if (kPrefetch && row0 < rows) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < vecs) {
        xv[i] = ld_f32x8(x + (long long)row0 * x_ld + vi * 8);
        dv[i] = ld_f32x8(dy + (long long)row0 * dy_ld + vi * 8);
      }
    }
  }
  for (int r = 0; r < rows_per_warp; ++r) {
    const int row = row0 + r;
    if (row >= rows) break;
    const float mean = mean_rstd[2 * row], rstd = mean_rstd[2 * row + 1];
    V8 xh[NV], dxh[NV], xn[NV], dn[NV];
    const bool has_next = kPrefetch && (r + 1 < rows_per_warp) && (row + 1 < rows);
    if (has_next) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = lane + 32 * i;
        if (vi < vecs) {
          xn[i] = ld_f32x8(x + (long long)(row + 1) * x_ld + vi * 8);
          dn[i] = ld_f32x8(dy + (long long)(row + 1) * dy_ld + vi * 8);
        }
      }
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < vecs) {
        if (!kPrefetch) {
          xv[i] = ld_f32x8(x + (long long)row * x_ld + vi * 8);
          dv[i] = ld_f32x8(dy + (long long)row * dy_ld + vi * 8);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i].v[j] = (xv[i].v[j] - mean) * rstd;
          dg[i].v[j] += dv[i].v[j] * xh[i].v[j];
          db[i].v[j] += dv[i].v[j];
          dxh[i].v[j] = dv[i].v[j] * g[i].v[j];
          s1 += dxh[i].v[j];
          s2 += dxh[i].v[j] * xh[i].v[j];
        }
      }
    }
    s1 = warp_sum(s1) / C;
    s2 = warp_sum(s2) / C;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < vecs) {
        V8 o;
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] = rstd * (dxh[i].v[j] - s1 - xh[i].v[j] * s2);
        if (dx_f32) st_f32x8(dx_f32 + (long long)row * dx_ld + vi * 8, o);
        if (dx_bf16) st_bf16x8(dx_bf16 + (long long)row * dx_ld + vi * 8, o);
      }
    }
    if (has_next) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        xv[i] = xn[i];
        dv[i] = dn[i];
      }
    }
  }
  // CTA-level reduction in shared memory, then ONE global atomic per channel per CTA
  extern __shared__ float s_red[];  // [2][C]
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) s_red[c] = 0.f;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < vecs) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&s_red[vi * 8 + j], dg[i].v[j]);
        atomicAdd(&s_red[C + vi * 8 + j], db[i].v[j]);
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(dgamma + c, s_red[c]);
    atomicAdd(dbeta + c, s_red[C + c]);
  }
}

// =================================================================================================== RoPE
// qkv (B, L, (H+2*KVH)*D) bf16; rotate the H q-slots and KVH k-slots in place, bf16 arithmetic as the reference
// (x*cos, rotate_half(x)*sin and their sum are each rounded to bf16: utils.py:25-32 under autocast).
__device__ __forceinline__ V8 ld_tab(const __nv_bfloat16* p) { return ld_bf16x8(p); }
__device__ __forceinline__ V8 ld_tab(const float* p) { return ld_f32x8(p); }

// TAB = __nv_bfloat16: tables in q's autocast dtype (attention.py:37-41) and bf16 intermediate rounding.
// TAB = float: q is fp32 in the reference (DoRA-adapted to_q promotes its output to fp32), so tables and arithmetic are fp32.
template <typename TAB>
__global__ void rope_fwd_kernel(__nv_bfloat16* qkv, long long ld, long long bs, int B, int L, int slots, int D,
                                const TAB* __restrict__ cosT, const TAB* __restrict__ sinT) {
  pdl_launch_dependents();
  pdl_wait();
  const int tps = D >> 4;  // threads per slot (each handles 8 + 8 paired channels)
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * L * slots * tps;
  if (idx >= total) return;
  const int tp = (int)(idx % tps);
  long long r = idx / tps;
  const int slot = (int)(r % slots);
  r /= slots;
  const int l = (int)(r % L), b = (int)(r / L);
  const int half = D >> 1;
  __nv_bfloat16* base = qkv + b * bs + (long long)l * ld + (long long)slot * D;
  V8 x1 = ld_bf16x8(base + tp * 8), x2 = ld_bf16x8(base + half + tp * 8);
  V8 c1 = ld_tab(cosT + (long long)l * D + tp * 8), c2 = ld_tab(cosT + (long long)l * D + half + tp * 8);
  V8 s1 = ld_tab(sinT + (long long)l * D + tp * 8), s2 = ld_tab(sinT + (long long)l * D + half + tp * 8);
  V8 o1, o2;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (sizeof(TAB) == 2) {
      o1.v[j] = bf16_round(x1.v[j] * c1.v[j]) + bf16_round(-x2.v[j] * s1.v[j]);
      o2.v[j] = bf16_round(x2.v[j] * c2.v[j]) + bf16_round(x1.v[j] * s2.v[j]);
    } else {
      o1.v[j] = x1.v[j] * c1.v[j] - x2.v[j] * s1.v[j];
      o2.v[j] = x2.v[j] * c2.v[j] + x1.v[j] * s2.v[j];
    }
  }
  st_bf16x8(base + tp * 8, o1);
  st_bf16x8(base + half + tp * 8, o2);
}

// dq32 (B,L,H*D), dk32/dv32 (B,L,KVH*D) fp32 -> dqkv16 (B,L,(H+2KVH)*D) bf16, undoing the rotation on q and k slots.
template <typename TAB>
__global__ void rope_bwd_kernel(const float* __restrict__ dq, long long dq_ld, long long dq_bs, const float* __restrict__ dk,
                                const float* __restrict__ dv, long long dkv_ld, long long dkv_bs, __nv_bfloat16* out,
                                long long out_ld, long long out_bs, int B, int L, int H, int KVH, int D,
                                const TAB* __restrict__ cosT, const TAB* __restrict__ sinT) {
  pdl_launch_dependents();
  pdl_wait();
  const int tps = D >> 4;
  const int slots = H + 2 * KVH;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * L * slots * tps;
  if (idx >= total) return;
  const int tp = (int)(idx % tps);
  long long r = idx / tps;
  const int slot = (int)(r % slots);
  r /= slots;
  const int l = (int)(r % L), b = (int)(r / L);
  const int half = D >> 1;
  const float* src;
  if (slot < H) src = dq + b * dq_bs + (long long)l * dq_ld + (long long)slot * D;
  else if (slot < H + KVH) src = dk + b * dkv_bs + (long long)l * dkv_ld + (long long)(slot - H) * D;
  else src = dv + b * dkv_bs + (long long)l * dkv_ld + (long long)(slot - H - KVH) * D;
  V8 g1 = ld_f32x8(src + tp * 8), g2 = ld_f32x8(src + half + tp * 8), o1, o2;
  if (slot < H + KVH) {
    V8 c1 = ld_tab(cosT + (long long)l * D + tp * 8), c2 = ld_tab(cosT + (long long)l * D + half + tp * 8);
    V8 s1 = ld_tab(sinT + (long long)l * D + tp * 8), s2 = ld_tab(sinT + (long long)l * D + half + tp * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // y1 = x1*c1 - x2*s1 ; y2 = x2*c2 + x1*s2  =>  dx1 = g1*c1 + g2*s2 ; dx2 = g2*c2 - g1*s1
      o1.v[j] = g1.v[j] * c1.v[j] + g2.v[j] * s2.v[j];
      o2.v[j] = g2.v[j] * c2.v[j] - g1.v[j] * s1.v[j];
    }
  } else {
    o1 = g1;
    o2 = g2;
  }
  __nv_bfloat16* dst = out + b * out_bs + (long long)l * out_ld + (long long)slot * D;
  st_bf16x8(dst + tp * 8, o1);
  st_bf16x8(dst + half + tp * 8, o2);
}

// =================================================================================================== small-M linear
constexpr int kMaxM = 16;
// one warp per output feature n; weights fp32 (master) rounded to bf16 on the fly when round_bf16 != 0.
template <int MM>
__global__ void __launch_bounds__(256) linear_small_fwd_kernel(const float* __restrict__ x, long long x_ld, int M, int N, int K,
                                                               const float* __restrict__ W, long long w_ld,
                                                               const float* __restrict__ bias, int act, int round_bf16,
                                                               float* __restrict__ y, long long y_ld,
                                                               float* __restrict__ ypre) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  if (n >= N) return;
  float acc[MM];
#pragma unroll
  for (int m = 0; m < MM; ++m) acc[m] = 0.f;
  const float* w = W + (long long)n * w_ld;
  if ((K & 3) == 0 && (w_ld & 3) == 0 && (x_ld & 3) == 0) {
#pragma unroll 4
    for (int k = lane * 4; k < K; k += 128) {
      float4 wv = *reinterpret_cast<const float4*>(w + k);
      if (round_bf16) { wv.x = bf16_round(wv.x); wv.y = bf16_round(wv.y); wv.z = bf16_round(wv.z); wv.w = bf16_round(wv.w); }
#pragma unroll
      for (int m = 0; m < MM; ++m) {
        if (m < M) {
          float4 xv = *reinterpret_cast<const float4*>(x + (long long)m * x_ld + k);
          if (round_bf16) { xv.x = bf16_round(xv.x); xv.y = bf16_round(xv.y); xv.z = bf16_round(xv.z); xv.w = bf16_round(xv.w); }
          acc[m] += wv.x * xv.x + wv.y * xv.y + wv.z * xv.z + wv.w * xv.w;
        }
      }
    }
  } else {
    for (int k = lane; k < K; k += 32) {
      float wv = w[k];
      if (round_bf16) wv = bf16_round(wv);
#pragma unroll
      for (int m = 0; m < MM; ++m) {
        if (m < M) {
          float xv = x[(long long)m * x_ld + k];
          if (round_bf16) xv = bf16_round(xv);
          acc[m] += wv * xv;
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MM; ++m) {
    if (m < M) {
      float v = warp_sum(acc[m]);
      if (lane == 0) {
        v += bias ? bias[n] : 0.f;
        if (round_bf16) v = bf16_round(v);
        if (ypre) ypre[(long long)m * y_ld + n] = v;
        if (act == 1) v = silu_acc(v);
        else if (act == 2) v = 1.0f / (1.0f + expf(-v));
        if (round_bf16 && act != 0) v = bf16_round(v);
        y[(long long)m * y_ld + n] = v;
      }
    }
  }
}

// dpre[m,n] = dy[m,n] * act'(ypre[m,n]);  dW[n,k] (+)= sum_m dpre*x ; db[n] (+)= sum_m dpre ; dx[m,k] += sum_n dpre*W (atomic)
// CTA = kNChunk output rows x a k-slice.  Threads are laid out as `kthreads` k-lanes (4 consecutive k each, or 1) times
// `nlanes` row-lanes, so small K (gate MLPs: K = C/2) still fills the CTA; every thread keeps 4 rows of W (and dW) in flight.
constexpr int kNChunk = 16;
template <int MM>
__global__ void __launch_bounds__(256) linear_small_bwd_kernel(const float* __restrict__ dy, long long dy_ld,
                                                               const float* __restrict__ ypre, int act,
                                                               const float* __restrict__ x, long long x_ld, int M, int N, int K,
                                                               const float* __restrict__ W, long long w_ld, int round_bf16,
                                                               float* __restrict__ dW, float* __restrict__ dbias,
                                                               float* __restrict__ dx, long long dx_ld, int kthreads, int accumulate) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sdpre[kNChunk][kMaxM];
  const int n0 = blockIdx.y * kNChunk;
  const int nn = min(kNChunk, N - n0);
  for (int i = threadIdx.x; i < kNChunk * kMaxM; i += blockDim.x) {
    const int j = i / kMaxM, m = i % kMaxM;
    float v = 0.f;
    if (j < nn && m < M) {
      v = dy[(long long)m * dy_ld + n0 + j];
      if (act == 1) v *= dsilu_acc(ypre[(long long)m * dy_ld + n0 + j]);
      else if (act == 2) {
        float s = 1.0f / (1.0f + expf(-ypre[(long long)m * dy_ld + n0 + j]));
        v *= s * (1.0f - s);
      }
    }
    sdpre[j][m] = v;
  }
  __syncthreads();
  if (dbias && blockIdx.x == 0 && threadIdx.x < nn) {
    float s = 0.f;
    for (int m = 0; m < M; ++m) s += sdpre[threadIdx.x][m];
    if (accumulate) dbias[n0 + threadIdx.x] += s;
    else dbias[n0 + threadIdx.x] = s;
  }
  const bool vec4 = ((K & 3) == 0) && ((w_ld & 3) == 0) && ((x_ld & 3) == 0) && ((dx_ld & 3) == 0);
  const int kw = vec4 ? 4 : 1;
  const int nlanes = blockDim.x / kthreads;
  const int kt = threadIdx.x % kthreads, nl = threadIdx.x / kthreads;
  const int k = (blockIdx.x * kthreads + kt) * kw;
  if (k >= K) return;
  float xa[MM][4];
  float dxa[MM][4];
#pragma unroll
  for (int m = 0; m < MM; ++m) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      dxa[m][e] = 0.f;
      xa[m][e] = 0.f;
      if (m < M && e < kw) {
        float v = x[(long long)m * x_ld + k + e];
        xa[m][e] = round_bf16 ? bf16_round(v) : v;
      }
    }
  }
  for (int j0 = nl; j0 < nn; j0 += 4 * nlanes) {
    float wv[4][4], gv[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int j = j0 + r * nlanes;
      const long long wo = (long long)(n0 + j) * w_ld + k;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        wv[r][e] = 0.f;
        gv[r][e] = 0.f;
      }
      if (j < nn) {
        if (vec4) {
          const float4 t = *reinterpret_cast<const float4*>(W + wo);
          wv[r][0] = t.x; wv[r][1] = t.y; wv[r][2] = t.z; wv[r][3] = t.w;
          if (dW && accumulate) {
            const float4 u = *reinterpret_cast<const float4*>(dW + wo);
            gv[r][0] = u.x; gv[r][1] = u.y; gv[r][2] = u.z; gv[r][3] = u.w;
          }
        } else {
          wv[r][0] = W[wo];
          if (dW && accumulate) gv[r][0] = dW[wo];
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int j = j0 + r * nlanes;
      if (j < nn) {
        const long long wo = (long long)(n0 + j) * w_ld + k;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (round_bf16) wv[r][e] = bf16_round(wv[r][e]);
#pragma unroll
        for (int m = 0; m < MM; ++m) {
          if (m < M) {
            const float d = sdpre[j][m];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              gv[r][e] = fmaf(d, xa[m][e], gv[r][e]);
              dxa[m][e] = fmaf(d, wv[r][e], dxa[m][e]);
            }
          }
        }
        if (dW) {
          if (vec4) *reinterpret_cast<float4*>(dW + wo) = make_float4(gv[r][0], gv[r][1], gv[r][2], gv[r][3]);
          else dW[wo] = gv[r][0];
        }
      }
    }
  }
  if (dx) {
#pragma unroll
    for (int m = 0; m < MM; ++m) {
      if (m < M) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (e < kw) atomicAdd(dx + (long long)m * dx_ld + k + e, dxa[m][e]);
      }
    }
  }
}

// =================================================================================================== column sum (bias grads)
// db[n] += sum_rows bf16 dy[row, n].  256 threads = vecs 16-byte column vectors x rpar row-lanes; shared-memory combine.
__global__ void __launch_bounds__(1024) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ dy, long long ld, long long rows,
                                                          int N, float* __restrict__ db, int rows_per_cta) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float s_cs[];  // [min(vecs,256) * 8]
  const int vecs_total = N >> 3;
  const int v0 = blockIdx.y * 256;
  const int vecs = min(256, vecs_total - v0);
  const int rpar = blockDim.x / vecs;      // row lanes: few, large CTAs keep the number of same-address global atomics per channel low
  const int vi = threadIdx.x % vecs, rsub = threadIdx.x / vecs;
  const bool active = rsub < rpar;
  for (int c = threadIdx.x; c < vecs * 8; c += blockDim.x) s_cs[c] = 0.f;
  __syncthreads();
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = min(r0 + rows_per_cta, rows);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (active) {
    const __nv_bfloat16* base = dy + (long long)(v0 + vi) * 8;
    long long r = r0 + rsub;
    for (; r + 3 * rpar < r1; r += 4 * rpar) {
      V8 a0 = ld_bf16x8(base + r * ld), a1 = ld_bf16x8(base + (r + rpar) * ld);
      V8 a2 = ld_bf16x8(base + (r + 2 * rpar) * ld), a3 = ld_bf16x8(base + (r + 3 * rpar) * ld);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += (a0.v[j] + a1.v[j]) + (a2.v[j] + a3.v[j]);
    }
    for (; r < r1; r += rpar) {
      V8 a0 = ld_bf16x8(base + r * ld);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += a0.v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&s_cs[vi * 8 + j], acc[j]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < vecs * 8; c += blockDim.x) atomicAdd(db + (long long)v0 * 8 + c, s_cs[c]);
}

// out[n] += sum_rows dy[row, n] * (y[row, n] - bias[n])   (DoRA magnitude gradient from activations: d mag = this / mag)
__global__ void __launch_bounds__(256) coldot_bf16_kernel(const __nv_bfloat16* __restrict__ dy, long long dy_ld,
                                                          const __nv_bfloat16* __restrict__ y, long long y_ld, long long rows, int N,
                                                          const float* __restrict__ bias, float* __restrict__ out, int rows_per_cta) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float s_cs[];  // [min(vecs,256) * 8]
  const int vecs_total = N >> 3;
  const int v0 = blockIdx.y * 256;
  const int vecs = min(256, vecs_total - v0);
  const int rpar = blockDim.x / vecs;
  const int vi = threadIdx.x % vecs, rsub = threadIdx.x / vecs;
  const bool active = rsub < rpar;
  for (int c = threadIdx.x; c < vecs * 8; c += blockDim.x) s_cs[c] = 0.f;
  __syncthreads();
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = min(r0 + rows_per_cta, rows);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (active) {
    const long long c0 = (long long)(v0 + vi) * 8;
    V8 bz;
#pragma unroll
    for (int j = 0; j < 8; ++j) bz.v[j] = bias ? bias[c0 + j] : 0.f;
    long long r = r0 + rsub;
    for (; r + rpar < r1; r += 2 * rpar) {
      V8 d0 = ld_bf16x8(dy + r * dy_ld + c0), y0 = ld_bf16x8(y + r * y_ld + c0);
      V8 d1 = ld_bf16x8(dy + (r + rpar) * dy_ld + c0), y1 = ld_bf16x8(y + (r + rpar) * y_ld + c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += d0.v[j] * (y0.v[j] - bz.v[j]) + d1.v[j] * (y1.v[j] - bz.v[j]);
    }
    for (; r < r1; r += rpar) {
      V8 d0 = ld_bf16x8(dy + r * dy_ld + c0), y0 = ld_bf16x8(y + r * y_ld + c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += d0.v[j] * (y0.v[j] - bz.v[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&s_cs[vi * 8 + j], acc[j]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < vecs * 8; c += blockDim.x) atomicAdd(out + (long long)v0 * 8 + c, s_cs[c]);
}

// =================================================================================================== layout / elementwise
// (B, C, N) fp32 channel-first -> (B, Lp, Cp) bf16 channels-last; out = bf16(ca[b]*x + cb[b]*noise) for l < N, pad_value for
// N <= l < Lp (channels >= C are zero).
__global__ void pack_input_kernel(const float* __restrict__ x, const float* __restrict__ noise, const float* __restrict__ ca,
                                  const float* __restrict__ cb, int B, int C, int N, __nv_bfloat16* __restrict__ out, int Lp,
                                  int Cp, float pad_value) {
  pdl_launch_dependents();
  pdl_wait();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cv = Cp >> 3;
  long long total = (long long)B * cv * Lp;
  if (idx >= total) return;
  const int l = (int)(idx % Lp);
  long long r = idx / Lp;
  const int v = (int)(r % cv), b = (int)(r / cv);
  const float a0 = ca ? ca[b] : 1.0f, b0 = cb ? cb[b] : 0.0f;
  V8 o;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = v * 8 + j;
    float val = 0.f;
    if (c < C) {
      if (l < N) {
        val = a0 * x[((long long)b * C + c) * N + l];
        if (noise) val += b0 * noise[((long long)b * C + c) * N + l];
      } else {
        val = pad_value;
      }
    }
    o.v[j] = val;
  }
  st_bf16x8(out + ((long long)b * Lp + l) * Cp + v * 8, o);
}

// (B, Lp, ld) bf16 channels-last -> (B, C, N) fp32 channel-first (first C channels, first N rows)
__global__ void unpack_output_kernel(const __nv_bfloat16* __restrict__ y, long long ld, long long bs, int B, int C, int N,
                                     float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * C * N;
  if (idx >= total) return;
  const int l = (int)(idx % N);
  long long r = idx / N;
  const int c = (int)(r % C), b = (int)(r / C);
  out[idx] = __bfloat162float(y[b * bs + (long long)l * ld + c]);
}

__global__ void upsample2x_fwd_kernel(const __nv_bfloat16* __restrict__ x, long long x_ld, long long x_bs, int B, int L, int C,
                                      __nv_bfloat16* __restrict__ out, long long o_ld, long long o_bs) {
  pdl_launch_dependents();
  pdl_wait();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cv = C >> 3;
  long long total = (long long)B * L * cv;
  if (idx >= total) return;
  const int v = (int)(idx % cv);
  long long r = idx / cv;
  const int l = (int)(r % L), b = (int)(r / L);
  uint4 u = *reinterpret_cast<const uint4*>(x + b * x_bs + (long long)l * x_ld + v * 8);
  *reinterpret_cast<uint4*>(out + b * o_bs + (long long)(2 * l) * o_ld + v * 8) = u;
  *reinterpret_cast<uint4*>(out + b * o_bs + (long long)(2 * l + 1) * o_ld + v * 8) = u;
}
// dx[b,l,:] = d[b,2l,:] + d[b,2l+1,:]   (fp32 in, fp32 and/or bf16 out)
__global__ void upsample2x_bwd_kernel(const float* __restrict__ d, long long d_ld, long long d_bs, int B, int L, int C,
                                      float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, long long o_ld,
                                      long long o_bs) {
  pdl_launch_dependents();
  pdl_wait();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cv = C >> 3;
  long long total = (long long)B * L * cv;
  if (idx >= total) return;
  const int v = (int)(idx % cv);
  long long r = idx / cv;
  const int l = (int)(r % L), b = (int)(r / L);
  V8 a = ld_f32x8(d + b * d_bs + (long long)(2 * l) * d_ld + v * 8);
  V8 c = ld_f32x8(d + b * d_bs + (long long)(2 * l + 1) * d_ld + v * 8), o;
#pragma unroll
  for (int j = 0; j < 8; ++j) o.v[j] = a.v[j] + c.v[j];
  if (out_f32) st_f32x8(out_f32 + b * o_bs + (long long)l * o_ld + v * 8, o);
  if (out_bf16) st_bf16x8(out_bf16 + b * o_bs + (long long)l * o_ld + v * 8, o);
}

// fp32 <-> bf16 strided copies / adds on (B, L, C) views
__global__ void cast_copy_kernel(const float* __restrict__ src32, const __nv_bfloat16* __restrict__ src16, long long s_ld,
                                 long long s_bs, int B, int L, int C, float* __restrict__ dst32, __nv_bfloat16* __restrict__ dst16,
                                 long long d_ld, long long d_bs, int accumulate) {
  pdl_launch_dependents();
  pdl_wait();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cv = C >> 3;
  long long total = (long long)B * L * cv;
  if (idx >= total) return;
  const int v = (int)(idx % cv);
  long long r = idx / cv;
  const int l = (int)(r % L), b = (int)(r / L);
  V8 a = src32 ? ld_f32x8(src32 + b * s_bs + (long long)l * s_ld + v * 8) : ld_bf16x8(src16 + b * s_bs + (long long)l * s_ld + v * 8);
  if (dst32) {
    float* p = dst32 + b * d_bs + (long long)l * d_ld + v * 8;
    if (accumulate) {
      V8 o = ld_f32x8(p);
#pragma unroll
      for (int j = 0; j < 8; ++j) a.v[j] += o.v[j];
    }
    st_f32x8(p, a);
  }
  if (dst16) st_bf16x8(dst16 + b * d_bs + (long long)l * d_ld + v * 8, a);
}

// sinusoidal embedding (unet.py:26-39): out[b, j] = sin(t_b f_j), out[b, half + j] = cos(t_b f_j), f_j = exp(-j ln(theta)/(half-1))
__global__ void time_embed_kernel(const float* __restrict__ t, int B, int dim, float theta, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2;
  if (idx >= B * half) return;
  const int j = idx % half, b = idx / half;
  const float step = logf(theta) / (float)(half - 1);
  const float f = expf((float)j * -step);
  const float ang = t[b] * f;
  out[(long long)b * dim + j] = sinf(ang);
  out[(long long)b * dim + half + j] = cosf(ang);
}

// elementwise SiLU on small fp32 tensors: fwd y = silu(x); bwd dx = dy * silu'(x)
__global__ void silu_small_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ out, long long n) {
  pdl_launch_dependents();
  pdl_wait();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = dy ? dy[i] * dsilu_acc(x[i]) : silu_acc(x[i]);
}

// masked MSE (diffusion.py:101-111): target = ta*x + tb*noise on (B,6,N); pred is bf16 channels-last (B, Lp, ld).
// accum[0] += sum mask*(pred-target)^2 ; accum[1] += sum mask
__global__ void __launch_bounds__(256) mse_fwd_kernel(const __nv_bfloat16* __restrict__ pred, long long ld, long long bs,
                                                      const float* __restrict__ x, const float* __restrict__ noise, float ta,
                                                      float tb, const long long* __restrict__ orig_len, int B, int C, int N,
                                                      float* __restrict__ accum) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sm[32];
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * C * N;
  float se = 0.f, cnt = 0.f;
  if (idx < total) {
    const int l = (int)(idx % N);
    long long r = idx / N;
    const int c = (int)(r % C), b = (int)(r / C);
    const bool on = orig_len ? (l < orig_len[b]) : true;
    if (on) {
      float tgt = tb * noise[idx] + (ta != 0.f ? ta * x[idx] : 0.f);
      float d = __bfloat162float(pred[b * bs + (long long)l * ld + c]) - tgt;
      se = d * d;
      cnt = 1.f;
    }
  }
  se = block_sum(se, sm);
  cnt = block_sum(cnt, sm);
  if (threadIdx.x == 0) {
    atomicAdd(accum, se);
    atomicAdd(accum + 1, cnt);
  }
}
__global__ void mse_finish_kernel(const float* __restrict__ accum, float* __restrict__ loss) {
  pdl_launch_dependents();
  pdl_wait(); loss[0] = accum[0] / accum[1]; }
// dpred (B, Lp, Cp) bf16 channels-last = gscale * 2 * mask * (pred - target) / count ; zero elsewhere
__global__ void mse_bwd_kernel(const __nv_bfloat16* __restrict__ pred, long long ld, long long bs, const float* __restrict__ x,
                               const float* __restrict__ noise, float ta, float tb, const long long* __restrict__ orig_len,
                               int B, int C, int N, int Lp, int Cp, const float* __restrict__ accum,
                               const float* __restrict__ gscale, __nv_bfloat16* __restrict__ dpred) {
  pdl_launch_dependents();
  pdl_wait();
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * Lp * Cp;
  if (idx >= total) return;
  const int c = (int)(idx % Cp);
  long long r = idx / Cp;
  const int l = (int)(r % Lp), b = (int)(r / Lp);
  float g = 0.f;
  if (c < C && l < N && (!orig_len || l < orig_len[b])) {
    const long long si = ((long long)b * C + c) * N + l;
    float tgt = tb * noise[si] + (ta != 0.f ? ta * x[si] : 0.f);
    float d = __bfloat162float(pred[b * bs + (long long)l * ld + c]) - tgt;
    g = 2.0f * d / accum[1] * (gscale ? gscale[0] : 1.0f);
  }
  dpred[idx] = __float2bfloat16_rn(g);
}

// sampler updates on (B,6,N) fp32 state with bf16 channels-last predictions (B, Lp, ld):
//   eps = null + (cond - null)*s  (bf16 arithmetic as unet.py:465 under autocast);  null == nullptr -> eps = cond
//   mode 0 (DDIM, diffusers 0.29.2 step with eta=0, clip_sample): x0 = clamp((x - c_eps*eps)/c_div, -1, 1); x = c_x0*x0 + c_dir*eps
//   mode 1 (axpy, midpoint half/full steps): out = y + c_eps * eps
// also emits the bf16 channels-last packed copy of the new state for the next denoiser call.
__global__ void sampler_update_kernel(const float* __restrict__ xin, const __nv_bfloat16* __restrict__ cond,
                                      const __nv_bfloat16* __restrict__ null_, long long ld, long long bs, float s, int mode,
                                      float c_eps, float c_div, float c_x0, float c_dir, int B, int C, int N,
                                      float* __restrict__ xout, __nv_bfloat16* __restrict__ packed, int Lp, int Cp,
                                      float pad_value, const float* __restrict__ coef_dev) {
  pdl_launch_dependents();
  pdl_wait();
  if (coef_dev) {   // per-step coefficients in device memory: the launch (and a CUDA graph holding it) is step-independent
    c_eps = coef_dev[0]; c_div = coef_dev[1]; c_x0 = coef_dev[2]; c_dir = coef_dev[3];
  }
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * Lp * Cp;
  if (idx >= total) return;
  const int c = (int)(idx % Cp);
  long long r = idx / Cp;
  const int l = (int)(r % Lp), b = (int)(r / Lp);
  float outv = 0.f;
  if (c < C) {
    if (l < N) {
      const long long si = ((long long)b * C + c) * N + l;
      float e = __bfloat162float(cond[b * bs + (long long)l * ld + c]);
      if (null_) {
        float nl = __bfloat162float(null_[b * bs + (long long)l * ld + c]);
        e = bf16_round(nl + bf16_round(bf16_round(e - nl) * s));
      }
      const float xv = xin[si];
      if (mode == 0) {
        float x0 = (xv - bf16_round(c_eps * e)) / c_div;
        x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
        outv = c_x0 * x0 + bf16_round(c_dir * e);
      } else {
        outv = xv + bf16_round(c_eps * e);
      }
      xout[si] = outv;
    } else {
      outv = pad_value;
    }
  }
  if (packed) packed[idx] = __float2bfloat16_rn(outv);
}

// weight repacking: torch Conv1d weight (Cout, Cin, k) fp32 -> [k][Cout][Cin_pad] bf16 (zero padded), and back-accumulation
// CTA = (output channel co, 256-wide ci chunk): the (ci, t) slab of one co is contiguous in the torch layout and each tap row
// is contiguous in the packed layout, so both sides are coalesced through a shared-memory transpose.
constexpr int kPackChunk = 256;
__global__ void __launch_bounds__(256) pack_conv_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int k,
                                                               __nv_bfloat16* __restrict__ out, int Cin_pad, int tap_offset,
                                                               int taps_total) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float s_w[];  // [kPackChunk * k]
  const int co = blockIdx.x;
  const int ci0 = blockIdx.y * kPackChunk;
  const int nci = min(kPackChunk, Cin_pad - ci0);
  const int nreal = max(0, min(kPackChunk, Cin - ci0));
  const float* src = w + ((long long)co * Cin + ci0) * k;
  for (int i = threadIdx.x; i < nreal * k; i += blockDim.x) s_w[i] = src[i];
  __syncthreads();
  for (int i = threadIdx.x; i < nci * k; i += blockDim.x) {
    const int t = i / nci, ci = i - t * nci;
    const float v = ci < nreal ? s_w[ci * k + t] : 0.f;
    out[((long long)(t + tap_offset) * Cout + co) * Cin_pad + ci0 + ci] = __float2bfloat16_rn(v);
  }
}
__global__ void __launch_bounds__(256) unpack_conv_wgrad_kernel(float* __restrict__ packed, int Cout, int Cin, int k,
                                                                int Cin_pad, int tap_offset, float* __restrict__ dw,
                                                                int accumulate, int rezero) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float s_w[];
  const int co = blockIdx.x;
  const int ci0 = blockIdx.y * kPackChunk;
  const int nreal = max(0, min(kPackChunk, Cin - ci0));
  for (int i = threadIdx.x; i < nreal * k; i += blockDim.x) {
    const int t = i / nreal, ci = i - t * nreal;
    float* src = packed + ((long long)(t + tap_offset) * Cout + co) * Cin_pad + ci0 + ci;
    s_w[ci * k + t] = *src;
    if (rezero) *src = 0.f;   // persistent split-K scratch: left clean for the next step (no per-step fill launch)
  }
  __syncthreads();
  float* dst = dw + ((long long)co * Cin + ci0) * k;
  for (int i = threadIdx.x; i < nreal * k; i += blockDim.x) dst[i] = accumulate ? dst[i] + s_w[i] : s_w[i];
}
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  pdl_launch_dependents();
  pdl_wait();
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    float4 v = *reinterpret_cast<const float4*>(src + i);
    uint2 u = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    *reinterpret_cast<uint2*>(dst + i) = u;
  } else {
    for (; i < n; ++i) dst[i] = __float2bfloat16_rn(src[i]);
  }
}

static unsigned blocks_for(long long total, int threads) { return (unsigned)((total + threads - 1) / threads); }

}  // namespace ofx

using namespace ofx;
#define STREAM reinterpret_cast<cudaStream_t>(stream)
#define DONE()                        \
  OF_CHECK_CUDA(cudaGetLastError()); \
  count_launch();                    \
  return OF_OK;

extern "C" int of_layernorm_fwd(const float* x, long long x_ld, int rows, int C, const float* gamma, const float* beta, float eps,
                                float* out_f32, void* out_bf16, long long out_ld, float* mean_rstd, void* stream) {
  OF_REQUIRE(x && gamma && beta && (out_f32 || out_bf16), "of_layernorm_fwd: null pointer");
  OF_REQUIRE(C % 8 == 0 && C <= 2048 && x_ld % 4 == 0 && out_ld % 8 == 0, "of_layernorm_fwd: unsupported C=%d", C);
  dim3 grid((rows + 7) / 8);
  const int nv = (C / 8 + 31) / 32;
  __nv_bfloat16* o16 = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  if (nv <= 1) OF_CHECK_CUDA(launch_pdl(layernorm_fwd_kernel<1>, dim3(grid), dim3(256), 0, STREAM, x, x_ld, rows, C, gamma, beta, eps, out_f32, o16, out_ld, mean_rstd));
  else if (nv <= 2) OF_CHECK_CUDA(launch_pdl(layernorm_fwd_kernel<2>, dim3(grid), dim3(256), 0, STREAM, x, x_ld, rows, C, gamma, beta, eps, out_f32, o16, out_ld, mean_rstd));
  else if (nv <= 4) OF_CHECK_CUDA(launch_pdl(layernorm_fwd_kernel<4>, dim3(grid), dim3(256), 0, STREAM, x, x_ld, rows, C, gamma, beta, eps, out_f32, o16, out_ld, mean_rstd));
  else OF_CHECK_CUDA(launch_pdl(layernorm_fwd_kernel<8>, dim3(grid), dim3(256), 0, STREAM, x, x_ld, rows, C, gamma, beta, eps, out_f32, o16, out_ld, mean_rstd));
  DONE()
}

extern "C" int of_layernorm_bwd(const float* dy, long long dy_ld, const float* x, long long x_ld, int rows, int C,
                                const float* gamma, const float* mean_rstd, float* dx_f32, void* dx_bf16, long long dx_ld,
                                float* dgamma, float* dbeta, void* stream) {
  OF_REQUIRE(dy && x && gamma && mean_rstd && dgamma && dbeta && (dx_f32 || dx_bf16), "of_layernorm_bwd: null pointer");
  OF_REQUIRE(C % 8 == 0 && C <= 2048, "of_layernorm_bwd: unsupported C=%d", C);
  int rpw = (rows + 8 * 296 - 1) / (8 * 296);   // ~2 CTAs per SM, each warp walks `rpw` consecutive rows
  if (rpw < 1) rpw = 1;
  dim3 grid((rows + 8 * rpw - 1) / (8 * rpw));
  const size_t red_smem = 2 * (size_t)C * sizeof(float);
  const int nv = (C / 8 + 31) / 32;
  __nv_bfloat16* o16 = reinterpret_cast<__nv_bfloat16*>(dx_bf16);
  if (nv <= 1) OF_CHECK_CUDA(launch_pdl(layernorm_bwd_kernel<1>, dim3(grid), dim3(256), red_smem, STREAM, dy, dy_ld, x, x_ld, rows, C, gamma, mean_rstd, dx_f32, o16, dx_ld, dgamma, dbeta, rpw));
  else if (nv <= 2) OF_CHECK_CUDA(launch_pdl(layernorm_bwd_kernel<2>, dim3(grid), dim3(256), red_smem, STREAM, dy, dy_ld, x, x_ld, rows, C, gamma, mean_rstd, dx_f32, o16, dx_ld, dgamma, dbeta, rpw));
  else if (nv <= 4) OF_CHECK_CUDA(launch_pdl(layernorm_bwd_kernel<4>, dim3(grid), dim3(256), red_smem, STREAM, dy, dy_ld, x, x_ld, rows, C, gamma, mean_rstd, dx_f32, o16, dx_ld, dgamma, dbeta, rpw));
  else OF_CHECK_CUDA(launch_pdl(layernorm_bwd_kernel<8>, dim3(grid), dim3(256), red_smem, STREAM, dy, dy_ld, x, x_ld, rows, C, gamma, mean_rstd, dx_f32, o16, dx_ld, dgamma, dbeta, rpw));
  DONE()
}

extern "C" int of_rope_fwd(void* qkv, long long ld, long long bs, int B, int L, int H, int KVH, int D, const void* cos_tab,
                           const void* sin_tab, int table_f32, void* stream) {
  OF_REQUIRE(qkv && cos_tab && sin_tab, "of_rope_fwd: null pointer");
  OF_REQUIRE(D % 16 == 0 && ld % 8 == 0, "of_rope_fwd: D=%d must be a multiple of 16", D);
  long long total = (long long)B * L * (H + KVH) * (D / 16);
  if (table_f32)
    OF_CHECK_CUDA(launch_pdl(rope_fwd_kernel<float>, dim3(blocks_for(total, 256)), dim3(256), 0, STREAM, reinterpret_cast<__nv_bfloat16*>(qkv), ld, bs, B, L, H + KVH, D,
                                                                       reinterpret_cast<const float*>(cos_tab),
                                                                       reinterpret_cast<const float*>(sin_tab)));
  else
    OF_CHECK_CUDA(launch_pdl(rope_fwd_kernel<__nv_bfloat16>, dim3(blocks_for(total, 256)), dim3(256), 0, STREAM, 
        reinterpret_cast<__nv_bfloat16*>(qkv), ld, bs, B, L, H + KVH, D, reinterpret_cast<const __nv_bfloat16*>(cos_tab),
        reinterpret_cast<const __nv_bfloat16*>(sin_tab)));
  DONE()
}

extern "C" int of_rope_bwd(const float* dq, long long dq_ld, long long dq_bs, const float* dk, const float* dv, long long dkv_ld,
                           long long dkv_bs, void* dqkv_bf16, long long out_ld, long long out_bs, int B, int L, int H, int KVH,
                           int D, const void* cos_tab, const void* sin_tab, int table_f32, void* stream) {
  OF_REQUIRE(dq && dk && dv && dqkv_bf16 && cos_tab && sin_tab, "of_rope_bwd: null pointer");
  OF_REQUIRE(D % 16 == 0, "of_rope_bwd: D=%d must be a multiple of 16", D);
  long long total = (long long)B * L * (H + 2 * KVH) * (D / 16);
  if (table_f32)
    OF_CHECK_CUDA(launch_pdl(rope_bwd_kernel<float>, dim3(blocks_for(total, 256)), dim3(256), 0, STREAM, dq, dq_ld, dq_bs, dk, dv, dkv_ld, dkv_bs,
                                                                       reinterpret_cast<__nv_bfloat16*>(dqkv_bf16), out_ld, out_bs, B, L,
                                                                       H, KVH, D, reinterpret_cast<const float*>(cos_tab),
                                                                       reinterpret_cast<const float*>(sin_tab)));
  else
    OF_CHECK_CUDA(launch_pdl(rope_bwd_kernel<__nv_bfloat16>, dim3(blocks_for(total, 256)), dim3(256), 0, STREAM, 
        dq, dq_ld, dq_bs, dk, dv, dkv_ld, dkv_bs, reinterpret_cast<__nv_bfloat16*>(dqkv_bf16), out_ld, out_bs, B, L, H, KVH, D,
        reinterpret_cast<const __nv_bfloat16*>(cos_tab), reinterpret_cast<const __nv_bfloat16*>(sin_tab)));
  DONE()
}

extern "C" int of_linear_small_fwd(const float* x, long long x_ld, int M, int N, int K, const float* W, long long w_ld,
                                   const float* bias, int act, int round_bf16, float* y, long long y_ld, float* ypre,
                                   void* stream) {
  OF_REQUIRE(x && W && y, "of_linear_small_fwd: null pointer");
  OF_REQUIRE(M >= 1 && M <= kMaxM, "of_linear_small_fwd: M=%d out of range (1..%d)", M, kMaxM);
  if (M <= 4) OF_CHECK_CUDA(launch_pdl(linear_small_fwd_kernel<4>, dim3((N + 7) / 8), dim3(256), 0, STREAM, x, x_ld, M, N, K, W, w_ld, bias, act, round_bf16, y, y_ld, ypre));
  else if (M <= 8) OF_CHECK_CUDA(launch_pdl(linear_small_fwd_kernel<8>, dim3((N + 7) / 8), dim3(256), 0, STREAM, x, x_ld, M, N, K, W, w_ld, bias, act, round_bf16, y, y_ld, ypre));
  else OF_CHECK_CUDA(launch_pdl(linear_small_fwd_kernel<16>, dim3((N + 7) / 8), dim3(256), 0, STREAM, x, x_ld, M, N, K, W, w_ld, bias, act, round_bf16, y, y_ld, ypre));
  DONE()
}

extern "C" int of_linear_small_bwd(const float* dy, long long dy_ld, const float* ypre, int act, const float* x, long long x_ld,
                                   int M, int N, int K, const float* W, long long w_ld, int round_bf16, float* dW, float* dbias,
                                   float* dx, long long dx_ld, int accumulate, void* stream) {
  OF_REQUIRE(dy && x && W, "of_linear_small_bwd: null pointer");
  OF_REQUIRE(M >= 1 && M <= kMaxM, "of_linear_small_bwd: M=%d out of range", M);
  OF_REQUIRE(act == 0 || ypre, "of_linear_small_bwd: ypre required for activation backward");
  const bool vec4 = ((K & 3) == 0) && ((w_ld & 3) == 0) && ((x_ld & 3) == 0) && (!dx || (dx_ld & 3) == 0);
  const int kw = vec4 ? 4 : 1;
  const int kcols = (K + kw - 1) / kw;
  int kthreads = 16;                       // power of two: at most 16 row-lanes (= kNChunk rows side by side)
  while (kthreads < kcols && kthreads < 256) kthreads <<= 1;
  dim3 grid((kcols + kthreads - 1) / kthreads, (N + kNChunk - 1) / kNChunk);
  const long long dxl = dx ? dx_ld : 4;
  if (M <= 4) OF_CHECK_CUDA(launch_pdl(linear_small_bwd_kernel<4>, dim3(grid), dim3(256), 0, STREAM, dy, dy_ld, ypre, act, x, x_ld, M, N, K, W, w_ld, round_bf16, dW, dbias, dx, dxl, kthreads, accumulate));
  else if (M <= 8) OF_CHECK_CUDA(launch_pdl(linear_small_bwd_kernel<8>, dim3(grid), dim3(256), 0, STREAM, dy, dy_ld, ypre, act, x, x_ld, M, N, K, W, w_ld, round_bf16, dW, dbias, dx, dxl, kthreads, accumulate));
  else OF_CHECK_CUDA(launch_pdl(linear_small_bwd_kernel<16>, dim3(grid), dim3(256), 0, STREAM, dy, dy_ld, ypre, act, x, x_ld, M, N, K, W, w_ld, round_bf16, dW, dbias, dx, dxl, kthreads, accumulate));
  DONE()
}

extern "C" int of_colsum_bf16(const void* dy, long long ld, long long rows, int N, float* db, void* stream) {
  OF_REQUIRE(dy && db && N % 8 == 0 && ld % 8 == 0, "of_colsum_bf16: bad args");
  // 256-thread CTAs, two per SM.  (One 1024-thread CTA per SM -- half as many same-address global atomics per channel -- was
  // measured slower: 1.90 vs 1.52 ms per training step over the 141 launches; OF_COLSUM_BIG=1 selects it.)
  static const bool big = [] {
    const char* e = getenv("OF_COLSUM_BIG");
    return e && e[0] == '1';
  }();
  const int ychunks = (N / 8 + 255) / 256;
  const int threads = big ? 1024 : 256;
  int ctas_x = (big ? 1 : 2) * device_sm_count() / ychunks;
  if (ctas_x < 1) ctas_x = 1;
  int rpc = (int)((rows + ctas_x - 1) / ctas_x);
  if (rpc < (big ? 64 : 16)) rpc = big ? 64 : 16;
  dim3 grid((unsigned)((rows + rpc - 1) / rpc), ychunks);
  OF_CHECK_CUDA(launch_pdl(colsum_bf16_kernel, dim3(grid), dim3(threads), 256 * 8 * sizeof(float), STREAM, reinterpret_cast<const __nv_bfloat16*>(dy), ld, rows, N, db, rpc));
  DONE()
}

extern "C" int of_coldot_bf16(const void* dy, long long dy_ld, const void* y, long long y_ld, long long rows, int N, const float* bias,
                              float* out, void* stream) {
  OF_REQUIRE(dy && y && out && N % 8 == 0 && dy_ld % 8 == 0 && y_ld % 8 == 0, "of_coldot_bf16: bad args");
  const int ychunks = (N / 8 + 255) / 256;
  int ctas_x = 2 * device_sm_count() / ychunks;
  if (ctas_x < 1) ctas_x = 1;
  int rpc = (int)((rows + ctas_x - 1) / ctas_x);
  if (rpc < 16) rpc = 16;
  dim3 grid((unsigned)((rows + rpc - 1) / rpc), ychunks);
  OF_CHECK_CUDA(launch_pdl(coldot_bf16_kernel, dim3(grid), dim3(256), 256 * 8 * sizeof(float), STREAM,
                           reinterpret_cast<const __nv_bfloat16*>(dy), dy_ld, reinterpret_cast<const __nv_bfloat16*>(y), y_ld, rows, N, bias,
                           out, rpc));
  DONE()
}

extern "C" int of_pack_input(const float* x, const float* noise, const float* ca, const float* cb, int B, int C, int N, void* out,
                             int Lp, int Cp, float pad_value, void* stream) {
  OF_REQUIRE(x && out && Cp % 8 == 0 && Cp >= C && Lp >= N, "of_pack_input: bad args");
  long long total = (long long)B * (Cp / 8) * Lp;
  OF_CHECK_CUDA(launch_pdl(pack_input_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, STREAM, x, noise, ca, cb, B, C, N, reinterpret_cast<__nv_bfloat16*>(out), Lp,
                                                                Cp, pad_value));
  DONE()
}

extern "C" int of_unpack_output(const void* y, long long ld, long long bs, int B, int C, int N, float* out, void* stream) {
  OF_REQUIRE(y && out, "of_unpack_output: null pointer");
  long long total = (long long)B * C * N;
  OF_CHECK_CUDA(launch_pdl(unpack_output_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, STREAM, reinterpret_cast<const __nv_bfloat16*>(y), ld, bs, B, C, N, out));
  DONE()
}

extern "C" int of_upsample2x_fwd(const void* x, long long x_ld, long long x_bs, int B, int L, int C, void* out, long long o_ld,
                                 long long o_bs, void* stream) {
  OF_REQUIRE(x && out && C % 8 == 0, "of_upsample2x_fwd: bad args");
  long long total = (long long)B * L * (C / 8);
  OF_CHECK_CUDA(launch_pdl(upsample2x_fwd_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, STREAM, reinterpret_cast<const __nv_bfloat16*>(x), x_ld, x_bs, B, L, C,
                                                                    reinterpret_cast<__nv_bfloat16*>(out), o_ld, o_bs));
  DONE()
}
extern "C" int of_upsample2x_bwd(const float* d, long long d_ld, long long d_bs, int B, int L, int C, float* out_f32, void* out_bf16,
                                 long long o_ld, long long o_bs, void* stream) {
  OF_REQUIRE(d && (out_f32 || out_bf16) && C % 8 == 0, "of_upsample2x_bwd: bad args");
  long long total = (long long)B * L * (C / 8);
  OF_CHECK_CUDA(launch_pdl(upsample2x_bwd_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, STREAM, d, d_ld, d_bs, B, L, C, out_f32,
                                                                    reinterpret_cast<__nv_bfloat16*>(out_bf16), o_ld, o_bs));
  DONE()
}

extern "C" int of_cast_copy(const float* src32, const void* src16, long long s_ld, long long s_bs, int B, int L, int C, float* dst32,
                            void* dst16, long long d_ld, long long d_bs, int accumulate, void* stream) {
  OF_REQUIRE((src32 || src16) && (dst32 || dst16) && C % 8 == 0, "of_cast_copy: bad args");
  long long total = (long long)B * L * (C / 8);
  OF_CHECK_CUDA(launch_pdl(cast_copy_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, STREAM, src32, reinterpret_cast<const __nv_bfloat16*>(src16), s_ld, s_bs, B, L, C,
                                                               dst32, reinterpret_cast<__nv_bfloat16*>(dst16), d_ld, d_bs, accumulate));
  DONE()
}

extern "C" int of_time_embed(const float* t, int B, int dim, float theta, float* out, void* stream) {
  OF_REQUIRE(t && out && dim >= 4 && dim % 2 == 0, "of_time_embed: bad args");
  OF_CHECK_CUDA(launch_pdl(time_embed_kernel, dim3(blocks_for((long long)B * (dim / 2), 256)), dim3(256), 0, STREAM, t, B, dim, theta, out));
  DONE()
}

extern "C" int of_silu_small(const float* x, const float* dy, float* out, long long n, void* stream) {
  OF_REQUIRE(x && out, "of_silu_small: null pointer");
  OF_CHECK_CUDA(launch_pdl(silu_small_kernel, dim3(blocks_for(n, 256)), dim3(256), 0, STREAM, x, dy, out, n));
  DONE()
}

extern "C" int of_mse_fwd(const void* pred, long long ld, long long bs, const float* x, const float* noise, float ta, float tb,
                          const long long* orig_len, int B, int C, int N, float* accum2, float* loss, void* stream) {
  OF_REQUIRE(pred && noise && accum2 && loss && (ta == 0.f || x), "of_mse_fwd: null pointer");
  OF_CHECK_CUDA(cudaMemsetAsync(accum2, 0, 2 * sizeof(float), STREAM));
  long long total = (long long)B * C * N;
  OF_CHECK_CUDA(launch_pdl(mse_fwd_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, STREAM, reinterpret_cast<const __nv_bfloat16*>(pred), ld, bs, x, noise, ta, tb,
                                                             orig_len, B, C, N, accum2));
  OF_CHECK_CUDA(launch_pdl(mse_finish_kernel, dim3(1), dim3(1), 0, STREAM, accum2, loss));
  count_launch();
  DONE()
}
extern "C" int of_mse_bwd(const void* pred, long long ld, long long bs, const float* x, const float* noise, float ta, float tb,
                          const long long* orig_len, int B, int C, int N, int Lp, int Cp, const float* accum2, const float* gscale,
                          void* dpred, void* stream) {
  OF_REQUIRE(pred && noise && accum2 && dpred, "of_mse_bwd: null pointer");
  long long total = (long long)B * Lp * Cp;
  OF_CHECK_CUDA(launch_pdl(mse_bwd_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, STREAM, reinterpret_cast<const __nv_bfloat16*>(pred), ld, bs, x, noise, ta, tb,
                                                             orig_len, B, C, N, Lp, Cp, accum2, gscale,
                                                             reinterpret_cast<__nv_bfloat16*>(dpred)));
  DONE()
}

extern "C" int of_sampler_update(const float* xin, const void* cond, const void* null_, long long ld, long long bs, float cond_scale,
                                 int mode, float c_eps, float c_div, float c_x0, float c_dir, int B, int C, int N, float* xout,
                                 void* packed, int Lp, int Cp, float pad_value, void* stream) {
  OF_REQUIRE(xin && cond && xout, "of_sampler_update: null pointer");
  long long total = (long long)B * Lp * Cp;
  OF_CHECK_CUDA(launch_pdl(sampler_update_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, STREAM, 
      xin, reinterpret_cast<const __nv_bfloat16*>(cond), reinterpret_cast<const __nv_bfloat16*>(null_), ld, bs, cond_scale, mode, c_eps,
      c_div, c_x0, c_dir, B, C, N, xout, reinterpret_cast<__nv_bfloat16*>(packed), Lp, Cp, pad_value, (const float*)nullptr));
  DONE()
}

extern "C" int of_sampler_update_dev(const float* xin, const void* cond, const void* null_, long long ld, long long bs, float cond_scale,
                                     int mode, const float* coef_dev, int B, int C, int N, float* xout, void* packed, int Lp, int Cp,
                                     float pad_value, void* stream) {
  OF_REQUIRE(xin && cond && xout && coef_dev, "of_sampler_update_dev: null pointer");
  long long total = (long long)B * Lp * Cp;
  OF_CHECK_CUDA(launch_pdl(sampler_update_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, STREAM,
      xin, reinterpret_cast<const __nv_bfloat16*>(cond), reinterpret_cast<const __nv_bfloat16*>(null_), ld, bs, cond_scale, mode, 0.f,
      1.f, 0.f, 0.f, B, C, N, xout, reinterpret_cast<__nv_bfloat16*>(packed), Lp, Cp, pad_value, coef_dev));
  DONE()
}

extern "C" int of_pack_conv_weight(const float* w, int Cout, int Cin, int k, void* out, int Cin_pad, int tap_offset, int taps_total,
                                   void* stream) {
  OF_REQUIRE(w && out && Cin_pad >= Cin && tap_offset + k <= taps_total, "of_pack_conv_weight: bad args");
  OF_REQUIRE(k <= 32, "of_pack_conv_weight: kernel size %d too large", k);
  dim3 grid(Cout, (Cin_pad + kPackChunk - 1) / kPackChunk);
  OF_CHECK_CUDA(launch_pdl(pack_conv_weight_kernel, dim3(grid), dim3(256), kPackChunk * k * sizeof(float), STREAM, w, Cout, Cin, k, reinterpret_cast<__nv_bfloat16*>(out),
                                                                                 Cin_pad, tap_offset, taps_total));
  DONE()
}
extern "C" int of_unpack_conv_wgrad(float* packed, int Cout, int Cin, int k, int Cin_pad, int tap_offset, float* dw,
                                    int accumulate, int rezero, void* stream) {
  OF_REQUIRE(packed && dw, "of_unpack_conv_wgrad: null pointer");
  OF_REQUIRE(k <= 32, "of_unpack_conv_wgrad: kernel size %d too large", k);
  dim3 grid(Cout, (Cin + kPackChunk - 1) / kPackChunk);
  OF_CHECK_CUDA(launch_pdl(unpack_conv_wgrad_kernel, dim3(grid), dim3(256), kPackChunk * k * sizeof(float), STREAM, packed, Cout, Cin, k, Cin_pad, tap_offset, dw,
                                                                                  accumulate, rezero));
  DONE()
}
extern "C" int of_cast_f32_bf16(const float* src, void* dst, long long n, void* stream) {
  OF_REQUIRE(src && dst, "of_cast_f32_bf16: null pointer");
  OF_CHECK_CUDA(launch_pdl(cast_f32_bf16_kernel, dim3(blocks_for((n + 3) / 4, 256)), dim3(256), 0, STREAM, src, reinterpret_cast<__nv_bfloat16*>(dst), n));
  DONE()
}
