// Bandwidth-bound kernels of the DiT / MMDiT backbones (reference osu_fusion/modules/dit.py, mmdit.py)
// that the UNet path does not already provide: adaLN-Zero gated residual, per-head q/k RMS norm, audio statistics pooling.
// Contracts and reference citations are in include/osufusion_b200.h next to each of_* declaration.
#include "host_common.h"
#include "rowops.cuh"

namespace ofx {

__device__ __forceinline__ float bf16r(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// out[b,l,:] = x[b,l,:] + gate[b,:] * y[b,l,:]   (16-byte vectors; x fp32 or bf16)
__global__ void __launch_bounds__(256) gate_residual_fwd_kernel(const float* __restrict__ x32, const __nv_bfloat16* __restrict__ x16,
                                                                long long x_ld, long long x_bs, const __nv_bfloat16* __restrict__ y16,
                                                                long long y_ld, long long y_bs, const float* __restrict__ gate,
                                                                long long gate_ld, int round_bf16, int B, int L, int C,
                                                                float* __restrict__ out32, long long o_ld, long long o_bs) {
  pdl_launch_dependents();
  pdl_wait();
  const int cv = C >> 3;
  const long long total = (long long)B * L * cv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(idx % cv);
    const long long r = idx / cv;
    const int l = (int)(r % L), b = (int)(r / L);
    const V8 xv = x32 ? ld_f32x8(x32 + b * x_bs + (long long)l * x_ld + v * 8) : ld_bf16x8(x16 + b * x_bs + (long long)l * x_ld + v * 8);
    const V8 yv = ld_bf16x8(y16 + b * y_bs + (long long)l * y_ld + v * 8);
    const V8 g = ld_f32x8(gate + b * gate_ld + v * 8);
    V8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float p = g.v[j] * yv.v[j];
      if (round_bf16) p = bf16r(p);
      o.v[j] = xv.v[j] + p;
    }
    st_f32x8(out32 + b * o_bs + (long long)l * o_ld + v * 8, o);
  }
}

// dx16 = bf16(dx);  dy16 = bf16(gate * dx)   (the product's incoming gradient is rounded to bf16 first when the forward product was)
__global__ void __launch_bounds__(256) gate_mul_bwd_kernel(const float* __restrict__ dx32, long long d_ld, long long d_bs,
                                                           const float* __restrict__ gate, long long gate_ld, int round_bf16, int B,
                                                           int L, int C, __nv_bfloat16* __restrict__ dx16,
                                                           __nv_bfloat16* __restrict__ dy16, long long o_ld, long long o_bs) {
  pdl_launch_dependents();
  pdl_wait();
  const int cv = C >> 3;
  const long long total = (long long)B * L * cv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(idx % cv);
    const long long r = idx / cv;
    const int l = (int)(r % L), b = (int)(r / L);
    const V8 d = ld_f32x8(dx32 + b * d_bs + (long long)l * d_ld + v * 8);
    const V8 g = ld_f32x8(gate + b * gate_ld + v * 8);
    V8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = g.v[j] * (round_bf16 ? bf16r(d.v[j]) : d.v[j]);
    const long long off = b * o_bs + (long long)l * o_ld + v * 8;
    st_bf16x8(dx16 + off, d);
    st_bf16x8(dy16 + off, o);
  }
}

// One thread per (row, head) of the fused [q | k | v] projection row: F.normalize(x, dim=-1) * gamma * sqrt(D) on the q and k heads
// with the reference's bf16 rounding points (norm, quotient, result); v heads are copied.  kMaxVec 16-byte vectors per head.
constexpr int kHeadVecs = 8;   // D <= 64

__global__ void __launch_bounds__(256) headnorm_fwd_kernel(const __nv_bfloat16* __restrict__ in, long long in_ld, long long in_bs, int B,
                                                           int L, int Hq, int Hk, int Hv, int D, const float* __restrict__ gq,
                                                           const float* __restrict__ gk, float scale,
                                                           __nv_bfloat16* __restrict__ out, long long o_ld, long long o_bs) {
  pdl_launch_dependents();
  pdl_wait();
  const int Ht = Hq + Hk + Hv, dv = D >> 3;
  const long long total = (long long)B * L * Ht;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int h = (int)(idx % Ht);
    const long long r = idx / Ht;
    const int l = (int)(r % L), b = (int)(r / L);
    const __nv_bfloat16* src = in + b * in_bs + (long long)l * in_ld + (long long)h * D;
    __nv_bfloat16* dst = out + b * o_bs + (long long)l * o_ld + (long long)h * D;
    V8 x[kHeadVecs];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kHeadVecs; ++i)
      if (i < dv) {
        x[i] = ld_bf16x8(src + i * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) ss += x[i].v[j] * x[i].v[j];
      }
    if (h >= Hq + Hk) {   // v: copy
#pragma unroll
      for (int i = 0; i < kHeadVecs; ++i)
        if (i < dv) st_bf16x8(dst + i * 8, x[i]);
      continue;
    }
    const float* g = h < Hq ? gq + (long long)h * D : gk + (long long)(h - Hq) * D;
    const float n = fmaxf(bf16r(sqrtf(ss)), 1e-12f);
#pragma unroll
    for (int i = 0; i < kHeadVecs; ++i)
      if (i < dv) {
        V8 o;
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] = (bf16r(x[i].v[j] / n) * g[i * 8 + j]) * scale;
        st_bf16x8(dst + i * 8, o);
      }
  }
}

// Backward of the above (roundings treated as identity): with xh = x/|x|, dxh = dy * gamma * scale,
//   dx = (dxh - xh <dxh, xh>) / |x|,   dgamma[h, d] += sum_rows dy * xh * scale.
// dq/dk/dv arrive as fp32 (attention backward accumulates them atomically); the result is the bf16 [dq | dk | dv] GEMM operand.
// gamma gradients: shared-memory accumulation per CTA (grid-stride rows), then one global atomic per element per CTA.
__global__ void __launch_bounds__(256) headnorm_bwd_kernel(const float* __restrict__ dq, long long dq_ld, long long dq_bs,
                                                           const float* __restrict__ dk, const float* __restrict__ dvp, long long dkv_ld,
                                                           long long dkv_bs, const __nv_bfloat16* __restrict__ in, long long in_ld,
                                                           long long in_bs, int B, int L, int Hq, int Hk, int Hv, int D,
                                                           const float* __restrict__ gq, const float* __restrict__ gk, float scale,
                                                           __nv_bfloat16* __restrict__ out, long long o_ld, long long o_bs,
                                                           float* __restrict__ dgq, float* __restrict__ dgk) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float s_dg[];   // [(Hq + Hk) * D]
  const int Ht = Hq + Hk + Hv, dv = D >> 3, ng = (Hq + Hk) * D;
  for (int i = threadIdx.x; i < ng; i += blockDim.x) s_dg[i] = 0.f;
  __syncthreads();
  const long long total = (long long)B * L * Ht;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int h = (int)(idx % Ht);
    const long long r = idx / Ht;
    const int l = (int)(r % L), b = (int)(r / L);
    __nv_bfloat16* dst = out + b * o_bs + (long long)l * o_ld + (long long)h * D;
    const float* dsrc;
    if (h < Hq) dsrc = dq + b * dq_bs + (long long)l * dq_ld + (long long)h * D;
    else if (h < Hq + Hk) dsrc = dk + b * dkv_bs + (long long)l * dkv_ld + (long long)(h - Hq) * D;
    else dsrc = dvp + b * dkv_bs + (long long)l * dkv_ld + (long long)(h - Hq - Hk) * D;
    V8 d[kHeadVecs];
#pragma unroll
    for (int i = 0; i < kHeadVecs; ++i)
      if (i < dv) d[i] = ld_f32x8(dsrc + i * 8);
    if (h >= Hq + Hk) {   // v: cast
#pragma unroll
      for (int i = 0; i < kHeadVecs; ++i)
        if (i < dv) st_bf16x8(dst + i * 8, d[i]);
      continue;
    }
    const __nv_bfloat16* src = in + b * in_bs + (long long)l * in_ld + (long long)h * D;
    const float* g = h < Hq ? gq + (long long)h * D : gk + (long long)(h - Hq) * D;
    V8 x[kHeadVecs];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kHeadVecs; ++i)
      if (i < dv) {
        x[i] = ld_bf16x8(src + i * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) ss += x[i].v[j] * x[i].v[j];
      }
    const float n = fmaxf(sqrtf(ss), 1e-12f), rn = 1.0f / n;
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < kHeadVecs; ++i)
      if (i < dv) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = x[i].v[j] * rn;
          const float dy = d[i].v[j];
          atomicAdd(&s_dg[h * D + i * 8 + j], dy * xh * scale);
          const float dxh = dy * g[i * 8 + j] * scale;
          dot += dxh * xh;
          x[i].v[j] = xh;
          d[i].v[j] = dxh;
        }
      }
#pragma unroll
    for (int i = 0; i < kHeadVecs; ++i)
      if (i < dv) {
        V8 o;
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] = (d[i].v[j] - x[i].v[j] * dot) * rn;
        st_bf16x8(dst + i * 8, o);
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ng; i += blockDim.x) {
    const float v = s_dg[i];
    if (v != 0.f) atomicAdd(i < Hq * D ? dgq + i : dgk + (i - Hq * D), v);
  }
}

// ---- variant 2 of the per-head RMS norm (head dims 8 / 16 / 32 / 64): thread = ONE 16-byte vector of one head, so a warp reads and
// writes whole 128-byte lines; the D/8 lanes of a head reduce |x|^2 and <dxh, xh> with shuffles; a thread keeps its channel column
// for every row it visits, so the gamma gradient accumulates in registers (variant 1 above: one thread per head, 128-byte strides
// between lanes and a 32-way bank conflict on its shared-memory gamma atomics).  Grid: (row chunks, column chunks of <= 256 vectors).
__global__ void __launch_bounds__(256) headnorm2_fwd_kernel(const __nv_bfloat16* __restrict__ in, long long in_ld, long long in_bs,
                                                            long long rows, int L, int Hq, int Hk, int Hv, int D,
                                                            const float* __restrict__ gq, const float* __restrict__ gk, float scale,
                                                            __nv_bfloat16* __restrict__ out, long long o_ld, long long o_bs, int cols,
                                                            int rows_per_cta) {
  pdl_launch_dependents();
  pdl_wait();
  const int gs = D >> 3;
  const int total_cols = (Hq + Hk + Hv) * gs;
  const int c0 = blockIdx.y * cols;
  const int ncols = min(cols, total_cols - c0);
  const int rpar = blockDim.x / ncols;
  const int c = threadIdx.x % ncols, rsub = threadIdx.x / ncols;
  const bool active = rsub < rpar;
  const int col = c0 + c;
  const int h = col / gs;
  const bool is_v = h >= Hq + Hk;
  V8 gv;
#pragma unroll
  for (int j = 0; j < 8; ++j) gv.v[j] = 1.f;
  if (!is_v) gv = ld_f32x8((h < Hq ? gq : gk - (long long)Hq * D) + (long long)col * 8);
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = min(r0 + (long long)rows_per_cta, rows);
  for (long long base = r0; base < r1; base += rpar) {
    const long long row = base + rsub;
    const bool valid = active && row < r1;
    const int b = (int)(row / L), l = (int)(row % L);
    V8 x;
#pragma unroll
    for (int j = 0; j < 8; ++j) x.v[j] = 0.f;
    if (valid) x = ld_bf16x8(in + b * in_bs + (long long)l * in_ld + (long long)col * 8);
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) ss += x.v[j] * x.v[j];
    for (int o = gs >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (valid) {
      if (!is_v) {
        const float n = fmaxf(bf16r(sqrtf(ss)), 1e-12f);
#pragma unroll
        for (int j = 0; j < 8; ++j) x.v[j] = (bf16r(x.v[j] / n) * gv.v[j]) * scale;
      }
      st_bf16x8(out + b * o_bs + (long long)l * o_ld + (long long)col * 8, x);
    }
  }
}

__global__ void __launch_bounds__(256) headnorm2_bwd_kernel(const float* __restrict__ dq, long long dq_ld, long long dq_bs,
                                                            const float* __restrict__ dk, const float* __restrict__ dvp, long long dkv_ld,
                                                            long long dkv_bs, const __nv_bfloat16* __restrict__ in, long long in_ld,
                                                            long long in_bs, long long rows, int L, int Hq, int Hk, int Hv, int D,
                                                            const float* __restrict__ gq, const float* __restrict__ gk, float scale,
                                                            __nv_bfloat16* __restrict__ out, long long o_ld, long long o_bs,
                                                            float* __restrict__ dgq, float* __restrict__ dgk, int cols, int rows_per_cta) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float s_dg[];   // [ncols * 8]
  const int gs = D >> 3;
  const int total_cols = (Hq + Hk + Hv) * gs;
  const int c0 = blockIdx.y * cols;
  const int ncols = min(cols, total_cols - c0);
  const int rpar = blockDim.x / ncols;
  const int c = threadIdx.x % ncols, rsub = threadIdx.x / ncols;
  const bool active = rsub < rpar;
  const int col = c0 + c;
  const int h = col / gs;
  const bool is_v = h >= Hq + Hk;
  for (int i = threadIdx.x; i < ncols * 8; i += blockDim.x) s_dg[i] = 0.f;
  __syncthreads();
  V8 gv;
#pragma unroll
  for (int j = 0; j < 8; ++j) gv.v[j] = 1.f;
  if (!is_v) gv = ld_f32x8((h < Hq ? gq : gk - (long long)Hq * D) + (long long)col * 8);
  // gradient source of this column: dq | dk | dv hold Hq*D | Hk*D | Hv*D channels
  const float* dsrc;
  long long s_ld, s_bs;
  if (h < Hq) { dsrc = dq + (long long)col * 8; s_ld = dq_ld; s_bs = dq_bs; }
  else if (!is_v) { dsrc = dk + ((long long)col * 8 - (long long)Hq * D); s_ld = dkv_ld; s_bs = dkv_bs; }
  else { dsrc = dvp + ((long long)col * 8 - (long long)(Hq + Hk) * D); s_ld = dkv_ld; s_bs = dkv_bs; }
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = min(r0 + (long long)rows_per_cta, rows);
  for (long long base = r0; base < r1; base += rpar) {
    const long long row = base + rsub;
    const bool valid = active && row < r1;
    const int b = (int)(row / L), l = (int)(row % L);
    V8 x, d;
#pragma unroll
    for (int j = 0; j < 8; ++j) x.v[j] = d.v[j] = 0.f;
    if (valid) {
      d = ld_f32x8(dsrc + b * s_bs + (long long)l * s_ld);
      if (!is_v) x = ld_bf16x8(in + b * in_bs + (long long)l * in_ld + (long long)col * 8);
    }
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) ss += x.v[j] * x.v[j];
    for (int o = gs >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rn = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    float dot = 0.f;
    V8 xh, dxh;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      xh.v[j] = x.v[j] * rn;
      dxh.v[j] = d.v[j] * gv.v[j] * scale;
      dot += dxh.v[j] * xh.v[j];
    }
    for (int o = gs >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (valid) {
      if (!is_v) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[j] += d.v[j] * xh.v[j] * scale;
          d.v[j] = (dxh.v[j] - xh.v[j] * dot) * rn;
        }
      }
      st_bf16x8(out + b * o_bs + (long long)l * o_ld + (long long)col * 8, d);
    }
  }
  if (active && !is_v) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&s_dg[c * 8 + j], acc[j]);
  }
  __syncthreads();
  const long long nq = (long long)Hq * D, nqk = (long long)(Hq + Hk) * D;
  for (int i = threadIdx.x; i < ncols * 8; i += blockDim.x) {
    const long long ch = (long long)c0 * 8 + i;      // channel index within [q | k | v]
    const float v = s_dg[i];
    if (ch < nqk && v != 0.f) atomicAdd(ch < nq ? dgq + ch : dgk + (ch - nq), v);
  }
}

// out[b, c] = mean_n a[b, c, n], out[b, C + c] = unbiased std_n a[b, c, n]; one CTA per (b, c) row, two passes (mean, then centred squares).
__global__ void __launch_bounds__(256) row_mean_std_kernel(const float* __restrict__ a, int C, int N, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sm[32];
  const int row = blockIdx.x, b = row / C, c = row % C;
  const float* p = a + (long long)row * N;
  float s = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s += p[i];
  const float mean = block_sum(s, sm) / N;
  float q = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float dlt = p[i] - mean;
    q += dlt * dlt;
  }
  q = block_sum(q, sm);
  if (threadIdx.x == 0) {
    out[(long long)b * 2 * C + c] = mean;
    out[(long long)b * 2 * C + C + c] = sqrtf(q / (float)(N - 1));
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Batched adaLN: LayerNorm without affine + per-sample modulation `xhat * (1 + scale_b) + shift_b` -> bf16 GEMM operand.
// Warp per row over all B*L rows; lane owns 16-byte vectors lane + 32*i (i < NV); the row stays in registers.
template <int NV>
__global__ void __launch_bounds__(256) adaln_fwd_kernel(const float* __restrict__ x, long long x_ld, long long x_bs, int B, int L, int C,
                                                        const float* __restrict__ scale1p, long long sc_ld,
                                                        const float* __restrict__ shift, long long sh_ld, float eps,
                                                        __nv_bfloat16* __restrict__ out, long long o_ld, long long o_bs,
                                                        float* __restrict__ mean_rstd) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;
  if (row >= (long long)B * L) return;
  const int b = (int)(row / L), l = (int)(row % L);
  const float* xp = x + b * x_bs + (long long)l * x_ld;
  const int vecs = C >> 3;
  V8 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < vecs) {
      v[i] = ld_f32x8(xp + vi * 8);
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
        const float d = v[i].v[j] - mean;
        q += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
  if (lane == 0) {
    mean_rstd[2 * row] = mean;
    mean_rstd[2 * row + 1] = rstd;
  }
  __nv_bfloat16* op = out + b * o_bs + (long long)l * o_ld;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < vecs) {
      const V8 g = ld_f32x8(scale1p + b * sc_ld + vi * 8), bt = ld_f32x8(shift + b * sh_ld + vi * 8);
      V8 o;
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] = (v[i].v[j] - mean) * rstd * g.v[j] + bt.v[j];
      st_bf16x8(op + vi * 8, o);
    }
  }
}

// Backward: grid (row chunks, B).  dx = rstd * (dxh - mean(dxh) - xh * mean(dxh * xh)) [+ dres], dxh = dy * (1 + scale_b);
// d scale_b[c] += sum_l dy * xh, d shift_b[c] += sum_l dy  (shared-memory reduction per CTA, one global atomic per channel per CTA).
template <int NV>
__global__ void __launch_bounds__(256) adaln_bwd_kernel(const float* __restrict__ dy, long long dy_ld, long long dy_bs,
                                                        const float* __restrict__ x, long long x_ld, long long x_bs, int L, int C,
                                                        const float* __restrict__ scale1p, long long sc_ld,
                                                        const float* __restrict__ mean_rstd, const float* __restrict__ dres,
                                                        long long r_ld, long long r_bs, float* __restrict__ dx, long long dx_ld,
                                                        long long dx_bs, float* __restrict__ dscale, long long ds_ld,
                                                        float* __restrict__ dshift, long long dsh_ld, int rows_per_warp) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float s_red[];  // [2][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int vecs = C >> 3;
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) s_red[c] = 0.f;
  __syncthreads();
  V8 g[NV], dg[NV], db[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + 32 * i;
    if (vi < vecs) g[i] = ld_f32x8(scale1p + b * sc_ld + vi * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) dg[i].v[j] = db[i].v[j] = 0.f;
  }
  const int row0 = (blockIdx.x * 8 + warp) * rows_per_warp;
  for (int r = 0; r < rows_per_warp; ++r) {
    const int l = row0 + r;
    if (l >= L) break;
    const long long grow = (long long)b * L + l;
    const float mean = mean_rstd[2 * grow], rstd = mean_rstd[2 * grow + 1];
    V8 xh[NV], dxh[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < vecs) {
        const V8 xv = ld_f32x8(x + b * x_bs + (long long)l * x_ld + vi * 8);
        const V8 dv = ld_f32x8(dy + b * dy_bs + (long long)l * dy_ld + vi * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i].v[j] = (xv.v[j] - mean) * rstd;
          dg[i].v[j] += dv.v[j] * xh[i].v[j];
          db[i].v[j] += dv.v[j];
          dxh[i].v[j] = dv.v[j] * g[i].v[j];
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
        if (dres) {
          const V8 rv = ld_f32x8(dres + b * r_bs + (long long)l * r_ld + vi * 8);
#pragma unroll
          for (int j = 0; j < 8; ++j) o.v[j] += rv.v[j];
        }
        st_f32x8(dx + b * dx_bs + (long long)l * dx_ld + vi * 8, o);
      }
    }
  }
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
    atomicAdd(dscale + b * ds_ld + c, s_red[c]);
    atomicAdd(dshift + b * dsh_ld + c, s_red[C + c]);
  }
}

// Backward of `gate_b * y` in one pass: dy16 = bf16(gate * d~), d gate[b, c] += sum_l d~ * y with d~ = bf16(d) when the forward product
// was a bf16 multiplication.  Grid (row chunks, B); thread = (16-byte channel vector, row lane); shared reduction, then global atomics.
__global__ void __launch_bounds__(256) gate_bwd_kernel(const float* __restrict__ d32, long long d_ld, long long d_bs,
                                                       const float* __restrict__ gate, long long gate_ld,
                                                       const __nv_bfloat16* __restrict__ y16, long long y_ld, long long y_bs,
                                                       int round_bf16, int L, int C, __nv_bfloat16* __restrict__ dy16, long long o_ld,
                                                       long long o_bs, float* __restrict__ dgate, long long dg_ld, int rows_per_cta) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float s_acc[];  // [C]
  const int b = blockIdx.y;
  const int vecs = C >> 3;
  const int rpar = blockDim.x / vecs;
  const int vi = threadIdx.x % vecs, rsub = threadIdx.x / vecs;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s_acc[c] = 0.f;
  __syncthreads();
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(r0 + rows_per_cta, L);
  if (rsub < rpar) {
    const V8 g = ld_f32x8(gate + b * gate_ld + vi * 8);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int l = r0 + rsub; l < r1; l += rpar) {
      const V8 d = ld_f32x8(d32 + b * d_bs + (long long)l * d_ld + vi * 8);
      const V8 y = ld_bf16x8(y16 + b * y_bs + (long long)l * y_ld + vi * 8);
      V8 o;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dr = round_bf16 ? bf16r(d.v[j]) : d.v[j];
        acc[j] += dr * y.v[j];
        o.v[j] = g.v[j] * dr;
      }
      st_bf16x8(dy16 + b * o_bs + (long long)l * o_ld + vi * 8, o);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&s_acc[vi * 8 + j], acc[j]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(dgate + b * dg_ld + c, s_acc[c]);
}

static unsigned grid_for(long long total, int threads) {
  long long need = (total + threads - 1) / threads;
  long long cap = (long long)device_sm_count() * 8;
  return (unsigned)(need < cap ? (need < 1 ? 1 : need) : cap);
}

}  // namespace ofx

using namespace ofx;
#define STREAM reinterpret_cast<cudaStream_t>(stream)
#define DONE()                        \
  OF_CHECK_CUDA(cudaGetLastError()); \
  count_launch();                    \
  return OF_OK;

extern "C" int of_gate_residual_fwd(const float* x32, const void* x16, long long x_ld, long long x_bs, const void* y16, long long y_ld,
                                    long long y_bs, const float* gate, long long gate_ld, int round_bf16, int B, int L, int C,
                                    float* out32, long long o_ld, long long o_bs, void* stream) {
  OF_REQUIRE((x32 || x16) && y16 && gate && out32, "of_gate_residual_fwd: null pointer");
  OF_REQUIRE(B >= 1 && L >= 1 && C >= 8 && C % 8 == 0, "of_gate_residual_fwd: C=%d must be a positive multiple of 8", C);
  OF_REQUIRE(x_ld % 8 == 0 && y_ld % 8 == 0 && o_ld % 4 == 0 && gate_ld % 4 == 0 && x_bs % 4 == 0 && y_bs % 8 == 0 && o_bs % 4 == 0,
             "of_gate_residual_fwd: strides must keep 16-byte alignment");
  OF_CHECK_CUDA(launch_pdl(gate_residual_fwd_kernel, dim3(grid_for((long long)B * L * (C / 8), 256)), dim3(256), 0, STREAM, x32,
                           reinterpret_cast<const __nv_bfloat16*>(x16), x_ld, x_bs, reinterpret_cast<const __nv_bfloat16*>(y16), y_ld, y_bs,
                           gate, gate_ld, round_bf16, B, L, C, out32, o_ld, o_bs));
  DONE()
}

extern "C" int of_gate_mul_bwd(const float* dx32, long long d_ld, long long d_bs, const float* gate, long long gate_ld, int round_bf16,
                               int B, int L, int C, void* dx16, void* dy16, long long o_ld, long long o_bs, void* stream) {
  OF_REQUIRE(dx32 && gate && dx16 && dy16, "of_gate_mul_bwd: null pointer");
  OF_REQUIRE(B >= 1 && L >= 1 && C >= 8 && C % 8 == 0, "of_gate_mul_bwd: C=%d must be a positive multiple of 8", C);
  OF_REQUIRE(d_ld % 4 == 0 && d_bs % 4 == 0 && gate_ld % 4 == 0 && o_ld % 8 == 0 && o_bs % 8 == 0,
             "of_gate_mul_bwd: strides must keep 16-byte alignment");
  OF_CHECK_CUDA(launch_pdl(gate_mul_bwd_kernel, dim3(grid_for((long long)B * L * (C / 8), 256)), dim3(256), 0, STREAM, dx32, d_ld, d_bs, gate,
                           gate_ld, round_bf16, B, L, C, reinterpret_cast<__nv_bfloat16*>(dx16), reinterpret_cast<__nv_bfloat16*>(dy16), o_ld,
                           o_bs));
  DONE()
}

static bool headnorm2_ok(int D) { return D == 8 || D == 16 || D == 32 || D == 64; }
// column chunk (<= 256 vectors, a multiple of the D/8 lanes of a head) and rows per CTA of the variant-2 kernels
static void headnorm2_shape(int Ht, int D, long long rows, int ctas_per_sm, int* cols, dim3* grid, int* rpc) {
  const int total_cols = Ht * (D / 8);
  *cols = total_cols < 256 ? total_cols : 256;
  const int ychunks = (total_cols + *cols - 1) / *cols;
  long long target = (long long)device_sm_count() * ctas_per_sm / ychunks;
  if (target < 1) target = 1;
  long long r = (rows + target - 1) / target;
  if (r < 16) r = 16;
  *rpc = (int)r;
  *grid = dim3((unsigned)((rows + r - 1) / r), (unsigned)ychunks);
}

extern "C" int of_headnorm_fwd(const void* in16, long long in_ld, long long in_bs, int B, int L, int Hq, int Hk, int Hv, int D,
                               const float* gamma_q, const float* gamma_k, float scale, void* out16, long long o_ld, long long o_bs,
                               int variant, void* stream) {
  OF_REQUIRE(in16 && out16 && gamma_q && gamma_k, "of_headnorm_fwd: null pointer");
  OF_REQUIRE(D >= 8 && D <= 8 * kHeadVecs && D % 8 == 0, "of_headnorm_fwd: head dim %d unsupported (8..64, multiple of 8)", D);
  OF_REQUIRE(B >= 1 && L >= 1 && Hq >= 1 && Hk >= 1 && Hv >= 0, "of_headnorm_fwd: bad sizes");
  OF_REQUIRE(in_ld % 8 == 0 && in_bs % 8 == 0 && o_ld % 8 == 0 && o_bs % 8 == 0, "of_headnorm_fwd: strides must be multiples of 8");
  OF_REQUIRE(variant >= 0 && variant <= 2 && (variant != 2 || headnorm2_ok(D)), "of_headnorm_fwd: variant %d unsupported for D=%d", variant, D);
  if (variant == 2 || (variant == 0 && headnorm2_ok(D))) {
    int cols, rpc;
    dim3 grid;
    headnorm2_shape(Hq + Hk + Hv, D, (long long)B * L, 4, &cols, &grid, &rpc);
    OF_CHECK_CUDA(launch_pdl(headnorm2_fwd_kernel, grid, dim3(256), 0, STREAM, reinterpret_cast<const __nv_bfloat16*>(in16), in_ld, in_bs,
                             (long long)B * L, L, Hq, Hk, Hv, D, gamma_q, gamma_k, scale, reinterpret_cast<__nv_bfloat16*>(out16), o_ld,
                             o_bs, cols, rpc));
    DONE()
  }
  OF_CHECK_CUDA(launch_pdl(headnorm_fwd_kernel, dim3(grid_for((long long)B * L * (Hq + Hk + Hv), 256)), dim3(256), 0, STREAM,
                           reinterpret_cast<const __nv_bfloat16*>(in16), in_ld, in_bs, B, L, Hq, Hk, Hv, D, gamma_q, gamma_k, scale,
                           reinterpret_cast<__nv_bfloat16*>(out16), o_ld, o_bs));
  DONE()
}

extern "C" int of_headnorm_bwd(const float* dq, long long dq_ld, long long dq_bs, const float* dk, const float* dv, long long dkv_ld,
                               long long dkv_bs, const void* in16, long long in_ld, long long in_bs, int B, int L, int Hq, int Hk, int Hv,
                               int D, const float* gamma_q, const float* gamma_k, float scale, void* dqkv16, long long o_ld,
                               long long o_bs, float* dgamma_q, float* dgamma_k, int variant, void* stream) {
  OF_REQUIRE(dq && dk && (dv || Hv == 0) && in16 && dqkv16 && gamma_q && gamma_k && dgamma_q && dgamma_k, "of_headnorm_bwd: null pointer");
  OF_REQUIRE(D >= 8 && D <= 8 * kHeadVecs && D % 8 == 0, "of_headnorm_bwd: head dim %d unsupported (8..64, multiple of 8)", D);
  OF_REQUIRE(B >= 1 && L >= 1 && Hq >= 1 && Hk >= 1 && Hv >= 0, "of_headnorm_bwd: bad sizes");
  OF_REQUIRE(dq_ld % 4 == 0 && dq_bs % 4 == 0 && dkv_ld % 4 == 0 && dkv_bs % 4 == 0 && in_ld % 8 == 0 && in_bs % 8 == 0 && o_ld % 8 == 0 &&
                 o_bs % 8 == 0, "of_headnorm_bwd: strides must keep 16-byte alignment");
  OF_REQUIRE(variant >= 0 && variant <= 2 && (variant != 2 || headnorm2_ok(D)), "of_headnorm_bwd: variant %d unsupported for D=%d", variant, D);
  if (variant == 2 || (variant == 0 && headnorm2_ok(D))) {
    int cols, rpc;
    dim3 grid;
    headnorm2_shape(Hq + Hk + Hv, D, (long long)B * L, 2, &cols, &grid, &rpc);
    OF_CHECK_CUDA(launch_pdl(headnorm2_bwd_kernel, grid, dim3(256), (size_t)cols * 8 * sizeof(float), STREAM, dq, dq_ld, dq_bs, dk, dv, dkv_ld,
                             dkv_bs, reinterpret_cast<const __nv_bfloat16*>(in16), in_ld, in_bs, (long long)B * L, L, Hq, Hk, Hv, D, gamma_q,
                             gamma_k, scale, reinterpret_cast<__nv_bfloat16*>(dqkv16), o_ld, o_bs, dgamma_q, dgamma_k, cols, rpc));
    DONE()
  }
  const size_t smem = (size_t)(Hq + Hk) * D * sizeof(float);
  OF_REQUIRE(smem <= 48 * 1024, "of_headnorm_bwd: (Hq + Hk) * D = %d too large", (Hq + Hk) * D);
  long long need = ((long long)B * L * (Hq + Hk + Hv) + 255) / 256;
  long long cap = (long long)device_sm_count() * 2;   // few CTAs: each ends with (Hq + Hk) * D global atomics
  const unsigned grid = (unsigned)(need < cap ? need : cap);
  OF_CHECK_CUDA(launch_pdl(headnorm_bwd_kernel, dim3(grid), dim3(256), smem, STREAM, dq, dq_ld, dq_bs, dk, dv, dkv_ld, dkv_bs,
                           reinterpret_cast<const __nv_bfloat16*>(in16), in_ld, in_bs, B, L, Hq, Hk, Hv, D, gamma_q, gamma_k, scale,
                           reinterpret_cast<__nv_bfloat16*>(dqkv16), o_ld, o_bs, dgamma_q, dgamma_k));
  DONE()
}

extern "C" int of_row_mean_std(const float* a, int B, int C, int N, float* out, void* stream) {
  OF_REQUIRE(a && out && B >= 1 && C >= 1 && N >= 2, "of_row_mean_std: bad args");
  OF_CHECK_CUDA(launch_pdl(row_mean_std_kernel, dim3((unsigned)(B * C)), dim3(256), 0, STREAM, a, C, N, out));
  DONE()
}

extern "C" int of_adaln_fwd(const float* x, long long x_ld, long long x_bs, int B, int L, int C, const float* scale1p, long long sc_ld,
                            const float* shift, long long sh_ld, float eps, void* out_bf16, long long o_ld, long long o_bs,
                            float* mean_rstd, void* stream) {
  OF_REQUIRE(x && scale1p && shift && out_bf16 && mean_rstd, "of_adaln_fwd: null pointer");
  OF_REQUIRE(B >= 1 && L >= 1 && C >= 8 && C % 8 == 0 && C <= 2048, "of_adaln_fwd: unsupported C=%d", C);
  OF_REQUIRE(x_ld % 4 == 0 && x_bs % 4 == 0 && sc_ld % 4 == 0 && sh_ld % 4 == 0 && o_ld % 8 == 0 && o_bs % 8 == 0,
             "of_adaln_fwd: strides must keep 16-byte alignment");
  const long long rows = (long long)B * L;
  const dim3 grid((unsigned)((rows + 7) / 8));
  const int nv = (C / 8 + 31) / 32;
  __nv_bfloat16* o16 = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  if (nv <= 1) OF_CHECK_CUDA(launch_pdl(adaln_fwd_kernel<1>, grid, dim3(256), 0, STREAM, x, x_ld, x_bs, B, L, C, scale1p, sc_ld, shift, sh_ld, eps, o16, o_ld, o_bs, mean_rstd));
  else if (nv <= 2) OF_CHECK_CUDA(launch_pdl(adaln_fwd_kernel<2>, grid, dim3(256), 0, STREAM, x, x_ld, x_bs, B, L, C, scale1p, sc_ld, shift, sh_ld, eps, o16, o_ld, o_bs, mean_rstd));
  else if (nv <= 4) OF_CHECK_CUDA(launch_pdl(adaln_fwd_kernel<4>, grid, dim3(256), 0, STREAM, x, x_ld, x_bs, B, L, C, scale1p, sc_ld, shift, sh_ld, eps, o16, o_ld, o_bs, mean_rstd));
  else OF_CHECK_CUDA(launch_pdl(adaln_fwd_kernel<8>, grid, dim3(256), 0, STREAM, x, x_ld, x_bs, B, L, C, scale1p, sc_ld, shift, sh_ld, eps, o16, o_ld, o_bs, mean_rstd));
  DONE()
}

extern "C" int of_adaln_bwd(const float* dy, long long dy_ld, long long dy_bs, const float* x, long long x_ld, long long x_bs, int B, int L,
                            int C, const float* scale1p, long long sc_ld, const float* mean_rstd, const float* dres, long long r_ld,
                            long long r_bs, float* dx, long long dx_ld, long long dx_bs, float* dscale, long long ds_ld, float* dshift,
                            long long dsh_ld, void* stream) {
  OF_REQUIRE(dy && x && scale1p && mean_rstd && dx && dscale && dshift, "of_adaln_bwd: null pointer");
  OF_REQUIRE(B >= 1 && L >= 1 && C >= 8 && C % 8 == 0 && C <= 2048, "of_adaln_bwd: unsupported C=%d", C);
  OF_REQUIRE(dy_ld % 4 == 0 && dy_bs % 4 == 0 && x_ld % 4 == 0 && x_bs % 4 == 0 && sc_ld % 4 == 0 && dx_ld % 4 == 0 && dx_bs % 4 == 0 &&
                 (!dres || (r_ld % 4 == 0 && r_bs % 4 == 0)), "of_adaln_bwd: strides must keep 16-byte alignment");
  int target = 2 * device_sm_count() / B;            // CTAs per sample: ~2 per SM over the whole grid
  if (target < 1) target = 1;
  int rpw = (L + 8 * target - 1) / (8 * target);
  if (rpw < 1) rpw = 1;
  const dim3 grid((unsigned)((L + 8 * rpw - 1) / (8 * rpw)), (unsigned)B);
  const size_t smem = 2 * (size_t)C * sizeof(float);
  const int nv = (C / 8 + 31) / 32;
  if (nv <= 1) OF_CHECK_CUDA(launch_pdl(adaln_bwd_kernel<1>, grid, dim3(256), smem, STREAM, dy, dy_ld, dy_bs, x, x_ld, x_bs, L, C, scale1p, sc_ld, mean_rstd, dres, r_ld, r_bs, dx, dx_ld, dx_bs, dscale, ds_ld, dshift, dsh_ld, rpw));
  else if (nv <= 2) OF_CHECK_CUDA(launch_pdl(adaln_bwd_kernel<2>, grid, dim3(256), smem, STREAM, dy, dy_ld, dy_bs, x, x_ld, x_bs, L, C, scale1p, sc_ld, mean_rstd, dres, r_ld, r_bs, dx, dx_ld, dx_bs, dscale, ds_ld, dshift, dsh_ld, rpw));
  else if (nv <= 4) OF_CHECK_CUDA(launch_pdl(adaln_bwd_kernel<4>, grid, dim3(256), smem, STREAM, dy, dy_ld, dy_bs, x, x_ld, x_bs, L, C, scale1p, sc_ld, mean_rstd, dres, r_ld, r_bs, dx, dx_ld, dx_bs, dscale, ds_ld, dshift, dsh_ld, rpw));
  else OF_CHECK_CUDA(launch_pdl(adaln_bwd_kernel<8>, grid, dim3(256), smem, STREAM, dy, dy_ld, dy_bs, x, x_ld, x_bs, L, C, scale1p, sc_ld, mean_rstd, dres, r_ld, r_bs, dx, dx_ld, dx_bs, dscale, ds_ld, dshift, dsh_ld, rpw));
  DONE()
}

extern "C" int of_gate_bwd(const float* d32, long long d_ld, long long d_bs, const float* gate, long long gate_ld, const void* y16,
                           long long y_ld, long long y_bs, int round_bf16, int B, int L, int C, void* dy16, long long o_ld, long long o_bs,
                           float* dgate, long long dg_ld, void* stream) {
  OF_REQUIRE(d32 && gate && y16 && dy16 && dgate, "of_gate_bwd: null pointer");
  OF_REQUIRE(B >= 1 && L >= 1 && C >= 8 && C % 8 == 0 && C <= 2048, "of_gate_bwd: unsupported C=%d", C);
  OF_REQUIRE(d_ld % 4 == 0 && d_bs % 4 == 0 && gate_ld % 4 == 0 && y_ld % 8 == 0 && y_bs % 8 == 0 && o_ld % 8 == 0 && o_bs % 8 == 0,
             "of_gate_bwd: strides must keep 16-byte alignment");
  int target = 2 * device_sm_count() / B;
  if (target < 1) target = 1;
  int rpc = (L + target - 1) / target;
  if (rpc < 16) rpc = 16;
  const dim3 grid((unsigned)((L + rpc - 1) / rpc), (unsigned)B);
  OF_CHECK_CUDA(launch_pdl(gate_bwd_kernel, grid, dim3(256), (size_t)C * sizeof(float), STREAM, d32, d_ld, d_bs, gate, gate_ld,
                           reinterpret_cast<const __nv_bfloat16*>(y16), y_ld, y_bs, round_bf16, L, C, reinterpret_cast<__nv_bfloat16*>(dy16),
                           o_ld, o_bs, dgate, dg_ld, rpc));
  DONE()
}
