// ResidualBlock bandwidth kernels: GroupNorm(1,C)+FiLM+SiLU recompute, GlobalContext pooling, gate+residual and
// their backward passes.  See include/osufusion_b200.h (of_rb_*) for the exact contracts and reference citations.
//
// Thread mapping ("channel-owner"): a CTA of 256 threads covers `vecs = C/8` 16-byte channel vectors times
// `rpar = 256/vecs` rows at once; each thread keeps its 8 channels' constants (gamma, beta, FiLM, gate, ...) in
// registers and walks down the rows of its chunk, so per-channel reductions are register-local until one final
// atomic per channel.  Row-wise dot products use the warp-per-row kernel (of_rb_rowdot) instead.
#include "host_common.h"
#include "rowops.cuh"

namespace ofx {

constexpr int kRbThreads = 256;

struct GnCtx {
  int B, L, C;
  float eps;
  const __nv_bfloat16* y;
  long long y_ld, y_bs;
  const double* stats;
  const float* gamma;
  const float* beta;
  const float* ss;
};

__device__ __forceinline__ GnCtx make_ctx(const of_rb_args& a) {
  GnCtx g;
  g.B = a.B; g.L = a.L; g.C = a.C; g.eps = a.eps;
  g.y = reinterpret_cast<const __nv_bfloat16*>(a.y); g.y_ld = a.y_ld; g.y_bs = a.y_bs;
  g.stats = a.stats; g.gamma = a.gamma; g.beta = a.beta; g.ss = a.ss;
  return g;
}

// Per-thread constants for its 8 channels, folded so that the per-element work is two FMAs:
//   xhat = y*rstd + nmr            (nmr = -mean*rstd)
//   z    = y*A + Bc                (A = rstd*gamma, Bc = beta - mean*rstd*gamma: the same folding torch's GroupNorm kernel uses)
template <bool FILM>
struct ChanConst {
  V8 A, Bc, sp1, shift;   // sp1/shift are dead (never materialised) when FILM is false
  float rstd, nmr;
};

template <bool FILM>
__device__ __forceinline__ ChanConst<FILM> load_consts(const GnCtx& g, int b, int c0) {
  ChanConst<FILM> k;
  V8 gamma = ld_f32x8(g.gamma + c0);
  V8 beta = ld_f32x8(g.beta + c0);
  if (FILM) {
    V8 sc = ld_f32x8(g.ss + (long long)b * 2 * g.C + c0);
    k.shift = ld_f32x8(g.ss + (long long)b * 2 * g.C + g.C + c0);
#pragma unroll
    for (int j = 0; j < 8; ++j) k.sp1.v[j] = bf16_round(sc.v[j] + 1.0f);  // reference: bf16 `scale + 1` under autocast
  }
  float mean;
  gn_mean_rstd(g.stats, b, (double)g.L * (double)g.C, g.eps, mean, k.rstd);
  k.nmr = -mean * k.rstd;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    k.A.v[j] = k.rstd * gamma.v[j];
    k.Bc.v[j] = fmaf(k.nmr, gamma.v[j], beta.v[j]);
  }
  return k;
}

// sigmoid via MUFU.EX2 + MUFU.RCP (relative error ~1e-6: far below the bf16 rounding of every tensor these kernels emit)
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// z = GroupNorm(y), f = FiLM(z), h = silu(f) = f*sg for one 8-channel vector (sg = sigmoid(f), reused by the backward kernels).
template <bool FILM>
__device__ __forceinline__ void gn_eval(const ChanConst<FILM>& k, const V8& y, V8& z, V8& f, V8& h, V8& sg) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    z.v[j] = fmaf(y.v[j], k.A.v[j], k.Bc.v[j]);
    f.v[j] = FILM ? fmaf(z.v[j], k.sp1.v[j], k.shift.v[j]) : z.v[j];
    sg.v[j] = sigmoid_fast(f.v[j]);
    h.v[j] = f.v[j] * sg.v[j];
  }
}
template <bool FILM>
__device__ __forceinline__ V8 gn_h(const ChanConst<FILM>& k, const V8& y) {
  V8 z, f, h, sg;
  gn_eval(k, y, z, f, h, sg);
  return h;
}
__device__ __forceinline__ V8 cvt_bf16x8(const uint4& u) {
  V8 r;
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y; r.v[4] = c.x; r.v[5] = c.y; r.v[6] = d.x; r.v[7] = d.y;
  return r;
}
__device__ __forceinline__ uint4 ldg16(const void* p) { return *reinterpret_cast<const uint4*>(p); }

struct Map {
  int vecs, rpar, vi, rsub, c0, b, l_begin, l_end;
  bool active;
};
__device__ __forceinline__ Map make_map(int C, int L, int rows_per_cta) {
  Map m;
  m.vecs = C >> 3;
  m.rpar = blockDim.x / m.vecs;   // the host sizes the CTA as vecs * rpar threads
  m.vi = threadIdx.x % m.vecs;
  m.rsub = threadIdx.x / m.vecs;
  m.active = m.rsub < m.rpar;
  m.c0 = m.vi * 8;
  m.b = blockIdx.y;
  m.l_begin = blockIdx.x * rows_per_cta;
  m.l_end = min(m.l_begin + rows_per_cta, L);
  return m;
}


// Up to 5 per-channel partial arrays reduced across the CTA with 3 barriers in total: sm holds n*C floats.
struct RedSlot {
  const float* part;  // this thread's 8 partials
  float* gdst;        // global [C] destination (nullptr = skip)
};
__device__ __forceinline__ void cta_channel_reduce_multi(const Map& m, int C, const RedSlot* slots, int n, float* sm) {
  for (int c = threadIdx.x; c < n * C; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  if (m.active) {
    for (int s = 0; s < n; ++s) {
      if (slots[s].gdst == nullptr) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&sm[s * C + m.c0 + j], slots[s].part[j]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < n * C; c += blockDim.x) {
    const int s = c / C;
    if (slots[s].gdst != nullptr) atomicAdd(slots[s].gdst + (c - s * C), sm[c]);
  }
}

// Sum the per-thread 8-channel partials of the `rpar` row-lanes of a CTA in shared memory, then one global atomic per channel.
// `sm` holds C floats; all threads of the CTA must call this.
__device__ __forceinline__ void cta_channel_reduce(const Map& m, int C, const float (&part)[8], float* sm, float* gdst) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  if (m.active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&sm[m.c0 + j], part[j]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(gdst + c, sm[c]);
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ forward
// All kernels walk their rows in batches of R: the 16-byte loads of a whole batch are issued before any arithmetic, so a
// thread keeps R (or 2-3 R) independent requests in flight instead of paying one memory latency per row.
#define RB_FOR_BATCH(R) for (int l0 = m.l_begin + m.rsub; l0 < m.l_end; l0 += (R) * m.rpar)
#define RB_ROW(r) (l0 + (r) * m.rpar)

template <bool FILM>
__global__ void __launch_bounds__(kRbThreads, 3) rb_apply_fwd_kernel(const of_rb_args a, const int rpc) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int R = 8;
  GnCtx g = make_ctx(a);
  Map m = make_map(a.C, a.L, rpc);
  if (!m.active) return;
  ChanConst<FILM> k = load_consts<FILM>(g, m.b, m.c0);
  const __nv_bfloat16* yb = g.y + m.b * g.y_bs + m.c0;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.out_bf16) + m.b * a.out_bf16_bs + m.c0;
  RB_FOR_BATCH(R) {
    uint4 raw[R];
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (RB_ROW(r) < m.l_end) raw[r] = ldg16(yb + (long long)RB_ROW(r) * g.y_ld);
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (RB_ROW(r) < m.l_end) st_bf16x8(out + (long long)RB_ROW(r) * a.out_bf16_ld, gn_h(k, cvt_bf16x8(raw[r])));
  }
}

// Row-wise dot products dot(h[b,l,:], vec): channel-owner mapping (constants loaded once per thread), per-row partials
// combined through shared memory (one shuffle-reduced atomic per warp when a warp lies inside one row).
__global__ void __launch_bounds__(kRbThreads, 3) rb_rowdot_kernel(const of_rb_args a, const int rpc) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int R = 8;
  extern __shared__ float s_red[];   // [rpc] row accumulators
  GnCtx g = make_ctx(a);
  Map m = make_map(a.C, a.L, rpc);
  for (int i = threadIdx.x; i < rpc; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();
  if (m.active) {
    ChanConst<false> k = load_consts<false>(g, m.b, m.c0);
    V8 w = ld_f32x8(a.vec + (long long)m.b * a.vec_bs + m.c0);
    if (a.mode == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) w.v[j] = bf16_round(w.v[j]);
    }
    const bool warp_in_row = (m.vecs & 31) == 0;
    const __nv_bfloat16* yb = g.y + m.b * g.y_bs + m.c0;
    RB_FOR_BATCH(R) {
      uint4 raw[R];
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (RB_ROW(r) < m.l_end) raw[r] = ldg16(yb + (long long)RB_ROW(r) * g.y_ld);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (RB_ROW(r) < m.l_end) {
          V8 h = gn_h(k, cvt_bf16x8(raw[r]));
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc = fmaf(bf16_round(h.v[j]), w.v[j], acc);
          if (warp_in_row) {
            acc = warp_sum(acc);
            if ((threadIdx.x & 31) == 0) atomicAdd(&s_red[RB_ROW(r) - m.l_begin], acc);
          } else {
            atomicAdd(&s_red[RB_ROW(r) - m.l_begin], acc);
          }
        }
      }
    }
  }
  __syncthreads();
  const float bias = a.vec_bias ? *a.vec_bias : 0.f;
  for (int i = threadIdx.x; i < rpc; i += blockDim.x) {
    const int l = blockIdx.x * rpc + i;
    if (l < a.L) a.out_rows[(long long)blockIdx.y * a.L + l] = (a.mode == 0) ? bf16_round(s_red[i] + bias) : s_red[i];
  }
}

// GlobalContext forward in ONE pass over y (residual.py:29-32): logits = to_k(h) per row, then softmax-weighted channel sums with a
// CTA-local ONLINE softmax (running maximum, rescaled partial sums) -- the separate pooling pass over y and the softmax launch
// disappear.  Each CTA writes (acc[C], m, z) to `part`; rb_pool_finish_kernel combines the CTAs of a sample and turns the logits
// into the fp32 probabilities the backward pass uses.
__global__ void __launch_bounds__(kRbThreads, 2) rb_logit_pool_kernel(const of_rb_args a, const int rpc, float* __restrict__ part) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int R = 4;
  extern __shared__ float s_red[];              // [C] channel combine | then [rpc] row logits | [2] z, spare
  float* s_log = s_red + a.C;
  float* s_z = s_log + rpc;
  GnCtx g = make_ctx(a);
  Map m = make_map(a.C, a.L, rpc);
  for (int i = threadIdx.x; i < rpc; i += blockDim.x) s_log[i] = 0.f;
  if (threadIdx.x == 0) s_z[0] = 0.f;
  __syncthreads();
  const float bias = a.vec_bias ? *a.vec_bias : 0.f;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float m_run = -INFINITY, z = 0.f;
  ChanConst<false> k;
  V8 w;
  if (m.active) {
    k = load_consts<false>(g, m.b, m.c0);
    w = ld_f32x8(a.vec + m.c0);
#pragma unroll
    for (int j = 0; j < 8; ++j) w.v[j] = bf16_round(w.v[j]);
  }
  const bool warp_in_row = (m.vecs & 31) == 0;
  const __nv_bfloat16* yb = g.y + m.b * g.y_bs + m.c0;
  const int rows_here = m.l_end - m.l_begin;
  for (int base = 0; base < rows_here; base += R * m.rpar) {     // all threads run the same number of batches (barriers inside)
    V8 h[R];
    if (m.active) {
      uint4 raw[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int l = m.l_begin + base + m.rsub + r * m.rpar;
        if (l < m.l_end) raw[r] = ldg16(yb + (long long)l * g.y_ld);
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int li = base + m.rsub + r * m.rpar;
        if (m.l_begin + li < m.l_end) {
          h[r] = gn_h(k, cvt_bf16x8(raw[r]));
          float d = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            h[r].v[j] = bf16_round(h[r].v[j]);       // einsum / conv operands are bf16 under autocast
            d = fmaf(h[r].v[j], w.v[j], d);
          }
          if (warp_in_row) {
            d = warp_sum(d);
            if ((threadIdx.x & 31) == 0) atomicAdd(&s_log[li], d);
          } else {
            atomicAdd(&s_log[li], d);
          }
        }
      }
    }
    __syncthreads();
    // logits of this batch are complete: running maximum over the batch (every thread reads the same values)
    const int nb = min(R * m.rpar, rows_here - base);
    float mb = -INFINITY;
    for (int i = 0; i < nb; ++i) mb = fmaxf(mb, bf16_round(s_log[base + i] + bias));
    const float m_new = fmaxf(m_run, mb);
    const float sc = __expf(m_run - m_new);           // 0 on the first batch (m_run = -inf)
    z *= sc;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= sc;
    if (m.active) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int li = base + m.rsub + r * m.rpar;
        if (m.l_begin + li < m.l_end) {
          const float lg = bf16_round(s_log[li] + bias);
          const float wgt = __expf(lg - m_new);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(wgt, h[r].v[j], acc[j]);
          if (m.vi == 0) {
            z += wgt;
            a.out_rows[(long long)m.b * a.L + m.l_begin + li] = lg;
          }
        }
      }
    }
    m_run = m_new;
  }
  // CTA combine: channel sums through shared memory, z of the row lanes, then plain stores of the partial record
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) s_red[c] = 0.f;
  __syncthreads();
  if (m.active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&s_red[m.c0 + j], acc[j]);
    if (m.vi == 0) atomicAdd(&s_z[0], z);
  }
  __syncthreads();
  float* rec = part + ((long long)m.b * gridDim.x + blockIdx.x) * (a.C + 2);
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) rec[c] = s_red[c];
  if (threadIdx.x == 0) {
    rec[a.C] = m_run;
    rec[a.C + 1] = s_z[0];
  }
}

// gridDim.y CTAs per sample: combine the per-CTA records -> pooled[C] (fp32) and logits -> probabilities in place.  Every CTA
// recomputes the (cheap) global max / normaliser and owns a slice of the channels and of the rows; the record loop keeps four
// independent loads in flight (the first version, one CTA per sample with a serial loop, took 17 us).
__global__ void __launch_bounds__(256) rb_pool_finish_kernel(const float* __restrict__ part, int nparts, int C, float* __restrict__ rows,
                                                             int L, float* __restrict__ pooled) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sm[32];
  __shared__ float s_scale[1024];
  const float* pb = part + (long long)blockIdx.x * nparts * (C + 2);
  const int stride = C + 2;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) mx = fmaxf(mx, pb[(long long)i * stride + C]);
  mx = block_max(mx, sm);
  float zz = 0.f;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) {
    const float sc = __expf(pb[(long long)i * stride + C] - mx);
    s_scale[i] = sc;
    zz += sc * pb[(long long)i * stride + C + 1];
  }
  zz = block_sum(zz, sm);       // (contains the barriers that publish s_scale)
  const float inv = 1.0f / zz;
  const int c_per = (C + gridDim.y - 1) / gridDim.y, c0 = blockIdx.y * c_per, c1 = min(c0 + c_per, C);
  for (int c = c0 + threadIdx.x; c < c1; c += blockDim.x) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int i = 0;
    for (; i + 3 < nparts; i += 4) {
      const float v0 = pb[(long long)i * stride + c], v1 = pb[(long long)(i + 1) * stride + c];
      const float v2 = pb[(long long)(i + 2) * stride + c], v3 = pb[(long long)(i + 3) * stride + c];
      s0 = fmaf(s_scale[i], v0, s0); s1 = fmaf(s_scale[i + 1], v1, s1);
      s2 = fmaf(s_scale[i + 2], v2, s2); s3 = fmaf(s_scale[i + 3], v3, s3);
    }
    for (; i < nparts; ++i) s0 = fmaf(s_scale[i], pb[(long long)i * stride + c], s0);
    pooled[(long long)blockIdx.x * C + c] = ((s0 + s1) + (s2 + s3)) * inv;
  }
  const int l_per = (L + gridDim.y - 1) / gridDim.y, l0 = blockIdx.y * l_per, l1 = min(l0 + l_per, L);
  float* r = rows + (long long)blockIdx.x * L;
  for (int i = l0 + threadIdx.x; i < l1; i += blockDim.x) r[i] = __expf(r[i] - mx) * inv;
}

// one CTA per sample: da = p * (rd - sum_l p*rd) in place on rd   (softmax backward with the fp32 probabilities)
__global__ void __launch_bounds__(1024) softmax_bwd_rows_kernel(const float* __restrict__ p, float* rd, int L) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sm[32];
  const float* pr = p + (long long)blockIdx.x * L;
  float* r = rd + (long long)blockIdx.x * L;
  float s = 0.f;
  for (int i = threadIdx.x; i < L; i += blockDim.x) s += pr[i] * r[i];
  s = block_sum(s, sm);
  for (int i = threadIdx.x; i < L; i += blockDim.x) r[i] = pr[i] * (r[i] - s);
}

// one CTA per sample: p = softmax(logits) in place (fp32; consumers round to bf16 where the reference does)
__global__ void __launch_bounds__(1024) softmax_rows_kernel(float* rows, int L) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sm[32];
  float* r = rows + (long long)blockIdx.x * L;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < L; i += blockDim.x) mx = fmaxf(mx, r[i]);
  mx = block_max(mx, sm);
  float s = 0.f;
  for (int i = threadIdx.x; i < L; i += blockDim.x) s += expf(r[i] - mx);
  s = block_sum(s, sm);
  const float inv = 1.0f / s;
  for (int i = threadIdx.x; i < L; i += blockDim.x) r[i] = expf(r[i] - mx) * inv;
}

__global__ void __launch_bounds__(kRbThreads, 3) rb_pool_kernel(const of_rb_args a, const int rpc) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int R = 8;
  extern __shared__ float s_red[];
  GnCtx g = make_ctx(a);
  Map m = make_map(a.C, a.L, rpc);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (m.active) {
    ChanConst<false> k = load_consts<false>(g, m.b, m.c0);
    const __nv_bfloat16* yb = g.y + m.b * g.y_bs + m.c0;
    const float* pb = a.p + (long long)m.b * a.L;
    RB_FOR_BATCH(R) {
      uint4 raw[R];
      float pl[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (RB_ROW(r) < m.l_end) {
          raw[r] = ldg16(yb + (long long)RB_ROW(r) * g.y_ld);
          pl[r] = pb[RB_ROW(r)];
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (RB_ROW(r) < m.l_end) {
          V8 h = gn_h(k, cvt_bf16x8(raw[r]));
          const float pw = bf16_round(pl[r]);  // einsum operand is cast to bf16 (autocast)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(bf16_round(h.v[j]), pw, acc[j]);
        }
      }
    }
  }
  cta_channel_reduce(m, a.C, acc, s_red, a.acc_bc + (long long)m.b * a.C);
}

__global__ void __launch_bounds__(kRbThreads, 2) rb_gate_fwd_kernel(const of_rb_args a, const int rpc) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int R = 4;
  GnCtx g = make_ctx(a);
  Map m = make_map(a.C, a.L, rpc);
  if (!m.active) return;
  ChanConst<false> k = load_consts<false>(g, m.b, m.c0);
  V8 gate = ld_f32x8(a.gate + (long long)m.b * a.C + m.c0);
  const __nv_bfloat16* yb = g.y + m.b * g.y_bs + m.c0;
  const __nv_bfloat16* r16 = a.res_bf16 ? reinterpret_cast<const __nv_bfloat16*>(a.res_bf16) + m.b * a.res_bf16_bs + m.c0 : nullptr;
  const float* r32 = a.res_f32 ? a.res_f32 + m.b * a.res_f32_bs + m.c0 : nullptr;
  __nv_bfloat16* o16 = a.out_bf16 ? reinterpret_cast<__nv_bfloat16*>(a.out_bf16) + m.b * a.out_bf16_bs + m.c0 : nullptr;
  float* o32 = a.out_f32 ? a.out_f32 + m.b * a.out_f32_bs + m.c0 : nullptr;
  RB_FOR_BATCH(R) {
    uint4 raw[R];
    V8 res[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (RB_ROW(r) < m.l_end) {
        raw[r] = ldg16(yb + (long long)RB_ROW(r) * g.y_ld);
        if (r32) res[r] = ld_f32x8(r32 + (long long)RB_ROW(r) * a.res_f32_ld);
        else res[r] = cvt_bf16x8(ldg16(r16 + (long long)RB_ROW(r) * a.res_bf16_ld));
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (RB_ROW(r) < m.l_end) {
        V8 h = gn_h(k, cvt_bf16x8(raw[r])), o;
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] = fmaf(h.v[j], gate.v[j], res[r].v[j]);
        if (o32) st_f32x8(o32 + (long long)RB_ROW(r) * a.out_f32_ld, o);
        if (o16) st_bf16x8(o16 + (long long)RB_ROW(r) * a.out_bf16_ld, o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward
__global__ void __launch_bounds__(kRbThreads, 3) rb_gate_bwd_reduce_kernel(const of_rb_args a, const int rpc) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int R = 4;
  extern __shared__ float s_red[];
  GnCtx g = make_ctx(a);
  Map m = make_map(a.C, a.L, rpc);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (m.active) {
    ChanConst<false> k = load_consts<false>(g, m.b, m.c0);
    const __nv_bfloat16* yb = g.y + m.b * g.y_bs + m.c0;
    const float* db = a.dout_f32 + m.b * a.dout_f32_bs + m.c0;
    RB_FOR_BATCH(R) {
      uint4 raw[R];
      V8 d[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (RB_ROW(r) < m.l_end) {
          raw[r] = ldg16(yb + (long long)RB_ROW(r) * g.y_ld);
          d[r] = ld_f32x8(db + (long long)RB_ROW(r) * a.dout_f32_ld);
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (RB_ROW(r) < m.l_end) {
          V8 h = gn_h(k, cvt_bf16x8(raw[r]));
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(d[r].v[j], h.v[j], acc[j]);
        }
      }
    }
  }
  cta_channel_reduce(m, a.C, acc, s_red, a.acc_bc + (long long)m.b * a.C);
}

template <int MODE, bool FILM>
__global__ void __launch_bounds__(kRbThreads, 2) rb_bwd_pass1_kernel(const of_rb_args a, const int rpc) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int R = MODE == 0 ? 2 : 4;
  __shared__ float sm[32];
  extern __shared__ float s_red[];
  GnCtx g = make_ctx(a);
  Map m = make_map(a.C, a.L, rpc);
  float s1 = 0.f, s2 = 0.f, sda = 0.f;
  float dgam[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dbet[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float dsc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dsh[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dwk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (m.active) {
    ChanConst<FILM> k = load_consts<FILM>(g, m.b, m.c0);
    const V8 gamma = ld_f32x8(g.gamma + m.c0);
    V8 gate, dpool, wk;
    if (MODE == 0) {
      gate = ld_f32x8(a.gate + (long long)m.b * a.C + m.c0);
      dpool = ld_f32x8(a.dpooled + (long long)m.b * a.C + m.c0);
      wk = ld_f32x8(a.wk + m.c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) wk.v[j] = bf16_round(wk.v[j]);
    }
    const __nv_bfloat16* yb = g.y + m.b * g.y_bs + m.c0;
    const __nv_bfloat16* dh16 = MODE == 1 ? reinterpret_cast<const __nv_bfloat16*>(a.dh_bf16) + m.b * a.dh_bs + m.c0 : nullptr;
    const float* d32 = MODE == 0 ? a.dout_f32 + m.b * a.dout_f32_bs + m.c0 : nullptr;
    __nv_bfloat16* dxh = reinterpret_cast<__nv_bfloat16*>(a.dxhat_bf16) + m.b * a.dxhat_bs + m.c0;
    __nv_bfloat16* do16 = (MODE == 0 && a.dout_bf16) ? reinterpret_cast<__nv_bfloat16*>(a.dout_bf16) + m.b * a.dout_bf16_bs + m.c0 : nullptr;
    const float* pb = MODE == 0 ? a.p + (long long)m.b * a.L : nullptr;
    const float* dab = MODE == 0 ? a.da + (long long)m.b * a.L : nullptr;
    RB_FOR_BATCH(R) {
      uint4 raw[R];
      V8 din[R];
      float pl[R], dal[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (RB_ROW(r) < m.l_end) {
          raw[r] = ldg16(yb + (long long)RB_ROW(r) * g.y_ld);
          if (MODE == 0) {
            din[r] = ld_f32x8(d32 + (long long)RB_ROW(r) * a.dout_f32_ld);
            pl[r] = pb[RB_ROW(r)];
            dal[r] = dab[RB_ROW(r)];
          } else {
            din[r] = cvt_bf16x8(ldg16(dh16 + (long long)RB_ROW(r) * a.dh_ld));
          }
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (RB_ROW(r) < m.l_end) {
          const long long l = RB_ROW(r);
          V8 y = cvt_bf16x8(raw[r]), z, f, h, sg, dh, dx;
          gn_eval(k, y, z, f, h, sg);
          if (MODE == 0) {
            const float pw = bf16_round(pl[r]);
            const float da = dal[r];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              dh.v[j] = fmaf(din[r].v[j], gate.v[j], fmaf(dpool.v[j], pw, da * wk.v[j]));
              dwk[j] = fmaf(da, bf16_round(h.v[j]), dwk[j]);
            }
            if (m.vi == 0) sda += da;
            if (do16) st_bf16x8(do16 + l * a.dout_bf16_ld, din[r]);
          } else {
            dh = din[r];
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            // silu'(f) = s*(1 + f*(1-s)), s = sigmoid(f) from the forward recompute
            const float s = sg.v[j];
            float df = dh.v[j] * (s * fmaf(f.v[j], 1.0f - s, 1.0f));
            float dz = df;
            if (FILM) {
              dsc[j] = fmaf(df, z.v[j], dsc[j]);
              dsh[j] += df;
              dz = df * k.sp1.v[j];
            }
            const float xh = fmaf(y.v[j], k.rstd, k.nmr);
            dgam[j] = fmaf(dz, xh, dgam[j]);
            dbet[j] += dz;
            dx.v[j] = dz * gamma.v[j];
            float dxr = bf16_round(dx.v[j]);  // what pass 2 will read back
            s1 += dxr;
            s2 = fmaf(dxr, xh, s2);
          }
          st_bf16x8(dxh + l * a.dxhat_ld, dx);
        }
      }
    }
  }
  // CTA-level combine of the per-channel partials in shared memory (accumulators stay in registers: no indirection)
  float* dst[5] = {a.dgamma, a.dbeta, FILM ? a.dss + (long long)blockIdx.y * 2 * a.C : nullptr,
                   FILM ? a.dss + (long long)blockIdx.y * 2 * a.C + a.C : nullptr, MODE == 0 ? a.dwk : nullptr};
  for (int c = threadIdx.x; c < 5 * a.C; c += blockDim.x) s_red[c] = 0.f;
  __syncthreads();
  if (m.active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&s_red[0 * a.C + m.c0 + j], dgam[j]);
      atomicAdd(&s_red[1 * a.C + m.c0 + j], dbet[j]);
      if (FILM) {
        atomicAdd(&s_red[2 * a.C + m.c0 + j], dsc[j]);
        atomicAdd(&s_red[3 * a.C + m.c0 + j], dsh[j]);
      }
      if (MODE == 0) atomicAdd(&s_red[4 * a.C + m.c0 + j], dwk[j]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 5 * a.C; c += blockDim.x) {
    const int sl = c / a.C;
    if (dst[sl] != nullptr) atomicAdd(dst[sl] + (c - sl * a.C), s_red[c]);
  }
  s1 = block_sum(s1, sm);
  s2 = block_sum(s2, sm);
  if (MODE == 0) sda = block_sum(sda, sm);
  if (threadIdx.x == 0) {
    atomicAdd(a.dstats + 2 * blockIdx.y, (double)s1);
    atomicAdd(a.dstats + 2 * blockIdx.y + 1, (double)s2);
    if (MODE == 0 && a.dbk) atomicAdd(a.dbk, sda);
  }
}

__global__ void __launch_bounds__(kRbThreads, 3) rb_bwd_apply_kernel(const of_rb_args a, const int rpc) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int R = 4;
  extern __shared__ float s_red[];
  GnCtx g = make_ctx(a);
  Map m = make_map(a.C, a.L, rpc);
  float mean, rstd;
  const double n = (double)a.L * (double)a.C;
  gn_mean_rstd(g.stats, m.b, n, g.eps, mean, rstd);
  const float nmr = -mean * rstd;
  const float m1 = (float)(a.dstats[2 * m.b] / n), m2 = (float)(a.dstats[2 * m.b + 1] / n);
  float db[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (m.active) {
    const __nv_bfloat16* yb = g.y + m.b * g.y_bs + m.c0;
    const __nv_bfloat16* dxh = reinterpret_cast<const __nv_bfloat16*>(a.dxhat_bf16) + m.b * a.dxhat_bs + m.c0;
    __nv_bfloat16* dy = reinterpret_cast<__nv_bfloat16*>(a.dy_bf16) + m.b * a.dy_bs + m.c0;
    RB_FOR_BATCH(R) {
      uint4 raw[R], rdx[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (RB_ROW(r) < m.l_end) {
          raw[r] = ldg16(yb + (long long)RB_ROW(r) * g.y_ld);
          rdx[r] = ldg16(dxh + (long long)RB_ROW(r) * a.dxhat_ld);
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (RB_ROW(r) < m.l_end) {
          V8 y = cvt_bf16x8(raw[r]), dx = cvt_bf16x8(rdx[r]), o;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xh = fmaf(y.v[j], rstd, nmr);
            o.v[j] = rstd * (dx.v[j] - m1 - xh * m2);
            db[j] += bf16_round(o.v[j]);
          }
          st_bf16x8(dy + (long long)RB_ROW(r) * a.dy_ld, o);
        }
      }
    }
  }
  if (a.dbias) cta_channel_reduce(m, a.C, db, s_red, a.dbias);
}

static int check_common(const of_rb_args* a, const char* who) {
  OF_REQUIRE(a != nullptr, "%s: null args", who);
  OF_REQUIRE(a->B >= 1 && a->L >= 1 && a->C >= 8 && a->C % 8 == 0 && a->C <= 2048, "%s: unsupported C=%d (need C%%8==0, C<=2048)",
             who, a->C);
  OF_REQUIRE(a->y && a->stats && a->gamma && a->beta, "%s: null GroupNorm input", who);
  OF_REQUIRE(a->y_ld % 8 == 0, "%s: y_ld %% 8", who);
  return OF_OK;
}
// CTA = vecs * rpar threads (rpar = rows processed side by side): 256 threads when C/8 divides 256, e.g. 192 for C = 1536.
static int rb_threads(const of_rb_args* a) {
  const int vecs = a->C / 8;
  const int rpar = kRbThreads / vecs > 0 ? kRbThreads / vecs : 1;
  return vecs * rpar;
}
// rows per CTA: one wave of ~4 CTAs per SM, a whole number of row-lanes, at most 64 rows per row-lane
static int rb_rows_per_cta(const of_rb_args* a, int ctas_per_sm) {
  const int vecs = a->C / 8;
  const int rpar = kRbThreads / vecs > 0 ? kRbThreads / vecs : 1;
  const long long slots = (long long)ctas_per_sm * device_sm_count();
  long long per_sample = (slots + a->B - 1) / a->B;           // CTAs available to one sample
  if (per_sample < 1) per_sample = 1;
  long long r = (a->L + per_sample - 1) / per_sample;
  r = (r + rpar - 1) / rpar * rpar;
  if (r < rpar) r = rpar;
  if (r > 64LL * rpar) r = 64LL * rpar;
  return (int)r;
}
static dim3 rb_grid(const of_rb_args* a, int rpc) { return dim3((a->L + rpc - 1) / rpc, a->B); }

}  // namespace ofx

using namespace ofx;

#define RB_LAUNCH(kernel, ctas_per_sm)                                                                                       \
  const int rpc = rb_rows_per_cta(a, ctas_per_sm);                                                                                      \
  OF_CHECK_CUDA(launch_pdl(kernel, rb_grid(a, rpc), dim3(rb_threads(a)), 5 * (size_t)a->C * sizeof(float),                     \
                           reinterpret_cast<cudaStream_t>(stream), *a, rpc));                                              \
  count_launch();                                                                 \
  return OF_OK;

extern "C" int of_rb_apply_fwd(const of_rb_args* a, void* stream) {
  int rc = check_common(a, "of_rb_apply_fwd");
  if (rc) return rc;
  OF_REQUIRE(a->out_bf16 && a->out_bf16_ld % 8 == 0, "of_rb_apply_fwd: bad out_bf16");
  if (a->ss) { RB_LAUNCH(rb_apply_fwd_kernel<true>, 3) }
  RB_LAUNCH(rb_apply_fwd_kernel<false>, 3)
}
extern "C" int of_rb_rowdot(const of_rb_args* a, void* stream) {
  int rc = check_common(a, "of_rb_rowdot");
  if (rc) return rc;
  OF_REQUIRE(a->ss == nullptr, "of_rb_rowdot: FiLM (ss) is only supported by of_rb_apply_fwd / of_rb_bwd_pass1(mode 1)");
  OF_REQUIRE(a->vec && a->out_rows, "of_rb_rowdot: null vec/out_rows");
  const int rpc = rb_rows_per_cta(a, 3);
  OF_CHECK_CUDA(launch_pdl(rb_rowdot_kernel, rb_grid(a, rpc), dim3(rb_threads(a)), (size_t)(rpc > 5 * a->C ? rpc : 5 * a->C) * sizeof(float),
                           reinterpret_cast<cudaStream_t>(stream), *a, rpc));
  count_launch();
  return OF_OK;
}
// Number of per-CTA partial records per sample of of_rb_logit_pool (the caller allocates part[B][n][C+2] floats).
extern "C" int of_rb_pool_parts(const of_rb_args* a) {
  if (a == nullptr || a->C < 8) return -1;
  const int rpc = rb_rows_per_cta(a, 2);
  return (a->L + rpc - 1) / rpc;
}
extern "C" int of_rb_logit_pool(const of_rb_args* a, float* part, float* pooled, void* stream) {
  int rc = check_common(a, "of_rb_logit_pool");
  if (rc) return rc;
  OF_REQUIRE(a->ss == nullptr, "of_rb_logit_pool: FiLM (ss) is not supported on block2");
  OF_REQUIRE(a->vec && a->out_rows && part && pooled, "of_rb_logit_pool: null vec/out_rows/part/pooled");
  const int rpc = rb_rows_per_cta(a, 2);
  const dim3 grid = rb_grid(a, rpc);
  OF_REQUIRE(grid.x <= 1024, "of_rb_logit_pool: too many partial records (%u)", grid.x);
  OF_CHECK_CUDA(launch_pdl(rb_logit_pool_kernel, grid, dim3(rb_threads(a)), ((size_t)a->C + rpc + 2) * sizeof(float),
                           reinterpret_cast<cudaStream_t>(stream), *a, rpc, part));
  count_launch();
  OF_CHECK_CUDA(launch_pdl(rb_pool_finish_kernel, dim3(a->B, 8), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
                           (const float*)part, (int)grid.x, a->C, a->out_rows, a->L, pooled));
  count_launch();
  return OF_OK;
}
extern "C" int of_softmax_rows(float* rows, int B, int L, void* stream) {
  OF_REQUIRE(rows && B >= 1 && L >= 1, "of_softmax_rows: bad args");
  OF_CHECK_CUDA(launch_pdl(softmax_rows_kernel, dim3(B), dim3(1024), 0, reinterpret_cast<cudaStream_t>(stream), rows, L));
  count_launch();
  return OF_OK;
}
extern "C" int of_softmax_bwd_rows(const float* p, float* rd, int B, int L, void* stream) {
  OF_REQUIRE(p && rd && B >= 1 && L >= 1, "of_softmax_bwd_rows: bad args");
  OF_CHECK_CUDA(launch_pdl(softmax_bwd_rows_kernel, dim3(B), dim3(1024), 0, reinterpret_cast<cudaStream_t>(stream), p, rd, L));
  count_launch();
  return OF_OK;
}
extern "C" int of_rb_pool(const of_rb_args* a, void* stream) {
  int rc = check_common(a, "of_rb_pool");
  if (rc) return rc;
  OF_REQUIRE(a->ss == nullptr, "of_rb_pool: FiLM (ss) is only supported by of_rb_apply_fwd / of_rb_bwd_pass1(mode 1)");
  OF_REQUIRE(a->p && a->acc_bc, "of_rb_pool: null p/acc_bc");
  RB_LAUNCH(rb_pool_kernel, 3)
}
extern "C" int of_rb_gate_fwd(const of_rb_args* a, void* stream) {
  int rc = check_common(a, "of_rb_gate_fwd");
  if (rc) return rc;
  OF_REQUIRE(a->ss == nullptr, "of_rb_gate_fwd: FiLM (ss) is only supported by of_rb_apply_fwd / of_rb_bwd_pass1(mode 1)");
  OF_REQUIRE(a->gate && (a->res_f32 || a->res_bf16) && (a->out_f32 || a->out_bf16), "of_rb_gate_fwd: null gate/res/out");
  RB_LAUNCH(rb_gate_fwd_kernel, 2)
}
extern "C" int of_rb_gate_bwd_reduce(const of_rb_args* a, void* stream) {
  int rc = check_common(a, "of_rb_gate_bwd_reduce");
  if (rc) return rc;
  OF_REQUIRE(a->ss == nullptr, "of_rb_gate_bwd_reduce: FiLM (ss) is only supported by of_rb_apply_fwd / of_rb_bwd_pass1(mode 1)");
  OF_REQUIRE(a->dout_f32 && a->acc_bc, "of_rb_gate_bwd_reduce: null dout/acc");
  RB_LAUNCH(rb_gate_bwd_reduce_kernel, 3)
}
extern "C" int of_rb_bwd_pass1(const of_rb_args* a, void* stream) {
  int rc = check_common(a, "of_rb_bwd_pass1");
  if (rc) return rc;
  OF_REQUIRE(a->dxhat_bf16 && a->dstats && a->dgamma && a->dbeta, "of_rb_bwd_pass1: null outputs");
  if (a->ss) OF_REQUIRE(a->dss, "of_rb_bwd_pass1: dss required with FiLM");
  if (a->mode == 0)
    OF_REQUIRE(a->dout_f32 && a->gate && a->dpooled && a->p && a->da && a->wk && a->dwk, "of_rb_bwd_pass1(mode 0): null inputs");
  else
    OF_REQUIRE(a->dh_bf16, "of_rb_bwd_pass1(mode 1): null dh");
  if (a->mode == 0) {
    OF_REQUIRE(a->ss == nullptr, "of_rb_bwd_pass1(mode 0): FiLM is not supported on block2");
    RB_LAUNCH((rb_bwd_pass1_kernel<0, false>), 2)
  }
  if (a->ss) { RB_LAUNCH((rb_bwd_pass1_kernel<1, true>), 2) }
  RB_LAUNCH((rb_bwd_pass1_kernel<1, false>), 2)
}
extern "C" int of_rb_bwd_apply(const of_rb_args* a, void* stream) {
  int rc = check_common(a, "of_rb_bwd_apply");
  if (rc) return rc;
  OF_REQUIRE(a->dxhat_bf16 && a->dstats && a->dy_bf16, "of_rb_bwd_apply: null pointers");
  RB_LAUNCH(rb_bwd_apply_kernel, 3)
}
