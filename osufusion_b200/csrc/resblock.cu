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

// Per-thread constants for its 8 channels.
struct ChanConst {
  V8 gamma, beta, sp1, shift;
  float mean, rstd;
  bool film;
};

__device__ __forceinline__ ChanConst load_consts(const GnCtx& g, int b, int c0) {
  ChanConst k;
  k.gamma = ld_f32x8(g.gamma + c0);
  k.beta = ld_f32x8(g.beta + c0);
  k.film = g.ss != nullptr;
  if (k.film) {
    V8 sc = ld_f32x8(g.ss + (long long)b * 2 * g.C + c0);
    k.shift = ld_f32x8(g.ss + (long long)b * 2 * g.C + g.C + c0);
#pragma unroll
    for (int j = 0; j < 8; ++j) k.sp1.v[j] = bf16_round(sc.v[j] + 1.0f);  // reference: bf16 `scale + 1` under autocast
  }
  gn_mean_rstd(g.stats, b, (double)g.L * (double)g.C, g.eps, k.mean, k.rstd);
  return k;
}

// xhat, z = xhat*gamma+beta, f = FiLM(z), h = silu(f) for one 8-channel vector.
__device__ __forceinline__ void gn_eval(const ChanConst& k, const V8& y, V8& xhat, V8& z, V8& f, V8& h) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    xhat.v[j] = (y.v[j] - k.mean) * k.rstd;
    z.v[j] = xhat.v[j] * k.gamma.v[j] + k.beta.v[j];
    f.v[j] = k.film ? (z.v[j] * k.sp1.v[j] + k.shift.v[j]) : z.v[j];
    h.v[j] = silu_acc(f.v[j]);
  }
}

struct Map {
  int vecs, rpar, vi, rsub, c0, b, l_begin, l_end;
  bool active;
};
__device__ __forceinline__ Map make_map(int C, int L, int rows_per_cta) {
  Map m;
  m.vecs = C >> 3;
  m.rpar = kRbThreads / m.vecs;
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
__global__ void __launch_bounds__(kRbThreads) rb_apply_fwd_kernel(const of_rb_args a, const int rpc) {
  GnCtx g = make_ctx(a);
  Map m = make_map(a.C, a.L, rpc);
  if (!m.active) return;
  ChanConst k = load_consts(g, m.b, m.c0);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.out_bf16);
  for (int l = m.l_begin + m.rsub; l < m.l_end; l += m.rpar) {
    V8 y = ld_bf16x8(g.y + m.b * g.y_bs + (long long)l * g.y_ld + m.c0), xh, z, f, h;
    gn_eval(k, y, xh, z, f, h);
    st_bf16x8(out + m.b * a.out_bf16_bs + (long long)l * a.out_bf16_ld + m.c0, h);
  }
}

// Row-wise dot products dot(h[b,l,:], vec): channel-owner mapping (constants loaded once per thread), per-row partials
// combined through shared memory (one shuffle-reduced atomic per warp when a warp lies inside one row).
__global__ void __launch_bounds__(kRbThreads) rb_rowdot_kernel(const of_rb_args a, const int rpc) {
  extern __shared__ float s_red[];   // [rpc] row accumulators
  GnCtx g = make_ctx(a);
  Map m = make_map(a.C, a.L, rpc);
  for (int i = threadIdx.x; i < rpc; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();
  if (m.active) {
    ChanConst k = load_consts(g, m.b, m.c0);
    V8 w = ld_f32x8(a.vec + (long long)m.b * a.vec_bs + m.c0);
    if (a.mode == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) w.v[j] = bf16_round(w.v[j]);
    }
    const bool warp_in_row = (m.vecs & 31) == 0;
    for (int l = m.l_begin + m.rsub; l < m.l_end; l += m.rpar) {
      V8 y = ld_bf16x8(g.y + m.b * g.y_bs + (long long)l * g.y_ld + m.c0), xh, z, f, h;
      gn_eval(k, y, xh, z, f, h);
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += bf16_round(h.v[j]) * w.v[j];
      if (warp_in_row) {
        acc = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) atomicAdd(&s_red[l - m.l_begin], acc);
      } else {
        atomicAdd(&s_red[l - m.l_begin], acc);
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

// one CTA per sample: da = p * (rd - sum_l p*rd) in place on rd   (softmax backward with the fp32 probabilities)
__global__ void __launch_bounds__(1024) softmax_bwd_rows_kernel(const float* __restrict__ p, float* rd, int L) {
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

__global__ void __launch_bounds__(kRbThreads) rb_pool_kernel(const of_rb_args a, const int rpc) {
  extern __shared__ float s_red[];
  GnCtx g = make_ctx(a);
  Map m = make_map(a.C, a.L, rpc);
  ChanConst k;
  if (m.active) k = load_consts(g, m.b, m.c0);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int l = m.l_begin + m.rsub; m.active && l < m.l_end; l += m.rpar) {
    V8 y = ld_bf16x8(g.y + m.b * g.y_bs + (long long)l * g.y_ld + m.c0), xh, z, f, h;
    gn_eval(k, y, xh, z, f, h);
    const float pl = bf16_round(a.p[(long long)m.b * a.L + l]);  // einsum operand is cast to bf16 (autocast)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += bf16_round(h.v[j]) * pl;
  }
  cta_channel_reduce(m, a.C, acc, s_red, a.acc_bc + (long long)m.b * a.C);
}

__global__ void __launch_bounds__(kRbThreads) rb_gate_fwd_kernel(const of_rb_args a, const int rpc) {
  GnCtx g = make_ctx(a);
  Map m = make_map(a.C, a.L, rpc);
  if (!m.active) return;
  ChanConst k = load_consts(g, m.b, m.c0);
  V8 gate = ld_f32x8(a.gate + (long long)m.b * a.C + m.c0);
  const __nv_bfloat16* r16 = reinterpret_cast<const __nv_bfloat16*>(a.res_bf16);
  __nv_bfloat16* o16 = reinterpret_cast<__nv_bfloat16*>(a.out_bf16);
  for (int l = m.l_begin + m.rsub; l < m.l_end; l += m.rpar) {
    V8 y = ld_bf16x8(g.y + m.b * g.y_bs + (long long)l * g.y_ld + m.c0), xh, z, f, h, o;
    gn_eval(k, y, xh, z, f, h);
    V8 r;
    if (a.res_f32) r = ld_f32x8(a.res_f32 + m.b * a.res_f32_bs + (long long)l * a.res_f32_ld + m.c0);
    else r = ld_bf16x8(r16 + m.b * a.res_bf16_bs + (long long)l * a.res_bf16_ld + m.c0);
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = h.v[j] * gate.v[j] + r.v[j];
    if (a.out_f32) st_f32x8(a.out_f32 + m.b * a.out_f32_bs + (long long)l * a.out_f32_ld + m.c0, o);
    if (o16) st_bf16x8(o16 + m.b * a.out_bf16_bs + (long long)l * a.out_bf16_ld + m.c0, o);
  }
}

// ------------------------------------------------------------------------------------------------ backward
__global__ void __launch_bounds__(kRbThreads) rb_gate_bwd_reduce_kernel(const of_rb_args a, const int rpc) {
  extern __shared__ float s_red[];
  GnCtx g = make_ctx(a);
  Map m = make_map(a.C, a.L, rpc);
  ChanConst k;
  if (m.active) k = load_consts(g, m.b, m.c0);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int l = m.l_begin + m.rsub; m.active && l < m.l_end; l += m.rpar) {
    V8 y = ld_bf16x8(g.y + m.b * g.y_bs + (long long)l * g.y_ld + m.c0), xh, z, f, h;
    gn_eval(k, y, xh, z, f, h);
    V8 d = ld_f32x8(a.dout_f32 + m.b * a.dout_f32_bs + (long long)l * a.dout_f32_ld + m.c0);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += d.v[j] * h.v[j];
  }
  cta_channel_reduce(m, a.C, acc, s_red, a.acc_bc + (long long)m.b * a.C);
}

template <int MODE>
__global__ void __launch_bounds__(kRbThreads, 2) rb_bwd_pass1_kernel(const of_rb_args a, const int rpc) {
  __shared__ float sm[32];
  extern __shared__ float s_red[];
  GnCtx g = make_ctx(a);
  Map m = make_map(a.C, a.L, rpc);
  float s1 = 0.f, s2 = 0.f, sda = 0.f;
  float dgam[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dbet[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float dsc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dsh[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dwk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (m.active) {
    ChanConst k = load_consts(g, m.b, m.c0);
    V8 gate, dpool, wk;
    if (MODE == 0) {
      gate = ld_f32x8(a.gate + (long long)m.b * a.C + m.c0);
      dpool = ld_f32x8(a.dpooled + (long long)m.b * a.C + m.c0);
      wk = ld_f32x8(a.wk + m.c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) wk.v[j] = bf16_round(wk.v[j]);
    }
    const __nv_bfloat16* dh16 = reinterpret_cast<const __nv_bfloat16*>(a.dh_bf16);
    __nv_bfloat16* dxh = reinterpret_cast<__nv_bfloat16*>(a.dxhat_bf16);
    __nv_bfloat16* do16 = reinterpret_cast<__nv_bfloat16*>(a.dout_bf16);
    for (int l = m.l_begin + m.rsub; l < m.l_end; l += m.rpar) {
      V8 y = ld_bf16x8(g.y + m.b * g.y_bs + (long long)l * g.y_ld + m.c0), xh, z, f, h, dh, dx;
      gn_eval(k, y, xh, z, f, h);
      if (MODE == 0) {
        V8 d = ld_f32x8(a.dout_f32 + m.b * a.dout_f32_bs + (long long)l * a.dout_f32_ld + m.c0);
        const float pl = bf16_round(a.p[(long long)m.b * a.L + l]);
        const float da = a.da[(long long)m.b * a.L + l];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          dh.v[j] = d.v[j] * gate.v[j] + dpool.v[j] * pl + da * wk.v[j];
          dwk[j] += da * bf16_round(h.v[j]);
        }
        if (m.vi == 0) sda += da;
        if (do16) st_bf16x8(do16 + m.b * a.dout_bf16_bs + (long long)l * a.dout_bf16_ld + m.c0, d);
      } else {
        dh = ld_bf16x8(dh16 + m.b * a.dh_bs + (long long)l * a.dh_ld + m.c0);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        // silu'(f) = s*(1 + f*(1-s)) with s = h/f reused from the forward recompute (s = sigmoid(f))
        const float sg = 1.0f / (1.0f + expf(-f.v[j]));
        float df = dh.v[j] * (sg * (1.0f + f.v[j] * (1.0f - sg)));
        float dz = df;
        if (MODE == 1 && k.film) {
          dsc[j] += df * z.v[j];
          dsh[j] += df;
          dz = df * k.sp1.v[j];
        }
        dgam[j] += dz * xh.v[j];
        dbet[j] += dz;
        dx.v[j] = dz * k.gamma.v[j];
        float dxr = bf16_round(dx.v[j]);  // what pass 2 will read back
        s1 += dxr;
        s2 += dxr * xh.v[j];
      }
      st_bf16x8(dxh + m.b * a.dxhat_bs + (long long)l * a.dxhat_ld + m.c0, dx);
    }
  }
  // CTA-level combine of the per-channel partials in shared memory (accumulators stay in registers: no indirection)
  float* dst[5] = {a.dgamma, a.dbeta, (MODE == 1 && a.ss) ? a.dss + (long long)blockIdx.y * 2 * a.C : nullptr,
                   (MODE == 1 && a.ss) ? a.dss + (long long)blockIdx.y * 2 * a.C + a.C : nullptr, MODE == 0 ? a.dwk : nullptr};
  for (int c = threadIdx.x; c < 5 * a.C; c += blockDim.x) s_red[c] = 0.f;
  __syncthreads();
  if (m.active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&s_red[0 * a.C + m.c0 + j], dgam[j]);
      atomicAdd(&s_red[1 * a.C + m.c0 + j], dbet[j]);
      if (MODE == 1 && a.ss) {
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

__global__ void __launch_bounds__(kRbThreads) rb_bwd_apply_kernel(const of_rb_args a, const int rpc) {
  extern __shared__ float s_red[];
  GnCtx g = make_ctx(a);
  Map m = make_map(a.C, a.L, rpc);
  float mean, rstd;
  const double n = (double)a.L * (double)a.C;
  gn_mean_rstd(g.stats, m.b, n, g.eps, mean, rstd);
  const float m1 = (float)(a.dstats[2 * m.b] / n), m2 = (float)(a.dstats[2 * m.b + 1] / n);
  const __nv_bfloat16* dxh = reinterpret_cast<const __nv_bfloat16*>(a.dxhat_bf16);
  __nv_bfloat16* dy = reinterpret_cast<__nv_bfloat16*>(a.dy_bf16);
  float db[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int l = m.l_begin + m.rsub; m.active && l < m.l_end; l += m.rpar) {
    V8 y = ld_bf16x8(g.y + m.b * g.y_bs + (long long)l * g.y_ld + m.c0);
    V8 dx = ld_bf16x8(dxh + m.b * a.dxhat_bs + (long long)l * a.dxhat_ld + m.c0), o;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float xh = (y.v[j] - mean) * rstd;
      o.v[j] = rstd * (dx.v[j] - m1 - xh * m2);
      db[j] += bf16_round(o.v[j]);
    }
    st_bf16x8(dy + m.b * a.dy_bs + (long long)l * a.dy_ld + m.c0, o);
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
// rows per CTA: enough CTAs for ~4 per SM, at least 2 rows per row-lane, at most 64 rows
static int rb_rows_per_cta(const of_rb_args* a) {
  const int rpar = kRbThreads / (a->C / 8) > 0 ? kRbThreads / (a->C / 8) : 1;
  long long want = ((long long)a->B * a->L + 4 * device_sm_count() - 1) / (4 * device_sm_count());
  int r = (int)want;
  if (r < 2 * rpar) r = 2 * rpar;
  if (r > 64) r = 64;
  return r;
}
static dim3 rb_grid(const of_rb_args* a, int rpc) { return dim3((a->L + rpc - 1) / rpc, a->B); }

}  // namespace ofx

using namespace ofx;

#define RB_LAUNCH(kernel)                                                                                                    \
  const int rpc = rb_rows_per_cta(a);                                                                                        \
  kernel<<<rb_grid(a, rpc), kRbThreads, 5 * (size_t)a->C * sizeof(float), reinterpret_cast<cudaStream_t>(stream)>>>(*a, rpc); \
  OF_CHECK_CUDA(cudaGetLastError());                                              \
  count_launch();                                                                 \
  return OF_OK;

extern "C" int of_rb_apply_fwd(const of_rb_args* a, void* stream) {
  int rc = check_common(a, "of_rb_apply_fwd");
  if (rc) return rc;
  OF_REQUIRE(a->out_bf16 && a->out_bf16_ld % 8 == 0, "of_rb_apply_fwd: bad out_bf16");
  RB_LAUNCH(rb_apply_fwd_kernel)
}
extern "C" int of_rb_rowdot(const of_rb_args* a, void* stream) {
  int rc = check_common(a, "of_rb_rowdot");
  if (rc) return rc;
  OF_REQUIRE(a->vec && a->out_rows, "of_rb_rowdot: null vec/out_rows");
  const int rpc = rb_rows_per_cta(a);
  rb_rowdot_kernel<<<rb_grid(a, rpc), kRbThreads, (size_t)(rpc > 5 * a->C ? rpc : 5 * a->C) * sizeof(float),
                     reinterpret_cast<cudaStream_t>(stream)>>>(*a, rpc);
  OF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OF_OK;
}
extern "C" int of_softmax_rows(float* rows, int B, int L, void* stream) {
  OF_REQUIRE(rows && B >= 1 && L >= 1, "of_softmax_rows: bad args");
  softmax_rows_kernel<<<B, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(rows, L);
  OF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OF_OK;
}
extern "C" int of_softmax_bwd_rows(const float* p, float* rd, int B, int L, void* stream) {
  OF_REQUIRE(p && rd && B >= 1 && L >= 1, "of_softmax_bwd_rows: bad args");
  softmax_bwd_rows_kernel<<<B, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, rd, L);
  OF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OF_OK;
}
extern "C" int of_rb_pool(const of_rb_args* a, void* stream) {
  int rc = check_common(a, "of_rb_pool");
  if (rc) return rc;
  OF_REQUIRE(a->p && a->acc_bc, "of_rb_pool: null p/acc_bc");
  RB_LAUNCH(rb_pool_kernel)
}
extern "C" int of_rb_gate_fwd(const of_rb_args* a, void* stream) {
  int rc = check_common(a, "of_rb_gate_fwd");
  if (rc) return rc;
  OF_REQUIRE(a->gate && (a->res_f32 || a->res_bf16) && (a->out_f32 || a->out_bf16), "of_rb_gate_fwd: null gate/res/out");
  RB_LAUNCH(rb_gate_fwd_kernel)
}
extern "C" int of_rb_gate_bwd_reduce(const of_rb_args* a, void* stream) {
  int rc = check_common(a, "of_rb_gate_bwd_reduce");
  if (rc) return rc;
  OF_REQUIRE(a->dout_f32 && a->acc_bc, "of_rb_gate_bwd_reduce: null dout/acc");
  RB_LAUNCH(rb_gate_bwd_reduce_kernel)
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
    RB_LAUNCH(rb_bwd_pass1_kernel<0>)
  }
  RB_LAUNCH(rb_bwd_pass1_kernel<1>)
}
extern "C" int of_rb_bwd_apply(const of_rb_args* a, void* stream) {
  int rc = check_common(a, "of_rb_bwd_apply");
  if (rc) return rc;
  OF_REQUIRE(a->dxhat_bf16 && a->dstats && a->dy_bf16, "of_rb_bwd_apply: null pointers");
  RB_LAUNCH(rb_bwd_apply_kernel)
}
