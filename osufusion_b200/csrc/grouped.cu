// Grouped ("multi-tensor") bandwidth kernels: one launch over MANY parameter tensors instead of one tiny launch per tensor.
//
//   of_film_fwd / of_film_bwd : every FiLM head `ResidualBlock.mlp = Sequential(SiLU, Linear(2*dim_emb, 2*C))`
//                               (reference residual.py:104-111) of the denoiser at once.  Their common input is the (B, 2*dim_emb)
//                               conditioning vector, so all heads are known right after `time_mlp`/`cond_mlp` have run: 35 heads,
//                               1.09 GB of fp32 weights at CFG-L, streamed once at HBM speed instead of 35 latency-bound launches.
//   of_pack_weights           : fp32 master weights (reference state_dict layout) -> bf16 GEMM operand layout for every
//                               Conv1d / Linear of the denoiser in one launch (what autocast's per-op weight cast does in the
//                               reference, trainer.py:295).
//
// Descriptor tables live in device memory and are built once per model by the host (osufusion_b200/engine.py).
#include "host_common.h"
#include "rowops.cuh"

namespace ofx {

constexpr int kFilmMaxM = 16;

__device__ __forceinline__ int find_group_by_row(const of_film_group* __restrict__ g, int n, int row) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (g[mid].row_start <= row) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

// warp = 4 consecutive output rows of one head; lane strides over K with float4; acc[4][MM] in registers.
template <int MM>
__global__ void __launch_bounds__(256) film_fwd_kernel(const of_film_group* __restrict__ groups, int num_groups, int total_rows,
                                                       const float* __restrict__ x, int M, int K, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = (blockIdx.x * 8 + warp) * 4;
  if (row0 >= total_rows) return;
  const int gi = find_group_by_row(groups, num_groups, row0);
  const of_film_group g = groups[gi];
  const int n0 = row0 - g.row_start;
  const float* w = g.W + (long long)n0 * K;
  float acc[4][MM];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int m = 0; m < MM; ++m) acc[r][m] = 0.f;
#pragma unroll 2
  for (int k = lane * 4; k < K; k += 128) {
    float4 wv[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) wv[r] = *reinterpret_cast<const float4*>(w + (long long)r * K + k);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      wv[r].x = bf16_round(wv[r].x); wv[r].y = bf16_round(wv[r].y); wv[r].z = bf16_round(wv[r].z); wv[r].w = bf16_round(wv[r].w);
    }
#pragma unroll
    for (int m = 0; m < MM; ++m) {
      if (m < M) {
        const float4 xv = *reinterpret_cast<const float4*>(x + (long long)m * K + k);
#pragma unroll
        for (int r = 0; r < 4; ++r)
          acc[r][m] = fmaf(wv[r].x, xv.x, fmaf(wv[r].y, xv.y, fmaf(wv[r].z, xv.z, fmaf(wv[r].w, xv.w, acc[r][m]))));
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
#pragma unroll
    for (int m = 0; m < MM; ++m) {
      if (m < M) {
        float v = warp_sum(acc[r][m]);
        if (lane == 0) {
          v += g.bias ? g.bias[n0 + r] : 0.f;
          out[g.out_off + (long long)m * g.N + n0 + r] = bf16_round(v);
        }
      }
    }
  }
}

// CTA = (1024-wide k slice, chunk of <= kFilmChunk rows of one head).  thread owns 4 consecutive k:
//   dW[n, k] = sum_m dss[m, n] * x[m, k]   (plain store: every element has exactly one writer)
//   d_emb[m, k] += sum_n dss[m, n] * bf16(W[n, k])   (register partials over the chunk, one atomic per element per CTA)
constexpr int kFilmChunk = 128;
template <int MM>
__global__ void __launch_bounds__(256) film_bwd_kernel(const of_film_group* __restrict__ groups, const int2* __restrict__ chunks,
                                                       const float* __restrict__ dss, const float* __restrict__ x, int M, int K,
                                                       float* __restrict__ d_emb) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sd[kFilmChunk][MM];
  const int2 ch = chunks[blockIdx.y];
  const of_film_group g = groups[ch.x];
  const int n0 = ch.y;
  const int nn = min(kFilmChunk, g.N - n0);
  for (int i = threadIdx.x; i < kFilmChunk * MM; i += blockDim.x) {
    const int j = i / MM, m = i - j * MM;
    sd[j][m] = (j < nn && m < M) ? dss[g.out_off + (long long)m * g.N + n0 + j] : 0.f;
  }
  __syncthreads();
  if (g.dbias && blockIdx.x == 0 && threadIdx.x < nn) {
    float s = 0.f;
#pragma unroll
    for (int m = 0; m < MM; ++m) s += sd[threadIdx.x][m];
    g.dbias[n0 + threadIdx.x] = s;
  }
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (k >= K) return;
  float4 xa[MM], dxa[MM];
#pragma unroll
  for (int m = 0; m < MM; ++m) {
    xa[m] = (m < M) ? *reinterpret_cast<const float4*>(x + (long long)m * K + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    dxa[m] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float* w = g.W + (long long)n0 * K + k;
  float* dw = g.dW ? g.dW + (long long)n0 * K + k : nullptr;
#pragma unroll 4
  for (int j = 0; j < nn; ++j) {
    float4 wv = *reinterpret_cast<const float4*>(w + (long long)j * K);
    wv.x = bf16_round(wv.x); wv.y = bf16_round(wv.y); wv.z = bf16_round(wv.z); wv.w = bf16_round(wv.w);
    float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int m = 0; m < MM; ++m) {
      const float d = sd[j][m];
      gv.x = fmaf(d, xa[m].x, gv.x); gv.y = fmaf(d, xa[m].y, gv.y); gv.z = fmaf(d, xa[m].z, gv.z); gv.w = fmaf(d, xa[m].w, gv.w);
      dxa[m].x = fmaf(d, wv.x, dxa[m].x); dxa[m].y = fmaf(d, wv.y, dxa[m].y);
      dxa[m].z = fmaf(d, wv.z, dxa[m].z); dxa[m].w = fmaf(d, wv.w, dxa[m].w);
    }
    if (dw) *reinterpret_cast<float4*>(dw + (long long)j * K) = gv;
  }
  if (d_emb) {
#pragma unroll
    for (int m = 0; m < MM; ++m) {
      if (m < M) {
        float* d = d_emb + (long long)m * K + k;
        atomicAdd(d, dxa[m].x); atomicAdd(d + 1, dxa[m].y); atomicAdd(d + 2, dxa[m].z); atomicAdd(d + 3, dxa[m].w);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ weight packing
constexpr int kPackFlat = 4096;   // elements per CTA of a flat cast segment
constexpr int kPackCi = 1024;     // input channels per CTA of a conv segment (k <= 4)

__global__ void __launch_bounds__(256) pack_weights_kernel(const of_pack_seg* __restrict__ segs, int num_segs) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_w[kPackCi * 4];
  int lo = 0, hi = num_segs - 1;
  const int cta = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (segs[mid].cta_begin <= cta) lo = mid;
    else hi = mid - 1;
  }
  const of_pack_seg sg = segs[lo];
  const int item = cta - sg.cta_begin;
  const float sc = sg.scale != 0.f ? sg.scale : 1.0f;      // 0 = plain cast (x * 1.0f is exact, so one code path serves both)
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(sg.dst);
  if (sg.k == 1 && sg.cin_pad == sg.Cin) {
    const long long n = (long long)sg.Cout * sg.Cin;
    const long long base = (long long)item * kPackFlat;
    if ((n & 3) == 0) {
      float4 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long e = base + (i * 256 + threadIdx.x) * 4;
        if (e < n) v[i] = *reinterpret_cast<const float4*>(sg.src + e);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long e = base + (i * 256 + threadIdx.x) * 4;
        if (e < n) *reinterpret_cast<uint2*>(dst + e) = make_uint2(pack_bf16x2(sc * v[i].x, sc * v[i].y), pack_bf16x2(sc * v[i].z, sc * v[i].w));
      }
    } else {
      for (long long e = base + threadIdx.x; e < min(base + kPackFlat, n); e += 256) dst[e] = __float2bfloat16_rn(sc * sg.src[e]);
    }
    return;
  }
  // conv: item = (co, ci chunk); the (ci, t) slab of one co is contiguous in the torch layout, each tap row in the packed one
  const int chunks = (sg.cin_pad + kPackCi - 1) / kPackCi;
  const int co = item / chunks;
  const int ci0 = (item - co * chunks) * kPackCi;
  const int nci = min(kPackCi, sg.cin_pad - ci0);
  const int nreal = max(0, min(kPackCi, sg.Cin - ci0));
  const int k = sg.k;
  const float* src = sg.src + ((long long)co * sg.Cin + ci0) * k;
  for (int i = threadIdx.x; i < nreal * k; i += blockDim.x) s_w[i] = src[i];
  __syncthreads();
  for (int i = threadIdx.x; i < nci * k; i += blockDim.x) {
    const int t = i / nci, ci = i - t * nci;
    const float v = ci < nreal ? sc * s_w[ci * k + t] : 0.f;
    dst[((long long)t * sg.Cout + co) * sg.cin_pad + ci0 + ci] = __float2bfloat16_rn(v);
  }
}

}  // namespace ofx

using namespace ofx;

#define STREAM reinterpret_cast<cudaStream_t>(stream)

extern "C" int of_film_fwd(const of_film_group* groups_dev, int num_groups, int total_rows, const float* x, int M, int K, float* out,
                           void* stream) {
  OF_REQUIRE(groups_dev && x && out, "of_film_fwd: null pointer");
  OF_REQUIRE(M >= 1 && M <= kFilmMaxM, "of_film_fwd: M=%d out of range (1..%d)", M, kFilmMaxM);
  OF_REQUIRE(K % 4 == 0 && total_rows % 4 == 0, "of_film_fwd: K=%d and every head's row count must be multiples of 4", K);
  const int grid = (total_rows + 31) / 32;
  if (M <= 4) OF_CHECK_CUDA(launch_pdl(film_fwd_kernel<4>, dim3(grid), dim3(256), 0, STREAM, groups_dev, num_groups, total_rows, x, M, K, out));
  else if (M <= 8) OF_CHECK_CUDA(launch_pdl(film_fwd_kernel<8>, dim3(grid), dim3(256), 0, STREAM, groups_dev, num_groups, total_rows, x, M, K, out));
  else OF_CHECK_CUDA(launch_pdl(film_fwd_kernel<16>, dim3(grid), dim3(256), 0, STREAM, groups_dev, num_groups, total_rows, x, M, K, out));
  OF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OF_OK;
}

extern "C" int of_film_bwd(const of_film_group* groups_dev, const int* chunks_dev, int num_chunks, const float* dss, const float* x,
                           int M, int K, float* d_emb, void* stream) {
  OF_REQUIRE(groups_dev && chunks_dev && dss && x, "of_film_bwd: null pointer");
  OF_REQUIRE(M >= 1 && M <= kFilmMaxM, "of_film_bwd: M=%d out of range (1..%d)", M, kFilmMaxM);
  OF_REQUIRE(K % 4 == 0, "of_film_bwd: K=%d must be a multiple of 4", K);
  dim3 grid((K / 4 + 255) / 256, num_chunks);
  const int2* ch = reinterpret_cast<const int2*>(chunks_dev);
  if (M <= 4) OF_CHECK_CUDA(launch_pdl(film_bwd_kernel<4>, dim3(grid), dim3(256), 0, STREAM, groups_dev, ch, dss, x, M, K, d_emb));
  else if (M <= 8) OF_CHECK_CUDA(launch_pdl(film_bwd_kernel<8>, dim3(grid), dim3(256), 0, STREAM, groups_dev, ch, dss, x, M, K, d_emb));
  else OF_CHECK_CUDA(launch_pdl(film_bwd_kernel<16>, dim3(grid), dim3(256), 0, STREAM, groups_dev, ch, dss, x, M, K, d_emb));
  OF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OF_OK;
}

extern "C" int of_film_chunk_rows(void) { return kFilmChunk; }

extern "C" int of_pack_weights(const of_pack_seg* segs_dev, int num_segs, int total_ctas, void* stream) {
  OF_REQUIRE(segs_dev && num_segs >= 1 && total_ctas >= 1, "of_pack_weights: bad args");
  OF_CHECK_CUDA(launch_pdl(pack_weights_kernel, dim3(total_ctas), dim3(256), 0, STREAM, segs_dev, num_segs));
  OF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OF_OK;
}

// CTAs a segment needs (the host builds `cta_begin` as the running sum of this).
extern "C" int of_pack_seg_ctas(int Cout, int Cin, int k, int cin_pad) {
  if (k == 1 && cin_pad == Cin) return (int)(((long long)Cout * Cin + kPackFlat - 1) / kPackFlat);
  if (k > 4) return -1;
  return Cout * ((cin_pad + kPackCi - 1) / kPackCi);
}

// ------------------------------------------------------------------------------------------------ fused optimizer step
// (`accelerator.clip_grad_norm_(params, 1.0)` + `torch.optim.AdamW.step()` of trainer.py:302-309 without the
//  per-tensor launches and the two host synchronisations.)  Gradients, exp_avg and exp_avg_sq are flat arenas with identical
//  layout (the engine's gradient arena); only the parameters are separate tensors.
namespace ofx {

// sum of squares of the whole arena (padding between tensors is zero) -> out[0] (double, atomically accumulated)
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sm[32];
  float s = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      const float4 v = *reinterpret_cast<const float4*>(g + i);
      s = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s))));
    } else {
      for (long long j = i; j < n; ++j) s = fmaf(g[j], g[j], s);
    }
  }
  s = block_sum(s, sm);
  if (threadIdx.x == 0) atomicAdd(out, (double)s);
}

constexpr int kOptChunk = 4096;   // elements per CTA (flat tensors)
constexpr int kOptPairs = 1024;   // (co, ci) pairs per CTA (packed conv tensors, k <= 4): 256 threads x 4

// torch.optim.AdamW (amsgrad=False, maximize=False) on every tensor of the table; gradients are scaled by
// clip = min(1, max_norm / (sqrt(sumsq) + 1e-6)) read from device memory (torch.nn.utils.clip_grad_norm_ semantics), without
// being modified in place.
__global__ void __launch_bounds__(256) adamw_kernel(const of_opt_tensor* __restrict__ tab, int num_tensors, const float* __restrict__ grads,
                                                    float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                    const double* __restrict__ sumsq, float max_norm, float lr, float beta1, float beta2,
                                                    float eps, float weight_decay, float bias_corr1, float bias_corr2_sqrt) {
  pdl_launch_dependents();
  pdl_wait();
  int lo = 0, hi = num_tensors - 1;
  const int cta = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab[mid].cta_begin <= cta) lo = mid;
    else hi = mid - 1;
  }
  const of_opt_tensor t = tab[lo];
  float clip = 1.0f;
  if (sumsq != nullptr && max_norm > 0.f) {
    const float nrm = (float)sqrt(*sumsq);
    clip = fminf(1.0f, max_norm / (nrm + 1e-6f));
  }
  const float step_size = lr / bias_corr1;
  const float decay = 1.0f - lr * weight_decay;
  __nv_bfloat16* op16 = reinterpret_cast<__nv_bfloat16*>(t.operand_bf16);
  if (t.k > 1) {
    // packed conv tensor: gradient / moments / operand are [k][Cout*Cin] planes; the parameter is (Cout*Cin, k) (torch layout).
    // With q = co*Cin + ci the two layouts are plane[tap][q] and param[q*k + tap]: a CTA takes kOptPairs consecutive q — a contiguous
    // parameter slab staged through shared memory, and k contiguous runs of the planes (16-byte vectors, all taps in flight).
    __shared__ float sp[kOptPairs * 4];
    const long long pairs = (long long)t.Cout * t.Cin;
    const long long q0 = (long long)(cta - t.cta_begin) * kOptPairs;
    const int nq = (int)min((long long)kOptPairs, pairs - q0);
    float* pslab = t.param + q0 * t.k;
    for (int i = threadIdx.x * 4; i < nq * t.k; i += blockDim.x * 4) {      // nq*k is a multiple of 4 (Cin % 8 == 0)
      *reinterpret_cast<float4*>(sp + i) = *reinterpret_cast<const float4*>(pslab + i);
    }
    __syncthreads();
    const int q = threadIdx.x * 4;
    if (q < nq) {
      float4 g4[4], m4[4], v4[4];
#pragma unroll
      for (int tap = 0; tap < 4; ++tap) {
        if (tap < t.k) {
          const long long a = t.arena_off + (long long)tap * pairs + q0 + q;
          g4[tap] = *reinterpret_cast<const float4*>(grads + a);
          m4[tap] = *reinterpret_cast<const float4*>(exp_avg + a);
          v4[tap] = *reinterpret_cast<const float4*>(exp_avg_sq + a);
        }
      }
#pragma unroll
      for (int tap = 0; tap < 4; ++tap) {
        if (tap < t.k) {
          const long long a = t.arena_off + (long long)tap * pairs + q0 + q;
          float gg[4] = {g4[tap].x, g4[tap].y, g4[tap].z, g4[tap].w}, mm[4] = {m4[tap].x, m4[tap].y, m4[tap].z, m4[tap].w},
                vv[4] = {v4[tap].x, v4[tap].y, v4[tap].z, v4[tap].w}, pp[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float g = gg[j] * clip;
            float pv = sp[(q + j) * t.k + tap] * decay;
            mm[j] = beta1 * mm[j] + (1.0f - beta1) * g;
            vv[j] = beta2 * vv[j] + (1.0f - beta2) * g * g;
            pv -= step_size * mm[j] / (sqrtf(vv[j]) / bias_corr2_sqrt + eps);
            sp[(q + j) * t.k + tap] = pv;
            pp[j] = pv;
          }
          *reinterpret_cast<float4*>(exp_avg + a) = make_float4(mm[0], mm[1], mm[2], mm[3]);
          *reinterpret_cast<float4*>(exp_avg_sq + a) = make_float4(vv[0], vv[1], vv[2], vv[3]);
          if (op16) {
            uint2 o;
            o.x = pack_bf16x2(pp[0], pp[1]);
            o.y = pack_bf16x2(pp[2], pp[3]);
            *reinterpret_cast<uint2*>(op16 + (long long)tap * pairs + q0 + q) = o;
          }
        }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x * 4; i < nq * t.k; i += blockDim.x * 4) {
      *reinterpret_cast<float4*>(pslab + i) = *reinterpret_cast<const float4*>(sp + i);
    }
    return;
  }
  const long long base = (long long)(cta - t.cta_begin) * kOptChunk;
#pragma unroll
  for (int it = 0; it < kOptChunk / 1024; ++it) {
    const long long e = base + (it * 256 + threadIdx.x) * 4;
    if (e >= t.numel) break;
    const long long a = t.arena_off + e;
    if (e + 3 < t.numel && (t.numel & 3) == 0) {
      float4 p = *reinterpret_cast<const float4*>(t.param + e);
      const float4 g4 = *reinterpret_cast<const float4*>(grads + a);
      float4 m = *reinterpret_cast<const float4*>(exp_avg + a), v = *reinterpret_cast<const float4*>(exp_avg_sq + a);
      float pp[4] = {p.x, p.y, p.z, p.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w}, mm[4] = {m.x, m.y, m.z, m.w}, vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float g = gg[j] * clip;
        pp[j] *= decay;
        mm[j] = beta1 * mm[j] + (1.0f - beta1) * g;
        vv[j] = beta2 * vv[j] + (1.0f - beta2) * g * g;
        pp[j] -= step_size * mm[j] / (sqrtf(vv[j]) / bias_corr2_sqrt + eps);
      }
      *reinterpret_cast<float4*>(t.param + e) = make_float4(pp[0], pp[1], pp[2], pp[3]);
      *reinterpret_cast<float4*>(exp_avg + a) = make_float4(mm[0], mm[1], mm[2], mm[3]);
      *reinterpret_cast<float4*>(exp_avg_sq + a) = make_float4(vv[0], vv[1], vv[2], vv[3]);
      if (op16) {
        uint2 o;
        o.x = pack_bf16x2(pp[0], pp[1]);
        o.y = pack_bf16x2(pp[2], pp[3]);
        *reinterpret_cast<uint2*>(op16 + e) = o;
      }
    } else {
      for (long long j = e; j < min(e + 4, t.numel); ++j) {
        const float g = grads[t.arena_off + j] * clip;
        float p = t.param[j] * decay;
        const float m = beta1 * exp_avg[t.arena_off + j] + (1.0f - beta1) * g;
        const float v = beta2 * exp_avg_sq[t.arena_off + j] + (1.0f - beta2) * g * g;
        p -= step_size * m / (sqrtf(v) / bias_corr2_sqrt + eps);
        t.param[j] = p;
        exp_avg[t.arena_off + j] = m;
        exp_avg_sq[t.arena_off + j] = v;
        if (op16) op16[j] = __float2bfloat16_rn(p);
      }
    }
  }
}

}  // namespace ofx

extern "C" int of_grad_sumsq(const float* grads, long long n, double* out, void* stream) {
  OF_REQUIRE(grads && out && n >= 0, "of_grad_sumsq: bad args");
  OF_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(double), STREAM));
  if (n > 0) {
    long long want = (n / 4 + 255) / 256;
    const int grid = (int)(want < 8LL * device_sm_count() ? (want < 1 ? 1 : want) : 8LL * device_sm_count());
    OF_CHECK_CUDA(launch_pdl(sumsq_kernel, dim3(grid), dim3(256), 0, STREAM, grads, n, out));
    OF_CHECK_CUDA(cudaGetLastError());
  }
  count_launch();
  return OF_OK;
}

extern "C" int of_opt_tensor_ctas(long long numel) { return (int)((numel + kOptChunk - 1) / kOptChunk); }
extern "C" int of_opt_tensor_ctas2(long long numel, int Cout, int Cin, int k) {
  if (k <= 1) return of_opt_tensor_ctas(numel);
  if (k > 4 || (long long)Cout * Cin * k != numel || Cin % 8 != 0) return -1;
  return (int)(((long long)Cout * Cin + kOptPairs - 1) / kOptPairs);
}

extern "C" int of_adamw_step(const of_opt_tensor* table_dev, int num_tensors, int total_ctas, const float* grads, float* exp_avg,
                             float* exp_avg_sq, const double* sumsq, float max_norm, float lr, float beta1, float beta2, float eps,
                             float weight_decay, int step, void* stream) {
  OF_REQUIRE(table_dev && grads && exp_avg && exp_avg_sq && num_tensors >= 1 && total_ctas >= 1 && step >= 1, "of_adamw_step: bad args");
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2s = sqrtf(1.0f - powf(beta2, (float)step));
  OF_CHECK_CUDA(launch_pdl(adamw_kernel, dim3(total_ctas), dim3(256), 0, STREAM, table_dev, num_tensors, grads, exp_avg, exp_avg_sq, sumsq, max_norm, lr, beta1, beta2, eps,
                                               weight_decay, bc1, bc2s));
  OF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OF_OK;
}
