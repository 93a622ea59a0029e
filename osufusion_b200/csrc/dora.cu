// LoRA / DoRA adapter kernels (reference: osu_fusion/modules/lora_layers.py:15-96,292-328 for Conv1d; peft 0.12.0
// DoraLinearLayer for nn.Linear).  The adapted layer is evaluated as ONE GEMM with the effective weight
//     W_eff[co, e] = s[co] * (W[co, e] + scaling * sum_r B[co, r] * A[r, e]),   s = magnitude / ||W + scaling*B*A||_2 (detached)
// (e runs over (ci, tap)), which is algebraically the reference's  base(x) + (s-1)*conv(x, W) + s*scaling*B(A(x)).
// Backward: the weight-gradient GEMM produces dW_eff; of_dora_grad projects it onto dA, dB and d(magnitude).
#include "host_common.h"
#include "rowops.cuh"

namespace ofx {

constexpr int kDoraE = 256;   // e-columns per CTA (one per thread)
constexpr int kDoraCo = 16;   // output channels per CTA tile
constexpr int kDoraEP = kDoraE + 4;   // padded row stride of the A / G tiles in shared memory (float4 reads down a column of rows)

struct DoraArgs {
  const float* W;   // (Cout, E)   E = Cin*k, torch layout (ci-major, tap-minor)
  const float* A;   // (r, E)
  const float* B;   // (Cout, r)
  const float* mag; // (Cout) or nullptr (plain LoRA: s = 1)
  float scaling;
  int Cout, Cin, k, r, Cin_pad;
};

constexpr int kDoraRMax = 128;   // rank limit (register tile of the A column is processed in chunks of 32)

// d[c] = sum_r B[co0+c, r] * A[r, e] for the thread's column e and all kDoraCo rows of the CTA: the A column is held in
// registers 32 ranks at a time and B is read as broadcast float4 (0.25 shared loads per FMA instead of 2).
__device__ __forceinline__ void dora_ba_column(const DoraArgs& a, const float* sA, const float* sB, float (&d)[kDoraCo]) {
#pragma unroll
  for (int c = 0; c < kDoraCo; ++c) d[c] = 0.f;
  for (int r0 = 0; r0 < a.r; r0 += 32) {
    float av[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) av[r] = (r0 + r < a.r) ? sA[(r0 + r) * kDoraEP + threadIdx.x] : 0.f;
    if ((a.r & 3) == 0) {
#pragma unroll
      for (int c = 0; c < kDoraCo; ++c) {
#pragma unroll
        for (int r4 = 0; r4 < 8; ++r4) {
          if (r0 + r4 * 4 < a.r) {
            const float4 b = *reinterpret_cast<const float4*>(sB + c * a.r + r0 + r4 * 4);
            d[c] = fmaf(b.x, av[r4 * 4], fmaf(b.y, av[r4 * 4 + 1], fmaf(b.z, av[r4 * 4 + 2], fmaf(b.w, av[r4 * 4 + 3], d[c]))));
          }
        }
      }
    } else {
#pragma unroll
      for (int c = 0; c < kDoraCo; ++c)
#pragma unroll
        for (int r = 0; r < 32; ++r)
          if (r0 + r < a.r) d[c] = fmaf(sB[c * a.r + r0 + r], av[r], d[c]);
    }
  }
}

__device__ __forceinline__ void dora_load_a_tile(const DoraArgs& a, int e0, float* sA) {
  const int E = a.Cin * a.k;
  const int e = e0 + threadIdx.x;            // blockDim.x == kDoraE: thread = column, loop = rank (16 loads in flight)
  for (int r0 = 0; r0 < a.r; r0 += 16) {
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = (r0 + j < a.r && e < E) ? a.A[(long long)(r0 + j) * E + e] : 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (r0 + j < a.r) sA[(r0 + j) * kDoraEP + threadIdx.x] = v[j];
  }
}
// B rows of the CTA's whole channel range [co_begin, co_begin + rows_pad) -> sB[rows_pad][r] (rows >= co_end are zero): loaded
// once, so the tile loop needs no barrier and no global-load latency for B.
__device__ __forceinline__ void dora_load_b_range(const DoraArgs& a, int co_begin, int co_end, int rows_pad, float* sB) {
  const int n = rows_pad * a.r;
  for (int i0 = threadIdx.x; i0 < n; i0 += 4 * blockDim.x) {
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = i0 + j * blockDim.x;
      const int c = i / a.r;
      v[j] = (i < n && co_begin + c < co_end) ? a.B[(long long)(co_begin + c) * a.r + (i - c * a.r)] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (i0 + j * blockDim.x < n) sB[i0 + j * blockDim.x] = v[j];
  }
}

// CTA = one 256-column e-tile x a range of output channels walked 16 at a time (the A tile is loaded once per CTA).
// n2[co] += sum_e (W + scaling*BA)^2
__global__ void __launch_bounds__(kDoraE) dora_norm_kernel(const DoraArgs a, float* __restrict__ n2, int co_per_cta) {
  extern __shared__ float sm[];
  float* sA = sm;
  float* sBall = sA + a.r * kDoraEP;              // [co_per_cta][r]
  float* sN = sBall + co_per_cta * a.r;           // [co_per_cta]
  const int E = a.Cin * a.k;
  const int e0 = blockIdx.x * kDoraE;
  const int co_begin = blockIdx.y * co_per_cta, co_end = min(co_begin + co_per_cta, a.Cout);
  const int e = e0 + threadIdx.x;
  dora_load_a_tile(a, e0, sA);
  dora_load_b_range(a, co_begin, co_end, co_per_cta, sBall);
  for (int i = threadIdx.x; i < co_per_cta; i += blockDim.x) sN[i] = 0.f;
  float w[kDoraCo], wn[kDoraCo];
#pragma unroll
  for (int c = 0; c < kDoraCo; ++c) wn[c] = (co_begin + c < co_end && e < E) ? a.W[(long long)(co_begin + c) * E + e] : 0.f;
  __syncthreads();
  for (int co0 = co_begin; co0 < co_end; co0 += kDoraCo) {
    float d[kDoraCo];
#pragma unroll
    for (int c = 0; c < kDoraCo; ++c) {
      w[c] = wn[c];                            // the next tile's rows are fetched while this tile is computed
      wn[c] = (co0 + kDoraCo + c < co_end && e < E) ? a.W[(long long)(co0 + kDoraCo + c) * E + e] : 0.f;
    }
    dora_ba_column(a, sA, sBall + (co0 - co_begin) * a.r, d);
#pragma unroll
    for (int c = 0; c < kDoraCo; ++c) {
      float sq = 0.f;
      if (co0 + c < co_end && e < E) {
        const float v = w[c] + a.scaling * d[c];
        sq = v * v;
      }
      sq = warp_sum(sq);
      if ((threadIdx.x & 31) == 0 && co0 + c < co_end) atomicAdd(&sN[co0 - co_begin + c], sq);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < co_end - co_begin; i += blockDim.x) atomicAdd(n2 + co_begin + i, sN[i]);
}

// packed[t][co][ci] = bf16(s[co] * (W + scaling*BA)[co, ci, t]);  s_out[co] = s[co]
__global__ void __launch_bounds__(kDoraE) dora_merge_kernel(const DoraArgs a, const float* __restrict__ n2,
                                                            __nv_bfloat16* __restrict__ packed, long long tap_stride,
                                                            float* __restrict__ s_out, int co_per_cta) {
  extern __shared__ float sm[];
  float* sA = sm;
  float* sB = sA + a.r * kDoraEP;
  const int E = a.Cin * a.k;
  const int e0 = blockIdx.x * kDoraE;
  const int co_begin = blockIdx.y * co_per_cta, co_end = min(co_begin + co_per_cta, a.Cout);
  const int e = e0 + threadIdx.x;
  const int ci = e / a.k, t = e - ci * a.k;
  dora_load_a_tile(a, e0, sA);
  dora_load_b_range(a, co_begin, co_end, co_per_cta, sB);
  float w[kDoraCo], wn[kDoraCo];
#pragma unroll
  for (int c = 0; c < kDoraCo; ++c) wn[c] = (co_begin + c < co_end && e < E) ? a.W[(long long)(co_begin + c) * E + e] : 0.f;
  __syncthreads();
  for (int co0 = co_begin; co0 < co_end; co0 += kDoraCo) {
    float d[kDoraCo];
#pragma unroll
    for (int c = 0; c < kDoraCo; ++c) {
      w[c] = wn[c];
      wn[c] = (co0 + kDoraCo + c < co_end && e < E) ? a.W[(long long)(co0 + kDoraCo + c) * E + e] : 0.f;
    }
    dora_ba_column(a, sA, sB + (co0 - co_begin) * a.r, d);
#pragma unroll
    for (int c = 0; c < kDoraCo; ++c) {
      const int co = co0 + c;
      if (co < co_end) {
        const float s = a.mag ? a.mag[co] * rsqrtf(n2[co]) : 1.0f;
        if (blockIdx.x == 0 && threadIdx.x == 0 && s_out) s_out[co] = s;
        if (e < E) {
          const float v = w[c] + a.scaling * d[c];
          packed[(long long)t * tap_stride + (long long)co * a.Cin_pad + ci] = __float2bfloat16_rn(s * v);
        }
      }
    }
  }
}

// From dWp = d(loss)/d(W_eff) in the packed [t][co][ci] fp32 layout:
//   dB[co, r] += scaling * s[co] * sum_e dWp[co,e] * A[r,e]
//   dA[r, e]  += scaling * sum_co B[co,r] * s[co] * dWp[co,e]
//   dmag[co]  += sum_e dWp[co,e] * (W + scaling*BA)[co,e] / n[co]          (n = ||W + scaling*BA||, detached)
// CTA = one 256-column e-tile x a RANGE of output channels (co_per_cta, walked 16 at a time): the dA partial of the whole range
// stays in registers (r <= 32: 32 accumulators per thread) and is added to global memory once per CTA -- the first version issued
// r*256 global atomics per 16-channel tile, ~2 atomics per weight element, and was bound by L2 atomic throughput.
__global__ void __launch_bounds__(kDoraE) dora_grad_kernel(const DoraArgs a, const float* __restrict__ n2,
                                                           const float* __restrict__ dWp, long long tap_stride,
                                                           float* __restrict__ dA, float* __restrict__ dB,
                                                           float* __restrict__ dmag, int co_per_cta) {
  extern __shared__ float sm[];
  float* sA = sm;                           // [r][kDoraEP]
  float* sBall = sA + a.r * kDoraEP;        // [co_per_cta][r]
  float* sG = sBall + co_per_cta * a.r;     // [kDoraCo][kDoraEP]  G = scaling * s * dWp
  float* sM = sG + kDoraCo * kDoraEP;       // [kDoraCo]
  const int E = a.Cin * a.k;
  const int e0 = blockIdx.x * kDoraE;
  const int co_begin = blockIdx.y * co_per_cta, co_end = min(co_begin + co_per_cta, a.Cout);
  const int e = e0 + threadIdx.x;
  const int ci = e / a.k, t = e - ci * a.k;
  dora_load_a_tile(a, e0, sA);
  dora_load_b_range(a, co_begin, co_end, co_per_cta, sBall);
  float dAacc[32];
#pragma unroll
  for (int r = 0; r < 32; ++r) dAacc[r] = 0.f;
  const bool fast_r = (a.r <= 32) && ((a.r & 3) == 0);
  for (int co0 = co_begin; co0 < co_end; co0 += kDoraCo) {
    __syncthreads();                        // previous tile's readers of sG / sM are done (also orders the sA / sBall fill)
    const float* sB = sBall + (co0 - co_begin) * a.r;
    if (threadIdx.x < kDoraCo) sM[threadIdx.x] = 0.f;
    float g[kDoraCo], w[kDoraCo], d[kDoraCo];
#pragma unroll
    for (int c = 0; c < kDoraCo; ++c) {
      const bool ok = co0 + c < co_end && e < E;
      g[c] = ok ? dWp[(long long)t * tap_stride + (long long)(co0 + c) * a.Cin_pad + ci] : 0.f;
      w[c] = (ok && a.mag) ? a.W[(long long)(co0 + c) * E + e] : 0.f;
    }
    __syncthreads();
    if (a.mag) dora_ba_column(a, sA, sB, d);
#pragma unroll
    for (int c = 0; c < kDoraCo; ++c) {
      const int co = co0 + c;
      float gv = 0.f, s = 1.f;
      if (co < co_end && a.mag) {
        const float inv_n = rsqrtf(n2[co]);
        s = a.mag[co] * inv_n;
        gv = g[c] * (w[c] + a.scaling * d[c]) * inv_n;
      }
      g[c] *= a.scaling * s;                       // G = scaling * s * dWp
      sG[c * kDoraEP + threadIdx.x] = g[c];
      if (a.mag) {
        gv = warp_sum(gv);
        if ((threadIdx.x & 31) == 0) atomicAdd(&sM[c], gv);
      }
    }
    // dA partial: thread e keeps its G column in registers, B read as broadcast float4 along r
    if (fast_r) {
#pragma unroll
      for (int r4 = 0; r4 < 8; ++r4) {
        if (r4 * 4 < a.r) {
#pragma unroll
          for (int c = 0; c < kDoraCo; ++c) {
            const float4 b = *reinterpret_cast<const float4*>(sB + c * a.r + r4 * 4);
            dAacc[r4 * 4] = fmaf(b.x, g[c], dAacc[r4 * 4]); dAacc[r4 * 4 + 1] = fmaf(b.y, g[c], dAacc[r4 * 4 + 1]);
            dAacc[r4 * 4 + 2] = fmaf(b.z, g[c], dAacc[r4 * 4 + 2]); dAacc[r4 * 4 + 3] = fmaf(b.w, g[c], dAacc[r4 * 4 + 3]);
          }
        }
      }
    } else if (e < E) {                       // general rank: per-tile atomics
      for (int r = 0; r < a.r; ++r) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < kDoraCo; ++c) acc = fmaf(sB[c * a.r + r], g[c], acc);
        atomicAdd(dA + (long long)r * E + e, acc);
      }
    }
    __syncthreads();
    // dB: kDoraCo x r outputs, each a dot product over the CTA's 256 e-columns.  thread = (channel pair cp, rank rr): a warp
    // shares its G rows (broadcast reads) and its lanes read 32 different A rows -- conflict-free thanks to the padded stride.
    for (int o = threadIdx.x; o < (kDoraCo / 2) * a.r; o += blockDim.x) {
      const int cp = o / a.r, rr = o - cp * a.r;
      float acc0 = 0.f, acc1 = 0.f;
      const float4* g0 = reinterpret_cast<const float4*>(sG + cp * kDoraEP);
      const float4* g1 = reinterpret_cast<const float4*>(sG + (cp + kDoraCo / 2) * kDoraEP);
      const float4* ap = reinterpret_cast<const float4*>(sA + rr * kDoraEP);
#pragma unroll 8
      for (int j = 0; j < kDoraE / 4; ++j) {
        const float4 y = ap[j], x0 = g0[j], x1 = g1[j];
        acc0 = fmaf(x0.x, y.x, fmaf(x0.y, y.y, fmaf(x0.z, y.z, fmaf(x0.w, y.w, acc0))));
        acc1 = fmaf(x1.x, y.x, fmaf(x1.y, y.y, fmaf(x1.z, y.z, fmaf(x1.w, y.w, acc1))));
      }
      if (co0 + cp < co_end) atomicAdd(dB + (long long)(co0 + cp) * a.r + rr, acc0);
      if (co0 + cp + kDoraCo / 2 < co_end) atomicAdd(dB + (long long)(co0 + cp + kDoraCo / 2) * a.r + rr, acc1);
    }
    if (a.mag && threadIdx.x < kDoraCo && co0 + threadIdx.x < co_end) atomicAdd(dmag + co0 + threadIdx.x, sM[threadIdx.x]);
  }
  if (fast_r && e < E) {
#pragma unroll
    for (int r = 0; r < 32; ++r)
      if (r < a.r) atomicAdd(dA + (long long)r * E + e, dAacc[r]);
  }
}

// ---- rank-r backward helpers (engine._adapter_backward_rank_r): the tiny per-module glue as two launches instead of ~10 torch ops
// rowscale[co] = scaling * s[co];  Bst[j][co] = bf16(rowscale[co] * B[co][j])   (the [N = r][K = Cout] operand of e = dy @ (s B))
__global__ void dora_rankr_prep_kernel(const float* __restrict__ B, const float* __restrict__ mag, const float* __restrict__ n2,
                                       float scaling, int Cout, int r, __nv_bfloat16* __restrict__ Bst, float* __restrict__ rowscale) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Cout * r) return;
  const int j = idx / Cout, co = idx - j * Cout;        // consecutive threads -> consecutive co: coalesced Bst stores
  const float rs = scaling * (mag ? mag[co] * rsqrtf(n2[co]) : 1.0f);
  Bst[(long long)j * Cout + co] = __float2bfloat16_rn(rs * B[(long long)co * r + j]);
  if (j == 0) rowscale[co] = rs;
}
// gB[co][j] += rowscale[co] * dBraw[co][j];  gmag[co] += dm[co] / mag[co]
__global__ void dora_rankr_finish_kernel(const float* __restrict__ dBraw, const float* __restrict__ rowscale, float* __restrict__ gB,
                                         const float* __restrict__ dm, const float* __restrict__ mag, float* __restrict__ gmag, int Cout,
                                         int r) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Cout * r) return;
  const int co = idx / r, j = idx - co * r;
  gB[idx] += rowscale[co] * dBraw[idx];
  if (j == 0 && mag) gmag[co] += dm[co] / mag[co];
}

// ---- tensor-core merge: V = W + scaling * B A comes from of_gemm (K = r, fp32 W as residual); this kernel turns one row of V
// into the GEMM operand: n2[co] = sum_e V^2, s = mag / sqrt(n2) (1 without magnitude), packed[t][co][ci] = bf16(s * V[co][ci*k + t]).
// One CTA per output channel; the row (<= 8192 elements) stays in registers between the norm and the store.
constexpr int kScalePackMaxPerThread = 32;
__global__ void __launch_bounds__(256) dora_scale_pack_kernel(const float* __restrict__ V, const float* __restrict__ mag, int Cin, int k,
                                                              float* __restrict__ n2_out, __nv_bfloat16* __restrict__ packed,
                                                              int cin_pad, long long tap_stride, const float* __restrict__ B,
                                                              float scaling, int r, __nv_bfloat16* __restrict__ Bst,
                                                              float* __restrict__ rowscale) {
  __shared__ float sm[32];
  const int co = blockIdx.x;
  const int E = Cin * k;
  const float* row = V + (long long)co * E;
  float v[kScalePackMaxPerThread];
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kScalePackMaxPerThread; ++i) {
    const int e = i * 256 + threadIdx.x;
    v[i] = e < E ? row[e] : 0.f;
    sq = fmaf(v[i], v[i], sq);
  }
  sq = block_sum(sq, sm);
  if (threadIdx.x == 0 && n2_out) n2_out[co] = sq;
  const float s = mag ? mag[co] * rsqrtf(sq) : 1.0f;
  if (rowscale) {      // operands of the rank-r backward (what of_dora_rankr_prep produces)
    const float rs = scaling * s;
    if (threadIdx.x == 0) rowscale[co] = rs;
    for (int j = threadIdx.x; j < r; j += blockDim.x) Bst[(long long)j * gridDim.x + co] = __float2bfloat16_rn(rs * B[(long long)co * r + j]);
  }
#pragma unroll
  for (int i = 0; i < kScalePackMaxPerThread; ++i) {
    const int e = i * 256 + threadIdx.x;
    if (e < E) {
      const int ci = e / k, t = e - ci * k;
      packed[(long long)t * tap_stride + (long long)co * cin_pad + ci] = __float2bfloat16_rn(s * v[i]);
    }
  }
}

// every adapted layer's of_dora_rankr_finish in one launch
__global__ void __launch_bounds__(256) lora_finish_all_kernel(const of_lora_finish_seg* __restrict__ segs, int num_segs) {
  int lo = 0, hi = num_segs - 1;
  const int cta = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (segs[mid].cta_begin <= cta) lo = mid;
    else hi = mid - 1;
  }
  const of_lora_finish_seg sg = segs[lo];
  const int idx = (cta - sg.cta_begin) * 256 + threadIdx.x;
  if (idx >= sg.Cout * sg.r) return;
  const int co = idx / sg.r, j = idx - co * sg.r;
  sg.gB[idx] += sg.rowscale[co] * sg.dBraw[idx];
  if (j == 0 && sg.mag && sg.gmag) sg.gmag[co] += sg.dm[co] / sg.mag[co];
}
// dst = bf16(scale * src)
__global__ void scale_cast_bf16_kernel(const float* __restrict__ src, float scale, __nv_bfloat16* __restrict__ dst, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16_rn(scale * src[i]);
}

// output-channel range per CTA: at most ONE wave of CTAs (register use allows one CTA per SM; a CTA pays a fixed ~2 us for its A
// tile), each walking its range 16 channels at a time
static dim3 dora_grid(int E, int Cout, int* co_per_cta) {
  const int etiles = (E + kDoraE - 1) / kDoraE;
  int ychunks = device_sm_count() / etiles;
  const int max_chunks = (Cout + kDoraCo - 1) / kDoraCo;
  if (ychunks > max_chunks) ychunks = max_chunks;
  if (ychunks < 1) ychunks = 1;
  *co_per_cta = ((Cout + ychunks - 1) / ychunks + kDoraCo - 1) / kDoraCo * kDoraCo;
  if (*co_per_cta > 512) *co_per_cta = 512;      // the range's B rows live in shared memory
  return dim3(etiles, (Cout + *co_per_cta - 1) / *co_per_cta);
}

static int dora_check(const float* W, const float* A, const float* B, int Cout, int Cin, int k, int r, const char* who) {
  OF_REQUIRE(W && A && B, "%s: null pointer", who);
  OF_REQUIRE(Cout >= 1 && Cin >= 1 && k >= 1 && r >= 1 && r <= kDoraRMax, "%s: bad sizes (r=%d)", who, r);
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
  int co_per_cta;
  dim3 grid = dora_grid(E, Cout, &co_per_cta);
  size_t smem = ((size_t)r * kDoraEP + (size_t)co_per_cta * r + co_per_cta) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    OF_CHECK_CUDA(cudaFuncSetAttribute(dora_norm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    OF_CHECK_CUDA(cudaFuncSetAttribute(dora_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    OF_CHECK_CUDA(cudaFuncSetAttribute(dora_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr = true;
  }
  if (mag) {
    OF_CHECK_CUDA(cudaMemsetAsync(n2_ws, 0, (size_t)Cout * sizeof(float), stream));
    dora_norm_kernel<<<grid, kDoraE, smem, stream>>>(a, n2_ws, co_per_cta);
    OF_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }
  dora_merge_kernel<<<grid, kDoraE, smem, stream>>>(a, n2_ws, reinterpret_cast<__nv_bfloat16*>(packed_bf16), tap_stride, s_out, co_per_cta);
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
  int co_per_cta;
  dim3 grid = dora_grid(E, Cout, &co_per_cta);
  size_t smem = ((size_t)r * kDoraEP + (size_t)co_per_cta * r + (size_t)kDoraCo * kDoraEP + kDoraCo) * sizeof(float);
  dora_grad_kernel<<<grid, kDoraE, smem, stream>>>(a, n2, dW_packed, tap_stride, dA, dB, dmag, co_per_cta);
  OF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OF_OK;
}

extern "C" int of_dora_rankr_prep(const float* B, const float* mag, const float* n2, float scaling, int Cout, int r, void* Bst_bf16,
                                  float* rowscale, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  OF_REQUIRE(B && Bst_bf16 && rowscale && Cout >= 1 && r >= 1 && (!mag || n2), "of_dora_rankr_prep: bad args");
  const int n = Cout * r;
  dora_rankr_prep_kernel<<<(n + 255) / 256, 256, 0, stream>>>(B, mag, n2, scaling, Cout, r, reinterpret_cast<__nv_bfloat16*>(Bst_bf16),
                                                              rowscale);
  OF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OF_OK;
}

extern "C" int of_dora_rankr_finish(const float* dBraw, const float* rowscale, float* gB, const float* dm, const float* mag, float* gmag,
                                    int Cout, int r, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  OF_REQUIRE(dBraw && rowscale && gB && Cout >= 1 && r >= 1 && (!mag || (dm && gmag)), "of_dora_rankr_finish: bad args");
  const int n = Cout * r;
  dora_rankr_finish_kernel<<<(n + 255) / 256, 256, 0, stream>>>(dBraw, rowscale, gB, dm, mag, gmag, Cout, r);
  OF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OF_OK;
}

extern "C" int of_dora_scale_pack(const float* V, const float* mag, int Cout, int Cin, int k, float* n2_out, void* packed_bf16, int cin_pad,
                                  long long tap_stride, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  OF_REQUIRE(V && packed_bf16 && Cout >= 1 && Cin >= 1 && k >= 1 && cin_pad >= Cin, "of_dora_scale_pack: bad args");
  OF_REQUIRE((long long)Cin * k <= 256LL * kScalePackMaxPerThread, "of_dora_scale_pack: row of %d x %d elements is too long", Cin, k);
  dora_scale_pack_kernel<<<Cout, 256, 0, stream>>>(V, mag, Cin, k, n2_out, reinterpret_cast<__nv_bfloat16*>(packed_bf16), cin_pad,
                                                    tap_stride, nullptr, 0.f, 0, nullptr, nullptr);
  OF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OF_OK;
}

extern "C" int of_dora_scale_pack_prep(const float* V, const float* mag, int Cout, int Cin, int k, float* n2_out, void* packed_bf16,
                                       int cin_pad, long long tap_stride, const float* B, float scaling, int r, void* Bst_bf16,
                                       float* rowscale, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  OF_REQUIRE(V && packed_bf16 && Cout >= 1 && Cin >= 1 && k >= 1 && cin_pad >= Cin, "of_dora_scale_pack_prep: bad args");
  OF_REQUIRE(B && Bst_bf16 && rowscale && r >= 1, "of_dora_scale_pack_prep: null rank-r operand");
  OF_REQUIRE((long long)Cin * k <= 256LL * kScalePackMaxPerThread, "of_dora_scale_pack_prep: row of %d x %d elements is too long", Cin, k);
  dora_scale_pack_kernel<<<Cout, 256, 0, stream>>>(V, mag, Cin, k, n2_out, reinterpret_cast<__nv_bfloat16*>(packed_bf16), cin_pad,
                                                    tap_stride, B, scaling, r, reinterpret_cast<__nv_bfloat16*>(Bst_bf16), rowscale);
  OF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OF_OK;
}

extern "C" int of_lora_finish_all(const of_lora_finish_seg* segs_dev, int num_segs, int total_ctas, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  OF_REQUIRE(segs_dev && num_segs >= 1 && total_ctas >= 1, "of_lora_finish_all: bad args");
  lora_finish_all_kernel<<<total_ctas, 256, 0, stream>>>(segs_dev, num_segs);
  OF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OF_OK;
}

extern "C" int of_scale_cast_f32_bf16(const float* src, float scale, void* dst, long long n, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  OF_REQUIRE(src && dst && n >= 0, "of_scale_cast_f32_bf16: bad args");
  if (n > 0) scale_cast_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, scale, reinterpret_cast<__nv_bfloat16*>(dst), n);
  OF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OF_OK;
}
