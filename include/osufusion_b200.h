/*
 * osufusion_b200 — C-ABI of the B200 (sm_100a) hot-path kernels.
 *
 * The reference (fauzanardh/OsuFusion) has no FFI layer of its own: its boundary is the Python
 * nn.Module protocol.  This header is the boundary *below* that protocol: every
 * entry point replaces a group of torch-eager library calls the reference makes, cited per function
 * as  <reference file>:<line>.  Conventions:
 *   - extern "C", plain pointers + sizes, no torch / C++ types;
 *   - every call is stream-ordered on `stream` (a cudaStream_t passed as void*), never synchronises,
 *     never allocates; outputs and workspaces are caller-owned device buffers;
 *   - returns 0 on success, a negative code on error; of_last_error() gives the message;
 *   - activations are channels-last: (batch, rows=L, channels) with an explicit leading dimension
 *     (`ld`, in elements) and batch stride so that channel-slices of wider buffers can be used
 *     (skip-connection concats are written in place, never materialised by a copy);
 *   - bf16 = raw uint16 storage (__nv_bfloat16), f32 = float.
 */
#ifndef OSUFUSION_B200_H_
#define OSUFUSION_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OF_OK 0
#define OF_ERR_INVALID (-1)
#define OF_ERR_CUDA (-2)
#define OF_ERR_UNSUPPORTED (-3)

const char* of_last_error(void);
int of_version(void);
/* Number of kernel launches issued through this library since load (for bench.py `gpu_launches`). */
long long of_launch_count(void);
void of_reset_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * of_gemm — tcgen05/TMEM/TMA GEMM + implicit-GEMM conv1d (forward, dgrad, wgrad) with fused epilogue.
 * Replaces: nn.Conv1d k=3 `Block.proj` (residual.py:70,75), 1x1 `res_conv` (residual.py:115,137),
 *   `Parallel` (unet.py:225,234), `Downsample`/`Upsample` convs (unet.py:64-69,81-86), CrossEmbedLayer
 *   (unet.py:42-58), nn.Linear to_q/to_kv/to_out (unet.py:118-123), FeedForward (unet.py:149-156),
 *   `final_conv` (unet.py:354,513), and their autograd backward (cuDNN/cuBLAS in the reference).
 *
 * mode OF_GEMM_FWD (also dgrad):  for each batch b, m in [0,rows), n in [0,N):
 *     acc[b,m,n] = sum_{t<taps} sum_{k<K}  A[b, m + shift0 + t*shift_step, k] * Bt[t](n,k)
 *   A: (batch, rows, K) bf16, K contiguous, rows outside [0,rows) read as zero (conv padding).
 *   b_mn_major = 0:  B is [taps][N][K] (K contiguous);   Bt[t](n,k) = B[t][n][k]
 *   b_mn_major = 1:  B is [taps][K][N] (N contiguous);   Bt[t](n,k) = B[t][k][n]   (dgrad: W used transposed)
 * mode OF_GEMM_WGRAD:  for each tap t, m in [0,M=rows_out), n in [0,N):
 *     acc[t,m,n] = sum_b sum_{l<red_rows}  A[b,l,m] * B[b, l + shift0 + t*shift_step, n]
 *   A: (batch, red_rows, M) bf16 (dY), B: (batch, red_rows, N) bf16 (X), both channel-contiguous.
 *   The result is ATOMICALLY ADDED (fp32) into out_f32[t][m][n]  (split-K over (batch,l); grads accumulate).
 *
 * Epilogue (FWD mode), in this order, all optional:
 *     v = acc;  v += bias[n];  v += aux_f32[b,m,n];  v += float(aux_bf16[b,m,n]) (if !aux_is_dsilu)
 *     if pre_bf16: pre_bf16[b,m,n] = bf16(v)
 *     if act == OF_ACT_SILU: v = silu(v)
 *     if aux_is_dsilu: v *= dsilu(float(aux_bf16[b,m,n]))
 *     out_bf16 = bf16(v); out_f32 = v;
 *     if stats: stats[b][0] += sum(bf16round(v)), stats[b][1] += sum(bf16round(v)^2)   (double; GroupNorm(1,C))
 * Constraints: K % 8 == 0 (A/B leading dims multiples of 8 elements), N % 8 == 0, 16-byte aligned pointers.
 * ------------------------------------------------------------------------------------------------ */
#define OF_GEMM_FWD 0
#define OF_GEMM_WGRAD 1
#define OF_ACT_NONE 0
#define OF_ACT_SILU 1

typedef struct {
  int mode;       /* OF_GEMM_FWD | OF_GEMM_WGRAD */
  int b_mn_major; /* FWD only */
  int batch;      /* number of independent samples (batch dim of A / outputs) */
  int rows;       /* FWD: output rows per sample (L).  WGRAD: reduction rows per sample (L) */
  int N;          /* output columns */
  int K;          /* FWD: reduction channels per tap.  WGRAD: output rows M (channels of A) */
  int taps;       /* >= 1 */
  int shift0, shift_step;
  const void* a;  /* bf16 */
  long long a_ld, a_batch_stride; /* elements */
  const void* b;  /* bf16 */
  long long b_ld, b_tap_stride;   /* FWD: per-tap matrix ld / stride.  WGRAD: b_ld, b_tap_stride = batch stride */
  /* epilogue */
  const float* bias;              /* [N] or NULL */
  const float* aux_f32;           /* (batch, rows, N) or NULL */
  long long aux_f32_ld, aux_f32_batch_stride;
  const void* aux_bf16;           /* bf16 or NULL */
  long long aux_bf16_ld, aux_bf16_batch_stride;
  int aux_is_dsilu;
  int act;
  void* pre_bf16;                 /* bf16 or NULL (same ld/stride as out_bf16) */
  void* out_bf16;                 /* bf16 or NULL */
  long long out_bf16_ld, out_bf16_batch_stride;
  float* out_f32;                 /* or NULL.  WGRAD: [taps][M][N] accumulated atomically */
  long long out_f32_ld, out_f32_batch_stride; /* WGRAD: ld and per-tap stride */
  double* stats;                  /* [batch][2] or NULL */
  /* tuning (0 = auto) */
  int block_n;
  int split_k;
} of_gemm_args;

int of_gemm(const of_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * of_attn_fwd / of_attn_bwd — multi-query flash attention (tcgen05/TMEM), non-causal, no mask, bf16.
 * Replaces: `Attend.forward` = F.scaled_dot_product_attention on bf16-cast q,k,v (attention.py:77-101) and the
 *   GQA expansion `repeat(k|v, "b h n d -> b (r h) n d")` (unet.py:135-137), plus their autograd backward.
 * Layout: q (B, L, H*D), k and v (B, L, KVH*D) channels-last views (ld / batch stride in elements), q head i uses
 *   kv head i % KVH (einops "(r h)" ordering).  D <= 64, D % 8 == 0.  scale <= 0 selects 1/sqrt(D).
 * fwd: out (B, L, H*D) bf16, lse (B, H, L) fp32 = log2-domain log-sum-exp (row max + log2 sum), saved for bwd.
 * bwd: inputs q,k,v,out,dout (bf16), lse; workspace delta (B,H,L) fp32;
 *      dq (B, L, H*D) fp32 and dkv = [dk | dv] (B, L, 2*KVH*D) fp32 are ACCUMULATED atomically: caller zero-fills.
 * variant: 0 = default (P operand kept in tensor memory), 1 = P staged through shared memory.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  int B, H, KVH, L, D;
  float scale;
  int variant;
  const void* q; long long q_ld, q_batch_stride;
  const void* k; const void* v; long long kv_ld, kv_batch_stride;
  void* out; long long out_ld, out_batch_stride;      /* fwd: output; bwd: forward output (input) */
  float* lse;
  /* backward only */
  const void* dout; long long dout_ld, dout_batch_stride;
  float* delta;                                        /* (B, H, L) workspace */
  float* dq; long long dq_ld, dq_batch_stride;
  float* dk; float* dv; long long dkv_ld, dkv_batch_stride;
} of_attn_args;

int of_attn_fwd(const of_attn_args* args, void* stream);
int of_attn_bwd(const of_attn_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OSUFUSION_B200_H_ */
