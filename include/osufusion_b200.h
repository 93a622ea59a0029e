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
/* Upper bound on the SMs persistent kernels may assume (0 = all).  Used by the data-parallel wrapper to leave room for the NCCL
 * all-reduce kernels that overlap backward (torch DDP gives the reference the same overlap: trainer.py:211-220,301). */
int of_set_sm_limit(int n);
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
 *      dq (B, L, H*D) fp32 and dkv = [dk | dv] (B, L, 2*KVH*D) fp32 are ACCUMULATED atomically: the caller zero-fills them, or sets
 *      zero_grads and the delta pre-pass (one pass over out / dout that every call runs anyway) does it: no fill launches.
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
  int zero_grads;                                      /* bwd: != 0 -> the delta pre-pass zero-fills dq / dk / dv itself */
} of_attn_args;

int of_attn_fwd(const of_attn_args* args, void* stream);
int of_attn_bwd(const of_attn_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * ResidualBlock bandwidth kernels (channels-last, GroupNorm(1,C) statistics come from of_gemm's epilogue).
 * Replace, per ResidualBlock (residual.py:118-137): nn.GroupNorm(1,C) + FiLM `x*(scale+1)+shift` + SiLU
 * (residual.py:71-83), GlobalContext `to_k` 1x1 conv + softmax over L + pooled einsum (residual.py:29-31),
 * `h * se(h) + res_conv(x)` (residual.py:135-137) and their autograd backward.
 * h = SiLU(FiLM(GN(y))) is never stored for block2: every consumer recomputes it from the bf16 conv output y.
 *
 *   of_rb_apply_fwd       out_bf16 = h                                   (block1: input of the second conv)
 *   of_rb_rowdot          mode 0: out_rows[b,l] = bf16r( sum_c bf16r(h)*bf16r(vec[c]) + vec_bias )     (to_k logits)
 *                         mode 1: out_rows[b,l] = sum_c bf16r(h)*vec[b,c]      (= d p, with vec = d pooled)
 *   of_softmax_rows       p[b,:] = softmax_L(logits[b,:]) in place, fp32 (consumers round to bf16 like the autocast einsum)
 *   of_softmax_bwd_rows   rd[b,:] <- p * (rd - sum_l p*rd)                 (softmax backward -> d logits)
 *   of_rb_pool            acc_bc[b,c] += sum_l bf16r(h[b,l,c]) * p[b,l]
 *   of_rb_gate_fwd        out = h * gate[b,c] + res   -> out_f32 and/or out_bf16
 *   of_rb_gate_bwd_reduce acc_bc[b,c] += sum_l dout_f32[b,l,c] * h[b,l,c]                              (d gate)
 *   of_rb_bwd_pass1       mode 0 (block2): dh = dout_f32*gate + dpooled[b,c]*p[b,l] + da[b,l]*wk[c]
 *                         mode 1 (block1): dh = dh_bf16
 *                         df = dh*silu'(f); dss += (df*z | df); dz = df*(scale+1); dgamma += dz*xhat; dbeta += dz;
 *                         dxhat_bf16 = dz*gamma; dstats[b] += (sum dxhat, sum dxhat*xhat);
 *                         mode 0 also: dwk[c] += da*bf16r(h), dbk += da, optional dout_bf16 = bf16(dout_f32)
 *   of_rb_bwd_apply       dy_bf16 = rstd*(dxhat - S1/n - xhat*S2/n);  dbias[c] += dy
 * Constraints: C % 8 == 0, C <= 2048.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  int B, L, C;
  float eps;
  int mode;
  const void* y; long long y_ld, y_bs;
  const double* stats;
  const float* gamma; const float* beta;
  const float* ss;
  const float* vec; long long vec_bs;
  const float* vec_bias;
  const float* p;
  const float* pooled;
  const float* gate;
  const float* res_f32; long long res_f32_ld, res_f32_bs;
  const void* res_bf16; long long res_bf16_ld, res_bf16_bs;
  float* out_f32; long long out_f32_ld, out_f32_bs;
  void* out_bf16; long long out_bf16_ld, out_bf16_bs;
  float* out_rows;
  float* acc_bc;
  const float* dout_f32; long long dout_f32_ld, dout_f32_bs;
  const void* dh_bf16; long long dh_ld, dh_bs;
  const float* dpooled;
  const float* da;
  const float* wk;
  double* dstats;
  float* dgamma; float* dbeta;
  float* dwk; float* dbk;
  float* dss;
  void* dxhat_bf16; long long dxhat_ld, dxhat_bs;
  void* dout_bf16; long long dout_bf16_ld, dout_bf16_bs;
  void* dy_bf16; long long dy_ld, dy_bs;
  float* dbias;
} of_rb_args;

int of_rb_apply_fwd(const of_rb_args* a, void* stream);
int of_rb_rowdot(const of_rb_args* a, void* stream);
/* GlobalContext forward fused (residual.py:29-32): logits = to_k(h) (bf16-rounded, written to out_rows), softmax over L and the
 * pooled channel sums in ONE pass over y with a CTA-local online softmax; a finishing launch combines the per-CTA records
 * part[B][of_rb_pool_parts()][C+2], writes pooled[B][C] and converts out_rows to the fp32 probabilities.  Replaces
 * of_rb_rowdot(mode 0) + of_softmax_rows + of_rb_pool (one pass over y less). */
int of_rb_pool_parts(const of_rb_args* a);
int of_rb_logit_pool(const of_rb_args* a, float* part, float* pooled, void* stream);
int of_rb_pool(const of_rb_args* a, void* stream);
int of_rb_gate_fwd(const of_rb_args* a, void* stream);
int of_rb_gate_bwd_reduce(const of_rb_args* a, void* stream);
int of_rb_bwd_pass1(const of_rb_args* a, void* stream);
int of_rb_bwd_apply(const of_rb_args* a, void* stream);
int of_softmax_rows(float* rows, int B, int L, void* stream);
int of_softmax_bwd_rows(const float* p, float* rd, int B, int L, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Transformer-block and glue kernels.
 * of_layernorm_fwd/bwd : nn.LayerNorm(C) of `Attention.norm` (unet.py:117,127) on the fp32 residual stream; emits the
 *                        fp32 normed tensor (the residual of unet.py:141) and its bf16 copy for the q/kv GEMM.
 * of_rope_fwd/bwd      : RotaryPositionEmbedding.forward + apply_rotary_pos_emb/rotate_half (attention.py:52-58,
 *                        utils.py:25-32) applied in place to the q and k slots of the fused qkv buffer; the cos/sin
 *                        tables (L, D) are generated by the host exactly as the reference does (attention.py:33-49): bf16 when
 *                        q is bf16 (autocast), fp32 (table_f32=1, fp32 arithmetic) when a DoRA-adapted to_q promotes q to fp32.
 *                        bwd also converts the fp32 attention gradients to the bf16 (B, L, (H+2KVH)*D) layout.
 * of_linear_small_*    : nn.Linear with M <= 16 rows: time_mlp / cond_mlp (unet.py:356-366), FiLM heads
 *                        (residual.py:104-111), GlobalContext.layers 1x1 convs on the pooled (B,C,1) vector
 *                        (residual.py:22-27).  act: 0 none, 1 SiLU, 2 Sigmoid.  round_bf16 mimics autocast rounding
 *                        of inputs, weights and outputs.  bwd: dW/dbias are written (accumulate=0) or accumulated (+=), dx accumulates atomically.
 * of_colsum_bf16       : bias gradients of the large Linear/Conv layers: db[n] += sum_rows dy[row, n].
 * of_pack_input        : F.pad(x, value=-1) / F.pad(a, value=-23) (unet.py:475-480) + channel-first -> channels-last
 *                        bf16, fused with `add_noise` (diffusion.py:96) / `t*x+(1-t)*noise` (rectified_flow.py:95).
 * of_unpack_output     : `final_conv(x)[:, :, :n]` slice + layout back to (B, 6, N) fp32 (unet.py:513).
 * of_upsample2x_*      : F.interpolate(scale_factor=2, mode="nearest") of `Upsample` (unet.py:65) and its backward.
 * of_cast_copy         : strided fp32/bf16 copy, cast or accumulate on (B, L, C) views (concat slices, grad sums).
 * of_time_embed        : SinusoidalPositionEmbedding (unet.py:26-39).
 * of_silu_small        : nn.SiLU on the (B, 2*dim_emb) conditioning vector (residual.py:105) and its backward.
 * of_mse_fwd/bwd       : F.mse_loss(pred, target, "none") + orig_len mask + mean (diffusion.py:101-111).
 * of_sampler_update    : CFG combine (unet.py:458-465) fused with DDIMScheduler.step (diffusers 0.29.2, eta=0,
 *                        clip_sample; call site diffusion.py:75) or the midpoint axpy of torchdiffeq's fixed-grid
 *                        solver (rectified_flow.py:78); also emits the packed bf16 input of the next denoiser call.
 * of_pack_conv_weight / of_unpack_conv_wgrad / of_cast_f32_bf16 : parameter layout conversion between the
 *                        reference's state_dict layout (Cout, Cin, k) fp32 and the [tap][Cout][Cin] bf16 GEMM operand.
 * ------------------------------------------------------------------------------------------------ */
int of_layernorm_fwd(const float* x, long long x_ld, int rows, int C, const float* gamma, const float* beta, float eps,
                     float* out_f32, void* out_bf16, long long out_ld, float* mean_rstd, void* stream);
int of_layernorm_bwd(const float* dy, long long dy_ld, const float* x, long long x_ld, int rows, int C, const float* gamma,
                     const float* mean_rstd, float* dx_f32, void* dx_bf16, long long dx_ld, float* dgamma, float* dbeta,
                     void* stream);
int of_rope_fwd(void* qkv, long long ld, long long bs, int B, int L, int H, int KVH, int D, const void* cos_tab,
                const void* sin_tab, int table_f32, void* stream);
int of_rope_bwd(const float* dq, long long dq_ld, long long dq_bs, const float* dk, const float* dv, long long dkv_ld,
                long long dkv_bs, void* dqkv_bf16, long long out_ld, long long out_bs, int B, int L, int H, int KVH, int D,
                const void* cos_tab, const void* sin_tab, int table_f32, void* stream);
int of_linear_small_fwd(const float* x, long long x_ld, int M, int N, int K, const float* W, long long w_ld, const float* bias,
                        int act, int round_bf16, float* y, long long y_ld, float* ypre, void* stream);
int of_linear_small_bwd(const float* dy, long long dy_ld, const float* ypre, int act, const float* x, long long x_ld, int M,
                        int N, int K, const float* W, long long w_ld, int round_bf16, float* dW, float* dbias, float* dx,
                        long long dx_ld, int accumulate, void* stream);
int of_colsum_bf16(const void* dy, long long ld, long long rows, int N, float* db, void* stream);
/* out[n] += sum_rows dy[row,n] * (y[row,n] - bias[n]): the DoRA magnitude gradient taken from the activations,
 * d mag[co] = sum_l dy[l,co] * conv(x, V)[l,co] / ||V_co|| = sum_l dy * (y - b) / mag   (lora_layers.py:76-90 with the norm detached). */
int of_coldot_bf16(const void* dy, long long dy_ld, const void* y, long long y_ld, long long rows, int N, const float* bias, float* out,
                   void* stream);
int of_pack_input(const float* x, const float* noise, const float* ca, const float* cb, int B, int C, int N, void* out, int Lp,
                  int Cp, float pad_value, void* stream);
int of_unpack_output(const void* y, long long ld, long long bs, int B, int C, int N, float* out, void* stream);
int of_upsample2x_fwd(const void* x, long long x_ld, long long x_bs, int B, int L, int C, void* out, long long o_ld,
                      long long o_bs, void* stream);
int of_upsample2x_bwd(const float* d, long long d_ld, long long d_bs, int B, int L, int C, float* out_f32, void* out_bf16,
                      long long o_ld, long long o_bs, void* stream);
int of_cast_copy(const float* src32, const void* src16, long long s_ld, long long s_bs, int B, int L, int C, float* dst32,
                 void* dst16, long long d_ld, long long d_bs, int accumulate, void* stream);
int of_time_embed(const float* t, int B, int dim, float theta, float* out, void* stream);
int of_silu_small(const float* x, const float* dy, float* out, long long n, void* stream);
int of_mse_fwd(const void* pred, long long ld, long long bs, const float* x, const float* noise, float ta, float tb,
               const long long* orig_len, int B, int C, int N, float* accum2, float* loss, void* stream);
int of_mse_bwd(const void* pred, long long ld, long long bs, const float* x, const float* noise, float ta, float tb,
               const long long* orig_len, int B, int C, int N, int Lp, int Cp, const float* accum2, const float* gscale,
               void* dpred, void* stream);
int of_sampler_update(const float* xin, const void* cond, const void* null_, long long ld, long long bs, float cond_scale,
                      int mode, float c_eps, float c_div, float c_x0, float c_dir, int B, int C, int N, float* xout,
                      void* packed, int Lp, int Cp, float pad_value, void* stream);
/* same, with the four per-step coefficients {c_eps, c_div, c_x0, c_dir} read from DEVICE memory, so that one captured CUDA graph
 * (denoiser evaluation + update) serves every step of the loop (diffusion.py:71-75, rectified_flow.py:69-79) */
int of_sampler_update_dev(const float* xin, const void* cond, const void* null_, long long ld, long long bs, float cond_scale,
                          int mode, const float* coef_dev, int B, int C, int N, float* xout, void* packed, int Lp, int Cp,
                          float pad_value, void* stream);
int of_pack_conv_weight(const float* w, int Cout, int Cin, int k, void* out, int Cin_pad, int tap_offset, int taps_total,
                        void* stream);
int of_unpack_conv_wgrad(float* packed, int Cout, int Cin, int k, int Cin_pad, int tap_offset, float* dw, int accumulate,
                         int rezero, void* stream);
int of_cast_f32_bf16(const float* src, void* dst, long long n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LoRA / DoRA adapters.  Replace `LoraConv1d.forward` + `DoraConv1dLayer.forward` (lora_layers.py:59-92,292-328: base conv,
 * a SECOND base conv scaled by (s-1), lora_B(lora_A(x))) and peft 0.12.0's DoRA nn.Linear path by one GEMM on the effective
 * weight W_eff = s * (W + scaling * B A), s = magnitude / ||W + scaling B A||_2 (norm detached, lora_layers.py:76-79).
 *   of_dora_merge : W (Cout,Cin,k) fp32, A (r,Cin,k), B (Cout,r), mag (Cout) or NULL (plain LoRA)  ->
 *                   packed bf16 [k][Cout][Cin_pad] GEMM operand (tap stride given), n2_ws[Cout] = ||.||^2, s_out[Cout] (optional)
 *   of_dora_grad  : dW_packed = d loss / d W_eff (fp32, same packed layout, from of_gemm WGRAD) -> dA, dB, dmag (accumulated)
 * ------------------------------------------------------------------------------------------------ */
int of_dora_merge(const float* W, const float* A, const float* B, const float* mag, float scaling, int Cout, int Cin, int k, int r,
                  float* n2_ws, void* packed_bf16, int Cin_pad, long long tap_stride, float* s_out, void* stream);
/* Rank-r backward glue (no full weight gradient of the frozen base; see engine._adapter_backward_rank_r):
 *   of_dora_rankr_prep   : rowscale[co] = scaling * s[co] (s = mag / sqrt(n2), 1 without magnitude);  Bst[j][co] = bf16(rowscale[co] * B[co][j])
 *   of_dora_rankr_finish : gB[co][j] += rowscale[co] * dBraw[co][j];  gmag[co] += dm[co] / mag[co] */
int of_dora_rankr_prep(const float* B, const float* mag, const float* n2, float scaling, int Cout, int r, void* Bst_bf16,
                       float* rowscale, void* stream);
int of_dora_rankr_finish(const float* dBraw, const float* rowscale, float* gB, const float* dm, const float* mag, float* gmag, int Cout,
                         int r, void* stream);
/* Tensor-core merge: V = W + scaling*B*A is an of_gemm call (M = Cout, N = Cin*k, K = r, fp32 W as aux_f32, fp32 V out);
 *   of_dora_scale_pack      : per output channel n2 = ||V_co||^2, s = mag/sqrt(n2) (1 without mag), packed[t][co][ci] = bf16(s*V)
 *   of_scale_cast_f32_bf16  : dst = bf16(scale * src)   (the scaling*B operand) */
int of_dora_scale_pack(const float* V, const float* mag, int Cout, int Cin, int k, float* n2_out, void* packed_bf16, int cin_pad,
                       long long tap_stride, void* stream);
int of_scale_cast_f32_bf16(const float* src, float scale, void* dst, long long n, void* stream);
/* of_dora_scale_pack + of_dora_rankr_prep in one launch (the CTA that owns output channel co knows s[co]): additionally writes
 *   rowscale[co] = scaling * s[co]  and  Bst[j][co] = bf16(rowscale[co] * B[co][j])  for the rank-r backward. */
int of_dora_scale_pack_prep(const float* V, const float* mag, int Cout, int Cin, int k, float* n2_out, void* packed_bf16, int cin_pad,
                            long long tap_stride, const float* B, float scaling, int r, void* Bst_bf16, float* rowscale, void* stream);
/* of_dora_rankr_finish for every adapted layer of the model in ONE launch at the end of backward (table in device memory; layer i
 * owns CTAs [cta_begin_i, cta_begin_{i+1}) of 256 threads, one thread per element of B). */
typedef struct {
  const float* dBraw;     /* (Cout, r) fp32 = dy^T u accumulated by the weight-gradient GEMM */
  const float* rowscale;  /* (Cout) */
  float* gB;              /* (Cout, r) gradient of lora_B (accumulated into) */
  const float* dm;        /* (Cout) sum_l dy (y - b), or NULL without magnitude */
  const float* mag;       /* (Cout) or NULL */
  float* gmag;            /* (Cout) gradient of the magnitude vector (accumulated into) or NULL */
  int Cout, r;
  int cta_begin;
  int _pad;
} of_lora_finish_seg;
int of_lora_finish_all(const of_lora_finish_seg* segs_dev, int num_segs, int total_ctas, void* stream);
int of_dora_grad(const float* W, const float* A, const float* B, const float* mag, float scaling, int Cout, int Cin, int k, int r,
                 const float* n2, const float* dW_packed, int Cin_pad, long long tap_stride, float* dA, float* dB, float* dmag,
                 void* stream);

/* ------------------------------------------------------------------------------------------------
 * Grouped ("multi-tensor") kernels: one launch over many parameter tensors.  Descriptor tables are arrays in DEVICE memory
 * built once per model by the host.
 *   of_film_fwd  : every FiLM head `ResidualBlock.mlp[1]` (nn.Linear(2*dim_emb, 2*C), residual.py:104-111) at once:
 *                  out[out_off_g + m*N_g + n] = bf16(bias_g[n] + sum_k bf16(W_g[n,k]) * x[m,k]);  x (M, K) fp32 already holds
 *                  bf16-rounded values (SiLU(cat(t, c)) under autocast).  Bytes: sum_g N_g*K*4 read once.
 *   of_film_bwd  : dW_g[n,k] = sum_m dss[m,n] x[m,k] and dbias_g[n] = sum_m dss[m,n] are WRITTEN (not accumulated; NULL =
 *                  frozen), d_emb[m,k] += sum_g sum_n dss[m,n] bf16(W_g[n,k]) (atomic).  chunks = (group, first row) pairs, each
 *                  covering at most of_film_chunk_rows() rows of one head.  Bytes: weights read once + gradients written once.
 *   of_pack_weights : fp32 (Cout, Cin, k) master weights -> bf16 [k][Cout][cin_pad] GEMM operands (k == 1: plain cast) for a
 *                  whole list of tensors; segment i owns CTAs [cta_begin_i, cta_begin_{i+1}), of_pack_seg_ctas() CTAs each
 *                  (-1: unsupported kernel size, use of_pack_conv_weight).
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  const float* W;      /* (N, K) fp32 */
  const float* bias;   /* (N) or NULL */
  float* dW;           /* (N, K) fp32 or NULL */
  float* dbias;        /* (N) or NULL */
  long long out_off;   /* offset (floats) of this head's (M, N) block in out / dss: M * row_start */
  int N;
  int row_start;       /* first row of this head in the concatenated row space */
} of_film_group;

typedef struct {
  const float* src;
  void* dst;
  int Cout, Cin, k, cin_pad;
  int cta_begin;
  float scale;         /* dst = bf16(scale * src); 0 means 1 (plain cast) */
} of_pack_seg;

int of_film_fwd(const of_film_group* groups_dev, int num_groups, int total_rows, const float* x, int M, int K, float* out,
                void* stream);
int of_film_bwd(const of_film_group* groups_dev, const int* chunks_dev, int num_chunks, const float* dss, const float* x, int M,
                int K, float* d_emb, void* stream);
int of_film_chunk_rows(void);
int of_pack_weights(const of_pack_seg* segs_dev, int num_segs, int total_ctas, void* stream);
int of_pack_seg_ctas(int Cout, int Cin, int k, int cin_pad);

/* ------------------------------------------------------------------------------------------------
 * Fused optimizer step (trainer.py:302-309: clip_grad_norm_(params, 1.0) + torch.optim.AdamW.step()).
 * Gradients / exp_avg / exp_avg_sq are flat fp32 arenas with one common layout; parameters are separate tensors.
 *   of_grad_sumsq : out[0] = sum g^2 over the arena (double; padding between tensors must be zero)
 *   of_adamw_step : torch.optim.AdamW semantics (decoupled weight decay, bias correction, amsgrad off); gradients are
 *                   scaled by min(1, max_norm / (sqrt(sumsq) + 1e-6)) read from DEVICE memory (sumsq NULL or max_norm <= 0:
 *                   no clipping) -- no host synchronisation.  Bytes per element: 16 read + 12 written (+ 2 written when the
 *                   bf16 GEMM operand is emitted).
 *   Two extensions remove a full-tensor pass each from the training step:
 *     - `operand_bf16` non-NULL: the updated parameter is also written as the bf16 GEMM operand the next forward pass consumes
 *       (what of_pack_weights produced: the fp32 -> bf16 re-cast torch autocast does on every iteration);
 *     - `k > 1` ("packed" Conv1d weight): gradient, moments and operand use the GEMM layout [k][Cout][Cin] (the weight-gradient
 *       GEMM accumulates straight into the arena, no unpack pass), the fp32 master parameter keeps torch's (Cout, Cin, k) layout;
 *       the kernel converts through a shared-memory slab (CTA = one output channel x 256 input channels, all taps).
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  float* param;          /* fp32 parameter tensor (contiguous, torch layout) */
  long long arena_off;   /* offset (floats) of its gradient / moments in the arenas */
  long long numel;
  void* operand_bf16;    /* NULL or the bf16 operand copy to refresh (same layout as the gradient) */
  int cta_begin;         /* running sum of of_opt_tensor_ctas2(numel, Cout, Cin, k) */
  int Cout, Cin, k;      /* k > 1: packed conv tensor (numel == Cout*Cin*k); k <= 1: flat */
} of_opt_tensor;

int of_grad_sumsq(const float* grads, long long n, double* out, void* stream);
int of_opt_tensor_ctas(long long numel);
int of_opt_tensor_ctas2(long long numel, int Cout, int Cin, int k);
int of_adamw_step(const of_opt_tensor* table_dev, int num_tensors, int total_ctas, const float* grads, float* exp_avg,
                  float* exp_avg_sq, const double* sumsq, float max_norm, float lr, float beta1, float beta2, float eps,
                  float weight_decay, int step, void* stream);

/* ------------------------------------------------------------------------------------------------
 * DiT / MMDiT backbones (osu_fusion/modules/dit.py, mmdit.py).  Their GEMMs, attention, adaLN LayerNorm
 * (of_layernorm_* called per sample with gamma = 1 + scale_b, beta = shift_b: `modulate(norm(x), shift, scale)`, dit.py:13-15,150-152)
 * and small conditioning MLPs reuse the entry points above; these cover what the UNet path has no counterpart for.
 *   of_gate_residual_fwd : `x + gate.unsqueeze(1) * y` (dit.py:151-152, mmdit.py:199-204): out (B,L,C) fp32 residual stream =
 *                          x (fp32 or bf16) + gate[b,:] * y (bf16); round_bf16 = the product is a bf16 x bf16 multiplication under
 *                          autocast.  Bytes/element: 4 (or 2) + 2 read, 4 written.
 *   of_gate_mul_bwd      : backward of the product: dx16 = bf16(d), dy16 = bf16(gate * d) for the incoming fp32 stream gradient d
 *                          (both are GEMM / of_coldot_bf16 operands: d gate[b,c] = sum_l dx16 * y).  Bytes/element: 4 read, 4 written.
 *   of_headnorm_fwd/bwd  : `MultiHeadRMSNorm` (dit.py:62-69, mmdit.py:55-62) = F.normalize(x, dim=-1) * gamma * sqrt(D) on the q and k
 *                          heads of a fused [Hq*D q | Hk*D k | Hv*D v] projection row (bf16), v copied through; strided in/out views
 *                          so the two modality streams of JointAttention (mmdit.py:98-129) are written straight into the joint
 *                          [audio ; beatmap] sequence.  bwd takes the fp32 dq/dk/dv of of_attn_bwd and the PRE-norm projection, emits
 *                          the bf16 [dq | dk | dv] operand and accumulates dgamma_q (Hq*D) / dgamma_k (Hk*D).
 *                          variant: 1 = one thread per head vector; 2 = one thread per 16-byte vector, the D/8 lanes of a head
 *                          reduce with shuffles (D in {8,16,32,64}; coalesced 128-byte lines, register gamma gradients);
 *                          0 = auto (2 where supported).  Bytes/element: fwd 2 + 2, bwd 4 + 2 read, 2 written.
 *   of_row_mean_std      : statistic audio pooling `cat([a.mean(-1), a.std(-1)])` (dit.py:279-282, mmdit.py:352-355) of the fp32
 *                          (B, C, N) spectrogram -> (B, 2C) fp32 (unbiased std).
 * ------------------------------------------------------------------------------------------------ */
int of_gate_residual_fwd(const float* x32, const void* x16, long long x_ld, long long x_bs, const void* y16, long long y_ld,
                         long long y_bs, const float* gate, long long gate_ld, int round_bf16, int B, int L, int C, float* out32,
                         long long o_ld, long long o_bs, void* stream);
int of_gate_mul_bwd(const float* dx32, long long d_ld, long long d_bs, const float* gate, long long gate_ld, int round_bf16, int B,
                    int L, int C, void* dx16, void* dy16, long long o_ld, long long o_bs, void* stream);
int of_headnorm_fwd(const void* in16, long long in_ld, long long in_bs, int B, int L, int Hq, int Hk, int Hv, int D,
                    const float* gamma_q, const float* gamma_k, float scale, void* out16, long long o_ld, long long o_bs, int variant,
                    void* stream);
int of_headnorm_bwd(const float* dq, long long dq_ld, long long dq_bs, const float* dk, const float* dv, long long dkv_ld,
                    long long dkv_bs, const void* in16, long long in_ld, long long in_bs, int B, int L, int Hq, int Hk, int Hv, int D,
                    const float* gamma_q, const float* gamma_k, float scale, void* dqkv16, long long o_ld, long long o_bs,
                    float* dgamma_q, float* dgamma_k, int variant, void* stream);
int of_row_mean_std(const float* a, int B, int C, int N, float* out, void* stream);
/* Batched forms (one launch for all samples; strides in elements, `*_ld` of the (B, C) vectors = their row stride):
 *   of_adaln_fwd : out_bf16[b,l,:] = LayerNorm_noaffine(x[b,l,:]) * scale1p[b,:] + shift[b,:]  (scale1p = 1 + scale, dit.py:13-15);
 *                  mean_rstd (B*L, 2) saved for backward.  Bytes/element: 4 read, 2 written.
 *   of_adaln_bwd : dx = LayerNorm backward of dy * scale1p[b,:]  (+ dres, the residual-stream gradient that bypasses the branch);
 *                  dscale[b,:] += sum_l dy * xhat, dshift[b,:] += sum_l dy (atomic; caller zero-fills).  Bytes/element: 8 (+4) read, 4 written.
 *   of_gate_bwd  : backward of `gate.unsqueeze(1) * y` in one pass: dy16 = bf16(gate[b,:] * d), dgate[b,:] += sum_l d * y (atomic), with
 *                  d rounded to bf16 first when round_bf16 (the forward product was a bf16 multiplication).  Bytes/element: 4 + 2 read, 2 written. */
int of_adaln_fwd(const float* x, long long x_ld, long long x_bs, int B, int L, int C, const float* scale1p, long long sc_ld,
                 const float* shift, long long sh_ld, float eps, void* out_bf16, long long o_ld, long long o_bs, float* mean_rstd,
                 void* stream);
int of_adaln_bwd(const float* dy, long long dy_ld, long long dy_bs, const float* x, long long x_ld, long long x_bs, int B, int L, int C,
                 const float* scale1p, long long sc_ld, const float* mean_rstd, const float* dres, long long r_ld, long long r_bs,
                 float* dx, long long dx_ld, long long dx_bs, float* dscale, long long ds_ld, float* dshift, long long dsh_ld,
                 void* stream);
int of_gate_bwd(const float* d32, long long d_ld, long long d_bs, const float* gate, long long gate_ld, const void* y16, long long y_ld,
                long long y_bs, int round_bf16, int B, int L, int C, void* dy16, long long o_ld, long long o_bs, float* dgate,
                long long dg_ld, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OSUFUSION_B200_H_ */
