// of_attn_bwd: multi-query flash attention backward on tcgen05/TMEM (sm_100a).
//
// Replaces autograd's backward of F.scaled_dot_product_attention (reference attention.py:94-99) and of the GQA
// `repeat` (unet.py:135-137): dK/dV of the single shared KV head are summed over all q heads.
//
// One CTA = one 128-key KV tile of one (batch, q head); it loops over all 128-row Q tiles:
//   S  = Q_i K^T,  dP = dO_i V^T                         (SS MMAs into TMEM)
//   P  = exp2(S*scale*log2e - lse2),  dS = P o (dP - delta) * scale     (one thread per q row, bf16 -> swizzled smem)
//   dV += P^T dO_i,  dK += dS^T Q_i                      (MN-major A straight from the same smem tiles; TMEM accum)
//   dQ_i = dS K                                          (TMEM -> fp32 red.global.add, summed over KV tiles)
// At the end dK/dV tiles are atomically added (fp32) to the shared-KV-head gradient (16 q heads contribute).
#include <stdlib.h>

#include "host_common.h"
#include "ptx.cuh"

namespace ofx {

int make_head_tmap(CUtensorMap* m, const void* base, int D, int heads, int L, int B, long long ld, long long bs,
                   unsigned box_rows);
int launch_attn_bwd2(const of_attn_args* a, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
                     cudaStream_t stream);

// Softmax / epilogue warps: kNP warps per TMEM lane quarter, each owning 128/kNP key columns of S and dP (and 64/kNP d-columns of
// dQ, dK, dV).  kNP = 4 (16 warps, twice the warps per scheduler at half the work per thread) was measured: 1063 us vs 1030 us for
// kNP = 2 at B4 H16 L4096 -- the kernel is not bound by softmax issue slots but by shared-memory bandwidth: every MMA takes both
// operands from shared memory (the N=64 ones need 192 B/clk against 128 B/clk), plus 64 KB of P/dS stores and 64 KB of dQ
// staging per 128x128 tile: ~2.7 k clk of shared-memory time per tile against 1.3 k clk of tensor work.
constexpr int kNP = 2;
constexpr int kCW = 128 / kNP;            // S / dP columns per thread
constexpr int kDW = 64 / kNP;             // dQ / dK / dV columns per thread
constexpr int kSoftWarps = 4 * kNP;
constexpr int kBThreads = 64 + 32 * kSoftWarps;  // warp 0 TMA, warp 1 MMA, then the softmax/epilogue warps
constexpr uint32_t kTile = 128 * 64 * 2;  // 16 KB

struct AttnBwdParams {
  int B, H, KVH, L, D;
  int n_q_tiles;
  float scale, scale_log2;
  const float* lse;
  const float* delta;
  float* dq;
  long long dq_ld, dq_bs;
  float* dk;
  float* dv;
  long long dkv_ld, dkv_bs;
};

__device__ __forceinline__ uint32_t mul_bf16x2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ float ex2b(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void red_add4(float* ptr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// A thread holds one ROW (32 fp32 columns) of a 32x32 accumulator block; adding it to global memory directly makes every
// red.v4 warp instruction touch 32 different 128-byte lines (measured: ~2.8 clk per line on the LSU -- this alone bounded the kernel
// at ~5.7k clk per Q tile).  Staged through a swizzled 4 KB shared tile, one instruction covers 4 full lines.
template <int W>
__device__ __forceinline__ void red_tile_32xW(uint32_t stg, int lane, const uint32_t (&v)[W], float* base, long long ld, int row0,
                                              int row_lim, int col_lim) {
  constexpr int G = W / 4;          // 16-byte groups per row (8 for 32 columns, 4 for 16)
  constexpr int RPI = 32 / G;       // rows covered by one warp instruction
#pragma unroll
  for (int g = 0; g < G; ++g) sts128(stg + lane * 128 + ((g ^ (lane & 7)) << 4), v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
  __syncwarp();
  const int rr = lane / G, gg = lane % G;
#pragma unroll
  for (int i = 0; i < 32 / RPI; ++i) {
    const int r = i * RPI + rr;
    if (row0 + r < row_lim && gg * 4 < col_lim) {
      const uint4 t = lds128(stg + r * 128 + ((gg ^ (r & 7)) << 4));
      red_add4(base + (long long)(row0 + r) * ld + gg * 4, __uint_as_float(t.x), __uint_as_float(t.y), __uint_as_float(t.z), __uint_as_float(t.w));
    }
  }
  __syncwarp();
}
template <int W>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[W]);
template <>
__device__ __forceinline__ void tmem_ld_cols<32>(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld_32x32b_x32(taddr, r); }
template <>
__device__ __forceinline__ void tmem_ld_cols<16>(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld_32x32b_x16(taddr, r); }
template <>
__device__ __forceinline__ void tmem_ld_cols<64>(uint32_t taddr, uint32_t (&r)[64]) {
  tmem_ld_32x32b_x32(taddr, *reinterpret_cast<uint32_t (*)[32]>(&r[0]));
  tmem_ld_32x32b_x32(taddr + 32, *reinterpret_cast<uint32_t (*)[32]>(&r[32]));
}

// Software pipeline (per Q tile i; tensor core and softmax threads overlap):
//   MMA   : dV+=P(i)^T dO | S(i+1)=Q K^T | dK+=dS(i)^T Q, dQ(i)=dS K | dP(i+1)=dO V^T
//   threads: A(i): S -> P (regs + smem) | C(i-1): dQ(i-1) -> global atomics | B(i): dP -> dS (smem)
__global__ void __launch_bounds__(kBThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                const AttnBwdParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + kTile;
  uint8_t* sQdO = sV + kTile;      // 2 stages x (Q 16 KB + dO 16 KB)
  uint8_t* sP = sQdO + 4 * kTile;  // 32 KB
  uint8_t* sdS = sP + 2 * kTile;   // 32 KB
  uint8_t* sStage = sdS + 2 * kTile;   // per softmax warp a 32-row x 128-byte coalescing buffer for the fp32 red.add tiles (dQ, dK, dV)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + kSoftWarps * 4096);
  uint64_t* kv_full = bars;
  uint64_t* qdo_full = bars + 1;    // [2]
  uint64_t* qdo_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* s_empty = bars + 6;
  uint64_t* dp_full = bars + 7;
  uint64_t* dp_empty = bars + 8;
  uint64_t* p_full = bars + 9;
  uint64_t* p_empty = bars + 10;
  uint64_t* ds_full = bars + 11;
  uint64_t* ds_empty = bars + 12;
  uint64_t* dq_full = bars + 13;
  uint64_t* dq_empty = bars + 14;
  uint64_t* acc_done = bars + 15;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int kvh = h % p.KVH;
  const int n = p.n_q_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_do);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qdo_full[i], 1);
      mbar_init(&qdo_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, kSoftWarps);
    mbar_init(dp_full, 1);
    mbar_init(dp_empty, kSoftWarps);
    mbar_init(p_full, kSoftWarps);
    mbar_init(p_empty, 1);
    mbar_init(ds_full, kSoftWarps);
    mbar_init(ds_empty, 1);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, kSoftWarps);
    mbar_init(acc_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();   // the prologue above overlapped the previous kernel's tail
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tdP = tmem + 128, tdV = tmem + 256, tdK = tmem + 320, tdQ = tmem + 384;

  if (warp == 0) {
    const bool leader = elect_one();   // warp-uniform loop; the elected lane issues the TMA instructions
    if (leader) {
      mbar_arrive_expect_tx(kv_full, 2 * kTile);
      tma_load_4d(sK, &tmap_k, kv_full, 0, kvh, k0, b);
      tma_load_4d(sV, &tmap_v, kv_full, 0, kvh, k0, b);
    }
    for (int i = 0; i < n; ++i) {
      const int st = i & 1, use = i >> 1;
      mbar_wait(&qdo_empty[st], (use & 1) ^ 1);
      uint8_t* sq = sQdO + st * 2 * kTile;
      if (leader) {
        mbar_arrive_expect_tx(&qdo_full[st], 2 * kTile);
        tma_load_4d(sq, &tmap_q, &qdo_full[st], 0, h, i * 128, b);
        tma_load_4d(sq + kTile, &tmap_do, &qdo_full[st], 0, h, i * 128, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // warp-uniform MMA loop: only tcgen05.mma / commit are predicated on the elected lane (operands stay in uniform registers)
    const bool leader = elect_one();
    {
      const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_kv = make_idesc_bf16(128, 64, 1, 1);  // dV, dK: A = P^T / dS^T (MN-major), B MN-major
      const uint32_t idesc_dq = make_idesc_bf16(128, 64, 0, 1);  // dQ: A = dS (K-major), B = K (MN-major)
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP), adS = smem_u32(sdS);
      auto issue_s = [&](int i) {
        const uint32_t aQ = smem_u32(sQdO + (i & 1) * 2 * kTile);
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(tS, make_smem_desc(aQ + k * 32, 16, 1024), make_smem_desc(aK + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
          umma_commit(s_full);
        }
        __syncwarp();
      };
      auto issue_dp = [&](int i) {
        const uint32_t adO = smem_u32(sQdO + (i & 1) * 2 * kTile) + kTile;
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(tdP, make_smem_desc(adO + k * 32, 16, 1024), make_smem_desc(aV + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
          umma_commit(dp_full);
        }
        __syncwarp();
      };
      mbar_wait(kv_full, 0);
      mbar_wait(&qdo_full[0], 0);
      tc_fence_after();
      issue_s(0);
      issue_dp(0);
      for (int i = 0; i < n; ++i) {
        const int st = i & 1;
        const uint32_t aQ = smem_u32(sQdO + st * 2 * kTile), adO = aQ + kTile;
        // dV += P^T dO   (K = 128 q rows, 16 per step)
        mbar_wait(p_full, i & 1);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_f16_ss(tdV, make_smem_desc(aP + k * 2048, kTile, 1024), make_smem_desc(adO + k * 2048, 8192, 1024), idesc_kv,
                        (i > 0 || k > 0) ? 1u : 0u);
          umma_commit(p_empty);
        }
        __syncwarp();
        // S(i+1)
        if (i + 1 < n) {
          mbar_wait(&qdo_full[(i + 1) & 1], ((i + 1) >> 1) & 1);
          mbar_wait(s_empty, i & 1);
          tc_fence_after();
          issue_s(i + 1);
        }
        // dK += dS^T Q ; dQ = dS K
        mbar_wait(ds_full, i & 1);
        mbar_wait(dq_empty, (i & 1) ^ 1);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_f16_ss(tdK, make_smem_desc(adS + k * 2048, kTile, 1024), make_smem_desc(aQ + k * 2048, 8192, 1024), idesc_kv,
                        (i > 0 || k > 0) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_f16_ss(tdQ, make_smem_desc(adS + (k >> 2) * kTile + (k & 3) * 32, 16, 1024),
                        make_smem_desc(aK + k * 2048, 8192, 1024), idesc_dq, k > 0 ? 1u : 0u);
          umma_commit(dq_full);
          umma_commit(ds_empty);
          umma_commit(&qdo_empty[st]);
        }
        __syncwarp();
        // dP(i+1)
        if (i + 1 < n) {
          mbar_wait(dp_empty, i & 1);
          tc_fence_after();
          issue_dp(i + 1);
        }
      }
      if (leader) umma_commit(acc_done);
    }
    __syncwarp();
  } else {
    const int qd = warp & 3;
    const int part = (warp - 2) >> 2;     // which kCW keys (S/dP columns) / which kDW d-columns (dQ, dK, dV) this warp owns
    const int row = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const long long bh = (long long)b * p.H + h;
    const int key_base = k0 + part * kCW;
    const uint32_t stg = smem_u32(sStage + (warp - 2) * 4096);
    // P / dS tiles in shared memory: [64-key half][row][128 bytes]; this thread's kCW keys are 16-byte chunks pc0 .. pc0+kCW/8-1
    const int khalf = (part * kCW) / 64, pc0 = ((part * kCW) % 64) / 8;
    constexpr int kChunks = kCW / 8;

    auto flush_dq = [&](int i) {   // stage C: dQ(i) tile -> global fp32 atomics
      mbar_wait(dq_full, i & 1);
      tc_fence_after();
      uint32_t dq[kDW];
      tmem_ld_cols<kDW>(tdQ + lane_off + part * kDW, dq);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_empty);
      // rows i*128 + qd*32 .. +31 of dQ, columns part*kDW .. of head h
      red_tile_32xW<kDW>(stg, lane, dq, p.dq + (long long)b * p.dq_bs + (long long)h * p.D + part * kDW, p.dq_ld, i * 128 + qd * 32, p.L,
                         p.D - part * kDW);
    };

    // per-row softmax statistics are fetched ONE tile ahead: a global load issued at the top of the tile that needs it put a full
    // L2/DRAM round trip on the critical path of every iteration (long-scoreboard was the top stall of the first capture)
    float lse_next = row < p.L ? p.lse[bh * p.L + row] : 0.f;
    float dlt_next = row < p.L ? p.delta[bh * p.L + row] : 0.f;
    for (int i = 0; i < n; ++i) {
      const int qrow = i * 128 + row;
      const bool q_ok = qrow < p.L;
      const bool tile_full = (i * 128 + 128 <= p.L) && (k0 + 128 <= p.L);
      const float lse2 = lse_next, dlt = dlt_next;
      if (i + 1 < n) {
        const int qn = qrow + 128;
        lse_next = qn < p.L ? p.lse[bh * p.L + qn] : 0.f;
        dlt_next = qn < p.L ? p.delta[bh * p.L + qn] : 0.f;
      }
      // ---- stage A: P = exp2(S*scale*log2e - lse2)
      uint32_t pk[kCW / 2];  // kCW probabilities, packed bf16x2
      {
        uint32_t s[kCW];
        mbar_wait(s_full, i & 1);
        tc_fence_after();
        tmem_ld_cols<kCW>(tS + lane_off + part * kCW, s);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty);
        if (tile_full) {   // whole 128x128 tile valid: no per-element masking on the hot path
#pragma unroll
          for (int e = 0; e < kCW / 2; ++e) {
            const float p0 = ex2b(fmaf(__uint_as_float(s[2 * e]), p.scale_log2, -lse2));
            const float p1 = ex2b(fmaf(__uint_as_float(s[2 * e + 1]), p.scale_log2, -lse2));
            pk[e] = pack_bf16x2(p0, p1);
          }
        } else {
#pragma unroll
          for (int e = 0; e < kCW / 2; ++e) {
            float p0 = ex2b(fmaf(__uint_as_float(s[2 * e]), p.scale_log2, -lse2));
            float p1 = ex2b(fmaf(__uint_as_float(s[2 * e + 1]), p.scale_log2, -lse2));
            if (!q_ok || key_base + 2 * e >= p.L) p0 = 0.f;
            if (!q_ok || key_base + 2 * e + 1 >= p.L) p1 = 0.f;
            pk[e] = pack_bf16x2(p0, p1);
          }
        }
      }
      mbar_wait(p_empty, (i & 1) ^ 1);   // dV MMA of the previous tile has finished reading sP
      {
        const uint32_t prow = smem_u32(sP + khalf * kTile + row * 128);
#pragma unroll
        for (int c = 0; c < kChunks; ++c)
          sts128(prow + (((pc0 + c) ^ (row & 7)) << 4), pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      // ---- stage C of the previous tile (its dK/dQ MMAs ran while we computed P)
      if (i > 0) flush_dq(i - 1);
      // ---- stage B: dS = P o (dP - delta) * scale
      {
        uint32_t dp[kCW];
        mbar_wait(dp_full, i & 1);
        tc_fence_after();
        tmem_ld_cols<kCW>(tdP + lane_off + part * kCW, dp);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dp_empty);
        // dS = P o ((dP - delta) * scale): the difference is formed in fp32 (it cancels), rounded to bf16 pairs and multiplied by
        // the bf16 P pair with one packed multiply -- 2 instructions per element instead of 4.5 (unpack, sub, 2 mul, pack).
        const float nds = -dlt * p.scale;
#pragma unroll
        for (int e = 0; e < kCW / 2; ++e) {
          const uint32_t t2 = pack_bf16x2(fmaf(__uint_as_float(dp[2 * e]), p.scale, nds), fmaf(__uint_as_float(dp[2 * e + 1]), p.scale, nds));
          pk[e] = mul_bf16x2(pk[e], t2);
        }
      }
      mbar_wait(ds_empty, (i & 1) ^ 1);  // dK/dQ MMAs of the previous tile have finished reading sdS
      {
        const uint32_t dsrow = smem_u32(sdS + khalf * kTile + row * 128);
#pragma unroll
        for (int c = 0; c < kChunks; ++c)
          sts128(dsrow + (((pc0 + c) ^ (row & 7)) << 4), pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
    }
    flush_dq(n - 1);
    // ---- dK / dV tiles -> global (fp32 atomics; summed over q heads); this warp owns d-columns [part*kDW, +kDW)
    mbar_wait(acc_done, 0);
    tc_fence_after();
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      uint32_t a[kDW];
      tmem_ld_cols<kDW>((which == 0 ? tdV : tdK) + lane_off + part * kDW, a);
      tmem_wait_ld();
      red_tile_32xW<kDW>(stg, lane, a, (which == 0 ? p.dv : p.dk) + (long long)b * p.dkv_bs + (long long)kvh * p.D + part * kDW, p.dkv_ld,
                         k0 + qd * 32, p.L, p.D - part * kDW);
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// delta[b,h,l] = sum_d dO[b,l,h,d] * O[b,l,h,d]; optionally zero-fills the fp32 gradient accumulators of of_attn_bwd in the same pass
// (dq: this warp's own (b, l, h) slice; dk / dv rows of kv head h from the warps with h < KVH) so the caller launches no fill kernels.
// warp = (b, h, 32 consecutive l).  8 lanes x 16 bytes cover one head row (D <= 64): 4 rows per iteration, all 8 iterations' loads
// are issued before any arithmetic; the 32 row sums are transposed onto the lanes with shuffles so delta leaves as one 128-byte store.
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ out, long long out_ld, long long out_bs,
                  const __nv_bfloat16* __restrict__ dout, long long do_ld, long long do_bs, float* __restrict__ delta,
                  float* __restrict__ dq, long long dq_ld, long long dq_bs, float* __restrict__ dk, float* __restrict__ dv,
                  long long dkv_ld, long long dkv_bs, int KVH, int B, int H, int L, int D) {
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int nlb = (L + 31) >> 5;
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= (long long)B * H * nlb) return;
  const int h = (int)(wid % H);
  const int lb = (int)((wid / H) % nlb);
  const int b = (int)(wid / ((long long)H * nlb));
  const int sub = lane >> 3, col = (lane & 7) * 8;
  const bool col_ok = col < D;
  uint4 uo[8], ud[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int l = lb * 32 + it * 4 + sub;
    uo[it] = make_uint4(0, 0, 0, 0);
    ud[it] = make_uint4(0, 0, 0, 0);
    if (l < L && col_ok) {
      uo[it] = *reinterpret_cast<const uint4*>(out + (long long)b * out_bs + (long long)l * out_ld + (long long)h * D + col);
      ud[it] = *reinterpret_cast<const uint4*>(dout + (long long)b * do_bs + (long long)l * do_ld + (long long)h * D + col);
    }
  }
  if (dq != nullptr) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int l = lb * 32 + it * 4 + sub;
      if (l < L && col_ok) {
        float* p = dq + (long long)b * dq_bs + (long long)l * dq_ld + (long long)h * D + col;
        *reinterpret_cast<float4*>(p) = z;
        *reinterpret_cast<float4*>(p + 4) = z;
        if (h < KVH) {
          float* pk = dk + (long long)b * dkv_bs + (long long)l * dkv_ld + (long long)h * D + col;
          float* pv = dv + (long long)b * dkv_bs + (long long)l * dkv_ld + (long long)h * D + col;
          *reinterpret_cast<float4*>(pk) = z;
          *reinterpret_cast<float4*>(pk + 4) = z;
          *reinterpret_cast<float4*>(pv) = z;
          *reinterpret_cast<float4*>(pv + 4) = z;
        }
      }
    }
  }
  float mine = 0.f;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const float2 a0 = unpack_bf16x2(uo[it].x), a1 = unpack_bf16x2(uo[it].y), a2 = unpack_bf16x2(uo[it].z), a3 = unpack_bf16x2(uo[it].w);
    const float2 b0 = unpack_bf16x2(ud[it].x), b1 = unpack_bf16x2(ud[it].y), b2 = unpack_bf16x2(ud[it].z), b3 = unpack_bf16x2(ud[it].w);
    float acc = a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y + a3.x * b3.x + a3.y * b3.y;
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    const float t = __shfl_sync(0xffffffffu, acc, (lane & 3) * 8);     // row it*4 + (lane & 3)
    if ((lane >> 2) == it) mine = t;
  }
  const int l = lb * 32 + lane;
  if (l < L) delta[((long long)b * H + h) * L + l] = mine;
}

}  // namespace ofx

using namespace ofx;

extern "C" int of_attn_bwd(const of_attn_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  OF_REQUIRE(a && a->q && a->k && a->v && a->out && a->dout && a->lse && a->delta && a->dq && a->dk && a->dv,
             "of_attn_bwd: null pointer");
  OF_REQUIRE(a->D >= 8 && a->D <= 64 && a->D % 8 == 0, "of_attn_bwd: head dim %d unsupported", a->D);
  OF_REQUIRE(a->H >= 1 && a->KVH >= 1 && a->H % a->KVH == 0, "of_attn_bwd: bad head counts");
  OF_REQUIRE(a->dq_ld % 4 == 0 && a->dkv_ld % 4 == 0, "of_attn_bwd: fp32 gradient lds must be multiples of 4");
  CUtensorMap tq, tk, tv, tdo;
  int rc;
  if ((rc = make_head_tmap(&tq, a->q, a->D, a->H, a->L, a->B, a->q_ld, a->q_batch_stride, 128)) != OF_OK) return rc;
  if ((rc = make_head_tmap(&tk, a->k, a->D, a->KVH, a->L, a->B, a->kv_ld, a->kv_batch_stride, 128)) != OF_OK) return rc;
  if ((rc = make_head_tmap(&tv, a->v, a->D, a->KVH, a->L, a->B, a->kv_ld, a->kv_batch_stride, 128)) != OF_OK) return rc;
  if ((rc = make_head_tmap(&tdo, a->dout, a->D, a->H, a->L, a->B, a->dout_ld, a->dout_batch_stride, 128)) != OF_OK)
    return rc;
  {
    const long long warps = (long long)a->B * a->H * ((a->L + 31) / 32);
    const int threads = 256;
    const long long blocks = (warps + 7) / 8;
    const bool zf = a->zero_grads != 0;
    if (zf) OF_REQUIRE(a->dq_ld % 4 == 0 && a->dq_batch_stride % 4 == 0 && a->dkv_batch_stride % 4 == 0 && a->D % 8 == 0,
                       "of_attn_bwd: zero_grads needs 16-byte aligned gradient rows");
    attn_delta_kernel<<<(unsigned)blocks, threads, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(a->out), a->out_ld, a->out_batch_stride,
        reinterpret_cast<const __nv_bfloat16*>(a->dout), a->dout_ld, a->dout_batch_stride, a->delta, zf ? a->dq : nullptr, a->dq_ld,
        a->dq_batch_stride, a->dk, a->dv, a->dkv_ld, a->dkv_batch_stride, a->KVH, a->B, a->H, a->L, a->D);
    OF_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }
  {
    // version 2 (attn_bwd2.cu: transposed formulation, P^T / dS^T as tensor-memory operands) unless disabled or L % 4 != 0;
    // variant 1 / OF_ATTN_BWD_VARIANT=1 selects the version-1 kernel below
    static const int dflt = [] { const char* e = getenv("OF_ATTN_BWD_VARIANT"); return e ? atoi(e) : 0; }();
    const int v = a->variant ? a->variant : dflt;
    if (v != 1 && a->L % 4 == 0) {
      int rc2 = launch_attn_bwd2(a, tq, tk, tv, tdo, stream);
      if (rc2 != OF_OK) return rc2;
      OF_CHECK_CUDA(cudaGetLastError());
      count_launch();
      return OF_OK;
    }
  }
  AttnBwdParams p;
  p.B = a->B; p.H = a->H; p.KVH = a->KVH; p.L = a->L; p.D = a->D;
  p.n_q_tiles = (a->L + 127) / 128;
  p.scale = a->scale > 0.f ? a->scale : 1.0f / sqrtf((float)a->D);
  p.scale_log2 = p.scale * 1.4426950408889634f;
  p.lse = a->lse;
  p.delta = a->delta;
  p.dq = a->dq; p.dq_ld = a->dq_ld; p.dq_bs = a->dq_batch_stride;
  p.dk = a->dk; p.dv = a->dv; p.dkv_ld = a->dkv_ld; p.dkv_bs = a->dkv_batch_stride;
  size_t smem_bytes = 1024 + kTile * 10 + kSoftWarps * 4096 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    OF_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  dim3 grid((a->L + 127) / 128, a->H, a->B);
  OF_CHECK_CUDA(launch_pdl<1>(attn_bwd_kernel, grid, dim3(kBThreads), smem_bytes, stream, tq, tk, tv, tdo, p));
  OF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OF_OK;
}
