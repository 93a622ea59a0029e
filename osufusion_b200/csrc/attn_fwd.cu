// of_attn_fwd: multi-query flash attention forward on tcgen05/TMEM (sm_100a).
//
// Replaces `Attend.forward` (reference osu_fusion/modules/attention.py:77-101: bf16 q,k,v, non-causal, no mask,
// scale 1/sqrt(D)) together with the GQA `repeat` of k,v (unet.py:135-137), which is never materialised here: every
// q head reads the single shared KV head through its own TMA descriptor coordinates.
//
// One CTA = TWO 128-row Q tiles of one (batch, head) ("ping-pong"); KV tiles of 128 keys stream through a shared TMA ring.
//   warp 0    : TMA producer (both Q tiles once; K/V ring, 3 stages)
//   warp 1    : MMA issuer: S_w(j) = Q_w K_j^T (SS, M128 N128 K64) into the S region of group w; O_w += P_w(j) V_j
//               (P from TMEM [TS mode] or from 128B-swizzled smem, V consumed MN-major straight from its row-major tile)
//   warps 2-5 : softmax group 0, warps 6-9: softmax group 1 — one thread per query row, two passes over S in tensor memory
//               (row max, then exp2 / row sum / bf16 P), lazy rescale of the TMEM O accumulator, final O/l and log-sum-exp.
//   While one group runs its softmax the tensor core works for the other, hiding mbarrier / TMEM round trips.
// Head dim D <= 64 (zero-padded to 64 by TMA out-of-bounds fill); any L (key tail masked).
#include <stdlib.h>

#include "host_common.h"
#include "ptx.cuh"

namespace ofx {

constexpr int kAThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2..5 softmax group 0, warps 6..9 softmax group 1
constexpr int kKvStages = 6;
constexpr int kKvTile = 64;                        // keys per KV tile
constexpr uint32_t kTileBytes = 128 * 64 * 2;      // 16 KB: 128 rows x 64 bf16 (Q tile)
constexpr uint32_t kKvBytes = kKvTile * 64 * 2;    // 8 KB: 64 keys x 64 bf16 (K or V tile)

struct AttnFwdParams {
  int B, H, KVH, L, D;
  int n_kv_tiles;
  float scale_log2;
  int p_in_tmem;
  __nv_bfloat16* out;
  long long out_ld, out_bs;
  float* lse;  // (B, H, L) log2-domain log-sum-exp
};

// One CTA owns TWO 128-row Q tiles (same batch / head), one softmax warpgroup each; K/V stream in 64-key tiles.  With the small
// KV tile, S and P are DOUBLE-buffered per group inside the 512 TMEM columns, so S_w(j+1), S_w(j+2) are computed while group w
// still works on tile j and the softmax threads never wait for the tensor core (nor the tensor core for them):
//   TMEM columns: S[w][b] = w*128 + b*64 (0..255) | O[w] = 256 + w*64 | P[w][b] = 384 + w*64 + b*32 (bf16x2 packed)
// kPoly: how many of the 32 element pairs of a row of a KV tile take the polynomial exp2 (the rest use MUFU).
// kAlt : the two softmax groups take turns in their exp phase (named barriers 2 / 3), so the MUFU unit of every SM sub-partition
//        serves one warp at a time while the other warp of that sub-partition does its MUFU-free work (TMEM load, row max, P store,
//        barrier traffic): left alone the symmetric groups fall into lockstep and the MUFU idles ~40 % of the time.
template <int kPoly, bool kAlt>
__global__ void __launch_bounds__(kAThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const AttnFwdParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                  // 2 x 16 KB
  uint8_t* sKV = sQ + 2 * kTileBytes;                  // kKvStages x (K 8 KB + V 8 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + kKvStages * 2 * kKvBytes);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;               // [kKvStages]
  uint64_t* kv_empty = kv_full + kKvStages;   // [kKvStages]
  uint64_t* s_full = kv_empty + kKvStages;    // [group][buf] = [4]
  uint64_t* s_empty = s_full + 4;             // [4]
  uint64_t* p_full = s_empty + 4;             // [4]
  uint64_t* p_empty = p_full + 4;             // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_empty + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256, h = blockIdx.y, b = blockIdx.z;
  const int n = p.n_kv_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < kKvStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&p_empty[i], 1);
    }
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

  if (warp == 0) {
    const bool leader = elect_one();
    const int kvh = h % p.KVH;
    if (leader) {
      mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
      tma_load_4d(sQ, &tmap_q, q_full, 0, h, q0, b);
      tma_load_4d(sQ + kTileBytes, &tmap_q, q_full, 0, h, q0 + 128, b);
    }
    for (int j = 0; j < n; ++j) {
      const int st = j % kKvStages, use = j / kKvStages;
      mbar_wait(&kv_empty[st], (use & 1) ^ 1);
      uint8_t* sk = sKV + st * 2 * kKvBytes;
      if (leader) {
        mbar_arrive_expect_tx(&kv_full[st], 2 * kKvBytes);
        tma_load_4d(sk, &tmap_k, &kv_full[st], 0, kvh, j * kKvTile, b);
        tma_load_4d(sk + kKvBytes, &tmap_v, &kv_full[st], 0, kvh, j * kKvTile, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // The whole warp runs the loop convergently (mbarrier waits, address arithmetic stay warp-uniform, so the tensor-core
    // operands live in uniform registers); only the tcgen05.mma / commit instructions are predicated on one elected lane.
    // (Issuing from a divergent `if (lane == 0)` region makes ptxas wrap every UTCHMMA in an ELECT / BRA.U.ANY waterfall loop,
    // ~100 cycles per MMA — enough to starve the softmax groups.)
    const bool leader = elect_one();
    {
      const uint32_t idesc_s = make_idesc_bf16(128, kKvTile, 0, 0);
      const uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);
      auto issue_s = [&](int w, int j) {   // S_w(j) = Q_w K_j^T into buffer j&1 (must have been drained by the softmax group)
        const int bf = j & 1;
        const uint32_t aQ = smem_u32(sQ + w * kTileBytes);
        const uint32_t aK = smem_u32(sKV + (j % kKvStages) * 2 * kKvBytes);
        mbar_wait(&s_empty[w * 2 + bf], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(tmem + w * 128 + bf * 64, make_smem_desc(aQ + k * 32, 16, 1024), make_smem_desc(aK + k * 32, 16, 1024),
                        idesc_s, k > 0 ? 1u : 0u);
          umma_commit(&s_full[w * 2 + bf]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int w, int j) {  // O_w += P_w(j) V_j   (P from tensor memory, V MN-major from its row-major tile)
        const int bf = j & 1;
        const uint32_t aV = smem_u32(sKV + (j % kKvStages) * 2 * kKvBytes + kKvBytes);
        const uint32_t tO = tmem + 256 + w * 64, tP = tmem + 384 + w * 64 + bf * 32;
        mbar_wait(&p_full[w * 2 + bf], (j >> 1) & 1);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int k = 0; k < kKvTile / 16; ++k)
            umma_f16_ts(tO, tP + k * 8, make_smem_desc(aV + k * 2048, 8192, 1024), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
          umma_commit(&p_empty[w * 2 + bf]);
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      for (int j = 0; j < 2 && j < n; ++j) {
        mbar_wait(&kv_full[j % kKvStages], 0);
        issue_s(0, j);
        issue_s(1, j);
      }
      for (int j = 0; j < n; ++j) {
        const bool more = j + 2 < n;
        if (more) mbar_wait(&kv_full[(j + 2) % kKvStages], ((j + 2) / kKvStages) & 1);
        issue_pv(0, j);
        if (more) issue_s(0, j + 2);
        issue_pv(1, j);
        if (leader) umma_commit(&kv_empty[j % kKvStages]);   // K_j and V_j are no longer needed by either group
        __syncwarp();
        if (more) issue_s(1, j + 2);
      }
    }
    __syncwarp();
  } else {
    const int w = (warp - 2) >> 2;        // softmax group = Q tile index
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const uint32_t tO = tmem + 256 + w * 64;
    float m_run = -INFINITY, m_used = -INFINITY, l = 0.f;
    if (kAlt && w == 1) named_bar_arrive(2, 256);   // group 0 takes the first turn
    for (int j = 0; j < n; ++j) {
      const int bf = j & 1;
      const uint32_t tS = tmem + w * 128 + bf * 64, tP = tmem + 384 + w * 64 + bf * 32;
      mbar_wait(&s_full[w * 2 + bf], (j >> 1) & 1);
      tc_fence_after();
      uint32_t s[64];
      {
        uint32_t (&s0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[0]);
        uint32_t (&s1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[32]);
        tmem_ld_32x32b_x32(tS + lane_off, s0);
        tmem_ld_32x32b_x32(tS + lane_off + 32, s1);
        tmem_wait_ld();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[w * 2 + bf]);   // S is in registers: the buffer can take S(j+2)
      const int kv_valid = p.L - j * kKvTile;
      const bool full = kv_valid >= kKvTile;
      float mx;
      if (full) {
        float m0 = fmaxf(__uint_as_float(s[0]), __uint_as_float(s[1])), m1 = fmaxf(__uint_as_float(s[2]), __uint_as_float(s[3]));
#pragma unroll
        for (int i = 4; i < 64; i += 4) {
          m0 = fmaxf(m0, fmaxf(__uint_as_float(s[i]), __uint_as_float(s[i + 1])));
          m1 = fmaxf(m1, fmaxf(__uint_as_float(s[i + 2]), __uint_as_float(s[i + 3])));
        }
        mx = fmaxf(m0, m1);
      } else {
        mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i < kv_valid) mx = fmaxf(mx, __uint_as_float(s[i]));
      }
      m_run = fmaxf(m_run, mx * p.scale_log2);
      if (j > 0) {
        const bool need = (m_run - m_used) > 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          // rare: rescale the O accumulator; every PV issued so far must have completed (they complete in order)
          mbar_wait(&p_empty[w * 2 + ((j - 1) & 1)], ((j - 1) >> 1) & 1);
          tc_fence_after();
          const float f = ex2(m_used - m_run);
          m_used = m_run;
          l *= f;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(tO + lane_off + c * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            uint32_t (&o0)[16] = *reinterpret_cast<uint32_t (*)[16]>(&o[0]);
            uint32_t (&o1)[16] = *reinterpret_cast<uint32_t (*)[16]>(&o[16]);
            tmem_st_32x32b_x16(tO + lane_off + c * 32, o0);
            tmem_st_32x32b_x16(tO + lane_off + c * 32 + 16, o1);
          }
          tmem_wait_st();
        }
      } else {
        m_used = m_run;
      }
      uint32_t pk[32];
      if (kAlt) named_bar_sync(2 + w, 256);        // my turn on the MUFU
      if (full) {
        const unsigned long long sc2 = pack_f32x2(p.scale_log2, p.scale_log2), nm2 = pack_f32x2(-m_used, -m_used);
        unsigned long long sum2 = 0ull;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const unsigned long long x2 = fma_f32x2(pack_f32x2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), sc2, nm2);
          float p0, p1;
          if (((i + 1) * kPoly) / 32 != (i * kPoly) / 32) {     // kPoly of the 32 pairs, evenly interleaved with the MUFU pairs
            exp2_poly2(x2, p0, p1);
          } else {
            float x0, x1;
            unpack_f32x2(x2, x0, x1);
            p0 = ex2(x0);
            p1 = ex2(x1);
          }
          sum2 = add_f32x2(sum2, pack_f32x2(p0, p1));
          pk[i] = pack_bf16x2(p0, p1);
        }
        float sa, sb;
        unpack_f32x2(sum2, sa, sb);
        l += sa + sb;
      } else {
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float p0 = ex2(fmaf(__uint_as_float(s[2 * i]), p.scale_log2, -m_used));
          float p1 = ex2(fmaf(__uint_as_float(s[2 * i + 1]), p.scale_log2, -m_used));
          if (2 * i >= kv_valid) p0 = 0.f;
          if (2 * i + 1 >= kv_valid) p1 = 0.f;
          sum += p0 + p1;
          pk[i] = pack_bf16x2(p0, p1);
        }
        l += sum;
      }
      if (kAlt && !(w == 1 && j == n - 1)) named_bar_arrive(2 + (w ^ 1), 256);   // hand the MUFU to the other group
      mbar_wait(&p_empty[w * 2 + bf], ((j >> 1) & 1) ^ 1);   // P V of tile j-2 has finished reading this P buffer
      tc_fence_after();
      {
        uint32_t (&p0)[16] = *reinterpret_cast<uint32_t (*)[16]>(&pk[0]);
        uint32_t (&p1)[16] = *reinterpret_cast<uint32_t (*)[16]>(&pk[16]);
        tmem_st_32x32b_x16(tP + lane_off, p0);
        tmem_st_32x32b_x16(tP + lane_off + 16, p1);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[w * 2 + bf]);
    }
    // ---- epilogue: wait for the last P V of this group
    mbar_wait(&p_empty[w * 2 + ((n - 1) & 1)], ((n - 1) >> 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const int qrow = q0 + w * 128 + row;
    uint32_t o[64];
    {
      uint32_t (&o0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&o[0]);
      uint32_t (&o1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&o[32]);
      tmem_ld_32x32b_x32(tO + lane_off, o0);
      tmem_ld_32x32b_x32(tO + lane_off + 32, o1);
      tmem_wait_ld();
    }
    if (qrow < p.L) {
      __nv_bfloat16* dst = p.out + (long long)b * p.out_bs + (long long)qrow * p.out_ld + (long long)h * p.D;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        if (g * 8 < p.D) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
          v.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
          v.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
          v.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + g * 8) = v;
        }
      }
      if (p.lse) p.lse[((long long)b * p.H + h) * p.L + qrow] = m_used + log2f(l);
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------------------------------
// Version 3: TWO CTAs per SM.  Measured on the B200 (tools/ubench/ubench_sm.cu): one warp can issue a MUFU.EX2 only every ~12 clk,
// two warps per SM sub-partition reach 8.8 clk, four reach the unit's 8 clk — and the serial, latency-bound sections of a softmax
// warp (TMEM load, row-max chain, TMEM store, barrier round trips, ~400 clk per tile) leave the MUFU idle unless other warps of the
// same sub-partition have exp work.  The kernel above has 2 softmax warps per sub-partition (MUFU 58 % busy under ncu); this one
// halves every per-CTA resource so that two CTAs (4 softmax warps per sub-partition) share an SM:
//   * 256 TMEM columns: S is single-buffered and P (bf16x2) is written over the first 32 columns of its own S tile — S_w(j+1) is
//     issued right after P_w(j) V_j (tcgen05.mma execute in issue order), so neither `s_empty` nor `p_empty` barriers exist;
//     layout: S/P[w] = w*64 | O[w] = 128 + w*64
//   * <= 102 registers: the softmax makes two passes over S in tensor memory (row max; then exp / sum / pack in two 32-column halves)
//   * 4 K/V stages (96 KB of shared memory per CTA).
// Tried and dropped (B4 L4096, 322 us for this kernel): three groups in one CTA with a separate P region so that S_w(j+1) overlaps
// the softmax of tile j (480 TMEM columns, event-driven MMA issue): 413-424 us — a group's tile takes ~2.7 k clk in both layouts
// (a latency chain of TMEM round trips, row-max / sum chains and barrier hand-offs), so what counts is how many groups share an SM.
constexpr int kKvStages2 = 4;

template <int kPoly>
__global__ void __launch_bounds__(kAThreads, 2)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v, const AttnFwdParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                  // 2 x 16 KB
  uint8_t* sKV = sQ + 2 * kTileBytes;                  // kKvStages2 x (K 8 KB + V 8 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + kKvStages2 * 2 * kKvBytes);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                // [kKvStages2]
  uint64_t* kv_empty = kv_full + kKvStages2;   // [kKvStages2]
  uint64_t* s_full = kv_empty + kKvStages2;    // [group]
  uint64_t* p_full = s_full + 2;               // [group]
  uint64_t* o_full = p_full + 2;               // [group]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256, h = blockIdx.y, b = blockIdx.z;
  const int n = p.n_kv_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < kKvStages2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    const bool leader = elect_one();
    const int kvh = h % p.KVH;
    if (leader) {
      mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
      tma_load_4d(sQ, &tmap_q, q_full, 0, h, q0, b);
      tma_load_4d(sQ + kTileBytes, &tmap_q, q_full, 0, h, q0 + 128, b);
    }
    for (int j = 0; j < n; ++j) {
      const int st = j % kKvStages2, use = j / kKvStages2;
      mbar_wait(&kv_empty[st], (use & 1) ^ 1);
      uint8_t* sk = sKV + st * 2 * kKvBytes;
      if (leader) {
        mbar_arrive_expect_tx(&kv_full[st], 2 * kKvBytes);
        tma_load_4d(sk, &tmap_k, &kv_full[st], 0, kvh, j * kKvTile, b);
        tma_load_4d(sk + kKvBytes, &tmap_v, &kv_full[st], 0, kvh, j * kKvTile, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc_s = make_idesc_bf16(128, kKvTile, 0, 0);
    const uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);
    auto issue_s = [&](int w, int j) {     // S_w(j) = Q_w K_j^T into the group's (single) S region, then signal the softmax group
      const uint32_t aQ = smem_u32(sQ + w * kTileBytes);
      const uint32_t aK = smem_u32(sKV + (j % kKvStages2) * 2 * kKvBytes);
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(tmem + w * 64, make_smem_desc(aQ + k * 32, 16, 1024), make_smem_desc(aK + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&s_full[w]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    issue_s(0, 0);
    issue_s(1, 0);
    for (int j = 0; j < n; ++j) {
      const bool more = j + 1 < n;
      if (more) mbar_wait(&kv_full[(j + 1) % kKvStages2], ((j + 1) / kKvStages2) & 1);
      const uint32_t aV = smem_u32(sKV + (j % kKvStages2) * 2 * kKvBytes + kKvBytes);
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        mbar_wait(&p_full[w], j & 1);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int k = 0; k < kKvTile / 16; ++k)     // O_w += P_w(j) V_j : P from tensor memory (over its own S tile), V MN-major
            umma_f16_ts(tmem + 128 + w * 64, tmem + w * 64 + k * 8, make_smem_desc(aV + k * 2048, 8192, 1024), idesc_pv,
                        (j > 0 || k > 0) ? 1u : 0u);
          if (!more) umma_commit(&o_full[w]);
        }
        __syncwarp();
        if (more) issue_s(w, j + 1);                // overwrites P_w(j): ordered after the MMAs above (in-order tensor pipe)
      }
      if (leader) umma_commit(&kv_empty[j % kKvStages2]);
      __syncwarp();
    }
  } else {
    const int w = (warp - 2) >> 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const uint32_t tS = tmem + w * 64 + lane_off, tO = tmem + 128 + w * 64 + lane_off;
    float m_run = -INFINITY, m_used = -INFINITY, l = 0.f;
    for (int j = 0; j < n; ++j) {
      mbar_wait(&s_full[w], j & 1);   // also implies P_w(j-1) V_(j-1) has completed (commit covers every earlier MMA)
      tc_fence_after();
      const int kv_valid = p.L - j * kKvTile;
      const bool full = kv_valid >= kKvTile;
      // ---- pass 1: row max
      float mx;
      {
        uint32_t a[32], c[32];
        tmem_ld_32x32b_x32(tS, a);
        tmem_ld_32x32b_x32(tS + 32, c);
        tmem_wait_ld();
        if (full) {
          float m0 = fmaxf(__uint_as_float(a[0]), __uint_as_float(a[1])), m1 = fmaxf(__uint_as_float(c[0]), __uint_as_float(c[1]));
#pragma unroll
          for (int i = 2; i < 32; i += 2) {
            m0 = fmaxf(m0, fmaxf(__uint_as_float(a[i]), __uint_as_float(a[i + 1])));
            m1 = fmaxf(m1, fmaxf(__uint_as_float(c[i]), __uint_as_float(c[i + 1])));
          }
          mx = fmaxf(m0, m1);
        } else {
          mx = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i < kv_valid) mx = fmaxf(mx, __uint_as_float(a[i]));
            if (32 + i < kv_valid) mx = fmaxf(mx, __uint_as_float(c[i]));
          }
        }
      }
      m_run = fmaxf(m_run, mx * p.scale_log2);
      if (j > 0) {
        const bool need = (m_run - m_used) > 8.0f;
        if (__any_sync(0xffffffffu, need)) {   // rare: rescale O (every P V issued so far has completed, see above)
          const float f = ex2(m_used - m_run);
          m_used = m_run;
          l *= f;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(tO + c * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            uint32_t (&o0)[16] = *reinterpret_cast<uint32_t (*)[16]>(&o[0]);
            uint32_t (&o1)[16] = *reinterpret_cast<uint32_t (*)[16]>(&o[16]);
            tmem_st_32x32b_x16(tO + c * 32, o0);
            tmem_st_32x32b_x16(tO + c * 32 + 16, o1);
          }
          tmem_wait_st();
        }
      } else {
        m_used = m_run;
      }
      // ---- pass 2: P = exp2(S*scale - m_used) in two 32-column halves; the bf16x2 pairs go over the S columns already consumed
      const unsigned long long sc2 = pack_f32x2(p.scale_log2, p.scale_log2), nm2 = pack_f32x2(-m_used, -m_used);
      unsigned long long sum2 = 0ull;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t s[32];
        tmem_ld_32x32b_x32(tS + hh * 32, s);
        tmem_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const unsigned long long x2 = fma_f32x2(pack_f32x2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), sc2, nm2);
          float p0, p1;
          if (((i + 1) * kPoly) / 16 != (i * kPoly) / 16) {     // kPoly of the 16 pairs of a half on the FMA pipe
            exp2_poly2(x2, p0, p1);
          } else {
            float x0, x1;
            unpack_f32x2(x2, x0, x1);
            p0 = ex2(x0);
            p1 = ex2(x1);
          }
          if (!full) {
            if (hh * 32 + 2 * i >= kv_valid) p0 = 0.f;
            if (hh * 32 + 2 * i + 1 >= kv_valid) p1 = 0.f;
          }
          sum2 = add_f32x2(sum2, pack_f32x2(p0, p1));
          pk[i] = pack_bf16x2(p0, p1);
        }
        tmem_st_32x32b_x16(tS + hh * 16, pk);
      }
      {
        float sa, sb;
        unpack_f32x2(sum2, sa, sb);
        l += sa + sb;
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[w]);
    }
    // ---- epilogue
    mbar_wait(&o_full[w], 0);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const int qrow = q0 + w * 128 + row;
    __nv_bfloat16* dst = p.out + (long long)b * p.out_bs + (long long)qrow * p.out_ld + (long long)h * p.D;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(tO + c * 32, o);
      tmem_wait_ld();
      if (qrow < p.L) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (c * 32 + g * 8 < p.D) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
            v.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
            v.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
            v.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = v;
          }
        }
      }
    }
    if (qrow < p.L && p.lse) p.lse[((long long)b * p.H + h) * p.L + qrow] = m_used + log2f(l);
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

// (D, heads, L, B) view of a (B, L, heads*D [+ more]) channels-last buffer; box (64, 1, 128, 1).
int make_head_tmap(CUtensorMap* m, const void* base, int D, int heads, int L, int B, long long ld, long long bs,
                   unsigned box_rows) {
  unsigned long long dims[4] = {(unsigned long long)D, (unsigned long long)heads, (unsigned long long)L,
                                (unsigned long long)B};
  unsigned long long bstride = B > 1 ? (unsigned long long)bs : (unsigned long long)L * ld;
  unsigned long long hstride = heads > 1 ? (unsigned long long)D : (unsigned long long)8;
  unsigned long long str[3] = {hstride * 2ull, (unsigned long long)ld * 2ull, bstride * 2ull};
  unsigned box[4] = {64, 1, box_rows, 1};
  return make_tmap_bf16(m, base, 4, dims, str, box);
}

}  // namespace ofx

using namespace ofx;

extern "C" int of_attn_fwd(const of_attn_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  OF_REQUIRE(a && a->q && a->k && a->v && a->out, "of_attn_fwd: null pointer");
  OF_REQUIRE(a->D >= 8 && a->D <= 64 && a->D % 8 == 0, "of_attn_fwd: head dim %d unsupported (8..64, multiple of 8)", a->D);
  OF_REQUIRE(a->H >= 1 && a->KVH >= 1 && a->H % a->KVH == 0, "of_attn_fwd: bad head counts");
  OF_REQUIRE(a->B >= 1 && a->L >= 1, "of_attn_fwd: bad sizes");
  OF_REQUIRE(a->out_ld % 8 == 0, "of_attn_fwd: out_ld %% 8");
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_head_tmap(&tq, a->q, a->D, a->H, a->L, a->B, a->q_ld, a->q_batch_stride, 128)) != OF_OK) return rc;
  if ((rc = make_head_tmap(&tk, a->k, a->D, a->KVH, a->L, a->B, a->kv_ld, a->kv_batch_stride, kKvTile)) != OF_OK) return rc;
  if ((rc = make_head_tmap(&tv, a->v, a->D, a->KVH, a->L, a->B, a->kv_ld, a->kv_batch_stride, kKvTile)) != OF_OK) return rc;
  AttnFwdParams p;
  p.B = a->B; p.H = a->H; p.KVH = a->KVH; p.L = a->L; p.D = a->D;
  p.n_kv_tiles = (a->L + kKvTile - 1) / kKvTile;
  p.scale_log2 = (a->scale > 0.f ? a->scale : 1.0f / sqrtf((float)a->D)) * 1.4426950408889634f;
  p.p_in_tmem = a->variant == 1 ? 0 : 1;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  p.out_ld = a->out_ld;
  p.out_bs = a->out_batch_stride;
  p.lse = a->lse;
  size_t smem_bytes = 1024 + 2 * kTileBytes + kKvStages * 2 * kKvBytes + 512;
  dim3 grid((a->L + 255) / 256, a->H, a->B);
  // variant (tuning / A-B probes; 0 = default): 10 + k -> k*4 polynomial pairs of 32 with alternation; 2 = round-1 schedule (no
  // alternation, MUFU only)
  //          30 + k -> version 3 (two CTAs per SM, see attn_fwd2_kernel) with 2k polynomial pairs of 16 per half row.
  // Measured (B200, H16 KVH1 D64, tools/probe_attn_variants.py): B4 L4096 439 us (round 1) -> 377 us (13) -> 318 us (33);
  // B1 L32768 6.52 ms -> 5.2 ms; grids of <= one CTA per SM (B2 L1024: 128 CTAs) are faster with one CTA per SM (20 us vs 23 us).
  int v = a->variant;
  if (v == 0) {
    static const int dflt = [] { const char* e = getenv("OF_ATTN_FWD_VARIANT"); return e ? atoi(e) : 0; }();
    v = dflt;
    if (v == 0) v = ((long long)grid.x * grid.y * grid.z > (long long)device_sm_count()) ? 33 : 13;
  }
#define OF_FWD_CASE(V, POLY, ALT)                                                                                              \
  if (v == (V)) {                                                                                                               \
    static bool attr_set = false;                                                                                               \
    if (!attr_set) {                                                                                                            \
      OF_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<POLY, ALT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
      attr_set = true;                                                                                                          \
    }                                                                                                                           \
    OF_CHECK_CUDA(launch_pdl<1>(attn_fwd_kernel<POLY, ALT>, grid, dim3(kAThreads), smem_bytes, stream, tq, tk, tv, p));         \
  } else
  OF_FWD_CASE(1, 0, false)
  OF_FWD_CASE(2, 0, false)
  OF_FWD_CASE(10, 0, true)
  OF_FWD_CASE(11, 4, true)
  OF_FWD_CASE(12, 8, true)
  OF_FWD_CASE(13, 12, true)
  OF_FWD_CASE(14, 16, true)
  OF_FWD_CASE(21, 4, false)
  OF_FWD_CASE(22, 8, false)
#define OF_FWD2_CASE(V, POLY)                                                                                                  \
  if (v == (V)) {                                                                                                               \
    static bool attr_set = false;                                                                                               \
    const size_t smem2 = 1024 + 2 * kTileBytes + kKvStages2 * 2 * kKvBytes + 256;                                               \
    if (!attr_set) {                                                                                                            \
      OF_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd2_kernel<POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));     \
      attr_set = true;                                                                                                          \
    }                                                                                                                           \
    OF_CHECK_CUDA(launch_pdl<1>(attn_fwd2_kernel<POLY>, grid, dim3(kAThreads), smem2, stream, tq, tk, tv, p));                  \
  } else
  OF_FWD2_CASE(30, 0)
  OF_FWD2_CASE(31, 2)
  OF_FWD2_CASE(32, 4)
  OF_FWD2_CASE(33, 6)
  OF_FWD2_CASE(34, 8)
  { OF_REQUIRE(false, "of_attn_fwd: unknown variant %d", v); }
#undef OF_FWD2_CASE
#undef OF_FWD_CASE
  OF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OF_OK;
}
