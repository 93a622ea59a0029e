// of_attn_bwd, version 2: multi-query flash attention backward in the TRANSPOSED formulation (keys on the TMEM lanes).
//
// Why: the version-1 kernel (attn_bwd.cu) is bound by shared-memory bandwidth — ncu: 1664 tensor-core operand wavefronts + 1468
// LSU wavefronts per 128x128 tile against 1280 clk of tensor work (every MMA takes both operands from shared memory, P and dS make
// a round trip through it).  With S^T = K Q^T the probabilities come out as [key lane][query column], which is exactly the layout a
// tcgen05.mma takes its A operand in FROM TENSOR MEMORY for the two key-major products:
//   S^T  = K Q_i^T            (SS)   dP^T = V dO_i^T            (SS)
//   P^T  = exp2(S^T*c - lse_q)                 -> bf16, written over its own S^T columns in TMEM
//   dS^T = P^T o (dP^T - delta_q)              -> bf16, written over its own dP^T columns in TMEM + ONE copy in shared memory
//   dV  += P^T dO_i           (TS)   dK  += dS^T Q_i            (TS)   dQ_i = dS K   (SS, A = the shared copy read MN-major)
// Operand traffic drops from 208 KB to 144 KB per tile and the P round trip (64 KB) disappears.  The softmax scale is applied to dQ
// and dK when they are flushed, not per element.  Per-query statistics (lse, delta) vary along a thread's COLUMNS here: the TMA warp
// brings them into shared memory with the Q / dO tiles (cp.async.bulk on the same mbarrier) and the threads read them as broadcasts.
// A share of the exponentials runs as a polynomial on the FMA pipe (ptx.cuh: exp2_poly2); FFMA2 / FADD2 handle element pairs.
//
// Requires L % 4 == 0 (16-byte bulk copies of the statistics); of_attn_bwd falls back to version 1 otherwise.
#include <stdlib.h>

#include <type_traits>

#include "host_common.h"
#include "ptx.cuh"

namespace ofx {

constexpr int kB2Threads = 64 + 32 * 8;       // warp 0 TMA, warp 1 MMA, warps 2..9 softmax / epilogue (two per TMEM lane quarter)
constexpr uint32_t kT = 128 * 64 * 2;         // 16 KB tile
constexpr int kB2Poly = 8;                    // of the 32 element pairs a thread handles per tile, this many use the polynomial exp2

struct AttnBwd2Params {
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

__device__ __forceinline__ uint32_t mul_bf16x2_(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ void red_add4_(float* ptr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// 32 rows x 32 fp32 columns held one row per lane -> coalesced red.global.add.v4 through a swizzled 4 KB staging tile (see attn_bwd.cu)
__device__ __forceinline__ void red_tile_32x32(uint32_t stg, int lane, const uint32_t (&v)[32], float mul, float* base, long long ld,
                                               int row0, int row_lim, int col_lim) {
#pragma unroll
  for (int g = 0; g < 8; ++g)
    sts128(stg + lane * 128 + ((g ^ (lane & 7)) << 4), __float_as_uint(__uint_as_float(v[4 * g]) * mul),
           __float_as_uint(__uint_as_float(v[4 * g + 1]) * mul), __float_as_uint(__uint_as_float(v[4 * g + 2]) * mul),
           __float_as_uint(__uint_as_float(v[4 * g + 3]) * mul));
  __syncwarp();
  const int rr = lane >> 3, gg = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + rr;
    if (row0 + r < row_lim && gg * 4 < col_lim) {
      const uint4 t = lds128(stg + r * 128 + ((gg ^ (r & 7)) << 4));
      red_add4_(base + (long long)(row0 + r) * ld + gg * 4, __uint_as_float(t.x), __uint_as_float(t.y), __uint_as_float(t.z),
                __uint_as_float(t.w));
    }
  }
  __syncwarp();
}

// TMEM columns: S^T / P^T = 0..127 | dP^T / dS^T = 128..255 | dV = 256 | dK = 320 | dQ = 384.
// A softmax warp of lane quarter qd and part pt (= which 64 query columns) reads S^T / dP^T columns [pt*64, +64) and writes its 32
// packed bf16x2 columns at [pt*96, +32) — inside its own, already consumed range, so the two warps of a quarter never race.
__global__ void __launch_bounds__(kB2Threads, 1)
attn_bwd2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do, const AttnBwd2Params p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + kT;
  uint8_t* sQdO = sV + kT;            // 2 stages x (Q 16 KB + dO 16 KB)
  uint8_t* sdS = sQdO + 4 * kT;       // 32 KB: dS^T as [query half][key row][64 queries] (A operand of dQ = dS K, MN-major)
  uint8_t* sStage = sdS + 2 * kT;     // 8 warps x 4 KB coalescing tiles for the fp32 red.add flushes
  float* sStat = reinterpret_cast<float*>(sStage + 8 * 4096);   // 2 stages x (128 lse | 128 delta)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStat + 2 * 256);
  uint64_t* kv_full = bars;
  uint64_t* qdo_full = bars + 1;    // [2]
  uint64_t* qdo_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* dp_full = bars + 7;
  uint64_t* ds_full = bars + 8;
  uint64_t* dq_full = bars + 9;
  uint64_t* dq_empty = bars + 10;
  uint64_t* acc_done = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int kvh = h % p.KVH;
  const int n = p.n_q_tiles;
  const long long bh = (long long)b * p.H + h;

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
    mbar_init(p_full, 8);
    mbar_init(dp_full, 1);
    mbar_init(ds_full, 8);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 8);
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
  pdl_wait();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tdP = tmem + 128, tdV = tmem + 256, tdK = tmem + 320, tdQ = tmem + 384;

  if (warp == 0) {
    const bool leader = elect_one();
    if (leader) {
      mbar_arrive_expect_tx(kv_full, 2 * kT);
      tma_load_4d(sK, &tmap_k, kv_full, 0, kvh, k0, b);
      tma_load_4d(sV, &tmap_v, kv_full, 0, kvh, k0, b);
    }
    for (int i = 0; i < n; ++i) {
      const int st = i & 1, use = i >> 1;
      mbar_wait(&qdo_empty[st], (use & 1) ^ 1);
      uint8_t* sq = sQdO + st * 2 * kT;
      if (leader) {
        const int nq = min(128, p.L - i * 128);
        const uint32_t sb = (uint32_t)nq * 4u;
        mbar_arrive_expect_tx(&qdo_full[st], 2 * kT + 2 * sb);
        tma_load_4d(sq, &tmap_q, &qdo_full[st], 0, h, i * 128, b);
        tma_load_4d(sq + kT, &tmap_do, &qdo_full[st], 0, h, i * 128, b);
        bulk_load_1d(sStat + st * 256, p.lse + bh * p.L + i * 128, sb, &qdo_full[st]);
        bulk_load_1d(sStat + st * 256 + 128, p.delta + bh * p.L + i * 128, sb, &qdo_full[st]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);    // S^T = K Q^T, dP^T = V dO^T : both operands K-major (d contiguous)
    const uint32_t idesc_ts = make_idesc_bf16(128, 64, 0, 1);    // dV, dK: A from tensor memory, B = dO / Q read MN-major
    const uint32_t idesc_dq = make_idesc_bf16(128, 64, 1, 1);    // dQ = dS K: A = shared dS^T copy (MN-major), B = K (MN-major)
    const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), adS = smem_u32(sdS);
    auto issue_s = [&](int i) {
      const uint32_t aQ = smem_u32(sQdO + (i & 1) * 2 * kT);
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(tS, make_smem_desc(aK + k * 32, 16, 1024), make_smem_desc(aQ + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_full);
      }
      __syncwarp();
    };
    auto issue_dp = [&](int i) {
      const uint32_t adO = smem_u32(sQdO + (i & 1) * 2 * kT) + kT;
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(tdP, make_smem_desc(aV + k * 32, 16, 1024), make_smem_desc(adO + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
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
      const uint32_t aQ = smem_u32(sQdO + st * 2 * kT), adO = aQ + kT;
      // dV += P^T dO   (K dimension = 128 queries, 16 per step; packed P^T columns at [0,32) and [96,128) of the S^T region)
      mbar_wait(p_full, i & 1);
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16_ts(tdV, tS + (k < 4 ? k * 8 : 96 + (k - 4) * 8), make_smem_desc(adO + k * 2048, 8192, 1024), idesc_ts,
                      (i > 0 || k > 0) ? 1u : 0u);
      }
      __syncwarp();
      // S^T(i+1) overwrites P^T(i): issued after the MMAs above (tcgen05.mma execute in issue order)
      if (i + 1 < n) {
        mbar_wait(&qdo_full[(i + 1) & 1], ((i + 1) >> 1) & 1);
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
          umma_f16_ts(tdK, tdP + (k < 4 ? k * 8 : 96 + (k - 4) * 8), make_smem_desc(aQ + k * 2048, 8192, 1024), idesc_ts,
                      (i > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16_ss(tdQ, make_smem_desc(adS + k * 2048, kT, 1024), make_smem_desc(aK + k * 2048, 8192, 1024), idesc_dq, k > 0 ? 1u : 0u);
        umma_commit(dq_full);
        umma_commit(&qdo_empty[st]);
      }
      __syncwarp();
      // dP^T(i+1) overwrites dS^T(i) (in order after dK(i)); its completion also tells the threads that dQ(i) has read the shared copy
      if (i + 1 < n) issue_dp(i + 1);
    }
    if (leader) umma_commit(acc_done);
    __syncwarp();
  } else {
    const int qd = warp & 3;
    const int pt = (warp - 2) >> 2;           // which 64 query columns of S^T / dP^T (and which 32 d-columns of dQ, dK, dV)
    const int row = qd * 32 + lane;           // key row of this thread within the KV tile
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const bool key_ok = k0 + row < p.L;
    const uint32_t stg = smem_u32(sStage + (warp - 2) * 4096);
    const uint32_t ds_row = smem_u32(sdS + pt * kT + row * 128);
    const unsigned long long sc2 = pack_f32x2(p.scale_log2, p.scale_log2);

    auto flush_dq = [&](int i) {   // dQ(i) tile [128 queries][64 d] -> global fp32 atomics (this warp: 32 query rows x 32 d columns)
      mbar_wait(dq_full, i & 1);
      tc_fence_after();
      uint32_t dq[32];
      tmem_ld_32x32b_x32(tdQ + lane_off + pt * 32, dq);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_empty);
      red_tile_32x32(stg, lane, dq, p.scale, p.dq + (long long)b * p.dq_bs + (long long)h * p.D + pt * 32, p.dq_ld, i * 128 + qd * 32, p.L,
                     p.D - pt * 32);
    };

    uint32_t pk[32];
    for (int i = 0; i < n; ++i) {
      const int st = i & 1;
      const float* stat = sStat + st * 256 + pt * 64;             // lse of this thread's 64 query columns; delta at +128
      const int q_valid = p.L - (i * 128 + pt * 64);               // columns [0, q_valid) of this thread's 64 are real queries
      const bool tile_full = q_valid >= 64 && k0 + 128 <= p.L;      // warp-uniform
      mbar_wait(&qdo_full[st], (i >> 1) & 1);                      // statistics of tile i have landed
      // ---- stage A: P^T = exp2(S^T * scale*log2e - lse_q).  Two instantiations: the mask-free one for tiles that lie completely
      // inside the sequence (predicated-off selects / compares still cost issue slots: they were 40 % of the first version's instructions)
      auto stage_a = [&](auto mask_tag) {
        constexpr bool kMask = decltype(mask_tag)::value;
        uint32_t s[64];
        mbar_wait(s_full, i & 1);
        tc_fence_after();
        tmem_ld_32x32b_x32(tS + lane_off + pt * 64, *reinterpret_cast<uint32_t (*)[32]>(&s[0]));
        tmem_ld_32x32b_x32(tS + lane_off + pt * 64 + 32, *reinterpret_cast<uint32_t (*)[32]>(&s[32]));
        tmem_wait_ld();
#pragma unroll
        for (int g = 0; g < 16; ++g) {
          const uint4 ls = lds128(smem_u32(stat + 4 * g));          // broadcast: every lane reads the same 4 lse values
          const unsigned long long lpair[2] = {pack_f32x2(__uint_as_float(ls.x), __uint_as_float(ls.y)),
                                               pack_f32x2(__uint_as_float(ls.z), __uint_as_float(ls.w))};
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = 4 * g + 2 * e;
            const unsigned long long x2 = sub_f32x2(mul_f32x2(pack_f32x2(__uint_as_float(s[c]), __uint_as_float(s[c + 1])), sc2), lpair[e]);
            float p0, p1;
            if ((((c >> 1) + 1) * kB2Poly) / 32 != ((c >> 1) * kB2Poly) / 32) {
              exp2_poly2(x2, p0, p1);
            } else {
              float x0, x1;
              unpack_f32x2(x2, x0, x1);
              p0 = ex2(x0);
              p1 = ex2(x1);
            }
            if (kMask) {
              if (c >= q_valid || !key_ok) p0 = 0.f;
              if (c + 1 >= q_valid || !key_ok) p1 = 0.f;
            }
            pk[c >> 1] = pack_bf16x2(p0, p1);
          }
        }
      };
      if (tile_full) stage_a(std::false_type{});
      else stage_a(std::true_type{});
      {
        uint32_t (&a0)[16] = *reinterpret_cast<uint32_t (*)[16]>(&pk[0]);
        uint32_t (&a1)[16] = *reinterpret_cast<uint32_t (*)[16]>(&pk[16]);
        tmem_st_32x32b_x16(tS + lane_off + pt * 96, a0);
        tmem_st_32x32b_x16(tS + lane_off + pt * 96 + 16, a1);
        tmem_wait_st();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      // ---- stage C of the previous tile (its dK / dQ MMAs ran while P^T was computed)
      if (i > 0) flush_dq(i - 1);
      // ---- stage B: dS^T = P^T o (dP^T - delta_q)      (the softmax scale is applied when dQ / dK are flushed)
      auto stage_b = [&](auto mask_tag) {
        constexpr bool kMask = decltype(mask_tag)::value;
        uint32_t dp[64];
        mbar_wait(dp_full, i & 1);     // also: dQ(i-1) has finished reading the shared dS^T copy
        tc_fence_after();
        tmem_ld_32x32b_x32(tdP + lane_off + pt * 64, *reinterpret_cast<uint32_t (*)[32]>(&dp[0]));
        tmem_ld_32x32b_x32(tdP + lane_off + pt * 64 + 32, *reinterpret_cast<uint32_t (*)[32]>(&dp[32]));
        tmem_wait_ld();
#pragma unroll
        for (int g = 0; g < 16; ++g) {
          const uint4 dl = lds128(smem_u32(stat + 128 + 4 * g));
          const unsigned long long dpair[2] = {pack_f32x2(__uint_as_float(dl.x), __uint_as_float(dl.y)),
                                               pack_f32x2(__uint_as_float(dl.z), __uint_as_float(dl.w))};
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = 4 * g + 2 * e;
            float t0, t1;
            unpack_f32x2(sub_f32x2(pack_f32x2(__uint_as_float(dp[c]), __uint_as_float(dp[c + 1])), dpair[e]), t0, t1);
            if (kMask) {   // invalid query columns carry undefined statistics: P is exactly 0 there, force the factor finite as well
              if (c >= q_valid) t0 = 0.f;
              if (c + 1 >= q_valid) t1 = 0.f;
            }
            pk[c >> 1] = mul_bf16x2_(pk[c >> 1], pack_bf16x2(t0, t1));
          }
        }
      };
      if (tile_full) stage_b(std::false_type{});
      else stage_b(std::true_type{});
      {
        uint32_t (&a0)[16] = *reinterpret_cast<uint32_t (*)[16]>(&pk[0]);
        uint32_t (&a1)[16] = *reinterpret_cast<uint32_t (*)[16]>(&pk[16]);
        tmem_st_32x32b_x16(tdP + lane_off + pt * 96, a0);
        tmem_st_32x32b_x16(tdP + lane_off + pt * 96 + 16, a1);
#pragma unroll
        for (int c = 0; c < 8; ++c)
          sts128(ds_row + ((c ^ (row & 7)) << 4), pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
        tmem_wait_st();
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
    }
    flush_dq(n - 1);
    // ---- dK / dV tiles [128 keys][64 d] -> global (fp32 atomics; 16 q heads add into the shared KV head)
    mbar_wait(acc_done, 0);
    tc_fence_after();
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      uint32_t a[32];
      tmem_ld_32x32b_x32((which == 0 ? tdV : tdK) + lane_off + pt * 32, a);
      tmem_wait_ld();
      red_tile_32x32(stg, lane, a, which == 0 ? 1.0f : p.scale,
                     (which == 0 ? p.dv : p.dk) + (long long)b * p.dkv_bs + (long long)kvh * p.D + pt * 32, p.dkv_ld, k0 + qd * 32, p.L,
                     p.D - pt * 32);
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

int make_head_tmap(CUtensorMap* m, const void* base, int D, int heads, int L, int B, long long ld, long long bs, unsigned box_rows);

// Launch of the version-2 kernel (tensor maps built by the caller, of_attn_bwd in attn_bwd.cu); attn_delta_kernel has run before.
int launch_attn_bwd2(const of_attn_args* a, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
                     cudaStream_t stream) {
  AttnBwd2Params p;
  p.B = a->B; p.H = a->H; p.KVH = a->KVH; p.L = a->L; p.D = a->D;
  p.n_q_tiles = (a->L + 127) / 128;
  p.scale = a->scale > 0.f ? a->scale : 1.0f / sqrtf((float)a->D);
  p.scale_log2 = p.scale * 1.4426950408889634f;
  p.lse = a->lse;
  p.delta = a->delta;
  p.dq = a->dq; p.dq_ld = a->dq_ld; p.dq_bs = a->dq_batch_stride;
  p.dk = a->dk; p.dv = a->dv; p.dkv_ld = a->dkv_ld; p.dkv_bs = a->dkv_batch_stride;
  const size_t smem_bytes = 1024 + kT * 8 + 8 * 4096 + 2 * 256 * 4 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    OF_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    attr_set = true;
  }
  dim3 grid((a->L + 127) / 128, a->H, a->B);
  OF_CHECK_CUDA(launch_pdl<1>(attn_bwd2_kernel, grid, dim3(kB2Threads), smem_bytes, stream, tq, tk, tv, tdo, p));
  return OF_OK;
}

}  // namespace ofx
