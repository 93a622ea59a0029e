// of_gemm: persistent, warp-specialised tcgen05 GEMM / implicit-GEMM conv1d for sm_100a.
//
//   warp 0 (1 lane)  : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1 (1 lane)  : MMA issuer     (tcgen05.mma kind::f16, fp32 accumulators in TMEM, 2 accumulator stages)
//   warps 2..5       : epilogue       (tcgen05.ld TMEM->regs, bias/residual/SiLU/dSiLU/GroupNorm-stats, global stores)
//
// One CTA per SM, tiles 128 x BN (BN <= 256) distributed round-robin; the epilogue of tile i overlaps the
// main loop of tile i+1 through the double-buffered TMEM accumulator.
//
// Reference ops replaced: see include/osufusion_b200.h (of_gemm).
#include <stdlib.h>

#include "host_common.h"
#include "ptx.cuh"

namespace ofx {

constexpr int kBM = 128;
constexpr int kBK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int kThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two warps per TMEM lane quarter)
constexpr int kMaxStages = 8;
constexpr uint32_t kABytes = kBM * kBK * 2;  // 16 KB
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccStride = 256;

struct GemmKParams {
  int mode, b_mn, batch, rows, N, K, taps, shift0, shift_step;
  int BN, num_stages;
  uint32_t stage_bytes, tx_bytes;
  int m_tiles, n_tiles, num_tiles;
  int k_chunks;       // FWD: ceil(K/64); WGRAD: ceil(rows/64)
  int k_iters_total;  // FWD: taps*k_chunks; WGRAD: batch*k_chunks
  int split_k, k_iters_per_split;
  // epilogue
  const float* bias;
  const float* aux_f32;
  long long aux_f32_ld, aux_f32_bs;
  const __nv_bfloat16* aux_bf16;
  long long aux_bf16_ld, aux_bf16_bs;
  int aux_is_dsilu, act;
  __nv_bfloat16* pre_bf16;
  __nv_bfloat16* out_bf16;
  long long out_bf16_ld, out_bf16_bs;
  float* out_f32;
  long long out_f32_ld, out_f32_bs;
  double* stats;
};

struct TileCoord {
  int b, m0, n0, tap, k_begin, k_end;
};

// Cluster mode: `tile` enumerates PAIRS of M-adjacent tiles (m_tiles is counted in pairs by the host); `rank` picks the tile.
__device__ __forceinline__ TileCoord decode_tile(const GemmKParams& p, int tile, int rank = 0, int mc = 0) {
  TileCoord tc;
  if (p.mode == OF_GEMM_FWD) {
    int n_blk = tile % p.n_tiles;
    int rest = tile / p.n_tiles;
    int m_blk = rest % p.m_tiles;
    tc.b = rest / p.m_tiles;
    tc.m0 = (mc ? 2 * m_blk + rank : m_blk) * kBM;
    tc.n0 = n_blk * p.BN;
    tc.tap = 0;
    tc.k_begin = 0;
    tc.k_end = p.k_iters_total;
  } else {
    int split = tile % p.split_k;
    int rest = tile / p.split_k;
    int n_blk = rest % p.n_tiles;
    rest /= p.n_tiles;
    int m_blk = rest % p.m_tiles;
    tc.tap = rest / p.m_tiles;
    tc.b = 0;
    tc.m0 = (mc ? 2 * m_blk + rank : m_blk) * kBM;
    tc.n0 = n_blk * p.BN;
    tc.k_begin = split * p.k_iters_per_split;
    tc.k_end = min(tc.k_begin + p.k_iters_per_split, p.k_iters_total);
  }
  return tc;
}

__device__ __forceinline__ void st_global_v4(void* ptr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* ptr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// ---- epilogue variants.  The epilogue is instruction-issue bound (8 warps, one accumulator row per thread), so the fused
// options are compile-time flags: each (flag set) the engine uses gets its own instantiation with the dead paths removed;
// anything else runs the kEpiRuntime instance, which reads the same flags from the parameter block.
enum : uint32_t {
  E_BIAS = 1u, E_AUX32 = 2u, E_AUX16_ADD = 4u, E_AUX16_DSILU = 8u, E_PRE16 = 16u, E_SILU = 32u, E_O16 = 64u, E_O32 = 128u,
  E_STATS = 256u, E_WGRAD = 512u, kEpiRuntime = 0x80000000u
};

template <uint32_t F>
struct EpiFlags {
  bool bias, aux32, aux16_add, aux16_dsilu, pre16, silu, o16, o32, stats;
  __device__ __forceinline__ explicit EpiFlags(const GemmKParams& p) {
    if (F == kEpiRuntime) {
      bias = p.bias != nullptr; aux32 = p.aux_f32 != nullptr;
      aux16_add = p.aux_bf16 != nullptr && !p.aux_is_dsilu; aux16_dsilu = p.aux_bf16 != nullptr && p.aux_is_dsilu;
      pre16 = p.pre_bf16 != nullptr; silu = p.act == OF_ACT_SILU; o16 = p.out_bf16 != nullptr; o32 = p.out_f32 != nullptr;
      stats = p.stats != nullptr;
    } else {
      bias = F & E_BIAS; aux32 = F & E_AUX32; aux16_add = F & E_AUX16_ADD; aux16_dsilu = F & E_AUX16_DSILU; pre16 = F & E_PRE16;
      silu = F & E_SILU; o16 = F & E_O16; o32 = F & E_O32; stats = F & E_STATS;
    }
  }
};

__device__ __forceinline__ void unpack8(const uint4& u, float* x) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; x[4] = c.x; x[5] = c.y; x[6] = d.x; x[7] = d.y;
}

// ---- epilogue staging tile: 32 rows x 128 bytes per warp, the 16-byte group index XOR-swizzled with (row & 7).
// A thread owns one accumulator ROW, so direct global accesses touch 32 different cache lines per warp instruction (measured:
// ~2.8 clk per line on the LSU, the bound of every epilogue with fp32 streams); staged, an instruction covers 4 full 128-byte
// lines (fp32) or 8 64-byte row segments (bf16).
__device__ __forceinline__ uint32_t stage_at(uint32_t stg, int row, int grp) { return stg + row * 128 + ((grp ^ (row & 7)) << 4); }

// f[32] = this lane's row of a 32-column chunk at column nb -> bf16 global rows (row_base + 0..31), coalesced.
template <bool FULL>
__device__ __forceinline__ void store_chunk_bf16(uint32_t stg, int lane, const float (&f)[32], __nv_bfloat16* base, long long ld,
                                                 int row_base, int row_lim, int nb, int N) {
#pragma unroll
  for (int g = 0; g < 4; ++g)
    sts128(stage_at(stg, lane, g), pack_bf16x2(f[8 * g], f[8 * g + 1]), pack_bf16x2(f[8 * g + 2], f[8 * g + 3]),
           pack_bf16x2(f[8 * g + 4], f[8 * g + 5]), pack_bf16x2(f[8 * g + 6], f[8 * g + 7]));
  __syncwarp();
  const int rr = lane >> 2, gg = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int mm = row_base + i * 8 + rr, nn = nb + gg * 8;
    if (mm < row_lim && (FULL || nn < N)) {
      const uint4 t = lds128(stage_at(stg, i * 8 + rr, gg));
      st_global_v4(base + (long long)mm * ld + nn, t.x, t.y, t.z, t.w);
    }
  }
  __syncwarp();
}
template <bool FULL, bool kRed>
__device__ __forceinline__ void store_chunk_f32(uint32_t stg, int lane, const float (&f)[32], float* base, long long ld, int row_base,
                                                int row_lim, int nb, int N) {
#pragma unroll
  for (int g = 0; g < 8; ++g)
    sts128(stage_at(stg, lane, g), __float_as_uint(f[4 * g]), __float_as_uint(f[4 * g + 1]), __float_as_uint(f[4 * g + 2]),
           __float_as_uint(f[4 * g + 3]));
  __syncwarp();
  const int rr = lane >> 3, gg = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int mm = row_base + i * 4 + rr, nn = nb + gg * 4;
    if (mm < row_lim && (FULL || nn < N)) {
      const uint4 u = lds128(stage_at(stg, i * 4 + rr, gg));
      const float4 t = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
      if (kRed) red_add_v4(base + (long long)mm * ld + nn, t.x, t.y, t.z, t.w);
      else *reinterpret_cast<float4*>(base + (long long)mm * ld + nn) = t;
    }
  }
  __syncwarp();
}
// Coalesced gather of a 32x32 fp32 / bf16 tile into this lane's row.
template <bool FULL>
__device__ __forceinline__ void load_chunk_f32(uint32_t stg, int lane, float (&x)[32], const float* base, long long ld, int row_base,
                                               int row_lim, int nb, int N) {
  const int rr = lane >> 3, gg = lane & 7;
  float4 t[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int mm = row_base + i * 4 + rr, nn = nb + gg * 4;
    t[i] = (mm < row_lim && (FULL || nn < N)) ? *reinterpret_cast<const float4*>(base + (long long)mm * ld + nn)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
    sts128(stage_at(stg, i * 4 + rr, gg), __float_as_uint(t[i].x), __float_as_uint(t[i].y), __float_as_uint(t[i].z), __float_as_uint(t[i].w));
  __syncwarp();
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const uint4 u = lds128(stage_at(stg, lane, g));
    x[4 * g] = __uint_as_float(u.x); x[4 * g + 1] = __uint_as_float(u.y); x[4 * g + 2] = __uint_as_float(u.z); x[4 * g + 3] = __uint_as_float(u.w);
  }
  __syncwarp();
}
template <bool FULL>
__device__ __forceinline__ void load_chunk_bf16(uint32_t stg, int lane, float (&x)[32], const __nv_bfloat16* base, long long ld,
                                                int row_base, int row_lim, int nb, int N) {
  const int rr = lane >> 2, gg = lane & 3;
  uint4 t[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int mm = row_base + i * 8 + rr, nn = nb + gg * 8;
    t[i] = (mm < row_lim && (FULL || nn < N)) ? *reinterpret_cast<const uint4*>(base + (long long)mm * ld + nn) : make_uint4(0u, 0u, 0u, 0u);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) sts128(stage_at(stg, i * 8 + rr, gg), t[i].x, t[i].y, t[i].z, t[i].w);
  __syncwarp();
#pragma unroll
  for (int g = 0; g < 4; ++g) unpack8(lds128(stage_at(stg, lane, g)), &x[8 * g]);
  __syncwarp();
}

// One 32-column chunk (columns nb..nb+31) of the 32 accumulator rows of this warp (thread = row row_base + lane).  Warp-collective.
// FULL: all 32 columns are < N (no column predicates).  Base pointers are per batch (row 0, column 0).
template <uint32_t F, bool FULL>
__device__ __forceinline__ void epi_chunk(const GemmKParams& p, const EpiFlags<F>& e, uint32_t stg, int lane, const uint32_t (&v)[32],
                                          int row_base, int nb, const float* aux32, const __nv_bfloat16* aux16,
                                          __nv_bfloat16* pre16, __nv_bfloat16* o16, float* o32, float& s1, float& s2) {
  const int row_lim = p.rows;
  float f[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
  if (e.bias) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      if (FULL || nb + g * 4 < p.N) {
        const float4 b = *reinterpret_cast<const float4*>(p.bias + nb + g * 4);   // same address in every lane: one line
        f[4 * g] += b.x; f[4 * g + 1] += b.y; f[4 * g + 2] += b.z; f[4 * g + 3] += b.w;
      }
    }
  }
  if (e.aux32) {
    float x[32];
    load_chunk_f32<FULL>(stg, lane, x, aux32, p.aux_f32_ld, row_base, row_lim, nb, p.N);
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] += x[j];
  }
  float x16[32];
  if (e.aux16_add || e.aux16_dsilu) {
    load_chunk_bf16<FULL>(stg, lane, x16, aux16, p.aux_bf16_ld, row_base, row_lim, nb, p.N);
    if (e.aux16_add) {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] += x16[j];
    }
  }
  if (e.pre16) store_chunk_bf16<FULL>(stg, lane, f, pre16, p.out_bf16_ld, row_base, row_lim, nb, p.N);
  if (e.silu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = silu_f(f[j]);
  }
  if (e.aux16_dsilu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] *= dsilu_f(x16[j]);
  }
  if (e.o16) store_chunk_bf16<FULL>(stg, lane, f, o16, p.out_bf16_ld, row_base, row_lim, nb, p.N);
  if (e.o32) store_chunk_f32<FULL, false>(stg, lane, f, o32, p.out_f32_ld, row_base, row_lim, nb, p.N);
  if (e.stats) {
    if (row_base + lane < row_lim) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (FULL || nb + j < p.N) {
          const float rv = bf16_round(f[j]);
          s1 += rv;
          s2 = fmaf(rv, rv, s2);
        }
      }
    }
  }
}

// kMC: launched as clusters of 2 CTAs that own two M-adjacent tiles with the same N tile; each CTA fetches half of the shared
// B (weight / X) tile and TMA-multicasts it to both, cutting L2->SM operand traffic per CTA from A+B to A+B/2.
// kMC == 2: CTA PAIR (tcgen05 cta_group::2).  The two CTAs of a cluster own M-adjacent tiles with the same N tile, as in mode 1, but
// the pair's leader issues ONE M=256 MMA for both: each CTA loads only its own 128 rows of A and its HALF of the B tile, so the
// per-SM operand ingress per k-step drops from 48 KB to 32 KB (the main loop is bound by exactly that at 128x256 tiles) and a
// pipeline stage shrinks from 48 KB to 32 KB (6 stages instead of 4).
template <int kMC, uint32_t F>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const GemmKParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte aligned tile ring (SWIZZLE_128B atoms), barriers after it.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)p.num_stages * p.stage_bytes);
  uint64_t* full_bar = bars;                       // [kMaxStages]
  uint64_t* empty_bar = bars + kMaxStages;         // [kMaxStages]
  uint64_t* tmem_full_bar = bars + 2 * kMaxStages; // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* stat_sm = reinterpret_cast<float*>(tmem_base_slot + 2);   // [2] per-tile GroupNorm partial sums
  uint8_t* staging = reinterpret_cast<uint8_t*>(bars) + 256;         // [8 epilogue warps][4 KB]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    stat_sm[0] = 0.f;
    stat_sm[1] = 0.f;
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < p.num_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kMC == 1 ? 2 : 1);   // multicast mode: both CTAs' MMA warps release a stage (it receives multicast data)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], kMC == 2 ? 16 : 8);   // pair mode: the leader's barrier collects the epilogue warps of BOTH CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (kMC == 2) {
      tmem_alloc_2sm(tmem_base_slot, kTmemCols);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_base_slot, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kMC) cluster_sync_all();   // peer barriers must be initialised before any multicast can arrive
  tc_fence_after();
  pdl_wait();                    // everything above overlapped the previous kernel's tail; its results are visible from here on
  const uint32_t tmem_base = *tmem_base_slot;
  const int crank = kMC ? (int)cluster_ctarank() : 0;
  const int tile_first = kMC ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = kMC ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  constexpr bool kWgrad = (F == E_WGRAD);
  const bool a_mn = kWgrad;
  const bool b_mn = kWgrad || (p.b_mn != 0);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // (whole warp runs the loop convergently; only the TMA / expect_tx instructions are predicated on the elected lane, so
    //  coordinates and descriptors stay in uniform registers instead of going through per-instruction ELECT waterfall loops)
    const bool leader = elect_one();
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile_first; tile < p.num_tiles; tile += tile_step) {
        TileCoord tc = decode_tile(p, tile, crank, kMC);
        for (int it = tc.k_begin; it < tc.k_end; ++it) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = ring + (size_t)stage * p.stage_bytes;
          uint8_t* sb = sa + kABytes;
          if (leader && kMC == 2) {
            // pair mode: both CTAs load (own A rows, own half of B) into their own ring; all bytes are credited to the LEADER's barrier
            if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * p.tx_bytes);
            const int nch2 = p.BN / 128;    // 64-column chunks of this CTA's half of an MN-major B tile
            if (!kWgrad) {
              int t = it / p.k_chunks;
              int kc = it - t * p.k_chunks;
              tma_load_3d_2sm(sa, &tmap_a, &full_bar[stage], kc * kBK, tc.m0 + p.shift0 + t * p.shift_step, tc.b);
              if (!b_mn) {
                tma_load_3d_2sm(sb, &tmap_b, &full_bar[stage], kc * kBK, tc.n0 + crank * (p.BN / 2), t);
              } else {
                for (int j = 0; j < nch2; ++j)
                  tma_load_3d_2sm(sb + j * 8192, &tmap_b, &full_bar[stage], tc.n0 + (crank * nch2 + j) * 64, kc * kBK, t);
              }
            } else {
              int b = it / p.k_chunks;
              int lc = it - b * p.k_chunks;
              for (int j = 0; j < 2; ++j)
                tma_load_3d_2sm(sa + j * 8192, &tmap_a, &full_bar[stage], tc.m0 + j * 64, lc * kBK, b);
              int shift = p.shift0 + tc.tap * p.shift_step;
              for (int j = 0; j < nch2; ++j)
                tma_load_3d_2sm(sb + j * 8192, &tmap_b, &full_bar[stage], tc.n0 + (crank * nch2 + j) * 64, lc * kBK + shift, b);
            }
          } else if (leader) {
          mbar_arrive_expect_tx(&full_bar[stage], p.tx_bytes);
          if (!kWgrad) {
            int t = it / p.k_chunks;
            int kc = it - t * p.k_chunks;
            tma_load_3d(sa, &tmap_a, &full_bar[stage], kc * kBK, tc.m0 + p.shift0 + t * p.shift_step, tc.b);
            if (!b_mn) {
              if (kMC) {   // this CTA's half of the B rows (box = BN/2 rows), multicast to both CTAs
                const int hr = p.BN / 2;
                tma_load_3d_mc(sb + crank * hr * 128, &tmap_b, &full_bar[stage], kc * kBK, tc.n0 + crank * hr, t, 3);
              } else {
                tma_load_3d(sb, &tmap_b, &full_bar[stage], kc * kBK, tc.n0, t);
              }
            } else {
              const int nch = p.BN / 64;
              if (kMC) {
                for (int j = crank * (nch / 2); j < (crank + 1) * (nch / 2); ++j)
                  tma_load_3d_mc(sb + j * 8192, &tmap_b, &full_bar[stage], tc.n0 + j * 64, kc * kBK, t, 3);
              } else {
                for (int j = 0; j < nch; ++j)
                  tma_load_3d(sb + j * 8192, &tmap_b, &full_bar[stage], tc.n0 + j * 64, kc * kBK, t);
              }
            }
          } else {
            int b = it / p.k_chunks;
            int lc = it - b * p.k_chunks;
            for (int j = 0; j < 2; ++j)
              tma_load_3d(sa + j * 8192, &tmap_a, &full_bar[stage], tc.m0 + j * 64, lc * kBK, b);
            int shift = p.shift0 + tc.tap * p.shift_step;
            const int nch = p.BN / 64;
            if (kMC) {
              for (int j = crank * (nch / 2); j < (crank + 1) * (nch / 2); ++j)
                tma_load_3d_mc(sb + j * 8192, &tmap_b, &full_bar[stage], tc.n0 + j * 64, lc * kBK + shift, b, 3);
            } else {
              for (int j = 0; j < nch; ++j)
                tma_load_3d(sb + j * 8192, &tmap_b, &full_bar[stage], tc.n0 + j * 64, lc * kBK + shift, b);
            }
          }
          }
          __syncwarp();
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform loop, elected lane issues)
    const bool leader = elect_one();
    if (kMC != 2 || crank == 0) {      // pair mode: only the leader CTA issues (for both CTAs' accumulators)
      const uint32_t idesc = make_idesc_bf16(kMC == 2 ? 2 * kBM : kBM, p.BN, a_mn ? 1u : 0u, b_mn ? 1u : 0u);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = tile_first; tile < p.num_tiles; tile += tile_step) {
        TileCoord tc = decode_tile(p, tile, crank, kMC);
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kAccStride;
        for (int it = tc.k_begin; it < tc.k_end; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          uint32_t sa = smem_u32(ring + (size_t)stage * p.stage_bytes);
          uint32_t sb = sa + kABytes;
          if (leader) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              uint64_t da = a_mn ? make_smem_desc(sa + k * 2048, 8192, 1024) : make_smem_desc(sa + k * 32, 16, 1024);
              uint64_t db = b_mn ? make_smem_desc(sb + k * 2048, 8192, 1024) : make_smem_desc(sb + k * 32, 16, 1024);
              if (kMC == 2) umma_f16_ss_2sm(tmem_d, da, db, idesc, (it > tc.k_begin || k > 0) ? 1u : 0u);
              else umma_f16_ss(tmem_d, da, db, idesc, (it > tc.k_begin || k > 0) ? 1u : 0u);
            }
            if (kMC == 2) umma_commit_2sm(&empty_bar[stage], 3);
            else if (kMC == 1) umma_commit_mc(&empty_bar[stage], 3);
            else umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (leader) {
          if (kMC == 2) umma_commit_2sm(&tmem_full_bar[acc], 3);
          else umma_commit(&tmem_full_bar[acc]);
        }
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (2..9)
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int chalf = (warp - 2) >> 2;   // the two warps of a quarter take alternate 32-column chunks
    const uint32_t stg = smem_u32(staging + (warp - 2) * 4096);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = tile_first; tile < p.num_tiles; tile += tile_step) {
      TileCoord tc = decode_tile(p, tile, crank, kMC);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const bool k_nonempty = tc.k_end > tc.k_begin;
      float s1 = 0.f, s2 = 0.f;
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kAccStride;
      const int nchunks = p.BN / 32;
      if (!kWgrad) {
        const EpiFlags<F> e(p);
        const long long brow = (long long)tc.b;
        const int row_base = tc.m0 + q * 32;
        const float* aux32 = e.aux32 ? p.aux_f32 + brow * p.aux_f32_bs : nullptr;
        const __nv_bfloat16* aux16 = (e.aux16_add || e.aux16_dsilu) ? p.aux_bf16 + brow * p.aux_bf16_bs : nullptr;
        __nv_bfloat16* o16 = e.o16 ? p.out_bf16 + brow * p.out_bf16_bs : nullptr;
        __nv_bfloat16* pre16 = e.pre16 ? p.pre_bf16 + brow * p.out_bf16_bs : nullptr;
        float* o32 = e.o32 ? p.out_f32 + brow * p.out_f32_bs : nullptr;
        // two 32-column chunks per iteration: both TMEM loads are in flight before the single wait
        for (int c = chalf; c < nchunks; c += 4) {
          uint32_t va[32], vb[32];
          const bool two = c + 2 < nchunks;
          tmem_ld_32x32b_x32(tacc + c * 32, va);
          if (two) tmem_ld_32x32b_x32(tacc + (c + 2) * 32, vb);
          tmem_wait_ld();
          if (row_base < p.rows) {   // warp-uniform: the staged epilogue is warp-collective
            const int na = tc.n0 + c * 32, nb2 = tc.n0 + (c + 2) * 32;
            if (na + 32 <= p.N) epi_chunk<F, true>(p, e, stg, lane, va, row_base, na, aux32, aux16, pre16, o16, o32, s1, s2);
            else if (na < p.N) epi_chunk<F, false>(p, e, stg, lane, va, row_base, na, aux32, aux16, pre16, o16, o32, s1, s2);
            if (two) {
              if (nb2 + 32 <= p.N) epi_chunk<F, true>(p, e, stg, lane, vb, row_base, nb2, aux32, aux16, pre16, o16, o32, s1, s2);
              else if (nb2 < p.N) epi_chunk<F, false>(p, e, stg, lane, vb, row_base, nb2, aux32, aux16, pre16, o16, o32, s1, s2);
            }
          }
        }
      } else {
        // WGRAD: atomically accumulate fp32 into out_f32[tap][m][n] (coalesced red.v4 through the staging tile)
        float* o32 = p.out_f32 + (long long)tc.tap * p.out_f32_bs;
        const int row_base = tc.m0 + q * 32;
        for (int c = chalf; c < nchunks; c += 2) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tacc + c * 32, v);
          tmem_wait_ld();
          const int nb = tc.n0 + c * 32;
          if (k_nonempty && nb < p.N && row_base < p.K) {
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            store_chunk_f32<false, true>(stg, lane, f, o32, p.out_f32_ld, row_base, p.K, nb, p.N);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {   // accumulator drained: the MMA warp may start the next tile into it
        if (kMC == 2) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
        else mbar_arrive(&tmem_empty_bar[acc]);
      }
      if (!kWgrad) {
      if (F == kEpiRuntime ? (p.stats != nullptr) : ((F & E_STATS) != 0)) {
        // one pair of global double atomics per TILE: the 8 epilogue warps first combine in shared memory
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane == 0) {
          atomicAdd(&stat_sm[0], s1);
          atomicAdd(&stat_sm[1], s2);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (warp == 2 && lane == 0) {
          atomicAdd(p.stats + 2 * tc.b, (double)stat_sm[0]);
          atomicAdd(p.stats + 2 * tc.b + 1, (double)stat_sm[1]);
          stat_sm[0] = 0.f;
          stat_sm[1] = 0.f;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kMC) cluster_sync_all();   // no CTA may exit while its peer can still multicast into it / arrive on its barriers
  tc_fence_after();
  if (warp == 1) {
    if (kMC == 2) tmem_dealloc_2sm(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

static int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace ofx

namespace ofx {

template <int kMC, uint32_t F>
static int launch_one(const GemmKParams& p, const CUtensorMap& ta, const CUtensorMap& tb, size_t smem_bytes, cudaStream_t stream) {
  static bool attr_set = false;   // per instantiation
  if (!attr_set) {
    OF_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<kMC, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const bool pdl = pdl_level() >= 1;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (!kMC) {
    cfg.gridDim = dim3(p.num_tiles < device_sm_count() ? p.num_tiles : device_sm_count());
  } else {
    int pairs = device_sm_count() / 2;
    if (p.num_tiles < pairs) pairs = p.num_tiles;
    cfg.gridDim = dim3(2 * pairs);
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl) {   // programmatic dependent launch: the kernel's prologue overlaps the previous kernel's tail (griddepcontrol.wait inside)
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  OF_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_kernel<kMC, F>, ta, tb, p));
  return OF_OK;
}

// The flag sets the engine produces (osufusion_b200/engine.py); everything else falls back to the runtime-flag instance.
#define OF_GEMM_EPILOGUES(X)                                                                                                   \
  X(E_WGRAD)                                                                                                                    \
  X(E_O16) X(E_BIAS | E_O16) X(E_BIAS | E_O16 | E_STATS)                               /* qkv / conv, res_conv / conv + GN stats */ \
  X(E_BIAS | E_AUX32 | E_O32 | E_O16)                                                  /* to_out, ff2: + fp32 residual */        \
  X(E_BIAS | E_SILU | E_PRE16 | E_O16) X(E_BIAS | E_SILU | E_O16)                      /* ff1 (training / inference) */          \
  X(E_O32) X(E_AUX32 | E_O32) X(E_AUX32 | E_O32 | E_O16) X(E_O32 | E_O16)              /* dgrad, gradient accumulated in place */ \
  X(E_AUX16_DSILU | E_O16)                                                             /* ff backward through SiLU */            \
  X(E_BIAS | E_AUX16_ADD | E_O16) X(E_AUX16_ADD | E_O16)                               /* Parallel sampler, Downsample fix-up */

static int launch_gemm(uint32_t mask, int mc, const GemmKParams& p, const CUtensorMap& ta, const CUtensorMap& tb, size_t smem_bytes,
                       cudaStream_t stream) {
#define OF_CASE(FLAGS)                                                                                   \
  if (mask == (uint32_t)(FLAGS))                                                                         \
    return mc == 2   ? launch_one<2, (uint32_t)(FLAGS)>(p, ta, tb, smem_bytes, stream)                   \
           : mc == 1 ? launch_one<1, (uint32_t)(FLAGS)>(p, ta, tb, smem_bytes, stream)                   \
                     : launch_one<0, (uint32_t)(FLAGS)>(p, ta, tb, smem_bytes, stream);
  OF_GEMM_EPILOGUES(OF_CASE)
#undef OF_CASE
  return mc == 2   ? launch_one<2, kEpiRuntime>(p, ta, tb, smem_bytes, stream)
         : mc == 1 ? launch_one<1, kEpiRuntime>(p, ta, tb, smem_bytes, stream)
                   : launch_one<0, kEpiRuntime>(p, ta, tb, smem_bytes, stream);
}

}  // namespace ofx

using namespace ofx;

extern "C" int of_gemm(const of_gemm_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  OF_REQUIRE(a != nullptr, "of_gemm: null args");
  OF_REQUIRE(a->mode == OF_GEMM_FWD || a->mode == OF_GEMM_WGRAD, "of_gemm: bad mode %d", a->mode);
  OF_REQUIRE(a->batch >= 1 && a->rows >= 1 && a->N >= 1 && a->K >= 1 && a->taps >= 1, "of_gemm: bad sizes");
  OF_REQUIRE(a->N % 8 == 0, "of_gemm: N=%d must be a multiple of 8", a->N);
  OF_REQUIRE(a->a && a->b, "of_gemm: null operand");
  OF_REQUIRE(a->a_ld % 8 == 0 && a->b_ld % 8 == 0, "of_gemm: leading dims must be multiples of 8");

  GemmKParams p;
  memset(&p, 0, sizeof(p));
  p.mode = a->mode;
  p.b_mn = a->b_mn_major;
  p.batch = a->batch;
  p.rows = a->rows;
  p.N = a->N;
  p.K = a->K;
  p.taps = a->taps;
  p.shift0 = a->shift0;
  p.shift_step = a->shift_step;
  const bool wgrad = a->mode == OF_GEMM_WGRAD;
  const bool b_mn = wgrad || a->b_mn_major;

  // ---- tile N
  int BN = a->block_n;
  if (BN <= 0) {
    if (a->N >= 256) BN = 256;
    else if (a->N > 128) BN = (a->N > 192) ? 256 : 192;
    else if (a->N > 64) BN = 128;
    else if (a->N > 32) BN = 64;
    else BN = b_mn ? 64 : 32;
    // Tile width by a small cost model: rounds of the persistent grid x time of one tile.  A CTA pair owns one 256-row tile, so a
    // paired launch has sm/2 slots.  Per FLOP a narrower tile needs more operand bytes from L2 (A is re-read per N tile: 24 KB per
    // k-step at BN=128 against 32 KB for twice the work at BN=256), hence the penalties; ties go to the wider tile.
    if (!wgrad && BN == 256) {
      const int mt = ceil_div(a->rows, kBM);
      const bool pair = mt >= 2;
      const long long rows_units = (long long)a->batch * (pair ? ceil_div(mt, 2) : mt);
      const int slots = pair ? device_sm_count() / 2 : device_sm_count();
      static const int legacy = [] { const char* e = getenv("OF_GEMM_BN_LEGACY"); return e ? atoi(e) : 0; }();
      if (legacy) {
        long long tiles256 = (long long)a->batch * mt * ceil_div(a->N, 256);
        if (tiles256 < device_sm_count()) BN = 128;
      } else {
        auto rounds = [&](int bn) { return (double)((rows_units * ceil_div(a->N, bn) + slots - 1) / slots); };
        double best = rounds(256) * 2.0;
        const double c128 = rounds(128) * 1.15;
        if (c128 < best) best = c128, BN = 128;
        if (!b_mn) {                                   // MN-major B pairs need an even number of 64-column swizzle atoms
          const double c192 = rounds(192) * 1.55;
          if (c192 < best) best = c192, BN = 192;
        }
      }
    }
  }
  OF_REQUIRE(BN % 32 == 0 && BN >= 32 && BN <= 256, "of_gemm: block_n=%d invalid", BN);
  if (b_mn) OF_REQUIRE(BN % 64 == 0, "of_gemm: block_n=%d must be a multiple of 64 for MN-major B", BN);
  p.BN = BN;
  // cluster of 2 CTAs on M-adjacent tiles: needs two M tiles to pair and a B tile that splits into two swizzle-aligned halves.
  //   mode 2 (default): CTA pair, tcgen05 cta_group::2 -- each CTA holds only its half of B;  mode 1: TMA multicast of the halves
  const int m_tiles_real = wgrad ? ceil_div(a->K, kBM) : ceil_div(a->rows, kBM);
  bool use_mc = (m_tiles_real >= 2) && (b_mn ? ((BN / 64) % 2 == 0) : (BN % 16 == 0 && BN >= 64));
  int mc_mode = use_mc ? 2 : 0;
  {
    const char* e = getenv("OF_GEMM_NO_MULTICAST");
    if (e && e[0] == '1') use_mc = false, mc_mode = 0;
    const char* e2 = getenv("OF_GEMM_2CTA");
    if (e2 && e2[0] == '0' && mc_mode == 2) mc_mode = 1;
  }
  const uint32_t b_bytes = (uint32_t)BN * 128u;
  p.stage_bytes = kABytes + (mc_mode == 2 ? b_bytes / 2 : b_bytes);
  p.tx_bytes = p.stage_bytes;
  int stages = (int)((192u * 1024u) / p.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  p.num_stages = stages;

  if (!wgrad) {
    p.m_tiles = use_mc ? ceil_div(m_tiles_real, 2) : m_tiles_real;
    p.n_tiles = ceil_div(a->N, BN);
    p.num_tiles = a->batch * p.m_tiles * p.n_tiles;
    p.k_chunks = ceil_div(a->K, kBK);
    p.k_iters_total = a->taps * p.k_chunks;
    p.split_k = 1;
    p.k_iters_per_split = p.k_iters_total;
  } else {
    OF_REQUIRE(a->out_f32 != nullptr, "of_gemm(wgrad): out_f32 required");
    OF_REQUIRE(a->K % 8 == 0, "of_gemm(wgrad): M=%d must be a multiple of 8", a->K);
    p.m_tiles = use_mc ? ceil_div(m_tiles_real, 2) : m_tiles_real;
    p.n_tiles = ceil_div(a->N, BN);
    p.k_chunks = ceil_div(a->rows, kBK);
    p.k_iters_total = a->batch * p.k_chunks;
    int base_tiles = a->taps * p.m_tiles * p.n_tiles;
    int split = a->split_k;
    if (split <= 0) {
      split = ceil_div((use_mc ? 1 : 2) * device_sm_count(), base_tiles);
      int max_split = ceil_div(p.k_iters_total, 4);  // at least 4 k-iterations per split
      if (split > max_split) split = max_split;
      if (split < 1) split = 1;
    }
    if (split > p.k_iters_total) split = p.k_iters_total;
    p.k_iters_per_split = ceil_div(p.k_iters_total, split);
    p.split_k = ceil_div(p.k_iters_total, p.k_iters_per_split);
    p.num_tiles = base_tiles * p.split_k;
  }

  p.bias = a->bias;
  p.aux_f32 = a->aux_f32;
  p.aux_f32_ld = a->aux_f32_ld;
  p.aux_f32_bs = a->aux_f32_batch_stride;
  p.aux_bf16 = reinterpret_cast<const __nv_bfloat16*>(a->aux_bf16);
  p.aux_bf16_ld = a->aux_bf16_ld;
  p.aux_bf16_bs = a->aux_bf16_batch_stride;
  p.aux_is_dsilu = a->aux_is_dsilu;
  p.act = a->act;
  p.pre_bf16 = reinterpret_cast<__nv_bfloat16*>(a->pre_bf16);
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(a->out_bf16);
  p.out_bf16_ld = a->out_bf16_ld;
  p.out_bf16_bs = a->out_bf16_batch_stride;
  p.out_f32 = a->out_f32;
  p.out_f32_ld = a->out_f32_ld;
  p.out_f32_bs = a->out_f32_batch_stride;
  p.stats = a->stats;
  if (!wgrad) {
    OF_REQUIRE(p.out_bf16 || p.out_f32 || p.pre_bf16, "of_gemm: no output");
    if (p.out_bf16 || p.pre_bf16) OF_REQUIRE(p.out_bf16_ld % 8 == 0, "of_gemm: out_bf16_ld %% 8");
    if (p.out_f32) OF_REQUIRE(p.out_f32_ld % 4 == 0, "of_gemm: out_f32_ld %% 4");
    if (p.aux_f32) OF_REQUIRE(p.aux_f32_ld % 4 == 0, "of_gemm: aux_f32_ld %% 4");
    if (p.aux_bf16) OF_REQUIRE(p.aux_bf16_ld % 8 == 0, "of_gemm: aux_bf16_ld %% 8");
  } else {
    OF_REQUIRE(p.out_f32_ld % 4 == 0, "of_gemm(wgrad): out_f32_ld %% 4");
  }

  // ---- tensor maps
  CUtensorMap ta, tb;
  int rc;
  if (!wgrad) {
    unsigned long long bs = a->batch > 1 ? (unsigned long long)a->a_batch_stride : (unsigned long long)a->rows * a->a_ld;
    unsigned long long dims[3] = {(unsigned long long)a->K, (unsigned long long)a->rows, (unsigned long long)a->batch};
    unsigned long long str[2] = {(unsigned long long)a->a_ld * 2ull, bs * 2ull};
    unsigned box[3] = {64, (unsigned)kBM, 1};
    if ((rc = make_tmap_bf16(&ta, a->a, 3, dims, str, box)) != OF_OK) return rc;
    if (!b_mn) {
      unsigned long long ts = a->taps > 1 ? (unsigned long long)a->b_tap_stride : (unsigned long long)a->N * a->b_ld;
      unsigned long long bd[3] = {(unsigned long long)a->K, (unsigned long long)a->N, (unsigned long long)a->taps};
      unsigned long long bstr[2] = {(unsigned long long)a->b_ld * 2ull, ts * 2ull};
      unsigned bbox[3] = {64, (unsigned)(use_mc ? BN / 2 : BN), 1};
      if ((rc = make_tmap_bf16(&tb, a->b, 3, bd, bstr, bbox)) != OF_OK) return rc;
    } else {
      unsigned long long ts = a->taps > 1 ? (unsigned long long)a->b_tap_stride : (unsigned long long)a->K * a->b_ld;
      unsigned long long bd[3] = {(unsigned long long)a->N, (unsigned long long)a->K, (unsigned long long)a->taps};
      unsigned long long bstr[2] = {(unsigned long long)a->b_ld * 2ull, ts * 2ull};
      unsigned bbox[3] = {64, 64, 1};
      if ((rc = make_tmap_bf16(&tb, a->b, 3, bd, bstr, bbox)) != OF_OK) return rc;
    }
  } else {
    unsigned long long abs_ = a->batch > 1 ? (unsigned long long)a->a_batch_stride : (unsigned long long)a->rows * a->a_ld;
    unsigned long long bbs = a->batch > 1 ? (unsigned long long)a->b_tap_stride : (unsigned long long)a->rows * a->b_ld;
    unsigned long long ad[3] = {(unsigned long long)a->K, (unsigned long long)a->rows, (unsigned long long)a->batch};
    unsigned long long astr[2] = {(unsigned long long)a->a_ld * 2ull, abs_ * 2ull};
    unsigned long long bd[3] = {(unsigned long long)a->N, (unsigned long long)a->rows, (unsigned long long)a->batch};
    unsigned long long bstr[2] = {(unsigned long long)a->b_ld * 2ull, bbs * 2ull};
    unsigned box[3] = {64, 64, 1};
    if ((rc = make_tmap_bf16(&ta, a->a, 3, ad, astr, box)) != OF_OK) return rc;
    if ((rc = make_tmap_bf16(&tb, a->b, 3, bd, bstr, box)) != OF_OK) return rc;
  }

  size_t smem_bytes = (size_t)p.num_stages * p.stage_bytes + 1024 /*align slack*/ + 256 /*barriers*/ + 8 * 4096 /*epilogue staging*/;
  uint32_t mask;
  if (wgrad) {
    mask = E_WGRAD;
  } else {
    mask = (p.bias ? E_BIAS : 0u) | (p.aux_f32 ? E_AUX32 : 0u) | (p.aux_bf16 ? (p.aux_is_dsilu ? E_AUX16_DSILU : E_AUX16_ADD) : 0u) |
           (p.pre_bf16 ? E_PRE16 : 0u) | (p.act == OF_ACT_SILU ? E_SILU : 0u) | (p.out_bf16 ? E_O16 : 0u) |
           (p.out_f32 ? E_O32 : 0u) | (p.stats ? E_STATS : 0u);
  }
  int rc2 = launch_gemm(mask, mc_mode, p, ta, tb, smem_bytes, stream);
  if (rc2 != OF_OK) return rc2;
  OF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OF_OK;
}
