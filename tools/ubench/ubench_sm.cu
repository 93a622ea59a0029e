// Micro-benchmarks of the per-SM resources the attention softmax warps compete for (sm_100a):
//   tcgen05.ld / tcgen05.st throughput vs number of warps, MUFU ex2, packed f32x2 FMA / ADD, 3-input max, cvt.bf16x2.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_sm ubench_sm.cu ; run: ./ubench_sm
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../osufusion_b200/csrc/ptx.cuh"
using namespace ofx;

__global__ void k_tmem_ld(long long* out, int iters, int mode) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t col0 = (warp >> 2) * 64;    // warps sharing a lane quarter read different columns
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t r[32], r2[32];
    if (mode == 0) {
      tmem_ld_32x32b_x32(tmem + lane_off + col0, r);
      tmem_wait_ld();
      acc += r[0] + r[31];
    } else {
      tmem_ld_32x32b_x32(tmem + lane_off + col0, r);
      tmem_ld_32x32b_x32(tmem + lane_off + col0 + 32, r2);
      tmem_wait_ld();
      acc += r[0] + r2[31];
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678) out[1000] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

__global__ void k_tmem_st(long long* out, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t col0 = (warp >> 2) * 64;
  uint32_t r[16];
  for (int i = 0; i < 16; ++i) r[i] = threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    tmem_st_32x32b_x16(tmem + lane_off + col0, r);
    tmem_st_32x32b_x16(tmem + lane_off + col0 + 16, r);
    tmem_wait_st();
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// mode 0: MUFU ex2; 1: fma.f32x2; 2: add.f32x2; 3: max3; 4: cvt bf16x2; 5: scalar FFMA
__global__ void k_alu(long long* out, float* sink, int iters, int mode) {
  float a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i * 0.01f;
  uint32_t u[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = ex2f(a[i]);
    } else if (mode == 1) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        unsigned long long x, y = 0x3f8000003f800000ull, z = 0x3a0000003a000000ull;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[i]), "f"(a[i + 1]));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(y), "l"(z));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(x));
      }
    } else if (mode == 2) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        unsigned long long x, z = 0x3a0000003a000000ull;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[i]), "f"(a[i + 1]));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(z));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(x));
      }
    } else if (mode == 3) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[i + 1]), "f"(a[(i + 3) & 15]));
    } else if (mode == 4) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        uint32_t p;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(a[i + 1]), "f"(a[i]));
        u[i >> 1] ^= p;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(1.0001f), "f"(0.0005f));
    }
  }
  long long t1 = clock64();
  __syncthreads();
  float s = 0;
  for (int i = 0; i < 16; ++i) s += a[i];
  for (int i = 0; i < 8; ++i) s += __uint_as_float(u[i]);
  if (s == 123.456f) sink[0] = s;
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

int main() {
  long long* d; float* sink;
  cudaMalloc(&d, 4096 * sizeof(long long)); cudaMalloc(&sink, 16);
  long long h[4];
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {4, 8, 16}) {
      k_tmem_ld<<<1, warps * 32>>>(d, iters, mode);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
      double bytes = (double)warps * iters * (mode ? 2 : 1) * 4096.0;
      printf("tmem_ld mode%d (x32%s) warps=%2d: %lld clk, %.1f B/clk/SM, %.1f clk per x32 load per warp  (%s)\n", mode, mode ? " x2 per wait" : "",
             warps, h[0], bytes / h[0], (double)h[0] / (iters * (mode ? 2 : 1)), cudaGetErrorString(e));
    }
  for (int warps : {4, 8, 16}) {
    k_tmem_st<<<1, warps * 32>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    double bytes = (double)warps * iters * 2 * 2048.0;
    printf("tmem_st (2 x16 per wait) warps=%2d: %lld clk, %.1f B/clk/SM  (%s)\n", warps, h[0], bytes / h[0], cudaGetErrorString(e));
  }
  const char* names[] = {"MUFU ex2", "fma.f32x2", "add.f32x2", "max3.f32", "cvt.bf16x2", "fma.f32"};
  for (int mode = 0; mode < 6; ++mode)
    for (int warps : {4, 8, 16}) {
      k_alu<<<1, warps * 32>>>(d, sink, iters, mode);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
      const int per_iter = (mode == 0 || mode == 5) ? 16 : 8;
      double ops = (double)warps * 32 * iters * per_iter;
      printf("%-10s warps=%2d: %lld clk, %.2f thread-instr/clk/SM (%.2f warp-instr/clk)  (%s)\n", names[mode], warps, h[0], ops / h[0],
             ops / h[0] / 32, cudaGetErrorString(e));
    }
  return 0;
}
